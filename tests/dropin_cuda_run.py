"""Helper of test_gpu_dropin_cuda.py (run as a subprocess on the GPU box): builds the UNMODIFIED reference networks
(staged at oracle/_ref/src by oracle/build_ref.py) on cuda with whatever `torch_utils.ops` PYTHONPATH resolves to, runs
forward + backward and saves outputs and gradients. LVG_REF_PLUGINS=1: the reference's own ops with its own prebuilt
CUDA plugins (oracle #2); otherwise this repository's ops must have resolved."""
import os
import sys
import types
import warnings

import torch

warnings.filterwarnings('ignore')
sys.modules.setdefault('imageio', types.ModuleType('imageio'))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

import torch_utils.custom_ops as _co            # noqa: E402
import torch_utils.ops.bias_act as _ba          # noqa: E402
if os.environ.get('LVG_REF_PLUGINS') == '1':
    sys.path.append(ROOT)
    from oracle import ref_cuda
    ref_cuda.patch_custom_ops(_co)
from model import generator_lres, discriminator_lres, generator_sres, discriminator_sres   # noqa: E402

if os.environ.get('LVG_REF_PLUGINS') != '1':
    # this repository's ops resolved: also hand the model files' F.conv3d / F.conv1d to the tensor-core engine (the files stay as they are)
    from torch_utils.ops import conv_nd
    patched = conv_nd.install_functional(generator_lres, discriminator_lres)
    assert len(patched) == 2
torch.backends.cudnn.allow_tf32 = False          # train_lres.py:269-270, train_sres.py
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device('cuda')
out = {'where': {'bias_act': _ba.__file__, 'plugin': type(_ba._plugin).__name__ if _ba._init() else None,
                 'conv3d': type(generator_lres.F).__name__}}


def grads(net):
    return torch.cat([p.grad.flatten().float() for p in net.parameters() if p.grad is not None])


torch.manual_seed(0)
G = generator_lres.VideoGenerator(out_height=36, out_width=64, num_fp16_layers=0, temporal_padding=8, temporal_emb_dim=1024).to(dev)
torch.manual_seed(1)
v = G(2, 32)
out['lres_G'] = v.detach().cpu()
tgt = torch.randn(v.shape, generator=torch.Generator().manual_seed(11)).to(dev)
(v * tgt).mean().backward()
out['lres_G_grad'] = grads(G).cpu()
del G, v

torch.manual_seed(2)
D = discriminator_lres.VideoDiscriminator(seq_length=32, max_edge=64, num_fp16_res=0).to(dev)
x = (torch.rand(2, 3, 32, 36, 64, generator=torch.Generator().manual_seed(12)) * 2 - 1).to(dev).requires_grad_(True)
logits = D(x)
out['lres_D'] = logits.detach().cpu()
# R1-style: gradient penalty needs double backward through bias_act / upfirdn2d (video_gan_lres.py:178-199)
gx, = torch.autograd.grad(logits.sum(), [x], create_graph=True)
(gx.square().sum() + torch.nn.functional.softplus(logits).mean()).backward()
out['lres_D_r1_gx'] = gx.detach().cpu()
out['lres_D_grad'] = grads(D).cpu()
del D, x, logits, gx

FP16_RES = 0 if os.environ.get('LVG_FP32') == '1' else 4      # LVG_FP32=1: every layer in fp32 (the yardstick of the fp16 runs)
torch.manual_seed(3)
S = generator_sres.VideoGenerator(hr_height=144, hr_width=256, lr_height=36, lr_width=64, temporal_context=4, num_fp16_res=FP16_RES,
                                  fourfeats=False).to(dev)
lr = (torch.rand(2, 3, 2 + 8, 36, 64, generator=torch.Generator().manual_seed(13)) * 2 - 1).to(dev)
hr = S(lr)
out['sres_G'] = hr.detach().float().cpu()
tgt = torch.randn(hr.shape, generator=torch.Generator().manual_seed(14)).to(dev)
(hr.float() * tgt).sum().backward()        # O(1) gradients: fp16 layers do not underflow (no loss scaling in the reference)
out['sres_G_grad'] = grads(S).cpu()
del S

torch.manual_seed(4)
SD = discriminator_sres.VideoDiscriminator(channels=3, seq_length=2, lr_height=36, lr_width=64, hr_height=144, hr_width=256,
                                           num_fp16_res=FP16_RES).to(dev)
# (an input of its own, so that the discriminator comparison does not inherit the generator's rounding)
hrv = (torch.rand(2, 3, 2, 144, 256, generator=torch.Generator().manual_seed(15)) * 2 - 1).to(dev).requires_grad_(True)
logits = SD(lr[:, :, 4:-4], hrv)
out['sres_D'] = logits.detach().float().cpu()
(torch.nn.functional.softplus(logits.float()).mean() * 4096.0).backward()   # scaled: the fp16 blocks' gradients stay normal numbers
out['sres_D_grad'] = grads(SD).cpu()
out['sres_D_gx'] = hrv.grad.cpu()
torch.save(out, sys.argv[1])
