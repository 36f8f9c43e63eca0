"""Helper of test_dropin_reference.py (run as a subprocess): builds reference networks on CPU with whatever
`torch_utils` the PYTHONPATH resolves to, runs one forward pass each and saves the outputs."""
import sys
import types
import warnings

import torch

warnings.filterwarnings('ignore')
sys.modules.setdefault('imageio', types.ModuleType('imageio'))

import torch_utils.ops.bias_act as _ba          # noqa: E402
import torch_utils.ops.upfirdn2d as _up         # noqa: E402
import torch_utils.misc as _misc                # noqa: E402
import dnnlib as _dnnlib                        # noqa: E402
from model import generator_lres, discriminator_lres, generator_sres   # noqa: E402

out = {'where': {'bias_act': _ba.__file__, 'upfirdn2d': _up.__file__, 'misc': _misc.__file__, 'dnnlib': _dnnlib.__file__}}
torch.manual_seed(0)
G = generator_lres.VideoGenerator(out_height=36, out_width=64, num_fp16_layers=0, temporal_padding=8, temporal_emb_dim=1024)
with torch.no_grad():
    torch.manual_seed(1)
    out['lres_G'] = G(1, 32)
del G
torch.manual_seed(2)
D = discriminator_lres.VideoDiscriminator(seq_length=32, max_edge=64, num_fp16_res=0)
with torch.no_grad():
    out['lres_D'] = D(torch.rand(1, 3, 32, 36, 64) * 2 - 1)
del D
torch.manual_seed(3)
S = generator_sres.VideoGenerator(hr_height=144, hr_width=256, lr_height=36, lr_width=64, temporal_context=4, num_fp16_res=0,
                                  fourfeats=False)
with torch.no_grad():
    out['sres_G'] = S(torch.rand(1, 3, 2 + 8, 36, 64) * 2 - 1)
torch.save(out, sys.argv[1])
