"""Generate tests/golden/*.npz from the REFERENCE implementation and pin the oracle to it.

Run in the authoring container only (needs the read-only reference checkout):

    python oracle/pin_against_reference.py [--reference /root/reference]

For every case it imports the reference's own pure-PyTorch ops
(torch_utils/ops/{bias_act,upfirdn2d,filtered_lrelu,conv2d_resample,fma}.py,
``impl='ref'`` on CPU tensors), evaluates outputs and -- through torch autograd on
those ``_ref`` functions -- first and second order gradients, stores inputs and
results as small float32 arrays, and asserts that oracle/oracle.py reproduces
every stored forward result. The fixtures travel to the GPU box; the reference
does not.

The reference has no tests or golden vectors of its own, so
these generated vectors are the pin.
"""
import argparse
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def kaiser_lowpass(taps, cutoff, width, fs):
    import scipy.signal
    return scipy.signal.firwin(numtaps=taps, cutoff=cutoff, width=width, fs=fs).astype(np.float32)


def rnd(gen, *shape, scale=1.0):
    return (torch.randn(*shape, generator=gen) * scale).to(torch.float32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reference', default='/root/reference')
    args = ap.parse_args()
    sys.path.insert(0, args.reference)
    sys.path.insert(0, ROOT)
    warnings.filterwarnings('ignore')
    from torch_utils.ops import bias_act as R_ba, upfirdn2d as R_up, filtered_lrelu as R_fl, conv2d_resample as R_cr, fma as R_fma
    assert R_ba.__file__.startswith(args.reference), R_ba.__file__
    from oracle import oracle as orc
    os.makedirs(GOLDEN, exist_ok=True)
    gen = torch.Generator().manual_seed(20260924)
    worst = 0.0

    def close(name, ref, got, tol=2e-6):
        nonlocal worst
        ref, got = np.asarray(ref, np.float64), np.asarray(got, np.float64)
        assert ref.shape == got.shape, (name, ref.shape, got.shape)
        err = np.abs(ref - got).max() / max(1.0, np.abs(ref).max())
        worst = max(worst, err)
        assert err <= tol, f'oracle disagrees with the reference on {name}: {err:.3e}'

    # ------------------------------------------------------------------ bias_act
    out = {}
    cases = []
    for act in R_ba.activation_funcs:
        for clamp in (None, 0.7):
            cases.append((act, clamp))
    for act, clamp in cases:
        spec = R_ba.activation_funcs[act]
        tag = f'{act}_c{"n" if clamp is None else "y"}'
        x = rnd(gen, 2, 5, 3, 4, scale=1.5).requires_grad_(True)
        b = rnd(gen, 5).requires_grad_(True)
        y = R_ba.bias_act(x, b, dim=1, act=act, clamp=clamp, impl='ref')
        dy = rnd(gen, *y.shape).requires_grad_(True)
        dx, db = torch.autograd.grad(y, [x, b], dy, create_graph=True)
        out[f'{tag}/x'], out[f'{tag}/b'], out[f'{tag}/y'], out[f'{tag}/dy'] = (t.detach().numpy() for t in (x, b, y, dy))
        out[f'{tag}/dx'], out[f'{tag}/db'] = dx.detach().numpy(), db.detach().numpy()
        # second order: J = <dx, v>; d J / d dy and d J / d x
        v = rnd(gen, *x.shape)
        out[f'{tag}/v'] = v.numpy()
        if dx.requires_grad:
            g_dy, g_x = torch.autograd.grad((dx * v).sum(), [dy, x], allow_unused=True)
            out[f'{tag}/ddy'] = g_dy.numpy()
            out[f'{tag}/ddx'] = (g_x if g_x is not None else torch.zeros_like(x)).numpy()
        close(f'bias_act {tag}', out[f'{tag}/y'], orc.bias_act(out[f'{tag}/x'], out[f'{tag}/b'], 1, act, clamp=clamp))
        # oracle gradient formula (through the saved y, or x for swish) against autograd of the reference
        keep_x = spec.ref == 'x'
        o_dx = orc.bias_act_grad(out[f'{tag}/dy'], x=out[f'{tag}/x'] if keep_x else None, b=out[f'{tag}/b'] if keep_x else None,
                                 y=None if (keep_x or spec.ref == '') else out[f'{tag}/y'], dim=1, act=act, clamp=clamp, order=1)
        if not (act == 'linear' and clamp is not None):   # linear keeps no y: the reference does not mask it (quirk)
            close(f'bias_act grad {tag}', out[f'{tag}/dx'], o_dx, tol=5e-6)
    # 2-D case, bias along the last dim (fully connected layers)
    x = rnd(gen, 7, 12)
    b = rnd(gen, 12)
    out['fc/x'], out['fc/b'] = x.numpy(), b.numpy()
    out['fc/y'] = R_ba.bias_act(x, b, dim=1, act='lrelu', impl='ref').numpy()
    close('bias_act fc', out['fc/y'], orc.bias_act(out['fc/x'], out['fc/b'], 1, 'lrelu'))
    np.savez_compressed(os.path.join(GOLDEN, 'bias_act.npz'), **out)

    # ------------------------------------------------------------------ upfirdn2d
    out = {}
    f4 = R_up.setup_filter([1, 3, 3, 1])                 # [4, 4] outer product (D resampling, U8)
    f4s = R_up.setup_filter([1, 3, 3, 1], separable=True)
    lin = torch.tensor([1, 3, 3, 1], dtype=torch.float32) / 8   # temporal linear up-sampling taps (U2)
    k12 = torch.from_numpy(kaiser_lowpass(12, 0.45, 0.3, 2.0))
    k24 = torch.from_numpy(kaiser_lowpass(24, 0.22, 0.15, 2.0))
    k8 = torch.from_numpy(kaiser_lowpass(8, 0.2, 0.2, 2.0))
    sym6 = torch.tensor([0.015404109327027373, 0.0034907120842174702, -0.11799011114819057, -0.048311742585633,
                         0.4910559419267466, 0.787641141030194, 0.3379294217276218, -0.07263752278646252,
                         -0.021060292512300564, 0.04472490177066578, 0.0017677118642428036, -0.007800708325034148])
    up_cases = {
        # name: (x shape, filter, kwargs)  -- scaled-down versions of SURVEY.md Appendix A signatures U1..U9
        'U1_tdown_k12': ((2, 6, 20, 1), k12[:, None], dict(down=[1, 2], padding=[0, 0, 5, 5])),
        'U2_tup_lin':   ((2, 4, 6, 12), lin[:, None], dict(up=[1, 2], padding=[0, 0, 2, 1], gain=2)),
        'U3_sup_bil':   ((1, 10, 5, 8), f4s, dict(up=2, padding=[2, 1, 2, 1], gain=4)),
        'U3_sup_tiny':  ((2, 16, 3, 4), f4s, dict(up=2, padding=[2, 1, 2, 1], gain=4)),
        'U4_sdown':     ((1, 6, 16, 16), f4s, dict(down=2, padding=[1, 1, 1, 1])),
        'U5_tdown':     ((2, 3, 8, 16), lin[:, None] * 2, dict(down=[1, 2], padding=[0, 0, 1, 1])),
        'U6_k24_down4': ((1, 3, 30, 30), k24, dict(down=4, padding=[10, 10, 10, 10])),
        'U6_k12_down2': ((1, 3, 22, 24), k12, dict(down=2, padding=[5, 5, 5, 5])),
        'U6_k12_up2':   ((1, 3, 9, 10), k12, dict(up=2, padding=[4, 3, 4, 3], gain=4)),
        'U6_k24_up4':   ((1, 3, 7, 8), k24, dict(up=4, padding=[9, 6, 9, 6], gain=16)),
        'U7_k8_up4':    ((2, 3, 6, 9), k8, dict(up=4, padding=[5, 2, 5, 2], gain=16)),
        'U8_full_pad':  ((2, 5, 10, 12), f4, dict(padding=[2, 2, 2, 2])),
        'U8_full_down': ((2, 5, 10, 12), f4, dict(down=2, padding=[1, 1, 1, 1])),
        'U9_sym6_up':   ((1, 3, 12, 12), sym6, dict(up=2, padding=[-6, -6, -6, -6], flip_filter=True, gain=4)),
        'U9_sym6_down': ((1, 3, 24, 24), sym6, dict(down=2, padding=[-6, -6, -6, -6], flip_filter=True)),
        'mixed_updown': ((2, 2, 7, 9), f4, dict(up=[3, 2], down=[2, 3], padding=[3, -1, 0, 2])),
        'crop_only':    ((1, 2, 9, 9), None, dict(padding=[-2, -1, -1, -3])),
        'full_5x3':     ((1, 3, 8, 9), torch.randn(5, 3, generator=gen), dict(up=2, down=1, padding=[1, 2, 3, 0], flip_filter=True)),
    }
    for name, (shape, f, kw) in up_cases.items():
        x = rnd(gen, *shape).requires_grad_(True)
        y = R_up.upfirdn2d(x, f, impl='ref', **kw)
        dy = rnd(gen, *y.shape)
        dx, = torch.autograd.grad(y, [x], dy)
        out[f'{name}/x'], out[f'{name}/y'], out[f'{name}/dy'], out[f'{name}/dx'] = x.detach().numpy(), y.detach().numpy(), dy.numpy(), dx.numpy()
        if f is not None:
            out[f'{name}/f'] = f.numpy()
        fn = None if f is None else f.numpy()
        close(f'upfirdn2d {name}', out[f'{name}/y'], orc.upfirdn2d(out[f'{name}/x'], fn, **kw))
        close(f'upfirdn2d adj {name}', out[f'{name}/dx'], orc.upfirdn2d_adjoint(out[f'{name}/dy'], fn, shape, **kw), tol=5e-6)
    out['__cases__'] = np.array(repr({k: (v[0], v[2]) for k, v in up_cases.items()}))
    np.savez_compressed(os.path.join(GOLDEN, 'upfirdn2d.npz'), **out)

    # ------------------------------------------------------------------ filtered_lrelu
    out = {}
    fl_cases = {
        # the four kernel configurations the sres generator hits + edge cases
        'up2_down2_k12':   ((2, 3, 9, 11), k12, k12, dict(up=2, down=2, padding=[9, 8, 9, 8])),
        'up4_down2_k24':   ((1, 3, 8, 9), k24, k12, dict(up=4, down=2, padding=[20, 19, 20, 19])),
        'up4_down2_crop':  ((1, 2, 14, 16), k24, k12, dict(up=4, down=2, padding=[-6, -9, -6, -9])),
        'up2_down4_k24':   ((1, 3, 20, 22), k12, k24, dict(up=2, down=4, padding=[16, 15, 16, 15])),
        'up2_down2_crop':  ((1, 2, 28, 30), k12, k12, dict(up=2, down=2, padding=[-11, -12, -11, -12], clamp=0.3)),
        'torgb_1x1':       ((2, 3, 6, 7), None, None, dict(up=1, down=1, padding=0, gain=1.0, slope=1.0, clamp=0.5)),
        'clamp_active':    ((1, 3, 9, 11), k12, k12, dict(up=2, down=2, padding=[9, 8, 9, 8], clamp=0.05)),
        'flip_full_fu':    ((1, 2, 6, 7), torch.randn(4, 4, generator=gen) / 4, k12, dict(up=2, down=2, padding=[7, 6, 7, 6], flip_filter=True)),
        'up1_down2':       ((1, 2, 18, 20), None, k12, dict(up=1, down=2, padding=[5, 5, 5, 5])),
    }
    for name, (shape, fu, fd, kw) in fl_cases.items():
        scale = 100.0 if name == 'clamp_active' else 1.0
        x = rnd(gen, *shape, scale=scale).requires_grad_(True)
        b = rnd(gen, shape[1]).requires_grad_(True)
        y = R_fl.filtered_lrelu(x, fu=fu, fd=fd, b=b, impl='ref', **kw)
        dy = rnd(gen, *y.shape)
        dx, db = torch.autograd.grad(y, [x, b], dy)
        for k, t in (('x', x), ('b', b), ('y', y), ('dy', dy), ('dx', dx), ('db', db)):
            out[f'{name}/{k}'] = t.detach().numpy()
        if fu is not None:
            out[f'{name}/fu'] = fu.numpy()
        if fd is not None:
            out[f'{name}/fd'] = fd.numpy()
        got = orc.filtered_lrelu(out[f'{name}/x'], None if fu is None else fu.numpy(), None if fd is None else fd.numpy(),
                                 out[f'{name}/b'], **kw)
        close(f'filtered_lrelu {name}', out[f'{name}/y'], got, tol=5e-6)
    out['__cases__'] = np.array(repr({k: (v[0], v[3]) for k, v in fl_cases.items()}))
    np.savez_compressed(os.path.join(GOLDEN, 'filtered_lrelu.npz'), **out)

    # ------------------------------------------------------------------ conv2d_resample / conv boundary / fma
    out = {}
    cr_cases = {
        'plain_3x3':      ((2, 4, 9, 10), (6, 4, 3, 3), dict(padding=1)),
        'down2_3x3':      ((2, 4, 12, 12), (6, 4, 3, 3), dict(f=f4, down=2, padding=1)),
        'skip_1x1_down2': ((2, 4, 12, 12), (5, 4, 1, 1), dict(f=f4, down=2)),
        'fromrgb_1x1':    ((2, 6, 8, 8), (4, 6, 1, 1), dict()),
        'up2_3x3':        ((1, 3, 6, 7), (4, 3, 3, 3), dict(f=f4, up=2, padding=1)),
        'up2_1x1':        ((1, 3, 6, 7), (4, 3, 1, 1), dict(f=f4, up=2)),
        'grouped_mod':    ((1, 6, 7, 8), (8, 3, 3, 3), dict(padding=2, groups=2)),
        'noflip':         ((1, 2, 7, 8), (3, 2, 3, 3), dict(padding=[1, 0, 2, 1], flip_weight=False)),
    }
    for name, (xs, ws, kw) in cr_cases.items():
        x = rnd(gen, *xs).requires_grad_(True)
        w = (rnd(gen, *ws) / np.sqrt(ws[1] * ws[2] * ws[3])).requires_grad_(True)
        y = R_cr.conv2d_resample(x, w, **kw)
        dy = rnd(gen, *y.shape)
        dx, dw = torch.autograd.grad(y, [x, w], dy)
        for k, t in (('x', x), ('w', w), ('y', y), ('dy', dy), ('dx', dx), ('dw', dw)):
            out[f'{name}/{k}'] = t.detach().numpy()
    out['f4'] = f4.numpy()
    close('conv2d grouped', out['grouped_mod/y'], orc.conv2d(out['grouped_mod/x'], out['grouped_mod/w'], padding=2, groups=2), tol=5e-6)
    close('conv2d plain', out['plain_3x3/y'], orc.conv2d(out['plain_3x3/x'], out['plain_3x3/w'], padding=1), tol=5e-6)
    a, b, c = rnd(gen, 3, 1, 5).requires_grad_(True), rnd(gen, 4, 5).requires_grad_(True), rnd(gen, 5).requires_grad_(True)
    o = R_fma.fma(a, b, c)
    do = rnd(gen, *o.shape)
    da, db_, dc = torch.autograd.grad(o, [a, b, c], do)
    for k, t in (('a', a), ('b', b), ('c', c), ('o', o), ('do', do), ('da', da), ('db', db_), ('dc', dc)):
        out[f'fma/{k}'] = t.detach().numpy()
    close('fma', out['fma/o'], orc.fma(out['fma/a'], out['fma/b'], out['fma/c']))
    out['__cases__'] = np.array(repr({k: (v[0], v[1], {kk: vv for kk, vv in v[2].items() if kk != 'f'}, 'f' in v[2]) for k, v in cr_cases.items()}))
    np.savez_compressed(os.path.join(GOLDEN, 'conv.npz'), **out)

    sizes = {f: os.path.getsize(os.path.join(GOLDEN, f)) for f in sorted(os.listdir(GOLDEN)) if f.endswith('.npz')}
    print('golden fixtures written:', sizes)
    print(f'oracle vs reference worst relative error: {worst:.3e}')


if __name__ == '__main__':
    main()
