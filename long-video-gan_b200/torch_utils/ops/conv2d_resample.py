"""2-D convolution with optional FIR up/down-sampling (``torch_utils.ops.conv2d_resample``).

Signature and padding conventions as in the reference (conv2d_resample.py:45-141). The operator is a chain of
``upfirdn2d`` and ``conv2d_gradfix.conv2d`` / ``conv_transpose2d`` stages; which chain is used for which case is part
of the numerics (fp16 rounding points) and is kept:

    1x1 kernel, down only : FIR-decimate, then convolve
    1x1 kernel, up only   : convolve, then zero-insert + FIR
    down only             : FIR (no decimation), then strided convolution
    up (optionally down)  : transposed strided convolution, then FIR (then FIR-decimate)
    no resampling         : plain convolution when the padding is symmetric and >= 0
    otherwise             : up-FIR, convolve, FIR-decimate

Here the case analysis is separated from the execution: ``_plan`` turns the arguments into a short list of stage
descriptions, ``_run`` executes them.
"""
import torch

from . import conv2d_gradfix
from . import upfirdn2d
from .upfirdn2d import _get_filter_size, _parse_padding


def _get_weight_shape(w):
    return [int(sz) for sz in w.shape]


def _conv2d_wrapper(x, w, stride=1, padding=0, groups=1, transpose=False, flip_weight=True):
    """conv2d is a correlation; ``flip_weight=False`` asks for a true convolution (mirrored kernel)."""
    kh, kw = _get_weight_shape(w)[2:]
    if (kh > 1 or kw > 1) and not flip_weight:
        w = w.flip([2, 3])
    conv = conv2d_gradfix.conv_transpose2d if transpose else conv2d_gradfix.conv2d
    return conv(x, w, stride=stride, padding=padding, groups=groups)


def _footprint(taps, factor, upsampling):
    """Padding (before, after) that centres a `taps`-tap resampling filter for the given factor."""
    if upsampling:
        return (taps + factor - 1) // 2, (taps - factor) // 2
    return (taps - factor + 1) // 2, (taps - factor) // 2


def _to_transposed_layout(w, groups):
    """[Cout, Cin/g, kh, kw] -> the [Cin, Cout/g, kh, kw] layout conv_transpose2d expects."""
    if groups == 1:
        return w.transpose(0, 1)
    cout, cin_g, kh, kw = _get_weight_shape(w)
    w = w.reshape(groups, cout // groups, cin_g, kh, kw).transpose(1, 2)
    return w.reshape(groups * cin_g, cout // groups, kh, kw)


def _plan(kw, kh, fw, fh, up, down, padding):
    """Stage list for one call. A stage is ('fir', kwargs of upfirdn2d) or ('conv', kwargs of _conv2d_wrapper);
    'conv' stages may carry 'transposed_weight': True (the weight is re-laid-out before the call)."""
    px0, px1, py0, py1 = _parse_padding(padding)
    if up > 1:          # fold the resampling filters' own footprint into the padding
        (ax, bx), (ay, by) = _footprint(fw, up, True), _footprint(fh, up, True)
        px0, px1, py0, py1 = px0 + ax, px1 + bx, py0 + ay, py1 + by
    if down > 1:
        (ax, bx), (ay, by) = _footprint(fw, down, False), _footprint(fh, down, False)
        px0, px1, py0, py1 = px0 + ax, px1 + bx, py0 + ay, py1 + by
    pad = [px0, px1, py0, py1]
    pointwise = kw == 1 and kh == 1
    gain = up ** 2

    if pointwise and up == 1 and down > 1:
        return [('fir', dict(down=down, padding=pad)), ('conv', dict())]
    if pointwise and down == 1 and up > 1:
        return [('conv', dict()), ('fir', dict(up=up, padding=pad, gain=gain))]
    if up == 1 and down > 1:
        return [('fir', dict(padding=pad)), ('conv', dict(stride=down))]
    if up > 1:
        # the transposed convolution absorbs as much of the (now possibly negative) padding as it can
        px0, px1, py0, py1 = px0 - (kw - 1), px1 - (kw - up), py0 - (kh - 1), py1 - (kh - up)
        pxt, pyt = max(min(-px0, -px1), 0), max(min(-py0, -py1), 0)
        stages = [('conv', dict(stride=up, padding=[pyt, pxt], transpose=True, transposed_weight=True)),
                  ('fir', dict(padding=[px0 + pxt, px1 + pxt, py0 + pyt, py1 + pyt], gain=gain))]
        if down > 1:
            stages.append(('fir', dict(down=down)))
        return stages
    if down == 1 and px0 == px1 and py0 == py1 and min(px0, py0) >= 0:
        return [('conv', dict(padding=[py0, px0]))]
    stages = [('fir', dict(up=up, padding=pad, gain=gain, no_filter=(up == 1))), ('conv', dict())]
    if down > 1:
        stages.append(('fir', dict(down=down)))
    return stages


def _run(stages, x, w, f, groups, flip_weight, flip_filter):
    for kind, kwargs in stages:
        kwargs = dict(kwargs)
        if kind == 'fir':
            filt = None if kwargs.pop('no_filter', False) else f
            x = upfirdn2d.upfirdn2d(x=x, f=filt, flip_filter=flip_filter, **kwargs)
        else:
            transposed = kwargs.pop('transposed_weight', False)
            weight = _to_transposed_layout(w, groups) if transposed else w
            # the transposed convolution mirrors the kernel once more, hence the inverted flag
            x = _conv2d_wrapper(x=x, w=weight, groups=groups, flip_weight=(flip_weight != transposed), **kwargs)
    return x


def conv2d_resample(x, w, f=None, up=1, down=1, padding=0, groups=1, flip_weight=True, flip_filter=False):
    """x ``[N, Cin, H, W]``, w ``[Cout, Cin/groups, kh, kw]`` (same dtype), f: filter from
    ``upfirdn2d.setup_filter`` or None. `padding` is relative to the up-sampled image and applied once."""
    assert isinstance(x, torch.Tensor) and (x.ndim == 4)
    assert isinstance(w, torch.Tensor) and (w.ndim == 4) and (w.dtype == x.dtype)
    assert f is None or (isinstance(f, torch.Tensor) and f.ndim in [1, 2] and f.dtype == torch.float32)
    assert isinstance(up, int) and (up >= 1)
    assert isinstance(down, int) and (down >= 1)
    assert isinstance(groups, int) and (groups >= 1)
    kh, kw = _get_weight_shape(w)[2:]
    fw, fh = _get_filter_size(f)
    return _run(_plan(kw, kh, fw, fh, up, down, padding), x, w, f, groups, flip_weight, flip_filter)
