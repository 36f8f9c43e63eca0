"""Plugin loader for the sm_100a operator library (replaces the JIT loader).

The reference builds three pybind11 modules at first use
(``torch_utils/custom_ops.py:59-157`` -> ``torch.utils.cpp_extension.load``).
Here there is ONE ahead-of-time built C-ABI shared library, ``liblvg_ops.so``
(``include/lvg_ops.h``), opened with ``ctypes``; ``get_plugin(name)`` returns an
object exposing the same functions, argument order and return values as the
reference's pybind module of that name:

    bias_act_plugin.bias_act            bias_act.cpp:32,96
    upfirdn2d_plugin.upfirdn2d          upfirdn2d.cpp:16,104
    filtered_lrelu_plugin.filtered_lrelu / filtered_lrelu_act_   filtered_lrelu.cpp:16,213,296

Tensors are allocated here with torch; the library only sees raw device
pointers, shapes, strides and the current CUDA stream. There is no CPU
implementation behind these objects: a missing library or a non-CUDA tensor
raises.
"""
import ctypes
import os

import torch

verbosity = 'brief'  # kept for API compatibility ('none' | 'brief' | 'full')

_LIB_NAME = 'liblvg_ops.so'
_lib = None
_plugins = {}

_c_void_p = ctypes.c_void_p
_c_int = ctypes.c_int
_c_i64 = ctypes.c_int64
_c_float = ctypes.c_float
_I64x4 = ctypes.c_int64 * 4
_I64x6 = ctypes.c_int64 * 6

_DTYPE_CODE = {torch.float32: 0, torch.float16: 1, torch.float64: 2}

# every exported symbol of include/lvg_ops.h with its C signature
_SIGNATURES = {
    'lvg_abi_version': (_c_int, []),
    'lvg_last_error': (ctypes.c_char_p, []),
    'lvg_build_info': (ctypes.c_char_p, []),
    'lvg_launch_count': (_c_i64, []),
    'lvg_grad_postprocess': (_c_int, [_c_void_p, _c_i64, _c_float, _c_float, _c_void_p]),
    'lvg_adam_step': (_c_int, [_c_void_p] * 5 + [_c_i64] + [_c_float] * 4 + [_c_i64, _c_float, _c_float, _c_int, _c_float, _c_void_p]),
    'lvg_lerp': (_c_int, [_c_void_p, _c_void_p, _c_i64, _c_float, _c_void_p]),
    'lvg_bias_act': (_c_int, [_c_void_p] * 6 + [_c_int, _c_i64, _c_i64, _c_i64, _c_int, _c_int, _c_float, _c_float, _c_float, _c_void_p]),
    'lvg_bias_act_grad_db': (_c_int, [_c_void_p] * 6 + [_c_int, _c_i64, _c_i64, _c_i64, _c_int, _c_float, _c_float, _c_float, _c_void_p]),
    'lvg_bias_act_fwd_codes': (_c_int, [_c_void_p] * 4 + [_c_int, _c_i64, _c_i64, _c_i64, _c_int, _c_float, _c_float, _c_float, _c_void_p]),
    'lvg_bias_act_codes_bytes': (_c_i64, [_c_int, _c_i64]),
    'lvg_bias_act_bwd_codes': (_c_int, [_c_void_p] * 4 + [_c_int, _c_i64, _c_i64, _c_i64, _c_int, _c_float, _c_float, _c_float, _c_void_p]),
    'lvg_upfirdn2d': (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _I64x4, _I64x4, _I64x4, _I64x4, _c_int, _c_int, _c_i64, _c_i64]
                      + [_c_int] * 7 + [_c_float, _c_void_p]),
    'lvg_upfirdn2d_sep': (_c_int, [_c_void_p] * 4 + [_c_int, _I64x4, _I64x4, _I64x4, _I64x4] + [_c_int] * 9 + [_c_float, _c_void_p]),
    'lvg_filtered_lrelu': (_c_int, [_c_void_p] * 7 + [_c_int, _I64x4, _I64x4, _I64x4, _I64x4] + [_c_int] * 12
                           + [_c_float, _c_float, _c_float, _c_int, _c_int, _c_void_p]),
    'lvg_filtered_lrelu_supported': (_c_int, [_c_int] * 7),
    'lvg_filtered_lrelu_act': (_c_int, [_c_void_p] * 3 + [_c_int, _I64x4, _I64x4] + [_c_int] * 4 + [_c_float] * 3 + [_c_int, _c_void_p]),
    'lvg_fma': (_c_int, [_c_void_p] * 4 + [_c_int, _c_int, _I64x6, _I64x6, _I64x6, _I64x6, _c_void_p]),
    'lvg_conv2d_fprop': (_c_int, [_c_void_p] * 3 + [_c_int] * 12 + [_c_void_p, _c_i64, _c_void_p]),
    'lvg_conv2d_fprop_workspace': (_c_i64, [_c_int] * 12),
    'lvg_conv2d_dgrad': (_c_int, [_c_void_p] * 3 + [_c_int] * 12 + [_c_void_p, _c_i64, _c_void_p]),
    'lvg_conv2d_wgrad': (_c_int, [_c_void_p] * 3 + [_c_int] * 12 + [_c_void_p]),
    'lvg_fir1d_depthwise_workspace': (_c_i64, [_c_int]),
    'lvg_fir1d_depthwise': (_c_int, [_c_void_p] * 3 + [_c_int] * 4 + [_c_void_p, _c_i64, _c_void_p]),
    'lvg_convnd_workspace': (_c_i64, [_c_int] * 14),
    'lvg_convnd_fprop': (_c_int, [_c_void_p] * 3 + [_c_int] * 15 + [_c_void_p, _c_int, _c_float, _c_float, _c_float, _c_void_p, _c_i64, _c_void_p]),
    'lvg_convnd_dgrad': (_c_int, [_c_void_p] * 3 + [_c_int] * 15 + [_c_void_p, _c_i64, _c_void_p]),
    'lvg_convnd_wgrad_workspace': (_c_i64, [_c_int] * 14),
    'lvg_convnd_wgrad': (_c_int, [_c_void_p] * 3 + [_c_int] * 15 + [_c_void_p, _c_i64, _c_void_p]),
    'lvg_convnd_plan': (_c_int, [_c_int] * 16 + [ctypes.POINTER(_c_int), _c_int]),
    'lvg_convnd_wgrad_plan': (_c_int, [_c_int] * 14 + [ctypes.POINTER(_c_int), _c_int]),
    'lvg_convnd_backward_workspace': (_c_i64, [_c_int] * 14),
    'lvg_convnd_backward': (_c_int, [_c_void_p] * 5 + [_c_int] * 15 + [_c_void_p, _c_i64, _c_void_p]),
}

LVG_UNSUPPORTED = -1


def library_path():
    """Absolute path where liblvg_ops.so is expected (next to this package)."""
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), _LIB_NAME)


def load_library():
    """Open liblvg_ops.so once and declare every entry point. Raises if it is missing."""
    global _lib
    if _lib is None:
        path = os.environ.get('LVG_OPS_LIBRARY', library_path())
        if not os.path.isfile(path):
            raise RuntimeError(
                f'{_LIB_NAME} not found at {path}: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                f'or `make -C long-video-gan_b200/csrc`. There is no fallback for CUDA tensors.')
        lib = ctypes.CDLL(path)
        for name, (restype, argtypes) in _SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here = header / library mismatch
            fn.restype = restype
            fn.argtypes = argtypes
        if lib.lvg_abi_version() != 1:
            raise RuntimeError(f'{path}: unexpected ABI version {lib.lvg_abi_version()}')
        _lib = lib
    return _lib


def exported_symbols():
    return sorted(_SIGNATURES)


def launch_count():
    """Kernels launched by liblvg_ops in this process so far."""
    return int(load_library().lvg_launch_count())


def _check(rc, what):
    if rc > 0:
        raise RuntimeError(f'{what}: {_lib.lvg_last_error().decode()}')
    return rc


def _ptr(t):
    """Device pointer of a tensor, or NULL for None / empty ("absent" in the reference API)."""
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


_raw_stream = getattr(torch._C, '_cuda_getCurrentRawStream', None)


def _stream(t):
    """cudaStream_t of torch's current stream on t's device (the raw-handle accessor skips building a Stream object)."""
    if _raw_stream is not None and t.device.index is not None:
        return _raw_stream(t.device.index)
    return torch.cuda.current_stream(t.device).cuda_stream


def _i4(seq):
    return _I64x4(*[int(v) for v in seq])


def _dtype_code(t, what):
    try:
        return _DTYPE_CODE[t.dtype]
    except KeyError:
        raise RuntimeError(f'{what}: unsupported dtype {t.dtype}') from None


def _absent(t):
    return t is None or t.numel() == 0


def _same_layout(a, b):
    return a.shape == b.shape and a.stride() == b.stride()


def _is_dense(x):
    if x.is_contiguous():
        return True
    dims = sorted((d for d in range(x.ndim) if x.shape[d] != 1), key=lambda d: x.stride(d))
    expect = 1
    for d in dims:
        if x.stride(d) != expect:
            return False
        expect *= x.shape[d]
    return True


class _DeviceGuard:
    """Makes t's device current for the duration of a launch (like OptionalCUDAGuard)."""

    def __init__(self, t):
        self.idx = t.device.index
        self.prev = None

    def __enter__(self):
        cur = torch.cuda.current_device()
        if self.idx is not None and self.idx != cur:
            self.prev = cur
            torch.cuda.set_device(self.idx)

    def __exit__(self, *exc):
        if self.prev is not None:
            torch.cuda.set_device(self.prev)


# ----------------------------------------------------------------------------

class BiasActPlugin:
    """``bias_act_plugin`` (bias_act.cpp:32-96)."""

    def __init__(self, lib):
        self._lib = lib

    def bias_act(self, x, b, xref, yref, dy, grad, dim, act, alpha, gain, clamp):
        if not x.is_cuda:
            raise RuntimeError('x must reside on CUDA device')
        code = _dtype_code(x, 'bias_act')
        for name, t in (('xref', xref), ('yref', yref), ('dy', dy)):
            if not _absent(t) and not (t.dtype == x.dtype and t.device == x.device and _same_layout(t, x)):
                raise RuntimeError(f'{name} must have the same shape, dtype, device and layout as x')
        if grad < 0:
            raise RuntimeError('grad must be non-negative')
        size_b, step_b = 1, 1
        if not _absent(b):
            if b.dtype != x.dtype or b.device != x.device:
                raise RuntimeError('b must have the same dtype and device as x')
            if b.ndim != 1:
                raise RuntimeError('b must have rank 1')
            if not (0 <= dim < x.ndim):
                raise RuntimeError('dim is out of bounds')
            if b.numel() != x.shape[dim]:
                raise RuntimeError('b has wrong number of elements')
            if not b.is_contiguous():
                raise RuntimeError('b must be contiguous')
            size_b, step_b = b.numel(), x.stride(dim)
        if not _is_dense(x):
            raise RuntimeError('x must be non-overlapping and dense')
        y = torch.empty_like(x)
        if not _same_layout(y, x):
            raise RuntimeError('y must have the same layout as x')
        with _DeviceGuard(x):
            _check(self._lib.lvg_bias_act(_ptr(x), _ptr(b), _ptr(xref), _ptr(yref), _ptr(dy), _ptr(y), code,
                                          x.numel(), size_b, max(step_b, 1), int(grad), int(act),
                                          float(alpha), float(gain), float(clamp), _stream(x)), 'bias_act')
        return y

    def bias_act_grad_db(self, dy, b, xref, yref, dim, act, alpha, gain, clamp):
        """Backward pass with the bias-gradient reduction fused in. Returns (dx, db) with db in dy.dtype,
        or None when the fused kernel does not cover the layout (bias along the contiguous dimension)."""
        if not dy.is_cuda:
            raise RuntimeError('dy must reside on CUDA device')
        code = _dtype_code(dy, 'bias_act_grad_db')
        if not _is_dense(dy):
            raise RuntimeError('dy must be non-overlapping and dense')
        for name, t in (('xref', xref), ('yref', yref)):
            if not _absent(t) and not (t.dtype == dy.dtype and _same_layout(t, dy)):
                raise RuntimeError(f'{name} must have the same shape, dtype and layout as dy')
        size_b, step_b = dy.shape[dim], max(dy.stride(dim), 1)
        dx = torch.empty_like(dy)
        db = torch.zeros([size_b], dtype=torch.float32, device=dy.device)
        with _DeviceGuard(dy):
            rc = _check(self._lib.lvg_bias_act_grad_db(_ptr(dy), _ptr(b), _ptr(xref), _ptr(yref), _ptr(dx), _ptr(db), code,
                                                       dy.numel(), size_b, step_b, int(act), float(alpha), float(gain),
                                                       float(clamp), _stream(dy)), 'bias_act_grad_db')
        if rc == LVG_UNSUPPORTED:
            return None     # layout not covered by the fused kernel: caller runs bias_act(grad=1) + sum
        return dx, db.to(dy.dtype)


    def bias_act_fwd_codes(self, x, b, dim, act, alpha, gain, clamp):
        """relu / lrelu forward that also emits 2-bit sign / clamp codes (uint8 [numel / 4]) for the backward pass.
        Returns (y, codes), or None when the kernel does not cover the call (other activations, odd sizes, fp64)."""
        if not x.is_cuda or x.dtype not in (torch.float16, torch.float32) or x.numel() == 0 or not _is_dense(x):
            return None
        size_b, step_b = 1, 1
        if not _absent(b):
            if b.dtype != x.dtype or b.device != x.device or b.ndim != 1 or b.numel() != x.shape[dim] or not b.is_contiguous():
                return None         # let the general entry point produce the reference's error message
            size_b, step_b = b.numel(), max(x.stride(dim), 1)
        y = torch.empty_like(x)
        nbytes = self._lib.lvg_bias_act_codes_bytes(_dtype_code(x, 'bias_act'), x.numel())
        if nbytes < 0:
            return None
        codes = torch.empty([int(nbytes)], dtype=torch.uint8, device=x.device)
        with _DeviceGuard(x):
            rc = _check(self._lib.lvg_bias_act_fwd_codes(_ptr(x), _ptr(b), _ptr(y), _ptr(codes), _dtype_code(x, 'bias_act'),
                                                         x.numel(), size_b, step_b, int(act), float(alpha), float(gain),
                                                         float(clamp), _stream(x)), 'bias_act_fwd_codes')
        return None if rc == LVG_UNSUPPORTED else (y, codes)

    def bias_act_bwd_codes(self, dy, codes, dim, act, alpha, gain, clamp, want_db):
        """dx (and the bias gradient when want_db and the layout allows fusing it) from dy and the forward's codes.
        dy must have the memory layout of the forward's x. Returns (dx, db or None)."""
        code = _dtype_code(dy, 'bias_act_bwd_codes')
        dx = torch.empty_like(dy)
        size_b, step_b = dy.shape[dim], max(dy.stride(dim), 1)
        pack = 8 if dy.dtype == torch.float16 else 4
        db = torch.zeros([size_b], dtype=torch.float32, device=dy.device) if (want_db and step_b % pack == 0) else None
        with _DeviceGuard(dy):
            rc = _check(self._lib.lvg_bias_act_bwd_codes(_ptr(dy), _ptr(codes), _ptr(dx), _ptr(db), code, dy.numel(), size_b, step_b,
                                                         int(act), float(alpha), float(gain), float(clamp), _stream(dy)),
                        'bias_act_bwd_codes')
        if rc == LVG_UNSUPPORTED:
            raise RuntimeError('bias_act_bwd_codes: ' + self._lib.lvg_last_error().decode())
        return dx, (db.to(dy.dtype) if db is not None else None)


class Upfirdn2dPlugin:
    """``upfirdn2d_plugin`` (upfirdn2d.cpp:16-104), plus the single-launch separable entry point."""

    def __init__(self, lib):
        self._lib = lib

    @staticmethod
    def _out_size(x, fw, fh, upx, upy, downx, downy, padx0, padx1, pady0, pady1):
        ow = (x.shape[3] * upx + padx0 + padx1 - fw + downx) // downx
        oh = (x.shape[2] * upy + pady0 + pady1 - fh + downy) // downy
        if ow < 1 or oh < 1:
            raise RuntimeError('output must be at least 1x1')
        return oh, ow

    @staticmethod
    def _alloc_like(x, oh, ow):
        fmt = torch.channels_last if (x.shape[1] > 1 and x.stride(1) == 1) else torch.contiguous_format
        return torch.empty([x.shape[0], x.shape[1], oh, ow], dtype=x.dtype, device=x.device, memory_format=fmt)

    @staticmethod
    def _validate(x, f):
        if not x.is_cuda:
            raise RuntimeError('x must reside on CUDA device')
        if f.device != x.device:
            raise RuntimeError('f must reside on the same device as x')
        if f.dtype != torch.float32:
            raise RuntimeError('f must be float32')
        if x.numel() == 0:
            raise RuntimeError('x has zero size')
        if f.numel() == 0:
            raise RuntimeError('f has zero size')
        if x.ndim != 4:
            raise RuntimeError('x must be rank 4')

    def upfirdn2d(self, x, f, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip, gain):
        self._validate(x, f)
        if f.ndim != 2:
            raise RuntimeError('f must be rank 2')
        if upx < 1 or upy < 1:
            raise RuntimeError('upsampling factor must be at least 1')
        if downx < 1 or downy < 1:
            raise RuntimeError('downsampling factor must be at least 1')
        code = _dtype_code(x, 'upfirdn2d')
        fh, fw = f.shape
        oh, ow = self._out_size(x, fw, fh, upx, upy, downx, downy, padx0, padx1, pady0, pady1)
        y = self._alloc_like(x, oh, ow)
        with _DeviceGuard(x):
            _check(self._lib.lvg_upfirdn2d(_ptr(x), _ptr(f), _ptr(y), code, _i4(x.shape), _i4(x.stride()), _i4(y.shape),
                                           _i4(y.stride()), fw, fh, f.stride(1), f.stride(0), upx, upy, downx, downy,
                                           padx0, pady0, int(bool(flip)), float(gain), _stream(x)), 'upfirdn2d')
        return y

    def upfirdn2d_sep(self, x, fx, fy, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip, gain):
        """Both passes of a separable filter in one launch (fx along W, fy along H; None = no filter
        on that axis). Returns None when the tiled kernel does not cover the configuration."""
        f_any = fx if fx is not None else fy
        self._validate(x, f_any)
        for f in (fx, fy):
            if f is not None and (f.ndim != 1 or not f.is_contiguous() or f.dtype != torch.float32):
                raise RuntimeError('separable filters must be contiguous float32 vectors')
        code = _dtype_code(x, 'upfirdn2d_sep')
        fw = fx.numel() if fx is not None else 1
        fh = fy.numel() if fy is not None else 1
        oh, ow = self._out_size(x, fw, fh, upx, upy, downx, downy, padx0, padx1, pady0, pady1)
        y = self._alloc_like(x, oh, ow)
        with _DeviceGuard(x):
            rc = _check(self._lib.lvg_upfirdn2d_sep(_ptr(x), _ptr(fx), _ptr(fy), _ptr(y), code, _i4(x.shape), _i4(x.stride()),
                                                    _i4(y.shape), _i4(y.stride()), fw, fh, upx, upy, downx, downy,
                                                    padx0, pady0, int(bool(flip)), float(gain), _stream(x)), 'upfirdn2d_sep')
        return None if rc == LVG_UNSUPPORTED else y


class FilteredLReluPlugin:
    """``filtered_lrelu_plugin`` (filtered_lrelu.cpp:16-297)."""

    def __init__(self, lib):
        self._lib = lib

    def filtered_lrelu(self, x, fu, fd, b, si, up, down, px0, px1, py0, py1, sx, sy, gain, slope, clamp, flip_filters, writeSigns):
        if not x.is_cuda:
            raise RuntimeError('x must reside on CUDA device')
        if not (fu.device == x.device and fd.device == x.device and b.device == x.device):
            raise RuntimeError('all input tensors must reside on the same device')
        if fu.dtype != torch.float32 or fd.dtype != torch.float32:
            raise RuntimeError('fu and fd must be float32')
        if b.dtype != x.dtype:
            raise RuntimeError('x and b must have the same dtype')
        if x.dtype not in (torch.float16, torch.float32):
            raise RuntimeError('x and b must be float16 or float32')
        if x.ndim != 4:
            raise RuntimeError('x must be rank 4')
        if x.numel() == 0:
            raise RuntimeError('x is empty')
        if fu.ndim not in (1, 2) or fd.ndim not in (1, 2):
            raise RuntimeError('fu and fd must be rank 1 or 2')
        if fu.numel() == 0 or fd.numel() == 0:
            raise RuntimeError('fu and fd must not be empty')
        if b.ndim != 1 or b.shape[0] != x.shape[1]:
            raise RuntimeError('b must be a vector with the same number of channels as x')
        if up < 1 or down < 1:
            raise RuntimeError('up and down must be at least 1')
        code = _dtype_code(x, 'filtered_lrelu')

        fu_w, fu_h = fu.shape[-1], (fu.shape[0] if fu.ndim == 2 else 0)   # height 0 marks a separable filter
        fd_w, fd_h = fd.shape[-1], (fd.shape[0] if fd.ndim == 2 else 0)
        if self._lib.lvg_filtered_lrelu_supported(code, fu_w, fu_h, fd_w, fd_h, up, down) != 0:
            return None, None, -1   # same contract as the reference: caller runs the generic path

        fut_w, fut_h = fu.shape[-1] - 1, fu.shape[0] - 1
        fdt_w, fdt_h = fd.shape[-1] - 1, fd.shape[0] - 1
        cw = x.shape[3] * up + (px0 + px1) - fut_w      # logical size of the up-sampled buffer
        ch = x.shape[2] * up + (py0 + py1) - fut_h
        if not (cw > fdt_w and ch > fdt_h):
            raise RuntimeError('upsampled buffer must be at least the size of downsampling filter')
        yw = (cw - fdt_w + (down - 1)) // down
        yh = (ch - fdt_h + (down - 1)) // down
        if yw < 1 or yh < 1:
            raise RuntimeError('output must be at least 1x1')
        fmt = torch.channels_last if (x.shape[1] > 1 and x.stride(1) == 1) else torch.contiguous_format
        y = torch.empty([x.shape[0], x.shape[1], yh, yw], dtype=x.dtype, device=x.device, memory_format=fmt)

        so = None
        s = si if not _absent(si) else None
        read_signs = s is not None
        if writeSigns:
            if read_signs:
                raise RuntimeError('cannot read and write signs in the same call')
            sw_active = yw * down - (down - 1) + fdt_w
            sh = yh * down - (down - 1) + fdt_h
            sw = (sw_active + 15) & ~15
            s = so = torch.empty([x.shape[0], x.shape[1], sh, sw >> 2], dtype=torch.uint8, device=x.device)
        if s is not None:
            if not (s.is_contiguous() and s.dtype == torch.uint8 and s.device == x.device and s.ndim == 4
                    and s.shape[0] == x.shape[0] and s.shape[1] == x.shape[1]):
                raise RuntimeError('signs must be a contiguous uint8 [N, C, H, W/4] tensor on the same device as x')
        s_h, s_wb = (s.shape[2], s.shape[3]) if s is not None else (0, 0)
        fu_c, fd_c, b_c = fu.contiguous(), fd.contiguous(), b.contiguous()
        with _DeviceGuard(x):
            rc = _check(self._lib.lvg_filtered_lrelu(
                _ptr(x), _ptr(fu_c), _ptr(fd_c), _ptr(b_c), _ptr(s) if read_signs else None, _ptr(y),
                _ptr(so) if writeSigns else None, code, _i4(x.shape), _i4(x.stride()), _i4(y.shape), _i4(y.stride()),
                fu_w, fu_h, fd_w, fd_h, up, down, px0, py0, s_h, s_wb, sx, sy, float(gain), float(slope),
                float(clamp), int(bool(flip_filters)), int(bool(writeSigns)), _stream(x)), 'filtered_lrelu')
        if rc == LVG_UNSUPPORTED:
            return None, None, -1
        return y, so, 0

    def filtered_lrelu_act_(self, x, si, sx, sy, gain, slope, clamp, writeSigns):
        if not x.is_cuda:
            raise RuntimeError('x must reside on CUDA device')
        if x.ndim != 4:
            raise RuntimeError('x must be rank 4')
        if x.numel() == 0:
            raise RuntimeError('x is empty')
        code = _dtype_code(x, 'filtered_lrelu_act_')
        so = None
        s = si if not _absent(si) else None
        read_signs = s is not None
        if writeSigns:
            sw = (x.shape[3] + 15) & ~15
            s = so = torch.empty([x.shape[0], x.shape[1], x.shape[2], sw >> 2], dtype=torch.uint8, device=x.device)
        if s is not None:
            if not (s.is_contiguous() and s.dtype == torch.uint8 and s.device == x.device and s.ndim == 4
                    and s.shape[0] == x.shape[0] and s.shape[1] == x.shape[1]):
                raise RuntimeError('signs must be a contiguous uint8 [N, C, H, W/4] tensor on the same device as x')
        s_h, s_wb = (s.shape[2], s.shape[3]) if s is not None else (0, 0)
        with _DeviceGuard(x):
            _check(self._lib.lvg_filtered_lrelu_act(_ptr(x), _ptr(s) if (read_signs and not writeSigns) else None,
                                                    _ptr(so) if writeSigns else None, code, _i4(x.shape), _i4(x.stride()),
                                                    s_h, s_wb, sx, sy, float(gain), float(slope), float(clamp),
                                                    int(bool(writeSigns)), _stream(x)), 'filtered_lrelu_act_')
        return so


class FmaPlugin:
    """Elementwise a * b + c with broadcasting (no reference plugin: fma.py uses torch.addcmul)."""

    def __init__(self, lib):
        self._lib = lib

    def fma(self, a, b, c):
        if not (a.is_cuda and b.device == a.device and c.device == a.device):
            raise RuntimeError('fma operands must reside on one CUDA device')
        dtype = torch.promote_types(torch.promote_types(a.dtype, b.dtype), c.dtype)
        a, b, c = a.to(dtype), b.to(dtype), c.to(dtype)
        code = _dtype_code(a, 'fma')
        shape = torch.broadcast_shapes(a.shape, b.shape, c.shape)
        if len(shape) > 6:
            raise RuntimeError('fma supports at most 6 dimensions')
        out = torch.empty(shape, dtype=dtype, device=a.device)
        if out.numel() == 0:
            return out
        ae, be, ce = a.expand(shape), b.expand(shape), c.expand(shape)
        pad = [0] * (6 - len(shape))
        with _DeviceGuard(out):
            _check(self._lib.lvg_fma(_ptr(ae), _ptr(be), _ptr(ce), _ptr(out), code, len(shape),
                                     _I64x6(*(list(shape) + [1] * len(pad))), _I64x6(*(list(ae.stride()) + pad)),
                                     _I64x6(*(list(be.stride()) + pad)), _I64x6(*(list(ce.stride()) + pad)),
                                     _stream(out)), 'fma')
        return out


class Conv2dPlugin:
    """Tensor-core convolution behind conv2d_gradfix.conv2d (no reference plugin: the reference calls cuDNN)."""

    def __init__(self, lib):
        self._lib = lib
        self._ws = {}

    def supported(self, x, w, stride, padding, dilation, groups):
        if not (x.is_cuda and x.dtype == torch.float16 and w.dtype == torch.float16 and x.ndim == 4 and w.ndim == 4):
            return False
        if tuple(stride) != (1, 1) or tuple(dilation) != (1, 1):
            return False
        kh, kw = w.shape[2], w.shape[3]
        # the C side's envelope, all three legs: 3x3 / 1x1, 0 <= pad <= k - 1 (lvg_conv2d_dgrad), at least one sample
        if (kh, kw) not in ((3, 3), (1, 1)) or min(padding) < 0 or padding[0] > kh - 1 or padding[1] > kw - 1 or x.shape[0] < 1 or x.numel() == 0:
            return False
        if x.shape[1] != w.shape[1] * groups or w.shape[0] % groups != 0:
            return False
        return x.shape[0] * groups <= 65535 and x.shape[2] + 2 * padding[0] >= kh and x.shape[3] + 2 * padding[1] >= kw

    def _workspace(self, x, n, groups, cin, cout, h, wd, kh, kw, ph, pw):
        need = self._lib.lvg_conv2d_fprop_workspace(1, n, groups, cin, cout, h, wd, kh, kw, 1, ph, pw)
        if need < 0:
            raise RuntimeError('conv2d: configuration outside the tensor-core kernel envelope')
        key = x.device
        buf = self._ws.get(key)
        if buf is None or buf.numel() < need:
            buf = torch.empty(max(int(need), 1 << 20), dtype=torch.uint8, device=x.device)
            self._ws[key] = buf     # stream-ordered reuse: every call repacks before it reads
        return buf, need

    def fprop(self, x, w, padding, groups):
        x, w = x.contiguous(), w.contiguous()
        n, ctot, h, wd = x.shape
        cout_tot, cin, kh, kw = w.shape
        ph, pw = padding
        cout = cout_tot // groups
        y = torch.empty([n, cout_tot, h + 2 * ph - kh + 1, wd + 2 * pw - kw + 1], dtype=x.dtype, device=x.device)
        ws, need = self._workspace(x, n, groups, cin, cout, h, wd, kh, kw, ph, pw)
        with _DeviceGuard(x):
            rc = _check(self._lib.lvg_conv2d_fprop(_ptr(x), _ptr(w), _ptr(y), 1, n, groups, cin, cout, h, wd, kh, kw, 1, ph, pw,
                                                   _ptr(ws), ws.numel(), _stream(x)), 'conv2d_fprop')
        if rc == LVG_UNSUPPORTED:
            raise RuntimeError('conv2d_fprop: ' + self._lib.lvg_last_error().decode())
        return y

    def dgrad(self, dy, w, x_shape, padding, groups):
        dy, w = dy.contiguous(), w.contiguous()
        n, ctot, h, wd = x_shape
        cout_tot, cin, kh, kw = w.shape
        ph, pw = padding
        cout = cout_tot // groups
        dx = torch.empty(list(x_shape), dtype=dy.dtype, device=dy.device)
        ws, need = self._workspace(dy, n, groups, cin, cout, h, wd, kh, kw, ph, pw)
        with _DeviceGuard(dy):
            rc = _check(self._lib.lvg_conv2d_dgrad(_ptr(dy), _ptr(w), _ptr(dx), 1, n, groups, cin, cout, h, wd, kh, kw, 1, ph, pw,
                                                   _ptr(ws), ws.numel(), _stream(dy)), 'conv2d_dgrad')
        if rc == LVG_UNSUPPORTED:
            raise RuntimeError('conv2d_dgrad: ' + self._lib.lvg_last_error().decode())
        return dx

    def wgrad(self, x, dy, w_shape, padding, groups):
        """dw [G*Cout, Cin, kh, kw] (fp16, summed over the batch) from x and dy."""
        x, dy = x.contiguous(), dy.contiguous()
        n, ctot, h, wd = x.shape
        cout_tot, cin, kh, kw = w_shape
        ph, pw = padding
        cout = cout_tot // groups
        dw = torch.empty(list(w_shape), dtype=x.dtype, device=x.device)
        with _DeviceGuard(x):
            rc = _check(self._lib.lvg_conv2d_wgrad(_ptr(x), _ptr(dy), _ptr(dw), 1, n, groups, cin, cout, h, wd, kh, kw, 1, ph, pw,
                                                   _stream(x)), 'conv2d_wgrad')
        if rc == LVG_UNSUPPORTED:
            raise RuntimeError('conv2d_wgrad: ' + self._lib.lvg_last_error().decode())
        return dw


class ConvNdPlugin:
    """TMA-fed tcgen05 implicit-GEMM convolution for 1-D / 2-D / 3-D NC(T)HW tensors, fp16 or fp32 (bf16 hi/lo split),
    stride 1: forward (optionally with the bias_act epilogue fused), input gradient, weight gradient. No reference
    plugin: the reference hands these to cuDNN (conv2d_gradfix.py:37-45, generator_lres.py:119, discriminator_lres.py:121,172)."""

    def __init__(self, lib):
        self._lib = lib
        self._ws = {}

    @staticmethod
    def _dims(x, w):
        """-> (n, c, t, h, w), (kt, kh, kw), spatial rank"""
        nd = x.ndim - 2
        sp = list(x.shape[2:])
        k = list(w.shape[2:])
        while len(sp) < 3:
            sp.insert(0, 1)
            k.insert(0, 1)
        return sp, k, nd

    @staticmethod
    def _pad3(padding, nd):
        p = list(padding) if isinstance(padding, (list, tuple)) else [padding] * nd
        return [0] * (3 - nd) + [int(v) for v in p]

    def supported(self, x, w, stride, padding, dilation, groups):
        if not (x.is_cuda and x.dtype in (torch.float16, torch.float32) and w.dtype == x.dtype and x.ndim == w.ndim and x.ndim in (3, 4, 5)):
            return False
        nd = x.ndim - 2
        as_t = lambda v: tuple(v) if isinstance(v, (list, tuple)) else (v,) * nd       # noqa: E731
        st = as_t(stride)
        if as_t(dilation) != (1,) * nd or len(set(st[-2:])) != 1 or not (1 <= st[-1] <= 4) or (nd == 3 and st[0] != 1) or (nd == 1 and st[0] != 1):
            return False
        sp, k, _ = self._dims(x, w)
        pad = self._pad3(padding, nd)
        if k[1] * k[2] > 9 or k[0] > 7 or k[2] > 3 or min(pad) < 0 or any(p > kk - 1 for p, kk in zip(pad, k)):
            return False
        if x.shape[1] != w.shape[1] * groups or w.shape[0] % groups != 0 or x.shape[0] * groups > 65535 or x.numel() == 0:
            return False
        return all(s + 2 * p - kk + 1 >= 1 for s, p, kk in zip(sp, pad, k))

    def _workspace(self, device, need):
        if need < 0:
            raise RuntimeError('convnd: configuration outside the tensor-core kernel envelope')
        buf = self._ws.get(device)
        if buf is None or buf.numel() < need:
            buf = torch.empty(max(int(need), 1 << 22), dtype=torch.uint8, device=device)
            self._ws[device] = buf      # stream-ordered reuse: every call re-tiles its operands before it reads them
        return buf

    def _args(self, x_shape, w_shape, padding, groups, dtype):
        nd = len(x_shape) - 2
        sp = [1] * (3 - nd) + list(x_shape[2:])
        k = [1] * (3 - nd) + list(w_shape[2:])
        pad = self._pad3(padding, nd)
        code = 1 if dtype == torch.float16 else 0
        return [code, x_shape[0], groups, w_shape[1], w_shape[0] // groups] + sp + k + pad, sp, k, pad

    def fprop(self, x, w, padding, groups, bias=None, act=0, alpha=0.2, gain=1.0, clamp=-1.0, stride=1):
        x, w = x.contiguous(), w.contiguous()
        a, sp, k, pad = self._args(tuple(x.shape), tuple(w.shape), padding, groups, x.dtype)
        st3 = [1, stride, stride] if x.ndim >= 4 else [1, 1, 1]
        out_sp = [(s + 2 * p - kk) // q + 1 for s, p, kk, q in zip(sp, pad, k, st3)][3 - (x.ndim - 2):]
        y = torch.empty([x.shape[0], w.shape[0]] + out_sp, dtype=x.dtype, device=x.device)
        ws = self._workspace(x.device, self._lib.lvg_convnd_workspace(*a))
        if bias is not None:
            bias = bias.to(torch.float32).contiguous()
        with _DeviceGuard(x):
            rc = _check(self._lib.lvg_convnd_fprop(_ptr(x), _ptr(w), _ptr(y), *a, int(stride), _ptr(bias), int(act), float(alpha), float(gain),
                                                   float(clamp), _ptr(ws), ws.numel(), _stream(x)), 'convnd_fprop')
        if rc == LVG_UNSUPPORTED:
            raise RuntimeError('convnd_fprop: ' + self._lib.lvg_last_error().decode())
        return y

    def dgrad(self, dy, w, x_shape, padding, groups, stride=1):
        dy, w = dy.contiguous(), w.contiguous()
        a, sp, k, pad = self._args(tuple(x_shape), tuple(w.shape), padding, groups, dy.dtype)
        dx = torch.empty(list(x_shape), dtype=dy.dtype, device=dy.device)
        ws = self._workspace(dy.device, self._lib.lvg_convnd_workspace(*a))
        with _DeviceGuard(dy):
            rc = _check(self._lib.lvg_convnd_dgrad(_ptr(dy), _ptr(w), _ptr(dx), *a, int(stride), _ptr(ws), ws.numel(), _stream(dy)), 'convnd_dgrad')
        if rc == LVG_UNSUPPORTED:
            raise RuntimeError('convnd_dgrad: ' + self._lib.lvg_last_error().decode())
        return dx

    def fir1d_depthwise(self, x, w):
        """y[n][g][t] = sum_k w[g][0][k] * x[n][g][t + k]: F.conv1d(x, w, groups = G) with one channel per group, fp32."""
        x, w = x.contiguous(), w.contiguous()
        n, g, lin = x.shape
        k = w.shape[2]
        y = torch.empty([n, g, lin - k + 1], dtype=x.dtype, device=x.device)
        ws = torch.empty([g + 4], dtype=torch.int32, device=x.device)
        with _DeviceGuard(x):
            _check(self._lib.lvg_fir1d_depthwise(_ptr(x), _ptr(w), _ptr(y), n, g, lin, k, _ptr(ws), ws.numel() * 4, _stream(x)), 'fir1d_depthwise')
        return y

    def wgrad(self, x, dy, w_shape, padding, groups, stride=1):
        x, dy = x.contiguous(), dy.contiguous()
        a, sp, k, pad = self._args(tuple(x.shape), tuple(w_shape), padding, groups, x.dtype)
        dw = torch.empty(list(w_shape), dtype=x.dtype, device=x.device)
        ws = self._workspace(x.device, self._lib.lvg_convnd_wgrad_workspace(*a))
        with _DeviceGuard(x):
            rc = _check(self._lib.lvg_convnd_wgrad(_ptr(x), _ptr(dy), _ptr(dw), *a, int(stride), _ptr(ws), ws.numel(), _stream(x)), 'convnd_wgrad')
        if rc == LVG_UNSUPPORTED:
            raise RuntimeError('convnd_wgrad: ' + self._lib.lvg_last_error().decode())
        return dw

    def backward(self, x, dy, w, padding, groups, stride=1):
        """(dx, dw) of y = conv(x, w) in one call: dy is re-tiled once for the input- and the weight-gradient kernels."""
        x, dy, w = x.contiguous(), dy.contiguous(), w.contiguous()
        a, sp, k, pad = self._args(tuple(x.shape), tuple(w.shape), padding, groups, x.dtype)
        dx = torch.empty_like(x)
        dw = torch.empty_like(w)
        ws = self._workspace(x.device, self._lib.lvg_convnd_backward_workspace(*a))
        with _DeviceGuard(x):
            rc = _check(self._lib.lvg_convnd_backward(_ptr(x), _ptr(dy), _ptr(w), _ptr(dx), _ptr(dw), *a, int(stride), _ptr(ws), ws.numel(),
                                                      _stream(x)), 'convnd_backward')
        if rc == LVG_UNSUPPORTED:
            raise RuntimeError('convnd_backward: ' + self._lib.lvg_last_error().decode())
        return dx, dw


_PLUGIN_CLASSES = {
    'convnd_plugin': ConvNdPlugin,
    'conv2d_plugin': Conv2dPlugin,
    'bias_act_plugin': BiasActPlugin,
    'upfirdn2d_plugin': Upfirdn2dPlugin,
    'filtered_lrelu_plugin': FilteredLReluPlugin,
    'fma_plugin': FmaPlugin,
}


def get_plugin(module_name, sources=None, headers=None, source_dir=None, **build_kwargs):
    """Same call shape as the reference loader (custom_ops.py:59); sources/headers/build flags
    are accepted and ignored because the library is prebuilt for sm_100a."""
    if module_name not in _plugins:
        if module_name not in _PLUGIN_CLASSES:
            raise RuntimeError(f'unknown plugin "{module_name}"')
        _plugins[module_name] = _PLUGIN_CLASSES[module_name](load_library())
    return _plugins[module_name]
