"""Small helpers shared by the op modules (kept local so that the ops do not
depend on the reference's ``torch_utils.misc`` / ``dnnlib``)."""
import torch


class AttrDict(dict):
    """dict with attribute access -- stands in for ``dnnlib.EasyDict`` (dnnlib/util.py:40-55)."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name) from None

    def __setattr__(self, name, value):
        self[name] = value

    def __delattr__(self, name):
        del self[name]


def null_tensor():
    return torch.empty([0])


def is_absent(t):
    return t is None or t.numel() == 0


def dense_like(t, like):
    """t laid out exactly like `like` (same strides); copies only when needed."""
    if t.stride() == like.stride() and t.shape == like.shape:
        return t
    return torch.empty_like(like).copy_(t)


def is_dense(x):
    """True when x occupies numel() consecutive elements in SOME dimension order (no gaps, no overlap)."""
    if x.is_contiguous():
        return True
    dims = sorted((d for d in range(x.ndim) if x.shape[d] != 1), key=lambda d: x.stride(d))
    expect = 1
    for d in dims:
        if x.stride(d) != expect:
            return False
        expect *= x.shape[d]
    return True


def as_dense(x):
    """x itself when it is dense in any dimension order, else a contiguous copy."""
    return x if is_dense(x) else x.contiguous()
