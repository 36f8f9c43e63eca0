"""``conv2d`` / ``conv_transpose2d`` entry points (``torch_utils.ops.conv2d_gradfix``).

Same functions and module switches as the reference (conv2d_gradfix.py:22-45):
``enabled``, ``weight_gradients_disabled``, ``no_weight_gradients()``. In the
reference the custom autograd path is inert on torch >= 1.11 (:49-58) and every
call lands in cuDNN. Here this module is the tensor-core boundary: CUDA calls
inside the envelope of the tcgen05 implicit-GEMM kernel (csrc/conv2d_tc.cu;
grouped "modulated" 3x3 / 1x1 convolutions in fp16) are routed to it through
``_native`` below; everything else goes to ``torch.nn.functional``.
"""
import contextlib

import torch

enabled = False                     # kept for API compatibility (train_lres.py:80 sets it)
weight_gradients_disabled = False   # forcefully skip weight gradients (R1 penalty, see no_weight_gradients)

# native convolution backend: an object with .conv2d(input, weight, bias, stride, padding, dilation, groups)
# returning a tensor or None ("outside the kernel's envelope"). Installed by conv2d_native.install().
_native = None


@contextlib.contextmanager
def no_weight_gradients(disable=True):
    global weight_gradients_disabled
    old = weight_gradients_disabled
    if disable:
        weight_gradients_disabled = True
    yield
    weight_gradients_disabled = old


def conv2d(input, weight, bias=None, stride=1, padding=0, dilation=1, groups=1):
    if _native is not None and input.device.type == 'cuda':
        out = _native.conv2d(input, weight, bias, stride, padding, dilation, groups)
        if out is not None:
            return out
    return torch.nn.functional.conv2d(input=input, weight=weight, bias=bias, stride=stride, padding=padding,
                                      dilation=dilation, groups=groups)


def conv_transpose2d(input, weight, bias=None, stride=1, padding=0, output_padding=0, groups=1, dilation=1):
    if _native is not None and input.device.type == 'cuda':
        out = _native.conv_transpose2d(input, weight, bias, stride, padding, output_padding, groups, dilation)
        if out is not None:
            return out
    return torch.nn.functional.conv_transpose2d(input=input, weight=weight, bias=bias, stride=stride, padding=padding,
                                                output_padding=output_padding, groups=groups, dilation=dilation)
