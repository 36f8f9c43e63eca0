"""``grid_sample`` with a double-backward that only propagates to the input image
(``torch_utils.ops.grid_sample_gradfix``, reference grid_sample_gradfix.py:22-84).

Out of the hot-path scope: only the ADA augmentation pipeline calls it
(model/ada_augment.py:300). Kept importable with the reference's semantics; the
arithmetic is ATen's grid sampler (no kernel of ours).
"""
import torch

enabled = False  # route through the custom autograd functions below when True


def grid_sample(input, grid):
    if enabled:
        return _Sample.apply(input, grid)
    return torch.nn.functional.grid_sample(input=input, grid=grid, mode='bilinear', padding_mode='zeros',
                                           align_corners=False)


class _Sample(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, grid):
        assert input.ndim == 4 and grid.ndim == 4
        ctx.save_for_backward(input, grid)
        return torch.nn.functional.grid_sample(input=input, grid=grid, mode='bilinear', padding_mode='zeros',
                                               align_corners=False)

    @staticmethod
    def backward(ctx, grad_output):
        input, grid = ctx.saved_tensors
        return _SampleGrad.apply(grad_output, input, grid)


class _SampleGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, grad_output, input, grid):
        mask = (ctx.needs_input_grad[1], ctx.needs_input_grad[2])
        # bilinear = 0, zeros padding = 0, align_corners = False
        grad_input, grad_grid = torch.ops.aten.grid_sampler_2d_backward(grad_output, input, grid, 0, 0, False, mask)
        ctx.save_for_backward(grid)
        return grad_input, grad_grid

    @staticmethod
    def backward(ctx, gg_input, gg_grid):
        grid, = ctx.saved_tensors
        gg_output = None
        if ctx.needs_input_grad[0]:
            gg_output = _Sample.apply(gg_input, grid)   # sampling is linear in the image
        assert not ctx.needs_input_grad[2]
        return gg_output, None, None
