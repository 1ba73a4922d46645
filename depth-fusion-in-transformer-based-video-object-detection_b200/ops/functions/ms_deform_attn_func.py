"""Autograd binding of the sm_100a multi-scale deformable attention op.

Mirror of /root/reference/models/ops/functions/ms_deform_attn_func.py:21-38: same class name,
same ``apply(value, value_spatial_shapes, value_level_start_index, sampling_locations,
attention_weights, im2col_step)`` signature, differentiable w.r.t. arguments 0, 3 and 4,
``None`` for the rest, once-differentiable.  The reference's debug path
``ms_deform_attn_core_pytorch`` (func.py:41-61) is NOT shipped here: it is the test oracle and
lives in oracle/msda_oracle.py; the product has no CPU or PyTorch fallback.
"""
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from ... import MultiScaleDeformableAttention as MSDA


class MSDeformAttnFunction(Function):
    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations,
                attention_weights, im2col_step):
        ctx.im2col_step = im2col_step
        output = MSDA.ms_deform_attn_forward(value, value_spatial_shapes, value_level_start_index,
                                             sampling_locations, attention_weights, im2col_step)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                              attention_weights)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, shapes, level_start, sampling_locations, attention_weights = ctx.saved_tensors
        grad_value, grad_loc, grad_attn = MSDA.ms_deform_attn_backward(
            value, shapes, level_start, sampling_locations, attention_weights, grad_output.contiguous(),
            ctx.im2col_step)
        return grad_value, None, None, grad_loc, grad_attn, None
