"""``MSDeformAttn`` -- multi-scale deformable attention module, B200 build.

Drop-in for /root/reference/models/ops/modules/ms_deform_attn.py:31-117: same constructor
``MSDeformAttn(d_model=256, n_levels=4, n_heads=8, n_points=4)``, same forward signature, same
parameter names (``sampling_offsets``, ``attention_weights``, ``value_proj``, ``output_proj`` --
reference checkpoints load unchanged), same initialisation (``_reset_parameters`` is called by the
reference transformers, deformable_transformer_single.py:96-98) and same ``im2col_step`` attribute.

What differs is underneath: the gather runs in the hand-written sm_100a kernels of
libmsda_b200.so; the four projections stay library GEMMs (tensor cores) around it.
The reference's per-call device->host sync (``assert ... .sum() == Len_in``, :92) is paid once
per distinct ``input_spatial_shapes`` tensor instead of once per layer call.

Fused path (default, ``self.fused``): the sampling-offset and attention-weight projections run as
ONE GEMM and a single kernel per direction does softmax, offset -> location arithmetic, the
multi-level bilinear gather and the weighted head reduction (ops/functions/
ms_deform_attn_fused_func.py); locations and attention weights never reach HBM.  Shapes the fused
kernels do not cover (head width not 16/32/64, more than 16 samples per head, fp64) take the
reference's unfused sequence around the drop-in op -- still on the GPU, never on the CPU.
"""
import math
import warnings
import weakref

import torch
import torch.nn.functional as F
from torch import nn
from torch.nn.init import constant_, xavier_uniform_

from ..functions import (MSDeformAttnFunction, MSDeformAttnFusedFunction, fused_supported, linear, linear_wb, linear_zero_rows,
                         zero_masked_rows_, head_major_supported, value_proj_head_major, fused_forward_head_major)

# bf16 inference: own value-projection GEMM with a head-major epilogue + the head-major fused gather (see forward).
# Opt-in: measured on B200 (tools/run_hm_layer.py, 6-layer encoder, batch 8) the route is slower than the library GEMM +
# reference-layout gather (5.25 vs 4.62 ms) -- see DESIGN.md section 3.11.
HEAD_MAJOR_INFERENCE = False


def _is_power_of_2(n):
    if (not isinstance(n, int)) or (n < 0):
        raise ValueError("invalid input for _is_power_of_2: {} (type: {})".format(n, type(n)))
    return (n & (n - 1) == 0) and n != 0


# host-side facts about spatial_shapes tensors already read back once.  Keyed by the identity of the
# tensor OBJECT (validated through a weak reference, plus its in-place version counter): a new
# tensor that merely re-uses a freed device address can never alias an old entry.
_SHAPE_CACHE = {}
_SHAPE_CACHE_MAX = 64


def host_shape_list(spatial_shapes):
    """[(H, W), ...] as python ints, device->host at most once per tensor object (the reference
    syncs on every call: ms_deform_attn.py:92, deformable_transformer_single.py:166-169)."""
    if not torch.is_tensor(spatial_shapes):
        return [(int(h), int(w)) for h, w in spatial_shapes]
    key = id(spatial_shapes)
    hit = _SHAPE_CACHE.get(key)
    if hit is not None and hit[0]() is spatial_shapes and hit[1] == spatial_shapes._version:
        return hit[2]
    if len(_SHAPE_CACHE) >= _SHAPE_CACHE_MAX:
        _SHAPE_CACHE.clear()
    shapes = [(int(h), int(w)) for h, w in spatial_shapes.tolist()]
    _SHAPE_CACHE[key] = (weakref.ref(spatial_shapes), spatial_shapes._version, shapes)
    return shapes


def _total_pixels(spatial_shapes):
    return sum(h * w for h, w in host_shape_list(spatial_shapes))


class MSDeformAttn(nn.Module):
    def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4):
        """
        :param d_model   hidden dimension
        :param n_levels  number of feature levels
        :param n_heads   number of attention heads
        :param n_points  number of sampling points per attention head per feature level
        """
        super().__init__()
        if d_model % n_heads != 0:
            raise ValueError('d_model must be divisible by n_heads, but got {} and {}'.format(d_model, n_heads))
        head_dim = d_model // n_heads
        # the vectorised sm_100a kernels need head_dim*itemsize to be a power-of-two multiple of
        # 16 bytes; anything else runs through the (slower) shape-generic kernels
        if not _is_power_of_2(head_dim):
            warnings.warn("You'd better set d_model in MSDeformAttn to make the dimension of each attention head "
                          "a power of 2 which is more efficient in our CUDA implementation.")

        self.im2col_step = 64
        self.fused = True          # one-kernel layer path when the shape allows it

        self.d_model = d_model
        self.n_levels = n_levels
        self.n_heads = n_heads
        self.n_points = n_points

        self.sampling_offsets = nn.Linear(d_model, n_heads * n_levels * n_points * 2)
        self.attention_weights = nn.Linear(d_model, n_heads * n_levels * n_points)
        self.value_proj = nn.Linear(d_model, d_model)
        self.output_proj = nn.Linear(d_model, d_model)

        self._reset_parameters()

    def _reset_parameters(self):
        """Reference :60-76 -- offsets start as a fixed compass pattern (head h looks along
        direction 2*pi*h/n_heads, point i at i+1 pixels), attention starts uniform."""
        constant_(self.sampling_offsets.weight.data, 0.)
        angles = torch.arange(self.n_heads, dtype=torch.float32) * (2.0 * math.pi / self.n_heads)
        compass = torch.stack([angles.cos(), angles.sin()], -1)
        compass = compass / compass.abs().max(-1, keepdim=True)[0]
        steps = torch.arange(1, self.n_points + 1, dtype=torch.float32).view(1, 1, self.n_points, 1)
        bias = compass.view(self.n_heads, 1, 1, 2).repeat(1, self.n_levels, self.n_points, 1) * steps
        with torch.no_grad():
            self.sampling_offsets.bias = nn.Parameter(bias.reshape(-1))
        constant_(self.attention_weights.weight.data, 0.)
        constant_(self.attention_weights.bias.data, 0.)
        xavier_uniform_(self.value_proj.weight.data)
        constant_(self.value_proj.bias.data, 0.)
        xavier_uniform_(self.output_proj.weight.data)
        constant_(self.output_proj.bias.data, 0.)

    def _sampling_locations(self, reference_points, offsets, input_spatial_shapes):
        """Reference :102-113.  offsets [N,Lq,M,L,P,2] in pixels of each level (2-d refs) or in
        units of half a box (4-d refs)."""
        if reference_points.shape[-1] == 2:
            normalizer = torch.stack([input_spatial_shapes[..., 1], input_spatial_shapes[..., 0]], -1)
            return reference_points[:, :, None, :, None, :] + offsets / normalizer[None, None, None, :, None, :]
        if reference_points.shape[-1] == 4:
            return reference_points[:, :, None, :, None, :2] \
                + offsets / self.n_points * reference_points[:, :, None, :, None, 2:] * 0.5
        raise ValueError(
            'Last dim of reference_points must be 2 or 4, but get {} instead.'.format(reference_points.shape[-1]))

    def forward(self, query, reference_points, input_flatten, input_spatial_shapes, input_level_start_index,
                input_padding_mask=None, project_output=True, precomputed_value=None):
        """
        :param query                    (N, Length_{query}, C)
        :param reference_points         (N, Length_{query}, n_levels, 2), range in [0, 1], top-left (0,0),
                                        bottom-right (1, 1), including padding area
                                        or (N, Length_{query}, n_levels, 4), add additional (w, h) to form boxes
        :param input_flatten            (N, sum_l H_l*W_l, C)
        :param input_spatial_shapes     (n_levels, 2), [(H_0, W_0), ..., (H_{L-1}, W_{L-1})]
        :param input_level_start_index  (n_levels, ), [0, H_0*W_0, H_0*W_0+H_1*W_1, ...]
        :param input_padding_mask       (N, sum_l H_l*W_l), True for padding elements

        :return output                  (N, Length_{query}, C)

        ``precomputed_value`` (private): this module's ``value_proj(input_flatten)`` with the padding rows already
        zeroed, as a ``[N, Len_in, n_heads, head_dim]`` tensor or pixel-strided view -- see :func:`project_values`.
        ``project_output=False`` (private to this repo's layer classes) returns the heads' output BEFORE
        ``output_proj``: the layer then runs projection + residual + LayerNorm as one kernel
        (ops.functions.proj_layer_norm).
        """
        N, Len_q, _ = query.shape
        N, Len_in, _ = input_flatten.shape
        assert _total_pixels(input_spatial_shapes) == Len_in

        if reference_points.shape[-1] not in (2, 4):
            raise ValueError(
                'Last dim of reference_points must be 2 or 4, but get {} instead.'.format(reference_points.shape[-1]))

        # bf16 inference at d_model 256 / 8 heads: value_proj as an own tcgen05 GEMM whose epilogue adds the bias, zeroes
        # the padding rows and writes the HEAD-MAJOR layout the gather likes best (3 instead of 4 L1 lines per sample);
        # the fused forward kernel then reads that layout.  Opt-in (HEAD_MAJOR_INFERENCE).
        if (self.fused and HEAD_MAJOR_INFERENCE and precomputed_value is None and input_flatten.is_cuda
                and input_flatten.dtype == torch.bfloat16
                and not (torch.is_grad_enabled() and (query.requires_grad or input_flatten.requires_grad
                                                      or self.value_proj.weight.requires_grad))):
            weight, bias = self._raw_projection_params()
            raw = linear_wb(query, weight, bias)
            if head_major_supported(self.value_proj, input_flatten, raw, reference_points.shape[-1], self.n_heads,
                                    self.n_levels, self.n_points):
                value_hm = value_proj_head_major(self.value_proj, input_flatten, input_padding_mask, self.n_heads)
                output = fused_forward_head_major(value_hm, input_spatial_shapes, input_level_start_index,
                                                  reference_points, raw, self.n_points)
                return linear(self.output_proj, output) if project_output else output

        # 2-d in, 2-d out: the projection result is a fresh tensor (not a view), so the padding rows
        # can be zeroed in place without autograd having to copy slices back
        if precomputed_value is not None:
            value = precomputed_value
        else:
            # `value` is consumed by the deformable-attention op below and by nothing else
            value = linear_zero_rows(self.value_proj, input_flatten.reshape(N * Len_in, -1),
                                     None if input_padding_mask is None else input_padding_mask.reshape(-1))
            value = value.view(N, Len_in, self.n_heads, self.d_model // self.n_heads)
        if reference_points.shape[-1] not in (2, 4):
            raise ValueError(
                'Last dim of reference_points must be 2 or 4, but get {} instead.'.format(reference_points.shape[-1]))

        if self.fused and value.is_cuda:
            # [ offsets | logits ] from one GEMM; parameters stay separate for checkpoint compatibility
            weight, bias = self._raw_projection_params()
            raw = linear_wb(query, weight, bias)
            if raw.dtype != value.dtype and value.dtype == torch.float32:
                raw = raw.float()
            if fused_supported(value, raw, reference_points.shape[-1], self.n_levels, self.n_points):
                output = MSDeformAttnFusedFunction.apply(
                    value, input_spatial_shapes, input_level_start_index, reference_points, raw, self.n_points)
                return linear(self.output_proj, output) if project_output else output
            split = self.n_heads * self.n_levels * self.n_points * 2
            offsets = raw[..., :split].reshape(N, Len_q, self.n_heads, self.n_levels, self.n_points, 2)
            attention = raw[..., split:].reshape(N, Len_q, self.n_heads, self.n_levels * self.n_points)
            attention = F.softmax(attention, -1).view(N, Len_q, self.n_heads, self.n_levels, self.n_points)
            sampling_locations = self._sampling_locations(reference_points, offsets, input_spatial_shapes)
            output = MSDeformAttnFunction.apply(
                value.contiguous(), input_spatial_shapes, input_level_start_index, sampling_locations.contiguous(),
                attention.contiguous(), self.im2col_step)
            return linear(self.output_proj, output) if project_output else output

        offsets = self.sampling_offsets(query).view(N, Len_q, self.n_heads, self.n_levels, self.n_points, 2)
        attention = self.attention_weights(query).view(N, Len_q, self.n_heads, self.n_levels * self.n_points)
        attention = F.softmax(attention, -1).view(N, Len_q, self.n_heads, self.n_levels, self.n_points)
        sampling_locations = self._sampling_locations(reference_points, offsets, input_spatial_shapes)
        output = MSDeformAttnFunction.apply(
            value.contiguous(), input_spatial_shapes, input_level_start_index, sampling_locations.contiguous(),
            attention.contiguous(), self.im2col_step)
        return linear(self.output_proj, output) if project_output else output


def _raw_projection_params(self):
    """[sampling_offsets | attention_weights] weight and bias as one matrix for the single projection GEMM of the fused
    path.  Training: concatenated inside autograd every call (the gradients must reach the two parameters).
    Inference: concatenated once and reused until a parameter changes (in-place update, .to(), load_state_dict) --
    two launches and a 300 KB copy per layer call otherwise."""
    params = (self.sampling_offsets.weight, self.attention_weights.weight, self.sampling_offsets.bias,
              self.attention_weights.bias)
    if torch.is_grad_enabled() and any(p.requires_grad for p in params):
        return torch.cat(params[:2], 0), torch.cat(params[2:], 0)
    key = tuple((p.data_ptr(), p._version, p.dtype, p.device) for p in params)
    cache = getattr(self, "_raw_cache", None)
    if cache is None or cache[0] != key:
        with torch.no_grad():
            cache = (key, torch.cat(params[:2], 0), torch.cat(params[2:], 0))
        # tensors made during a CUDA-graph capture live in the graph's private pool: never cache those
        if not (params[0].is_cuda and torch.cuda.is_current_stream_capturing()):
            self._raw_cache = cache
    return cache[1], cache[2]


MSDeformAttn._raw_projection_params = _raw_projection_params


def project_values(modules, input_flatten, input_padding_mask=None):
    """``value_proj`` of SEVERAL ``MSDeformAttn`` modules that attend to the same ``input_flatten`` (the decoder
    layers all re-project the encoder memory, reference deformable_transformer_single.py:617-628 ->
    ms_deform_attn.py:94-96) as ONE GEMM: the memory is read once instead of once per layer, padding rows are zeroed
    once.  Returns one pixel-strided ``[N, Len_in, n_heads, head_dim]`` view per module for ``precomputed_value``
    (the fused forward kernel reads the slice in place).  Inference only; returns ``None`` when gradients are
    needed or the modules differ in shape."""
    first = modules[0]
    if len(modules) < 2 or not input_flatten.is_cuda or any(
            m.d_model != first.d_model or m.n_heads != first.n_heads or
            m.value_proj.weight.dtype != input_flatten.dtype for m in modules):
        return None
    if torch.is_grad_enabled() and (input_flatten.requires_grad or
                                    any(m.value_proj.weight.requires_grad for m in modules)):
        return None
    N, Len_in, _ = input_flatten.shape
    weight = torch.cat([m.value_proj.weight for m in modules], 0)
    bias = torch.cat([m.value_proj.bias for m in modules], 0)
    values = linear_wb(input_flatten.reshape(N * Len_in, -1), weight, bias)          # [N*Len_in, layers*C]
    if input_padding_mask is not None:
        values = zero_masked_rows_(values, input_padding_mask.reshape(-1))
    c, heads = first.d_model, first.n_heads
    values = values.view(N, Len_in, len(modules), heads, c // heads)
    return [values[:, :, i] for i in range(len(modules))]
