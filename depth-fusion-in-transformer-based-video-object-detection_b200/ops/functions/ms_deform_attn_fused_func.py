"""Autograd binding of the FUSED layer op (C ABI msda_fused_forward / msda_fused_backward).

One kernel per direction replaces, inside ``MSDeformAttn.forward``, the reference's chain
    softmax(attention_weights(query))                       ms_deform_attn.py:99-100
    reference_points + sampling_offsets(query) / normaliser :102-110
    MSDeformAttnFunction.apply(...)                         :114
so the sampling locations [N,Lq,M,L,P,2] and attention weights [N,Lq,M,L,P] (and their
gradients) are never written to HBM.  ``raw`` is the output of ONE projection GEMM whose rows are
``[ sampling offsets (M*L*P*2) | attention logits (M*L*P) ]``.
"""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from ... import _lib
from ... import MultiScaleDeformableAttention as _msda

_DTYPES = {torch.float32: _lib.DTYPE_F32, torch.bfloat16: _lib.DTYPE_BF16}


def fused_supported(value, raw, ref_dim, n_levels, n_points):
    """True when the fused kernels cover this call (else use the unfused op)."""
    if not (value.is_cuda and raw.is_cuda) or value.dtype not in _DTYPES or raw.dtype not in _DTYPES:
        return False
    if value.dtype == torch.float32 and raw.dtype != torch.float32:
        return False
    _, s, m, d = value.shape
    return bool(_lib.load().msda_fused_supported(_DTYPES[value.dtype], _DTYPES[raw.dtype], int(ref_dim), int(s),
                                                 int(m), int(d), int(n_levels), int(n_points)))


def _pixel_strided(value):
    """Pixel stride (elements) of a [N,S,M,D] value view whose pixels are dense rows of M*D elements spaced evenly
    -- a column slice of a wider projection output -- or 0 when the view is not of that form."""
    n, s, m, d = value.shape
    ld = value.stride(1)
    if value.stride(3) != 1 or value.stride(2) != d or ld < m * d or (n > 1 and value.stride(0) != s * ld):
        return 0
    vec = 16 // value.element_size()
    if ld % vec != 0 or value.data_ptr() % 16 != 0 or s * ld >= 2 ** 31:
        return 0
    return ld


def _aligned(t, nbytes=16):
    return t if t.data_ptr() % nbytes == 0 else t.clone()


class MSDeformAttnFusedFunction(Function):
    """apply(value[N,S,M,D], spatial_shapes[L,2], level_start_index[L], reference_points[N,Lq,L,2|4],
             raw[N,Lq,3*M*L*P], n_points) -> [N, Lq, M*D]
    Differentiable w.r.t. value, reference_points and raw."""

    @staticmethod
    def forward(ctx, value, spatial_shapes, level_start_index, reference_points, raw, n_points):
        n, s, m, d = value.shape
        nl = spatial_shapes.size(0)
        lq = raw.size(1)
        p = int(n_points)
        mlp = m * nl * p
        if raw.size(-1) != 3 * mlp or raw.size(0) != n:
            raise RuntimeError(f"raw must be [N, Lq, 3*M*L*P={3 * mlp}], got {tuple(raw.shape)}")
        if tuple(reference_points.shape[:3]) != (n, lq, nl) or reference_points.size(-1) not in (2, 4):
            raise RuntimeError(f"reference_points must be [N={n}, Lq={lq}, L={nl}, 2|4], got "
                               f"{tuple(reference_points.shape)}")
        if spatial_shapes.dtype != torch.int64 or level_start_index.dtype != torch.int64:
            raise RuntimeError("spatial_shapes and level_start_index must be int64 (torch.long)")
        # a strided column slice is read in place when no gradient will be asked for (backward needs it dense)
        ld = 0
        if not value.is_contiguous() and not any(ctx.needs_input_grad):
            ld = _pixel_strided(value)
        if ld == 0:
            value = _aligned(value.contiguous())
        raw = _aligned(raw.contiguous())
        ref = _aligned(reference_points.detach().to(torch.float32).contiguous())
        spatial_shapes = spatial_shapes.contiguous()
        level_start_index = level_start_index.contiguous()
        lib = _lib.load()
        with torch.cuda.device(value.device):
            out = torch.empty((n, lq, m * d), dtype=value.dtype, device=value.device)
            esz = raw.element_size()
            common = (spatial_shapes.data_ptr(), level_start_index.data_ptr(), ref.data_ptr(), ref.size(-1),
                      raw.data_ptr(), 3 * mlp, raw.data_ptr() + 2 * mlp * esz, 3 * mlp, n, s, m, d, nl, lq, p,
                      out.data_ptr())
            stream = torch.cuda.current_stream().cuda_stream
            if ld:
                code = lib.msda_fused_forward_strided(_DTYPES[value.dtype], _DTYPES[raw.dtype], value.data_ptr(), ld,
                                                      *common, stream)
            else:
                code = lib.msda_fused_forward(_DTYPES[value.dtype], _DTYPES[raw.dtype], value.data_ptr(),
                                              *common, stream)
        _lib.check(code, "msda_fused_forward")
        ctx.save_for_backward(value, spatial_shapes, level_start_index, ref, raw)
        ctx.n_points = p
        ctx.ref_dtype = reference_points.dtype
        ctx.ref_needs_grad = reference_points.requires_grad
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, spatial_shapes, level_start_index, ref, raw = ctx.saved_tensors
        n, s, m, d = value.shape
        nl = spatial_shapes.size(0)
        lq = raw.size(1)
        p = ctx.n_points
        mlp = m * nl * p
        grad_output = _aligned(grad_output.contiguous())
        if grad_output.dtype != value.dtype:
            grad_output = grad_output.to(value.dtype)
        lib = _lib.load()
        with torch.cuda.device(value.device):
            grad_value = torch.empty_like(value)
            grad_raw = torch.empty_like(raw)
            grad_ref = torch.zeros_like(ref) if ctx.ref_needs_grad else None
            accum = torch.empty(value.shape, dtype=torch.float32, device=value.device) \
                if value.dtype != torch.float32 else None
            esz = raw.element_size()
            code = lib.msda_fused_backward(
                _DTYPES[value.dtype], _DTYPES[raw.dtype], grad_output.data_ptr(), value.data_ptr(),
                spatial_shapes.data_ptr(), level_start_index.data_ptr(), ref.data_ptr(), ref.size(-1),
                raw.data_ptr(), 3 * mlp, raw.data_ptr() + 2 * mlp * esz, 3 * mlp,
                n, s, m, d, nl, lq, p, grad_value.data_ptr(), grad_raw.data_ptr(),
                grad_raw.data_ptr() + 2 * mlp * esz,
                grad_ref.data_ptr() if grad_ref is not None else None,
                accum.data_ptr() if accum is not None else None, torch.cuda.current_stream().cuda_stream)
        _lib.check(code, "msda_fused_backward")
        if grad_ref is not None and grad_ref.dtype != ctx.ref_dtype:
            grad_ref = grad_ref.to(ctx.ref_dtype)
        return grad_value, None, None, grad_ref, grad_raw, None
