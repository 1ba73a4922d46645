"""Input projections in front of the transformer: the reference projects every backbone level
with ``nn.Sequential(nn.Conv2d(in_channels, hidden_dim, kernel_size=1), nn.GroupNorm(32, hidden_dim))`` (extra levels:
a 3x3 stride-2 convolution), and the depth map with the same pair, on NCHW maps
(/root/reference/models/deformable_detr_single.py:101-150, used at :262-267, :270-281, :288-292); the transformer then
flattens and transposes each level to tokens (deformable_transformer_single.py:190-206).

``InputProjection`` IS that ``nn.Sequential`` (same sub-module indices, so the checkpoint keys ``input_proj.<l>.0.weight``
... ``input_proj.<l>.1.bias`` are unchanged and ``forward`` on an NCHW map gives the reference's NCHW result), plus
``forward_tokens``: the step written for the layout the deformable attention wants --

  * a 1x1 convolution is a GEMM; computed as ``X^T W^T`` it reads the NCHW backbone map in place (transposed operand of
    the library GEMM) and writes tokens ``[N, H*W, hidden]`` directly;
  * a k x k convolution runs channels-last, whose output is token-major memory already;
  * GroupNorm runs on the tokens in place (csrc/layer_epilogue.cu: group_norm_tokens, C ABI
    ``msda_layer_group_norm_tokens``): two HBM passes, no NCHW intermediate, no transpose, no concatenation.

``DeformableTransformer.forward`` accepts such ``[N, H*W, C]`` levels next to (or instead of) NCHW maps.
"""
import torch
import torch.nn.functional as F
from torch import nn

from . import _lib

_DTYPES = {torch.float32: _lib.DTYPE_F32, torch.bfloat16: _lib.DTYPE_BF16, torch.float16: _lib.DTYPE_F16}


def _token_slice_ok(t):
    """[N, S, C] with contiguous rows (a level's slice of the flattened multi-level token tensor qualifies)."""
    return t.dim() == 3 and t.stride(2) == 1 and t.stride(1) == t.shape[2] and t.stride(0) >= t.shape[1] * t.shape[2]


def group_norm_tokens(tokens, num_groups, weight, bias, eps=1e-5, inplace=False, pre_bias=None):
    """GroupNorm of token-major activations ``[N, S, C]`` (== ``F.group_norm`` on the ``[N, C, S]`` view), applied to
    ``tokens + pre_bias`` when a per-channel ``pre_bias`` is given (the convolution's bias).  One statistics kernel +
    one apply kernel when no gradient is needed; otherwise the PyTorch op.  ``inplace`` writes into ``tokens``, which
    may be a level's slice ``big[:, start:start + S]`` of a larger token tensor."""
    n, s, c = tokens.shape
    needs_grad = torch.is_grad_enabled() and any(
        t is not None and t.requires_grad for t in (tokens, weight, bias, pre_bias))
    slabs = 0
    if (not needs_grad and tokens.is_cuda and tokens.dtype in _DTYPES and weight is not None and bias is not None
            and s > 0 and n > 0):
        slabs = _lib.load().msda_layer_group_norm_tokens_slabs(_DTYPES[tokens.dtype], s, c, int(num_groups))
    if slabs == 0:
        x = tokens if pre_bias is None else tokens + pre_bias
        y = F.group_norm(x.transpose(1, 2), num_groups, weight, bias, eps).transpose(1, 2)
        return tokens.copy_(y) if inplace and not needs_grad else y
    vec = 16 // tokens.element_size()
    x = tokens
    if not (_token_slice_ok(x) and x.data_ptr() % 16 == 0 and x.stride(0) % vec == 0):
        x = tokens.contiguous()
        if x.data_ptr() % 16 != 0:
            x = x.clone()
    y = x if (inplace or x is not tokens) else torch.empty((n, s, c), dtype=x.dtype, device=x.device)
    if y is not x and x.stride(0) != s * c:          # out of place needs one item stride for both: densify
        x = x.contiguous()
    gamma = weight.detach().to(x.dtype).contiguous()
    beta = bias.detach().to(x.dtype).contiguous()
    pre = None if pre_bias is None else pre_bias.detach().to(x.dtype).contiguous()
    with torch.cuda.device(x.device):
        partial = torch.empty((n, slabs, int(num_groups), 2), dtype=torch.float32, device=x.device)
        code = _lib.load().msda_layer_group_norm_tokens(
            _DTYPES[x.dtype], x.data_ptr(), None if pre is None else pre.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
            n, s, c, int(num_groups), float(eps), x.stride(0), partial.data_ptr(), slabs, y.data_ptr(),
            torch.cuda.current_stream().cuda_stream)
    _lib.check(code, "msda_layer_group_norm_tokens")
    if inplace and y is not tokens:
        tokens.copy_(y)
        return tokens
    return y


class InputProjection(nn.Sequential):
    """``nn.Sequential(nn.Conv2d(in_channels, hidden_dim, kernel_size, stride, padding), nn.GroupNorm(32, hidden_dim))``
    with a token-major fast path.  Initialised like the reference (:167-170: xavier_uniform_ weight, zero bias)."""

    def __init__(self, in_channels, hidden_dim=256, kernel_size=1, stride=1, padding=0, num_groups=32):
        super().__init__(nn.Conv2d(in_channels, hidden_dim, kernel_size=kernel_size, stride=stride, padding=padding),
                         nn.GroupNorm(num_groups, hidden_dim))
        nn.init.xavier_uniform_(self[0].weight, gain=1)
        nn.init.constant_(self[0].bias, 0)

    def output_hw(self, h, w):
        conv = self[0]
        f = lambda size, i: (size + 2 * conv.padding[i] - conv.dilation[i] * (conv.kernel_size[i] - 1) - 1) // conv.stride[i] + 1
        return f(h, 0), f(w, 1)

    def _is_pointwise(self):
        conv = self[0]
        return conv.kernel_size == (1, 1) and conv.stride == (1, 1) and conv.padding == (0, 0) and conv.groups == 1

    def forward_tokens(self, x, out=None):
        """``self(x).flatten(2).transpose(1, 2)`` without the NCHW intermediate: ``([N, H'*W', hidden], (H', W'))``.
        ``out``: where to put the tokens (``[N, H'*W', hidden]``, rows contiguous), e.g. a level's slice of the
        flattened multi-level token tensor -- then no concatenation is needed either."""
        conv, norm = self
        n, cin, h, w = x.shape
        hw = self.output_hw(h, w)
        fused = not (torch.is_grad_enabled() and (x.requires_grad or conv.weight.requires_grad))
        if self._is_pointwise():
            # 1x1 convolution = GEMM; X^T is a transposed operand of the library GEMM (no copy of the backbone map),
            # the bias is added inside the GroupNorm kernels
            xt = x.reshape(n, cin, h * w).transpose(1, 2)
            wt = conv.weight.view(conv.out_channels, cin).t().expand(n, cin, conv.out_channels)
            if fused and out is not None and _token_slice_ok(out):
                tokens = torch.bmm(xt, wt, out=out)
            else:
                tokens = torch.bmm(xt, wt)
            pre_bias = conv.bias
        else:
            y = conv(x.contiguous(memory_format=torch.channels_last))          # channels-last result = token-major memory
            tokens = y.permute(0, 2, 3, 1).reshape(n, hw[0] * hw[1], conv.out_channels)
            pre_bias = None
        tokens = group_norm_tokens(tokens, norm.num_groups, norm.weight, norm.bias, norm.eps, inplace=fused,
                                   pre_bias=pre_bias)
        if out is not None and tokens.data_ptr() != out.data_ptr():
            out.copy_(tokens)
            tokens = out
        return tokens, hw


def project_levels(projections, feature_maps):
    """All levels of a pyramid through their ``InputProjection`` into ONE flattened token tensor
    ``[N, sum_l H_l*W_l, hidden]`` (what ``DeformableTransformer.forward`` builds with flatten / transpose / cat from
    the reference's per-level NCHW projections): returns ``(tokens, per-level token views, shapes)``."""
    shapes = [proj.output_hw(x.shape[2], x.shape[3]) for proj, x in zip(projections, feature_maps)]
    n = feature_maps[0].shape[0]
    hidden = projections[0][0].out_channels
    total = sum(h * w for h, w in shapes)
    big = torch.empty((n, total, hidden), dtype=feature_maps[0].dtype, device=feature_maps[0].device)
    views, start = [], 0
    needs_grad = torch.is_grad_enabled() and any(
        x.requires_grad or proj[0].weight.requires_grad for proj, x in zip(projections, feature_maps))
    if needs_grad:                                   # autograd-visible composition
        levels = [proj.forward_tokens(x)[0] for proj, x in zip(projections, feature_maps)]
        big = torch.cat(levels, 1)
    for (h, w), proj, x in zip(shapes, projections, feature_maps):
        view = big[:, start:start + h * w]
        if not needs_grad:
            proj.forward_tokens(x, out=view)
        views.append(view)
        start += h * w
    return big, views, shapes
