"""Sine position embedding in the layout the deformable attention wants (the step that feeds
``pos`` / ``lvl_pos_embed_flatten`` of every encoder layer).

The reference builds ``[N, 2F, H, W]`` per level with ~12 elementwise launches (``PositionEmbeddingSine.forward``,
/root/reference/models/position_encoding.py:35-56), the backbone joiner casts it to the feature dtype
(/root/reference/models/backbone_scratch.py:185), and ``DeformableTransformer.forward`` flattens,
transposes, adds ``level_embed[l]`` and concatenates the levels (deformable_transformer_single.py:190-206).

``PositionEmbeddingSine`` here has the reference's constructor and ``forward(tensor_list)`` (same ``[N, 2F, H, W]``
values; on a CUDA mask it is a permuted view of the token-major kernel output), plus ``forward_tokens``: all levels
written by two kernels per level (coordinates, embedding) straight into ``lvl_pos_embed_flatten [N, sum_l H_l*W_l, 2F]`` (C ABI
``msda_layer_sine_position_tokens``, csrc/layer_epilogue.cu).  ``DeformableTransformer.forward`` accepts that tensor in
place of the per-level ``pos_embeds`` list.  The cumulative coordinates (``[N, H, W]`` fp32, 1/2F of the output) come
from ``msda_layer_sine_coordinates``, which reproduces the reference's cumsum / normalise ops bit for bit.
"""
import math

import torch
from torch import nn

from . import _lib

_DTYPES = {torch.float32: _lib.DTYPE_F32, torch.bfloat16: _lib.DTYPE_BF16, torch.float16: _lib.DTYPE_F16}


def _mask_of(tensor_list):
    """NestedTensor-like (``.mask``), ``(tensors, mask)`` pair, or the bool mask ``[N, H, W]`` itself."""
    if isinstance(tensor_list, torch.Tensor):
        mask = tensor_list
    elif isinstance(tensor_list, (tuple, list)):
        mask = tensor_list[1]
    else:
        mask = tensor_list.mask
    assert mask is not None
    return mask


class PositionEmbeddingSine(nn.Module):
    """Drop-in for /root/reference/models/position_encoding.py:20-56 (no parameters, no state-dict entries)."""

    def __init__(self, num_pos_feats=64, temperature=10000, normalize=False, scale=None):
        super().__init__()
        self.num_pos_feats = num_pos_feats
        self.temperature = temperature
        self.normalize = normalize
        if scale is not None and normalize is False:
            raise ValueError("normalize should be True if scale is passed")
        if scale is None:
            scale = 2 * math.pi
        self.scale = scale

    def _coordinates(self, mask):
        """position_encoding.py:39-46."""
        not_mask = ~mask
        y_embed = not_mask.cumsum(1, dtype=torch.float32)
        x_embed = not_mask.cumsum(2, dtype=torch.float32)
        if self.normalize:
            eps = 1e-6
            y_embed = (y_embed - 0.5) / (y_embed[:, -1:, :] + eps) * self.scale
            x_embed = (x_embed - 0.5) / (x_embed[:, :, -1:] + eps) * self.scale
        return y_embed, x_embed

    def _device_coordinates(self, mask, stream):
        """`_coordinates` as one kernel (C ABI ``msda_layer_sine_coordinates``; bit-identical maps)."""
        n, h, w = mask.shape
        mask = mask.contiguous()
        y_embed = torch.empty((n, h, w), dtype=torch.float32, device=mask.device)
        x_embed = torch.empty_like(y_embed)
        code = _lib.load().msda_layer_sine_coordinates(mask.data_ptr(), n, h, w, int(bool(self.normalize)),
                                                       float(self.scale), y_embed.data_ptr(), x_embed.data_ptr(), stream)
        _lib.check(code, "msda_layer_sine_coordinates")
        return y_embed, x_embed

    def _dim_t(self, device):
        """position_encoding.py:48-49 (a constant of the module: computed once per device)."""
        key = (str(device), self.num_pos_feats, self.temperature)
        cached = getattr(self, "_dim_t_cache", None)
        if cached is not None and cached[0] == key:
            return cached[1]
        dim_t = torch.arange(self.num_pos_feats, dtype=torch.float32, device=device)
        dim_t = self.temperature ** (2 * (dim_t // 2) / self.num_pos_feats)
        self._dim_t_cache = (key, dim_t)
        return dim_t

    def _host_composition(self, mask):
        """What position_encoding.py:51-56 evaluates, for masks that live on the host: channel k of the y block is
        sin(y_embed / dim_t[k]) for even k and cos for odd k (the reference interleaves the 0::2 sines with the 1::2
        cosines), then the x block.  Returns [N, H, W, 2F]."""
        odd = (torch.arange(self.num_pos_feats, device=mask.device) % 2).bool()
        dim_t = self._dim_t(mask.device)
        blocks = []
        for coord in self._coordinates(mask):                       # y first, then x
            arg = coord.unsqueeze(-1) / dim_t
            blocks.append(torch.where(odd, arg.cos(), arg.sin()))
        return torch.cat(blocks, dim=-1)

    def forward_tokens(self, masks, level_embed=None, dtype=torch.float32):
        """masks: one bool ``[N, H_l, W_l]`` per level (True = padding).  Returns ``[N, sum_l H_l*W_l, 2F]`` of `dtype`:
        level l's slice is ``cast(pos_l).flatten(2).transpose(1, 2) + level_embed[l]``, i.e. the reference's
        ``lvl_pos_embed_flatten`` (deformable_transformer_single.py:196-206).  Not differentiable (``level_embed`` is
        read as a constant: callers that train it keep the per-level list)."""
        masks = [_mask_of(m) for m in masks]
        if not masks[0].is_cuda:
            flat = []
            for lvl, mask in enumerate(masks):
                pos = self._host_composition(mask).to(dtype).flatten(1, 2)
                flat.append(pos if level_embed is None else pos + level_embed[lvl].detach().to(dtype).view(1, 1, -1))
            return flat[0] if len(flat) == 1 else torch.cat(flat, 1)
        if dtype not in _DTYPES:
            raise TypeError(f"sine position tokens: unsupported dtype {dtype}")
        lib = _lib.load()
        n, f = masks[0].shape[0], int(self.num_pos_feats)
        sizes = [m.shape[1] * m.shape[2] for m in masks]
        total = sum(sizes)
        device = masks[0].device
        with torch.cuda.device(device):
            dim_t = self._dim_t(device).contiguous()
            out = torch.empty((n, total, 2 * f), dtype=dtype, device=device)
            stream = torch.cuda.current_stream().cuda_stream
            start = 0
            for lvl, (mask, hw) in enumerate(zip(masks, sizes)):
                if mask.dtype == torch.bool:
                    y_embed, x_embed = self._device_coordinates(mask, stream)
                else:
                    y_embed, x_embed = (t.contiguous() for t in self._coordinates(mask))
                add = None if level_embed is None else level_embed[lvl].detach().to(dtype).contiguous()
                code = lib.msda_layer_sine_position_tokens(
                    _DTYPES[dtype], y_embed.data_ptr(), x_embed.data_ptr(), dim_t.data_ptr(), f,
                    None if add is None else add.data_ptr(), n, hw, out.data_ptr(), total, start, stream)
                _lib.check(code, "msda_layer_sine_position_tokens")
                start += hw
        return out

    def forward(self, tensor_list):
        """``[N, 2F, H, W]`` fp32 as the reference returns it (position_encoding.py:35-56)."""
        mask = _mask_of(tensor_list)
        n, h, w = mask.shape
        if mask.is_cuda:
            tokens = self.forward_tokens([mask])
        else:
            tokens = self._host_composition(mask).flatten(1, 2)
        return tokens.view(n, h, w, 2 * self.num_pos_feats).permute(0, 3, 1, 2)


def build_position_encoding(args):
    """position_encoding.py:86-97 for the sine variants ('v2' / 'sine'); the learned table is a plain nn.Embedding
    pair with no kernel of ours behind it and stays with the reference."""
    n_steps = args.hidden_dim // 2
    if args.position_embedding in ("v2", "sine"):
        return PositionEmbeddingSine(n_steps, normalize=True)
    raise ValueError(f"not supported {args.position_embedding}")
