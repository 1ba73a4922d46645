"""CPU oracle for multi-scale deformable attention (MSDA).

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this file;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs use it, and only as the checker or the CPU arm.

Two independent restatements of the reference algorithm live here:

* :func:`core_pytorch` -- the reference's own debug path
  ``ms_deform_attn_core_pytorch`` (``models/ops/functions/ms_deform_attn_func.py:41-61``):
  per level, view the value slab as an image, ``grid_sample`` (bilinear, zero
  padding, ``align_corners=False``) at ``2*loc-1``, weight and sum.  Backward is
  PyTorch autograd through it.  This is what the reference's ``test.py`` compares
  its CUDA op against (``models/ops/test.py:38-40``), so it is *the* parity oracle.
* :func:`forward_np` / :func:`backward_np` -- a numpy restatement of the arithmetic
  of the reference CUDA kernels (forward ``cuda/ms_deform_im2col_cuda.cuh:272-297``
  with the bilinear helper ``:33-84``; backward ``:87-158``), written with explicit
  corner gathers and explicit gradient formulas, no autograd.  It exists so that the
  gradient formulas the CUDA kernels implement are pinned by something other than
  PyTorch's grid_sample backward.

Pinning: both are checked against golden vectors produced by importing the *real*
reference module in the dev container (``oracle/gen_golden.py`` ->
``tests/golden/*.npz``), including the known-answer vector of the reference test
recipe (``models/ops/test.py:21-36``, seed 3).
"""
from __future__ import annotations

import numpy as np

try:  # torch is needed only by core_pytorch; numpy paths work without it
    import torch
    import torch.nn.functional as F
except Exception:  # pragma: no cover
    torch = None


# --------------------------------------------------------------------------- #
# 1. grid_sample restatement (reference func.py:41-61)
# --------------------------------------------------------------------------- #
def core_pytorch(value, spatial_shapes, sampling_locations, attention_weights):
    """value [N,S,M,D]; spatial_shapes iterable of (H,W); sampling_locations
    [N,Lq,M,L,P,2] (x,y in [0,1]); attention_weights [N,Lq,M,L,P] -> [N,Lq,M*D]."""
    n, _, m, d = value.shape
    _, lq, _, nl, npt, _ = sampling_locations.shape
    hw = [(int(h), int(w)) for h, w in spatial_shapes]
    per_level = value.split([h * w for h, w in hw], dim=1)
    grids = 2 * sampling_locations - 1
    sampled = []
    for lvl, (h, w) in enumerate(hw):
        # [N, H*W, M, D] -> [N*M, D, H, W]
        img = per_level[lvl].flatten(2).transpose(1, 2).reshape(n * m, d, h, w)
        # [N, Lq, M, P, 2] -> [N*M, Lq, P, 2]
        g = grids[:, :, :, lvl].transpose(1, 2).flatten(0, 1)
        sampled.append(F.grid_sample(img, g, mode="bilinear", padding_mode="zeros",
                                     align_corners=False))      # [N*M, D, Lq, P]
    aw = attention_weights.transpose(1, 2).reshape(n * m, 1, lq, nl * npt)
    out = (torch.stack(sampled, dim=-2).flatten(-2) * aw).sum(-1).view(n, m * d, lq)
    return out.transpose(1, 2).contiguous()


def core_pytorch_fwd_bwd(value, spatial_shapes, sampling_locations, attention_weights, grad_output):
    """Forward + autograd backward through :func:`core_pytorch`.  Returns
    (out, grad_value, grad_loc, grad_attn)."""
    v = value.detach().clone().requires_grad_(True)
    loc = sampling_locations.detach().clone().requires_grad_(True)
    aw = attention_weights.detach().clone().requires_grad_(True)
    out = core_pytorch(v, spatial_shapes, loc, aw)
    gv, gl, ga = torch.autograd.grad(out, (v, loc, aw), grad_output)
    return out.detach(), gv, gl, ga


# --------------------------------------------------------------------------- #
# 2. numpy restatement of the CUDA kernel arithmetic
# --------------------------------------------------------------------------- #
def _corner_terms(loc, h, w, dtype):
    """Pixel-space quantities for one level.  loc [...,2] normalised (x,y)."""
    x = loc[..., 0].astype(dtype) * dtype(w) - dtype(0.5)        # cuh:286
    y = loc[..., 1].astype(dtype) * dtype(h) - dtype(0.5)        # cuh:285
    inside = (y > -1) & (x > -1) & (y < h) & (x < w)             # cuh:288
    x0 = np.floor(x)
    y0 = np.floor(y)
    lw = x - x0
    lh = y - y0
    hw_ = 1 - lw
    hh = 1 - lh
    x0 = x0.astype(np.int64)
    y0 = y0.astype(np.int64)
    return inside, x0, y0, lw, lh, hw_, hh


def forward_np(value, spatial_shapes, level_start_index, loc, attn):
    """Arithmetic of ``ms_deformable_im2col_gpu_kernel`` (cuh:237-299).  Returns
    [N,Lq,M*D] in value.dtype."""
    value = np.asarray(value)
    dtype = value.dtype.type
    n, s, m, d = value.shape
    _, lq, _, nl, npt, _ = loc.shape
    out = np.zeros((n, lq, m, d), dtype=value.dtype)
    ni = np.arange(n)[:, None, None, None]
    mi = np.arange(m)[None, None, :, None]
    for lvl in range(nl):
        h, w = int(spatial_shapes[lvl][0]), int(spatial_shapes[lvl][1])
        st = int(level_start_index[lvl])
        img = value[:, st:st + h * w].reshape(n, h, w, m, d)
        inside, x0, y0, lw, lh, hw_, hh = _corner_terms(loc[:, :, :, lvl], h, w, dtype)
        a = attn[:, :, :, lvl].astype(dtype)
        for dy, dx, wt in ((0, 0, hh * hw_), (0, 1, hh * lw), (1, 0, lh * hw_), (1, 1, lh * lw)):
            yy = y0 + dy
            xx = x0 + dx
            ok = inside & (yy >= 0) & (yy <= h - 1) & (xx >= 0) & (xx <= w - 1)   # cuh:57-80
            vals = img[ni, np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1), mi]       # [N,Lq,M,P,D]
            coef = np.where(ok, wt * a, dtype(0))
            out += (coef[..., None] * vals).sum(axis=3)
    return out.reshape(n, lq, m * d)


def backward_np(value, spatial_shapes, level_start_index, loc, attn, grad_output):
    """Arithmetic of ``ms_deform_attn_col2im_bilinear`` (cuh:87-158) summed over
    channels the way every backward kernel variant reduces it (cuh:376-394).
    Returns (grad_value, grad_loc, grad_attn)."""
    value = np.asarray(value)
    dtype = value.dtype.type
    n, s, m, d = value.shape
    _, lq, _, nl, npt, _ = loc.shape
    g = np.asarray(grad_output).reshape(n, lq, m, d).astype(value.dtype)
    grad_value = np.zeros_like(value)
    grad_loc = np.zeros(loc.shape, dtype=value.dtype)
    grad_attn = np.zeros(attn.shape, dtype=value.dtype)
    ni = np.arange(n)[:, None, None, None]
    mi = np.arange(m)[None, None, :, None]
    nb = np.broadcast_to(ni, (n, lq, m, npt))
    mb = np.broadcast_to(mi, (n, lq, m, npt))
    for lvl in range(nl):
        h, w = int(spatial_shapes[lvl][0]), int(spatial_shapes[lvl][1])
        st = int(level_start_index[lvl])
        img = value[:, st:st + h * w].reshape(n, h, w, m, d)
        gimg = np.zeros_like(img)
        inside, x0, y0, lw, lh, hw_, hh = _corner_terms(loc[:, :, :, lvl], h, w, dtype)
        a = attn[:, :, :, lvl].astype(dtype)
        gh = np.zeros((n, lq, m, npt, d), dtype=value.dtype)   # d val / d h   per channel
        gw = np.zeros_like(gh)                                   # d val / d w
        val = np.zeros_like(gh)
        tg = g[:, :, :, None, :] * a[..., None]                  # top_grad * attn  (cuh:113)
        corners = (
            (0, 0, hh * hw_, -hw_, -hh),      # v1: grad_h -= hw*v1 ; grad_w -= hh*v1  (cuh:121-124)
            (0, 1, hh * lw, -lw, +hh),        # v2                                        (cuh:130-133)
            (1, 0, lh * hw_, +hw_, -lh),      # v3                                        (cuh:139-142)
            (1, 1, lh * lw, +lw, +lh),        # v4                                        (cuh:148-151)
        )
        for dy, dx, wt, ch, cw in corners:
            yy = y0 + dy
            xx = x0 + dx
            ok = inside & (yy >= 0) & (yy <= h - 1) & (xx >= 0) & (xx <= w - 1)
            yc = np.clip(yy, 0, h - 1)
            xc = np.clip(xx, 0, w - 1)
            vals = np.where(ok[..., None], img[ni, yc, xc, mi], dtype(0))
            gh += ch[..., None] * vals
            gw += cw[..., None] * vals
            val += wt[..., None] * vals
            contrib = np.where(ok[..., None], wt[..., None] * tg, dtype(0))
            np.add.at(gimg, (nb, yc, xc, mb), contrib)          # atomicAdd(grad_value+ptr, w*top_grad_value)
        grad_attn[:, :, :, lvl] = np.where(inside, (g[:, :, :, None, :] * val).sum(-1), dtype(0))   # cuh:156
        grad_loc[:, :, :, lvl, :, 0] = np.where(inside, dtype(w) * (gw * tg).sum(-1), dtype(0))      # cuh:157
        grad_loc[:, :, :, lvl, :, 1] = np.where(inside, dtype(h) * (gh * tg).sum(-1), dtype(0))      # cuh:158
        grad_value[:, st:st + h * w] += gimg.reshape(n, h * w, m, d)
    return grad_value, grad_loc, grad_attn


# --------------------------------------------------------------------------- #
# helpers shared by tests and bench
# --------------------------------------------------------------------------- #
def level_start_index_of(spatial_shapes):
    """[0, H0*W0, H0*W0+H1*W1, ...] (reference test.py:24)."""
    sizes = [int(h) * int(w) for h, w in spatial_shapes]
    out, acc = [], 0
    for sz in sizes:
        out.append(acc)
        acc += sz
    return out


def normalised_errors(x, ref):
    """(max|x-ref| / max|ref|, ||x-ref||_2 / ||ref||_2) -- the tolerance metric of
    BASELINE.md section 4 / SURVEY.md 8(c)."""
    x = np.asarray(x, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    scale = max(np.abs(ref).max(), 1e-300)
    l2 = max(np.sqrt((ref ** 2).sum()), 1e-300)
    return float(np.abs(x - ref).max() / scale), float(np.sqrt(((x - ref) ** 2).sum()) / l2)
