"""Shared helpers for the test-suite (and only the test-suite)."""
import ctypes
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

COCO_SHAPES = [(100, 167), (50, 84), (25, 42), (13, 21)]      # 800x1333 pyramid, S = 22223


def load_golden(name):
    return dict(np.load(os.path.join(GOLD, name + ".npz")))


def lsi_of(shapes):
    out, acc = [], 0
    for h, w in shapes:
        out.append(acc)
        acc += int(h) * int(w)
    return out


def nudge_off_pixel_boundaries(loc, shapes, eps=2e-3):
    """grad_loc is discontinuous where a sample crosses a pixel boundary (and where it enters the
    map), so an fp32 evaluation and the fp64 oracle may legitimately pick different cells for a
    sample that sits within rounding distance of an integer pixel coordinate.  Move such samples
    `eps` pixels away so both sides differentiate the same bilinear cell."""
    loc = loc.clone().double()
    for lvl, (h, w) in enumerate(shapes):
        for axis, size in ((0, w), (1, h)):
            px = loc[:, :, :, lvl, :, axis] * size - 0.5
            near = (px - px.round()).abs() < eps
            loc[:, :, :, lvl, :, axis] = torch.where(near, (px.round() + 2 * eps + 0.5) / size,
                                                     loc[:, :, :, lvl, :, axis])
    return loc.float()


def make_inputs(shapes, n, m, d, lq, p, seed, dist="random", loc_range=(0.0, 1.0), dtype=torch.float32):
    """CPU-generated (so CPU oracle and GPU op see identical bits) op-level inputs.
    dist="random": loc ~ U[loc_range) as in the reference test (models/ops/test.py:34).
    dist="grid":   what MSDeformAttn.forward produces in the encoder -- the pixel-centre grid
                   plus the init compass offsets (+ N(0,1) px noise), per level."""
    g = torch.Generator().manual_seed(seed)
    nl = len(shapes)
    s = sum(h * w for h, w in shapes)
    value = torch.randn(n, s, m, d, generator=g)
    attn = torch.softmax(torch.randn(n, lq, m, nl * p, generator=g), -1).view(n, lq, m, nl, p)
    if dist == "random":
        lo, hi = loc_range
        loc = torch.rand(n, lq, m, nl, p, 2, generator=g) * (hi - lo) + lo
    else:
        assert lq == s, "grid distribution is the encoder self-attention case (Lq == S)"
        ref = []
        for h, w in shapes:
            ys = (torch.arange(h, dtype=torch.float32) + 0.5) / h
            xs = (torch.arange(w, dtype=torch.float32) + 0.5) / w
            yy, xx = torch.meshgrid(ys, xs, indexing="ij")
            ref.append(torch.stack([xx.reshape(-1), yy.reshape(-1)], -1))
        ref = torch.cat(ref, 0)                                          # [S,2]
        ang = torch.arange(m, dtype=torch.float32) * (2.0 * np.pi / m)
        comp = torch.stack([ang.cos(), ang.sin()], -1)
        comp = comp / comp.abs().max(-1, keepdim=True)[0]                # [M,2]
        steps = torch.arange(1, p + 1, dtype=torch.float32)
        off = comp[:, None, None, :] * steps[None, None, :, None]        # [M,1,P,2]
        off = off.expand(m, nl, p, 2) + torch.randn(n, lq, m, nl, p, 2, generator=g)
        norm = torch.tensor([[w, h] for h, w in shapes], dtype=torch.float32)   # [L,2]
        loc = ref[None, :, None, None, None, :] + off / norm[None, None, None, :, None, :]
    grad_out = torch.randn(n, lq, m * d, generator=g)
    loc = nudge_off_pixel_boundaries(loc, shapes)
    return (value.to(dtype).contiguous(), loc.to(dtype).contiguous(), attn.to(dtype).contiguous(),
            grad_out.to(dtype).contiguous())


def shapes_tensors(shapes, device):
    st = torch.as_tensor(shapes, dtype=torch.long, device=device)
    ls = torch.as_tensor(lsi_of(shapes), dtype=torch.long, device=device)
    return st, ls


def nerr(x, ref):
    """normalised max error and relative L2."""
    x = np.asarray(x, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    den = max(float(np.abs(ref).max()), 1e-300)
    l2 = max(float(np.sqrt((ref ** 2).sum())), 1e-300)
    return float(np.abs(x - ref).max() / den), float(np.sqrt(((x - ref) ** 2).sum()) / l2)


# ---------------------------------------------------------------------------
# the reference's own CUDA kernels recompiled for sm_100a (oracle/_ref, built by oracle/Makefile)
# ---------------------------------------------------------------------------
_REF = None


def ref_cuda_lib():
    global _REF
    if _REF is None:
        path = os.path.join(ROOT, "oracle", "_ref", "libmsda_ref_cuda.so")
        if not os.path.exists(path):
            return None
        _REF = ctypes.CDLL(path)
    return _REF


def ref_cuda_forward(value, shapes_t, lsi_t, loc, attn):
    lib = ref_cuda_lib()
    n, s, m, d = value.shape
    _, lq, _, nl, p, _ = loc.shape
    out = torch.zeros(n, lq, m * d, dtype=value.dtype, device=value.device)   # reference: at::zeros
    fn = lib.ref_msda_forward_f32 if value.dtype == torch.float32 else lib.ref_msda_forward_f64
    vp = ctypes.c_void_p
    code = fn(vp(torch.cuda.current_stream().cuda_stream), vp(value.data_ptr()), vp(shapes_t.data_ptr()),
              vp(lsi_t.data_ptr()), vp(loc.data_ptr()), vp(attn.data_ptr()), n, s, m, d, nl, lq, p,
              vp(out.data_ptr()))
    assert code == 0
    return out


def ref_cuda_backward(value, shapes_t, lsi_t, loc, attn, grad_out):
    lib = ref_cuda_lib()
    n, s, m, d = value.shape
    _, lq, _, nl, p, _ = loc.shape
    gv, gl, ga = torch.zeros_like(value), torch.zeros_like(loc), torch.zeros_like(attn)
    fn = lib.ref_msda_backward_f32 if value.dtype == torch.float32 else lib.ref_msda_backward_f64
    vp = ctypes.c_void_p
    code = fn(vp(torch.cuda.current_stream().cuda_stream), vp(grad_out.data_ptr()), vp(value.data_ptr()),
              vp(shapes_t.data_ptr()), vp(lsi_t.data_ptr()), vp(loc.data_ptr()), vp(attn.data_ptr()),
              n, s, m, d, nl, lq, p, vp(gv.data_ptr()), vp(gl.data_ptr()), vp(ga.data_ptr()))
    assert code == 0
    return gv, gl, ga
