"""Tensor-core (tcgen05) deformable-attention kernels against the lane-group kernels and the fp64 oracle, plus timing.

    python tools/run_tc_check.py [fwd|bwd|all] [--time]

Run on a B200 (under `timeout`: a wrong barrier protocol hangs).  Cases: the bench's encoder geometry (grid + compass +
N(0,1) px), a freshly initialised layer (exact compass offsets), uniform-random locations (every tile falls back to the
in-kernel gather), one level 50x84, odd shapes with out-of-map samples, P = 2, linear tiles (Lq != S).
"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dfvod_b200  # noqa: E402
from dfvod_b200 import _lib  # noqa: E402
from oracle import msda_oracle  # noqa: E402

DT = {torch.float32: _lib.DTYPE_F32, torch.bfloat16: _lib.DTYPE_BF16}


def level_start(shapes):
    lsi, acc = [], 0
    for h, w in shapes:
        lsi.append(acc)
        acc += h * w
    return lsi, acc


def make_case(shapes, n, m, p, dist, seed, lq=None, d=32):
    g = torch.Generator().manual_seed(seed)
    lsi, s = level_start(shapes)
    nl = len(shapes)
    pyramid = lq is None
    lq = s if lq is None else lq
    value = torch.randn(n, s, m, d, generator=g)
    attn = torch.softmax(torch.randn(n, lq, m, nl * p, generator=g), -1).view(n, lq, m, nl, p)
    if dist == "random":
        loc = torch.rand(n, lq, m, nl, p, 2, generator=g) * 1.3 - 0.15
    else:
        if pyramid:
            ref = []
            for h, w in shapes:
                ys = (torch.arange(h, dtype=torch.float32) + 0.5) / h
                xs = (torch.arange(w, dtype=torch.float32) + 0.5) / w
                yy, xx = torch.meshgrid(ys, xs, indexing="ij")
                ref.append(torch.stack([xx.reshape(-1), yy.reshape(-1)], -1))
            ref = torch.cat(ref, 0)
        else:   # a smooth curve of reference points: neighbouring queries are neighbours in the map
            tq = torch.arange(lq, dtype=torch.float32) / lq
            ref = torch.stack([0.5 + 0.45 * torch.cos(40 * tq), tq], -1)
        ang = torch.arange(m, dtype=torch.float32) * (2.0 * math.pi / m)
        comp = torch.stack([ang.cos(), ang.sin()], -1)
        comp = comp / comp.abs().max(-1, keepdim=True)[0]
        steps = torch.arange(1, p + 1, dtype=torch.float32)
        off = comp[:, None, None, :] * steps[None, None, :, None]
        noise = torch.randn(n, lq, m, nl, p, 2, generator=g)
        if dist == "init":
            noise = noise * 0.0
        off = off.expand(m, nl, p, 2) + noise
        norm = torch.tensor([[w, h] for h, w in shapes], dtype=torch.float32)
        loc = ref[None, :, None, None, None, :] + off / norm[None, None, None, :, None, :]
    gout = torch.randn(n, lq, m * d, generator=g)
    return value, loc.contiguous(), attn.contiguous(), gout, lsi


def fwd_call(value, st, ls, loc, attn, flags):
    n, s, m, d = value.shape
    lq, nl, p = loc.shape[1], loc.shape[3], loc.shape[4]
    out = torch.empty(n, lq, m * d, dtype=value.dtype, device=value.device)
    code = _lib.load().msda_forward(DT[value.dtype], value.data_ptr(), st.data_ptr(), ls.data_ptr(), loc.data_ptr(),
                                    attn.data_ptr(), n, s, m, d, nl, lq, p, out.data_ptr(), flags,
                                    torch.cuda.current_stream().cuda_stream)
    _lib.check(code, "msda_forward")
    return out


def bwd_call(value, st, ls, loc, attn, gout, flags):
    n, s, m, d = value.shape
    lq, nl, p = loc.shape[1], loc.shape[3], loc.shape[4]
    gv = torch.empty_like(value)
    gl = torch.empty_like(loc)
    ga = torch.empty_like(attn)
    accum = torch.empty(value.shape, dtype=torch.float32, device=value.device) if value.dtype != torch.float32 else None
    code = _lib.load().msda_backward(DT[value.dtype], gout.data_ptr(), value.data_ptr(), st.data_ptr(), ls.data_ptr(),
                                     loc.data_ptr(), attn.data_ptr(), n, s, m, d, nl, lq, p, gv.data_ptr(), gl.data_ptr(),
                                     ga.data_ptr(), accum.data_ptr() if accum is not None else None, flags,
                                     torch.cuda.current_stream().cuda_stream)
    _lib.check(code, "msda_backward")
    return gv, gl, ga


def nerr(x, ref):
    x, ref = x.double().cpu().reshape(-1), ref.double().cpu().reshape(-1)
    return float((x - ref).abs().max() / ref.abs().max().clamp_min(1e-300)), \
        float((x - ref).norm() / ref.norm().clamp_min(1e-300))


def time_ms(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


COCO = [(100, 167), (50, 84), (25, 42), (13, 21)]
CASES = [
    ("coco grid n=2", COCO, 2, 8, 4, "grid", 0, None),
    ("coco init n=1", COCO, 1, 8, 4, "init", 1, None),
    ("coco random n=1", COCO, 1, 8, 4, "random", 2, None),
    ("one level 50x84 n=4", [(50, 84)], 4, 8, 4, "grid", 3, None),
    ("odd shapes n=3 m=3 p=2", [(37, 53), (19, 27)], 3, 3, 2, "grid", 4, None),
    ("linear tiles lq=3000", [(40, 60), (20, 30)], 2, 8, 4, "grid", 5, 3000),
    ("small 3 levels p=3", [(30, 45), (15, 23), (8, 12)], 2, 4, 3, "grid", 6, None),
]


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    timing = "--time" in sys.argv
    dev = torch.device("cuda:0")
    ok = True
    for name, shapes, n, m, p, dist, seed, lq in CASES:
        value, loc, attn, gout, lsi = make_case(shapes, n, m, p, dist, seed, lq)
        st = torch.as_tensor(shapes, dtype=torch.long, device=dev)
        ls = torch.as_tensor(lsi, dtype=torch.long, device=dev)
        vb = value.to(torch.bfloat16)
        gb = gout.to(torch.bfloat16)
        vd, ld, ad, gd = vb.to(dev), loc.to(dev), attn.to(dev), gb.to(dev)
        if what in ("fwd", "all"):
            ref = msda_oracle.core_pytorch(vb.double(), shapes, loc.double(), attn.double())
            o_tc = fwd_call(vd, st, ls, ld, ad, _lib.FLAG_TC)
            o_lg = fwd_call(vd, st, ls, ld, ad, 0)
            torch.cuda.synchronize()
            e_tc, e_lg = nerr(o_tc, ref), nerr(o_lg, ref)
            good = e_tc[0] <= 2.0 ** -7 and e_tc[1] <= 4e-3
            ok &= good
            print(f"fwd {name:28s} tc max {e_tc[0]:.2e} l2 {e_tc[1]:.2e} | lane-group max {e_lg[0]:.2e} l2 {e_lg[1]:.2e}"
                  f"  {'OK' if good else 'FAIL'}", flush=True)
        if what in ("bwd", "all"):
            v64 = vb.double().requires_grad_(True)
            l64 = loc.double().requires_grad_(True)
            a64 = attn.double().requires_grad_(True)
            msda_oracle.core_pytorch(v64, shapes, l64, a64).backward(gb.double())
            refs = (v64.grad, l64.grad, a64.grad)
            g_tc = bwd_call(vd, st, ls, ld, ad, gd, _lib.FLAG_TC)
            g_lg = bwd_call(vd, st, ls, ld, ad, gd, 0)
            torch.cuda.synchronize()
            for nm, x, y, r in zip(("grad_value", "grad_loc", "grad_attn"), g_tc, g_lg, refs):
                e_tc, e_lg = nerr(x, r), nerr(y, r)
                good = e_tc[0] <= 2.0 ** -7 and e_tc[1] <= 4e-3
                ok &= good
                print(f"bwd {name:28s} {nm:10s} tc max {e_tc[0]:.2e} l2 {e_tc[1]:.2e} | lane-group max {e_lg[0]:.2e} "
                      f"l2 {e_lg[1]:.2e}  {'OK' if good else 'FAIL'}", flush=True)
    if timing:
        for dist in ("grid", "init", "random"):
            value, loc, attn, gout, lsi = make_case(COCO, 8, 8, 4, dist, 0)
            st = torch.as_tensor(COCO, dtype=torch.long, device=dev)
            ls = torch.as_tensor(lsi, dtype=torch.long, device=dev)
            vd, ld, ad = value.to(torch.bfloat16).to(dev), loc.to(dev), attn.to(dev)
            gd = gout.to(torch.bfloat16).to(dev)
            if what in ("fwd", "all"):
                t_tc = time_ms(lambda: fwd_call(vd, st, ls, ld, ad, _lib.FLAG_TC))
                t_lg = time_ms(lambda: fwd_call(vd, st, ls, ld, ad, 0))
                print(f"time fwd bf16 batch 8 {dist:6s}: tc {t_tc:.3f} ms   lane-group {t_lg:.3f} ms", flush=True)
            if what in ("bwd", "all"):
                t_tc = time_ms(lambda: bwd_call(vd, st, ls, ld, ad, gd, _lib.FLAG_TC))
                t_lg = time_ms(lambda: bwd_call(vd, st, ls, ld, ad, gd, 0))
                print(f"time bwd bf16 batch 8 {dist:6s}: tc {t_tc:.3f} ms   lane-group {t_lg:.3f} ms", flush=True)
    print("ALL OK" if ok else "SOME FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
