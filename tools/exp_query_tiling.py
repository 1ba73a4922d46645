"""Experiment: does ordering the encoder's queries in 2-D tiles (instead of raster order) speed the
gather up?  The op is equivariant to a permutation of the queries, so the permutation is applied to
loc / attn / grad_out on the host side and the kernels are untouched."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from dfvod_b200 import MultiScaleDeformableAttention as MSDA

dev = torch.device("cuda:0")
n = 8
value, loc, attn, gout = bench.make_inputs(torch, n, 0, "grid")
lsi, s = bench.level_start(bench.COCO_SHAPES)
st = torch.as_tensor(bench.COCO_SHAPES, dtype=torch.long, device=dev)
ls = torch.as_tensor(lsi, dtype=torch.long, device=dev)


def tile_perm(tw, th):
    perm = []
    for (h, w), start in zip(bench.COCO_SHAPES, lsi):
        for ty in range(0, h, th):
            for tx in range(0, w, tw):
                for y in range(ty, min(ty + th, h)):
                    for x in range(tx, min(tx + tw, w)):
                        perm.append(start + y * w + x)
    return torch.tensor(perm, dtype=torch.long)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for dt in (torch.float32, torch.bfloat16):
    v = value.to(dev, dt)
    for name, perm in [("raster", None)] + [(f"tile{tw}x{th}", tile_perm(tw, th)) for tw, th in ((8, 8), (8, 4), (16, 4), (4, 8), (16, 8), (4, 4))]:
        l, a, g = (loc, attn, gout) if perm is None else (loc[:, perm], attn[:, perm], gout[:, perm])
        l, a, g = l.contiguous().to(dev), a.contiguous().to(dev), g.contiguous().to(dev, dt)
        f = timeit(lambda: MSDA.ms_deform_attn_forward(v, st, ls, l, a, 64))
        b = timeit(lambda: MSDA.ms_deform_attn_backward(v, st, ls, l, a, g, 64))
        print(f"{str(dt)[6:]:9s} {name:9s} fwd {f:.4f} ms  bwd {b:.4f} ms")
