"""Does a 2-D query tile per CTA help the lane-group kernels?  The kernels give a CTA 32 CONSECUTIVE queries of one head
(a 32 x 1 pixel run of the pyramid when the queries are the pyramid's pixels).  Here the query order of the SAME problem
is permuted on the host so that consecutive query ids form tw x th tiles of each level, and the unchanged kernels are
timed on it: same work, different query -> CTA assignment."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import bench
import run_tc_check as tc

dev = torch.device("cuda:0")
shapes = bench.COCO_SHAPES
lsi, s = bench.level_start(shapes)


def tile_perm(tw, th):
    idx = []
    for (h, w), st in zip(shapes, lsi):
        ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
        key = ((ys // th) * ((w + tw - 1) // tw) + xs // tw) * (tw * th) + (ys % th) * tw + xs % tw
        order = torch.argsort(key.reshape(-1), stable=True)
        idx.append(order + st)
    return torch.cat(idx)


def main():
    n = 8
    value, loc, attn, gout = bench.make_inputs(torch, n, 0, sys.argv[1] if len(sys.argv) > 1 else "grid")
    st = torch.as_tensor(shapes, dtype=torch.long, device=dev)
    ls = torch.as_tensor(lsi, dtype=torch.long, device=dev)
    for dt in (torch.float32, torch.bfloat16):
        vd = value.to(dt).to(dev)
        for name, perm in [("32x1 (as is)", None), ("16x2", tile_perm(16, 2)), ("8x4", tile_perm(8, 4)),
                           ("4x8", tile_perm(4, 8)), ("8x8", tile_perm(8, 8)), ("16x4", tile_perm(16, 4))]:
            l, a, g = (loc, attn, gout) if perm is None else (loc[:, perm], attn[:, perm], gout[:, perm])
            ld, ad, gd = l.contiguous().to(dev), a.contiguous().to(dev), g.to(dt).contiguous().to(dev)
            t_f = tc.time_ms(lambda: tc.fwd_call(vd, st, ls, ld, ad, 0))
            t_b = tc.time_ms(lambda: tc.bwd_call(vd, st, ls, ld, ad, gd, 0))
            print(f"{str(dt):16s} tile {name:14s} fwd {t_f:.3f} ms  bwd {t_b:.3f} ms", flush=True)


if __name__ == "__main__":
    main()
