"""A few launches of the tensor-core forward (or backward) at the bench shape -- the command ncu wraps."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import run_tc_check as c

dev = torch.device("cuda:0")
what = sys.argv[1] if len(sys.argv) > 1 else "fwd"
dist = sys.argv[2] if len(sys.argv) > 2 else "grid"
n = int(sys.argv[3]) if len(sys.argv) > 3 else 8
value, loc, attn, gout, lsi = c.make_case(c.COCO, n, 8, 4, dist, 0)
st = torch.as_tensor(c.COCO, dtype=torch.long, device=dev)
ls = torch.as_tensor(lsi, dtype=torch.long, device=dev)
vd, ld, ad = value.to(torch.bfloat16).to(dev), loc.to(dev), attn.to(dev)
gd = gout.to(torch.bfloat16).to(dev)
for _ in range(3):
    if what == "fwd":
        c.fwd_call(vd, st, ls, ld, ad, c._lib.FLAG_TC)
    else:
        c.bwd_call(vd, st, ls, ld, ad, gd, c._lib.FLAG_TC)
    torch.cuda.synchronize()
print("done")
