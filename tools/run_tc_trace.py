"""One traced launch of the tensor-core forward at the bench shape (MSDA_TC_TRACE=1 prints per-role cycle sums)."""
import os
import sys
os.environ["MSDA_TC_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import run_tc_check as c

dev = torch.device("cuda:0")
dist = sys.argv[1] if len(sys.argv) > 1 else "grid"
value, loc, attn, gout, lsi = c.make_case(c.COCO, 8, 8, 4, dist, 0)
st = torch.as_tensor(c.COCO, dtype=torch.long, device=dev)
ls = torch.as_tensor(lsi, dtype=torch.long, device=dev)
vd, ld, ad = value.to(torch.bfloat16).to(dev), loc.to(dev), attn.to(dev)
for _ in range(2):
    c.fwd_call(vd, st, ls, ld, ad, c._lib.FLAG_TC)
    torch.cuda.synchronize()
