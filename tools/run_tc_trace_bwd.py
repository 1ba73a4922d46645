"""One traced launch of the tensor-core grad_value kernel at the bench shape (MSDA_TC_TRACE=1, -DMSDA_TC_TRACE build)."""
import os
import sys
os.environ["MSDA_TC_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import run_tc_check as c

dev = torch.device("cuda:0")
value, loc, attn, gout, lsi = c.make_case(c.COCO, 8, 8, 4, "grid", 0)
st = torch.as_tensor(c.COCO, dtype=torch.long, device=dev)
ls = torch.as_tensor(lsi, dtype=torch.long, device=dev)
vd, ld, ad = value.to(torch.bfloat16).to(dev), loc.to(dev), attn.to(dev)
gd = gout.to(torch.bfloat16).to(dev)
for _ in range(2):
    c.bwd_call(vd, st, ls, ld, ad, gd, c._lib.FLAG_TC)
    torch.cuda.synchronize()
