"""Head-major value layout [N, M, S, 32] (csrc/msda_forward_hm.cu) against the default kernel and the fp64 oracle, + timing."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import run_tc_check as c
from oracle import msda_oracle

dev = torch.device("cuda:0")
ok = True
for name, shapes, n, m, p, dist, seed, lq in c.CASES:
    value, loc, attn, gout, lsi = c.make_case(shapes, n, m, p, dist, seed, lq)
    st = torch.as_tensor(shapes, dtype=torch.long, device=dev)
    ls = torch.as_tensor(lsi, dtype=torch.long, device=dev)
    vb = value.to(torch.bfloat16)
    ref = msda_oracle.core_pytorch(vb.double(), shapes, loc.double(), attn.double())
    vd, ld, ad = vb.to(dev), loc.to(dev), attn.to(dev)
    vhm = vd.permute(0, 2, 1, 3).contiguous()
    o_hm = c.fwd_call(vhm.view(vd.shape), st, ls, ld, ad, c._lib.FLAG_VALUE_HEAD_MAJOR)
    o_lg = c.fwd_call(vd, st, ls, ld, ad, 0)
    torch.cuda.synchronize()
    e_hm, e_lg = c.nerr(o_hm, ref), c.nerr(o_lg, ref)
    good = e_hm[0] <= 2.0 ** -7 and e_hm[1] <= 4e-3
    ok &= good
    print(f"{name:28s} head-major max {e_hm[0]:.2e} l2 {e_hm[1]:.2e} | default max {e_lg[0]:.2e} l2 {e_lg[1]:.2e} {'OK' if good else 'FAIL'}", flush=True)
for dist in ("grid", "init", "random"):
    value, loc, attn, gout, lsi = c.make_case(c.COCO, 8, 8, 4, dist, 0)
    st = torch.as_tensor(c.COCO, dtype=torch.long, device=dev)
    ls = torch.as_tensor(lsi, dtype=torch.long, device=dev)
    vd, ld, ad = value.to(torch.bfloat16).to(dev), loc.to(dev), attn.to(dev)
    vhm = vd.permute(0, 2, 1, 3).contiguous().view(vd.shape)
    t_hm = c.time_ms(lambda: c.fwd_call(vhm, st, ls, ld, ad, c._lib.FLAG_VALUE_HEAD_MAJOR))
    t_lg = c.time_ms(lambda: c.fwd_call(vd, st, ls, ld, ad, 0))
    t_tr = c.time_ms(lambda: vd.permute(0, 2, 1, 3).contiguous())
    print(f"time fwd bf16 batch 8 {dist:6s}: head-major {t_hm:.3f} ms   default {t_lg:.3f} ms   (torch transpose pass {t_tr:.3f} ms)", flush=True)
print("ALL OK" if ok else "SOME FAILED")
