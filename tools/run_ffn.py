"""Runs the tcgen05 feed-forward kernel a few times at the bench shape (ncu target)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dfvod_b200.ops.functions import ffn_layer_norm
dev = torch.device("cuda:0"); bf = torch.bfloat16
rows, c, f = 8 * 22223, 256, 1024
lin1, lin2 = torch.nn.Linear(c, f).to(dev).to(bf), torch.nn.Linear(f, c).to(dev).to(bf)
norm = torch.nn.LayerNorm(c).to(dev).to(bf)
x, p = torch.randn(rows, c, device=dev).to(bf), torch.randn(rows, c, device=dev).to(bf)
with torch.no_grad():
    for _ in range(3):
        ffn_layer_norm(lin1, lin2, norm, x, p)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ffn_layer_norm(lin1, lin2, norm, x, p)
    e1.record(); torch.cuda.synchronize()
    print("ffn_layer_norm", e0.elapsed_time(e1) / 10, "ms")
