"""The tcgen05 output-projection + residual + LayerNorm kernel against the composition it replaces
(library GEMM + fused add-LayerNorm kernel), 8 x 22223 rows, bf16."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from dfvod_b200.ops.functions import add_layer_norm, linear, proj_layer_norm

dev = torch.device("cuda:0")
rows, c = 8 * 22223, 256
lin, norm = torch.nn.Linear(c, c).to(dev).bfloat16(), torch.nn.LayerNorm(c).to(dev).bfloat16()
x = torch.randn(rows, c, device=dev).bfloat16()
res = torch.randn(rows, c, device=dev).bfloat16()
pos = torch.randn(rows, c, device=dev).bfloat16()
with torch.no_grad():
    for name, fn in (("tcgen05 proj+res+LN", lambda: proj_layer_norm(lin, norm, x, res)),
                     ("tcgen05 proj+res+LN+pos", lambda: proj_layer_norm(lin, norm, x, res, pos)),
                     ("GEMM + add_layer_norm", lambda: add_layer_norm(norm, linear(lin, x), res)),
                     ("GEMM + add_layer_norm + pos", lambda: add_layer_norm(norm, linear(lin, x), res, None, pos)),
                     ("GEMM alone", lambda: linear(lin, x))):
        ms = bench._time_events(torch, fn, 30, 5)
        print(f"{name:30s} {ms * 1e3:8.1f} us")
