"""One fp32-grade linear layer [8 * 22223, K] -> N through the tcgen05 kernel (for ncu)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dfvod_b200.ops.functions import layer_epilogue_func as L

n, k = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (256, 256)
dev = torch.device("cuda:0")
x = torch.randn(8 * 22223, k, device=dev)
w = torch.randn(n, k, device=dev) / 16
b = torch.randn(n, device=dev)
for _ in range(3):
    L.linear_tf32x3(x, w, b, route="kernel")
torch.cuda.synchronize()
