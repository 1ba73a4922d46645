"""fp32 linear layers at the encoder's shapes: IEEE SGEMM, one TF32 GEMM, the split pass + library GEMM (round 1), and
the one-kernel tcgen05 route (csrc/linear_tf32x3.cu); then the 6-layer fp32 encoder in each mode."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import bench
from dfvod_b200.ops.functions import layer_epilogue_func as L

dev = torch.device("cuda:0")
rows = 8 * 22223


def t(fn, iters=10):
    return bench._time_events(torch, fn, iters, 3)


def split_pass(x, w, b):
    k = x.shape[-1]
    a = L._tf32_split(x)
    ws = L._tf32_split(w)
    ws = torch.cat([ws[:, k:2 * k], ws[:, :k], ws[:, 2 * k:]], 1)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        return torch.addmm(b, a, ws.t())
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


for n, k in ((256, 256), (384, 256), (1024, 256), (256, 1024)):
    torch.manual_seed(0)
    x = torch.randn(rows, k, device=dev)
    w = torch.randn(n, k, device=dev) / 16
    b = torch.randn(n, device=dev)
    ref = F.linear(x[:4096].double(), w.double(), b.double())
    torch.backends.cuda.matmul.allow_tf32 = False
    t_ieee = t(lambda: F.linear(x, w, b))
    torch.backends.cuda.matmul.allow_tf32 = True
    t_tf32 = t(lambda: F.linear(x, w, b))
    torch.backends.cuda.matmul.allow_tf32 = False
    t_pass = t(lambda: split_pass(x, w, b))
    t_own = t(lambda: L.linear_tf32x3(x, w, b, route="kernel"))
    err = float((L.linear_tf32x3(x[:4096], w, b, route="kernel").double() - ref).abs().max() / ref.abs().max())
    flops = 2.0 * rows * n * k * 3
    print(f"[{rows} x {k}] -> {n}: IEEE {t_ieee:.3f} ms | TF32 {t_tf32:.3f} | split pass + library {t_pass:.3f} | one kernel "
          f"{t_own:.3f} ms ({flops / t_own / 1e9:.0f} TFLOP/s of TF32 work), max err / max ref {err:.2e}", flush=True)
