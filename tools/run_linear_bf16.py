"""bf16 linear layers at the encoder's shapes: the library GEMM (+ the row-zeroing kernel for value_proj) against the
TMA / tcgen05 kernel of csrc/linear_bf16.cu."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import bench
from dfvod_b200.ops.functions import layer_epilogue_func as L

dev = torch.device("cuda:0")
rows = 8 * 22223


def t(fn, iters=20):
    return bench._time_events(torch, fn, iters, 3)


for n, k, relu, masked in ((256, 256, False, False), (256, 256, False, True), (384, 256, False, False), (1024, 256, True, False),
                           (256, 1024, False, False), (1536, 256, False, True)):
    torch.manual_seed(0)
    x = torch.randn(rows, k, device=dev).bfloat16()
    w = (torch.randn(n, k, device=dev) / 16).bfloat16()
    b = torch.randn(n, device=dev).bfloat16()
    mask = (torch.rand(rows, device=dev) < 0.1) if masked else None

    def lib():
        y = torch._addmm_activation(b, x, w.t(), use_gelu=False) if relu else torch.addmm(b, x, w.t())
        if mask is not None:
            y = L.zero_masked_rows_(y, mask, exclusive=True)
        return y

    t_lib = t(lib)
    t_own = t(lambda: L.linear_bf16(x, w, b, relu=relu, zero_rows=mask))
    gb = (rows * k * 2 + rows * n * 2) / 1e9
    print(f"[{rows} x {k}] -> {n}{' + ReLU' if relu else ''}{' + mask' if masked else ''}: library {t_lib * 1e3:.1f} us | kernel "
          f"{t_own * 1e3:.1f} us ({gb / t_own * 1e3:.0f} GB/s of x + y)", flush=True)
