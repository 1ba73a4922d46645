"""bf16 encoder-layer / encoder inference with and without the head-major value path (CUDA-graph replay), and the
value-projection kernels alone."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from dfvod_b200 import transformer_layers as tl
from dfvod_b200.data_parallel import GraphedInference
from dfvod_b200.ops.functions import value_proj_head_major
from dfvod_b200.ops.modules import ms_deform_attn as msda_module

dev = torch.device("cuda:0")
lsi, s = bench.level_start(bench.COCO_SHAPES)
st = torch.as_tensor(bench.COCO_SHAPES, dtype=torch.long, device=dev)
ls = torch.as_tensor(lsi, dtype=torch.long, device=dev)
bf = torch.bfloat16
n = 8
torch.manual_seed(1)
layer = tl.DeformableTransformerEncoderLayer(256, 1024, 0.0, "relu", 4, 8, 4).to(dev).to(bf).eval()
enc = tl.DeformableTransformerEncoder(layer, 6).to(dev).to(bf).eval()
src = torch.randn(n, s, 256, device=dev, dtype=bf)
pos = torch.randn(n, s, 256, device=dev, dtype=bf)
vr = torch.ones(n, 4, 2, device=dev)


def t(fn, iters=20):
    return bench._time_events(torch, fn, iters, 3)


with torch.no_grad():
    lin = layer.self_attn.value_proj
    x2 = src.reshape(n * s, 256)
    t_lib = t(lambda: torch.addmm(lin.bias, x2, lin.weight.t()))
    t_own = t(lambda: value_proj_head_major(lin, src, None, 8))
    print(f"value_proj alone: library GEMM {t_lib * 1e3:.1f} us   tcgen05 head-major epilogue {t_own * 1e3:.1f} us", flush=True)
    for flag in (False, True, False, True):
        msda_module.HEAD_MAJOR_INFERENCE = flag
        run = GraphedInference(lambda: enc(src, st, ls, vr, pos, None))
        ms = t(run, 10)
        print(f"6-layer encoder bf16 batch 8, head-major={flag}: {ms:.3f} ms = {n / ms * 1e3:.0f} frames/s", flush=True)
