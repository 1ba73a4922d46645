"""CUDA-event timing of the fused MSDeformAttn layer (bf16, encoder shape): inference forward at batch 8 and
training forward + backward at batch 4.  MSDA_B200_LIB selects the library build (A/B of kernel variants)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from dfvod_b200 import MSDeformAttn
from dfvod_b200.transformer_layers import encoder_reference_points

dev = torch.device("cuda:0")
bf = torch.bfloat16
torch.manual_seed(0)
shapes = bench.COCO_SHAPES
lsi, s = bench.level_start(shapes)
st = torch.as_tensor(shapes, dtype=torch.long, device=dev)
ls = torch.as_tensor(lsi, dtype=torch.long, device=dev)
attn = MSDeformAttn(256, 4, 8, 4).to(dev)
with torch.no_grad():
    attn.sampling_offsets.weight.normal_(0, 1 / 16)
    attn.attention_weights.weight.normal_(0, 1 / 16)
attn = attn.to(bf)
ref4 = encoder_reference_points(shapes, torch.ones(4, len(shapes), 2, device=dev), dev)
q = torch.randn(4, s, 256, device=dev, dtype=bf, requires_grad=True)
x = torch.randn(4, s, 256, device=dev, dtype=bf, requires_grad=True)


def step():
    out = attn(q, ref4, x, st, ls, None)
    out.backward(torch.ones_like(out))


for _ in range(5):
    step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(30):
    step()
b.record()
torch.cuda.synchronize()
print(os.environ.get("MSDA_B200_LIB", "in-tree"), "fused layer fwd+bwd (batch 4, bf16): %.4f ms" % (a.elapsed_time(b) / 30))
