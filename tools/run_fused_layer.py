"""Launch the FUSED layer kernels at the encoder shape (for ncu): MSDeformAttn(256, 4, 8, 4) in bf16 on the 800x1333
pyramid -- inference forward at batch 8 (msda_fwd_fast_kernel<bf16, 32, true, bf16>, half of the 6+6 transformer's GPU
time) and training forward + backward at batch 4 (msda_bwd_fast_kernel<bf16, 32, true, bf16>, 37 % of the training step).
The offset projection is perturbed so that sampling offsets scatter ~N(0, 1) px around the compass pattern, as in
bench.py's "grid" distribution."""
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


def inputs(n, grad):
    q = torch.randn(n, s, 256, device=dev, dtype=bf, requires_grad=grad)
    x = torch.randn(n, s, 256, device=dev, dtype=bf, requires_grad=grad)
    ref = encoder_reference_points(shapes, torch.ones(n, len(shapes), 2, device=dev), dev)
    return q, ref, x


q8, ref8, x8 = inputs(8, False)
q4, ref4, x4 = inputs(4, True)
for _ in range(3):
    with torch.no_grad():
        attn(q8, ref8, x8, st, ls, None)
    out = attn(q4, ref4, x4, st, ls, None)
    out.backward(torch.ones_like(out))
torch.cuda.synchronize()
print("ok")
