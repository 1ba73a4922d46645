"""One bf16 encoder layer at the bench geometry, HEAD_MAJOR_INFERENCE from the environment (for an ncu launch list)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from dfvod_b200 import transformer_layers as tl
from dfvod_b200.ops.modules import ms_deform_attn as msda_module

msda_module.HEAD_MAJOR_INFERENCE = os.environ.get("HM", "1") == "1"
dev = torch.device("cuda:0")
lsi, s = bench.level_start(bench.COCO_SHAPES)
st = torch.as_tensor(bench.COCO_SHAPES, dtype=torch.long, device=dev)
ls = torch.as_tensor(lsi, dtype=torch.long, device=dev)
bf = torch.bfloat16
n = 8
torch.manual_seed(1)
layer = tl.DeformableTransformerEncoderLayer(256, 1024, 0.0, "relu", 4, 8, 4).to(dev).to(bf).eval()
enc = tl.DeformableTransformerEncoder(layer, 1).to(dev).to(bf).eval()
src = torch.randn(n, s, 256, device=dev, dtype=bf)
pos = torch.randn(n, s, 256, device=dev, dtype=bf)
vr = torch.ones(n, 4, 2, device=dev)
with torch.no_grad():
    for _ in range(3):
        enc(src, st, ls, vr, pos, None)
torch.cuda.synchronize()
