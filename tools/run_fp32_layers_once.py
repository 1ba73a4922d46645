"""Two fp32 encoder layers at the bench geometry (for an ncu launch list of the fp32 inference path)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from dfvod_b200 import transformer_layers as tl

dev = torch.device("cuda:0")
lsi, s = bench.level_start(bench.COCO_SHAPES)
st = torch.as_tensor(bench.COCO_SHAPES, dtype=torch.long, device=dev)
ls = torch.as_tensor(lsi, dtype=torch.long, device=dev)
n = 8
torch.manual_seed(1)
layer = tl.DeformableTransformerEncoderLayer(256, 1024, 0.0, "relu", 4, 8, 4).to(dev).eval()
enc = tl.DeformableTransformerEncoder(layer, 2).to(dev).eval()
src = torch.randn(n, s, 256, device=dev)
pos = torch.randn(n, s, 256, device=dev)
vr = torch.ones(n, 4, 2, device=dev)
with torch.no_grad():
    for _ in range(3):
        enc(src, st, ls, vr, pos, None)
torch.cuda.synchronize()
