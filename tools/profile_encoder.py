"""torch.profiler breakdown of the bf16 6-layer encoder (inference, batch 8, COCO pyramid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from dfvod_b200 import transformer_layers as tl

dev = torch.device("cuda:0")
bf = torch.bfloat16
n = int(os.environ.get("BATCH", "8"))
lsi, s = bench.level_start(bench.COCO_SHAPES)
st = torch.as_tensor(bench.COCO_SHAPES, dtype=torch.long, device=dev)
ls = torch.as_tensor(lsi, dtype=torch.long, device=dev)
torch.manual_seed(1)
enc = tl.DeformableTransformerEncoder(tl.DeformableTransformerEncoderLayer(256, 1024, 0.1, "relu", 4, 8, 4), 6)
enc = enc.to(dev).eval().bfloat16()
src = torch.randn(n, s, 256, device=dev, dtype=bf)
pos = torch.randn(n, s, 256, device=dev, dtype=bf)
vr = torch.ones(n, 4, 2, device=dev)

def run():
    with torch.no_grad():
        return enc(src, st, ls, vr, pos, None)

for _ in range(3):
    run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        run()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=90))
