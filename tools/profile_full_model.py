"""torch.profiler breakdown of the bf16 6+6 DeformableTransformer (inference, batch 8, COCO pyramid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from dfvod_b200.deformable_transformer import DeformableTransformer

dev = torch.device("cuda:0")
bf = torch.bfloat16
torch.manual_seed(1)
model = DeformableTransformer(num_feature_levels=4, return_intermediate_dec=True).to(dev).eval().bfloat16()
srcs, masks, poss = bench._pyramid(torch, dev, bench.COCO_SHAPES, 8, bf, 2)
query = torch.randn(300, 512, device=dev, dtype=bf)

def run():
    with torch.no_grad():
        return model(srcs, masks, poss, None, None, None, query)

for _ in range(3):
    run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        run()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=90))
