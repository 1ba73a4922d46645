"""torch.profiler breakdown of the Encoder-Cross-Fusion training step and the bf16 encoder."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from dfvod_b200 import data_parallel
from dfvod_b200.deformable_transformer import DeformableTransformer

dev = torch.device("cuda:0")
bf = torch.bfloat16
torch.manual_seed(4)
model = DeformableTransformer(num_feature_levels=4, return_intermediate_dec=True, use_depth=True, dropout=0.0,
                              depth_type="DepthDeform_encoder_cf_dformer").to(dev).bfloat16()
opt = torch.optim.AdamW(model.parameters(), lr=1e-5)
n = 4
srcs, masks, poss = bench._pyramid(torch, dev, bench.COCO_SHAPES, n, bf, 40)
dsrcs, dmasks, dposs = bench._pyramid(torch, dev, bench.COCO_SHAPES, n, bf, 140)
query = torch.randn(300, 512, device=dev, dtype=bf)

def step():
    opt.zero_grad(set_to_none=True)
    hs = model(srcs, masks, poss, dsrcs, dmasks, dposs, query)[0]
    hs.float().square().mean().backward()
    opt.step()

for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
