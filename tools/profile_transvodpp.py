"""torch.profiler breakdown of the bf16 TransVOD++ clip transformer with its temporal query stage
(8 clips x 4 frames, one 50x84 level, Late Fusion): which kernels the temporal stage adds."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch import nn
from torch.profiler import profile, ProfilerActivity
import bench
from dfvod_b200 import temporal_stage

dev = torch.device("cuda:0")
bf = torch.bfloat16
shapes, clip, clips = [(50, 84)], 4, 8
n = clip * clips
torch.manual_seed(33)
tr = temporal_stage.DeformableTransformer(num_feature_levels=1, return_intermediate_dec=True, use_depth=True,
                                          num_ref_frames=clip - 1, depth_type="DepthDeform_latefusion_dformer")
mlp = lambda: nn.Sequential(nn.Linear(256, 256), nn.ReLU(), nn.Linear(256, 256), nn.ReLU(), nn.Linear(256, 4))
heads = nn.ModuleDict(dict(cls=nn.Linear(256, 31), box=nn.ModuleList(mlp() for _ in range(6)),
                           tcls=nn.ModuleList(nn.Linear(256, 31) for _ in range(3)),
                           tbox=nn.ModuleList(mlp() for _ in range(3))))
tr.decoder.bbox_embed = heads["box"]
tr = tr.to(dev).eval().to(bf)
heads = heads.to(dev).eval().to(bf)
srcs, masks, poss = bench._pyramid(torch, dev, shapes, n, bf, 5)
dsrcs, dmasks, dposs = bench._pyramid(torch, dev, shapes, n, bf, 6)
query = torch.randn(300, 512, device=dev, dtype=bf)
h, w = shapes[0]
whwh = torch.tensor([[w * 32, h * 32, w * 32, h * 32]], dtype=torch.long, device=dev)


def run():
    with torch.no_grad():
        return tr(srcs, masks, poss, dsrcs, dmasks, dposs, whwh, query, heads["cls"], heads["box"][-1],
                  heads["tcls"], heads["tbox"])


for _ in range(3):
    run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        run()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=100))
