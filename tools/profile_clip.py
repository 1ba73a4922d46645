"""Kernel-time breakdown of the TransVOD++ clip transformer with its temporal stage (bench configs[3]) -- torch profiler,
eager, bf16, 8 clips of 4 frames, one 50x84 level."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch import nn
import bench
from dfvod_b200 import temporal_stage

dev = torch.device("cuda:0")
bf = torch.bfloat16
shapes, clip, clips = [(50, 84)], 4, 8
n = clip * clips
torch.manual_seed(33)
tr = temporal_stage.DeformableTransformer(num_feature_levels=1, return_intermediate_dec=True, use_depth=True,
                                          num_ref_frames=clip - 1, depth_type="DepthDeform_latefusion_dformer")


def mlp():
    return nn.Sequential(nn.Linear(256, 256), nn.ReLU(), nn.Linear(256, 256), nn.ReLU(), nn.Linear(256, 4))


heads = nn.ModuleDict(dict(cls=nn.Linear(256, 31), box=nn.ModuleList(mlp() for _ in range(6)),
                           tcls=nn.ModuleList(nn.Linear(256, 31) for _ in range(3)),
                           tbox=nn.ModuleList(mlp() for _ in range(3))))
tr.decoder.bbox_embed = heads["box"]
tr = tr.to(dev).eval().to(bf)
heads = heads.to(dev).eval().to(bf)
with torch.no_grad():
    srcs, masks, poss = bench._pyramid(torch, dev, shapes, n, bf, 5)
    dsrcs, dmasks, dposs = bench._pyramid(torch, dev, shapes, n, bf, 6)
    query = torch.randn(300, 512, device=dev, dtype=bf)
    h, w = shapes[0]
    whwh = torch.tensor([[w * 32, h * 32, w * 32, h * 32]], dtype=torch.long, device=dev)
    run = lambda: tr(srcs, masks, poss, dsrcs, dmasks, dposs, whwh, query, heads["cls"], heads["box"][-1],
                     heads["tcls"], heads["tbox"])
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            run()
        torch.cuda.synchronize()
    rows = [(e.key, e.device_time_total / 3.0, e.count / 3) for e in prof.key_averages() if e.device_time_total > 0]
    rows.sort(key=lambda r: -r[1])
    total = sum(r[1] for r in rows)
    print(f"total kernel time per call {total / 1e3:.3f} ms, {sum(r[2] for r in rows):.0f} launches")
    for k, t, c in rows[:28]:
        print(f"{t:9.1f} us {100 * t / total:5.1f} %  x{c:<5.0f} {k[:110]}")
