"""Launch the kernels either side of the attention at the sizes bench.py quotes (for ncu):
roi_align forward / backward, norm_act (64 and 256 channels), group_norm_tokens, flatten_level."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dfvod_b200 import temporal_stage
from dfvod_b200.input_projection import group_norm_tokens
from dfvod_b200.ops.functions import flatten_levels, norm_act

dev = torch.device("cuda:0")
bf = torch.bfloat16
n, h, w = 32, 50, 84
g = torch.Generator().manual_seed(5)
tokens = torch.randn(n, h * w, 256, device=dev, dtype=bf).requires_grad_(True)
cxy = torch.rand(n * 300, 2, generator=g) * torch.tensor([w * 32.0, h * 32.0])
wh = torch.rand(n * 300, 2, generator=g) * torch.tensor([w * 16.0, h * 16.0]) + 8
rois = torch.cat([torch.arange(n).repeat_interleave(300)[:, None].float(), cxy - wh / 2, cxy + wh / 2], -1).to(dev)
ln64, ln256 = torch.nn.LayerNorm(64).to(dev).to(bf), torch.nn.LayerNorm(256).to(dev).to(bf)
gn = torch.nn.GroupNorm(32, 256).to(dev).to(bf)
x64 = torch.randn(9600 * 49, 64, device=dev, dtype=bf)
x256 = torch.randn(9600 * 49, 256, device=dev, dtype=bf)
tok8 = torch.randn(8, 16700, 256, device=dev, dtype=bf)
maps = [torch.randn(8, 256, hh, ww, device=dev, dtype=bf) for hh, ww in ((100, 167), (50, 84), (25, 42), (13, 21))]
for _ in range(3):
    pooled = temporal_stage.roi_align_tokens(tokens, rois, h, w, 7, 1 / 32, 2, True)
    pooled.backward(torch.ones_like(pooled))
    with torch.no_grad():
        norm_act(ln64, x64, "relu", inplace=True)
        norm_act(ln256, x256, "relu", inplace=True)
        group_norm_tokens(tok8, 32, gn.weight, gn.bias, 1e-5, inplace=True, pre_bias=gn.bias)
        flatten_levels(maps)
torch.cuda.synchronize()
print("ok")
