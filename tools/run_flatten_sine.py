"""Launch the layout kernels in front of the encoder at the sizes bench.py quotes (for ncu): flatten_level (bf16 / fp32,
with the level embedding), sine_coordinates and sine_position_tokens (bf16) on the 800x1333 pyramid, batch 8."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dfvod_b200.ops.functions import flatten_levels
from dfvod_b200.position_encoding import PositionEmbeddingSine

dev = torch.device("cuda:0")
bf = torch.bfloat16
shapes = ((100, 167), (50, 84), (25, 42), (13, 21))
embed = torch.randn(4, 256, device=dev)
maps = {dt: [torch.randn(8, 256, h, w, device=dev).to(dt) for h, w in shapes] for dt in (bf, torch.float32)}
masks = [torch.zeros(8, h, w, dtype=torch.bool, device=dev) for h, w in shapes]
sine = PositionEmbeddingSine(128, normalize=True)
for _ in range(3):
    for dt, levels in maps.items():
        flatten_levels(levels, [embed[i].to(dt) for i in range(4)])
    sine.forward_tokens(masks, embed, dtype=bf)
torch.cuda.synchronize()
print("ok")
