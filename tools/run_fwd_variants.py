"""Runs the bf16 drop-in forward once per layout (plain / paired) at the bench workload -- the
target of `ncu --set full` captures comparing the two gathers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from dfvod_b200 import MultiScaleDeformableAttention as MSDA

dev = torch.device("cuda:0")
value, loc, attn, _ = bench.make_inputs(torch, int(os.environ.get("BATCH", "8")), 0, "grid")
lsi, s = bench.level_start(bench.COCO_SHAPES)
st = torch.as_tensor(bench.COCO_SHAPES, dtype=torch.long, device=dev)
ls = torch.as_tensor(lsi, dtype=torch.long, device=dev)
v = value.to(dev, torch.bfloat16)
loc, attn = loc.to(dev), attn.to(dev)
for mode, wb in ((False, False), (True, False), (True, True)):
    MSDA.PAIRED_FORWARD = mode
    MSDA.PAIRED_BF16_WEIGHTS = wb
    for _ in range(3):
        out = MSDA.ms_deform_attn_forward(v, st, ls, loc, attn, 64)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        out = MSDA.ms_deform_attn_forward(v, st, ls, loc, attn, 64)
    e1.record()
    torch.cuda.synchronize()
    print(("paired+bf16w" if wb else "paired") if mode else "plain", e0.elapsed_time(e1) / 10, "ms")
