"""Diagnose per-step outliers seen in bench.py: host time per call, allocator activity."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from dfvod_b200 import MultiScaleDeformableAttention as MSDA
dt = torch.bfloat16 if "bf16" in sys.argv else torch.float32
dev = torch.device("cuda:0")
lsi, s = bench.level_start(bench.COCO_SHAPES)
v, l, a, g = bench.make_inputs(torch, 8, 1000, "grid")
v, g = v.to(dt), g.to(dt)
v, l, a, g = (t.to(dev) for t in (v, l, a, g))
st = torch.as_tensor(bench.COCO_SHAPES, dtype=torch.long, device=dev)
ls = torch.as_tensor(lsi, dtype=torch.long, device=dev)
use_sampler = "sampler" in sys.argv
import contextlib
ctx = bench.ClockSampler(0) if use_sampler else contextlib.nullcontext()
for _ in range(5):
    MSDA.ms_deform_attn_forward(v, st, ls, l, a, 64); MSDA.ms_deform_attn_backward(v, st, ls, l, a, g, 64)
torch.cuda.synchronize()
K = 300
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
host = []
stats0 = torch.cuda.memory_stats()
with ctx:
    for k in range(K):
        t0 = time.perf_counter()
        ev[k][0].record()
        out = MSDA.ms_deform_attn_forward(v, st, ls, l, a, 64)
        t1 = time.perf_counter()
        ev[k][1].record()
        grads = MSDA.ms_deform_attn_backward(v, st, ls, l, a, g, 64)
        t2 = time.perf_counter()
        ev[k][2].record()
        host.append((t1 - t0, t2 - t1, torch.cuda.memory_stats()["num_device_alloc"]))
    torch.cuda.synchronize()
stats1 = torch.cuda.memory_stats()
print("device allocs during loop:", stats1["num_device_alloc"] - stats0["num_device_alloc"],
      "frees:", stats1["num_device_free"] - stats0["num_device_free"], "retries:", stats1["num_alloc_retries"])
f = [e[0].elapsed_time(e[1]) for e in ev]; b = [e[1].elapsed_time(e[2]) for e in ev]
import statistics
print("fwd med %.4f max %.4f | bwd med %.4f max %.4f" % (statistics.median(f), max(f), statistics.median(b), max(b)))
for k in range(K):
    if f[k] > 1.5 * statistics.median(f) or b[k] > 1.5 * statistics.median(b):
        print(f"step {k}: gpu fwd {f[k]:.3f} bwd {b[k]:.3f} ms | host fwd-call {host[k][0]*1e3:.3f} bwd-call {host[k][1]*1e3:.3f} ms | allocs {host[k][2]}")
