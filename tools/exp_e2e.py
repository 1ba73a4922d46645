"""PCIe ceilings on this box and the host-buffer pipeline against them."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from dfvod_b200.host_pipeline import HostPipelinedMSDA

dev = torch.device("cuda:0")
mb = 256
h = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
h2 = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
d = torch.empty(mb << 20, dtype=torch.uint8, device=dev)
d2 = torch.empty(mb << 20, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def wall(fn, iters=10):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / iters

t = wall(lambda: d.copy_(h, non_blocking=True)); print(f"H2D alone   {mb / 1024 / t:6.1f} GiB/s")
t = wall(lambda: h2.copy_(d2, non_blocking=True)); print(f"D2H alone   {mb / 1024 / t:6.1f} GiB/s")
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
t = wall(both); print(f"both ways   {mb / 1024 / t:6.1f} GiB/s each")

n = 8
lsi, s = bench.level_start(bench.COCO_SHAPES)
st = torch.as_tensor(bench.COCO_SHAPES, dtype=torch.long)
ls = torch.as_tensor(lsi, dtype=torch.long)
for dtype in (torch.float32, torch.bfloat16):
    value_h, loc_h, attn_h, gout_h = bench.make_inputs(torch, n, 1000, "grid")
    value_h, gout_h = value_h.to(dtype), gout_h.to(dtype)
    host = [x.pin_memory() for x in (value_h, loc_h, attn_h, gout_h)]
    outs = [torch.empty((n, s, 256), dtype=dtype).pin_memory(), torch.empty_like(value_h).pin_memory(),
            torch.empty_like(loc_h).pin_memory(), torch.empty_like(attn_h).pin_memory()]
    nbytes = sum(x.numel() * x.element_size() for x in host)
    for chunk, depth in ((1, 3), (1, 4), (2, 3), (2, 4), (4, 3)):
        pipe = HostPipelinedMSDA(dev, st, ls, 8, 32, 4, s, dtype=dtype, chunk_frames=chunk, depth=depth)
        t = wall(lambda: pipe.forward_backward(*host, *outs), 10)
        print(f"{str(dtype)[6:]:9s} chunk {chunk} depth {depth}: {t * 1e3:6.2f} ms/step  {n * s / t / 1e6:6.2f} M q/s  "
              f"{nbytes / t / 2**30:5.1f} GiB/s each way")
        del pipe
