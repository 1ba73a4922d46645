"""Developer benchmark: MSDeformAttn module (fused vs unfused) and the 6-layer encoder at the
COCO-scale pyramid.  Prints one JSON object.  (bench.py carries the judged numbers.)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from dfvod_b200 import MSDeformAttn
from dfvod_b200 import transformer_layers as tl

dev = torch.device("cuda:0")
shapes = bench.COCO_SHAPES
lsi, s = bench.level_start(shapes)
st = torch.as_tensor(shapes, dtype=torch.long, device=dev)
ls = torch.as_tensor(lsi, dtype=torch.long, device=dev)
N = int(os.environ.get("BATCH", "8"))


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


res = {"batch": N, "tokens_per_frame": s}
torch.manual_seed(0)
vr = torch.ones(N, len(shapes), 2, device=dev)
ref = tl.encoder_reference_points(shapes, vr, dev)
for dtype_name, dtype in (("f32", torch.float32), ("bf16", torch.bfloat16)):
    mod = MSDeformAttn(256, 4, 8, 4).to(dev)
    with torch.no_grad():
        for prm in mod.parameters():
            prm.add_(torch.randn_like(prm) * 0.02)
    mod = mod.to(dtype)
    q = torch.randn(N, s, 256, device=dev, dtype=dtype, requires_grad=True)
    x = torch.randn(N, s, 256, device=dev, dtype=dtype, requires_grad=True)
    for fused in (True, False):
        mod.fused = fused

        def fwd():
            with torch.no_grad():
                return mod(q, ref, x, st, ls, None)

        def fwdbwd():
            out = mod(q, ref, x, st, ls, None)
            out.backward(torch.ones_like(out))
            mod.zero_grad(set_to_none=True)
            q.grad = None
            x.grad = None

        res[f"module_{dtype_name}_{'fused' if fused else 'unfused'}_fwd_ms"] = timeit(fwd)
        res[f"module_{dtype_name}_{'fused' if fused else 'unfused'}_fwdbwd_ms"] = timeit(fwdbwd)

# 6-layer encoder, inference
enc = tl.DeformableTransformerEncoder(tl.DeformableTransformerEncoderLayer(256, 1024, 0.1, "relu", 4, 8, 4), 6)
enc = enc.to(dev).eval()
src = torch.randn(N, s, 256, device=dev)
pos = torch.randn(N, s, 256, device=dev)
import copy
enc16 = copy.deepcopy(enc).bfloat16()
src16, pos16, vr16 = src.bfloat16(), pos.bfloat16(), vr
for name, ctx, conv in (("f32", torch.autocast("cuda", enabled=False), lambda t: t),
                        ("bf16_autocast", torch.autocast("cuda", dtype=torch.bfloat16), lambda t: t),
                        ("bf16_pure", torch.autocast("cuda", enabled=False), None)):
    def run():
        with torch.no_grad(), ctx:
            return enc(src, shapes, ls, vr, pos, None)
    # note: spatial_shapes passed as a python list -> reference-point builder does not sync
    def run_t():
        with torch.no_grad(), ctx:
            if conv is None:
                return enc16(src16, st, ls, vr16, pos16, None)
            return enc(src, st, ls, vr, pos, None)
    ms = timeit(run_t, iters=5, warm=2)
    res[f"encoder6_{name}_ms"] = ms
    res[f"encoder6_{name}_fps"] = N / ms * 1e3
    # CUDA graph
    try:
        static_out = None
        gph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            run_t()
        torch.cuda.current_stream().wait_stream(side)
        with torch.cuda.graph(gph):
            static_out = run_t()
        ms = timeit(gph.replay, iters=5, warm=2)
        res[f"encoder6_{name}_graph_ms"] = ms
        res[f"encoder6_{name}_graph_fps"] = N / ms * 1e3
    except Exception as e:  # noqa
        res[f"encoder6_{name}_graph_error"] = repr(e)[:200]
print(json.dumps(res, indent=1))
