"""GPU: time this repo's kernels next to the reference's own CUDA kernels recompiled for
sm_100a (oracle/_ref) on BASELINE.json config 1 shapes, and record the numbers under
gpurun_out/perf_vs_ref_cuda.json.  The only assertion is the sanity bar "not slower than the
recompiled reference"; the numbers themselves are reported, not gated."""
import json
import os

import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu

from dfvod_b200 import MultiScaleDeformableAttention as MSDA


def _time(fn, iters=10, warm=3):
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


@pytest.mark.parametrize("dist", ["grid", "random"])
def test_time_against_recompiled_reference(dist):
    if util.ref_cuda_lib() is None:
        pytest.skip("oracle/_ref/libmsda_ref_cuda.so not built")
    shapes = util.COCO_SHAPES
    s = sum(h * w for h, w in shapes)
    n = 8
    value, loc, attn, gout = util.make_inputs(shapes, n, 8, 32, s, 4, seed=0, dist=dist)
    st, ls = util.shapes_tensors(shapes, "cuda")
    v, l, a, g = (t.cuda() for t in (value, loc, attn, gout))
    res = {
        "ours_fwd_ms": _time(lambda: MSDA.ms_deform_attn_forward(v, st, ls, l, a, 64)),
        "ours_bwd_ms": _time(lambda: MSDA.ms_deform_attn_backward(v, st, ls, l, a, g, 64)),
        "ref_cuda_fwd_ms": _time(lambda: util.ref_cuda_forward(v, st, ls, l, a)),
        "ref_cuda_bwd_ms": _time(lambda: util.ref_cuda_backward(v, st, ls, l, a, g)),
        "config": f"N={n} frames, COCO pyramid S=Lq={s}, M=8 D=32 L=4 P=4, fp32, loc={dist}; "
                  "reference timings include its zero-fills (at::zeros in the reference wrapper)",
    }
    res["speedup_fwd"] = res["ref_cuda_fwd_ms"] / res["ours_fwd_ms"]
    res["speedup_bwd"] = res["ref_cuda_bwd_ms"] / res["ours_bwd_ms"]
    out_dir = os.path.join(util.ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, "perf_vs_ref_cuda.json")
    blob = {}
    if os.path.exists(path):
        try:
            blob = json.load(open(path))
        except Exception:
            blob = {}
    blob[dist] = res
    json.dump(blob, open(path, "w"), indent=1)
    print(json.dumps(res))
    assert res["ours_fwd_ms"] <= res["ref_cuda_fwd_ms"] * 1.05
    assert res["ours_bwd_ms"] <= res["ref_cuda_bwd_ms"] * 1.05
