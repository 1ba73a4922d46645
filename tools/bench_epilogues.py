"""Developer benchmark: the layer-epilogue kernels against the HBM roofline and against the
PyTorch composition they replace (batch 8 x 22223 tokens x 256 channels).  Prints JSON."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from dfvod_b200.ops.functions import add_layer_norm, linear_relu, zero_masked_rows_

dev = torch.device("cuda:0")
PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))).get("hbm_gbs", 6552.3) \
    if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6552.3
rows, c = 8 * 22223, 256


def timeit(fn, iters=20, warm=3):
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


res = {"rows": rows, "channels": c}
for name, dt in (("bf16", torch.bfloat16), ("f32", torch.float32)):
    e = 2 if dt == torch.bfloat16 else 4
    norm = torch.nn.LayerNorm(c).to(dev).to(dt)
    b, r, p = (torch.randn(rows, c, device=dev).to(dt) for _ in range(3))
    # flush: operands (3 x 91 MB bf16) exceed L2 together with the outputs
    with torch.no_grad():
        ms = timeit(lambda: add_layer_norm(norm, b, r))
        res[f"{name}_add_ln_fwd_ms"] = ms
        res[f"{name}_add_ln_fwd_gbps"] = 3 * rows * c * e / ms / 1e6
        ms = timeit(lambda: add_layer_norm(norm, b, r, None, p))
        res[f"{name}_add_ln_pos_fwd_ms"] = ms
        res[f"{name}_add_ln_pos_fwd_gbps"] = 5 * rows * c * e / ms / 1e6
        ms = timeit(lambda: add_layer_norm(norm, b, r, "gelu"))
        res[f"{name}_add_gelu_ln_fwd_ms"] = ms
        res[f"{name}_torch_add_ln_fwd_ms"] = timeit(lambda: norm(r + b))
        res[f"{name}_torch_add_ln_pos_fwd_ms"] = timeit(lambda: norm(r + b) + p)
    bg, rg = b.clone().requires_grad_(True), r.clone().requires_grad_(True)
    go = torch.randn_like(b)

    def fb(fused):
        y = add_layer_norm(norm, bg, rg) if fused else norm(rg + bg)
        y.backward(go)
        bg.grad = rg.grad = None
        norm.zero_grad(set_to_none=True)
    res[f"{name}_add_ln_fwdbwd_ms"] = timeit(lambda: fb(True))
    res[f"{name}_torch_add_ln_fwdbwd_ms"] = timeit(lambda: fb(False))
    # backward alone: reads dy, branch, residual, writes d (4 x rows x C) + statistics
    y = add_layer_norm(norm, bg, rg)
    ms = timeit(lambda: torch.autograd.grad(y, (bg, rg), go, retain_graph=True))
    res[f"{name}_add_ln_bwd_ms"] = ms
    res[f"{name}_add_ln_bwd_gbps"] = 4 * rows * c * e / ms / 1e6
    del y
    mask = torch.rand(8, 22223, device=dev) < 0.05
    v = torch.randn(8, 22223, c, device=dev).to(dt)
    with torch.no_grad():
        res[f"{name}_zero_masked_rows_ms"] = timeit(lambda: zero_masked_rows_(v, mask))
        res[f"{name}_torch_masked_fill_ms"] = timeit(lambda: v.masked_fill(mask[..., None], 0.0))
    lin = torch.nn.Linear(c, 1024).to(dev).to(dt)
    with torch.no_grad():
        res[f"{name}_linear_relu_ms"] = timeit(lambda: linear_relu(lin, b))
        res[f"{name}_torch_linear_relu_ms"] = timeit(lambda: F.relu(lin(b)))
# the whole feed-forward block: tcgen05 kernel vs the three-launch composition it replaces
from dfvod_b200.ops.functions import ffn_layer_norm, linear
bf = torch.bfloat16
lin1, lin2 = torch.nn.Linear(c, 1024).to(dev).to(bf), torch.nn.Linear(1024, c).to(dev).to(bf)
norm = torch.nn.LayerNorm(c).to(dev).to(bf)
xb, pb = torch.randn(rows, c, device=dev).to(bf), torch.randn(rows, c, device=dev).to(bf)
with torch.no_grad():
    ms = timeit(lambda: ffn_layer_norm(lin1, lin2, norm, xb, pb))
    res["bf16_ffn_ln_tcgen05_ms"] = ms
    res["bf16_ffn_ln_tcgen05_tflops"] = 4.0 * rows * c * 1024 / ms / 1e9
    res["bf16_ffn_ln_composition_ms"] = timeit(
        lambda: add_layer_norm(norm, linear(lin2, linear_relu(lin1, xb)), xb, None, pb))
res["hbm_peak_gbps"] = PEAK
print(json.dumps(res, indent=1))
