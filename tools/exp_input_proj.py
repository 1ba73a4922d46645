"""Which library call computes the 1x1 input projection straight into token-major memory fastest?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

dev = torch.device("cuda:0")
bf = torch.bfloat16
n = 8
for cin, h, w in [(512, 100, 167), (1024, 50, 84), (2048, 25, 42)]:
    x = torch.randn(n, cin, h, w, device=dev).to(bf)
    wgt = torch.randn(256, cin, device=dev).to(bf) * 0.02
    bias = torch.randn(256, device=dev).to(bf)
    conv = torch.nn.Conv2d(cin, 256, 1).to(dev).to(bf)
    out = torch.empty(n, h * w, 256, device=dev, dtype=bf)
    big = torch.empty(n, 22223, 256, device=dev, dtype=bf)
    xt = x.flatten(2).transpose(1, 2)
    wt = wgt.t()
    cands = {
        "matmul(view^T, W^T)+bias": lambda: torch.matmul(xt, wt) + bias,
        "baddbmm(bias, view^T, W^T expand)": lambda: torch.baddbmm(bias, xt, wt.expand(n, cin, 256)),
        "bmm out=": lambda: torch.bmm(xt, wt.expand(n, cin, 256), out=out),
        "bmm out=slice of big": lambda: torch.bmm(xt, wt.expand(n, cin, 256), out=big[:, :h * w]),
        "conv2d NCHW (reference)": lambda: conv(x),
        "conv2d NCHW + flatten/transpose copy": lambda: conv(x).flatten(2).transpose(1, 2).contiguous(),
        "W @ X (NCHW out, bmm)": lambda: torch.bmm(wgt.expand(n, 256, cin), x.flatten(2)),
    }
    with torch.no_grad():
        for name, fn in cands.items():
            try:
                ms = bench._time_events(torch, fn, 20, 3)
                print(f"{cin:5d}x{h}x{w}  {name:40s} {ms*1e3:8.1f} us")
            except Exception as exc:
                print(f"{cin:5d}x{h}x{w}  {name:40s} failed: {repr(exc)[:100]}")
