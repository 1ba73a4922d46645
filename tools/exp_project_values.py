import sys; sys.path.insert(0, "/root/repo")
import torch, bench
from dfvod_b200.ops.modules import MSDeformAttn, project_values
from dfvod_b200.ops.functions import linear
dev = torch.device("cuda:0")
for n, s in ((8, 22223), (32, 4200)):
    mods = [MSDeformAttn(256, 4, 8, 4).to(dev).bfloat16() for _ in range(6)]
    x = torch.randn(n, s, 256, device=dev).bfloat16()
    with torch.no_grad():
        a = bench._time_events(torch, lambda: project_values(mods, x), 20, 3)
        b = bench._time_events(torch, lambda: [linear(m.value_proj, x.reshape(n * s, 256)) for m in mods], 20, 3)
    print(n, s, "one GEMM %.1f us" % (a * 1e3), "six GEMMs %.1f us" % (b * 1e3))
