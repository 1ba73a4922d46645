"""fp32 errors of every golden module / layer case on the GPU (normalised max, relative L2) for outputs, input
gradients and parameter gradients -- the numbers the tolerances in tests/test_gpu_modules.py are set from."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import module_cases  # noqa: E402
from tests.util import load_golden, nerr  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
worst = {"out": 0.0, "grad_in": 0.0, "grad_param": 0.0}
for name in sorted(module_cases.CASES):
    gold = load_golden(name)
    out, gin, gpar = module_cases.run_case(name, gold, "cuda", torch.float32)
    e_out = nerr(out, gold["out"])
    e_in = max((nerr(g, gold["grad_in." + k]) for k, g in gin.items()
                if g is not None and gold["grad_in." + k].shape != ()), default=(0.0, 0.0))
    e_par = max((nerr(g, gold["grad_param." + k]) for k, g in gpar.items()
                 if g is not None and gold.get("grad_param." + k) is not None and gold["grad_param." + k].shape != ()),
                default=(0.0, 0.0))
    worst["out"] = max(worst["out"], *e_out)
    worst["grad_in"] = max(worst["grad_in"], *e_in)
    worst["grad_param"] = max(worst["grad_param"], *e_par)
    print(f"{name:32s} out {e_out[0]:.2e} {e_out[1]:.2e} | grad_in {e_in[0]:.2e} {e_in[1]:.2e} | grad_param {e_par[0]:.2e} {e_par[1]:.2e}")
print("worst", worst)
