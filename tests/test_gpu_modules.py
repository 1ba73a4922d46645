"""GPU: MSDeformAttn and every layer class that owns one, running the real sm_100a kernels,
against outputs and gradients of the REAL reference classes (tests/golden/, made by
oracle/gen_golden.py).  fp64 runs the shape-generic kernels (tolerance 1e-7: the reference
builds its reference points in fp32); fp32 runs the vectorised kernels (head width 8)."""
import numpy as np
import pytest
import torch

from tests import module_cases
from tests.util import load_golden, nerr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(module_cases.CASES))
def test_fp64_matches_reference_class(name):
    gold = load_golden(name)
    out, gin, gpar = module_cases.run_case(name, gold, "cuda", torch.float64)
    np.testing.assert_allclose(out, gold["out"], rtol=1e-7, atol=1e-9)
    for k, g in gin.items():
        if gold["grad_in." + k].shape == ():
            assert g is None or not np.any(g)
            continue
        np.testing.assert_allclose(g, gold["grad_in." + k], rtol=1e-6, atol=1e-8, err_msg=k)
    for k, g in gpar.items():
        ref = gold["grad_param." + k]
        if ref.shape == ():
            assert g is None or not np.any(g)
            continue
        np.testing.assert_allclose(g, ref, rtol=1e-6, atol=1e-8, err_msg=k)


@pytest.mark.parametrize("name", sorted(module_cases.CASES))
def test_fp32_matches_reference_class(name):
    gold = load_golden(name)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        out, gin, gpar = module_cases.run_case(name, gold, "cuda", torch.float32)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    emax, el2 = nerr(out, gold["out"])
    assert emax <= 2e-5 and el2 <= 2e-5, f"out: {emax:.2e} {el2:.2e}"
    for k, g in gin.items():
        if gold["grad_in." + k].shape == () or g is None:
            continue
        emax, el2 = nerr(g, gold["grad_in." + k])
        assert emax <= 1e-4 and el2 <= 1e-4, f"{k}: {emax:.2e} {el2:.2e}"
    for k, g in gpar.items():
        ref = gold["grad_param." + k]
        if ref.shape == () or g is None:
            continue
        emax, el2 = nerr(g, ref)
        assert emax <= 1e-4 and el2 <= 1e-4, f"{k}: {emax:.2e} {el2:.2e}"
