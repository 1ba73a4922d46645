"""GPU: the opt-in tensor-core (tcgen05 + TMA + TMEM) formulation of the forward gather (csrc/msda_tc_forward.cu,
``MSDA_FLAG_TC``) against the fp64 oracle and against the default lane-group kernel, through the C ABI.

Stated bf16 tolerance: normalised max error <= 2^-7, relative L2 <= 4e-3 against the oracle
evaluated in fp64 on the same bf16 values.  The tensor-core path additionally rounds the per-corner coefficients
(bilinear weight x attention weight) to bf16 before the product, so it sits a little above the lane-group kernel
(measured 2.5e-3 vs 1.7e-3 relative L2) and inside the same bound."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools"))
import run_tc_check as tc  # noqa: E402
from dfvod_b200 import _lib  # noqa: E402
from oracle import msda_oracle  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", tc.CASES, ids=[c[0] for c in tc.CASES])
def test_tc_forward_matches_oracle_and_lane_group_kernel(case):
    name, shapes, n, m, p, dist, seed, lq = case
    dev = torch.device("cuda:0")
    value, loc, attn, _, lsi = tc.make_case(shapes, n, m, p, dist, seed, lq)
    st = torch.as_tensor(shapes, dtype=torch.long, device=dev)
    ls = torch.as_tensor(lsi, dtype=torch.long, device=dev)
    vb = value.to(torch.bfloat16)
    ref = msda_oracle.core_pytorch(vb.double(), shapes, loc.double(), attn.double())
    vd, ld, ad = vb.to(dev), loc.to(dev), attn.to(dev)
    out_tc = tc.fwd_call(vd, st, ls, ld, ad, _lib.FLAG_TC)
    out_lg = tc.fwd_call(vd, st, ls, ld, ad, 0)
    torch.cuda.synchronize()
    emax, el2 = tc.nerr(out_tc, ref)
    assert emax <= 2.0 ** -7 and el2 <= 4e-3, f"{name}: tensor-core path vs oracle {emax:.2e} {el2:.2e}"
    dmax, dl2 = tc.nerr(out_tc, out_lg.double())
    assert dmax <= 2.0 ** -6 and dl2 <= 6e-3, f"{name}: tensor-core path vs lane-group kernel {dmax:.2e} {dl2:.2e}"


def test_tc_forward_all_samples_outside_the_map_gives_zeros():
    dev = torch.device("cuda:0")
    shapes = [(24, 40), (12, 20)]
    value, loc, attn, _, lsi = tc.make_case(shapes, 2, 8, 4, "grid", 9)
    loc = loc + 3.0                                   # every sample far outside [0, 1)
    st = torch.as_tensor(shapes, dtype=torch.long, device=dev)
    ls = torch.as_tensor(lsi, dtype=torch.long, device=dev)
    out = tc.fwd_call(value.to(torch.bfloat16).to(dev), st, ls, loc.to(dev), attn.to(dev), _lib.FLAG_TC)
    torch.cuda.synchronize()
    assert not bool(out.float().abs().max() > 0)


def test_tc_flag_is_ignored_where_the_formulation_does_not_apply():
    """fp32 values, a head width other than 32, or a small problem: the flag falls through to the default kernels."""
    dev = torch.device("cuda:0")
    shapes = [(10, 12)]
    value, loc, attn, _, lsi = tc.make_case(shapes, 1, 4, 4, "random", 10, d=16)
    st = torch.as_tensor(shapes, dtype=torch.long, device=dev)
    ls = torch.as_tensor(lsi, dtype=torch.long, device=dev)
    for dtype in (torch.float32, torch.bfloat16):
        v = value.to(dtype)
        ref = msda_oracle.core_pytorch(v.double(), shapes, loc.double(), attn.double())
        out = tc.fwd_call(v.to(dev), st, ls, loc.to(dev), attn.to(dev), _lib.FLAG_TC)
        torch.cuda.synchronize()
        emax, _ = tc.nerr(out, ref)
        assert emax <= (1e-5 if dtype == torch.float32 else 2.0 ** -7)


@pytest.mark.parametrize("case", tc.CASES, ids=[c[0] for c in tc.CASES])
def test_tc_backward_grad_value_matches_oracle(case):
    """``msda_backward`` with MSDA_FLAG_TC: grad_value of every (tile, level) whose window fits is accumulated by
    msda_tc_dv_kernel (C^T . G on tcgen05, one bulk tensor reduction per window row), the rest and grad_loc / grad_attn
    by the lane-group kernel.  grad_value against the fp64 oracle within the bf16 bound; grad_loc / grad_attn must be
    BIT-IDENTICAL to the default path (the same kernel computes them)."""
    name, shapes, n, m, p, dist, seed, lq = case
    dev = torch.device("cuda:0")
    value, loc, attn, gout, lsi = tc.make_case(shapes, n, m, p, dist, seed, lq)
    st = torch.as_tensor(shapes, dtype=torch.long, device=dev)
    ls = torch.as_tensor(lsi, dtype=torch.long, device=dev)
    vb, gb = value.to(torch.bfloat16), gout.to(torch.bfloat16)
    v64 = vb.double().requires_grad_(True)
    msda_oracle.core_pytorch(v64, shapes, loc.double(), attn.double()).backward(gb.double())
    vd, ld, ad, gd = vb.to(dev), loc.to(dev), attn.to(dev), gb.to(dev)
    gv_tc, gl_tc, ga_tc = tc.bwd_call(vd, st, ls, ld, ad, gd, _lib.FLAG_TC)
    gv_lg, gl_lg, ga_lg = tc.bwd_call(vd, st, ls, ld, ad, gd, 0)
    torch.cuda.synchronize()
    emax, el2 = tc.nerr(gv_tc, v64.grad)
    assert emax <= 2.0 ** -7 and el2 <= 4e-3, f"{name}: grad_value {emax:.2e} {el2:.2e}"
    assert torch.equal(gl_tc, gl_lg) and torch.equal(ga_tc, ga_lg)
