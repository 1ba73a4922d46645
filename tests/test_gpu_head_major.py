"""GPU: the head-major inference path -- value_proj as an own tcgen05 GEMM with a head-major epilogue
(csrc/value_proj_hm.cu), the gather kernels reading value_hm [N, M, S, D] (csrc/msda_forward_hm.cu) -- against the
library composition and the default kernels."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools"))
import run_tc_check as tc  # noqa: E402
from dfvod_b200 import _lib  # noqa: E402
from dfvod_b200.ops import modules as ops_modules  # noqa: E402
from dfvod_b200.ops.functions import value_proj_head_major  # noqa: E402
from dfvod_b200.ops.modules import MSDeformAttn  # noqa: E402
from dfvod_b200.ops.modules import ms_deform_attn as msda_module  # noqa: E402
from oracle import msda_oracle  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


@pytest.mark.parametrize("case", tc.CASES, ids=[c[0] for c in tc.CASES])
def test_plain_forward_head_major_matches_oracle(case):
    name, shapes, n, m, p, dist, seed, lq = case
    value, loc, attn, _, lsi = tc.make_case(shapes, n, m, p, dist, seed, lq)
    st = torch.as_tensor(shapes, dtype=torch.long, device=DEV)
    ls = torch.as_tensor(lsi, dtype=torch.long, device=DEV)
    vb = value.to(torch.bfloat16)
    ref = msda_oracle.core_pytorch(vb.double(), shapes, loc.double(), attn.double())
    vd, ld, ad = vb.to(DEV), loc.to(DEV), attn.to(DEV)
    vhm = vd.permute(0, 2, 1, 3).contiguous().view(vd.shape)      # [N, M, S, D] bytes behind the [N, S, M, D] sizes
    out = tc.fwd_call(vhm, st, ls, ld, ad, _lib.FLAG_VALUE_HEAD_MAJOR)
    torch.cuda.synchronize()
    emax, el2 = tc.nerr(out, ref)
    assert emax <= 2.0 ** -7 and el2 <= 4e-3, f"{name}: {emax:.2e} {el2:.2e}"


@pytest.mark.parametrize("rows_per_frame,frames,masked", [(300, 3, True), (128, 1, False), (22223, 2, True), (1, 5, True)])
def test_value_proj_head_major_matches_linear(rows_per_frame, frames, masked):
    torch.manual_seed(3)
    lin = torch.nn.Linear(256, 256).to(DEV).bfloat16()
    x = torch.randn(frames, rows_per_frame, 256, device=DEV).bfloat16()
    mask = (torch.rand(frames, rows_per_frame, device=DEV) < 0.2) if masked else None
    got = value_proj_head_major(lin, x, mask, 8)                   # [N, 8, S, 32]
    want = torch.nn.functional.linear(x.float(), lin.weight.float(), lin.bias.float())
    if mask is not None:
        want = want.masked_fill(mask[..., None], 0.0)
    want = want.view(frames, rows_per_frame, 8, 32).permute(0, 2, 1, 3)
    torch.cuda.synchronize()
    err = (got.float() - want).abs().max() / want.abs().max().clamp_min(1e-6)
    assert float(err) <= 2.0 ** -8, float(err)                     # one bf16 rounding of an fp32-accumulated product
    if mask is not None:
        assert not bool(got.permute(0, 2, 1, 3)[mask].any())


@pytest.mark.parametrize("ref_dim", [2, 4])
def test_module_inference_head_major_equals_reference_layout(ref_dim):
    """MSDeformAttn under no_grad in bf16 at d_model 256: the head-major route (opt-in) against the same module with
    HEAD_MAJOR_INFERENCE off -- the two differ only in the rounding order of the value GEMM / the gather."""
    torch.manual_seed(7)
    shapes = [(20, 30), (10, 15)]
    s = sum(h * w for h, w in shapes)
    st = torch.as_tensor(shapes, dtype=torch.long, device=DEV)
    ls = torch.as_tensor([0, 600], dtype=torch.long, device=DEV)
    mod = MSDeformAttn(256, 2, 8, 4).to(DEV).bfloat16().eval()
    with torch.no_grad():
        for prm in mod.parameters():
            prm.add_(0.03 * torch.randn_like(prm))
    n, lq = 2, s if ref_dim == 2 else 77
    query = torch.randn(n, lq, 256, device=DEV).bfloat16()
    feat = torch.randn(n, s, 256, device=DEV).bfloat16()
    mask = torch.zeros(n, s, dtype=torch.bool, device=DEV)
    mask[1, -50:] = True
    if ref_dim == 2:
        ref = torch.rand(n, lq, 2, 2, device=DEV) * 0.9 + 0.05
    else:
        ref = torch.cat([torch.rand(n, lq, 2, 2, device=DEV) * 0.6 + 0.2, torch.rand(n, lq, 2, 2, device=DEV) * 0.3 + 0.05], -1)
    outs = []
    before = msda_module.HEAD_MAJOR_INFERENCE
    try:
        for flag in (True, False):
            msda_module.HEAD_MAJOR_INFERENCE = flag
            with torch.no_grad():
                outs.append(mod(query, ref, feat, st, ls, mask).float())
    finally:
        msda_module.HEAD_MAJOR_INFERENCE = before
    torch.cuda.synchronize()
    err = (outs[0] - outs[1]).abs().max() / outs[1].abs().max()
    assert float(err) <= 2.0 ** -6, float(err)
