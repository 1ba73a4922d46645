"""GPU: the paired value layout for the 16-bit forward (C ABI msda_pack_value_pairs,
msda_forward_paired, msda_fused_forward_paired) against
  * the layout definition  pairs[n,r,m,0|1,:] = value[n,r-1|r,m,:]  built with plain tensor ops,
  * the CPU oracle (fp64, on the 16-bit-rounded inputs) -- same tolerance as the unpaired 16-bit op,
  * the unpaired kernels on the same inputs (same arithmetic, different fp32 summation order)."""
import numpy as np
import pytest
import torch

from oracle import msda_oracle
from tests import util
from tests.util import nerr

pytestmark = pytest.mark.gpu

from dfvod_b200 import MultiScaleDeformableAttention as MSDA
from dfvod_b200.ops.functions import MSDeformAttnFusedFunction
from tests.test_gpu_fused import make_case, run_oracle64

DEV = "cuda"
BF16_MAX, BF16_L2 = 2.0 ** -7, 4e-3


@pytest.fixture
def paired(monkeypatch):
    def set_mode(mode, bf16_weights=False):
        monkeypatch.setattr(MSDA, "PAIRED_FORWARD", mode)
        monkeypatch.setattr(MSDA, "PAIRED_BF16_WEIGHTS", bf16_weights)
    return set_mode


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("d", [16, 32, 64])
def test_pack_layout(dtype, d):
    torch.manual_seed(d)
    n, s, m = 2, 37, 3
    value = torch.randn(n, s, m, d, device=DEV).to(dtype)
    pairs = MSDA.pack_value_pairs(value)
    assert pairs.shape == (n, s + 1, m, 2, d)
    want = torch.zeros_like(pairs)
    want[:, 1:, :, 0] = value
    want[:, :-1, :, 1] = value
    assert torch.equal(pairs, want)


CASES = [
    # shapes, N, M, D, Lq, P, loc_range
    ([(20, 30), (10, 15), (5, 8), (3, 4)], 2, 8, 32, 257, 4, (-0.2, 1.2)),   # production head layout + OOB
    ([(50, 84)], 2, 8, 32, 300, 4, (0.0, 1.0)),                     # shipped 1-level config
    ([(13, 21), (7, 11)], 3, 4, 64, 65, 4, (-0.1, 1.1)),            # D=64
    ([(9, 9)], 1, 16, 16, 33, 8, (-0.5, 1.5)),                      # D=16, P=8
    ([(5, 5)] * 5, 1, 2, 32, 9, 4, (0.0, 1.0)),                     # L*P = 20 > one 16-sample chunk
    ([(1, 1), (1, 7), (7, 1)], 1, 2, 32, 11, 4, (-0.3, 1.3)),       # degenerate 1-pixel-wide maps
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"L{len(c[0])}_M{c[2]}_D{c[3]}_Lq{c[4]}_P{c[5]}")
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_paired_op_vs_oracle_and_unpaired(case, dtype, paired):
    shapes, n, m, d, lq, p, rng = case
    value, loc, attn, _ = util.make_inputs(shapes, n, m, d, lq, p, seed=200 + d + lq, loc_range=rng)
    st, ls = util.shapes_tensors(shapes, DEV)
    v16 = value.to(dtype)
    args = (v16.to(DEV), st, ls, loc.to(DEV), attn.to(DEV), 64)
    paired(True)
    got = MSDA.ms_deform_attn_forward(*args)
    paired(False)
    plain = MSDA.ms_deform_attn_forward(*args)
    ref = msda_oracle.forward_np(v16.double().numpy(), shapes, util.lsi_of(shapes), loc.double().numpy(), attn.double().numpy())
    mx, l2 = nerr(got.double().cpu().numpy().reshape(ref.shape), ref)
    tol_mx, tol_l2 = (BF16_MAX, BF16_L2) if dtype == torch.bfloat16 else (2.0 ** -10, 5e-4)
    assert mx <= tol_mx and l2 <= tol_l2, (mx, l2)
    # same products, different fp32 summation order, one 16-bit rounding at the end
    mx, _ = nerr(got.double().cpu().numpy(), plain.double().cpu().numpy())
    assert mx <= (2.0 ** -7 if dtype == torch.bfloat16 else 2.0 ** -10)


def test_paired_is_opt_in_and_encoder_grid(paired):
    """The paired layout is opt-in (it ties with the plain gather on B200); when enabled
    it covers 16-bit values only."""
    assert MSDA.PAIRED_FORWARD is False
    assert not MSDA.use_paired_forward(torch.bfloat16, 32, 22223, 22223, 4, 4)
    paired(True)
    assert MSDA.use_paired_forward(torch.bfloat16, 32, 22223, 22223, 4, 4)
    assert not MSDA.use_paired_forward(torch.float32, 32, 22223, 22223, 4, 4)
    assert not MSDA.use_paired_forward(torch.bfloat16, 30, 22223, 22223, 4, 4)
    shapes = [(24, 40), (12, 20), (6, 10), (3, 5)]
    s = sum(h * w for h, w in shapes)
    value, loc, attn, _ = util.make_inputs(shapes, 2, 8, 32, s, 4, seed=6, dist="grid")
    st, ls = util.shapes_tensors(shapes, DEV)
    v16 = value.bfloat16()
    got = MSDA.ms_deform_attn_forward(v16.to(DEV), st, ls, loc.to(DEV), attn.to(DEV), 64)
    ref = msda_oracle.forward_np(v16.double().numpy(), shapes, util.lsi_of(shapes), loc.double().numpy(), attn.double().numpy())
    mx, l2 = nerr(got.double().cpu().numpy().reshape(ref.shape), ref)
    assert mx <= BF16_MAX and l2 <= BF16_L2


FUSED_CASES = [
    # shapes, N, M, D, Lq, P, ref_dim
    ([(20, 30), (10, 15), (5, 8), (3, 4)], 2, 8, 32, 131, 4, 2),
    ([(20, 30), (10, 15), (5, 8), (3, 4)], 1, 8, 32, 50, 4, 4),
    ([(50, 84)], 2, 8, 32, 77, 4, 2),
    ([(12, 9), (6, 5)], 2, 4, 16, 39, 4, 2),
    ([(12, 9), (6, 5)], 1, 4, 64, 21, 4, 4),
]


@pytest.mark.parametrize("case", FUSED_CASES, ids=lambda c: f"L{len(c[0])}_M{c[2]}_D{c[3]}_P{c[5]}_ref{c[6]}")
@pytest.mark.parametrize("rdtype", [torch.float32, torch.bfloat16])
def test_paired_fused_forward(case, rdtype, paired):
    shapes, n, m, d, lq, p, ref_dim = case
    nl = len(shapes)
    value, raw, ref, gout = make_case(shapes, n, m, d, lq, p, ref_dim, seed=31 + d)
    st, ls = util.shapes_tensors(shapes, DEV)
    v = value.to(DEV, torch.bfloat16)
    r = raw.to(DEV, rdtype)
    rf = ref.to(DEV)
    with torch.no_grad():
        paired(True)
        got = MSDeformAttnFusedFunction.apply(v, st, ls, rf, r, p)
        paired(False)
        plain = MSDeformAttnFusedFunction.apply(v, st, ls, rf, r, p)
    want = run_oracle64(v.float().cpu(), shapes, ref, r.float().cpu(), gout, m, nl, p)[0]
    mx, l2 = nerr(got.double().cpu().numpy(), want)
    assert mx <= BF16_MAX and l2 <= BF16_L2, (mx, l2)
    mx, _ = nerr(got.double().cpu().numpy(), plain.double().cpu().numpy())
    assert mx <= 2.0 ** -8


def test_paired_fused_backward_unchanged(paired):
    """Training: the forward may read the paired layout, the backward uses the saved plain value."""
    shapes, n, m, d, lq, p, ref_dim = FUSED_CASES[0]
    value, raw, ref, gout = make_case(shapes, n, m, d, lq, p, ref_dim, seed=5)
    st, ls = util.shapes_tensors(shapes, DEV)
    grads = []
    for mode in (True, False):
        paired(mode)
        v = value.to(DEV, torch.bfloat16).requires_grad_(True)
        r = raw.to(DEV, torch.bfloat16).requires_grad_(True)
        out = MSDeformAttnFusedFunction.apply(v, st, ls, ref.to(DEV), r, p)
        out.backward(gout.to(DEV, torch.bfloat16))
        grads.append((v.grad.clone(), r.grad.clone()))
    assert torch.equal(grads[0][0], grads[1][0]) or nerr(grads[0][0].double().cpu().numpy(),
                                                          grads[1][0].double().cpu().numpy())[0] <= 2.0 ** -7
    assert nerr(grads[0][1].double().cpu().numpy(), grads[1][1].double().cpu().numpy())[0] <= 2.0 ** -7


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"L{len(c[0])}_M{c[2]}_D{c[3]}_Lq{c[4]}_P{c[5]}")
def test_paired_bf16_weights_op(case, paired):
    """Paired layout + bf16-rounded per-corner weights + mixed-precision FMA (FHFMA.BF16): still inside the
    stated bf16 tolerance against the fp64 oracle (normalised max 2^-7, relative L2 4e-3)."""
    shapes, n, m, d, lq, p, rng = case
    value, loc, attn, _ = util.make_inputs(shapes, n, m, d, lq, p, seed=300 + d + lq, loc_range=rng)
    st, ls = util.shapes_tensors(shapes, DEV)
    v16 = value.bfloat16()
    paired(True, True)
    got = MSDA.ms_deform_attn_forward(v16.to(DEV), st, ls, loc.to(DEV), attn.to(DEV), 64)
    ref = msda_oracle.forward_np(v16.double().numpy(), shapes, util.lsi_of(shapes), loc.double().numpy(), attn.double().numpy())
    mx, l2 = nerr(got.double().cpu().numpy().reshape(ref.shape), ref)
    assert mx <= BF16_MAX and l2 <= BF16_L2, (mx, l2)


@pytest.mark.parametrize("rdtype", [torch.float32, torch.bfloat16])
def test_paired_bf16_weights_fused(rdtype, paired):
    shapes, n, m, d, lq, p, ref_dim = FUSED_CASES[0]
    nl = len(shapes)
    value, raw, ref, gout = make_case(shapes, n, m, d, lq, p, ref_dim, seed=77)
    st, ls = util.shapes_tensors(shapes, DEV)
    v = value.to(DEV, torch.bfloat16)
    r = raw.to(DEV, rdtype)
    with torch.no_grad():
        paired(True, True)
        got = MSDeformAttnFusedFunction.apply(v, st, ls, ref.to(DEV), r, p)
    want = run_oracle64(v.float().cpu(), shapes, ref, r.float().cpu(), gout, m, nl, p)[0]
    mx, l2 = nerr(got.double().cpu().numpy(), want)
    assert mx <= BF16_MAX and l2 <= BF16_L2, (mx, l2)
