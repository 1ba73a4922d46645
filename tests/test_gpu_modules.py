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


@pytest.mark.parametrize("name", sorted(set(module_cases.CASES) - module_cases.WIDE))
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
        ref = gold.get("grad_param." + k)
        if ref is None:
            continue
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
    # BASELINE.json's fp32 tolerance, 1e-5 (normalised max and relative L2 against the reference class in fp64), for
    # the output AND every gradient.  Measured on B200 (tools/measure_module_errors.py, round 2): outputs <= 1.0e-6,
    # input gradients <= 1.9e-6, parameter gradients <= 4.6e-6 over all cases.
    emax, el2 = nerr(out, gold["out"])
    assert emax <= 1e-5 and el2 <= 1e-5, f"out: {emax:.2e} {el2:.2e}"
    for k, g in gin.items():
        if gold["grad_in." + k].shape == () or g is None:
            continue
        emax, el2 = nerr(g, gold["grad_in." + k])
        assert emax <= 1e-5 and el2 <= 1e-5, f"{k}: {emax:.2e} {el2:.2e}"
    for k, g in gpar.items():
        ref = gold.get("grad_param." + k)
        if ref is None or ref.shape == () or g is None:
            continue
        emax, el2 = nerr(g, ref)
        assert emax <= 1e-5 and el2 <= 1e-5, f"{k}: {emax:.2e} {el2:.2e}"


@pytest.mark.parametrize("name", sorted(module_cases.CASES))
def test_fp32_inference_with_tensor_core_gemms_matches_reference_class(name):
    """set_fp32_gemm_mode("tf32x3") (the default), forced on for every row count: every nn.Linear of an fp32 inference pass runs as the three-term TF32 product
    (csrc/linear_tf32x3.cu where the shape allows, the split pass + library GEMM otherwise) -- same 1e-5 bound against the
    reference class in fp64 as the IEEE path."""
    from dfvod_b200.ops.functions import layer_epilogue_func as L
    gold = load_golden(name)
    prev_tf32, prev_rows = torch.backends.cuda.matmul.allow_tf32, L.TF32X3_MIN_ROWS
    torch.backends.cuda.matmul.allow_tf32 = False
    prev_mode = L.set_fp32_gemm_mode("tf32x3")
    L.TF32X3_MIN_ROWS = 1
    try:
        out = module_cases.run_case(name, gold, "cuda", torch.float32, forward_only=True)
    finally:
        L.set_fp32_gemm_mode(prev_mode)
        L.TF32X3_MIN_ROWS = prev_rows
        torch.backends.cuda.matmul.allow_tf32 = prev_tf32
    emax, el2 = nerr(out, gold["out"])
    assert emax <= 1e-5 and el2 <= 1e-5, f"out: {emax:.2e} {el2:.2e}"


def test_bf16_production_width_layer_matches_reference_class():
    """d_model 256 / 8 heads of 32 / 4 points in bf16: the fused bf16 deformable-attention kernels,
    the fused residual + LayerNorm (+ next query) kernels and the ReLU-epilogue GEMM inside the
    encoder layer, against the real reference class evaluated in fp64.  Stated bf16 tolerance for a
    whole layer (4 GEMMs + attention + 2 LayerNorms, every intermediate rounded to bf16):
    normalised max error 2^-5, relative L2 2^-6 on the output; gradients 2^-3 / 2^-5 (the gradient
    w.r.t. the sampling locations is discontinuous at pixel boundaries, and bf16 rounding of the query
    moves samples across them: measured 8e-2 / 2.7e-2 on src).  The gradient w.r.t. ``pos`` flows ONLY
    through that discontinuous path (query -> offsets / logits): relative L2 2^-3 (measured 8e-2)."""
    gold = load_golden("layer_encoder_c256")
    out, gin, gpar = module_cases.run_case("layer_encoder_c256", gold, "cuda", torch.bfloat16)
    emax, el2 = nerr(out, gold["out"])
    assert emax <= 2.0 ** -5 and el2 <= 2.0 ** -6, f"out: {emax:.2e} {el2:.2e}"
    for k, g in gin.items():
        emax, el2 = nerr(g, gold["grad_in." + k])
        if k == "pos":
            assert el2 <= 2.0 ** -3, f"{k}: {emax:.2e} {el2:.2e}"
        else:
            assert emax <= 2.0 ** -3 and el2 <= 2.0 ** -5, f"{k}: {emax:.2e} {el2:.2e}"
    for k, g in gpar.items():
        ref = gold.get("grad_param." + k)
        if ref is None or ref.shape == () or g is None:
            continue
        emax, el2 = nerr(g, ref)
        loose = "sampling_offsets" in k or "attention_weights" in k      # same discontinuous path
        assert el2 <= (2.0 ** -3 if loose else 2.0 ** -4), f"{k}: {emax:.2e} {el2:.2e}"


FUSION_C256 = ["layer_fusion_v2_c256", "layer_fusion_v2_c256_l2", "layer_late_fusion_c256"]


@pytest.mark.parametrize("name", FUSION_C256)
def test_bf16_production_width_fusion_layers_match_reference_class(name):
    """BASELINE.json configs[2] / [3]: Encoder Cross Fusion V2 (one- and two-level depth pyramid) and the Late Fusion
    layer at d_model 256 / 8 heads of 32 / 4 points in bf16, forward + backward, against the real reference classes
    (deformable_transformer_single.py:406-461, :341-402) evaluated in fp64.  Same stated bf16 layer tolerance as the
    encoder layer above: output 2^-5 max / 2^-6 L2; gradients 2^-3 max / 2^-5 L2, and 2^-3 L2 for what flows only
    through the sampling locations (query_pos; the offset / logit projections)."""
    gold = load_golden(name)
    out, gin, gpar = module_cases.run_case(name, gold, "cuda", torch.bfloat16)
    emax, el2 = nerr(out, gold["out"])
    assert emax <= 2.0 ** -5 and el2 <= 2.0 ** -6, f"out: {emax:.2e} {el2:.2e}"
    for k, g in gin.items():
        emax, el2 = nerr(g, gold["grad_in." + k])
        if k == "query_pos":
            assert el2 <= 2.0 ** -3, f"{k}: {emax:.2e} {el2:.2e}"
        else:
            assert emax <= 2.0 ** -3 and el2 <= 2.0 ** -5, f"{k}: {emax:.2e} {el2:.2e}"
    for k, g in gpar.items():
        ref = gold.get("grad_param." + k)
        if ref is None or ref.shape == () or g is None:
            continue
        emax, el2 = nerr(g, ref)
        loose = "sampling_offsets" in k or "attention_weights" in k
        assert el2 <= (2.0 ** -3 if loose else 2.0 ** -4), f"{k}: {emax:.2e} {el2:.2e}"


@pytest.mark.parametrize("name", FUSION_C256 + ["layer_encoder_c256"])
def test_bf16_production_width_inference_path_matches_reference_class(name):
    """The same layers under ``torch.no_grad()``: the route an inference takes -- the tcgen05 output-projection +
    residual + LayerNorm kernel (csrc/proj_fused.cu), GELU / ReLU folded into the norm kernels, the tcgen05 feed-forward
    block in the encoder layer -- inside the reference-shaped class, against the reference class's fp64 output."""
    gold = load_golden(name)
    case = module_cases.CASES[name]
    module = case["build"]().to(torch.bfloat16)
    module.load_state_dict({k[len("state."):]: torch.from_numpy(v) for k, v in gold.items() if k.startswith("state.")},
                           strict=True)
    module = module.cuda().eval()
    tensors = {}
    for k, v in gold.items():
        if k.startswith("in."):
            t = torch.from_numpy(v).cuda()
            tensors[k[len("in."):]] = t.to(torch.bfloat16) if t.is_floating_point() else t
    with torch.no_grad():
        out = case["call"](module, tensors)
    emax, el2 = nerr(out.double().cpu().numpy(), gold["out"])
    assert emax <= 2.0 ** -5 and el2 <= 2.0 ** -6, f"out: {emax:.2e} {el2:.2e}"
