"""CPU: the host-side mirror (MSDeformAttn, the fusion / encoder / decoder layer classes,
fuse_layers) reproduces the REAL reference classes' outputs and gradients
(tests/golden/{module,layer,encoder,backbone}_*.npz, made by oracle/gen_golden.py) when the CUDA
op is replaced by the oracle -- i.e. everything around the kernel is parity-checked without a
GPU.  (The product never does this substitution; it happens here, in tests/, only.)"""
import numpy as np
import pytest
import torch

from oracle import msda_oracle, roi_align_oracle
from tests import module_cases
from tests.util import load_golden

import dfvod_b200.ops.modules.ms_deform_attn as msda_module


class _OracleFunction:
    @staticmethod
    def apply(value, shapes, lsi, loc, attn, im2col_step):
        return msda_oracle.core_pytorch(value, shapes.tolist(), loc, attn)


@pytest.fixture()
def oracle_op(monkeypatch):
    from dfvod_b200 import temporal_stage
    monkeypatch.setattr(msda_module, "MSDeformAttnFunction", _OracleFunction)
    # the op has no gradient w.r.t. the boxes (mmcv RoIAlignFunction.backward returns None for them)
    monkeypatch.setattr(temporal_stage, "roi_align_tokens",
                        lambda tokens, rois, *a, **k: roi_align_oracle.roi_align_tokens(tokens, rois.detach(), *a, **k))


@pytest.mark.parametrize("name", sorted(module_cases.CASES))
def test_module_matches_reference_class(name, oracle_op):
    gold = load_golden(name)
    out, gin, gpar = module_cases.run_case(name, gold, "cpu")
    # wide cases keep fp32 copies of the reference's fp64 results (oracle/gen_golden.py)
    wide = name in module_cases.WIDE
    rt_o, at_o = (2e-6, 2e-6) if wide else (1e-9, 1e-11)
    rt_g, at_g = (2e-6, 2e-6) if wide else (1e-8, 1e-10)
    np.testing.assert_allclose(out, gold["out"], rtol=rt_o, atol=at_o)
    for k, g in gin.items():
        if gold["grad_in." + k].shape == ():          # input unused by this variant
            assert g is None or not np.any(g)
            continue
        np.testing.assert_allclose(g, gold["grad_in." + k], rtol=rt_g, atol=at_g, err_msg=k)
    for k, g in gpar.items():
        ref = gold.get("grad_param." + k)
        if ref is None:              # large matrices of the wide cases are not stored
            continue
        if ref.shape == ():          # unused parameter in the reference
            assert g is None or not np.any(g)
            continue
        np.testing.assert_allclose(g, ref, rtol=rt_g, atol=at_g, err_msg=k)


def test_state_dict_keys_and_init_match_reference():
    """Parameter names are the checkpoint contract; init follows
    ms_deform_attn.py:60-76."""
    m = msda_module.MSDeformAttn(256, 4, 8, 4)
    assert sorted(m.state_dict()) == sorted(
        f"{n}.{p}" for n in ("sampling_offsets", "attention_weights", "value_proj", "output_proj")
        for p in ("weight", "bias"))
    assert sum(p.numel() for p in m.parameters()) == 230272
    assert m.im2col_step == 64
    assert not m.sampling_offsets.weight.any() and not m.attention_weights.weight.any()
    assert not m.attention_weights.bias.any() and not m.value_proj.bias.any() and not m.output_proj.bias.any()
    bias = m.sampling_offsets.bias.view(8, 4, 4, 2)
    # head 0 looks along +x, head 2 along +y; point i sits at i+1 pixels; same on every level
    np.testing.assert_allclose(bias[0, :, :, 0].detach().numpy(), np.tile([1., 2., 3., 4.], (4, 1)), atol=1e-6)
    np.testing.assert_allclose(bias[0, :, :, 1].detach().numpy(), 0, atol=1e-6)
    np.testing.assert_allclose(bias[2, :, :, 1].detach().numpy(), np.tile([1., 2., 3., 4.], (4, 1)), atol=1e-6)
    np.testing.assert_allclose(bias[1, 0, 0].detach().numpy(), [1., 1.], atol=1e-6)     # 45 deg normalised by max
    gold = load_golden("module_init")
    ref = msda_module.MSDeformAttn(32, 2, 4, 3)
    np.testing.assert_allclose(ref.sampling_offsets.bias.detach().numpy(), gold["sampling_offsets.bias"], atol=1e-7)


def test_bad_reference_point_width_raises(oracle_op):
    m = msda_module.MSDeformAttn(32, 1, 4, 2)
    shapes = torch.tensor([[2, 3]])
    with pytest.raises(ValueError, match="Last dim of reference_points must be 2 or 4"):
        m(torch.zeros(1, 6, 32), torch.zeros(1, 6, 1, 3), torch.zeros(1, 6, 32), shapes, torch.tensor([0]))


def test_length_mismatch_asserts(oracle_op):
    m = msda_module.MSDeformAttn(32, 1, 4, 2)
    with pytest.raises(AssertionError):
        m(torch.zeros(1, 6, 32), torch.zeros(1, 6, 1, 2), torch.zeros(1, 7, 32), torch.tensor([[2, 3]]),
          torch.tensor([0]))


def test_constructor_validation():
    with pytest.raises(ValueError, match="d_model must be divisible by n_heads"):
        msda_module.MSDeformAttn(30, 1, 4, 2)
    with pytest.warns(UserWarning, match="power of 2"):
        msda_module.MSDeformAttn(24, 1, 2, 2)


def test_transvodpp_clips_batch_like_single_clips(oracle_op):
    """Several clips in one batch (clip-major) give, per clip, what the reference-shaped single-clip call gives."""
    torch.manual_seed(21)
    model = module_cases._transvodpp("DepthDeform_latefusion_dformer", True, 2).double().eval()
    frames, c, (fh, fw) = 3, 16, (4, 6)
    mk = lambda *s: torch.randn(*s, dtype=torch.float64)
    ins = dict(src0=mk(2 * frames, c, fh, fw), pos0=mk(2 * frames, c, fh, fw), depth_src0=mk(2 * frames, c, fh, fw),
               depth_pos0=mk(2 * frames, c, fh, fw), query_embed=mk(80, 2 * c))
    mask = torch.zeros(2 * frames, fh, fw, dtype=torch.bool)
    mask[4, :, 4:] = True
    ins["mask0"] = mask
    ins["depth_mask0"] = mask.clone()
    sizes = torch.tensor([[fw * 32, fh * 32, fw * 32, fh * 32], [fw * 30, fh * 28, fw * 30, fh * 28]])

    def run(sel, whwh):
        t = {k: (v[sel] if k != "query_embed" else v) for k, v in ins.items()}
        t["imgs_whwh"] = whwh
        hs, init_ref, inter_ref, _, _, final_hs, final_ref, out = model.transformer(
            [t["src0"]], [t["mask0"]], [t["pos0"]], [t["depth_src0"]], [t["depth_mask0"]], [t["depth_pos0"]],
            t["imgs_whwh"], t["query_embed"], model.class_embed[-1], model.bbox_embed[-1],
            model.temp_class_embed_list, model.temp_bbox_embed_list)
        return hs, init_ref, inter_ref, final_hs, final_ref, out["aux_outputs"][1]["pred_boxes"]

    with torch.no_grad():
        both = run(slice(0, 2 * frames), sizes)
        for clip in range(2):
            one = run(slice(clip * frames, (clip + 1) * frames), sizes[clip:clip + 1])
            for name, b, o in zip(("hs", "init_ref", "inter_ref", "final_hs", "final_ref", "aux_boxes"), both, one):
                b = b[:, clip:clip + 1] if name in ("hs", "inter_ref") else b[clip:clip + 1]
                np.testing.assert_allclose(b.numpy(), o.numpy(), rtol=1e-9, atol=1e-11, err_msg=name)


def test_input_projection_is_the_reference_sequential():
    """input_projection.InputProjection keeps the reference's nn.Sequential(Conv2d, GroupNorm(32, hidden)) state-dict
    keys (deformable_detr_single.py:101-123) and NCHW forward; forward_tokens is its flatten(2).transpose(1, 2)."""
    from dfvod_b200.input_projection import InputProjection
    torch.manual_seed(0)
    for kwargs in (dict(kernel_size=1), dict(kernel_size=3, stride=2, padding=1)):
        proj = InputProjection(24, 64, **kwargs).double()
        ref = torch.nn.Sequential(torch.nn.Conv2d(24, 64, **kwargs), torch.nn.GroupNorm(32, 64)).double()
        assert sorted(proj.state_dict()) == sorted(ref.state_dict()) == ["0.bias", "0.weight", "1.bias", "1.weight"]
        with torch.no_grad():
            for prm in proj.parameters():
                prm.add_(torch.randn_like(prm) * 0.2)
        ref.load_state_dict(proj.state_dict())
        x = torch.randn(2, 24, 7, 9, dtype=torch.float64)
        want = ref(x)
        np.testing.assert_allclose(proj(x).detach().numpy(), want.detach().numpy(), rtol=1e-12, atol=1e-12)
        tokens, (h, w) = proj.forward_tokens(x)
        assert (h, w) == tuple(want.shape[2:])
        np.testing.assert_allclose(tokens.detach().numpy(), want.flatten(2).transpose(1, 2).detach().numpy(),
                                   rtol=1e-10, atol=1e-12)


def test_transformer_accepts_token_major_levels(oracle_op):
    """Levels handed over as [N, H*W, C] tokens (InputProjection.forward_tokens) give what NCHW maps give."""
    from dfvod_b200.deformable_transformer import DeformableTransformer
    torch.manual_seed(6)
    shapes = [(5, 7), (3, 4)]
    model = DeformableTransformer(d_model=32, nhead=4, num_encoder_layers=1, num_decoder_layers=1, dim_feedforward=64,
                                  dropout=0.0, num_feature_levels=2, return_intermediate_dec=True).double().eval()
    srcs = [torch.randn(2, 32, h, w, dtype=torch.float64) for h, w in shapes]
    poss = [torch.randn(2, 32, h, w, dtype=torch.float64) for h, w in shapes]
    masks = [torch.zeros(2, h, w, dtype=torch.bool) for h, w in shapes]
    masks[0][1, :, -2:] = True
    query = torch.randn(6, 64, dtype=torch.float64)
    with torch.no_grad():
        a = model(srcs, masks, poss, None, None, None, query)[0]
        b = model([s.flatten(2).transpose(1, 2).contiguous() for s in srcs], masks, poss, None, None, None, query)[0]
        c = model([srcs[0].flatten(2).transpose(1, 2), srcs[1]], masks, poss, None, None, None, query)[0]
    assert torch.equal(a, b) and torch.equal(a, c)
    with pytest.raises(AssertionError, match="disagree"):
        model([srcs[0][:, :, :-1]] + srcs[1:], masks, poss, None, None, None, query)


def test_tf32_split_definition_and_three_term_product():
    """Host restatement of the split behind set_fp32_gemm_mode('tf32x3'): hi keeps 10 mantissa bits, lo + hi == x
    exactly, |lo| <= 2^-11 |x|, and [lo | hi | hi] x [hi | lo | hi]^T equals x W^T minus the (dropped) lo_x lo_W^T term."""
    import torch
    from dfvod_b200.ops.functions import layer_epilogue_func as L
    torch.manual_seed(0)
    x = torch.randn(64, 24) * torch.exp(torch.randn(64, 24) * 3)
    w = torch.randn(10, 24)
    parts = L._tf32_split(x)
    lo, hi, hi2 = parts[:, :24], parts[:, 24:48], parts[:, 48:]
    assert torch.equal(hi, hi2) and torch.equal(lo + hi, x)
    assert int((hi.contiguous().view(torch.int32) & 0x1fff).abs().max()) == 0
    assert bool((lo.abs() <= x.abs() * 2.0 ** -11).all())
    wp = L._tf32_split(w)
    wlo, whi = wp[:, :24], wp[:, 24:48]
    got = L.linear_tf32x3(x, w, None).double()
    want = x.double() @ w.double().t() - lo.double() @ wlo.double().t()
    scale = (x.double().abs() @ w.double().abs().t()).max()
    assert float((got - want).abs().max() / scale) <= 1e-6        # fp32 accumulation of exact terms (on the host)
    assert L.FP32_GEMM_MODE == "tf32x3" and not L._tf32x3_wanted(x, w)   # host tensors never take the split path
