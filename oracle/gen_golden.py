"""Generate tests/golden/*.npz by importing the REAL reference from /root/reference.

TEST INFRASTRUCTURE ONLY.  Run in the dev container (the reference tree does not travel
to the GPU box):   python oracle/gen_golden.py

What is executed is the reference's own code, unmodified:
  * ``ms_deform_attn_core_pytorch`` (models/ops/functions/ms_deform_attn_func.py:41-61)
    and autograd through it -- op-level vectors;
  * the reference ``MSDeformAttn`` module (models/ops/modules/ms_deform_attn.py) and the
    layer classes of models/deformable_transformer_single.py /
    models/dformer_crossfusion_backbone.py, with ``MSDeformAttnFunction.apply`` replaced by
    the reference oracle (the CUDA extension cannot be imported without a GPU build) --
    module/layer-level vectors.
Import shims: an empty ``MultiScaleDeformableAttention`` stub so
func.py:18 imports; a bare ``models`` namespace package to bypass models/__init__.py
(which pulls mmcv); torchvision.__version__ patched while util/misc.py:30 parses it.

Everything is float64 unless the case says otherwise; inputs are drawn with the CPU
generator in the order the reference test recipe uses (models/ops/test.py:28,33-36).
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("MSDA_REFERENCE", "/root/reference")
ONLY = [a for a in sys.argv[1:] if not a.startswith("-")]      # python oracle/gen_golden.py [case-name-prefix ...]


def wanted(name):
    return not ONLY or any(name.startswith(o) for o in ONLY)

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def import_reference():
    sys.modules.setdefault("MultiScaleDeformableAttention", types.ModuleType("MultiScaleDeformableAttention"))
    ops = os.path.join(REF, "models", "ops")
    models_pkg = types.ModuleType("models")
    models_pkg.__path__ = [os.path.join(REF, "models")]
    sys.modules["models"] = models_pkg
    ops_pkg = types.ModuleType("models.ops")
    ops_pkg.__path__ = [ops]
    sys.modules["models.ops"] = ops_pkg
    sys.path.insert(0, REF)
    func = importlib.import_module("models.ops.functions.ms_deform_attn_func")

    class OracleFunction:  # stands in for the CUDA autograd.Function (func.py:21-38)
        @staticmethod
        def apply(value, shapes, lsi, loc, attn, im2col_step):
            return func.ms_deform_attn_core_pytorch(value, shapes, loc, attn)

    func.MSDeformAttnFunction = OracleFunction
    sys.modules["models.ops.functions"].MSDeformAttnFunction = OracleFunction
    mod = importlib.import_module("models.ops.modules.ms_deform_attn")
    mod.MSDeformAttnFunction = OracleFunction
    import torchvision
    real_version = torchvision.__version__
    torchvision.__version__ = "0.9.0"
    try:
        importlib.import_module("util.misc")
    finally:
        torchvision.__version__ = real_version
    single = importlib.import_module("models.deformable_transformer_single")
    backbone_cf = importlib.import_module("models.dformer_crossfusion_backbone")
    return func, mod, single, backbone_cf


def import_transvod_plusplus():
    """models/deformable_transformer_multi_plusplus.py does ``from mmcv import ops`` (:25) for the
    RoIAlign of the temporal query stage; mmcv-full 1.7.0 is not installed, so a stub module stands
    in whose RoIAlign has mmcv's constructor (output_size, spatial_scale, sampling_ratio,
    pool_mode='avg', aligned=True) and computes with torchvision.ops.roi_align on CPU -- the same
    Detectron algorithm (oracle/roi_align_oracle.py restates it and is pinned against it)."""
    import torchvision
    mmcv = types.ModuleType("mmcv")
    ops = types.ModuleType("mmcv.ops")

    class RoIAlign(torch.nn.Module):
        def __init__(self, output_size, spatial_scale=1.0, sampling_ratio=0, pool_mode="avg", aligned=True,
                     use_torchvision=False):
            super().__init__()
            assert pool_mode == "avg"
            self.args = (output_size, spatial_scale, sampling_ratio, aligned)

        def forward(self, input, rois):
            return torchvision.ops.roi_align(input, rois.to(input.dtype), *self.args)

    ops.RoIAlign = RoIAlign
    mmcv.ops = ops
    sys.modules.setdefault("mmcv", mmcv)
    sys.modules.setdefault("mmcv.ops", ops)
    return importlib.import_module("models.deformable_transformer_multi_plusplus")


def lsi_of(shapes):
    return torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))


def op_case(func, name, shapes, n, m, d, lq, p, seed, loc_range=(0.0, 1.0), dtype=torch.float64):
    """The recipe of models/ops/test.py:33-36 (fp32 CPU draws, then cast)."""
    if not wanted("op_" + name):
        return None
    torch.manual_seed(seed)
    shapes_t = torch.as_tensor(shapes, dtype=torch.long)
    nl = len(shapes)
    s = int(shapes_t.prod(1).sum())
    f32 = torch.float32
    value = torch.rand(n, s, m, d, dtype=f32) * 0.01
    loc = torch.rand(n, lq, m, nl, p, 2, dtype=f32)
    attn = torch.rand(n, lq, m, nl, p, dtype=f32) + 1e-5
    attn /= attn.sum(-1, keepdim=True).sum(-2, keepdim=True)
    grad_out = torch.rand(n, lq, m * d, dtype=f32) - 0.5
    lo, hi = loc_range
    loc = loc * (hi - lo) + lo
    value, loc, attn, grad_out = (t.to(dtype) for t in (value, loc, attn, grad_out))
    v = value.clone().requires_grad_(True)
    lc = loc.clone().requires_grad_(True)
    aw = attn.clone().requires_grad_(True)
    out = func.ms_deform_attn_core_pytorch(v, shapes_t, lc, aw)
    gv, gl, ga = torch.autograd.grad(out, (v, lc, aw), grad_out)
    np.savez_compressed(
        os.path.join(OUT, f"op_{name}.npz"),
        shapes=shapes_t.numpy(), level_start_index=lsi_of(shapes_t).numpy(),
        value=value.numpy(), loc=loc.numpy(), attn=attn.numpy(), grad_out=grad_out.numpy(),
        out=out.detach().numpy(), grad_value=gv.numpy(), grad_loc=gl.numpy(), grad_attn=ga.numpy())
    print(f"op_{name}: out {tuple(out.shape)} |out|max {out.abs().max():.3e}")
    return out


def state_np(module):
    return {k: v.detach().numpy() for k, v in module.state_dict().items()}


def perturb(module, seed, std=0.05):
    """Default init makes offsets query-independent and attention uniform
    (ms_deform_attn.py:60-76); perturb so every parameter matters."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for prm in module.parameters():
            prm.add_(torch.randn(prm.shape, generator=g, dtype=prm.dtype) * std)


def grid_reference_points(single, shapes_t, n, dtype):
    vr = torch.ones(n, shapes_t.shape[0], 2, dtype=dtype)
    ref = single.DeformableTransformerEncoder.get_reference_points(
        [(int(h), int(w)) for h, w in shapes_t], vr, device="cpu")
    return ref.to(dtype)


def round_to_f32(module, inputs):
    """Make every parameter and floating-point input exactly fp32-representable (so that an fp32
    copy of the case is lossless) while the reference still computes in fp64."""
    with torch.no_grad():
        for prm in module.parameters():
            prm.copy_(prm.float().double())
    return {k: (v.float().double() if torch.is_tensor(v) and v.dtype == torch.float64 else v)
            for k, v in inputs.items()}


def save_module_case(name, module, inputs, call, wrt, store=None, grad_limit=None, state_store=None):
    """Run ``call(module, **inputs)``, save inputs/state/output and the gradients of
    sum(output * gout) w.r.t. the tensors named in ``wrt`` and every parameter.
    ``store=np.float32`` (wide cases): parameters and inputs are first rounded to fp32-representable
    values, the reference computes in fp64, and the file keeps fp32 copies (lossless for inputs and
    state, 1e-7 for outputs / gradients -- these cases serve the fp32 / bf16 tests only).
    ``grad_limit``: parameter gradients larger than this many elements are not stored."""
    if not wanted(name):
        return
    if store is not None:
        inputs = round_to_f32(module, inputs)
    if state_store is not None:          # parameters made exactly representable in `state_store` (e.g. fp16): the
        with torch.no_grad():            # state is then kept in that type without loss
            for prm in module.parameters():
                prm.copy_(prm.to(getattr(torch, np.dtype(state_store).name)).double())
    tensors = {k: (v.clone().requires_grad_(True) if k in wrt else v) for k, v in inputs.items()}
    out = call(module, tensors)
    g = torch.Generator().manual_seed(1234)
    gout = torch.randn(out.shape, generator=g, dtype=out.dtype)
    params = list(module.parameters())
    grads = torch.autograd.grad(out, [tensors[k] for k in wrt] + params, gout, allow_unused=True)
    blob = {f"in.{k}": (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in inputs.items()}
    blob.update({f"state.{k}": v for k, v in state_np(module).items()})
    blob["out"] = out.detach().numpy()
    blob["gout"] = gout.numpy()
    for k, gr in zip(wrt, grads[:len(wrt)]):
        blob[f"grad_in.{k}"] = (gr if gr is not None else torch.zeros(())).numpy()
    for (pname, _), gr in zip(module.named_parameters(), grads[len(wrt):]):
        if grad_limit is not None and gr is not None and gr.numel() > grad_limit and "value_proj" not in pname:
            continue
        blob[f"grad_param.{pname}"] = (gr if gr is not None else torch.zeros(())).numpy()
    if state_store is not None:
        blob = {k: (v.astype(state_store) if k.startswith("state.") and v.dtype == np.float64 else v)
                for k, v in blob.items()}
    if store is not None:
        blob = {k: (v.astype(store) if v.dtype == np.float64 else v) for k, v in blob.items()}
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **blob)
    print(f"{name}: out {tuple(out.shape)} |out|max {out.abs().max():.3e}")


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_default_dtype(torch.float64)
    func, mod, single, backbone_cf = import_reference()

    # ---------------- op-level ------------------------------------------------
    # (1) the reference test's own configuration and seed (test.py:21-28)
    out = op_case(func, "toy_seed3", [(6, 4), (3, 2)], n=1, m=2, d=2, lq=2, p=2, seed=3)
    known = [0.001899378416, 0.004602827533, 0.004671175247, 0.004384399819,
             0.003795097174, 0.002512764199, 0.001844426151, 0.003634679248]
    assert out is None or np.allclose(out.detach().numpy().ravel(), known, rtol=0, atol=1e-11), \
        "RNG stream differs from survey"
    # (2) locations outside [0,1): exercises the zero padding and the in-range test (cuh:288)
    op_case(func, "oob", [(5, 7), (3, 4), (2, 2)], n=2, m=3, d=8, lq=5, p=3, seed=11, loc_range=(-0.3, 1.3))
    # (3) production head width D=32, M=8, one level (shipped configs use 1 level)
    op_case(func, "d32_l1", [(7, 9)], n=2, m=8, d=32, lq=6, p=4, seed=12, loc_range=(-0.1, 1.1))
    # (4) four levels, D=32, M=8, P=4 (the COCO-scale head layout, tiny maps)
    op_case(func, "d32_l4", [(10, 13), (5, 7), (3, 4), (2, 2)], n=1, m=8, d=32, lq=9, p=4, seed=13,
            loc_range=(-0.05, 1.05))
    # (5) odd channel count as in the reference gradcheck list (test.py:85 uses 30, 71)
    op_case(func, "d30", [(6, 4), (3, 2)], n=1, m=2, d=30, lq=3, p=2, seed=14)
    # (6) fp32 evaluation of case 3 (reference float check, test.py:47-60)
    op_case(func, "d32_l1_f32", [(7, 9)], n=2, m=8, d=32, lq=6, p=4, seed=12, loc_range=(-0.1, 1.1),
            dtype=torch.float32)

    # ---------------- MSDeformAttn module (ms_deform_attn.py:31-117) ----------
    init_mod = mod.MSDeformAttn(32, 2, 4, 3)            # default initialisation (:60-76)
    np.savez_compressed(os.path.join(OUT, "module_init.npz"),
                        **{k: v.detach().numpy() for k, v in init_mod.state_dict().items()
                           if k.startswith("sampling_offsets") or k.startswith("attention_weights")})
    shapes_t = torch.as_tensor([(6, 5), (3, 3)], dtype=torch.long)
    lsi = lsi_of(shapes_t)
    s = int(shapes_t.prod(1).sum())
    n, c, heads, pts = 2, 32, 4, 3
    torch.manual_seed(21)
    attn_mod = mod.MSDeformAttn(c, 2, heads, pts).double()
    perturb(attn_mod, 22)
    query = torch.randn(n, s, c)
    feat = torch.randn(n, s, c)
    mask = torch.zeros(n, s, dtype=torch.bool)
    mask[1, -4:] = True
    ref2 = grid_reference_points(single, shapes_t, n, torch.float64)
    save_module_case(
        "module_msda_ref2", attn_mod,
        dict(query=query, reference_points=ref2, input_flatten=feat, spatial_shapes=shapes_t,
             level_start_index=lsi, padding_mask=mask),
        lambda m_, t: m_(t["query"], t["reference_points"], t["input_flatten"], t["spatial_shapes"],
                        t["level_start_index"], t["padding_mask"]),
        wrt=["query", "reference_points", "input_flatten"])
    # 4-d reference boxes (decoder with box refinement, ms_deform_attn.py:108-110), Lq != S
    torch.manual_seed(23)
    q7 = torch.randn(n, 7, c)
    ref4 = torch.rand(n, 7, 2, 4) * 0.6 + 0.2
    save_module_case(
        "module_msda_ref4", attn_mod,
        dict(query=q7, reference_points=ref4, input_flatten=feat, spatial_shapes=shapes_t,
             level_start_index=lsi),
        lambda m_, t: m_(t["query"], t["reference_points"], t["input_flatten"], t["spatial_shapes"],
                        t["level_start_index"], None),
        wrt=["query", "reference_points", "input_flatten"])

    # head width 16 (d_model 64 / 4 heads): the shapes the FUSED sm_100a layer kernels cover
    torch.manual_seed(24)
    c16 = 64
    attn16 = mod.MSDeformAttn(c16, 2, heads, 4).double()
    perturb(attn16, 25)
    query16 = torch.randn(n, s, c16)
    feat16 = torch.randn(n, s, c16)
    save_module_case(
        "module_msda_ref2_d16", attn16,
        dict(query=query16, reference_points=ref2, input_flatten=feat16, spatial_shapes=shapes_t,
             level_start_index=lsi, padding_mask=mask),
        lambda m_, t: m_(t["query"], t["reference_points"], t["input_flatten"], t["spatial_shapes"],
                        t["level_start_index"], t["padding_mask"]),
        wrt=["query", "reference_points", "input_flatten"])
    q7_16 = torch.randn(n, 7, c16)
    save_module_case(
        "module_msda_ref4_d16", attn16,
        dict(query=q7_16, reference_points=ref4, input_flatten=feat16, spatial_shapes=shapes_t,
             level_start_index=lsi),
        lambda m_, t: m_(t["query"], t["reference_points"], t["input_flatten"], t["spatial_shapes"],
                        t["level_start_index"], None),
        wrt=["query", "reference_points", "input_flatten"])

    # ---------------- layer classes (deformable_transformer_single.py) --------
    torch.manual_seed(31)
    pos = torch.randn(n, s, c)
    enc_layer = single.DeformableTransformerEncoderLayer(c, 64, 0.0, "relu", 2, heads, pts).double()
    perturb(enc_layer, 32)
    save_module_case(
        "layer_encoder", enc_layer,
        dict(src=feat, pos=pos, reference_points=ref2, spatial_shapes=shapes_t, level_start_index=lsi,
             padding_mask=mask),
        lambda m_, t: m_(t["src"], t["pos"], t["reference_points"], t["spatial_shapes"],
                        t["level_start_index"], t["padding_mask"]),
        wrt=["src", "pos"])

    # Encoder Cross Fusion layer (:406-461): RGB queries over a depth pyramid of its own shapes
    dshapes = torch.as_tensor([(4, 6)], dtype=torch.long)
    dlsi = lsi_of(dshapes)
    sd = int(dshapes.prod(1).sum())
    torch.manual_seed(33)
    depth = torch.randn(n, sd, c)
    dmask = torch.zeros(n, sd, dtype=torch.bool)
    dmask[0, -3:] = True
    ref_d1 = ref2[:, :, :1].contiguous()             # one depth level
    fusion = single.DeformableTransformerFusionLayerV2(c, 64, 0.0, "gelu", 1, heads, pts).double()
    perturb(fusion, 34)
    save_module_case(
        "layer_fusion_v2", fusion,
        dict(tgt=feat, query_pos=pos, reference_points=ref_d1, src=depth, src_spatial_shapes=dshapes,
             level_start_index=dlsi, src_padding_mask=dmask),
        lambda m_, t: m_(t["tgt"], t["query_pos"], t["reference_points"], t["src"], t["src_spatial_shapes"],
                        t["level_start_index"], t["src_padding_mask"]),
        wrt=["tgt", "query_pos", "src"])

    fusion16 = single.DeformableTransformerFusionLayerV2(c16, 64, 0.0, "gelu", 1, heads, 4).double()
    perturb(fusion16, 44)
    torch.manual_seed(45)
    pos16 = torch.randn(n, s, c16)
    depth16 = torch.randn(n, sd, c16)
    save_module_case(
        "layer_fusion_v2_d16", fusion16,
        dict(tgt=feat16, query_pos=pos16, reference_points=ref_d1, src=depth16, src_spatial_shapes=dshapes,
             level_start_index=dlsi, src_padding_mask=dmask),
        lambda m_, t: m_(t["tgt"], t["query_pos"], t["reference_points"], t["src"], t["src_spatial_shapes"],
                        t["level_start_index"], t["src_padding_mask"]),
        wrt=["tgt", "query_pos", "src"])

    # Late Fusion layer (:341-402)
    late = single.DepthDeformableTransformerEncoderLayer(c, 64, 0.0, "relu", 1, heads, pts, True, True, True).double()
    perturb(late, 35)
    save_module_case(
        "layer_late_fusion", late,
        dict(tgt=feat, query_pos=pos, reference_points=ref_d1, src=depth, src_spatial_shapes=dshapes,
             frame_start_index=dlsi, src_padding_mask=dmask),
        lambda m_, t: m_(t["tgt"], t["query_pos"], None, None, t["reference_points"], None, t["src"],
                        t["src_spatial_shapes"], t["frame_start_index"], None, t["src_padding_mask"]),
        wrt=["tgt", "query_pos", "src"])

    # decoder layer (:596-648), 4-d refs
    torch.manual_seed(36)
    tgt7 = torch.randn(n, 7, c)
    qpos7 = torch.randn(n, 7, c)
    dec = single.DeformableTransformerDecoderLayer(c, 64, 0.0, "relu", 2, heads, pts).double()
    perturb(dec, 37)
    save_module_case(
        "layer_decoder", dec,
        dict(tgt=tgt7, query_pos=qpos7, reference_points=ref4, src=feat, src_spatial_shapes=shapes_t,
             level_start_index=lsi, src_padding_mask=mask),
        lambda m_, t: m_(t["tgt"], t["query_pos"], t["reference_points"], t["src"], t["src_spatial_shapes"],
                        t["level_start_index"], t["src_padding_mask"]),
        wrt=["tgt", "src"])

    # Encoder-CF encoder (:465-518): 3 encoder layers, 2 fusion layers after layers 0 and 1.
    # One RGB level so that the fusion output (RGB length) can serve as the next fusion's
    # value with the depth shapes -- the only way the shipped model runs.
    one = torch.as_tensor([(4, 6)], dtype=torch.long)
    one_lsi = lsi_of(one)
    torch.manual_seed(38)
    src1 = torch.randn(n, sd, c)
    pos1 = torch.randn(n, sd, c)
    enc1 = single.DeformableTransformerEncoderLayer(c, 64, 0.0, "relu", 1, heads, pts).double()
    fus1 = single.DeformableTransformerFusionLayerV2(c, 64, 0.0, "gelu", 1, heads, pts).double()
    rgbd = single.RGBDDeformableTransformerEncoderV2(enc1, fus1, 3, 2, 2, [0, 1]).double()
    perturb(rgbd, 39)
    vr = torch.ones(n, 1, 2)
    vr[1, 0, 0] = 0.75
    save_module_case(
        "encoder_rgbd_v2", rgbd,
        dict(src=src1, spatial_shapes=one, level_start_index=one_lsi, valid_ratios=vr, pos=pos1,
             padding_mask=dmask, depth_src=depth, depth_spatial_shapes=dshapes, depth_level_start_index=dlsi),
        lambda m_, t: m_(t["src"], t["spatial_shapes"], t["level_start_index"], t["valid_ratios"], t["pos"],
                        t["padding_mask"], None, t["depth_src"], t["depth_spatial_shapes"],
                        t["depth_level_start_index"], None, None, None),
        wrt=["src", "depth_src"])

    # plain encoder (:551-593), 2 levels, 2 layers
    enc2 = single.DeformableTransformerEncoder(
        single.DeformableTransformerEncoderLayer(c, 64, 0.0, "relu", 2, heads, pts), 2).double()
    perturb(enc2, 40)
    vr2 = torch.ones(n, 2, 2)
    vr2[0, :, 1] = 0.8
    save_module_case(
        "encoder_plain", enc2,
        dict(src=feat, spatial_shapes=shapes_t, level_start_index=lsi, valid_ratios=vr2, pos=pos,
             padding_mask=mask),
        lambda m_, t: m_(t["src"], t["spatial_shapes"], t["level_start_index"], t["valid_ratios"], t["pos"],
                        t["padding_mask"]),
        wrt=["src"])

    # whole single-frame transformer (deformable_transformer_single.py:23-337), three variants
    def transformer_case(name, depth_type, use_depth, levels, seed):
        torch.manual_seed(seed)
        model = single.DeformableTransformer(
            d_model=c, nhead=heads, num_encoder_layers=5 if "encoder_cf" in depth_type else 2,
            num_decoder_layers=2, dim_feedforward=64, dropout=0.0, activation="relu",
            return_intermediate_dec=True, num_feature_levels=len(levels), dec_n_points=pts, enc_n_points=pts,
            use_depth=use_depth, depth_type=depth_type, dpth_n_points=pts).double()
        perturb(model, seed + 1)
        ins = {}
        for i, (h, w) in enumerate(levels):
            ins[f"src{i}"] = torch.randn(n, c, h, w)
            ins[f"pos{i}"] = torch.randn(n, c, h, w)
            mk = torch.zeros(n, h, w, dtype=torch.bool)
            # valid width = a power of two: torch's CUDA division by a python scalar multiplies by the
            # reciprocal, so only then is valid_ratio = valid_W / W bit-identical on CPU and GPU
            mk[1, :, 1 << ((w - 1).bit_length() - 1):] = True
            ins[f"mask{i}"] = mk
        ins["depth_src0"] = torch.randn(n, c, levels[0][0], levels[0][1])
        ins["depth_pos0"] = torch.randn(n, c, levels[0][0], levels[0][1])
        ins["depth_mask0"] = ins["mask0"].clone()
        ins["query_embed"] = torch.randn(5, 2 * c)
        nl_ = len(levels)

        def call(m_, t):
            hs, init_ref, inter_ref, _, _ = m_(
                [t[f"src{i}"] for i in range(nl_)], [t[f"mask{i}"] for i in range(nl_)],
                [t[f"pos{i}"] for i in range(nl_)], [t["depth_src0"]], [t["depth_mask0"]], [t["depth_pos0"]],
                t["query_embed"])
            return torch.cat([hs.flatten(), init_ref.flatten(), inter_ref.flatten()])
        save_module_case(name, model, ins, call, wrt=["src0", "depth_src0", "query_embed"])

    # RGB-D features as the encoder's query ("concat" depth types: rgbd_query replaces src + pos as the query of every
    # encoder layer, single.py:185-189, :556-559)
    def rgbd_query_case(name, seed, levels=((4, 6), (2, 3))):
        if not wanted(name):
            return
        torch.manual_seed(seed)
        model = single.DeformableTransformer(
            d_model=c, nhead=heads, num_encoder_layers=2, num_decoder_layers=1, dim_feedforward=64, dropout=0.0,
            activation="relu", return_intermediate_dec=True, num_feature_levels=len(levels), dec_n_points=pts,
            enc_n_points=pts, use_depth=False, depth_type="Baseline_concat").double()
        perturb(model, seed + 1)
        ins = {"query_embed": torch.randn(5, 2 * c)}
        for i, (h, w) in enumerate(levels):
            ins[f"src{i}"] = torch.randn(n, c, h, w)
            ins[f"pos{i}"] = torch.randn(n, c, h, w)
            ins[f"rgbd{i}"] = torch.randn(n, c, h, w)
            ins[f"mask{i}"] = torch.zeros(n, h, w, dtype=torch.bool)
        nl_ = len(levels)

        def call(m_, t):
            hs, init_ref, inter_ref, _, _ = m_(
                [t[f"src{i}"] for i in range(nl_)], [t[f"mask{i}"] for i in range(nl_)],
                [t[f"pos{i}"] for i in range(nl_)], None, None, None, t["query_embed"],
                [t[f"rgbd{i}"] for i in range(nl_)])
            return torch.cat([hs.flatten(), init_ref.flatten(), inter_ref.flatten()])
        save_module_case(name, model, ins, call, wrt=["src0", "rgbd0", "query_embed"])

    rgbd_query_case("transformer_rgbd_query", 56)
    transformer_case("transformer_baseline", "Baseline_rgb", False, [(6, 5), (3, 3)], 50)
    transformer_case("transformer_latefusion", "DepthDeform_latefusion_dformer", True, [(4, 6)], 52)
    transformer_case("transformer_encoder_cf", "DepthDeform_encoder_cf_dformer", True, [(4, 6)], 54)

    # ---------------- head width 16 at d_model 128 / 8 heads / 4 points: the narrowest layers for which
    # BOTH the fused deformable-attention kernels and the fused layer-epilogue kernels (residual +
    # LayerNorm [+ activation] [+ next query], fp32: 128 channels) run in the fp32 tests ------------
    cw, hw_, pw, ffw = 128, 8, 4, 128
    torch.manual_seed(61)
    feat_w = torch.randn(n, s, cw)
    pos_w = torch.randn(n, s, cw)
    enc_w = single.DeformableTransformerEncoderLayer(cw, ffw, 0.0, "relu", 2, hw_, pw).double()
    perturb(enc_w, 62)
    save_module_case(
        "layer_encoder_c128", enc_w,
        dict(src=feat_w, pos=pos_w, reference_points=ref2, spatial_shapes=shapes_t, level_start_index=lsi,
             padding_mask=mask),
        lambda m_, t: m_(t["src"], t["pos"], t["reference_points"], t["spatial_shapes"],
                        t["level_start_index"], t["padding_mask"]),
        wrt=["src", "pos"], store=np.float32, grad_limit=20000)
    enc_stack_w = single.DeformableTransformerEncoder(
        single.DeformableTransformerEncoderLayer(cw, ffw, 0.0, "relu", 2, hw_, pw), 2).double()
    perturb(enc_stack_w, 63)
    save_module_case(
        "encoder_plain_c128", enc_stack_w,
        dict(src=feat_w, spatial_shapes=shapes_t, level_start_index=lsi, valid_ratios=vr2, pos=pos_w,
             padding_mask=mask),
        lambda m_, t: m_(t["src"], t["spatial_shapes"], t["level_start_index"], t["valid_ratios"], t["pos"],
                        t["padding_mask"]),
        wrt=["src"], store=np.float32, grad_limit=20000)
    torch.manual_seed(64)
    depth_w = torch.randn(n, sd, cw)
    fusion_w = single.DeformableTransformerFusionLayerV2(cw, ffw, 0.0, "gelu", 1, hw_, pw).double()
    perturb(fusion_w, 65)
    save_module_case(
        "layer_fusion_v2_c128", fusion_w,
        dict(tgt=feat_w, query_pos=pos_w, reference_points=ref_d1, src=depth_w, src_spatial_shapes=dshapes,
             level_start_index=dlsi, src_padding_mask=dmask),
        lambda m_, t: m_(t["tgt"], t["query_pos"], t["reference_points"], t["src"], t["src_spatial_shapes"],
                        t["level_start_index"], t["src_padding_mask"]),
        wrt=["tgt", "query_pos", "src"], store=np.float32, grad_limit=20000)
    torch.manual_seed(66)
    tgt7_w = torch.randn(n, 7, cw)
    qpos7_w = torch.randn(n, 7, cw)
    dec_stack_w = single.DeformableTransformerDecoder(
        single.DeformableTransformerDecoderLayer(cw, ffw, 0.0, "relu", 2, hw_, pw), 2, return_intermediate=True).double()
    perturb(dec_stack_w, 67)
    ref_q2 = torch.rand(n, 7, 2) * 0.8 + 0.1

    def dec_call(m_, t):
        hs, refs = m_(t["tgt"], t["reference_points"], t["src"], t["src_spatial_shapes"], t["level_start_index"],
                      t["valid_ratios"], t["query_pos"], t["src_padding_mask"])
        return torch.cat([hs.flatten(), refs.flatten()])
    save_module_case(
        "decoder_c128", dec_stack_w,
        dict(tgt=tgt7_w, reference_points=ref_q2, src=feat_w, src_spatial_shapes=shapes_t, level_start_index=lsi,
             valid_ratios=vr2, query_pos=qpos7_w, src_padding_mask=mask),
        dec_call, wrt=["tgt", "src", "query_pos"], store=np.float32, grad_limit=20000)
    udf_w = backbone_cf.DepthDeformableTransformerEncoderLayer(cw, ffw, 0.0, "relu", 1, hw_, pw).double()
    perturb(udf_w, 68)
    torch.manual_seed(69)
    maps_w = dict(src=torch.randn(n, cw, 3, 4), target=torch.randn(n, cw, 6, 8), pos_src=torch.randn(n, cw, 3, 4),
                  pos_target=torch.randn(n, cw, 6, 8))
    m_rgb_w = torch.zeros(n, 3, 4, dtype=torch.bool)
    m_dep_w = torch.zeros(n, 6, 8, dtype=torch.bool)
    m_dep_w[1, :, -2:] = True
    m_rgb_w[1, :, -1:] = True
    save_module_case(
        "backbone_udf_fuse_c128", udf_w, dict(maps_w, mask_src=m_rgb_w, mask_target=m_dep_w),
        lambda m_, t: backbone_cf.FusionBackboneBase.fuse_layers(
            t["src"], t["target"], t["pos_src"], t["pos_target"], t["mask_src"], t["mask_target"], m_),
        wrt=["src", "target"], store=np.float32, grad_limit=20000)
    # production width (d_model 256, 8 heads of 32, 4 points): the bf16 kernels' shape; stored as fp32
    torch.manual_seed(71)
    feat_p = torch.randn(n, s, 256)
    pos_p = torch.randn(n, s, 256)
    enc_p = single.DeformableTransformerEncoderLayer(256, 256, 0.0, "relu", 2, 8, 4).double()
    perturb(enc_p, 72, std=0.03)
    save_module_case(
        "layer_encoder_c256", enc_p,
        dict(src=feat_p, pos=pos_p, reference_points=ref2, spatial_shapes=shapes_t, level_start_index=lsi,
             padding_mask=mask),
        lambda m_, t: m_(t["src"], t["pos"], t["reference_points"], t["spatial_shapes"],
                        t["level_start_index"], t["padding_mask"]),
        wrt=["src", "pos"], store=np.float32, grad_limit=20000)

    # BASELINE.json configs[2] / [3] at the production width: Encoder Cross Fusion V2 (:406-461) with a one-level and a
    # two-level depth pyramid, and the Late Fusion layer (:341-402); d_model 256, 8 heads of 32, 4 points, d_ffn 256
    torch.manual_seed(73)
    depth_p = torch.randn(n, sd, 256)
    fusion_p = single.DeformableTransformerFusionLayerV2(256, 256, 0.0, "gelu", 1, 8, 4).double()
    perturb(fusion_p, 74, std=0.03)
    save_module_case(
        "layer_fusion_v2_c256", fusion_p,
        dict(tgt=feat_p, query_pos=pos_p, reference_points=ref_d1, src=depth_p, src_spatial_shapes=dshapes,
             level_start_index=dlsi, src_padding_mask=dmask),
        lambda m_, t: m_(t["tgt"], t["query_pos"], t["reference_points"], t["src"], t["src_spatial_shapes"],
                        t["level_start_index"], t["src_padding_mask"]),
        wrt=["tgt", "query_pos", "src"], store=np.float32, grad_limit=20000)
    torch.manual_seed(75)
    depth_p2 = torch.randn(n, s, 256)              # a depth pyramid with the RGB shapes
    fusion_p2 = single.DeformableTransformerFusionLayerV2(256, 256, 0.0, "gelu", 2, 8, 4).double()
    perturb(fusion_p2, 76, std=0.03)
    save_module_case(
        "layer_fusion_v2_c256_l2", fusion_p2,
        dict(tgt=feat_p, query_pos=pos_p, reference_points=ref2, src=depth_p2, src_spatial_shapes=shapes_t,
             level_start_index=lsi, src_padding_mask=mask),
        lambda m_, t: m_(t["tgt"], t["query_pos"], t["reference_points"], t["src"], t["src_spatial_shapes"],
                        t["level_start_index"], t["src_padding_mask"]),
        wrt=["tgt", "query_pos", "src"], store=np.float32, grad_limit=20000)
    late_p = single.DepthDeformableTransformerEncoderLayer(256, 256, 0.0, "relu", 1, 8, 4, True, True, True).double()
    perturb(late_p, 77, std=0.03)
    save_module_case(
        "layer_late_fusion_c256", late_p,
        dict(tgt=feat_p, query_pos=pos_p, reference_points=ref_d1, src=depth_p, src_spatial_shapes=dshapes,
             frame_start_index=dlsi, src_padding_mask=dmask),
        lambda m_, t: m_(t["tgt"], t["query_pos"], None, None, t["reference_points"], None, t["src"],
                        t["src_spatial_shapes"], t["frame_start_index"], None, t["src_padding_mask"]),
        wrt=["tgt", "query_pos", "src"], store=np.float32, grad_limit=20000)

    # ---------------- temporal layers ----------------------------------------------------------------
    # frames-as-levels layer (deformable_transformer_single.py:650-700): 3 reference frames of (4,5)
    fshapes = torch.as_tensor([(4, 5)] * 3, dtype=torch.long)
    flsi = lsi_of(fshapes)
    fs = int(fshapes.prod(1).sum())
    torch.manual_seed(81)
    temporal = single.TemporalDeformableTransformerEncoderLayer(c, 64, 0.0, "relu", 3, heads, pts).double()
    perturb(temporal, 82)
    tq = torch.randn(n, 20, c)
    tqpos = torch.randn(n, 20, c)
    tmem = torch.randn(n, fs, c)
    tref = torch.rand(n, 20, 3, 2) * 0.9 + 0.05
    tmask = torch.zeros(n, fs, dtype=torch.bool)
    tmask[0, -5:] = True
    save_module_case(
        "layer_temporal_encoder", temporal,
        dict(tgt=tq, query_pos=tqpos, reference_points=tref, src=tmem, src_spatial_shapes=fshapes,
             frame_start_index=flsi, src_padding_mask=tmask),
        lambda m_, t: m_(t["tgt"], t["query_pos"], t["reference_points"], t["src"], t["src_spatial_shapes"],
                        t["frame_start_index"], t["src_padding_mask"]),
        wrt=["tgt", "query_pos", "src"])
    # TransVOD++ temporal decoder TDTD (deformable_transformer_multi_plusplus.py:1030-1076) over the
    # current frame's one-level memory, called with valid_ratios[:, 0:1]
    pp = import_transvod_plusplus()
    pp.MSDeformAttn = mod.MSDeformAttn
    torch.manual_seed(83)
    tdtd = pp.TemporalDeformableTransformerDecoder(
        pp.DeformableTransformerDecoderLayer(c, 64, 0.0, "relu", 1, heads, pts), 2, False).double()
    perturb(tdtd, 84)
    mem1 = torch.randn(1, sd, c)
    tgt_t = torch.randn(1, 9, c)
    ref_t = torch.rand(1, 9, 2) * 0.8 + 0.1
    vr_t = torch.ones(1, 1, 2)
    vr_t[0, 0, 0] = 0.5

    def tdtd_call(m_, t):
        hs, refs = m_(t["tgt"], t["reference_points"], t["src"], t["src_spatial_shapes"], t["level_start_index"],
                      t["valid_ratios"], None, None)
        return torch.cat([hs.flatten(), refs.flatten()])
    save_module_case(
        "temporal_decoder_pp", tdtd,
        dict(tgt=tgt_t, reference_points=ref_t, src=mem1, src_spatial_shapes=dshapes, level_start_index=dlsi,
             valid_ratios=vr_t),
        tdtd_call, wrt=["tgt", "src"])

    # two-stage variant of the single-frame transformer (single.py:82-86, :112-153, :308-322): encoder proposals,
    # top-k, proposal position embedding.  get_proposal_pos_embed hard-codes 128 features per coordinate, so the
    # model must be 256 wide; parameters are made fp16-representable and the state is stored as fp16.
    def two_stage_case(name, seed):
        if not wanted(name):
            return
        torch.manual_seed(seed)
        wide, levels, k = 256, [(5, 6), (3, 3)], 6
        model = single.DeformableTransformer(
            d_model=wide, nhead=8, num_encoder_layers=1, num_decoder_layers=1, dim_feedforward=64, dropout=0.0,
            activation="relu", return_intermediate_dec=True, num_feature_levels=len(levels), dec_n_points=2,
            enc_n_points=2, two_stage=True, two_stage_num_proposals=k).double()
        model.decoder.class_embed = torch.nn.ModuleList(torch.nn.Linear(wide, 3) for _ in range(2)).double()
        model.decoder.bbox_embed = torch.nn.ModuleList(
            torch.nn.Sequential(torch.nn.Linear(wide, 32), torch.nn.ReLU(), torch.nn.Linear(32, 4)) for _ in range(2)).double()
        perturb(model, seed + 1, std=0.03)
        ins = {}
        for i, (h, w) in enumerate(levels):
            ins[f"src{i}"] = torch.randn(n, wide, h, w)
            ins[f"pos{i}"] = torch.randn(n, wide, h, w)
            mk = torch.zeros(n, h, w, dtype=torch.bool)
            mk[1, :, 1 << ((w - 1).bit_length() - 1):] = True
            ins[f"mask{i}"] = mk

        def call(m_, t):
            hs, init_ref, inter_ref, enc_cls, enc_coord = m_(
                [t["src0"], t["src1"]], [t["mask0"], t["mask1"]], [t["pos0"], t["pos1"]], None, None, None, None)
            enc_coord = torch.where(torch.isinf(enc_coord), torch.zeros_like(enc_coord), enc_coord)
            return torch.cat([hs.flatten(), init_ref.flatten(), inter_ref.flatten(), enc_cls.flatten(), enc_coord.flatten()])
        save_module_case(name, model, ins, call, wrt=["src0"], store=np.float32, grad_limit=3000, state_store=np.float16)

    two_stage_case("transformer_two_stage", 96)

    # ---------------- TransVOD++ multi-frame transformer with its temporal query stage -----------------
    # (deformable_transformer_multi_plusplus.py:70-603): per-frame encoder/decoder with box refinement, RoIAlign of
    # every decoder box (mmcv stand-in above), QRF head, three TQE + TDTD rounds.  80 queries, because the first
    # round takes the top 80 * num_ref_frames of the reference-frame queries (:531).
    # "transvodpp_f1": the UNMODIFIED reference (one reference frame: its [1, F, 2] valid-ratio expansion (:425) is
    #                  then consistent with the one-level memory).
    # "transvodpp_f2": two reference frames, for which the reference's temporal decoder call is mis-shaped
    #                  (the oracle raises, the CUDA op mis-indexes); the temporal decoder is given
    #                  valid_ratios[:, 0:1] instead -- the one documented deviation of this case.
    class MLP(torch.nn.Module):                      # deformable_detr_multi_plusplus.py:585-597
        def __init__(self, input_dim, hidden_dim, output_dim, num_layers):
            super().__init__()
            self.num_layers = num_layers
            hdim = [hidden_dim] * (num_layers - 1)
            self.layers = torch.nn.ModuleList(torch.nn.Linear(a, b) for a, b in zip([input_dim] + hdim, hdim + [output_dim]))

        def forward(self, x):
            for i, layer in enumerate(self.layers):
                x = torch.relu(layer(x)) if i < self.num_layers - 1 else layer(x)
            return x

    class TransVODPP(torch.nn.Module):               # the detector's wiring, deformable_detr_multi_plusplus.py:78-81,183-194,320
        def __init__(self, transformer, width, n_dec, n_cls=3):
            super().__init__()
            self.transformer = transformer
            self.class_embed = torch.nn.ModuleList(torch.nn.Linear(width, n_cls) for _ in range(n_dec))
            self.bbox_embed = torch.nn.ModuleList(MLP(width, width, 4, 3) for _ in range(n_dec))
            self.temp_class_embed_list = torch.nn.ModuleList(torch.nn.Linear(width, n_cls) for _ in range(3))
            self.temp_bbox_embed_list = torch.nn.ModuleList(MLP(width, width, 4, 3) for _ in range(3))
            self.transformer.decoder.bbox_embed = self.bbox_embed          # with_box_refine (:194)

        def forward(self, t, n_levels=1):
            hs, init_ref, inter_ref, _, _, final_hs, final_ref, out = self.transformer(
                [t["src0"]], [t["mask0"]], [t["pos0"]], [t["depth_src0"]], [t["depth_mask0"]], [t["depth_pos0"]],
                t["imgs_whwh"], t["query_embed"], self.class_embed[-1], self.bbox_embed[-1],
                self.temp_class_embed_list, self.temp_bbox_embed_list)
            parts = [hs, init_ref, inter_ref, final_hs, final_ref]
            for aux in out["aux_outputs"]:
                parts += [aux["pred_logits"], aux["pred_boxes"]]
            return torch.cat([p_.flatten() for p_ in parts])

    def transvodpp_case(name, depth_type, use_depth, ref_frames, seed, fix_valid_ratios):
        if not wanted(name):
            return
        torch.manual_seed(seed)
        nq, (fh, fw) = 80, (4, 6)
        c = 16                                       # narrow: three QRF heads are 40 k parameters each even so
        model = TransVODPP(pp.DeformableTransformer(
            d_model=c, nhead=heads, num_encoder_layers=2, num_decoder_layers=2, dim_feedforward=64, dropout=0.0,
            activation="relu", return_intermediate_dec=True, num_feature_levels=1, dec_n_points=pts, enc_n_points=pts,
            num_query=nq, n_temporal_decoder_layers=1, num_ref_frames=ref_frames, use_depth=use_depth,
            depth_type=depth_type, dpth_n_points=pts), c, 2).double()
        perturb(model, seed + 1)
        frames = ref_frames + 1
        ins = dict(src0=torch.randn(frames, c, fh, fw), pos0=torch.randn(frames, c, fh, fw),
                   depth_src0=torch.randn(frames, c, fh, fw), depth_pos0=torch.randn(frames, c, fh, fw),
                   query_embed=torch.randn(nq, 2 * c))
        mk = torch.zeros(frames, fh, fw, dtype=torch.bool)
        mk[-1, :, 4:] = True                         # a padded reference frame (valid width 4: a power of two)
        ins["mask0"] = mk
        ins["depth_mask0"] = mk.clone()
        ins["imgs_whwh"] = torch.tensor([[fw * 32, fh * 32, fw * 32, fh * 32]])
        original = pp.TemporalDeformableTransformerDecoder.forward
        if fix_valid_ratios:
            def fixed(self, tgt, reference_points, src, shapes_, lsi_, valid_ratios, query_pos=None, mask=None):
                return original(self, tgt, reference_points, src, shapes_, lsi_, valid_ratios[:, 0:1], query_pos, mask)
            pp.TemporalDeformableTransformerDecoder.forward = fixed
        try:
            save_module_case(name, model, ins, lambda m_, t: m_(t), wrt=["src0", "depth_src0", "query_embed"],
                             grad_limit=6000)
        finally:
            pp.TemporalDeformableTransformerDecoder.forward = original

    transvodpp_case("transvodpp_f1", "Baseline_rgb", False, 1, 90, False)
    transvodpp_case("transvodpp_f2_latefusion", "DepthDeform_latefusion_dformer", True, 2, 92, True)

    # ---------------- TransVOD multi-frame transformer (deformable_transformer_multi.py:24-379) -------------------------
    # "transvod_f1": the UNMODIFIED reference class, one reference frame.  "transvod_f2_tdam": two reference frames,
    # with the frames-as-levels temporal encoder layer switched on (``TDAM = True`` set on the instance; the reference
    # hard-codes False, :46) and the temporal decoder given valid_ratios[:, 0:1] (as for TransVOD++).
    tv = importlib.import_module("models.deformable_transformer_multi")
    tv.MSDeformAttn = mod.MSDeformAttn

    class TransVOD(torch.nn.Module):
        def __init__(self, transformer, width, n_cls=4):
            super().__init__()
            self.transformer = transformer
            self.class_embed = torch.nn.Linear(width, n_cls)

        def forward(self, t):
            hs, init_ref, inter_ref, _, _, final_hs, final_ref = self.transformer(
                [t["src0"]], [t["mask0"]], [t["pos0"]], [t["depth_src0"]], [t["depth_mask0"]], [t["depth_pos0"]],
                t["query_embed"], self.class_embed)
            return torch.cat([x.flatten() for x in (hs, init_ref, inter_ref, final_hs, final_ref)])

    def transvod_case(name, depth_type, use_depth, ref_frames, seed, tdam):
        if not wanted(name):
            return
        torch.manual_seed(seed)
        nq, (fh, fw), cw_ = 80, (4, 6), 16
        model = TransVOD(tv.DeformableTransformer(
            d_model=cw_, nhead=heads, num_encoder_layers=2, num_decoder_layers=2, dim_feedforward=64, dropout=0.0,
            activation="relu", return_intermediate_dec=True, num_feature_levels=1, dec_n_points=pts, enc_n_points=pts,
            n_temporal_decoder_layers=1, num_ref_frames=ref_frames, use_depth=use_depth, depth_type=depth_type,
            dpth_n_points=pts), cw_).double()
        model.transformer.TDAM = tdam
        perturb(model, seed + 1)
        frames = ref_frames + 1
        ins = dict(src0=torch.randn(frames, cw_, fh, fw), pos0=torch.randn(frames, cw_, fh, fw),
                   depth_src0=torch.randn(frames, cw_, fh, fw), depth_pos0=torch.randn(frames, cw_, fh, fw),
                   query_embed=torch.randn(nq, 2 * cw_))
        mk = torch.zeros(frames, fh, fw, dtype=torch.bool)
        mk[-1, :, 4:] = True
        ins["mask0"] = mk
        ins["depth_mask0"] = mk.clone()
        original = tv.TemporalDeformableTransformerDecoder.forward
        if ref_frames > 1:
            def fixed(self, tgt, reference_points, src, shapes_, lsi_, valid_ratios, query_pos=None, mask=None):
                return original(self, tgt, reference_points, src, shapes_, lsi_, valid_ratios[:, 0:1], query_pos, mask)
            tv.TemporalDeformableTransformerDecoder.forward = fixed
        try:
            save_module_case(name, model, ins, lambda m_, t: m_(t), wrt=["src0", "depth_src0", "query_embed"])
        finally:
            tv.TemporalDeformableTransformerDecoder.forward = original

    transvod_case("transvod_f1", "Baseline_rgb", False, 1, 100, False)
    transvod_case("transvod_f2_tdam", "DepthDeform_latefusion_dformer", True, 2, 102, True)

    # Backbone Cross Fusion U-DF: fuse_layers + its layer (dformer_crossfusion_backbone.py:387-428,120-181)
    torch.manual_seed(41)
    udf = backbone_cf.DepthDeformableTransformerEncoderLayer(c, 64, 0.0, "relu", 1, heads, pts).double()
    perturb(udf, 42)
    rgb_map = torch.randn(n, c, 3, 4)
    dep_map = torch.randn(n, c, 6, 8)
    pos_rgb = torch.randn(n, c, 3, 4)
    pos_dep = torch.randn(n, c, 6, 8)
    m_rgb = torch.zeros(n, 3, 4, dtype=torch.bool)
    m_dep = torch.zeros(n, 6, 8, dtype=torch.bool)
    m_dep[1, :, -2:] = True
    m_rgb[1, :, -1:] = True
    save_module_case(
        "backbone_udf_fuse", udf,
        dict(src=rgb_map, target=dep_map, pos_src=pos_rgb, pos_target=pos_dep, mask_src=m_rgb, mask_target=m_dep),
        lambda m_, t: backbone_cf.FusionBackboneBase.fuse_layers(
            t["src"], t["target"], t["pos_src"], t["pos_target"], t["mask_src"], t["mask_target"], m_),
        wrt=["src", "target"])

    # Sine position embedding (position_encoding.py:20-56) on padded masks, the shipped configuration
    # (N_steps = hidden_dim // 2, normalize=True: position_encoding.py:88-91) and the un-normalised default
    if wanted("position_sine"):
        pe = importlib.import_module("models.position_encoding")
        misc = sys.modules["util.misc"]
        pmasks = []
        for (h, w), (ph, pw) in zip([(7, 9), (4, 5), (2, 3)], [(2, 3), (1, 2), (1, 1)]):
            pm = torch.zeros(2, h, w, dtype=torch.bool)
            pm[1, h - ph:, :] = True
            pm[1, :, w - pw:] = True
            pmasks.append(pm)
        store = {}
        for tag, module in (("norm", pe.PositionEmbeddingSine(16, normalize=True)),
                            ("raw", pe.PositionEmbeddingSine(16, temperature=20, normalize=False))):
            for lvl, pm in enumerate(pmasks):
                feat = torch.zeros(2, 1, *pm.shape[1:], dtype=torch.float32)
                store[f"{tag}_pos{lvl}"] = module(misc.NestedTensor(feat, pm)).numpy()
        for lvl, pm in enumerate(pmasks):
            store[f"mask{lvl}"] = pm.numpy()
        torch.manual_seed(43)
        store["level_embed"] = torch.randn(3, 32, dtype=torch.float32).numpy()
        np.savez_compressed(os.path.join(OUT, "position_sine.npz"), **store)
        print("wrote position_sine")


if __name__ == "__main__":
    main()
