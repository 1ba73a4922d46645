"""How each golden module/layer case (oracle/gen_golden.py) is rebuilt from this repo's classes.
Shared by the CPU host-logic tests (oracle patched in) and the GPU tests (real kernels)."""
import torch

from dfvod_b200 import backbone_fusion, transformer_layers as tl
from dfvod_b200.deformable_transformer import DeformableTransformer
from dfvod_b200.ops.modules import MSDeformAttn

C, HEADS, PTS = 32, 4, 3


def _enc(levels):
    return tl.DeformableTransformerEncoderLayer(C, 64, 0.0, "relu", levels, HEADS, PTS)


CASES = {
    "module_msda_ref2": dict(
        build=lambda: MSDeformAttn(C, 2, HEADS, PTS),
        call=lambda m, t: m(t["query"], t["reference_points"], t["input_flatten"], t["spatial_shapes"],
                            t["level_start_index"], t["padding_mask"]),
        wrt=["query", "reference_points", "input_flatten"]),
    "module_msda_ref4": dict(
        build=lambda: MSDeformAttn(C, 2, HEADS, PTS),
        call=lambda m, t: m(t["query"], t["reference_points"], t["input_flatten"], t["spatial_shapes"],
                            t["level_start_index"], None),
        wrt=["query", "reference_points", "input_flatten"]),
    "module_msda_ref2_d16": dict(
        build=lambda: MSDeformAttn(64, 2, HEADS, 4),
        call=lambda m, t: m(t["query"], t["reference_points"], t["input_flatten"], t["spatial_shapes"],
                            t["level_start_index"], t["padding_mask"]),
        wrt=["query", "reference_points", "input_flatten"]),
    "module_msda_ref4_d16": dict(
        build=lambda: MSDeformAttn(64, 2, HEADS, 4),
        call=lambda m, t: m(t["query"], t["reference_points"], t["input_flatten"], t["spatial_shapes"],
                            t["level_start_index"], None),
        wrt=["query", "reference_points", "input_flatten"]),
    "layer_fusion_v2_d16": dict(
        build=lambda: tl.DeformableTransformerFusionLayerV2(64, 64, 0.0, "gelu", 1, HEADS, 4),
        call=lambda m, t: m(t["tgt"], t["query_pos"], t["reference_points"], t["src"], t["src_spatial_shapes"],
                            t["level_start_index"], t["src_padding_mask"]),
        wrt=["tgt", "query_pos", "src"]),
    "layer_encoder": dict(
        build=lambda: _enc(2),
        call=lambda m, t: m(t["src"], t["pos"], t["reference_points"], t["spatial_shapes"],
                            t["level_start_index"], t["padding_mask"]),
        wrt=["src", "pos"]),
    "layer_fusion_v2": dict(
        build=lambda: tl.DeformableTransformerFusionLayerV2(C, 64, 0.0, "gelu", 1, HEADS, PTS),
        call=lambda m, t: m(t["tgt"], t["query_pos"], t["reference_points"], t["src"], t["src_spatial_shapes"],
                            t["level_start_index"], t["src_padding_mask"]),
        wrt=["tgt", "query_pos", "src"]),
    "layer_late_fusion": dict(
        build=lambda: tl.DepthDeformableTransformerEncoderLayer(C, 64, 0.0, "relu", 1, HEADS, PTS, True, True, True),
        call=lambda m, t: m(t["tgt"], t["query_pos"], None, None, t["reference_points"], None, t["src"],
                            t["src_spatial_shapes"], t["frame_start_index"], None, t["src_padding_mask"]),
        wrt=["tgt", "query_pos", "src"]),
    "layer_decoder": dict(
        build=lambda: tl.DeformableTransformerDecoderLayer(C, 64, 0.0, "relu", 2, HEADS, PTS),
        call=lambda m, t: m(t["tgt"], t["query_pos"], t["reference_points"], t["src"], t["src_spatial_shapes"],
                            t["level_start_index"], t["src_padding_mask"]),
        wrt=["tgt", "src"]),
    "encoder_rgbd_v2": dict(
        build=lambda: tl.RGBDDeformableTransformerEncoderV2(
            _enc(1), tl.DeformableTransformerFusionLayerV2(C, 64, 0.0, "gelu", 1, HEADS, PTS), 3, 2, 2, [0, 1]),
        call=lambda m, t: m(t["src"], t["spatial_shapes"], t["level_start_index"], t["valid_ratios"], t["pos"],
                            t["padding_mask"], None, t["depth_src"], t["depth_spatial_shapes"],
                            t["depth_level_start_index"], None, None, None),
        wrt=["src", "depth_src"]),
    "encoder_plain": dict(
        build=lambda: tl.DeformableTransformerEncoder(_enc(2), 2),
        call=lambda m, t: m(t["src"], t["spatial_shapes"], t["level_start_index"], t["valid_ratios"], t["pos"],
                            t["padding_mask"]),
        wrt=["src"]),
    "backbone_udf_fuse": dict(
        build=lambda: backbone_fusion.DepthDeformableTransformerEncoderLayer(C, 64, 0.0, "relu", 1, HEADS, PTS),
        call=lambda m, t: backbone_fusion.fuse_layers(t["src"], t["target"], t["pos_src"], t["pos_target"],
                                                      t["mask_src"], t["mask_target"], m),
        wrt=["src", "target"]),
}


# wide cases (stored as fp32, see oracle/gen_golden.py): d_model 128 / 8 heads / 4 points is the
# narrowest shape for which the fused deformable-attention kernels AND the fused layer-epilogue
# kernels run in fp32; d_model 256 is the production width (bf16 kernels).
CW, HW, PW, FFW = 128, 8, 4, 128
WIDE = {"layer_encoder_c128", "encoder_plain_c128", "layer_fusion_v2_c128", "decoder_c128",
        "backbone_udf_fuse_c128", "layer_encoder_c256", "layer_fusion_v2_c256", "layer_fusion_v2_c256_l2",
        "layer_late_fusion_c256"}


def _decoder_call(m, t):
    hs, refs = m(t["tgt"], t["reference_points"], t["src"], t["src_spatial_shapes"], t["level_start_index"],
                 t["valid_ratios"], t.get("query_pos"), t.get("src_padding_mask"))
    return torch.cat([hs.flatten(), refs.flatten()])


CASES.update({
    "layer_encoder_c128": dict(
        build=lambda: tl.DeformableTransformerEncoderLayer(CW, FFW, 0.0, "relu", 2, HW, PW),
        call=CASES["layer_encoder"]["call"], wrt=["src", "pos"]),
    "layer_encoder_c256": dict(
        build=lambda: tl.DeformableTransformerEncoderLayer(256, 256, 0.0, "relu", 2, 8, 4),
        call=CASES["layer_encoder"]["call"], wrt=["src", "pos"]),
    # BASELINE.json configs[2] / [3] at the production width (bf16 kernels incl. the tcgen05 projection + LayerNorm path)
    "layer_fusion_v2_c256": dict(
        build=lambda: tl.DeformableTransformerFusionLayerV2(256, 256, 0.0, "gelu", 1, 8, 4),
        call=CASES["layer_fusion_v2"]["call"], wrt=["tgt", "query_pos", "src"]),
    "layer_fusion_v2_c256_l2": dict(
        build=lambda: tl.DeformableTransformerFusionLayerV2(256, 256, 0.0, "gelu", 2, 8, 4),
        call=CASES["layer_fusion_v2"]["call"], wrt=["tgt", "query_pos", "src"]),
    "layer_late_fusion_c256": dict(
        build=lambda: tl.DepthDeformableTransformerEncoderLayer(256, 256, 0.0, "relu", 1, 8, 4, True, True, True),
        call=CASES["layer_late_fusion"]["call"], wrt=["tgt", "query_pos", "src"]),
    "encoder_plain_c128": dict(
        build=lambda: tl.DeformableTransformerEncoder(
            tl.DeformableTransformerEncoderLayer(CW, FFW, 0.0, "relu", 2, HW, PW), 2),
        call=CASES["encoder_plain"]["call"], wrt=["src"]),
    "layer_fusion_v2_c128": dict(
        build=lambda: tl.DeformableTransformerFusionLayerV2(CW, FFW, 0.0, "gelu", 1, HW, PW),
        call=CASES["layer_fusion_v2"]["call"], wrt=["tgt", "query_pos", "src"]),
    "decoder_c128": dict(
        build=lambda: tl.DeformableTransformerDecoder(
            tl.DeformableTransformerDecoderLayer(CW, FFW, 0.0, "relu", 2, HW, PW), 2, return_intermediate=True),
        call=_decoder_call, wrt=["tgt", "src", "query_pos"]),
    "backbone_udf_fuse_c128": dict(
        build=lambda: backbone_fusion.DepthDeformableTransformerEncoderLayer(CW, FFW, 0.0, "relu", 1, HW, PW),
        call=CASES["backbone_udf_fuse"]["call"], wrt=["src", "target"]),
    # temporal layers
    "layer_temporal_encoder": dict(
        build=lambda: tl.TemporalDeformableTransformerEncoderLayer(C, 64, 0.0, "relu", 3, HEADS, PTS),
        call=lambda m, t: m(t["tgt"], t["query_pos"], t["reference_points"], t["src"], t["src_spatial_shapes"],
                            t["frame_start_index"], t["src_padding_mask"]),
        wrt=["tgt", "query_pos", "src"]),
    "temporal_decoder_pp": dict(
        build=lambda: tl.TemporalDeformableTransformerDecoder(
            tl.DeformableTransformerDecoderLayer(C, 64, 0.0, "relu", 1, HEADS, PTS), 2, False),
        call=_decoder_call, wrt=["tgt", "src"]),
})


def _transformer(depth_type, use_depth, n_levels):
    return DeformableTransformer(
        d_model=C, nhead=HEADS, num_encoder_layers=5 if "encoder_cf" in depth_type else 2, num_decoder_layers=2,
        dim_feedforward=64, dropout=0.0, activation="relu", return_intermediate_dec=True,
        num_feature_levels=n_levels, dec_n_points=PTS, enc_n_points=PTS, use_depth=use_depth,
        depth_type=depth_type, dpth_n_points=PTS)


def _transformer_call(n_levels):
    def call(m, t):
        hs, init_ref, inter_ref, _, _ = m(
            [t[f"src{i}"] for i in range(n_levels)], [t[f"mask{i}"] for i in range(n_levels)],
            [t[f"pos{i}"] for i in range(n_levels)], [t["depth_src0"]], [t["depth_mask0"]], [t["depth_pos0"]],
            t["query_embed"])
        return torch.cat([hs.flatten(), init_ref.flatten(), inter_ref.flatten()])
    return call


CASES.update({
    "transformer_baseline": dict(build=lambda: _transformer("Baseline_rgb", False, 2), call=_transformer_call(2),
                                 wrt=["src0", "depth_src0", "query_embed"]),
    "transformer_latefusion": dict(build=lambda: _transformer("DepthDeform_latefusion_dformer", True, 1),
                                   call=_transformer_call(1), wrt=["src0", "depth_src0", "query_embed"]),
    "transformer_encoder_cf": dict(build=lambda: _transformer("DepthDeform_encoder_cf_dformer", True, 1),
                                   call=_transformer_call(1), wrt=["src0", "depth_src0", "query_embed"]),
})


def run_case(name, gold, device, dtype=torch.float64, forward_only=False):
    """Instantiate, load the reference state_dict (strict), run forward + backward.
    Returns (out, {input grads}, {param grads}) as CPU float64 numpy; forward_only: the output of an inference pass
    (no gradients anywhere), as CPU float64 numpy."""
    case = CASES[name]
    module = case["build"]().to(dtype)
    state = {k[len("state."):]: torch.from_numpy(v) for k, v in gold.items() if k.startswith("state.")}
    module.load_state_dict(state, strict=True)          # checkpoint-key compatibility with the reference
    module = module.to(device).eval()
    tensors = {}
    for k, v in gold.items():
        if not k.startswith("in."):
            continue
        t = torch.from_numpy(v).to(device)
        if t.is_floating_point():
            # fp32 inputs of the fp64 cases (valid ratios) stay fp32; wide cases are stored as fp32 throughout
            t = t.to(dtype) if (t.dtype == torch.float64 or name in WIDE) else t
        name_in = k[len("in."):]
        if name_in in case["wrt"] and not forward_only:
            t = t.requires_grad_(True)
        tensors[name_in] = t
    if forward_only:
        with torch.no_grad():
            return case["call"](module, tensors).double().cpu().numpy()
    out = case["call"](module, tensors)
    gout = torch.from_numpy(gold["gout"]).to(device=device, dtype=out.dtype)
    params = dict(module.named_parameters())
    grads = torch.autograd.grad(out, [tensors[k] for k in case["wrt"]] + list(params.values()), gout,
                                allow_unused=True)
    gin = {k: (g.detach().double().cpu().numpy() if g is not None else None) for k, g in zip(case["wrt"], grads)}
    gpar = {k: (g.detach().double().cpu().numpy() if g is not None else None)
            for k, g in zip(params.keys(), grads[len(case["wrt"]):])}
    return out.detach().double().cpu().numpy(), gin, gpar


# ---------------------------------------------------------------------------------------------------
# TransVOD++ multi-frame transformer with its temporal query stage (dfvod_b200.temporal_stage), wired to the
# detector's heads exactly like the golden generator wires the reference
# (oracle/gen_golden.py: TransVODPP; reference deformable_detr_multi_plusplus.py:78-81, 183-194, 320).
class _MLP(torch.nn.Module):
    def __init__(self, input_dim, hidden_dim, output_dim, num_layers):
        super().__init__()
        self.num_layers = num_layers
        hdim = [hidden_dim] * (num_layers - 1)
        self.layers = torch.nn.ModuleList(torch.nn.Linear(a, b) for a, b in zip([input_dim] + hdim, hdim + [output_dim]))

    def forward(self, x):
        for i, layer in enumerate(self.layers):
            x = torch.relu(layer(x)) if i < self.num_layers - 1 else layer(x)
        return x


class TransVODPP(torch.nn.Module):
    def __init__(self, transformer, width, n_dec, n_cls=3):
        super().__init__()
        self.transformer = transformer
        self.class_embed = torch.nn.ModuleList(torch.nn.Linear(width, n_cls) for _ in range(n_dec))
        self.bbox_embed = torch.nn.ModuleList(_MLP(width, width, 4, 3) for _ in range(n_dec))
        self.temp_class_embed_list = torch.nn.ModuleList(torch.nn.Linear(width, n_cls) for _ in range(3))
        self.temp_bbox_embed_list = torch.nn.ModuleList(_MLP(width, width, 4, 3) for _ in range(3))
        self.transformer.decoder.bbox_embed = self.bbox_embed          # with_box_refine

    def forward(self, t):
        hs, init_ref, inter_ref, _, _, final_hs, final_ref, out = self.transformer(
            [t["src0"]], [t["mask0"]], [t["pos0"]], [t["depth_src0"]], [t["depth_mask0"]], [t["depth_pos0"]],
            t["imgs_whwh"], t["query_embed"], self.class_embed[-1], self.bbox_embed[-1],
            self.temp_class_embed_list, self.temp_bbox_embed_list)
        parts = [hs, init_ref, inter_ref, final_hs, final_ref]
        for aux in out["aux_outputs"]:
            parts += [aux["pred_logits"], aux["pred_boxes"]]
        return torch.cat([p.flatten() for p in parts])


def _transvodpp(depth_type, use_depth, ref_frames, width=16):
    from dfvod_b200 import temporal_stage
    return TransVODPP(temporal_stage.DeformableTransformer(
        d_model=width, nhead=HEADS, num_encoder_layers=2, num_decoder_layers=2, dim_feedforward=64, dropout=0.0,
        activation="relu", return_intermediate_dec=True, num_feature_levels=1, dec_n_points=PTS, enc_n_points=PTS,
        num_query=80, n_temporal_decoder_layers=1, num_ref_frames=ref_frames, use_depth=use_depth,
        depth_type=depth_type, dpth_n_points=PTS), width, 2)


CASES.update({
    "transvodpp_f1": dict(build=lambda: _transvodpp("Baseline_rgb", False, 1), call=lambda m, t: m(t),
                          wrt=["src0", "depth_src0", "query_embed"]),
    "transvodpp_f2_latefusion": dict(build=lambda: _transvodpp("DepthDeform_latefusion_dformer", True, 2),
                                     call=lambda m, t: m(t), wrt=["src0", "depth_src0", "query_embed"]),
})


# two-stage transformer (reference single.py:82-86, :112-153, :308-322); 256 wide because get_proposal_pos_embed
# hard-codes 128 features per box coordinate.  Stored like the wide cases (fp32 results, fp16-representable state).
def _two_stage():
    model = DeformableTransformer(
        d_model=256, nhead=8, num_encoder_layers=1, num_decoder_layers=1, dim_feedforward=64, dropout=0.0,
        activation="relu", return_intermediate_dec=True, num_feature_levels=2, dec_n_points=2, enc_n_points=2,
        two_stage=True, two_stage_num_proposals=6)
    model.decoder.class_embed = torch.nn.ModuleList(torch.nn.Linear(256, 3) for _ in range(2))
    model.decoder.bbox_embed = torch.nn.ModuleList(
        torch.nn.Sequential(torch.nn.Linear(256, 32), torch.nn.ReLU(), torch.nn.Linear(32, 4)) for _ in range(2))
    return model


def _two_stage_call(m, t):
    hs, init_ref, inter_ref, enc_cls, enc_coord = m(
        [t["src0"], t["src1"]], [t["mask0"], t["mask1"]], [t["pos0"], t["pos1"]], None, None, None, None)
    enc_coord = torch.where(torch.isinf(enc_coord), torch.zeros_like(enc_coord), enc_coord)
    return torch.cat([hs.flatten(), init_ref.flatten(), inter_ref.flatten(), enc_cls.flatten(), enc_coord.flatten()])


CASES["transformer_two_stage"] = dict(build=_two_stage, call=_two_stage_call, wrt=["src0"])
WIDE.add("transformer_two_stage")


def _rgbd_query_call(m, t):
    hs, init_ref, inter_ref, _, _ = m(
        [t["src0"], t["src1"]], [t["mask0"], t["mask1"]], [t["pos0"], t["pos1"]], None, None, None, t["query_embed"],
        [t["rgbd0"], t["rgbd1"]])
    return torch.cat([hs.flatten(), init_ref.flatten(), inter_ref.flatten()])


CASES["transformer_rgbd_query"] = dict(
    build=lambda: DeformableTransformer(
        d_model=C, nhead=HEADS, num_encoder_layers=2, num_decoder_layers=1, dim_feedforward=64, dropout=0.0,
        activation="relu", return_intermediate_dec=True, num_feature_levels=2, dec_n_points=PTS, enc_n_points=PTS,
        use_depth=False, depth_type="Baseline_concat"),
    call=_rgbd_query_call, wrt=["src0", "rgbd0", "query_embed"])


# TransVOD multi-frame transformer (reference deformable_transformer_multi.py), wired like oracle/gen_golden.py: TransVOD
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


def _transvod(depth_type, use_depth, ref_frames, tdam, width=16):
    from dfvod_b200 import temporal_stage
    model = TransVOD(temporal_stage.TransVODDeformableTransformer(
        d_model=width, nhead=HEADS, num_encoder_layers=2, num_decoder_layers=2, dim_feedforward=64, dropout=0.0,
        activation="relu", return_intermediate_dec=True, num_feature_levels=1, dec_n_points=PTS, enc_n_points=PTS,
        n_temporal_decoder_layers=1, num_ref_frames=ref_frames, use_depth=use_depth, depth_type=depth_type,
        dpth_n_points=PTS), width)
    model.transformer.TDAM = tdam
    return model


CASES.update({
    "transvod_f1": dict(build=lambda: _transvod("Baseline_rgb", False, 1, False), call=lambda m, t: m(t),
                        wrt=["src0", "depth_src0", "query_embed"]),
    "transvod_f2_tdam": dict(build=lambda: _transvod("DepthDeform_latefusion_dformer", True, 2, True),
                             call=lambda m, t: m(t), wrt=["src0", "depth_src0", "query_embed"]),
})
