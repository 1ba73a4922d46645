"""``DeformableTransformer`` -- the single-frame transformer that assembles the MSDeformAttn layers
(host-side mirror of /root/reference/models/deformable_transformer_single.py:23-337).

Same constructor, same ``forward(srcs, masks, pos_embeds, depth_srcs, depth_masks,
depth_pos_embeds, query_embed, rgbd_query)``, same parameter names (``level_embed``,
``reference_points``, ``encoder.layers.*``, ``encoder.fusion_layers.*``, ``decoder.layers.*``,
``depth_encoder_layer.*``) so reference checkpoints load.  The fusion variant is selected, as in the
reference, by substrings of ``depth_type``:

    "latefusion"  one DepthDeformableTransformerEncoderLayer before the encoder      (:49-52, :212-244)
    "encoder_cf"  RGBDDeformableTransformerEncoderV2, fusion after layers 0..3       (:55-66, :269-302)
    otherwise     plain DeformableTransformerEncoder                                 (:68-73)

TransVOD / TransVOD++ (deformable_transformer_multi*.py) run exactly this stack with the frames
of a clip as the batch dimension; their temporal query stage (mmcv RoIAlign, RCNNHead) is outside
the hot path.
"""
import math

import torch
from torch import nn
from torch.nn.init import constant_, normal_, xavier_uniform_

from .ops.functions import flatten_levels, flatten_levels_supported
from .ops.modules import MSDeformAttn
from .transformer_layers import (DeformableTransformerDecoder, DeformableTransformerDecoderLayer,
                                 DeformableTransformerEncoder, DeformableTransformerEncoderLayer,
                                 DeformableTransformerFusionLayerV2, DepthDeformableTransformerEncoderLayer,
                                 RGBDDeformableTransformerEncoderV2, encoder_reference_points)


def _as_tokens(level):
    """A level as tokens [N, H*W, C]: NCHW maps are flattened and transposed (single.py:195-198); levels that already
    are token-major (input_projection.InputProjection.forward_tokens) pass through."""
    return level if level.dim() == 3 else level.flatten(2).transpose(1, 2)


def _flatten_levels(maps, masks, pos_embeds, level_embed=None):
    """[N,C,H,W] (or token-major [N,H*W,C]) per level -> tokens [N, sum HW, C], mask [N, sum HW], pos (+ level
    embedding), shapes (python list of (H, W), read from the masks)."""
    shapes = [(mask.shape[1], mask.shape[2]) for mask in masks]
    for feat, (h, w) in zip(maps, shapes):
        assert (feat.shape[1] == h * w) if feat.dim() == 3 else (tuple(feat.shape[2:]) == (h, w)), \
            "feature level and its mask disagree on (H, W)"
    flat_mask = masks[0].flatten(1) if len(masks) == 1 else torch.cat([m.flatten(1) for m in masks], 1)
    embed_trains = level_embed is not None and torch.is_grad_enabled() and level_embed.requires_grad
    nchw = [m for m in maps if m.dim() == 4]
    if all(m.dim() == 3 for m in maps):
        tokens = maps[0] if len(maps) == 1 else torch.cat(list(maps), 1)
    elif len(nchw) == len(maps) and flatten_levels_supported(list(maps)):
        tokens = flatten_levels(list(maps))          # inference: one transposing kernel per level, no cat
    else:
        tokens = torch.cat([_as_tokens(m) for m in maps], 1)
    if isinstance(pos_embeds, torch.Tensor):
        # already lvl_pos_embed_flatten [N, sum HW, C], level embedding included
        # (position_encoding.PositionEmbeddingSine.forward_tokens)
        assert pos_embeds.dim() == 3 and pos_embeds.shape[1] == flat_mask.shape[1], \
            "flattened position embedding and masks disagree on the token count"
        pos = pos_embeds
    elif not embed_trains and flatten_levels_supported(list(pos_embeds)):
        adds = None if level_embed is None else [level_embed[lvl] for lvl in range(len(maps))]
        pos = flatten_levels(list(pos_embeds), adds)
    else:
        flat_pos = []
        for lvl, p in enumerate(pos_embeds):
            p = _as_tokens(p)
            flat_pos.append(p if level_embed is None else p + level_embed[lvl].view(1, 1, -1))
        pos = flat_pos[0] if len(flat_pos) == 1 else torch.cat(flat_pos, 1)
    return tokens, flat_mask, pos, shapes


_SHAPE_TENSORS = {}


def _shape_tensors(shapes, device):
    """(spatial_shapes [L,2], level_start_index [L]) int64 on `device`.  Built once per
    (shapes, device): the reference re-creates them from a python list every forward
    (single.py:207-208), an H2D copy that also prevents CUDA-graph capture."""
    key = (tuple(shapes), str(device))
    hit = _SHAPE_TENSORS.get(key)
    if hit is None:
        if len(_SHAPE_TENSORS) >= 64:
            _SHAPE_TENSORS.clear()
        # built outside inference mode: an inference tensor cached here by a first call under torch.inference_mode()
        # could not be saved for backward by a later training step
        with torch.inference_mode(False):
            st = torch.as_tensor(shapes, dtype=torch.long, device=device)
            hit = (st, torch.cat((st.new_zeros((1,)), st.prod(1).cumsum(0)[:-1])))
        _SHAPE_TENSORS[key] = hit
    return hit


class DeformableTransformer(nn.Module):
    def __init__(self, d_model=256, nhead=8, num_encoder_layers=6, num_decoder_layers=6, dim_feedforward=1024,
                 dropout=0.1, activation="relu", return_intermediate_dec=False, num_feature_levels=4,
                 dec_n_points=4, enc_n_points=4, two_stage=False, two_stage_num_proposals=300,
                 use_depth=False, depth_type="Baseline_rgb", dpth_feature_levels=1, dpth_n_points=4):
        super().__init__()
        self.use_depth = use_depth
        self.depth_type = depth_type
        self.residual_fusion = "noresidual" not in depth_type
        self.rgbd_query = "concat" in depth_type
        self.d_model = d_model
        self.nhead = nhead
        self.two_stage = two_stage
        self.two_stage_num_proposals = two_stage_num_proposals
        self.depth_self_attn = True
        self.late_fusion_layers = 1
        self.adaptation_layers = True
        self.gate = True
        self.encoder_cross_fusion = True

        if "latefusion" in depth_type:
            self.depth_encoder_layer = DepthDeformableTransformerEncoderLayer(
                d_model, dim_feedforward, dropout, activation, dpth_feature_levels, nhead, dpth_n_points,
                self.depth_self_attn, self.gate, self.adaptation_layers)

        encoder_layer = DeformableTransformerEncoderLayer(d_model, dim_feedforward, dropout, activation,
                                                          num_feature_levels, nhead, enc_n_points)
        if "encoder_cf" in depth_type:
            self.num_depth_encoder_layers = 4
            self.num_enc_fusion_layers = 4
            self.enc_fusion_layers_order = [0, 1, 2, 3]
            fusion_layer = DeformableTransformerFusionLayerV2(d_model, dim_feedforward, dropout, activation,
                                                              num_feature_levels, nhead, enc_n_points)
            self.encoder = RGBDDeformableTransformerEncoderV2(
                encoder_layer, fusion_layer, num_encoder_layers, self.num_depth_encoder_layers,
                self.num_enc_fusion_layers, self.enc_fusion_layers_order)
        else:
            self.encoder = DeformableTransformerEncoder(encoder_layer, num_encoder_layers)

        decoder_layer = DeformableTransformerDecoderLayer(d_model, dim_feedforward, dropout, activation,
                                                          num_feature_levels, nhead, dec_n_points)
        self.decoder = DeformableTransformerDecoder(decoder_layer, num_decoder_layers, return_intermediate_dec)

        self.level_embed = nn.Parameter(torch.Tensor(num_feature_levels, d_model))
        self._build_extra_modules()
        if two_stage:
            self.enc_output = nn.Linear(d_model, d_model)
            self.enc_output_norm = nn.LayerNorm(d_model)
            self.pos_trans = nn.Linear(d_model * 2, d_model * 2)
            self.pos_trans_norm = nn.LayerNorm(d_model * 2)
        else:
            self.reference_points = nn.Linear(d_model, 2)
        self._reset_parameters()

    def _build_extra_modules(self):
        """Hook for the multi-frame transformer (temporal_stage.py): modules registered here take part in
        ``_reset_parameters`` exactly as in the reference constructor."""

    def _reset_parameters(self):
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        for m in self.modules():
            if isinstance(m, MSDeformAttn):
                m._reset_parameters()
            elif isinstance(m, nn.LayerNorm):
                constant_(m.weight, 1.0)
                constant_(m.bias, 0.0)
        if not self.two_stage:
            xavier_uniform_(self.reference_points.weight.data, gain=1.0)
            constant_(self.reference_points.bias.data, 0.)
        normal_(self.level_embed)

    # ------------------------------------------------------------------ two-stage helpers (:112-153)
    def get_proposal_pos_embed(self, proposals):
        num_pos_feats, temperature = 128, 10000
        dim_t = torch.arange(num_pos_feats, dtype=torch.float32, device=proposals.device)
        dim_t = temperature ** (2 * (dim_t // 2) / num_pos_feats)
        pos = (proposals.sigmoid() * (2 * math.pi))[:, :, :, None] / dim_t              # N, L, 4, 128
        return torch.stack((pos[:, :, :, 0::2].sin(), pos[:, :, :, 1::2].cos()), dim=4).flatten(2)

    def gen_encoder_output_proposals(self, memory, memory_padding_mask, spatial_shapes):
        N_ = memory.shape[0]
        proposals, cursor = [], 0
        for lvl, (H_, W_) in enumerate(spatial_shapes):
            H_, W_ = int(H_), int(W_)
            level_mask = memory_padding_mask[:, cursor:cursor + H_ * W_].view(N_, H_, W_, 1)
            valid_H = torch.sum(~level_mask[:, :, 0, 0], 1)
            valid_W = torch.sum(~level_mask[:, 0, :, 0], 1)
            grid_y, grid_x = torch.meshgrid(
                torch.linspace(0, H_ - 1, H_, dtype=torch.float32, device=memory.device),
                torch.linspace(0, W_ - 1, W_, dtype=torch.float32, device=memory.device), indexing="ij")
            grid = torch.cat([grid_x.unsqueeze(-1), grid_y.unsqueeze(-1)], -1)
            scale = torch.cat([valid_W.unsqueeze(-1), valid_H.unsqueeze(-1)], 1).view(N_, 1, 1, 2)
            grid = (grid.unsqueeze(0).expand(N_, -1, -1, -1) + 0.5) / scale
            wh = torch.ones_like(grid) * 0.05 * (2.0 ** lvl)
            proposals.append(torch.cat((grid, wh), -1).view(N_, -1, 4))
            cursor += H_ * W_
        output_proposals = torch.cat(proposals, 1)
        valid = ((output_proposals > 0.01) & (output_proposals < 0.99)).all(-1, keepdim=True)
        output_proposals = torch.log(output_proposals / (1 - output_proposals))
        output_proposals = output_proposals.masked_fill(memory_padding_mask.unsqueeze(-1), float('inf'))
        output_proposals = output_proposals.masked_fill(~valid, float('inf'))
        output_memory = memory.masked_fill(memory_padding_mask.unsqueeze(-1), float(0))
        output_memory = output_memory.masked_fill(~valid, float(0))
        output_memory = self.enc_output_norm(self.enc_output(output_memory))
        return output_memory, output_proposals

    @staticmethod
    def get_valid_ratio(mask):
        """(w, h) fraction of each [N,H,W] mask that is not padding (:155-162)."""
        _, H, W = mask.shape
        valid_h = torch.sum(~mask[:, :, 0], 1).float() / H
        valid_w = torch.sum(~mask[:, 0, :], 1).float() / W
        return torch.stack([valid_w, valid_h], -1)

    get_reference_points = staticmethod(encoder_reference_points)

    # ------------------------------------------------------------------ forward (:179-337)
    def _flatten_depth(self, depth_srcs, depth_masks, depth_pos_embeds):
        assert depth_srcs is not None and depth_masks is not None and depth_pos_embeds is not None, \
            "Depth information is required for Deformable DETR with depth"
        assert len(depth_srcs) == len(depth_masks) == len(depth_pos_embeds), \
            "The number of depth sources, masks and pos_embeds should be the same"
        tokens, mask, pos, shapes = _flatten_levels(depth_srcs, depth_masks, depth_pos_embeds, None)
        st, ls = _shape_tensors(shapes, tokens.device)
        ratios = torch.stack([self.get_valid_ratio(m) for m in depth_masks], 1)
        return tokens, mask, pos, shapes, st, ls, ratios

    def forward(self, srcs, masks, pos_embeds, depth_srcs, depth_masks, depth_pos_embeds, query_embed=None,
                rgbd_query=[], _return_state=False):
        assert self.two_stage or query_embed is not None
        src_flatten, mask_flatten, lvl_pos_embed_flatten, shapes = _flatten_levels(
            srcs, masks, pos_embeds, self.level_embed)
        rgbd_flatten = None
        if len(rgbd_query) > 0:
            rgbd_flatten = torch.cat([q.flatten(2).transpose(1, 2) for q in rgbd_query], 1)
        spatial_shapes, level_start_index = _shape_tensors(shapes, src_flatten.device)
        valid_ratios = torch.stack([self.get_valid_ratio(m) for m in masks], 1)
        rgbd_arg = rgbd_flatten if self.rgbd_query else None

        if "latefusion" in self.depth_type and self.use_depth:
            d_tok, d_mask, d_pos, d_shapes, d_st, d_ls, d_ratios = self._flatten_depth(
                depth_srcs, depth_masks, depth_pos_embeds)
            rgb_ref = self.get_reference_points(shapes, valid_ratios, device=src_flatten.device)
            depth_ref = self.get_reference_points(d_shapes, d_ratios, device=d_tok.device)
            fused = self.depth_encoder_layer(src_flatten, lvl_pos_embed_flatten, d_pos, spatial_shapes, rgb_ref,
                                             depth_ref, d_tok, d_st, d_ls, mask_flatten, d_mask)
            src_flatten = src_flatten + fused

        if "encoder_cf" in self.depth_type and self.use_depth:
            d_tok, d_mask, d_pos, d_shapes, d_st, d_ls, d_ratios = self._flatten_depth(
                depth_srcs, depth_masks, depth_pos_embeds)
            memory = self.encoder(src_flatten, spatial_shapes, level_start_index, valid_ratios, lvl_pos_embed_flatten,
                                  mask_flatten, rgbd_arg, d_tok, d_st, d_ls, d_ratios, d_pos, d_mask)
        else:
            memory = self.encoder(src_flatten, spatial_shapes, level_start_index, valid_ratios, lvl_pos_embed_flatten,
                                  mask_flatten, rgbd_arg)

        bs, _, c = memory.shape
        enc_outputs_class = enc_outputs_coord_unact = None
        if self.two_stage:
            output_memory, output_proposals = self.gen_encoder_output_proposals(memory, mask_flatten, shapes)
            enc_outputs_class = self.decoder.class_embed[self.decoder.num_layers](output_memory)
            enc_outputs_coord_unact = self.decoder.bbox_embed[self.decoder.num_layers](output_memory) + output_proposals
            topk_proposals = torch.topk(enc_outputs_class[..., 0], self.two_stage_num_proposals, dim=1)[1]
            topk_coords_unact = torch.gather(enc_outputs_coord_unact, 1,
                                             topk_proposals.unsqueeze(-1).repeat(1, 1, 4)).detach()
            reference_points = topk_coords_unact.sigmoid()
            pos_trans_out = self.pos_trans_norm(self.pos_trans(self.get_proposal_pos_embed(topk_coords_unact)))
            query_embed, tgt = torch.split(pos_trans_out, c, dim=2)
        else:
            query_embed, tgt = torch.split(query_embed, c, dim=1)
            query_embed = query_embed.unsqueeze(0).expand(bs, -1, -1)
            tgt = tgt.unsqueeze(0).expand(bs, -1, -1)
            reference_points = self.reference_points(query_embed).sigmoid()
        init_reference_out = reference_points

        hs, inter_references = self.decoder(tgt, reference_points, memory, spatial_shapes, level_start_index,
                                            valid_ratios, query_embed, mask_flatten)
        if _return_state:         # what the TransVOD++ temporal stage continues from (temporal_stage.py)
            return (hs, init_reference_out, inter_references, enc_outputs_class, enc_outputs_coord_unact,
                    (memory, lvl_pos_embed_flatten, spatial_shapes, level_start_index, valid_ratios, shapes))
        return hs, init_reference_out, inter_references, enc_outputs_class, enc_outputs_coord_unact


def build_deforamble_transformer(args):
    """Same (misspelt) name and argument mapping as the reference builder (single.py:766-785)."""
    return DeformableTransformer(
        d_model=args.hidden_dim, nhead=args.nheads, num_encoder_layers=args.enc_layers,
        num_decoder_layers=args.dec_layers, dim_feedforward=args.dim_feedforward, dropout=args.dropout,
        activation="relu", return_intermediate_dec=True, num_feature_levels=args.num_feature_levels,
        dec_n_points=args.dec_n_points, enc_n_points=args.enc_n_points, two_stage=args.two_stage,
        two_stage_num_proposals=args.num_queries, use_depth=args.use_depth, depth_type=args.depth_type,
        dpth_n_points=args.dpth_n_points)
