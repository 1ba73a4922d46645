"""GPU: msda_layer_sine_position_tokens (one kernel per level writing lvl_pos_embed_flatten) against the golden
vectors of the reference's PositionEmbeddingSine (tests/golden/position_sine.npz, oracle/gen_golden.py) and, at the
COCO pyramid, against the reference composition evaluated by PyTorch on the same device.
Tolerance: fp32 2e-6 absolute (values in [-1, 1] + level embedding: sinf / cosf of the CUDA math library on both
sides, arguments up to 2*pi); bf16 / fp16 one rounding of the fp32 value plus one of the sum."""
import pytest
import torch

from dfvod_b200.position_encoding import PositionEmbeddingSine
from tests.util import COCO_SHAPES, load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("tag,kwargs", [("norm", dict(num_pos_feats=16, normalize=True)),
                                        ("raw", dict(num_pos_feats=16, temperature=20))])
def test_kernel_matches_reference_golden(tag, kwargs):
    gold = load_golden("position_sine")
    module = PositionEmbeddingSine(**kwargs)
    masks = [torch.from_numpy(gold[f"mask{lvl}"]).to(DEV) for lvl in range(3)]
    level_embed = torch.from_numpy(gold["level_embed"]).to(DEV)
    pos = [torch.from_numpy(gold[f"{tag}_pos{lvl}"]).to(DEV) for lvl in range(3)]
    want = torch.cat([p.flatten(2).transpose(1, 2) + level_embed[lvl].view(1, 1, -1) for lvl, p in enumerate(pos)], 1)
    got = module.forward_tokens(masks, level_embed)
    assert got.shape == want.shape and got.dtype == torch.float32
    assert float((got - want).abs().max()) <= 2e-6
    for lvl, p in enumerate(pos):                           # the NCHW drop-in forward (a view of the tokens)
        nchw = module((None, masks[lvl]))
        assert nchw.shape == p.shape and float((nchw - p).abs().max()) <= 2e-6


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_coco_pyramid_against_device_composition(dtype):
    torch.manual_seed(0)
    module = PositionEmbeddingSine(128, normalize=True)
    n = 3
    masks = []
    for h, w in COCO_SHAPES:
        m = torch.zeros(n, h, w, dtype=torch.bool, device=DEV)
        m[1, :, w - w // 5:] = True
        m[2, h - h // 3:, :] = True
        masks.append(m)
    level_embed = torch.randn(len(masks), 256, device=DEV).to(dtype)
    got = module.forward_tokens(masks, level_embed, dtype=dtype)
    want = torch.cat([module._host_composition(m).to(dtype).flatten(1, 2) + level_embed[lvl].view(1, 1, -1)
                      for lvl, m in enumerate(masks)], 1)
    assert got.shape == (n, 22223, 256) and got.dtype == dtype
    err = float((got.float() - want.float()).abs().max())
    if dtype == torch.float32:
        assert err <= 2e-6
    else:                                                   # an fp32 ulp may flip the rounding of the fp32 -> 16-bit cast
        ulp = 2.0 ** -7 if dtype == torch.bfloat16 else 2.0 ** -10
        assert err <= 4 * ulp
        assert float((got.float() - want.float()).abs().gt(0).float().mean()) < 1e-3
    # no level embedding: plain cast of the embedding
    plain = module.forward_tokens(masks[1:2], None, dtype=dtype)
    assert float((plain.float() - module._host_composition(masks[1]).to(dtype).flatten(1, 2).float()).abs().max()) \
        <= (2e-6 if dtype == torch.float32 else 2.0 ** -8)


def test_transformer_accepts_flattened_position_tokens():
    """DeformableTransformer.forward(srcs, masks, <lvl_pos_embed_flatten>, ...) == the per-level list."""
    from dfvod_b200.deformable_transformer import DeformableTransformer
    torch.manual_seed(5)
    shapes = [(12, 17), (6, 9)]
    model = DeformableTransformer(d_model=64, nhead=4, num_encoder_layers=2, num_decoder_layers=2, dim_feedforward=128,
                                  dropout=0.0, num_feature_levels=2, return_intermediate_dec=True).to(DEV).eval()
    srcs = [torch.randn(2, 64, h, w, device=DEV) for h, w in shapes]
    masks = [torch.zeros(2, h, w, dtype=torch.bool, device=DEV) for h, w in shapes]
    masks[0][1, :, -3:] = True
    masks[1][1, :, -2:] = True
    sine = PositionEmbeddingSine(32, normalize=True)
    query = torch.randn(10, 128, device=DEV)
    with torch.no_grad():
        per_level = model(srcs, masks, [sine((None, m)) for m in masks], None, None, None, query)[0]
        flattened = model(srcs, masks, sine.forward_tokens(masks, model.level_embed), None, None, None, query)[0]
    assert float((per_level - flattened).abs().max() / per_level.abs().max()) <= 1e-6


def test_bad_arguments_are_errors():
    from dfvod_b200 import _lib
    lib = _lib.load()
    buf = torch.zeros(64, device=DEV)
    code = lib.msda_layer_sine_position_tokens(_lib.DTYPE_F32, buf.data_ptr(), buf.data_ptr(), buf.data_ptr(), 4, None, 1,
                                               8, buf.data_ptr(), 4, 0, None)      # level does not fit the item
    assert code != 0
    code = lib.msda_layer_sine_position_tokens(_lib.DTYPE_F64, buf.data_ptr(), buf.data_ptr(), buf.data_ptr(), 4, None, 1,
                                               8, buf.data_ptr(), 8, 0, None)      # fp64 not offered
    assert code != 0
    torch.cuda.synchronize()


@pytest.mark.parametrize("normalize", [False, True])
@pytest.mark.parametrize("n,h,w", [(3, 100, 167), (2, 13, 21), (1, 1, 1), (2, 5, 64), (2, 33, 32)])
def test_coordinates_kernel_is_bitwise_the_reference_ops(normalize, n, h, w):
    """msda_layer_sine_coordinates == cumsum / (c - 0.5) / (last + eps) * scale of position_encoding.py:39-46 evaluated
    by PyTorch on the device, bit for bit -- ragged padding, a fully padded row and column included."""
    torch.manual_seed(h * w)
    module = PositionEmbeddingSine(8, normalize=normalize, scale=3.0 if normalize else None)
    mask = torch.zeros(n, h, w, dtype=torch.bool, device=DEV)
    mask[0] = torch.rand(h, w, device=DEV) < 0.3                 # holes anywhere (the op only sees a byte map)
    if n > 1:
        mask[1, h - h // 3:, :] = True
        mask[1, :, w - w // 4:] = True
    if n > 2:
        mask[2, 0, :] = True                                     # a row / column with no valid pixel at all
        mask[2, :, 0] = True
    want_y, want_x = module._coordinates(mask)
    got_y, got_x = module._device_coordinates(mask, torch.cuda.current_stream().cuda_stream)
    assert torch.equal(got_y, want_y) and torch.equal(got_x, want_x)
