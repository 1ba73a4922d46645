"""CPU: the PositionEmbeddingSine mirror (host composition, constructor contract, token-major flattening) against
tests/golden/position_sine.npz, which oracle/gen_golden.py wrote by running the reference class
(/root/reference/models/position_encoding.py:20-56) on padded masks."""
import pytest
import torch

from dfvod_b200.position_encoding import PositionEmbeddingSine, build_position_encoding
from tests.util import load_golden

CASES = {"norm": dict(num_pos_feats=16, normalize=True), "raw": dict(num_pos_feats=16, temperature=20)}


@pytest.mark.parametrize("tag", sorted(CASES))
def test_host_composition_is_the_reference(tag):
    gold = load_golden("position_sine")
    module = PositionEmbeddingSine(**CASES[tag])
    for lvl in range(3):
        mask = torch.from_numpy(gold[f"mask{lvl}"])
        want = torch.from_numpy(gold[f"{tag}_pos{lvl}"])
        assert want.dtype == torch.float32
        got = module((torch.zeros(2, 1, *mask.shape[1:]), mask))
        assert got.shape == want.shape and got.dtype == torch.float32
        assert torch.equal(got, want)                       # same ops in the same order: bit for bit


def test_forward_tokens_is_flatten_transpose_plus_level_embed():
    gold = load_golden("position_sine")
    module = PositionEmbeddingSine(16, normalize=True)
    masks = [torch.from_numpy(gold[f"mask{lvl}"]) for lvl in range(3)]
    level_embed = torch.from_numpy(gold["level_embed"])
    want = torch.cat([torch.from_numpy(gold[f"norm_pos{lvl}"]).flatten(2).transpose(1, 2) + level_embed[lvl].view(1, 1, -1)
                      for lvl in range(3)], 1)               # deformable_transformer_single.py:196-206
    got = module.forward_tokens(masks, level_embed)
    assert torch.equal(got, want)


def test_constructor_contract():
    with pytest.raises(ValueError):
        PositionEmbeddingSine(8, normalize=False, scale=3.0)  # position_encoding.py:30-31

    class Args:
        hidden_dim = 256
        position_embedding = "sine"
    built = build_position_encoding(Args)
    assert built.num_pos_feats == 128 and built.normalize and len(list(built.parameters())) == 0
    Args.position_embedding = "bogus"
    with pytest.raises(ValueError):
        build_position_encoding(Args)
