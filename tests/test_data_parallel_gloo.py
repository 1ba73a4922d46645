"""CPU, world_size 2 over gloo: the frame-sharded data-parallel plumbing.
Two ranks each run an Encoder-Cross-Fusion layer on their shard of the frames; after
GradientAllReducer the gradients on both ranks equal those of ONE process running the whole batch
(mean loss over frames) -- the DDP contract of /root/reference/main.py:440-442.
The deformable-attention op itself is replaced by the oracle (CPU) in this host-logic test."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import msda_oracle
from dfvod_b200 import data_parallel
from dfvod_b200 import transformer_layers as tl
import dfvod_b200.ops.modules.ms_deform_attn as msda_module


class _OracleFunction:
    @staticmethod
    def apply(value, shapes, lsi, loc, attn, im2col_step):
        return msda_oracle.core_pytorch(value, shapes.tolist(), loc, attn)


def _build(seed=0):
    torch.manual_seed(seed)
    layer = tl.DeformableTransformerFusionLayerV2(32, 64, 0.0, "gelu", 1, 4, 2).double()
    with torch.no_grad():
        for p in layer.parameters():
            p.add_(torch.randn_like(p) * 0.05)
    return layer


def _data(n_frames, seed=1):
    g = torch.Generator().manual_seed(seed)
    shapes = torch.tensor([[4, 5]])
    lsi = torch.tensor([0])
    s = 20
    tgt = torch.randn(n_frames, s, 32, generator=g, dtype=torch.float64)
    pos = torch.randn(n_frames, s, 32, generator=g, dtype=torch.float64)
    depth = torch.randn(n_frames, s, 32, generator=g, dtype=torch.float64)
    ref = tl.encoder_reference_points([(4, 5)], torch.ones(n_frames, 1, 2, dtype=torch.float64), "cpu").double()
    return shapes, lsi, tgt, pos, depth, ref


def _loss(layer, shapes, lsi, tgt, pos, depth, ref):
    out = layer(tgt, pos, ref, depth, shapes, lsi, None)
    return out.square().mean(dim=(1, 2)).sum()          # sum over frames of a per-frame loss


def _worker(rank, world, port, n_frames, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    msda_module.MSDeformAttnFunction = _OracleFunction
    try:
        layer = _build()
        shapes, lsi, tgt, pos, depth, ref = _data(n_frames)
        a, b = data_parallel.shard_range(n_frames, world, rank)
        reducer = data_parallel.GradientAllReducer(layer.parameters(), bucket_bytes=4096)
        assert len(reducer.buckets) > 1                                # exercise the bucketing
        # each rank normalises by ITS frame count times world/N ... use sum loss scaled so that the
        # average over ranks equals the global mean over frames
        loss = _loss(layer, shapes, lsi, tgt[a:b], pos[a:b], depth[a:b], ref[a:b]) * (world / n_frames)
        loss.backward()
        reducer.finish()
        grads = {k: p.grad.clone() for k, p in layer.named_parameters()}
        gathered = [None] * world
        dist.all_gather_object(gathered, {k: v.numpy() for k, v in grads.items()})
        if rank == 0:
            for other in gathered[1:]:
                for k in gathered[0]:
                    assert np.array_equal(gathered[0][k], other[k]), f"ranks disagree on {k}"
            np.savez(out_path, **gathered[0])
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 8, 9, 33):
        for world in (1, 2, 3, 8):
            spans = [data_parallel.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        data_parallel.shard_range(4, 2, 2)


@pytest.mark.timeout(300)
def test_two_rank_gradients_equal_single_process(tmp_path, monkeypatch):
    n_frames = 5                                  # ragged split: 3 + 2 frames
    out_path = str(tmp_path / "grads.npz")
    mp.spawn(_worker, args=(2, _free_port(), n_frames, out_path), nprocs=2, join=True)
    got = dict(np.load(out_path))
    monkeypatch.setattr(msda_module, "MSDeformAttnFunction", _OracleFunction)
    layer = _build()
    shapes, lsi, tgt, pos, depth, ref = _data(n_frames)
    (_loss(layer, shapes, lsi, tgt, pos, depth, ref) / n_frames).backward()
    for k, p in layer.named_parameters():
        np.testing.assert_allclose(got[k], p.grad.numpy(), rtol=1e-10, atol=1e-12, err_msg=k)


def test_single_process_reducer_is_identity(monkeypatch):
    monkeypatch.setattr(msda_module, "MSDeformAttnFunction", _OracleFunction)
    layer = _build()
    shapes, lsi, tgt, pos, depth, ref = _data(2)
    reducer = data_parallel.GradientAllReducer(layer.parameters())
    _loss(layer, shapes, lsi, tgt, pos, depth, ref).backward()
    before = {k: p.grad.clone() for k, p in layer.named_parameters()}
    reducer.finish()
    for k, p in layer.named_parameters():
        assert torch.equal(before[k], p.grad)
    assert reducer.gradient_bytes == sum(p.numel() * 8 for p in layer.parameters())
