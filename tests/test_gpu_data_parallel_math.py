"""GPU: SURVEY.md C5 on the real kernels with ONE device -- a data-parallel step over two shards (gradients of each
shard's mean loss, averaged: what the all-reduce of dfvod_b200.data_parallel computes) must equal the step on the
concatenated batch.  (Cross-rank equality of the parameters after real NCCL steps is asserted by bench.py's
train-step extra; the reducer's multi-process logic by tests/test_data_parallel_gloo.py.)"""
import pytest
import torch

from dfvod_b200 import transformer_layers as tl

pytestmark = pytest.mark.gpu


def test_two_shard_average_equals_the_concatenated_batch_step():
    torch.manual_seed(11)
    dev = torch.device("cuda:0")
    shapes = [(12, 16), (6, 8)]
    s = sum(h * w for h, w in shapes)
    st = torch.as_tensor(shapes, dtype=torch.long, device=dev)
    ls = torch.as_tensor([0, shapes[0][0] * shapes[0][1]], dtype=torch.long, device=dev)
    layer = tl.DeformableTransformerEncoderLayer(128, 128, 0.0, "relu", 2, 8, 4).to(dev)
    frames = 4
    src = torch.randn(frames, s, 128, device=dev)
    pos = torch.randn(frames, s, 128, device=dev)
    ref = []
    for h, w in shapes:
        ys, xs = torch.meshgrid((torch.arange(h, device=dev) + 0.5) / h, (torch.arange(w, device=dev) + 0.5) / w, indexing="ij")
        ref.append(torch.stack([xs.reshape(-1), ys.reshape(-1)], -1))
    ref = torch.cat(ref, 0)[None, :, None, :].expand(frames, s, 2, 2).contiguous()

    def grads(sel):
        layer.zero_grad(set_to_none=True)
        out = layer(src[sel], pos[sel], ref[sel], st, ls, None)
        out.square().mean().backward()
        return [p.grad.detach().clone() for p in layer.parameters()]

    whole = grads(slice(0, frames))
    a, b = grads(slice(0, frames // 2)), grads(slice(frames // 2, frames))
    for gw, ga, gb, (name, _) in zip(whole, a, b, layer.named_parameters()):
        avg = (ga + gb) / 2
        scale = gw.abs().max().clamp_min(1e-12)
        assert float((avg - gw).abs().max() / scale) <= 2e-5, name
