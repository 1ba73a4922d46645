"""GPU: GraphedTrainStep (forward + backward replayed from a CUDA graph, gradients in one flat
buffer) takes the same optimisation step as the eager loop."""
import copy

import pytest
import torch

from dfvod_b200 import data_parallel
from dfvod_b200 import transformer_layers as tl

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_graphed_step_matches_eager():
    torch.manual_seed(0)
    shapes = [(12, 16), (6, 8)]
    s = sum(h * w for h, w in shapes)
    st = torch.as_tensor(shapes, dtype=torch.long, device=DEV)
    ls = torch.as_tensor([0, 12 * 16], dtype=torch.long, device=DEV)
    enc = tl.DeformableTransformerEncoder(tl.DeformableTransformerEncoderLayer(256, 512, 0.0, "relu", 2, 8, 4), 2).to(DEV)
    with torch.no_grad():
        for p in enc.parameters():
            p.add_(torch.randn_like(p) * 0.02)
    enc2 = copy.deepcopy(enc)
    src = torch.randn(2, s, 256, device=DEV)
    pos = torch.randn(2, s, 256, device=DEV)
    vr = torch.ones(2, 2, 2, device=DEV)
    loss_of = lambda m: m(src, st, ls, vr, pos, None).square().mean()

    opt = torch.optim.SGD(enc.parameters(), lr=0.5)
    losses = []
    for _ in range(2):
        opt.zero_grad(set_to_none=True)
        loss = loss_of(enc)
        loss.backward()
        opt.step()
        losses.append(float(loss))

    opt2 = torch.optim.SGD(enc2.parameters(), lr=0.5)
    step = data_parallel.GraphedTrainStep(enc2, opt2, lambda: loss_of(enc2), warmup=2)
    assert step.gradient_bytes == sum(p.numel() * 4 for p in enc2.parameters())
    glosses = [float(step()) for _ in range(2)]
    torch.cuda.synchronize()
    for a, b in zip(losses, glosses):
        assert abs(a - b) <= 1e-5 * abs(a)
    for (name, p), q in zip(enc.named_parameters(), enc2.parameters()):
        err = float((p - q).abs().max() / p.abs().max().clamp_min(1e-12))
        assert err <= 2e-5, (name, err)


def test_graphed_inference_matches_eager_and_takes_new_inputs():
    """GraphedInference replays the whole 2+2 transformer (bf16, all fused kernels) from one CUDA graph: same
    numbers as the eager call, for the captured inputs and for new ones copied into the static buffers."""
    from dfvod_b200.deformable_transformer import DeformableTransformer
    torch.manual_seed(3)
    shapes = [(20, 30), (10, 15)]
    model = DeformableTransformer(num_encoder_layers=2, num_decoder_layers=2, num_feature_levels=2,
                                  return_intermediate_dec=True).to(DEV).eval().bfloat16()
    mk = lambda: [torch.randn(2, 256, h, w, device=DEV).bfloat16() for h, w in shapes]
    srcs, poss = mk(), mk()
    masks = [torch.zeros(2, h, w, dtype=torch.bool, device=DEV) for h, w in shapes]
    query = torch.randn(50, 512, device=DEV).bfloat16()
    call = lambda: model(srcs, masks, poss, None, None, None, query)[0]
    with torch.no_grad():
        eager = call().clone()
    run = data_parallel.GraphedInference(call, inputs=(srcs, poss))
    assert torch.equal(run(), eager)
    new_srcs, new_poss = mk(), mk()
    got = run(new_srcs, new_poss).clone()
    with torch.no_grad():
        for dst, src in zip(srcs + poss, new_srcs + new_poss):
            dst.copy_(src)
        want = call()
    assert torch.equal(got, want) and not torch.equal(got, eager)
    with pytest.raises(ValueError, match="expected 4 input tensors"):
        run(new_srcs)
