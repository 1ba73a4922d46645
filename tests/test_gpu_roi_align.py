"""GPU: the token-major RoIAlign kernel (C ABI msda_roi_align_forward / _backward, csrc/roi_align.cu) that
replaces mmcv.ops.RoIAlign in the TransVOD++ temporal stage
(/root/reference/models/deformable_transformer_multi_plusplus.py:129-132, :499, :514) against the CPU oracle
(oracle/roi_align_oracle.py, itself pinned to torchvision.ops.roi_align in tests/test_oracle_golden.py).
Tolerances (normalised max error): fp64 1e-12, fp32 1e-5, bf16 / fp16 2^-7 vs the fp64 oracle on the same
(rounded) inputs."""
import pytest
import torch

from dfvod_b200.temporal_stage import RoIAlign, bbox2roi, roi_align_tokens
from oracle import roi_align_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = {torch.float64: 1e-12, torch.float32: 1e-5, torch.bfloat16: 2.0 ** -7, torch.float16: 2.0 ** -9}


def nerr(x, ref):
    return float((x.double().cpu() - ref).abs().max() / ref.abs().max().clamp_min(1e-300))


def make_rois(n, k, height, width, scale, seed):
    """Boxes in image units: ordinary ones, some hanging over the border, one fully outside, one degenerate."""
    g = torch.Generator().manual_seed(seed)
    img_w, img_h = width / scale, height / scale
    cx = torch.rand(k, generator=g) * img_w
    cy = torch.rand(k, generator=g) * img_h
    bw = torch.rand(k, generator=g) * img_w * 0.6 + 2
    bh = torch.rand(k, generator=g) * img_h * 0.6 + 2
    boxes = torch.stack([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2], -1)
    boxes[0] = torch.tensor([img_w * 1.5, img_h * 1.5, img_w * 1.8, img_h * 1.9])     # outside the map
    boxes[1] = torch.tensor([img_w * 0.3, img_h * 0.3, img_w * 0.3, img_h * 0.3])     # zero area
    boxes[2] = torch.tensor([-img_w * 0.2, -img_h * 0.1, img_w * 1.2, img_h * 1.1])   # larger than the image
    idx = torch.randint(0, n, (k, 1), generator=g).double()
    return torch.cat([idx, boxes.double()], -1)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("channels", [256, 30])          # 16-byte vector path / scalar path
@pytest.mark.parametrize("sampling_ratio,aligned,pooled", [(2, True, 7), (0, True, 7), (3, False, (3, 5))])
def test_roi_align_forward_backward(dtype, channels, sampling_ratio, aligned, pooled):
    torch.manual_seed(channels + sampling_ratio)
    n, height, width, scale, k = 3, 9, 13, 1 / 32, 23
    tokens = torch.randn(n, height * width, channels).to(dtype)
    rois = make_rois(n, k, height, width, scale, 7)
    ph, pw = (pooled, pooled) if isinstance(pooled, int) else pooled
    gout = torch.randn(k, ph * pw, channels).to(dtype)

    t64 = tokens.double().clone().requires_grad_(True)
    ref = roi_align_oracle.roi_align_tokens(t64, rois, height, width, pooled, scale, sampling_ratio, aligned)
    (ref_grad,) = torch.autograd.grad(ref, t64, gout.double())

    td = tokens.to(DEV).requires_grad_(True)
    out = roi_align_tokens(td, rois.to(DEV), height, width, pooled, scale, sampling_ratio, aligned)
    assert out.shape == (k, ph * pw, channels) and out.dtype == dtype
    out.backward(gout.to(DEV))
    assert nerr(out.detach(), ref.detach()) <= TOL[dtype]
    assert nerr(td.grad, ref_grad) <= TOL[dtype]


def test_roi_align_module_matches_torchvision_nchw():
    """mmcv-style module interface: NCHW in, [K, C, 7, 7] out, the reference's construction arguments."""
    import torchvision
    torch.manual_seed(3)
    layer = RoIAlign(output_size=7, sampling_ratio=2, spatial_scale=1 / 32)
    feat = torch.randn(2, 64, 12, 20)
    boxes = [torch.tensor([[10., 20., 300., 200.], [0., 0., 640., 384.]]), torch.tensor([[100., 50., 180., 90.]])]
    rois = bbox2roi(boxes)
    want = torchvision.ops.roi_align(feat, rois, 7, 1 / 32, 2, True)
    got = layer(feat.to(DEV), rois.to(DEV))
    assert got.shape == want.shape
    assert nerr(got, want.double()) <= 1e-5
    # a channels-last view of token-major memory goes through without a copy and gives the same
    tokens = feat.flatten(2).transpose(1, 2).contiguous().to(DEV)
    view = tokens.permute(0, 2, 1).view(2, 64, 12, 20)
    assert nerr(layer(view, rois.to(DEV)), want.double()) <= 1e-5


def test_roi_align_empty_and_errors():
    tokens = torch.randn(1, 12, 8, device=DEV)
    out = roi_align_tokens(tokens, torch.zeros(0, 5, device=DEV), 3, 4, 7, 1.0, 2, True)
    assert out.shape == (0, 49, 8)
    with pytest.raises(RuntimeError, match="do not form"):
        roi_align_tokens(tokens, torch.zeros(1, 5, device=DEV), 5, 5, 7)
    with pytest.raises(RuntimeError, match=r"\[K, 5\]"):
        roi_align_tokens(tokens, torch.zeros(1, 4, device=DEV), 3, 4, 7)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        roi_align_tokens(tokens.cpu(), torch.zeros(1, 5), 3, 4, 7)
    with pytest.raises(NotImplementedError):
        RoIAlign(7, pool_mode="max")


def test_roi_align_linear_in_the_features_at_full_size():
    """Size-independent property at the TransVOD++ size (300 boxes x 5 frames, 50 x 84 map, 256 channels):
    pooling is linear in the feature map, and its backward is the adjoint of its forward."""
    torch.manual_seed(9)
    n, height, width, c, k = 5, 50, 84, 256, 1500
    a = torch.randn(n, height * width, c, device=DEV)
    b = torch.randn(n, height * width, c, device=DEV)
    rois = make_rois(n, k, height, width, 1 / 32, 11).to(DEV)
    f = lambda t: roi_align_tokens(t, rois, height, width, 7, 1 / 32, 2, True)
    lhs = f(a * 0.5 + b * 2.0)
    rhs = f(a) * 0.5 + f(b) * 2.0
    assert float((lhs - rhs).abs().max() / rhs.abs().max()) <= 1e-5
    x = a.clone().requires_grad_(True)
    y = f(x)
    g = torch.randn_like(y)
    y.backward(g)
    dot_fwd = float((f(b).double() * g.double()).sum())
    dot_bwd = float((b.double() * x.grad.double()).sum())
    assert abs(dot_fwd - dot_bwd) <= 1e-5 * max(abs(dot_fwd), 1.0)
