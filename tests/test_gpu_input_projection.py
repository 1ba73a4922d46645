"""GPU: token-major GroupNorm kernels (C ABI msda_layer_group_norm_tokens) and the InputProjection module against
the reference's nn.Sequential(Conv2d, GroupNorm(32, hidden)) composition
(/root/reference/models/deformable_detr_single.py:101-150) evaluated in fp64.
Tolerances (normalised max error): fp32 1e-5, bf16 / fp16 2^-7 for the norm alone; the projection + norm in bf16 is
bounded by the bf16 GEMM (2^-6)."""
import pytest
import torch
import torch.nn.functional as F

from dfvod_b200.input_projection import InputProjection, group_norm_tokens

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = {torch.float32: 1e-5, torch.bfloat16: 2.0 ** -7, torch.float16: 2.0 ** -9}


def nerr(x, ref):
    return float((x.double() - ref).abs().max() / ref.abs().max().clamp_min(1e-300))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("n,s,c,g", [(3, 4200, 256, 32), (2, 273, 256, 32), (1, 1, 256, 32), (2, 1050, 512, 32),
                                     (2, 77, 64, 4)])
def test_group_norm_tokens(dtype, n, s, c, g):
    torch.manual_seed(s)
    x = (torch.randn(n, s, c, device=DEV) * 1.7 + torch.randn(1, 1, c, device=DEV)).to(dtype)
    gamma = (1 + 0.3 * torch.randn(c, device=DEV)).to(dtype)
    beta = (0.3 * torch.randn(c, device=DEV)).to(dtype)
    ref = F.group_norm(x.double().transpose(1, 2), g, gamma.double(), beta.double(), 1e-5).transpose(1, 2)
    keep = x.clone()
    y = group_norm_tokens(x, g, gamma, beta, 1e-5)
    assert torch.equal(x, keep) and y.shape == x.shape and y.dtype == dtype
    assert nerr(y, ref) <= TOL[dtype]
    z = group_norm_tokens(x, g, gamma, beta, 1e-5, inplace=True)
    assert z.data_ptr() == x.data_ptr() and torch.equal(z, y)


def test_group_norm_tokens_fallbacks():
    x = torch.randn(2, 9, 48, device=DEV)                       # 48 / 32 groups: not a whole vector per group
    w, b = torch.ones(48, device=DEV), torch.zeros(48, device=DEV)
    ref = F.group_norm(x.transpose(1, 2), 16, w, b, 1e-5).transpose(1, 2)
    assert torch.equal(group_norm_tokens(x, 16, w, b), ref)
    xg = torch.randn(2, 9, 256, device=DEV, requires_grad=True)   # gradient needed -> PyTorch op
    wg = torch.ones(256, device=DEV, requires_grad=True)
    group_norm_tokens(xg, 32, wg, torch.zeros(256, device=DEV)).square().sum().backward()
    assert xg.grad is not None and wg.grad is not None


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 2.0 ** -6)])
@pytest.mark.parametrize("kwargs", [dict(kernel_size=1), dict(kernel_size=3, stride=2, padding=1)])
def test_input_projection_tokens(dtype, tol, kwargs):
    torch.manual_seed(2)
    prev = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    try:
        proj = InputProjection(96, 256, **kwargs).to(DEV)
        with torch.no_grad():
            for prm in proj.parameters():
                prm.add_(torch.randn_like(prm) * 0.1)
        ref = torch.nn.Sequential(torch.nn.Conv2d(96, 256, **kwargs), torch.nn.GroupNorm(32, 256)).to(DEV).double()
        proj = proj.to(dtype)
        ref.load_state_dict({k: v.double() for k, v in proj.state_dict().items()})
        x = torch.randn(2, 96, 25, 42, device=DEV).to(dtype)
        want = ref(x.double())
        with torch.no_grad():
            tokens, (h, w) = proj.forward_tokens(x)
            assert (h, w) == tuple(want.shape[2:]) and tokens.shape == (2, h * w, 256)
            assert nerr(tokens, want.flatten(2).transpose(1, 2)) <= tol
            assert nerr(proj(x), want) <= tol                   # the NCHW drop-in forward
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = prev


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 2.0 ** -6)])
def test_project_levels_writes_one_token_tensor(dtype, tol):
    """Three 1x1 levels + the extra 3x3 stride-2 level of a 4-level Deformable DETR (deformable_detr_single.py:101-117,
    :262-281), every level written into its slice of ONE [N, S, 256] tensor (bias folded into the GroupNorm kernels,
    strided in-place normalisation) == the reference composition + flatten / transpose / cat."""
    from dfvod_b200.input_projection import project_levels
    torch.manual_seed(4)
    prev = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    try:
        chans = [(64, 25, 42), (128, 13, 21), (160, 7, 11)]
        projs = [InputProjection(c, 256) for c, _, _ in chans] + [InputProjection(160, 256, 3, 2, 1)]
        projs = torch.nn.ModuleList(projs).to(DEV)
        with torch.no_grad():
            for prm in projs.parameters():
                prm.add_(torch.randn_like(prm) * 0.1)
        projs = projs.to(dtype)
        feats = [torch.randn(2, c, h, w, device=DEV).to(dtype) for c, h, w in chans]
        feats.append(feats[-1])                                   # the extra level projects the last backbone map
        refs = []
        for p, x in zip(projs, feats):
            r = torch.nn.Sequential(torch.nn.Conv2d(x.shape[1], 256, p[0].kernel_size, p[0].stride, p[0].padding),
                                    torch.nn.GroupNorm(32, 256)).to(DEV).double()
            r.load_state_dict({k: v.double() for k, v in p.state_dict().items()})
            refs.append(r(x.double()).flatten(2).transpose(1, 2))
        want = torch.cat(refs, 1)
        with torch.no_grad():
            tokens, views, shapes = project_levels(projs, feats)
        assert shapes == [(25, 42), (13, 21), (7, 11), (4, 6)]
        assert tokens.shape == want.shape and tokens.is_contiguous()
        assert all(v.data_ptr() >= tokens.data_ptr() for v in views)
        assert nerr(tokens, want) <= tol
        # training: same numbers through the autograd-visible composition
        tokens_g, _, _ = project_levels(projs, feats)
        assert tokens_g.requires_grad and nerr(tokens_g.detach(), want) <= tol
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = prev
