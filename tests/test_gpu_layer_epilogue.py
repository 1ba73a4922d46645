"""GPU: the fused layer-epilogue kernels (C ABI msda_layer_*) against a plain PyTorch fp64
composition of the reference's element-wise chain
    norm(residual + act(branch)) [+ pos]      /root/reference/models/deformable_transformer_single.py:538-548, :393-400
Tolerances (normalised max error, max|x-ref| / max|ref|): fp32 1e-5, bf16 / fp16 2^-7 against the
fp64 composition evaluated on the same (rounded) inputs."""
import pytest
import torch
import torch.nn.functional as F

from dfvod_b200.ops.functions import add_layer_norm, column_sum, linear, linear_relu, zero_masked_rows_
from dfvod_b200.ops.functions import layer_epilogue_func as lef

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = {torch.float32: 1e-5, torch.bfloat16: 2.0 ** -7, torch.float16: 2.0 ** -9}


def nerr(x, ref):
    return float((x.double() - ref).abs().max() / ref.abs().max().clamp_min(1e-300))


def ref_chain(norm64, branch, residual, act, pos):
    h = branch.double()
    if act == "relu":
        h = F.relu(h)
    elif act == "gelu":
        h = F.gelu(h)
    v = h if residual is None else residual.double() + h
    y = norm64(v)
    return y, (None if pos is None else y + pos.double())


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("c", [256, 512, 1024])
@pytest.mark.parametrize("act", [None, "relu", "gelu"])
@pytest.mark.parametrize("with_res,with_pos", [(True, True), (True, False), (False, False)])
def test_add_layer_norm_forward_backward(dtype, c, act, with_res, with_pos):
    if dtype == torch.float32 and c == 1024:
        c = 128
    torch.manual_seed(c + (0 if act is None else len(act)))
    rows = (3, 347)                      # not a multiple of the warps per CTA
    norm = torch.nn.LayerNorm(c).to(DEV)
    with torch.no_grad():
        norm.weight.add_(torch.randn(c, device=DEV) * 0.3)
        norm.bias.add_(torch.randn(c, device=DEV) * 0.3)
    norm = norm.to(dtype)
    norm64 = torch.nn.LayerNorm(c).to(DEV).double()
    norm64.load_state_dict({k: v.double() for k, v in norm.state_dict().items()})
    mk = lambda: (torch.randn(*rows, c, device=DEV) * 1.5).to(dtype)
    branch, residual, pos = mk(), (mk() if with_res else None), (mk() if with_pos else None)
    leaves = [t.clone().requires_grad_(True) for t in (branch, residual, pos) if t is not None]
    it = iter(leaves)
    b = next(it)
    r = next(it) if with_res else None
    p = next(it) if with_pos else None
    assert lef.add_layer_norm_supported(b, c)
    out = add_layer_norm(norm, b, r, act, p)
    y, y_pos = out if with_pos else (out, None)

    leaves64 = [t.detach().double().requires_grad_(True) for t in leaves]
    it = iter(leaves64)
    b64 = next(it)
    r64 = next(it) if with_res else None
    p64 = next(it) if with_pos else None
    y64, y_pos64 = ref_chain(norm64, b64, r64, act, p64)
    tol = TOL[dtype]
    assert nerr(y, y64.detach()) <= tol
    if with_pos:
        assert nerr(y_pos, y_pos64.detach()) <= tol

    gy = torch.randn_like(y)
    gyp = torch.randn_like(y) if with_pos else None
    loss = (y.float() * gy.float()).sum() + ((y_pos.float() * gyp.float()).sum() if with_pos else 0)
    loss.backward()
    loss64 = (y64 * gy.double()).sum() + ((y_pos64 * gyp.double()).sum() if with_pos else 0)
    loss64.backward()
    gtol = tol * 4 if dtype != torch.float32 else 2e-5
    for got, want in zip(leaves, leaves64):
        assert nerr(got.grad, want.grad) <= gtol
    # parameter gradients are sums over 1041 rows of rounded products
    assert nerr(norm.weight.grad, norm64.weight.grad) <= (gtol if dtype == torch.float32 else 2.0 ** -6)
    assert nerr(norm.bias.grad, norm64.bias.grad) <= (gtol if dtype == torch.float32 else 2.0 ** -6)


def test_add_layer_norm_matches_torch_composition_bf16_bitwise_pos():
    """y_pos is defined on the ROUNDED y (like the unfused chain): y_pos == bf16(y) + pos exactly."""
    torch.manual_seed(0)
    norm = torch.nn.LayerNorm(256).to(DEV).bfloat16()
    b, r, p = (torch.randn(2, 100, 256, device=DEV).bfloat16() for _ in range(3))
    y, y_pos = add_layer_norm(norm, b, r, None, p)
    assert torch.equal(y_pos, y + p)


def test_add_layer_norm_unsupported_width_uses_torch():
    norm = torch.nn.LayerNorm(96).to(DEV)
    b, r = torch.randn(4, 96, device=DEV), torch.randn(4, 96, device=DEV)
    assert not lef.add_layer_norm_supported(b, 96)
    assert torch.allclose(add_layer_norm(norm, b, r), norm(b + r))


def test_add_layer_norm_empty():
    norm = torch.nn.LayerNorm(256).to(DEV)
    b = torch.zeros(0, 256, device=DEV, requires_grad=True)
    y = add_layer_norm(norm, b, None)
    assert y.shape == (0, 256)
    y.sum().backward()
    assert float(norm.weight.grad.abs().max()) == 0.0


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_zero_masked_rows(dtype):
    torch.manual_seed(1)
    n, s, c = 3, 1237, 256
    mask = torch.rand(n, s, device=DEV) < 0.3
    lin = torch.nn.Linear(c, c).to(DEV).to(dtype)
    x = torch.randn(n, s, c, device=DEV).to(dtype)
    ref = lin(x).masked_fill(mask[..., None], 0.0)
    got = zero_masked_rows_(lin(x), mask)
    assert torch.equal(got, ref)
    # gradient: masked rows receive nothing
    x1 = x.clone().requires_grad_(True)
    g = torch.randn_like(ref)
    zero_masked_rows_(lin(x1), mask).backward(g.clone())
    x2 = x.clone().requires_grad_(True)
    lin(x2).masked_fill(mask[..., None], 0.0).backward(g.clone())
    assert torch.equal(x1.grad, x2.grad)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_linear_relu(dtype):
    torch.manual_seed(2)
    lin = torch.nn.Linear(256, 1024).to(DEV).to(dtype)
    x = torch.randn(5, 77, 256, device=DEV).to(dtype).requires_grad_(True)
    out = linear_relu(lin, x)
    ref = F.relu(lin.double()(x.detach().double())) if False else None
    lin64 = torch.nn.Linear(256, 1024).to(DEV).double()
    lin64.load_state_dict({k: v.double() for k, v in lin.state_dict().items()})
    x64 = x.detach().double().requires_grad_(True)
    ref = F.relu(lin64(x64))
    tol = 1e-5 if dtype == torch.float32 else 2.0 ** -7
    assert nerr(out, ref.detach()) <= tol
    g = torch.randn_like(out)
    out.backward(g)
    ref.backward(g.double())
    assert nerr(x.grad, x64.grad) <= tol * 4
    assert nerr(lin.weight.grad, lin64.weight.grad) <= tol * 4
    assert nerr(lin.bias.grad, lin64.bias.grad) <= tol * 4


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("rows,c", [(1041, 256), (5, 384), (88892, 1024), (300, 128), (7, 2048), (33, 96)])
def test_column_sum(dtype, rows, c):
    torch.manual_seed(rows + c)
    x = torch.randn(rows, c, device=DEV).to(dtype)
    got = column_sum(x)
    ref = x.double().sum(0)
    # the result is rounded once to the storage type; accumulation is fp32
    tol = 1e-5 if dtype == torch.float32 else (2.0 ** -8 if dtype == torch.bfloat16 else 2.0 ** -10)
    assert got.shape == (c,) and got.dtype == dtype
    assert float((got.double() - ref).abs().max() / ref.abs().max()) <= tol


def test_column_sum_unsupported_width_uses_torch():
    x = torch.randn(10, 30, device=DEV)
    assert torch.allclose(column_sum(x), x.sum(0))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_linear_custom_backward(dtype):
    torch.manual_seed(3)
    lin = torch.nn.Linear(256, 384).to(DEV).to(dtype)
    x = torch.randn(4, 513, 256, device=DEV).to(dtype).requires_grad_(True)
    out = linear(lin, x)
    g = torch.randn_like(out)
    out.backward(g)
    got = (x.grad.clone(), lin.weight.grad.clone(), lin.bias.grad.clone())
    x.grad = None
    lin.zero_grad()
    ref_out = lin(x)
    ref_out.backward(g)
    assert torch.equal(out, ref_out)
    assert torch.equal(got[0], x.grad) and torch.equal(got[1], lin.weight.grad)
    tol = 1e-5 if dtype == torch.float32 else 2.0 ** -7
    assert nerr(got[2], lin.bias.grad.double()) <= tol
    # no autograd: the plain module call -- bit for bit in "library" mode; fp32 inference defaults to the fp32-grade
    # tensor-core product (set_fp32_gemm_mode, 2052 rows >= TF32X3_MIN_ROWS), equal within the SGEMM's own error
    from dfvod_b200.ops.functions import layer_epilogue_func as L
    with torch.no_grad():
        prev = L.set_fp32_gemm_mode("library")
        try:
            assert torch.equal(linear(lin, x), ref_out)
        finally:
            L.set_fp32_gemm_mode(prev)
        # bf16 inference takes csrc/linear_bf16.cu from BF16_KERNEL_MIN_ROWS rows up: same fp32 accumulation, another
        # summation order, so a result can land on the neighbouring bf16 value
        assert nerr(linear(lin, x), ref_out.double()) <= (2e-6 if dtype == torch.float32 else 2.0 ** -7)


@pytest.mark.parametrize("rows,f,with_pos", [(1000, 1024, True), (128, 64, False), (37, 256, True), (5000, 512, False)])
def test_ffn_layer_norm_tcgen05(rows, f, with_pos):
    """The tensor-core feed-forward block (csrc/ffn_fused.cu) against the fp64 composition
    norm(x + linear2(relu(linear1(x)))) [+ pos] on the same bf16 parameters and inputs.  Stated bf16
    tolerance: normalised max error 2^-6, relative L2 2^-7 (the hidden activation and the
    pre-norm sum are rounded to bf16 once each, like the unfused bf16 chain)."""
    from dfvod_b200.ops.functions import ffn_layer_norm, ffn_layer_norm_supported
    torch.manual_seed(rows + f)
    c = 256
    lin1 = torch.nn.Linear(c, f).to(DEV).bfloat16()
    lin2 = torch.nn.Linear(f, c).to(DEV).bfloat16()
    norm = torch.nn.LayerNorm(c).to(DEV)
    with torch.no_grad():
        norm.weight.add_(torch.randn(c, device=DEV) * 0.3)
        norm.bias.add_(torch.randn(c, device=DEV) * 0.3)
    norm = norm.bfloat16()
    x = torch.randn(rows, c, device=DEV).bfloat16()
    pos = torch.randn(rows, c, device=DEV).bfloat16() if with_pos else None
    with torch.no_grad():
        assert ffn_layer_norm_supported(x, lin1, lin2, norm)
        out = ffn_layer_norm(lin1, lin2, norm, x, pos)
        y, y_pos = out if with_pos else (out, None)
        d = lambda m: {k: v.double() for k, v in m.state_dict().items()}
        l1, l2, nm = torch.nn.Linear(c, f).to(DEV).double(), torch.nn.Linear(f, c).to(DEV).double(), \
            torch.nn.LayerNorm(c).to(DEV).double()
        l1.load_state_dict(d(lin1)); l2.load_state_dict(d(lin2)); nm.load_state_dict(d(norm))
        ref = nm(x.double() + l2(F.relu(l1(x.double()))))
    e = (y.double() - ref).abs().max() / ref.abs().max()
    l2err = (y.double() - ref).norm() / ref.norm()
    assert float(e) <= 2.0 ** -6 and float(l2err) <= 2.0 ** -7, (float(e), float(l2err))
    if with_pos:
        assert torch.equal(y_pos, y + pos)
    # training (autograd) never takes this kernel
    x.requires_grad_(True)
    assert not ffn_layer_norm_supported(x, lin1, lin2, norm)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("with_add", [False, True])
@pytest.mark.parametrize("c", [96, 256, 6])
def test_flatten_levels_bitwise(dtype, with_add, c):
    """msda_layer_flatten_level == flatten(2).transpose(1, 2) (+ level_embed[l]) + cat of
    /root/reference/models/deformable_transformer_single.py:190-206, bit for bit (a copy and one rounded add)."""
    from dfvod_b200.ops.functions import flatten_levels, flatten_levels_supported
    torch.manual_seed(11)
    n = 3            # 96: partial channel tile of the 16-byte-vector kernel; 6: no whole vector -> the 32 x 32 tile kernel
    shapes = [(13, 21), (7, 11), (1, 1), (5, 33)]
    maps = [torch.randn(n, c, h, w, device=DEV).to(dtype) for h, w in shapes]
    adds = [torch.randn(c, device=DEV).to(dtype) for _ in shapes] if with_add else None
    assert flatten_levels_supported(maps)
    got = flatten_levels(maps, adds)
    want = torch.cat([m.flatten(2).transpose(1, 2) + (adds[i].view(1, 1, -1) if with_add else 0)
                      for i, m in enumerate(maps)], 1)
    assert got.shape == want.shape and got.dtype == want.dtype
    assert torch.equal(got, want)


def test_flatten_levels_refuses_when_gradients_are_needed():
    from dfvod_b200.ops.functions import flatten_levels_supported
    m = torch.randn(1, 8, 2, 2, device=DEV, requires_grad=True)
    assert not flatten_levels_supported([m])
    with torch.no_grad():
        assert flatten_levels_supported([m])
    assert not flatten_levels_supported([m.detach().double()])


def test_transformer_inference_path_equals_training_path():
    """DeformableTransformer.forward under no_grad takes the transposing-kernel flatten; it must give what the
    autograd-visible PyTorch composition gives."""
    from dfvod_b200.deformable_transformer import DeformableTransformer
    torch.manual_seed(5)
    shapes = [(12, 17), (6, 9)]
    model = DeformableTransformer(d_model=64, nhead=4, num_encoder_layers=2, num_decoder_layers=2, dim_feedforward=128,
                                  dropout=0.0, num_feature_levels=2, return_intermediate_dec=True).to(DEV).eval()
    srcs = [torch.randn(2, 64, h, w, device=DEV) for h, w in shapes]
    poss = [torch.randn(2, 64, h, w, device=DEV) for h, w in shapes]
    masks = [torch.zeros(2, h, w, dtype=torch.bool, device=DEV) for h, w in shapes]
    masks[0][1, :, -3:] = True
    masks[1][1, :, -2:] = True
    query = torch.randn(10, 128, device=DEV)
    hs_train = model(srcs, masks, poss, None, None, None, query)[0]
    with torch.no_grad():
        hs_infer = model(srcs, masks, poss, None, None, None, query)[0]
    assert nerr(hs_infer, hs_train.detach().double()) <= 1e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("row_bytes", [128, 256, 512, 1024, 2048])
@pytest.mark.parametrize("act", [None, "relu", "gelu"])
def test_norm_act_forward(dtype, row_bytes, act):
    """msda_layer_norm_act_forward = act(LayerNorm(x)) of the dynamic interaction head
    (/root/reference/models/sparse_roi_head/head.py:156-170) vs the fp64 composition; rows not a multiple of the
    rows a warp / CTA owns; out of place and in place."""
    from dfvod_b200.ops.functions import norm_act
    c = row_bytes // torch.empty((), dtype=dtype).element_size()
    torch.manual_seed(c)
    norm = torch.nn.LayerNorm(c).to(DEV)
    with torch.no_grad():
        norm.weight.add_(torch.randn(c, device=DEV) * 0.3)
        norm.bias.add_(torch.randn(c, device=DEV) * 0.3)
    norm = norm.to(dtype)
    x = (torch.randn(7, 53, c, device=DEV) * 1.5 + 0.3).to(dtype)
    norm64 = torch.nn.LayerNorm(c).to(DEV).double()
    norm64.load_state_dict({k: v.double() for k, v in norm.state_dict().items()})
    ref = norm64(x.double())
    ref = ref if act is None else getattr(F, act)(ref)
    with torch.no_grad():
        before = x.clone()
        y = norm_act(norm, x, act)
        assert torch.equal(x, before) and y.data_ptr() != x.data_ptr()
        assert nerr(y, ref) <= TOL[dtype]
        z = norm_act(norm, x, act, inplace=True)
        assert z.data_ptr() == x.data_ptr() and torch.equal(z, y)


def test_norm_act_keeps_autograd_and_odd_widths():
    from dfvod_b200.ops.functions import norm_act
    norm = torch.nn.LayerNorm(48).to(DEV)
    x = torch.randn(5, 48, device=DEV, requires_grad=True)
    y = norm_act(norm, x, "relu")                       # gradient needed -> PyTorch composition
    y.sum().backward()
    assert x.grad is not None and norm.weight.grad is not None
    with torch.no_grad():
        assert torch.equal(norm_act(norm, x, "relu"), F.relu(norm(x)))      # 192-byte rows: composition as well


@pytest.mark.parametrize("rows", [1, 127, 128, 129, 5 * 128 + 77, 148 * 128 * 2 + 31])
@pytest.mark.parametrize("with_res,with_pos", [(True, False), (True, True), (False, False)])
def test_proj_layer_norm_tcgen05(rows, with_res, with_pos):
    """msda_layer_proj_layernorm_forward (weight-stationary tcgen05 kernel, csrc/proj_fused.cu) =
    norm(residual + linear(x)) [+ pos] of ms_deform_attn.py:116 + deformable_transformer_single.py:538-541, against
    the fp64 composition on the same bf16 inputs.  Tolerance 2^-7 normalised max (bf16 output)."""
    from dfvod_b200.ops.functions import proj_layer_norm, proj_layer_norm_supported
    torch.manual_seed(rows)
    c = 256
    lin = torch.nn.Linear(c, c).to(DEV)
    norm = torch.nn.LayerNorm(c).to(DEV)
    with torch.no_grad():
        lin.weight.mul_(2.0)
        lin.bias.add_(torch.randn(c, device=DEV) * 0.2)
        norm.weight.add_(torch.randn(c, device=DEV) * 0.3)
        norm.bias.add_(torch.randn(c, device=DEV) * 0.3)
    lin, norm = lin.bfloat16(), norm.bfloat16()
    x = torch.randn(rows, c, device=DEV).bfloat16()
    res = torch.randn(rows, c, device=DEV).bfloat16() if with_res else None
    pos = torch.randn(rows, c, device=DEV).bfloat16() if with_pos else None
    with torch.no_grad():
        assert proj_layer_norm_supported(x, lin, norm)
        if with_pos:          # the Python helper prefers GEMM + fused norm when `pos` is asked for: call the kernel directly
            from dfvod_b200 import _lib
            y, y_pos = torch.empty_like(x), torch.empty_like(x)
            code = _lib.load().msda_layer_proj_layernorm_forward(
                _lib.DTYPE_BF16, x.data_ptr(), lin.weight.data_ptr(), lin.bias.data_ptr(), res.data_ptr(),
                norm.weight.data_ptr(), norm.bias.data_ptr(), pos.data_ptr(), rows, c, float(norm.eps), y.data_ptr(),
                y_pos.data_ptr(), torch.cuda.current_stream().cuda_stream)
            _lib.check(code, "msda_layer_proj_layernorm_forward")
            out = (y, y_pos)
            # and the helper's own route gives the same
            y2, y2_pos = proj_layer_norm(lin, norm, x, res, pos)
            assert nerr(y2, y.double()) <= 2.0 ** -6 and nerr(y2_pos, y_pos.double()) <= 2.0 ** -6
        else:
            out = proj_layer_norm(lin, norm, x, res, pos)
    y, y_pos = out if with_pos else (out, None)
    lin64, norm64 = torch.nn.Linear(c, c).to(DEV).double(), torch.nn.LayerNorm(c).to(DEV).double()
    lin64.load_state_dict({k: v.double() for k, v in lin.state_dict().items()})
    norm64.load_state_dict({k: v.double() for k, v in norm.state_dict().items()})
    proj = lin64(x.double()).bfloat16().double()              # the unfused chain rounds the GEMM output to bf16
    ref = norm64(proj if res is None else res.double() + proj)
    assert nerr(y, ref) <= 2.0 ** -7
    if with_pos:
        assert nerr(y_pos, ref + pos.double()) <= 2.0 ** -7


def test_proj_layer_norm_falls_back_with_gradients():
    from dfvod_b200.ops.functions import proj_layer_norm
    lin, norm = torch.nn.Linear(256, 256).to(DEV).bfloat16(), torch.nn.LayerNorm(256).to(DEV).bfloat16()
    x = torch.randn(9, 256, device=DEV).bfloat16().requires_grad_(True)
    res = torch.randn(9, 256, device=DEV).bfloat16()
    proj_layer_norm(lin, norm, x, res).float().square().sum().backward()
    assert x.grad is not None and lin.weight.grad is not None


def test_tf32x3_linear_is_fp32_grade():
    """The error-compensated three-GEMM split (ops/functions/layer_epilogue_func.py: linear_tf32x3) against fp64:
    fp32-grade (<= 2e-6 normalised, the IEEE SGEMM's own order of magnitude), where one TF32 GEMM is >= 20x worse."""
    from dfvod_b200.ops.functions import layer_epilogue_func as L
    torch.manual_seed(0)
    x = torch.randn(20000, 256, device=DEV) * 3
    w = torch.randn(1024, 256, device=DEV) / 16
    b = torch.randn(1024, device=DEV)
    ref = F.linear(x.double(), w.double(), b.double())
    prev = torch.backends.cuda.matmul.allow_tf32
    try:
        torch.backends.cuda.matmul.allow_tf32 = False
        e_ieee = nerr(F.linear(x, w, b), ref)
        torch.backends.cuda.matmul.allow_tf32 = True
        e_tf32 = nerr(F.linear(x, w, b), ref)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    e_split = nerr(L.linear_tf32x3(x, w, b), ref)
    print(f"normalised error vs fp64: ieee {e_ieee:.2e}  tf32 {e_tf32:.2e}  tf32x3 {e_split:.2e}")
    assert e_split <= 2e-6 and e_split <= 10 * e_ieee
    assert e_tf32 >= 20 * e_split
    assert torch.backends.cuda.matmul.allow_tf32 == prev          # the switch is restored


@pytest.mark.parametrize("rows,n,k,relu,has_bias", [
    (20000, 256, 256, False, True),      # value_proj / output_proj
    (1001, 384, 256, False, True),       # [offsets | logits]: a second column tile of 128, ragged rows
    (4097, 1024, 256, True, True),       # linear1 + ReLU
    (333, 256, 1024, False, True),       # linear2: 32 K blocks
    (128, 32, 32, False, False),         # smallest shape, no bias
    (1, 64, 96, True, False),
])
def test_tf32x3_kernel_matches_fp64(rows, n, k, relu, has_bias):
    """csrc/linear_tf32x3.cu (the split of the activation tile happens in shared memory) against fp64: the same
    fp32-grade bound as the split-pass route, for every tile shape the layers use."""
    from dfvod_b200.ops.functions import layer_epilogue_func as L
    torch.manual_seed(rows + n)
    x = torch.randn(rows, k, device=DEV) * 3
    w = torch.randn(n, k, device=DEV) / 16
    b = torch.randn(n, device=DEV) if has_bias else None
    assert L.linear_tf32x3_kernel_supported(x, w)
    ref = F.linear(x.double(), w.double(), None if b is None else b.double())
    if relu:
        ref = ref.relu()
    got = L.linear_tf32x3(x, w, b, relu=relu, route="kernel")
    torch.cuda.synchronize()
    assert got.shape == (rows, n) and got.dtype == torch.float32
    err = nerr(got, ref)
    # the main accumulator takes K / 8 truncating tensor-core additions (like one TF32 GEMM): measured 5.7e-7 at K = 256,
    # 3.0e-6 at K = 1024; the IEEE SGEMM itself is 7e-7 at K = 256
    assert err <= 2e-6 * max(1.0, k / 512.0), err
    assert nerr(L.linear_tf32x3(x, w, b, relu=relu, route="library"), ref) <= 2e-6 * max(1.0, k / 512.0)
    # a 3-d input with a weight the kernel does not take (K % 32 != 0) falls back to the split pass
    x3 = torch.randn(4, 300, 40, device=DEV)
    w3 = torch.randn(64, 40, device=DEV)
    assert not L.linear_tf32x3_kernel_supported(x3, w3)
    assert nerr(L.linear_tf32x3(x3, w3, None), F.linear(x3.double(), w3.double())) <= 2e-6


@pytest.mark.parametrize("rows,n,k,relu,has_bias,masked", [
    (20000, 256, 256, False, True, True),    # value_proj with its padding mask
    (1001, 384, 256, False, True, False),    # [offsets | logits]: a second column tile of 128, ragged rows
    (4097, 1024, 256, True, True, False),    # linear1 + ReLU
    (333, 256, 1024, False, True, False),    # linear2: 16 K blocks
    (128, 64, 64, False, False, False),      # smallest shape, no bias
    (1, 128, 192, True, False, True),
])
def test_bf16_linear_kernel_matches_fp64(rows, n, k, relu, has_bias, masked):
    """csrc/linear_bf16.cu against the fp64 product of the same bf16 operands: one bf16 rounding of an fp32-accumulated
    result (2^-8 of the output range), bias / ReLU / zeroed rows in the epilogue."""
    from dfvod_b200.ops.functions import layer_epilogue_func as L
    torch.manual_seed(rows + n)
    x = (torch.randn(rows, k, device=DEV) * 3).bfloat16()
    w = (torch.randn(n, k, device=DEV) / 16).bfloat16()
    b = torch.randn(n, device=DEV).bfloat16() if has_bias else None
    mask = (torch.rand(rows, device=DEV) < 0.3) if masked else None
    ref = F.linear(x.double(), w.double(), None if b is None else b.double())
    if relu:
        ref = ref.relu()
    if mask is not None:
        ref = ref.masked_fill(mask[:, None], 0.0)
    got = L.linear_bf16(x, w, b, relu=relu, zero_rows=mask)
    torch.cuda.synchronize()
    assert got.shape == (rows, n) and got.dtype == torch.bfloat16
    assert nerr(got, ref) <= 2.0 ** -8
    if mask is not None:
        assert not bool(got[mask].any())


def test_bf16_linear_routes():
    """linear / linear_relu / linear_zero_rows pick the kernel for bf16 inference from BF16_KERNEL_MIN_ROWS rows up and
    keep the library GEMM (and the custom autograd function) when gradients flow."""
    from dfvod_b200.ops.functions import layer_epilogue_func as L
    from dfvod_b200.ops.functions import linear_relu, linear_zero_rows
    torch.manual_seed(1)
    lin = torch.nn.Linear(256, 384).to(DEV).bfloat16()
    x = torch.randn(3000, 256, device=DEV).bfloat16()
    mask = torch.rand(3000, device=DEV) < 0.2
    calls = []
    original, prev_rows = L.linear_bf16, L.BF16_KERNEL_MIN_ROWS
    L.linear_bf16 = lambda *a, **kw: (calls.append(kw), original(*a, **kw))[1]
    L.BF16_KERNEL_MIN_ROWS = 1024
    try:
        with torch.no_grad():
            y = linear(lin, x)
            yr = linear_relu(lin, x)
            ym = linear_zero_rows(lin, x, mask)
            assert len(calls) == 3 and calls[1].get("relu") and calls[2].get("zero_rows") is mask
            ref = lin(x)
            assert nerr(y, ref.double()) <= 2.0 ** -7 and nerr(yr, ref.relu().double()) <= 2.0 ** -7
            assert nerr(ym, ref.masked_fill(mask[:, None], 0).double()) <= 2.0 ** -7 and not bool(ym[mask].any())
            linear(lin, x[:100])                                   # few rows: library
            assert len(calls) == 3
        xg = x.clone().requires_grad_(True)
        out = linear_zero_rows(lin, xg, mask)                     # gradients flow: library GEMM + row-zeroing function
        assert len(calls) == 3
        out.float().sum().backward()
        assert xg.grad is not None and not bool(xg.grad[mask].any())
    finally:
        L.linear_bf16 = original
        L.BF16_KERNEL_MIN_ROWS = prev_rows


@pytest.mark.parametrize("rows", [1, 127, 129, 1000])
def test_linear_kernels_write_only_their_rows(rows):
    """The two GEMM kernels store whole 32-row boxes by TMA; the tensor map clips them at `rows`.  Outputs placed in the
    middle of a sentinel-filled buffer: nothing before the first or after the last row may change (C ABI called
    directly, ragged row counts around the 128-row tile)."""
    from dfvod_b200 import _lib
    lib = _lib.load()
    stream = torch.cuda.current_stream().cuda_stream
    torch.manual_seed(rows)
    n, k, pad = 128, 64, 160
    # bf16
    x = torch.randn(rows, k, device=DEV).bfloat16()
    w = torch.randn(n, k, device=DEV).bfloat16()
    big = torch.full((pad + rows + pad, n), 7.0, device=DEV, dtype=torch.bfloat16)
    y = big[pad:pad + rows]
    _lib.check(lib.msda_layer_linear_bf16(x.data_ptr(), w.data_ptr(), None, None, rows, n, k, 0, y.data_ptr(), stream), "bf16")
    torch.cuda.synchronize()
    assert bool((big[:pad] == 7.0).all()) and bool((big[pad + rows:] == 7.0).all())
    assert nerr(y, F.linear(x.double(), w.double())) <= 2.0 ** -8
    # fp32-grade
    x32 = torch.randn(rows, k, device=DEV)
    w32 = torch.randn(n, k, device=DEV)
    hi = ((w32.view(torch.int32) + 0x1000) & -0x2000).view(torch.float32)
    lo = w32 - hi
    big32 = torch.full((pad + rows + pad, n), 7.0, device=DEV)
    y32 = big32[pad:pad + rows]
    _lib.check(lib.msda_layer_linear_tf32x3(x32.data_ptr(), hi.data_ptr(), lo.data_ptr(), None, rows, n, k, 0,
                                            y32.data_ptr(), stream), "tf32x3")
    torch.cuda.synchronize()
    assert bool((big32[:pad] == 7.0).all()) and bool((big32[pad + rows:] == 7.0).all())
    assert nerr(y32, F.linear(x32.double(), w32.double())) <= 2e-6


def test_tf32x3_mode_routes_the_layer_gemms():
    """set_fp32_gemm_mode('tf32x3'): the fp32 transformer gives the library-SGEMM result within the fp32 parity
    tolerance (1e-5 normalised); gradients-needed calls and 16-bit inputs keep the library path.  "tf32x3" is the
    package default; "library" restores the IEEE SGEMMs."""
    from dfvod_b200.deformable_transformer import DeformableTransformer
    from dfvod_b200.ops.functions import layer_epilogue_func as L
    torch.manual_seed(7)
    shapes = [(24, 31), (12, 16)]
    model = DeformableTransformer(d_model=64, nhead=4, num_encoder_layers=2, num_decoder_layers=2, dim_feedforward=128,
                                  dropout=0.0, num_feature_levels=2, return_intermediate_dec=True).to(DEV).eval()
    srcs = [torch.randn(2, 64, h, w, device=DEV) for h, w in shapes]
    poss = [torch.randn(2, 64, h, w, device=DEV) for h, w in shapes]
    masks = [torch.zeros(2, h, w, dtype=torch.bool, device=DEV) for h, w in shapes]
    query = torch.randn(10, 128, device=DEV)
    prev_tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    prev_rows = L.TF32X3_MIN_ROWS
    prev_mode = L.set_fp32_gemm_mode("library")
    assert prev_mode == "tf32x3"                                   # the package default
    try:
        with torch.no_grad():
            want = model(srcs, masks, poss, None, None, None, query)[0]
            assert L.set_fp32_gemm_mode("tf32x3") == "library"
            L.TF32X3_MIN_ROWS = 1
            calls = []
            original = L.linear_tf32x3
            L.linear_tf32x3 = lambda *a, **kw: (calls.append(1), original(*a, **kw))[1]
            try:
                got = model(srcs, masks, poss, None, None, None, query)[0]
            finally:
                L.linear_tf32x3 = original
        assert len(calls) >= 10                                   # the projections did take the split path
        assert nerr(got, want.double()) <= 1e-5
        xg = torch.randn(2048, 64, device=DEV, requires_grad=True)
        lin = torch.nn.Linear(64, 64).to(DEV)
        assert not L._tf32x3_wanted(xg, lin.weight) and not L._tf32x3_wanted(xg.detach().bfloat16(), lin.weight)
        with torch.no_grad():
            assert L._tf32x3_wanted(xg.detach(), lin.weight)
            with torch.autocast("cuda", dtype=torch.bfloat16):       # the caller asked for 16-bit GEMMs
                assert not L._tf32x3_wanted(xg.detach(), lin.weight)
                assert L.linear(lin, xg.detach()).dtype == torch.bfloat16
        with pytest.raises(ValueError):
            L.set_fp32_gemm_mode("fp8")
    finally:
        L.set_fp32_gemm_mode(prev_mode)
        L.TF32X3_MIN_ROWS = prev_rows
        torch.backends.cuda.matmul.allow_tf32 = prev_tf32


@pytest.mark.parametrize("rows,cols", [(177, 256), (5, 1024), (1, 4), (3000, 64)])
def test_tf32_split_kernel_bitwise(rows, cols):
    """msda_layer_tf32_split == the integer-arithmetic definition (round the bit pattern to 10 mantissa bits, subtract),
    bit for bit; hi is TF32-representable and lo + hi == x exactly."""
    from dfvod_b200.ops.functions import layer_epilogue_func as L
    torch.manual_seed(rows)
    x = torch.randn(rows, cols, device=DEV) * torch.exp(torch.randn(rows, cols, device=DEV) * 4)
    got = L._tf32_split(x)
    hi = ((x.view(torch.int32) + 0x1000) & -0x2000).view(torch.float32)
    want = torch.cat([x - hi, hi, hi], 1)
    assert got.shape == (rows, 3 * cols) and torch.equal(got, want)
    assert int((got[:, cols:2 * cols].contiguous().view(torch.int32) & 0x1fff).abs().max()) == 0
    assert torch.equal(got[:, :cols] + got[:, cols:2 * cols], x)
    odd = torch.randn(7, 6, device=DEV)                       # cols % 4 != 0 -> torch composition, same definition
    hi6 = ((odd.view(torch.int32) + 0x1000) & -0x2000).view(torch.float32)
    assert torch.equal(L._tf32_split(odd), torch.cat([odd - hi6, hi6, hi6], 1))
