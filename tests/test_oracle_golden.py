"""CPU: the three oracle restatements against vectors produced by the real reference
(tests/golden/op_*.npz, written by oracle/gen_golden.py) and against the known-answer
vector of the reference test recipe (models/ops/test.py:21-36, seed 3)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import c_oracle, msda_oracle

GOLD = os.path.join(os.path.dirname(__file__), "golden")
OP_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "op_*.npz")))

KNOWN_SEED3 = [0.001899378416, 0.004602827533, 0.004671175247, 0.004384399819,
               0.003795097174, 0.002512764199, 0.001844426151, 0.003634679248]


def load(name):
    return dict(np.load(os.path.join(GOLD, name + ".npz")))


def tol(dtype):
    return dict(rtol=1e-9, atol=1e-13) if dtype == np.float64 else dict(rtol=2e-4, atol=2e-7)


def test_have_cases():
    assert "op_toy_seed3" in OP_CASES and len(OP_CASES) >= 6


def test_known_answer_reference_recipe():
    """Regenerate the reference test's inputs from its seed and check the survey's digits."""
    torch.manual_seed(3)
    shapes = [(6, 4), (3, 2)]
    value = torch.rand(1, 30, 2, 2) * 0.01
    loc = torch.rand(1, 2, 2, 2, 2, 2)
    attn = torch.rand(1, 2, 2, 2, 2) + 1e-5
    attn /= attn.sum(-1, keepdim=True).sum(-2, keepdim=True)
    assert abs(float(value[0, 0, 0, 0]) - 4.263520168e-05) < 1e-12
    out = msda_oracle.core_pytorch(value.double(), shapes, loc.double(), attn.double())
    np.testing.assert_allclose(out.numpy().ravel(), KNOWN_SEED3, rtol=0, atol=1e-11)
    lsi = msda_oracle.level_start_index_of(shapes)
    out_c = c_oracle.forward(value.double().numpy(), shapes, lsi, loc.double().numpy(), attn.double().numpy())
    np.testing.assert_allclose(out_c.ravel(), KNOWN_SEED3, rtol=0, atol=1e-11)
    out_np = msda_oracle.forward_np(value.double().numpy(), shapes, lsi, loc.double().numpy(), attn.double().numpy())
    np.testing.assert_allclose(out_np.ravel(), KNOWN_SEED3, rtol=0, atol=1e-11)


@pytest.mark.parametrize("name", OP_CASES)
def test_core_pytorch_matches_reference(name):
    g = load(name)
    t = {k: torch.from_numpy(v) for k, v in g.items()}
    out, gv, gl, ga = msda_oracle.core_pytorch_fwd_bwd(t["value"], g["shapes"], t["loc"], t["attn"], t["grad_out"])
    k = tol(g["value"].dtype)
    np.testing.assert_allclose(out.numpy(), g["out"], **k)
    np.testing.assert_allclose(gv.numpy(), g["grad_value"], **k)
    np.testing.assert_allclose(gl.numpy(), g["grad_loc"], **k)
    np.testing.assert_allclose(ga.numpy(), g["grad_attn"], **k)


@pytest.mark.parametrize("impl", ["numpy", "c"])
@pytest.mark.parametrize("name", OP_CASES)
def test_kernel_arithmetic_restatements_match_reference(name, impl):
    """The scalar CUDA-kernel arithmetic (forward AND the explicit gradient formulas)
    equals grid_sample + autograd, including out-of-range samples."""
    g = load(name)
    args = (g["value"], g["shapes"], g["level_start_index"], g["loc"], g["attn"])
    if impl == "numpy":
        out = msda_oracle.forward_np(*args)
        gv, gl, ga = msda_oracle.backward_np(*args, g["grad_out"])
    else:
        out = c_oracle.forward(*args)
        gv, gl, ga = c_oracle.backward(*args, g["grad_out"])
    k = tol(g["value"].dtype)
    np.testing.assert_allclose(out, g["out"], **k)
    np.testing.assert_allclose(gv, g["grad_value"], **k)
    np.testing.assert_allclose(gl, g["grad_loc"], **k)
    np.testing.assert_allclose(ga, g["grad_attn"], **k)


def test_c_oracle_thread_count_independent():
    g = load("op_d32_l4")
    args = (g["value"], g["shapes"], g["level_start_index"], g["loc"], g["attn"])
    c_oracle.set_threads(1)
    a = c_oracle.backward(*args, g["grad_out"])
    c_oracle.set_threads(5)
    b = c_oracle.backward(*args, g["grad_out"])
    c_oracle.set_threads(0)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


@pytest.mark.parametrize("aligned", [True, False])
@pytest.mark.parametrize("sampling_ratio", [2, 0, 3])
def test_roi_align_oracle_pinned_to_torchvision(aligned, sampling_ratio):
    """mmcv is absent (and the reference has no vector for its RoIAlign), so the restatement of its published
    algorithm is pinned against torchvision.ops.roi_align on CPU: same Detectron algorithm, same ``aligned`` switch,
    forward and gradient, incl. boxes hanging over / outside the map and a zero-area box."""
    import torchvision
    from oracle import roi_align_oracle
    torch.manual_seed(0)
    feat = torch.randn(2, 5, 9, 13, dtype=torch.float64)
    rois = torch.tensor([[0, 10., 20., 200., 150.], [1, -50., -30., 100., 400.], [0, 300., 100., 310., 104.],
                         [1, 0., 0., 416., 288.], [1, 500., 500., 600., 600.], [0, 5., 5., 5., 5.]], dtype=torch.float64)
    f1 = feat.clone().requires_grad_(True)
    f2 = feat.clone().requires_grad_(True)
    a = roi_align_oracle.roi_align(f1, rois, 7, 1 / 32, sampling_ratio, aligned)
    b = torchvision.ops.roi_align(f2, rois, 7, 1 / 32, sampling_ratio, aligned)
    assert float((a - b).abs().max()) <= 1e-13
    g = torch.randn(a.shape, dtype=torch.float64)
    a.backward(g)
    b.backward(g)
    assert float((f1.grad - f2.grad).abs().max()) <= 1e-12
    tok = roi_align_oracle.roi_align_tokens(feat.flatten(2).transpose(1, 2), rois, 9, 13, 7, 1 / 32, sampling_ratio, aligned)
    assert torch.equal(tok, a.detach().flatten(2).transpose(1, 2))
