"""GPU: the pipelined host-buffer entry point returns exactly what the device op returns."""
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu

from dfvod_b200 import MultiScaleDeformableAttention as MSDA
from dfvod_b200.host_pipeline import HostPipelinedMSDA


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("chunk", [1, 2])
def test_pipeline_equals_device_op(dtype, chunk):
    shapes = [(12, 20), (6, 10)]
    n, m, d, p = 5, 8, 32, 4                       # 5 frames: ragged last chunk when chunk=2
    s = sum(h * w for h, w in shapes)
    value, loc, attn, gout = util.make_inputs(shapes, n, m, d, s, p, seed=21, dist="grid")
    value, gout = value.to(dtype), gout.to(dtype)
    st, ls = util.shapes_tensors(shapes, "cuda")
    host = [t.pin_memory() for t in (value, loc, attn, gout)]
    outs = [torch.empty(n, s, m * d, dtype=dtype).pin_memory(), torch.empty_like(value).pin_memory(),
            torch.empty_like(loc).pin_memory(), torch.empty_like(attn).pin_memory()]
    pipe = HostPipelinedMSDA("cuda", st.cpu(), ls.cpu(), m, d, p, s, dtype=dtype, chunk_frames=chunk, depth=2)
    for _ in range(2):                              # second pass re-uses the ring
        pipe.forward_backward(*host, *outs)
    pipe.synchronize()
    v, l, a, g = (t.cuda() for t in (value, loc, attn, gout))
    ref_out = MSDA.ms_deform_attn_forward(v, st, ls, l, a, 64)
    ref_gv, ref_gl, ref_ga = MSDA.ms_deform_attn_backward(v, st, ls, l, a, g, 64)
    assert torch.equal(outs[0], ref_out.cpu())
    assert torch.equal(outs[2], ref_gl.cpu()) and torch.equal(outs[3], ref_ga.cpu())
    # grad_value is accumulated with atomics: order-dependent rounding only
    tol = 1e-5 if dtype == torch.float32 else 2 ** -7
    assert (outs[1].float() - ref_gv.cpu().float()).abs().max() <= tol * ref_gv.float().abs().max().cpu()
    with pytest.raises(RuntimeError, match="pinned host tensors"):
        pipe.forward_backward(v, *host[1:], *outs)
