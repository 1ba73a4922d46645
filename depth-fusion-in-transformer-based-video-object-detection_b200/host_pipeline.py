"""Host-buffer entry point of the op: forward + backward on pinned HOST tensors, with the
host<->device copies overlapped with the kernels.

The reference's op only takes device tensors (cuda/ms_deform_attn_cuda.cu:34-38); a caller whose
data lives in host memory pays H2D of the four inputs and D2H of the four results around every
call.  Every MSDA call is independent per batch element (the kernel only offsets into ``value`` by
the batch index, cuda/ms_deform_im2col_cuda.cuh:263,269), so the batch is cut into frame chunks
that flow through a three-stage pipeline on three CUDA streams

    copy-in stream : H2D  value / loc / attn / grad_out of chunk i+1
    compute stream : msda_forward + msda_backward (C ABI) of chunk i
    copy-out stream: D2H  out / grad_value / grad_loc / grad_attn of chunk i-1

with a ring of device staging buffers guarded by events.  PCIe is full duplex, so the step costs
about max(H2D, D2H) instead of H2D + compute + D2H.  Consecutive calls pipeline into each other: the
inputs are HOST tensors (ready when the call is made), so the copy-in of call k+1 does not wait for the
copy-out of call k; only the ring slots are re-used under their events.
"""
import torch

from . import _lib
from .MultiScaleDeformableAttention import _DTYPES


def bind_host_thread_near_device(device_index):
    """Pin the calling process to the CPU cores NVML reports as local to the GPU (its NUMA node), so that pinned
    host buffers allocated afterwards are first-touched on that node and the H2D / D2H copies do not cross sockets.
    One process per GPU (torchrun) should call this before allocating its pinned buffers.  Returns the number of
    cores bound to, or 0 when NVML / the affinity call is unavailable (nothing is changed then)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        physical = device_index
        if vis:
            try:
                physical = int(vis.split(",")[device_index])
            except Exception:
                physical = device_index
        handle = pynvml.nvmlDeviceGetHandleByIndex(physical)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (n_cpu + 63) // 64)
        local = {i for i in range(n_cpu) if (words[i // 64] >> (i % 64)) & 1}
        cpus = sorted(local & os.sched_getaffinity(0))
        if not cpus:
            return 0
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


class HostPipelinedMSDA:
    def __init__(self, device, spatial_shapes, level_start_index, n_heads, head_dim, n_points, num_query,
                 dtype=torch.float32, chunk_frames=1, depth=3):
        self.device = torch.device(device)
        self.shapes = spatial_shapes.to(self.device, torch.long).contiguous()
        self.lsi = level_start_index.to(self.device, torch.long).contiguous()
        hw = spatial_shapes.tolist()
        self.S = int(sum(h * w for h, w in hw))
        self.L, self.M, self.D, self.P, self.Lq = len(hw), n_heads, head_dim, n_points, num_query
        self.dtype, self.chunk, self.depth = dtype, chunk_frames, depth
        c, dev = chunk_frames, self.device
        f32 = torch.float32
        self.ring = []
        for _ in range(depth):
            self.ring.append(dict(
                value=torch.empty(c, self.S, self.M, self.D, dtype=dtype, device=dev),
                loc=torch.empty(c, self.Lq, self.M, self.L, self.P, 2, dtype=f32, device=dev),
                attn=torch.empty(c, self.Lq, self.M, self.L, self.P, dtype=f32, device=dev),
                gout=torch.empty(c, self.Lq, self.M * self.D, dtype=dtype, device=dev),
                out=torch.empty(c, self.Lq, self.M * self.D, dtype=dtype, device=dev),
                gv=torch.empty(c, self.S, self.M, self.D, dtype=dtype, device=dev),
                gl=torch.empty(c, self.Lq, self.M, self.L, self.P, 2, dtype=f32, device=dev),
                ga=torch.empty(c, self.Lq, self.M, self.L, self.P, dtype=f32, device=dev),
                accum=(torch.empty(c, self.S, self.M, self.D, dtype=f32, device=dev)
                       if dtype in (torch.bfloat16, torch.float16) else None),
                loaded=torch.cuda.Event(), computed=torch.cuda.Event(), drained=torch.cuda.Event()))
        self.s_in = torch.cuda.Stream(self.device)
        self.s_run = torch.cuda.Stream(self.device)
        self.s_out = torch.cuda.Stream(self.device)
        self.launches = 0
        self._uses = 0                 # chunks pushed through the ring so far (across calls)

    def forward_backward(self, value, loc, attn, grad_out, out, grad_value, grad_loc, grad_attn):
        """All arguments are pinned host tensors shaped like the op's device tensors with batch N;
        the last four are written.  Returns after everything has been enqueued; call
        ``synchronize()`` (or record an event on ``self.s_out``) before reading the results."""
        for t in (value, loc, attn, grad_out, out, grad_value, grad_loc, grad_attn):
            if t.is_cuda or not t.is_pinned():
                raise RuntimeError("HostPipelinedMSDA takes pinned host tensors")
        lib = _lib.load()
        n = value.shape[0]
        cur = torch.cuda.current_stream(self.device)
        n_chunks = (n + self.chunk - 1) // self.chunk
        for i in range(n_chunks):
            a, b = i * self.chunk, min(n, (i + 1) * self.chunk)
            k = b - a
            use = self._uses
            self._uses += 1
            slot = self.ring[use % self.depth]
            with torch.cuda.stream(self.s_in):
                if use >= self.depth:
                    self.s_in.wait_event(slot["drained"])          # slot's previous results are out
                slot["value"][:k].copy_(value[a:b], non_blocking=True)
                slot["loc"][:k].copy_(loc[a:b], non_blocking=True)
                slot["attn"][:k].copy_(attn[a:b], non_blocking=True)
                slot["gout"][:k].copy_(grad_out[a:b], non_blocking=True)
                slot["loaded"].record(self.s_in)
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(slot["loaded"])
                stream = self.s_run.cuda_stream
                dt = _DTYPES[self.dtype]
                code = lib.msda_forward(dt, slot["value"].data_ptr(), self.shapes.data_ptr(), self.lsi.data_ptr(),
                                        slot["loc"].data_ptr(), slot["attn"].data_ptr(), k, self.S, self.M, self.D,
                                        self.L, self.Lq, self.P, slot["out"].data_ptr(), 0, stream)
                _lib.check(code, "msda_forward")
                code = lib.msda_backward(dt, slot["gout"].data_ptr(), slot["value"].data_ptr(), self.shapes.data_ptr(),
                                         self.lsi.data_ptr(), slot["loc"].data_ptr(), slot["attn"].data_ptr(), k,
                                         self.S, self.M, self.D, self.L, self.Lq, self.P, slot["gv"].data_ptr(),
                                         slot["gl"].data_ptr(), slot["ga"].data_ptr(),
                                         slot["accum"].data_ptr() if slot["accum"] is not None else None, 0, stream)
                _lib.check(code, "msda_backward")
                self.launches += 2 if slot["accum"] is None else 3
                slot["computed"].record(self.s_run)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(slot["computed"])
                out[a:b].copy_(slot["out"][:k], non_blocking=True)
                grad_value[a:b].copy_(slot["gv"][:k], non_blocking=True)
                grad_loc[a:b].copy_(slot["gl"][:k], non_blocking=True)
                grad_attn[a:b].copy_(slot["ga"][:k], non_blocking=True)
                slot["drained"].record(self.s_out)
        cur.wait_stream(self.s_out)

    def synchronize(self):
        self.s_out.synchronize()
