"""TEST INFRASTRUCTURE ONLY -- CPU restatement of RoIAlign as the reference's TransVOD++ temporal stage uses it:
``mmcv.ops.RoIAlign(output_size=7, sampling_ratio=2, spatial_scale=1/32)`` (pool_mode 'avg', aligned=True),
constructed at /root/reference/models/deformable_transformer_multi_plusplus.py:129-132, called at :499 and :514.

mmcv-full==1.7.0 (pinned in /root/reference/supporting_files/detailed_requirements.txt:44) is a third-party
dependency that is NOT under /root/reference and not installed here, so this file restates its published
algorithm (mmcv/ops/csrc/common/cuda/roi_align_cuda_kernel.cuh: roi_align_forward_cuda_kernel, pool_mode 1, and
bilinear_interpolate of common_cuda_helper.hpp).  PINNING: the reference holds no test or golden vector for this
op; tests/test_oracle_golden.py checks this restatement against torchvision.ops.roi_align on CPU (the same
Detectron algorithm with the same ``aligned`` switch) -- the mmcv binding itself stays unpinned.

Plain torch indexing, one python iteration per RoI; differentiable w.r.t. ``feat`` so autograd through it is
the backward oracle (mmcv's backward kernel scatters ``weight * grad / count`` to the same four corners).
Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline may import this module.
"""
import math

import torch


def _interp_terms(y, x, height, width):
    """bilinear_interpolate: (valid, y_low, x_low, y_high, x_high, ly, lx) for coordinate tensors y, x."""
    valid = ~((y < -1.0) | (y > height) | (x < -1.0) | (x > width))
    y = y.clamp(min=0)
    x = x.clamp(min=0)
    y_low = y.floor().long()
    x_low = x.floor().long()
    top = y_low >= height - 1
    left = x_low >= width - 1
    y_low = torch.where(top, torch.full_like(y_low, height - 1), y_low)
    x_low = torch.where(left, torch.full_like(x_low, width - 1), x_low)
    y_high = torch.where(top, y_low, y_low + 1)
    x_high = torch.where(left, x_low, x_low + 1)
    y = torch.where(top, y_low.to(y.dtype), y)
    x = torch.where(left, x_low.to(x.dtype), x)
    return valid, y_low, x_low, y_high, x_high, y - y_low.to(y.dtype), x - x_low.to(x.dtype)


def roi_align(feat, rois, output_size, spatial_scale=1.0, sampling_ratio=0, aligned=True):
    """feat [N, C, H, W], rois [K, 5] (batch index, x1, y1, x2, y2) -> [K, C, PH, PW] (average pooling)."""
    ph, pw = (output_size, output_size) if isinstance(output_size, int) else output_size
    n, c, height, width = feat.shape
    dt = feat.dtype
    out = []
    offset = 0.5 if aligned else 0.0
    for roi in rois.to(dt):
        b = int(roi[0])
        x1 = roi[1] * spatial_scale - offset
        y1 = roi[2] * spatial_scale - offset
        x2 = roi[3] * spatial_scale - offset
        y2 = roi[4] * spatial_scale - offset
        rw, rh = x2 - x1, y2 - y1
        if not aligned:
            rw, rh = rw.clamp(min=1.0), rh.clamp(min=1.0)
        bin_h, bin_w = rh / ph, rw / pw
        gh = sampling_ratio if sampling_ratio > 0 else int(math.ceil(float(rh) / ph))
        gw = sampling_ratio if sampling_ratio > 0 else int(math.ceil(float(rw) / pw))
        count = max(gh * gw, 1)
        if gh <= 0 or gw <= 0:
            out.append(feat.new_zeros(c, ph, pw))
            continue
        iy = torch.arange(gh, dtype=dt)
        ix = torch.arange(gw, dtype=dt)
        p_h = torch.arange(ph, dtype=dt)
        p_w = torch.arange(pw, dtype=dt)
        ys = y1 + p_h[:, None] * bin_h + (iy[None, :] + 0.5) * bin_h / gh            # [PH, gh]
        xs = x1 + p_w[:, None] * bin_w + (ix[None, :] + 0.5) * bin_w / gw            # [PW, gw]
        y = ys[:, None, :, None].expand(ph, pw, gh, gw)
        x = xs[None, :, None, :].expand(ph, pw, gh, gw)
        valid, y_low, x_low, y_high, x_high, ly, lx = _interp_terms(y, x, height, width)
        hy, hx = 1.0 - ly, 1.0 - lx
        fmap = feat[b]                                                               # [C, H, W]
        v1 = fmap[:, y_low, x_low]
        v2 = fmap[:, y_low, x_high]
        v3 = fmap[:, y_high, x_low]
        v4 = fmap[:, y_high, x_high]
        val = hy * hx * v1 + hy * lx * v2 + ly * hx * v3 + ly * lx * v4              # [C, PH, PW, gh, gw]
        val = val * valid.to(dt)
        out.append(val.sum((-1, -2)) / count)
    if not out:
        return feat.new_zeros(0, c, ph, pw)
    return torch.stack(out)


def roi_align_tokens(tokens, rois, height, width, output_size, spatial_scale=1.0, sampling_ratio=0, aligned=True):
    """Token-major form of the same: tokens [N, H*W, C] -> [K, PH*PW, C] (the layout of the product kernel)."""
    n, hw, c = tokens.shape
    feat = tokens.transpose(1, 2).reshape(n, c, height, width)
    out = roi_align(feat, rois, output_size, spatial_scale, sampling_ratio, aligned)
    return out.flatten(2).transpose(1, 2)
