"""Restatement of the reference's hot path with the SAME PyTorch (ATen) CPU ops the reference calls --
TEST / BASELINE INFRASTRUCTURE ONLY, like oracle.py (never imported by the product path).

The reference is pure Python over torch ops, so its CPU performance *is* the performance of this op
sequence; `bench.py --impl reference` times it next to the C/OpenMP oracle and reports the faster of the
two as the CPU arm.  Each function names the reference lines it follows (paths relative to
/root/reference); the code is written from the call stacks in SURVEY.md section 3, not copied.
"""
import math

import torch
import torch.nn.functional as F


def warp_grid(flow_bhw2):
    """optical_flow/operator/operator.py:36-56 -- base grid linspace(-1, 1) in x and y, plus the flow."""
    b, h, w, _ = flow_bhw2.shape
    xs = torch.linspace(-1, 1, w, device=flow_bhw2.device)
    ys = torch.linspace(-1, 1, h, device=flow_bhw2.device)
    gy, gx = torch.meshgrid(ys, xs, indexing="ij")
    base = torch.stack((gx, gy), dim=-1)[None].expand(b, -1, -1, -1)
    return base + flow_bhw2


def warp(frame, flow, mode="bilinear", padding_mode="border", align_corners=False):
    """operator.py:8-33 -- grid_sample on the warp grid."""
    return F.grid_sample(frame, warp_grid(flow.permute(0, 2, 3, 1)), mode=mode, padding_mode=padding_mode,
                         align_corners=align_corners)


def normalize(flow):
    """operator.py:117-130 (through scale, :59-82): x * 2/max(W-1,1), y * 2/max(H-1,1)."""
    h, w = flow.shape[-2:]
    fac = torch.tensor([2.0 / max(w - 1, 1), 2.0 / max(h - 1, 1)], dtype=flow.dtype, device=flow.device)
    return flow * fac.view(1, 2, 1, 1)


def resize(flow, size):
    """operator.py:85-114 -- F.interpolate(bilinear, align_corners=False) then scale by (W'/W, H'/H)."""
    h, w = flow.shape[-2:]
    out = F.interpolate(flow, size=size, mode="bilinear")
    fac = torch.tensor([size[1] / w, size[0] / h], dtype=flow.dtype, device=flow.device)
    return out * fac.view(1, 2, 1, 1)


def upflow8(flow):
    """methods/raft/model/utils.py:89-91 -- 8 * F.interpolate(bilinear, align_corners=True)."""
    h, w = flow.shape[-2:]
    return 8 * F.interpolate(flow, size=(8 * h, 8 * w), mode="bilinear", align_corners=True)


def bilinear_sampler(img, coords):
    """methods/raft/model/utils.py:64-80 -- pixel coordinates -> [-1, 1] -> grid_sample(align_corners=True)."""
    h, w = img.shape[-2:]
    xg = 2 * coords[..., 0:1] / (w - 1) - 1
    yg = 2 * coords[..., 1:2] / (h - 1) - 1
    return F.grid_sample(img, torch.cat([xg, yg], dim=-1), align_corners=True)


def corr_pyramid(fmap1, fmap2, num_levels=4):
    """methods/raft/model/corr.py:38-54,79-87 -- matmul / sqrt(C), then repeated 2x2 average pooling."""
    b, c, h, w = fmap1.shape
    corr = torch.matmul(fmap1.view(b, c, h * w).transpose(1, 2), fmap2.view(b, c, h * w))
    corr = corr.view(b * h * w, 1, h, w) / math.sqrt(c)
    pyr = [corr]
    for _ in range(num_levels - 1):
        corr = F.avg_pool2d(corr, 2, stride=2)
        pyr.append(corr)
    return pyr


def corr_lookup(pyr, coords, radius=4):
    """corr.py:56-77 -- per level a (2r+1)^2 window around coords / 2^l; the window offsets come from
    meshgrid(dy, dx) stacked as (dy, dx) and are added to (x, y): the slow window index moves x."""
    b, _, h, w = coords.shape
    xy = coords.permute(0, 2, 3, 1).reshape(b * h * w, 1, 1, 2)
    d = torch.linspace(-radius, radius, 2 * radius + 1)
    delta = torch.stack(torch.meshgrid(d, d, indexing="ij"), dim=-1).view(1, 2 * radius + 1, 2 * radius + 1, 2)
    out = []
    for lvl, corr in enumerate(pyr):
        sampled = bilinear_sampler(corr, xy / 2 ** lvl + delta)
        out.append(sampled.view(b, h, w, -1))
    return torch.cat(out, dim=-1).permute(0, 3, 1, 2).contiguous().float()


def upsample_flow(flow, mask):
    """methods/raft/model/raft.py:73-85 -- convex combination of the 3x3 neighbourhood of 8 * flow."""
    n, _, h, w = flow.shape
    mask = torch.softmax(mask.view(n, 1, 9, 8, 8, h, w), dim=2)
    up = F.unfold(8 * flow, [3, 3], padding=1).view(n, 2, 9, 1, 1, h, w)
    up = torch.sum(mask * up, dim=2)
    return up.permute(0, 1, 4, 2, 5, 3).reshape(n, 2, 8 * h, 8 * w)


def epe_sum_count(pred, target, valid=None):
    """optical_flow/metrics/epe.py:25-35 -- L2 norm over the flow dimension, masked sum and count."""
    epe = torch.norm(pred - target, p=2, dim=1).view(-1)
    if valid is not None:
        epe = epe[valid.view(-1) >= 0.5]
    return float(epe.sum()), int(epe.numel())


def sequence_loss(flow_preds, flow_gt, valid, gamma=0.8, max_flow=400.0):
    """methods/raft/model/raft.py:231-260 -- gamma-weighted masked L1 over the prediction sequence, plus the
    fractions of kept pixels whose final end-point error is below 1 / 3 / 5 px."""
    n = len(flow_preds)
    keep = (valid >= 0.5) & (torch.sum(flow_gt ** 2, dim=1).sqrt() < max_flow)
    loss = 0.0
    for i, pred in enumerate(flow_preds):
        loss = loss + gamma ** (n - i - 1) * (keep[:, None] * (pred - flow_gt).abs()).mean()
    epe = torch.sum((flow_preds[-1] - flow_gt) ** 2, dim=1).sqrt().view(-1)[keep.view(-1)]
    return loss, {f"{t}px": (epe < t).float().mean().item() for t in (1, 3, 5)}
