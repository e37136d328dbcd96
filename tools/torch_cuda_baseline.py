"""Comparator asked for in SURVEY.md section 8d: the reference's op sequence for the hot path run on the SAME B200 through
stock PyTorch CUDA ops (matmul, avg_pool2d, grid_sample, softmax / unfold, norm) -- what a user of the reference gets
today by calling `.cuda()`.  Written directly against torch (reference call sites: methods/raft/model/corr.py:38-87,
utils.py:64-80, raft.py:73-85, optical_flow/operator/operator.py:8-56, optical_flow/metrics/epe.py:25-35); it does not
touch this repo's kernels or its oracle.  Prints one JSON line; not part of bench.py.

    python tools/torch_cuda_baseline.py [pairs_per_step=1] [steps=5]
"""
import json
import math
import sys
import time

import torch
import torch.nn.functional as F

H, W, C, ITERS, RADIUS, LEVELS = 1088, 1920, 256, 12, 4, 4
h8, w8 = H // 8, W // 8


def corr_pyramid(f1, f2):
    b, c, h, w = f1.shape
    corr = torch.matmul(f1.view(b, c, h * w).transpose(1, 2), f2.view(b, c, h * w)).view(b * h * w, 1, h, w) / math.sqrt(c)
    pyr = [corr]
    for _ in range(LEVELS - 1):
        corr = F.avg_pool2d(corr, 2, stride=2)
        pyr.append(corr)
    return pyr


def sampler(img, coords):
    h, w = img.shape[-2:]
    xg = 2 * coords[..., 0:1] / (w - 1) - 1
    yg = 2 * coords[..., 1:2] / (h - 1) - 1
    return F.grid_sample(img, torch.cat([xg, yg], dim=-1), align_corners=True)


def lookup(pyr, coords, delta):
    b, _, h, w = coords.shape
    xy = coords.permute(0, 2, 3, 1).reshape(b * h * w, 1, 1, 2)
    out = [sampler(corr, xy / 2 ** lvl + delta).view(b, h, w, -1) for lvl, corr in enumerate(pyr)]
    return torch.cat(out, dim=-1).permute(0, 3, 1, 2).contiguous().float()


def upsample_flow(flow, mask):
    n, _, h, w = flow.shape
    mask = torch.softmax(mask.view(n, 1, 9, 8, 8, h, w), dim=2)
    up = F.unfold(8 * flow, [3, 3], padding=1).view(n, 2, 9, 1, 1, h, w)
    return torch.sum(mask * up, dim=2).permute(0, 1, 4, 2, 5, 3).reshape(n, 2, 8 * h, 8 * w)


def warp(frame, flow_px):
    b, _, h, w = flow_px.shape
    fac = torch.tensor([2.0 / max(w - 1, 1), 2.0 / max(h - 1, 1)], device=flow_px.device).view(1, 2, 1, 1)
    flow = (flow_px * fac).permute(0, 2, 3, 1)
    gy, gx = torch.meshgrid(torch.linspace(-1, 1, h, device=flow.device), torch.linspace(-1, 1, w, device=flow.device), indexing="ij")
    grid = torch.stack((gx, gy), dim=-1)[None].expand(b, -1, -1, -1) + flow
    return F.grid_sample(frame, grid, mode="bilinear", padding_mode="border", align_corners=False)


def one_pass(bt, delta):
    pyr = corr_pyramid(bt["fmap1"], bt["fmap2"])
    for it in range(ITERS):
        lookup(pyr, bt["coords"][it], delta)
    up = upsample_flow(bt["flow_lo"], bt["up_mask"])
    warp(bt["frame"], up)
    epe = torch.norm(up - bt["target"], p=2, dim=1).view(-1)[bt["valid"].view(-1) >= 0.5]
    return epe.sum(), epe.numel()


def main():
    pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(1234)
    rn = lambda *s: torch.randn(s, device=dev, generator=g)  # noqa: E731
    ys, xs = torch.meshgrid(torch.arange(h8, device=dev), torch.arange(w8, device=dev), indexing="ij")
    base = torch.stack((xs, ys), 0).float()[None]
    bt = {"fmap1": rn(pairs, C, h8, w8), "fmap2": rn(pairs, C, h8, w8),
          "coords": (base[None] + 4 * rn(ITERS, pairs, 2, h8, w8)).contiguous(), "flow_lo": 1.5 * rn(pairs, 2, h8, w8),
          "up_mask": rn(pairs, 576, h8, w8), "frame": torch.rand((pairs, 3, H, W), device=dev, generator=g),
          "target": 12 * rn(pairs, 2, H, W), "valid": (torch.rand((pairs, H, W), device=dev, generator=g) > 0.1).float()}
    d = torch.linspace(-RADIUS, RADIUS, 2 * RADIUS + 1, device=dev)
    delta = torch.stack(torch.meshgrid(d, d, indexing="ij"), dim=-1).view(1, 2 * RADIUS + 1, 2 * RADIUS + 1, 2)
    with torch.no_grad():
        for _ in range(2):
            one_pass(bt, delta)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            one_pass(bt, delta)
        e1.record()
        torch.cuda.synchronize()
    ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3) / steps
    print(json.dumps({"impl": "reference op sequence on stock PyTorch CUDA ops (fp32 volume)", "metric": "image-pairs/sec (corr+lookup, warp)",
                      "value": round(pairs / (ms * 1e-3), 2), "unit": "pairs/s", "ms_per_step": round(ms, 3), "pairs_per_step": pairs,
                      "steps": steps, "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 1e9, 2), "torch": torch.__version__,
                      "gpu": torch.cuda.get_device_name(0)}))


if __name__ == "__main__":
    main()
