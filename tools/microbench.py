"""Per-kernel micro-benchmarks at the BASELINE.json shapes (CUDA events on the launching stream,
3 warm-ups, inputs rotated / larger than L2).  Prints one JSON line per kernel with the algorithmic
bytes / flops of SURVEY.md section 8d and the achieved fraction of the measured peaks."""
import ctypes
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "torch-optical-flow_b200"))
import torch  # noqa: E402

import ofb200  # noqa: E402
from model.corr import CorrBlock  # noqa: E402
from model.raft import upsample_flow  # noqa: E402
from model.utils import coords_grid, upflow8  # noqa: E402
from optical_flow import normalize, warp  # noqa: E402
from optical_flow.metrics import AverageEndPointError  # noqa: E402

PEAKS = {"hbm_gbs": 6565.5, "bf16_tflops": 1665.4, "bf16_tflops_sustained": 1391.6}
try:
    PEAKS.update(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))))
except Exception:
    pass


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def report(name, ms, nbytes=None, flops=None, **extra):
    rec = {"kernel": name, "ms": round(ms, 4)}
    if nbytes is not None:
        gbs = nbytes / ms / 1e6
        rec.update(gbs=round(gbs, 1), hbm_frac=round(gbs / PEAKS["hbm_gbs"], 3))
    if flops is not None:
        tf = flops / ms / 1e9
        rec.update(tflops=round(tf, 1), tc_frac_sustained=round(tf / PEAKS["bf16_tflops_sustained"], 3),
                   tc_frac_burst=round(tf / PEAKS["bf16_tflops"], 3))
    rec.update(extra)
    print(json.dumps(rec), flush=True)


def smooth_flow(b, h, w, sigma, gen):
    low = torch.randn((b, 2, h // 16 + 2, w // 16 + 2), device="cuda", generator=gen) * sigma
    return torch.nn.functional.interpolate(low, size=(h, w), mode="bilinear", align_corners=True).contiguous()


def bench_warp(which):
    b, c, h, w = 32, 3, 436, 1024
    gen = torch.Generator(device="cuda").manual_seed(1234)
    frame = torch.rand((b, c, h, w), device="cuda", generator=gen)
    flows = {
        "white5px": normalize(5 * torch.randn((b, 2, h, w), device="cuda", generator=gen)),
        "smooth5px": normalize(smooth_flow(b, h, w, 5.0, gen)),
    }
    px = b * h * w
    lib = ofb200.load()
    out = torch.empty_like(frame)
    mask = torch.empty((b, h, w), dtype=torch.uint8, device="cuda")
    for fname, flow in flows.items():
        for variant in (1, 2):
            for with_mask in (0, 1):
                def run():
                    rc = lib.ofb_warp_f32(ofb200.ptr(frame), ofb200.ptr(flow), ofb200.ptr(out),
                                          ofb200.ptr(mask) if with_mask else None, b, c, h, w, 0, 1, 0, 0, variant,
                                          ofb200.stream_ptr())
                    assert rc == 0
                ms = timeit(run)
                report(f"K1 warp C2 v{variant} {fname} mask={with_mask}", ms, nbytes=px * (32 + with_mask),
                       img_per_s=round(b / ms * 1e3))


def bench_corr(which):
    shapes = {"C3 sintel B16": (16, 256, 55, 128), "C4 kitti B16": (16, 256, 47, 156), "C5 1080p B4": (4, 256, 136, 240)}
    lib = ofb200.load()
    for name, (b, c, h, w) in shapes.items():
        gen = torch.Generator(device="cuda").manual_seed(1)
        f1 = torch.randn((b, c, h, w), device="cuda", generator=gen)
        f2 = torch.randn((b, c, h, w), device="cuda", generator=gen)
        n = h * w
        a_km = torch.empty((b, n, c), dtype=torch.bfloat16, device="cuda")
        b_km = torch.empty((b, n, c), dtype=torch.bfloat16, device="cuda")
        st = ofb200.stream_ptr()
        ms = timeit(lambda: (lib.ofb_corr_prep_bf16(ofb200.ptr(f1), ofb200.ptr(a_km), b, c, n, st),
                             lib.ofb_corr_prep_bf16(ofb200.ptr(f2), ofb200.ptr(b_km), b, c, n, st)))
        report(f"K2 prep {name}", ms, nbytes=2 * b * c * n * 6)
        blk = CorrBlock(f1, f2)
        elems = sum(int(blk._pyr.lvl_h[l]) * int(blk._pyr.lvl_w[l]) for l in range(4)) * b * n
        for cg in (1, 2):
            def run():
                rc = lib.ofb_corr_pyramid_bf16(ofb200.ptr(a_km), ofb200.ptr(b_km), ctypes.byref(blk._pyr), b, c, h, w,
                                               1.0 / math.sqrt(c), cg, st)
                assert rc == 0
            ms = timeit(run, reps=10)
            report(f"K2 pyramid {name} cta_group={cg}", ms, nbytes=elems * 2 + 2 * b * n * c * 2, flops=2.0 * b * n * n * c,
                   pairs_per_s=round(b / ms * 1e3, 1))
        # K3 on the same pyramid
        coords = coords_grid(b, h, w).cuda() + 4 * torch.randn((b, 2, h, w), device="cuda", generator=gen)
        out = torch.empty((b, 324, h, w), device="cuda")
        def look():
            rc = lib.ofb_corr_lookup(ctypes.byref(blk._pyr), ofb200.ptr(coords), ofb200.ptr(out), None, None, b, h, w, 4, st)
            assert rc == 0
        ms = timeit(look, reps=12)
        report(f"K3 lookup bf16 {name} (per iteration)", ms, nbytes=b * n * 2104, iters_12_ms=round(12 * ms, 3))
        del blk
        if h * w <= 7400 and b >= 16:
            blk32 = CorrBlock(f1[:8], f2[:8], pyramid_dtype=torch.float32)
            c8, o8 = coords[:8].contiguous(), torch.empty((8, 324, h, w), device="cuda")
            def look32():
                rc = lib.ofb_corr_lookup(ctypes.byref(blk32._pyr), ofb200.ptr(c8), ofb200.ptr(o8), None, None, 8, h, w, 4, st)
                assert rc == 0
            ms = timeit(look32, reps=12)
            report(f"K3 lookup fp32 {name} B8 (per iteration)", ms, nbytes=8 * n * 2904)
            del blk32


def bench_upsample(which):
    n, h, w = 16, 47, 156
    gen = torch.Generator(device="cuda").manual_seed(3)
    masks = [torch.randn((n, 576, h, w), device="cuda", generator=gen) for _ in range(2)]
    flow = torch.randn((n, 2, h, w), device="cuda", generator=gen)
    lib = ofb200.load()
    out = torch.empty((n, 2, 8 * h, 8 * w), device="cuda")
    k = [0]
    def run():
        k[0] ^= 1
        rc = lib.ofb_convex_upsample_f32(ofb200.ptr(flow), ofb200.ptr(masks[k[0]]), ofb200.ptr(out), n, h, w, ofb200.stream_ptr())
        assert rc == 0
    ms = timeit(run)
    report("K4b convex upsample C4", ms, nbytes=n * h * w * (578 * 4 + 128 * 4), pairs_per_s=round(n / ms * 1e3))
    flows = [torch.randn((n, 2, h, w), device="cuda", generator=gen) for _ in range(4)]
    outs = [torch.empty((n, 2, 8 * h, 8 * w), device="cuda") for _ in range(4)]
    def up():
        k[0] = (k[0] + 1) % 4
        rc = lib.ofb_resize_bilinear_f32(ofb200.ptr(flows[k[0]]), ofb200.ptr(outs[k[0]]), n, 2, h, w, 8 * h, 8 * w, 1, 8.0, 8.0, ofb200.stream_ptr())
        assert rc == 0
    ms = timeit(up)
    report("K4a upflow8 C4", ms, nbytes=n * 2 * h * w * 4 * 65)
    preds = [torch.randn((n, 2, 8 * h, 8 * w), device="cuda", generator=gen) for _ in range(3)]
    tgts = [torch.randn((n, 2, 8 * h, 8 * w), device="cuda", generator=gen) for _ in range(3)]
    valid = (torch.rand((n, 8 * h, 8 * w), device="cuda", generator=gen) > 0.1).float()
    acc = torch.zeros(2, dtype=torch.float64, device="cuda")
    def epe():
        k[0] = (k[0] + 1) % 3
        rc = lib.ofb_epe_reduce_f32(ofb200.ptr(preds[k[0]]), ofb200.ptr(tgts[k[0]]), ofb200.ptr(valid), ofb200.ptr(acc), n, 8 * h, 8 * w, ofb200.stream_ptr())
        assert rc == 0
    ms = timeit(epe)
    report("K4c EPE reduce C4 (valid map)", ms, nbytes=n * 8 * h * 8 * w * 20)


if __name__ == "__main__":
    which = sys.argv[1:] or ["warp", "corr", "upsample"]
    print(json.dumps({"peaks": {k: PEAKS[k] for k in ("hbm_gbs", "bf16_tflops", "bf16_tflops_sustained")}}))
    if "warp" in which:
        bench_warp(which)
    if "upsample" in which:
        bench_upsample(which)
    if "corr" in which:
        bench_corr(which)
