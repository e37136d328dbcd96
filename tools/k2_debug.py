"""Debug driver for the tcgen05 correlation-pyramid kernel: one configuration per process so a
trapped launch cannot poison other runs.  usage: k2_debug.py B C h w cta_group [levels]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "torch-optical-flow_b200"))
import torch  # noqa: E402

from model.corr import CorrBlock  # noqa: E402

B, C, h, w, cg = [int(x) for x in sys.argv[1:6]]
levels = int(sys.argv[6]) if len(sys.argv) > 6 else 4
gen = torch.Generator(device="cuda").manual_seed(1)
f1 = torch.randn((B, C, h, w), device="cuda", generator=gen)
f2 = torch.randn((B, C, h, w), device="cuda", generator=gen)
ref = CorrBlock(f1, f2, num_levels=levels, pyramid_dtype=torch.float32, builder="simt")
torch.cuda.synchronize()
print("simt ok", flush=True)
blk = CorrBlock(f1, f2, num_levels=levels, cta_group=cg)
torch.cuda.synchronize()
print("tcgen05 launched ok", flush=True)
for l in range(levels):
    a = blk.corr_pyramid[l].float()
    r = ref.corr_pyramid[l]
    rel = float((a - r).norm() / r.norm())
    mx = float((a - r).abs().max())
    nan = int(torch.isnan(a).sum())
    print(f"level {l}: shape {tuple(a.shape)} rel {rel:.3e} maxabs {mx:.3e} nan {nan}", flush=True)
    if rel > 1e-2 and l == 0:
        n = h * w
        d = (a - r).abs().view(B, n, h, w)
        bad = d > 0.1
        print("  bad frac", float(bad.float().mean()))
        print("  bad by batch", bad.float().mean(dim=(1, 2, 3)).tolist())
        print("  bad by query block(128)", [round(float(x), 3) for x in bad[0].float().mean(dim=(1, 2)).view(-1)[: 128 * 4].view(-1, 128).mean(1)] if n >= 512 else "")
        print("  bad by target row", [round(float(x), 3) for x in bad[0].float().mean(dim=(0, 2))][:32])
        print("  bad by target col", [round(float(x), 3) for x in bad[0].float().mean(dim=(0, 1))][:64])
        print("  sample got", a.view(B, n, h, w)[0, 0, 0, :8].tolist())
        print("  sample ref", r.view(B, n, h, w)[0, 0, 0, :8].tolist())
# timing
if os.environ.get("K2_TIME"):
    for _ in range(3):
        CorrBlock(f1, f2, num_levels=levels, cta_group=cg)
    torch.cuda.synchronize()
    import ctypes, math
    import ofb200
    lib = ofb200.load()
    n = h * w
    a_km = torch.empty((B, n, C), dtype=torch.bfloat16, device="cuda")
    b_km = torch.empty((B, n, C), dtype=torch.bfloat16, device="cuda")
    st = ofb200.stream_ptr()
    lib.ofb_corr_prep_bf16(ofb200.ptr(f1), ofb200.ptr(a_km), B, C, n, st)
    lib.ofb_corr_prep_bf16(ofb200.ptr(f2), ofb200.ptr(b_km), B, C, n, st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        lib.ofb_corr_pyramid_bf16(ofb200.ptr(a_km), ofb200.ptr(b_km), ctypes.byref(blk._pyr), B, C, h, w, 1.0 / math.sqrt(C), cg, st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flops = 2.0 * B * n * n * C
    elems = sum(int(blk._pyr.lvl_h[l]) * int(blk._pyr.lvl_w[l]) for l in range(levels)) * B * n
    print(f"K2 time {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s  write {elems * 2 / ms / 1e6:.1f} GB/s", flush=True)
