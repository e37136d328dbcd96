"""Pure-write / pure-read / copy bandwidth of this GPU with torch ops (context for the rooflines)."""
import json
import torch

def t(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

n = 1 << 30
x = torch.empty(n, dtype=torch.bfloat16, device="cuda")
y = torch.empty(n, dtype=torch.bfloat16, device="cuda")
gb = n * 2 / 1e9
print(json.dumps({"fill_write_gbs": round(gb / t(lambda: x.fill_(1.0)) * 1e3, 1),
                  "copy_rw_gbs": round(2 * gb / t(lambda: y.copy_(x)) * 1e3, 1),
                  "sum_read_gbs": round(gb / t(lambda: x.view(torch.int16).sum()) * 1e3, 1)}))

h = torch.empty(1 << 28, dtype=torch.float32, pin_memory=True)
d = torch.empty(1 << 28, dtype=torch.float32, device="cuda")
print(json.dumps({"h2d_pinned_gbs": round(h.numel() * 4 / 1e9 / t(lambda: d.copy_(h, non_blocking=True), reps=5) * 1e3, 1),
                  "d2h_pinned_gbs": round(h.numel() * 4 / 1e9 / t(lambda: h.copy_(d, non_blocking=True), reps=5) * 1e3, 1)}))
