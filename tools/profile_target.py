"""Small launch sequence for `ncu --set full`: each hot kernel twice at its BASELINE shape."""
import ctypes
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "torch-optical-flow_b200"))
import torch  # noqa: E402

import ofb200  # noqa: E402
from model.corr import CorrBlock  # noqa: E402
from model.raft import upsample_flow  # noqa: E402
from model.utils import coords_grid  # noqa: E402
from optical_flow import normalize, warp  # noqa: E402

which = sys.argv[1:] or ["warp", "corr", "lookup", "upsample"]
gen = torch.Generator(device="cuda").manual_seed(0)
reps = 2
if "warp" in which:
    b, c, h, w = 32, 3, 436, 1024
    frame = torch.rand((b, c, h, w), device="cuda", generator=gen)
    flow = normalize(5 * torch.randn((b, 2, h, w), device="cuda", generator=gen))
    for v in (1, 2):
        for _ in range(reps):
            warp(frame, flow, return_mask=True, variant=v)
if "corr" in which or "lookup" in which:
    b, c, h, w = (4, 256, 136, 240) if "c5" in which else (16, 256, 55, 128)
    f1 = torch.randn((b, c, h, w), device="cuda", generator=gen)
    f2 = torch.randn((b, c, h, w), device="cuda", generator=gen)
    for cg in (1, 2):
        for _ in range(reps):
            blk = CorrBlock(f1, f2, cta_group=cg)
    blk = CorrBlock(f1, f2)
    coords = coords_grid(b, h, w).cuda() + 4 * torch.randn((b, 2, h, w), device="cuda", generator=gen)
    if "lookup" in which:
        for _ in range(reps):
            blk(coords)
if "upsample" in which:
    n, h, w = 16, 47, 156
    mask = torch.randn((n, 576, h, w), device="cuda", generator=gen)
    flow = torch.randn((n, 2, h, w), device="cuda", generator=gen)
    for _ in range(reps):
        upsample_flow(flow, mask)
torch.cuda.synchronize()
print("done")
