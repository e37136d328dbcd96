"""One C2 warp launch per kernel variant (3 = rows, 4 = TMA window) on a smooth flow -- the ncu target for K1."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "torch-optical-flow_b200"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch

import ofb200
from microbench import smooth_flow
from optical_flow import normalize

b, c, h, w = 32, 3, 436, 1024
gen = torch.Generator(device="cuda").manual_seed(1234)
frame = torch.rand((b, c, h, w), device="cuda", generator=gen)
flow = normalize(smooth_flow(b, h, w, 5.0, gen, int(os.environ.get("SPACING", "64"))))
out = torch.empty_like(frame)
mask = torch.empty((b, h, w), dtype=torch.uint8, device="cuda")
lib = ofb200.load()
for variant in (3, 4, 3, 4):
    ofb200.check(lib.ofb_warp_f32(ofb200.ptr(frame), ofb200.ptr(flow), ofb200.ptr(out), ofb200.ptr(mask), b, c, h, w,
                                  0, 1, 0, 0, variant, 1.0, 1.0, ofb200.stream_ptr()), "warp")
torch.cuda.synchronize()
print("ok")
