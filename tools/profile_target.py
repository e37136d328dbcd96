"""Small launch sequence for `ncu --set full`: the kernels of one bench.py pass at the bench shapes
(C5: 8 pairs of 1088x1920, fmaps 256x136x240), each launched twice.
    python tools/profile_target.py [corr] [lookup] [warp] [upsample] [epe]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "torch-optical-flow_b200"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from ofb200.runner import hot_path  # noqa: E402
from optical_flow.metrics.epe import AverageEndPointError  # noqa: E402

pairs = int(os.environ.get("PAIRS", "8"))
batch = bench.make_batch(pairs, torch.device("cuda", 0), 1234, torch)
batch["coords"] = batch["coords"][:2].contiguous()          # 2 lookup launches are enough for the profiler
metric = AverageEndPointError()
for _ in range(2):
    hot_path(batch, metric)
torch.cuda.synchronize()
print("done", float(metric.compute()))
