"""Role-level cycle breakdown of the K2 kernel (ofb_corr_pyramid_bf16_profile): where the TMA warp,
the MMA warp and the epilogue warps wait.  python tools/k2_profile.py [shape ...]  shape = c3|c4|c5"""
import ctypes
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "torch-optical-flow_b200"))
import torch  # noqa: E402

import ofb200  # noqa: E402
from model.corr import CorrBlock, prepare_operands  # noqa: E402

SHAPES = {"c3": (16, 256, 55, 128), "c4": (16, 256, 47, 156), "c5": (4, 256, 136, 240), "c5b8": (8, 256, 136, 240)}
NAMES = ["tma_wait_b_empty", "tma_wait_a_empty", "mma_wait_a_full", "mma_wait_t_empty", "mma_wait_b_full",
         "epi_wait_t_full", "epi_store_section", "tiles", "kernel_cycles", "epi_tmem_ld", "epi_bulk_wait"]


def main():
    lib = ofb200.load()
    for key in sys.argv[1:] or ["c3", "c5"]:
        b, c, h, w = SHAPES[key] if key in SHAPES else tuple(int(v) for v in key.split("x"))
        gen = torch.Generator(device="cuda").manual_seed(1)
        f1 = torch.randn((b, c, h, w), device="cuda", generator=gen)
        f2 = torch.randn((b, c, h, w), device="cuda", generator=gen)
        blk = CorrBlock(f1, f2)
        st = ofb200.stream_ptr()
        a_km, b_km, q_km = prepare_operands(f1, f2, 4)
        for cg in tuple(int(x) for x in os.environ.get("K2_PROF_CG", "1,2").split(",")):
            prof = torch.zeros((2, 148, 16), dtype=torch.int64, device="cuda")
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for it in range(2):
                if it == 1:
                    e0.record()
                rc = lib.ofb_corr_pyramid_bf16_profile(ofb200.ptr(a_km), ofb200.ptr(b_km), ofb200.ptr(q_km),
                                                       ctypes.byref(blk._pyr), b, c, h, w, 1.0, cg, ofb200.ptr(prof), st)
                assert rc == 0, rc
            e1.record()
            torch.cuda.synchronize()
            p1 = prof[1].cpu().double()
            p = prof[0].cpu().double()                 # slot 0 = the full-resolution run (levels 0, 1)
            p = p[p[:, 8] > 0]
            rec = {"shape": key, "cta_group": cg, "dbg": os.environ.get("OFB_K2_DBG", "0"), "ctas": int(p.shape[0]),
                   "epi": os.environ.get("OFB_K2_EPI", "default")}
            for i, nm in enumerate(NAMES):
                rec[nm] = round(float(p[:, i].mean()), 1)
            rec["ms_both_runs"] = round(e0.elapsed_time(e1), 4)
            rec["run2_kernel_cycles_max"] = float(p1[:, 8].max())
            rec["run1_kernel_cycles_max"] = float(p[:, 8].max())
            rec["mhz_if_serial"] = round((rec["run1_kernel_cycles_max"] + rec["run2_kernel_cycles_max"]) / rec["ms_both_runs"] / 1e3, 1)
            t = max(rec["tiles"], 1.0)
            rec["cycles_per_tile"] = round(rec["kernel_cycles"] / t, 1)
            for nm in NAMES[:7] + NAMES[9:]:
                rec[nm + "_per_tile"] = round(rec[nm] / t, 1)
            for nm in NAMES:
                rec.pop(nm, None)
            print(json.dumps(rec), flush=True)
        del blk


if __name__ == "__main__":
    main()
