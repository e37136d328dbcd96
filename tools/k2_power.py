"""Is K2 power-bound?  Runs the pyramid builder back to back for a few seconds per configuration while sampling
nvidia-smi (power.draw, clocks.sm, throttle reasons) every 50 ms, and prints time per launch, median power and clock.

    python tools/k2_power.py [c5b8] [seconds]          K2_AB_MODES / K2_AB_CG as in tools/k2_ab.py
"""
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

import microbench  # noqa: E402
import ofb200  # noqa: E402
from model.corr import CorrBlock, prepare_operands  # noqa: E402

SHAPES = {"c5b8": (8, 256, 136, 240), "c3": (16, 256, 55, 128)}


class Sampler:
    def __init__(self):
        self.rows = []
        self.p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=power.draw,clocks.sm,clocks.mem,clocks_event_reasons.sw_power_cap",
                                   "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
        threading.Thread(target=self._pump, daemon=True).start()

    def _pump(self):
        for line in self.p.stdout:
            f = [x.strip() for x in line.split(",")]
            try:
                self.rows.append((time.time(), float(f[0]), float(f[1]), float(f[2]), f[3]))
            except (ValueError, IndexError):
                pass

    def window(self, t0, t1):
        r = [x for x in self.rows if t0 + 0.3 <= x[0] <= t1]
        if not r:
            return {}
        return {"power_w_median": statistics.median(x[1] for x in r), "power_w_max": max(x[1] for x in r),
                "sm_mhz_median": statistics.median(x[2] for x in r), "mem_mhz": r[-1][3],
                "power_cap_active_frac": sum(x[4].lower().startswith("active") for x in r) / len(r), "samples": len(r)}


def main():
    key = next((a for a in sys.argv[1:] if a in SHAPES), "c5b8")
    secs = float(next((a for a in sys.argv[1:] if a.replace(".", "").isdigit()), "3"))
    b, c, h, w = SHAPES[key]
    lib = ofb200.load()
    gen = torch.Generator(device="cuda").manual_seed(1)
    f1 = torch.randn((b, c, h, w), device="cuda", generator=gen)
    f2 = torch.randn((b, c, h, w), device="cuda", generator=gen)
    blk = CorrBlock(f1, f2)
    a_km, b_km, q_km = prepare_operands(f1, f2, 4)
    st = ofb200.stream_ptr()
    sm = Sampler()
    time.sleep(1.0)
    idle = sm.window(time.time() - 1.0, time.time())
    print(json.dumps({"idle": idle}), flush=True)
    for mode in os.environ.get("K2_AB_MODES", "direct").split(","):
        os.environ["OFB_K2_EPI"] = mode
        for cg in (int(x) for x in os.environ.get("K2_AB_CG", "1,2").split(",")):
            def run():
                rc = lib.ofb_corr_pyramid_bf16(ofb200.ptr(a_km), ofb200.ptr(b_km), ofb200.ptr(q_km), ctypes.byref(blk._pyr),
                                               b, c, h, w, 1.0, cg, st)
                assert rc == 0
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            t0 = time.time()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 0
            e0.record()
            while time.time() - t0 < secs:
                for _ in range(20):
                    run()
                n += 20
                torch.cuda.synchronize()
            e1.record()
            torch.cuda.synchronize()
            t1 = time.time()
            ms = e0.elapsed_time(e1) / n
            rec = {"shape": key, "epi": mode, "cta_group": cg, "ms_per_launch": round(ms, 4), "launches": n}
            rec.update(sm.window(t0, t1))
            if "power_w_median" in rec:
                rec["joule_per_launch"] = round(rec["power_w_median"] * ms * 1e-3, 3)
            print(json.dumps(rec), flush=True)
            time.sleep(1.0)
    sm.p.terminate()


if __name__ == "__main__":
    main()
