"""A/B of the K2 epilogue store modes (OFB_K2_EPI=direct|bulk) and cta_group settings at the BASELINE shapes, same
process, alternating, CUDA events; also checks that both modes write bit-identical pyramids.

    python tools/k2_ab.py [c5b8] [c5b4] [c3] [c4]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

import microbench  # noqa: E402

SHAPES = {"c5b8": ("C5 B8 136x240", 8, 256, 136, 240), "c5b4": ("C5 B4 136x240", 4, 256, 136, 240),
          "c3": ("C3 B16 55x128", 16, 256, 55, 128), "c4": ("C4 B16 47x156", 16, 256, 47, 156)}


def main():
    names = [a for a in sys.argv[1:] if a in SHAPES] or ["c5b8", "c3"]
    modes = os.environ.get("K2_AB_MODES", "direct,bulk").split(",")
    cgs = tuple(int(x) for x in os.environ.get("K2_AB_CG", "1").split(","))
    for nm in names:
        ref = None
        for rep in range(2):
            for mode in modes:
                os.environ["OFB_K2_EPI"] = mode
                recs, blk, _ = microbench.bench_corr(SHAPES[nm][0], *SHAPES[nm][1:], cta_groups=cgs, prep=False)
                for r in recs:
                    r["epi"] = mode
                    print(json.dumps(r), flush=True)
                if rep == 0:
                    torch.cuda.synchronize()
                    bufs = [b.clone() for b in blk._buffers]
                    if ref is None:
                        ref = bufs
                    else:
                        same = all(torch.equal(a, b) for a, b in zip(ref, bufs))
                        print(json.dumps({"shape": nm, "mode": mode, "bit_identical_to_first_mode": bool(same)}), flush=True)
                    del bufs
                del blk
                torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
