"""Regenerate the SASS evidence from the SHIPPED library: `cuobjdump -sass libofb200.so`, per-kernel counts of the
mnemonics that prove a Blackwell-native path (B200_PROFILING.md: UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st,
UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, SYNCS = mbarrier,
LDGSTS = cp.async) plus registers / shared memory from `cuobjdump -res-usage`.

    python tools/sass_evidence.py [out.txt]        (default profiles/r02_sass_evidence.txt; needs no GPU)
The product kernels (the ones bench.py's pass launches by default) are listed first and marked `*`."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "torch-optical-flow_b200", "ofb200", "libofb200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "UTCATOMSWS", "SYNCS", "LDGSTS",
        "HMMA", "LDG", "STG", "LDS", "STS", "RED", "REDG", "ATOM", "ATOMG", "REDUX", "FFMA", "MUFU", "SHFL", "BAR"]
PRODUCT = ["prep_kernelILi1EfE", "prep_kernelILi4EfE", "prep_kernelILi1E13__nv_bfloat16E", "prep_kernelILi4E13__nv_bfloat16E",
           "corr_pyramid_kernelILi1ELb0ELi2ELi0ELb1E", "lookup_tile_kernelILi4ELi3E", "convex_upsample_kernel",
           "warp_rows_kernelILi1ELb0ELi2E", "epe_reduce_kernelILi0E", "gemm_nt_kernel", "lookup_bwd", "warp_bwd"]


def demangle(names):
    try:
        out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
        return dict(zip(names, out))
    except OSError:
        return {n: n for n in names}


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_evidence.txt")
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+).*?SHARED:(\d+)", line)
        if m and cur:
            usage[cur] = (int(m.group(1)), int(m.group(2)))
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur]["instrs"] += 1
            base = op.split(".")[0]
            if base in KEYS:
                kernels[cur][base] += 1
            if op.startswith("STG.E.ENL2.256") or ".256" in op and base == "STG":
                kernels[cur]["STG.256"] += 1
            if base in ("LDG", "STG") and ".128" in op:
                kernels[cur][base + ".128"] += 1
    dm = demangle(list(kernels))
    is_prod = lambda n: any(p in n for p in PRODUCT)
    order = sorted(kernels, key=lambda n: (not is_prod(n), n))
    with open(out_path, "w") as fh:
        fh.write("# SASS evidence: cuobjdump -sass torch-optical-flow_b200/ofb200/libofb200.so (tools/sass_evidence.py)\n")
        fh.write("# mnemonic counts per kernel; UTCHMMA = tcgen05.mma kind::f16, LDTM = tcgen05.ld, UTMALDG = TMA tensor load,\n")
        fh.write("# UBLKCP = cp.async.bulk (TMA bulk store of the EPI_BULK experiment), UTCBAR = tcgen05.commit, SYNCS = mbarrier,\n")
        fh.write("# LDGSTS = cp.async, STG.256 = 32-byte sector stores (st.global.v8.b32).  `*` = launched by the default bench pass.\n\n")
        for n in order:
            c = kernels[n]
            reg, sh = usage.get(n, (None, None))
            parts = [f"instrs={c['instrs']}"] + [f"{k}={c[k]}" for k in KEYS + ["STG.256", "LDG.128", "STG.128"] if c[k]]
            fh.write(f"{'*' if is_prod(n) else ' '} {dm.get(n, n)[:150]}\n      regs={reg} smem_static={sh}  " + " ".join(parts) + "\n")
    print(f"{len(kernels)} kernels -> {out_path}")


if __name__ == "__main__":
    main()
