"""Summarise `ncu --page source --csv` output: stall totals and the hottest SASS lines.
usage: ncu -i rep --page source --csv --kernel-name regex:X --launch-skip N --launch-count 1 | python tools/ncu_hot.py [top]"""
import csv
import sys

top_n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rows = list(csv.reader(sys.stdin))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ci = {n: i for i, n in enumerate(hdr)}


def num(r, k):
    try:
        return int(float(r[ci[k]] or 0))
    except ValueError:
        return 0


data = [r for r in rows[h + 1:] if len(r) == len(hdr) and r[0] != "Address"]
tot = sum(num(r, "# Samples") for r in data)
print("kernel:", rows[0][1][:100] if rows[0] else "")
print("total samples", tot, "sass lines", len(data), "warp instr executed", sum(num(r, "Instructions Executed") for r in data))
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
agg = {s: sum(num(r, s) for r in data) for s in stalls}
print("stalls:", [(s, v) for s, v in sorted(agg.items(), key=lambda x: -x[1])[:8]])
for r in sorted(data, key=lambda r: -num(r, "# Samples"))[:top_n]:
    st = sorted([(num(r, s), s[6:]) for s in stalls], reverse=True)[:2]
    print(str(num(r, "# Samples")).rjust(6), str(num(r, "Instructions Executed")).rjust(9), r[ci["Source"]][:100].ljust(100), st)
