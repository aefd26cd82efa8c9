"""Hot instructions of one captured kernel: python tools/ncu_hot.py x.ncu-rep [top]"""
import csv
import subprocess
import sys


def f(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = [i for i, r in enumerate(rows) if "Source" in r][0]
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
key = "Warp Stall Sampling (All Samples)"
body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
tot = sum(f(r[ci[key]]) for r in body)
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
agg = {}
for r in body:
    op = r[ci["Source"]].split()[0] if r[ci["Source"]].split() else "?"
    if op.startswith("@"):
        op = r[ci["Source"]].split()[1]
    agg[op] = agg.get(op, 0.0) + f(r[ci[key]])
print("by opcode (share of stall samples at that instruction):")
for op, v in sorted(agg.items(), key=lambda kv: -kv[1])[:14]:
    print(f"  {100 * v / tot:6.2f}%  {op}")
print("instructions executed:", sum(f(r[ci['Instructions Executed']]) for r in body))
for r in sorted(body, key=lambda r: -f(r[ci[key]]))[:top]:
    print(f"{100 * f(r[ci[key]]) / tot:6.2f}%  {r[ci['Source']].strip()[:100]}")
