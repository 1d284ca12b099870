"""Summarise `ncu --page source --csv` (SASS view) of ONE kernel: executed warp instructions by opcode, the hottest
instructions by stall samples, and the stall-reason totals.
    ncu -i rep.ncu-rep --page source --csv --kernel-name <name> > k.csv ; python tools/ncu_sass_summary.py k.csv [top]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) >= len(hdr) - 2]
def num(r, k):
    try:
        return float(r[ix[k]])
    except (ValueError, KeyError, IndexError):
        return 0.0
tot_inst = sum(num(r, "Instructions Executed") for r in data)
tot_samp = sum(num(r, "# Samples") for r in data)
print(f"instructions executed (warp-level): {tot_inst:,.0f}   samples: {tot_samp:,.0f}   static instructions: {len(data)}")
ops, ops_s = Counter(), Counter()
for r in data:
    src = r[ix["Source"]].strip()
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    op = op.split(".")[0]
    ops[op] += num(r, "Instructions Executed")
    ops_s[op] += num(r, "# Samples")
print("by opcode (executed %, samples %):")
for op, n in ops.most_common(22):
    print(f"  {op:10s} {100 * n / tot_inst:5.1f} %   {100 * ops_s[op] / max(tot_samp, 1):5.1f} %")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("stall reasons (all samples):")
st = {s: sum(num(r, s) for r in data) for s in stalls}
for s, n in sorted(st.items(), key=lambda kv: -kv[1])[:10]:
    print(f"  {s:24s} {100 * n / max(tot_samp, 1):5.1f} %")
print(f"hottest {top} instructions by samples:")
order = sorted(range(len(data)), key=lambda i: -num(data[i], "# Samples"))[:top]
for i in sorted(order):
    r = data[i]
    best = max(stalls, key=lambda s: num(r, s))
    print(f"  #{i:5d} {num(r, '# Samples'):7.0f} smp  {num(r, 'Instructions Executed'):11,.0f} exe  {best[6:]:14s} {r[ix['Source']].strip()[:90]}")
