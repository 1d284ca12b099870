"""Summarise `ncu --page source --csv` output: per kernel, the SASS instructions with the most stall samples."""
import csv
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1]
        hdr = rows[i + 1]
        j = i + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            if len(rows[j]) == len(hdr):
                body.append(rows[j])
            j += 1
        si = hdr.index("# Samples")
        stall_cols = [k for k, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        total = sum(int(r[si] or 0) for r in body)
        print(f"\n=== {name}  total samples {total}")
        agg = {}
        for r in body:
            for k in stall_cols:
                agg[hdr[k]] = agg.get(hdr[k], 0) + int(r[k] or 0)
        print("   stall totals:", ", ".join(f"{k[6:]}={v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
        for idx, r in sorted(enumerate(body), key=lambda t: -int(t[1][si] or 0))[:topn]:
            st = ", ".join(f"{hdr[k][6:]}={r[k]}" for k in stall_cols if int(r[k] or 0) > 0)
            print(f"   {int(r[si]):6d}  #{idx:4d} {r[1].strip()[:70]:70s} {st}")
        i = j
    else:
        i += 1
