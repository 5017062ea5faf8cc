"""Summarise an `ncu --page source --csv` export: top instructions by warp-stall samples, and samples grouped by
source-line ranges.  usage: ncu -i X.ncu-rep --page source --csv [--kernel-name ...] > src.csv; python scripts/ncu_top_stalls.py src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) >= len(hdr) - 2 and r[0].startswith("0x")]
S = ix["# Samples"]
tot = sum(int(r[S]) for r in data)
print("kernel:", rows[0][1] if rows[0] else "?", "| total samples", tot, "| instructions", len(data))
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[ix[h]]) for r in data) for h in stall_cols}
print("stall mix:", ", ".join(f"{k[6:]} {v / max(tot, 1):.1%}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for pos, r in sorted(enumerate(data), key=lambda pr: -int(pr[1][S]))[:n]:
    st = sorted(((h[6:], int(r[ix[h]])) for h in stall_cols if int(r[ix[h]]) > 0), key=lambda kv: -kv[1])[:3]
    print(f"#{pos:5d} {int(r[S]):7d} {int(r[S]) / max(tot, 1):6.1%}  {r[ix['Source']].strip()[:72]:72s} {st}")
