"""Top stalled SASS instructions of an ncu source-page CSV: python tools/ncu_stalls.py src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[ix[h]] or 0) for r in data) for h in stall_cols}
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]] or 0))[:N]
for i in sorted(order):
    r = data[i]
    top = sorted(((int(r[ix[h]] or 0), h) for h in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {int(r[ix['# Samples']]):6d}  {r[ix['Source']].strip()[:90]:90s} {top}")
