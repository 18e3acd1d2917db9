"""Aggregate an ncu 'cuda,sass' source-page CSV per CUDA source line: samples, instrs, stalls."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None
hdr = None
agg = collections.OrderedDict()
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if len(r) >= 2 and r[0] == "Function Name":
        continue
    if r and r[0] == "Line No":
        hdr = r; ix = {}
        for i, h in enumerate(hdr):
            ix.setdefault(h, i)
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    if r[ix["Address"]] != "-":      # SASS rows carry an address; CUDA rows hold the aggregate
        continue
    key = (cur_file, line)
    s = int(r[ix["# Samples"]] or 0)
    inst = int(r[ix["Instructions Executed"]] or 0)
    st = {h: int(r[i] or 0) for h, i in ix.items() if h.startswith("stall_") and "Not Issued" not in h}
    agg[key] = (s, inst, r[1].strip()[:80], st)
tot = sum(v[0] for v in agg.values()); toti = sum(v[1] for v in agg.values())
print("total samples", tot, "instr", toti)
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:N]:
    top = sorted(((c, h[6:]) for h, c in v[3].items()), reverse=True)[:3]
    print(f"{key[0]}:{key[1]:4d} {v[0]:6d} {100*v[0]/tot:5.1f}% inst {100*v[1]/toti:5.1f}%  {v[2]:80s} {top}")
