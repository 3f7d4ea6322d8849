"""Top stall sites of an `ncu --page source --csv --print-source sass` export."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]; col = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[col["# Samples"]] or 0) for r in data)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("total samples", tot, "instructions", len(data))
idx = sorted(range(len(data)), key=lambda i: -int(data[i][col["# Samples"]] or 0))[:topn]
for i in sorted(idx):
    r = data[i]
    n = int(r[col["# Samples"]] or 0)
    top = sorted(((int(r[col[s]] or 0), s) for s in stalls), reverse=True)[:2]
    print(f"{i:5d} {100.0*n/tot:5.1f}%  {r[col['Source']].strip()[:70]:70s} exec={r[col['Instructions Executed']]:>8s} " + " ".join(f"{s[6:]}={v}" for v, s in top if v))
