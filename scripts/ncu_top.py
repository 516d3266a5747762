"""Top stall-sample SASS lines of one kernel from `ncu -i X.ncu-rep --page source --csv` (helper for reading profiles)."""
import csv, sys
path, kern, topn = sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25
rows = list(csv.reader(open(path)))
cur = None; hdr = None; data = []
for r in rows:
    if len(r) >= 2 and r[0] == "Kernel Name":
        cur = r[1]; hdr = None; continue
    if r and r[0] == "Address":
        hdr = r; continue
    if cur and kern in cur and hdr and len(r) == len(hdr):
        data.append(r)
if not data:
    sys.exit("kernel not found")
si = hdr.index("# Samples"); src = hdr.index("Source"); ie = hdr.index("Instructions Executed")
stall_cols = [i for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
tot = sum(int(r[si] or 0) for r in data)
print("total samples", tot, "instructions", len(data), "warp-inst executed", sum(int(r[ie] or 0) for r in data))
agg = {}
for r in data:
    for i in stall_cols:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
print("stall mix:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
idx = sorted(range(len(data)), key=lambda i: -int(data[i][si] or 0))[:topn]
for i in sorted(idx):
    r = data[i]
    st = sorted(((int(r[c] or 0), hdr[c]) for c in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {int(r[si]):6d} {100*int(r[si])/tot:5.1f}%  exec={r[ie]:>7s}  {r[src].strip()[:70]:70s} {st}")
