"""Aggregate ncu per-instruction samples (`ncu -i X.ncu-rep --page source --csv`) by CUDA source line, using the line
table of the cubin (`nvdisasm -g -c file.cubin`).  usage: ncu_by_line.py src.csv file.sass kernel_substring [topn]"""
import csv, re, sys, collections
src_csv, sass, kern = sys.argv[1:4]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 30
# --- line table from nvdisasm
lines = []; cur = None; in_k = False
for ln in open(sass, errors="replace"):
    if ln.startswith(".text."):
        in_k = kern in ln; continue
    if ln.startswith(".section") or ln.startswith("\t.section"):
        in_k = False if not ln.startswith(".text.") else in_k
    if not in_k: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
        lines.append(cur)
# --- ncu rows
rows = list(csv.reader(open(src_csv)))
name = None; hdr = None; data = []
for r in rows:
    if len(r) >= 2 and r[0] == "Kernel Name": name = r[1]; hdr = None; continue
    if r and r[0] == "Address": hdr = r; continue
    if name and kern in name and hdr and len(r) == len(hdr): data.append(r)
n = len(lines)
print(f"nvdisasm instructions {n}, ncu rows {len(data)} ({len(data)/max(n,1):.2f} launches)")
si = hdr.index("# Samples"); ie = hdr.index("Instructions Executed")
agg = collections.defaultdict(lambda: [0, 0])
for i, r in enumerate(data):
    key = lines[i % n] if n else None
    agg[key][0] += int(r[si] or 0); agg[key][1] += int(r[ie] or 0)
tot = sum(v[0] for v in agg.values()); tote = sum(v[1] for v in agg.values())
print("total samples", tot, "warp-inst", tote)
for key, (s_, e_) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"{str(key):34s} samples {s_:6d} {100*s_/max(tot,1):5.1f}%   inst {e_:9d} {100*e_/max(tote,1):5.1f}%")
