import csv, re, collections, sys
src_csv, sass, kern, base, lo, hi = sys.argv[1:7]; lo=int(lo); hi=int(hi)
lines=[];cur=None;in_k=False
for ln in open(sass,errors='replace'):
    if ln.startswith('.text.'):
        in_k=kern in ln; continue
    if not in_k: continue
    m=re.search(r'//## File "([^"]+)", line (\d+)(.*)',ln)
    if m: cur=(m.group(1).split('/')[-1],int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S",ln): lines.append(cur)
rows=list(csv.reader(open(src_csv)))
name=None;hdr=None;data=[]
for r in rows:
    if len(r)>=2 and r[0]=='Kernel Name': name=r[1];hdr=None;continue
    if r and r[0]=='Address': hdr=r;continue
    if name and kern in name and hdr and len(r)==len(hdr): data.append(r)
n=len(lines)
cols=[i for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
agg=collections.Counter(); last=None; tot=0
for i,r in enumerate(data[:n]):
    key=lines[i]
    if key and key[0]==base: last=key[1]
    if last and lo<=last<=hi:
        for c in cols: agg[hdr[c]]+=int(r[c] or 0)
s=sum(agg.values())
print({k:round(100*v/s,1) for k,v in agg.most_common(10)}, 'samples',s)
