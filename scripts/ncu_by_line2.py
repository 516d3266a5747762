import csv, re, collections, sys
src_csv, sass, kern, srcfile, minline = sys.argv[1:6]
minline=int(minline)
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
n=len(lines); print('sass instr',n,'ncu rows',len(data))
si=hdr.index('# Samples'); ie=hdr.index('Instructions Executed')
src=open(srcfile).read().split('\n'); base=srcfile.split('/')[-1]
agg=collections.defaultdict(lambda:[0,0]); last=None
for i,r in enumerate(data[:n]):
    key=lines[i]
    if key and key[0]==base and key[1]>=minline: last=key[1]
    agg[last][0]+=int(r[si] or 0); agg[last][1]+=int(r[ie] or 0)
tot=sum(v[0] for v in agg.values()); tote=sum(v[1] for v in agg.values())
print('samples',tot,'inst',tote)
for k in sorted(k for k in agg if k):
    v=agg[k]
    if v[1]>0.012*tote or v[0]>0.012*tot:
        print(f"{k:4d} smp {v[0]:5d} {100*v[0]/tot:4.1f}% inst {v[1]:8d} {100*v[1]/tote:4.1f}% | {src[k-1].strip()[:120]}")
