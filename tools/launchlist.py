import csv, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
steps=float(sys.argv[2]) if len(sys.argv)>2 else 1.0
hdr=None; acc=collections.Counter(); cnt=collections.Counter()
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        name=r[hdr.index('Kernel Name')].split('(')[0]; v=float(r[hdr.index('Metric Value')].replace(',',''))
        unit=r[hdr.index('Metric Unit')]
        if unit in ('usecond','us'): v/=1000
        elif unit in ('nsecond','ns'): v/=1e6
        elif unit in ('second','s'): v*=1000
        acc[name]+=v; cnt[name]+=1
tot=sum(acc.values())
print(f"total {tot/steps:.3f} ms per step over {steps:g} steps")
for n,v in acc.most_common(30): print(f"{n:45s} {cnt[n]/steps:7.1f} launches/step {v/steps:9.3f} ms/step {100*v/tot:5.1f}%")
