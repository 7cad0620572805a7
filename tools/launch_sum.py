import csv,collections,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=[r for r in rows if "Kernel Name" in r][0]
ix={h:i for i,h in enumerate(hdr)}
tot=collections.Counter(); cnt=collections.Counter()
for r in rows:
    if len(r)==len(hdr) and r[ix["Metric Name"]]=="gpu__time_duration.sum":
        k=r[ix["Kernel Name"]][:40]; v=float(r[ix["Metric Value"]].replace(",","")); u=r[ix["Metric Unit"]]
        v*= {"ns":1e-6,"us":1e-3,"ms":1.0,"usecond":1e-3,"nsecond":1e-6,"msecond":1.0}.get(u,1e-6)
        tot[k]+=v; cnt[k]+=1
s=sum(tot.values())
for k,v in tot.most_common(8): print(f"{k:42s} n={cnt[k]:3d} total {v:9.3f} ms share {v/s:.4f}")
