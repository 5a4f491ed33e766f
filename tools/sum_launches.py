import csv,collections,sys
rows=list(csv.reader(open(sys.argv[1])))
hi=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
h=rows[hi]; ix={k:i for i,k in enumerate(h)}
tot=collections.Counter(); cnt=collections.Counter()
for r in rows[hi+1:]:
    if len(r)<len(h): continue
    n=r[ix["Kernel Name"]].split("(")[0]; v=float(r[ix["Metric Value"]].replace(",","")); u=r[ix["Metric Unit"]]
    v = v/1e3 if u=="ns" else v*1e3 if u=="ms" else v
    tot[n]+=v; cnt[n]+=1
for n,v in tot.most_common(): print("%-34s n=%3d avg %9.2f us"%(n[:34],cnt[n],v/cnt[n]))
