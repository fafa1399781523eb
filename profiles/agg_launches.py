import csv, collections, re, sys
with open(sys.argv[1]) as f:
    lines=[l for l in f if not l.startswith('==')]
r=csv.DictReader(lines)
agg=collections.OrderedDict(); tot=0
for row in r:
    name=row['Kernel Name']; v=float(row['Metric Value'].replace(',','')); unit=row['Metric Unit']
    if unit=='ns': v/=1e6
    elif unit=='us': v/=1e3
    name=re.sub(r'\(.*','',name)[:72]
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=v; tot+=v
print('total ms %.2f launches %d'%(tot, sum(a[0] for a in agg.values())))
for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:int(sys.argv[2]) if len(sys.argv)>2 else 45]:
    print('%8.2f ms %5.1f%% x%-4d %s'%(t,100*t/tot,n,k))
