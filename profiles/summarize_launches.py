import csv, re, collections, sys
rows=[]
with open(sys.argv[1]) as f:
    lines=[l for l in f if not l.startswith('==')]
for row in csv.DictReader(lines): rows.append(row)
agg=collections.OrderedDict()
for row in rows:
    name=re.sub(r'\(.*','',row['Kernel Name']).replace('void ','').replace('b200::','')
    v=float(row['Metric Value'].replace(',',''))
    unit=row['Metric Unit']
    if unit=='ns': v/=1000
    elif unit=='ms': v*=1000
    key=(name,row.get('Grid Size',''))
    a=agg.setdefault(key,[0,0.0]); a[0]+=1; a[1]+=v
tot=sum(a[1] for a in agg.values())
for k,a in sorted(agg.items(), key=lambda x:-x[1][1]):
    print(f"{a[1]/tot*100:5.1f}%  n={a[0]:4d}  avg={a[1]/a[0]:8.1f}us  {k[0][:60]}  grid={k[1]}")
print("total us",tot, "kernels", len(rows))
