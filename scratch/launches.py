import csv, collections, sys
with open(sys.argv[1]) as f:
    lines=[l for l in f if not l.startswith('==')]
r=csv.DictReader(lines)
agg=collections.OrderedDict()
for row in r:
    name=row['Kernel Name'].split('(')[0].replace('<unnamed>::',''); v=float(row['Metric Value'].replace(',','')); u=row['Metric Unit']
    if u=='ns': v/=1e3
    elif u=='ms': v*=1e3
    agg.setdefault(name,[]).append(v)
tot=sum(sum(v) for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])):
    print(f"{k[:50]:50s} n={len(v):4d} total={sum(v):10.1f}us  mean={sum(v)/len(v):9.1f}us min={min(v):9.1f} share={100*sum(v)/tot:5.1f}%")
