import csv, sys, collections
path = sys.argv[1]; skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
lines=[l for l in open(path) if not l.startswith('==')]
seq=[]
for row in csv.DictReader(lines):
    if row.get('Metric Name')!='gpu__time_duration.sum': continue
    v=float(row['Metric Value'].replace(',','')); u=row['Metric Unit']
    v = v/1000.0 if u=='ns' else (v*1000.0 if u=='ms' else v)
    seq.append((row['Kernel Name'][:64], v, row.get('Grid Size','')))
agg=collections.OrderedDict()
for n,v,g in seq[skip:]:
    a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=v
tot=sum(v for _,v,_ in seq[skip:])
for n,(c,v) in sorted(agg.items(), key=lambda kv:-kv[1][1]): print(f"{v:9.1f} us {c:4d}x  {n}")
print("total", tot, "launches", len(seq[skip:]))
