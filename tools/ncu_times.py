import sys,csv,collections
rows=list(csv.reader(l for l in sys.stdin if l.startswith('"')))
acc=collections.defaultdict(list)
for r in rows[1:]:
    if 'gpu__time_duration' in r[-3]: acc[r[4][:40]].append(float(r[-1]))
for k,v in acc.items(): print(k, len(v), 'mean us %.1f'%(sum(v)/len(v)/1e3))
