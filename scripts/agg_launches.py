import csv, collections, sys
lines=[l for l in open(sys.argv[1]) if not l.startswith('==')]
rows=list(csv.DictReader(lines))
idx=[i for i,x in enumerate(rows) if 'k_sum_dups' in x['Kernel Name']]
seg=rows[idx[-1]:]
tot=collections.Counter(); cnt=collections.Counter()
for x in seg:
    name=x['Kernel Name'].split('(')[0].split('<')[0]
    t=float(x['Metric Value'].replace(',',''))
    tot[name]+=t; cnt[name]+=1
T=sum(tot.values())
for k,v in tot.most_common():
    print("%-28s n=%4d  %10.1f us  %5.1f%%  avg %7.1f us"%(k,cnt[k],v/1e3,100*v/T,v/1e3/cnt[k]))
print("total us",T/1e3)
