import csv, collections, re, sys
path = sys.argv[1]
lines=[l for l in open(path) if not l.startswith('==')]
agg=collections.defaultdict(lambda:[0,0.0,0.0,0.0])
for row in csv.DictReader(lines):
    name=row['Kernel Name']; m=row['Metric Name']; v=float(row['Metric Value'].replace(',',''))
    key=re.sub(r'\(.*','',name)[:70]
    a=agg[key]
    if m=='gpu__time_duration.sum': a[0]+=1; a[1]+=v
    elif m=='dram__bytes_read.sum': a[2]+=v
    elif m=='dram__bytes_write.sum': a[3]+=v
tot=sum(a[1] for a in agg.values())
print("total kernel time %.3f ms, %d launches" % (tot/1e6, sum(a[0] for a in agg.values())))
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1])[:int(sys.argv[2]) if len(sys.argv)>2 else 30]:
    print("%-72s n=%4d  %8.3f ms  %5.1f%%  rd %7.1f MB wr %7.1f MB  %6.0f GB/s" % (k,a[0],a[1]/1e6,100*a[1]/tot,a[2]/1e6,a[3]/1e6,(a[2]+a[3])/max(a[1],1)))
