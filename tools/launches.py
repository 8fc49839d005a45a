import csv, collections, sys
path = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/launches.csv'
lines=[l for l in open(path) if not l.startswith('==')]
r=list(csv.DictReader(lines))
seq=collections.OrderedDict()
for row in r:
    seq.setdefault(row['ID'], {'name':row['Kernel Name'][:40]})[row['Metric Name']]=float(row['Metric Value'].replace(',',''))
L=list(seq.values())
convs=[x for x in L if 'conv3x3_tc' in x['name']]
T='gpu__time_duration.sum'; P='sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'
print(len(convs))
for x in convs[3:9]: print(x['name'][:28], round(x[T]/1e3,1), round(x.get(P,0),1))
print('tail')
for x in convs[-7:]: print(x['name'][:28], round(x[T]/1e3,1), round(x.get(P,0),1))
agg=collections.defaultdict(float)
for x in L: agg[x['name']]+=x[T]
for k,v in sorted(agg.items(), key=lambda kv:-kv[1]): print(k, round(v/1e6,3),'ms')
