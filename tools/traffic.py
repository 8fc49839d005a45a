"""Sums DRAM bytes / time per kernel family from an ncu --csv log with dram__bytes_* metrics."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, data = None, []
for r in rows:
    if 'Kernel Name' in r: hdr = r; continue
    if hdr and len(r) == len(hdr): data.append(dict(zip(hdr, r)))
K = collections.OrderedDict()
for d in data:
    i = int(d['ID']); K.setdefault(i, {'name': d['Kernel Name'].split('(')[0]})
    v = float(d['Metric Value'].replace(',', ''))
    if d['Metric Name'].startswith('dram'):
        v *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[d['Metric Unit']]
    K[i][d['Metric Name']] = v
agg = collections.OrderedDict()
for i, k in K.items():
    a = agg.setdefault(k['name'], collections.Counter()); a['n'] += 1; a['t'] += k['gpu__time_duration.sum'] / 1e3
    a['rd'] += k.get('dram__bytes_read.sum', 0); a['wr'] += k.get('dram__bytes_write.sum', 0)
tot = t = 0
for n, a in agg.items():
    print('%-46s n=%4d %9.1f us  rd %8.1f MB wr %8.1f MB' % (n[-46:], a['n'], a['t'], a['rd'] / 1e6, a['wr'] / 1e6))
    if 'conv3x3' in n: tot += a['rd'] + a['wr']; t += a['t']
print('conv: %.1f GB DRAM, %.2f ms' % (tot / 1e9, t / 1e3))
