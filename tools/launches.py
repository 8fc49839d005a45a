"""Summarises an ncu --csv launch list (gpu__time_duration [+ tensor pipe %]) per kernel and per layer position."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, data = None, []
for r in rows:
    if 'Kernel Name' in r: hdr = r; continue
    if hdr and len(r) == len(hdr): data.append(dict(zip(hdr, r)))
t, p, names = {}, {}, {}
for d in data:
    i = int(d['ID']); names[i] = d['Kernel Name'].split('(')[0][-40:]
    v = float(d['Metric Value'].replace(',', ''))
    if d['Metric Name'].startswith('gpu__time'): t[i] = v / 1000
    else: p[i] = v
ids = sorted(t)
tot = collections.Counter(); cnt = collections.Counter()
for i in ids: tot[names[i]] += t[i]; cnt[names[i]] += 1
for k, v in tot.most_common(): print('%-42s n=%3d total %9.1f us' % (k, cnt[k], v))
print('total', sum(t.values()))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 14
for i in ids[:n] + ids[-n:]: print(i, names[i], round(t[i], 1), p.get(i))
