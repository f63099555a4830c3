"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum[,dram__bytes_*] --csv` launch list (cold-cache, serialised
per-launch times: compare SHARES, not absolutes).   python tools/summarize_launches.py launches.csv"""
import collections, csv, re, sys

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
r = csv.reader(lines)
hdr = next(r)
ix = {h: i for i, h in enumerate(hdr)}
data = collections.OrderedDict()
for row in r:
    if len(row) < len(hdr):
        continue
    d = data.setdefault(row[ix['ID']], {'name': row[ix['Kernel Name']], 'grid': row[ix['Grid Size']]})
    val, unit = float(row[ix['Metric Value']].replace(',', '')), row[ix['Metric Unit']]
    if row[ix['Metric Name']] == 'gpu__time_duration.sum':
        d['us'] = val / 1000 if unit in ('ns', 'nsecond') else val
    else:
        d[row[ix['Metric Name']]] = val
agg, tot = collections.OrderedDict(), 0.0
for d in data.values():
    key = (re.sub(r'\(.*', '', d['name']).replace('void ', '').replace('gitb200::', '')[:60], d['grid'])
    a = agg.setdefault(key, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += d.get('us', 0.0)
    a[2] += d.get('dram__bytes_read.sum', 0.0) + d.get('dram__bytes_write.sum', 0.0)
    tot += d.get('us', 0.0)
print('%d launches, %.1f us in total' % (len(data), tot))
for (n, g), a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    extra = '  dram %.1f MB/launch' % (a[2] / a[0] / 1e6) if a[2] else ''
    print('  %-60s grid=%-14s n=%4d total=%9.1f us avg=%8.2f us %5.1f%%%s' % (n, g, a[0], a[1], a[1] / a[0], 100 * a[1] / tot, extra))
