"""Summarise a long-format `ncu --metrics ... --csv --log-file X.csv` launch list: one line per launch (in order) and
per-kernel totals.  usage: python tools/ncu_launches.py X.csv [--per-launch [--min-us T]]"""
import csv, sys, collections, re
path = sys.argv[1]
per_launch = '--per-launch' in sys.argv
min_us = float(sys.argv[sys.argv.index('--min-us') + 1]) if '--min-us' in sys.argv else 0.0     # per-launch lines only for launches at least this long
lines = [l for l in open(path) if l.startswith('"')]
rows = list(csv.DictReader(lines))
L = collections.OrderedDict()
for r in rows:
    k = int(r['ID'])
    d = L.setdefault(k, {'name': re.sub(r'\(.*', '', r['Kernel Name']), 'grid': r['Grid Size'], 'block': r['Block Size']})
    v = float(r['Metric Value'].replace(',', ''))
    if r['Metric Unit'] in ('Kbyte',): v *= 1e3
    if r['Metric Unit'] in ('Mbyte',): v *= 1e6
    if r['Metric Unit'] in ('Gbyte',): v *= 1e9
    if r['Metric Name'] == 'gpu__time_duration.sum':
        v = v * {'ns': 1e-3, 'us': 1.0, 'usecond': 1.0, 'ms': 1e3, 'msecond': 1e3, 'nsecond': 1e-3, 'second': 1e6}[r['Metric Unit']]
    d[r['Metric Name']] = v
tot = collections.OrderedDict()
T = sum(d.get('gpu__time_duration.sum', 0) for d in L.values())
for k, d in L.items():
    us = d.get('gpu__time_duration.sum', 0)
    by = d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0)
    if per_launch and us >= min_us:
        print(f"{k:4d} {d['name'][:44]:44s} grid {d['grid']:>12s} {us:9.1f} us  dram {by/1e6:9.1f} MB  {by/us/1e3 if us else 0:7.0f} GB/s  "
              f"l1tex {d.get('l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 0):5.1f}%  lts {d.get('lts__throughput.avg.pct_of_peak_sustained_elapsed', 0):5.1f}%  "
              f"dram {d.get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 0):5.1f}%")
    t = tot.setdefault(d['name'], [0, 0.0, 0.0])
    t[0] += 1; t[1] += us; t[2] += by
print(f"-- {len(L)} launches, {T/1e3:.3f} ms of kernel time (cold-cache, serialised)")
for n, (c, us, by) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{n[:44]:44s} x{c:3d} {us/1e3:8.3f} ms {100*us/T:5.1f}%  dram {by/1e9:7.3f} GB  {by/us/1e3 if us else 0:7.0f} GB/s")
