"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_*] --csv` launch list: per kernel count,
mean duration, share of the step, DRAM bytes per launch."""
import collections, csv, sys
for fn in sys.argv[1:]:
    with open(fn) as f:
        lines = [l for l in f if l.startswith('"')]
    per = collections.OrderedDict()
    for row in csv.DictReader(lines):
        k = (row['ID'], row['Kernel Name'])
        v = float(row['Metric Value'].replace(',', ''))
        u = row['Metric Unit']
        name = row['Metric Name']
        if name == 'gpu__time_duration.sum':
            v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(u, 1.0)
        elif name.startswith('dram'):
            v *= {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1.0)
        per.setdefault(k, {})[name] = v
    agg = collections.OrderedDict()
    for (i, k), d in per.items():
        a = agg.setdefault(k[:70], [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += d.get('gpu__time_duration.sum', 0.0)
        a[2] += d.get('dram__bytes_read.sum', 0.0) + d.get('dram__bytes_write.sum', 0.0)
        a[3] += d.get('smsp__inst_executed.sum', 0.0)
    tot = sum(a[1] for a in agg.values())
    print('%s   total %.1f us' % (fn, tot))
    for k, a in agg.items():
        print('  %3d x %10.1f us  %5.1f%%  dram %8.1f MB/launch  %s%s' % (a[0], a[1] / a[0], 100 * a[1] / tot, a[2] / a[0] / 1e6,
              ('%7.1f Minst  ' % (a[3] / a[0] / 1e6)) if a[3] else '', k))
