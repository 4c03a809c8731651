"""Kernel shares of one timed step from an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv CMD`).
usage: python tools/kernel_shares.py LAUNCHES.csv N_LAUNCHES [--from FIRST_ROW] [--title "..."]
The timed step is rows [FIRST_ROW, FIRST_ROW + N_LAUNCHES) of the list (default: the last N rows); rows are counted over
all kernel launches of the process, torch's included.  `--find KERNEL` prints the rows at which KERNEL was launched."""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path, last_n = sys.argv[1], int(sys.argv[2])
    title = sys.argv[sys.argv.index('--title') + 1] if '--title' in sys.argv else ''
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith('=='))]
    hdr = rows[0]
    ci = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[1:] if len(r) == len(hdr) and r[ci['Metric Name']] == 'gpu__time_duration.sum']
    if '--find' in sys.argv:
        k = sys.argv[sys.argv.index('--find') + 1]
        print([i for i, r in enumerate(data) if k in r[ci['Kernel Name']]])
        return
    if '--from' in sys.argv:
        a = int(sys.argv[sys.argv.index('--from') + 1])
        data = data[a:a + last_n]
    else:
        data = data[-last_n:]
    agg = OrderedDict()
    for r in data:
        name = re.sub(r'^void\s+', '', r[ci['Kernel Name']])
        name = re.sub(r'^nbc::', '', name)
        name = re.sub(r'[<(].*$', '', name)
        unit, v = r[ci['Metric Unit']], float(r[ci['Metric Value']].replace(',', ''))
        ms = v * {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'second': 1e3}.get(unit, 1e-6)
        n, t = agg.get(name, (0, 0.0))
        agg[name] = (n + 1, t + ms)
    tot = sum(t for _, t in agg.values())
    if title:
        print(title)
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-34s %4d launches  %8.3f ms  %4.1f %%' % (name[:34], n, t, 100 * t / tot))
    print('total %.3f ms, %d launches' % (tot, len(data)))


if __name__ == '__main__':
    main()
