"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import collections
import csv
import re
import sys


def short(name):
    name = re.sub(r'\(.*', '', name)
    name = re.sub(r'^void ', '', name)
    name = name.replace('gd3::', '').replace('(anonymous namespace)::', '')
    return name[:90]


def main(path, skip=0):
    rows = list(csv.reader(open(path, errors='replace')))
    hdr = None
    agg = collections.OrderedDict()
    n = 0
    for r in rows:
        if 'Kernel Name' in r:
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        if d.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        n += 1
        if n <= skip:
            continue
        val = float(d['Metric Value'].replace(',', ''))
        a = agg.setdefault(short(d['Kernel Name']), [0, 0.0])
        a[0] += 1
        a[1] += val
    tot = sum(v[1] for v in agg.values())
    print(f'{n} launches, {tot / 1e6:.3f} ms total (cold-cache, serialised)')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
        print(f'{v[1] / 1e6:9.3f} ms {v[0]:6d}x avg {v[1] / v[0] / 1e3:9.1f} us {100 * v[1] / tot:5.1f}%  {k}')


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
