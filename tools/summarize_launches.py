"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel:
   python tools/summarize_launches.py gpurun_out/launches.csv > profiles/<name>.md
Times under ncu are serialised and cold-cache: compare SHARES, not absolutes."""
import collections
import csv
import re
import sys


def short(name):
    name = re.sub(r'\(.*$', '', name)
    name = name.replace('void ', '').replace('lvg::<unnamed>::', '').replace('at::native::', 'aten::')
    return name[:110]


def main(path):
    rows = [r for r in csv.reader(open(path, errors='replace')) if len(r) > 5]
    hdr = next(r for r in rows if 'Kernel Name' in r)
    ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows:
        if r is hdr or len(r) < len(hdr) or r[ix['Metric Name']] != 'gpu__time_duration.sum':
            continue
        val = float(r[ix['Metric Value']].replace(',', ''))
        if val != val:          # ncu occasionally reports nan for a launch
            continue
        unit = r[ix['Metric Unit']]
        us = val / 1e3 if unit in ('ns', 'nsecond') else val * 1e3 if unit in ('ms', 'msecond') else val
        k = short(r[ix['Kernel Name']])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(a[1] for a in agg.values())
    print(f'| kernel | launches | total us | share |')
    print('|---|---:|---:|---:|')
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'| `{k}` | {n} | {us:.1f} | {us / total * 100:.1f}% |')
    print(f'| **all** | {sum(a[0] for a in agg.values())} | {total:.1f} | 100% |')


if __name__ == '__main__':
    main(sys.argv[1])
