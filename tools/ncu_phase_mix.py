"""Dynamic instruction mix of one kernel from `ncu -i X.ncu-rep --page source --csv`:
splits the SASS at BAR.SYNC instructions into phases and prints executed warp instructions,
stall samples and the top opcodes per phase.  python tools/ncu_phase_mix.py file.csv"""
import csv
import sys
from collections import Counter


def main(path):
    rows = list(csv.reader(open(path)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
    hdr = rows[hdr_i]
    c_src, c_exec, c_smp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
    c_sh = hdr.index('L1 Wavefronts Shared') if 'L1 Wavefronts Shared' in hdr else None
    c_shi = hdr.index('L1 Wavefronts Shared Ideal') if 'L1 Wavefronts Shared Ideal' in hdr else None
    phases, cur = [], dict(n=0, smp=0, ops=Counter(), sh=0, shi=0, first=None)
    total = 0
    for r in rows[hdr_i + 1:]:
        if len(r) <= c_exec:
            continue
        src = r[c_src].strip()
        ex = int(r[c_exec] or 0)
        toks = src.split()
        op = toks[1] if toks and toks[0].startswith('@') and len(toks) > 1 else (toks[0] if toks else '?')
        op = op.split('.')[0].rstrip(';')
        if cur['first'] is None:
            cur['first'] = r[0]
        cur['n'] += ex
        cur['smp'] += int(r[c_smp] or 0)
        cur['ops'][op] += ex
        if c_sh is not None:
            cur['sh'] += int(r[c_sh] or 0)
            cur['shi'] += int(r[c_shi] or 0)
        total += ex
        if op == 'BAR':
            phases.append(cur)
            cur = dict(n=0, smp=0, ops=Counter(), sh=0, shi=0, first=None)
    phases.append(cur)
    tot_smp = sum(p['smp'] for p in phases) or 1
    print(f'total warp instructions {total}')
    for i, p in enumerate(phases):
        top = ', '.join(f'{k} {v * 100 // max(p["n"], 1)}%' for k, v in p['ops'].most_common(8))
        print(f'phase {i}: {p["n"]:>12d} inst ({p["n"] * 100 / max(total, 1):5.1f}%)  samples {p["smp"] * 100 / tot_smp:5.1f}%  '
              f'smem wavefronts {p["sh"]} (ideal {p["shi"]})\n    {top}')


if __name__ == '__main__':
    main(sys.argv[1])
