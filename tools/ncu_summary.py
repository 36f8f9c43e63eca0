"""Markdown summary of an ncu report (`ncu --set full`): per kernel instance the duration, DRAM traffic,
pipe utilisation, occupancy limits and top stall reasons.
   python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.md"""
import csv
import io
import subprocess
import sys

METRICS = [
    ('gpu__time_duration.sum', 'duration'),
    ('dram__bytes_read.sum', 'DRAM read'),
    ('dram__bytes_write.sum', 'DRAM write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput % of ncu peak'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput %'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy %'),
    ('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'FMA pipe active %'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe active %'),
    ('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'LSU pipe %'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy %'),
    ('launch__registers_per_thread', 'registers/thread'),
    ('launch__occupancy_limit_registers', 'CTAs/SM limit (registers)'),
    ('launch__occupancy_limit_shared_mem', 'CTAs/SM limit (shared memory)'),
    ('launch__grid_size', 'grid'),
    ('smsp__inst_executed.sum', 'warp instructions'),
    ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'shared-memory wavefronts'),
    ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'shared-memory bank conflicts'),
]


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    print(f'# ncu summary of `{path}`\n')
    for r in rows[2:]:
        print(f"## `{r[ix['Kernel Name']][:120]}`\n")
        print('| metric | value |\n|---|---|')
        for key, label in METRICS:
            if key in ix and r[ix[key]] not in ('', 'n/a'):
                print(f'| {label} | {r[ix[key]]} {units[ix[key]]} |')
        stalls = [(float(r[i]), h) for h, i in ix.items()
                  if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio') and r[i] not in ('', 'n/a')]
        top = ', '.join(f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {v:.2f}"
                        for v, h in sorted(stalls, reverse=True)[:6])
        print(f'| top stall reasons (warps per issue) | {top} |\n')


if __name__ == '__main__':
    main(sys.argv[1])
