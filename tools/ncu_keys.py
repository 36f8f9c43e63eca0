"""Key metrics of an ncu report: python tools/ncu_keys.py file.ncu-rep [more.ncu-rep ...]  (reads with `ncu -i ... --page raw --csv`)"""
import csv
import io
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'l1tex__m_xbar2l1tex_read_bytes.sum.per_second', 'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'smsp__inst_executed.sum', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers']


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            print(f'## {path}: {d.get("Kernel Name", "")[:100]}')
            for k in WANT:
                if k in d:
                    print(f'  {k:95s} {d[k]:>16s} {units[hdr.index(k)]}')


if __name__ == '__main__':
    main()
