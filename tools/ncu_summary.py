#!/usr/bin/env python3
"""Print the metrics we track from an .ncu-rep (raw page), one kernel launch per block.
    python tools/ncu_summary.py gpurun_out/prof_filter_r1.ncu-rep [extra-metric-substring ...]"""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_bytes.sum', 'sm__cycles_elapsed.avg', 'smsp__cycles_active.avg',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct']


def source_shas():
    """sha of the kernel sources at capture time, in the form bench.py's ncu_traffic() checks"""
    import os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
    import bench
    return ['# source_sha[%s]: %s' % (k, bench.source_sha(k)) for k in bench.KERNEL_SOURCES]


def main():
    rep = sys.argv[1]
    extra = sys.argv[2:]
    print('\n'.join(source_shas()))
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u = rows[0], rows[1]
    for v in rows[2:]:
        print('--- %s' % v[h.index('Kernel Name')][:100])
        for i, n in enumerate(h):
            if n in WANT or 'pipe_fma' in n or 'pipe_fp' in n or any(e in n for e in extra):
                print('%-84s %-14s %s' % (n, u[i], v[i]))


if __name__ == '__main__':
    main()
