#!/usr/bin/env python3
"""Print the metrics we track from an .ncu-rep (raw page), one kernel launch per block.
    python tools/ncu_summary.py gpurun_out/prof_filter_r1.ncu-rep [extra-metric-substring ...]
    python tools/ncu_summary.py gpurun_out/prof_step.ncu-rep --split profiles/r2d   # one profiles/r2d_ncu_<kernel>.txt per kernel"""
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


SHORT = (('filter_sym', 'filter'), ('sym_gather', 'gather'), ('prepass_kernel', 'prepass'), ('accumulate', 'accum'),
         ('filter_warp', 'filter_stream'), ('nonfinite_fixup', 'fixup'))


def main():
    rep = sys.argv[1]
    extra = sys.argv[2:]
    split = None
    if '--split' in extra:
        k = extra.index('--split')
        split = extra[k + 1]
        extra = extra[:k] + extra[k + 2:]
    head = '\n'.join(source_shas())
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u = rows[0], rows[1]
    blocks = []
    for v in rows[2:]:
        name = v[h.index('Kernel Name')]
        lines = ['--- %s' % name[:100]]
        for i, n in enumerate(h):
            if n in WANT or 'pipe_fma' in n or 'pipe_fp' in n or 'pcsamp_warps_issue_stalled' in n or any(e in n for e in extra):
                lines.append('%-84s %-14s %s' % (n, u[i], v[i]))
        blocks.append((name, '\n'.join(lines)))
    if not split:
        print(head)
        for _, b in blocks:
            print(b)
        return
    done = set()
    for name, b in blocks:
        short = next((s for key, s in SHORT if key in name), None)
        if short is None or short in done:
            continue  # one launch per kernel
        done.add(short)
        path = '%s_ncu_%s.txt' % (split, short)
        with open(path, 'w') as f:
            f.write(head + '\n# ncu --set full --clock-control none --import-source on, one launch of the default 4K step '
                    '(`tools/gpu_session.sh ncu_step`), from %s\n' % rep.split('/')[-1] + b + '\n')
        print('wrote', path)


if __name__ == '__main__':
    main()
