#!/usr/bin/env python
"""Print the handful of ncu metrics the gather kernels are judged by from one or more .ncu-rep files.

    python tools/ncu_keys.py gpurun_out/x/prof_*.ncu-rep
"""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__cycles_elapsed.max',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__warps_eligible.avg.per_cycle_active', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum']
for path in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('==', path, r[hdr.index('Kernel Name')][:60] if 'Kernel Name' in hdr else '')
        for i, h in enumerate(hdr):
            if h in KEYS or ('pcsamp_warps_issue_stalled' in h and 'not_issued' not in h and r[i] not in ('0', '')):
                print('  %-75s %-14s %s' % (h, units[i], r[i]))
