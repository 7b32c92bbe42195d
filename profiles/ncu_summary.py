#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, without a GPU): python profiles/ncu_summary.py gpurun_out/x.ncu-rep"""
import csv, subprocess, sys, io
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__grid_size', 'launch__block_size',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max',
        'smsp__inst_executed_pipe_fp64.sum', 'smsp__inst_executed_pipe_alu.sum', 'smsp__inst_executed_pipe_fma.sum',
        'smsp__inst_executed_pipe_lsu.sum', 'smsp__inst_executed_pipe_xu.sum', 'smsp__inst_executed_pipe_cbu.sum',
        'smsp__inst_executed_pipe_uniform.sum', 'smsp__inst_executed_pipe_adu.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct']
def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index('Kernel Name')
    for r in data:
        print('==', r[name_i][:80])
        for i, h in enumerate(hdr):
            if h in KEYS or (h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')):
                v = r[i]
                try:
                    if float(v.replace(',', '')) == 0: continue
                except ValueError:
                    pass
                print(f'  {h:90s} {v:>18s} {units[i]}')
if __name__ == '__main__':
    main(sys.argv[1])
