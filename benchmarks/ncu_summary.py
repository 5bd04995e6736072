#!/usr/bin/env python
"""Summarises an .ncu-rep (raw page) into the handful of metrics DESIGN.md and
bench.py quote.  Usage: python benchmarks/ncu_summary.py file.ncu-rep"""
import csv
import subprocess
import sys

KEYS = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'dram__bytes_read.sum.per_second',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__warps_active.avg.pct_of_peak_sustained_active',
    'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
    'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
    'launch__shared_mem_per_block_dynamic', 'smsp__inst_executed.sum',
    'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.max',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
]


def main(path):
  out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'],
                       capture_output=True, text=True, check=True).stdout
  rows = list(csv.reader(out.splitlines()))
  hdr, units = rows[0], rows[1]
  for r in rows[2:]:
    name = r[hdr.index('Kernel Name')]
    print(f'== {name}')
    for k in KEYS:
      if k in hdr:
        i = hdr.index(k)
        print(f'{k:92s} {r[i]:>18s} {units[i]}')


if __name__ == '__main__':
  main(sys.argv[1])
