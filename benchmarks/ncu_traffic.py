#!/usr/bin/env python
"""Reads the ncu captures made by benchmarks/ncu_traffic.sh (gpurun_out/
r2_ncu_<workload>.ncu-rep) and writes
  profiles/r2_traffic.json            dram bytes per launch etc. per workload
                                      (bench.py puts it into roofline.traffic)
  profiles/r2_ncu_<workload>_summary.txt   the metrics DESIGN.md quotes
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
    'sm__warps_active.avg.pct_of_peak_sustained_active',
    'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
    'launch__shared_mem_per_block_dynamic', 'sm__cycles_elapsed.avg',
    'lts__t_sectors.avg.pct_of_peak_sustained_elapsed',
]


def to_bytes(value, unit):
  scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}
  return float(value.replace(',', '')) * scale.get(unit, 1)


def to_us(value, unit):
  scale = {'ns': 1e-3, 'us': 1, 'ms': 1e3, 's': 1e6}
  return float(value.replace(',', '')) * scale.get(unit, 1)


def main():
  out = {}
  for name in ('rmse_acc', 'crps_sweep', 'regrid', 'spectrum_sweep',
               'spectrum_latsum'):
    rep = os.path.join(ROOT, 'gpurun_out', f'r2_ncu_{name}.ncu-rep')
    if not os.path.exists(rep):
      print('missing', rep)
      continue
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    r = rows[2]
    col = {h: i for i, h in enumerate(hdr)}
    rd = to_bytes(r[col['dram__bytes_read.sum']], units[col['dram__bytes_read.sum']])
    wr = to_bytes(r[col['dram__bytes_write.sum']], units[col['dram__bytes_write.sum']])
    dur = to_us(r[col['gpu__time_duration.sum']], units[col['gpu__time_duration.sum']])
    out[name] = {'kernel': r[col['Kernel Name']], 'dram_bytes_read': rd,
                 'dram_bytes_write': wr, 'dram_bytes_per_launch': rd + wr,
                 'duration_us_under_ncu': dur,
                 'source': f'ncu --set full --clock-control none, {os.path.basename(rep)}'}
    lines = [f'# {name}: {r[col["Kernel Name"]]}',
             f'# from {os.path.basename(rep)} (ncu --set full --clock-control none)']
    for k in KEYS:
      if k in col:
        lines.append(f'{k} [{units[col[k]]}] = {r[col[k]]}')
    for h, i in col.items():
      if 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
        try:
          if float(r[i]) >= 0.3:
            lines.append(f'{h} = {r[i]}')
        except ValueError:
          pass
    with open(os.path.join(ROOT, 'profiles', f'r2_ncu_{name}_summary.txt'), 'w') as fh:
      fh.write('\n'.join(lines) + '\n')
  with open(os.path.join(ROOT, 'profiles', 'r2_traffic.json'), 'w') as fh:
    json.dump(out, fh, indent=1)
  print(json.dumps(out, indent=1))


if __name__ == '__main__':
  sys.exit(main())
