#!/usr/bin/env python
"""Top source lines of a kernel by executed warp instructions / stall samples
from an .ncu-rep captured with --import-source on (needs -lineinfo).
Usage: python benchmarks/ncu_hot_lines.py file.ncu-rep [N]"""
import csv
import subprocess
import sys


def main(path, top=25):
  out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv',
                        '--print-source', 'cuda,sass'],
                       capture_output=True, text=True).stdout
  agg = {}
  hdr = None
  for r in csv.reader(out.splitlines()):
    if r and r[0] == 'Line No':
      hdr = r
      iexe = hdr.index('Instructions Executed')
      ist = hdr.index('Warp Stall Sampling (All Samples)')
      continue
    if hdr is None or len(r) != len(hdr):
      continue
    key = (r[0], r[1].strip())
    a = agg.setdefault(key, [0.0, 0.0])
    try:
      a[0] += float(r[iexe])
      a[1] += float(r[ist])
    except ValueError:
      pass
  tot = sum(v[0] for v in agg.values()) or 1.0
  tots = sum(v[1] for v in agg.values()) or 1.0
  print(f'total warp instructions {tot:.4g}, stall samples {tots:.4g}')
  for (line, src), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f'{100 * v[0] / tot:5.1f}% inst {100 * v[1] / tots:5.1f}% stall | '
          f'{line:>4s}: {src[:100]}')


if __name__ == '__main__':
  main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
