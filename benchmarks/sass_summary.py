#!/usr/bin/env python
"""profiles/sass_summary.txt: per-kernel counts of the SASS mnemonics that show
what each kernel is made of (run here, no GPU needed: `python
benchmarks/sass_summary.py` after `make -C weatherbench2_b200/csrc`).

  UBLKCP            cp.async.bulk (1-D TMA bulk copy global -> shared)
  SYNCS             mbarrier arrive / try_wait (TMA completion)
  FADD2/FMUL2/FFMA2 packed f32x2 arithmetic (Blackwell)
  FMNMX             float min / max (the sorting networks)
  LDG / LDS / STS   global / shared traffic
  HMMA/UTC*MMA      tensor cores (expected: none -- no dense contraction here)
"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, 'weatherbench2_b200', 'csrc')
COLS = ['UBLKCP', 'SYNCS', 'FADD2', 'FMUL2', 'FFMA2', 'FMNMX', 'FFMA', 'FADD',
        'LDG', 'LDS', 'STS', 'BAR', 'HMMA', 'UTCMMA']


def demangle(names):
  try:
    out = subprocess.run(['cu++filt'] + names, capture_output=True, text=True,
                         check=True).stdout.split('\n')
    return [o.strip() for o in out[:len(names)]]
  except Exception:  # pylint: disable=broad-except
    return names


def main():
  rows = []
  for obj in sorted(glob.glob(os.path.join(CSRC, '*.o'))):
    txt = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True,
                         text=True, check=True).stdout
    cur, counts = None, None
    for line in txt.splitlines():
      m = re.match(r'\s*Function : (\S+)', line)
      if m:
        if cur:
          rows.append((os.path.basename(obj), cur, counts))
        cur, counts = m.group(1), collections.Counter()
        continue
      m = re.match(r'\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
      if m and cur:
        op = m.group(1)
        counts['total'] += 1
        for c in COLS:
          if op == c or (c == 'UTCMMA' and op.startswith('UTC') and 'MMA' in op):
            counts[c] += 1
    if cur:
      rows.append((os.path.basename(obj), cur, counts))
  names = demangle([r[1] for r in rows])
  lines = ['# SASS mnemonic counts per kernel (cuobjdump -sass of the sm_100a objects '
           'in weatherbench2_b200/csrc; static counts)',
           '# written by benchmarks/sass_summary.py',
           '%-16s %7s ' % ('object', 'total') + ' '.join('%6s' % c for c in COLS) +
           '  kernel']
  for (obj, _, c), name in zip(rows, names):
    name = re.sub(r'\((?:int|bool)\)', '', name).replace('wb2::', '')
    name = re.sub(r'\([^()]*\)\s*$', '', name)
    lines.append('%-16s %7d ' % (obj, c['total']) +
                 ' '.join('%6d' % c[k] for k in COLS) + '  ' + name[:110])
  out = os.path.join(ROOT, 'profiles', 'sass_summary.txt')
  with open(out, 'w') as fh:
    fh.write('\n'.join(lines) + '\n')
  print(f'{out}: {len(rows)} kernels')
  for l in lines[:3]:
    print(l)
  for l in lines:
    if any(k in l for k in ('det_tma_kernel<', 'spectrum_pfa_kernel<', 'regrid_tma_kernel<',
                            'ens_pair_kernel<50')):
      print(l[:200])


if __name__ == '__main__':
  main()
