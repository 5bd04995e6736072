"""NumPy model of the index logic of csrc/spectrum_pfa.cu (CPU): Good's input
map, the CRT output map, the in-place [RB][RA][RC] layout, the mirrored-pair
task table of the last stage and the real-input split -- every bin of the rfft
power spectrum must come out exactly once.  The kernel's host code builds the
same tables (build_tables / Plan in spectrum_pfa.cu)."""
import numpy as np
import pytest


def _idem(n2, r):
  m = n2 // r
  x = m
  while x % r != 1 % r:
    x += m
  return x % n2


@pytest.mark.parametrize('ra,rb,rc', [(9, 16, 5), (9, 8, 5), (3, 8, 5)])
def test_prime_factor_fft_and_mirrored_split(ra, rb, rc):
  n2 = ra * rb * rc
  n = 2 * n2
  sa, sb, sc = n2 // ra, n2 // rb, n2 // rc
  ea, eb, ec = _idem(n2, ra), _idem(n2, rb), _idem(n2, rc)
  rs = np.random.RandomState(n2)
  x = rs.standard_normal(n)
  z = x[0::2] + 1j * x[1::2]
  # stage input: gather by Good's map into the [RB][RA][RC] work array
  w = np.zeros((rb, ra, rc), complex)
  for nb in range(rb):
    for na in range(ra):
      for nc in range(rc):
        w[nb, na, nc] = z[(sa * na + sb * nb + sc * nc) % n2]
  # three plain DFTs, in place, no twiddles
  w = np.fft.fft(w, axis=1)
  w = np.fft.fft(w, axis=0)
  w = np.fft.fft(w, axis=2)
  zf = np.fft.fft(z)
  for kb in range(rb):
    for ka in range(ra):
      for kc in range(rc):
        k = (ka * ea + kb * eb + kc * ec) % n2
        assert abs(w[kb, ka, kc] - zf[k]) < 1e-9 * n2
  # mirrored-pair tasks of the last stage (the order build_tables() uses)
  tasks, seen = [], set()
  for ia in range(1, ra + 1):
    ka = ia % ra
    for kb in range(rb):
      if (ka, kb) in seen:
        continue
      q = ((ra - ka) % ra, (rb - kb) % rb)
      seen.add((ka, kb))
      seen.add(q)
      tasks.append((ka, kb, q[0], q[1], q == (ka, kb)))
  nself = (2 if ra % 2 == 0 else 1) * (2 if rb % 2 == 0 else 1)
  assert len(tasks) == (ra * rb - nself) // 2 + nself
  power = np.zeros(n2 + 1)
  count = np.zeros(n2 + 1, int)
  for ka, kb, qa, qb, self_ in tasks:
    k0 = (ka * ea + kb * eb) % n2
    for kc in range(rc):
      p = (k0 + kc * ec) % n2
      zp, zq = w[kb, ka, kc], w[qb, qa, (rc - kc) % rc]
      e, d = zp + np.conj(zq), zp - np.conj(zq)
      wn = np.exp(-2j * np.pi * p / n)
      xa, xb = e + wn * (-1j * d), e - wn * (-1j * d)
      va = (not self_) or kc <= (rc - kc) % rc
      vb = va and 2 * p != n2
      if va:
        power[p] = abs(xa) ** 2 / 4
        count[p] += 1
      if vb:
        power[n2 - p] = abs(xb) ** 2 / 4
        count[n2 - p] += 1
  assert (count == 1).all()
  np.testing.assert_allclose(power, np.abs(np.fft.rfft(x)) ** 2, rtol=1e-9)
