"""CPU: pin oracle/ (the parity checker) against the reference's own outputs.

Goldens in tests/golden/small_*.npz were produced by the UNMODIFIED reference Python
layer (tests/golden/make_golden.py); counts_N360_*.npz are the reference's shipped
triangle-count caches (pyspectrum/dat/counts.Ngrid360.*.pyfftw) as exact integers.
"""
import os

import numpy as np
import pytest

from oracle import pyspec_oracle as O

CASES = ['A', 'B', 'C']


def _load(golden_dir, tag):
    return dict(np.load(os.path.join(golden_dir, 'small_%s.npz' % tag)))


@pytest.mark.parametrize('tag', CASES)
def test_delta_half_matches_reference(golden_dir, tag):
    g = _load(golden_dir, tag)
    d = O.FFT_periodic(g['xyz'], w=g.get('w'), Lbox=float(g['Lbox']), Ngrid=int(g['Ngrid']))
    assert np.array_equal(np.ascontiguousarray(d), g['delta_half'])      # same code path: bit-identical


@pytest.mark.parametrize('tag', CASES)
def test_pk_periodic_matches_reference(golden_dir, tag):
    g = _load(golden_dir, tag)
    pk = O.Pk_periodic(g['xyz'], w=g.get('w'), Lbox=float(g['Lbox']), Ngrid=int(g['Ngrid']))
    assert np.array_equal(pk['counts'], g['pk_counts'])
    np.testing.assert_allclose(pk['k'], g['pk_k'], rtol=1e-14)
    # float32 pairwise sums in a different gather order: 1e-6 of the raw power
    raw, raw_g = pk['p0k'] + pk['p0k_sn'], g['pk_p0k'] + g['pk_p0k_sn']
    np.testing.assert_allclose(raw, raw_g, rtol=2e-6)


@pytest.mark.parametrize('tag', CASES)
@pytest.mark.parametrize('rsd', [0, 1, 2])
@pytest.mark.parametrize('nmu', [5, 10])
def test_pk_rsd_matches_reference(golden_dir, tag, rsd, nmu):
    g = _load(golden_dir, tag)
    pr = O.Pk_periodic_rsd(g['xyz'], w=g.get('w'), Lbox=float(g['Lbox']), Ngrid=int(g['Ngrid']), rsd=rsd, Nmubin=nmu)
    pre = 'rsd%d_mu%d_' % (rsd, nmu)
    for key in ['k', 'p0k', 'p2k', 'p4k', 'p_sn', 'counts', 'k_kmu', 'mu_kmu', 'p_kmu', 'counts_kmu']:
        assert np.array_equal(np.asarray(pr[key]), g[pre + key]), key


@pytest.mark.parametrize('tag', CASES)
def test_bk_matches_reference(golden_dir, tag):
    g = _load(golden_dir, tag)
    N, L = int(g['Ngrid']), float(g['Lbox'])
    for (step, Ncut, Nmax) in [(3, 3, 4), (2, 3, 6), (1, 1, 8)]:
        pre = 'bk_s%d_c%d_m%d_' % (step, Ncut, Nmax)
        if pre + 'b123' not in g:
            continue
        raw = g[pre + 'rawcounts'].astype(np.float64) * N ** 3
        bk = O.Bk_periodic(g['xyz'], w=g.get('w'), Lbox=L, Ngrid=N, step=step, Ncut=Ncut, Nmax=Nmax, counts=raw)
        for key in ['i_k1', 'i_k2', 'i_k3']:
            assert np.array_equal(bk[key], g[pre + key])
        np.testing.assert_allclose(bk['counts'], g[pre + 'counts'], rtol=1e-12)
        np.testing.assert_allclose(bk['p0k1'] + bk['p0k_sn'], g[pre + 'p0k1'] + bk['p0k_sn'], rtol=1e-9)
        scale = np.abs(g[pre + 'b123'] + g[pre + 'b123_sn'])
        assert np.all(np.abs(bk['b123'] - g[pre + 'b123']) <= 1e-9 * scale + 1e-12 * scale.max())
        # the oracle's own counts (double-precision FFT of delta==1) reproduce the reference's integers
        mine = O._counts_Bk123(Ngrid=N, Nmax=Nmax, Ncut=Ncut, step=step)
        assert np.array_equal(np.rint(mine / N ** 3).astype(np.int64), g[pre + 'rawcounts'])


def test_counts_bruteforce_definition():
    """counts = N^3 * #{closed triangles mod N}; low shells are grid independent."""
    cb = O.counts_bruteforce(24, 3, 3, 3)
    assert [cb[0, 0, 0], cb[1, 0, 0], cb[1, 1, 0], cb[1, 1, 1]] == [54066, 71166, 225724, 470232]
    cf = O._counts_Bk123(24, 3, 3, 3)
    assert np.array_equal(np.rint(cf / 24 ** 3).astype(np.int64), cb)


@pytest.mark.parametrize('nmax', [10, 40, 50])
def test_shipped_counts_low_shells(golden_dir, nmax):
    """The reference's shipped N=360 caches agree with the oracle where 3*kmax < N (no wrap):
    checked on the shells a 48^3 grid can hold (k <= 7 shells of step 3 -> |k|<=22.5 < 24)."""
    g = np.load(os.path.join(golden_dir, 'counts_N360_Nmax%d_Ncut3_step3.npz' % nmax))
    ship = {tuple(t): n for t, n in zip(g['ijl'].tolist(), g['n'].tolist())}
    mine = np.rint(O._counts_Bk123(72, 7, 3, 3) / 72 ** 3).astype(np.int64)     # 3*22.5 < 72
    for (i, j, l) in O.triangle_list(7, 3, 3):
        assert mine[i - 1, j - 1, l - 1] == ship[(i, j, l)], (i, j, l)
