"""Coarse transform grids for the shell / triangle stage (pyspectrum.coarse_levels): sum_x I_i I_j I_l / Ngrid^3 and the number
of closed mode triples are the same on every grid that holds the triangle without wrap-around, so triangles are summed on the
coarsest compiled grid that holds them.  Checked against (a) the oracle, which works on the one grid like the reference,
(b) the single-grid GPU path, (c) the reference's shipped counts (tests/test_gpu_parity.py runs with the default levels)."""
import numpy as np
import pytest

RTOL = 1e-5


@pytest.fixture(scope='module')
def mods():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from pyspectrum_b200 import pyspectrum as pySpec
    from oracle import pyspec_oracle as O
    return pySpec, O


def _cat(seed, Np, L):
    rng = np.random.default_rng(seed)
    npar = max(Np // 40, 1)
    par = rng.uniform(0, L, (3, npar))
    kids = par[:, rng.integers(0, npar, Np // 2)] + rng.normal(0, 0.02 * L, (3, Np // 2))
    return np.ascontiguousarray(np.concatenate([kids, rng.uniform(0, L, (3, Np - Np // 2))], axis=1) % L)


def test_level_split_is_a_partition():
    from pyspectrum_b200 import pyspectrum as P
    for N, step, Nmax, sizes in [(360, 3, 40, [256]), (512, 2, 80, [256, 320, 400]), (1024, 3, 40, [256, 400]), (96, 1, 24, [48, 64])]:
        tri = P.triangle_list(Nmax, 3 if step > 1 else 1, step)
        lev = P.coarse_levels(N, step, tri, sizes)
        allidx = np.sort(np.concatenate([idx for _, idx, _ in lev]))
        assert np.array_equal(allidx, np.arange(len(tri)))
        for Nc, idx, smax in lev:
            R = P.shell_reach(tri[idx], step)
            assert smax == tri[idx].max()
            if Nc < N:
                assert np.all(R.sum(axis=1) < Nc) and np.all(2 * R.max(axis=1) < Nc)


@pytest.mark.gpu
@pytest.mark.parametrize('levels', ['48,64', '64'])
def test_small_grid_levels_match_oracle(mods, monkeypatch, levels):
    pySpec, O = mods
    N, L, Np, step, Ncut, Nmax = 96, 400., 150000, 1, 1, 24
    xyz = _cat(96, Np, L)
    monkeypatch.setenv('PSB_BK_LEVELS', levels)
    pipe = pySpec.PeriodicPipeline.get(N)
    tri, lev = pipe.bk_levels(step, Ncut, Nmax)
    assert len(lev) == len(levels.split(',')) + 1 and lev[-1][0].N == N      # the largest triangles (R_i+R_j+R_l >= 64) stay on the fine grid
    pipe._counts.pop((Nmax, Ncut, step), None)
    monkeypatch.setattr(pySpec, '_DAT_DIR', '/tmp/psb_levels_test_%s' % levels.replace(',', '_'))
    bk = pySpec.Bk_periodic(xyz, Lbox=L, Ngrid=N, step=step, Ncut=Ncut, Nmax=Nmax)
    rb = O.Bk_periodic(xyz, Lbox=L, Ngrid=N, step=step, Ncut=Ncut, Nmax=Nmax, workers=8, pool_threads=8)
    for key in ('i_k1', 'i_k2', 'i_k3'):
        assert np.array_equal(bk[key], rb[key])
    np.testing.assert_allclose(bk['counts'], rb['counts'], rtol=1e-12)       # exact integers / fac on both sides
    np.testing.assert_allclose(bk['p0k1'] + bk['p0k_sn'], rb['p0k1'] + rb['p0k_sn'], rtol=RTOL)
    scale = np.abs(rb['b123'] + rb['b123_sn'])
    assert np.all(np.abs(bk['b123'] - rb['b123']) <= RTOL * scale + 1e-7 * scale.max())
    pipe._counts.pop((Nmax, Ncut, step), None)


@pytest.mark.gpu
def test_compiled_coarse_grids_match_single_grid(mods, monkeypatch):
    """Ngrid=512, step=3, Nmax=40 (every triangle is alias-free below 400): levels 256 / 320 / 400 against the one-grid path on
    the same delta(k); triangle sums, shell powers and exact counts."""
    pySpec, _ = mods
    N, L, step, Ncut, Nmax = 512, 2600., 3, 3, 40
    xyz = _cat(5, 2 * 10 ** 6, L)
    pipe = pySpec.PeriodicPipeline.get(N)
    half, _ = pipe.fft_periodic(xyz, None, L)
    out = {}
    for spec in ('off', '256,320,400', '400'):
        monkeypatch.setenv('PSB_BK_LEVELS', spec)
        tri, lev = pipe.bk_levels(step, Ncut, Nmax)
        assert [pc.N for pc, _, _, _ in lev] == {'off': [512], '256,320,400': [256, 320, 400], '400': [400]}[spec]
        out[spec] = pipe.bispectrum_sums(half, step, Ncut, Nmax)
        out[spec + 'c'] = pipe.compute_counts(Nmax, Ncut, step)
    s0, q0 = out['off']
    for spec in ('256,320,400', '400'):
        s1, q1 = out[spec]
        np.testing.assert_allclose(q1, q0, rtol=RTOL)
        # the float32 transforms on two grids round differently; bound on |S| plus the tiny absolute floor of sums that cancel
        assert np.all(np.abs(s1 - s0) <= RTOL * np.abs(s0) + 1e-7 * np.abs(s0).max()), np.abs((s1 - s0) / s0).max()
        assert np.array_equal(out[spec + 'c'], out['offc'])
