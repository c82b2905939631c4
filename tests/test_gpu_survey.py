"""GPU parity tests of the survey-geometry monopole bispectrum (pyspectrum.py:13-282, 731-826) through
the CUDA path: FFT_survey_mono / _B0_survey / B0_survey / _Bk_periodic of pyspectrum_b200.pyspectrum against goldens made
by the UNMODIFIED reference (tests/golden/survey_{A,B}.npz) and against the CPU oracle on seeded catalogues.

Tolerances.  Index lists and normalised counts: exact / 1e-13.  delta_0(k): 3e-6 of max|delta| (float32 scatter order).
p0k: rtol 1e-5 on the pre-shot-noise value p + (1+alpha) I12/I22.  b123: |diff| <= 1e-5 * scale + 1e-6 * max(scale) with
scale = |b123 + shot-noise term| -- delta = delta_d - alpha * delta_r cancels the survey window, which amplifies the
float32 rounding of the two meshes: re-running the reference algorithm itself with the particles in a different order moves
b123 by up to 0.3 of this bound (measured on these catalogues with the oracle), so nothing tighter is meaningful.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-5
SURVEY_CFGS = {'A': [(3, 3, 4), (2, 3, 6)], 'B': [(2, 2, 7), (1, 1, 8)]}


@pytest.fixture(scope='module')
def mods():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from pyspectrum_b200 import pyspectrum as pySpec
    from oracle import pyspec_oracle as O
    return pySpec, O


def _g(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name)))


def _check_b0(bk, ref, alpha, I12, I13, I22, I23, I33):
    for key in ['i_k1', 'i_k2', 'i_k3']:
        assert np.array_equal(bk[key], ref[key]), key
    np.testing.assert_allclose(bk['counts'], ref['counts'], rtol=1e-13)
    sn_p = (1. + alpha) * I12 / I22
    for key in ['p0k1', 'p0k2', 'p0k3']:
        np.testing.assert_allclose(bk[key] + sn_p, ref[key] + sn_p, rtol=RTOL)
    sn_b = ((ref['p0k1'] + ref['p0k2'] + ref['p0k3']) * I23 + (1. - alpha ** 2) * I13) / I33
    scale = np.abs(ref['b123'] + sn_b)
    assert np.all(np.abs(bk['b123'] - ref['b123']) <= RTOL * scale + 1e-6 * scale.max())


@pytest.mark.parametrize('tag', ['A', 'B'])
def test_survey_matches_reference_goldens(mods, golden_dir, tag):
    pySpec, O = mods
    g = _g(golden_dir, 'survey_%s.npz' % tag)
    N, L, P0 = int(g['Ngrid']), float(g['Lbox']), float(g['P0_fkp'])
    w = g.get('w')
    radecz0 = g['radecz'].copy()
    w0 = None if w is None else w.copy()
    # ---- FFT_survey_mono: delta_0(k) and the normalisation sums, data and randoms
    dd = pySpec.FFT_survey_mono(g['radecz'], g['nbar'], w=w, P0_fkp=P0, Lbox=L, Ngrid=N)
    dr = pySpec.FFT_survey_mono(g['radecz_r'], g['nbar_r'], P0_fkp=P0, Lbox=L, Ngrid=N)
    for out, dk, sk in [(dd, 'delta_d', 'sums_d'), (dr, 'delta_r', 'sums_r')]:
        assert out[0].shape == g[dk].shape and out[0].dtype == np.complex64
        assert np.abs(out[0] - g[dk]).max() <= 3e-6 * np.abs(g[dk]).max()
        np.testing.assert_allclose(np.array(out[1:]), g[sk], rtol=1e-12)
    assert np.array_equal(g['radecz'], radecz0) and (w is None or np.array_equal(w, w0))     # inputs are not modified
    alpha = g['sums_d'][0] / g['sums_r'][0]
    I12, I13, I22, I23, I33 = alpha * g['sums_r'][1:]
    for (step, Ncut, Nmax) in SURVEY_CFGS[tag]:
        pre = 'b0_s%d_c%d_m%d_' % (step, Ncut, Nmax)
        ref = {k: g[pre + k] for k in ['i_k1', 'i_k2', 'i_k3', 'p0k1', 'p0k2', 'p0k3', 'b123', 'q123', 'counts']}
        # ---- the public entry point
        bk = pySpec.B0_survey(g['radecz'], g['nbar'], w=w, radecz_r=g['radecz_r'], nbar_r=g['nbar_r'], P0_fkp=P0, Lbox=L,
                              Ngrid=N, step=step, Ncut=Ncut, Nmax=Nmax)
        _check_b0(bk, ref, alpha, I12, I13, I22, I23, I33)
        assert bk['meta']['Ngrid'] == N and bk['meta']['N'] == g['radecz'].shape[1] and bk['meta']['kf'] == 2 * np.pi / L
        # ---- _B0_survey on the reference's own full field (reflect_delta of its half fields)
        deltak = O.reflect_delta(g['delta_d'], N) - alpha * O.reflect_delta(g['delta_r'], N)
        bk2 = pySpec._B0_survey(deltak, alpha, I12, I13, I22, I23, I33, Nmax=Nmax, Ncut=Ncut, step=step)
        _check_b0(bk2, ref, alpha, I12, I13, I22, I23, I33)


def test_survey_argument_errors(mods, golden_dir):
    pySpec, _ = mods
    g = _g(golden_dir, 'survey_A.npz')
    N, L = int(g['Ngrid']), float(g['Lbox'])
    with pytest.raises(ValueError):
        pySpec.B0_survey(g['radecz'], g['nbar'], radecz_r=g['radecz_r'], Lbox=L, Ngrid=N)          # nbar_r missing (py:82-83)
    with pytest.raises(ValueError):
        pySpec.B0_survey(g['radecz'], g['nbar'], Lbox=L, Ngrid=N)                                  # no randoms
    with pytest.raises(AssertionError):
        pySpec.FFT_survey_mono(g['radecz'], g['nbar'], Lbox=1000., Ngrid=N)                        # 'box not big enough!'


def _shell_amplification(O, half_d, half, N, step, Nmax):
    """Per shell j: rms|delta_d| / rms|delta_d - alpha delta_r| over the modes of the shell -- how much the subtraction of
    the survey window amplifies the relative float32 rounding of the data (and random) mesh."""
    irk = O.shell_index(N, step)[:N // 2 + 1]
    amp = np.ones(Nmax + 1)
    for j in range(1, Nmax + 1):
        m = irk == j
        if m.any():
            amp[j] = max(1., np.sqrt(np.sum(np.abs(half_d[m]) ** 2) / np.sum(np.abs(half[m]) ** 2)))
    return amp


@pytest.mark.parametrize('N,Nd,Nr', [(48, 20000, 100000), (64, 50000, 200000)])
def test_survey_matches_oracle_seeded(mods, N, Nd, Nr):
    """Larger seeded cone against the oracle, in three steps, because delta = delta_d - alpha delta_r cancels the survey
    window (here by a factor ~45 in the lowest shell) and with it amplifies the float32 summation-order noise of the two
    meshes: the ORACLE itself, fed the same catalogue in a different particle order, moves b123 by up to 5.5x the plain
    1e-5 bound on these catalogues (measured).  (1) the uncancelled fields delta_d, delta_r agree to 3e-6 of max|delta|;
    (2) the shell / triangle stage on the oracle's own delta(k) agrees to the plain bound; (3) the whole of B0_survey agrees
    to the bound scaled by the largest window amplification of the triangle's three shells."""
    pySpec, O = mods
    rng = np.random.default_rng(N)

    def cone(n):
        return np.array([rng.uniform(100., 260., n), np.degrees(np.arcsin(rng.uniform(-0.1, 0.9, n))),
                         rng.uniform(0.15 ** 3, 0.6 ** 3, n) ** (1. / 3.)])
    data, rand = cone(Nd), cone(Nr)
    nz = lambda z: 2e-4 * (1. + z)
    w = rng.uniform(0.9, 1.2, Nd)
    L, P0, step, Ncut, Nmax = 3600., 2e4, 2, 3, 10
    kw = dict(P0_fkp=P0, Lbox=L, Ngrid=N, step=step, Ncut=Ncut, Nmax=Nmax)
    # (1) uncancelled fields
    od = O.FFT_survey_mono(data, nz(data[2]), w=w, P0_fkp=P0, Lbox=L, Ngrid=N)
    orr = O.FFT_survey_mono(rand, nz(rand[2]), P0_fkp=P0, Lbox=L, Ngrid=N)
    gd = pySpec.FFT_survey_mono(data, nz(data[2]), w=w, P0_fkp=P0, Lbox=L, Ngrid=N)
    gr = pySpec.FFT_survey_mono(rand, nz(rand[2]), P0_fkp=P0, Lbox=L, Ngrid=N)
    for got, ref in [(gd, od), (gr, orr)]:
        assert np.abs(got[0] - ref[0]).max() <= 3e-6 * np.abs(ref[0]).max()
        np.testing.assert_allclose(np.array(got[1:]), np.array(ref[1:]), rtol=1e-12)
    alpha = od[1] / orr[1]
    Is = [alpha * x for x in orr[2:]]
    # (2) shell + triangle stage on the same delta(k)
    counts = pySpec._counts_Bk123(Ngrid=N, Nmax=Nmax, Ncut=Ncut, step=step)
    deltak = O.reflect_delta(od[0], N) - alpha * O.reflect_delta(orr[0], N)
    ref2 = O._B0_survey(deltak, alpha, *Is, Nmax=Nmax, Ncut=Ncut, step=step, counts=counts)
    _check_b0(pySpec._B0_survey(deltak, alpha, *Is, Nmax=Nmax, Ncut=Ncut, step=step), ref2, alpha, *Is)
    # (3) end to end
    bk = pySpec.B0_survey(data, nz(data[2]), w=w, radecz_r=rand, nbar_r=nz(rand[2]), **kw)
    ref = O.B0_survey(data, nz(data[2]), w=w, radecz_r=rand, nbar_r=nz(rand[2]), counts=counts, **kw)
    half = np.asarray(od[0]) - alpha * np.asarray(orr[0])
    amp = _shell_amplification(O, np.asarray(od[0]), half, N, step, Nmax)
    tri = O.triangle_list(Nmax, Ncut, step)
    amp3 = np.maximum(np.maximum(amp[tri[:, 0]], amp[tri[:, 1]]), amp[tri[:, 2]])
    assert amp3.max() > 5.                                            # the catalogue does exercise the cancellation
    for key in ['i_k1', 'i_k2', 'i_k3']:
        assert np.array_equal(bk[key], ref[key])
    sn_p = (1. + alpha) * Is[0] / Is[2]
    assert np.all(np.abs(bk['p0k1'] - ref['p0k1']) <= RTOL * amp[tri[:, 0]] * np.abs(ref['p0k1'] + sn_p))
    sn_b = ((ref['p0k1'] + ref['p0k2'] + ref['p0k3']) * Is[3] + (1. - alpha ** 2) * Is[1]) / Is[4]
    scale = np.abs(ref['b123'] + sn_b)
    assert np.all(np.abs(bk['b123'] - ref['b123']) <= amp3 * (RTOL * scale + 1e-6 * scale.max()))


def test_bk_periodic_from_full_field(mods, golden_dir):
    """_Bk_periodic (pyspectrum.py:359-457) on a host full field: the reference golden's raw (unit-less) bispectrum, and a
    non-Hermitian input, for which the reference keeps Re(FFT) = the transform of the Hermitian part."""
    pySpec, O = mods
    g = _g(golden_dir, 'small_A.npz')
    N, L = int(g['Ngrid']), float(g['Lbox'])
    full = O.reflect_delta(g['delta_half'], N)
    step, Ncut, Nmax = 2, 3, 6
    raw = pySpec._counts_Bk123(Ngrid=N, Nmax=Nmax, Ncut=Ncut, step=step)
    ref = O._Bk_periodic(full, Nmax=Nmax, Ncut=Ncut, step=step, counts=raw)
    bk = pySpec._Bk_periodic(full, Nmax=Nmax, Ncut=Ncut, step=step)
    assert np.array_equal(bk['i_k1'], ref['i_k1']) and np.array_equal(bk['i_k3'], ref['i_k3'])
    np.testing.assert_allclose(bk['p0k1'], ref['p0k1'], rtol=RTOL)
    assert np.all(np.abs(bk['b123'] - ref['b123']) <= RTOL * np.abs(ref['b123']) + 1e-7 * np.abs(ref['b123']).max())
    rng = np.random.default_rng(3)
    noise = (rng.normal(size=full.shape) + 1j * rng.normal(size=full.shape)).astype(np.complex64) * np.abs(full).mean()
    ref = O._Bk_periodic(full + noise, Nmax=Nmax, Ncut=Ncut, step=step, counts=raw)
    bk = pySpec._Bk_periodic(full + noise, Nmax=Nmax, Ncut=Ncut, step=step)
    np.testing.assert_allclose(bk['p0k1'], ref['p0k1'], rtol=RTOL)
    assert np.all(np.abs(bk['b123'] - ref['b123']) <= RTOL * np.abs(ref['b123']) + 1e-6 * np.abs(ref['b123']).max())


def test_pk_periodic_rsd_from_half_field(mods, golden_dir):
    """_Pk_periodic_rsd (pyspectrum.py:541-641, code='fortran') on the reference golden's own half field."""
    pySpec, O = mods
    g = _g(golden_dir, 'small_B.npz')
    N, L = int(g['Ngrid']), float(g['Lbox'])
    for rsd in (0, 1, 2):
        ks, p0k, p2k, p4k, nk, k_kmu, mu_kmu, p_kmu, n_kmu = pySpec._Pk_periodic_rsd(g['delta_half'], Lbox=L, rsd=rsd, Nmubin=5)
        pre = 'rsd%d_mu5_' % rsd
        sn = g[pre + 'p_sn'][0]
        assert np.array_equal(nk, g[pre + 'counts']) and np.array_equal(n_kmu, g[pre + 'counts_kmu'])
        np.testing.assert_allclose(ks, g[pre + 'k'], rtol=1e-6)
        np.testing.assert_allclose(p0k, g[pre + 'p0k'] + sn, rtol=RTOL)            # the golden is shot-noise corrected
        scale = np.abs(g[pre + 'p0k'] + sn)
        assert np.all(np.abs(p2k - g[pre + 'p2k']) <= 5 * RTOL * scale)
        assert np.all(np.abs(p4k - g[pre + 'p4k']) <= 9 * RTOL * scale)
        m = n_kmu > 0
        np.testing.assert_allclose(p_kmu[m], (g[pre + 'p_kmu'] + sn)[m], rtol=RTOL)
    with pytest.raises(ValueError):                          # code='python' takes the FULL field (py:546)
        pySpec._Pk_periodic_rsd(g['delta_half'], Lbox=L, code='python')
    with pytest.raises(ValueError):
        pySpec._Pk_periodic_rsd(g['delta_half'][:, :, :-1], Lbox=L)


def test_counts_f77_entry_point(mods, tmp_path, monkeypatch):
    """_counts_Bk123_f77 (pyspectrum.py:1033-1057): estimator.bk_counts layout coun(i<=j<=l), cached as a Fortran record."""
    pySpec, O = mods
    monkeypatch.setattr(pySpec, '_DAT_DIR', str(tmp_path))
    N, nmax = 24, 3
    c = pySpec._counts_Bk123_f77(Ngrid=N, Nmax=nmax, Ncut=3, step=3)
    cb = O.counts_bruteforce(N, nmax, 3, 3)
    for (i, j, l) in O.triangle_list(nmax, 3, 3):
        assert c[l - 1, j - 1, i - 1] == cb[i - 1, j - 1, l - 1] * N ** 3
    assert os.path.isfile(os.path.join(str(tmp_path), 'counts.Ngrid24.Nmax3.Ncut3.step3.fort77'))
    assert np.array_equal(pySpec._counts_Bk123_f77(Ngrid=N, Nmax=nmax, Ncut=3, step=3), c)      # cache hit: same array


@pytest.mark.parametrize('Np,weighted', [(1, False), (777, True), (300000, True)])
def test_survey_prepare_kernel(mods, Np, weighted):
    """psb_survey_prepare (one pass on the device) against util.radecz_to_cartesian + the numpy lines of py:789-806."""
    pySpec, O = mods
    from pyspectrum_b200 import util as UT
    rng = np.random.default_rng(Np)
    radecz = np.array([rng.uniform(0., 360., Np), rng.uniform(-90., 90., Np), rng.uniform(0., 1.5, Np)])
    nb = rng.uniform(1e-5, 5e-4, Np)
    w = rng.uniform(0.5, 2., Np) if weighted else None
    cosmo = UT.FlatLambdaCDM(H0=70., Om0=0.3)
    xyz, wf, out = pySpec.PeriodicPipeline.get(24).survey_prepare(radecz, nb, w, 1e4, cosmo)
    ref = O.radecz_to_cartesian(radecz, O.FlatLambdaCDM(70., 0.3))
    xyz = xyz.cpu().numpy()
    assert xyz.dtype == np.float32 and xyz.shape == (3, Np)
    # float32 cast of a float64 value that agrees to ~1e-12: identical or one float32 ulp apart
    assert np.all(np.abs(xyz - ref.astype(np.float32)) <= np.spacing(np.abs(ref).astype(np.float32)) + 1e-8)
    w0 = np.ones(Np) if w is None else w
    wref = w0 * (1. / (1. + nb * 1e4))
    assert np.array_equal(wf.cpu().numpy(), wref.astype(np.float32))
    sums = [w0.sum(), (wref ** 2).sum(), (wref ** 3).sum(), (nb * wref ** 2).sum(), (nb * wref ** 3).sum(), (nb ** 2 * wref ** 3).sum()]
    np.testing.assert_allclose(out[:6], sums, rtol=1e-12)
    np.testing.assert_allclose(out[6:9], ref.min(axis=1), rtol=1e-10, atol=1e-8)
    np.testing.assert_allclose(out[9:12], ref.max(axis=1), rtol=1e-10, atol=1e-8)
    with pytest.raises(ValueError):
        pySpec.PeriodicPipeline.get(24).survey_prepare(-radecz, nb, w, 1e4, cosmo)               # negative redshift


@pytest.mark.parametrize('LOS', ['x', 'y', 'z'])
def test_apply_rsd_on_device_equals_numpy_bit_for_bit(mods, LOS):
    """util.applyRSD (util.py:54-75) with torch CUDA tensors runs psb_apply_rsd: float64 products, sums and np.remainder reproduced
    exactly (the numpy path itself is pinned to the reference's output in tests/golden/util.npz on CPU)."""
    import torch
    from pyspectrum_b200 import util as UT
    rng = np.random.default_rng(ord(LOS))
    L, n = 2600., 200001
    xyz = rng.uniform(0, L, (3, n))
    v = rng.normal(0, 600., (3, n))
    v[:, :4] = [[-5e5, 5e5, 0., -1e-300]] * 3                       # several box lengths either way, zero, denormal shift
    xyz[:, 4] = 0.
    v[:, 4] = 0.                                                     # lands exactly on L -> remainder 0
    want = UT.applyRSD(xyz, v, 0.5, h=0.7, omega0_m=0.3, LOS=LOS, Lbox=L)
    got = UT.applyRSD(torch.from_numpy(xyz).cuda(), torch.from_numpy(v).cuda(), 0.5, h=0.7, omega0_m=0.3, LOS=LOS, Lbox=L)
    assert got.is_cuda and got.dtype == torch.float64
    g = got.cpu().numpy()
    assert np.array_equal(g.view(np.int64), want.view(np.int64))
    assert g.min() >= 0. and g['xyz'.index(LOS)].max() < L
