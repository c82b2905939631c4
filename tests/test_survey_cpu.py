"""CPU: the survey-geometry restatement in oracle/ and the host utilities (pyspectrum_b200/util.py) against goldens made
by the UNMODIFIED reference (tests/golden/make_golden.py survey: survey_{A,B}.npz, util.npz) and against known answers
for the cosmology helper that stands in for astropy (not installed; un-pinned third-party code)."""
import os

import numpy as np
import pytest

from oracle import pyspec_oracle as O
from pyspectrum_b200 import util as UT

SURVEY_CFGS = {'A': [(3, 3, 4), (2, 3, 6)], 'B': [(2, 2, 7), (1, 1, 8)]}


def _load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name)))


# ------------------------------------------------------------------------------ cosmology stand-in
def test_comoving_distance_known_answers():
    c = 299792.458
    # Einstein-de Sitter (Om0 = 1): D = 2c/H0 (1 - 1/sqrt(1+z)), exact
    z = np.array([0., 0.1, 0.5, 1., 3., 9.])
    eds = 2. * c / 70. * (1. - 1. / np.sqrt(1. + z))
    np.testing.assert_allclose(UT.FlatLambdaCDM(H0=70., Om0=1.).comoving_distance(z), eds, rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(O.FlatLambdaCDM(H0=70., Om0=1.).comoving_distance(z), eds, rtol=1e-11, atol=1e-9)
    # de Sitter (Om0 = 0): D = c z / H0
    np.testing.assert_allclose(UT.FlatLambdaCDM(H0=50., Om0=0.).comoving_distance(z), c * z / 50., rtol=1e-13)
    # product (Gauss-Legendre) vs oracle (adaptive quad): two independent quadratures of the same integrand
    zz = np.linspace(0.01, 2.5, 40)
    a = UT.FlatLambdaCDM(H0=67.6, Om0=0.31).comoving_distance(zz)
    b = O.FlatLambdaCDM(H0=67.6, Om0=0.31).comoving_distance(zz)
    np.testing.assert_allclose(a, b, rtol=1e-12)
    assert UT.FlatLambdaCDM(H0=67.6, Om0=0.31).h == pytest.approx(0.676)
    # catalogue-sized input: cubic Hermite table (exact slopes) vs the direct quadrature, absolute error in Mpc
    cos = UT.FlatLambdaCDM(H0=67.6, Om0=0.31)
    zbig = np.random.default_rng(1).uniform(0., 1.2, 20000)
    assert np.abs(cos.comoving_distance(zbig) - (c / 67.6) * cos._integral(zbig)).max() < 1e-9
    assert UT.FlatLambdaCDM(H0=67.6, Om0=0.31).efunc(0.) == pytest.approx(1.)


@pytest.mark.parametrize('tag', ['A', 'B'])
def test_radecz_to_cartesian_matches_reference(golden_dir, tag):
    g = _load(golden_dir, 'survey_%s.npz' % tag)
    before = g['radecz'].copy()
    xyz = UT.radecz_to_cartesian(g['radecz'])
    assert np.array_equal(g['radecz'], before)                       # the caller's array is left alone (ut:41-42 is not)
    np.testing.assert_allclose(xyz, g['xyz'], rtol=1e-11, atol=1e-8)
    np.testing.assert_allclose(O.radecz_to_cartesian(g['radecz'], O.FlatLambdaCDM(67.6, 0.31)), g['xyz'], rtol=1e-14)

    class Quantity(object):                                          # astropy returns a Quantity: .value is unwrapped
        def __init__(self, v):
            self.value = v

    class AstropyLike(object):
        h = 0.676

        def comoving_distance(self, z):
            return Quantity(UT.FlatLambdaCDM(67.6, 0.31).comoving_distance(z))
    np.testing.assert_allclose(UT.radecz_to_cartesian(g['radecz'], cosmo=AstropyLike()), xyz, rtol=1e-15)
    with pytest.raises(AssertionError):
        UT.radecz_to_cartesian(np.zeros((2, 5)))


# ------------------------------------------------------------------------------ util.py
def test_ijl_order_matches_reference(golden_dir):
    g = _load(golden_dir, 'util.npz')
    t = g['ijl_in']
    got = UT.ijl_order(t[:, 0], t[:, 1], t[:, 2], typ='GM')
    assert got.shape == g['ijl_out'].shape and np.array_equal(got, g['ijl_out'])
    assert np.array_equal(O.ijl_order(t[:, 0], t[:, 1], t[:, 2]), g['ijl_out'])
    s = t[got.ravel()]                                               # l slowest, then j, then i
    key = s[:, 2] * 10000 + s[:, 1] * 100 + s[:, 0]
    assert np.all(np.diff(key) > 0)
    with pytest.raises(NotImplementedError):
        UT.ijl_order(t[:, 0], t[:, 1], t[:, 2], typ='other')
    # duplicates: one row per distinct triple holding all its positions, like the reference's boolean masks
    d = UT.ijl_order(np.array([3, 3, 2, 2]), np.array([2, 2, 2, 2]), np.array([1, 1, 1, 1]))
    assert np.array_equal(d, np.array([[2, 3], [0, 1]]))


def test_apply_rsd_matches_reference(golden_dir):
    g = _load(golden_dir, 'util.npz')
    for los in 'xyz':
        out = UT.applyRSD(g['rsd_xyz'], g['rsd_vxyz'], 0.5, h=0.7, omega0_m=0.3, LOS=los, Lbox=500.)
        assert np.array_equal(out, g['rsd_out_' + los])
        assert np.array_equal(O.applyRSD(g['rsd_xyz'], g['rsd_vxyz'], 0.5, LOS=los, Lbox=500.), g['rsd_out_' + los])
        assert out.min() >= 0. and out.max() < 500.
    with pytest.raises(ValueError):
        UT.applyRSD(g['rsd_xyz'], g['rsd_vxyz'], 0.5, Lbox=500.)
    with pytest.raises(ValueError):
        UT.applyRSD(g['rsd_xyz'], g['rsd_vxyz'], 0.5, LOS='z')


def test_read_fortfft_roundtrip(tmp_path):
    N, h = 8, 4
    rng = np.random.default_rng(5)
    half = (rng.normal(size=(h + 1, N, N)) + 1j * rng.normal(size=(h + 1, N, N))).astype(np.complex64)
    f = tmp_path / 'fft.dat'
    b1 = np.array([N], '<i4').tobytes()
    b2 = np.asfortranarray(half).tobytes(order='F')
    with open(f, 'wb') as fh:
        for b in (b1, b2):
            m = np.array([len(b)], '<i4').tobytes()
            fh.write(m + b + m)
    full = UT.read_fortFFT(str(f))
    assert np.array_equal(full, O.reflect_delta(half, N))


# ------------------------------------------------------------------------------ survey oracle vs the reference
@pytest.mark.parametrize('tag', ['A', 'B'])
def test_oracle_fft_survey_mono_matches_reference(golden_dir, tag):
    g = _load(golden_dir, 'survey_%s.npz' % tag)
    N, L, P0 = int(g['Ngrid']), float(g['Lbox']), float(g['P0_fkp'])
    w = g.get('w')
    w_before = None if w is None else w.copy()
    out = O.FFT_survey_mono(g['radecz'], g['nbar'], w=w, P0_fkp=P0, Lbox=L, Ngrid=N)
    assert np.array_equal(np.ascontiguousarray(out[0]), g['delta_d'])         # same code path: bit-identical
    np.testing.assert_allclose(np.array(out[1:]), g['sums_d'], rtol=1e-14)
    if w is not None:
        assert np.array_equal(w, w_before)
    out = O.FFT_survey_mono(g['radecz_r'], g['nbar_r'], P0_fkp=P0, Lbox=L, Ngrid=N)
    assert np.array_equal(np.ascontiguousarray(out[0]), g['delta_r'])
    np.testing.assert_allclose(np.array(out[1:]), g['sums_r'], rtol=1e-14)
    with pytest.raises(AssertionError):                                        # 'box not big enough!' (py:785)
        O.FFT_survey_mono(g['radecz'], g['nbar'], P0_fkp=P0, Lbox=1000., Ngrid=N)


@pytest.mark.parametrize('tag', ['A', 'B'])
def test_oracle_b0_survey_matches_reference(golden_dir, tag):
    g = _load(golden_dir, 'survey_%s.npz' % tag)
    N, L, P0 = int(g['Ngrid']), float(g['Lbox']), float(g['P0_fkp'])
    for (step, Ncut, Nmax) in SURVEY_CFGS[tag]:
        pre = 'b0_s%d_c%d_m%d_' % (step, Ncut, Nmax)
        bk = O.B0_survey(g['radecz'], g['nbar'], w=g.get('w'), radecz_r=g['radecz_r'], nbar_r=g['nbar_r'], P0_fkp=P0,
                         Lbox=L, Ngrid=N, step=step, Ncut=Ncut, Nmax=Nmax)
        for key in ['i_k1', 'i_k2', 'i_k3']:
            assert np.array_equal(bk[key], g[pre + key])
        np.testing.assert_allclose(bk['counts'], g[pre + 'counts'], rtol=1e-12)
        for key in ['p0k1', 'p0k2', 'p0k3']:
            np.testing.assert_allclose(bk[key], g[pre + key], rtol=1e-9, atol=1e-9 * np.abs(g[pre + key]).max())
        np.testing.assert_allclose(bk['b123'], g[pre + 'b123'], rtol=1e-8, atol=1e-9 * np.abs(g[pre + 'b123']).max())
        assert bk['meta']['Ngrid'] == N and bk['meta']['N'] == g['radecz'].shape[1]


# ------------------------------------------------------------------------------ host logic of the product's survey API
class _StubPipe(object):
    """Stands in for PeriodicPipeline on a machine without a GPU: the device stages (K1-K3, K5, K6) are answered by the
    oracle so that the HOST code of pyspectrum_b200.pyspectrum's survey functions (layout conversions, Hermitian part,
    data - alpha * randoms, normalisation algebra, output dictionaries) runs on CPU.  The kernels themselves are covered by
    tests/test_gpu_survey.py."""

    def __init__(self, N):
        import torch
        from pyspectrum_b200.pyspectrum import PeriodicPipeline
        self.N, self.h, self.dev = N, N // 2, torch.device('cpu')
        self._real = PeriodicPipeline

    def half_from_full(self, delta):
        return self._real.half_from_full(self, delta)

    def _to_half_tensor(self, half_f):                       # (h+1,N,N) [kx,ky,kz] -> float32 [kz][ky][kx][2]
        import torch
        arr = np.ascontiguousarray(np.asarray(half_f).transpose(2, 1, 0))
        return torch.from_numpy(arr.view(np.float32).reshape(self.N, self.N, self.h + 1, 2))

    def survey_prepare(self, radecz, nb, w, P0_fkp, cosmo):
        """numpy emulation of k_survey_prepare (csrc/psb_survey.cu) on the product's own Hermite table."""
        import torch
        radecz = np.asarray(radecz, dtype=np.float64)
        zmax = max(float(radecz[2].max()), 1e-6)
        tab = self._real.survey_distance_table(cosmo, zmax)
        nn = tab.shape[0] - 1
        u = radecz[2] * (nn / zmax)
        k = np.clip(u.astype(np.int64), 0, nn - 1)
        t = u - k
        t2, t3 = t * t, t * t * t
        rad = (2 * t3 - 3 * t2 + 1) * tab[k, 0] + (t3 - 2 * t2 + t) * tab[k, 1] + (-2 * t3 + 3 * t2) * tab[k + 1, 0] + (t3 - t2) * tab[k + 1, 1]
        ra, dec = radecz[0] * (np.pi / 180.), radecz[1] * (np.pi / 180.)
        xyz = np.array([rad * np.cos(dec) * np.cos(ra), rad * np.cos(dec) * np.sin(ra), rad * np.sin(dec)])
        w0 = np.ones(radecz.shape[1]) if w is None else np.asarray(w, dtype=np.float64)
        nb = np.asarray(nb, dtype=np.float64)
        wf = w0 * (1. / (1. + nb * P0_fkp))
        out = np.array([w0.sum(), (wf ** 2).sum(), (wf ** 3).sum(), (nb * wf ** 2).sum(), (nb * wf ** 3).sum(), (nb ** 2 * wf ** 3).sum()]
                       + list(xyz.min(axis=1)) + list(xyz.max(axis=1)))
        return torch.from_numpy(xyz.astype(np.float32)), torch.from_numpy(wf.astype(np.float32)), out

    def fft_survey(self, xyz, w, Lbox):
        N = self.N
        xyz, w = xyz.numpy(), w.numpy()
        xyzs = np.zeros([3, xyz.shape[1]], dtype=np.float32, order='F')
        xyzs[:] = xyz
        _delta = np.zeros([2 * N, N, N], dtype=np.float32, order='F')
        O.assign_quad(xyzs, w, _delta, np.float32(float(N) / Lbox), 0.5 * N, 0, 0, 0, 0)
        d = O._FFT(_delta, N)
        O.fcomb_survey(d, N)
        return self._to_half_tensor(d[:N // 2 + 1])

    def shell_mode_counts(self, step, Nmax):
        irk = O.shell_index(self.N, step)
        return np.array([np.sum(irk == i) for i in range(Nmax + 1)])

    def counts(self, Nmax, Ncut, step, fft='pyfftw', silent=True):
        return O._counts_Bk123(Ngrid=self.N, Nmax=Nmax, Ncut=Ncut, step=step)

    def bispectrum_sums(self, half, step, Ncut, Nmax):
        N = self.N
        hf = half.numpy().reshape(N, N, self.h + 1, 2).view(np.complex64)[..., 0].transpose(2, 1, 0)
        full = O.reflect_delta(hf, N)
        s0 = Ncut // step
        fields = O.shell_fields(full, O.shell_index(N, step), range(s0, Nmax + 1))
        tri = O.triangle_list(Nmax, Ncut, step)
        sums = np.array([O._triple(fields[i], fields[j], fields[l]) for (i, j, l) in tri])
        sumsq = np.array([np.dot(fields[j].astype(np.float64), fields[j].astype(np.float64)) for j in range(s0, Nmax + 1)])
        return sums, sumsq


@pytest.mark.parametrize('tag', ['A', 'B'])
def test_product_survey_host_logic_with_stubbed_device(golden_dir, tag, monkeypatch):
    from pyspectrum_b200 import pyspectrum as pySpec
    g = _load(golden_dir, 'survey_%s.npz' % tag)
    N, L, P0 = int(g['Ngrid']), float(g['Lbox']), float(g['P0_fkp'])
    stub = _StubPipe(N)
    monkeypatch.setattr(pySpec.PeriodicPipeline, 'get', classmethod(lambda cls, Ngrid: stub))
    w = g.get('w')
    w0, r0 = (None if w is None else w.copy()), g['radecz'].copy()
    out = pySpec.FFT_survey_mono(g['radecz'], g['nbar'], w=w, P0_fkp=P0, Lbox=L, Ngrid=N)
    # positions go through the Hermite distance table (<= 1 float32 ulp away from the reference's): not bit-identical
    assert out[0].shape == g['delta_d'].shape and np.abs(out[0] - g['delta_d']).max() <= 1e-6 * np.abs(g['delta_d']).max()
    np.testing.assert_allclose(np.array(out[1:]), g['sums_d'], rtol=1e-13)
    assert np.array_equal(g['radecz'], r0) and (w is None or np.array_equal(w, w0))
    alpha = g['sums_d'][0] / g['sums_r'][0]
    Is = alpha * g['sums_r'][1:]
    for (step, Ncut, Nmax) in SURVEY_CFGS[tag]:
        pre = 'b0_s%d_c%d_m%d_' % (step, Ncut, Nmax)
        bk = pySpec.B0_survey(g['radecz'], g['nbar'], w=w, radecz_r=g['radecz_r'], nbar_r=g['nbar_r'], P0_fkp=P0, Lbox=L,
                              Ngrid=N, step=step, Ncut=Ncut, Nmax=Nmax)
        deltak = O.reflect_delta(g['delta_d'], N) - alpha * O.reflect_delta(g['delta_r'], N)
        bk2 = pySpec._B0_survey(deltak, alpha, *Is, Nmax=Nmax, Ncut=Ncut, step=step)
        for b in (bk, bk2):
            for key in ['i_k1', 'i_k2', 'i_k3']:
                assert np.array_equal(b[key], g[pre + key])
            np.testing.assert_allclose(b['counts'], g[pre + 'counts'], rtol=1e-12)
            tol = 1e-8 if b is bk2 else 1e-5                          # bk: positions through the Hermite table (see above)
            for key in ['p0k1', 'p0k2', 'p0k3', 'b123', 'q123']:
                np.testing.assert_allclose(b[key], g[pre + key], rtol=tol, atol=tol * np.abs(g[pre + key]).max(), err_msg=key)
        assert sorted(bk.keys()) == sorted(['i_k1', 'i_k2', 'i_k3', 'p0k1', 'p0k2', 'p0k3', 'b123', 'q123', 'counts', 'meta'])
    with pytest.raises(ValueError):
        pySpec.B0_survey(g['radecz'], g['nbar'], radecz_r=g['radecz_r'], Lbox=L, Ngrid=N)
    with pytest.raises(AssertionError):
        pySpec.FFT_survey_mono(g['radecz'], g['nbar'], Lbox=1000., Ngrid=N)


def test_half_from_full_is_the_hermitian_part():
    from pyspectrum_b200 import pyspectrum as pySpec
    N = 12
    stub = _StubPipe(N)
    rng = np.random.default_rng(9)
    full = (rng.normal(size=(N, N, N)) + 1j * rng.normal(size=(N, N, N))).astype(np.complex64)
    half = stub.half_from_full(full).numpy().view(np.complex64)[..., 0].transpose(2, 1, 0)        # back to [kx,ky,kz]
    # Re(FFT(d)) == FFT(Hermitian part of d): the identity the shell stage relies on
    re = np.real(np.fft.fftn(full.astype(np.complex128)))
    herm = O.reflect_delta(half, N).astype(np.complex128)
    idx = (-np.arange(N)) % N
    assert np.abs(herm - np.conj(herm[idx][:, idx][:, :, idx])).max() == 0.                       # exactly Hermitian
    got = np.fft.fftn(herm)
    assert np.abs(got.imag).max() < 1e-4 and np.abs(got.real - re).max() < 1e-4 * np.abs(re).max()
    # an already Hermitian field passes through unchanged
    h2 = stub.half_from_full(O.reflect_delta(half, N)).numpy().view(np.complex64)[..., 0].transpose(2, 1, 0)
    assert np.array_equal(h2, O.reflect_delta(half, N)[:N // 2 + 1])
    with pytest.raises(ValueError):
        stub.half_from_full(np.zeros((N, N, N // 2 + 1), np.complex64))


def test_survey_distance_table_accuracy():
    """The Hermite table the device kernel interpolates: node slopes from the spline vs the exact integrand, and the
    interpolant vs the direct quadrature between the nodes."""
    from pyspectrum_b200.pyspectrum import PeriodicPipeline
    cos = UT.FlatLambdaCDM(67.6, 0.31)
    for zmax in (0.05, 0.7, 3.0):
        tab = PeriodicPipeline.survey_distance_table(cos, zmax)
        nn = tab.shape[0] - 1
        zn = np.linspace(0., zmax, nn + 1)
        exact_slope = 299792.458 / cos.H0 / cos.efunc(zn) * cos.h * (zmax / nn)
        assert np.abs(tab[:, 1] / exact_slope - 1.).max() < 1e-9
        z = np.random.default_rng(0).uniform(0., zmax, 5000)
        u = z * (nn / zmax)
        k = np.clip(u.astype(np.int64), 0, nn - 1)
        t = u - k
        t2, t3 = t * t, t * t * t
        d = (2 * t3 - 3 * t2 + 1) * tab[k, 0] + (t3 - 2 * t2 + t) * tab[k, 1] + (-2 * t3 + 3 * t2) * tab[k + 1, 0] + (t3 - t2) * tab[k + 1, 1]
        ref = cos.comoving_distance(z) * cos.h
        assert np.abs(d - ref).max() < 1e-9 * ref.max()


def test_product_bk_periodic_host_logic_with_stubbed_device(golden_dir, monkeypatch):
    """_Bk_periodic (pyspectrum.py:359-457) and its epilogue on the reference golden's field, device stages stubbed as above."""
    from pyspectrum_b200 import pyspectrum as pySpec
    g = dict(np.load(os.path.join(golden_dir, 'small_A.npz')))
    N, L = int(g['Ngrid']), float(g['Lbox'])
    stub = _StubPipe(N)
    monkeypatch.setattr(pySpec.PeriodicPipeline, 'get', classmethod(lambda cls, Ngrid: stub))
    full = O.reflect_delta(g['delta_half'], N)
    for (step, Ncut, Nmax) in [(3, 3, 4), (2, 3, 6), (1, 1, 8)]:
        pre = 'bk_s%d_c%d_m%d_' % (step, Ncut, Nmax)
        bk = pySpec._Bk_periodic(full, Nmax=Nmax, Ncut=Ncut, step=step)
        nbar = g['xyz'].shape[1] / L ** 3
        kf = 2 * np.pi / L
        for key in ['i_k1', 'i_k2', 'i_k3']:
            assert np.array_equal(bk[key], g[pre + key])
        np.testing.assert_allclose(bk['counts'], g[pre + 'counts'], rtol=1e-12)
        np.testing.assert_allclose(bk['p0k1'] * (2 * np.pi) ** 3 / kf ** 3 - 1. / nbar, g[pre + 'p0k1'], rtol=1e-9)
        raw = (g[pre + 'b123'] + g[pre + 'b123_sn']) * kf ** 6 / (2 * np.pi) ** 6
        np.testing.assert_allclose(bk['b123'], raw, rtol=1e-8, atol=1e-10 * np.abs(raw).max())
    with pytest.raises(ValueError):
        pySpec._Bk_periodic(full, Nmax=4, Ncut=1, step=3)           # Ncut//step == 0
