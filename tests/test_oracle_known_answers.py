"""CPU: known-answer tests for the C restatement of estimator.f."""
import numpy as np
import pytest

from oracle import pyspec_oracle as O


def test_mesh_mass_is_216_sum_w_per_grid():
    rng = np.random.default_rng(0)
    N, L = 20, 50.
    xyz = rng.uniform(0, L, (3, 777))
    w = rng.uniform(0.5, 1.5, 777)
    m = O.assign_mesh(xyz, w, L, N)
    assert abs(m[::2].sum() / w.sum() - 216.) < 1e-4
    assert abs(m[1::2].sum() / w.sum() - 216.) < 1e-4


def test_single_particle_stencil_values():
    """Particle at u = 5.25 cells: grid A weights per axis ((1-h)^3, 4+(3h-6)h^2, ., h^3), h=.25."""
    N, L = 16, 16.
    xyz = np.array([[5.25], [5.25], [5.25]])
    m = O.assign_mesh(xyz, None, L, N)
    A = m[::2]
    h = np.float32(0.25)
    wx = np.array([(1 - h) ** 3, 4 + (3 * h - 6) * h * h, 0, h ** 3], dtype=np.float64)
    wx[2] = 6 - wx[0] - wx[1] - wx[3]
    for a, ca in enumerate(range(4, 8)):
        for b, cb in enumerate(range(4, 8)):
            for c, cc in enumerate(range(4, 8)):
                assert abs(A[ca, cb, cc] - wx[a] * wx[b] * wx[c]) < 1e-4
    assert np.count_nonzero(A) == 64
    # grid B is shifted by half a cell: u+0.5 = 5.75 -> base cell 5, h = 0.75
    B = m[1::2]
    assert np.count_nonzero(B) == 64 and B[4, 4, 4] > 0 and B[3, 4, 4] == 0 and B[7, 7, 7] > 0


def test_clip_not_wrap():
    """py:938-941: out-of-box particles are clipped onto the faces."""
    N, L = 16, 16.
    a = O.assign_mesh(np.array([[-3.0], [20.0], [8.0]]), None, L, N)
    b = O.assign_mesh(np.array([[0.0], [L * (1. - 1e-6)], [8.0]]), None, L, N)
    assert np.array_equal(a, b)


def test_delta_k0_is_one_and_plane_wave():
    N, L = 24, 100.
    x1 = np.array([[13.3], [47.7], [88.1]])
    d = O.FFT_periodic(x1, None, L, N)
    assert abs(d[0, 0, 0] - 1.) < 1e-6
    k = np.arange(N // 2 + 1)
    kk = np.array([i if i <= N // 2 else i - N for i in range(N)])
    ph = np.exp(1j * 2 * np.pi / L * (k[:, None, None] * x1[0] + kk[None, :, None] * x1[1] + kk[None, None, :] * x1[2]))
    kmag = np.sqrt(k[:, None, None] ** 2 + kk[None, :, None] ** 2 + kk[None, None, :] ** 2)
    assert np.abs(d - ph)[kmag < N / 4].max() < 1e-3          # interlaced PCS: aliasing ~1e-4 here


def test_fcomb_self_conjugate_planes_are_hermitian_pairs():
    """After fcomb the kx=0 and kx=N/2 planes hold exact conjugate pairs (the later write wins,
    estimator.f:658-665) -- except the 4 self-conjugate points of each plane, which keep a complex
    value (reflect_delta, py:1149-1156, is what makes 7 of them real)."""
    rng = np.random.default_rng(3)
    N, L = 12, 30.
    d = np.ascontiguousarray(O.FFT_periodic(rng.uniform(0, L, (3, 300)), None, L, N))
    for ix in (0, N // 2):
        P = d[ix]
        M = np.conj(P[(-np.arange(N)) % N][:, (-np.arange(N)) % N])
        bad = {tuple(b) for b in np.argwhere(P != M).tolist()}
        assert bad <= {(0, 0), (0, N // 2), (N // 2, 0), (N // 2, N // 2)}


def test_randoms_give_unit_shot_noise_power():
    rng = np.random.default_rng(5)
    N, L = 32, 100.
    pk = O.Pk_periodic(rng.uniform(0, L, (3, 50000)), None, L, N)
    r = (pk['p0k'] + pk['p0k_sn']) / pk['p0k_sn']
    assert abs(np.average(r[3:], weights=pk['counts'][3:]) - 1.) < 0.02


def test_f2py_shape_and_order_checks():
    with pytest.raises(ValueError):
        O.assign_quad(np.zeros((3, 1), np.float32), np.ones(1, np.float32), np.zeros((8, 4, 4), np.float32), 1., 0, 0, 0, 0, 0)
    with pytest.raises(ValueError):
        O.fcomb_periodic(np.zeros((4, 4, 4), np.complex64), 1.)


# ---------------------------------------------------------------------------- survey-geometry pieces (estimator.f:677-745, py:817-823)
def test_fcomb_survey_is_fcomb_periodic_with_unit_weight():
    """The two subroutines differ only in cf: 1/(6^3 4 N) (f:614) against 1/(6^3 4) (f:686)."""
    rng = np.random.default_rng(8)
    N = 12
    F = np.asfortranarray((rng.normal(size=(N, N, N)) + 1j * rng.normal(size=(N, N, N))).astype(np.complex64))
    a, b = F.copy(order='F'), F.copy(order='F')
    O.fcomb_survey(a, N)
    O.fcomb_periodic(b, 1.)
    assert np.array_equal(a, b)
    c = F.copy(order='F')
    O.fcomb_periodic(c, 4.)                                   # exact power of two: the same values scaled
    assert np.array_equal(c * np.float32(4.), a)


def test_assign_offset_is_a_half_box_shift():
    """assign_quad(offset = N/2) on positions in (-L/2, L/2) (py:819) = assign_quad(offset = 0) on positions + L/2,
    when the shift is exact in float32 (kf_ks = 1, half-integer offsets)."""
    rng = np.random.default_rng(4)
    N = 16
    r = np.asfortranarray(rng.uniform(-7.5, 7.5, (3, 500)).astype(np.float32))
    w = rng.uniform(0.5, 2., 500).astype(np.float32)
    a = np.zeros((2 * N, N, N), np.float32, order='F')
    b = np.zeros((2 * N, N, N), np.float32, order='F')
    O.assign_quad(r, w, a, np.float32(1.), 0.5 * N, 0, 0, 0, 0)
    O.assign_quad(np.asfortranarray(r + np.float32(8.)), w, b, np.float32(1.), 0., 0, 0, 0, 0)
    assert np.abs(a - b).max() <= 2e-6 * np.abs(b).max()     # (r + 1) + 8 vs (r + 8) + 1: one float32 rounding apart
    assert abs(a[::2].sum(dtype=np.float64) / w.sum(dtype=np.float64) - 216.) < 1e-3


def test_survey_delta_of_one_object_is_a_plane_wave():
    """One object at x (relative to the box centre): delta_0(k) = w_fkp e^{+i k.(x + L/2)} up to window/aliasing (py:817-823)."""
    N, L = 24, 1000.
    cos = O.FlatLambdaCDM(67.6, 0.31)
    radecz = np.array([[35.], [12.], [0.1]])
    nb = np.array([2e-4])
    d, Ntot, I12, I13, I22, I23, I33 = O.FFT_survey_mono(radecz, nb, P0_fkp=1e4, Lbox=L, Ngrid=N, cosmo=cos)
    wf = 1. / (1. + 2e-4 * 1e4)
    assert Ntot == 1. and np.isclose(I12, wf ** 2) and np.isclose(I33, nb[0] ** 2 * wf ** 3)
    x = O.radecz_to_cartesian(radecz, cos)[:, 0] + 0.5 * L
    assert abs(d[0, 0, 0] - wf) < 1e-6
    k = np.arange(N // 2 + 1)
    kk = np.array([i if i <= N // 2 else i - N for i in range(N)])
    ph = wf * np.exp(1j * 2 * np.pi / L * (k[:, None, None] * x[0] + kk[None, :, None] * x[1] + kk[None, None, :] * x[2]))
    kmag = np.sqrt(k[:, None, None] ** 2 + kk[None, :, None] ** 2 + kk[None, None, :] ** 2)
    assert np.abs(d - ph)[kmag < N / 4].max() < 1e-3 * wf


def test_pk_pbox_rsd_single_mode_lands_in_its_bin():
    """One mode pair +-k with |delta|^2 = 1 on the half grid: nk counts both partners, k = kf |k|, P0 kf^3 = mean |delta|^2 and
    the quadrupole follows L2(mu) with mu = kz/|k| for the line of sight along z (estimator.f:196-262)."""
    N, Lbox, nbin, nmu = 16, 100, 8, 4
    d = np.zeros((N // 2 + 1, N, N), np.complex64, order='F')
    kx, ky, kz = 2, 1, 2                                        # |k| = 3 exactly
    d[kx, ky, kz] = 1.
    k, p0, p2, p4, nk, km, mk, pkm, nkm = O.pk_pbox_rsd(d, 2, Lbox, nbin, nmu)
    kf = 2 * np.pi / Lbox
    b = 2                                                       # nint(nbin * 3 / (N/2)) - 1 = bin index 3 -> slot 2
    assert nk[b] > 0 and np.isclose(k[b], kf * 3., rtol=0.1)    # mean |k| over all 98 modes with 2.5 <= |k| < 3.5
    tot = p0 * nk * np.float64(np.float32(kf) ** 3)
    assert np.isclose(tot[b], 2., rtol=1e-5) and np.allclose(np.delete(tot, b), 0.)      # both partners of the pair counted
    mu = kz / 3.
    assert np.isclose(p2[b] / p0[b], 5. * (-0.5 + 1.5 * mu * mu), rtol=1e-5)
    assert np.isclose(p4[b] / p0[b], 9. * (0.375 - 3.75 * mu ** 2 + 4.375 * mu ** 4), rtol=1e-4)


# ------------------------------------------------------------------------------ quadrupole-field pieces (f:294-300, 514-603)
def _half_fields(N, n, seed):
    rng = np.random.default_rng(seed)
    return [np.asfortranarray((rng.normal(size=(N // 2 + 1, N, N)) + 1j * rng.normal(size=(N // 2 + 1, N, N))).astype(np.complex64))
            for _ in range(n)]


def _kgrid(N):
    s = lambda i: np.where(i <= N // 2, i, i - N)
    kx, ky, kz = np.meshgrid(s(np.arange(N // 2 + 1)), s(np.arange(N)), s(np.arange(N)), indexing='ij')
    rk = np.sqrt((kx ** 2 + ky ** 2 + kz ** 2).astype(np.float32))
    return kx, ky, kz, rk


@pytest.mark.parametrize('N', [8, 12])
def test_quadrupole_field_combinations(N):
    """FiveDelta2g_1/_2 and build_quad against an independent numpy evaluation.  The unit-vector components of FiveDelta2g_*
    are implicitly INTEGER in estimator.f, so only on-axis modes pick up a Q_ii term and every cross term vanishes."""
    kx, ky, kz, rk = _kgrid(N)
    nz = rk > 0
    with np.errstate(divide='ignore', invalid='ignore'):
        ih = [np.where(nz, np.trunc(k.astype(np.float32) / rk), 0).astype(np.float32) for k in (kx, ky, kz)]
    xx, yy, zz = _half_fields(N, 3, N)
    want = np.where(nz, np.float32(7.5) * (xx * ih[0] ** 2 + yy * ih[1] ** 2 + zz * ih[2] ** 2), xx).astype(np.complex64)
    got = xx.copy(order='F')
    O.fivedelta2g_1(got, yy, zz)
    assert np.array_equal(got, want)
    on_axis = ((kx != 0).astype(int) + (ky != 0) + (kz != 0)) == 1
    assert np.all(got[nz & ~on_axis] == 0) and np.all(got[on_axis] != 0) and got[0, 0, 0] == xx[0, 0, 0]
    dcg, xy, yz, zx = _half_fields(N, 4, N + 1)
    want = np.where(nz, xx - np.float32(2.5) * dcg, xx).astype(np.complex64)      # integer kxh kyh = 0 everywhere
    got = xx.copy(order='F')
    O.fivedelta2g_2(dcg, got, xy, yz, zx)
    assert np.array_equal(got, want)
    for irsd, k in ((1, kx), (2, ky), (3, kz)):
        with np.errstate(divide='ignore', invalid='ignore'):
            amu = (k.astype(np.float32) / rk).astype(np.float32)
        fac = (np.float32(7.5) * (amu * amu) - np.float32(2.5)).astype(np.float32)
        want = np.where(nz, np.stack([fac * dcg.real, fac * dcg.imag], -1).astype(np.float32).view(np.complex64)[..., 0], xx)
        got = xx.copy(order='F')
        O.build_quad(dcg, got, irsd)
        assert np.array_equal(got, want.astype(np.complex64)), irsd


def test_quadrupole_weights_of_assign_quad():
    """f:294-300: the Q_ij mesh equals the delta mesh of the same particles with weights w r_ia r_ib / r^2 (float32, left to
    right), likewise the four-index form; Q_xx + Q_yy + Q_zz gives back the delta mesh."""
    rng = np.random.default_rng(3)
    N, L, Np = 12, 30., 500
    r = np.asfortranarray(rng.uniform(1., L, (3, Np)).astype(np.float32))
    w = rng.uniform(0.5, 2., Np).astype(np.float32)
    kf_ks = np.float32(N / L)
    mesh = lambda wt, *idx: (lambda d: (O.assign_quad(r, wt, d, kf_ks, 0, *idx), d)[1])(np.zeros((2 * N, N, N), np.float32, order='F'))
    rn = (r[0] * r[0] + r[1] * r[1]) + r[2] * r[2]
    for ia, ib in ((1, 1), (2, 3), (3, 1)):
        assert np.array_equal(mesh(w, ia, ib, 0, 0), mesh((w * r[ia - 1] * r[ib - 1] / rn).astype(np.float32), 0, 0, 0, 0))
    assert np.array_equal(mesh(w, 1, 2, 3, 3), mesh((w * r[0] * r[1] * r[2] * r[2] / (rn * rn)).astype(np.float32), 0, 0, 0, 0))
    tr = mesh(w, 1, 1, 0, 0) + mesh(w, 2, 2, 0, 0) + mesh(w, 3, 3, 0, 0)
    d0 = mesh(w, 0, 0, 0, 0)
    assert np.abs(tr - d0).max() < 1e-5 * np.abs(d0).max()


@pytest.mark.parametrize('N,irsd,Nmu', [(12, 0, 5), (12, 1, 5), (12, 2, 5), (16, 2, 10), (20, 1, 7)])
def test_pk_pbox_rsd_against_an_independent_numpy_restatement(N, irsd, Nmu):
    """estimator.f:155-264 restated a second time (tests/pk_pbox_rsd_numpy.py: vectorised numpy written from the Fortran text) against
    the C restatement the rest of the suite relies on: mode counts and (k,mu) counts exact, every sum to 1e-12 (same float32 / float64
    typing, same accumulation order).  Two independent translations agreeing narrows the 'Fortran layer unpinned' gap for a6."""
    import pk_pbox_rsd_numpy as PN
    rng = np.random.default_rng(N + irsd)
    dtl = np.asfortranarray((rng.normal(size=(N // 2 + 1, N, N)) + 1j * rng.normal(size=(N // 2 + 1, N, N))).astype(np.complex64))
    Nbin = N // 2
    a = O.pk_pbox_rsd(dtl, irsd, 100, Nbin, Nmu)
    b = PN.pk_pbox_rsd(dtl, irsd, 100, Nbin, Nmu)
    names = ('k', 'p0', 'p2', 'p4', 'nk', 'km', 'mk', 'pkm', 'nkm')
    for name, x, y in zip(names, a, b):
        x, y = np.asarray(x), np.asarray(y)
        if name in ('nk', 'nkm'):
            assert np.array_equal(x, y), name
        else:
            np.testing.assert_allclose(x, y, rtol=1e-12, atol=0, err_msg=name)
    assert a[4].sum() > 0 and a[8].sum() > 0


@pytest.mark.parametrize('N', [8, 12])
@pytest.mark.parametrize('periodic', [True, False])
def test_fcomb_against_an_independent_python_restatement(N, periodic):
    """estimator.f:605-745 restated a second time (tests/fcomb_numpy.py: a sequential Python loop written from the Fortran text, with its
    implicit typing and in-place update order) against the C restatement: every element of the combined field, self-conjugate planes
    included, bit for bit."""
    import fcomb_numpy as FN
    rng = np.random.default_rng(N)
    F = np.asfortranarray((rng.normal(size=(N, N, N)) + 1j * rng.normal(size=(N, N, N))).astype(np.complex64))
    a = F.copy(order='F')
    b = F.copy(order='F')
    if periodic:
        O.fcomb_periodic(a, 1234.5)
        FN.fcomb(b, 1234.5)
    else:
        O.fcomb_survey(a)
        FN.fcomb(b, None)
    assert np.abs(a).max() > 0 and a.tobytes(order='F') == b.tobytes(order='F')                 # bit for bit, all 8 mirror images


@pytest.mark.parametrize('N,offset,idx', [(8, 0., (0, 0, 0, 0)), (12, 0., (0, 0, 0, 0)), (8, 4., (0, 0, 0, 0)), (8, 0., (1, 2, 0, 0)), (8, 0., (3, 3, 1, 2))])
def test_assign_quad_against_an_independent_python_restatement(N, offset, idx):
    """estimator.f:284-512 restated a second time (tests/assign_quad_numpy.py: a sequential Python loop over float32 scalars written from
    the Fortran text) against the C restatement: the interlaced mesh bit for bit -- face particles (stencil wrap), the unwrapped base
    cell of grid A, the wrapped one of grid B, the survey offset, the Q_ij / Q_ijkl weights."""
    import assign_quad_numpy as AN
    rng = np.random.default_rng(N + int(offset) + sum(idx))
    L, Np = 50., 300
    r = np.asfortranarray(rng.uniform(0.5 if offset == 0. else -0.45 * L, L * (1 - 1e-6) if offset == 0. else 0.45 * L, (3, Np)).astype(np.float32))
    if offset == 0.:
        r[:, 0] = [0.5, L * (1 - 1e-6), 0.5 * L]
        r[:, 1] = [L * (1 - 1e-6), L * (1 - 1e-6), L * (1 - 1e-6)]
    w = rng.uniform(0.5, 2., Np).astype(np.float32)
    a = np.zeros((2 * N, N, N), np.float32, order='F')
    b = np.zeros((2 * N, N, N), np.float32, order='F')
    O.assign_quad(r, w, a, np.float32(N / L), offset, *idx)
    AN.assign_quad(r, w, b, np.float32(N / L), offset, *idx)
    assert np.abs(a).max() > 0 and a.tobytes(order='F') == b.tobytes(order='F')
