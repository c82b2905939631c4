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
