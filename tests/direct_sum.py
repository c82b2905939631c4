"""Independent check of a1+a2+a3 (assign_quad -> backward FFT -> fcomb_periodic; estimator.f:284-512, py:1060-1080, f:605-675)
that does NOT go through the C restatement: the interlaced, window-corrected delta(k) of a PCS mesh has a closed form,

    delta(k) = 1/sum(w) * sum_p w_p * sum_{n in Z^3, nx+ny+nz even}  prod_a  [ sinc^4((t_a + 2 pi n_a)/2) / sinc^4(t_a/2) ]
                                                                             * exp(+i (t_a + 2 pi n_a) v_pa),
    t_a = 2 pi k_a / N (signed k_a, |k_a| < N/2),   v_p = kf_ks * r_p + offset  (0-based continuous grid coordinate),

because the four weights of f:316-320 are 6 M4(v - g) (cubic B-spline, Fourier transform 6 sinc^4(q/2)), Poisson summation turns
the sum over grid points into the sum over aliases n, grid B (shifted by half a cell, f:305-307) carries (-1)^(nx+ny+nz) after the
phase rec = exp(-i pi (kx+ky+kz)/N) of f:619-623, and 2 (A^ + rec B^) / (6^3 4 sum(w) Wx Wy Wz) (f:615, 632-657) keeps the even
aliases only.  The parity constraint factorises: sum_{even} prod_a T_a(n_a) = (prod_a S+_a + prod_a S-_a) / 2 with
S+-_a = sum_n (+-1)^n T_a(n).  Evaluated in float64 by direct summation over particles: O(Np * Nk * aliases), toy sizes only."""
import numpy as np


def grid_coordinate_f32(r, kf_ks, offset=0.):
    """v = (kf_ks * r + 1 + offset) - 1 with the float32 rounding of estimator.f:302-304, returned as float64."""
    rp = (np.float32(kf_ks) * np.asarray(r, dtype=np.float32) + np.float32(1.)) + np.float32(offset)
    return rp.astype(np.float64) - 1.0


def direct_delta(xyz, w, Lbox, N, kvecs, M=60, offset=0.):
    """delta(k) for the signed integer wave vectors `kvecs` [nk,3] (|k_a| < N/2): float64 complex array [nk]."""
    xyz = np.asarray(xyz, dtype=np.float64)
    Np = xyz.shape[1]
    w = np.ones(Np) if w is None else np.asarray(w, dtype=np.float64)
    kf_ks = np.float32(float(N) / Lbox)
    r32 = np.clip(xyz, 0., Lbox * (1. - 1e-6)).astype(np.float32)             # py:938-941
    v = grid_coordinate_f32(r32, kf_ks, offset)                                # [3, Np]
    n = np.arange(-M, M + 1)
    sgn = np.where(n % 2 == 0, 1.0, -1.0)
    out = np.zeros(len(kvecs), dtype=np.complex128)
    for q, kv in enumerate(np.asarray(kvecs)):
        Sp, Sm = [], []
        for a in range(3):
            t = 2 * np.pi * kv[a] / N
            arg = 0.5 * (t + 2 * np.pi * n)                                    # [2M+1]
            with np.errstate(divide='ignore', invalid='ignore'):
                s = np.where(arg == 0., 1., np.sin(arg) / arg) ** 4
            s0 = 1.0 if t == 0 else (np.sin(0.5 * t) / (0.5 * t)) ** 4
            T = (s / s0)[:, None] * np.exp(1j * np.outer(t + 2 * np.pi * n, v[a]))   # [2M+1, Np]
            Sp.append(T.sum(axis=0))
            Sm.append((sgn[:, None] * T).sum(axis=0))
        out[q] = np.sum(w * 0.5 * (Sp[0] * Sp[1] * Sp[2] + Sm[0] * Sm[1] * Sm[2])) / w.sum()
    return out


def test_wavevectors(N, seed=0, nk=40):
    """A reproducible sample of signed wave vectors away from the Nyquist planes (incl. the axes and k = 0)."""
    rng = np.random.default_rng(seed)
    h = N // 2
    ks = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (h - 1, 0, 0), (0, -(h - 1), 0), (h - 1, h - 1, -(h - 1)), (2, -3, 1)]
    while len(ks) < nk:
        ks.append((int(rng.integers(0, h)), int(rng.integers(-h + 1, h)), int(rng.integers(-h + 1, h))))
    return np.array(ks)


def pick(delta_half, N, kvecs):
    """Values of a half field indexed [kx, ky mod N, kz mod N] (what FFT_periodic returns) at signed k with kx >= 0."""
    kv = np.asarray(kvecs)
    return np.array([delta_half[k[0], k[1] % N, k[2] % N] for k in kv])
