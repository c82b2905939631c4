"""A second, independent restatement of estimator.f:155-264 (pk_pbox_rsd), written straight from the Fortran text as vectorised
numpy -- not derived from oracle/estimator_oracle.c.  Used by tests/test_oracle_known_answers.py to cross-check the C
restatement (two translations of the same 110 lines agreeing mode for mode), since no Fortran compiler is available to run the
original.  float32 where the Fortran is REAL, float64 where it is REAL*8; np.add.at accumulates in loop order (iz, iy, ix)."""
import math

import numpy as np

f32 = np.float32


def pk_pbox_rsd(dtl, irsd, Lbox, Nbin, Nmu):
    """dtl: complex64 (Ngrid/2+1, Ngrid, Ngrid) indexed [ix, iy, iz] (Fortran order or not).  Returns k,p0,p2,p4,nk,km,mk,pkm,nkm."""
    N = dtl.shape[1]
    h = N // 2
    pi = f32(3.141592654)                                         # f:163
    tpi = f32(2.) * pi
    kf = tpi / f32(int(Lbox))                                     # f:169, Lbox INTEGER
    mubin = f32(1.) / f32(Nmu)                                    # f:171
    thetaobs, phiobs = {0: (f32(0.5) * pi, f32(0.)), 1: (f32(0.5) * pi, f32(0.5) * pi), 2: (f32(0.), f32(0.))}[irsd]
    sf = lambda x: f32(math.sin(float(x)))                        # single-precision sin / cos of a single-precision argument
    cf = lambda x: f32(math.cos(float(x)))
    i1 = np.arange(1, N + 1)
    rk1 = ((i1 + h - 2) % N - h + 1).astype(f32)                  # f:198, 201, 204 (all three axes)
    ic1 = (N + 1 - i1) % N + 1                                    # f:197, 200, 203
    rkz, rky, rkx = np.meshgrid(rk1, rk1, rk1, indexing='ij')     # arrays [iz][iy][ix]: C order = the Fortran loop order
    iz, iy, ix = np.meshgrid(i1, i1, i1, indexing='ij')
    rk = np.sqrt((rkx * rkx + rky * rky) + rkz * rkz)             # f32
    q = (f32(Nbin) * rk) / f32(h)
    imk = np.floor(q.astype(np.float64) + 0.5).astype(np.int64)   # nint of a non-negative value
    sel = (imk <= Nbin) & (imk != 0)
    with np.errstate(divide='ignore', invalid='ignore'):
        cot1 = rkz / rk
        sit1 = np.sqrt(f32(1.) - cot1 * cot1)
        den = rk * sit1
        cp, sp = rkx / den, rky / den
        cc = np.where(sit1 > 0, sf(phiobs) * sp + cf(phiobs) * cp, f32(0.)).astype(f32)
        mu = (cf(thetaobs) * cot1 + (sf(thetaobs) * sit1) * cc).astype(np.float64)
    amu = np.abs(mu)
    with np.errstate(invalid='ignore'):
        imu = np.where(sel, np.trunc((amu + np.float64(mubin)) / np.float64(mubin)), 0).astype(np.int64)
    mu2 = mu * mu
    Le2 = -0.5 + 1.5 * mu2
    Le4 = 0.375 - 3.75 * mu2 + 4.375 * (mu2 * mu2)
    half = ix <= h + 1
    jx = np.where(half, ix, (N + 1 - ix) % N + 1) - 1
    jy = np.where(half, iy, ic1[iy - 1]) - 1
    jz = np.where(half, iz, ic1[iz - 1]) - 1
    ct = np.asarray(dtl)[jx, jy, jz]
    # cabs: libm's cabsf / hypotf evaluates in double and rounds once (numpy's own complex64 abs loop differs in the last bit)
    ab = np.sqrt(ct.real.astype(np.float64) ** 2 + ct.imag.astype(np.float64) ** 2).astype(f32)
    pk = (ab * ab).astype(np.float64)                             # (cabs(ct))**2 in single, then dble
    kk = (kf * rk).astype(np.float64)
    k, p0, p2, p4, nk = (np.zeros(Nbin) for _ in range(5))
    km, mk, pkm, nkm = (np.zeros((Nbin, Nmu)) for _ in range(4))
    s = sel.ravel()
    b = imk.ravel()[s] - 1
    np.add.at(nk, b, 1.0)
    np.add.at(k, b, kk.ravel()[s])
    np.add.at(p0, b, pk.ravel()[s])
    np.add.at(p2, b, (pk * 5.0 * Le2).ravel()[s])
    np.add.at(p4, b, (pk * 9.0 * Le4).ravel()[s])
    t = s & ((imu.ravel() <= Nmu) & (imu.ravel() > 0))
    bt, mt = imk.ravel()[t] - 1, imu.ravel()[t] - 1
    np.add.at(nkm, (bt, mt), 1.0)
    np.add.at(km, (bt, mt), kk.ravel()[t])
    np.add.at(mk, (bt, mt), amu.ravel()[t])
    np.add.at(pkm, (bt, mt), pk.ravel()[t])
    kf3 = np.float64(kf * kf * kf)                                # dble(kf**3): the cube in single
    ok = nk > 0
    k[ok] /= nk[ok]
    for a in (p0, p2, p4):
        a[ok] = a[ok] / nk[ok] / kf3
    ok2 = nkm > 0
    km[ok2] /= nkm[ok2]
    mk[ok2] /= nkm[ok2]
    pkm[ok2] = pkm[ok2] / nkm[ok2] / kf3
    return k, p0, p2, p4, nk, km, mk, pkm, nkm
