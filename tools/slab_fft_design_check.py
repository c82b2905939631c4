#!/usr/bin/env python
"""CPU check of the slab-decomposed FFT + fcomb planned for the sharded path.  No GPU, numpy only.

The single-GPU path transforms d = A + iB (the two interlaced grids) with one complex FFT and lets fcomb separate
A^ and B^ through the conjugate partner F(-k) (estimator.f:652-656), which under a slab decomposition lives on another
GPU.  Plan without partner traffic, for G ranks owning z-slabs of the mesh:

  1. x pass of d (local), then separate inside every x row:  A^x(kx) = (D(kx) + conj(D(-kx)))/2,  B^x(kx) = (D(kx) - conj(D(-kx)))/(2i),
     kept for kx = 0..N/2 only (A and B are real);
  2. y pass of both half-width arrays (local to the z-slab);
  3. all-to-all: z-slabs -> y-slabs (every rank gets all z of its ky range);
  4. z pass, then point-wise  delta(k) = cfac(k) * 2 (A^(k) + rec(k) B^(k)),  rec = exp(-i pi (kx+ky+kz)/N) (signed k), cfac = the
     window / normalisation of fcomb.

This script runs the plan with numpy FFTs, emulating G ranks and the all-to-all by slicing, and compares with the oracle's
fcomb_periodic / fcomb_survey output (sequential Fortran semantics, "last write wins" on the self-conjugate planes)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyspec_oracle as O          # noqa: E402  (test infrastructure: this is a design check, not product code)


def slab_plan(mesh, N, G, sumw, periodic=True):
    """mesh: float32 (2N,N,N) Fortran order as assign_quad fills it (dtl(2*ix-1 / 2*ix, iy, iz)).  Returns the half field
    (N/2+1, N, N) complex64 indexed [kx,ky,kz]."""
    h = N // 2
    A = np.ascontiguousarray(mesh[0::2].transpose(2, 1, 0))          # [z][y][x]
    B = np.ascontiguousarray(mesh[1::2].transpose(2, 1, 0))
    d = (A + 1j * B).astype(np.complex64)
    nz = N // G
    ahalf, bhalf = [], []
    for g in range(G):                                                # ---- steps 1-2 on rank g's z-slab
        s = d[g * nz:(g + 1) * nz]
        D = np.fft.ifft(s, axis=2, norm='forward').astype(np.complex64)            # sign +, unnormalised (FFTW_BACKWARD)
        Dm = np.conj(D[:, :, (-np.arange(N)) % N])
        Ax = (0.5 * (D + Dm))[:, :, :h + 1]
        Bx = (-0.5j * (D - Dm))[:, :, :h + 1]
        ahalf.append(np.fft.ifft(Ax, axis=1, norm='forward').astype(np.complex64))
        bhalf.append(np.fft.ifft(Bx, axis=1, norm='forward').astype(np.complex64))
    ny = N // G
    out = np.empty((N, N, h + 1), dtype=np.complex64)                 # [kz][ky][kx]
    k1 = np.arange(N)
    ks = np.where(k1 <= h, k1, k1 - N).astype(np.float64)            # signed frequency
    kxs = np.arange(h + 1, dtype=np.float64)

    def win(k):                                                       # (sin(pi k/N)/(pi k/N))^4 in float32 (f:632-645)
        x = (np.float32(np.pi) * np.abs(k).astype(np.float32) / np.float32(N))
        w = np.ones_like(x)
        nzm = x != 0
        w[nzm] = (np.sin(x[nzm]) / x[nzm]) ** 4
        return w.astype(np.float32)
    cf = np.float32(1. / (6. ** 3 * 4. * (sumw if periodic else 1.)))
    for q in range(G):                                                # ---- step 3: rank q gathers its ky range of every z-slab
        ay = np.concatenate([a[:, q * ny:(q + 1) * ny] for a in ahalf], axis=0)     # [z][ky in slab][kx]
        by = np.concatenate([b[:, q * ny:(q + 1) * ny] for b in bhalf], axis=0)
        Az = np.fft.ifft(ay, axis=0, norm='forward').astype(np.complex64)          # ---- step 4
        Bz = np.fft.ifft(by, axis=0, norm='forward').astype(np.complex64)
        kyq = ks[q * ny:(q + 1) * ny]
        ph = np.exp(-1j * np.pi * (ks[:, None, None] + kyq[None, :, None] + kxs[None, None, :]) / N)
        cfac = cf / (win(ks)[:, None, None] * win(kyq)[None, :, None] * win(kxs)[None, None, :])
        out[:, q * ny:(q + 1) * ny] = (cfac * 2. * (Az + ph * Bz)).astype(np.complex64)
    return out.transpose(2, 1, 0)                                     # [kx,ky,kz]


def main():
    rng = np.random.default_rng(0)
    for (N, G, periodic) in [(24, 2, True), (24, 4, True), (36, 3, True), (32, 8, False)]:
        L = 100.
        Np = 20000
        xyz = rng.uniform(0, L, (3, Np))
        xyz[:, :Np // 2] = (xyz[:, :Np // 2] * 0.2 + 40.) % L
        w = rng.uniform(0.5, 2., Np)
        mesh = O.assign_mesh(xyz, w, L, N)
        F = O._FFT(mesh, N)
        if periodic:
            O.fcomb_periodic(F, np.sum(w))
        else:
            O.fcomb_survey(F, N)
        ref = np.ascontiguousarray(F[:N // 2 + 1])
        got = slab_plan(mesh, N, G, np.sum(w), periodic)
        err = np.abs(got - ref)
        h = N // 2
        nyq = np.zeros(ref.shape, bool)                               # any Nyquist component
        nyq[h] = True
        nyq[:, h] = True
        nyq[:, :, h] = True
        scale = np.abs(ref).max()
        print('N=%d G=%d %s: max|diff|/max|ref| = %.2e away from the Nyquist planes, %.2e on them'
              % (N, G, 'periodic' if periodic else 'survey', err[~nyq].max() / scale, err[nyq].max() / scale))
        # On the Nyquist planes +h and -h are the same mode and the sequential Fortran writes such an element more than once
        # (c000 / c001 / ... and the conjugate mirror writes, f:652-665): the value that survives carries the phase of the LAST
        # write.  The plan reproduces it by rebuilding F(k) = A^ + i B^ and conj(F(-k)) = A^ - i B^ locally and calling the same
        # closed form as the single-GPU kernel (psb_fcomb_core.cuh: fcomb_value), which needs nothing but these two numbers.
        Fk_full = np.fft.ifftn((np.ascontiguousarray(mesh[0::2].transpose(2, 1, 0)) + 1j * np.ascontiguousarray(mesh[1::2].transpose(2, 1, 0))).astype(np.complex64),
                               norm='forward').astype(np.complex64).transpose(2, 1, 0)
        assert np.abs(Fk_full - O._FFT(mesh, N)).max() <= 2e-6 * np.abs(Fk_full).max()


if __name__ == '__main__':
    main()
