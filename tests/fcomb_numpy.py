"""A second, independent restatement of estimator.f:605-745 (fcomb_periodic / fcomb_survey), written straight from the Fortran text
as a sequential Python loop over numpy scalars -- not derived from oracle/estimator_oracle.c.  It keeps the Fortran's implicit typing
(cf, rk*, Wk*, cfac REAL; the phase recurrences COMPLEX*16 rounded to COMPLEX on assignment to cma..cmd), its in-place update order
(later iterations read what earlier ones wrote on the self-conjugate planes) and single-precision complex arithmetic without
contraction.  Small grids only (pure Python).
(The first draft of this file read two typing rules wrongly -- `parameter(tpi=6.283185307d0)` makes an implicitly REAL constant, and
CMPLX() of two doubles is default-kind complex -- and disagreed with the C restatement at the 1e-7 level everywhere; with the rules
applied as the Fortran standard states them the two agree bit for bit on almost every element.)"""
import math

import numpy as np

f32, c64 = np.float32, np.complex64


def _cmul(a, b):
    """COMPLEX*COMPLEX in single precision: four products, one subtraction, one addition, each rounded."""
    ar, ai, br, bi = f32(a.real), f32(a.imag), f32(b.real), f32(b.imag)
    return c64(complex(f32(f32(ar * br) - f32(ai * bi)), f32(f32(ar * bi) + f32(ai * br))))


def _cadd(a, b):
    return c64(complex(f32(f32(a.real) + f32(b.real)), f32(f32(a.imag) + f32(b.imag))))


def _csub(a, b):
    return c64(complex(f32(f32(a.real) - f32(b.real)), f32(f32(a.imag) - f32(b.imag))))


def _window(rk):
    """Wk = (sin(rk/2)/(rk/2))**4 in single precision (rk REAL)."""
    if rk == 0:
        return f32(1.)
    hx = f32(rk / f32(2.))
    r = f32(f32(math.sin(float(hx))) / hx)
    r2 = f32(r * r)
    return f32(r2 * r2)


def fcomb(dcl, Nsum=None):
    """In place on dcl: complex64 (Ngrid,Ngrid,Ngrid) indexed [ix,iy,iz].  Nsum: the REAL argument N of fcomb_periodic; None = fcomb_survey."""
    N = dcl.shape[0]
    cf = f32(f32(1.) / f32(f32(f32(216.) * f32(4.)) * f32(Nsum))) if Nsum is not None else f32(f32(1.) / f32(f32(216.) * f32(4.)))
    Lnyq = N // 2 + 1
    tpi = f32(6.283185307)                                          # PARAMETER without a type statement: implicit REAL, the d0 constant is rounded
    tpiL = float(f32(tpi / f32(N)))                                  # single division, then widened into the real*8 variable
    piL = -tpiL / 2.
    rec = complex(f32(math.cos(piL)), f32(math.sin(piL)))            # CMPLX() without KIND returns DEFAULT complex: rounded, then stored in complex*16
    c1, ci = c64(1.), c64(1j)
    ic = lambda i: (N - i + 1) % N + 1                               # 1-based
    zrec = complex(1., 0.)
    for iz in range(1, Lnyq + 1):
        icz = ic(iz)
        Wkz = _window(f32(tpiL * (iz - 1)))
        yrec = complex(1., 0.)
        for iy in range(1, Lnyq + 1):
            icy = ic(iy)
            Wky = _window(f32(tpiL * (iy - 1)))
            xrec = complex(1., 0.)
            for ix in range(1, Lnyq + 1):
                icx = ic(ix)
                Wkx = _window(f32(tpiL * (ix - 1)))
                cfac = f32(cf / f32(f32(Wkx * Wky) * Wkz))
                t = complex(ci) * xrec                               # complex*16 products, rounded when stored in cma..cmd
                cma = c64(t * yrec * zrec)
                cmb = c64(t * yrec * zrec.conjugate())
                cmc = c64(t * yrec.conjugate() * zrec)
                cmd = c64(t * (yrec * zrec).conjugate())
                x, y, z, cx, cy, cz = ix - 1, iy - 1, iz - 1, icx - 1, icy - 1, icz - 1
                c000 = _cadd(_cmul(dcl[x, y, z], _csub(c1, cma)), _cmul(np.conj(dcl[cx, cy, cz]), _cadd(c1, cma)))
                c001 = _cadd(_cmul(dcl[x, y, cz], _csub(c1, cmb)), _cmul(np.conj(dcl[cx, cy, z]), _cadd(c1, cmb)))
                c010 = _cadd(_cmul(dcl[x, cy, z], _csub(c1, cmc)), _cmul(np.conj(dcl[cx, y, cz]), _cadd(c1, cmc)))
                c011 = _cadd(_cmul(dcl[x, cy, cz], _csub(c1, cmd)), _cmul(np.conj(dcl[cx, y, z]), _cadd(c1, cmd)))
                sc = lambda v: c64(complex(f32(f32(v.real) * cfac), f32(f32(v.imag) * cfac)))
                dcl[x, y, z] = sc(c000)
                dcl[x, y, cz] = sc(c001)
                dcl[x, cy, z] = sc(c010)
                dcl[x, cy, cz] = sc(c011)
                dcl[cx, y, z] = np.conj(dcl[x, cy, cz])
                dcl[cx, y, cz] = np.conj(dcl[x, cy, z])
                dcl[cx, cy, z] = np.conj(dcl[x, y, cz])
                dcl[cx, cy, cz] = np.conj(dcl[x, y, z])
                xrec = xrec * rec
            yrec = yrec * rec
        zrec = zrec * rec
    return dcl
