// psb_fcomb_core.cuh -- closed form of estimator.f:605-675 (fcomb_periodic) / 677-745 (fcomb_survey)
// for ONE half-space mode, given F(k) and F(-k) of the packed two-grid transform.
//
// The Fortran sweeps (ix,iy,iz) in [1,N/2+1]^3 and writes 8 mirror images in a fixed order
// (f:658-665); where images coincide (index 0 or N/2 on an axis) the LAST write wins.  Working
// that order through gives, for the value finally stored at half-space index (ix,iy,iz):
//
//   own(ix,iy,iz; Fk,Fm) = [ Fk (1 - cm) + conj(Fm) (1 + cm) ] * cfac ,   cm = i * X(ix) * Y~ * Z~
//     Y~ = rec[iy]            if 0 < iy < N/2      (the "+ky" image, written by c000/c001)
//        = conj(rec[min(iy,N-iy)]) otherwise       (the "-ky" image c010/c011 is written later)
//     Z~ likewise.
//   ix not in {0,N/2}                      : value = own(ix,iy,iz; F(k), F(-k))
//   ix in {0,N/2} and (iy,iz) "kept"       : value = own(...)                (written by c010/c011, re-written unchanged)
//   ix in {0,N/2} and not kept             : value = conj( own(ix,-iy,-iz; F(-k), F(k)) )   (f:662-665)
//     kept := iy > N/2  or  (iy in {0,N/2} and (iz > N/2 or iz in {0,N/2}))
//
// rec[j] is the Fortran's double-precision phase recurrence whose base was rounded to single
// (f:619-623); Wk[j] the single-precision sinc^4 window (f:630-645).  Both are 1-D
// host-built tables of length N/2+1 (capi: psb_build_fcomb_tables).
#pragma once
#include "psb_common.cuh"

namespace psb {

PSB_HD Cx<float> fcomb_own(int N, int ix, int iy, int iz, Cx<float> Fk, Cx<float> Fm,
                           const Cx<double>* rec, const float* Wk, float cf)
{
    const int h = N / 2;
    const int jy = iy <= h ? iy : N - iy, jz = iz <= h ? iz : N - iz;
    Cx<double> X = rec[ix];
    Cx<double> Y = rec[jy];
    if (!(iy > 0 && iy < h)) Y.y = -Y.y;
    Cx<double> Z = rec[jz];
    if (!(iz > 0 && iz < h)) Z.y = -Z.y;
    Cx<double> iX = mk<double>(-X.y, X.x);                 // ci*xrec
    Cx<double> ph = (iX * Y) * Z;
    Cx<float> cm = mk<float>((float)ph.x, (float)ph.y);     // rounded to single on assignment (f:647-650)
    Cx<float> a = mk<float>(1.f - cm.x, 0.f - cm.y);
    Cx<float> b = mk<float>(1.f + cm.x, 0.f + cm.y);
    Cx<float> c = Fk * a + conj(Fm) * b;
    const float cfac = cf / ((Wk[ix] * Wk[jy]) * Wk[jz]);
    return mk<float>(c.x * cfac, c.y * cfac);
}

PSB_HD bool fcomb_kept(int N, int iy, int iz)
{
    const int h = N / 2;
    const bool ys = (iy == 0 || iy == h), zs = (iz == 0 || iz == h);
    return (iy > h) || (ys && (iz > h || zs));
}

// Fk = F(ix,iy,iz), Fm = F(-ix,-iy,-iz) of the unnormalised backward transform of A + iB.
PSB_HD Cx<float> fcomb_value(int N, int ix, int iy, int iz, Cx<float> Fk, Cx<float> Fm,
                             const Cx<double>* rec, const float* Wk, float cf)
{
    const int h = N / 2;
    if ((ix == 0 || ix == h) && !fcomb_kept(N, iy, iz))
        return conj(fcomb_own(N, ix, kneg(iy, N), kneg(iz, N), Fm, Fk, rec, Wk, cf));
    return fcomb_own(N, ix, iy, iz, Fk, Fm, rec, Wk, cf);
}

// Host-side table builders (same libm calls as the Fortran would make at run time).
inline void fcomb_build_tables(int N, Cx<double>* rec, float* Wk)
{
    const float tpi = (float)6.283185307;                   // implicit REAL parameter, f:607
    const double tpiL = (double)(tpi / (float)N);
    const double piL = -tpiL / 2.0;
    const Cx<double> base = mk<double>((double)(float)cos(piL), (double)(float)sin(piL));   // cmplx() -> single, f:621
    Cx<double> r = mk<double>(1.0, 0.0);
    for (int j = 0; j <= N / 2; ++j) {
        rec[j] = r;
        r = r * base;
        float rk = (float)(tpiL * (double)j);
        float W = 1.f;
        if (rk != 0.f) { float q = sinf(rk / 2.f) / (rk / 2.f); float q2 = q * q; W = q2 * q2; }
        Wk[j] = W;
    }
}

}  // namespace psb
