/*
 * oracle/estimator_oracle.c  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, single-threaded CPU restatement of the Fortran-77 routines that the
 * reference's hot path calls through f2py (module `estimator`):
 *
 *   oracle_assign_quad      <-  pyspectrum/estimator.f:284-512  (assign_quad)
 *   oracle_fcomb            <-  pyspectrum/estimator.f:605-675  (fcomb_periodic)
 *                               pyspectrum/estimator.f:677-745  (fcomb_survey, periodic=0)
 *   oracle_pk_pbox_rsd      <-  pyspectrum/estimator.f:155-264  (pk_pbox_rsd)
 *   oracle_quad_fields      <-  pyspectrum/estimator.f:514-603  (FiveDelta2g_1, FiveDelta2g_2, build_quad)
 *
 * The restatement follows the Fortran's *implicit typing* to the letter (every
 * undeclared a-h/o-z name is a 4-byte real, i-n a 4-byte integer) and its
 * evaluation order, because several outputs (mode counts, mu bins, the interlacing
 * phase drift) depend on IEEE rounding details.  Build with
 *   gcc -O2 -ffp-contract=off -fno-fast-math
 * so that no multiply-add is fused (x86-64 gfortran without -mfma does not fuse).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product path (pyspectrum_b200) never does.
 *
 * PARITY STATUS: gfortran is not available in the build container, so this file
 * could not be diffed against a compiled estimator.f.  It is pinned instead by the
 * known-answer tests in tests/test_oracle_*.py (mesh mass = 216*sum(w) per grid,
 * single-particle plane wave, delta(k=0)=1, Hermitian pairing on self-conjugate
 * planes), by a closed-form direct-summation delta(k) (tests/direct_sum.py), by SECOND
 * restatements written independently from the Fortran text that agree with this file
 * bit for bit (tests/assign_quad_numpy.py, tests/fcomb_numpy.py) or to 1e-12
 * (tests/pk_pbox_rsd_numpy.py) and, through the Python layer, by the reference's shipped
 * triangle-count files.  "Parity unpinned" against a compiled estimator.f still holds.
 * See DESIGN.md "Oracle".
 *
 * Array conventions: all arrays are Fortran (column-major) as f2py passes them.
 *   r   : float  (3,Np)              -> r[3*i + a]
 *   dtl : float  (2*Ngrid,Ngrid,Ngrid)-> dtl[row + 2*Ngrid*(iy + Ngrid*iz)], row=2*ix(+1)
 *   dcl : complex(Ngrid,Ngrid,Ngrid) -> dcl[ix + Ngrid*(iy + Ngrid*iz)]  (re,im floats)
 */
#include <math.h>
#include <complex.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* x**3 and x**4 as gfortran expands integer powers: (x*x)*x and (x*x)*(x*x). */
static inline float pow3f(float x) { float x2 = x * x; return x2 * x; }
static inline float pow4f(float x) { float x2 = x * x; return x2 * x2; }

/* 0-based wrap helpers mirroring mod(.,Ngrid) for non-negative arguments. */
static inline int imod(int a, int n) { return a % n; }

/* The four piecewise-cubic weights of estimator.f:316-320 (and 352-356 etc.).
 * h in [0,1): wm2 -> cell c-1, wm1 -> cell c, wp1 -> cell c+1, wp2 -> cell c+2. */
static inline void pcs_weights(float h, float *wm2, float *wm1, float *wp1, float *wp2)
{
    float h2 = h * h;
    *wm2 = pow3f(1.f - h);
    *wm1 = 4.f + (3.f * h - 6.f) * h2;
    *wp2 = h2 * h;
    *wp1 = 6.f - *wm2 - *wm1 - *wp2;
}

/* estimator.f:284-512.  ia..id select the weight variant (f:292-300); the hot path
 * always passes 0,0,0,0 but the Q_ij / Q_ijkl branches are restated for completeness. */
void oracle_assign_quad(const float *r, const float *w, float *dtl,
                        int64_t Np, int Ngrid, float kf_ks, float offset,
                        int ia, int ib, int ic, int id)
{
    const int N = Ngrid;
    const int64_t ldx = 2 * (int64_t)N;          /* rows per (iy,iz) column */
    for (int64_t i = 0; i < Np; ++i) {
        const float *ri = r + 3 * i;
        float we;
        if (ia == 0 && ib == 0 && ic == 0 && id == 0) {
            we = w[i];
        } else if (ic == 0 && id == 0) {
            float rnorm = ri[0] * ri[0] + ri[1] * ri[1] + ri[2] * ri[2];
            we = w[i] * ri[ia - 1] * ri[ib - 1] / rnorm;
        } else {
            float rnorm = ri[0] * ri[0] + ri[1] * ri[1] + ri[2] * ri[2];
            /* rnorm**2 is a single-precision square */
            we = w[i] * ri[ia - 1] * ri[ib - 1] * ri[ic - 1] * ri[id - 1] / (rnorm * rnorm);
        }

        /* f:302-307: 1-based continuous coordinates on grid A (rx) and the
         * half-cell shifted grid B (tx). */
        float rp[3], tp[3];
        for (int a = 0; a < 3; ++a) {
            rp[a] = kf_ks * ri[a] + 1.f + offset;
            tp[a] = rp[a] + 0.5f;
        }
        /* per axis: 0-based cell indices of the 4 stencil points and their weights */
        int   ca[3][4], cb[3][4];
        float ha[3][4], hb[3][4];
        for (int a = 0; a < 3; ++a) {
            /* grid A, f:308-340 */
            int im1 = (int)rp[a];                         /* 1-based cell */
            float h = rp[a] - (float)im1;
            ca[a][0] = imod(im1 - 2 + N, N);              /* m2 */
            ca[a][1] = im1 - 1;                           /* m1 (no wrap in the Fortran) */
            ca[a][2] = imod(im1, N);                      /* p1 */
            ca[a][3] = imod(im1 + 1, N);                  /* p2 */
            pcs_weights(h, &ha[a][0], &ha[a][1], &ha[a][2], &ha[a][3]);
            /* grid B, f:342-378 */
            int nm1 = (int)tp[a];
            float g = tp[a] - (float)nm1;
            nm1 = imod(nm1 - 1, N) + 1;                   /* wrap cell N+1 -> 1 */
            cb[a][0] = imod(nm1 - 2 + N, N);
            cb[a][1] = nm1 - 1;
            cb[a][2] = imod(nm1, N);
            cb[a][3] = imod(nm1 + 1, N);
            pcs_weights(g, &hb[a][0], &hb[a][1], &hb[a][2], &hb[a][3]);
        }
        /* f:380-508: z outer, y middle, x inner; product order ((hx*hy)*hz)*we */
        for (int kz = 0; kz < 4; ++kz)
            for (int ky = 0; ky < 4; ++ky)
                for (int kx = 0; kx < 4; ++kx) {
                    int64_t col = (int64_t)ca[1][ky] + (int64_t)N * ca[2][kz];
                    dtl[2 * ca[0][kx] + ldx * col] += ha[0][kx] * ha[1][ky] * ha[2][kz] * we;
                }
        for (int kz = 0; kz < 4; ++kz)
            for (int ky = 0; ky < 4; ++ky)
                for (int kx = 0; kx < 4; ++kx) {
                    int64_t col = (int64_t)cb[1][ky] + (int64_t)N * cb[2][kz];
                    dtl[2 * cb[0][kx] + 1 + ldx * col] += hb[0][kx] * hb[1][ky] * hb[2][kz] * we;
                }
    }
}

typedef struct { float re, im; } cf32;
typedef struct { double re, im; } cf64;

static inline cf64 zmul(cf64 a, cf64 b) { cf64 c = { a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re }; return c; }
static inline cf64 zconj(cf64 a) { cf64 c = { a.re, -a.im }; return c; }
static inline cf32 cmulf_(cf32 a, cf32 b) { cf32 c = { a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re }; return c; }
static inline cf32 caddf_(cf32 a, cf32 b) { cf32 c = { a.re + b.re, a.im + b.im }; return c; }
static inline cf32 csubf_(cf32 a, cf32 b) { cf32 c = { a.re - b.re, a.im - b.im }; return c; }
static inline cf32 cconjf_(cf32 a) { cf32 c = { a.re, -a.im }; return c; }
static inline cf32 z2c(cf64 a) { cf32 c = { (float)a.re, (float)a.im }; return c; }

/* estimator.f:605-675 (periodic != 0: cf carries 1/N) and f:677-745 (periodic == 0).
 * Sequential, in place, including the overwrite order on self-conjugate planes. */
void oracle_fcomb(float *dcl_, float Nsum, int Ngrid, int periodic)
{
    cf32 *dcl = (cf32 *)dcl_;
    const int N = Ngrid;
    const float tpi = (float)6.283185307;          /* implicit REAL parameter (f:607) */
    float cf;
    if (periodic) cf = 1.f / (pow3f(6.f) * 4.f * Nsum);   /* f:615 */
    else          cf = 1.f / (pow3f(6.f) * 4.f);          /* f:686 */
    const int Lnyq = N / 2 + 1;
    const double tpiL = (double)(tpi / (float)N);  /* single division, widened (f:619) */
    const double piL = -tpiL / 2.0;
    /* cmplx() without KIND returns default (single) complex: f:621-623 */
    cf64 rec = { (double)(float)cos(piL), (double)(float)sin(piL) };
    const cf32 c1 = { 1.f, 0.f };
    const cf64 ci = { 0.0, 1.0 };

#define D(ix, iy, iz) dcl[(int64_t)(ix) + (int64_t)N * ((int64_t)(iy) + (int64_t)N * (iz))]
    cf64 zrec = { 1.0, 0.0 };
    for (int iz = 0; iz < Lnyq; ++iz) {
        int icz = (N - iz) % N;                     /* 0-based of mod(Ngrid-iz+1,Ngrid)+1 */
        float rkz = (float)(tpiL * (double)iz);
        float Wkz = 1.f;
        if (rkz != 0.f) Wkz = pow4f(sinf(rkz / 2.f) / (rkz / 2.f));
        cf64 yrec = { 1.0, 0.0 };
        for (int iy = 0; iy < Lnyq; ++iy) {
            int icy = (N - iy) % N;
            float rky = (float)(tpiL * (double)iy);
            float Wky = 1.f;
            if (rky != 0.f) Wky = pow4f(sinf(rky / 2.f) / (rky / 2.f));
            cf64 xrec = { 1.0, 0.0 };
            for (int ix = 0; ix < Lnyq; ++ix) {
                int icx = (N - ix) % N;
                float rkx = (float)(tpiL * (double)ix);
                float Wkx = 1.f;
                if (rkx != 0.f) Wkx = pow4f(sinf(rkx / 2.f) / (rkx / 2.f));
                float cfac = cf / (Wkx * Wky * Wkz);

                cf64 cix = zmul(ci, xrec);
                cf32 cma = z2c(zmul(zmul(cix, yrec), zrec));
                cf32 cmb = z2c(zmul(zmul(cix, yrec), zconj(zrec)));
                cf32 cmc = z2c(zmul(zmul(cix, zconj(yrec)), zrec));
                cf32 cmd = z2c(zmul(cix, zconj(zmul(yrec, zrec))));

                cf32 c000 = caddf_(cmulf_(D(ix, iy, iz),   csubf_(c1, cma)), cmulf_(cconjf_(D(icx, icy, icz)), caddf_(c1, cma)));
                cf32 c001 = caddf_(cmulf_(D(ix, iy, icz),  csubf_(c1, cmb)), cmulf_(cconjf_(D(icx, icy, iz)),  caddf_(c1, cmb)));
                cf32 c010 = caddf_(cmulf_(D(ix, icy, iz),  csubf_(c1, cmc)), cmulf_(cconjf_(D(icx, iy, icz)),  caddf_(c1, cmc)));
                cf32 c011 = caddf_(cmulf_(D(ix, icy, icz), csubf_(c1, cmd)), cmulf_(cconjf_(D(icx, iy, iz)),   caddf_(c1, cmd)));

                D(ix, iy, iz).re   = c000.re * cfac; D(ix, iy, iz).im   = c000.im * cfac;
                D(ix, iy, icz).re  = c001.re * cfac; D(ix, iy, icz).im  = c001.im * cfac;
                D(ix, icy, iz).re  = c010.re * cfac; D(ix, icy, iz).im  = c010.im * cfac;
                D(ix, icy, icz).re = c011.re * cfac; D(ix, icy, icz).im = c011.im * cfac;
                D(icx, iy, iz)   = cconjf_(D(ix, icy, icz));
                D(icx, iy, icz)  = cconjf_(D(ix, icy, iz));
                D(icx, icy, iz)  = cconjf_(D(ix, iy, icz));
                D(icx, icy, icz) = cconjf_(D(ix, iy, iz));

                xrec = zmul(xrec, rec);
            }
            yrec = zmul(yrec, rec);
        }
        zrec = zmul(zrec, rec);
    }
#undef D
}

/* estimator.f:155-264.  dtl is complex (Ngrid/2+1,Ngrid,Ngrid) Fortran order.
 * Outputs are Fortran order: k,p0,p2,p4,nk (Nbin); km,mk,pkm,nkm (Nbin,Nmu). */
void oracle_pk_pbox_rsd(const float *dtl_, double *k, double *p0, double *p2, double *p4,
                        double *nk, double *km, double *mk, double *pkm, double *nkm,
                        int irsd, int Lbox, int Nbin, int Nmu, int Ngrid)
{
    const cf32 *dtl = (const cf32 *)dtl_;
    const int N = Ngrid, Nh = N / 2 + 1;
    const float pi = 3.141592654f, tpi = 2.f * pi;
    const float kf = tpi / (float)Lbox;
    const float mubin = 1.f / (float)Nmu;
    float thetaobs = 0.f, phiobs = 0.f;
    if (irsd == 0)      { thetaobs = 0.5f * pi; phiobs = 0.f; }
    else if (irsd == 1) { thetaobs = 0.5f * pi; phiobs = 0.5f * pi; }
    else if (irsd == 2) { thetaobs = 0.f;       phiobs = 0.f; }
    const float sinph = sinf(phiobs), cosph = cosf(phiobs);
    const float sinth = sinf(thetaobs), costh = cosf(thetaobs);

    for (int i = 0; i < Nbin; ++i) {
        k[i] = p0[i] = p2[i] = p4[i] = nk[i] = 0.0;
        for (int j = 0; j < Nmu; ++j) {
            int64_t o = i + (int64_t)Nbin * j;
            km[o] = mk[o] = pkm[o] = nkm[o] = 0.0;
        }
    }
    for (int iz = 1; iz <= N; ++iz) {
        int icz = (N + 1 - iz) % N + 1;
        float rkz = (float)((iz + N / 2 - 2) % N - N / 2 + 1);
        for (int iy = 1; iy <= N; ++iy) {
            int icy = (N + 1 - iy) % N + 1;
            float rky = (float)((iy + N / 2 - 2) % N - N / 2 + 1);
            for (int ix = 1; ix <= N; ++ix) {
                int icx = (N + 1 - ix) % N + 1;
                float rkx = (float)((ix + N / 2 - 2) % N - N / 2 + 1);
                float rk = sqrtf(rkx * rkx + rky * rky + rkz * rkz);
                int imk = (int)lroundf((float)Nbin * rk / (float)(N / 2));
                if (imk <= Nbin && imk != 0) {
                    float cot1 = rkz / rk;
                    float sit1 = sqrtf(1.f - cot1 * cot1);
                    float cc;
                    if (sit1 > 0.f) {
                        float cp = rkx / (rk * sit1);
                        float sp = rky / (rk * sit1);
                        cc = sinph * sp + cosph * cp;
                    } else {
                        cc = 0.f;
                    }
                    double mu = (double)(costh * cot1 + sinth * sit1 * cc);
                    int imu = (int)((fabs(mu) + (double)mubin) / (double)mubin);
                    double mu2 = mu * mu;
                    double Le2 = -5.e-1 + 1.5e0 * mu2;
                    double Le4 = 3.75e-1 - 3.75e0 * mu2 + 4.375e0 * (mu2 * mu2);
                    nk[imk - 1] += 1.0;
                    cf32 ct;
                    if (ix <= Nh) ct = dtl[(ix - 1) + (int64_t)Nh * ((iy - 1) + (int64_t)N * (iz - 1))];
                    else          ct = dtl[(icx - 1) + (int64_t)Nh * ((icy - 1) + (int64_t)N * (icz - 1))];
                    float ab = hypotf(ct.re, ct.im);      /* cabs() */
                    float pk = ab * ab;
                    k[imk - 1]  += (double)(kf * rk);
                    p0[imk - 1] += (double)pk;
                    p2[imk - 1] += (double)pk * 5.e0 * Le2;
                    p4[imk - 1] += (double)pk * 9.e0 * Le4;
                    if (imu <= Nmu && imu > 0) {
                        int64_t o = (imk - 1) + (int64_t)Nbin * (imu - 1);
                        nkm[o] += 1.0;
                        km[o]  += (double)(kf * rk);
                        mk[o]  += fabs(mu);
                        pkm[o] += (double)pk;
                    }
                }
            }
        }
    }
    const double kf3 = (double)(kf * kf * kf);
    for (int i = 0; i < Nbin; ++i) {
        if (nk[i] > 0) {
            k[i]  = k[i] / nk[i];
            p0[i] = p0[i] / nk[i] / kf3;
            p2[i] = p2[i] / nk[i] / kf3;
            p4[i] = p4[i] / nk[i] / kf3;
        }
    }
    for (int i = 0; i < Nbin; ++i)
        for (int j = 0; j < Nmu; ++j) {
            int64_t o = i + (int64_t)Nbin * j;
            if (nkm[o] > 0) {
                km[o]  = km[o] / nkm[o];
                mk[o]  = mk[o] / nkm[o];
                pkm[o] = pkm[o] / nkm[o] / kf3;
            }
        }
}

/* Streaming helper for the Python oracle's triangle loop (pyspectrum.py:427-430):
 * sum_x a(x) b(x) c(x) with float32-valued inputs widened to double, serial order.
 * (np.einsum('i,i,i') on float64 views does the same arithmetic; this avoids the
 * reference's 8-byte-per-cell shell store so the oracle fits small hosts.) */
double oracle_triple_sum_f32(const float *a, const float *b, const float *c, int64_t n)
{
    /* 8 independent partial sums: mirrors a SIMD-reassociated double accumulation */
    double s[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
    int64_t i = 0;
    for (; i + 8 <= n; i += 8)
        for (int j = 0; j < 8; ++j)
            s[j] += (double)a[i + j] * (double)b[i + j] * (double)c[i + j];
    double t = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
    for (; i < n; ++i) t += (double)a[i] * (double)b[i] * (double)c[i];
    return t;
}


/* ---------------------------------------------------------------------------------------------
 * estimator.f:514-603: FiveDelta2g_1, FiveDelta2g_2, build_quad on (Ngrid/2+1,Ngrid,Ngrid) complex arrays (x fastest).
 * The file has no `implicit none`: kxh, kyh, kzh are undeclared, start with `k`, and are therefore INTEGER -- the quotient
 * float(ikx)/rk is truncated on assignment (f:531-533, 563-565).  `amu` in build_quad is declared real (f:578).
 * mode 1: out = dcgxx, a = dcgyy, b = dcgzz;  mode 2: out = dcgxx, a = dcg, b = dcgxy, c = dcgyz, d = dcgzx;
 * mode 3: out = dclr2, a = dclr1, irsd in 1..3 (anything else: the Fortran stops; here: nothing is written). */
static inline int signed_k(int i0, int N) { return (i0 + N / 2 - 1) % N - N / 2 + 1; }      /* f:524 with i = i0 + 1 */

void oracle_quad_fields(int mode, const float *a_, const float *b_, const float *c_, const float *d_, float *out_, int N, int irsd)
{
    const cf32 *a = (const cf32 *)a_, *b = (const cf32 *)b_, *c = (const cf32 *)c_, *d = (const cf32 *)d_;
    cf32 *out = (cf32 *)out_;
    const int hx = N / 2 + 1;
    if (mode == 3 && (irsd < 1 || irsd > 3)) return;
    for (int iz = 0; iz < N; ++iz)
        for (int iy = 0; iy < N; ++iy)
            for (int ix = 0; ix < hx; ++ix) {
                const int64_t e = ((int64_t)iz * N + iy) * hx + ix;
                const int ikx = signed_k(ix, N), iky = signed_k(iy, N), ikz = signed_k(iz, N);
                const float rk = sqrtf((float)(ikx * ikx + iky * iky + ikz * ikz));
                if (!(rk > 0.f)) continue;
                if (mode == 3) {
                    const float amu = (float)(irsd == 3 ? ikz : (irsd == 2 ? iky : ikx)) / rk;
                    const float amu2 = amu * amu;
                    const float t = 7.5f * amu2;
                    const float fac = t - 2.5f;
                    out[e].re = fac * a[e].re; out[e].im = fac * a[e].im;
                    continue;
                }
                const int kxh = (int)((float)ikx / rk), kyh = (int)((float)iky / rk), kzh = (int)((float)ikz / rk);
                if (mode == 1) {
                    const float sx = (float)(kxh * kxh), sy = (float)(kyh * kyh), sz = (float)(kzh * kzh);
                    cf32 s = { out[e].re * sx, out[e].im * sx };
                    cf32 t = { a[e].re * sy, a[e].im * sy };
                    s = caddf_(s, t);
                    t.re = b[e].re * sz; t.im = b[e].im * sz;
                    s = caddf_(s, t);
                    out[e].re = 7.5f * s.re; out[e].im = 7.5f * s.im;
                } else {
                    const cf32 *src[3] = { b, c, d };
                    const int f1[3] = { kxh, kyh, kzh }, f2[3] = { kyh, kzh, kxh };
                    cf32 s = { 0.f, 0.f };
                    for (int q = 0; q < 3; ++q) {
                        cf32 t = { 2.f * src[q][e].re, 2.f * src[q][e].im };
                        t.re = t.re * (float)f1[q]; t.im = t.im * (float)f1[q];
                        t.re = t.re * (float)f2[q]; t.im = t.im * (float)f2[q];
                        s = q == 0 ? t : caddf_(s, t);
                    }
                    cf32 u = { 7.5f * s.re, 7.5f * s.im };
                    u = caddf_(out[e], u);
                    cf32 v = { 2.5f * a[e].re, 2.5f * a[e].im };
                    out[e] = csubf_(u, v);
                }
            }
}
