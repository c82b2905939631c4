// psb_fft_core.cuh -- radix planner, small in-register DFTs and the Stockham stage used by
// every FFT pass (mesh -> delta(k): replaces the FFTW_BACKWARD call of pyspectrum.py:1073-1075;
// shell fields: replaces the per-shell FFTW forward of pyspectrum.py:397-399).
//
// Line FFT = sequence of Stockham autosort stages executed in shared memory.  A stage with
// radix R and Ns = product of the previous radices does, for butterfly j in [0, N/R):
//     k  = j mod Ns
//     v[q] = x[j + q*N/R] * W_N^{DIR * q*k*N/(Ns*R)}          q = 0..R-1
//     v    = DFT_R(v)                                          (sign DIR)
//     x'[(j/Ns)*Ns*R + k + q*Ns] = v[q]
// After all stages x' holds the DFT in natural order.  W comes from a table
// tw[i] = exp(+2*pi*i/N) computed in double on the host (pyspectrum_b200/plan.py).
#pragma once
#include "psb_common.cuh"

namespace psb {

constexpr int PSB_MAX_STAGES = 12;

struct FftPlan {
    int N;
    int nstages;
    int radix[PSB_MAX_STAGES];
    int nb_max;      // max butterflies per line over the stages (= N / min radix)
};

// Factor N into radices from {8,9,5,4,3,2}; returns false if N has another prime factor.
inline bool make_plan(int N, FftPlan* p)
{
    if (N < 2) return false;
    int e2 = 0, e3 = 0, e5 = 0, n = N;
    while (n % 2 == 0) { n /= 2; ++e2; }
    while (n % 3 == 0) { n /= 3; ++e3; }
    while (n % 5 == 0) { n /= 5; ++e5; }
    if (n != 1) return false;
    int ns = 0;
    int r[64];
    // powers of two: as many 8s as possible, never a lone radix-2 when a 4 can absorb it
    int n8 = e2 / 3, rem = e2 % 3;
    if (rem == 1 && n8 >= 1) { n8 -= 1; r[ns++] = 4; r[ns++] = 4; rem = 0; }
    for (int i = 0; i < n8; ++i) r[ns++] = 8;
    if (rem == 2) r[ns++] = 4;
    if (rem == 1) r[ns++] = 2;
    for (int i = 0; i < e3 / 2; ++i) r[ns++] = 9;
    if (e3 % 2) r[ns++] = 3;
    for (int i = 0; i < e5; ++i) r[ns++] = 5;
    if (ns > PSB_MAX_STAGES) return false;
    p->N = N;
    p->nstages = ns;
    int rmin = 1 << 30;
    // largest radices first: the early stages (small Ns) have the worst smem write strides
    for (int i = 0; i < ns; ++i)
        for (int j = i + 1; j < ns; ++j)
            if (r[j] > r[i]) { int t = r[i]; r[i] = r[j]; r[j] = t; }
    for (int i = 0; i < PSB_MAX_STAGES; ++i) p->radix[i] = (i < ns) ? r[i] : 1;
    for (int i = 0; i < ns; ++i) if (r[i] < rmin) rmin = r[i];
    p->nb_max = N / rmin;
    return true;
}

// ---------------------------------------------------------------------------------------
// small DFTs, out[p] = sum_q v[q] exp(DIR*2*pi*i*p*q/R), in place
// ---------------------------------------------------------------------------------------
template <int R, int DIR, typename T> struct Dft;

template <int DIR, typename T> struct Dft<2, DIR, T> {
    static PSB_HD void run(Cx<T>* v) { Cx<T> a = v[0], b = v[1]; v[0] = a + b; v[1] = a - b; }
};

template <int DIR, typename T> struct Dft<3, DIR, T> {
    static PSB_HD void run(Cx<T>* v) {
        const T h = (T)0.86602540378443864676;     // sqrt(3)/2
        Cx<T> s = v[1] + v[2], d = v[1] - v[2];
        Cx<T> t = mk<T>(v[0].x - (T)0.5 * s.x, v[0].y - (T)0.5 * s.y);
        Cx<T> u = mul_i<DIR>(h * d);
        v[0] = v[0] + s; v[1] = t + u; v[2] = t - u;
    }
};

template <int DIR, typename T> struct Dft<4, DIR, T> {
    static PSB_HD void run(Cx<T>* v) {
        Cx<T> a = v[0] + v[2], b = v[0] - v[2], c = v[1] + v[3], d = mul_i<DIR>(v[1] - v[3]);
        v[0] = a + c; v[1] = b + d; v[2] = a - c; v[3] = b - d;
    }
};

template <int DIR, typename T> struct Dft<5, DIR, T> {
    static PSB_HD void run(Cx<T>* v) {
        const T c1 = (T)0.30901699437494742410, c2 = (T)-0.80901699437494742410;
        const T s1 = (T)0.95105651629515357212, s2 = (T)0.58778525229247312917;
        Cx<T> t1 = v[1] + v[4], t2 = v[2] + v[3], t3 = v[1] - v[4], t4 = v[2] - v[3];
        Cx<T> m1 = mk<T>(v[0].x + c1 * t1.x + c2 * t2.x, v[0].y + c1 * t1.y + c2 * t2.y);
        Cx<T> m2 = mk<T>(v[0].x + c2 * t1.x + c1 * t2.x, v[0].y + c2 * t1.y + c1 * t2.y);
        Cx<T> n1 = mul_i<DIR>(mk<T>(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y));
        Cx<T> n2 = mul_i<DIR>(mk<T>(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y));
        v[0] = v[0] + t1 + t2;
        v[1] = m1 + n1; v[4] = m1 - n1; v[2] = m2 + n2; v[3] = m2 - n2;
    }
};

template <int DIR, typename T> struct Dft<8, DIR, T> {
    static PSB_HD void run(Cx<T>* v) {
        const T r = (T)0.70710678118654752440;
        Cx<T> e[4] = { v[0], v[2], v[4], v[6] }, o[4] = { v[1], v[3], v[5], v[7] };
        Dft<4, DIR, T>::run(e);
        Dft<4, DIR, T>::run(o);
        // W8^1 = r(1 + DIR i), W8^2 = DIR i, W8^3 = r(-1 + DIR i)
        Cx<T> o1 = r * (o[1] + mul_i<DIR>(o[1]));
        Cx<T> o2 = mul_i<DIR>(o[2]);
        Cx<T> o3 = r * (mul_i<DIR>(o[3]) - o[3]);
        v[0] = e[0] + o[0]; v[4] = e[0] - o[0];
        v[1] = e[1] + o1;   v[5] = e[1] - o1;
        v[2] = e[2] + o2;   v[6] = e[2] - o2;
        v[3] = e[3] + o3;   v[7] = e[3] - o3;
    }
};

template <int DIR, typename T> struct Dft<9, DIR, T> {
    static PSB_HD void run(Cx<T>* v) {
        // W9^1, W9^2, W9^4 = exp(DIR*2*pi*i*{1,2,4}/9)
        const T c1 = (T)0.76604444311897803520, s1 = (T)0.64278760968653932632;
        const T c2 = (T)0.17364817766693034885, s2 = (T)0.98480775301220805937;
        const T c4 = (T)-0.93969262078590838405, s4 = (T)0.34202014332566873304;
        Cx<T> a[3] = { v[0], v[3], v[6] }, b[3] = { v[1], v[4], v[7] }, c[3] = { v[2], v[5], v[8] };
        Dft<3, DIR, T>::run(a); Dft<3, DIR, T>::run(b); Dft<3, DIR, T>::run(c);
        const T sg = (T)DIR;
        b[1] = b[1] * mk<T>(c1, sg * s1); b[2] = b[2] * mk<T>(c2, sg * s2);
        c[1] = c[1] * mk<T>(c2, sg * s2); c[2] = c[2] * mk<T>(c4, sg * s4);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int p0 = 0; p0 < 3; ++p0) {
            Cx<T> f[3] = { a[p0], b[p0], c[p0] };
            Dft<3, DIR, T>::run(f);
            v[p0] = f[0]; v[p0 + 3] = f[1]; v[p0 + 6] = f[2];
        }
    }
};


// ---------------------------------------------------------------------------------------
// composite in-register DFTs (R = RA*RB, natural order in and out), built from the small ones above:
//   X[c + RA d] = sum_b W_RB^{bd} W_R^{bc} sum_a W_RA^{ac} v[RB a + b]        (all indices compile-time)
// Used for the two-stage plans (360 = 20 x 18, 256 = 16 x 16): one shared-memory exchange per line FFT.
// ---------------------------------------------------------------------------------------
template <int R> struct RootTable;
template <> struct RootTable<16> {
    static PSB_HD double c(int m) { constexpr double t[16] = { 1.00000000000000000000, 0.92387953251128673848, 0.70710678118654757274, 0.38268343236508983729, 0.00000000000000006123, -0.38268343236508972627, -0.70710678118654746172, -0.92387953251128673848, -1.00000000000000000000, -0.92387953251128684951, -0.70710678118654768376, -0.38268343236509033689, -0.00000000000000018370, 0.38268343236509000382, 0.70710678118654735069, 0.92387953251128651644 }; return t[m]; }
    static PSB_HD double s(int m) { constexpr double t[16] = { 0.00000000000000000000, 0.38268343236508978178, 0.70710678118654746172, 0.92387953251128673848, 1.00000000000000000000, 0.92387953251128673848, 0.70710678118654757274, 0.38268343236508989280, 0.00000000000000012246, -0.38268343236508967076, -0.70710678118654746172, -0.92387953251128651644, -1.00000000000000000000, -0.92387953251128662746, -0.70710678118654768376, -0.38268343236509039240 }; return t[m]; }
};
template <> struct RootTable<18> {
    static PSB_HD double c(int m) { constexpr double t[18] = { 1.00000000000000000000, 0.93969262078590842791, 0.76604444311897801345, 0.50000000000000011102, 0.17364817766693041445, -0.17364817766693030343, -0.49999999999999977796, -0.76604444311897790243, -0.93969262078590831688, -1.00000000000000000000, -0.93969262078590842791, -0.76604444311897834652, -0.50000000000000044409, -0.17364817766693033119, 0.17364817766692997036, 0.49999999999999933387, 0.76604444311897779141, 0.93969262078590842791 }; return t[m]; }
    static PSB_HD double s(int m) { constexpr double t[18] = { 0.00000000000000000000, 0.34202014332566871291, 0.64278760968653925190, 0.86602540378443859659, 0.98480775301220802032, 0.98480775301220802032, 0.86602540378443870761, 0.64278760968653947394, 0.34202014332566887944, 0.00000000000000012246, -0.34202014332566865740, -0.64278760968653891883, -0.86602540378443837454, -0.98480775301220802032, -0.98480775301220813134, -0.86602540378443904068, -0.64278760968653958496, -0.34202014332566860189 }; return t[m]; }
};
template <> struct RootTable<20> {
    static PSB_HD double c(int m) { constexpr double t[20] = { 1.00000000000000000000, 0.95105651629515353118, 0.80901699437494745126, 0.58778525229247313710, 0.30901699437494745126, 0.00000000000000006123, -0.30901699437494734024, -0.58778525229247302608, -0.80901699437494734024, -0.95105651629515353118, -1.00000000000000000000, -0.95105651629515375323, -0.80901699437494756229, -0.58778525229247324813, -0.30901699437494756229, -0.00000000000000018370, 0.30901699437494722922, 0.58778525229247291506, 0.80901699437494734024, 0.95105651629515353118 }; return t[m]; }
    static PSB_HD double s(int m) { constexpr double t[20] = { 0.00000000000000000000, 0.30901699437494739575, 0.58778525229247313710, 0.80901699437494745126, 0.95105651629515353118, 1.00000000000000000000, 0.95105651629515364220, 0.80901699437494745126, 0.58778525229247324813, 0.30901699437494750677, 0.00000000000000012246, -0.30901699437494689615, -0.58778525229247302608, -0.80901699437494734024, -0.95105651629515353118, -1.00000000000000000000, -0.95105651629515364220, -0.80901699437494756229, -0.58778525229247335915, -0.30901699437494761780 }; return t[m]; }
};
template <> struct RootTable<32> {
    static PSB_HD double c(int m) { constexpr double t[32] = { 1.00000000000000000000, 0.98078528040323043058, 0.92387953251128673848, 0.83146961230254523567, 0.70710678118654757274, 0.55557023301960228867, 0.38268343236508983729, 0.19509032201612833135, 0.00000000000000006123, -0.19509032201612819257, -0.38268343236508972627, -0.55557023301960195560, -0.70710678118654746172, -0.83146961230254534669, -0.92387953251128673848, -0.98078528040323043058, -1.00000000000000000000, -0.98078528040323043058, -0.92387953251128684951, -0.83146961230254545772, -0.70710678118654768376, -0.55557023301960217765, -0.38268343236509033689, -0.19509032201612866442, -0.00000000000000018370, 0.19509032201612830359, 0.38268343236509000382, 0.55557023301960184458, 0.70710678118654735069, 0.83146961230254523567, 0.92387953251128651644, 0.98078528040323031956 }; return t[m]; }
    static PSB_HD double s(int m) { constexpr double t[32] = { 0.00000000000000000000, 0.19509032201612824808, 0.38268343236508978178, 0.55557023301960217765, 0.70710678118654746172, 0.83146961230254523567, 0.92387953251128673848, 0.98078528040323043058, 1.00000000000000000000, 0.98078528040323043058, 0.92387953251128673848, 0.83146961230254545772, 0.70710678118654757274, 0.55557023301960217765, 0.38268343236508989280, 0.19509032201612860891, 0.00000000000000012246, -0.19509032201612835911, -0.38268343236508967076, -0.55557023301960195560, -0.70710678118654746172, -0.83146961230254523567, -0.92387953251128651644, -0.98078528040323031956, -1.00000000000000000000, -0.98078528040323043058, -0.92387953251128662746, -0.83146961230254545772, -0.70710678118654768376, -0.55557023301960217765, -0.38268343236509039240, -0.19509032201612871993 }; return t[m]; }
};

template <int RA, int RB, int DIR, typename T> struct DftComposite {
    static constexpr int R = RA * RB;
    static PSB_HD void run(Cx<T>* v) {
        Cx<T> y[RB][RA];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int b = 0; b < RB; ++b) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int a = 0; a < RA; ++a) y[b][a] = v[RB * a + b];
            Dft<RA, DIR, T>::run(y[b]);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int c = 1; c < RA; ++c) {
                if (b > 0) {
                    const int m = (b * c) % R;
                    y[b][c] = y[b][c] * mk<T>((T)RootTable<R>::c(m), (T)((double)DIR * RootTable<R>::s(m)));
                }
            }
        }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int c = 0; c < RA; ++c) {
            Cx<T> w[RB];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int b = 0; b < RB; ++b) w[b] = y[b][c];
            Dft<RB, DIR, T>::run(w);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int d = 0; d < RB; ++d) v[c + RA * d] = w[d];
        }
    }
};
template <int DIR, typename T> struct Dft<16, DIR, T> { static PSB_HD void run(Cx<T>* v) { DftComposite<4, 4, DIR, T>::run(v); } };
template <int DIR, typename T> struct Dft<32, DIR, T> { static PSB_HD void run(Cx<T>* v) { DftComposite<8, 4, DIR, T>::run(v); } };
template <int DIR, typename T> struct Dft<18, DIR, T> { static PSB_HD void run(Cx<T>* v) { DftComposite<9, 2, DIR, T>::run(v); } };
template <int DIR, typename T> struct Dft<20, DIR, T> { static PSB_HD void run(Cx<T>* v) { DftComposite<5, 4, DIR, T>::run(v); } };

// ---------------------------------------------------------------------------------------
// one butterfly of a Stockham stage, split in its read and compute+write halves so that the
// kernel can run all reads of a stage, barrier, then all writes (in-place in shared memory)
// ---------------------------------------------------------------------------------------
template <int R, typename T>
PSB_HD void stage_read(const Cx<T>* s, int estride, int N, int j, Cx<T>* v)
{
    const int M = N / R;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int q = 0; q < R; ++q) v[q] = s[(j + q * M) * estride];
}

template <int R, int DIR, typename T>
PSB_HD void stage_write(Cx<T>* s, int estride, int N, int Ns, int j, const Cx<T>* tw, Cx<T>* v)
{
    const int k = j % Ns;
    if (Ns > 1) {
        const int tstep = k * (N / (Ns * R));            // q*tstep < N
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int q = 1; q < R; ++q) {
            Cx<T> w = tw[q * tstep];
            if (DIR < 0) w.y = -w.y;
            v[q] = v[q] * w;
        }
    }
    Dft<R, DIR, T>::run(v);
    const int j0 = (j / Ns) * Ns * R + k;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int q = 0; q < R; ++q) s[(j0 + q * Ns) * estride] = v[q];
}

}  // namespace psb
