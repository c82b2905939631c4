// psb_fft_lines.cuh -- hand-written 3-D FFT passes for sm_100a (kernel + IO functors + dispatch).
//
// One generic kernel, fft_lines_kernel<T,DIR,MAXB,MAXT,IO>: a CTA stages LPC lines of length N
// in shared memory as s[idx][line] (line fastest, row padded to LPC+1 -> conflict-free for the
// line-fastest thread mapping and a 2-way worst case on the transposing loads), runs the
// Stockham stages of psb_fft_core.cuh in place (read all / barrier / write all), and hands the
// result to an IO functor that owns the global-memory side (coalesced 64-128 B segments).
//
// IO functors:
//   IoRows     contiguous lines, in place                     mesh pass 1 (x)
//   IoCols     strided lines, optional input pruning,         mesh pass 2 (y); shell passes 2 (y), 3 (z)
//              optional planar-real output (+ sum of squares)
//   IoZFcomb   z lines of a kx tile and of its mirror tile,   mesh pass 3 (z) fused with
//              epilogue = fcomb closed form -> half field      estimator.f:605-675
//   IoShellX   builds the packed shell pair  Z = D 1[a] + i D 1[b]  from the half field on load
//              (Hermitian completion + reflect_delta's realification, pyspectrum.py:1134-1157,
//              shell mask pyspectrum.py:378,393), pruned to |k| <= R              shell pass 1 (x)
//
// Replaces: pyfftw ifftn (pyspectrum.py:1073-1075), fcomb_periodic (estimator.f:605-675),
//           per-shell pyfftw fftn + np.real (pyspectrum.py:387-400), reflect_delta (py:1134-1157).
#pragma once
#include <cuda_runtime.h>
#include <utility>
#include <type_traits>
#include "psb_fft_core.cuh"
#include "psb_fcomb_core.cuh"
#include <cuda_fp16.h>
#include "psb_kernels.h"

namespace psb {

template <typename T, int R, int DIR, int MAXB>
__device__ __forceinline__ void run_stage(Cx<T>* sl, int estride, int N, int Ns, int t, int TPL, const Cx<T>* tw)
{
    Cx<T> v[MAXB][R];
    const int M = N / R;
#pragma unroll
    for (int b = 0; b < MAXB; ++b) {
        const int j = t + b * TPL;
        if (j < M) stage_read<R, T>(sl, estride, N, j, v[b]);
    }
    __syncthreads();
#pragma unroll
    for (int b = 0; b < MAXB; ++b) {
        const int j = t + b * TPL;
        if (j < M) stage_write<R, DIR, T>(sl, estride, N, Ns, j, tw, v[b]);
    }
    __syncthreads();
}

// Fused edge stages: the first stage takes its inputs straight from the IO functor (global memory -> registers: no staging
// pass through shared memory, and all loads of a thread are in flight together), the last one hands its outputs to the
// functor's epilogue from registers.  Only the exchanges between stages go through shared memory.
template <typename T, int R, int DIR, int MAXB, int LPC, class IO>
__device__ __forceinline__ void first_stage_fused(Cx<T>* sl, int estride, int N, int t, int TPL, const IO& io, const typename IO::Acc& acc)
{
    Cx<T> v[MAXB][R];
    const int M = N / R;
#pragma unroll
    for (int b = 0; b < MAXB; ++b) {
        const int j = t + b * TPL;
        if (j < M) {
#pragma unroll
            for (int q = 0; q < R; ++q) v[b][q] = io.fetch(j + q * M, N, acc);
        }
    }
#pragma unroll
    for (int b = 0; b < MAXB; ++b) {
        const int j = t + b * TPL;
        if (j < M) {
            Dft<R, DIR, T>::run(v[b]);                     // Ns = 1: no twiddles, outputs at j*R + q
#pragma unroll
            for (int q = 0; q < R; ++q) sl[(j * R + q) * estride] = v[b][q];
        }
    }
}

template <typename T, int R, int DIR, int MAXB, int LPC, class IO>
__device__ __forceinline__ void last_stage_fused(const Cx<T>* sl, int estride, int N, int t, int TPL, const Cx<T>* tw, const IO& io,
                                                 int line, typename IO::Acc& acc)
{
    const int M = N / R;                                   // == Ns of the last stage
#pragma unroll
    for (int b = 0; b < MAXB; ++b) {
        const int j = t + b * TPL;
        const bool act = j < M;
        Cx<T> v[R];
        if (act) {
            stage_read<R, T>(sl, estride, N, j, v);
#pragma unroll
            for (int q = 1; q < R; ++q) {
                Cx<T> w = tw[q * j];
                if (DIR < 0) w.y = -w.y;
                v[q] = v[q] * w;
            }
            Dft<R, DIR, T>::run(v);
        } else {
#pragma unroll
            for (int q = 0; q < R; ++q) v[q] = mk<T>(0, 0);
        }
#pragma unroll
        for (int q = 0; q < R; ++q) io.emit(j + q * M, line, v[q], act, acc);     // every lane calls (warp shuffles inside)
    }
}

// Stage policies: StaticStages<RS...> expands the radix sequence at compile time (straight-line
// stages, N a compile-time constant); DynStages walks a runtime plan with a switch.  The switch form
// makes ptxas speculate the shared-memory reads of every case (245 registers unconstrained), so it is
// only the fallback for grid sizes without a compiled plan and runs with <= 256 threads per CTA.
template <int... RS> struct StaticStages {
    static constexpr int N = (RS * ...);
    static constexpr int NSTAGES = sizeof...(RS);
    static constexpr int min_radix() { int m = 1 << 30; int r[] = { RS... }; for (int i = 0; i < NSTAGES; ++i) if (r[i] < m) m = r[i]; return m; }
    static constexpr int NB_MAX = N / min_radix();
    static constexpr bool IS_STATIC = true;
    template <typename T, int DIR, int MAXB>
    static __device__ __forceinline__ void run(const FftPlan&, Cx<T>* sl, int estride, int t, int TPL, const Cx<T>* tw) {
        int Ns = 1;
        ((run_stage<T, RS, DIR, MAXB>(sl, estride, N, Ns, t, TPL, tw), Ns *= RS), ...);
    }
    // fused variant: stage 0 reads through io.fetch, the last stage writes through io.emit
    template <typename T, int DIR, int MAXB, int LPC, class IO, int I, int R>
    static __device__ __forceinline__ void fused_stage(Cx<T>* sl, int estride, int& Ns, int t, int TPL, const Cx<T>* tw, const IO& io,
                                                       int line, typename IO::Acc& acc) {
        if constexpr (I == 0) {
            first_stage_fused<T, R, DIR, MAXB, LPC, IO>(sl, estride, N, t, TPL, io, acc);
            __syncthreads();
        } else if constexpr (I == NSTAGES - 1) {
            last_stage_fused<T, R, DIR, MAXB, LPC, IO>(sl, estride, N, t, TPL, tw, io, line, acc);
        } else {
            run_stage<T, R, DIR, MAXB>(sl, estride, N, Ns, t, TPL, tw);
        }
        Ns *= R;
    }
    template <typename T, int DIR, int MAXB, int LPC, class IO, int... I>
    static __device__ __forceinline__ void run_fused_impl(Cx<T>* sl, int estride, int t, int TPL, const Cx<T>* tw, const IO& io, int line,
                                                          typename IO::Acc& acc, std::integer_sequence<int, I...>) {
        int Ns = 1;
        (fused_stage<T, DIR, MAXB, LPC, IO, I, RS>(sl, estride, Ns, t, TPL, tw, io, line, acc), ...);
    }
    template <typename T, int DIR, int MAXB, int LPC, class IO>
    static __device__ __forceinline__ void run_fused(Cx<T>* sl, int estride, int t, int TPL, const Cx<T>* tw, const IO& io, int line,
                                                     typename IO::Acc& acc) {
        static_assert(NSTAGES >= 2, "fused edge stages need at least two stages");
        run_fused_impl<T, DIR, MAXB, LPC, IO>(sl, estride, t, TPL, tw, io, line, acc, std::make_integer_sequence<int, NSTAGES>{});
    }
};

struct DynStages {
    static constexpr bool IS_STATIC = false;
    static constexpr int N = 0;
    template <typename T, int DIR, int MAXB>
    static __device__ __forceinline__ void run(const FftPlan& plan, Cx<T>* sl, int estride, int t, int TPL, const Cx<T>* tw) {
        const int N = plan.N;
        int Ns = 1;
        for (int st = 0; st < plan.nstages; ++st) {
            const int R = plan.radix[st];
            switch (R) {
                case 2: run_stage<T, 2, DIR, MAXB>(sl, estride, N, Ns, t, TPL, tw); break;
                case 3: run_stage<T, 3, DIR, MAXB>(sl, estride, N, Ns, t, TPL, tw); break;
                case 4: run_stage<T, 4, DIR, MAXB>(sl, estride, N, Ns, t, TPL, tw); break;
                case 5: run_stage<T, 5, DIR, MAXB>(sl, estride, N, Ns, t, TPL, tw); break;
                case 8: run_stage<T, 8, DIR, MAXB>(sl, estride, N, Ns, t, TPL, tw); break;
                default: run_stage<T, 9, DIR, MAXB>(sl, estride, N, Ns, t, TPL, tw); break;
            }
            Ns *= R;
        }
    }
};

template <typename T, int DIR, int MAXB, int NTHR, int MINB, int LPC_T, class STAGES, class IO>
__global__ void __launch_bounds__(NTHR, MINB) fft_lines_kernel(FftPlan plan, int lpc_rt, int TPL, const Cx<T>* __restrict__ tw_g, IO io)
{
    extern __shared__ __align__(16) unsigned char psb_smem[];
    const int N = STAGES::IS_STATIC ? STAGES::N : plan.N;
    const int LPC = LPC_T > 0 ? LPC_T : lpc_rt;          // compile-time for the compiled plans: index math folds to shifts
    const int LPCP = LPC + 2;                            // even row pitch: two adjacent lines form an aligned 16-byte pair
    Cx<T>* s = reinterpret_cast<Cx<T>*>(psb_smem);
    Cx<T>* tw = s + (size_t)N * LPCP;
    for (int i = threadIdx.x; i < N; i += blockDim.x) tw[i] = tw_g[i];
    io.template load<LPC_T>(s, LPCP, LPC, N);
    __syncthreads();
    const int line = threadIdx.x % LPC, t = threadIdx.x / LPC;
    STAGES::template run<T, DIR, MAXB>(plan, s + line, LPCP, t, TPL, tw);
    io.template store<LPC_T>(s, LPCP, LPC, N);
}

template <typename T, int DIR, int MAXB, int NTHR, int MINB, int LPC, class STAGES, class IO>
__global__ void __launch_bounds__(NTHR, MINB) fft_lines_fused_kernel(int TPL, const Cx<T>* __restrict__ tw_g, IO io)
{
    extern __shared__ __align__(16) unsigned char psb_smem[];
    constexpr int N = STAGES::N;
    constexpr int LPCP = LPC + 2;
    Cx<T>* s = reinterpret_cast<Cx<T>*>(psb_smem);
    Cx<T>* tw = s + (size_t)N * LPCP;
    for (int i = threadIdx.x; i < N; i += NTHR) tw[i] = tw_g[i];           // first used after the stage-0 barrier
    const int line = threadIdx.x % LPC, t = threadIdx.x / LPC;
    typename IO::Acc acc;
    io.template acc_init_t<LPC>(acc, line, N);
    STAGES::template run_fused<T, DIR, MAXB, LPC, IO>(s + line, LPCP, t, TPL, tw, io, line, acc);
    io.finish(acc);
}

// ---------------------------------------------------------------------------------------
// IO functors
// ---------------------------------------------------------------------------------------
template <typename T> struct IoRows {
    Cx<T>* g;
    long long nrows;
    template <int LPC_T> __device__ void load(Cx<T>* s, int LPCP, int LPC, int N) const {
        const long long row0 = (long long)blockIdx.x * LPC;
        const Cx<T>* base = g + row0 * N;
        const int nvalid = (int)((nrows - row0 < LPC ? nrows - row0 : LPC)) * N;
        for (int e = threadIdx.x; e < LPC * N; e += blockDim.x) {
            const int line = e / N, idx = e - line * N;
            s[idx * LPCP + line] = e < nvalid ? base[e] : mk<T>(0, 0);
        }
    }
    template <int LPC_T> __device__ void store(const Cx<T>* s, int LPCP, int LPC, int N) const {
        const long long row0 = (long long)blockIdx.x * LPC;
        Cx<T>* base = g + row0 * N;
        const int nvalid = (int)((nrows - row0 < LPC ? nrows - row0 : LPC)) * N;
        for (int e = threadIdx.x; e < nvalid; e += blockDim.x) {
            const int line = e / N, idx = e - line * N;
            base[e] = s[idx * LPCP + line];
        }
    }
};

// lines = `nlines` consecutive elements; line element idx lives at  base(b) + cidx*in_istride + line
// where cidx = idx (dense) or the compact index of signed k in [-Rm, Rp] (pruned, zero outside).
template <typename T> struct Pair;                     // two adjacent complex elements as one vector access
template <> struct alignas(16) Pair<float> { Cx<float> a, b; };
template <> struct alignas(32) Pair<double> { Cx<double> a, b; };
template <typename T> struct Real2;
template <> struct alignas(8) Real2<float> { float a, b; };
template <> struct alignas(16) Real2<double> { double a, b; };

// (a,b) = values of two adjacent cells -> {half2 hi(a,b), half2 lo(a,b)} with hi + lo = value to ~2^-22 (bit patterns in a Real2)
__device__ __forceinline__ Real2<float> pack_hilo(Real2<float> v)
{
    const __half2 hh = __floats2half2_rn(v.a, v.b);
    const float2 back = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(v.a - back.x, v.b - back.y);
    Real2<float> r;
    r.a = __uint_as_float(*reinterpret_cast<const unsigned int*>(&hh));
    r.b = __uint_as_float(*reinterpret_cast<const unsigned int*>(&ll));
    return r;
}
__device__ __forceinline__ Real2<double> pack_hilo(Real2<double> v) { return v; }       // float64 path is never packed

template <typename T, bool PRUNED, bool REALOUT, bool ROUTED = false> struct IoCols {
    const Cx<T>* in;
    Cx<T>* out;            // complex output (REALOUT == false)
    T* outa; T* outb;      // planar real outputs (REALOUT == true): re -> outa, im -> outb (outb may be null)
    double* sumsq;         // REALOUT: sumsq[0] += sum re^2, sumsq[1] += sum im^2 (of the stored, scaled values)
    const float* scale2;   // REALOUT: optional power-of-two scales for (re, im) applied before the store
    unsigned int* maxabs2; // REALOUT: optional running max |stored value| per plane (float bits, atomicMax)
    int halfpack;          // REALOUT, float only: store each aligned cell pair (x, x+1) as {half2 hi(x,x+1), half2 lo(x,x+1)} (same 4 B/cell)
    int nlines;            // even; line strides and batch strides are even too (N is even) -> line pairs are 16-byte aligned
    long long in_bstride, in_istride, out_bstride, out_istride;
    int Rm, Rp;
    // REALOUT, fused kernels only -- routed output (multi-GPU): the output planes (line index idx = z) are cut into slabs
    // of `rplanes` planes, slab q belongs to rank q and is written straight into that rank's memory (peer pointer over NVLink):
    // route[q] / route[rranks + q] = address that plane a / plane b of this pair would start at on rank q if the slab buffer held
    // all planes (owner's row of the slab buffer minus q slabs), 0 = plane not stored.  null: everything goes to outa / outb.
    const long long* route;
    int rplanes, rranks;
    unsigned rmagic;       // floor(2^32 / rplanes) + 1: idx / rplanes == __umulhi(idx, rmagic) for idx < 2^16
    // ---- fused edge stages (fft_lines_fused_kernel): element-wise access from registers
    static constexpr bool FUSED = true;
    struct Acc {
        T m, q;                                            // running max |value| and sum of squares of this thread's plane
        T sa, sb;                                          // plane scales (loaded once)
        T* outp;                                           // this thread's output plane (even line: a, odd line: b), offset to its cell pair
                                                           // (routed output: the offset alone, as a pointer from null)
        const Cx<T>* inp;                                  // this thread's input column
        unsigned istr, ostr;                               // element strides between line indices (32-bit: every offset is < N^3 <= 2^30)
        int lo_max, hi_min, hi_sub;                        // pruned input: idx <= lo_max -> row idx + Rm; idx >= hi_min -> row idx - hi_sub
        bool live, st, even, pack;                         // line inside the grid; this thread stores; line parity; packed-half output
    };
    template <int LPC> __device__ __forceinline__ void acc_init_t(Acc& a, int line, int N) const {
        a.m = (T)0; a.q = (T)0;
        const int l = blockIdx.x * LPC + line;
        a.live = l < nlines;
        a.inp = in + (long long)blockIdx.y * in_bstride + l;
        a.istr = (unsigned)in_istride; a.ostr = (unsigned)out_istride;
        a.lo_max = a.live ? Rp : -1;                       // dead lines: no index passes the range tests
        a.hi_min = a.live ? N - Rm : (1 << 30);
        a.hi_sub = N - Rm;
        a.even = (line & 1) == 0;
        a.pack = halfpack != 0;
        if (REALOUT) {
            a.sa = scale2 ? (T)scale2[0] : (T)1;
            a.sb = scale2 ? (T)scale2[1] : (T)1;
            if (ROUTED) {
                a.st = a.live && route[(line & 1) ? rranks : 0] != 0;
                a.outp = static_cast<T*>(nullptr) + ((long long)blockIdx.y * out_bstride + (l & ~1));
            } else {
                T* pl = (line & 1) ? outb : outa;
                a.st = a.live && pl != nullptr;
                a.outp = (pl ? pl : outa) + (long long)blockIdx.y * out_bstride + (l & ~1);
            }
        } else {
            a.sa = a.sb = (T)1;
            a.st = a.live;
            a.outp = reinterpret_cast<T*>(out + (long long)blockIdx.y * out_bstride + l);
        }
    }
    __device__ __forceinline__ Cx<T> fetch(int idx, int N, const Acc& a) const {
        if (PRUNED) {
            // signed frequency k of idx lies in [-Rm, Rp]  <=>  idx <= Rp  or  idx >= N - Rm; compact row = k + Rm
            const bool lo = idx <= a.lo_max, hi = idx >= a.hi_min;
            if (!(lo || hi)) return mk<T>(0, 0);
            const unsigned row = (unsigned)(lo ? idx + Rm : idx - a.hi_sub);
            return a.inp[row * a.istr];
        }
        if (!a.live) return mk<T>(0, 0);
        return a.inp[(unsigned)idx * a.istr];
    }
    __device__ __forceinline__ void emit(int idx, int line, Cx<T> v, bool act, Acc& acc) const {
        const bool ok = act && acc.live;
        if (!REALOUT) {
            if (ok) reinterpret_cast<Cx<T>*>(acc.outp)[(unsigned)idx * acc.ostr] = v;
        } else {
            // lanes (line, line^1) are neighbours: the even one takes the pair of real parts (plane a), the odd one the pair of
            // imaginary parts (plane b), so every thread packs and stores one aligned cell pair
            const T ra = v.x * acc.sa, rb = v.y * acc.sb;
            const T recv = __shfl_xor_sync(0xffffffffu, acc.even ? rb : ra, 1);
            Real2<T> r;
            r.a = acc.even ? ra : recv;
            r.b = acc.even ? recv : rb;
            if (ok) {
                acc.m = fmax(acc.m, fmax(fabs(r.a), fabs(r.b)));
                acc.q += r.a * r.a + r.b * r.b;
                if (acc.st) {
                    if (halfpack) r = pack_hilo(r);
                    T* dst = acc.outp + (unsigned)idx * acc.ostr;
                    if (ROUTED) {                          // plane idx lives on rank idx / rplanes (compile-time: the local kernel pays nothing)
                        const unsigned q = __umulhi((unsigned)idx, rmagic);
                        dst = reinterpret_cast<T*>(route[(acc.even ? 0u : (unsigned)rranks) + q] + reinterpret_cast<long long>(dst));
                    }
                    *reinterpret_cast<Real2<T>*>(dst) = r;
                }
            }
        }
    }
    __device__ __forceinline__ void finish(Acc& acc) const {
        if (!REALOUT) return;
        T m = acc.m;
        double q = (double)acc.q;
        for (int o = 2; o < 32; o <<= 1) {                 // keeps the lane parity: lane 0 ends with plane a, lane 1 with plane b
            m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
            q += __shfl_xor_sync(0xffffffffu, q, o);
        }
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
        if (maxabs2 && lane < 2) atomicMax(&maxabs2[lane], __float_as_uint((float)m));
        __shared__ double red[2][32];
        if (lane < 2) red[lane][w] = q;
        __syncthreads();
        if (threadIdx.x < 64) {
            const int pl = threadIdx.x >> 5;
            q = lane < nw ? red[pl][lane] : 0.0;
            for (int o = 16; o > 0; o >>= 1) q += __shfl_down_sync(0xffffffffu, q, o);
            if (lane == 0) atomicAdd(&sumsq[pl], q);
        }
    }
    template <int LPC_T> __device__ void load(Cx<T>* s, int LPCP, int LPC, int N) const {
        const int l0 = blockIdx.x * LPC, HP = LPC / 2;
        const Cx<T>* base = in + (long long)blockIdx.y * in_bstride + l0;
        for (int e = threadIdx.x; e < HP * N; e += blockDim.x) {
            const int idx = e / HP, line = 2 * (e - idx * HP);
            Pair<T> v;
            v.a = mk<T>(0, 0); v.b = v.a;
            if (l0 + line < nlines) {
                if (PRUNED) {
                    const int k = kfreq(idx, N);
                    if (k >= -Rm && k <= Rp) v = *reinterpret_cast<const Pair<T>*>(base + (long long)(k + Rm) * in_istride + line);
                } else {
                    v = *reinterpret_cast<const Pair<T>*>(base + (long long)idx * in_istride + line);
                }
            }
            *reinterpret_cast<Pair<T>*>(s + idx * LPCP + line) = v;
        }
    }
    template <int LPC_T> __device__ void store(const Cx<T>* s, int LPCP, int LPC, int N) const {
        const int l0 = blockIdx.x * LPC, HP = LPC / 2;
        if (!REALOUT) {
            Cx<T>* base = out + (long long)blockIdx.y * out_bstride + l0;
            for (int e = threadIdx.x; e < HP * N; e += blockDim.x) {
                const int idx = e / HP, line = 2 * (e - idx * HP);
                if (l0 + line < nlines)
                    *reinterpret_cast<Pair<T>*>(base + (long long)idx * out_istride + line) = *reinterpret_cast<const Pair<T>*>(s + idx * LPCP + line);
            }
        } else {
            const long long boff = (long long)blockIdx.y * out_bstride + l0;
            double qa = 0.0, qb = 0.0;
            const T sa_ = scale2 ? (T)scale2[0] : (T)1, sb_ = scale2 ? (T)scale2[1] : (T)1;
            T ma = 0, mb = 0;
            for (int e = threadIdx.x; e < HP * N; e += blockDim.x) {
                const int idx = e / HP, line = 2 * (e - idx * HP);
                if (l0 + line < nlines) {
                    const Pair<T> v = *reinterpret_cast<const Pair<T>*>(s + idx * LPCP + line);
                    Real2<T> ra, rb;
                    ra.a = v.a.x * sa_; ra.b = v.b.x * sa_; rb.a = v.a.y * sb_; rb.b = v.b.y * sb_;
                    ma = fmax(ma, fmax(fabs(ra.a), fabs(ra.b))); mb = fmax(mb, fmax(fabs(rb.a), fabs(rb.b)));
                    qa += (double)ra.a * (double)ra.a + (double)ra.b * (double)ra.b;
                    qb += (double)rb.a * (double)rb.a + (double)rb.b * (double)rb.b;
                    const long long o = boff + (long long)idx * out_istride + line;
                    if (halfpack) { ra = pack_hilo(ra); rb = pack_hilo(rb); }
                    *reinterpret_cast<Real2<T>*>(outa + o) = ra;
                    if (outb) *reinterpret_cast<Real2<T>*>(outb + o) = rb;
                }
            }
            if (maxabs2) {
                for (int o = 16; o > 0; o >>= 1) { ma = fmax(ma, __shfl_down_sync(0xffffffffu, ma, o)); mb = fmax(mb, __shfl_down_sync(0xffffffffu, mb, o)); }
                if ((threadIdx.x & 31) == 0) { atomicMax(&maxabs2[0], __float_as_uint((float)ma)); atomicMax(&maxabs2[1], __float_as_uint((float)mb)); }
            }
            // block reduction of the two sums of squares (shell power, pyspectrum.py:404)
            __shared__ double red[2][32];
            for (int o = 16; o > 0; o >>= 1) { qa += __shfl_down_sync(0xffffffffu, qa, o); qb += __shfl_down_sync(0xffffffffu, qb, o); }
            const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
            if ((threadIdx.x & 31) == 0) { red[0][w] = qa; red[1][w] = qb; }
            __syncthreads();
            if (threadIdx.x < 32) {
                qa = threadIdx.x < nw ? red[0][threadIdx.x] : 0.0;
                qb = threadIdx.x < nw ? red[1][threadIdx.x] : 0.0;
                for (int o = 16; o > 0; o >>= 1) { qa += __shfl_down_sync(0xffffffffu, qa, o); qb += __shfl_down_sync(0xffffffffu, qb, o); }
                if (threadIdx.x == 0) { atomicAdd(&sumsq[0], qa); atomicAdd(&sumsq[1], qb); }
            }
        }
    }
};

// mesh pass 3: lines [0,HW) = columns (kx0+l, ky), lines [HW,2HW) = mirror columns (-kx,-ky); idx = z.
struct IoZFcomb {
    const Cx<float>* g;        // full complex grid after the x and y passes, [z][y][x]
    Cx<float>* half;           // output half field [kz][ky][kx], kx in [0,N/2]
    const Cx<double>* rec;     // fcomb phase table, N/2+1
    const float* Wk;           // window table, N/2+1
    const double* sumw;        // device scalar: sum of weights (periodic) -- ignored if !periodic
    int periodic;
    template <int LPC_T> __device__ void load(Cx<float>* s, int LPCP, int LPC, int N) const {
        const int HW = LPC / 2, h = N / 2;
        const int kx0 = blockIdx.x * HW, ky = blockIdx.y;
        for (int e = threadIdx.x; e < LPC * N; e += blockDim.x) {
            const int idx = e / LPC, line = e - idx * LPC;
            const int l = line < HW ? line : line - HW;
            const int kx = kx0 + l;
            Cx<float> v = mk<float>(0.f, 0.f);
            if (kx <= h) {
                const int x = line < HW ? kx : kneg(kx, N), y = line < HW ? ky : kneg(ky, N);
                v = g[((long long)idx * N + y) * N + x];
            }
            s[idx * LPCP + line] = v;
        }
    }
    template <int LPC_T> __device__ void store(const Cx<float>* s, int LPCP, int LPC, int N) const {
        const int HW = LPC / 2, h = N / 2;
        const int kx0 = blockIdx.x * HW, ky = blockIdx.y;
        const float cf = periodic ? 1.f / (864.f * (float)(*sumw)) : 1.f / 864.f;     // f:615 / f:686
        for (int e = threadIdx.x; e < HW * N; e += blockDim.x) {
            const int kz = e / HW, l = e - kz * HW;
            const int kx = kx0 + l;
            if (kx <= h) {
                const Cx<float> Fk = s[kz * LPCP + l];
                const Cx<float> Fm = s[kneg(kz, N) * LPCP + HW + l];
                half[((long long)kz * N + ky) * (h + 1) + kx] = fcomb_value(N, kx, ky, kz, Fk, Fm, rec, Wk, cf);
            }
        }
    }
};

// shell pass 1: line = compact (ky',kz'), idx = kx index; input built on the fly from the half field.
template <typename T> struct IoShellX {
    const Cx<float>* half;     // delta half field [kz][ky][kx] on the SOURCE grid Ns; null -> delta == 1 (triangle counts, py:977)
    const unsigned short* irk; // shell index of m = kx^2+ky^2+kz^2 (host table, pyspectrum.py:378)
    Cx<T>* out;                // T1 [kz'][ky'][x]
    int sa, sb;                // shell indices packed as real / imaginary part (sb < 0: none)
    int Rm, Rp, W;             // signed k range [-Rm,Rp], W = Rm+Rp+1
    int Ns;                    // grid of the half field (>= N).  Ns > N: the shells are transformed on a COARSER grid N than the one
                               // delta(k) was measured on -- the same band-limited field sampled at fewer points (needs R < N/2)
    template <int LPC_T> __device__ void load(Cx<T>* s, int LPCP, int LPC, int N) const {
        const int hs = Ns / 2;
        const int kyp0 = blockIdx.x * LPC, kzp = blockIdx.y;
        const int kz = kzp - Rm;
        for (int e = threadIdx.x; e < LPC * N; e += blockDim.x) {
            const int line = e / N, idx = e - line * N;
            Cx<T> v = mk<T>(0, 0);
            const int kyp = kyp0 + line;
            const int kx = kfreq(idx, N);
            if (kyp < W && kx >= -Rm && kx <= Rp) {
                const int ky = kyp - Rm;
                const int m = kx * kx + ky * ky + kz * kz;
                const int sh = irk[m];
                if (sh == sa || sh == sb) {
                    Cx<float> d = mk<float>(1.f, 0.f);
                    if (half) {
                        if (kx >= 0) {
                            const int kyi = ky < 0 ? ky + Ns : ky, kzi = kz < 0 ? kz + Ns : kz;
                            d = half[((long long)kzi * Ns + kyi) * (hs + 1) + kx];
                            // reflect_delta (py:1149-1156): the self-conjugate points are made real
                            if ((kx == 0 || kx == hs) && (kyi == 0 || kyi == hs) && (kzi == 0 || kzi == hs)) d.y = 0.f;
                        } else {
                            const int kyn = ky > 0 ? Ns - ky : -ky, kzn = kz > 0 ? Ns - kz : -kz;      // indices of -ky, -kz
                            d = conj(half[((long long)kzn * Ns + kyn) * (hs + 1) + (-kx)]);
                        }
                    }
                    v = (sh == sa) ? mk<T>((T)d.x, (T)d.y) : mk<T>(-(T)d.y, (T)d.x);
                }
            }
            s[idx * LPCP + line] = v;
        }
    }
    template <int LPC_T> __device__ void store(const Cx<T>* s, int LPCP, int LPC, int N) const {
        const int kyp0 = blockIdx.x * LPC, kzp = blockIdx.y;
        for (int e = threadIdx.x; e < LPC * N; e += blockDim.x) {
            const int line = e / N, idx = e - line * N;
            const int kyp = kyp0 + line;
            if (kyp < W) out[((long long)kzp * W + kyp) * N + idx] = s[idx * LPCP + line];
        }
    }
};

// ---------------------------------------------------------------------------------------
// launch: compiled plans for the production grids (BASELINE configs: 256, 360, 512, 1024) and the
// small grids the parity tests use; any other N = 2^a 3^b 5^c takes the runtime-plan fallback.
// ---------------------------------------------------------------------------------------
template <class STAGES, int LPC_, int MAXB_, int MINB_ = 1> struct Cfg {
    using Stages = STAGES;
    static constexpr int LPC = LPC_, MAXB = MAXB_, MINB = MINB_;
    static constexpr int TPL = (STAGES::NB_MAX + MAXB_ - 1) / MAXB_;
    static constexpr int NTHR = LPC * TPL;
    static_assert(NTHR <= 1024, "too many threads");
};
struct CfgDyn { using Stages = DynStages; static constexpr int MAXB = 2, NTHR = 256; };

template <class IO, class = void> struct io_is_fused : std::false_type {};
template <class IO> struct io_is_fused<IO, std::enable_if_t<IO::FUSED>> : std::true_type {};

template <typename T, int DIR, class CFG, class IO>
static int launch_static(dim3 grid, const Cx<T>* tw, const IO& io, cudaStream_t st)
{
    constexpr int N = CFG::Stages::N;
    const size_t smem = ((size_t)N * (CFG::LPC + 2) + N) * sizeof(Cx<T>);
    if constexpr (io_is_fused<IO>::value && CFG::Stages::NSTAGES >= 2) {
        static_assert(CFG::LPC % 2 == 0 && 32 % CFG::LPC == 0, "line pairs must sit in one warp");
        auto kf = fft_lines_fused_kernel<T, DIR, CFG::MAXB, CFG::NTHR, (sizeof(T) == 4 ? CFG::MINB : 1), CFG::LPC, typename CFG::Stages, IO>;
        if (cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return PSB_ERR_CUDA;
        cudaFuncSetAttribute(kf, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        kf<<<grid, CFG::NTHR, smem, st>>>(CFG::TPL, tw, io);
        return cudaGetLastError() == cudaSuccess ? PSB_OK : PSB_ERR_CUDA;
    }
    auto kern = fft_lines_kernel<T, DIR, CFG::MAXB, CFG::NTHR, (sizeof(T) == 4 ? CFG::MINB : 1), CFG::LPC, typename CFG::Stages, IO>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return PSB_ERR_CUDA;
    // ask for the largest shared-memory carveout so that several CTAs co-reside and overlap load / FFT / store
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    FftPlan p; p.N = N; p.nstages = 0; p.nb_max = 0;
    kern<<<grid, CFG::NTHR, smem, st>>>(p, CFG::LPC, CFG::TPL, tw, io);
    return cudaGetLastError() == cudaSuccess ? PSB_OK : PSB_ERR_CUDA;
}

template <typename T, int DIR, class IO>
static int launch_dyn(const FftPlan& p, int LPC, dim3 grid, const Cx<T>* tw, const IO& io, cudaStream_t st)
{
    const int TPL = (p.nb_max + 1) / 2;
    if (LPC * TPL > 256) return PSB_ERR_UNSUPPORTED_N;
    const size_t smem = ((size_t)p.N * (LPC + 2) + p.N) * sizeof(Cx<T>);
    if (smem > 200 * 1024) return PSB_ERR_UNSUPPORTED_N;
    auto kern = fft_lines_kernel<T, DIR, 2, 256, 1, 0, DynStages, IO>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return PSB_ERR_CUDA;
    kern<<<grid, LPC * TPL, smem, st>>>(p, LPC, TPL, tw, io);
    return cudaGetLastError() == cudaSuccess ? PSB_OK : PSB_ERR_CUDA;
}

// lines-per-CTA for the fallback: largest even value keeping the CTA within 256 threads
static inline int dyn_lpc(const FftPlan& p) { int tpl = (p.nb_max + 1) / 2; int l = 256 / tpl; if (l > 16) l = 16; l &= ~1; return l; }

// F is a generic callable: f(cfg_tag) -> int, where cfg_tag is Cfg<...>{} or CfgDyn{}
#define PSB_PLAN_CASE(N_, LPC_, MAXB_, ...) case N_: return f(Cfg<StaticStages<__VA_ARGS__>, LPC_, MAXB_>{});
#define PSB_PLAN_CASE_OCC(N_, LPC_, MAXB_, MINB_, ...) case N_: return f(Cfg<StaticStages<__VA_ARGS__>, LPC_, MAXB_, MINB_>{});
template <class F> static int dispatch_plan(int N, F&& f)
{
    switch (N) {
        PSB_PLAN_CASE(24, 16, 1, 8, 3)
        PSB_PLAN_CASE(32, 16, 1, 8, 4)
        PSB_PLAN_CASE(36, 16, 1, 9, 4)
        PSB_PLAN_CASE(48, 16, 1, 4, 4, 3)
        PSB_PLAN_CASE(64, 16, 1, 8, 8)
        PSB_PLAN_CASE(128, 16, 1, 8, 4, 4)
        PSB_PLAN_CASE_OCC(256, 16, 1, 3, 16, 16)
        PSB_PLAN_CASE_OCC(320, 16, 1, 3, 20, 16)         // 320, 400: coarse shell grids (pyspectrum.py coarse_levels)
        PSB_PLAN_CASE_OCC(360, 16, 1, 3, 20, 18)
        PSB_PLAN_CASE_OCC(400, 16, 1, 3, 20, 20)
        PSB_PLAN_CASE_OCC(512, 16, 2, 2, 8, 8, 8)
        PSB_PLAN_CASE(1024, 8, 4, 8, 8, 4, 4)
        default: break;
    }
    return f(CfgDyn{});
}

template <typename T, int DIR, class CFG, class IO>
static int launch_any(CFG, const FftPlan& p, int lpc, dim3 grid, const Cx<T>* tw, const IO& io, cudaStream_t st)
{
    if constexpr (CFG::Stages::IS_STATIC) return launch_static<T, DIR, CFG, IO>(grid, tw, io, st);
    else return launch_dyn<T, DIR, IO>(p, lpc, grid, tw, io, st);
}
// whether launch_static takes the fused-edge-stage kernel for this configuration (compiled plans with >= 2 stages)
template <class CFG> static bool plan_is_fused(CFG)
{
    if constexpr (CFG::Stages::IS_STATIC) return CFG::Stages::NSTAGES >= 2; else return false;
}
template <class CFG> static int cfg_lpc(CFG, const FftPlan& p)
{
    if constexpr (CFG::Stages::IS_STATIC) return CFG::LPC; else return dyn_lpc(p);
}

}  // namespace psb
