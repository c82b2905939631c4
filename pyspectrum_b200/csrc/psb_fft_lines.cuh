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
#include "psb_fft_core.cuh"
#include "psb_fcomb_core.cuh"
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

template <typename T, int DIR, int MAXB, int NTHR, class STAGES, class IO>
__global__ void __launch_bounds__(NTHR) fft_lines_kernel(FftPlan plan, int LPC, int TPL, const Cx<T>* __restrict__ tw_g, IO io)
{
    extern __shared__ __align__(16) unsigned char psb_smem[];
    const int N = STAGES::IS_STATIC ? STAGES::N : plan.N;
    const int LPCP = LPC + 1;
    Cx<T>* s = reinterpret_cast<Cx<T>*>(psb_smem);
    Cx<T>* tw = s + (size_t)N * LPCP;
    for (int i = threadIdx.x; i < N; i += blockDim.x) tw[i] = tw_g[i];
    io.load(s, LPCP, LPC, N);
    __syncthreads();
    const int line = threadIdx.x % LPC, t = threadIdx.x / LPC;
    STAGES::template run<T, DIR, MAXB>(plan, s + line, LPCP, t, TPL, tw);
    io.store(s, LPCP, LPC, N);
}

// ---------------------------------------------------------------------------------------
// IO functors
// ---------------------------------------------------------------------------------------
template <typename T> struct IoRows {
    Cx<T>* g;
    long long nrows;
    __device__ void load(Cx<T>* s, int LPCP, int LPC, int N) const {
        const long long row0 = (long long)blockIdx.x * LPC;
        for (int e = threadIdx.x; e < LPC * N; e += blockDim.x) {
            const int line = e / N, idx = e - line * N;
            const long long row = row0 + line;
            s[(size_t)idx * LPCP + line] = row < nrows ? g[row * N + idx] : mk<T>(0, 0);
        }
    }
    __device__ void store(const Cx<T>* s, int LPCP, int LPC, int N) const {
        const long long row0 = (long long)blockIdx.x * LPC;
        for (int e = threadIdx.x; e < LPC * N; e += blockDim.x) {
            const int line = e / N, idx = e - line * N;
            const long long row = row0 + line;
            if (row < nrows) g[row * N + idx] = s[(size_t)idx * LPCP + line];
        }
    }
};

// lines = `nlines` consecutive elements; line element idx lives at  base(b) + cidx*in_istride + line
// where cidx = idx (dense) or the compact index of signed k in [-Rm, Rp] (pruned, zero outside).
template <typename T, bool PRUNED, bool REALOUT> struct IoCols {
    const Cx<T>* in;
    Cx<T>* out;            // complex output (REALOUT == false)
    T* outa; T* outb;      // planar real outputs (REALOUT == true): re -> outa, im -> outb (outb may be null)
    double* sumsq;         // REALOUT: sumsq[0] += sum re^2, sumsq[1] += sum im^2 (of the stored, scaled values)
    const float* scale2;   // REALOUT: optional power-of-two scales for (re, im) applied before the store
    unsigned int* maxabs2; // REALOUT: optional running max |stored value| per plane (float bits, atomicMax)
    int nlines;
    long long in_bstride, in_istride, out_bstride, out_istride;
    int Rm, Rp;
    __device__ void load(Cx<T>* s, int LPCP, int LPC, int N) const {
        const int l0 = blockIdx.x * LPC;
        const Cx<T>* base = in + (long long)blockIdx.y * in_bstride;
        for (int e = threadIdx.x; e < LPC * N; e += blockDim.x) {
            const int idx = e / LPC, line = e - idx * LPC;
            Cx<T> v = mk<T>(0, 0);
            if (l0 + line < nlines) {
                if (PRUNED) {
                    const int k = kfreq(idx, N);
                    if (k >= -Rm && k <= Rp) v = base[(long long)(k + Rm) * in_istride + l0 + line];
                } else {
                    v = base[(long long)idx * in_istride + l0 + line];
                }
            }
            s[(size_t)idx * LPCP + line] = v;
        }
    }
    __device__ void store(const Cx<T>* s, int LPCP, int LPC, int N) const {
        const int l0 = blockIdx.x * LPC;
        if (!REALOUT) {
            Cx<T>* base = out + (long long)blockIdx.y * out_bstride;
            for (int e = threadIdx.x; e < LPC * N; e += blockDim.x) {
                const int idx = e / LPC, line = e - idx * LPC;
                if (l0 + line < nlines) base[(long long)idx * out_istride + l0 + line] = s[(size_t)idx * LPCP + line];
            }
        } else {
            const long long boff = (long long)blockIdx.y * out_bstride;
            double qa = 0.0, qb = 0.0;
            const T sa_ = scale2 ? (T)scale2[0] : (T)1, sb_ = scale2 ? (T)scale2[1] : (T)1;
            T ma = 0, mb = 0;
            for (int e = threadIdx.x; e < LPC * N; e += blockDim.x) {
                const int idx = e / LPC, line = e - idx * LPC;
                if (l0 + line < nlines) {
                    Cx<T> v = s[(size_t)idx * LPCP + line];
                    v.x *= sa_; v.y *= sb_;
                    ma = fmax(ma, fabs(v.x)); mb = fmax(mb, fabs(v.y));
                    const long long o = boff + (long long)idx * out_istride + l0 + line;
                    outa[o] = v.x;
                    if (outb) outb[o] = v.y;
                    qa += (double)v.x * (double)v.x;
                    qb += (double)v.y * (double)v.y;
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
    __device__ void load(Cx<float>* s, int LPCP, int LPC, int N) const {
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
            s[(size_t)idx * LPCP + line] = v;
        }
    }
    __device__ void store(const Cx<float>* s, int LPCP, int LPC, int N) const {
        const int HW = LPC / 2, h = N / 2;
        const int kx0 = blockIdx.x * HW, ky = blockIdx.y;
        const float cf = periodic ? 1.f / (864.f * (float)(*sumw)) : 1.f / 864.f;     // f:615 / f:686
        for (int e = threadIdx.x; e < HW * N; e += blockDim.x) {
            const int kz = e / HW, l = e - kz * HW;
            const int kx = kx0 + l;
            if (kx <= h) {
                const Cx<float> Fk = s[(size_t)kz * LPCP + l];
                const Cx<float> Fm = s[(size_t)kneg(kz, N) * LPCP + HW + l];
                half[((long long)kz * N + ky) * (h + 1) + kx] = fcomb_value(N, kx, ky, kz, Fk, Fm, rec, Wk, cf);
            }
        }
    }
};

// shell pass 1: line = compact (ky',kz'), idx = kx index; input built on the fly from the half field.
template <typename T> struct IoShellX {
    const Cx<float>* half;     // delta half field [kz][ky][kx]; null -> delta == 1 (triangle counts, py:977)
    const unsigned short* irk; // shell index of m = kx^2+ky^2+kz^2 (host table, pyspectrum.py:378)
    Cx<T>* out;                // T1 [kz'][ky'][x]
    int sa, sb;                // shell indices packed as real / imaginary part (sb < 0: none)
    int Rm, Rp, W;             // signed k range [-Rm,Rp], W = Rm+Rp+1
    __device__ void load(Cx<T>* s, int LPCP, int LPC, int N) const {
        const int h = N / 2;
        const int kyp0 = blockIdx.x * LPC, kzp = blockIdx.y;
        const int kz = kzp - Rm;
        const int kzi = kz < 0 ? kz + N : kz;
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
                    const int kyi = ky < 0 ? ky + N : ky;
                    Cx<float> d = mk<float>(1.f, 0.f);
                    if (half) {
                        if (idx <= h) {
                            d = half[((long long)kzi * N + kyi) * (h + 1) + idx];
                            // reflect_delta (py:1149-1156): the self-conjugate points are made real
                            if ((idx == 0 || idx == h) && (kyi == 0 || kyi == h) && (kzi == 0 || kzi == h)) d.y = 0.f;
                        } else {
                            d = conj(half[((long long)kneg(kzi, N) * N + kneg(kyi, N)) * (h + 1) + (N - idx)]);
                        }
                    }
                    v = (sh == sa) ? mk<T>((T)d.x, (T)d.y) : mk<T>(-(T)d.y, (T)d.x);
                }
            }
            s[(size_t)idx * LPCP + line] = v;
        }
    }
    __device__ void store(const Cx<T>* s, int LPCP, int LPC, int N) const {
        const int kyp0 = blockIdx.x * LPC, kzp = blockIdx.y;
        for (int e = threadIdx.x; e < LPC * N; e += blockDim.x) {
            const int line = e / N, idx = e - line * N;
            const int kyp = kyp0 + line;
            if (kyp < W) out[((long long)kzp * W + kyp) * N + idx] = s[(size_t)idx * LPCP + line];
        }
    }
};

// ---------------------------------------------------------------------------------------
// launch: compiled plans for the production grids (BASELINE configs: 256, 360, 512, 1024) and the
// small grids the parity tests use; any other N = 2^a 3^b 5^c takes the runtime-plan fallback.
// ---------------------------------------------------------------------------------------
template <class STAGES, int LPC_, int MAXB_> struct Cfg {
    using Stages = STAGES;
    static constexpr int LPC = LPC_, MAXB = MAXB_;
    static constexpr int TPL = (STAGES::NB_MAX + MAXB_ - 1) / MAXB_;
    static constexpr int NTHR = LPC * TPL;
    static_assert(NTHR <= 1024, "too many threads");
};
struct CfgDyn { using Stages = DynStages; static constexpr int MAXB = 2, NTHR = 256; };

template <typename T, int DIR, class CFG, class IO>
static int launch_static(dim3 grid, const Cx<T>* tw, const IO& io, cudaStream_t st)
{
    constexpr int N = CFG::Stages::N;
    const size_t smem = ((size_t)N * (CFG::LPC + 1) + N) * sizeof(Cx<T>);
    auto kern = fft_lines_kernel<T, DIR, CFG::MAXB, CFG::NTHR, typename CFG::Stages, IO>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return PSB_ERR_CUDA;
    FftPlan p; p.N = N; p.nstages = 0; p.nb_max = 0;
    kern<<<grid, CFG::NTHR, smem, st>>>(p, CFG::LPC, CFG::TPL, tw, io);
    return cudaGetLastError() == cudaSuccess ? PSB_OK : PSB_ERR_CUDA;
}

template <typename T, int DIR, class IO>
static int launch_dyn(const FftPlan& p, int LPC, dim3 grid, const Cx<T>* tw, const IO& io, cudaStream_t st)
{
    const int TPL = (p.nb_max + 1) / 2;
    if (LPC * TPL > 256) return PSB_ERR_UNSUPPORTED_N;
    const size_t smem = ((size_t)p.N * (LPC + 1) + p.N) * sizeof(Cx<T>);
    if (smem > 200 * 1024) return PSB_ERR_UNSUPPORTED_N;
    auto kern = fft_lines_kernel<T, DIR, 2, 256, DynStages, IO>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return PSB_ERR_CUDA;
    kern<<<grid, LPC * TPL, smem, st>>>(p, LPC, TPL, tw, io);
    return cudaGetLastError() == cudaSuccess ? PSB_OK : PSB_ERR_CUDA;
}

// lines-per-CTA for the fallback: largest even value keeping the CTA within 256 threads
static inline int dyn_lpc(const FftPlan& p) { int tpl = (p.nb_max + 1) / 2; int l = 256 / tpl; if (l > 16) l = 16; l &= ~1; return l; }

// F is a generic callable: f(cfg_tag) -> int, where cfg_tag is Cfg<...>{} or CfgDyn{}
#define PSB_PLAN_CASE(N_, LPC_, MAXB_, ...) case N_: return f(Cfg<StaticStages<__VA_ARGS__>, LPC_, MAXB_>{});
template <class F> static int dispatch_plan(int N, F&& f)
{
    switch (N) {
        PSB_PLAN_CASE(24, 16, 1, 8, 3)
        PSB_PLAN_CASE(32, 16, 1, 8, 4)
        PSB_PLAN_CASE(36, 16, 1, 9, 4)
        PSB_PLAN_CASE(48, 16, 1, 4, 4, 3)
        PSB_PLAN_CASE(64, 16, 1, 8, 8)
        PSB_PLAN_CASE(128, 16, 1, 8, 4, 4)
        PSB_PLAN_CASE(256, 16, 2, 8, 8, 4)
        PSB_PLAN_CASE(360, 16, 2, 9, 8, 5)
        PSB_PLAN_CASE(512, 16, 2, 8, 8, 8)
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
template <class CFG> static int cfg_lpc(CFG, const FftPlan& p)
{
    if constexpr (CFG::Stages::IS_STATIC) return CFG::LPC; else return dyn_lpc(p);
}

}  // namespace psb
