// psb_fft_shell.cuh -- packed shell-pair transform: half field (or delta == 1) -> two real fields.
// Replaces, per pair of shells, pyspectrum.py:387-404 (mask, complex64 FFTW forward, np.real, shell
// power) and reflect_delta (py:1134-1157); with T = double and half == nullptr it is the counts cold
// path (py:975-1011 / estimator.f:27-86).  All three passes are pruned to |k_i| <= R.
#pragma once
#include <cstdlib>
#include "psb_fft_lines.cuh"

namespace psb {

template <typename T>
int fft_shell_pair(const Cx<float>* half, const unsigned short* irk, int N, int Ns, int sa, int sb, int R,
                   Cx<T>* t1, Cx<T>* t2, T* fa, T* fb, double* sumsq, const float* scale2, unsigned int* maxabs2, int halfpack,
                   const Cx<T>* tw, cudaStream_t st, const long long* route, int rplanes, int rranks)
{
    FftPlan p;
    if (N % 2 || !make_plan(N, &p)) return PSB_ERR_UNSUPPORTED_N;
    if (Ns <= 0) Ns = N;
    if (Ns % 2 || Ns < N || (Ns > N && 2 * R >= N)) return PSB_ERR_ARG;      // a coarser transform grid must hold the shells without wrap
    if (route && (rplanes < 1 || rranks < 1 || rplanes * rranks != N || N >= 65536)) return PSB_ERR_ARG;
    const int Rp = R < N / 2 ? R : N / 2;
    const int Rm = R < (N - 1) / 2 ? R : (N - 1) / 2;
    const int W = Rm + Rp + 1;
    return dispatch_plan(N, [&](auto cfg) -> int {
        const int LPC = cfg_lpc(cfg, p);
        if (LPC < 1) return (int)PSB_ERR_UNSUPPORTED_N;
        IoShellX<T> io1{ half, irk, t1, sa, sb, Rm, Rp, W, Ns };
        int rc = launch_any<T, -1>(cfg, p, LPC, dim3((W + LPC - 1) / LPC, W), tw, io1, st);
        if (rc) return rc;
        // pass 2 (y): batch = kz', T1[kz'][ky'][x] -> T2[kz'][y][x]
        IoCols<T, true, false> io2{ t1, t2, nullptr, nullptr, nullptr, nullptr, nullptr, 0, N, (long long)W * N, N, (long long)N * N, N, Rm, Rp };
        rc = launch_any<T, -1>(cfg, p, LPC, dim3((N + LPC - 1) / LPC, W), tw, io2, st);
        if (rc) return rc;
        // pass 3 (z): batch = y, T2[kz'][y][x] -> real planes [z][y][x] (+ sums of squares)
        if (!route) {
            IoCols<T, true, true> io3{ t2, nullptr, fa, fb, sumsq, scale2, maxabs2, halfpack, N, N, (long long)N * N, N, (long long)N * N, Rm, Rp,
                                       nullptr, 0, 0, 0u };
            return launch_any<T, -1>(cfg, p, LPC, dim3((N + LPC - 1) / LPC, N), tw, io3, st);
        }
        // routed output (multi-GPU): its own instantiation of the epilogue; stores exist in the fused kernels only
        if (!plan_is_fused(cfg)) return (int)PSB_ERR_UNSUPPORTED_N;
        IoCols<T, true, true, true> io3{ t2, nullptr, fa, fb, sumsq, scale2, maxabs2, halfpack, N, N, (long long)N * N, N, (long long)N * N, Rm, Rp,
                                         route, rplanes, rranks, (unsigned)(4294967296ULL / (unsigned)rplanes) + 1u };
        // the routed z pass runs with 32 lines per CTA: a CTA's (plane, z) run is 128 bytes instead of 64, which NVLink stores need
        // (2 GPUs, C2: shell stage 5.78 -> 4.97 ms; with local stores the 16-line kernel wins, 5.3 vs 6.0 ms for the 20 pairs of C2:
        // one CTA of 640 threads per SM instead of three of 320).  PSB_ROUTED_LPC=16 restores it.
        using CFG = decltype(cfg);
        if constexpr (CFG::Stages::IS_STATIC && sizeof(T) == 4) {
            if constexpr (CFG::Stages::NSTAGES >= 2 && CFG::LPC == 16 && 32 * CFG::TPL <= 1024) {
                static const int wide = [] { const char* e = getenv("PSB_ROUTED_LPC"); return e ? atoi(e) : 32; }();
                if (wide == 32) {
                    using C32 = Cfg<typename CFG::Stages, 32, CFG::MAXB, 1>;
                    return launch_static<T, -1, C32, IoCols<T, true, true, true>>(dim3((N + 31) / 32, N), tw, io3, st);
                }
            }
        }
        return launch_any<T, -1>(cfg, p, LPC, dim3((N + LPC - 1) / LPC, N), tw, io3, st);
    });
}

}  // namespace psb
