// psb_common.cuh -- shared types for the pyspectrum_b200 CUDA kernels (sm_100a).
// Everything marked PSB_HD also compiles with a plain host compiler so the per-thread
// math (DFT butterflies, index maps, combine rule, weights) is unit-tested on CPU
// (tests/host_emu/); the kernels themselves only run on the GPU.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define PSB_HD __host__ __device__ __forceinline__
#define PSB_D __device__ __forceinline__
#else
#define PSB_HD inline
#define PSB_D inline
#endif

namespace psb {

template <typename T> struct Cx;
template <> struct alignas(8) Cx<float> { float x, y; };
template <> struct alignas(16) Cx<double> { double x, y; };

template <typename T> PSB_HD Cx<T> mk(T x, T y) { Cx<T> c; c.x = x; c.y = y; return c; }
template <typename T> PSB_HD Cx<T> operator+(Cx<T> a, Cx<T> b) { return mk<T>(a.x + b.x, a.y + b.y); }
template <typename T> PSB_HD Cx<T> operator-(Cx<T> a, Cx<T> b) { return mk<T>(a.x - b.x, a.y - b.y); }
template <typename T> PSB_HD Cx<T> operator*(Cx<T> a, Cx<T> b) { return mk<T>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
template <typename T> PSB_HD Cx<T> operator*(T s, Cx<T> a) { return mk<T>(s * a.x, s * a.y); }
template <typename T> PSB_HD Cx<T> conj(Cx<T> a) { return mk<T>(a.x, -a.y); }
// multiply by (DIR * i): rotation by +-90 degrees
template <int DIR, typename T> PSB_HD Cx<T> mul_i(Cx<T> a) { return DIR > 0 ? mk<T>(-a.y, a.x) : mk<T>(a.y, -a.x); }

// signed wave number of FFT index i on an N-grid: 0..N/2 -> itself, above -> i-N
PSB_HD int kfreq(int i, int N) { return (i <= N / 2) ? i : i - N; }
PSB_HD int kneg(int i, int N) { return i == 0 ? 0 : N - i; }     // index of -k

}  // namespace psb

// Error codes returned by every C-ABI entry point (same values as include/psb200.h)
#ifndef PSB_OK
#define PSB_OK 0
#define PSB_ERR_ARG (-1)
#define PSB_ERR_UNSUPPORTED_N (-2)
#define PSB_ERR_CUDA (-3)
#define PSB_ERR_WORKSPACE (-4)
#endif
