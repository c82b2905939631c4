// psb_spectra.cu -- K4: one pass over the half field -> binned power spectra.
//   mode 0  replaces the Python bin loop of Pk_periodic (pyspectrum.py:690-716) on reflect_delta's field
//   mode 1  replaces estimator.f:155-264 (pk_pbox_rsd): multipoles + (k,mu) table
// A half-space mode with 0 < kx < N/2 stands for itself and its conjugate (weight 2); the kx = 0 and
// kx = N/2 planes hold both partners explicitly (weight 1).  The reference visits all N^3 modes; under
// k -> -k every float32 quantity of f:206-223 flips sign exactly, so |mu|, mu^2 and the bins agree.
// Bin indices come from a host table indexed by m = kx^2+ky^2+kz^2 that evaluates the reference's own
// expression (float64 int(x+0.5) for mode 0, float32 nint() for mode 1) -> mode counts are bit exact.
// The mu bin needs IEEE float32 sqrt/div with no contraction: explicit _rn intrinsics.
#include <cuda_runtime.h>
#include "psb_kernels.h"

namespace psb {

// one mode of the estimator.f:196-244 loop body with signed integer wave numbers (rkx,rky,rkz): adds the five per-k-bin
// contributions to acc[0..4] and the four (k,mu)-table contributions straight into `tab` (shared or global memory)
// tab == nullptr: the table contributions are handed back instead (tkey = b * 1024 + imu, 0 = none; tv[4]) for the caller's
// segmented warp reduction.
__device__ __forceinline__ void rsd_mode(const SpectraIn& in, double* acc, double* tab, int b, float rkx, float rky, float rkz, double pk, double wgt,
                                         int* tkey = nullptr, double* tv = nullptr)
{
    const int Nbin = in.Nbin;
    const float rk = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(rkx, rkx), __fmul_rn(rky, rky)), __fmul_rn(rkz, rkz)));
    const float cot1 = __fdiv_rn(rkz, rk);
    const float sit1 = __fsqrt_rn(__fsub_rn(1.f, __fmul_rn(cot1, cot1)));
    float cc = 0.f;
    if (sit1 > 0.f) {
        const float den = __fmul_rn(rk, sit1);
        const float cp = __fdiv_rn(rkx, den), sp = __fdiv_rn(rky, den);
        cc = __fadd_rn(__fmul_rn(in.sinph, sp), __fmul_rn(in.cosph, cp));
    }
    const double mu = (double)__fadd_rn(__fmul_rn(in.costh, cot1), __fmul_rn(__fmul_rn(in.sinth, sit1), cc));
    const double mubin = (double)__fdiv_rn(1.f, (float)in.Nmu);
    const double amu = fabs(mu);
    const int imu = (int)__ddiv_rn(__dadd_rn(amu, mubin), mubin);
    const double mu2 = mu * mu;
    const double Le2 = -0.5 + 1.5 * mu2;
    const double Le4 = 0.375 - 3.75 * mu2 + 4.375 * (mu2 * mu2);
    const double kk = (double)__fmul_rn(in.kf32, rk);
    acc[0] += wgt;
    acc[1] += wgt * kk;
    acc[2] += wgt * pk;
    acc[3] += wgt * (pk * 5.0 * Le2);
    acc[4] += wgt * (pk * 9.0 * Le4);
    if (imu <= in.Nmu && imu > 0 && tab == nullptr) {
        *tkey = b * 1024 + imu;
        tv[0] = wgt; tv[1] = wgt * kk; tv[2] = wgt * amu; tv[3] = wgt * pk;
    } else if (imu <= in.Nmu && imu > 0) {
        double* t = tab + (long long)(imu - 1) * Nbin + (b - 1);
        const long long tb = (long long)Nbin * in.Nmu;
        atomicAdd(t, wgt);
        atomicAdd(t + tb, wgt * kk);
        atomicAdd(t + 2 * tb, wgt * amu);
        atomicAdd(t + 3 * tb, wgt * pk);
    }
}

// One warp per (kz,ky) row of the half field.  Along a row the k-bin is non-decreasing in kx, so the per-bin sums are first
// reduced inside the warp (segmented shuffle reduction), then added to the CTA's bins in shared memory; one global float64
// atomic per CTA and bin at the end.  Rows that start beyond the last bin (the corners of the cube) are skipped without a load.
template <int MODE>
__global__ void __launch_bounds__(256) k_spectra(SpectraIn in, double* out, int table_in_smem)
{
    extern __shared__ double sbin[];
    constexpr int NV = MODE == 0 ? 3 : 5;
    const int N = in.N, h = N / 2, Nbin = in.Nbin, lane = threadIdx.x & 31;
    const int nsm = NV * Nbin + ((MODE == 1 && table_in_smem) ? 4 * in.Nmu * Nbin : 0);
    for (int i = threadIdx.x; i < nsm; i += blockDim.x) sbin[i] = 0.0;
    __syncthreads();
    double* tab = (MODE == 1 && table_in_smem) ? sbin + NV * Nbin : out + 5 * (long long)Nbin;
    // rows of the array: all kz, ky in [ky0, ky0 + ny) (the whole half field: ky0 = 0, ny = N; a rank's ky-slab otherwise)
    const int ny = in.ny, nrow = N * ny, nwarp = gridDim.x * (blockDim.x >> 5);
    for (int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < nrow; r += nwarp) {
        const int iz = r / ny, iy = in.ky0 + (r - iz * ny);
        const int ky = kfreq(iy, N), kz = kfreq(iz, N);
        const int m0 = ky * ky + kz * kz;
        if (in.bin[m0] > Nbin) continue;                         // bins are non-decreasing in m
        const Cx<float>* row = in.half + (long long)r * (h + 1);
        for (int x0 = 0; x0 <= h; x0 += 32) {
            const int ix = x0 + lane;
            int b = 0;
            if (ix <= h) { b = in.bin[m0 + ix * ix]; if (b > Nbin) b = 0; }
            if (__ballot_sync(0xffffffffu, b != 0) == 0u) { if (in.bin[m0 + x0 * x0] > Nbin) break; else continue; }
            double v[NV];
#pragma unroll
            for (int i = 0; i < NV; ++i) v[i] = 0.0;
            int tkey = 0;                                        // mode 1: (k bin, mu bin) of this lane's table contribution
            double tv[4] = { 0.0, 0.0, 0.0, 0.0 };
            if (b) {
                Cx<float> d = row[ix];
                const bool edge = ix == 0 || ix == h;
                if (MODE == 0) {
                    // reflect_delta (py:1149-1156) makes the self-conjugate points real before |.|^2
                    if (edge && (iy == 0 || iy == h) && (iz == 0 || iz == h)) d.y = 0.f;
                    const float pw = d.x * d.x + d.y * d.y;
                    const double wgt = edge ? 1.0 : 2.0;
                    v[0] = wgt;
                    v[1] = wgt * (in.kf * sqrt((double)(m0 + ix * ix)));
                    v[2] = wgt * (double)pw;
                } else {
                    const float ab = (float)sqrt((double)d.x * (double)d.x + (double)d.y * (double)d.y);    // cabs()
                    const double pk = (double)__fmul_rn(ab, ab);
                    const bool warp_reduce = in.Nmu < 1024;          // table contributions leave through the segmented reduction below
                    if (edge) {
                        rsd_mode(in, v, warp_reduce ? nullptr : tab, b, (float)ix, (float)ky, (float)kz, pk, 1.0, &tkey, tv);
                    } else if (iy != h && iz != h) {
                        rsd_mode(in, v, warp_reduce ? nullptr : tab, b, (float)ix, (float)ky, (float)kz, pk, 2.0, &tkey, tv);      // partner is exactly -k
                    } else {
                        // the conjugate partner visited by the Fortran loop is (-kx, -ky, -kz) with a Nyquist
                        // component folded back to +N/2 (f:198-204), so it is not the exact negation: do both
                        rsd_mode(in, v, tab, b, (float)ix, (float)ky, (float)kz, pk, 1.0);
                        rsd_mode(in, v, tab, b, (float)(-ix), (float)(iy == h ? h : -ky), (float)(iz == h ? h : -kz), pk, 1.0);
                    }
                }
            }
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int tb = __shfl_down_sync(0xffffffffu, b, o);
                const bool take = lane + o < 32 && tb == b;
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    const double tv = __shfl_down_sync(0xffffffffu, v[i], o);
                    if (take) v[i] += tv;
                }
            }
            const int prev = __shfl_up_sync(0xffffffffu, b, 1);
            if (b != 0 && (lane == 0 || prev != b)) {
#pragma unroll
                for (int i = 0; i < NV; ++i) atomicAdd(&sbin[i * Nbin + b - 1], v[i]);
            }
            if (MODE == 1) {
                // (k,mu) table: |mu| is monotonic in kx along a row up to float32 rounding, so equal (bin, mu bin) keys form runs --
                // but a one-ulp wiggle at a bin edge may split a key into two runs: the reduction is segmented by RUN (head flags),
                // not by key equality, so every run is summed exactly once whatever the key order.  One atomic per run and quantity
                // instead of one per mode (float64 shared-memory atomics are CAS loops, and the lanes of a run hit one address).
                const int pk_ = __shfl_up_sync(0xffffffffu, tkey, 1);
                const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || pk_ != tkey);
                if (__any_sync(0xffffffffu, tkey != 0)) {
                    const int rid = __popc(heads & (0xffffffffu >> (31 - lane)));
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int tr = __shfl_down_sync(0xffffffffu, rid, o);
                        const bool take = lane + o < 32 && tr == rid;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const double t = __shfl_down_sync(0xffffffffu, tv[i], o);
                            if (take) tv[i] += t;
                        }
                    }
                    if (tkey != 0 && ((heads >> lane) & 1u)) {
                        const int imu = tkey & 1023, bb = tkey >> 10;
                        double* t = tab + (long long)(imu - 1) * Nbin + (bb - 1);
                        const long long tb = (long long)Nbin * in.Nmu;
                        atomicAdd(t, tv[0]); atomicAdd(t + tb, tv[1]); atomicAdd(t + 2 * tb, tv[2]); atomicAdd(t + 3 * tb, tv[3]);
                    }
                }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nsm; i += blockDim.x)
        if (sbin[i] != 0.0) atomicAdd(&out[i], sbin[i]);
}

int binned_spectra(const SpectraIn& in, double* out, cudaStream_t st)
{
    if (in.N < 2 || in.N % 2 || in.Nbin < 1 || in.ny < 1 || in.ky0 < 0 || in.ky0 + in.ny > in.N) return PSB_ERR_ARG;
    const size_t nout = in.mode == 0 ? 3 * (size_t)in.Nbin : (5 + 4 * (size_t)in.Nmu) * in.Nbin;
    if (cudaMemsetAsync(out, 0, nout * sizeof(double), st) != cudaSuccess) return PSB_ERR_CUDA;
    if (in.mode == 0) {
        const size_t smem = 3 * (size_t)in.Nbin * sizeof(double);
        if (smem > 200 * 1024) return PSB_ERR_ARG;
        if (cudaFuncSetAttribute(k_spectra<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return PSB_ERR_CUDA;
        k_spectra<0><<<sm_count() * 4, 256, smem, st>>>(in, out, 0);
    } else {
        size_t smem = nout * sizeof(double);
        int tab_smem = 1;
        // the (k,mu) table stays in shared memory only while it leaves room for 8 CTAs per SM: the per-mode arithmetic (IEEE float32
        // sqrt / div, float64 Legendre terms) is latency bound at low occupancy, and after the warp-level run reduction the table sees
        // few atomics, which the L2 handles natively for float64
        if (smem > 24 * 1024) { smem = 5 * (size_t)in.Nbin * sizeof(double); tab_smem = 0; }
        if (smem > 200 * 1024) return PSB_ERR_ARG;
        if (cudaFuncSetAttribute(k_spectra<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return PSB_ERR_CUDA;
        k_spectra<1><<<sm_count() * 8, 256, smem, st>>>(in, out, tab_smem);
    }
    return cudaGetLastError() == cudaSuccess ? PSB_OK : PSB_ERR_CUDA;
}

// code='python' variant of _Pk_periodic_rsd (pyspectrum.py:545-626): a float64 loop over ALL N^3 modes of a full (reflected) field
// indexed [kx][ky][kz] with the absolute wave numbers |k_a| = min(i, N-i) (so mu >= 0), mu bin imu = ceil(mu/dmu) (mu = 0 falls in
// no bin and, as in the reference, contributes to neither the (k,mu) table nor p2k/p4k), Legendre sums over the mu-binned modes
// only.  Every float64 operation is written with a round-to-nearest intrinsic in numpy's evaluation order (no FMA contraction):
// the bin counts are bit exact.  A debugging variant in the reference (it prints every bin) -> plain global float64 atomics for
// the (k,mu) table, per-CTA shared bins for the five per-k sums.
struct KmuPyIn {
    const Cx<float>* full; int N; const unsigned short* bin; int Nbin, Nmu;
    double kf, dmu, cos_th, sin_th, cos_ph, sin_ph;
};

__global__ void __launch_bounds__(256) k_kmu_python(KmuPyIn in, double* out)
{
    extern __shared__ double sbin[];                       // nk, ksum, p0, p2, p4 [Nbin]
    const int N = in.N, Nbin = in.Nbin, Nmu = in.Nmu;
    for (int i = threadIdx.x; i < 5 * Nbin; i += blockDim.x) sbin[i] = 0.0;
    __syncthreads();
    double* tab = out + 5 * (long long)Nbin;               // Nkmu, kkmu, mukmu, pkmu [Nbin][Nmu]
    const long long tb = (long long)Nbin * Nmu, ntot = (long long)N * N * N;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < ntot; e += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(e % N);
        const long long r = e / N;
        const int b_ = (int)(r % N), a = (int)(r / N);
        const int ia = a < N - a ? a : N - a, ib = b_ < N - b_ ? b_ : N - b_, ic = c < N - c ? c : N - c;
        const int m = ia * ia + ib * ib + ic * ic;
        const int bin = in.bin[m];
        if (bin < 1 || bin > Nbin) continue;
        const double rk = __dmul_rn(in.kf, __dsqrt_rn((double)m));
        const double rkx = __dmul_rn(in.kf, (double)ia), rky = __dmul_rn(in.kf, (double)ib), rkz = __dmul_rn(in.kf, (double)ic);
        const double cos_t = __ddiv_rn(rkz, rk);
        const double sin_t = __dsqrt_rn(__dsub_rn(1.0, __dmul_rn(cos_t, cos_t)));
        double cc = 0.0;
        if (sin_t > 0.0) {
            const double den = __dmul_rn(rk, sin_t);
            cc = __dadd_rn(__dmul_rn(in.sin_ph, __ddiv_rn(rky, den)), __dmul_rn(in.cos_ph, __ddiv_rn(rkx, den)));
        }
        const double mu = __dadd_rn(__dmul_rn(in.cos_th, cos_t), __dmul_rn(__dmul_rn(in.sin_th, sin_t), cc));
        const int imu = (int)ceil(__ddiv_rn(mu, in.dmu));
        const Cx<float> d = in.full[e];
        const float ab = (float)sqrt((double)d.x * (double)d.x + (double)d.y * (double)d.y);       // np.absolute(complex64)
        const double pk = (double)__fmul_rn(ab, ab);
        atomicAdd(&sbin[bin - 1], 1.0);
        atomicAdd(&sbin[Nbin + bin - 1], rk);
        atomicAdd(&sbin[2 * Nbin + bin - 1], pk);
        if (imu >= 1 && imu <= Nmu) {
            const double mu2 = __dmul_rn(mu, mu);
            const double L2 = __dadd_rn(-0.5, __dmul_rn(1.5, mu2));
            const double L4 = __dadd_rn(__dsub_rn(0.375, __dmul_rn(3.75, mu2)), __dmul_rn(4.375, __dmul_rn(mu2, mu2)));
            atomicAdd(&sbin[3 * Nbin + bin - 1], __dmul_rn(pk, L2));
            atomicAdd(&sbin[4 * Nbin + bin - 1], __dmul_rn(pk, L4));
            double* t = tab + (long long)(bin - 1) * Nmu + (imu - 1);
            atomicAdd(t, 1.0);
            atomicAdd(t + tb, rk);
            atomicAdd(t + 2 * tb, mu);
            atomicAdd(t + 3 * tb, pk);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 5 * Nbin; i += blockDim.x)
        if (sbin[i] != 0.0) atomicAdd(&out[i], sbin[i]);
}

int kmu_python(const Cx<float>* full, int N, const unsigned short* bin, int Nbin, int Nmu, double kf, const double* trig4, double* out,
               cudaStream_t st)
{
    if (!full || !bin || !out || !trig4 || N < 2 || N % 2 || Nbin < 1 || Nmu < 1) return PSB_ERR_ARG;
    const size_t nout = (5 + 4 * (size_t)Nmu) * Nbin, smem = 5 * (size_t)Nbin * sizeof(double);
    if (smem > 200 * 1024) return PSB_ERR_ARG;
    if (cudaMemsetAsync(out, 0, nout * sizeof(double), st) != cudaSuccess) return PSB_ERR_CUDA;
    if (cudaFuncSetAttribute(k_kmu_python, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return PSB_ERR_CUDA;
    KmuPyIn in;
    in.full = full; in.N = N; in.bin = bin; in.Nbin = Nbin; in.Nmu = Nmu; in.kf = kf; in.dmu = 1.0 / (double)Nmu;
    in.cos_th = trig4[0]; in.sin_th = trig4[1]; in.cos_ph = trig4[2]; in.sin_ph = trig4[3];
    k_kmu_python<<<sm_count() * 4, 256, smem, st>>>(in, out);
    return cudaGetLastError() == cudaSuccess ? PSB_OK : PSB_ERR_CUDA;
}

// Nk[j] = #{k on the full grid : irk(k) == j}, j = 0..nshell-1   (pyspectrum.py:380)
__global__ void __launch_bounds__(256) k_shell_counts(int N, const unsigned short* irk, int nshell, unsigned long long* nk)
{
    extern __shared__ unsigned int hist[];
    for (int i = threadIdx.x; i < nshell; i += blockDim.x) hist[i] = 0u;
    __syncthreads();
    const int h = N / 2;
    const long long nmode = (long long)(h + 1) * N * N;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < nmode; e += (long long)gridDim.x * blockDim.x) {
        const int ix = (int)(e % (h + 1));
        const long long r = e / (h + 1);
        const int ky = kfreq((int)(r % N), N), kz = kfreq((int)(r / N), N);
        const int s = irk[ix * ix + ky * ky + kz * kz];
        if (s < nshell) atomicAdd(&hist[s], (ix == 0 || ix == h) ? 1u : 2u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nshell; i += blockDim.x)
        if (hist[i]) atomicAdd(&nk[i], (unsigned long long)hist[i]);
}

int shell_mode_counts(int N, const unsigned short* irk, int nshell, unsigned long long* nk, cudaStream_t st)
{
    if (N < 2 || N % 2 || nshell < 1 || nshell > 8192) return PSB_ERR_ARG;
    if (cudaMemsetAsync(nk, 0, nshell * sizeof(unsigned long long), st) != cudaSuccess) return PSB_ERR_CUDA;
    k_shell_counts<<<sm_count() * 4, 256, nshell * sizeof(unsigned int), st>>>(N, irk, nshell, nk);
    return cudaGetLastError() == cudaSuccess ? PSB_OK : PSB_ERR_CUDA;
}

// psum[j] = sum over the FULL grid of |delta(k)|^2 for modes in shell j+1 (the self-conjugate points are made real first, as
// reflect_delta does).  One warp per (kz,ky) row of the half field; rows whose smallest |k| already lies beyond the last shell
// are skipped without a load.  Along a row the shell index is non-decreasing in kx, so a 5-step segmented shuffle reduction
// leaves one shared-memory atomic per (warp iteration, shell); blocks flush their bins with one global atomic per shell.
constexpr int SHELL_POWER_MAXBINS = 1024;
__global__ void __launch_bounds__(256) k_shell_power(const Cx<float>* __restrict__ half, int N, const unsigned short* __restrict__ irk,
                                                    int nshell, double* psum)
{
    __shared__ double bins[SHELL_POWER_MAXBINS];
    for (int i = threadIdx.x; i < nshell; i += blockDim.x) bins[i] = 0.0;
    __syncthreads();
    const int h = N / 2, lane = threadIdx.x & 31;
    const int nrow = N * N, nwarp = gridDim.x * (blockDim.x >> 5);
    for (int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < nrow; r += nwarp) {
        const int iz = r / N, iy = r - iz * N;
        const int ky = kfreq(iy, N), kz = kfreq(iz, N);
        const int m0 = ky * ky + kz * kz;
        if (irk[m0] > nshell) continue;                          // irk is non-decreasing in m: the whole row is outside
        const bool selfyz = (iy == 0 || iy == h) && (iz == 0 || iz == h);
        const Cx<float>* row = half + (long long)r * (h + 1);
        for (int x0 = 0; x0 <= h; x0 += 32) {
            const int ix = x0 + lane;
            int sh = 0;
            double v = 0.0;
            if (ix <= h) {
                sh = irk[m0 + ix * ix];
                if (sh >= 1 && sh <= nshell) {
                    Cx<float> d = row[ix];
                    const bool edge = ix == 0 || ix == h;
                    if (edge && selfyz) d.y = 0.f;
                    v = (edge ? 1.0 : 2.0) * ((double)d.x * d.x + (double)d.y * d.y);
                } else {
                    sh = 0;
                }
            }
            if (__ballot_sync(0xffffffffu, sh != 0) == 0u) { if (irk[m0 + x0 * x0] > nshell) break; else continue; }
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double tv = __shfl_down_sync(0xffffffffu, v, o);
                const int ts = __shfl_down_sync(0xffffffffu, sh, o);
                if (lane + o < 32 && ts == sh) v += tv;
            }
            const int prev = __shfl_up_sync(0xffffffffu, sh, 1);
            if (sh != 0 && (lane == 0 || prev != sh)) atomicAdd(&bins[sh - 1], v);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nshell; i += blockDim.x)
        if (bins[i] != 0.0) atomicAdd(&psum[i], bins[i]);
}

int shell_power(const Cx<float>* half, int N, const unsigned short* irk, int nshell, double* psum, cudaStream_t st)
{
    if (!half || !irk || !psum || N < 2 || N % 2 || nshell < 1 || nshell > SHELL_POWER_MAXBINS) return PSB_ERR_ARG;
    if (cudaMemsetAsync(psum, 0, nshell * sizeof(double), st) != cudaSuccess) return PSB_ERR_CUDA;
    k_shell_power<<<sm_count() * 4, 256, 0, st>>>(half, N, irk, nshell, psum);
    return cudaGetLastError() == cudaSuccess ? PSB_OK : PSB_ERR_CUDA;
}

// scales[j] = 2^round(log2(target / sqrt(psum[j]))): exact power-of-two normalisation of shell field j, whose rms
// over the grid is sqrt(sum_{k in shell} |delta|^2) by Parseval.  Empty shells get 1.
__global__ void k_shell_scales(const double* psum, int nshell, float target, float* scales)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nshell) return;
    const double p = psum[j];
    float s = 1.f;
    if (p > 0.0) s = exp2f(rintf(log2f(target / (float)sqrt(p))));
    scales[j] = s;
}

// Low-|k| modes of a ky-slab of the half field on grid N -> the half field of a (coarser or equal) carrier grid Ng, zero elsewhere.
// Ng < N keeps |k_a| < Ng/2 (the carrier's Nyquist planes stay empty); Ng == N copies the slab's rows.  Every rank writes the
// modes it owns into its own zeroed copy; the sum over ranks (disjoint supports) is the carrier field.
__global__ void __launch_bounds__(256) k_half_extract(const Cx<float>* __restrict__ src, int N, int ky0, int ny, Cx<float>* __restrict__ dst, int Ng)
{
    const int h = N / 2, hg = Ng / 2;
    const long long n = (long long)Ng * Ng * (hg + 1);
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int kx = (int)(e % (hg + 1));
        const long long r = e / (hg + 1);
        const int jy = (int)(r % Ng), jz = (int)(r / Ng);
        const int ky = kfreq(jy, Ng), kz = kfreq(jz, Ng);
        if (Ng != N && (kx >= hg || jy == hg || jz == hg)) continue;
        const int iy = ky < 0 ? ky + N : ky, iz = kz < 0 ? kz + N : kz;
        if (iy < ky0 || iy >= ky0 + ny) continue;
        dst[e] = src[((long long)iz * ny + (iy - ky0)) * (h + 1) + kx];
    }
}

int half_extract(const Cx<float>* src, int N, int ky0, int ny, Cx<float>* dst, int Ng, cudaStream_t st)
{
    if (!src || !dst || N < 2 || N % 2 || Ng < 2 || Ng % 2 || Ng > N || ny < 1 || ky0 < 0 || ky0 + ny > N) return PSB_ERR_ARG;
    if (cudaMemsetAsync(dst, 0, sizeof(Cx<float>) * (size_t)Ng * Ng * (Ng / 2 + 1), st) != cudaSuccess) return PSB_ERR_CUDA;
    k_half_extract<<<sm_count() * 8, 256, 0, st>>>(src, N, ky0, ny, dst, Ng);
    return cudaGetLastError() == cudaSuccess ? PSB_OK : PSB_ERR_CUDA;
}

int shell_scales(const double* psum, int nshell, float target_rms, float* scales, cudaStream_t st)
{
    if (!psum || !scales || nshell < 1) return PSB_ERR_ARG;
    k_shell_scales<<<(nshell + 127) / 128, 128, 0, st>>>(psum, nshell, target_rms, scales);
    return cudaGetLastError() == cudaSuccess ? PSB_OK : PSB_ERR_CUDA;
}

}  // namespace psb
