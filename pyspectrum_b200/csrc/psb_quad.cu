// psb_quad.cu -- the quadrupole-field pieces of estimator.f: the Q_ij / Q_ijkl particle weights of
// assign_quad (f:294-300) and the k-space combinations FiveDelta2g_1, FiveDelta2g_2, build_quad (f:514-603).
// All of it is element-wise float32 work on HBM-resident arrays; every operation is an explicit round-to-nearest
// intrinsic in the Fortran's evaluation order (no FMA contraction), so the results equal the C restatement bit for bit.
//
// estimator.f has no `implicit none`: in FiveDelta2g_1/_2 the unit-vector components kxh, kyh, kzh are never declared and
// start with `k`, hence INTEGER -- `kxh=float(ikx)/rk` truncates to -1, 0 or 1 (non-zero only for modes on the kx axis).
// The reference's arithmetic is what is reproduced here; `amu` in build_quad is declared real.
#include <cuda_runtime.h>
#include "psb_kernels.h"

namespace psb {

// we(i) of f:292-300 for ia..id != 0; r is the Fortran (3,np) array (12 bytes per particle)
__global__ void __launch_bounds__(256) k_quad_weights(const float* __restrict__ r, const float* __restrict__ w, long long np,
                                                      int ia, int ib, int ic, int id, float* __restrict__ we)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < np; i += (long long)gridDim.x * blockDim.x) {
        const float x = r[3 * i], y = r[3 * i + 1], z = r[3 * i + 2];
        const float c[3] = { x, y, z };
        const float rnorm = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
        float v = __fmul_rn(__fmul_rn(w[i], c[ia - 1]), c[ib - 1]);
        if (ic == 0 && id == 0) {
            v = __fdiv_rn(v, rnorm);
        } else {
            v = __fmul_rn(__fmul_rn(v, c[ic - 1]), c[id - 1]);
            v = __fdiv_rn(v, __fmul_rn(rnorm, rnorm));
        }
        we[i] = v;
    }
}

int quad_weights(const float* r, const float* w, long long np, int ia, int ib, int ic, int id, float* we, cudaStream_t st)
{
    const bool two = ic == 0 && id == 0;
    if (!r || !w || !we || np < 0 || ia < 1 || ia > 3 || ib < 1 || ib > 3) return PSB_ERR_ARG;
    if (!two && (ic < 1 || ic > 3 || id < 1 || id > 3)) return PSB_ERR_ARG;
    if (np == 0) return PSB_OK;
    const long long nb = (np + 255) / 256;
    k_quad_weights<<<(unsigned)(nb < 16LL * sm_count() ? nb : 16LL * sm_count()), 256, 0, st>>>(r, w, np, ia, ib, ic, id, we);
    return cudaGetLastError() == cudaSuccess ? PSB_OK : PSB_ERR_CUDA;
}

// signed wavenumber of 0-based index i (f:524-529: mod(i+N/2-2,N)-N/2+1 with the 1-based i): 0..N/2, then -N/2+1..-1
__device__ __forceinline__ int signed_k(int i, int N) { return i <= N / 2 ? i : i - N; }

struct QuadIn {
    const Cx<float>* a;      // FiveDelta2g_1: dcgyy        FiveDelta2g_2: dcg        build_quad: dclr1
    const Cx<float>* b;      //                dcgzz                       dcgxy
    const Cx<float>* c;      //                                            dcgyz
    const Cx<float>* d;      //                                            dcgzx
    Cx<float>* out;          //                dcgxx (inout)               dcgxx (inout)             dclr2 (inout)
    int N, mode, irsd;       // mode 1, 2: FiveDelta2g_1 / _2; 3: build_quad
};

__device__ __forceinline__ Cx<float> cscale(Cx<float> v, float s) { return Cx<float>{ __fmul_rn(v.x, s), __fmul_rn(v.y, s) }; }
__device__ __forceinline__ Cx<float> cadd(Cx<float> u, Cx<float> v) { return Cx<float>{ __fadd_rn(u.x, v.x), __fadd_rn(u.y, v.y) }; }
__device__ __forceinline__ Cx<float> csub(Cx<float> u, Cx<float> v) { return Cx<float>{ __fsub_rn(u.x, v.x), __fsub_rn(u.y, v.y) }; }

// one thread per element of the half arrays [iz][iy][ix], ix = 0..N/2 fastest (the Fortran (N/2+1,N,N) memory)
__global__ void __launch_bounds__(256) k_quad_fields(QuadIn q)
{
    const int N = q.N, hx = N / 2 + 1;
    const long long n = (long long)hx * N * N;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int ix = (int)(e % hx);
        const long long t = e / hx;
        const int iy = (int)(t % N), iz = (int)(t / N);
        const int kx = signed_k(ix, N), ky = signed_k(iy, N), kz = signed_k(iz, N);
        const float rk = __fsqrt_rn((float)(kx * kx + ky * ky + kz * kz));
        if (!(rk > 0.f)) continue;
        if (q.mode == 3) {                                                    // f:590-599
            const float amu = __fdiv_rn((float)(q.irsd == 3 ? kz : (q.irsd == 2 ? ky : kx)), rk);
            const float fac = __fsub_rn(__fmul_rn(7.5f, __fmul_rn(amu, amu)), 2.5f);
            q.out[e] = cscale(q.a[e], fac);
            continue;
        }
        const int kxh = (int)__fdiv_rn((float)kx, rk), kyh = (int)__fdiv_rn((float)ky, rk), kzh = (int)__fdiv_rn((float)kz, rk);   // INTEGER (implicit)
        if (q.mode == 1) {                                                    // f:530-537
            Cx<float> s = cscale(q.out[e], (float)(kxh * kxh));
            s = cadd(s, cscale(q.a[e], (float)(kyh * kyh)));
            s = cadd(s, cscale(q.b[e], (float)(kzh * kzh)));
            q.out[e] = cscale(s, 7.5f);
        } else {                                                              // f:562-571
            Cx<float> s = cscale(cscale(cscale(q.b[e], 2.f), (float)kxh), (float)kyh);
            s = cadd(s, cscale(cscale(cscale(q.c[e], 2.f), (float)kyh), (float)kzh));
            s = cadd(s, cscale(cscale(cscale(q.d[e], 2.f), (float)kzh), (float)kxh));
            q.out[e] = csub(cadd(q.out[e], cscale(s, 7.5f)), cscale(q.a[e], 2.5f));
        }
    }
}

int quad_fields(int mode, const Cx<float>* a, const Cx<float>* b, const Cx<float>* c, const Cx<float>* d, Cx<float>* out, int N, int irsd,
                cudaStream_t st)
{
    if (N < 2 || N % 2 || N > 4096 || !out || !a) return PSB_ERR_ARG;
    if (mode == 1 && !b) return PSB_ERR_ARG;
    if (mode == 2 && (!b || !c || !d)) return PSB_ERR_ARG;
    if (mode == 3 && (irsd < 1 || irsd > 3)) return PSB_ERR_ARG;             // the Fortran stops
    if (mode < 1 || mode > 3) return PSB_ERR_ARG;
    QuadIn q{ a, b, c, d, out, N, mode, irsd };
    const long long n = (long long)(N / 2 + 1) * N * N, nb = (n + 255) / 256;
    k_quad_fields<<<(unsigned)(nb < 16LL * sm_count() ? nb : 16LL * sm_count()), 256, 0, st>>>(q);
    return cudaGetLastError() == cudaSuccess ? PSB_OK : PSB_ERR_CUDA;
}

}  // namespace psb
