// psb_triangles.cu -- K6: triangle triple-product sums  S[t] = sum_x I_i(x) I_j(x) I_l(x)
// Replaces the einsum loop of pyspectrum.py:415-430 (and 1017-1023 / estimator.f:90-100 for the counts).
//
// FFMA formulation: the (i,j,l) index space is cut into 4x4x4 tiles (host-built list of the tiles that
// contain at least one wanted triangle).  A CTA stages an x-chunk of all shell fields in shared memory;
// each warp owns a subset of tiles; lanes run over x: 12 conflict-free LDS, 16 pair products and
// 64 FMAs per x per tile, accumulators in registers, then a 62-shuffle transposing reduction and a
// CTA-private read-modify-write of the tile's 64 partial sums in an L2-resident float64 buffer.
// A second kernel folds the per-CTA partials into the output.  Cross-chunk accumulation is float64;
// within a chunk a lane adds XC/32 float32 products (float64 throughout for the counts path).
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "psb_kernels.h"

namespace psb {

constexpr int TILE_INTS = 4 + 64;     // i0, j0, l0, pad, slot[64] (output index or -1); i0.. are 0-based field slots

// cell x of a field row; packed rows hold, per aligned cell pair (x, x+1), the words {half2 hi(x,x+1), half2 lo(x,x+1)}
__device__ __forceinline__ float load_cell(const float* src, int x, int packed_half)
{
    if (!packed_half) return src[x];
    const unsigned int hw = __float_as_uint(src[x & ~1]), lw = __float_as_uint(src[x | 1]);
    const unsigned short h = (x & 1) ? (unsigned short)(hw >> 16) : (unsigned short)(hw & 0xffffu);
    const unsigned short l = (x & 1) ? (unsigned short)(lw >> 16) : (unsigned short)(lw & 0xffffu);
    return __half2float(__ushort_as_half(h)) + __half2float(__ushort_as_half(l));
}
__device__ __forceinline__ double load_cell(const double* src, int x, int) { return src[x]; }

template <typename T, int NT>
__global__ void __launch_bounds__(NT, 1) k_tri(const T* const* __restrict__ fields, int S, long long ncell, int XC,
                                             const int* __restrict__ tiles, int ntiles, double* __restrict__ partial, int packed_half)
{
    extern __shared__ __align__(16) unsigned char tri_smem[];
    T* fs = reinterpret_cast<T*>(tri_smem);                    // [S][XC]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = NT / 32;
    double* mypart = partial + (size_t)blockIdx.x * ntiles * 64;
    const long long nchunk = (ncell + XC - 1) / XC;
    for (long long chunk = blockIdx.x; chunk < nchunk; chunk += gridDim.x) {
        const long long x0 = chunk * XC;
        __syncthreads();
        for (int f = 0; f < S; ++f) {
            const T* src = fields[f] + x0;
            for (int x = threadIdx.x; x < XC; x += NT) fs[f * XC + x] = (x0 + x < ncell) ? load_cell(src, x, packed_half) : (T)0;
        }
        __syncthreads();
        for (int tile = warp; tile < ntiles; tile += nwarp) {
            const int* td = tiles + (size_t)tile * TILE_INTS;
            const T* pi = fs + td[0] * XC;
            const T* pj = fs + td[1] * XC;
            const T* pl = fs + td[2] * XC;
            T acc[64];
#pragma unroll
            for (int q = 0; q < 64; ++q) acc[q] = (T)0;
            for (int x = lane; x < XC; x += 32) {
                T fi[4], fj[4], fl[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) { fi[a] = pi[a * XC + x]; fj[a] = pj[a * XC + x]; fl[a] = pl[a * XC + x]; }
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const T pr = fi[a] * fj[b];
#pragma unroll
                        for (int c = 0; c < 4; ++c) acc[(a * 4 + b) * 4 + c] += pr * fl[c];
                    }
            }
            // transposing butterfly: after offsets 16,8,4,2,1 lane L holds the totals of entries 2L, 2L+1
#pragma unroll
            for (int o = 16, n = 64; o > 0; o >>= 1, n >>= 1) {
                const bool up = (lane & o) != 0;
#pragma unroll
                for (int k = 0; k < n / 2; ++k) {
                    const T send = up ? acc[k] : acc[k + n / 2];
                    const T keep = up ? acc[k + n / 2] : acc[k];
                    acc[k] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                }
            }
            double2* dst = reinterpret_cast<double2*>(mypart + (size_t)tile * 64) + lane;
            double2 cur = *dst;
            cur.x += (double)acc[0];
            cur.y += (double)acc[1];
            *dst = cur;
        }
    }
}

__global__ void k_tri_fold(const double* __restrict__ partial, int ncta, const int* __restrict__ tiles, int ntiles, double* sums)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ntiles * 64) return;
    const int tile = e >> 6, q = e & 63;
    const int slot = tiles[(size_t)tile * TILE_INTS + 4 + q];
    if (slot < 0) return;
    double s = 0.0;
    for (int c = 0; c < ncta; ++c) s += partial[(size_t)c * ntiles * 64 + e];
    sums[slot] = s;
}

static int tri_ctas() { return sm_count(); }

size_t triangle_workspace_bytes(int ntiles)
{
    return (size_t)tri_ctas() * ntiles * 64 * sizeof(double);
}

template <typename T>
int triangle_sums_tiles(const T* const* fields, int S, long long ncell, const int* tiles, int ntiles,
                        double* sums, void* ws, size_t ws_bytes, int packed_half, cudaStream_t st)
{
    if (S < 4 || ntiles < 1 || ncell < 1) return PSB_ERR_ARG;
    if (ws_bytes < triangle_workspace_bytes(ntiles)) return PSB_ERR_WORKSPACE;
    constexpr int NT = sizeof(T) == 4 ? 512 : 256;
    int XC = 1024;
    while ((size_t)S * XC * sizeof(T) > 200 * 1024 && XC > 32) XC >>= 1;
    const size_t smem = (size_t)S * XC * sizeof(T);
    auto kern = k_tri<T, NT>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return PSB_ERR_CUDA;
    double* partial = static_cast<double*>(ws);
    if (cudaMemsetAsync(partial, 0, triangle_workspace_bytes(ntiles), st) != cudaSuccess) return PSB_ERR_CUDA;
    const int ncta = tri_ctas();
    kern<<<ncta, NT, smem, st>>>(fields, S, ncell, XC, tiles, ntiles, partial, packed_half);
    k_tri_fold<<<(ntiles * 64 + 255) / 256, 256, 0, st>>>(partial, ncta, tiles, ntiles, sums);
    return cudaGetLastError() == cudaSuccess ? PSB_OK : PSB_ERR_CUDA;
}

template int triangle_sums_tiles<float>(const float* const*, int, long long, const int*, int, double*, void*, size_t, int, cudaStream_t);
template int triangle_sums_tiles<double>(const double* const*, int, long long, const int*, int, double*, void*, size_t, int, cudaStream_t);

}  // namespace psb
