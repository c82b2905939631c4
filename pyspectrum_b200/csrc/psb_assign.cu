// psb_assign.cu -- K1: PCS (4th-order) mass assignment onto two half-cell interlaced grids.
// Replaces estimator.f:284-512 (assign_quad, delta branch) and the clip/cast of pyspectrum.py:938-941.
//
// Design (sm_100a): shared-memory float atomics are CAS spin loops on this part (ATOMS.CAST.SPIN), so nothing here uses them.
//   dense catalogues (>= 0.1 particles per cell): counting sort by 8 x 8 x 4 TILE of the particle's cell, then k_assign_tile -- one
//       warp owns one tile (plus halo) in its own shared memory, lanes own distinct stencil points, plain LDS / FFMA / STS, and the
//       finished tile leaves with coalesced red.global.add.v4.f32 (REDG.E.ADD.F32x4); runs at the shared-memory data pipe;
//   sparse catalogues / chunks of a streamed upload: counting sort by (z,y) row, then k_assign_tri -- one vector reduction per
//       aligned pair of interlaced cells {A(c),B(c),A(c+1),B(c+1)} and row, three lanes per particle (<= 75, typically ~56
//       reductions per particle instead of 128 scalar read-modify-writes);
//   the sort is one pass (k_sort_scatter) or, for unordered input on large grids, two (k_partition_coarse + k_bucket_scatter),
//       chosen on the device from how ordered the input is (k_hist / k_sort_decide);
//   slab mode (multi-GPU): the same kernels on a window of planes; k_route_count / k_route_scatter(_peer) deal the particles (with
//       ghost copies) to the ranks that own the planes they touch, the peer variant by NVLink stores into the receive buffers.
#include <cuda_runtime.h>
#include <cstdlib>
#include "psb_kernels.h"

namespace psb {

__device__ __forceinline__ void load_particle(const AssignIn& a, long long i, float& x, float& y, float& z, float& w, double& wd)
{
    if (a.pos_aos == 2) {            // routed particles of the slab path: float4 {x, y, z, w}, already clipped and cast
        const float4 p = static_cast<const float4*>(a.pos)[i];
        x = p.x; y = p.y; z = p.z; w = p.w; wd = (double)p.w;
        return;
    }
    const long long s = a.pos_aos ? 1 : a.Np, b = a.pos_aos ? 3 * i : i;
    if (a.pos_f64) {
        const double* p = static_cast<const double*>(a.pos);
        double dx = p[b], dy = p[b + s], dz = p[b + 2 * s];
        if (a.do_clip) {
            dx = fmin(fmax(dx, 0.0), a.clip_hi); dy = fmin(fmax(dy, 0.0), a.clip_hi); dz = fmin(fmax(dz, 0.0), a.clip_hi);
        }
        x = (float)dx; y = (float)dy; z = (float)dz;
    } else {
        const float* p = static_cast<const float*>(a.pos);
        x = p[b]; y = p[b + s]; z = p[b + 2 * s];
        if (a.do_clip) {
            const float hi = (float)a.clip_hi;
            x = fminf(fmaxf(x, 0.f), hi); y = fminf(fmaxf(y, 0.f), hi); z = fminf(fmaxf(z, 0.f), hi);
        }
    }
    if (a.w) {
        if (a.w_f64) { wd = static_cast<const double*>(a.w)[i]; w = (float)wd; }
        else { w = static_cast<const float*>(a.w)[i]; wd = (double)w; }
    } else { w = 1.f; wd = 1.0; }
}

// 1-based continuous grid coordinate of estimator.f:302-304 (no fused multiply-add: the integer
// cell must come out as in the reference)
__device__ __forceinline__ float grid_coord(float kf_ks, float r, float offset)
{
    return __fadd_rn(__fadd_rn(__fmul_rn(kf_ks, r), 1.f), offset);
}

__device__ __forceinline__ int wrapN(int c, int N) { c %= N; return c < 0 ? c + N : c; }

// Sort key = (z plane, y row) of the particle's cell.  Full grid: N*N keys.  Slab [zbase, zbase+nzs): planes are counted from
// zbase - 4 (a particle in cell c touches planes c-1 .. c+3, so cells zbase-3 .. zbase+nzs reach the slab): (nzs+8)*N keys plus
// one trailing bucket for particles that cannot touch the slab (they are sorted behind the valid ones and never assigned).
__device__ __forceinline__ int row_key(const AssignIn& a, float y, float z)
{
    const int cy = (int)grid_coord(a.kf_ks, y, a.offset) - 1, cz = (int)grid_coord(a.kf_ks, z, a.offset) - 1;
    if (a.nzs >= a.N) return wrapN(cz, a.N) * a.N + wrapN(cy, a.N);
    const int zr = wrapN(cz - (a.zbase - 4), a.N);
    return zr < a.nzs + 8 ? zr * a.N + wrapN(cy, a.N) : (a.nzs + 8) * a.N;
}

// Tile scatter (k_assign_tile): particles are sorted by the TX x TY x TZ tile of their cell instead of by (z,y) row.
constexpr int TX = 8, TY = 8, TZ = 4;
constexpr int AX = TX + 4, AY = TY + 4, AZ = TZ + 4;        // tile + halo: a particle in cell c touches cells c-1 .. c+3
constexpr int SY = AX, SZ = AX * AY;                        // strides 12 and 144 (== 16 mod 32): a 4 x 4 x 2 block of lanes hits 32 banks
constexpr int TILE_WORDS = AX * AY * AZ;                    // 1152 floats per grid
constexpr int STG = 44;                                     // staging words per particle: 11 float4 (row starts an odd number of
                                                            // 16-byte units apart: conflict-free 128-bit stores)
constexpr int TILE_WARPS = 5;                               // warps per CTA: 5 x 14.7 KB, three CTAs per SM

struct TileGeom { int ntx, nty, ntz, zlo, nkeys; };         // zlo: first plane the z tiles are counted from; nkeys = valid tiles

__host__ __device__ __forceinline__ TileGeom tile_geom(int N, int zbase, int nzs)
{
    TileGeom g;
    g.ntx = (N + TX - 1) / TX; g.nty = (N + TY - 1) / TY;
    const bool slab = nzs < N;
    g.zlo = slab ? zbase - 4 : 0;
    g.ntz = ((slab ? nzs + 8 : N) + TZ - 1) / TZ;
    g.nkeys = g.ntx * g.nty * g.ntz;
    return g;
}

__device__ __forceinline__ int tile_key(const AssignIn& a, float x, float y, float z)
{
    const TileGeom g = tile_geom(a.N, a.zbase, a.nzs);
    const int cx = wrapN((int)grid_coord(a.kf_ks, x, a.offset) - 1, a.N), cy = wrapN((int)grid_coord(a.kf_ks, y, a.offset) - 1, a.N);
    const int zr = wrapN((int)grid_coord(a.kf_ks, z, a.offset) - 1 - g.zlo, a.N);
    if (a.nzs < a.N && zr >= a.nzs + 8) return g.nkeys;      // cannot touch the slab: trailing bucket
    return ((zr / TZ) * g.nty + cy / TY) * g.ntx + cx / TX;
}

__device__ __forceinline__ int sort_key(const AssignIn& a, float x, float y, float z) { return a.tiles ? tile_key(a, x, y, z) : row_key(a, y, z); }

// `ordered` (optional): number of particles whose key is within 1 of their predecessor's in the input -- how spatially ordered
// the catalogue already is, which decides between the one-pass and the two-pass sort ON THE DEVICE (k_sort_decide: no host sync)
__global__ void k_hist(AssignIn a, unsigned int* hist, double* sumw, unsigned long long* ordered)
{
    double acc = 0.0;
    unsigned int near = 0;
    const long long nround = (a.Np + (long long)gridDim.x * blockDim.x - 1) / ((long long)gridDim.x * blockDim.x);
    for (long long it = 0; it < nround; ++it) {           // warp-uniform trip count: the shuffle below needs every lane
        const long long i = (it * gridDim.x + blockIdx.x) * (long long)blockDim.x + threadIdx.x;
        int key = -0x40000000;
        if (i < a.Np) {
            float x, y, z, w; double wd;
            load_particle(a, i, x, y, z, w, wd);
            acc += wd;
            key = sort_key(a, x, y, z);
            atomicAdd(&hist[key], 1u);
        }
        if (ordered) {
            const int prev = __shfl_up_sync(0xffffffffu, key, 1);
            if ((threadIdx.x & 31) && i < a.Np && (unsigned)(key - prev + 1) <= 2u) ++near;
        }
    }
    if (ordered) {
        for (int o = 16; o > 0; o >>= 1) near += __shfl_down_sync(0xffffffffu, near, o);
        if ((threadIdx.x & 31) == 0 && near) atomicAdd(ordered, (unsigned long long)near);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    __shared__ double red[32];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        acc = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        if (threadIdx.x == 0) atomicAdd(sumw, acc);
    }
}

// exclusive scan of the n = N^2 row counters in three coalesced steps: per-tile sums, scan of the tile sums (one CTA), per-tile
// exclusive scan + offset.  Tiles of SCAN_TILE counters, one CTA of 256 threads each (16 counters per thread).
constexpr int SCAN_TILE = 4096;

__device__ __forceinline__ unsigned int block_exclusive_scan_256(unsigned int v, unsigned int* total)
{
    __shared__ unsigned int wsum[8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) wsum[w] = x;
    __syncthreads();
    unsigned int base = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { const unsigned int sw = wsum[k]; if (k < w) base += sw; tot += sw; }
    __syncthreads();
    *total = tot;
    return base + x - v;
}

__global__ void __launch_bounds__(256) k_scan_tile_sums(const unsigned int* __restrict__ data, int n, unsigned int* __restrict__ tile_sum)
{
    const int t0 = blockIdx.x * SCAN_TILE;
    unsigned int s = 0;
    for (int i = threadIdx.x; i < SCAN_TILE; i += 256) { const int e = t0 + i; if (e < n) s += data[e]; }
    unsigned int tot;
    block_exclusive_scan_256(s, &tot);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(256) k_scan_tiles(unsigned int* tile_sum, int ntile)
{
    // ntile <= 256 * 16 (N <= 4096): every thread owns a contiguous run of 16 tile sums
    unsigned int loc[16], s = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) { const int e = threadIdx.x * 16 + k; loc[k] = e < ntile ? tile_sum[e] : 0u; s += loc[k]; }
    unsigned int tot;
    unsigned int run = block_exclusive_scan_256(s, &tot);
#pragma unroll
    for (int k = 0; k < 16; ++k) { const int e = threadIdx.x * 16 + k; if (e < ntile) tile_sum[e] = run; run += loc[k]; }
}

__global__ void __launch_bounds__(256) k_scan_apply(unsigned int* __restrict__ data, int n, const unsigned int* __restrict__ tile_off)
{
    const int t0 = blockIdx.x * SCAN_TILE + threadIdx.x * 16;            // 16 consecutive counters per thread (64-byte runs)
    unsigned int loc[16], s = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) { loc[k] = t0 + k < n ? data[t0 + k] : 0u; s += loc[k]; }
    unsigned int tot;
    unsigned int run = tile_off[blockIdx.x] + block_exclusive_scan_256(s, &tot);
#pragma unroll
    for (int k = 0; k < 16; ++k) { if (t0 + k < n) data[t0 + k] = run; run += loc[k]; }
}

// flag[0] = 1: two-pass sort, 0: one pass.  Fewer than a quarter of the particles next to their predecessor = unordered input.
__global__ void k_sort_decide(const unsigned long long* ordered, long long np, int* flag) { flag[0] = 4 * ordered[0] < (unsigned long long)np ? 1 : 0; }

__global__ void k_sort_scatter(AssignIn a, unsigned int* cursor, float4* sorted, const int* flag = nullptr, int want = 0)
{
    if (flag && *flag != want) return;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < a.Np; i += (long long)gridDim.x * blockDim.x) {
        float x, y, z, w; double wd;
        load_particle(a, i, x, y, z, w, wd);
        const unsigned int slot = atomicAdd(&cursor[sort_key(a, x, y, z)], 1u);
        sorted[slot] = make_float4(x, y, z, w);
    }
}

// Two-pass sort (randomly ordered catalogues on large grids).  The one-pass scatter above writes 16-byte particles to as many output
// streams as there are keys.  When the input is not already spatially ordered and the keys number 5*10^5 or more, the half-written
// 32-byte sectors of all streams do not survive in L2 until their other half arrives: ncu at 1024^3 / 5e8 particles shows 20.6 GB of
// DRAM reads (sector fills) + 16.6 GB of writes for 6 GB in / 8 GB out, 50 ps per particle instead of 14 (33 ps at 5.6e5 keys).
//   pass 0  k_partition_coarse: a CTA takes 4096 consecutive particles, ranks them per coarse bucket (2^gshift consecutive keys,
//           <= 1024 buckets) with shared-memory counters, scans the counts, reserves the tile's range of every bucket with one global
//           atomic, REORDERS the particles in shared memory and writes them out bucket by bucket: neighbouring lanes store
//           neighbouring 16 bytes (runs of ~8-16 particles), full sectors, ~10^3 streams;
//   pass 1  k_bucket_scatter: bucket by bucket (a few MB of output: L2 resident) in chunks of 8192 particles: count per key in shared
//           memory (the returned count is the rank inside the chunk), reserve the chunk's range of every key with one atomic on the
//           key cursor, store.  hist[] ends as after the one-pass scatter.
constexpr int PC_THREADS = 512, PC_ITEMS = 8, PC_TILE = PC_THREADS * PC_ITEMS, PC_MAXB = 1024;
constexpr int BS_THREADS = 512, BS_ITEMS = 16, BS_CHUNK = BS_THREADS * BS_ITEMS, BS_MAXD = 4352;

__global__ void __launch_bounds__(256) k_bucket_starts(const unsigned int* __restrict__ key_start, long long nrow, int gshift, int nb,
                                                       unsigned int np, unsigned int* __restrict__ bucket_start)
{
    const int b = blockIdx.x * 256 + threadIdx.x;
    if (b <= nb) bucket_start[b] = ((long long)b << gshift) < nrow ? key_start[(size_t)b << gshift] : np;
}

__global__ void __launch_bounds__(PC_THREADS, 2) k_partition_coarse(AssignIn a, int gshift, int nb, const unsigned int* __restrict__ bucket_start,
                                                                    unsigned int* coarse_cursor, float4* __restrict__ out, const int* flag)
{
    if (flag && *flag != 1) return;
    extern __shared__ __align__(16) unsigned char pc_smem[];
    float4* sp = reinterpret_cast<float4*>(pc_smem);                                  // [PC_TILE] particles in bucket order
    unsigned short* sb = reinterpret_cast<unsigned short*>(sp + PC_TILE);            // [PC_TILE] bucket of every slot
    unsigned int* cnt = reinterpret_cast<unsigned int*>(sb + PC_TILE);               // [PC_MAXB] count, then local start
    unsigned int* gof = cnt + PC_MAXB;                                                // [PC_MAXB] global start - local start
    __shared__ unsigned int wsum[PC_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long ntile = (a.Np + PC_TILE - 1) / PC_TILE;
    for (long long tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
        const long long lo = tile * PC_TILE;
        const int n = (int)(a.Np - lo < PC_TILE ? a.Np - lo : PC_TILE);
        for (int b = threadIdx.x; b < PC_MAXB; b += PC_THREADS) cnt[b] = 0u;
        __syncthreads();
        float4 p[PC_ITEMS];
        unsigned int br[PC_ITEMS];                        // bucket (low 10 bits) | rank inside the tile (high bits)
#pragma unroll
        for (int it = 0; it < PC_ITEMS; ++it) {
            const int i = it * PC_THREADS + threadIdx.x;
            if (i < n) {
                float x, y, z, w; double wd;
                load_particle(a, lo + i, x, y, z, w, wd);
                p[it] = make_float4(x, y, z, w);
                const unsigned int b = (unsigned)(sort_key(a, x, y, z) >> gshift);
                br[it] = b | (atomicAdd(&cnt[b], 1u) << 10);
            }
        }
        __syncthreads();
        // exclusive scan of the PC_MAXB counts (two per thread) and the global reservation of every non-empty bucket
        const unsigned int c0 = cnt[2 * threadIdx.x], c1 = cnt[2 * threadIdx.x + 1];
        unsigned int x = c0 + c1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) wsum[warp] = x;
        __syncthreads();
        unsigned int base = 0;
        for (int k = 0; k < warp; ++k) base += wsum[k];
        const unsigned int s0 = base + x - (c0 + c1), s1 = s0 + c0;
        __syncthreads();
        cnt[2 * threadIdx.x] = s0; cnt[2 * threadIdx.x + 1] = s1;
        if (c0) gof[2 * threadIdx.x] = bucket_start[2 * threadIdx.x] + atomicAdd(&coarse_cursor[2 * threadIdx.x], c0) - s0;
        if (c1) gof[2 * threadIdx.x + 1] = bucket_start[2 * threadIdx.x + 1] + atomicAdd(&coarse_cursor[2 * threadIdx.x + 1], c1) - s1;
        __syncthreads();
#pragma unroll
        for (int it = 0; it < PC_ITEMS; ++it) {
            const int i = it * PC_THREADS + threadIdx.x;
            if (i < n) {
                const unsigned int b = br[it] & 1023u, slot = cnt[b] + (br[it] >> 10);
                sp[slot] = p[it];
                sb[slot] = (unsigned short)b;
            }
        }
        __syncthreads();
        for (int slot = threadIdx.x; slot < n; slot += PC_THREADS) out[gof[sb[slot]] + slot] = sp[slot];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(BS_THREADS, 2) k_bucket_scatter(AssignIn a, int gshift, long long nrow, int pieces,
                                                                  const unsigned int* __restrict__ bucket_start, unsigned int* key_cursor,
                                                                  float4* out, const int* flag)
{
    if (flag && *flag != 1) return;
    __shared__ unsigned int cnt[BS_MAXD];
    __shared__ unsigned int off[BS_MAXD];
    const int b = blockIdx.x / pieces, j = blockIdx.x % pieces;
    const int kbase = b << gshift;
    const long long rest = nrow - kbase;
    const int nd = (int)(rest < (1LL << gshift) ? rest : (1LL << gshift));
    const long long hi = bucket_start[b + 1], stride = (long long)pieces * BS_CHUNK;
    for (long long lo = (long long)bucket_start[b] + (long long)j * BS_CHUNK; lo < hi; lo += stride) {
        const long long end = lo + BS_CHUNK < hi ? lo + BS_CHUNK : hi;
        for (int d = threadIdx.x; d < nd; d += BS_THREADS) cnt[d] = 0u;
        __syncthreads();
        unsigned int dr[BS_ITEMS];                       // key inside the bucket (low 16 bits) and rank inside the chunk (high 16 bits)
#pragma unroll
        for (int it = 0; it < BS_ITEMS; ++it) {
            const long long i = lo + it * BS_THREADS + threadIdx.x;
            if (i < end) {
                float x, y, z, w; double wd;
                load_particle(a, i, x, y, z, w, wd);
                const unsigned int d = (unsigned)(sort_key(a, x, y, z) - kbase);
                dr[it] = d | (atomicAdd(&cnt[d], 1u) << 16);
            }
        }
        __syncthreads();
        for (int d = threadIdx.x; d < nd; d += BS_THREADS) {
            const unsigned int c = cnt[d];
            if (c) off[d] = atomicAdd(&key_cursor[kbase + d], c);
        }
        __syncthreads();
#pragma unroll
        for (int it = 0; it < BS_ITEMS; ++it) {           // the chunk is re-read (L2 resident) rather than held in registers
            const long long i = lo + it * BS_THREADS + threadIdx.x;
            if (i < end) {
                float x, y, z, w; double wd;
                load_particle(a, i, x, y, z, w, wd);
                out[off[dr[it] & 0xffffu] + (dr[it] >> 16)] = make_float4(x, y, z, w);
            }
        }
        __syncthreads();
    }
}

// the four weights of estimator.f:316-320 for fractional offset h (cells c-1, c, c+1, c+2)
__device__ __forceinline__ void pcs4(float h, float& m2, float& m1, float& p1, float& p2)
{
    const float h2 = __fmul_rn(h, h);
    const float omh = __fsub_rn(1.f, h);
    m2 = __fmul_rn(__fmul_rn(omh, omh), omh);
    m1 = __fadd_rn(4.f, __fmul_rn(__fsub_rn(__fmul_rn(3.f, h), 6.f), h2));
    p2 = __fmul_rn(h2, h);
    p1 = __fsub_rn(__fsub_rn(__fsub_rn(6.f, m2), m1), p2);
}

// per-axis window of 5 cells starting at c-1: A weights in slots 0..3, B weights in slots sB..sB+3
struct AxisWin { int c0; float a[5], b[5]; };

__device__ __forceinline__ AxisWin axis_window(float kf_ks, float r, float offset)
{
    AxisWin wv;
    const float rx = grid_coord(kf_ks, r, offset), tx = __fadd_rn(rx, 0.5f);
    const int im1 = (int)rx, nm1 = (int)tx;
    float a0, a1, a2, a3, b0, b1, b2, b3;
    pcs4(__fsub_rn(rx, (float)im1), a0, a1, a2, a3);
    pcs4(__fsub_rn(tx, (float)nm1), b0, b1, b2, b3);
    const bool sh = nm1 != im1;            // grid-B base cell is c or c+1
    wv.c0 = im1 - 2;                       // 0-based cell c-1
    wv.a[0] = a0; wv.a[1] = a1; wv.a[2] = a2; wv.a[3] = a3; wv.a[4] = 0.f;
    wv.b[0] = sh ? 0.f : b0; wv.b[1] = sh ? b0 : b1; wv.b[2] = sh ? b1 : b2; wv.b[3] = sh ? b2 : b3; wv.b[4] = sh ? b3 : 0.f;
    return wv;
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(256) k_assign(const float4* __restrict__ sorted, long long Np, int N, float kf_ks, float offset, float* mesh)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= Np) return;
    const float4 p = sorted[i];
    const AxisWin X = axis_window(kf_ks, p.x, offset), Y = axis_window(kf_ks, p.y, offset), Z = axis_window(kf_ks, p.z, offset);
    // x: three aligned cell pairs cover the 5-cell window; slot q = 2*pair + {0,1} holds window cell q - off
    const int P0 = X.c0 >> 1, off = X.c0 - 2 * P0, Nh = N / 2;
    float xa[6], xb[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        const int k0 = q, k1 = q - 1;      // window cell if off == 0 / off == 1
        const float a0 = (k0 < 5) ? X.a[k0 < 5 ? k0 : 0] : 0.f, a1 = (k1 >= 0 && k1 < 5) ? X.a[k1 >= 0 && k1 < 5 ? k1 : 0] : 0.f;
        const float b0 = (k0 < 5) ? X.b[k0 < 5 ? k0 : 0] : 0.f, b1 = (k1 >= 0 && k1 < 5) ? X.b[k1 >= 0 && k1 < 5 ? k1 : 0] : 0.f;
        xa[q] = off ? a1 : a0;
        xb[q] = off ? b1 : b0;
    }
    long long xo[3];
#pragma unroll
    for (int pr = 0; pr < 3; ++pr) { int P = (P0 + pr) % Nh; if (P < 0) P += Nh; xo[pr] = 4LL * P; }
#pragma unroll
    for (int rz = 0; rz < 5; ++rz) {
        if (Z.a[rz] == 0.f && Z.b[rz] == 0.f) continue;
        const long long zo = (long long)wrapN(Z.c0 + rz, N) * N;
#pragma unroll
        for (int ry = 0; ry < 5; ++ry) {
            const float wa = __fmul_rn(__fmul_rn(Y.a[ry], Z.a[rz]), p.w), wb = __fmul_rn(__fmul_rn(Y.b[ry], Z.b[rz]), p.w);
            if (wa == 0.f && wb == 0.f) continue;
            float* row = mesh + (zo + wrapN(Y.c0 + ry, N)) * (2LL * N);
#pragma unroll
            for (int pr = 0; pr < 3; ++pr) {
                const float v0 = xa[2 * pr] * wa, v1 = xb[2 * pr] * wb, v2 = xa[2 * pr + 1] * wa, v3 = xb[2 * pr + 1] * wb;
                if (v0 != 0.f || v1 != 0.f || v2 != 0.f || v3 != 0.f) red_add_v4(row + xo[pr], v0, v1, v2, v3);
            }
        }
    }
}

// Lane-per-pair variant (PSB_ASSIGN_VARIANT=1; also the path for N > 1024): four consecutive lanes share a particle, lane pr < 3 owns the aligned cell pair P0 + pr of every row of the
// window (lane 3 idles).  The three vector reductions of a row then leave in ONE instruction from adjacent lanes, and the
// load/store unit merges their 48 contiguous bytes into two 32-byte sectors instead of three separate sector packets: a third
// fewer packets on the L1 -> L2 reduction path that bounds this kernel.  The windows are recomputed per lane (cheap).
__device__ __forceinline__ float sel5(const float* a, int k)
{
    float v = 0.f;
#pragma unroll
    for (int q = 0; q < 5; ++q) v = (k == q) ? a[q] : v;
    return v;
}

__global__ void __launch_bounds__(256) k_assign_pairs(const float4* __restrict__ sorted, long long Np, int N, float kf_ks, float offset, float* mesh)
{
    const long long gt = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long i = gt >> 2;
    const int pr = (int)(gt & 3);
    if (i >= Np || pr == 3) return;
    const float4 p = sorted[i];
    const AxisWin X = axis_window(kf_ks, p.x, offset), Y = axis_window(kf_ks, p.y, offset), Z = axis_window(kf_ks, p.z, offset);
    const int P0 = X.c0 >> 1, off = X.c0 - 2 * P0, Nh = N / 2;
    const int k0 = 2 * pr - off, k1 = k0 + 1;                  // window cells held by this lane's pair
    const float xa0 = sel5(X.a, k0), xa1 = sel5(X.a, k1), xb0 = sel5(X.b, k0), xb1 = sel5(X.b, k1);
    if (xa0 == 0.f && xa1 == 0.f && xb0 == 0.f && xb1 == 0.f) return;
    int P = (P0 + pr) % Nh;
    if (P < 0) P += Nh;
    const long long xo = 4LL * P;
#pragma unroll
    for (int rz = 0; rz < 5; ++rz) {
        if (Z.a[rz] == 0.f && Z.b[rz] == 0.f) continue;
        const long long zo = (long long)wrapN(Z.c0 + rz, N) * N;
#pragma unroll
        for (int ry = 0; ry < 5; ++ry) {
            const float wa = __fmul_rn(__fmul_rn(Y.a[ry], Z.a[rz]), p.w), wb = __fmul_rn(__fmul_rn(Y.b[ry], Z.b[rz]), p.w);
            if (wa == 0.f && wb == 0.f) continue;
            float* row = mesh + (zo + wrapN(Y.c0 + ry, N)) * (2LL * N);
            const float v0 = xa0 * wa, v1 = xb0 * wb, v2 = xa1 * wa, v3 = xb1 * wb;
            if (v0 != 0.f || v1 != 0.f || v2 != 0.f || v3 != 0.f) red_add_v4(row + xo, v0, v1, v2, v3);
        }
    }
}

// Three-lanes-per-particle variant (default): lanes 3g, 3g+1, 3g+2 of a warp own the three aligned cell pairs of particle g
// (10 particles per warp, lanes 30/31 idle: 94 % of the lanes work instead of 75 % with four lanes per particle), so the merged
// 48-byte row reductions of k_assign_pairs are kept while a warp instruction serves 10 particles instead of 8.  The inner loop
// is trimmed to what ncu showed it spends its issue slots on (k_assign_pairs: 78 % issue utilisation, 1435 instructions per
// lane and particle, REDG only 1.7 % of them): periodic wraps by compare-and-subtract from one wrapped base cell instead of a
// modulo per row, 32-bit element offsets (2 N^3 < 2^32), weights pre-multiplied per z row, and the four "product != 0" tests of
// a reduction replaced by tests on its factors.
// Slab mode (nzs < N): `mesh` holds the planes zbase .. zbase+nzs-1 only, contributions to other planes are dropped (their owner
// rank receives the same particle as a ghost), and only the first *nvalid sorted particles can touch the slab.
__global__ void __launch_bounds__(256) k_assign_tri(const float4* __restrict__ sorted, long long Np, int N, float kf_ks, float offset, float* mesh,
                                                    int zbase, int nzs, const unsigned int* __restrict__ nvalid)
{
    const int lane = threadIdx.x & 31;
    const int g = lane / 3, pr = lane - 3 * g;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long i = warp * 10 + g;
    if (g >= 10 || i >= Np) return;
    if (nvalid && i >= (long long)*nvalid) return;
    const float4 p = sorted[i];
    const AxisWin X = axis_window(kf_ks, p.x, offset), Y = axis_window(kf_ks, p.y, offset), Z = axis_window(kf_ks, p.z, offset);
    const int P0 = X.c0 >> 1, off = X.c0 - 2 * P0, Nh = N / 2;
    const int k0 = 2 * pr - off, k1 = k0 + 1;                  // window cells held by this lane's pair
    const float xa0 = sel5(X.a, k0), xa1 = sel5(X.a, k1), xb0 = sel5(X.b, k0), xb1 = sel5(X.b, k1);
    const bool nza = xa0 != 0.f || xa1 != 0.f, nzb = xb0 != 0.f || xb1 != 0.f;
    if (!nza && !nzb) return;
    int P = (P0 + pr) % Nh;
    if (P < 0) P += Nh;
    const int y0 = wrapN(Y.c0, N), z0 = wrapN(Z.c0, N);
    const unsigned rowlen = 2u * (unsigned)N;
    unsigned yoff[5];
    float za[5], zb[5];
#pragma unroll
    for (int r = 0; r < 5; ++r) {
        int yy = y0 + r;
        yy = yy >= N ? yy - N : yy;
        yoff[r] = (unsigned)yy * rowlen + 4u * (unsigned)P;
        za[r] = Z.a[r] * p.w;
        zb[r] = Z.b[r] * p.w;
    }
#pragma unroll
    for (int rz = 0; rz < 5; ++rz) {
        if (za[rz] == 0.f && zb[rz] == 0.f) continue;
        int zz = z0 + rz;
        zz = zz >= N ? zz - N : zz;
        zz -= zbase;                                           // plane inside the slab (full grid: zbase = 0, nzs = N)
        zz = zz < 0 ? zz + N : zz;
        if (zz >= nzs) continue;
        const unsigned zoff = (unsigned)zz * (unsigned)N * rowlen;
#pragma unroll
        for (int ry = 0; ry < 5; ++ry) {
            const float wa = Y.a[ry] * za[rz], wb = Y.b[ry] * zb[rz];
            if (!((nza && wa != 0.f) || (nzb && wb != 0.f))) continue;
            red_add_v4(mesh + (zoff + yoff[ry]), xa0 * wa, xb0 * wb, xa1 * wa, xb1 * wb);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------------
// Tile scatter: the shared-memory accumulation north_star asks for, without atomics.  One WARP owns one tile of TX x TY x TZ
// cells at a time (plus the 1 + 3 halo cells its particles reach) as two planar float arrays (grid A, grid B) in shared memory
// that no other warp touches.  Per batch of 32 particles: (1) every lane evaluates ONE particle -- cell, the 2 x 12 cubic weights
// of f:316-320, the 2 x 16 products wx*wy, wz*w -- and leaves them in a staging row; (2) the warp walks the batch, and for each
// particle lane (x, y, zl) adds the four values (x, y, zl | zl+2) of grid A and of grid B with plain LDS / FADD / STS: 32 distinct
// banks per instruction, no two lanes share an address, and a warp's shared-memory operations execute in order, so consecutive
// particles may overlap freely.  The finished tile leaves with vector reductions whose neighbouring lanes cover neighbouring
// 16 bytes (full sectors); all-zero pairs are skipped.  The L1 -> L2 reduction traffic that bounded the per-particle scatter
// (k_assign_tri: ~56 vector reductions per particle) becomes ~2.3 per CELL, independent of the particle density.
__global__ void __launch_bounds__(32 * TILE_WARPS) k_assign_tile(const float4* __restrict__ sorted, const unsigned int* __restrict__ key_end,
                                                                  int N, float kf_ks, float offset, float* mesh, int zbase, int nzs,
                                                                  unsigned int* tile_counter)
{
    extern __shared__ __align__(16) float tsm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* tA = tsm + (size_t)warp * (2 * TILE_WORDS + 32 * STG);
    float* tB = tA + TILE_WORDS;
    float* stg = tB + TILE_WORDS;
    const TileGeom g = tile_geom(N, zbase, nzs);
    const int lx = lane & 3, ly = (lane >> 2) & 3, lz = lane >> 4;
    const int loff = lx + ly * SY + lz * SZ;
    for (;;) {
        unsigned int t = 0;
        if (lane == 0) t = atomicAdd(tile_counter, 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= (unsigned)g.nkeys) break;
        const unsigned int beg = t ? key_end[t - 1] : 0u, end = key_end[t];
        if (beg == end) continue;
        const int tx = (int)(t % g.ntx), ty = (int)((t / g.ntx) % g.nty), tz = (int)(t / ((unsigned)g.ntx * g.nty));
        const int ox = tx * TX, oy = ty * TY, ozr = tz * TZ;          // ozr: z origin relative to g.zlo
        for (int i = lane; i < 2 * TILE_WORDS / 4; i += 32) reinterpret_cast<float4*>(tA)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncwarp();
        for (unsigned int b0 = beg; b0 < end; b0 += 32) {
            const int n = (int)min(32u, end - b0);
            if (lane < n) {
                const float4 p = sorted[b0 + lane];
                float* row = stg + lane * STG;
                float wa[3][4], wb[3][4];
                int cl[3], sh[3];
                const float pos[3] = { p.x, p.y, p.z };
#pragma unroll
                for (int ax = 0; ax < 3; ++ax) {
                    const float rx = grid_coord(kf_ks, pos[ax], offset), tx_ = __fadd_rn(rx, 0.5f);
                    const int im1 = (int)rx, nm1 = (int)tx_;
                    pcs4(__fsub_rn(rx, (float)im1), wa[ax][0], wa[ax][1], wa[ax][2], wa[ax][3]);
                    pcs4(__fsub_rn(tx_, (float)nm1), wb[ax][0], wb[ax][1], wb[ax][2], wb[ax][3]);
                    sh[ax] = nm1 - im1;                               // grid B's base cell is c or c+1
                    int c = im1 - 1;
                    if (ax == 2) c = wrapN(c - g.zlo, N) - ozr; else c = wrapN(c, N) - (ax == 0 ? ox : oy);
                    cl[ax] = c;                                        // array index of cell c-1 (index 0 = cell origin-1)
                }
                // staging row (float4 stores): [0,32) {wxyA[k], wxyB[k]} interleaved, k = 4 y + x; [32,36) z weights (times w) of the
                // lanes with lz = 0: {A0, A2, B0, B2}; [36,40) lz = 1: {A1, A3, B1, B3}; [40,42) array offsets of the two windows
                float4* r4 = reinterpret_cast<float4*>(row);
#pragma unroll
                for (int yy = 0; yy < 4; ++yy) {
                    r4[2 * yy] = make_float4(wa[0][0] * wa[1][yy], wb[0][0] * wb[1][yy], wa[0][1] * wa[1][yy], wb[0][1] * wb[1][yy]);
                    r4[2 * yy + 1] = make_float4(wa[0][2] * wa[1][yy], wb[0][2] * wb[1][yy], wa[0][3] * wa[1][yy], wb[0][3] * wb[1][yy]);
                }
                r4[8] = make_float4(wa[2][0] * p.w, wa[2][2] * p.w, wb[2][0] * p.w, wb[2][2] * p.w);
                r4[9] = make_float4(wa[2][1] * p.w, wa[2][3] * p.w, wb[2][1] * p.w, wb[2][3] * p.w);
                const int baseA = cl[0] + cl[1] * SY + cl[2] * SZ;
                r4[10] = make_float4(__int_as_float(baseA), __int_as_float(baseA + sh[0] + sh[1] * SY + sh[2] * SZ), 0.f, 0.f);
            }
            __syncwarp();
            for (int q = 0; q < n; ++q) {
                const float* row = stg + q * STG;
                const float2 bs = *reinterpret_cast<const float2*>(row + 40);
                const float2 wxy = *reinterpret_cast<const float2*>(row + 2 * (lane & 15));
                const float4 wz = *reinterpret_cast<const float4*>(row + 32 + 4 * lz);
                float* pa = tA + __float_as_int(bs.x) + loff;
                float* pb = tB + __float_as_int(bs.y) + loff;
                const float a0 = pa[0], a1 = pa[2 * SZ], b0 = pb[0], b1 = pb[2 * SZ];
                pa[0] = a0 + wxy.x * wz.x;
                pa[2 * SZ] = a1 + wxy.x * wz.y;
                pb[0] = b0 + wxy.y * wz.z;
                pb[2 * SZ] = b1 + wxy.y * wz.w;
                __syncwarp();
            }
        }
        // ---- flush: 4 rows of 12 cells per step; lane = (row in step, slot); slot s covers the ALIGNED cell pair ox + 2s - 2, + 1
        //      = array indices 2s - 1, 2s (index ix <-> cell ox + ix - 1): slot 0 only has its right cell (ox - 1), slot 6 only its
        //      left one (ox + 10), slot 7 is idle.  Every lane issues the same vector reduction (the missing half adds zero), and all
        //      wrapping is one modulo per tile and axis plus compare-and-subtract inside the loops (the first version spent as many
        //      instructions on % N here as the whole particle loop).
        const int fr = lane >> 3, slot = lane & 7;
        const bool hasl = slot >= 1 && slot <= 6, hasr = slot <= 5;
        const int ixl = 2 * slot - 1;
        const int cxl = wrapN(ox + ixl - 1, N);                          // even: N and ox are
        const int y0 = wrapN(oy - 1, N), z0 = wrapN(g.zlo + ozr - 1 - zbase, N);      // plane inside the slab (full grid: zbase = 0)
        const size_t rowf = (size_t)N * 2;
        for (int iz = 0; iz < AZ; ++iz) {
            int zl = z0 + iz;
            zl -= zl >= N ? N : 0;
            if (zl >= nzs) continue;                                      // warp-uniform
            float* mz = mesh + (size_t)zl * N * rowf + 2 * cxl;
#pragma unroll
            for (int j = 0; j < AY / 4; ++j) {
                const int iy = 4 * j + fr;
                int gy = y0 + iy;
                gy -= gy >= N ? N : 0;
                const float* ra = tA + iy * SY + iz * SZ + ixl;
                const float* rb = tB + iy * SY + iz * SZ + ixl;
                const float a0 = hasl ? ra[0] : 0.f, b0 = hasl ? rb[0] : 0.f, a1 = hasr ? ra[1] : 0.f, b1 = hasr ? rb[1] : 0.f;
                if (a0 != 0.f || a1 != 0.f || b0 != 0.f || b1 != 0.f) red_add_v4(mz + gy * rowf, a0, b0, a1, b1);
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------------------------------
// Slab routing (multi-GPU): rank q owns the mesh planes [q*nzr, (q+1)*nzr).  A particle in cell c touches planes
// c-1 .. c+3 (grid A: c-1..c+2, the half-cell shifted grid B up to c+3), so it is sent to the owner of plane c-1 and, if
// different, to the owner of plane c+3 (ghost copy).  Positions leave as float32 after the float64 clip of py:938-941, exactly
// what assign_quad receives, packed with the weight as float4.  Two kernels: counts per destination (+ sum of weights over the
// un-duplicated particles), then a scatter into a destination-major send buffer; chunk-local shared-memory ranks keep the
// global atomics at one per destination and 256-particle chunk.
__device__ __forceinline__ void slab_dests(const AssignIn& a, float z, int nzr, int& d0, int& d1)
{
    const int c = (int)grid_coord(a.kf_ks, z, a.offset) - 1;
    d0 = wrapN(c - 1, a.N) / nzr;
    d1 = wrapN(c + 3, a.N) / nzr;
}

constexpr int ROUTE_MAXR = 64;

__global__ void __launch_bounds__(256) k_route_count(AssignIn a, int nzr, int nranks, unsigned long long* counts, double* sumw)
{
    __shared__ unsigned int h[ROUTE_MAXR];
    if (threadIdx.x < ROUTE_MAXR) h[threadIdx.x] = 0u;
    __syncthreads();
    double acc = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < a.Np; i += (long long)gridDim.x * blockDim.x) {
        float x, y, z, w; double wd;
        load_particle(a, i, x, y, z, w, wd);
        acc += wd;
        int d0, d1;
        slab_dests(a, z, nzr, d0, d1);
        atomicAdd(&h[d0], 1u);
        if (d1 != d0) atomicAdd(&h[d1], 1u);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    __shared__ double red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < nranks && h[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)h[threadIdx.x]);
    if (threadIdx.x == 0) { double t = 0.0; for (int k = 0; k < 8; ++k) t += red[k]; atomicAdd(sumw, t); }
}

__global__ void __launch_bounds__(256) k_route_scatter(AssignIn a, int nzr, int nranks, const unsigned long long* __restrict__ base,
                                                       unsigned long long* cursor, float4* send)
{
    __shared__ unsigned int cnt[ROUTE_MAXR];
    __shared__ unsigned long long off[ROUTE_MAXR];
    const long long nchunk = (a.Np + blockDim.x - 1) / blockDim.x;
    for (long long ch = blockIdx.x; ch < nchunk; ch += gridDim.x) {
        if (threadIdx.x < ROUTE_MAXR) cnt[threadIdx.x] = 0u;
        __syncthreads();
        const long long i = ch * blockDim.x + threadIdx.x;
        float x = 0.f, y = 0.f, z = 0.f, w = 0.f; double wd;
        int d0 = -1, d1 = -1;
        unsigned int r0 = 0, r1 = 0;
        if (i < a.Np) {
            load_particle(a, i, x, y, z, w, wd);
            slab_dests(a, z, nzr, d0, d1);
            r0 = atomicAdd(&cnt[d0], 1u);
            if (d1 != d0) r1 = atomicAdd(&cnt[d1], 1u); else d1 = -1;
        }
        __syncthreads();
        if (threadIdx.x < nranks) off[threadIdx.x] = base[threadIdx.x] + atomicAdd(&cursor[threadIdx.x], (unsigned long long)cnt[threadIdx.x]);
        __syncthreads();
        if (d0 >= 0) send[off[d0] + r0] = make_float4(x, y, z, w);
        if (d1 >= 0) send[off[d1] + r1] = make_float4(x, y, z, w);
        __syncthreads();
    }
}

// The same scatter with the particle exchange fused in (NVLink peer stores instead of a send buffer + all-to-all): dest[d] is the
// address at which THIS rank's segment starts in rank d's receive buffer (peer pointer + the particles of the ranks in front, from
// the all-gathered counts).  A CTA takes 2048 consecutive particles, ranks their <= 4096 copies per destination in shared memory,
// reserves the tile's range of every destination with one atomic on the local cursor, REORDERS the copies in shared memory and
// writes them out destination by destination: neighbouring lanes store neighbouring 16 bytes, runs of ~2048/G particles.
constexpr int RP_THREADS = 512, RP_ITEMS = 4, RP_TILE = RP_THREADS * RP_ITEMS;

__global__ void __launch_bounds__(RP_THREADS, 2) k_route_scatter_peer(AssignIn a, int nzr, int nranks, const long long* __restrict__ dest,
                                                                      unsigned long long* cursor)
{
    extern __shared__ __align__(16) unsigned char rp_smem[];
    float4* sp = reinterpret_cast<float4*>(rp_smem);                                  // [2 * RP_TILE] copies in destination order
    unsigned char* sd = reinterpret_cast<unsigned char*>(sp + 2 * RP_TILE);          // [2 * RP_TILE] destination of every slot
    __shared__ unsigned int cnt[ROUTE_MAXR], lstart[ROUTE_MAXR];
    __shared__ long long gaddr[ROUTE_MAXR];                                          // address of slot lstart[d] in rank d's buffer
    __shared__ unsigned int total;
    const long long ntile = (a.Np + RP_TILE - 1) / RP_TILE;
    for (long long tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
        const long long lo = tile * RP_TILE;
        const int n = (int)(a.Np - lo < RP_TILE ? a.Np - lo : RP_TILE);
        if (threadIdx.x < ROUTE_MAXR) cnt[threadIdx.x] = 0u;
        __syncthreads();
        float4 p[RP_ITEMS];
        int d0[RP_ITEMS], d1[RP_ITEMS];
        unsigned int r0[RP_ITEMS], r1[RP_ITEMS];
#pragma unroll
        for (int it = 0; it < RP_ITEMS; ++it) {
            const int i = it * RP_THREADS + threadIdx.x;
            d0[it] = d1[it] = -1;
            if (i < n) {
                float x, y, z, w; double wd;
                load_particle(a, lo + i, x, y, z, w, wd);
                p[it] = make_float4(x, y, z, w);
                slab_dests(a, z, nzr, d0[it], d1[it]);
                r0[it] = atomicAdd(&cnt[d0[it]], 1u);
                if (d1[it] != d0[it]) r1[it] = atomicAdd(&cnt[d1[it]], 1u); else d1[it] = -1;
            }
        }
        __syncthreads();
        if (threadIdx.x < 32) {                            // exclusive scan over the destinations (<= 64: two per lane) + reservation
            const int e0 = 2 * threadIdx.x, e1 = e0 + 1;
            const unsigned int c0 = e0 < nranks ? cnt[e0] : 0u, c1 = e1 < nranks ? cnt[e1] : 0u;
            unsigned int x = c0 + c1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned int y = __shfl_up_sync(0xffffffffu, x, o); if (threadIdx.x >= o) x += y; }
            const unsigned int s0 = x - (c0 + c1), s1 = s0 + c0;
            if (e0 < nranks) { lstart[e0] = s0; if (c0) gaddr[e0] = dest[e0] + 16LL * (long long)atomicAdd(&cursor[e0], (unsigned long long)c0); }
            if (e1 < nranks) { lstart[e1] = s1; if (c1) gaddr[e1] = dest[e1] + 16LL * (long long)atomicAdd(&cursor[e1], (unsigned long long)c1); }
            if (threadIdx.x == 31) total = x;
        }
        __syncthreads();
#pragma unroll
        for (int it = 0; it < RP_ITEMS; ++it) {
            if (d0[it] >= 0) { const unsigned int s = lstart[d0[it]] + r0[it]; sp[s] = p[it]; sd[s] = (unsigned char)d0[it]; }
            if (d1[it] >= 0) { const unsigned int s = lstart[d1[it]] + r1[it]; sp[s] = p[it]; sd[s] = (unsigned char)d1[it]; }
        }
        __syncthreads();
        const unsigned int ne = total;
        for (unsigned int s = threadIdx.x; s < ne; s += RP_THREADS) {
            const int d = sd[s];
            *reinterpret_cast<float4*>(gaddr[d] + 16LL * (long long)(s - lstart[d])) = sp[s];
        }
        __syncthreads();
    }
}

int slab_route_scatter_peer(const AssignIn& in, int nzr, int nranks, const long long* dest, unsigned long long* cursor, cudaStream_t st)
{
    if (in.N < 4 || in.N % 2 || in.Np < 0 || nranks < 1 || nranks > ROUTE_MAXR || nzr < 8 || nzr * nranks != in.N || !dest || !cursor) return PSB_ERR_ARG;
    if (cudaMemsetAsync(cursor, 0, nranks * sizeof(unsigned long long), st) != cudaSuccess) return PSB_ERR_CUDA;
    if (in.Np > 0) {
        const size_t smem = (size_t)2 * RP_TILE * (sizeof(float4) + 1);
        if (cudaFuncSetAttribute(k_route_scatter_peer, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return PSB_ERR_CUDA;
        const long long nt = (in.Np + RP_TILE - 1) / RP_TILE;
        k_route_scatter_peer<<<(unsigned)(nt < 2LL * sm_count() ? nt : 2LL * sm_count()), RP_THREADS, smem, st>>>(in, nzr, nranks, dest, cursor);
    }
    return cudaGetLastError() == cudaSuccess ? PSB_OK : PSB_ERR_CUDA;
}

int slab_route_count(const AssignIn& in, int nzr, int nranks, unsigned long long* counts, double* sumw, cudaStream_t st)
{
    if (in.N < 4 || in.N % 2 || in.Np < 0 || nranks < 1 || nranks > ROUTE_MAXR || nzr < 8 || nzr * nranks != in.N) return PSB_ERR_ARG;
    if (cudaMemsetAsync(counts, 0, nranks * sizeof(unsigned long long), st) != cudaSuccess) return PSB_ERR_CUDA;
    if (cudaMemsetAsync(sumw, 0, sizeof(double), st) != cudaSuccess) return PSB_ERR_CUDA;
    if (in.Np > 0) {
        const long long nb = (in.Np + 255) / 256;
        k_route_count<<<(unsigned)(nb < sm_count() * 16 ? nb : sm_count() * 16), 256, 0, st>>>(in, nzr, nranks, counts, sumw);
    }
    return cudaGetLastError() == cudaSuccess ? PSB_OK : PSB_ERR_CUDA;
}

int slab_route_scatter(const AssignIn& in, int nzr, int nranks, const unsigned long long* base, unsigned long long* cursor, float4* send,
                       cudaStream_t st)
{
    if (in.N < 4 || in.N % 2 || in.Np < 0 || nranks < 1 || nranks > ROUTE_MAXR || nzr < 8 || nzr * nranks != in.N) return PSB_ERR_ARG;
    if (cudaMemsetAsync(cursor, 0, nranks * sizeof(unsigned long long), st) != cudaSuccess) return PSB_ERR_CUDA;
    if (in.Np > 0) {
        const long long nb = (in.Np + 255) / 256;
        k_route_scatter<<<(unsigned)(nb < sm_count() * 16 ? nb : sm_count() * 16), 256, 0, st>>>(in, nzr, nranks, base, cursor, send);
    }
    return cudaGetLastError() == cudaSuccess ? PSB_OK : PSB_ERR_CUDA;
}

// counter area: sort keys = N*N rows (any slab: (nzs+8)*N + 1, nzs + 8 <= N) or the tiles of the tile scatter (+ trailing bucket)
static size_t assign_hist_bytes(int N)
{
    const TileGeom g = tile_geom(N, 0, N);
    const size_t nkey = (size_t)N * N > (size_t)g.nkeys ? (size_t)N * N : (size_t)g.nkeys;
    return ((nkey + 2) * sizeof(unsigned int) + 255) / 256 * 256 + 131072;     // + scan tile sums (4096), the tile work counter, coarse cursors (8192), bucket starts (16384)
}

// PSB_ASSIGN_TWOPASS (tests / A-B runs): 0 = never, 1 = always, n >= 2 = always with coarse buckets of 2^n keys; unset = by size
static int two_pass_knob()
{
    static const int v = [] { const char* e = getenv("PSB_ASSIGN_TWOPASS"); return e ? atoi(e) : -1; }();
    return v;
}

// whether the two-pass sort is a candidate (many keys AND a particle array far beyond L2); whether it RUNS is then decided on the
// device from how ordered the input already is (measured: cell-ordered 1e8 particles at 512^3 sort 12 % faster in one pass,
// shuffled ones 7 % faster in two; 1024^3 / 5e8 random particles 62.5 -> 49.9 ms in two)
static bool sort_two_pass(long long Np, size_t nkeys)
{
    const int force = two_pass_knob();
    if (force >= 0) return force != 0 && Np > 0;
    return nkeys >= ((size_t)1 << 19) && (size_t)Np * sizeof(float4) >= ((size_t)512 << 20);
}

size_t assign_workspace_bytes(long long Np, int N)
{
    const TileGeom g = tile_geom(N, 0, N);
    const size_t copies = sort_two_pass(Np, (size_t)g.nkeys) ? 2 : 1;      // + the bucket-ordered copy
    return 2 * assign_hist_bytes(N) + copies * ((size_t)Np * sizeof(float4) + 256) + 256;
}

int assign_pcs_interlaced(const AssignIn& in_, float* mesh, int zero_mesh, void* ws, size_t ws_bytes, double* sumw, cudaStream_t st)
{
    if (in_.N < 4 || (in_.N % 2) || in_.Np < 0) return PSB_ERR_ARG;
    if (ws_bytes < assign_workspace_bytes(in_.Np, in_.N)) return PSB_ERR_WORKSPACE;
    const bool slab = in_.nzs < in_.N;
    if (in_.nzs < 1 || in_.nzs > in_.N || (slab && (in_.nzs + 8 > in_.N || in_.zbase < 0 || in_.zbase >= in_.N))) return PSB_ERR_ARG;
    if (slab && 2.0 * in_.nzs * in_.N * in_.N >= 4294967296.0) return PSB_ERR_ARG;
    // 3 (default): tile scatter in shared memory; 2: per-particle vector reductions, three lanes per particle; 1, 0: older variants
    static const int variant = [] { const char* e = getenv("PSB_ASSIGN_VARIANT"); return e ? atoi(e) : 3; }();
    AssignIn in = in_;
    // the tile scatter pays a fixed price per tile (zero, flush): below ~0.1 particles per cell the per-particle reductions win
    // (measured: 256^3 with 1e6 particles 0.26 vs 0.48 ms; 360^3 with 1e7 2.17 vs 1.80 ms; 512^3 with 1e8 21.6 vs 10.3 ms)
    static const double min_density = [] { const char* e = getenv("PSB_ASSIGN_TILE_DENSITY"); return e ? atof(e) : 0.1; }();
    in.tiles = (variant == 3 && in.N >= 16 && (double)in.Np >= min_density * (double)in.nzs * in.N * in.N) ? 1 : 0;
    const TileGeom tg = tile_geom(in.N, in.zbase, in.nzs);
    // sort keys; in slab mode one trailing bucket takes the particles that cannot touch the slab
    const size_t nrow = in.tiles ? (size_t)tg.nkeys + (slab ? 1 : 0) : (slab ? (size_t)(in.nzs + 8) * in.N + 1 : (size_t)in.N * in.N);
    const size_t mesh_rows = (size_t)in.nzs * in.N;
    const size_t hist_b = assign_hist_bytes(in.N);
    unsigned int* hist = static_cast<unsigned int*>(ws);
    unsigned int* tile_sum = hist + hist_b / sizeof(unsigned int);      // second counter area: scan tile sums (<= 4096 entries)
    unsigned int* tile_counter = tile_sum + 4096;                         // work counter of the tile scatter
    float4* sorted = reinterpret_cast<float4*>(static_cast<char*>(ws) + 2 * hist_b);
    if (cudaMemsetAsync(hist, 0, 2 * hist_b, st) != cudaSuccess) return PSB_ERR_CUDA;
    if (cudaMemsetAsync(sumw, 0, sizeof(double), st) != cudaSuccess) return PSB_ERR_CUDA;
    if (zero_mesh && cudaMemsetAsync(mesh, 0, sizeof(float) * 2 * mesh_rows * in.N, st) != cudaSuccess) return PSB_ERR_CUDA;
    if (in.Np > 0) {
        const int blk = 256;
        const int grid = (int)((in.Np + blk - 1) / blk < sm_count() * 16 ? (in.Np + blk - 1) / blk : sm_count() * 16);
        // candidate for the two-pass sort by size: k_hist also measures how ordered the input is and k_sort_decide picks the path
        // on the device (both paths are enqueued, the kernels of the other one return at once); the knob forces either path
        const bool adaptive = in.tiles && two_pass_knob() < 0 && sort_two_pass(in.Np, nrow);
        unsigned long long* ordered = reinterpret_cast<unsigned long long*>(tile_sum + 24576);      // zeroed with the counter areas
        int* path_flag = reinterpret_cast<int*>(tile_sum + 24580);
        k_hist<<<grid, blk, 0, st>>>(in, hist, sumw, adaptive ? ordered : nullptr);
        if (adaptive) k_sort_decide<<<1, 1, 0, st>>>(ordered, in.Np, path_flag);
        const int* flag = adaptive ? path_flag : nullptr;
        const int ntile = (int)((nrow + SCAN_TILE - 1) / SCAN_TILE);
        if (ntile > 4096) return PSB_ERR_UNSUPPORTED_N;
        k_scan_tile_sums<<<ntile, 256, 0, st>>>(hist, (int)nrow, tile_sum);
        k_scan_tiles<<<1, 256, 0, st>>>(tile_sum, ntile);
        k_scan_apply<<<ntile, 256, 0, st>>>(hist, (int)nrow, tile_sum);
        int gshift = two_pass_knob() >= 2 ? two_pass_knob() : 10;        // coarse buckets of >= 1024 keys, at most PC_MAXB of them
        while ((((long long)nrow + (1LL << gshift) - 1) >> gshift) > PC_MAXB) ++gshift;
        const size_t copy_b = ((size_t)in.Np * sizeof(float4) + 255) / 256 * 256;
        const bool two_pass = in.tiles && sort_two_pass(in.Np, nrow) && (1 << gshift) <= BS_MAXD && in.Np < 4294967295LL &&
                              ws_bytes >= 2 * hist_b + 2 * copy_b;
        if (two_pass) {
            const int nb = (int)(((long long)nrow + (1LL << gshift) - 1) >> gshift);
            float4* bucketed = reinterpret_cast<float4*>(reinterpret_cast<char*>(sorted) + copy_b);
            unsigned int* coarse_cursor = tile_sum + 8192;                 // zeroed with the counter areas
            unsigned int* bucket_start = tile_sum + 16384;
            k_bucket_starts<<<(PC_MAXB + 256) / 256, 256, 0, st>>>(hist, (long long)nrow, gshift, PC_MAXB, (unsigned)in.Np, bucket_start);
            const size_t pc_smem = (size_t)PC_TILE * (sizeof(float4) + sizeof(unsigned short)) + 2 * PC_MAXB * sizeof(unsigned int);
            if (cudaFuncSetAttribute(k_partition_coarse, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pc_smem) != cudaSuccess) return PSB_ERR_CUDA;
            const long long ntile = (in.Np + PC_TILE - 1) / PC_TILE;
            k_partition_coarse<<<(unsigned)(ntile < 2LL * sm_count() ? ntile : 2LL * sm_count()), PC_THREADS, pc_smem, st>>>(
                in, gshift, nb, bucket_start, coarse_cursor, bucketed, flag);
            AssignIn in2 = in;                                            // second pass: float4 {x,y,z,w}, already clipped and cast
            in2.pos = bucketed; in2.pos_aos = 2; in2.pos_f64 = 0; in2.w = nullptr; in2.do_clip = 0;
            long long pieces = (in.Np / nb + BS_CHUNK - 1) / BS_CHUNK;   // CTAs per bucket: one chunk each at uniform density
            pieces = pieces < 1 ? 1 : (pieces > 64 ? 64 : pieces);
            k_bucket_scatter<<<(unsigned)(nb * pieces), BS_THREADS, 0, st>>>(in2, gshift, (long long)nrow, (int)pieces, bucket_start, hist, sorted,
                                                                          flag);
            if (adaptive) k_sort_scatter<<<grid, blk, 0, st>>>(in, hist, sorted, flag, 0);
        } else {
            k_sort_scatter<<<grid, blk, 0, st>>>(in, hist, sorted);
        }
        // after the scatter hist[k] = end of key k = start of key k+1
        if (in.tiles) {
            const size_t smem = (size_t)TILE_WARPS * (2 * TILE_WORDS + 32 * STG) * sizeof(float);
            if (cudaFuncSetAttribute(k_assign_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return PSB_ERR_CUDA;
            const long long want = ((long long)tg.nkeys + TILE_WARPS - 1) / TILE_WARPS;
            const int ncta = (int)(want < 3LL * sm_count() ? want : 3LL * sm_count());
            k_assign_tile<<<ncta, 32 * TILE_WARPS, smem, st>>>(sorted, hist, in.N, in.kf_ks, in.offset, mesh, slab ? in.zbase : 0,
                                                             slab ? in.nzs : in.N, tile_counter);
        } else if (slab || (variant >= 2 && in.N <= 1024)) {
            k_assign_tri<<<(unsigned)((in.Np + 79) / 80), 256, 0, st>>>(sorted, in.Np, in.N, in.kf_ks, in.offset, mesh, slab ? in.zbase : 0,
                                                                       slab ? in.nzs : in.N, slab ? hist + (nrow - 2) : nullptr);
        } else if (variant == 1) {
            k_assign_pairs<<<(unsigned)((4 * in.Np + 255) / 256), 256, 0, st>>>(sorted, in.Np, in.N, in.kf_ks, in.offset, mesh);
        } else {
            k_assign<<<(unsigned)((in.Np + 255) / 256), 256, 0, st>>>(sorted, in.Np, in.N, in.kf_ks, in.offset, mesh);
        }
    }
    return cudaGetLastError() == cudaSuccess ? PSB_OK : PSB_ERR_CUDA;
}

}  // namespace psb
