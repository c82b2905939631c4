// psb_triangles_tc.cu -- K6 on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a only.
//
//   S[i,j,l] = sum_x I_i(x) I_j(x) I_l(x)      (pyspectrum.py:427-430)
//
// as a split-K GEMM  C[p,l] = sum_x A[p,x] B[x,l]  with  A[p,x] = I_i(x) I_j(x)  formed on the fly for the
// pair rows p=(i,j) that own at least one wanted triangle (j >= i/2: 440 rows instead of 820 at Nmax=40),
// B[x,l] = I_l(x).  Operands are split into fp16 hi + fp16 lo (A = Ah+Al, B = Bh+Bl, each 11 significant
// bits) and three MMAs  Ah.Bh + Ah.Bl + Al.Bh  recover ~2^-21 relative accuracy per product -- the
// "3x split" scheme north_star allows, with fp16 rather than TF32 pieces (kind::f16 runs at twice the
// kind::tf32 rate).  The shell fields arrive already split (K5's epilogue stores {half2 hi, half2 lo} per cell
// pair, 4 bytes per cell as before), so B needs no arithmetic and the A products are formed and re-split in
// packed half2 FMAs (prod_split: 2 instructions per product).  Every CTA keeps the whole C (MT x 128 rows x NT columns) resident in TMEM and owns a
// contiguous x range; tensor-core accumulation rounds toward zero, so TMEM is drained into float64
// every p.flush_ksteps*16 cells (double-buffered accumulator sets: the drain overlaps the next MMAs).
//
// Warp roles (NG = 4 former groups: 20 warps = 640 threads at 96 registers, 1 CTA/SM; Roles<NG> below):
//   warps 0..15  A formers, warp = 4*g + q (group g, TMEM lane quarter q).  Default lane layout 1 (MT == 4): a thread owns one TMEM
//                lane = a 2x2 block of pair rows (i0,i1) x (j0,j1), one row in each of the four M tiles; a work unit is one K-step
//                (16 cells) of ALL four tiles, group g takes operand stage g of every 64-cell sub-chunk: four field vectors give four
//                rows of products (1 shared-memory word per product), formed in packed half2 already hi/lo split and written
//                straight into TMEM as the A operand (one tcgen05.st.x16 per tile).  Lane layout 0 (more than 64 shells: 2-3
//                accumulator tiles): a lane holds one i and one j per tile, a unit is one K-step of a tile pair.  The groups take
//                turns draining the accumulator tiles.  (PSB_TC_GROUPS=6 runs 24 former warps in layout 0: same speed.)
//   warps 16,17  B formers (two operand stages each): packed fields -> fp16 hi/lo K-major core-matrix tiles in shared memory
//   warp 18      TMA producer: cp.async.bulk (UBLKCP) of [S][XCH] field chunks, one row per lane
//   warp 19      MMA issuer (one elected lane): tcgen05.mma.cta_group::1.kind::f16 with A from TMEM, B from smem; tcgen05.commit
// Pipelines: chunk ring (TMA -> formers/B), four operand stages = the four K-steps of a sub-chunk (formers/B -> MMA),
// accumulators (MMA -> drain, one period late so nobody waits).
// Truncation: the tensor core rounds its fp32 accumulation toward zero; between two drains an accumulator takes
// 12*flush_chunks MMAs and comes out short by ~1.8e-8 per MMA relative to its own value (measured: mean -8.8e-7 of |S| at 48
// MMAs).  The drain adds the TMEM value times (1 + bias_comp) to the fp32 shared accumulators (one FFMA, round to nearest),
// which removes the systematic part; what is left is the zero-mean spread.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdlib>
#include <cstdio>
#include "psb_kernels.h"

namespace psb {

namespace tc {

// XCH = cells per staged chunk (template parameter: 256 -> one 1 KB bulk copy per field row; 64 when many shells would not fit)
constexpr int SUB = 64;              // cells per sub-chunk = 4 K-steps of 16 = one round of the four operand stages
constexpr int NCHUNKBUF = 2;
constexpr int TMEM_COLS = 512;
constexpr unsigned WATCHDOG = 1u << 22;       // x 20 us suspend hint

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// try_wait suspends the thread in hardware up to the hinted time, so waiting warps do not burn the issue
// slots the MMA-issuing warp needs (polling formers made that warp the bottleneck: profiles/r1)
// backoff_ns > 0: roles with slack (B formers, TMA producer) sleep between probes -- the hardware try_wait returns every ~85 cycles,
// and eight spinning warps were issuing a third of all instructions of the SM (profiles/r1_summary.md)
constexpr uint32_t WAIT_HINT_NS = 20000u;

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, unsigned backoff_ns = 0)
{
    const uint32_t hint = WAIT_HINT_NS;
    const uint32_t a = smem_u32(bar);
    uint32_t done = 0;
    unsigned spins = 0;
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity), "r"(hint) : "memory");
        if (done) break;
        if (backoff_ns) __nanosleep(backoff_ns);
        if (++spins > WATCHDOG) asm volatile("trap;");      // a protocol bug must not hang the GPU
    }
}
// non-blocking probe (issued early, consumed later: hides the ~150-cycle barrier read behind useful work)
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity)
{
    uint32_t done = 0;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return done != 0;
}
// one lane of a converged warp (the CUTLASS elect_one_sync idiom): keeps the warp's control flow uniform so that the
// uniform-datapath instructions (UTCHMMA, UTCBAR, UBLKCP) are issued directly instead of through a per-lane waterfall loop
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred px;\n\telect.sync _|px, 0xffffffff;\n\tselp.u32 %0, 1, 0, px;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1):
// core matrix = 8 rows x 16 B contiguous; SBO = byte step between 8-row groups, LBO = byte step between
// the two 16-byte K chunks of one K=16 MMA.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                                           // descriptor version (Blackwell)
    return d;
}

// kind::f16 instruction descriptor: D=f32, A=B=f16, both K-major, M=128, N=n
__device__ __forceinline__ uint32_t umma_idesc_f16(int n)
{
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}
// A operand from TMEM (128 lanes x 8 columns of packed half2 per K=16 step), B from shared memory
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}"
                 ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
// hi (8 columns) and lo (8 columns) of one tile and K-step are adjacent in TMEM: one 16-column store instead of two 8-column ones
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* a, const uint32_t* b)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
                   "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]) : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t* r)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
// Pair product of two hi/lo-split values, itself hi/lo split, entirely in packed half2 arithmetic (two cells per instruction):
//   hi = fp16(ih*jh);  lo = fp16(il*jh + fp16(ih*jl + (ih*jh - hi)))      -- the inner residual ih*jh - hi is exact (one FMA),
// so hi + lo = (ih+il)(jh+jl) - il*jl to ~2^-22 relative: 4 FMA-pipe instructions per two products, no conversions.
__device__ __forceinline__ void prod_split(uint32_t ih, uint32_t il, uint32_t jh, uint32_t jl, uint32_t& hi, uint32_t& lo)
{
    const __half2 a = *reinterpret_cast<const __half2*>(&ih), al = *reinterpret_cast<const __half2*>(&il);
    const __half2 b = *reinterpret_cast<const __half2*>(&jh), bl = *reinterpret_cast<const __half2*>(&jl);
    const __half2 h = __hmul2(a, b);
    __half2 r = __hfma2(a, b, __hneg2(h));
    r = __hfma2(a, bl, r);
    r = __hfma2(al, b, r);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&r);
}
__device__ __forceinline__ void umma_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t* r)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

struct Params {
    const float* const* fields;   // S device pointers, each ncell 32-bit words of packed hi/lo halves (16-byte aligned)
    int S;                        // real shells (rows of a chunk)
    int NT;                       // MMA N = shells padded to a multiple of 16 (<= 128)
    int MT;                       // M tiles (128 rows each) in this pass (<= 256 / tile_cols)
    int tile_cols;                // TMEM column spacing of one accumulator tile: 64 (NT <= 64) or NT
    const int* lane_ij;           // [128][5]: field slot i of the lane and j of its row in tiles 0..3 (-1: padding row)
    long long nchunk;             // ncell / XCH (chunk size of the launched instantiation)
    int flush_chunks;             // TMEM accumulators are drained (fp32, round-to-nearest) every flush_chunks 64-cell sub-chunks
    int gflush_drains;            // the fp32 accumulators go to float64 global every gflush_drains drains
    float bias_gain;              // 1 + bias_comp: multiplies every drained TMEM value (compensates the truncating accumulation)
    double* partial;              // [gridDim.x][NT][MT*128] float64 partial sums (zeroed by the host)
    long long* trace;             // optional clock64 timeline of CTA 0 (profiling only): [role 0..4][event][sub-chunk]
    unsigned backoff_ns;          // sleep between barrier probes of the roles with slack (B formers; x5 for the TMA producer)
    int debug;                    // ablation bits (profiling only): 1 no MMAs, 2 no forming math, 4 no tcgen05.st, 8 no drain body,
                                  //                                  16 no TMA copies, 32 no B forming, 256 no tcgen05 fences in the formers, 512 no I_i loads
};

// Warp roles for NG former groups (NG*4 former warps: warp = 4*group + lane quarter).
//   NG = 4 (640 threads, 96 registers): formers 0..15, B formers 16, 17 (two stages each), TMA 18, MMA 19.
//   NG = 6 (960 threads, 64 registers): formers 0..23, B formers 24..27, TMA 28, MMA 29.
template <int NG> struct Roles {
    static constexpr int NFW = NG * 4;
    static constexpr int NBW = NG == 4 ? 2 : 4;          // B-operand warps (stages b, b + NBW, ...): the regrouping is light, two suffice
    static constexpr int NTHR = (NFW + NBW + 2) * 32;    // NG = 4: 640 threads -> 96 registers per thread
    static constexpr int W_TMA = NFW + NBW;
    static constexpr int W_MMA = NFW + NBW + 1;          // highest warp id: favoured by the sub-partition arbiter
};
constexpr int TMEM_A0 = 256;

// TMEM map (512 columns): [0,256) accumulator tiles (tile m at m*tile_cols);
//                         [256,512) A-operand stages: stage g (= K-step g of a chunk) at 256 + g*64, tile m: hi at +m*16,
//                         lo at +m*16+8 (one K-step of one tile = 128 lanes x 16 fp16 = 8 columns of packed half2).

// TMEM accumulators of tile `m` -> fp32 shared accumulators (round-to-nearest adds); every gflush_drains-th drain
// (and the last one) moves them on to the float64 partial sums in global memory.
__device__ __forceinline__ void drain_accumulators(const Params& p, uint64_t* acc_full, uint64_t* acc_empty, int period, bool final_drain,
                                                   bool live, uint32_t t_acc, float4* accs, int MR, int NT, int row, int lane)
{
    mbar_wait(acc_full, (uint32_t)(period & 1));
    tc_fence_after();
    const bool to_global = ((period + 1) % p.gflush_drains == 0) || final_drain;
    const float bg = p.bias_gain;
    if (live && !(p.debug & 8)) {
        double* dst = p.partial + (size_t)blockIdx.x * MR * NT + (size_t)row;     // [col][row]: coalesced
        // one 16-column TMEM load per round trip: keeping two or three in flight needs 32-48 more live registers (measured on the
        // 704-thread / 80-register predecessor of this kernel: spills, 11.7 -> 15.6 ms; at 640 threads / 96 registers the forming
        // loop already holds 64 of them)
        for (int c0 = 0; c0 < NT; c0 += 16) {
            uint32_t r[16];
            tmem_ld16_issue(t_acc + (uint32_t)c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
                float4* sp = &accs[(size_t)(c0 / 4 + q4) * MR + row];
                float4 cur = *sp;
                cur.x = fmaf(__uint_as_float(r[4 * q4]), bg, cur.x); cur.y = fmaf(__uint_as_float(r[4 * q4 + 1]), bg, cur.y);
                cur.z = fmaf(__uint_as_float(r[4 * q4 + 2]), bg, cur.z); cur.w = fmaf(__uint_as_float(r[4 * q4 + 3]), bg, cur.w);
                if (to_global) {
                    dst[(size_t)(c0 + 4 * q4 + 0) * MR] += (double)cur.x;
                    dst[(size_t)(c0 + 4 * q4 + 1) * MR] += (double)cur.y;
                    dst[(size_t)(c0 + 4 * q4 + 2) * MR] += (double)cur.z;
                    dst[(size_t)(c0 + 4 * q4 + 3) * MR] += (double)cur.w;
                    cur = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                *sp = cur;
            }
        }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(acc_empty);
}

// shared memory carve-up (dynamic):
//   chunk[NCHUNKBUF][S][ROWF] fp32 | B_hi[4][2][NT][8] fp16 | B_lo[...] | accs[NT/4][MT*128][4] fp32 | barriers
template <int XCH, int NG, bool BLK2>
__global__ void __launch_bounds__(Roles<NG>::NTHR, 1) k_tri_tc(Params p)
{
    using R = Roles<NG>;
    constexpr int NTHR = R::NTHR, NFWARPS = R::NFW, NBW = R::NBW, W_MMA = R::W_MMA, W_TMA = R::W_TMA;
    constexpr int NSUB = XCH / SUB;
    constexpr int ROWF = XCH + 4;    // padded fp32 row stride (== 4 mod 32 words: conflict-free LDS.128 across rows)
    extern __shared__ __align__(1024) unsigned char smem[];
    const int S = p.S, NT = p.NT, MT = p.MT, MR = MT * 128;
    uint32_t* chunk = reinterpret_cast<uint32_t*>(smem);      // one 32-bit word per cell (packed hi/lo pairs, see psb_fft_lines.cuh pack_hilo)
    const size_t chunk_bytes = (size_t)S * ROWF * 4;
    unsigned char* bop = smem + ((NCHUNKBUF * chunk_bytes + 1023) / 1024) * 1024;
    const uint32_t b_stage_bytes = 2u * NT * 16u;
    unsigned char* b_hi = bop;
    unsigned char* b_lo = b_hi + 4 * b_stage_bytes;
    float4* accs = reinterpret_cast<float4*>(b_lo + 4 * b_stage_bytes);      // [NT/4][MR]
    uint64_t* bars = reinterpret_cast<uint64_t*>(accs + (size_t)(NT / 4) * MR);
    uint64_t* chunk_full = bars;                  // [NCHUNKBUF]  TMA -> formers + B warp (tx bytes)
    uint64_t* chunk_empty = bars + NCHUNKBUF;     // [NCHUNKBUF]  16 former warps + B warp -> TMA
    uint64_t* st_full = chunk_empty + NCHUNKBUF;  // [4]          8 former warps (two tile-pair groups) + B warp -> MMA
    uint64_t* st_empty = st_full + 4;             // [4]          MMA (commit) -> formers of the group, B warp
    uint64_t* acc_full = st_empty + 4;            // [1]          MMA (commit) -> drain
    uint64_t* acc_empty = acc_full + 1;           // [1]          drain (16 warps) -> MMA
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(acc_empty + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int i = 0; i < NCHUNKBUF; ++i) { mbar_init(&chunk_full[i], 1); mbar_init(&chunk_empty[i], NFWARPS + NBW); }
        for (int i = 0; i < 4; ++i) { mbar_init(&st_full[i], BLK2 ? 5 : 9); mbar_init(&st_empty[i], 1); }     // former warps of a stage + its B warp
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, 16);                  // the four groups on drain duty
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == W_MMA) {       // TMEM allocation by one full warp
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_smem)), "r"((uint32_t)TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int e = tid; e < (NT / 4) * MR; e += NTHR) accs[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;

    // contiguous chunk range of this CTA; everything below counts in 32 bits
    const long long c_begin = p.nchunk * blockIdx.x / gridDim.x, c_end = p.nchunk * (blockIdx.x + 1) / gridDim.x;
    const int nch = (int)(c_end - c_begin);
    const int FC = p.flush_chunks;

    if (warp == W_TMA) {
        // ------------------------------------------------------------------ TMA producer (rows spread over the lanes)
        const float* src0 = lane < S ? p.fields[lane] : nullptr;
        const float* src1 = lane + 32 < S ? p.fields[lane + 32] : nullptr;
        int buf = 0;
        uint32_t ph = 0;
        for (int c = 0; c < nch; ++c) {
            mbar_wait(&chunk_empty[buf], ph ^ 1, 5u * p.backoff_ns);    // whole warp waits: uniform control flow; a chunk lasts ~8 us
            if (elect_one()) mbar_arrive_expect_tx(&chunk_full[buf], (p.debug & 16) ? 0u : (uint32_t)(S * XCH * 4));
            __syncwarp();
            if (p.debug & 16) { if (++buf == NCHUNKBUF) { buf = 0; ph ^= 1; } continue; }
            const long long x0 = (c_begin + c) * XCH;
            uint32_t* dst = chunk + (size_t)buf * S * ROWF;
            if (src0) bulk_g2s(dst + (size_t)lane * ROWF, src0 + x0, XCH * 4, &chunk_full[buf]);
            if (src1) bulk_g2s(dst + (size_t)(lane + 32) * ROWF, src1 + x0, XCH * 4, &chunk_full[buf]);
            for (int f = lane + 64; f < S; f += 32) bulk_g2s(dst + (size_t)f * ROWF, p.fields[f] + x0, XCH * 4, &chunk_full[buf]);
            if (++buf == NCHUNKBUF) { buf = 0; ph ^= 1; }
        }
    } else if (warp == W_MMA) {
        // ------------------------------------------------------------------ MMA issuer
        {
            // the whole warp runs the loop (uniform control flow); one elected lane issues
            const uint32_t idesc = umma_idesc_f16(NT);
            const uint32_t b_lbo = (uint32_t)NT * 16u;
            const uint64_t dbh0 = umma_desc(smem_u32(b_hi), b_lbo, 128), dbl0 = umma_desc(smem_u32(b_lo), b_lbo, 128);
            const uint64_t dstep = (uint64_t)(b_stage_bytes >> 4);          // start-address field advances by one stage
            const uint32_t tc_ = (uint32_t)p.tile_cols;
            int cf = 0;
            uint32_t eph = 0;
            bool ready = false;
            const int nsub = nch * NSUB;
            for (int c = 0; c < nsub; ++c) {                                       // c counts 64-cell sub-chunks here
                const uint32_t ph = (uint32_t)(c & 1);
                if (cf == 0 && c > 0) { mbar_wait(acc_empty, eph); eph ^= 1; }      // previous period drained
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const uint32_t acc = (cf == 0 && g == 0) ? 0u : 1u;
                    if (p.trace && blockIdx.x == 0 && c < 64 && lane == 0) p.trace[(0 * 8 + 4 + g) * 64 + c] = clock64();
                    if (!ready) mbar_wait(&st_full[g], ph);
                    if (p.trace && blockIdx.x == 0 && c < 64 && lane == 0) p.trace[(0 * 8 + g) * 64 + c] = clock64();
                    tc_fence_after();
                    // probe the next stage now, look at the answer after this stage's MMAs are issued
                    ready = mbar_test(&st_full[(g + 1) & 3], g == 3 ? (ph ^ 1u) : ph);
                    const uint64_t dbh = dbh0 + dstep * (uint64_t)g, dbl = dbl0 + dstep * (uint64_t)g;
                    const uint32_t a0 = tmem_base + (uint32_t)(TMEM_A0 + g * 64);
                    if (elect_one()) {
#pragma unroll
                        for (int m = 0; m < 4; ++m) {
                            if (m < MT && !(p.debug & 1)) {
                                const uint32_t d = tmem_base + (uint32_t)m * tc_;
                                const uint32_t ah = a0 + (uint32_t)(m * 16);
                                umma_f16_ts(d, ah, dbh, idesc, acc);
                                umma_f16_ts(d, ah, dbl, idesc, 1u);
                                umma_f16_ts(d, ah + 8u, dbh, idesc, 1u);
                            }
                        }
                        umma_commit(&st_empty[g]);                               // operands of this stage consumed
                        if (g == 3 && (cf + 1 == FC || c == nsub - 1)) umma_commit(acc_full);
                    }
                    __syncwarp();
                }
                if (++cf == FC) cf = 0;
            }
        }
    } else if (warp >= NFWARPS) {
        // ------------------------------------------------------------------ B-operand former: fields -> fp16 hi/lo tiles
        int buf = 0;
        uint32_t cph = 0;
        const int nbcell = 2 * NT;                               // (l, kchunk) cells of the B tile per K-step
        for (int c = 0; c < nch; ++c) {
            mbar_wait(&chunk_full[buf], cph, p.backoff_ns);
            for (int sub = 0; sub < NSUB; ++sub) {
            const uint32_t* ch = chunk + (size_t)buf * S * ROWF + sub * SUB;
            const uint32_t sph = (uint32_t)((c * NSUB + sub) & 1);
            for (int g = warp - NFWARPS; g < 4; g += NBW) {       // this warp's operand stages (K-step g of every sub-chunk)
                const int tt = c * NSUB + sub;
                if (p.trace && blockIdx.x == 0 && tt < 64 && lane == 0 && g == 0) p.trace[(2 * 8 + 0) * 64 + tt] = clock64();
                mbar_wait(&st_empty[g], sph ^ 1, p.backoff_ns);     // ~2600 cycles of slack per stage
                if (p.trace && blockIdx.x == 0 && tt < 64 && lane == 0 && g == 0) p.trace[(2 * 8 + 1) * 64 + tt] = clock64();
                // lane <-> field: rows are 260 words apart (== 4 mod 32), so 8 consecutive fields hit 8 distinct bank groups
                for (int cell = lane; cell < ((p.debug & 32) ? 0 : nbcell); cell += 32) {
                    const int kc = cell >= NT ? 1 : 0, l = cell - kc * NT;
                    uint4 h4 = make_uint4(0, 0, 0, 0), l4 = h4;
                    if (l < S) {                       // 8 cells = 4 x {hi2, lo2}: pure data movement, K5 already split the fields
                        const uint32_t* pl = ch + (size_t)l * ROWF + g * 16 + kc * 8;
                        const uint4 a = *reinterpret_cast<const uint4*>(pl), b = *reinterpret_cast<const uint4*>(pl + 4);
                        h4 = make_uint4(a.x, a.z, b.x, b.z);
                        l4 = make_uint4(a.y, a.w, b.y, b.w);
                    }
                    const size_t off = (size_t)g * b_stage_bytes + (size_t)kc * NT * 16 + (size_t)l * 16;
                    *reinterpret_cast<uint4*>(b_hi + off) = h4;
                    *reinterpret_cast<uint4*>(b_lo + off) = l4;
                }
                if (p.trace && blockIdx.x == 0 && tt < 64 && lane == 0 && g == 0) p.trace[(2 * 8 + 2) * 64 + tt] = clock64();
                fence_proxy_async();                  // generic-proxy smem writes -> visible to the tensor core
                __syncwarp();
                if (lane == 0) mbar_arrive(&st_full[g]);
                if (p.trace && blockIdx.x == 0 && tt < 64 && lane == 0 && g == 0) p.trace[(2 * 8 + 3) * 64 + tt] = clock64();
            }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&chunk_empty[buf]);
            if (++buf == NCHUNKBUF) { buf = 0; cph ^= 1; }
        }
    } else {
        // ------------------------------------------------------------------ A-operand formers (+ TMEM drain)
        // Work unit n = (sub-chunk t = n >> 3, K-step k = (n >> 1) & 3, tile pair tp = n & 1): tiles {2tp, 2tp+1} of operand stage k.
        // Group g = warp >> 2 takes the units n = g (mod NG).  With NG = 4 that is the fixed pairing (k, k+2; tp) of one tile pair
        // with two alternating stages; with NG = 6 the 24 former warps rotate through all (stage, tile pair) combinations.
        // A thread is one TMEM lane (32*q + lane) in every tile: the rows of a lane share the field i, which is loaded once per unit.
        const int g = warp >> 2, q = warp & 3;
        const int tl = q * 32 + lane;                            // TMEM lane
        const int* lij = p.lane_ij + tl * 5;
        const int fi = lij[0];
        const uint32_t fio = (uint32_t)(fi < 0 ? 0 : fi) * ROWF;
        uint32_t fjo[4];
        bool rv[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) { const int fj = lij[1 + m]; rv[m] = fi >= 0 && fj >= 0 && m < MT; fjo[m] = (uint32_t)(fj < 0 ? 0 : fj) * ROWF; }
        const uint32_t lane_quarter = (uint32_t)(q * 32) << 16;
        const uint32_t t_a = tmem_base + lane_quarter + (uint32_t)TMEM_A0;
        const int nsub = nch * NSUB;
        const int nperiods = (nsub + FC - 1) / FC;
        int cur_chunk = -1, buf = 0, dP = 0;                     // dP = periods this warp has passed (drained or skipped)
        // drain duty of period P: tile m is drained by group (m + P) % NG (all four groups for NG = 4: tile g)
        auto drain_period = [&](int P, bool fin) {
            int m = g - P % NG; if (m < 0) m += NG;
            if (m < 4) drain_accumulators(p, acc_full, acc_empty, P, fin, m < MT, tmem_base + lane_quarter + (uint32_t)(m * p.tile_cols),
                                          accs, MR, NT, m * 128 + tl, lane);
        };
        if constexpr (BLK2) {
            // 2x2 row blocks: the lane's rows are (i0,j0), (i0,j1), (i1,j0), (i1,j1) in tiles 0..3 (lane_ij = {i0, i1, j0, j1}); a unit is
            // one K-step of ALL four tiles: four field vectors for four rows of products (1 shared-memory word per product instead
            // of 1.5).  Group g = n % NG with n = 4 t + k.
            const int fi1 = lij[1];
            const uint32_t fio1 = (uint32_t)(fi1 < 0 ? 0 : fi1) * ROWF;
            const int j0 = lij[2], j1 = lij[3];
            const uint32_t fj0o = (uint32_t)(j0 < 0 ? 0 : j0) * ROWF, fj1o = (uint32_t)(j1 < 0 ? 0 : j1) * ROWF;
            const bool v00 = fi >= 0 && j0 >= 0, v01 = fi >= 0 && j1 >= 0, v10 = fi1 >= 0 && j0 >= 0, v11 = fi1 >= 0 && j1 >= 0;
            for (int n = g; n < nsub * 4; n += NG) {
                const int t = n >> 2, kk = n & 3;
                const int c = t / NSUB, sub = t - c * NSUB;
                if (c != cur_chunk) {
                    if (cur_chunk >= 0) { __syncwarp(); if (lane == 0) mbar_arrive(&chunk_empty[buf]); }
                    cur_chunk = c;
                    buf = c % NCHUNKBUF;
                    mbar_wait(&chunk_full[buf], (uint32_t)((c / NCHUNKBUF) & 1));
                }
                const uint32_t* ch = chunk + (size_t)buf * S * ROWF + sub * SUB + kk * 16;
                mbar_wait(&st_empty[kk], (uint32_t)(t & 1) ^ 1);
                tc_fence_after();
                uint4 va[4], vb[4];
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) { va[q4] = *reinterpret_cast<const uint4*>(ch + fio + q4 * 4); vb[q4] = *reinterpret_cast<const uint4*>(ch + fio1 + q4 * 4); }
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                    uint4 vj[4];
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) vj[q4] = *reinterpret_cast<const uint4*>(ch + (jj ? fj1o : fj0o) + q4 * 4);
#pragma unroll
                    for (int ii = 0; ii < 2; ++ii) {
                        uint32_t hi[8], lo[8];
                        const bool ok = ii ? (jj ? v11 : v10) : (jj ? v01 : v00);
                        if (ok) {
#pragma unroll
                            for (int q4 = 0; q4 < 4; ++q4) {
                                const uint4 a = ii ? vb[q4] : va[q4];
                                prod_split(a.x, a.y, vj[q4].x, vj[q4].y, hi[2 * q4], lo[2 * q4]);
                                prod_split(a.z, a.w, vj[q4].z, vj[q4].w, hi[2 * q4 + 1], lo[2 * q4 + 1]);
                            }
                        } else {
#pragma unroll
                            for (int k = 0; k < 8; ++k) { hi[k] = 0u; lo[k] = 0u; }
                        }
                        tmem_st16(t_a + (uint32_t)(kk * 64 + (2 * ii + jj) * 16), hi, lo);
                    }
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&st_full[kk]);
                if (dP < t / FC) { drain_period(dP, false); ++dP; }
            }
        } else {
        for (int n = g; n < nsub * 8; n += NG) {
            const int t = n >> 3, kk = (n >> 1) & 3, tp = n & 1;
            const int c = t / NSUB, sub = t - c * NSUB;
            if (c != cur_chunk) {
                if (cur_chunk >= 0) { __syncwarp(); if (lane == 0) mbar_arrive(&chunk_empty[buf]); }     // all reads of the old buffer are done
                cur_chunk = c;
                buf = c % NCHUNKBUF;
                mbar_wait(&chunk_full[buf], (uint32_t)((c / NCHUNKBUF) & 1));
            }
            const uint32_t* ch = chunk + (size_t)buf * S * ROWF + sub * SUB;
            const bool tr = p.trace && blockIdx.x == 0 && t < 64 && lane == 0 && (warp == 0 || warp == 12);
            const int tre = ((warp == 0 ? 1 : 3) * 8 + 4 * (kk >> 1)) * 64 + t;
            if (tr) p.trace[tre] = clock64();
            mbar_wait(&st_empty[kk], (uint32_t)(t & 1) ^ 1);
            if (tr) p.trace[tre + 64] = clock64();
            if (!(p.debug & 256)) tc_fence_after();
            // (Issuing the operand loads before the barrier wait, or in longer bursts, shortens this warp's K-step but slows the
            // whole kernel by 3-20%: measured several ways, profiles/r1_summary.md.)
            uint4 vi[4];                       // 16 cells of I_i: per 4 cells {hi2(0,1), lo2(0,1), hi2(2,3), lo2(2,3)}
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) vi[q4] = (p.debug & 512) ? make_uint4(0, 0, 0, 0) : *reinterpret_cast<const uint4*>(ch + fio + kk * 16 + q4 * 4);
#pragma unroll
            for (int mm = 0; mm < 2; ++mm) {
                const int m = 2 * tp + mm;
                if (m < MT) {
                    uint32_t hi[8], lo[8];
                    const bool rvm = tp ? (mm ? rv[3] : rv[2]) : (mm ? rv[1] : rv[0]);
                    if (rvm && !(p.debug & 2)) {
                        const uint32_t* pj = ch + (tp ? (mm ? fjo[3] : fjo[2]) : (mm ? fjo[1] : fjo[0])) + kk * 16;
#pragma unroll
                        for (int q4 = 0; q4 < 4; ++q4) {
                            const uint4 b = *reinterpret_cast<const uint4*>(pj + q4 * 4);
                            prod_split(vi[q4].x, vi[q4].y, b.x, b.y, hi[2 * q4], lo[2 * q4]);
                            prod_split(vi[q4].z, vi[q4].w, b.z, b.w, hi[2 * q4 + 1], lo[2 * q4 + 1]);
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < 8; ++k) { hi[k] = 0u; lo[k] = 0u; }       // padding rows
                    }
                    if (!(p.debug & 4)) {
                        if (p.debug & 1024) {
                            tmem_st8(t_a + (uint32_t)(kk * 64 + m * 16), hi);
                            tmem_st8(t_a + (uint32_t)(kk * 64 + m * 16 + 8), lo);
                        } else {
                            tmem_st16(t_a + (uint32_t)(kk * 64 + m * 16), hi, lo);
                        }
                    }
                }
            }
            if (tr) p.trace[tre + 3 * 64] = clock64();
            if (!(p.debug & 4)) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            if (!(p.debug & 256)) tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&st_full[kk]);
            if (tr) p.trace[tre + 2 * 64] = clock64();
            // Drain the previous period once this warp has formed a unit of the next one: its MMAs are done by now (no wait on
            // acc_full) while the tensor core still has this sub-chunk's K-steps queued.  Every group owns a unit of every sub-chunk.
            if (dP < t / FC) { drain_period(dP, false); ++dP; }
        }
        }
        if (cur_chunk >= 0) { __syncwarp(); if (lane == 0) mbar_arrive(&chunk_empty[buf]); }
        while (dP < nperiods) { drain_period(dP, dP == nperiods - 1); ++dP; }
    }
    // ---------------------------------------------------------------------- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

__global__ void k_tri_tc_fold(const double* __restrict__ partial, int ncta, int MR, int NT, const int* __restrict__ tri_rc,
                              int ntri, double* sums)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntri) return;
    const int r = tri_rc[2 * t], c = tri_rc[2 * t + 1];
    if (r < 0) return;                                   // triangle belongs to another pass
    double s = 0.0;
    for (int b = 0; b < ncta; ++b) s += partial[((size_t)b * NT + c) * MR + r];
    sums[t] = s;
}

}  // namespace tc

// tuning / profiling knobs, read from the environment ONCE (not per launch)
struct TcEnv {
    int groups = 4, debug = 0, flush_chunks = 4, gflush_drains = 64;
    unsigned backoff_ns = 0;
    double bias_per_mma = 1.83e-8;           // measured: -8.8e-7 of |S| after 48 accumulating MMAs (profiles/r1_summary.md)
    const char* trace_path = nullptr;
    TcEnv()
    {
        if (const char* e = getenv("PSB_TC_GROUPS")) { if (atoi(e) == 6) groups = 6; }
        if (const char* e = getenv("PSB_TC_BACKOFF")) backoff_ns = (unsigned)strtoul(e, nullptr, 10);
        if (const char* e = getenv("PSB_TC_DEBUG")) debug = atoi(e);
        if (const char* e = getenv("PSB_TC_FLUSH")) { int v = atoi(e); if (v >= 1 && v <= 4096) flush_chunks = v; }
        if (const char* e = getenv("PSB_TC_GFLUSH")) { int v = atoi(e); if (v >= 1 && v <= 65536) gflush_drains = v; }
        if (const char* e = getenv("PSB_TC_BIAS")) bias_per_mma = atof(e);
        trace_path = getenv("PSB_TC_TRACE");
    }
};
static const TcEnv& tc_env() { static const TcEnv e; return e; }

size_t triangle_tc_workspace_bytes(int MT, int NT) { return (size_t)sm_count() * MT * 128 * NT * sizeof(double); }

// One pass: 128 lanes x MT tiles of pair rows (lane_ij) against all NT columns.  tri_rc[t] = (row, col) or (-1,-1).
int triangle_sums_tc_pass(const float* const* fields, int S, long long ncell, const int* lane_ij, int lane_layout, int MT, int NT,
                          const int* tri_rc, int ntri, double* sums, void* ws, size_t ws_bytes, cudaStream_t st)
{
    using namespace tc;
    const int tile_cols = NT <= 64 ? 64 : NT;           // accumulator tiles packed at NT columns: 3 tiles for 65..80 shells, 2 above
    if (S < 1 || S > NT || NT % 16 || NT > 128 || MT < 1 || MT > 256 / tile_cols || ncell % 64) return PSB_ERR_ARG;
    const TcEnv& env = tc_env();
    const int ncta = sm_count();
    if (ncell / 64 / ncta >= (1LL << 28)) return PSB_ERR_ARG;
    if (ws_bytes < triangle_tc_workspace_bytes(MT, NT)) return PSB_ERR_WORKSPACE;
    const int MR = MT * 128;
    auto smem_for = [&](int xch) {
        const size_t chunk_bytes = (size_t)S * (xch + 4) * 4;
        size_t b = ((NCHUNKBUF * chunk_bytes + 1023) / 1024) * 1024;
        b += 2 * (size_t)4 * (2 * NT * 16);
        b += (size_t)NT * MR * 4;
        b += (2 * NCHUNKBUF + 8 + 2) * sizeof(uint64_t) + 16;
        return b;
    };
    int XCH = 256;
    if (smem_for(256) > 227 * 1024 || ncell % 256) XCH = 64;
    const size_t smem = smem_for(XCH);
    if (smem > 227 * 1024 || ncell % XCH) return PSB_ERR_ARG;
    int NG = env.groups;                                   // former groups: 4 (640 threads) or 6 (960 threads, 64 registers)
    void (*kern)(Params) = nullptr;
    if (lane_layout == 1) { if (MT != 4) return PSB_ERR_ARG; NG = 4; kern = XCH == 256 ? k_tri_tc<256, 4, true> : k_tri_tc<64, 4, true>; }
    else if (NG == 6) kern = XCH == 256 ? k_tri_tc<256, 6, false> : k_tri_tc<64, 6, false>;
    else kern = XCH == 256 ? k_tri_tc<256, 4, false> : k_tri_tc<64, 4, false>;
    const int NTHR = NG == 6 ? Roles<6>::NTHR : Roles<4>::NTHR;
    const long long nchunk_launch = ncell / XCH;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return PSB_ERR_CUDA;
    if (cudaMemsetAsync(ws, 0, triangle_tc_workspace_bytes(MT, NT), st) != cudaSuccess) return PSB_ERR_CUDA;
    Params p;
    p.fields = fields; p.S = S; p.NT = NT; p.MT = MT; p.tile_cols = tile_cols; p.lane_ij = lane_ij; p.nchunk = nchunk_launch;
    p.partial = static_cast<double*>(ws);
    p.debug = env.debug;
    p.backoff_ns = env.backoff_ns;
    p.flush_chunks = env.flush_chunks;          // default 4: 16 K-steps = 48 accumulating MMAs per accumulator between round-to-nearest drains
    p.gflush_drains = env.gflush_drains;
    p.bias_gain = (float)(1.0 + env.bias_per_mma * 12.0 * env.flush_chunks);
    p.trace = nullptr;
    const char* trace_path = env.trace_path;
    if (trace_path) { cudaMalloc(&p.trace, 5 * 8 * 64 * sizeof(long long)); cudaMemset(p.trace, 0, 5 * 8 * 64 * sizeof(long long)); }
    kern<<<ncta, NTHR, smem, st>>>(p);
    if (trace_path) {
        static long long host_trace[5 * 8 * 64];
        cudaMemcpy(host_trace, p.trace, sizeof(host_trace), cudaMemcpyDeviceToHost);
        if (FILE* f = fopen(trace_path, "wb")) { fwrite(host_trace, 1, sizeof(host_trace), f); fclose(f); }
        cudaFree(p.trace);
    }
    k_tri_tc_fold<<<(ntri + 255) / 256, 256, 0, st>>>(p.partial, ncta, MR, NT, tri_rc, ntri, sums);
    return cudaGetLastError() == cudaSuccess ? PSB_OK : PSB_ERR_CUDA;
}

}  // namespace psb
