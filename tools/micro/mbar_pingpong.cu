// Microbenchmark (not product code): round-trip latency of an mbarrier hand-off between two warps of one CTA,
// for suspended try_wait (with / without a time hint) and for test_wait polling.  Build: nvcc -arch=sm_100a -shared.
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t sa(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void arrive(uint64_t* b) { asm volatile("{\n\t.reg .b64 s;\n\tmbarrier.arrive.shared::cta.b64 s, [%0];\n\t}" ::"r"(sa(b)) : "memory"); }
template <int MODE> __device__ __forceinline__ void wait(uint64_t* b, uint32_t par)
{
    uint32_t done = 0;
    while (!done) {
        if (MODE == 0) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0,1,0,p;\n\t}" : "=r"(done) : "r"(sa(b)), "r"(par) : "memory");
        if (MODE == 1) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0,1,0,p;\n\t}" : "=r"(done) : "r"(sa(b)), "r"(par), "r"(20000u) : "memory");
        if (MODE == 2) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0,1,0,p;\n\t}" : "=r"(done) : "r"(sa(b)), "r"(par) : "memory");
    }
}
template <int MODE> __global__ void pingpong(int iters, int nextra, long long* out)
{
    __shared__ uint64_t A, B, C;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sa(&A)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sa(&B)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sa(&C)));
    }
    __syncthreads();
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (w == 0 && l == 0) {
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) { arrive(&A); wait<MODE>(&B, i & 1); }
        out[blockIdx.x] = (clock64() - t0) / iters;
        arrive(&C);
    } else if (w == 1 && l == 0) {
        for (int i = 0; i < iters; ++i) { wait<MODE>(&A, i & 1); arrive(&B); }
    } else if (w >= 2 && w < 2 + nextra && l == 0) {
        wait<MODE>(&C, 0);                       // bystander warps parked on an unrelated barrier
    }
}
extern "C" int run_pingpong(int mode, int iters, int nextra, long long* host_out)
{
    long long* d; cudaMalloc(&d, 148 * sizeof(long long));
    if (mode == 0) pingpong<0><<<148, 32 * (2 + nextra)>>>(iters, nextra, d);
    if (mode == 1) pingpong<1><<<148, 32 * (2 + nextra)>>>(iters, nextra, d);
    if (mode == 2) pingpong<2><<<148, 32 * (2 + nextra)>>>(iters, nextra, d);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(host_out, d, 148 * sizeof(long long), cudaMemcpyDeviceToHost); cudaFree(d);
    return (int)e;
}
