// How the tcgen05 pipe reacts to what the issuing thread does between MMAs (B200).
// Per "slab": 12 MMAs 128xNx16 on one accumulator, then `commits` tcgen05.commit, then a gap of
// `gap` dependent IMADs.  Prints cycles per slab.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr)
{
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ bool elect()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(d), "l"(da), "l"(db), "r"(idesc) : "memory");
}
__device__ __forceinline__ uint32_t idesc_n(int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24); }
// pattern 0: 12 x N      pattern 1: 12 x (N=128 at col 384, N=64 at col 0)     pattern 2: 12 x N alternating two accumulators
extern "C" __global__ void __launch_bounds__(128) k_gap(int N, int pattern, int commits, int gap, int nslabs, long long *out)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar, dummy[16];
    __shared__ uint32_t holder;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u + i * 7;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        for (int i = 0; i < 16; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&dummy[i])));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&holder)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t taddr = holder;
    if (warp == 1) {
        const uint32_t w0 = smem_u32(smem), in0 = smem_u32(smem) + 72 * 1024;
        const uint32_t iN = idesc_n(N), i128 = idesc_n(128), i64 = idesc_n(64);
        unsigned x = 1;
        long long t0 = clock64(), tg = 0;
        for (int s = 0; s < nslabs; s++) {
            const uint64_t da0 = make_desc(in0 + (s & 3) * 18432 + 7 * 128);
            const uint64_t db0 = make_desc(w0);
            if (elect()) {
#pragma unroll
                for (int m = 0; m < 12; m++) {
                    const uint64_t da = da0 + (m >> 2) * 8 + (m & 3) * 2, db = db0 + (m >> 2) * 192 * 8 + (m & 3) * 2;
                    if (pattern == 0) mma(taddr + (s & 1) * 256, da, db, iN);
                    else if (pattern == 1) { mma(taddr + 384, da, db, i128); mma(taddr, da, db + 128 * 8, i64); }
                    else mma(taddr + (m & 1) * 256, da, db, iN);
                }
                if (commits >= 1)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&dummy[s & 7])));
                if (commits >= 2)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&dummy[8 + (s & 7)])));
            }
            __syncwarp();
            long long g0 = clock64();
            for (int i = 0; i < gap; i++) x = x * 1664525u + 1013904223u;
            if (x == 12345u) out[200] = 1;
            tg += clock64() - g0;
        }
        if (elect())
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u));
        long long t1 = clock64();
        if (tid == 32) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = tg; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(512u));
}
int main()
{
    long long *out; cudaMallocManaged(&out, 512 * 8);
    cudaFuncSetAttribute(k_gap, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int nslabs = 300;
    auto run = [&](int N, int pattern, int commits, int gap) {
        k_gap<<<148, 128, 200 * 1024>>>(N, pattern, commits, gap, nslabs, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); exit(1); }
        printf("N=%3d pattern %d commits %d gap %3d (%4.0f cycles): %7.1f cycles per slab\n", N, pattern, commits, gap,
               (double)out[1] / nslabs, (double)out[0] / nslabs);
    };
    for (int N : {64, 128, 192}) run(N, 0, 0, 0);
    run(192, 1, 0, 0);
    run(192, 2, 0, 0);
    for (int commits : {0, 1, 2}) for (int gap : {0, 20, 40, 80, 160}) run(192, 0, commits, gap);
    for (int commits : {1, 2}) run(192, 1, commits, 0);
    return 0;
}
