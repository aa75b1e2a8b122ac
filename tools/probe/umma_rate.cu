// cycles per tcgen05.mma 128xNx16 (bf16, cta_group::1), back-to-back issue
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr)
{
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
extern "C" __global__ void __launch_bounds__(128) k_rate(int N, int reps, int ntiles, long long *out, int rowoff)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t holder;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&holder)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t taddr = holder;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 96 * 1024;
        long long t0 = clock64();
        for (int r = 0; r < reps; r++) {
            for (int m = 0; m < ntiles; m++) {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    uint64_t da = make_desc(a0 + ((r * 7 + m) % 5) * 128 * 128 + rowoff * 128 * (1 + (r % 3)) + k * 32), db = make_desc(b0 + k * 32);
                    uint32_t acc = 1;
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(taddr + m * N),
                                 "l"(da), "l"(db), "r"(idesc), "r"(acc));
                }
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u));
        long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(512u));
}
int main()
{
    long long *out; cudaMallocManaged(&out, 148 * 8);
    cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int rowoff : {0, 1, 5}) for (int grid : {148}) for (int N : {64, 128, 192, 256}) {
        int ntiles = 512 / N > 4 ? 4 : 512 / N, reps = 200;
        k_rate<<<grid, 128, 200 * 1024>>>(N, reps, ntiles, out, rowoff);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
        double cyc = (double)out[0] / (reps * ntiles * 4);
        printf("rowoff %d grid %3d N=%3d: %.1f cycles per 128xNx16 MMA -> %.0f FLOP/cycle/SM\n", rowoff, grid, N, cyc, 2.0 * 128 * N * 16 / cyc);
    }
    return 0;
}
