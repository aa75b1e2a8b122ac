// does a tcgen05.mma whose D range runs past TMEM column 511 wrap to column 0?
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr)
{
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
extern "C" __global__ void __launch_bounds__(128) k_wrap(int dcol, int N, float *out)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t holder;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x3f803f80u;   // bf16 1.0
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
    const uint32_t tl = taddr + ((uint32_t)(warp * 32) << 16);
    const uint32_t z = 0;
    for (int c = 0; c < 512; c += 8)
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(tl + c), "r"(z));
    asm volatile("tcgen05.wait::st.sync.aligned;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(taddr + dcol),
                     "l"(make_desc(smem_u32(smem))), "l"(make_desc(smem_u32(smem) + 32 * 1024)), "r"(idesc));
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
    }
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u));
    asm volatile("tcgen05.fence::after_thread_sync;");
    for (int c = 0; c < 512; c += 8) {
        uint32_t v[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(tl + c));
        asm volatile("tcgen05.wait::ld.sync.aligned;");
        for (int q = 0; q < 8; q++) out[tid * 512 + c + q] = __uint_as_float(v[q]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(512u));
}
int main()
{
    float *out; cudaMallocManaged(&out, 128 * 512 * 4);
    cudaFuncSetAttribute(k_wrap, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int dcol : {0, 384, 448}) {
        k_wrap<<<1, 128, 100 * 1024>>>(dcol, 192, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("dcol %d: err %s\n", dcol, cudaGetErrorString(e)); return 1; }
        printf("dcol %3d N=192: nonzero column runs (lane 0 / lane 127):", dcol);
        for (int lane : {0, 127}) {
            int start = -1;
            for (int c = 0; c <= 512; c++) {
                bool nz = c < 512 && out[lane * 512 + c] != 0.f;
                if (nz && start < 0) start = c;
                if (!nz && start >= 0) { printf(" [%d,%d)=%g", start, c, out[lane * 512 + start]); start = -1; }
            }
            printf(" |");
        }
        printf("\n");
    }
    return 0;
}
