// Stage-0 probe for the tcgen05 path: D[128x64] = A[r0:r0+128, 0:64] * B[64x64]^T
// A lives in a taller SMEM buffer so that the descriptor start can be an
// arbitrary row (the implicit-GEMM tap shift).  bf16 in, fp32 out.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t base_offset)
{
    // K-major, SWIZZLE_128B: LBO = 1 (unused), SBO = 1024 B (8 rows x 128 B), version 1
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_offset & 7) << 49;
    d |= (uint64_t)2 << 61;
    return d;
}

// instruction descriptor: D f32, A/B bf16, K-major both, N=64, M=128
__device__ __forceinline__ uint32_t make_idesc(int M, int N)
{
    uint32_t d = 0;
    d |= 1u << 4;               // c_format f32
    d |= 1u << 7;               // a bf16
    d |= 1u << 10;              // b bf16
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;
}

extern "C" __global__ void __launch_bounds__(128)
k_umma_probe(const uint16_t *A, int a_rows, int r0, const uint16_t *B, float *D, int use_base_offset)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *sA = smem;                              // a_rows x 128 B (swizzled by absolute row)
    uint8_t *sB = smem + ((a_rows * 128 + 1023) & ~1023);   // 64 x 128 B
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_holder;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < a_rows * 8; i += 128) {
        int r = i >> 3, c = i & 7;
        *reinterpret_cast<uint4 *>(sA + r * 128 + ((c ^ (r & 7)) << 4)) =
            *reinterpret_cast<const uint4 *>(A + r * 64 + c * 8);
    }
    for (int i = tid; i < 64 * 8; i += 128) {
        int r = i >> 3, c = i & 7;
        *reinterpret_cast<uint4 *>(sB + r * 128 + ((c ^ (r & 7)) << 4)) =
            *reinterpret_cast<const uint4 *>(B + r * 64 + c * 8);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)), "r"(64u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");     // st.shared -> visible to the tensor core (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t taddr = tmem_holder;
    if (tid == 0) {
        const uint32_t idesc = make_idesc(128, 64);
        const uint32_t a0 = smem_u32(sA) + r0 * 128, b0 = smem_u32(sB);
        for (int k = 0; k < 4; k++) {
            uint64_t da = make_desc(a0 + k * 32, use_base_offset ? (uint32_t)(r0 & 7) : 0u);
            uint64_t db = make_desc(b0 + k * 32, 0u);
            uint32_t acc = k > 0;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(taddr),
                "l"(da), "l"(db), "r"(idesc), "r"(acc));
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
    }
    // wait for the MMAs
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u));
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    uint32_t v[64];
    const uint32_t ta = taddr + ((uint32_t)(warp * 32) << 16);
#pragma unroll
    for (int c = 0; c < 64; c += 16) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(v[c + 0]), "=r"(v[c + 1]), "=r"(v[c + 2]), "=r"(v[c + 3]), "=r"(v[c + 4]), "=r"(v[c + 5]),
              "=r"(v[c + 6]), "=r"(v[c + 7]), "=r"(v[c + 8]), "=r"(v[c + 9]), "=r"(v[c + 10]), "=r"(v[c + 11]),
              "=r"(v[c + 12]), "=r"(v[c + 13]), "=r"(v[c + 14]), "=r"(v[c + 15])
            : "r"(ta + c));
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    const int row = warp * 32 + lane;
    for (int c = 0; c < 64; c++) D[row * 64 + c] = __uint_as_float(v[c]);
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(64u));
}

extern "C" int umma_probe(const void *A, int a_rows, int r0, const void *B, float *D, int use_base_offset)
{
    size_t smem = ((a_rows * 128 + 1023) & ~1023) + 64 * 128 + 1024;
    cudaFuncSetAttribute(k_umma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_umma_probe<<<1, 128, smem>>>((const uint16_t *)A, a_rows, r0, (const uint16_t *)B, D, use_base_offset);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return -1; }
    return 0;
}
