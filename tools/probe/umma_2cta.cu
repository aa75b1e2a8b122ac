// tcgen05.mma.cta_group::2 on B200: operand split, result layout and rate (groundwork for a
// fused residual block: both layers' weights fit only if the pair shares the B operand).
//   A: each CTA of the pair holds its own 128 rows (M = 256 in total)
//   B: each CTA holds N/2 rows of the N x K operand
// Part 1 (1 cluster): one 256 x N x 16 MMA on small integers, D read back from both CTAs' TMEM and
// compared with the host.  Part 2 (74 clusters): cycles per MMA, 12 back-to-back per "slab".
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
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
__device__ __forceinline__ void cluster_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cta_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }

// A rows: [128][64] bf16 (128-byte rows, SW128), B rows: [N/2][64] bf16.  aval/bval: device arrays
// with the global operands A[256][16], B[N][16] (floats holding small integers); out: D [256][N].
extern "C" __global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
k_2cta(int N, int nslabs, const float *aval, const float *bval, float *out, long long *cycles)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t holder;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cta_rank();
    uint8_t *sa = smem, *sb = smem + 64 * 1024;
    for (int i = tid; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0;
    __syncthreads();
    // fill k = 0..15 of every row: 16-byte chunk j of row r at chunk j ^ (r & 7)
    for (int i = tid; i < 128 * 16; i += 128) {
        const int r = i >> 4, k = i & 15;
        const float v = aval ? aval[(rank * 128 + r) * 16 + k] : 1.0f;
        reinterpret_cast<__nv_bfloat16 *>(sa + r * 128 + (((k >> 3) ^ (r & 7)) << 4))[k & 7] = __float2bfloat16(v);
    }
    for (int i = tid; i < (N / 2) * 16; i += 128) {
        const int r = i >> 4, k = i & 15;
        const float v = bval ? bval[(rank * (N / 2) + r) * 16 + k] : 1.0f;
        reinterpret_cast<__nv_bfloat16 *>(sb + r * 128 + (((k >> 3) ^ (r & 7)) << 4))[k & 7] = __float2bfloat16(v);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&holder)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    cluster_sync();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t taddr = holder;
    // zero the accumulator columns
    {
        const uint32_t tl = taddr + ((uint32_t)(warp * 32) << 16);
        const uint32_t z = 0;
        for (int c = 0; c < 512; c += 8)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(tl + c), "r"(z));
        asm volatile("tcgen05.wait::st.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    cluster_sync();
    asm volatile("tcgen05.fence::after_thread_sync;");
    long long t0 = clock64();
    if (warp == 1 && rank == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((256u >> 4) << 24);
        const uint64_t da0 = make_desc(smem_u32(sa)), db0 = make_desc(smem_u32(sb));
        for (int s = 0; s < nslabs; s++) {
            if (elect()) {
                const int reps = nslabs == 1 ? 1 : 12;
#pragma unroll 1
                for (int m = 0; m < reps; m++)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
                                 "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                                 ::"r"(taddr), "l"(da0), "l"(db0), "r"(idesc) : "memory");
            }
            __syncwarp();
        }
        if (elect())
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                         ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
    }
    // both CTAs wait for the multicast arrival on their own barrier
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u));
    long long t1 = clock64();
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (tid == 32 && rank == 0 && cycles) cycles[blockIdx.x >> 1] = t1 - t0;
    if (out) {
        const uint32_t tl = taddr + ((uint32_t)(warp * 32) << 16);
        for (int c = 0; c < N; c += 8) {
            uint32_t v[8];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(tl + c));
            asm volatile("tcgen05.wait::ld.sync.aligned;");
            for (int q = 0; q < 8; q++) out[(size_t)(rank * 128 + tid) * N + c + q] = __uint_as_float(v[q]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    cluster_sync();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(512u));
}
int main()
{
    const int N = 192;
    float *a, *b, *out; long long *cyc;
    cudaMallocManaged(&a, 256 * 16 * 4); cudaMallocManaged(&b, 256 * 16 * 4);
    cudaMallocManaged(&out, 256 * 256 * 4); cudaMallocManaged(&cyc, 148 * 8);
    for (int r = 0; r < 256; r++) for (int k = 0; k < 16; k++) a[r * 16 + k] = (float)((r * 3 + k) % 7 - 3);
    for (int n = 0; n < 256; n++) for (int k = 0; k < 16; k++) b[n * 16 + k] = (float)((n * 5 + k * 2) % 5 - 2);
    cudaFuncSetAttribute(k_2cta, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    k_2cta<<<2, 128, 100 * 1024>>>(N, 1, a, b, out, nullptr);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("part 1: %s\n", cudaGetErrorString(e)); return 1; }
    int bad = 0, first = -1;
    for (int r = 0; r < 256; r++) for (int n = 0; n < N; n++) {
        float want = 0; for (int k = 0; k < 16; k++) want += a[r * 16 + k] * b[n * 16 + k];
        if (out[r * N + n] != want) { if (first < 0) first = r * N + n; bad++; }
    }
    printf("part 1: 256x%dx16 cta_group::2 MMA, A rows split by CTA, B rows [0,N/2) | [N/2,N) by CTA: %d mismatches of %d", N, bad, 256 * N);
    if (bad) printf(" (first at row %d col %d: got %g)", first / N, first % N, out[first]);
    printf("\n");
    for (int n : {128, 192, 256}) {
        const int nslabs = 300;
        k_2cta<<<148, 128, 100 * 1024>>>(n, nslabs, nullptr, nullptr, nullptr, cyc);
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("part 2: %s\n", cudaGetErrorString(e)); return 1; }
        printf("part 2: N=%3d: %.1f cycles per 256xNx16 MMA (cluster 0), %.1f (cluster 50) -> %.0f FLOP/cycle/SM\n", n,
               (double)cyc[0] / (nslabs * 12), (double)cyc[50] / (nslabs * 12), 2.0 * 256 * n * 16 / ((double)cyc[0] / (nslabs * 12)) / 2);
    }
    return 0;
}
