// tcgen05.mma.ws (weight-stationary, B held in a collector buffer) on B200: does it take N = 192,
// does "::use" really skip the shared-memory read of B, and at what rate does it run?
// Part 1 (1 CTA): D1 = A1 * B^T with collector::b0::fill, then D2 = A2 * B^T with collector::b0::use
// and a B descriptor that points at ZEROS; both read back and compared with the host.
// Part 2 (148 CTAs): cycles per MMA, back-to-back, fill every time | fill + (R-1) x use.
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
#define MMA_WS(op, d, da, db, idesc, acc)                                                              \
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"                                     \
                 "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::" op " [%0], %1, %2, %3, p;\n\t}\n" \
                 ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory")
#define MMA_PLAIN(d, da, db, idesc, acc)                                                               \
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"                                     \
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"                      \
                 ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory")

// smem: A1 at 0 (16 KB), A2 at 16 KB, B at 32 KB (32 KB), zeros at 64 KB (32 KB)
extern "C" __global__ void __launch_bounds__(128)
k_ws(int N, int mode, int reps, int reuse, const float *a1, const float *a2, const float *bv, float *out, long long *cycles)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t holder;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0;
    __syncthreads();
    for (int i = tid; i < 128 * 16; i += 128) {
        const int r = i >> 4, k = i & 15;
        const int off = r * 128 + (((k >> 3) ^ (r & 7)) << 4);
        reinterpret_cast<__nv_bfloat16 *>(smem + off)[k & 7] = __float2bfloat16(a1 ? a1[r * 16 + k] : 1.0f);
        reinterpret_cast<__nv_bfloat16 *>(smem + 16384 + off)[k & 7] = __float2bfloat16(a2 ? a2[r * 16 + k] : 1.0f);
    }
    for (int i = tid; i < N * 16; i += 128) {
        const int r = i >> 4, k = i & 15;
        reinterpret_cast<__nv_bfloat16 *>(smem + 32768 + r * 128 + (((k >> 3) ^ (r & 7)) << 4))[k & 7] = __float2bfloat16(bv ? bv[r * 16 + k] : 1.0f);
    }
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
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t dA1 = make_desc(smem_u32(smem)), dA2 = make_desc(smem_u32(smem) + 16384);
    const uint64_t dB = make_desc(smem_u32(smem) + 32768), dZ = make_desc(smem_u32(smem) + 65536);
    if (tid == 0) {
        const long long t0 = clock64();
        if (mode == 0) {            // correctness: fill, then use with a B descriptor of zeros
            MMA_WS("fill", taddr, dA1, dB, idesc, 0);
            MMA_WS("lastuse", taddr + 256, dA2, dZ, idesc, 0);
        } else if (mode == 1) {     // plain MMA, back to back
            for (int r = 0; r < reps; r++) MMA_PLAIN(taddr + (r & 1) * 256, (r & 1) ? dA2 : dA1, dB, idesc, 1);
        } else if (mode == 2) {     // ws, fill every time
            for (int r = 0; r < reps; r++) MMA_WS("fill", taddr + (r & 1) * 256, (r & 1) ? dA2 : dA1, dB, idesc, 1);
        } else {                    // ws, one fill then reuse-1 uses
            for (int r = 0; r < reps; r += reuse) {
                MMA_WS("fill", taddr, dA1, dB, idesc, 1);
                for (int u = 1; u < reuse - 1; u++) MMA_WS("use", taddr + (u & 1) * 256, (u & 1) ? dA2 : dA1, dZ, idesc, 1);
                MMA_WS("lastuse", taddr + 256, dA2, dZ, idesc, 1);
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u));
        cycles[blockIdx.x] = clock64() - t0;
    }
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (mode == 0 && out) {
        for (int half = 0; half < 2; half++)
            for (int c = 0; c < N; c += 16) {
                uint32_t v[16];
                const uint32_t ta = taddr + ((uint32_t)(warp * 32) << 16) + half * 256 + c;
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                               "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                             : "r"(ta));
                asm volatile("tcgen05.wait::ld.sync.aligned;");
                for (int q = 0; q < 16; q++) out[(half * 128 + warp * 32 + lane) * 256 + c + q] = __uint_as_float(v[q]);
            }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(512u));
}

int main()
{
    float *a1, *a2, *b, *out; long long *cyc;
    cudaMallocManaged(&a1, 128 * 16 * 4); cudaMallocManaged(&a2, 128 * 16 * 4); cudaMallocManaged(&b, 256 * 16 * 4);
    cudaMallocManaged(&out, 256 * 256 * 4); cudaMallocManaged(&cyc, 148 * 8);
    srand(1);
    for (int i = 0; i < 128 * 16; i++) { a1[i] = (float)(rand() % 7 - 3); a2[i] = (float)(rand() % 5 - 2); }
    for (int i = 0; i < 256 * 16; i++) b[i] = (float)(rand() % 9 - 4);
    cudaFuncSetAttribute(k_ws, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    for (int N : {64, 128, 256}) {
        for (int i = 0; i < 256 * 256; i++) out[i] = -12345.f;
        k_ws<<<1, 128, 96 * 1024>>>(N, 0, 0, 0, a1, a2, b, out, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("N=%d: error %s\n", N, cudaGetErrorString(e)); return 1; }
        int bad1 = 0, bad2 = 0;
        for (int r = 0; r < 128; r++)
            for (int c = 0; c < N; c++) {
                float s1 = 0, s2 = 0;
                for (int k = 0; k < 16; k++) { s1 += a1[r * 16 + k] * b[c * 16 + k]; s2 += a2[r * 16 + k] * b[c * 16 + k]; }
                bad1 += out[r * 256 + c] != s1;
                bad2 += out[(128 + r) * 256 + c] != s2;
            }
        printf("ws N=%3d: fill: %d of %d wrong; use (B descriptor -> zeros): %d wrong%s\n", N, bad1, 128 * N, bad2,
               bad2 == 0 ? "  => B came from the collector" : "");
        if (bad1) { printf("   sample row 0:"); for (int c = 0; c < 8; c++) printf(" %g", out[c]); printf("\n"); }
    }
    for (int N : {128, 256}) {
        const int reps = 960;
        const char *names[] = {"", "plain", "ws fill each", "ws fill + 1 use", "ws fill + 3 uses", "ws fill + 11 uses"};
        for (int m = 1; m <= 5; m++) {
            const int mode = m <= 2 ? m : 3, reuse = m == 3 ? 2 : m == 4 ? 4 : 12;
            k_ws<<<148, 128, 96 * 1024>>>(N, mode, reps, reuse, nullptr, nullptr, nullptr, nullptr, cyc);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("N=%d mode %d: error %s\n", N, m, cudaGetErrorString(e)); return 1; }
            printf("N=%3d %-18s %.1f cycles per 128xNx16 MMA\n", N, names[m], (double)cyc[0] / reps);
        }
    }
    // N = 192 last: the launch faults ("an illegal instruction was encountered")
    k_ws<<<1, 128, 96 * 1024>>>(192, 0, 0, 0, a1, a2, b, out, cyc);
    printf("ws N=192: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
