// az_tower.cuh -- 3x3 convolution (64 -> 64 channels) + bias (+ residual) +
// ReLU for the evaluator's tower on tcgen05 tensor cores (sm_100a).
//
// network.py:17-39 (Resblock) spends 99 % of the evaluator's FLOPs in twelve
// of these.  The kernel is an implicit GEMM with no im2col, no tensor maps
// and no cross-thread traffic in the epilogue.  Three ideas:
//
// (1) Slab layout.  Activations are 128-byte rows (64 bf16 channels).  One
// 128-row MMA tile ("slab") holds ONE board row y of `bpg` boards side by
// side, each board row padded with one zero cell:
//
//     slab q = group * n + y
//     row  l = b_local * (n+1) + x        (x == n: zero pad cell; l >= bpg*(n+1): zero)
//     R      = 8 + 128 * q + l            (global row; 8 zero rows in front)
//
// (bpg = 128 / (n+1): 10 boards at 11x11).  The x-neighbours of a cell are
// the rows R +- 1 and its y-neighbours are the SAME row of the slabs q +- 1
// of the same group.
//
// (2) Pointer-shift taps, N = 192.  The input of tap dx for all 128 rows of
// a slab is the same shared-memory tile with the UMMA descriptor start moved
// by dx rows (the hardware swizzles on absolute address bits, any row offset
// works: tools/probe/umma_probe.cu).  A 128xNx16 UMMA takes 64 cycles for
// any N <= 128 (tools/probe/umma_seq.cu), so 64 output channels alone idle
// half the tensor pipe; here the three dy-taps of one dx share the A operand
// against their stacked weights, N = 192 at full rate (96 cycles): 12 MMAs
// per slab.
//
// (3) TMEM accumulator ring.  Block dy of MMA(q) is the contribution of input
// slab q to output slab q + 1 - dy -- at the same TMEM lane.  Accumulators are
// a ring of eight 64-column blocks, output slab u at block (-u) mod 8, so the
// 192 columns of MMA(q) are exactly the accumulators of u = q+1, q, q-1 (two
// MMAs where the range wraps past column 511: the hardware faults instead of
// wrapping, tools/probe/tmem_wrap.cu).  The first and last board row of a
// group have no neighbour slab on one side: those MMAs are N = 128 over two
// blocks.  Every MMA accumulates; blocks are zeroed by the epilogue that
// retires them.  Output slab u is complete after MMA(u+1): the epilogue reads
// ONE block, adds bias (+ residual), ReLU, zeroes the pad cells, and the
// finished slab leaves as one 16 KB bulk store.  The ring gives the epilogue
// ~5 slabs of slack, so the tensor pipe never waits for it.
//
// Input and output move by cp.async.bulk: the layout is stored pre-swizzled
// in global memory (16-byte chunk j of row R at chunk j ^ (R & 7)), so an
// input slab plus its two neighbour rows is one contiguous 18 KB block.  The
// residual is read by the epilogue threads themselves (one sector per
// request); bulk-loading it into the staging tiles that the bulk stores read
// from faulted when two instances of the kernel ran concurrently
// (tools/probe/conv_concurrent.py), so those tiles only ever see generic
// writes and bulk-store reads.
//
// (4) Nothing but MMAs on the issuing thread.  The tcgen05 pipe does not
// queue ahead: every cycle the issuing thread spends on anything else
// (barrier polls, descriptor arithmetic) is a cycle the tensor cores idle,
// switching accumulators between consecutive MMAs costs ~33 cycles and a
// commit ~29 (tools/probe/umma_gap.cu).  So two issuer warps alternate slabs:
// while one issues its 12 back-to-back MMAs (one commit), the other waits for
// the next slab's barriers and computes its descriptors, then takes over
// through a named barrier.  The epilogue is two groups of eight warps on
// alternate output slabs for the same reason: one slab's read -> convert ->
// store chain is longer than a slab's MMA time.
//
// Roles (672 threads, one persistent CTA per SM, a contiguous range of board
// groups per CTA):
//   warps 0-7   epilogue group 0 (even output slabs): warp = (32 channels, 32 rows)
//   warps 8-15  epilogue group 1 (odd output slabs)
//   warps 16,17 MMA issuers (even / odd input slabs)
//   warp 18     one thread streams input slabs through a 5-stage ring
//   warp 19     spare
//   warp 20     one thread bulk-stores finished staging tiles
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define AZT_C 64                    // channels = one 128-byte swizzle row
#define AZT_ROW 128                 // bytes per row
#define AZT_HALO 8                  // zero rows in front of slab 0
#define AZT_WBYTES (9 * 64 * 128)   // one layer's weights: [dx][dy][c_out][c_in]
#define AZT_STAGES 5                // input ring
#define AZT_CHUNK_ROWS 144          // 8 + 128 + 8
#define AZT_CHUNK_BYTES (AZT_CHUNK_ROWS * AZT_ROW)
#define AZT_OUT_BYTES (128 * AZT_ROW)
#define AZT_OUT_STAGES 3            // staging slabs for finished output
#define AZT_SMEM_BYTES (AZT_WBYTES + AZT_STAGES * AZT_CHUNK_BYTES + AZT_OUT_STAGES * AZT_OUT_BYTES)
#define AZT_THREADS 672
#define AZT_BLOCKS 8                // TMEM ring: 8 x 64 columns

struct azt_params {
    const uint8_t *x;       // input activations, slab layout, pre-swizzled
    const uint8_t *w;       // [3 dx][192 = dy*64 + c_out][128 B] pre-swizzled weights
    const float *bias;      // [64]
    const uint8_t *resid;   // residual input (same layout) or NULL
    uint8_t *out;           // output activations
    int n;                  // board size
    int bpg;                // boards per group = 128 / (n+1)
    long long groups;       // board groups
    int debug;              // probe only: 2 = skip the output stores, 4 = skip the epilogue after the
                            // block retirement, 8 = skip the MMAs, 16 = skip the input loads
};

__device__ __forceinline__ uint32_t azt_smem(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ReLU + round to bf16 of two values in one instruction (lo in bits 0..15); the bits of
// __floats2bfloat162_rn(fmaxf(lo, 0), fmaxf(hi, 0)) for every finite input
__device__ __forceinline__ uint32_t azt_relu_bf16x2(float lo, float hi)
{
    uint32_t r;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

__device__ __forceinline__ void azt_mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(azt_smem(bar)), "r"(count));
}

__device__ __forceinline__ void azt_mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done) : "r"(azt_smem(bar)), "r"(parity) : "memory");
    }
}

__device__ __forceinline__ bool azt_elect()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void azt_mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(azt_smem(bar)) : "memory");
}

__device__ __forceinline__ void azt_mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(azt_smem(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void azt_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(azt_smem(dst)), "l"(src), "r"(bytes), "r"(azt_smem(bar)) : "memory");
}

__device__ __forceinline__ void azt_bulk_s2g(void *dst, const void *src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(dst), "r"(azt_smem(src)), "r"(bytes) : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start >> 4 | LBO 1 << 16 | SBO (1024 B >> 4) << 32 | version 1 << 46 | layout 2 << 61
__device__ __forceinline__ uint64_t azt_desc(uint32_t saddr)
{
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

// cute::UMMA::InstrDescriptor: D f32 (1 << 4), A bf16 (1 << 7), B bf16 (1 << 10),
// both K-major, N >> 3 at bit 17, M >> 4 at bit 24;  M = 128, N = 64 * blocks
#define AZT_IDESC(blocks) ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(blocks) << 20) | ((128u >> 4) << 24))

__device__ __forceinline__ void azt_mma(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc)
{
    // every MMA accumulates: ring blocks are zeroed by the epilogue that retires them
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(d), "l"(da), "l"(db), "r"(idesc) : "memory");
}

#define AZT_TMEM_LD16(v, addr)                                                                     \
    asm volatile(                                                                                  \
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                  \
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"                          \
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),      \
          "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),  \
          "=r"(v[14]), "=r"(v[15])                                                                 \
        : "r"(addr))

// 32 bytes (one sector) in one request: the two 16-byte chunks 2m, 2m+1 of a row
__device__ __forceinline__ void azt_ldg_sector(const void *p, uint4 &lo, uint4 &hi)
{
    unsigned long long a, b, c, d;
    asm volatile("ld.global.nc.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    lo = make_uint4((uint32_t)a, (uint32_t)(a >> 32), (uint32_t)b, (uint32_t)(b >> 32));
    hi = make_uint4((uint32_t)c, (uint32_t)(c >> 32), (uint32_t)d, (uint32_t)(d >> 32));
}

// logical chunks 4 half .. 4 half + 3 of a swizzled row (chunk j at j ^ sw) -> r[0..3]: two
// whole sectors, each requested once (four 16-byte loads would ask L2 for every sector twice)
__device__ __forceinline__ void azt_load_resid(const uint8_t *row, int half, int sw, uint4 (&r)[4])
{
#pragma unroll
    for (int m = 0; m < 2; m++) {
        const int pair = ((half * 4 + 2 * m) ^ sw) >> 1;            // physical sector of logical chunks 2k, 2k+1
        uint4 lo, hi;
        azt_ldg_sector(row + pair * 32, lo, hi);
        r[2 * m] = (sw & 1) ? hi : lo;
        r[2 * m + 1] = (sw & 1) ? lo : hi;
    }
}

__device__ __forceinline__ void azt_tmem_zero16(uint32_t addr)
{
    const uint32_t z = 0u;
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};"
        ::"r"(addr), "r"(z) : "memory");
}

template <bool RESID>
__global__ void __launch_bounds__(AZT_THREADS, 1)
k_conv3x3(const azt_params p)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *s_w = smem;                                            // 72 KB
    uint8_t *s_in = smem + AZT_WBYTES;                              // ring of input slabs (+ 8 rows each side)
    uint8_t *s_out = s_in + AZT_STAGES * AZT_CHUNK_BYTES;           // staging slabs
    __shared__ uint64_t bar_w, bar_in_full[AZT_STAGES];
    __shared__ uint64_t bar_mma_done[8];            // MMA(j) retired, by j & 7 (the MMA runs at most 7 slabs ahead)
    __shared__ uint64_t bar_blk_free[AZT_BLOCKS];   // ring block read, zeroed and free for its next output slab
    __shared__ uint64_t bar_out_empty[AZT_OUT_STAGES], bar_out_done[AZT_OUT_STAGES];
    __shared__ uint32_t tmem_holder;
    __shared__ __align__(16) float s_bias[AZT_C];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        azt_mbar_init(&bar_w, 1);
        for (int i = 0; i < AZT_STAGES; i++) azt_mbar_init(&bar_in_full[i], 1);
        for (int i = 0; i < 8; i++) azt_mbar_init(&bar_mma_done[i], 1);
        for (int i = 0; i < AZT_BLOCKS; i++) azt_mbar_init(&bar_blk_free[i], 8);    // one arrival per warp of a group
        for (int i = 0; i < AZT_OUT_STAGES; i++) {
            azt_mbar_init(&bar_out_empty[i], 1);        // released by the storer
            azt_mbar_init(&bar_out_done[i], 8);         // one arrival per warp of the group that wrote the slab
        }
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (tid < AZT_C) s_bias[tid] = p.bias[tid];
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(azt_smem(&tmem_holder)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_holder;

    // this CTA's contiguous range of groups -> slabs [q0, q0 + nslabs), local index j = q - q0;
    // board row y = j % n
    const int n = p.n;
    const long long g0 = p.groups * blockIdx.x / gridDim.x, g1 = p.groups * (blockIdx.x + 1) / gridDim.x;
    const long long q0 = g0 * n;
    const int nslabs = (int)((g1 - g0) * n);
    // ring block of the output slab with local index u
#define AZT_RING(u) ((8 - ((u) & 7)) & 7)

    if (warp < 16) {
        // zero the whole accumulator ring once: every MMA accumulates
        const uint32_t tq = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        for (int c = (warp >> 2) * 16; c < 512; c += 64) azt_tmem_zero16(tq + c);
        asm volatile("tcgen05.wait::st.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");

    if (warp >= 20) {
        // -------------------------------------------------------- storer --
        if (lane == 0 && !(p.debug & 4)) {
            for (int j = 0; j < nslabs; j++) {
                const int sb = j % AZT_OUT_STAGES;
                azt_mbar_wait(&bar_out_done[sb], (j / AZT_OUT_STAGES) & 1);
                if (!(p.debug & 2))
                    azt_bulk_s2g(p.out + (size_t)(AZT_HALO + (q0 + j) * 128) * AZT_ROW, s_out + sb * AZT_OUT_BYTES,
                                 AZT_OUT_BYTES);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                // the previous store has finished reading its staging slab: hand it back
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                if (j >= 1) azt_mbar_arrive(&bar_out_empty[(j - 1) % AZT_OUT_STAGES]);
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    } else if (warp == 19) {
        // (spare warp)
    } else if (warp == 18) {
        // -------------------------------------------------- input loader --
        if (lane == 0) {
            azt_mbar_expect_tx(&bar_w, AZT_WBYTES);
            for (int t = 0; t < 9; t++) azt_bulk_g2s(s_w + t * 8192, p.w + t * 8192, 8192, &bar_w);
            for (int j = 0; j < nslabs; j++) {
                const int st = j % AZT_STAGES;
                // the stage is free once the MMAs of the slab that used it last have retired
                if (j >= AZT_STAGES) azt_mbar_wait(&bar_mma_done[(j - AZT_STAGES) & 7], ((j - AZT_STAGES) >> 3) & 1);
                if (p.debug & 16) { azt_mbar_arrive(&bar_in_full[st]); continue; }
                // slab rows plus 8 rows on each side: global rows [128 q, 128 q + 144)
                azt_mbar_expect_tx(&bar_in_full[st], AZT_CHUNK_BYTES);
                azt_bulk_g2s(s_in + st * AZT_CHUNK_BYTES, p.x + (size_t)((q0 + j) * 128) * AZT_ROW,
                             AZT_CHUNK_BYTES, &bar_in_full[st]);
            }
        }
    } else if (warp >= 16) {
        // -------------------------------------------------- MMA issuers --
        // Warp 16 issues the even slabs, warp 17 the odd ones.  A whole warp runs the loop
        // (uniform control flow keeps descriptors in uniform registers: the 12 MMAs of a
        // slab are back-to-back instructions); one elected lane issues.  The turn passes
        // through named barriers 3 (-> even) and 4 (-> odd).
        const int w = warp - 16;
        azt_mbar_wait(&bar_w, 0);
        const uint64_t db_base = azt_desc(azt_smem(s_w));
        if (w == 1 && nslabs > 0) asm volatile("bar.arrive 3, 64;" ::: "memory");   // slab 0 has the first turn
        for (int j = w, y = w % n; j < nslabs; j += 2, y = (y + 2) % n) {
            const int st = j % AZT_STAGES;
            // input slab j feeds output slabs j+1 (dy 0), j (dy 1), j-1 (dy 2) of the same group
            const int dy0 = y + 1 < n ? 0 : 1, dy1 = y > 0 ? 2 : 1;
            const int top = j + 1 - dy0;
            // highest output slab entered by the slabs before this one
            const int entered = j == 0 ? -1 : (y == 0 ? j - 1 : j);
            azt_mbar_wait(&bar_in_full[st], (j / AZT_STAGES) & 1);
            // a block entered for a new output slab t: its previous tenant t-8 must have been retired
            for (int t = entered + 1; t <= top; t++)
                if (t >= 8) azt_mbar_wait(&bar_blk_free[AZT_RING(t)], ((t >> 3) - 1) & 1);
            const int blk = AZT_RING(top), nb = dy1 - dy0 + 1;
            const int first = nb < 8 - blk ? nb : 8 - blk, second = nb - first;     // split where the ring wraps
            const uint32_t d0 = tmem + blk * 64, i0 = AZT_IDESC(first), i1 = AZT_IDESC(second);
            // descriptors advance in 16-byte units: one row = 8, one K step (32 B) = 2
            const uint64_t da0 = azt_desc(azt_smem(s_in + st * AZT_CHUNK_BYTES) + 7 * AZT_ROW);    // row l-1 of the slab
            const uint64_t db0 = db_base + (uint64_t)(dy0 * 64 * (AZT_ROW / 16));
            const uint64_t db1 = db0 + (uint64_t)(first * 64 * (AZT_ROW / 16));
            // take the turn: the other warp has issued slab j-1
            if (w == 0) asm volatile("bar.sync 3, 64;" ::: "memory");
            else asm volatile("bar.sync 4, 64;" ::: "memory");
            asm volatile("tcgen05.fence::after_thread_sync;");
            if (azt_elect()) {
                if (!(p.debug & 8)) {
                    if (second == 0) {
#pragma unroll
                        for (int dx = 0; dx < 3; dx++)
#pragma unroll
                            for (int k = 0; k < 4; k++)
                                azt_mma(d0, da0 + (dx * 8 + k * 2), db0 + (dx * 192 * 8 + k * 2), i0);
                    } else {
                        // one accumulator range after the other: alternating between two
                        // costs ~33 cycles per switch (tools/probe/umma_gap.cu)
#pragma unroll
                        for (int dx = 0; dx < 3; dx++)
#pragma unroll
                            for (int k = 0; k < 4; k++)
                                azt_mma(d0, da0 + (dx * 8 + k * 2), db0 + (dx * 192 * 8 + k * 2), i0);
#pragma unroll
                        for (int dx = 0; dx < 3; dx++)
#pragma unroll
                            for (int k = 0; k < 4; k++)
                                azt_mma(tmem, da0 + (dx * 8 + k * 2), db1 + (dx * 192 * 8 + k * 2), i1);
                    }
                }
                // one commit per slab: output slabs wait for it, and so does the input stage
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                             ::"r"(azt_smem(&bar_mma_done[j & 7])) : "memory");
            }
            __syncwarp();
            asm volatile("tcgen05.fence::before_thread_sync;");
            if (j + 1 < nslabs) {
                if (w == 0) asm volatile("bar.arrive 4, 64;" ::: "memory");
                else asm volatile("bar.arrive 3, 64;" ::: "memory");
            }
        }
    } else {
        // ----------------------------------------------------- epilogue --
        // group = warp >> 3 takes the output slabs j = group (mod 2); inside a group a warp
        // owns 32 channels (half) of the 32 rows its TMEM lane quadrant holds:
        // thread = TMEM lane = row l of the slab
        const int grp = warp >> 3, half = (warp >> 2) & 1, wq = warp & 3;
        const int l = wq * 32 + lane;
        const bool real = l < p.bpg * (n + 1) && (l % (n + 1)) != n;    // not a pad cell
        const uint32_t keep = real ? 0xffffffffu : 0u;
        const int sw = l & 7;                                       // == R & 7 (8 + 128 q + l)
        uint4 rnext[4];
        if (RESID && grp < nslabs) {
            azt_load_resid(p.resid + (size_t)(AZT_HALO + (q0 + grp) * 128 + l) * AZT_ROW, half, sw, rnext);
        }
        for (int j = grp, y = grp % n; j < nslabs; j += 2, y = (y + 2) % n) {
            const int sb = j % AZT_OUT_STAGES;
            uint4 *srow = reinterpret_cast<uint4 *>(s_out + sb * AZT_OUT_BYTES + l * AZT_ROW);
            // the residual of this thread's row and channels comes straight from global memory,
            // one slab of the group ahead: the loads for slab j + 2 fly while slab j is finished
            uint4 rv[4];
            if (RESID) {
#pragma unroll
                for (int c = 0; c < 4; c++) rv[c] = rnext[c];
                if (j + 2 < nslabs) {
                    azt_load_resid(p.resid + (size_t)(AZT_HALO + (q0 + j + 2) * 128 + l) * AZT_ROW, half, sw, rnext);
                }
            }
            // the staging slab must have been drained by the bulk store that used it last
            if (!(p.debug & 4)) azt_mbar_wait(&bar_out_empty[sb], ((j / AZT_OUT_STAGES) & 1) ^ 1);
            // output slab j is complete once MMA(j+1) retired (MMA(j) for the last board row)
            const int last = y + 1 < n ? j + 1 : j;
            azt_mbar_wait(&bar_mma_done[last & 7], (last >> 3) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;");
            const int blk = AZT_RING(j);
            const uint32_t ta = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)blk * 64u + (uint32_t)half * 32u;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                uint32_t acc[16];
                AZT_TMEM_LD16(acc, ta + h * 16);
                asm volatile("tcgen05.wait::ld.sync.aligned;");
                azt_tmem_zero16(ta + h * 16);       // retire: zero for the block's next output slab
                if (p.debug & 4) continue;
#pragma unroll
                for (int g = 0; g < 2; g++) {
                    const int c8 = half * 4 + h * 2 + g;            // 8-channel chunk of the row
                    const float4 b0 = *reinterpret_cast<const float4 *>(&s_bias[c8 * 8]);
                    const float4 b1 = *reinterpret_cast<const float4 *>(&s_bias[c8 * 8 + 4]);
                    float f[8];
                    f[0] = __uint_as_float(acc[g * 8 + 0]) + b0.x; f[1] = __uint_as_float(acc[g * 8 + 1]) + b0.y;
                    f[2] = __uint_as_float(acc[g * 8 + 2]) + b0.z; f[3] = __uint_as_float(acc[g * 8 + 3]) + b0.w;
                    f[4] = __uint_as_float(acc[g * 8 + 4]) + b1.x; f[5] = __uint_as_float(acc[g * 8 + 5]) + b1.y;
                    f[6] = __uint_as_float(acc[g * 8 + 6]) + b1.z; f[7] = __uint_as_float(acc[g * 8 + 7]) + b1.w;
                    if (RESID) {
                        const uint4 r = rv[h * 2 + g];
                        const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            f[2 * q] += __uint_as_float(rw[q] << 16);
                            f[2 * q + 1] += __uint_as_float(rw[q] & 0xffff0000u);
                        }
                    }
                    uint32_t ow[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        ow[q] = azt_relu_bf16x2(f[2 * q], f[2 * q + 1]) & keep;
                    }
                    srow[c8 ^ sw] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                }
            }
            // hand the ring block back
            asm volatile("tcgen05.wait::st.sync.aligned;");
            asm volatile("tcgen05.fence::before_thread_sync;");
            __syncwarp();
            if (lane == 0) azt_mbar_arrive(&bar_blk_free[blk]);
            if (p.debug & 4) continue;
            // staging slab complete: the storer sends it to global memory
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) azt_mbar_arrive(&bar_out_done[sb]);
        }
    }
#undef AZT_RING
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}
