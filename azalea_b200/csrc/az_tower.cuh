// az_tower.cuh -- 3x3 convolution (64 -> 64 channels) + bias (+ residual) +
// ReLU for the evaluator's tower on tcgen05 tensor cores (sm_100a).
//
// network.py:17-39 (Resblock) spends 99 % of the evaluator's FLOPs in twelve
// of these.  The kernel is an implicit GEMM with NO im2col and no tensor
// maps, built on two observations.
//
// (1) Pointer-shift convolution.  Activations are stored as a tall "image"
// of 128-byte rows (64 bf16 channels) in which every board row carries one
// zero pad cell and every board one zero pad row,
//
//     row(b, r, c) = HALO + b * (n+1)^2 + r * (n+1) + c,   pad: c == n or r == n
//
// so the input of tap (dy, dx) for output row p is row p + dy*(n+1) + dx for
// EVERY row of a 128-row MMA tile at once: the A operand of a tap is the
// same shared-memory buffer with the UMMA descriptor's start address moved
// by some rows.  The hardware applies the 128-byte swizzle on absolute
// address bits, so any row offset works (tools/probe/umma_probe.cu).  The
// zero pads are the convolution's zero padding and the halo between boards.
// Cost: (n+1)^2 / n^2 = 19 % more rows at 11x11.
//
// (2) N = 192.  A 128xNx16 UTCHMMA takes ~68 cycles for N = 64 AND for
// N = 128 (tools/probe/umma_rate.cu: 68 / 68 / 128 cycles at N = 64 / 128 /
// 256), so 64 output channels alone leave half the tensor pipe idle -- that
// is the ceiling of the library kernel (~900 TFLOP/s).  Here the three taps
// of one kernel row (dx = -1, 0, +1) share ONE A operand (shift dy*(n+1)
// only) against their three weight blocks stacked into N = 192:
//     E_dx[m] = sum_dy A[m + dy*(n+1)] . W(dy,dx)        (12 MMAs per tile)
//     D[m]    = E_0[m] + E_-1[m-1] + E_+1[m+1]
// and the epilogue recombines neighbouring rows (warp shuffles; a 2 KB
// shared-memory exchange across the four epilogue warps).  Tiles advance by
// 126 rows so every tile's inner 126 rows have both neighbours.
//
// Data movement: the layout is stored PRE-SWIZZLED in global memory (16-byte
// chunk j of row R lives at chunk j ^ (R & 7)), so a group of boards plus
// halo is one contiguous block that plain bulk async copies (cp.async.bulk,
// no tensor map) drop into shared memory exactly as UMMA wants it.
//
// Roles (576 threads, one persistent CTA per SM):
//   warps 0-15 epilogue.  Two groups of 8 warps take alternate tiles (TMEM
//              slot = group); inside a group warp (half, quadrant) owns 32 of
//              the 64 output channels of 32 rows: tcgen05.ld -> recombine ->
//              +bias (+residual) -> ReLU -> zero the pad rows -> bf16 ->
//              global (pre-swizzled).  The epilogue is a long dependent
//              chain per warp, so it needs the thread-level parallelism.
//   warp 16    one thread issues tcgen05.mma; accumulators in TMEM, two
//              slots of 192 columns
//   warp 17    one thread streams board groups into a 2-stage smem ring
// mbarriers: in_full/in_empty per stage, acc_full/acc_empty per TMEM slot.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define AZT_C 64                    // channels = one 128-byte swizzle row
#define AZT_ROW 128                 // bytes per row
#define AZT_WBYTES (9 * 64 * 128)   // one layer's weights: [dy][dx][c_out][c_in]
#define AZT_TSTRIDE 126             // rows a tile advances
#define AZT_SLOTS 2                 // TMEM accumulator slots (256 columns apart)

struct azt_params {
    const uint8_t *x;       // input activations, padded pre-swizzled layout
    const uint8_t *w;       // [3 dy][192 = dx*64 + c_out][128 B] pre-swizzled weights
    const float *bias;      // [64]
    const uint8_t *resid;   // residual input (same layout) or NULL
    uint8_t *out;           // output activations
    int n;                  // board size
    int halo;               // rows of zero halo (multiple of 8, >= n + 2)
    int rpb;                // rows per board = (n+1)^2
    int rows_group;         // boards per group * rpb (multiple of 8)
    int tiles;              // ceil(rows_group / 126)
    int stage_rows;         // halo + rows_group
    long long groups;       // number of board groups
    int debug;              // probe only: 1 skip input copies, 2 skip output stores, 4 skip epilogue math
};

__device__ __forceinline__ uint32_t azt_smem(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

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

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start >> 4 | LBO 1 << 16 | SBO (1024 B >> 4) << 32 | version 1 << 46 | layout 2 << 61
__device__ __forceinline__ uint64_t azt_desc(uint32_t saddr)
{
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

// cute::UMMA::InstrDescriptor: D f32 (1 << 4), A bf16 (1 << 7), B bf16 (1 << 10),
// both K-major, N >> 3 at bit 17, M >> 4 at bit 24;  M = 128, N = 192
#define AZT_IDESC ((1u << 4) | (1u << 7) | (1u << 10) | ((192u >> 3) << 17) | ((128u >> 4) << 24))

#define AZT_TMEM_LD16(v, addr)                                                                     \
    asm volatile(                                                                                  \
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                  \
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"                          \
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),      \
          "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),  \
          "=r"(v[14]), "=r"(v[15])                                                                 \
        : "r"(addr))

template <bool RESID>
__global__ void __launch_bounds__(576, 1)
k_conv3x3(const azt_params p)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    // [stage 0][stage 1][weights]: the last tile of a group reads some rows
    // past its stage (rows that are never written back); keeping the weights
    // last keeps those reads inside the allocation
    const uint32_t stage_bytes = (uint32_t)p.stage_rows * AZT_ROW;
    uint8_t *s_in[2] = {smem, smem + stage_bytes};
    uint8_t *s_w = smem + 2 * stage_bytes;
    __shared__ uint64_t bar_w, bar_in_full[2], bar_in_empty[2];
    __shared__ uint64_t bar_acc_full[AZT_SLOTS], bar_acc_empty[AZT_SLOTS];
    __shared__ uint32_t tmem_holder;
    __shared__ __align__(16) float s_bias[AZT_C];
    __shared__ uint8_t s_pad[1024];             // 1 = pad row (by row within a group)
    __shared__ __align__(16) float s_edge[2][4][2][64];       // [group][quadrant][first row's E+1 | last row's E-1][channel]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        azt_mbar_init(&bar_w, 1);
        for (int i = 0; i < 2; i++) { azt_mbar_init(&bar_in_full[i], 1); azt_mbar_init(&bar_in_empty[i], 1); }
        for (int i = 0; i < AZT_SLOTS; i++) { azt_mbar_init(&bar_acc_full[i], 1); azt_mbar_init(&bar_acc_empty[i], 256); }
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (tid < AZT_C) s_bias[tid] = p.bias[tid];
    for (int i = tid; i < p.rows_group; i += blockDim.x) {
        const int ib = i % p.rpb, r = ib / (p.n + 1), c = ib % (p.n + 1);
        s_pad[i] = (r == p.n || c == p.n) ? 1 : 0;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(azt_smem(&tmem_holder)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_holder;
    const long long first = blockIdx.x, step = gridDim.x;

    if (warp == 17) {
        // ------------------------------------------------------- loader --
        if (lane == 0) {
            azt_mbar_expect_tx(&bar_w, AZT_WBYTES);
            for (int t = 0; t < 9; t++) azt_bulk_g2s(s_w + t * 8192, p.w + t * 8192, 8192, &bar_w);
            int it = 0;
            for (long long g = first; g < p.groups; g += step, it++) {
                const int st = it & 1;
                azt_mbar_wait(&bar_in_empty[st], ((it >> 1) & 1) ^ 1);
                if ((p.debug & 1) && it >= 2) { azt_mbar_arrive(&bar_in_full[st]); continue; }
                azt_mbar_expect_tx(&bar_in_full[st], stage_bytes);
                // the group's rows plus a halo on both sides: one contiguous block
                const uint8_t *src = p.x + (size_t)g * p.rows_group * AZT_ROW;
                uint32_t off = 0;
                while (off < stage_bytes) {
                    uint32_t nbytes = stage_bytes - off < 16384u ? stage_bytes - off : 16384u;
                    azt_bulk_g2s(s_in[st] + off, src + off, nbytes, &bar_in_full[st]);
                    off += nbytes;
                }
            }
        }
    } else if (warp == 16) {
        // --------------------------------------------------- MMA issuer --
        if (lane == 0) {
            azt_mbar_wait(&bar_w, 0);
            const int rs = p.n + 1;
            const uint32_t b_base = azt_smem(s_w);
            int it = 0, tcount = 0;
            for (long long g = first; g < p.groups; g += step, it++) {
                const int st = it & 1;
                azt_mbar_wait(&bar_in_full[st], (it >> 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;");
                // tile j covers group rows [126 j - 1, 126 j + 127)
                const uint32_t a_base = azt_smem(s_in[st]) + (uint32_t)(p.halo - 1) * AZT_ROW;
                for (int j = 0; j < p.tiles; j++, tcount++) {
                    const int slot = tcount & 1;
                    azt_mbar_wait(&bar_acc_empty[slot], ((tcount >> 1) & 1) ^ 1);
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    const uint32_t d = tmem + (uint32_t)slot * 256u;
#pragma unroll
                    for (int dy = 0; dy < 3; dy++) {
                        const uint32_t a_dy = a_base + (uint32_t)((j * AZT_TSTRIDE + (dy - 1) * rs) * AZT_ROW);
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const uint64_t da = azt_desc(a_dy + k * 32);
                            const uint64_t db = azt_desc(b_base + dy * (192 * AZT_ROW) + k * 32);
                            const uint32_t acc = (dy | k) != 0;
                            asm volatile(
                                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                                ::"r"(d), "l"(da), "l"(db), "r"(AZT_IDESC), "r"(acc) : "memory");
                        }
                    }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                                 ::"r"(azt_smem(&bar_acc_full[slot])) : "memory");
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                             ::"r"(azt_smem(&bar_in_empty[st])) : "memory");
            }
        }
    } else {
        // ----------------------------------------------------- epilogue --
        // thread = TMEM lane = row (126 j - 1 + t) of the group; columns
        // [0,64) = E_-1, [64,128) = E_0, [128,192) = E_+1
        const int wg = warp >> 3, half = (warp >> 2) & 1, wq = warp & 3;   // group, channel half, TMEM lane quadrant
        const int barid = 1 + wg * 2 + half;            // named barrier of the 4 warps sharing (group, half)
        int tcount = 0;
        const int t = wq * 32 + lane;
        for (long long g = first; g < p.groups; g += step) {
            for (int j = 0; j < p.tiles; j++, tcount++) {
                if ((tcount & 1) != wg) continue;
                const int slot = wg;
                azt_mbar_wait(&bar_acc_full[slot], (tcount >> 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;");
                // this warp's 32 channels: columns [32 half, 32 half + 32) of each E block
                const uint32_t ta = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)slot * 256u + (uint32_t)half * 32u;
                const int row = j * AZT_TSTRIDE - 1 + t;        // row within the group
                const bool valid = t >= 1 && t <= AZT_TSTRIDE && row < p.rows_group;
                const size_t grow = (size_t)p.halo + (size_t)g * p.rows_group + (valid ? row : 0);
                const uint32_t keep = (valid && s_pad[row] == 0) ? 0xffffffffu : 0u;
                const int sw = (int)(grow & 7);
                uint4 *orow = reinterpret_cast<uint4 *>(p.out + grow * AZT_ROW);
                const uint4 *rrow = reinterpret_cast<const uint4 *>(p.resid + grow * AZT_ROW);
                // TMEM reads are slow while the tensor core works on the other
                // slot, and the slot cannot be reused until they are done: fetch
                // the warp's 32 channels of all three blocks up front (E_-1 / E_+1
                // squeezed to fp16 pairs -- their rounding, 2^-12, is far below
                // the bf16 rounding of the output), release the slot, then do the
                // arithmetic and the stores while the next tile's MMAs run.
                if (p.debug & 4) {
                    asm volatile("tcgen05.fence::before_thread_sync;");
                    azt_mbar_arrive(&bar_acc_empty[slot]);
                    continue;
                }
                // Rows t-1 / t+1 live in the neighbouring lanes -- except across
                // the four quadrant warps.  Once per tile lane 31 publishes its
                // E_-1 row and lane 0 its E_+1 row through shared memory.
                float *edge = &s_edge[wg][0][0][0];              // [quadrant][2][64]
#pragma unroll
                for (int cb = 0; cb < 2; cb++) {
                    uint32_t em[16], ep[16];
                    AZT_TMEM_LD16(em, ta + cb * 16);
                    AZT_TMEM_LD16(ep, ta + 128 + cb * 16);
                    asm volatile("tcgen05.wait::ld.sync.aligned;");
                    if (lane == 0 || lane == 31) {
                        float *dst = edge + (wq * 2 + (lane == 0 ? 0 : 1)) * 64 + half * 32 + cb * 16;
#pragma unroll
                        for (int q = 0; q < 16; q += 4) {
                            const uint32_t s0 = lane == 0 ? ep[q] : em[q], s1 = lane == 0 ? ep[q + 1] : em[q + 1];
                            const uint32_t s2 = lane == 0 ? ep[q + 2] : em[q + 2], s3 = lane == 0 ? ep[q + 3] : em[q + 3];
                            *reinterpret_cast<float4 *>(dst + q) =
                                make_float4(__uint_as_float(s0), __uint_as_float(s1), __uint_as_float(s2), __uint_as_float(s3));
                        }
                    }
                }
                asm volatile("bar.sync %0, 128;" ::"r"(barid) : "memory");
                const float mask_up = lane == 0 ? 0.f : 1.f, mask_dn = lane == 31 ? 0.f : 1.f;
                const float *eu = edge + ((wq > 0 ? wq - 1 : 0) * 2 + 1) * 64 + half * 32;    // E_-1 of row 32 wq - 1
                const float *ed = edge + ((wq < 3 ? wq + 1 : 3) * 2 + 0) * 64 + half * 32;    // E_+1 of row 32 wq + 32
#pragma unroll 1
                for (int cb = 0; cb < 2; cb++) {                 // 16 output channels at a time
                    uint4 res[2];
                    if (RESID && valid) {                        // issue early: overlaps the TMEM loads
                        res[0] = rrow[(half * 4 + cb * 2) ^ sw];
                        res[1] = rrow[(half * 4 + cb * 2 + 1) ^ sw];
                    }
                    uint32_t em[16], e0[16], ep[16];
                    AZT_TMEM_LD16(em, ta + cb * 16);
                    AZT_TMEM_LD16(e0, ta + 64 + cb * 16);
                    AZT_TMEM_LD16(ep, ta + 128 + cb * 16);
                    asm volatile("tcgen05.wait::ld.sync.aligned;");
                    float f[16];
#pragma unroll
                    for (int q = 0; q < 16; q += 4) {
                        const float4 bq = *reinterpret_cast<const float4 *>(s_bias + half * 32 + cb * 16 + q);
                        const float bb[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
                        for (int r = 0; r < 4; r++) {
                            const float up = __shfl_up_sync(0xffffffffu, __uint_as_float(em[q + r]), 1);    // E_-1[t-1]
                            const float dn = __shfl_down_sync(0xffffffffu, __uint_as_float(ep[q + r]), 1);  // E_+1[t+1]
                            // lanes 0 / 31 got their own value back: masked out here, patched below
                            f[q + r] = fmaf(up, mask_up, fmaf(dn, mask_dn, __uint_as_float(e0[q + r]) + bb[r]));
                        }
                    }
                    if (lane == 0 || lane == 31) {
                        // the neighbour row of the warp's first / last lane is in another warp
                        const float *ex = (lane == 0 ? eu : ed) + cb * 16;
#pragma unroll
                        for (int q = 0; q < 16; q += 4) {
                            const float4 xq = *reinterpret_cast<const float4 *>(ex + q);
                            f[q] += xq.x; f[q + 1] += xq.y; f[q + 2] += xq.z; f[q + 3] += xq.w;
                        }
                    }
                    if (valid && !(p.debug & 2)) {
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            if (RESID) {
                                const uint32_t rw[4] = {res[h].x, res[h].y, res[h].z, res[h].w};
#pragma unroll
                                for (int q = 0; q < 4; q++) {
                                    f[h * 8 + 2 * q] += __uint_as_float(rw[q] << 16);
                                    f[h * 8 + 2 * q + 1] += __uint_as_float(rw[q] & 0xffff0000u);
                                }
                            }
                            uint32_t ow[4];
#pragma unroll
                            for (int q = 0; q < 4; q++) {
                                __nv_bfloat162 hh = __floats2bfloat162_rn(fmaxf(f[h * 8 + 2 * q], 0.f),
                                                                          fmaxf(f[h * 8 + 2 * q + 1], 0.f));
                                ow[q] = *reinterpret_cast<uint32_t *>(&hh) & keep;
                            }
                            orow[(half * 4 + cb * 2 + h) ^ sw] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                        }
                    }
                }
                // all four quadrant warps are done with this tile's edge rows
                asm volatile("bar.sync %0, 128;" ::"r"(barid) : "memory");
                // the accumulator slot has been read: hand TMEM back
                asm volatile("tcgen05.fence::before_thread_sync;");
                azt_mbar_arrive(&bar_acc_empty[slot]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}
