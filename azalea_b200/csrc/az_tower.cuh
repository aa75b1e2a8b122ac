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
// N = 128 (tools/probe/umma_rate.cu: 68 / 68 / 96 / 128 cycles at N = 64 /
// 128 / 192 / 256), so 64 output channels alone leave half the tensor pipe
// idle -- that is the ceiling of the library kernel (~900 TFLOP/s).  Here
// the three taps of one kernel row (dx = -1, 0, +1) share ONE A operand
// (shift dy*(n+1) only) against their three weight blocks stacked into
// N = 192:
//     E_dx[m] = sum_dy A[m + dy*(n+1)] . W(dy,dx)        (12 MMAs per tile)
//     D[m]    = E_0[m] + E_-1[m-1] + E_+1[m+1]
// and the epilogue recombines neighbouring rows (warp shuffles; a small
// shared-memory exchange across the four quadrant warps).  Tiles advance by
// 126 rows so every tile's inner 126 rows have both neighbours.
//
// Data movement is bulk async copies only (cp.async.bulk, no tensor maps):
// the layout is stored PRE-SWIZZLED in global memory (16-byte chunk j of row
// R lives at chunk j ^ (R & 7)), so the rows a tile needs are one contiguous
// block that lands in shared memory exactly as UMMA wants it; the 126 output
// rows of a tile are staged in shared memory and leave as one bulk store; the
// residual rows are bulk-loaded into that same staging buffer.
//
// Roles (608 threads, one persistent CTA per SM, tiles round-robin over CTAs):
//   warps 0-15 epilogue, all on the same tile: warp (quarter, quadrant) owns
//              16 of the 64 output channels of 32 rows: one batch of
//              tcgen05.ld, release the TMEM slot, recombine -> +bias
//              (+residual) -> ReLU -> zero the pad rows -> bf16 -> staging
//   warp 16    one thread issues tcgen05.mma; accumulators in TMEM, two
//              slots of 192 columns
//   warp 17    one thread streams input chunks through a 4-stage ring
//   warp 18    one thread bulk-loads the residual rows into the staging tiles
// mbarriers: in_full/in_empty per ring stage, acc_full/acc_empty per TMEM
// slot, out_full (residual landed) / out_empty (store drained) per staging
// buffer.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define AZT_C 64                    // channels = one 128-byte swizzle row
#define AZT_ROW 128                 // bytes per row
#define AZT_WBYTES (9 * 64 * 128)   // one layer's weights: [dy][dx][c_out][c_in]
#define AZT_TSTRIDE 126             // rows a tile advances
#define AZT_SLOTS 2                 // TMEM accumulator slots (256 columns apart)
#define AZT_STAGES 4                // input ring
#define AZT_CHUNK_ROWS 176          // 7 (alignment) + 128 + 2 * (19 + 1) + 1
#define AZT_CHUNK_BYTES (AZT_CHUNK_ROWS * AZT_ROW)
#define AZT_OUT_BYTES (128 * AZT_ROW)
#define AZT_SMEM_BYTES (AZT_WBYTES + AZT_STAGES * AZT_CHUNK_BYTES + 2 * AZT_OUT_BYTES)
#define AZT_THREADS 608

struct azt_params {
    const uint8_t *x;       // input activations, padded pre-swizzled layout
    const uint8_t *w;       // [3 dy][192 = dx*64 + c_out][128 B] pre-swizzled weights
    const float *bias;      // [64]
    const uint8_t *resid;   // residual input (same layout) or NULL
    uint8_t *out;           // output activations
    int n;                  // board size
    int halo;               // rows of zero halo (multiple of 8, >= n + 2)
    int rpb;                // rows per board = (n+1)^2
    long long rows;         // boards * rpb: rows that carry data
    long long tiles;        // ceil(rows / 126)
    int debug;              // probe only: 2 = skip the output stores
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

// first global row of the input chunk of tile T: the lowest row any tap
// reads, rounded down to the 8-row swizzle period
__device__ __forceinline__ long long azt_chunk_row0(const azt_params &p, long long T)
{
    return (p.halo + T * AZT_TSTRIDE - 1 - (p.n + 1)) & ~7ll;
}

template <bool RESID>
__global__ void __launch_bounds__(AZT_THREADS, 1)
k_conv3x3(const azt_params p)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *s_w = smem;                                            // 72 KB
    uint8_t *s_in = smem + AZT_WBYTES;                              // ring of input chunks
    uint8_t *s_out = s_in + AZT_STAGES * AZT_CHUNK_BYTES;           // two staging tiles
    __shared__ uint64_t bar_w, bar_in_full[AZT_STAGES], bar_in_empty[AZT_STAGES];
    __shared__ uint64_t bar_acc_full[AZT_SLOTS], bar_acc_empty[AZT_SLOTS];
    __shared__ uint64_t bar_out_full[2], bar_out_empty[2];
    __shared__ uint32_t tmem_holder;
    __shared__ __align__(16) float s_bias[AZT_C];
    __shared__ uint8_t s_pad[512];              // 1 = pad row (by row within a board)
    __shared__ __align__(16) float s_edge[2][4][2][64];   // [tile parity][quadrant][first row's E+1 | last row's E-1][channel]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        azt_mbar_init(&bar_w, 1);
        for (int i = 0; i < AZT_STAGES; i++) { azt_mbar_init(&bar_in_full[i], 1); azt_mbar_init(&bar_in_empty[i], 1); }
        for (int i = 0; i < AZT_SLOTS; i++) { azt_mbar_init(&bar_acc_full[i], 1); azt_mbar_init(&bar_acc_empty[i], 512); }
        for (int i = 0; i < 2; i++) { azt_mbar_init(&bar_out_full[i], 1); azt_mbar_init(&bar_out_empty[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (tid < AZT_C) s_bias[tid] = p.bias[tid];
    for (int i = tid; i < p.rpb; i += blockDim.x) {
        const int r = i / (p.n + 1), c = i % (p.n + 1);
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
    const long long last_row = p.halo + p.rows;             // one past the last data row

    if (warp == 18) {
        // ----------------------------------------------- residual loader --
        if (RESID && lane == 0) {
            int it = 0;
            for (long long T = first; T < p.tiles; T += step, it++) {
                // the tile's residual rows go straight into its staging buffer
                const int sb = it & 1;
                azt_mbar_wait(&bar_out_empty[sb], ((it >> 1) & 1) ^ 1);
                const long long g0 = p.halo + T * AZT_TSTRIDE;
                const long long nrows = last_row - g0 < AZT_TSTRIDE ? last_row - g0 : AZT_TSTRIDE;
                azt_mbar_expect_tx(&bar_out_full[sb], (uint32_t)nrows * AZT_ROW);
                azt_bulk_g2s(s_out + sb * AZT_OUT_BYTES, p.resid + (size_t)g0 * AZT_ROW,
                             (uint32_t)nrows * AZT_ROW, &bar_out_full[sb]);
            }
        }
    } else if (warp == 17) {
        // -------------------------------------------------- input loader --
        if (lane == 0) {
            azt_mbar_expect_tx(&bar_w, AZT_WBYTES);
            for (int t = 0; t < 9; t++) azt_bulk_g2s(s_w + t * 8192, p.w + t * 8192, 8192, &bar_w);
            int it = 0;
            for (long long T = first; T < p.tiles; T += step, it++) {
                const int st = it % AZT_STAGES;
                azt_mbar_wait(&bar_in_empty[st], ((it / AZT_STAGES) & 1) ^ 1);
                // the buffer carries AZT_CHUNK_ROWS spare rows after its data rows
                const uint8_t *src = p.x + (size_t)azt_chunk_row0(p, T) * AZT_ROW;
                uint8_t *dst = s_in + st * AZT_CHUNK_BYTES;
                azt_mbar_expect_tx(&bar_in_full[st], AZT_CHUNK_BYTES);
                azt_bulk_g2s(dst, src, AZT_CHUNK_BYTES / 2, &bar_in_full[st]);
                azt_bulk_g2s(dst + AZT_CHUNK_BYTES / 2, src + AZT_CHUNK_BYTES / 2, AZT_CHUNK_BYTES / 2, &bar_in_full[st]);
            }
        }
    } else if (warp == 16) {
        // --------------------------------------------------- MMA issuer --
        if (lane == 0) {
            azt_mbar_wait(&bar_w, 0);
            const int rs = p.n + 1;
            const uint32_t b_base = azt_smem(s_w);
            int it = 0;
            for (long long T = first; T < p.tiles; T += step, it++) {
                const int st = it % AZT_STAGES, slot = it & 1;
                azt_mbar_wait(&bar_in_full[st], (it / AZT_STAGES) & 1);
                azt_mbar_wait(&bar_acc_empty[slot], ((it >> 1) & 1) ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;");
                // tile T covers global rows [halo + 126 T - 1, + 128)
                const long long r0 = azt_chunk_row0(p, T);
                const uint32_t a_base = azt_smem(s_in + st * AZT_CHUNK_BYTES) +
                                        (uint32_t)((p.halo + T * AZT_TSTRIDE - 1 - r0) * AZT_ROW);
                const uint32_t d = tmem + (uint32_t)slot * 256u;
#pragma unroll
                for (int dy = 0; dy < 3; dy++) {
                    const uint32_t a_dy = a_base + (uint32_t)((dy - 1) * rs * AZT_ROW);
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
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                             ::"r"(azt_smem(&bar_in_empty[st])) : "memory");
            }
        }
    } else {
        // ----------------------------------------------------- epilogue --
        // thread = TMEM lane = global row halo + 126 T - 1 + t; columns
        // [0,64) = E_-1, [64,128) = E_0, [128,192) = E_+1
        const int cq = warp >> 2, wq = warp & 3;        // channel quarter, TMEM lane quadrant
        const int barid = 1 + cq;                       // named barrier of the 4 quadrant warps of a quarter
        const int t = wq * 32 + lane;
        const float mask_up = lane == 0 ? 0.f : 1.f, mask_dn = lane == 31 ? 0.f : 1.f;
        // row within its board, advanced by (126 * gridDim) mod rpb per tile
        float bias[16];
#pragma unroll
        for (int q = 0; q < 16; q++) bias[q] = s_bias[cq * 16 + q];
        int rmod = (int)((first * AZT_TSTRIDE + t - 1 + p.rpb) % p.rpb);
        const int rstep = (int)((step * AZT_TSTRIDE) % p.rpb);
        int it = 0;
        for (long long T = first; T < p.tiles; T += step, it++) {
            const int slot = it & 1, sb = it & 1;
            const long long grow = p.halo + T * AZT_TSTRIDE - 1 + t;        // this thread's global row
            const bool valid = t >= 1 && t <= AZT_TSTRIDE && grow < last_row;
            const uint32_t keep = (valid && s_pad[rmod] == 0) ? 0xffffffffu : 0u;
            rmod += rstep;
            if (rmod >= p.rpb) rmod -= p.rpb;
            const int sw = (int)(grow & 7);
            uint4 *srow = reinterpret_cast<uint4 *>(s_out + sb * AZT_OUT_BYTES + (t - 1) * AZT_ROW);
            azt_mbar_wait(&bar_acc_full[slot], (it >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;");
            // this warp's 16 channels of the three blocks: columns 16 cq of E_-1 | E_0 | E_+1
            const uint32_t ta = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)slot * 256u + (uint32_t)cq * 16u;
            uint32_t em[16], e0[16], ep[16];
            AZT_TMEM_LD16(em, ta);
            AZT_TMEM_LD16(e0, ta + 64);
            AZT_TMEM_LD16(ep, ta + 128);
            asm volatile("tcgen05.wait::ld.sync.aligned;");
            // the accumulators are in registers: hand the TMEM slot back
            asm volatile("tcgen05.fence::before_thread_sync;");
            azt_mbar_arrive(&bar_acc_empty[slot]);
            // Rows t-1 / t+1 live in the neighbouring lanes -- except across the
            // four quadrant warps: lane 31 publishes its E_-1 row and lane 0 its
            // E_+1 row through shared memory (double-buffered by tile parity).
            float *edge = &s_edge[it & 1][0][0][0];          // [quadrant][2][64]
            if (lane == 0) {
                uint4 *dst = reinterpret_cast<uint4 *>(edge + (wq * 2 + 0) * 64 + cq * 16);
#pragma unroll
                for (int q = 0; q < 4; q++) dst[q] = make_uint4(ep[4 * q], ep[4 * q + 1], ep[4 * q + 2], ep[4 * q + 3]);
            } else if (lane == 31) {
                uint4 *dst = reinterpret_cast<uint4 *>(edge + (wq * 2 + 1) * 64 + cq * 16);
#pragma unroll
                for (int q = 0; q < 4; q++) dst[q] = make_uint4(em[4 * q], em[4 * q + 1], em[4 * q + 2], em[4 * q + 3]);
            }
            asm volatile("bar.sync %0, 128;" ::"r"(barid) : "memory");
            float f[16];
#pragma unroll
            for (int q = 0; q < 16; q++) {
                const float up = __shfl_up_sync(0xffffffffu, __uint_as_float(em[q]), 1);    // E_-1[t-1]
                const float dn = __shfl_down_sync(0xffffffffu, __uint_as_float(ep[q]), 1);  // E_+1[t+1]
                // lanes 0 / 31 got their own value back: masked out here, patched below
                f[q] = fmaf(up, mask_up, fmaf(dn, mask_dn, __uint_as_float(e0[q]) + bias[q]));
            }
            if (lane == 0 || lane == 31) {
                // the neighbour row of the warp's first / last lane is in another warp
                const float *ex = edge + (lane == 0 ? ((wq > 0 ? wq - 1 : 0) * 2 + 1) : ((wq < 3 ? wq + 1 : 3) * 2)) * 64 + cq * 16;
#pragma unroll
                for (int q = 0; q < 16; q += 4) {
                    const float4 xq = *reinterpret_cast<const float4 *>(ex + q);
                    f[q] += xq.x; f[q + 1] += xq.y; f[q + 2] += xq.z; f[q + 3] += xq.w;
                }
            }
            // the staging buffer holds the residual rows (RESID) or must have
            // been drained by the bulk store of two tiles ago
            if (RESID) azt_mbar_wait(&bar_out_full[sb], (it >> 1) & 1);
            else azt_mbar_wait(&bar_out_empty[sb], ((it >> 1) & 1) ^ 1);
            if (valid) {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    if (RESID) {
                        const uint4 r = srow[(cq * 2 + h) ^ sw];
                        const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
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
                    srow[(cq * 2 + h) ^ sw] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                }
            }
            // staging tile complete: one thread sends it to global memory
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("bar.sync 5, 512;" ::: "memory");
            if (tid == 0) {
                const long long g0 = p.halo + T * AZT_TSTRIDE;
                const long long nrows = last_row - g0 < AZT_TSTRIDE ? last_row - g0 : AZT_TSTRIDE;
                if (!(p.debug & 2))
                    azt_bulk_s2g(p.out + (size_t)g0 * AZT_ROW, s_out + sb * AZT_OUT_BYTES, (uint32_t)nrows * AZT_ROW);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                // the previous tile's store has finished reading its buffer
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                if (it > 0) azt_mbar_arrive(&bar_out_empty[sb ^ 1]);
            }
        }
        if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}
