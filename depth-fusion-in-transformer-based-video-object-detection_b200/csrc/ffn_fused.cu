// Fused feed-forward block of the encoder / decoder layers on the 5th-generation tensor cores (sm_100a):
//
//     y     = LayerNorm( x + linear2( relu( linear1(x) ) ) ) * gamma + beta        d_model = 256, bf16
//     y_pos = y + pos                                                             (optional second output)
//
// i.e. forward_ffn + norm2 of DeformableTransformerEncoderLayer (reference
// models/deformable_transformer_single.py:544-548) and the query of the next layer (:530-531, :538), in ONE
// kernel: the [rows, d_ffn] hidden activation (364 MB at batch 8) never leaves the SM.
//
// One CTA owns a 128-row tile (persistent over tiles).  Per 64-wide chunk c of the hidden dimension:
//     GEMM1(c):  Hacc[c&1] (TMEM, 128 x 64 fp32)  = X (smem, 128 x 256)  @ W1[c]^T (smem, 64 x 256)
//     epilogue1: Hacc -> + b1 -> ReLU -> bf16 -> sH[c&1] (smem, K-major SWIZZLE_128B: the A operand of GEMM2)
//     GEMM2(c):  Yacc (TMEM, 128 x 256 fp32)     += sH[c&1] (128 x 64)  @ W2[:, c]^T (smem, 256 x 64)
// Warp roles (14 warps):
//   warps 0-7   epilogue warps (thread = half a row).  Per chunk: Hacc -> +b1 -> ReLU -> bf16 -> sH.  Per tile:
//               pass 1: Yacc + b2 + x -> bf16 into the staging tile + row sums; then straight on to the next tile.
//   warp 8      MMA issuer.  Warp-uniform control flow, tcgen05 instructions issued by one elected lane.  The X
//               tile is copied once from shared memory into tensor memory (tcgen05.cp) and GEMM1 takes its A
//               operand from there: no 64 KB re-read of X per chunk (GEMM1 was shared-memory-bandwidth bound),
//               and the shared-memory X buffer is free for the whole main loop.  Issue order GEMM1(c+1),
//               GEMM2(c): the tensor pipe has work queued while the epilogue of chunk c runs.
//   warp 9      TMA producer: X / W1 / W2 boxes (SWIZZLE_128B boxes land directly in the UMMA operand layout,
//               completion by mbarrier transaction bytes); every buffer is refilled as soon as its last reader
//               (a committed MMA group) has finished.  X lands in the (then idle) W2 buffers.
//   warps 10-13 store warps: normalise the staged tile and write y (+ pos) with coalesced 16-byte stores while
//               the other warps are already in the next tile's main loop (the SM's store egress, 32 B/clk, makes
//               this the longest phase of a tile: it must not be on the critical path).
// TMEM (512 columns): Yacc 256 | Hacc 2 x 64 | X tile 128 (bf16 pairs).
// Shared memory: staging 64 KB + W1 2 x 32 KB + W2 2 x 32 KB (X landing zone) + H 2 x 16 KB + row statistics.
#include <cuda.h>

#include "msda_common.cuh"
#include "msda_launch.h"
#include "umma.cuh"

namespace msda {

using namespace umma;

constexpr int kFfnC = 256;            // d_model
constexpr int kFfnTM = 128;           // rows per tile
constexpr int kFfnCH = 64;            // hidden columns per chunk (one 128-byte k-block of GEMM2)
constexpr int kFfnSplit = 2;          // epilogue warps per 32-row group: each takes 1/SPLIT of the columns
constexpr int kFfnEpiWarps = 4 * kFfnSplit;
constexpr int kFfnEpiThreads = 32 * kFfnEpiWarps;
constexpr int kFfnStoreWarps = 4;
constexpr int kFfnStoreThreads = 32 * kFfnStoreWarps;
constexpr int kFfnThreads = kFfnEpiThreads + 64 + kFfnStoreThreads;
constexpr int kSmemX = 4 * kFfnTM * 128;          // 65536
constexpr int kSmemW1 = 4 * kFfnCH * 128;         // 32768 per buffer
constexpr int kSmemW2 = kFfnC * 128;              // 32768 per buffer
constexpr int kSmemH = kFfnTM * 128;              // 16384 per buffer
constexpr int kSmemStat = 2 * kFfnSplit * kFfnTM * 4;   // partial row statistics (sum, sum of squares)
constexpr int kSmemBars = 256;
constexpr int kFfnSmem = kSmemX + 2 * kSmemW1 + 2 * kSmemW2 + 2 * kSmemH + kSmemStat + kSmemBars;
// CTAs of a cluster that work on neighbouring row tiles in lockstep and share every weight chunk: each loads 1 / C of a
// W1 / W2 chunk and multicasts it into all of them (TMA multicast).  The kernel streams 1 MB of weights per 128-row tile
// out of L2 -- 1.64 GB per call at the encoder's 8 x 22 223 rows, 9 TB/s (profiles/ncu_r1e_ffn.json: tensor pipe 52-57 %
// active) -- and a cluster of two halves that.  MEASURED (round 2, same box, alternating runs, tools/run_ffn.py):
// 0.1896 ms with clusters of two against 0.1815 ms without: the L2 -> SM traffic is not what holds this kernel, and the
// lockstep of two CTAs costs 4 %.  Default 1 (no clusters); 2 stays as a build option (parity-tested).
#ifndef MSDA_FFN_CLUSTER
#define MSDA_FFN_CLUSTER 1
#endif
constexpr int kFfnCluster = MSDA_FFN_CLUSTER;
static_assert(kFfnCluster == 1 || kFfnCluster == 2, "W1 chunks are split by K block pairs, W2 chunks by row halves");
static_assert(kFfnSmem <= 232448, "exceeds the 227 KB of shared memory a CTA can opt in to");

#ifdef MSDA_FFN_TRACE
__device__ long long g_ffn_trace[2][40][8];
#define FFN_TRACE(role, c, k) do { if (blockIdx.x == 0 && it == 1) g_ffn_trace[role][c][k] = clock64(); } while (0)
#else
#define FFN_TRACE(role, c, k) do { } while (0)
#endif

struct FfnBars {
    unsigned long long x, xcopied, xy_free, stage_full, stage_free, w1[2], w2[2], g1[2], g2[2], hfull[2];
    // cluster-wide versions of g1 / g2 / xcopied (one arrival per CTA, multicast tcgen05.commit): a weight buffer may be
    // refilled -- the refill lands in EVERY CTA of the cluster -- once its readers in every CTA are done
    unsigned long long f1[2], f2[2], xc_all;
    unsigned tmem_base;
};

// phase parity of the (c >> 1)-th use, in tile iteration `it`, of a per-buffer barrier that is used once
// per chunk of parity b = c & 1 (nb = number of such chunks per tile)
__device__ __forceinline__ unsigned chunk_parity(unsigned it, int c, int NC)
{
    const int b = c & 1;
    const unsigned nb = (unsigned)((NC + 1 - b) >> 1);
    return (it * nb + (unsigned)(c >> 1)) & 1u;
}

__global__ void __cluster_dims__(kFfnCluster, 1, 1) __launch_bounds__(kFfnThreads, 1)
ffn_layernorm_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w1,
                     const __grid_constant__ CUtensorMap tm_w2,
                     const __nv_bfloat16* __restrict__ b1, const __nv_bfloat16* __restrict__ b2,
                     const __nv_bfloat16* __restrict__ gamma, const __nv_bfloat16* __restrict__ beta,
                     const __nv_bfloat16* __restrict__ pos, __nv_bfloat16* __restrict__ y,
                     __nv_bfloat16* __restrict__ y_pos, long long rows, int F, float eps)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sStage = smem;                  // pre-norm tile of the previous tile, read by the store warps
    unsigned char* sW1 = sStage + kSmemX;          // [2][kSmemW1]
    unsigned char* sW2 = sW1 + 2 * kSmemW1;        // [2][kSmemW2]; at a tile boundary: landing zone of the X tile
    unsigned char* sH = sW2 + 2 * kSmemW2;         // [2][kSmemH]
    float* s_stat = reinterpret_cast<float*>(sH + 2 * kSmemH);
    FfnBars* bars = reinterpret_cast<FfnBars*>(sH + 2 * kSmemH + kSmemStat);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NC = F / kFfnCH;
    // work unit of a cluster: kFfnCluster consecutive row tiles; CTA `rank` takes tile unit * C + rank (past the last
    // tile: TMA reads zeros, nothing is stored)
    const long long units = ((rows + kFfnTM - 1) / kFfnTM + kFfnCluster - 1) / kFfnCluster;
    const unsigned rank = kFfnCluster > 1 ? cluster_ctarank() : 0u;
    const long long first = blockIdx.x / kFfnCluster, stride = gridDim.x / kFfnCluster;
    constexpr unsigned short kAll = (unsigned short)((1u << kFfnCluster) - 1u);

    if (warp == 0) tmem_alloc(&bars->tmem_base, 512);
    if (tid == kFfnEpiThreads) {
        mbar_init(&bars->x, 1);
        mbar_init(&bars->xcopied, 1);
        mbar_init(&bars->xy_free, kFfnEpiThreads);
        mbar_init(&bars->stage_full, kFfnEpiThreads);
        mbar_init(&bars->stage_free, kFfnStoreThreads);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars->w1[i], 1); mbar_init(&bars->w2[i], 1);
            mbar_init(&bars->g1[i], 1); mbar_init(&bars->g2[i], 1);
            mbar_init(&bars->hfull[i], kFfnEpiThreads);
            mbar_init(&bars->f1[i], kFfnCluster); mbar_init(&bars->f2[i], kFfnCluster);
        }
        mbar_init(&bars->xc_all, kFfnCluster);
        fence_mbar_init();
        tma_prefetch_desc(&tm_x); tma_prefetch_desc(&tm_w1); tma_prefetch_desc(&tm_w2);
    }
    tcgen05_fence_before();
    __syncthreads();
    if (kFfnCluster > 1) cluster_sync_all();       // every CTA's barriers exist before anything arrives on them
    tcgen05_fence_after();
    const unsigned tmem = bars->tmem_base;
    const unsigned tmem_y = tmem;                  // columns [0, 256)
    const unsigned tmem_h = tmem + 256;            // columns [256, 384): two 64-column buffers
    const unsigned tmem_x = tmem + 384;            // columns [384, 512): the X tile, two bf16 per column

    if (warp == kFfnEpiWarps) {
        // ======================================= MMA issuer =======================================
        const unsigned idesc1 = make_idesc_bf16(kFfnTM, kFfnCH);
        const unsigned idesc2 = make_idesc_bf16(kFfnTM, kFfnC);
        const unsigned long long dW1 = make_desc_sw128(sW1), dW2 = make_desc_sw128(sW2), dH = make_desc_sw128(sH);
        unsigned it = 0;
        for (long long unit = first; unit < units; unit += stride, ++it) {
            // GEMM1(buf): Hacc[buf] = X (TMEM) @ W1buf[buf]^T
            auto gemm1 = [&](int buf) {
                tcgen05_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb)
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            mma_bf16_ts(tmem_h + buf * kFfnCH, tmem_x + (kb * 4 + j) * 8,
                                        desc_advance(dW1, buf * kSmemW1 + kb * kFfnCH * 128 + j * 32), idesc1, (kb | j) != 0);
                    mma_commit(&bars->g1[buf]);
                    if (kFfnCluster > 1) mma_commit_multicast(&bars->f1[buf], kAll);
                }
                __syncwarp();
            };
            mbar_wait(&bars->x, it & 1);                                     // X tile landed in the W2 buffers
            if (it > 0) mbar_wait(&bars->xy_free, (it - 1) & 1);             // previous tile's pass 1 is done with Yacc / X
            tcgen05_fence_after();
            if (elect_one()) {               // X: shared memory -> tensor memory, 16 slabs of 128 rows x 16 columns
#pragma unroll
                for (int kb = 0; kb < 4; ++kb)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        tmem_cp_128x256b(tmem_x + (kb * 4 + j) * 8, desc_advance(dW2, kb * kFfnTM * 128 + j * 32));
                mma_commit(&bars->xcopied);                                  // the W2 buffers may be refilled
                if (kFfnCluster > 1) mma_commit_multicast(&bars->xc_all, kAll);
            }
            __syncwarp();
            mbar_wait(&bars->w1[0], chunk_parity(it, 0, NC));
            gemm1(0);
            for (int c = 0; c < NC; ++c) {
                const int b = c & 1;
                FFN_TRACE(0, c, 0);
                if (c + 1 < NC) {                // queue GEMM1(c+1): Hacc[b^1] was drained by epilogue1(c-1)
                    mbar_wait(&bars->w1[b ^ 1], chunk_parity(it, c + 1, NC));
                    FFN_TRACE(0, c, 1);
                    gemm1(b ^ 1);
                }
                FFN_TRACE(0, c, 2);
                mbar_wait(&bars->hfull[b], chunk_parity(it, c, NC));         // sH[b] written (and Hacc[b] drained)
                FFN_TRACE(0, c, 3);
                mbar_wait(&bars->w2[b], chunk_parity(it, c, NC));
                FFN_TRACE(0, c, 4);
                tcgen05_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        mma_bf16(tmem_y, desc_advance(dH, b * kSmemH + j * 32), desc_advance(dW2, b * kSmemW2 + j * 32),
                                 idesc2, (c | j) != 0);
                    mma_commit(&bars->g2[b]);
                    if (kFfnCluster > 1) mma_commit_multicast(&bars->f2[b], kAll);
                }
                __syncwarp();
                FFN_TRACE(0, c, 5);
            }
        }
    } else if (warp == kFfnEpiWarps + 1) {
        // ======================================= TMA producer =======================================
        unsigned it = 0;
        for (long long unit = first; unit < units; unit += stride, ++it) {
            const long long tile = unit * kFfnCluster + rank;
            // a chunk of W1 is 4 K blocks of [64 rows x 128 B], a chunk of W2 one box of [256 rows x 128 B]: CTA `rank`
            // loads K blocks 2 rank, 2 rank + 1 and rows [128 rank, 128 rank + 128) and multicasts them to the cluster
            auto load_w1 = [&](int c, int buf) {
                if (elect_one()) {
                    mbar_expect_tx(&bars->w1[buf], kSmemW1);
                    if (kFfnCluster > 1) {
#pragma unroll
                        for (int kb = 2 * (int)rank; kb < 2 * (int)rank + 2; ++kb)
                            tma_load_2d_multicast(sW1 + buf * kSmemW1 + kb * kFfnCH * 128, &tm_w1, kb * 64, c * kFfnCH,
                                                  &bars->w1[buf], kAll);
                    } else {
#pragma unroll
                        for (int kb = 0; kb < 4; ++kb)
                            tma_load_2d(sW1 + buf * kSmemW1 + kb * kFfnCH * 128, &tm_w1, kb * 64, c * kFfnCH, &bars->w1[buf]);
                    }
                }
                __syncwarp();
            };
            auto load_w2 = [&](int c, int buf) {
                if (elect_one()) {
                    mbar_expect_tx(&bars->w2[buf], kSmemW2);
                    if (kFfnCluster > 1)
                        tma_load_2d_multicast(sW2 + buf * kSmemW2 + rank * (kSmemW2 / kFfnCluster), &tm_w2, c * kFfnCH,
                                              (int)rank * (kFfnC / kFfnCluster), &bars->w2[buf], kAll);
                    else
                        tma_load_2d(sW2 + buf * kSmemW2, &tm_w2, c * kFfnCH, 0, &bars->w2[buf]);
                }
                __syncwarp();
            };
            // every MMA of MY previous tile has completed: the W2 buffers may take my X tile
            if (it > 0) mbar_wait(&bars->g2[(NC - 1) & 1], chunk_parity(it - 1, NC - 1, NC));
            if (elect_one()) {
                mbar_expect_tx(&bars->x, kSmemX);
#pragma unroll
                for (int kb = 0; kb < 4; ++kb)
                    tma_load_2d(sW2 + kb * kFfnTM * 128, &tm_x, kb * 64, (int)(tile * kFfnTM), &bars->x);
            }
            __syncwarp();
            if (kFfnCluster > 1 && it > 0) {
                // the W1 buffers of EVERY CTA are free: the last GEMM1 on each buffer, cluster-wide
                if (NC > 1) mbar_wait(&bars->f1[(NC - 2) & 1], chunk_parity(it - 1, NC - 2, NC));
                mbar_wait(&bars->f1[(NC - 1) & 1], chunk_parity(it - 1, NC - 1, NC));
            }
            load_w1(0, 0);
            if (NC > 1) load_w1(1, 1);
            mbar_wait(&bars->xcopied, it & 1);       // X has been copied into tensor memory
            if (kFfnCluster > 1) mbar_wait(&bars->xc_all, it & 1);   // ... in every CTA: their W2 buffers held their X tiles
            load_w2(0, 0);
            for (int c = 0; c < NC; ++c) {
                const int b = c & 1;
                if (c + 2 < NC) {                // W1 buffer b is free once GEMM1(c) has completed (in every CTA)
                    mbar_wait(kFfnCluster > 1 ? &bars->f1[b] : &bars->g1[b], chunk_parity(it, c, NC));
                    load_w1(c + 2, b);
                }
                if (c + 1 < NC) {                // W2 buffer b^1 is free once GEMM2(c-1) has completed (in every CTA)
                    if (c >= 1) mbar_wait(kFfnCluster > 1 ? &bars->f2[b ^ 1] : &bars->g2[b ^ 1], chunk_parity(it, c - 1, NC));
                    load_w2(c + 1, b ^ 1);
                }
            }
        }
    } else if (warp < kFfnEpiWarps) {
        // ======================================= epilogue warps =======================================
        const int r = (warp & 3) * 32 + lane;          // row of the tile = TMEM lane
        const int hsel = warp >> 2;                    // which 1/SPLIT of the columns this warp handles
        const unsigned lane_base = (unsigned)((warp & 3) * 32) << 16;
        unsigned it = 0;
        for (long long unit = first; unit < units; unit += stride, ++it) {
            for (int c = 0; c < NC; ++c) {
                const int b = c & 1;
                constexpr int kCols1 = kFfnCH / kFfnSplit;         // Hacc columns per thread: 32 or 16
                uint4 braw[kCols1 / 8];
                const __nv_bfloat16* bias = b1 + c * kFfnCH + hsel * kCols1;
#pragma unroll
                for (int q = 0; q < kCols1 / 8; ++q) braw[q] = *reinterpret_cast<const uint4*>(bias + q * 8);
                if (tid == 0) FFN_TRACE(1, c, 0);
                mbar_wait(&bars->g1[b], chunk_parity(it, c, NC));            // Hacc[b] = X @ W1[c]^T done
                if (tid == 0) FFN_TRACE(1, c, 1);
                tcgen05_fence_after();
                float v[kCols1];
                if constexpr (kCols1 == 32) tmem_ld32(tmem_h + b * kFfnCH + hsel * kCols1 + lane_base, v);
                else tmem_ld16(tmem_h + b * kFfnCH + hsel * kCols1 + lane_base, v);
                if (tid == 0) FFN_TRACE(1, c, 2);
                if (c >= 2) mbar_wait(&bars->g2[b], chunk_parity(it, c - 2, NC));   // GEMM2(c-2) done reading sH[b]
                if (tid == 0) FFN_TRACE(1, c, 3);
                unsigned char* hrow = sH + b * kSmemH + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
                for (int q = 0; q < kCols1 / 8; ++q) {
                    float bb[8], o[8];
                    unpack<__nv_bfloat16>(braw[q], bb);
#pragma unroll
                    for (int e = 0; e < 8; ++e) o[e] = fmaxf(v[q * 8 + e] + bb[e], 0.f);
                    const int ch = hsel * (kCols1 / 8) + q;
                    *reinterpret_cast<uint4*>(hrow + ((ch ^ (r & 7)) << 4)) = pack<__nv_bfloat16>(o);
                }
                fence_proxy_async();
                tcgen05_fence_before();
                mbar_arrive(&bars->hfull[b]);
                if (tid == 0) FFN_TRACE(1, c, 4);
            }
            // ---- all MMAs of the tile done (tcgen05.commit covers every earlier MMA) ----
            mbar_wait(&bars->g2[(NC - 1) & 1], chunk_parity(it, NC - 1, NC));
            if (it > 0) mbar_wait(&bars->stage_free, (it - 1) & 1);          // store warps are done with the staging tile
            if (tid == 0) FFN_TRACE(1, 32, 0);
            tcgen05_fence_after();
            // pass 1 (thread = 1/SPLIT of a row): v = Yacc + b2 + x -> bf16 into the staging tile; partial sums of
            // v and v^2 of the rounded values (one pass: |mean| is of the order of the deviation for these
            // activations, so E[v^2] - mean^2 loses nothing at bf16 output precision)
            float sum = 0.f, sumsq = 0.f;
            constexpr int kColsY = kFfnC / kFfnSplit;             // Yacc columns per thread: 128 or 64
#pragma unroll 1
            for (int cb = 0; cb < kColsY / 32; ++cb) {
                float v[32], xp[16];
                tmem_ld32(tmem_y + hsel * kColsY + cb * 32 + lane_base, v);
                tmem_ld16(tmem_x + (hsel * kColsY + cb * 32) / 2 + lane_base, xp);      // 32 bf16 of x as 16 pairs
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int ch = hsel * (kColsY / 8) + cb * 4 + q;
                    float xr[8], bb[8], o[8];
                    unpack<__nv_bfloat16>(make_uint4(__float_as_uint(xp[q * 4]), __float_as_uint(xp[q * 4 + 1]),
                                                     __float_as_uint(xp[q * 4 + 2]), __float_as_uint(xp[q * 4 + 3])), xr);
                    unpack<__nv_bfloat16>(*reinterpret_cast<const uint4*>(b2 + ch * 8), bb);
#pragma unroll
                    for (int e = 0; e < 8; ++e) o[e] = v[q * 8 + e] + bb[e] + xr[e];
                    const uint4 packed = pack<__nv_bfloat16>(o);
                    *reinterpret_cast<uint4*>(sStage + sw128_offset(r, ch, kFfnTM)) = packed;
                    unpack<__nv_bfloat16>(packed, o);
#pragma unroll
                    for (int e = 0; e < 8; ++e) { sum += o[e]; sumsq = fmaf(o[e], o[e], sumsq); }
                }
            }
            s_stat[hsel * kFfnTM + r] = sum;
            s_stat[(kFfnSplit + hsel) * kFfnTM + r] = sumsq;
            tcgen05_fence_before();
            mbar_arrive(&bars->xy_free);             // Yacc and the X columns may be overwritten by the next tile
            mbar_arrive(&bars->stage_full);          // hand the staged tile to the store warps
            if (tid == 0) FFN_TRACE(1, 32, 2);
        }
    } else {
        // ======================================= store warps =======================================
        // pass 2 (coalesced: thread = one 16-byte column chunk of 32 rows): normalise, store y (+ pos).
        // The chunk index of a thread never changes, so its gamma / beta slice lives in registers.
        const int stid = tid - (kFfnEpiThreads + 64);
        const int my_ch = stid & 31;
        float my_gamma[8], my_beta[8];
        unpack<__nv_bfloat16>(*reinterpret_cast<const uint4*>(gamma + my_ch * 8), my_gamma);
        unpack<__nv_bfloat16>(*reinterpret_cast<const uint4*>(beta + my_ch * 8), my_beta);
        constexpr int kRowsPerSweep = kFfnStoreThreads >> 5;     // rows covered by one sweep of the store warps
        constexpr int kSweeps = kFfnTM / kRowsPerSweep;
        unsigned it = 0;
        for (long long unit = first; unit < units; unit += stride, ++it) {
            const long long row0 = (unit * kFfnCluster + rank) * kFfnTM;
            mbar_wait(&bars->stage_full, it & 1);
#pragma unroll 1
            for (int batch = 0; batch < kSweeps / 8; ++batch) {
                uint4 val[8], pp[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int rr = (stid >> 5) + (batch * 8 + u) * kRowsPerSweep;
                    const long long gr = row0 + rr;
                    val[u] = *reinterpret_cast<const uint4*>(sStage + sw128_offset(rr, my_ch, kFfnTM));
                    pp[u] = make_uint4(0u, 0u, 0u, 0u);
                    if (y_pos != nullptr && gr < rows) pp[u] = ldg_stream_v4(pos + gr * kFfnC + my_ch * 8);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int rr = (stid >> 5) + (batch * 8 + u) * kRowsPerSweep;
                    const long long gr = row0 + rr;
                    if (gr < rows) {
                        float s1 = 0.f, s2 = 0.f;
#pragma unroll
                        for (int h = 0; h < kFfnSplit; ++h) {
                            s1 += s_stat[h * kFfnTM + rr];
                            s2 += s_stat[(kFfnSplit + h) * kFfnTM + rr];
                        }
                        const float mean = s1 * (1.f / kFfnC);
                        const float var = fmaxf(s2 * (1.f / kFfnC) - mean * mean, 0.f);
                        const float rstd = rsqrtf(var + eps);
                        float o[8];
                        unpack<__nv_bfloat16>(val[u], o);
#pragma unroll
                        for (int e = 0; e < 8; ++e) o[e] = fmaf((o[e] - mean) * rstd, my_gamma[e], my_beta[e]);
                        const uint4 outv = pack<__nv_bfloat16>(o);
                        stg_stream_v4(y + gr * kFfnC + my_ch * 8, outv);
                        if (y_pos != nullptr) {
                            float pf[8];
                            unpack<__nv_bfloat16>(outv, o);          // y_pos is defined on the rounded y
                            unpack<__nv_bfloat16>(pp[u], pf);
#pragma unroll
                            for (int e = 0; e < 8; ++e) o[e] += pf[e];
                            stg_stream_v4(y_pos + gr * kFfnC + my_ch * 8, pack<__nv_bfloat16>(o));
                        }
                    }
                }
            }
            mbar_arrive(&bars->stage_free);          // the staging tile and the row statistics may be rewritten
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (kFfnCluster > 1) cluster_sync_all();       // no CTA leaves while a peer may still write its shared memory / barriers
    if (warp == 0) tmem_free(tmem, 512);
}

// ------------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// row-major bf16 matrix [n_rows, n_cols]; box = 64 columns (128 bytes, SWIZZLE_128B) x box_rows rows
static bool make_map(CUtensorMap* map, const void* base, unsigned long long n_rows, unsigned long long n_cols,
                     unsigned box_rows)
{
    EncodeTiledFn fn = encode_tiled_fn();
    if (fn == nullptr) return false;
    const cuuint64_t dims[2] = {n_cols, n_rows};
    const cuuint64_t strides[1] = {n_cols * 2};
    const cuuint32_t box[2] = {64, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool ffn_layernorm_supported(int dtype, int d_model, int d_ffn)
{
    return dtype == kBF16 && d_model == kFfnC && d_ffn >= kFfnCH && d_ffn % kFfnCH == 0;
}

cudaError_t ffn_layernorm_forward(const FfnArgs& a, cudaStream_t stream)
{
    if (!ffn_layernorm_supported(a.dtype, a.C, a.F) || (a.y_pos != nullptr) != (a.pos != nullptr))
        return cudaErrorInvalidValue;
    if (a.rows == 0) return cudaSuccess;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    static bool configured[64] = {};          // the attribute is per device: one process may drive several GPUs
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(ffn_layernorm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFfnSmem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    alignas(64) CUtensorMap tm_x, tm_w1, tm_w2;
    if (!make_map(&tm_x, a.x, (unsigned long long)a.rows, kFfnC, kFfnTM) ||
        !make_map(&tm_w1, a.w1, (unsigned long long)a.F, kFfnC, kFfnCH) ||
        !make_map(&tm_w2, a.w2, kFfnC, (unsigned long long)a.F, kFfnC / kFfnCluster))
        return cudaErrorNotSupported;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long tiles = (a.rows + kFfnTM - 1) / kFfnTM;
    const long long units = (tiles + kFfnCluster - 1) / kFfnCluster, clusters = sms / kFfnCluster;
    const int grid = (int)(units < clusters ? units : clusters) * kFfnCluster;
    using bf = __nv_bfloat16;
    ffn_layernorm_kernel<<<grid, kFfnThreads, kFfnSmem, stream>>>(
        tm_x, tm_w1, tm_w2, (const bf*)a.b1, (const bf*)a.b2, (const bf*)a.gamma, (const bf*)a.beta,
        (const bf*)a.pos, (bf*)a.y, (bf*)a.y_pos, a.rows, a.F, a.eps);
    return cudaGetLastError();
}

}  // namespace msda
