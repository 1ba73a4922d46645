// Output projection + residual + LayerNorm (+ next query) of the attention blocks on the 5th-generation tensor
// cores (sm_100a), bf16, d_model 256:
//
//     y     = LayerNorm( residual + (a @ W^T + b) ) * gamma + beta
//     y_pos = y + pos                                                          (optional second output)
//
// i.e. `output_proj` of MSDeformAttn (reference models/ops/modules/ms_deform_attn.py:116) followed by
// `src = norm1(src + dropout1(src2))` of the layer that owns it (models/deformable_transformer_single.py:538-541,
// :385-394 for the fusion layers' adapt Linear, :617-628 for the decoder), ONE kernel instead of a GEMM that writes
// [rows, 256] and an add + LayerNorm kernel that reads it back.
//
// Weights-stationary, transposed: the kernel computes Y^T = W @ A^T so that
//   * W (256 x 256 bf16) is the A operand of tcgen05.mma and lives in TENSOR MEMORY for the whole kernel (copied
//     once: TMA -> shared memory -> tcgen05.cp; 256 columns), not 128 KB of shared memory;
//   * the token tile (128 rows x 256, K-major SWIZZLE_128B straight from TMA, double buffered) is the B operand;
//   * the accumulator Y^T (2 blocks of 128 channels x 128 tokens fp32 = 256 TMEM columns) puts a CHANNEL in each
//     TMEM lane, so the epilogue thread adds its (per-lane constant) bias and writes bf16 into a token-major
//     staging tile with conflict-free 64-byte warp stores -- the transpose costs nothing;
//   * the staging tile IS the token buffer the MMAs have just finished reading, and the residual rows arrive by TMA
//     in a third buffer, so the LayerNorm warps (one warp per token row, statistics by shuffles) read everything
//     from shared memory and overlap with the next tile's MMA and epilogue.
// The projection result is rounded to bf16 before the residual add, exactly like the unfused chain (GEMM output in
// bf16, then add + LayerNorm in fp32).
// Warp roles (26 warps): 0-7 epilogue (warp & 3 = TMEM lane quadrant, warp >> 2 = channel block), 8 MMA issuer,
// 9 TMA producer, 10-25 LayerNorm + store.  TMEM: W 256 | Y^T 256.  Shared memory: token / staging tiles 2 x 64 KB
// (W lands there first) + residual tile 64 KB.
#include <cuda.h>

#include "msda_common.cuh"
#include "msda_launch.h"
#include "umma.cuh"

namespace msda {

using namespace umma;

constexpr int kPjC = 256;             // d_model = K = N
constexpr int kPjTM = 128;            // tokens per tile
constexpr int kPjEpiWarps = 8;
constexpr int kPjEpiThreads = 32 * kPjEpiWarps;
constexpr int kPjStoreWarps = 16;     // the LayerNorm is ALU-latency bound per warp: 4 warps per scheduler hide it
constexpr int kPjStoreThreads = 32 * kPjStoreWarps;
constexpr int kPjThreads = kPjEpiThreads + 64 + kPjStoreThreads;
constexpr int kPjSmemA = 4 * kPjTM * 128;         // 65536: 4 k-blocks of [128 rows x 128 B]
constexpr int kPjSmemRes = 4 * kPjTM * 128;       // 65536: the residual tile, same SWIZZLE_128B box layout
// 1: the residual is added BY THE TENSOR CORE (Y^T += I R^T with a 128 x 128 identity in shared memory as A operand and
// the residual tile as B operand): exact (the products are the residual's bf16 values, accumulated in fp32 with the
// projection before anything is rounded), the residual buffer is free again as soon as the MMAs are done (so the next
// tile's residual loads while this one is normalised), and the LayerNorm warps unpack one staged operand, not two.
// 0 (default): the residual is added by the LayerNorm warps from its shared-memory tile.
// Measured (8 x 22223 rows, two-pass LayerNorm at the time): 67.0 us with 0, 68.0 us with 1 (129.6 vs 117.0 us with the
// `pos` output).  Neither the instruction count of the LayerNorm phase nor its global stores were what a tile waited
// for (removing the stores left the phase at 7 k cycles); its shared-memory passes were -- the tensor-core residual
// only swaps LayerNorm reads for operand reads, hence the tie.  It also costs 32 KB more shared memory and 60 bytes of
// spills at the 72-register cap of an 832-thread CTA.  It stays as a build option (parity-tested on the B200 with
// either setting); the default path was then cut to ONE pass over the staged rows (61 us).
#ifndef MSDA_PROJ_TC_RESIDUAL
#define MSDA_PROJ_TC_RESIDUAL 0
#endif
constexpr int kPjSmemPart = MSDA_PROJ_TC_RESIDUAL ? 0 : kPjStoreWarps * 2 * (kPjTM / kPjStoreWarps) * 33 * 4;   // per-lane partial row statistics
constexpr int kPjSmemId = MSDA_PROJ_TC_RESIDUAL ? 2 * kPjTM * 128 : 0;   // 32768: identity, 2 k-blocks of [128 rows x 128 B]
constexpr int kPjSmemBars = 256;
constexpr int kPjSmem = 2 * kPjSmemA + kPjSmemRes + kPjSmemPart + kPjSmemId + kPjSmemBars;
static_assert(kPjSmem <= 232448, "exceeds the 227 KB of shared memory a CTA can opt in to");

#ifdef MSDA_PROJ_TRACE
__device__ long long g_proj_trace[4][8][8];      // [role][tile iteration][event]
#define PJ_TRACE(role, k) do { if (blockIdx.x == 0 && it < 8 && (threadIdx.x & 31) == 0) g_proj_trace[role][it][k] = clock64(); } while (0)
#else
#define PJ_TRACE(role, k) do { } while (0)
#endif

struct ProjBars {
    unsigned long long w_full, w_copied, a_full[2], mma_done[2], yacc_free, stage_full, res_full, store_done, store_done2[2];
    unsigned tmem_base;
};

__global__ void __launch_bounds__(kPjThreads, 1)
proj_layernorm_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w,
                      const __grid_constant__ CUtensorMap tm_r, const int has_residual,
                      const __nv_bfloat16* __restrict__ bias,
                      const __nv_bfloat16* __restrict__ gamma, const __nv_bfloat16* __restrict__ beta,
                      const __nv_bfloat16* __restrict__ pos, __nv_bfloat16* __restrict__ y,
                      __nv_bfloat16* __restrict__ y_pos, long long rows, float eps)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sA = smem;                              // [2][kPjSmemA]: token tile, then (MMAs done) its staged result
    unsigned char* sRes = smem + 2 * kPjSmemA;
    float* s_part = reinterpret_cast<float*>(sRes + kPjSmemRes);
    unsigned char* sId = sRes + kPjSmemRes + kPjSmemPart;
    ProjBars* bars = reinterpret_cast<ProjBars*>(sRes + kPjSmemRes + kPjSmemPart + kPjSmemId);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long tiles = (rows + kPjTM - 1) / kPjTM;

    if (warp == 0) tmem_alloc(&bars->tmem_base, 512);
    if (tid == kPjEpiThreads) {
        mbar_init(&bars->w_full, 1);
        mbar_init(&bars->w_copied, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&bars->a_full[i], 1); mbar_init(&bars->mma_done[i], 1); }
        mbar_init(&bars->yacc_free, kPjEpiThreads);
        mbar_init(&bars->stage_full, kPjEpiThreads);
        mbar_init(&bars->res_full, 1);
        mbar_init(&bars->store_done, kPjStoreThreads);
        mbar_init(&bars->store_done2[0], kPjStoreThreads);
        mbar_init(&bars->store_done2[1], kPjStoreThreads);
        fence_mbar_init();
        tma_prefetch_desc(&tm_a); tma_prefetch_desc(&tm_w); tma_prefetch_desc(&tm_r);
    }
#if MSDA_PROJ_TC_RESIDUAL
    // identity [128 x 128] bf16 in the K-major SWIZZLE_128B operand layout (16-byte chunk c of row r: ones at k = r)
    for (int idx = tid; idx < kPjTM * 16; idx += kPjThreads) {
        const int r = idx >> 4, c = idx & 15;
        const unsigned one = ((r >> 3) == c) ? ((r & 1) ? 0x3f800000u : 0x00003f80u) : 0u;   // bf16 1.0 at column k = r
        const int word = (r & 7) >> 1;
        *reinterpret_cast<uint4*>(sId + sw128_offset(r, c, kPjTM)) =
            make_uint4(word == 0 ? one : 0u, word == 1 ? one : 0u, word == 2 ? one : 0u, word == 3 ? one : 0u);
    }
    fence_proxy_async();
#endif
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const unsigned tmem = bars->tmem_base;
    const unsigned tmem_w = tmem;                  // columns [0, 256): W as A operand, block mb at mb * 128, K16 step s at s * 8
    const unsigned tmem_y = tmem + 256;            // columns [256, 512): Y^T block mb at mb * 128 (128 token columns)

    if (warp == kPjEpiWarps) {
        // ======================================= MMA issuer =======================================
        const unsigned idesc = make_idesc_bf16(kPjTM, kPjTM);            // M = 128 channels, N = 128 tokens
        const unsigned long long dA = make_desc_sw128(sA);
        // W: shared memory (both token buffers, one 128-channel block each) -> tensor memory, 32 slabs of 128 x 16
        mbar_wait(&bars->w_full, 0);
        tcgen05_fence_after();
        if (elect_one()) {
#pragma unroll
            for (int mb = 0; mb < 2; ++mb)
#pragma unroll
                for (int kb = 0; kb < 4; ++kb)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        tmem_cp_128x256b(tmem_w + mb * 128 + (kb * 4 + j) * 8,
                                         desc_advance(dA, mb * kPjSmemA + kb * kPjTM * 128 + j * 32));
            mma_commit(&bars->w_copied);                                 // the token buffers may be filled
        }
        __syncwarp();
        unsigned it = 0;
        for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            PJ_TRACE(0, 0);
            mbar_wait(&bars->a_full[buf], (it >> 1) & 1);                // token tile landed
            PJ_TRACE(0, 1);
            if (it > 0) mbar_wait(&bars->yacc_free, (it - 1) & 1);       // previous tile's epilogue has drained Y^T
            PJ_TRACE(0, 2);
            tcgen05_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int mb = 0; mb < 2; ++mb)
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb)
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            mma_bf16_ts(tmem_y + mb * 128, tmem_w + mb * 128 + (kb * 4 + j) * 8,
                                        desc_advance(dA, buf * kPjSmemA + kb * kPjTM * 128 + j * 32), idesc, (kb | j) != 0);
#if !MSDA_PROJ_TC_RESIDUAL
                mma_commit(&bars->mma_done[buf]);
#else
                if (!has_residual) mma_commit(&bars->mma_done[buf]);
#endif
            }
            __syncwarp();
#if MSDA_PROJ_TC_RESIDUAL
            if (has_residual) {
                mbar_wait(&bars->res_full, it & 1);                      // residual tile landed
                tcgen05_fence_after();
                if (elect_one()) {
                    const unsigned long long dI = make_desc_sw128(sId), dR = make_desc_sw128(sRes);
#pragma unroll
                    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
                        for (int s8 = 0; s8 < 8; ++s8)                   // Y^T[mb] += I (128 x 128) @ R[:, mb*128 : +128]^T
                            mma_bf16(tmem_y + mb * 128, desc_advance(dI, (s8 >> 2) * kPjTM * 128 + (s8 & 3) * 32),
                                     desc_advance(dR, (2 * mb + (s8 >> 2)) * kPjTM * 128 + (s8 & 3) * 32), idesc, true);
                    mma_commit(&bars->mma_done[buf]);                    // covers the projection MMAs issued above too
                }
                __syncwarp();
            }
#endif
            PJ_TRACE(0, 3);
        }
    } else if (warp == kPjEpiWarps + 1) {
        // ======================================= TMA producer =======================================
        if (elect_one()) {
            mbar_expect_tx(&bars->w_full, 2 * kPjSmemA);
#pragma unroll
            for (int mb = 0; mb < 2; ++mb)
#pragma unroll
                for (int kb = 0; kb < 4; ++kb)
                    tma_load_2d(sA + mb * kPjSmemA + kb * kPjTM * 128, &tm_w, kb * 64, mb * 128, &bars->w_full);
        }
        __syncwarp();
        mbar_wait(&bars->w_copied, 0);
        unsigned it = 0;
        for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
#if MSDA_PROJ_TC_RESIDUAL
            // buffer `buf` was the staging tile of tile it-2
            if (it >= 2) mbar_wait(&bars->store_done2[buf], ((it - 2) >> 1) & 1);
#endif
            PJ_TRACE(1, 0);
            if (elect_one()) {
                mbar_expect_tx(&bars->a_full[buf], kPjSmemA);
#pragma unroll
                for (int kb = 0; kb < 4; ++kb)
                    tma_load_2d(sA + buf * kPjSmemA + kb * kPjTM * 128, &tm_a, kb * 64, (int)(tile * kPjTM), &bars->a_full[buf]);
                if (tile + gridDim.x < tiles) {      // the next token tile: as far as L2 already
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb)
                        tma_prefetch_l2_2d(&tm_a, kb * 64, (int)((tile + gridDim.x) * kPjTM));
                }
            }
            __syncwarp();
#if MSDA_PROJ_TC_RESIDUAL
            // the residual buffer is free once the MMAs of the previous tile (its only readers) are done
            if (it > 0) mbar_wait(&bars->mma_done[buf ^ 1], ((it - 1) >> 1) & 1);
#else
            // (buffer `buf` of the NEXT iteration was the staging tile of tile it-1: awaited here)
            if (it > 0) mbar_wait(&bars->store_done, (it - 1) & 1);      // the residual buffer (and buffer buf^1) are free
#endif
            PJ_TRACE(1, 1);
            if (has_residual) {
                if (elect_one()) {
                    mbar_expect_tx(&bars->res_full, kPjSmemRes);
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb)
                        tma_load_2d(sRes + kb * kPjTM * 128, &tm_r, kb * 64, (int)(tile * kPjTM), &bars->res_full);
                    if (tile + gridDim.x < tiles) {  // and the next residual tile as far as L2
#pragma unroll
                        for (int kb = 0; kb < 4; ++kb)
                            tma_prefetch_l2_2d(&tm_r, kb * 64, (int)((tile + gridDim.x) * kPjTM));
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp < kPjEpiWarps) {
        // ======================================= epilogue warps =======================================
        // thread = one output channel (TMEM lane); 128 token columns -> + bias -> bf16 -> staging[token][channel]
        const int mb = warp >> 2;
        const int ch = mb * 128 + (warp & 3) * 32 + lane;
        const unsigned lane_base = (unsigned)((warp & 3) * 32) << 16;
        const float my_bias = __bfloat162float(bias[ch]);
        unsigned it = 0;
        for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            if (warp == 0) PJ_TRACE(2, 0);
            mbar_wait(&bars->mma_done[buf], (it >> 1) & 1);              // Y^T complete; the token buffer is dead -> staging
            if (warp == 0) PJ_TRACE(2, 1);
            __nv_bfloat16* stage = reinterpret_cast<__nv_bfloat16*>(sA + buf * kPjSmemA) + ch;
            tcgen05_fence_after();
#pragma unroll 1
            for (int tb = 0; tb < kPjTM / 32; ++tb) {
                float v[32];
                tmem_ld32(tmem_y + mb * 128 + tb * 32 + lane_base, v);
#pragma unroll
                for (int t = 0; t < 32; ++t) stage[(tb * 32 + t) * kPjC] = __float2bfloat16_rn(v[t] + my_bias);
            }
            tcgen05_fence_before();
            mbar_arrive(&bars->yacc_free);           // Y^T may be overwritten by the next tile's MMAs
            mbar_arrive(&bars->stage_full);          // hand the staged tile to the store warps
            if (warp == 0) PJ_TRACE(2, 2);
        }
    } else {
        // ======================================= LayerNorm + store warps =======================================
#if MSDA_PROJ_TC_RESIDUAL
        // The staged tile already holds bf16(projection + bias + residual).  A warp owns kPjTM / kPjStoreWarps token
        // rows, lane = one 16-byte chunk (8 channels) of the 256-wide row: every shared / global access is a
        // coalesced 512-byte row.  Pass 1: per-row partial (sum, sum of squares), the rows' 2 x 8 butterflies are
        // independent of one another (the other three warps of the scheduler fill the shuffle latency); pass 2
        // re-reads the rows, normalises with 2 FMAs per element and stores.
        const int swarp = warp - (kPjEpiWarps + 2);
        float g[8], b[8];
        unpack<__nv_bfloat16>(*reinterpret_cast<const uint4*>(gamma + lane * 8), g);
        unpack<__nv_bfloat16>(*reinterpret_cast<const uint4*>(beta + lane * 8), b);
        constexpr int kRowsPerWarp = kPjTM / kPjStoreWarps;      // 8
        unsigned it = 0;
        for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
            const long long row0 = tile * kPjTM;
            const unsigned char* sStage = sA + (it & 1) * kPjSmemA + (swarp * kRowsPerWarp) * (kPjC * 2) + lane * 16;
            if (swarp == 0) PJ_TRACE(3, 0);
            mbar_wait(&bars->stage_full, it & 1);
            if (swarp == 0) PJ_TRACE(3, 2);
            constexpr int kHalf = kRowsPerWarp / 2;      // 4 rows at a time: 8 statistics registers instead of 16
#pragma unroll 1
            for (int h0 = 0; h0 < kRowsPerWarp; h0 += kHalf) {
                float sum[kHalf], sq[kHalf];
#pragma unroll
                for (int r = 0; r < kHalf; ++r) {
                    float v[8];
                    unpack<__nv_bfloat16>(*reinterpret_cast<const uint4*>(sStage + (h0 + r) * (kPjC * 2)), v);
                    sum[r] = 0.f; sq[r] = 0.f;
#pragma unroll
                    for (int e = 0; e < 8; ++e) { sum[r] += v[e]; sq[r] = fmaf(v[e], v[e], sq[r]); }
                }
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) {
#pragma unroll
                    for (int r = 0; r < kHalf; ++r) {
                        sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], off);
                        sq[r] += __shfl_xor_sync(0xffffffffu, sq[r], off);
                    }
                }
#pragma unroll
                for (int r = 0; r < kHalf; ++r) {
                    const long long gr = row0 + swarp * kRowsPerWarp + h0 + r;
                    const float mean = sum[r] * (1.f / kPjC);
                    const float rstd = rsqrtf(fmaxf(sq[r] * (1.f / kPjC) - mean * mean, 0.f) + eps);
                    const float shift = -mean * rstd;
                    if (gr < rows) {
                        float v[8], o[8];
                        unpack<__nv_bfloat16>(*reinterpret_cast<const uint4*>(sStage + (h0 + r) * (kPjC * 2)), v);
#pragma unroll
                        for (int e = 0; e < 8; ++e) o[e] = fmaf(fmaf(v[e], rstd, shift), g[e], b[e]);
                        const uint4 outv = pack<__nv_bfloat16>(o);
                        stg_stream_v4(y + gr * kPjC + lane * 8, outv);
                        if (y_pos != nullptr) {
                            float pf[8];
                            unpack<__nv_bfloat16>(outv, o);          // y_pos is defined on the rounded y
                            unpack<__nv_bfloat16>(ldg_stream_v4(pos + gr * kPjC + lane * 8), pf);
#pragma unroll
                            for (int e = 0; e < 8; ++e) o[e] += pf[e];
                            stg_stream_v4(y_pos + gr * kPjC + lane * 8, pack<__nv_bfloat16>(o));
                        }
                    }
                }
            }
            if (swarp == 0) PJ_TRACE(3, 3);
            fence_proxy_async();                     // generic-proxy reads of a buffer the TMA (async proxy) refills next
            mbar_arrive(&bars->store_done2[it & 1]); // the staging buffer may be refilled with a token tile
        }
    }
#else
        // A warp owns kPjTM / kPjStoreWarps token rows; lane = one 16-byte chunk (8 channels) of the 256-wide row, so
        // every shared / global access is a coalesced 512-byte row.  The kernel is bound by the shared-memory pipe, so
        // each row is read ONCE (4 rows at a time stay in registers between the statistics and the normalisation),
        // and the statistics avoid shuffle chains (a warp cannot hide ten dependent shuffles per row): each lane
        // parks its partial (sum, sum of squares) of a row in shared memory, 8 lanes fold a row's 32 partials
        // (4 each, conflict-free, + 3 butterfly steps).
        const int swarp = warp - (kPjEpiWarps + 2);
        float g[8], b[8];
        unpack<__nv_bfloat16>(*reinterpret_cast<const uint4*>(gamma + lane * 8), g);
        unpack<__nv_bfloat16>(*reinterpret_cast<const uint4*>(beta + lane * 8), b);
        constexpr int kRowsPerWarp = kPjTM / kPjStoreWarps;      // 8
        constexpr int kGroup = 4;                                // rows kept in registers at a time
        float* part_sum = s_part + swarp * (2 * kRowsPerWarp * 33);
        float* part_sq = part_sum + kRowsPerWarp * 33;
        unsigned it = 0;
        for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
            const long long row0 = tile * kPjTM;
            const unsigned char* sStage = sA + (it & 1) * kPjSmemA;
            if (swarp == 0) PJ_TRACE(3, 0);
            mbar_wait(&bars->stage_full, it & 1);
            if (swarp == 0) PJ_TRACE(3, 1);
            if (has_residual) mbar_wait(&bars->res_full, it & 1);
            if (swarp == 0) PJ_TRACE(3, 2);
#pragma unroll 1
            for (int h0 = 0; h0 < kRowsPerWarp; h0 += kGroup) {
                float v[kGroup][8];
                // ---- v = bf16(projection) + residual, partial statistics ----
#pragma unroll
                for (int r = 0; r < kGroup; ++r) {
                    const int rr = swarp * kRowsPerWarp + h0 + r;
                    unpack<__nv_bfloat16>(*reinterpret_cast<const uint4*>(sStage + rr * (kPjC * 2) + lane * 16), v[r]);
                    if (has_residual) {
                        float rs[8];
                        unpack<__nv_bfloat16>(*reinterpret_cast<const uint4*>(sRes + sw128_offset(rr, lane, kPjTM)), rs);
#pragma unroll
                        for (int e = 0; e < 8; ++e) v[r][e] += rs[e];
                    }
                    float s = 0.f, q = 0.f;
#pragma unroll
                    for (int e = 0; e < 8; ++e) { s += v[r][e]; q = fmaf(v[r][e], v[r][e], q); }
                    part_sum[(h0 + r) * 33 + lane] = s;
                    part_sq[(h0 + r) * 33 + lane] = q;
                }
                __syncwarp();
                // ---- lanes 8r .. 8r+7 fold row r: 4 partials each, then 3 butterfly steps ----
                const int fr = h0 + (lane >> 3), seg = (lane & 7) * 4;
                float s = 0.f, q = 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i) { s += part_sum[fr * 33 + seg + i]; q += part_sq[fr * 33 + seg + i]; }
#pragma unroll
                for (int off = 1; off <= 4; off <<= 1) {
                    s += __shfl_xor_sync(0xffffffffu, s, off);
                    q += __shfl_xor_sync(0xffffffffu, q, off);
                }
                const float my_mean = s * (1.f / kPjC);
                const float my_rstd = rsqrtf(fmaxf(q * (1.f / kPjC) - my_mean * my_mean, 0.f) + eps);
                __syncwarp();
                // ---- normalise from registers, store y (+ pos) ----
#pragma unroll
                for (int r = 0; r < kGroup; ++r) {
                    const long long gr = row0 + swarp * kRowsPerWarp + h0 + r;
                    const float mean = __shfl_sync(0xffffffffu, my_mean, r * 8), rstd = __shfl_sync(0xffffffffu, my_rstd, r * 8);
                    if (gr < rows) {
                        float o[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) o[e] = fmaf((v[r][e] - mean) * rstd, g[e], b[e]);
                        const uint4 outv = pack<__nv_bfloat16>(o);
                        stg_stream_v4(y + gr * kPjC + lane * 8, outv);
                        if (y_pos != nullptr) {
                            float pf[8];
                            unpack<__nv_bfloat16>(outv, o);          // y_pos is defined on the rounded y
                            unpack<__nv_bfloat16>(ldg_stream_v4(pos + gr * kPjC + lane * 8), pf);
#pragma unroll
                            for (int e = 0; e < 8; ++e) o[e] += pf[e];
                            stg_stream_v4(y_pos + gr * kPjC + lane * 8, pack<__nv_bfloat16>(o));
                        }
                    }
                }
            }
            if (swarp == 0) PJ_TRACE(3, 3);
            fence_proxy_async();                     // generic-proxy reads of buffers the TMA (async proxy) refills next
            mbar_arrive(&bars->store_done);          // staging buffer and residual buffer may be refilled
        }
    }
#endif
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free(tmem, 512);
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*PjEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PjEncodeTiledFn pj_encode_tiled_fn()
{
    static PjEncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PjEncodeTiledFn>(p);
    }
    return fn;
}

// row-major bf16 matrix [n_rows, 256]; box = 64 columns (128 bytes, SWIZZLE_128B) x 128 rows
static bool pj_make_map(CUtensorMap* map, const void* base, unsigned long long n_rows)
{
    PjEncodeTiledFn fn = pj_encode_tiled_fn();
    if (fn == nullptr) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)kPjC, n_rows};
    const cuuint64_t strides[1] = {(cuuint64_t)kPjC * 2};
    const cuuint32_t box[2] = {64, (cuuint32_t)kPjTM};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool proj_layernorm_supported(int dtype, int d_in, int d_out)
{
    return dtype == kBF16 && d_in == kPjC && d_out == kPjC;
}

cudaError_t proj_layernorm_forward(const ProjArgs& a, cudaStream_t stream)
{
    if (!proj_layernorm_supported(a.dtype, a.C, a.C) || (a.y_pos != nullptr) != (a.pos != nullptr))
        return cudaErrorInvalidValue;
    if (a.rows == 0) return cudaSuccess;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    static bool configured[64] = {};          // the attribute is per device: one process may drive several GPUs
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(proj_layernorm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPjSmem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    alignas(64) CUtensorMap tm_a, tm_w, tm_r;
    const void* res_base = a.residual != nullptr ? a.residual : a.x;     // a valid map either way; unused without residual
    if (!pj_make_map(&tm_a, a.x, (unsigned long long)a.rows) || !pj_make_map(&tm_w, a.w, kPjC) ||
        !pj_make_map(&tm_r, res_base, (unsigned long long)a.rows))
        return cudaErrorNotSupported;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long tiles = (a.rows + kPjTM - 1) / kPjTM;
    const int grid = (int)(tiles < sms ? tiles : sms);
    using bf = __nv_bfloat16;
    proj_layernorm_kernel<<<grid, kPjThreads, kPjSmem, stream>>>(
        tm_a, tm_w, tm_r, a.residual != nullptr ? 1 : 0, (const bf*)a.b, (const bf*)a.gamma, (const bf*)a.beta, (const bf*)a.pos,
        (bf*)a.y, (bf*)a.y_pos, a.rows, a.eps);
    return cudaGetLastError();
}

}  // namespace msda
