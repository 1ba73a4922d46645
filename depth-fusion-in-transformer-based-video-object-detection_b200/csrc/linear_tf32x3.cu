// fp32-grade linear layer on the 5th-generation tensor cores (sm_100a):
//
//     y = x W^T + b   (optionally ReLU),   x [rows, K], W [N, K], y [rows, N], everything fp32 in HBM.
//
// The reference model is fp32 end to end (every nn.Linear of models/ops/modules/ms_deform_attn.py:94-116 and
// models/deformable_transformer_single.py:544-548 is an IEEE SGEMM there); on B200 the SIMT SGEMM is 40 of the 45 ms of a
// 6-layer fp32 encoder.  This kernel evaluates the product with the error-compensated three-term TF32 split
//
//     x = x_hi + x_lo,  W = W_hi + W_lo      (*_hi = the operand on TF32's 10 mantissa bits, *_lo = the exact rest)
//     x W^T ~= x_lo W_hi^T + x_hi W_lo^T + x_hi W_hi^T          (fp32 accumulation in tensor memory)
//
// whose dropped term and operand roundings are O(2^-22) -- fp32-grade results (tests/test_gpu_layer_epilogue.py) at a
// third of the TF32 tensor rate.  Round 1 ran the same split as one library TF32 GEMM over a 3K-long reduction, which
// needs a separate pass that writes [lo | hi | hi] (3x the activation) to HBM first: 6.6 of its 20.1 ms.  Here the split
// of the activation tile happens in SHARED MEMORY between the TMA load and the MMA; W_hi / W_lo are split once on the
// host side of the C ABI (weights are static during inference).
//
// One CTA per SM, clusters of two, persistent over (row tile, column tile) pairs, 128 x 256 output tile per CTA, K in blocks of 16 fp32
// (64-byte SWIZZLE_64B rows).  Warp roles (14 warps):
//   0-3, 10-13  epilogue   tcgen05.ld of the finished accumulators (lane = output row), + bias, ReLU, staged 32 x 32 and written
//                   with TMA stores
//   4-7  converter  x tile in shared memory -> x_lo (second tile, same swizzled layout: the split is elementwise, so it
//                   never has to know the layout), then fence.proxy.async; x_hi is the x tile as the tensor core reads it
//   8    MMA issuer 2 K-steps x 3 terms of tcgen05.mma kind::tf32 per K block; TWO accumulators in TMEM (see there)
//   9    TMA producer  x tile and this CTA's half of the W_hi / W_lo tiles (multicast to the cluster) per K block, four
//                      stages of 48 KB
#include <cuda.h>

#include "msda_common.cuh"
#include "msda_launch.h"
#include "umma.cuh"

namespace msda {

using namespace umma;

constexpr int kLtBM = 128;                 // output rows per tile (TMEM lanes)
constexpr int kLtBN = 256;                 // output columns per tile (TMEM columns of one accumulator)
// K block: 16 fp32 = 64-byte SWIZZLE_64B rows, four stages of 48 KB (default), or 32 fp32 = one 128-byte SWIZZLE_128B row,
// two stages of 96 KB (-DMSDA_TF32_BK=32).  The same bytes in flight in finer grains: measured at 8 x 22 223 rows, BK 32 |
// BK 16, ms: 256 <- 256: 0.122 | 0.116; 384 <- 256: 0.190 | 0.175; 1024 <- 256: 0.412 | 0.417; 256 <- 1024: 0.380 | 0.370.
#ifndef MSDA_TF32_BK
#define MSDA_TF32_BK 16
#endif
constexpr int kLtBK = MSDA_TF32_BK;
constexpr int kLtRowBytes = kLtBK * 4;     // bytes of a K block row = swizzle span
constexpr int kLtKSteps = kLtBK / 8;       // tcgen05.mma kind::tf32 takes K = 8 per instruction
constexpr int kLtStages = kLtBK == 32 ? 2 : 4;
static_assert(kLtBK == 32 || kLtBK == 16, "K block = one SWIZZLE_128B or SWIZZLE_64B row");
constexpr int kLtCluster = 2;              // CTAs that share every W tile (TMA multicast)
constexpr int kLtABytes = kLtBM * kLtRowBytes;     // 16 KB
constexpr int kLtBBytes = kLtBN * kLtRowBytes;     // 32 KB
constexpr int kLtStageBytes = 2 * kLtABytes + 2 * kLtBBytes;      // x_hi | x_lo | W_hi | W_lo = 96 KB
constexpr int kLtEpiThreads = 256, kLtConvThreads = 128;      // epilogue: warps 0-3 and 10-13 (two per TMEM lane quadrant)
constexpr int kLtThreads = kLtEpiThreads + kLtConvThreads + 64;
constexpr int kLtOutBytes = 32 * 128;      // one epilogue warp's staging tile: 32 rows x 32 fp32, SWIZZLE_128B
constexpr int kLtSmem = kLtStages * kLtStageBytes + 8 * kLtOutBytes + kLtBN * 4 + 256;
static_assert(kLtSmem <= 232448, "exceeds the 227 KB of shared memory a CTA can opt in to");

struct LtBars {
    unsigned long long full[kLtStages], conv[kLtStages], empty[kLtStages], acc_full, acc_free;
    unsigned tmem_base;
};

// instruction descriptor, kind::tf32: D fp32, A / B TF32 (format 2), both K-major, M x N tile
__device__ __forceinline__ unsigned make_idesc_tf32(int m, int n)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(n >> 3) << 17) | ((unsigned)(m >> 4) << 24);
}
// shared-memory matrix descriptor of a K-major operand tile whose rows are one swizzle span (128 or 64 bytes): 8-row groups
// of 8 x span bytes
__device__ __forceinline__ unsigned long long lt_desc(const void* smem_ptr)
{
    if constexpr (kLtBK == 32) return make_desc_sw128(smem_ptr);
    const unsigned addr = smem_u32(smem_ptr);
    unsigned long long d = 0;
    d |= (unsigned long long)((addr & 0x3FFFF) >> 4);            // start address
    d |= (unsigned long long)1 << 16;                            // leading byte offset (unused for swizzled K-major)
    d |= (unsigned long long)(512 >> 4) << 32;                   // stride between 8-row groups: 8 x 64 bytes
    d |= (unsigned long long)1 << 46;                            // descriptor version (sm_100)
    d |= (unsigned long long)4 << 61;                            // layout: SWIZZLE_64B
    return d;
}
__device__ __forceinline__ void mma_tf32(unsigned tmem_d, unsigned long long desc_a, unsigned long long desc_b,
                                         unsigned idesc, bool accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"((unsigned)accumulate) : "memory");
}

// W tiles are what the kernel streams: 64 of the 80 KB a K block brings in, re-read from L2 by every CTA for every tile --
// at full tensor rate 148 CTAs would ask L2 for 14 TB/s, and the single-CTA version of this kernel sat at the ~6-8 TB/s
// the L2 -> SM fabric delivers (ncu: 911 MB in 153 us, tensor pipe 40 % active).  Two CTAs of a cluster therefore work on
// two ROW tiles of the same column tile in lockstep, each loads HALF of every W tile and multicasts it into both CTAs'
// shared memory: 48 KB per K block and CTA.  A stage is released by the MMAs of BOTH CTAs (multicast tcgen05.commit on
// the `empty` barrier of each).
__global__ void __cluster_dims__(kLtCluster, 1, 1) __launch_bounds__(kLtThreads, 1)
linear_tf32x3_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_wh,
                     const __grid_constant__ CUtensorMap tm_wl, const __grid_constant__ CUtensorMap tm_y,
                     const float* __restrict__ bias, long long rows, int N, int K, int relu, int bn)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* s_out = smem + kLtStages * kLtStageBytes;            // [8 epilogue warps][kLtOutBytes], 1024-byte aligned
    float* s_bias = reinterpret_cast<float*>(s_out + 8 * kLtOutBytes);
    LtBars* bars = reinterpret_cast<LtBars*>(s_out + 8 * kLtOutBytes + kLtBN * 4);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long tiles_m = (rows + kLtBM - 1) / kLtBM;
    // bn: column tile width chosen by the host (<= kLtBN, a multiple of 32 that splits N evenly: 384 -> 2 x 192)
    const int tiles_n = (N + bn - 1) / bn;
    // work unit of a cluster: kLtCluster consecutive row tiles x one column tile; CTA `rank` takes row tile group * C + rank
    // (past the last row tile: TMA reads zeros, the epilogue stores nothing)
    const long long tiles = ((tiles_m + kLtCluster - 1) / kLtCluster) * tiles_n;
    const int kblocks = K / kLtBK;
    const unsigned rank = cluster_ctarank();
    const long long first = blockIdx.x / kLtCluster, stride = gridDim.x / kLtCluster;
    constexpr unsigned short kAll = (unsigned short)((1u << kLtCluster) - 1u);

    if (warp == 0) tmem_alloc(&bars->tmem_base, 512);
    if (tid == 128) {
        for (int i = 0; i < kLtStages; ++i) {
            mbar_init(&bars->full[i], 1);
            mbar_init(&bars->conv[i], kLtConvThreads);
            mbar_init(&bars->empty[i], kLtCluster);
        }
        mbar_init(&bars->acc_full, 1);
        mbar_init(&bars->acc_free, kLtEpiThreads);
        fence_mbar_init();
        tma_prefetch_desc(&tm_x); tma_prefetch_desc(&tm_wh); tma_prefetch_desc(&tm_wl); tma_prefetch_desc(&tm_y);
    }
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();                            // every CTA's barriers exist before anything arrives on them
    tcgen05_fence_after();
    const unsigned tmem = bars->tmem_base;

    if (warp == 9) {
        // ======================================= TMA producer =======================================
        if (elect_one()) {
            unsigned kiter = 0;
            const int kSlice = bn / kLtCluster, kSliceBytes = kSlice * kLtRowBytes;
            for (long long t = first; t < tiles; t += stride) {
                const int m0 = (int)((t / tiles_n) * kLtCluster + rank) * kLtBM, n0 = (int)(t % tiles_n) * bn;
                for (int kb = 0; kb < kblocks; ++kb, ++kiter) {
                    const int s = kiter % kLtStages;
                    // both CTAs are done with the stage: my slices land in the peer's shared memory too
                    if (kiter >= kLtStages) mbar_wait(&bars->empty[s], ((kiter / kLtStages) - 1) & 1);
                    unsigned char* st = smem + s * kLtStageBytes;
                    mbar_expect_tx(&bars->full[s], kLtABytes + 2 * bn * kLtRowBytes);   // my x tile + every CTA's W slices
                    tma_load_2d(st, &tm_x, kb * kLtBK, m0, &bars->full[s]);
                    tma_load_2d_multicast(st + 2 * kLtABytes + rank * kSliceBytes, &tm_wh, kb * kLtBK,
                                          n0 + (int)rank * kSlice, &bars->full[s], kAll);
                    tma_load_2d_multicast(st + 2 * kLtABytes + kLtBBytes + rank * kSliceBytes, &tm_wl, kb * kLtBK,
                                          n0 + (int)rank * kSlice, &bars->full[s], kAll);
                }
            }
        }
        __syncwarp();
    } else if (warp == 8) {
        // ======================================= MMA issuer =======================================
        unsigned kiter = 0, it = 0;
        for (long long t = first; t < tiles; t += stride, ++it) {
            const int n0 = (int)(t % tiles_n) * bn;
            const int n_cur = min(bn, N - n0);
            const unsigned idesc = make_idesc_tf32(kLtBM, n_cur);
            // Two accumulators: the hi x hi products and the two correction terms.  The tensor core truncates when it adds
            // into the accumulator, about half an ulp of the RUNNING SUM per instruction and always towards zero; 96
            // instructions into one accumulator measured 2.3e-6 of the output range (K = 256; 8.5e-6 at K = 1024).  The
            // correction terms are 2^-11 of the sum, so in an accumulator of their own their truncations do not matter,
            // and the main accumulator sees K / 8 instructions, like one plain TF32 GEMM.
            const unsigned acc_big = tmem, acc_small = tmem + kLtBN;
            if (it >= 1) mbar_wait(&bars->acc_free, (it - 1) & 1);               // the epilogue has drained the accumulators
            for (int kb = 0; kb < kblocks; ++kb, ++kiter) {
                const int s = kiter % kLtStages;
                unsigned char* st = smem + s * kLtStageBytes;
                const unsigned long long d_xh = lt_desc(st), d_xl = lt_desc(st + kLtABytes);
                const unsigned long long d_wh = lt_desc(st + 2 * kLtABytes), d_wl = lt_desc(st + 2 * kLtABytes + kLtBBytes);
                mbar_wait(&bars->full[s], (kiter / kLtStages) & 1);              // x and W tiles landed
                tcgen05_fence_after();
                // the two terms with x_hi need nothing from the converter: the tensor core reads the top 19 bits of the
                // fp32 words TMA wrote (x_hi = x truncated to TF32), so they go first and cover the converter's latency
                if (elect_one()) {
#pragma unroll
                    for (int j = 0; j < kLtKSteps; ++j) {                        // K = 8 per instruction: 32 bytes of the row
                        mma_tf32(acc_small, desc_advance(d_xh, j * 32), desc_advance(d_wl, j * 32), idesc, (kb | j) != 0);
                        mma_tf32(acc_big, desc_advance(d_xh, j * 32), desc_advance(d_wh, j * 32), idesc, (kb | j) != 0);
                    }
                }
                __syncwarp();
                mbar_wait(&bars->conv[s], (kiter / kLtStages) & 1);              // x_lo written
                tcgen05_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int j = 0; j < kLtKSteps; ++j)
                        mma_tf32(acc_small, desc_advance(d_xl, j * 32), desc_advance(d_wh, j * 32), idesc, true);
                    mma_commit_multicast(&bars->empty[s], kAll);                 // stage free (here and in the peer) once these MMAs have read it
                    if (kb == kblocks - 1) mma_commit(&bars->acc_full);
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ======================================= converter warps =======================================
        const int ct = tid - 128;
        unsigned kiter = 0;
        for (long long t = first; t < tiles; t += stride) {
            for (int kb = 0; kb < kblocks; ++kb, ++kiter) {
                const int s = kiter % kLtStages;
                mbar_wait(&bars->full[s], (kiter / kLtStages) & 1);
                const uint4* xh = reinterpret_cast<const uint4*>(smem + s * kLtStageBytes);
                uint4* xl = reinterpret_cast<uint4*>(smem + s * kLtStageBytes + kLtABytes);
                // x_hi is what the tensor core sees of x: the word with its 13 low mantissa bits cleared (the x tile is not
                // rewritten).  x_lo = x - x_hi is exact in fp32 (at most 13 significant bits); it is rounded to nearest on
                // TF32's 10 mantissa bits here, so that the tensor core's own truncation of it is a no-op.
                auto lo = [](unsigned v) -> unsigned {
                    const float l = __uint_as_float(v) - __uint_as_float(v & 0xffffe000u);
                    return (__float_as_uint(l) + 0x1000u) & 0xffffe000u;
                };
#pragma unroll
                for (int i = 0; i < kLtABytes / 16 / kLtConvThreads; ++i) {
                    const int idx = i * kLtConvThreads + ct;                     // consecutive lanes, consecutive chunks
                    const uint4 v = xh[idx];
                    xl[idx] = make_uint4(lo(v.x), lo(v.y), lo(v.z), lo(v.w));
                }
                fence_proxy_async();                 // generic-proxy writes -> tensor-core (async proxy) reads
                mbar_arrive(&bars->conv[s]);
            }
        }
    } else {
        // ======================================= epilogue warps =======================================
        // two warps per TMEM lane quadrant (warp % 4): warps 0-3 take the even 32-column chunks, warps 10-13 the odd ones
        const int quad = warp & 3, half = warp < 4 ? 0 : 1, et = half * 128 + quad * 32 + lane;
        const unsigned lane_base = (unsigned)(quad * 32) << 16;
        unsigned it = 0;
        for (long long t = first; t < tiles; t += stride, ++it) {
            const long long m0 = ((t / tiles_n) * kLtCluster + rank) * kLtBM;
            const int n0 = (int)(t % tiles_n) * bn;
            const int n_cur = min(bn, N - n0);
            // bias slice of the tile (the previous tile's readers are past their last read: barrier below)
            named_bar_sync(1, kLtEpiThreads);
            for (int i = et; i < kLtBN; i += kLtEpiThreads) s_bias[i] = (bias != nullptr && i < n_cur) ? bias[n0 + i] : 0.f;
            named_bar_sync(1, kLtEpiThreads);
            mbar_wait(&bars->acc_full, it & 1);
            tcgen05_fence_after();
            // 32 rows x 32 columns at a time: TMEM -> registers (lane = row) -> this warp's staging tile in the SWIZZLE_128B
            // layout (16-byte chunk c of row r at chunk c ^ (r & 7): conflict-free for lane = row) -> ONE TMA store of
            // the box (full 128-byte lines; rows past the end of y are clipped by the tensor map).  The staging tile is
            // reused once the previous store has read it.
            const int row0 = (int)m0 + quad * 32;
            unsigned char* stage = s_out + (half * 4 + quad) * kLtOutBytes;
            for (int c = half * 32; c < n_cur; c += 64) {
                float v[32], w[32];
                tmem_ld32(tmem + c + lane_base, v);
                tmem_ld32(tmem + kLtBN + c + lane_base, w);
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 b4 = *reinterpret_cast<const float4*>(s_bias + c + 4 * q);
                    float4 o = make_float4(v[4 * q] + w[4 * q] + b4.x, v[4 * q + 1] + w[4 * q + 1] + b4.y,
                                           v[4 * q + 2] + w[4 * q + 2] + b4.z, v[4 * q + 3] + w[4 * q + 3] + b4.w);
                    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                    *reinterpret_cast<float4*>(stage + lane * 128 + ((q ^ (lane & 7)) << 4)) = o;
                }
                fence_proxy_async();                 // generic-proxy writes -> TMA (async proxy) read
                __syncwarp();
                if (lane == 0) {
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];"
                                 :: "l"(&tm_y), "r"(smem_u32(stage)), "r"(n0 + c), "r"(row0) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
            tcgen05_fence_before();
            mbar_arrive(&bars->acc_free);
        }
    }
    if ((warp < 4 || warp >= 10) && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // my stores have left shared memory
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();                            // no CTA leaves while a peer may still write its shared memory / barriers
    if (warp == 0) tmem_free(tmem, 512);
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*LtEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static LtEncodeTiledFn lt_encode_tiled_fn()
{
    static LtEncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<LtEncodeTiledFn>(p);
    }
    return fn;
}

// row-major fp32 matrix [n_rows, n_cols]; box = 32 columns (128 bytes, SWIZZLE_128B) x box_rows rows; rows past the end
// read as zeros
static bool lt_make_map(CUtensorMap* map, const void* base, unsigned long long n_rows, unsigned long long n_cols, int box_rows,
                        int box_cols = kLtBK)
{
    LtEncodeTiledFn fn = lt_encode_tiled_fn();
    if (fn == nullptr) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)n_cols, n_rows};
    const cuuint64_t strides[1] = {(cuuint64_t)n_cols * 4};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
              CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool linear_tf32x3_supported(int n, int k)
{
    return n >= 32 && n % 32 == 0 && k >= 32 && k % 32 == 0;
}

cudaError_t linear_tf32x3(const float* x, const float* w_hi, const float* w_lo, const float* bias, long long rows, int n,
                          int k, int relu, float* y, cudaStream_t stream)
{
    if (!linear_tf32x3_supported(n, k) || rows < 0 || rows >= (1ll << 31) - 1024) return cudaErrorInvalidValue;   // row coordinates are int
    if (rows == 0) return cudaSuccess;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w_hi) | reinterpret_cast<uintptr_t>(w_lo) |
         reinterpret_cast<uintptr_t>(y)) % 16 != 0)
        return cudaErrorInvalidValue;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    static bool configured[64] = {};          // the attribute is per device: one process may drive several GPUs
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(linear_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kLtSmem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    // column tile width: as few tiles as 256 columns allow, of equal width (a multiple of 32; each CTA of the cluster
    // loads bn / 2 weight rows, a multiple of the 8-row swizzle group)
    const int tiles_n = (n + kLtBN - 1) / kLtBN;
    const int bn = ((n + tiles_n - 1) / tiles_n + 31) / 32 * 32;
    alignas(64) CUtensorMap tm_x, tm_wh, tm_wl, tm_y;
    if (!lt_make_map(&tm_x, x, (unsigned long long)rows, (unsigned long long)k, kLtBM) ||
        !lt_make_map(&tm_wh, w_hi, (unsigned long long)n, (unsigned long long)k, bn / kLtCluster) ||
        !lt_make_map(&tm_wl, w_lo, (unsigned long long)n, (unsigned long long)k, bn / kLtCluster) ||
        !lt_make_map(&tm_y, y, (unsigned long long)rows, (unsigned long long)n, 32, 32))
        return cudaErrorNotSupported;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long tiles_m = (rows + kLtBM - 1) / kLtBM;
    const long long units = ((tiles_m + kLtCluster - 1) / kLtCluster) * ((n + bn - 1) / bn);
    const long long clusters = sms / kLtCluster;
    const int grid = (int)(units < clusters ? units : clusters) * kLtCluster;
    linear_tf32x3_kernel<<<grid, kLtThreads, kLtSmem, stream>>>(tm_x, tm_wh, tm_wl, tm_y, bias, rows, n, k, relu, bn);
    return cudaGetLastError();
}

}  // namespace msda
