// bf16 linear layer of an inference pass on the 5th-generation tensor cores (sm_100a):
//
//     y = x W^T + b   (optionally ReLU, optionally rows zeroed by a padding mask),   x [rows, K], W [N, K], y [rows, N], bf16.
//
// `value_proj` of MSDeformAttn with its `masked_fill(padding_mask, 0)` (reference models/ops/modules/ms_deform_attn.py:94-96),
// the [sampling_offsets | attention_weights] projection (:98-100) and every other nn.Linear of a bf16 inference pass.  The
// product is memory-bound (91 MB in, 91-137 MB out at the encoder's 8 x 22 223 rows against 23-35 GFLOP), so the kernel is
// built around the HBM streams: x tiles by TMA through a four-stage ring, the weight tile shared by the two CTAs of a
// cluster (each loads half, TMA multicast), fp32 accumulators double-buffered in tensor memory so that the epilogue of a
// tile -- bias, ReLU, padding rows, bf16, 32 x 64 staging tiles in the SWIZZLE_128B layout, TMA stores of full lines --
// runs under the next tile's loads and MMAs.  Same skeleton as linear_tf32x3.cu without the operand split.
// Warp roles (10 warps): 0-7 epilogue (warp % 4 = TMEM lane quadrant, warp / 4 = which 64-column chunks), 8 MMA issuer,
// 9 TMA producer.
#include <cuda.h>

#include "msda_common.cuh"
#include "msda_launch.h"
#include "umma.cuh"

namespace msda {

using namespace umma;

constexpr int kLbBM = 128;                 // output rows per tile (TMEM lanes)
constexpr int kLbBN = 256;                 // output columns per tile (TMEM columns of one accumulator)
// K block: 64 bf16 = one 128-byte SWIZZLE_128B row, four stages of 48 KB (default), or 32 bf16 = 64-byte SWIZZLE_64B rows,
// eight stages of 24 KB (-DMSDA_LINEAR_BF16_BK=32: what helped the TF32 kernel is slower here -- 38.2 -> 41.9 us at
// 256 <- 256, 125 -> 131 us at 1024 <- 256)
#ifndef MSDA_LINEAR_BF16_BK
#define MSDA_LINEAR_BF16_BK 64
#endif
constexpr int kLbBK = MSDA_LINEAR_BF16_BK;
constexpr int kLbRowBytes = kLbBK * 2;
constexpr int kLbKSteps = kLbBK / 16;      // tcgen05.mma kind::f16 takes K = 16 per instruction
constexpr int kLbStages = kLbBK == 64 ? 4 : 8;
static_assert(kLbBK == 64 || kLbBK == 32, "K block = one SWIZZLE_128B or SWIZZLE_64B row");
#ifndef MSDA_LINEAR_BF16_CLUSTER
#define MSDA_LINEAR_BF16_CLUSTER 2
#endif
constexpr int kLbCluster = MSDA_LINEAR_BF16_CLUSTER;   // CTAs that share every weight tile (TMA multicast); 1 = none
constexpr int kLbABytes = kLbBM * kLbRowBytes;     // 16 KB
constexpr int kLbBBytes = kLbBN * kLbRowBytes;     // 32 KB
constexpr int kLbStageBytes = kLbABytes + kLbBBytes;
constexpr int kLbEpiThreads = 256;
constexpr int kLbThreads = kLbEpiThreads + 64;
constexpr int kLbOutBytes = 32 * 128;      // one epilogue warp's staging tile: 32 rows x 64 bf16
constexpr int kLbSmem = kLbStages * kLbStageBytes + 8 * kLbOutBytes + kLbBN * 4 + 256;
static_assert(kLbSmem <= 232448, "exceeds the 227 KB of shared memory a CTA can opt in to");

struct LbBars {
    unsigned long long full[kLbStages], empty[kLbStages], acc_full[2], acc_free[2];
    unsigned tmem_base;
};

// shared-memory matrix descriptor of a K-major operand tile whose rows are one swizzle span (128 or 64 bytes)
__device__ __forceinline__ unsigned long long lb_desc(const void* smem_ptr)
{
    if constexpr (kLbBK == 64) return make_desc_sw128(smem_ptr);
    const unsigned addr = smem_u32(smem_ptr);
    unsigned long long d = 0;
    d |= (unsigned long long)((addr & 0x3FFFF) >> 4);            // start address
    d |= (unsigned long long)1 << 16;                            // leading byte offset (unused for swizzled K-major)
    d |= (unsigned long long)(512 >> 4) << 32;                   // stride between 8-row groups: 8 x 64 bytes
    d |= (unsigned long long)1 << 46;                            // descriptor version (sm_100)
    d |= (unsigned long long)4 << 61;                            // layout: SWIZZLE_64B
    return d;
}

__global__ void __cluster_dims__(kLbCluster, 1, 1) __launch_bounds__(kLbThreads, 1)
linear_bf16_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                   const __grid_constant__ CUtensorMap tm_y, const __nv_bfloat16* __restrict__ bias,
                   const unsigned char* __restrict__ row_mask, long long rows, int N, int K, int relu, int bn)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* s_out = smem + kLbStages * kLbStageBytes;            // [8 epilogue warps][kLbOutBytes], 1024-byte aligned
    float* s_bias = reinterpret_cast<float*>(s_out + 8 * kLbOutBytes);
    LbBars* bars = reinterpret_cast<LbBars*>(s_out + 8 * kLbOutBytes + kLbBN * 4);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long tiles_m = (rows + kLbBM - 1) / kLbBM;
    // bn: column tile width chosen by the host (<= kLbBN, a multiple of 64 that splits N evenly: 384 -> 2 x 192)
    const int tiles_n = (N + bn - 1) / bn;
    // work unit of a cluster: kLbCluster consecutive row tiles x one column tile; CTA `rank` takes row tile group * C + rank
    // (past the last row tile: TMA reads zeros, the stores are clipped)
    const long long tiles = ((tiles_m + kLbCluster - 1) / kLbCluster) * tiles_n;
    const int kblocks = K / kLbBK;
    const unsigned rank = cluster_ctarank();
    const long long first = blockIdx.x / kLbCluster, stride = gridDim.x / kLbCluster;
    constexpr unsigned short kAll = (unsigned short)((1u << kLbCluster) - 1u);

    if (warp == 0) tmem_alloc(&bars->tmem_base, 512);
    if (tid == kLbEpiThreads) {
        for (int i = 0; i < kLbStages; ++i) {
            mbar_init(&bars->full[i], 1);
            mbar_init(&bars->empty[i], kLbCluster);
        }
        for (int i = 0; i < 2; ++i) { mbar_init(&bars->acc_full[i], 1); mbar_init(&bars->acc_free[i], kLbEpiThreads); }
        fence_mbar_init();
        tma_prefetch_desc(&tm_x); tma_prefetch_desc(&tm_w); tma_prefetch_desc(&tm_y);
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
            const int slice = bn / kLbCluster, slice_bytes = slice * kLbRowBytes;
            for (long long t = first; t < tiles; t += stride) {
                const int m0 = (int)((t / tiles_n) * kLbCluster + rank) * kLbBM, n0 = (int)(t % tiles_n) * bn;
                // (an L2 prefetch of the next unit's x rows was measured: 38.0 -> 40.9 us at 256 <- 256, 96 -> 120 us at
                // 256 <- 1024 -- it only competes with the loads that are needed now)
                for (int kb = 0; kb < kblocks; ++kb, ++kiter) {
                    const int s = kiter % kLbStages;
                    // both CTAs are done with the stage: my slice lands in the peer's shared memory too
                    if (kiter >= kLbStages) mbar_wait(&bars->empty[s], ((kiter / kLbStages) - 1) & 1);
                    unsigned char* st = smem + s * kLbStageBytes;
                    mbar_expect_tx(&bars->full[s], kLbABytes + bn * kLbRowBytes);   // my x tile + every CTA's W slice
                    tma_load_2d(st, &tm_x, kb * kLbBK, m0, &bars->full[s]);
                    tma_load_2d_multicast(st + kLbABytes + rank * slice_bytes, &tm_w, kb * kLbBK, n0 + (int)rank * slice,
                                          &bars->full[s], kAll);
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
            const unsigned idesc = make_idesc_bf16(kLbBM, n_cur);
            const int buf = it & 1;
            const unsigned acc = tmem + buf * kLbBN;
            if (it >= 2) mbar_wait(&bars->acc_free[buf], ((it >> 1) - 1) & 1);     // the epilogue has drained this accumulator
            for (int kb = 0; kb < kblocks; ++kb, ++kiter) {
                const int s = kiter % kLbStages;
                mbar_wait(&bars->full[s], (kiter / kLbStages) & 1);
                tcgen05_fence_after();
                if (elect_one()) {
                    unsigned char* st = smem + s * kLbStageBytes;
                    const unsigned long long d_x = lb_desc(st), d_w = lb_desc(st + kLbABytes);
#pragma unroll
                    for (int j = 0; j < kLbKSteps; ++j)                          // K = 16 per instruction: 32 bytes of the row
                        mma_bf16(acc, desc_advance(d_x, j * 32), desc_advance(d_w, j * 32), idesc, (kb | j) != 0);
                    mma_commit_multicast(&bars->empty[s], kAll);                 // stage free, here and in the peer
                    if (kb == kblocks - 1) mma_commit(&bars->acc_full[buf]);
                }
                __syncwarp();
            }
        }
    } else {
        // ======================================= epilogue warps =======================================
        // two warps per TMEM lane quadrant: warps 0-3 take the even 64-column chunks, warps 4-7 the odd ones
        const int quad = warp & 3, half = warp >> 2;
        const unsigned lane_base = (unsigned)(quad * 32) << 16;
        unsigned it = 0;
        for (long long t = first; t < tiles; t += stride, ++it) {
            const long long m0 = ((t / tiles_n) * kLbCluster + rank) * kLbBM;
            const int n0 = (int)(t % tiles_n) * bn;
            const int n_cur = min(bn, N - n0);
            const int buf = it & 1;
            // bias slice of the tile (the previous tile's readers are past their last read: barrier below)
            named_bar_sync(1, kLbEpiThreads);
            for (int i = tid; i < kLbBN; i += kLbEpiThreads)
                s_bias[i] = (bias != nullptr && i < n_cur) ? __bfloat162float(bias[n0 + i]) : 0.f;
            named_bar_sync(1, kLbEpiThreads);
            const long long row = m0 + quad * 32 + lane;
            const bool dead = row_mask != nullptr && row < rows && row_mask[row] != 0;     // padding row: zeros
            mbar_wait(&bars->acc_full[buf], (it >> 1) & 1);
            tcgen05_fence_after();
            // 32 rows x 64 columns at a time: TMEM -> registers (lane = row) -> this warp's staging tile in the SWIZZLE_128B
            // layout (16-byte chunk c of row r at chunk c ^ (r & 7)) -> one TMA store of the box
            unsigned char* stage = s_out + (half * 4 + quad) * kLbOutBytes;
            for (int c = half * 64; c < n_cur; c += 128) {
                float v[64];
                tmem_ld32(tmem + buf * kLbBN + c + lane_base, v);
                tmem_ld32(tmem + buf * kLbBN + c + 32 + lane_base, v + 32);
                if (lane == 0) tma_store_wait_read();
                __syncwarp();
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    float o[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        o[e] = v[8 * q + e] + s_bias[c + 8 * q + e];
                        if (relu) o[e] = fmaxf(o[e], 0.f);
                        if (dead) o[e] = 0.f;
                    }
                    *reinterpret_cast<uint4*>(stage + lane * 128 + ((q ^ (lane & 7)) << 4)) = pack<__nv_bfloat16>(o);
                }
                fence_proxy_async();                 // generic-proxy writes -> TMA (async proxy) read
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&tm_y, stage, n0 + c, (int)m0 + quad * 32);
                    tma_store_commit();
                }
            }
            tcgen05_fence_before();
            mbar_arrive(&bars->acc_free[buf]);
        }
        if (lane == 0) tma_store_wait_all();         // my stores have left shared memory
    }
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();                            // no CTA leaves while a peer may still write its shared memory / barriers
    if (warp == 0) tmem_free(tmem, 512);
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*LbEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static LbEncodeTiledFn lb_encode_tiled_fn()
{
    static LbEncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<LbEncodeTiledFn>(p);
    }
    return fn;
}

// row-major bf16 matrix [n_rows, n_cols]; box = 64 columns (128 bytes, SWIZZLE_128B) x box_rows rows
static bool lb_make_map(CUtensorMap* map, const void* base, unsigned long long n_rows, unsigned long long n_cols, int box_rows,
                        int box_cols = kLbBK)
{
    LbEncodeTiledFn fn = lb_encode_tiled_fn();
    if (fn == nullptr) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)n_cols, n_rows};
    const cuuint64_t strides[1] = {(cuuint64_t)n_cols * 2};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
              CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool linear_bf16_supported(int n, int k)
{
    return n >= 64 && n % 64 == 0 && k >= 64 && k % 64 == 0;
}

cudaError_t linear_bf16(const void* x, const void* w, const void* bias, const unsigned char* row_mask, long long rows, int n,
                        int k, int relu, void* y, cudaStream_t stream)
{
    if (!linear_bf16_supported(n, k) || rows < 0 || rows >= (1ll << 31) - 1024) return cudaErrorInvalidValue;
    if (rows == 0) return cudaSuccess;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(y)) % 16 != 0)
        return cudaErrorInvalidValue;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    static bool configured[64] = {};          // the attribute is per device: one process may drive several GPUs
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(linear_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kLbSmem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    // column tile width: as few tiles as 256 columns allow, of equal width (a multiple of 64; each CTA of the cluster
    // loads bn / 2 weight rows, a multiple of the 8-row swizzle group)
    const int tiles_n = (n + kLbBN - 1) / kLbBN;
    const int bn = ((n + tiles_n - 1) / tiles_n + 63) / 64 * 64;
    alignas(64) CUtensorMap tm_x, tm_w, tm_y;
    if (!lb_make_map(&tm_x, x, (unsigned long long)rows, (unsigned long long)k, kLbBM) ||
        !lb_make_map(&tm_w, w, (unsigned long long)n, (unsigned long long)k, bn / kLbCluster) ||
        !lb_make_map(&tm_y, y, (unsigned long long)rows, (unsigned long long)n, 32, 64))
        return cudaErrorNotSupported;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long tiles_m = (rows + kLbBM - 1) / kLbBM;
    const long long units = ((tiles_m + kLbCluster - 1) / kLbCluster) * ((n + bn - 1) / bn);
    const long long clusters = sms / kLbCluster;
    const int grid = (int)(units < clusters ? units : clusters) * kLbCluster;
    linear_bf16_kernel<<<grid, kLbThreads, kLbSmem, stream>>>(tm_x, tm_w, tm_y, (const __nv_bfloat16*)bias, row_mask, rows, n, k,
                                                               relu, bn);
    return cudaGetLastError();
}

}  // namespace msda
