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
// One CTA per SM, persistent over (row tile, column tile) pairs, 128 x 256 output tile, K in blocks of 32 fp32 (one
// 128-byte SWIZZLE_128B row).  Warp roles (10 warps):
//   0-3  epilogue   tcgen05.ld of the finished accumulator (lane = output row), + bias, ReLU, fp32 stores
//   4-7  converter  x tile in shared memory -> x_lo (second tile, same swizzled layout: the split is elementwise, so it
//                   never has to know the layout), then fence.proxy.async; x_hi is the x tile as the tensor core reads it
//   8    MMA issuer 4 K-steps x 3 terms of tcgen05.mma kind::tf32 per K block; TWO accumulators in TMEM (see there)
//   9    TMA producer  x tile, W_hi tile, W_lo tile per K block, two stages of 96 KB
#include <cuda.h>

#include "msda_common.cuh"
#include "msda_launch.h"
#include "umma.cuh"

namespace msda {

using namespace umma;

constexpr int kLtBM = 128;                 // output rows per tile (TMEM lanes)
constexpr int kLtBN = 256;                 // output columns per tile (TMEM columns of one accumulator)
constexpr int kLtBK = 32;                  // fp32 per K block = 128 bytes = one swizzle row
constexpr int kLtStages = 2;
constexpr int kLtABytes = kLtBM * 128;     // 16 KB
constexpr int kLtBBytes = kLtBN * 128;     // 32 KB
constexpr int kLtStageBytes = 2 * kLtABytes + 2 * kLtBBytes;      // x_hi | x_lo | W_hi | W_lo = 96 KB
constexpr int kLtEpiThreads = 128, kLtConvThreads = 128;
constexpr int kLtThreads = kLtEpiThreads + kLtConvThreads + 64;
constexpr int kLtSmem = kLtStages * kLtStageBytes + kLtBN * 4 + 256;
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
__device__ __forceinline__ void mma_tf32(unsigned tmem_d, unsigned long long desc_a, unsigned long long desc_b,
                                         unsigned idesc, bool accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"((unsigned)accumulate) : "memory");
}

__global__ void __launch_bounds__(kLtThreads, 1)
linear_tf32x3_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_wh,
                     const __grid_constant__ CUtensorMap tm_wl, const float* __restrict__ bias, float* __restrict__ y,
                     long long rows, int N, int K, int relu)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    float* s_bias = reinterpret_cast<float*>(smem + kLtStages * kLtStageBytes);
    LtBars* bars = reinterpret_cast<LtBars*>(smem + kLtStages * kLtStageBytes + kLtBN * 4);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long tiles_m = (rows + kLtBM - 1) / kLtBM;
    const int tiles_n = (N + kLtBN - 1) / kLtBN;
    const long long tiles = tiles_m * tiles_n;
    const int kblocks = K / kLtBK;

    if (warp == 0) tmem_alloc(&bars->tmem_base, 512);
    if (tid == kLtEpiThreads) {
        for (int i = 0; i < kLtStages; ++i) {
            mbar_init(&bars->full[i], 1);
            mbar_init(&bars->conv[i], kLtConvThreads);
            mbar_init(&bars->empty[i], 1);
        }
        mbar_init(&bars->acc_full, 1);
        mbar_init(&bars->acc_free, kLtEpiThreads);
        fence_mbar_init();
        tma_prefetch_desc(&tm_x); tma_prefetch_desc(&tm_wh); tma_prefetch_desc(&tm_wl);
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const unsigned tmem = bars->tmem_base;

    if (warp == 9) {
        // ======================================= TMA producer =======================================
        if (elect_one()) {
            unsigned kiter = 0;
            for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
                const int m0 = (int)(t / tiles_n) * kLtBM, n0 = (int)(t % tiles_n) * kLtBN;
                // x is the only HBM stream (W stays in L2): bring the NEXT tile's rows as far as L2 while this tile computes
                const long long tn = t + gridDim.x;
                if (tn < tiles && tn / tiles_n != t / tiles_n) {
                    for (int kb = 0; kb < kblocks; ++kb) tma_prefetch_l2_2d(&tm_x, kb * kLtBK, (int)(tn / tiles_n) * kLtBM);
                }
                for (int kb = 0; kb < kblocks; ++kb, ++kiter) {
                    const int s = kiter % kLtStages;
                    if (kiter >= kLtStages) mbar_wait(&bars->empty[s], ((kiter / kLtStages) - 1) & 1);
                    unsigned char* st = smem + s * kLtStageBytes;
                    mbar_expect_tx(&bars->full[s], kLtABytes + 2 * kLtBBytes);
                    tma_load_2d(st, &tm_x, kb * kLtBK, m0, &bars->full[s]);
                    tma_load_2d(st + 2 * kLtABytes, &tm_wh, kb * kLtBK, n0, &bars->full[s]);
                    tma_load_2d(st + 2 * kLtABytes + kLtBBytes, &tm_wl, kb * kLtBK, n0, &bars->full[s]);
                }
            }
        }
        __syncwarp();
    } else if (warp == 8) {
        // ======================================= MMA issuer =======================================
        unsigned kiter = 0, it = 0;
        for (long long t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
            const int n0 = (int)(t % tiles_n) * kLtBN;
            const int n_cur = min(kLtBN, N - n0);
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
                const unsigned long long d_xh = make_desc_sw128(st), d_xl = make_desc_sw128(st + kLtABytes);
                const unsigned long long d_wh = make_desc_sw128(st + 2 * kLtABytes),
                                         d_wl = make_desc_sw128(st + 2 * kLtABytes + kLtBBytes);
                mbar_wait(&bars->full[s], (kiter / kLtStages) & 1);              // x and W tiles landed
                tcgen05_fence_after();
                // the two terms with x_hi need nothing from the converter: the tensor core reads the top 19 bits of the
                // fp32 words TMA wrote (x_hi = x truncated to TF32), so they go first and cover the converter's latency
                if (elect_one()) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {                                // K = 8 per instruction: 32 bytes of the row
                        mma_tf32(acc_small, desc_advance(d_xh, j * 32), desc_advance(d_wl, j * 32), idesc, (kb | j) != 0);
                        mma_tf32(acc_big, desc_advance(d_xh, j * 32), desc_advance(d_wh, j * 32), idesc, (kb | j) != 0);
                    }
                }
                __syncwarp();
                mbar_wait(&bars->conv[s], (kiter / kLtStages) & 1);              // x_lo written
                tcgen05_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        mma_tf32(acc_small, desc_advance(d_xl, j * 32), desc_advance(d_wh, j * 32), idesc, true);
                    mma_commit(&bars->empty[s]);                                 // stage free once these MMAs have read it
                    if (kb == kblocks - 1) mma_commit(&bars->acc_full);
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // ======================================= converter warps =======================================
        const int ct = tid - kLtEpiThreads;
        unsigned kiter = 0;
        for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
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
        const unsigned lane_base = (unsigned)(warp * 32) << 16;
        unsigned it = 0;
        for (long long t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
            const long long m0 = (t / tiles_n) * kLtBM;
            const int n0 = (int)(t % tiles_n) * kLtBN;
            const int n_cur = min(kLtBN, N - n0);
            // bias slice of the tile (the previous tile's readers are past their last read: barrier below)
            named_bar_sync(1, kLtEpiThreads);
            for (int i = tid; i < kLtBN; i += kLtEpiThreads) s_bias[i] = (bias != nullptr && i < n_cur) ? bias[n0 + i] : 0.f;
            named_bar_sync(1, kLtEpiThreads);
            mbar_wait(&bars->acc_full, it & 1);
            tcgen05_fence_after();
            const long long row = m0 + warp * 32 + lane;
            float* yr = y + row * N + n0;
            for (int c = 0; c < n_cur; c += 32) {
                float v[32], w[32];
                tmem_ld32(tmem + c + lane_base, v);
                tmem_ld32(tmem + kLtBN + c + lane_base, w);
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] += w[i];
                if (row < rows) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 b4 = *reinterpret_cast<const float4*>(s_bias + c + 4 * q);
                        float4 o = make_float4(v[4 * q] + b4.x, v[4 * q + 1] + b4.y, v[4 * q + 2] + b4.z, v[4 * q + 3] + b4.w);
                        if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                        *reinterpret_cast<float4*>(yr + c + 4 * q) = o;
                    }
                }
            }
            tcgen05_fence_before();
            mbar_arrive(&bars->acc_free);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
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
static bool lt_make_map(CUtensorMap* map, const void* base, unsigned long long n_rows, unsigned long long n_cols, int box_rows)
{
    LtEncodeTiledFn fn = lt_encode_tiled_fn();
    if (fn == nullptr) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)n_cols, n_rows};
    const cuuint64_t strides[1] = {(cuuint64_t)n_cols * 4};
    const cuuint32_t box[2] = {(cuuint32_t)kLtBK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool linear_tf32x3_supported(int n, int k)
{
    return n >= 32 && n % 32 == 0 && k >= kLtBK && k % kLtBK == 0;
}

cudaError_t linear_tf32x3(const float* x, const float* w_hi, const float* w_lo, const float* bias, long long rows, int n,
                          int k, int relu, float* y, cudaStream_t stream)
{
    if (!linear_tf32x3_supported(n, k) || rows < 0 || rows >= (1ll << 31)) return cudaErrorInvalidValue;
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
    alignas(64) CUtensorMap tm_x, tm_wh, tm_wl;
    if (!lt_make_map(&tm_x, x, (unsigned long long)rows, (unsigned long long)k, kLtBM) ||
        !lt_make_map(&tm_wh, w_hi, (unsigned long long)n, (unsigned long long)k, kLtBN) ||
        !lt_make_map(&tm_wl, w_lo, (unsigned long long)n, (unsigned long long)k, kLtBN))
        return cudaErrorNotSupported;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long tiles = ((rows + kLtBM - 1) / kLtBM) * ((n + kLtBN - 1) / kLtBN);
    const int grid = (int)(tiles < sms ? tiles : sms);
    linear_tf32x3_kernel<<<grid, kLtThreads, kLtSmem, stream>>>(tm_x, tm_wh, tm_wl, bias, y, rows, n, k, relu);
    return cudaGetLastError();
}

}  // namespace msda
