// value_proj of MSDeformAttn on the tensor cores with the epilogue the gather wants (bf16, d_model 256, 8 heads of 32):
//
//     value_hm[n, m, s, :] = (x[n, s, :] @ W^T + b)[32 m : 32 m + 32],   zero where padding_mask[n, s]
//
// i.e. `value = self.value_proj(input_flatten)`, `value.masked_fill(input_padding_mask[..., None], 0)` and the view to
// [N, S, M, D] (reference models/ops/modules/ms_deform_attn.py:94-97) -- but written HEAD-MAJOR, the layout in which the
// x-neighbours of a bilinear footprint are adjacent (csrc/msda_forward_hm.cu: 3 instead of 4 L1 lines per sample).
//
// Same weights-stationary transposed product as proj_fused.cu (Y^T = W A^T: W lives in tensor memory for the whole
// kernel as the A operand, the 128-token tile is the B operand straight from TMA, the accumulator has a CHANNEL in each
// TMEM lane).  That orientation makes the head-major store free: an epilogue warp owns 32 consecutive channels = one
// head, so for every token its 32 lanes write one contiguous 64-byte row of value_hm.  No staging tile, no LayerNorm
// warps: 8 epilogue warps, 1 MMA issuer, 1 TMA producer.
#include <cuda.h>

#include "msda_common.cuh"
#include "msda_launch.h"
#include "umma.cuh"

namespace msda {

using namespace umma;

constexpr int kVpC = 256;             // d_model = K = N
constexpr int kVpTM = 128;            // tokens per tile
constexpr int kVpHeads = 8;
constexpr int kVpEpiWarps = 8;
constexpr int kVpEpiThreads = 32 * kVpEpiWarps;
constexpr int kVpThreads = kVpEpiThreads + 64;
constexpr int kVpSmemA = 4 * kVpTM * 128;         // 65536: 4 k-blocks of [128 rows x 128 B]
constexpr int kVpSmem = 2 * kVpSmemA + 1024;

struct VpBars {
    unsigned long long w_full, w_copied, a_full[2], mma_done[2], yacc_free;
    unsigned tmem_base;
};

__global__ void __launch_bounds__(kVpThreads, 1)
value_proj_hm_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w,
                     const __nv_bfloat16* __restrict__ bias, const unsigned char* __restrict__ mask,
                     __nv_bfloat16* __restrict__ out_hm, long long rows, int S)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* sA = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);     // [2][kVpSmemA]
    __shared__ __align__(8) VpBars bars;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long tiles = (rows + kVpTM - 1) / kVpTM;

    if (warp == 0) tmem_alloc(&bars.tmem_base, 512);
    if (tid == kVpEpiThreads) {
        mbar_init(&bars.w_full, 1);
        mbar_init(&bars.w_copied, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&bars.a_full[i], 1); mbar_init(&bars.mma_done[i], 1); }
        mbar_init(&bars.yacc_free, kVpEpiThreads);
        fence_mbar_init();
        tma_prefetch_desc(&tm_a); tma_prefetch_desc(&tm_w);
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const unsigned tmem = bars.tmem_base;
    const unsigned tmem_w = tmem;                  // columns [0, 256): W, block mb at mb * 128, K16 step s at s * 8
    const unsigned tmem_y = tmem + 256;            // columns [256, 512): Y^T block mb at mb * 128 (128 token columns)

    if (warp == kVpEpiWarps) {
        // ======================================= MMA issuer =======================================
        const unsigned idesc = make_idesc_bf16(kVpTM, kVpTM);            // M = 128 channels, N = 128 tokens
        const unsigned long long dA = make_desc_sw128(sA);
        mbar_wait(&bars.w_full, 0);
        tcgen05_fence_after();
        if (elect_one()) {
#pragma unroll
            for (int mb = 0; mb < 2; ++mb)
#pragma unroll
                for (int kb = 0; kb < 4; ++kb)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        tmem_cp_128x256b(tmem_w + mb * 128 + (kb * 4 + j) * 8,
                                         desc_advance(dA, mb * kVpSmemA + kb * kVpTM * 128 + j * 32));
            mma_commit(&bars.w_copied);
        }
        __syncwarp();
        unsigned it = 0;
        for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            mbar_wait(&bars.a_full[buf], (it >> 1) & 1);
            if (it > 0) mbar_wait(&bars.yacc_free, (it - 1) & 1);        // previous tile's epilogue has drained Y^T
            tcgen05_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int mb = 0; mb < 2; ++mb)
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb)
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            mma_bf16_ts(tmem_y + mb * 128, tmem_w + mb * 128 + (kb * 4 + j) * 8,
                                        desc_advance(dA, buf * kVpSmemA + kb * kVpTM * 128 + j * 32), idesc, (kb | j) != 0);
                mma_commit(&bars.mma_done[buf]);
            }
            __syncwarp();
        }
    } else if (warp == kVpEpiWarps + 1) {
        // ======================================= TMA producer =======================================
        if (elect_one()) {
            mbar_expect_tx(&bars.w_full, 2 * kVpSmemA);
#pragma unroll
            for (int mb = 0; mb < 2; ++mb)
#pragma unroll
                for (int kb = 0; kb < 4; ++kb)
                    tma_load_2d(sA + mb * kVpSmemA + kb * kVpTM * 128, &tm_w, kb * 64, mb * 128, &bars.w_full);
        }
        __syncwarp();
        mbar_wait(&bars.w_copied, 0);
        unsigned it = 0;
        for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            // the buffer was the B operand of tile it - 2
            if (it >= 2) mbar_wait(&bars.mma_done[buf], ((it - 2) >> 1) & 1);
            if (elect_one()) {
                mbar_expect_tx(&bars.a_full[buf], kVpSmemA);
#pragma unroll
                for (int kb = 0; kb < 4; ++kb)
                    tma_load_2d(sA + buf * kVpSmemA + kb * kVpTM * 128, &tm_a, kb * 64, (int)(tile * kVpTM), &bars.a_full[buf]);
                if (tile + gridDim.x < tiles) {
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb)
                        tma_prefetch_l2_2d(&tm_a, kb * 64, (int)((tile + gridDim.x) * kVpTM));
                }
            }
            __syncwarp();
        }
    } else {
        // ======================================= epilogue warps =======================================
        // thread = one output channel (TMEM lane); the warp = one head; per token one 64-byte row of value_hm
        const int mb = warp >> 2;
        const int head = mb * 4 + (warp & 3);
        const unsigned lane_base = (unsigned)((warp & 3) * 32) << 16;
        const float my_bias = __bfloat162float(bias[head * 32 + lane]);
        unsigned it = 0;
        for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            const long long row0 = tile * kVpTM;
            long long n = row0 / S;
            int s = (int)(row0 - n * S);
            mbar_wait(&bars.mma_done[buf], (it >> 1) & 1);
            tcgen05_fence_after();
#pragma unroll 1
            for (int tb = 0; tb < kVpTM / 32; ++tb) {
                float v[32];
                tmem_ld32(tmem_y + mb * 128 + tb * 32 + lane_base, v);
                // padding mask of the 32 tokens: one byte per lane, broadcast below
                const long long my_row = row0 + tb * 32 + lane;
                const unsigned my_masked = (mask != nullptr && my_row < rows) ? (unsigned)(mask[my_row] != 0) : 0u;
                const unsigned masked = __ballot_sync(0xffffffffu, my_masked);
#pragma unroll
                for (int t = 0; t < 32; ++t) {
                    if (row0 + tb * 32 + t < rows) {
                        const float val = ((masked >> t) & 1u) ? 0.f : v[t] + my_bias;
                        out_hm[((n * kVpHeads + head) * (long long)S + s) * 32 + lane] = __float2bfloat16_rn(val);
                    }
                    if (++s == S) { s = 0; ++n; }
                }
            }
            tcgen05_fence_before();
            mbar_arrive(&bars.yacc_free);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free(tmem, 512);
}

typedef CUresult (*VpEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static VpEncodeTiledFn vp_encode_fn()
{
    static VpEncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<VpEncodeTiledFn>(p);
    }
    return fn;
}
static bool vp_make_map(CUtensorMap* map, const void* base, unsigned long long n_rows)
{
    VpEncodeTiledFn fn = vp_encode_fn();
    if (fn == nullptr) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)kVpC, n_rows};
    const cuuint64_t strides[1] = {(cuuint64_t)kVpC * 2};
    const cuuint32_t box[2] = {64, (cuuint32_t)kVpTM};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool value_proj_hm_supported(int dtype, int d_model, int n_heads)
{
    return dtype == kBF16 && d_model == kVpC && n_heads == kVpHeads;
}

// x [rows = N * S, 256], w [256, 256] (nn.Linear layout), b [256], mask [rows] bytes or null -> out_hm [N, 8, S, 32]
cudaError_t value_proj_hm(int dtype, const void* x, const void* w, const void* b, const unsigned char* mask, void* out_hm,
                          long long rows, int S, cudaStream_t stream)
{
    if (!value_proj_hm_supported(dtype, kVpC, kVpHeads) || S <= 0 || rows < 0 || rows % S != 0 ||
        ((size_t)x % 16) != 0 || ((size_t)w % 16) != 0)
        return cudaErrorInvalidValue;
    if (rows == 0) return cudaSuccess;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    static bool configured[64] = {};
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(value_proj_hm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kVpSmem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    alignas(64) CUtensorMap tm_a, tm_w;
    if (!vp_make_map(&tm_a, x, (unsigned long long)rows) || !vp_make_map(&tm_w, w, kVpC)) return cudaErrorNotSupported;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long tiles = (rows + kVpTM - 1) / kVpTM;
    const int grid = (int)(tiles < sms ? tiles : sms);
    value_proj_hm_kernel<<<grid, kVpThreads, kVpSmem, stream>>>(
        tm_a, tm_w, (const __nv_bfloat16*)b, mask, (__nv_bfloat16*)out_hm, rows, S);
    return cudaGetLastError();
}

}  // namespace msda
