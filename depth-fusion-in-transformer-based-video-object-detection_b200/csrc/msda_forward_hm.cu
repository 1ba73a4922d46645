// Forward gather out of a HEAD-MAJOR bf16 value tensor  value_hm[N, M, S, 32]  (16-bit values, 32 channels per head).
//
// Why another layout: in the reference layout [N, S, M, D] a head's slice of a pixel is 64 bytes, and the two x-neighbours
// of a bilinear footprint lie 512 bytes apart -- every corner row is its own L1 wavefront although it fills only half of
// a 128-byte line, and that half-empty wavefront is what bounds the bf16 forward (l1tex data pipe 96 %,
// profiles/ncu_r2b.json).  Head-major, the x-neighbours (x0, x0+1) of one head are ADJACENT: 128 contiguous bytes, one
// line when x0's pixel index is even, two otherwise -- 3 lines per sample instead of 4 on average, and every line that
// is fetched is useful to the CTA (which works on one head).
//
// Mapping: 8 lanes per (query, head).  lane = half * 4 + sub: `half` selects the footprint COLUMN (x0 or x0 + 1), `sub`
// the 16-byte channel slice.  One load instruction of the group fetches both x-neighbours of a footprint row; a lane
// accumulates its column's two corners, the two halves are added by one shuffle step at the end.
// Phase 1 (as in msda_forward.cu): the lanes of a group split the L*P samples, compute each footprint once and park a
// 32-byte record per sample in shared memory -- 16 bytes per column: the two pixel indices and the two weights
// (bilinear x attention, zero for corners outside the map), so a lane reads ONE 16-byte record per sample.
#include "msda_common.cuh"
#include "msda_launch.h"

namespace msda {

constexpr int kHmD = 32;
constexpr int kHmG = 8;                 // lanes per pair
constexpr int kHmPairs = 32 / kHmG;     // pairs per warp
constexpr int kHmWarps = 8;             // 32 neighbouring queries of one head per CTA
constexpr int kHmChunk = 16;            // samples per pass

template <bool FUSED, typename RT>
__global__ void __launch_bounds__(kHmWarps * 32, 5)
msda_fwd_hm_kernel(const __nv_bfloat16* __restrict__ value_hm, const int64_t* __restrict__ shapes,
                   const int64_t* __restrict__ lsi, const SampleSrc src, __nv_bfloat16* __restrict__ out,
                   int S, int M, int L, int Lq, int P, int p_magic, long long nq_total)
{
    __shared__ int s_meta[3 * kMaxLevelsFast];
    __shared__ float s_inv[FUSED ? 2 * kMaxLevelsFast : 2];       // fused: 1 / H, 1 / W per level
    // per sample: [column 0 | column 1], each {pix row0, pix row1, weight row0, weight row1}
    __shared__ __align__(16) uint4 s_rec[kHmWarps][kHmPairs][kHmChunk + 1][2];

    if (threadIdx.x < L) {
        s_meta[3 * threadIdx.x + 0] = (int)shapes[2 * threadIdx.x];
        s_meta[3 * threadIdx.x + 1] = (int)shapes[2 * threadIdx.x + 1];
        s_meta[3 * threadIdx.x + 2] = (int)lsi[threadIdx.x];
        if constexpr (FUSED) {
            s_inv[2 * threadIdx.x + 0] = 1.f / (float)shapes[2 * threadIdx.x];
            s_inv[2 * threadIdx.x + 1] = 1.f / (float)shapes[2 * threadIdx.x + 1];
        }
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane / kHmG, sub8 = lane % kHmG;
    const int half = sub8 >> 2, sub = sub8 & 3;
    const int m = (int)(blockIdx.x % M);
    const long long nq_raw = ((long long)(blockIdx.x / M) * kHmWarps + warp) * kHmPairs + grp;
    const bool active = nq_raw < nq_total;
    const long long nq = active ? nq_raw : nq_total - 1;
    const long long pair = nq * M + m;
    const long long n = nq / Lq;
    const int LP = L * P;
    const __nv_bfloat16* vbase = value_hm + ((n * M + m) * (long long)S) * kHmD + sub * 8;
    const float* lp = nullptr;
    const float* ap = nullptr;
    const RT* op = nullptr;
    const RT* gp = nullptr;
    if constexpr (FUSED) {
        op = static_cast<const RT*>(src.loc) + nq * src.loc_stride + (long long)m * LP * 2;
        gp = static_cast<const RT*>(src.attn) + nq * src.attn_stride + (long long)m * LP;
    } else {
        lp = static_cast<const float*>(src.loc) + pair * LP * 2;
        ap = static_cast<const float*>(src.attn) + pair * LP;
    }

    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = 0.f;

    constexpr int K = kHmChunk / kHmG;          // samples a lane prepares per pass: j = sub8 + k * 8
    // fused: softmax over the pair's L*P (<= 16) logits
    float prob[K];
    float inv_sum = 1.f;
    if constexpr (FUSED) {
        float mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int s = sub8 + k * kHmG;
            prob[k] = s < LP ? load_raw1<RT>(gp + s) : -INFINITY;
            mx = fmaxf(mx, prob[k]);
        }
        mx = group_max<kHmG>(mx);
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            prob[k] = prob[k] == -INFINITY ? 0.f : expf(prob[k] - mx);
            sum += prob[k];
        }
        inv_sum = 1.f / group_sum<kHmG>(sum);
    }
    const float half_inv_p = 0.5f / (float)P;

    for (int s0 = 0; s0 < LP; s0 += kHmChunk) {
        const int cnt = min(kHmChunk, LP - s0);
        const int cnt2 = (cnt + 1) & ~1;
        // ---- phase 1: footprints ----
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int j = sub8 + k * kHmG;
            if (j >= cnt2) break;
            uint4 c0 = make_uint4(0u, 0u, 0u, 0u), c1 = make_uint4(0u, 0u, 0u, 0u);
            if (j < cnt) {
                const int s = s0 + j;
                const int l = div_by_points(s, p_magic);
                float2 xy;
                float a;
                if constexpr (FUSED) {
                    xy = fused_location(load_raw2<RT>(op + 2 * s), src.ref + (nq * L + l) * src.ref_dim, src.ref_dim,
                                        s_inv[2 * l], s_inv[2 * l + 1], half_inv_p);
                    a = prob[k] * inv_sum;
                } else {
                    xy = ldg_stream_f32x2(lp + 2 * s);
                    a = ldg_stream_f32(ap + s);
                }
                const Footprint f = footprint<float>(xy.x, xy.y, s_meta[3 * l], s_meta[3 * l + 1], s_meta[3 * l + 2]);
                const float hw = 1.f - f.lw, hh = 1.f - f.lh;
                // invalid corners: weight 0 and a harmless in-range pixel (0)
                c0.x = (f.ok & 1u) ? (unsigned)f.pix00 : 0u;
                c0.y = (f.ok & 4u) ? (unsigned)(f.pix00 + f.rowstep) : 0u;
                c0.z = __float_as_uint((f.ok & 1u) ? hh * hw * a : 0.f);
                c0.w = __float_as_uint((f.ok & 4u) ? f.lh * hw * a : 0.f);
                c1.x = (f.ok & 2u) ? (unsigned)(f.pix00 + 1) : 0u;
                c1.y = (f.ok & 8u) ? (unsigned)(f.pix00 + f.rowstep + 1) : 0u;
                c1.z = __float_as_uint((f.ok & 2u) ? hh * f.lw * a : 0.f);
                c1.w = __float_as_uint((f.ok & 8u) ? f.lh * f.lw * a : 0.f);
            }
            s_rec[warp][grp][j][0] = c0;
            s_rec[warp][grp][j][1] = c1;
        }
        __syncwarp();
        // ---- phase 2: gather, 2 samples x 2 rows in flight per lane ----
        for (int j0 = 0; j0 < cnt2; j0 += 2) {
            uint4 raw[2][2];
            uint4 rec[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                rec[u] = s_rec[warp][grp][j0 + u][half];
                raw[u][0] = ldg_v4(vbase + (long long)rec[u].x * kHmD);
                raw[u][1] = ldg_v4(vbase + (long long)rec[u].y * kHmD);
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const float w0 = __uint_as_float(rec[u].z), w1 = __uint_as_float(rec[u].w);
                float v0[8], v1[8];
                unpack<__nv_bfloat16>(raw[u][0], v0);
                unpack<__nv_bfloat16>(raw[u][1], v1);
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[c] = fmaf(w1, v1[c], fmaf(w0, v0[c], acc[c]));
            }
        }
        __syncwarp();
    }
    // the two columns of the footprints
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 4);
    if (active && half == 0) stg_stream_v4(out + pair * kHmD + sub * 8, pack<__nv_bfloat16>(acc));
}

bool forward_hm_supported(int dtype, int D, int L, int P)
{
    return dtype == kBF16 && D == kHmD && L >= 1 && L <= kMaxLevelsFast && P >= 1 && P <= 64 &&
           (long long)L * P * P < 65536;
}

cudaError_t forward_hm(const FwdArgs& a, cudaStream_t stream)
{
    if (!forward_hm_supported(a.dtype, a.D, a.L, a.P) || (long long)a.S * a.D >= (1ll << 31)) return cudaErrorInvalidValue;
    const long long nq_total = (long long)a.N * a.Lq;
    if (nq_total * a.M == 0) return cudaSuccess;
    const long long blocks = ((nq_total + kHmWarps * kHmPairs - 1) / (kHmWarps * kHmPairs)) * a.M;
    if (blocks > 0x7fffffffll) return cudaErrorInvalidConfiguration;
    SampleSrc src;
    src.loc = a.loc; src.attn = a.attn; src.ref = nullptr; src.loc_stride = 0; src.attn_stride = 0; src.ref_dim = 0;
    msda_fwd_hm_kernel<false, float><<<(unsigned)blocks, kHmWarps * 32, 0, stream>>>(
        (const __nv_bfloat16*)a.value, a.shapes, a.lsi, src, (__nv_bfloat16*)a.out, a.S, a.M, a.L, a.Lq, a.P,
        (65536 + a.P - 1) / a.P, nq_total);
    return cudaGetLastError();
}

cudaError_t fused_forward_hm(const FusedArgs& a, cudaStream_t stream)
{
    if (!fused_supported(a) || !fused_raw_layout_ok(a) || a.dtype != kBF16 || a.D != kHmD || a.value_ld != 0 ||
        (long long)a.S * a.D >= (1ll << 31))
        return cudaErrorInvalidValue;
    const long long nq_total = (long long)a.N * a.Lq;
    if (nq_total * a.M == 0) return cudaSuccess;
    const long long blocks = ((nq_total + kHmWarps * kHmPairs - 1) / (kHmWarps * kHmPairs)) * a.M;
    if (blocks > 0x7fffffffll) return cudaErrorInvalidConfiguration;
    SampleSrc src;
    src.loc = a.offsets; src.attn = a.logits; src.ref = a.ref;
    src.loc_stride = a.off_stride; src.attn_stride = a.logit_stride; src.ref_dim = a.ref_dim;
    const int p_magic = (65536 + a.P - 1) / a.P;
    if (a.raw_dtype == kF32)
        msda_fwd_hm_kernel<true, float><<<(unsigned)blocks, kHmWarps * 32, 0, stream>>>(
            (const __nv_bfloat16*)a.value, a.shapes, a.lsi, src, (__nv_bfloat16*)a.out, a.S, a.M, a.L, a.Lq, a.P, p_magic, nq_total);
    else
        msda_fwd_hm_kernel<true, __nv_bfloat16><<<(unsigned)blocks, kHmWarps * 32, 0, stream>>>(
            (const __nv_bfloat16*)a.value, a.shapes, a.lsi, src, (__nv_bfloat16*)a.out, a.S, a.M, a.L, a.Lq, a.P, p_magic, nq_total);
    return cudaGetLastError();
}

}  // namespace msda
