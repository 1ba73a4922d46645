// Forward multi-scale deformable attention for sm_100a.
//
//   out[n,q,m,:] = sum_{l,p} A[n,q,m,l,p] * bilinear(value_l[n,:,m,:], loc[n,q,m,l,p])
//
// Replaces the reference's ms_deformable_im2col_gpu_kernel (cuda/ms_deform_im2col_cuda.cuh:237-299,
// one thread per output scalar, 4-byte loads, per-thread address arithmetic).
//
// Fast kernel (D*sizeof(VT) a power-of-two multiple of 16 B):
//   * a group of G = D*sizeof(VT)/16 lanes owns one pair (n,q,m); a warp owns 32/G consecutive
//     pairs, so loc / attn / out accesses of a warp are one contiguous block;
//   * phase 1 (setup): the lanes of a group split the L*P samples between them; each computes
//     one sample's bilinear footprint ONCE (4 clamped pixel indices + 4 weights already
//     multiplied by the attention weight and zeroed for out-of-map corners) and parks it in
//     shared memory - 32 B per sample.  No per-channel address arithmetic is left;
//   * phase 2 (gather): every lane walks the samples, reads the 32-byte record with two
//     broadcast LDS.128 and issues four independent 128-bit loads (one per corner) of its
//     16-byte slice of the head's channel vector; 4 samples are unrolled so 16 loads are in
//     flight per lane.  Accumulation is fp32 in registers; one 128-bit store per lane.
// Generic kernel: any D, any dtype (fp64 for gradcheck), one thread per output scalar.
#include "msda_common.cuh"
#include "msda_launch.h"

namespace msda {

template <int PAIRS> struct FwdWarps { static constexpr int value = PAIRS >= 16 ? 2 : (PAIRS >= 8 ? 4 : 8); };

// ptxas only keeps the 16 gather loads of a 4-sample batch in flight together when it is told the
// occupancy target (otherwise it minimises registers and sinks each load next to its FMAs):
// measured best on B200 (tools/ab_variants.sh): 40 resident warps per SM (5 CTAs x 8 warps for fp32
// D=32, 10 CTAs x 4 warps for bf16 D=32), i.e. <= 48 registers.
#ifndef MSDA_FWD_MINWARPS
#define MSDA_FWD_MINWARPS 40
#endif
// CTA -> work mapping.  1: a CTA owns WARPS*PAIRS consecutive queries of ONE head (neighbouring
// queries of a head sample overlapping pixels -> L1 reuse); 0: consecutive pairs (all heads of a
// few queries).
#ifndef MSDA_CTA_PER_HEAD
#define MSDA_CTA_PER_HEAD 1
#endif

// FUSED = false: the drop-in op (sampling_loc / attn_weight materialised, fp32).
// FUSED = true : the layer kernel - reads the raw sampling_offsets / attention_weights projection
//                outputs (type RT) and the reference points; softmax over the L*P logits of the
//                pair (group shuffles) and the offset -> location arithmetic happen in phase 1.
template <typename VT, int D, bool FUSED, typename RT>
__global__ void __launch_bounds__(FwdWarps<32 / (D / Traits<VT>::kEpl)>::value * 32,
                                  MSDA_FWD_MINWARPS / FwdWarps<32 / (D / Traits<VT>::kEpl)>::value)
msda_fwd_fast_kernel(const VT* __restrict__ value, const int64_t* __restrict__ shapes,
                     const int64_t* __restrict__ lsi, const SampleSrc src, VT* __restrict__ out,
                     int S, int M, int L, int Lq, int P, int p_magic, long long total_pairs, int value_ld)
{
    constexpr int EPL = Traits<VT>::kEpl;
    constexpr int G = D / EPL;
    constexpr int PAIRS = 32 / G;
    constexpr int WARPS = FwdWarps<PAIRS>::value;
    static_assert(G >= 1 && G <= 32 && (G & (G - 1)) == 0, "head width must map to a power-of-two lane group");
    // samples per pass.  Plain op: 8 instead of 16 halves the record shared memory, which leaves more of the
    // SM's 256 KB to the L1 cache the gather lives on (bf16: 0.484 -> 0.460 ms, fp32: 0.641 -> 0.633 ms).
    // Fused op: one pass of 16 -- two passes plus the softmax numerators kept across them spill at the
    // 48-register budget and measured slower (bf16 encoder 4.84 -> 5.03 ms).
#ifndef MSDA_FWD_CHUNK
#define MSDA_FWD_CHUNK 8
#endif
#ifndef MSDA_FWD_CHUNK_FUSED
#define MSDA_FWD_CHUNK_FUSED 16
#endif
    constexpr int kChunkWanted = FUSED ? MSDA_FWD_CHUNK_FUSED : MSDA_FWD_CHUNK;
    constexpr int kChunk = kChunkWanted > G ? kChunkWanted : G;
    constexpr int K = kChunk / G;                    // samples per lane per pass: j = sub + k*G
    constexpr int FCH = (msda::kChunk + kChunk - 1) / kChunk;   // passes of the fused op (L*P <= msda::kChunk)

    __shared__ int s_meta[3 * kMaxLevelsFast];
    __shared__ float s_inv[FUSED ? 2 * kMaxLevelsFast : 2];       // fused: 1 / H, 1 / W per level
    __shared__ __align__(16) int4   s_pix[WARPS][PAIRS][kChunk + 1];   // +1 record: group stride 272 B, so the groups of a warp hit distinct banks
    __shared__ __align__(16) float4 s_wgt[WARPS][PAIRS][kChunk + 1];

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
    const int grp = lane / G, sub = lane % G;
#if MSDA_CTA_PER_HEAD
    const int m = (int)(blockIdx.x % M);
    const long long nq_total = total_pairs / M;
    const long long nq_raw = ((long long)(blockIdx.x / M) * WARPS + warp) * PAIRS + grp;
    const bool active = nq_raw < nq_total;
    const long long nq = active ? nq_raw : nq_total - 1;          // clamp: loads stay in bounds
    const long long pair = nq * M + m;
#else
    const long long pair_raw = ((long long)blockIdx.x * WARPS + warp) * PAIRS + grp;
    const bool active = pair_raw < total_pairs;
    const long long pair = active ? pair_raw : total_pairs - 1;   // clamp: loads stay in bounds
    const int m = (int)(pair % M);
    const long long nq = pair / M;
#endif
    const long long n = nq / Lq;
    const int LP = L * P;
    // value_ld: elements between consecutive pixels of `value` (M*D when dense; larger when the caller hands in a
    // column slice of a wider projection output, e.g. the decoder's six value projections computed as one GEMM)
    const int MD = value_ld;
    const VT* vbase = value + n * S * (long long)value_ld + (long long)m * D + sub * EPL;
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

    float acc[EPL];
#pragma unroll
    for (int c = 0; c < EPL; ++c) acc[c] = 0.f;

    // fused: softmax over the pair's L*P logits up front; this lane keeps the numerators of its own samples
    float prob[FCH * K];
    float inv_sum = 1.f;
    if constexpr (FUSED) {
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < FCH; ++c)
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int s = c * kChunk + sub + k * G;
                prob[c * K + k] = s < LP ? load_raw1<RT>(gp + s) : -INFINITY;
                mx = fmaxf(mx, prob[c * K + k]);
            }
        mx = group_max<G>(mx);
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < FCH * K; ++i) {
            prob[i] = prob[i] == -INFINITY ? 0.f : expf(prob[i] - mx);
            sum += prob[i];
        }
        inv_sum = 1.f / group_sum<G>(sum);            // one division per lane; the samples multiply
    }
    const float half_inv_p = 0.5f / (float)P;

#pragma unroll
    for (int c = 0; c < (FUSED ? FCH : 1 << 30); ++c) {
        const int s0 = c * kChunk;
        if (s0 >= LP) break;
        const int cnt = min(kChunk, LP - s0);
        const int cnt4 = (cnt + 3) & ~3;
        // ---- phase 1: footprints, one sample per lane of the group -------------------------
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int j = sub + k * G;
            if (j >= cnt4) break;
            int4 px = make_int4(0, 0, 0, 0);
            float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j < cnt) {
                const int s = s0 + j;
                const int l = div_by_points(s, p_magic);
                float2 xy;
                float a;
                if constexpr (FUSED) {
                    xy = fused_location(load_raw2<RT>(op + 2 * s), src.ref + (nq * L + l) * src.ref_dim, src.ref_dim,
                                        s_inv[2 * l], s_inv[2 * l + 1], half_inv_p);
                    a = prob[c * K + k] * inv_sum;       // softmax: exp(x - max) / sum
                } else {
                    xy = ldg_stream_f32x2(lp + 2 * s);
                    a = ldg_stream_f32(ap + s);
                }
                const Footprint f = footprint<float>(xy.x, xy.y, s_meta[3 * l], s_meta[3 * l + 1], s_meta[3 * l + 2]);
                const float hw = 1.f - f.lw, hh = 1.f - f.lh;
                // invalid corners: weight 0 and a harmless in-range address (pixel 0)
                px.x = (f.ok & 1u) ? f.pix00 : 0;
                px.y = (f.ok & 2u) ? f.pix00 + 1 : 0;
                px.z = (f.ok & 4u) ? f.pix00 + f.rowstep : 0;
                px.w = (f.ok & 8u) ? f.pix00 + f.rowstep + 1 : 0;
                w.x = (f.ok & 1u) ? hh * hw * a : 0.f;
                w.y = (f.ok & 2u) ? hh * f.lw * a : 0.f;
                w.z = (f.ok & 4u) ? f.lh * hw * a : 0.f;
                w.w = (f.ok & 8u) ? f.lh * f.lw * a : 0.f;
            }
            s_pix[warp][grp][j] = px;
            s_wgt[warp][grp][j] = w;
        }
        __syncwarp();
        // ---- phase 2: gather, 4 samples x 4 corners in flight ------------------------------
        for (int j0 = 0; j0 < cnt4; j0 += 4) {
            uint4 raw[4][4];
            float4 w[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int4 px = s_pix[warp][grp][j0 + u];
                w[u] = s_wgt[warp][grp][j0 + u];
                raw[u][0] = ldg_v4(vbase + (long long)px.x * MD);
                raw[u][1] = ldg_v4(vbase + (long long)px.y * MD);
                raw[u][2] = ldg_v4(vbase + (long long)px.z * MD);
                raw[u][3] = ldg_v4(vbase + (long long)px.w * MD);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float wk[4] = {w[u].x, w[u].y, w[u].z, w[u].w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float v[EPL];
                    unpack<VT>(raw[u][k], v);
#pragma unroll
                    for (int c = 0; c < EPL; ++c) acc[c] = fmaf(wk[k], v[c], acc[c]);
                }
            }
        }
        __syncwarp();
    }
    if (active) stg_stream_v4(out + pair * D + sub * EPL, pack<VT>(acc));
}

// ------------------------------------------------------------------------------------------
// Generic path: any channel count, any dtype.  One thread per output scalar; metadata from
// global memory (no level limit).  Accumulates in acc_t (fp32 for 16-bit storage).
// ------------------------------------------------------------------------------------------
template <typename VT> struct Scalar {
    using acc_t = typename Traits<VT>::acc_t;
    static __device__ __forceinline__ acc_t load(const VT* p) { return (acc_t)to_f32<VT>(*p); }
    static __device__ __forceinline__ VT store(acc_t v) { return from_f32<VT>((float)v); }
};
template <> struct Scalar<double> {
    using acc_t = double;
    static __device__ __forceinline__ double load(const double* p) { return *p; }
    static __device__ __forceinline__ double store(double v) { return v; }
};

template <typename VT>
__global__ void __launch_bounds__(256)
msda_fwd_generic_kernel(const VT* __restrict__ value, const int64_t* __restrict__ shapes,
                        const int64_t* __restrict__ lsi, const typename Traits<VT>::loc_t* __restrict__ loc,
                        const typename Traits<VT>::loc_t* __restrict__ attn, VT* __restrict__ out,
                        int S, int M, int D, int L, int Lq, int P, long long total)
{
    using acc_t = typename Traits<VT>::acc_t;
    using loc_t = typename Traits<VT>::loc_t;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % D);
        const long long pair = idx / D;
        const int m = (int)(pair % M);
        const long long n = (pair / M) / Lq;
        const VT* vbase = value + (n * S * M + m) * (long long)D + c;
        const loc_t* lp = loc + pair * L * P * 2;
        const loc_t* ap = attn + pair * L * P;
        const long long MD = (long long)M * D;
        acc_t acc = 0;
        for (int l = 0; l < L; ++l) {
            const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1], start = (int)lsi[l];
            for (int p = 0; p < P; ++p) {
                const loc_t x = lp[(l * P + p) * 2], y = lp[(l * P + p) * 2 + 1];
                const acc_t a = (acc_t)ap[l * P + p];
                const Footprint f = footprint<loc_t>(x, y, H, W, start);
                if (!f.ok) continue;
                // recompute the fractions in acc_t so that fp64 keeps fp64 weights
                const loc_t w_im = x * (loc_t)W - (loc_t)0.5, h_im = y * (loc_t)H - (loc_t)0.5;
                const acc_t lw = (acc_t)(w_im - floor(w_im)), lh = (acc_t)(h_im - floor(h_im));
                const acc_t hw = 1 - lw, hh = 1 - lh;
                const VT* p00 = vbase + (long long)f.pix00 * MD;
                acc_t v1 = 0, v2 = 0, v3 = 0, v4 = 0;
                if (f.ok & 1u) v1 = Scalar<VT>::load(p00);
                if (f.ok & 2u) v2 = Scalar<VT>::load(p00 + MD);
                if (f.ok & 4u) v3 = Scalar<VT>::load(p00 + (long long)f.rowstep * MD);
                if (f.ok & 8u) v4 = Scalar<VT>::load(p00 + (long long)f.rowstep * MD + MD);
                acc += (hh * hw * v1 + hh * lw * v2 + lh * hw * v3 + lh * lw * v4) * a;
            }
        }
        out[idx] = Scalar<VT>::store(acc);
    }
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
template <typename VT, int D>
static cudaError_t launch_fwd_fast(const FwdArgs& a, cudaStream_t stream)
{
    constexpr int G = D / Traits<VT>::kEpl;
    constexpr int PAIRS = 32 / G;
    constexpr int WARPS = FwdWarps<PAIRS>::value;
    const long long total_pairs = (long long)a.N * a.Lq * a.M;
#if MSDA_CTA_PER_HEAD
    const long long nq_total = (long long)a.N * a.Lq;
    const long long blocks = ((nq_total + WARPS * PAIRS - 1) / (WARPS * PAIRS)) * a.M;
#else
    const long long blocks = (total_pairs + WARPS * PAIRS - 1) / (WARPS * PAIRS);
#endif
    if (blocks > 0x7fffffffll) return cudaErrorInvalidConfiguration;
    const int p_magic = (65536 + a.P - 1) / a.P;
    SampleSrc src;
    src.loc = a.loc; src.attn = a.attn; src.ref = nullptr; src.loc_stride = 0; src.attn_stride = 0; src.ref_dim = 0;
    msda_fwd_fast_kernel<VT, D, false, float><<<(unsigned)blocks, WARPS * 32, 0, stream>>>(
        (const VT*)a.value, a.shapes, a.lsi, src, (VT*)a.out, a.S, a.M, a.L, a.Lq, a.P, p_magic, total_pairs, a.M * D);
    return cudaGetLastError();
}

template <typename VT, int D, typename RT>
static cudaError_t launch_fwd_fused(const FusedArgs& a, cudaStream_t stream)
{
    constexpr int G = D / Traits<VT>::kEpl;
    constexpr int PAIRS = 32 / G;
    constexpr int WARPS = FwdWarps<PAIRS>::value;
    const long long total_pairs = (long long)a.N * a.Lq * a.M;
#if MSDA_CTA_PER_HEAD
    const long long nq_total = (long long)a.N * a.Lq;
    const long long blocks = ((nq_total + WARPS * PAIRS - 1) / (WARPS * PAIRS)) * a.M;
#else
    const long long blocks = (total_pairs + WARPS * PAIRS - 1) / (WARPS * PAIRS);
#endif
    if (blocks > 0x7fffffffll) return cudaErrorInvalidConfiguration;
    const int p_magic = (65536 + a.P - 1) / a.P;
    SampleSrc src;
    src.loc = a.offsets; src.attn = a.logits; src.ref = a.ref;
    src.loc_stride = a.off_stride; src.attn_stride = a.logit_stride; src.ref_dim = a.ref_dim;
    msda_fwd_fast_kernel<VT, D, true, RT><<<(unsigned)blocks, WARPS * 32, 0, stream>>>(
        (const VT*)a.value, a.shapes, a.lsi, src, (VT*)a.out, a.S, a.M, a.L, a.Lq, a.P, p_magic, total_pairs,
        a.value_ld > 0 ? (int)a.value_ld : a.M * D);
    return cudaGetLastError();
}

template <typename VT>
static cudaError_t launch_fwd_generic(const FwdArgs& a, cudaStream_t stream)
{
    using loc_t = typename Traits<VT>::loc_t;
    const long long total = (long long)a.N * a.Lq * a.M * a.D;
    long long blocks = (total + 255) / 256;
    if (blocks > (1ll << 30)) blocks = 1ll << 30;
    msda_fwd_generic_kernel<VT><<<(unsigned)blocks, 256, 0, stream>>>(
        (const VT*)a.value, a.shapes, a.lsi, (const loc_t*)a.loc, (const loc_t*)a.attn, (VT*)a.out,
        a.S, a.M, a.D, a.L, a.Lq, a.P, total);
    return cudaGetLastError();
}

static bool fast_shape_ok(const FwdArgs& a)
{
    return !a.force_generic && a.L <= kMaxLevelsFast && a.P <= 64 && (long long)a.L * a.P * a.P < 65536 &&
           (long long)a.S * a.M * a.D < (1ll << 31);
}

template <typename VT>
static cudaError_t dispatch_fwd_16or32(const FwdArgs& a, cudaStream_t stream)
{
    if (fast_shape_ok(a)) {
        switch (a.D) {
            case 8:   return launch_fwd_fast<VT, 8>(a, stream);
            case 16:  return launch_fwd_fast<VT, 16>(a, stream);
            case 32:  return launch_fwd_fast<VT, 32>(a, stream);
            case 64:  return launch_fwd_fast<VT, 64>(a, stream);
            case 128: return launch_fwd_fast<VT, 128>(a, stream);
            default: break;
        }
    }
    return launch_fwd_generic<VT>(a, stream);
}

bool fused_supported(const FusedArgs& a)
{
    const bool dtype_ok = (a.dtype == kF32 && a.raw_dtype == kF32) ||
                          (a.dtype == kBF16 && (a.raw_dtype == kF32 || a.raw_dtype == kBF16));
    return dtype_ok && (a.D == 16 || a.D == 32 || a.D == 64) && a.L >= 1 && a.P >= 1 && a.L * a.P <= kChunk &&
           a.L <= kMaxLevelsFast && (a.ref_dim == 2 || a.ref_dim == 4) &&
           (long long)a.S * a.M * a.D < (1ll << 31) && (a.P % 2 == 0 || a.raw_dtype == kF32) &&
           // offset pairs are moved with one 8-byte (fp32) / 4-byte (16-bit) access: every query row of the raw
           // projection output must start on such a boundary.  The module's row is [offsets (2 MLP) | logits (MLP)],
           // 3*M*L*P elements, so an odd M*L*P would leave every odd row misaligned.
           ((long long)a.M * a.L * a.P) % 2 == 0;
}

// run-time layout of the raw projection outputs handed to the fused kernels (see fused_supported)
bool fused_raw_layout_ok(const FusedArgs& a)
{
    const size_t pair_bytes = a.raw_dtype == kF32 ? 8 : 4;
    const auto aligned = [&](const void* p) { return p == nullptr || ((size_t)p % pair_bytes) == 0; };
    return a.off_stride % 2 == 0 && aligned(a.offsets) && aligned(a.grad_offsets);
}

template <typename VT, typename RT>
static cudaError_t dispatch_fwd_fused(const FusedArgs& a, cudaStream_t stream)
{
    switch (a.D) {
        case 16: return launch_fwd_fused<VT, 16, RT>(a, stream);
        case 32: return launch_fwd_fused<VT, 32, RT>(a, stream);
        case 64: return launch_fwd_fused<VT, 64, RT>(a, stream);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t fused_forward(const FusedArgs& a, cudaStream_t stream)
{
    if (!fused_supported(a) || !fused_raw_layout_ok(a)) return cudaErrorInvalidValue;
    if (a.value_ld != 0 && (a.value_ld < (long long)a.M * a.D || a.value_ld % (16 / (a.dtype == kF32 ? 4 : 2)) != 0 ||
                            (long long)a.S * a.value_ld >= (1ll << 31)))
        return cudaErrorInvalidValue;
    if ((long long)a.N * a.Lq * a.M * a.D == 0) return cudaSuccess;
    if (a.dtype == kF32) return dispatch_fwd_fused<float, float>(a, stream);
    if (a.raw_dtype == kF32) return dispatch_fwd_fused<__nv_bfloat16, float>(a, stream);
    return dispatch_fwd_fused<__nv_bfloat16, __nv_bfloat16>(a, stream);
}

cudaError_t forward(const FwdArgs& a, cudaStream_t stream)
{
    if ((long long)a.N * a.Lq * a.M * a.D == 0) return cudaSuccess;
    if (!a.no_tc && tc_forward_supported(a)) return tc_forward(a, stream);
    switch (a.dtype) {
        case kF32:  return dispatch_fwd_16or32<float>(a, stream);
        case kBF16: return dispatch_fwd_16or32<__nv_bfloat16>(a, stream);
        case kF16:  return dispatch_fwd_16or32<__half>(a, stream);
        case kF64:  return launch_fwd_generic<double>(a, stream);
    }
    return cudaErrorInvalidValue;
}

}  // namespace msda
