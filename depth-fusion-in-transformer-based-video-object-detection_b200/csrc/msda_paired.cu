// Forward multi-scale deformable attention over a PAIRED value layout (16-bit values), sm_100a.
//
// Why: the gather is bound by L1 wavefronts (one per distinct 128-byte line an instruction touches).
// In the reference layout value[N,S,M,D] a head's slice of a pixel is D*2 = 64 bytes (bf16, D=32):
// half a line per bilinear corner, 4 lines per sample.  The two x-adjacent corners of a sample are
// consecutive pixels, so a layout that stores, per pixel index r and head m, the slices of pixels
// r-1 and r next to each other
//     pairs[n, r, m, 0, :] = value[n, r-1, m, :]      r = 0 .. S      (zeros for r-1 = -1)
//     pairs[n, r, m, 1, :] = value[n, r,   m, :]                      (zeros for r   =  S)
// makes one (corner-pair, head) record exactly one 128-byte line: 2 lines per sample instead of 4.
// The layout costs one streaming pass (pack_value_pairs_kernel: read S*C, write 2*(S+1)*C) and is
// only worth it when the gather dominates (encoder-sized query sets); the host side decides.
//
// Kernel: a group of G = D/4 lanes owns one (n,q,m) pair; the lower half of the group holds the
// x0 corner's channels, the upper half the x1 corner's (8 channels = 16 bytes per lane).  Phase 1
// (one sample per lane) parks two 16-byte records per sample in shared memory, one per half:
// {element offset of the y0 row's record, of the y1 row's, weight(y0, this x), weight(y1, this x)}.  Phase 2:
// one LDS.128 + two LDG.128 + 16 FFMA per sample per lane, 4 samples unrolled.  The halves are
// combined with one xor-shuffle per channel at the end.
// Semantics are those of msda_fwd_fast_kernel (msda_forward.cu), i.e. of the reference's
// ms_deformable_im2col_gpu_kernel (cuda/ms_deform_im2col_cuda.cuh:237-299).
#include "msda_common.cuh"
#include "msda_launch.h"

namespace msda {

// one thread per (record r, 16-byte chunk of the C-wide pixel row)
template <typename VT>
__global__ void __launch_bounds__(256)
pack_value_pairs_kernel(const VT* __restrict__ value, VT* __restrict__ pairs, int S, int M, int D,
                        long long total_chunks)
{
    constexpr int V = 16 / (int)sizeof(VT);
    const int cph = D / V;                       // chunks per head
    const int cpr = M * cph;                     // chunks per pixel row
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total_chunks; idx += stride) {
        const long long rec = idx / cpr;
        const int ch = (int)(idx - rec * cpr);
        const long long n = rec / (S + 1);
        const int r = (int)(rec - n * (S + 1));
        const int m = ch / cph, j = ch - m * cph;
        const VT* row = value + ((n * S + r) * (long long)cpr + ch) * V;       // pixel r
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        const uint4 lo = r > 0 ? ldg_v4(row - (long long)cpr * V) : z;          // pixel r-1
        const uint4 hi = r < S ? ldg_stream_v4(row) : z;
        VT* dst = pairs + ((rec * M + m) * 2) * (long long)D + j * V;
        *reinterpret_cast<uint4*>(dst) = lo;
        *reinterpret_cast<uint4*>(dst + D) = hi;
    }
}

template <int PAIRS> struct PairedWarps { static constexpr int value = PAIRS >= 16 ? 2 : (PAIRS >= 8 ? 4 : 8); };
constexpr int kRecPad = 2;       // records: group stride (2*kChunk + 2) * 16 B -> the groups of a warp hit distinct banks

// 1: accumulate channel pairs with packed fp32 FMAs (msda_common.cuh) -- 8 FFMA2 instead of 16 FFMA per sample per lane in
// the kernel that ncu showed issue-bound (83 % issue-active).  Bit-identical arithmetic.  Compile-checked (no spills, 64
// FFMA2 in the D = 32 bf16 kernel) but NOT yet run on a GPU: off by default, first A/B of the next round
// (make EXTRA=-DMSDA_PAIRED_F32X2=1; tests/test_gpu_paired.py; MultiScaleDeformableAttention.PAIRED_FORWARD = True).
#ifndef MSDA_PAIRED_F32X2
#define MSDA_PAIRED_F32X2 0
#endif

#ifndef MSDA_PAIRED_MINWARPS
#define MSDA_PAIRED_MINWARPS 40
#endif

// acc += v * w with v, w bf16 and acc fp32 in ONE instruction (FHFMA.BF16; the halves of a packed register are
// selected by the instruction, no unpack)
__device__ __forceinline__ float fma_bf16_f32(unsigned short v, unsigned short w, float acc)
{
    float d;
    asm("fma.rn.f32.bf16 %0, %1, %2, %3;" : "=f"(d) : "h"(v), "h"(w), "f"(acc));
    return d;
}
__device__ __forceinline__ void split_halves(unsigned x, unsigned short& lo, unsigned short& hi)
{
    asm("mov.b32 {%0,%1}, %2;" : "=h"(lo), "=h"(hi) : "r"(x));
}

// WB = true (bf16 only): the per-corner weights (attention weight x bilinear weight) are rounded to bf16 and
// the accumulation uses the mixed-precision FMA above -- 16 instead of 32 math instructions per sample per
// lane, which is what bounds this kernel (ncu: 83 % issue-active).  It is the rounding every bf16 tensor-core
// attention applies to its probabilities; results stay inside the stated bf16 tolerance
// (tests/test_gpu_paired.py).  WB = false keeps fp32 weights.
template <typename VT, int D, bool FUSED, typename RT, bool WB>
__global__ void __launch_bounds__(PairedWarps<32 / (D / 4)>::value * 32,
                                  MSDA_PAIRED_MINWARPS / PairedWarps<32 / (D / 4)>::value)
msda_fwd_paired_kernel(const VT* __restrict__ pairs, const int64_t* __restrict__ shapes,
                       const int64_t* __restrict__ lsi, const SampleSrc src, VT* __restrict__ out,
                       int S, int M, int L, int Lq, int P, int p_magic, long long total_pairs)
{
    static_assert(sizeof(VT) == 2, "paired layout is for 16-bit values");
    constexpr int EPL = 8;                       // channels per lane (16 bytes)
    constexpr int G = D / 4;                     // lanes per pair: D/8 for each x corner
    constexpr int HALF = G / 2;
    constexpr int PAIRS = 32 / G;
    constexpr int WARPS = PairedWarps<PAIRS>::value;
    static_assert(G >= 2 && G <= 32 && (G & (G - 1)) == 0, "head width must map to a power-of-two lane group");

    __shared__ int s_meta[3 * kMaxLevelsFast];
    __shared__ __align__(16) int4 s_rec[WARPS][PAIRS][2 * kChunk + kRecPad];

    if (threadIdx.x < L) {
        s_meta[3 * threadIdx.x + 0] = (int)shapes[2 * threadIdx.x];
        s_meta[3 * threadIdx.x + 1] = (int)shapes[2 * threadIdx.x + 1];
        s_meta[3 * threadIdx.x + 2] = (int)lsi[threadIdx.x];
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane / G, sub = lane % G;
    const int half = sub / HALF, jch = sub % HALF;
    // a CTA owns WARPS*PAIRS consecutive queries of ONE head (see msda_forward.cu)
    const int m = (int)(blockIdx.x % M);
    const long long nq_total = total_pairs / M;
    const long long nq_raw = ((long long)(blockIdx.x / M) * WARPS + warp) * PAIRS + grp;
    const bool active = nq_raw < nq_total;
    const long long nq = active ? nq_raw : nq_total - 1;
    const long long pair = nq * M + m;
    const long long n = nq / Lq;
    const int LP = L * P;
    const unsigned rec_stride = (unsigned)(M * 2 * D);                        // elements between records r and r+1
    const VT* vbase = pairs + (n * (S + 1) * M + m) * (long long)(2 * D) + sub * EPL;
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

    for (int s0 = 0; s0 < LP; s0 += kChunk) {
        const int cnt = min(kChunk, LP - s0);
        const int cnt4 = (cnt + 3) & ~3;
        // ---- phase 1: footprints, one sample per lane of the group -------------------------
        constexpr int K = (kChunk + G - 1) / G;
        float prob[K];
        float inv_sum = 1.f;
        if constexpr (FUSED) {                           // host guarantees L*P <= kChunk: one chunk
            float mx = -INFINITY;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int j = sub + k * G;
                prob[k] = j < cnt ? load_raw1<RT>(gp + j) : -INFINITY;
                mx = fmaxf(mx, prob[k]);
            }
            mx = group_max<G>(mx);
            float sum = 0.f;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                prob[k] = (sub + k * G) < cnt ? expf(prob[k] - mx) : 0.f;
                sum += prob[k];
            }
            inv_sum = group_sum<G>(sum);
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int j = sub + k * G;
            if (j >= cnt4) break;
            int4 r0 = make_int4(0, 0, 0, 0), r1 = make_int4(0, 0, 0, 0);
            if (j < cnt) {
                const int s = s0 + j;
                const int l = div_by_points(s, p_magic);
                float2 xy;
                float a;
                if constexpr (FUSED) {
                    xy = fused_location(load_raw2<RT>(op + 2 * s), src.ref + (nq * L + l) * src.ref_dim, src.ref_dim,
                                        s_meta[3 * l], s_meta[3 * l + 1], P);
                    a = prob[k] / inv_sum;
                } else {
                    xy = ldg_stream_f32x2(lp + 2 * s);
                    a = ldg_stream_f32(ap + s);
                }
                const Footprint f = footprint<float>(xy.x, xy.y, s_meta[3 * l], s_meta[3 * l + 1], s_meta[3 * l + 2]);
                const float hw = 1.f - f.lw, hh = 1.f - f.lh;
                // record r holds pixels (r-1, r): corners (x0, x1) of row y live in record pix(y, x0) + 1.
                // Rows outside the map carry zero weights; their index is clamped into [0, S].
                const int ry0 = min(max(f.pix00 + 1, 0), S);
                const int ry1 = min(max(f.pix00 + f.rowstep + 1, 0), S);
                const float w00 = (f.ok & 1u) ? hh * hw * a : 0.f;
                const float w01 = (f.ok & 2u) ? hh * f.lw * a : 0.f;
                const float w10 = (f.ok & 4u) ? f.lh * hw * a : 0.f;
                const float w11 = (f.ok & 8u) ? f.lh * f.lw * a : 0.f;
                // element offsets (< 2^32, checked by the launcher): one IMAD.WIDE.U32 per load in phase 2
                const int o0 = (int)((unsigned)ry0 * rec_stride), o1 = (int)((unsigned)ry1 * rec_stride);
                if constexpr (WB) {
                    const __nv_bfloat162 p0 = __floats2bfloat162_rn(w00, w10), p1 = __floats2bfloat162_rn(w01, w11);
                    r0 = make_int4(o0, o1, (int)*reinterpret_cast<const unsigned*>(&p0), 0);   // x0 half: {w(y0), w(y1)}
                    r1 = make_int4(o0, o1, (int)*reinterpret_cast<const unsigned*>(&p1), 0);   // x1 half
                } else {
                    r0 = make_int4(o0, o1, __float_as_int(w00), __float_as_int(w10));   // x0 half
                    r1 = make_int4(o0, o1, __float_as_int(w01), __float_as_int(w11));   // x1 half
                }
            }
            s_rec[warp][grp][2 * j] = r0;
            s_rec[warp][grp][2 * j + 1] = r1;
        }
        __syncwarp();
        // ---- phase 2: gather, 4 samples x 2 row-pairs in flight ----------------------------
        for (int j0 = 0; j0 < cnt4; j0 += 4) {
            uint4 raw[4][2];
            int wz[4], ww[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int4 rec = s_rec[warp][grp][2 * (j0 + u) + half];
                wz[u] = rec.z;
                ww[u] = rec.w;
                raw[u][0] = ldg_v4(vbase + (unsigned)rec.x);
                raw[u][1] = ldg_v4(vbase + (unsigned)rec.y);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if constexpr (WB) {
                    unsigned short w0, w1;
                    split_halves((unsigned)wz[u], w0, w1);
                    const unsigned a0[4] = {raw[u][0].x, raw[u][0].y, raw[u][0].z, raw[u][0].w};
                    const unsigned a1[4] = {raw[u][1].x, raw[u][1].y, raw[u][1].z, raw[u][1].w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        unsigned short lo, hi;
                        split_halves(a0[i], lo, hi);
                        acc[2 * i] = fma_bf16_f32(lo, w0, acc[2 * i]);
                        acc[2 * i + 1] = fma_bf16_f32(hi, w0, acc[2 * i + 1]);
                        split_halves(a1[i], lo, hi);
                        acc[2 * i] = fma_bf16_f32(lo, w1, acc[2 * i]);
                        acc[2 * i + 1] = fma_bf16_f32(hi, w1, acc[2 * i + 1]);
                    }
                } else {
                    const float wy0 = __int_as_float(wz[u]), wy1 = __int_as_float(ww[u]);
                    float v0[EPL], v1[EPL];
                    unpack<VT>(raw[u][0], v0);
                    unpack<VT>(raw[u][1], v1);
#if MSDA_PAIRED_F32X2
                    const F2 W0 = f2_dup(wy0), W1 = f2_dup(wy1);
#pragma unroll
                    for (int c = 0; c < EPL; c += 2) {                     // the same two fmas per channel, two channels at once
                        F2 a, x0, x1;
                        a.x = acc[c]; a.y = acc[c + 1];
                        x0.x = v0[c]; x0.y = v0[c + 1];
                        x1.x = v1[c]; x1.y = v1[c + 1];
                        a = fma2(W1, x1, fma2(W0, x0, a));
                        acc[c] = a.x;
                        acc[c + 1] = a.y;
                    }
#else
#pragma unroll
                    for (int c = 0; c < EPL; ++c) acc[c] = fmaf(wy1, v1[c], fmaf(wy0, v0[c], acc[c]));
#endif
                }
            }
        }
        __syncwarp();
    }
    // x0 half + x1 half
#pragma unroll
    for (int c = 0; c < EPL; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], HALF);
    if (active && half == 0) stg_stream_v4(out + pair * D + jch * EPL, pack<VT>(acc));
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
bool paired_supported(int dtype, int D)
{
    return (dtype == kBF16 || dtype == kF16) && (D == 16 || D == 32 || D == 64);
}

cudaError_t pack_value_pairs(int dtype, const void* value, void* pairs, int N, int S, int M, int D, cudaStream_t stream)
{
    if (!paired_supported(dtype, D)) return cudaErrorInvalidValue;
    const long long total = (long long)N * (S + 1) * M * (D / 8);
    if (total == 0) return cudaSuccess;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    if (dtype == kBF16)
        pack_value_pairs_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, stream>>>(
            (const __nv_bfloat16*)value, (__nv_bfloat16*)pairs, S, M, D, total);
    else
        pack_value_pairs_kernel<__half><<<(unsigned)blocks, 256, 0, stream>>>(
            (const __half*)value, (__half*)pairs, S, M, D, total);
    return cudaGetLastError();
}

template <typename VT, int D, bool FUSED, typename RT, bool WB>
static cudaError_t launch_paired(const VT* pairs, const int64_t* shapes, const int64_t* lsi, const SampleSrc& src,
                                 VT* out, int N, int S, int M, int L, int Lq, int P, cudaStream_t stream)
{
    constexpr int G = D / 4;
    constexpr int PAIRS = 32 / G;
    constexpr int WARPS = PairedWarps<PAIRS>::value;
    const long long total_pairs = (long long)N * Lq * M;
    const long long nq_total = (long long)N * Lq;
    const long long blocks = ((nq_total + WARPS * PAIRS - 1) / (WARPS * PAIRS)) * M;
    if (blocks > 0x7fffffffll) return cudaErrorInvalidConfiguration;
    const int p_magic = (65536 + P - 1) / P;
    msda_fwd_paired_kernel<VT, D, FUSED, RT, WB><<<(unsigned)blocks, WARPS * 32, 0, stream>>>(
        pairs, shapes, lsi, src, out, S, M, L, Lq, P, p_magic, total_pairs);
    return cudaGetLastError();
}

template <typename VT, bool FUSED, typename RT, bool WB>
static cudaError_t dispatch_paired(int D, const void* pairs, const int64_t* shapes, const int64_t* lsi,
                                   const SampleSrc& src, void* out, int N, int S, int M, int L, int Lq, int P,
                                   cudaStream_t stream)
{
    switch (D) {
        case 16: return launch_paired<VT, 16, FUSED, RT, WB>((const VT*)pairs, shapes, lsi, src, (VT*)out, N, S, M, L, Lq, P, stream);
        case 32: return launch_paired<VT, 32, FUSED, RT, WB>((const VT*)pairs, shapes, lsi, src, (VT*)out, N, S, M, L, Lq, P, stream);
        case 64: return launch_paired<VT, 64, FUSED, RT, WB>((const VT*)pairs, shapes, lsi, src, (VT*)out, N, S, M, L, Lq, P, stream);
        default: return cudaErrorInvalidValue;
    }
}

// a.value is the PAIRED tensor [N, S+1, M, 2, D]; a.force_generic carries the flags (bit 1: bf16 weights)
cudaError_t forward_paired(const FwdArgs& a, cudaStream_t stream)
{
    if (!paired_supported(a.dtype, a.D) || a.L > kMaxLevelsFast || a.P > 64 || (long long)a.L * a.P * a.P >= 65536 ||
        (long long)(a.S + 1) * a.M * a.D * 2 >= (1ll << 32))
        return cudaErrorInvalidValue;
    if ((long long)a.N * a.Lq * a.M == 0) return cudaSuccess;
    SampleSrc src;
    src.loc = a.loc; src.attn = a.attn; src.ref = nullptr; src.loc_stride = 0; src.attn_stride = 0; src.ref_dim = 0;
    if (a.dtype == kBF16 && (a.force_generic & 2))
        return dispatch_paired<__nv_bfloat16, false, float, true>(a.D, a.value, a.shapes, a.lsi, src, a.out, a.N, a.S,
                                                                  a.M, a.L, a.Lq, a.P, stream);
    if (a.dtype == kBF16)
        return dispatch_paired<__nv_bfloat16, false, float, false>(a.D, a.value, a.shapes, a.lsi, src, a.out, a.N, a.S,
                                                                   a.M, a.L, a.Lq, a.P, stream);
    return dispatch_paired<__half, false, float, false>(a.D, a.value, a.shapes, a.lsi, src, a.out, a.N, a.S, a.M, a.L,
                                                        a.Lq, a.P, stream);
}

// a.value is the PAIRED tensor; supported: what fused_supported covers with a 16-bit dtype
cudaError_t fused_forward_paired(const FusedArgs& a, int flags, cudaStream_t stream)
{
    if (!fused_supported(a) || a.dtype != kBF16 || (long long)(a.S + 1) * a.M * a.D * 2 >= (1ll << 32))
        return cudaErrorInvalidValue;
    if ((long long)a.N * a.Lq * a.M * a.D == 0) return cudaSuccess;
    SampleSrc src;
    src.loc = a.offsets; src.attn = a.logits; src.ref = a.ref;
    src.loc_stride = a.off_stride; src.attn_stride = a.logit_stride; src.ref_dim = a.ref_dim;
    const bool wb = (flags & 2) != 0;
    if (a.raw_dtype == kF32)
        return wb ? dispatch_paired<__nv_bfloat16, true, float, true>(a.D, a.value, a.shapes, a.lsi, src, a.out, a.N, a.S,
                                                                      a.M, a.L, a.Lq, a.P, stream)
                  : dispatch_paired<__nv_bfloat16, true, float, false>(a.D, a.value, a.shapes, a.lsi, src, a.out, a.N,
                                                                       a.S, a.M, a.L, a.Lq, a.P, stream);
    return wb ? dispatch_paired<__nv_bfloat16, true, __nv_bfloat16, true>(a.D, a.value, a.shapes, a.lsi, src, a.out, a.N,
                                                                          a.S, a.M, a.L, a.Lq, a.P, stream)
              : dispatch_paired<__nv_bfloat16, true, __nv_bfloat16, false>(a.D, a.value, a.shapes, a.lsi, src, a.out,
                                                                           a.N, a.S, a.M, a.L, a.Lq, a.P, stream);
}

}  // namespace msda
