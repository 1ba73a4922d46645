// EXPERIMENT, NOT BUILT (kept as the record of a measured negative result; see DESIGN.md section 3.2 and
// profiles/r2_bwd_cta_merge.txt).  msda_backward.cu with the grad_value rows of a CTA (32 neighbouring queries of one head)
// merged per pixel in shared memory before they leave the SM: hash table keyed by pixel (atomicCAS), per-slot linked lists
// (atomicExch), one red.global.add.v4.f32 per distinct pixel and pass.  Parity: the 111 tests of test_gpu_op_parity /
// test_gpu_fused / test_gpu_tc_forward pass.  Measured on B200, fp32 batch 8: reduction rows 69 M -> 22 M (request path
// 84 % -> 23 % busy) but 2.29 ms instead of 1.57 ms -- the merge moves every row through shared memory once more, and the
// L1 data pipe (128 B/clk/SM, shared by global loads, shared-memory traffic and shuffles) goes to 88 % busy with 440 M
// shared-memory wavefronts.
// Backward multi-scale deformable attention for sm_100a.
//
//   grad_value[n, pix, m, :] += w_corner * A * g            (scatter over the 4 corners)
//   grad_attn [n,q,m,l,p]     = sum_c g_c * bilinear_c
//   grad_loc  [n,q,m,l,p]     = ( W * A * sum_c g_c * d bilinear_c / dx ,
//                                 H * A * sum_c g_c * d bilinear_c / dy )
//
// Replaces the reference's six col2im kernels + switch(channels) (cuda/ms_deform_im2col_cuda.cuh:
// 301-920, 956-1327; helper :87-159).  The production reference kernel for D=32 runs one
// 32-thread block per (n,q,m), issues 4 scalar atomicAdd per channel per sample, and reduces
// grad_loc / grad_attn through shared memory with thread 0 summing serially between two
// __syncthreads per sample (:376-394).
//
// Fast kernel (same lane mapping as the forward: a group of G lanes owns a pair, each lane a
// 16-byte channel slice):
//   * phase 1: one lane per sample computes the footprint ONCE and parks everything phase 2 needs in shared memory,
//     already masked for corners outside the map: the four row offsets (clamped to a harmless in-range row), the
//     grad_value coefficient of each row, and the coefficients that turn the four corner dot products
//     t_k = <grad_output, value_k> into grad_attn and the two grad_loc partials
//         pa = sum_k ca_k t_k      ca = (hh*hw, hh*lw, lh*hw, lh*lw)
//         px = sum_k cx_k t_k      cx = a * (-hh, +hh, -lh, +lh)
//         py = sum_k cy_k t_k      cy = a * (-hw, -lw, +hw, +lw)
//     (the same sums as cuh:113-158, with the channel sum taken first);
//   * phase 2, per sample and lane: 4 unconditional loads, 8 packed FMAs for the dots, 6 for the three outputs, three
//     8-lane butterflies, 4 predicated reductions -- ~80 instructions where the round-1 kernel needed 164 (64-bit
//     address arithmetic per corner, zero-filled predicated loads, branches around every reduction);
//   * grad_value uses ONE 16-byte vector reduction (red.global.add.v4.f32 -> REDG.E.ADD.F32x4)
//     per lane per corner instead of 4 scalar atomics, always into an fp32 buffer.  A lane
//     always owns 4 channels here (16-bit values are read with 8-byte loads) so that the 8 lanes
//     of a group cover one whole 128-byte fp32 row per instruction: measured on B200, the
//     SM->L2 reduction path costs ~5.5 cycles per (instruction, row) whether the row is written
//     whole or in halves (tools/microbench_red.cu), and it is what bounds this kernel;
//   * grad_loc / grad_attn are written for every sample (zeros for samples outside the map),
//     so they need no zero-fill pass; only grad_value is memset.
// Generic kernel: any D / dtype (fp64 for gradcheck): one warp per pair, lanes stride channels.
#include <type_traits>
#include "msda_common.cuh"
#include "msda_launch.h"

namespace msda {

// warps per CTA: at most 32 queries of a head per CTA (the merge tables are sized by queries x samples per pass)
template <int PAIRS> struct BwdWarps { static constexpr int value = PAIRS >= 32 ? 1 : (PAIRS >= 16 ? 2 : (PAIRS >= 8 ? 4 : 8)); };

// resident CTAs per SM the register allocation must allow.  Measured on B200 (tools/ab_variants.sh,
// fp32 / bf16 backward, batch 8): 2 x 8 warps at 128 registers 1.90 / 1.88 ms; 3 x 8 warps at 80
// registers 1.74 / 1.76 ms; 4 x 8 warps at 64 registers (immediate reduction, no spills) 1.70 / 1.75 ms.
#ifndef MSDA_BWD_MINBLOCKS
#define MSDA_BWD_MINBLOCKS 4
#endif

// Where the per-sample gradients go.  Plain: grad_sampling_loc / grad_attn_weight.  Fused:
// gradients w.r.t. the raw projection outputs (through the location arithmetic and the softmax)
// and, optionally, w.r.t. the reference points.
struct GradDst {
    void* loc;        // plain: grad_loc [pairs,LP,2] fp32 ; fused: grad_offsets (addressing of SampleSrc)
    void* attn;       // plain: grad_attn [pairs,LP] fp32  ; fused: grad_logits
    float* ref;       // fused: grad_ref [N*Lq, L, ref_dim] or nullptr
};

// FUSED: see msda_forward.cu.  Backward of the fused layer op additionally applies
//   d loc / d offset  (1/W, 1/H  or  0.5*wh/P)           reference modules/ms_deform_attn.py:102-110
//   softmax backward   g_logit = a * (g_a - sum_j a_j g_a_j)                      :99-100
// to the finished per-sample gradients in phase 3, so neither the locations / weights nor their
// gradients ever exist in HBM.
// Samples per pass.  8 (instead of the forward's 16) halves the per-lane partial-sum registers
// (3 per sample) and lets a lane keep its own samples' weights in registers between phases: the
// kernel is bound by load latency under RED traffic (ncu: long-scoreboard stalls, 15 resident
// warps at 128 registers), so registers buy resident warps.  3*CH must be divisible by G.
#ifndef MSDA_BWD_CHUNK
#define MSDA_BWD_CHUNK 8
#endif
// grad_value rows are merged ACROSS THE CTA before they leave the SM.  A CTA owns Q neighbouring queries of one head; their
// samples of a level land on heavily overlapping pixels (bench distribution, Q = 32: 3.6 corner rows per distinct pixel
// and pass; 5 at initialisation), and the SM -> L2 reduction path (5.5 cycles per 128-byte row, tools/microbench_red.cu)
// is what bounded the round-1 kernel.  Per pass:
//   * phase 1: the lane that owns a sample files each of its corner rows under the row's pixel in a shared-memory hash
//     table (open addressing, atomicCAS on the key; a slot's rows form a linked list, atomicExch on its head); a slot's
//     first row also appends the slot to a dense list.  Only integer shared-memory atomics -- fp32 atomics on shared
//     memory are a CAS loop on sm_100a and cost as much as the global reduction they would replace;
//   * phase C, after a block barrier: each lane group takes slots off the dense list, sums coefficient x grad_output row
//     over the slot's rows (the CTA's grad_output rows sit in shared memory) and issues ONE red.global.add.v4.f32 per
//     lane for the pixel; the group then clears the slot for the next pass.
// 1: the fused kernel's passes run as a real loop instead of FCH unrolled copies.  Unrolled, the D = 32 bf16 kernel is
// 4832 instructions (77 KB of SASS) that every warp walks end to end, and ncu shows 14 % of its stall samples waiting for
// instructions (`no_instructions`, profiles/ncu_r1k.json); the per-pass register arrays are then indexed through
// reg_pick / reg_put (selects over a static index) so that they stay in registers.
#ifndef MSDA_BWD_FUSED_ROLLED
#define MSDA_BWD_FUSED_ROLLED 1
#endif
template <int V> struct Log2 { static constexpr int value = 1 + Log2<V / 2>::value; };
template <> struct Log2<1> { static constexpr int value = 0; };

template <int K>
__device__ __forceinline__ float reg_pick(const float (&a)[K], int idx)
{
    float r = a[0];
#pragma unroll
    for (int k = 1; k < K; ++k) r = idx == k ? a[k] : r;
    return r;
}
template <int K>
__device__ __forceinline__ void reg_put(float (&a)[K], int idx, float v)
{
#pragma unroll
    for (int k = 0; k < K; ++k) a[k] = idx == k ? v : a[k];
}

template <typename VT, int D, bool FUSED, typename RT>
__global__ void __launch_bounds__(BwdWarps<32 / (D / 4)>::value * 32, MSDA_BWD_MINBLOCKS)
msda_bwd_fast_kernel(const VT* __restrict__ grad_out, const VT* __restrict__ value,
                     const int64_t* __restrict__ shapes, const int64_t* __restrict__ lsi,
                     const SampleSrc src, float* __restrict__ gv_accum, const GradDst dst,
                     int S, int M, int L, int Lq, int P, int p_magic, long long total_pairs,
                     const unsigned char* __restrict__ red_levels)
{
    // red_levels (optional): per pair, bit l set = this kernel issues the grad_value reductions of level l; clear =
    // another kernel accumulates that level (msda_tc_backward.cu).  nullptr = every level.
    constexpr int EPL = 4;                       // channels per lane: one red.v4.f32 per corner
    using SliceT = Slice<VT, EPL>;
    constexpr int G = D / EPL;
    constexpr int PAIRS = 32 / G;
    constexpr int WARPS = BwdWarps<PAIRS>::value;
    constexpr int CH = MSDA_BWD_CHUNK > G ? MSDA_BWD_CHUNK : G;   // samples per pass
    constexpr int SPL = CH / G;                  // samples a lane owns per pass: j = sub*SPL + i, in phase 1
                                                 // (footprints) and again after the reduce-scatter (gradients)
    constexpr int FCH = (kChunk + CH - 1) / CH;  // passes of the fused op (host guarantees L*P <= kChunk)
    static_assert(G >= 1 && G <= 32 && (G & (G - 1)) == 0, "fast backward needs 1..32 lanes per head");
    static_assert(CH % G == 0, "pass size must split evenly over the lanes");

    constexpr int Q = WARPS * PAIRS;             // queries (of one head) per CTA
    constexpr int E = Q * CH * 4;                // corner rows of a pass
    constexpr int T = 2 * E;                     // hash slots: load factor <= 0.5
    constexpr int kLogT = Log2<T>::value;
    constexpr int kLogRowsPerQuery = Log2<CH * 4>::value;
    constexpr unsigned kEmpty = 0xffffffffu;
    constexpr unsigned short kEnd = 0xffffu;
    static_assert(E <= 0xffff, "row ids are 16-bit");

    // per-sample records of a pass (+1 record: the groups of a warp start in distinct banks)
    __shared__ int s_meta[3 * kMaxLevelsFast];
    // record of a sample, 4 x 16 bytes: [0] byte offset (pixel * M*D * sizeof(VT)) of each corner row (corner outside the
    // map -> 0), [1..3] corner dots -> grad_attn, grad_loc.x / W, grad_loc.y / H
    constexpr int kRec = 4;
    __shared__ __align__(16) uint4 s_rec[WARPS][PAIRS][CH * kRec + 1];
    // the merge tables (see the top of the file)
    __shared__ unsigned s_key[T];                // pixel of the slot: (frame - first frame of the CTA) * S + pixel
    __shared__ int s_head[T];                    // newest row of the slot, -1 = none
    __shared__ unsigned short s_next[E];         // next row of the same slot
    __shared__ unsigned short s_list[E];         // slots in use
    __shared__ float s_coef[E];                  // grad_value coefficient of a row
    __shared__ int s_nslots[2];                  // per pass parity
    __shared__ __align__(16) float4 s_g[Q][G];   // grad_output rows of the CTA's queries

    if (threadIdx.x < L) {
        s_meta[3 * threadIdx.x + 0] = (int)shapes[2 * threadIdx.x];
        s_meta[3 * threadIdx.x + 1] = (int)shapes[2 * threadIdx.x + 1];
        s_meta[3 * threadIdx.x + 2] = (int)lsi[threadIdx.x];
    }
    for (int i = threadIdx.x; i < T; i += WARPS * 32) {
        s_key[i] = kEmpty;
        s_head[i] = -1;
    }
    if (threadIdx.x < 2) s_nslots[threadIdx.x] = 0;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane / G, sub = lane % G;
    const int qloc = warp * PAIRS + grp;         // this group's query inside the CTA
    // a CTA owns Q consecutive queries of ONE head: neighbouring queries of a head sample overlapping pixels (L1 reuse
    // for the gather, shared rows for the merge)
    const int m = (int)(blockIdx.x % M);
    const long long nq_total = total_pairs / M;
    const long long nq_first = (long long)(blockIdx.x / M) * Q;
    const long long nq_raw = nq_first + qloc;
    const bool active = nq_raw < nq_total;
    const long long nq = active ? nq_raw : nq_total - 1;
    const long long pair = nq * M + m;
    const long long n = nq / Lq;
    const long long n_first = nq_first / Lq;      // first frame the CTA touches
    const int LP = L * P;
    const int MD = M * D;
    const long long head_off = (n * S * M + m) * (long long)D + sub * EPL;
    // this lane's slice of pixel 0 of its (frame, head), as opaque addresses: a corner address is then one 64-bit add
    // (two instructions) of the 32-bit BYTE offset parked in shared memory -- scaled by 4 / sizeof(VT) for grad_value
    const unsigned long long vaddr = opaque_addr(value + head_off);
    // phase C: this lane's channels of pixel 0 of the CTA's FIRST frame; a slot's key is the pixel offset from there
    const unsigned long long gaddr = opaque_addr(gv_accum + (n_first * S * M + m) * (long long)D + sub * EPL);
    const unsigned key_frame = (unsigned)(n - n_first) * (unsigned)S;
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

    const unsigned red_mask = red_levels != nullptr ? (unsigned)red_levels[pair] : 0xffffffffu;
    float g[EPL];
    SliceT::unpack(SliceT::load_stream(grad_out + pair * D + sub * EPL), g);
    if (!active) {
#pragma unroll
        for (int c = 0; c < EPL; ++c) g[c] = 0.f;     // clamped duplicate pair contributes nothing
    }
    F2 G01, G23;
    G01.x = g[0]; G01.y = g[1]; G23.x = g[2]; G23.y = g[3];
    s_g[qloc][sub] = make_float4(g[0], g[1], g[2], g[3]);
    __syncthreads();                              // level metadata, empty merge tables, grad_output rows

    // fused: softmax over the pair's L*P logits; this lane keeps the numerators of its own samples
    float prob[FCH * SPL];
    float inv_sum = 1.f;
    if constexpr (FUSED) {
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < FCH; ++c)
#pragma unroll
            for (int i = 0; i < SPL; ++i) {
                const int s = c * CH + sub * SPL + i;
                prob[c * SPL + i] = s < LP ? load_raw1<RT>(gp + s) : -INFINITY;
                mx = fmaxf(mx, prob[c * SPL + i]);
            }
        mx = group_max<G>(mx);
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < FCH * SPL; ++k) {
            prob[k] = prob[k] == -INFINITY ? 0.f : expf(prob[k] - mx);
            sum += prob[k];
        }
        inv_sum = group_sum<G>(sum);
    }
    // fused: finished per-sample gradients of this lane's own samples, kept until the softmax
    // backward can be applied (it needs sum_j a_j * g_a_j over the whole pair)
    float fin_x[FCH * SPL], fin_y[FCH * SPL], fin_a[FCH * SPL], own_a[FCH * SPL];

    const int passes = FUSED ? FCH : (LP + CH - 1) / CH;
#if MSDA_BWD_FUSED_ROLLED
#pragma unroll 1
#else
#pragma unroll
#endif
    for (int c = 0; c < (FUSED ? FCH : 1 << 30); ++c) {
        if (c >= passes) break;
        const int s0 = c * CH;
        if (FUSED && s0 >= LP) break;
        const int cnt = min(CH, LP - s0);
        // ---- phase 1: footprints of this lane's own samples ------------------------------------
        float a_own[SPL];
#pragma unroll
        for (int i = 0; i < SPL; ++i) {
            const int j = sub * SPL + i;
            int4 geo = make_int4(0, 0, 0, 0);
            float4 fr = make_float4(0.f, 0.f, 0.f, 0.f);
            a_own[i] = 0.f;
            if (j < cnt) {
                const int s = s0 + j;
                const int l = div_by_points(s, p_magic);
                float2 xy;
                float a;
                if constexpr (FUSED) {
                    xy = fused_location(load_raw2<RT>(op + 2 * s), src.ref + (nq * L + l) * src.ref_dim, src.ref_dim,
                                        s_meta[3 * l], s_meta[3 * l + 1], P);
                    a = reg_pick(prob, c * SPL + i) / inv_sum;
                } else {
                    xy = ldg_stream_f32x2(lp + 2 * s);
                    a = ldg_stream_f32(ap + s);
                }
                const Footprint f = footprint<float>(xy.x, xy.y, s_meta[3 * l], s_meta[3 * l + 1], s_meta[3 * l + 2]);
                geo = make_int4(f.pix00, f.rowstep, (int)f.ok, 0);
                fr = make_float4(f.lw, f.lh, a, 0.f);
                a_own[i] = a;
            }
            // coefficient of each corner row in grad_value: bilinear weight x attention weight (cuh:113-116); every
            // coefficient set is zero for corners outside the map (and for samples past cnt: ok = 0)
            const float lw = fr.x, lh = fr.y, hw = 1.f - fr.x, hh = 1.f - fr.y, a = fr.z;
            const bool k0 = geo.z & 1, k1 = geo.z & 2, k2 = geo.z & 4, k3 = geo.z & 8;
            const float4 ca = make_float4(k0 ? hh * hw : 0.f, k1 ? hh * lw : 0.f, k2 ? lh * hw : 0.f, k3 ? lh * lw : 0.f);
            const int my_level = div_by_points(s0 + j, p_magic);
            // grad_value coefficient of each corner row: bilinear weight x attention weight (cuh:113-116)
            const float ck[4] = {ca.x * a, ca.y * a, ca.z * a, ca.w * a};
            const int pixk[4] = {geo.x, geo.x + 1, geo.x + geo.y, geo.x + geo.y + 1};
            const unsigned MDu = (unsigned)MD * (unsigned)sizeof(VT);   // bytes between neighbouring pixels
            if (active && ((red_mask >> my_level) & 1u)) {
                // file the rows under their pixels (zero-coefficient rows -- corners outside the map, exact-zero
                // bilinear weights -- are dropped)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (ck[k] != 0.f) {
                        const unsigned key = key_frame + (unsigned)pixk[k];
                        unsigned h = (key * 0x9E3779B1u) >> (32 - kLogT);
                        while (true) {
                            const unsigned prev = atomicCAS(&s_key[h], kEmpty, key);
                            if (prev == kEmpty) {
                                s_list[atomicAdd(&s_nslots[c & 1], 1)] = (unsigned short)h;
                                break;
                            }
                            if (prev == key) break;
                            h = (h + 1) & (T - 1);
                        }
                        const int e = ((qloc * CH + j) << 2) + k;
                        s_coef[e] = ck[k];
                        s_next[e] = (unsigned short)atomicExch(&s_head[h], e);
                    }
                }
            }
            uint4* rec = &s_rec[warp][grp][j * kRec];
            rec[0] = make_uint4(k0 ? (unsigned)pixk[0] * MDu : 0u, k1 ? (unsigned)pixk[1] * MDu : 0u,
                                k2 ? (unsigned)pixk[2] * MDu : 0u, k3 ? (unsigned)pixk[3] * MDu : 0u);
            const float ahh = a * hh, alh = a * lh, ahw = a * hw, alw = a * lw;
            reinterpret_cast<float4*>(rec)[1] = ca;
            reinterpret_cast<float4*>(rec)[2] = make_float4(k0 ? -ahh : 0.f, k1 ? ahh : 0.f, k2 ? -alh : 0.f, k3 ? alh : 0.f);
            reinterpret_cast<float4*>(rec)[3] = make_float4(k0 ? -ahw : 0.f, k1 ? -alw : 0.f, k2 ? ahw : 0.f, k3 ? alw : 0.f);
        }
        __syncwarp();

        // ---- phase 2: per-sample gather, corner dots -> grad_loc / grad_attn partials -------------
        float part[3 * SPL];
#pragma unroll
        for (int i = 0; i < 3 * SPL; ++i) part[i] = 0.f;

#pragma unroll
        for (int j0 = 0; j0 < CH; ++j0) {
            if (j0 < cnt) {
                const uint4* rec = &s_rec[warp][grp][j0 * kRec];
                const uint4 o = rec[0];
                const unsigned ok[4] = {o.x, o.y, o.z, o.w};
                typename SliceT::raw_t raw[4];
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    raw[k] = SliceT::load(reinterpret_cast<const VT*>(vaddr + ok[k]));
                const float4 ca = reinterpret_cast<const float4*>(rec)[1], cx = reinterpret_cast<const float4*>(rec)[2],
                             cy = reinterpret_cast<const float4*>(rec)[3];
                float t[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float v[EPL];
                    SliceT::unpack(raw[k], v);
                    F2 V01, V23;
                    V01.x = v[0]; V01.y = v[1]; V23.x = v[2]; V23.y = v[3];
                    const F2 d = fma2(G23, V23, mul2(G01, V01));      // this lane's 4 channels of <grad_output, value_k>
                    t[k] = d.x + d.y;
                }
                F2 T01, T23;
                T01.x = t[0]; T01.y = t[1]; T23.x = t[2]; T23.y = t[3];
                auto combine = [&](const float4& cf) -> float {
                    F2 C01, C23;
                    C01.x = cf.x; C01.y = cf.y; C23.x = cf.z; C23.y = cf.w;
                    const F2 r = fma2(C23, T23, mul2(C01, T01));
                    return r.x + r.y;
                };
                const float px = group_sum<G>(combine(cx));           // cuh:119-151 (grad_w_weight * top_grad_value)
                const float py = group_sum<G>(combine(cy));           // (grad_h_weight)
                const float pa = group_sum<G>(combine(ca));           // cuh:156
                if (j0 / SPL == sub) {
                    part[3 * (j0 % SPL) + 0] = px;
                    part[3 * (j0 % SPL) + 1] = py;
                    part[3 * (j0 % SPL) + 2] = pa;
                }
            }
        }
        __syncwarp();

        // ---- phase 3: combine the group's partials; each lane finishes its own SPL samples ----
        if constexpr (!FUSED) {
            float* grad_loc = static_cast<float*>(dst.loc);
            float* grad_attn = static_cast<float*>(dst.attn);
            if (active) {
#pragma unroll
                for (int i = 0; i < SPL; ++i) {
                    const int j = sub * SPL + i;
                    if (j < cnt) {
                        const int s = s0 + j;
                        const int l = div_by_points(s, p_magic);
                        const float Hf = (float)s_meta[3 * l], Wf = (float)s_meta[3 * l + 1];
                        float2 gl = make_float2(Wf * part[3 * i + 0], Hf * part[3 * i + 1]);   // cuh:157-158
                        *reinterpret_cast<float2*>(grad_loc + (pair * LP + s) * 2) = gl;
                        grad_attn[pair * LP + s] = part[3 * i + 2];                            // cuh:156
                    }
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < SPL; ++i) {
                reg_put(fin_x, c * SPL + i, part[3 * i + 0]);
                reg_put(fin_y, c * SPL + i, part[3 * i + 1]);
                reg_put(fin_a, c * SPL + i, part[3 * i + 2]);
                reg_put(own_a, c * SPL + i, a_own[i]);
            }
        }

        // ---- phase C: one reduction row per distinct pixel the CTA touched in this pass -----------
        __syncthreads();                              // every row of the pass is filed
        {
            const int n_slots = s_nslots[c & 1];
            if (threadIdx.x == 0) s_nslots[(c + 1) & 1] = 0;      // the next pass counts here (after the barrier below)
            const unsigned gmask = G == 32 ? 0xffffffffu : ((1u << G) - 1u) << (grp * G);
            const unsigned long long row_bytes = (unsigned long long)MD * 4ull;
            for (int i = qloc; i < n_slots; i += Q) {
                const int h = s_list[i];
                const unsigned key = s_key[h];
                int e = s_head[h];
                F2 a01 = f2_dup(0.f), a23 = f2_dup(0.f);
                while (e >= 0) {
                    const F2 cf = f2_dup(s_coef[e]);
                    const float4 gq = s_g[e >> kLogRowsPerQuery][sub];
                    F2 q01, q23;
                    q01.x = gq.x; q01.y = gq.y; q23.x = gq.z; q23.y = gq.w;
                    a01 = fma2(cf, q01, a01);
                    a23 = fma2(cf, q23, a23);
                    const unsigned short nx = s_next[e];
                    e = nx == kEnd ? -1 : (int)nx;
                }
                red_add_f32x4(reinterpret_cast<float*>(gaddr + key * row_bytes), a01.x, a01.y, a23.x, a23.y);
                __syncwarp(gmask);                    // the whole group has read the slot
                if (sub == 0) {
                    s_key[h] = kEmpty;
                    s_head[h] = -1;
                }
            }
        }
        __syncthreads();                              // tables are empty again
    }

    if constexpr (FUSED) {
        // softmax backward needs sum_j a_j * g_a_j over the pair's samples
        float dot = 0.f;
#pragma unroll
        for (int k = 0; k < FCH * SPL; ++k) {
            const int s = (k / SPL) * CH + sub * SPL + (k % SPL);
            if (s < LP) dot = fmaf(own_a[k], fin_a[k], dot);
        }
        dot = group_sum<G>(dot);
        if (active) {
            RT* gop = static_cast<RT*>(dst.loc) + nq * src.loc_stride + (long long)m * LP * 2;
            RT* ggp = static_cast<RT*>(dst.attn) + nq * src.attn_stride + (long long)m * LP;
            int run_l = -1;                       // lane-local run of samples on one level -> one grad_ref update
            float rx = 0.f, ry = 0.f, rw = 0.f, rh = 0.f;
            auto flush_ref = [&]() {
                if (dst.ref != nullptr && run_l >= 0) {
                    float* gr = dst.ref + (nq * L + run_l) * src.ref_dim;
                    atomicAdd(gr, rx);
                    atomicAdd(gr + 1, ry);
                    if (src.ref_dim == 4) { atomicAdd(gr + 2, rw); atomicAdd(gr + 3, rh); }
                }
            };
#pragma unroll
            for (int k = 0; k < FCH * SPL; ++k) {
                const int s = (k / SPL) * CH + sub * SPL + (k % SPL);
                if (s < LP) {
                    const int l = div_by_points(s, p_magic);
                    const float Hf = (float)s_meta[3 * l], Wf = (float)s_meta[3 * l + 1];
                    const float glx = Wf * fin_x[k], gly = Hf * fin_y[k];                     // d/d loc
                    const float a = own_a[k];
                    const float glogit = a * (fin_a[k] - dot);
                    float gox, goy, gwx = 0.f, gwy = 0.f;
                    if (src.ref_dim == 2) {
                        gox = glx / Wf;
                        goy = gly / Hf;
                    } else {
                        const float4 r = *reinterpret_cast<const float4*>(src.ref + (nq * L + l) * 4);
                        gox = glx * (r.z * 0.5f / (float)P);
                        goy = gly * (r.w * 0.5f / (float)P);
                        if (dst.ref != nullptr) {
                            const float2 off = load_raw2<RT>(op + 2 * s);
                            gwx = glx * (off.x / (float)P * 0.5f);
                            gwy = gly * (off.y / (float)P * 0.5f);
                        }
                    }
                    store_raw2<RT>(gop + 2 * s, gox, goy);
                    ggp[s] = from_f32<RT>(glogit);
                    if (dst.ref != nullptr) {
                        if (l != run_l) { flush_ref(); run_l = l; rx = ry = rw = rh = 0.f; }
                        rx += glx; ry += gly; rw += gwx; rh += gwy;
                    }
                }
            }
            flush_ref();
        }
    }
}

// ------------------------------------------------------------------------------------------
// Generic path: one warp per pair, lanes stride the channels, scalar atomics.
// ------------------------------------------------------------------------------------------
template <typename VT> struct ScalarLoad {
    using acc_t = typename Traits<VT>::acc_t;
    static __device__ __forceinline__ acc_t load(const VT* p) { return (acc_t)to_f32<VT>(*p); }
};
template <> struct ScalarLoad<double> {
    using acc_t = double;
    static __device__ __forceinline__ double load(const double* p) { return *p; }
};

template <typename VT>
__global__ void __launch_bounds__(256)
msda_bwd_generic_kernel(const VT* __restrict__ grad_out, const VT* __restrict__ value,
                        const int64_t* __restrict__ shapes, const int64_t* __restrict__ lsi,
                        const typename Traits<VT>::loc_t* __restrict__ loc,
                        const typename Traits<VT>::loc_t* __restrict__ attn,
                        typename Traits<VT>::acc_t* __restrict__ gv_accum,
                        typename Traits<VT>::loc_t* __restrict__ grad_loc,
                        typename Traits<VT>::loc_t* __restrict__ grad_attn,
                        int S, int M, int D, int L, int Lq, int P, long long total_pairs)
{
    using acc_t = typename Traits<VT>::acc_t;
    using loc_t = typename Traits<VT>::loc_t;
    const int lane = threadIdx.x & 31;
    const long long warps_total = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long MD = (long long)M * D;
    for (long long pair = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; pair < total_pairs;
         pair += warps_total) {
        const int m = (int)(pair % M);
        const long long n = (pair / M) / Lq;
        const long long head_off = (n * S * M + m) * (long long)D;
        const VT* go = grad_out + pair * D;
        const loc_t* lp = loc + pair * L * P * 2;
        const loc_t* ap = attn + pair * L * P;
        for (int l = 0; l < L; ++l) {
            const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1], start = (int)lsi[l];
            for (int p = 0; p < P; ++p) {
                const int s = l * P + p;
                const loc_t x = lp[2 * s], y = lp[2 * s + 1];
                const acc_t a = (acc_t)ap[s];
                const Footprint f = footprint<loc_t>(x, y, H, W, start);
                acc_t px = 0, py = 0, pa = 0;
                if (f.ok) {
                    const loc_t w_im = x * (loc_t)W - (loc_t)0.5, h_im = y * (loc_t)H - (loc_t)0.5;
                    const acc_t lw = (acc_t)(w_im - floor(w_im)), lh = (acc_t)(h_im - floor(h_im));
                    const acc_t hw = 1 - lw, hh = 1 - lh;
                    const acc_t w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
                    const long long o00 = head_off + (long long)f.pix00 * MD;
                    const long long row = (long long)f.rowstep * MD;
                    for (int c = lane; c < D; c += 32) {
                        const acc_t gc = ScalarLoad<VT>::load(go + c);
                        const acc_t tg = a * gc;
                        acc_t v1 = 0, v2 = 0, v3 = 0, v4 = 0;
                        if (f.ok & 1u) { v1 = ScalarLoad<VT>::load(value + o00 + c);            atomicAdd(gv_accum + o00 + c, w1 * tg); }
                        if (f.ok & 2u) { v2 = ScalarLoad<VT>::load(value + o00 + MD + c);       atomicAdd(gv_accum + o00 + MD + c, w2 * tg); }
                        if (f.ok & 4u) { v3 = ScalarLoad<VT>::load(value + o00 + row + c);      atomicAdd(gv_accum + o00 + row + c, w3 * tg); }
                        if (f.ok & 8u) { v4 = ScalarLoad<VT>::load(value + o00 + row + MD + c); atomicAdd(gv_accum + o00 + row + MD + c, w4 * tg); }
                        pa += gc * (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4);
                        px += tg * (hh * (v2 - v1) + lh * (v4 - v3));
                        py += tg * (hw * (v3 - v1) + lw * (v4 - v2));
                    }
                }
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) {
                    px += __shfl_xor_sync(0xffffffffu, px, off);
                    py += __shfl_xor_sync(0xffffffffu, py, off);
                    pa += __shfl_xor_sync(0xffffffffu, pa, off);
                }
                if (lane == 0) {
                    grad_loc[(pair * L * P + s) * 2] = (loc_t)((acc_t)W * px);
                    grad_loc[(pair * L * P + s) * 2 + 1] = (loc_t)((acc_t)H * py);
                    grad_attn[pair * L * P + s] = (loc_t)pa;
                }
            }
        }
    }
}

// fp32 accumulation buffer -> 16-bit grad_value
template <typename VT>
__global__ void __launch_bounds__(256)
msda_cast_accum_kernel(const float* __restrict__ src, VT* __restrict__ dst, long long count)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long vec = count / 8;
    for (long long i = i0; i < vec; i += stride) {
        float f[8];
        const uint4 a = ldg_stream_v4(src + i * 8), b = ldg_stream_v4(src + i * 8 + 4);
        unpack<float>(a, f);
        unpack<float>(b, f + 4);
        stg_stream_v4(dst + i * 8, pack<VT>(f));
    }
    for (long long i = vec * 8 + i0; i < count; i += stride) dst[i] = from_f32<VT>(src[i]);
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
template <typename VT, int D>
static cudaError_t launch_bwd_fast(const BwdArgs& a, float* accum, cudaStream_t stream,
                                   const unsigned char* red_levels = nullptr)
{
    constexpr int G = D / 4;
    constexpr int PAIRS = 32 / G;
    constexpr int WARPS = BwdWarps<PAIRS>::value;
    const long long total_pairs = (long long)a.N * a.Lq * a.M;
    const long long nq_total = (long long)a.N * a.Lq;
    const long long blocks = ((nq_total + WARPS * PAIRS - 1) / (WARPS * PAIRS)) * a.M;
    if (blocks > 0x7fffffffll) return cudaErrorInvalidConfiguration;
    const int p_magic = (65536 + a.P - 1) / a.P;
    SampleSrc src;
    src.loc = a.loc; src.attn = a.attn; src.ref = nullptr; src.loc_stride = 0; src.attn_stride = 0; src.ref_dim = 0;
    GradDst dst;
    dst.loc = a.grad_loc; dst.attn = a.grad_attn; dst.ref = nullptr;
    msda_bwd_fast_kernel<VT, D, false, float><<<(unsigned)blocks, WARPS * 32, 0, stream>>>(
        (const VT*)a.grad_out, (const VT*)a.value, a.shapes, a.lsi, src, accum, dst,
        a.S, a.M, a.L, a.Lq, a.P, p_magic, total_pairs, red_levels);
    return cudaGetLastError();
}

template <typename VT, int D, typename RT>
static cudaError_t launch_bwd_fused(const FusedArgs& a, float* accum, cudaStream_t stream)
{
    constexpr int G = D / 4;
    constexpr int PAIRS = 32 / G;
    constexpr int WARPS = BwdWarps<PAIRS>::value;
    const long long total_pairs = (long long)a.N * a.Lq * a.M;
    const long long nq_total = (long long)a.N * a.Lq;
    const long long blocks = ((nq_total + WARPS * PAIRS - 1) / (WARPS * PAIRS)) * a.M;
    if (blocks > 0x7fffffffll) return cudaErrorInvalidConfiguration;
    const int p_magic = (65536 + a.P - 1) / a.P;
    SampleSrc src;
    src.loc = a.offsets; src.attn = a.logits; src.ref = a.ref;
    src.loc_stride = a.off_stride; src.attn_stride = a.logit_stride; src.ref_dim = a.ref_dim;
    GradDst dst;
    dst.loc = a.grad_offsets; dst.attn = a.grad_logits; dst.ref = a.grad_ref;
    msda_bwd_fast_kernel<VT, D, true, RT><<<(unsigned)blocks, WARPS * 32, 0, stream>>>(
        (const VT*)a.grad_out, (const VT*)a.value, a.shapes, a.lsi, src, accum, dst,
        a.S, a.M, a.L, a.Lq, a.P, p_magic, total_pairs, nullptr);
    return cudaGetLastError();
}

template <typename VT>
static cudaError_t launch_bwd_generic(const BwdArgs& a, typename Traits<VT>::acc_t* accum, cudaStream_t stream)
{
    using loc_t = typename Traits<VT>::loc_t;
    const long long total_pairs = (long long)a.N * a.Lq * a.M;
    long long blocks = (total_pairs + 7) / 8;
    if (blocks > (1ll << 30)) blocks = 1ll << 30;
    msda_bwd_generic_kernel<VT><<<(unsigned)blocks, 256, 0, stream>>>(
        (const VT*)a.grad_out, (const VT*)a.value, a.shapes, a.lsi, (const loc_t*)a.loc, (const loc_t*)a.attn,
        accum, (loc_t*)a.grad_loc, (loc_t*)a.grad_attn, a.S, a.M, a.D, a.L, a.Lq, a.P, total_pairs);
    return cudaGetLastError();
}

static bool fast_shape_ok(const BwdArgs& a)
{
    return !a.force_generic && a.L <= kMaxLevelsFast && a.P <= 64 && (long long)a.L * a.P * a.P < 65536 &&
           (long long)a.S * a.M * a.D < (1ll << 30) &&    // 32-bit byte offsets inside one frame
           (long long)a.S < (1ll << 25);                   // merge keys: (frame of the CTA) * S + pixel in 32 bits
}

template <typename VT>
static cudaError_t run_bwd_16or32(const BwdArgs& a, cudaStream_t stream)
{
    constexpr bool k16 = sizeof(VT) == 2;
    const size_t count = (size_t)a.N * a.S * a.M * a.D;
    float* accum = k16 ? a.grad_value_accum : (float*)a.grad_value;
    if (k16 && accum == nullptr) return cudaErrorInvalidValue;
    cudaError_t err = cudaMemsetAsync(accum, 0, count * sizeof(float), stream);
    if (err != cudaSuccess) return err;
    const long long total_pairs = (long long)a.N * a.Lq * a.M;
    if constexpr (std::is_same<VT, __nv_bfloat16>::value) {
        if (!a.no_tc && total_pairs > 0 && fast_shape_ok(a) && tc_backward_supported(a)) {
            // opt-in: grad_value of every (tile, level) whose window fits is accumulated on the tensor cores
            // (msda_tc_backward.cu); the lane-group kernel keeps grad_loc / grad_attn and the levels left over
            unsigned char* red_levels = nullptr;
            err = cudaMallocAsync((void**)&red_levels, (size_t)total_pairs, stream);
            if (err != cudaSuccess) return err;
            err = tc_backward_dv(a, red_levels, stream);
            if (err == cudaSuccess) err = launch_bwd_fast<VT, 32>(a, accum, stream, red_levels);
            const cudaError_t e2 = cudaFreeAsync(red_levels, stream);
            if (err != cudaSuccess) return err;
            if (e2 != cudaSuccess) return e2;
            long long blocks = (long long)((count / 8 + 255) / 256);
            if (blocks < 1) blocks = 1;
            if (blocks > 148 * 16) blocks = 148 * 16;
            msda_cast_accum_kernel<VT><<<(unsigned)blocks, 256, 0, stream>>>(accum, (VT*)a.grad_value, (long long)count);
            return cudaGetLastError();
        }
    }
    if (total_pairs > 0 && a.D > 0) {
        bool done = false;
        if (fast_shape_ok(a)) {
            done = true;
            switch (a.D) {   // G = D/4 lanes per head must be a power of two <= 32
                case 4:   err = launch_bwd_fast<VT, 4>(a, accum, stream); break;
                case 8:   err = launch_bwd_fast<VT, 8>(a, accum, stream); break;
                case 16:  err = launch_bwd_fast<VT, 16>(a, accum, stream); break;
                case 32:  err = launch_bwd_fast<VT, 32>(a, accum, stream); break;
                case 64:  err = launch_bwd_fast<VT, 64>(a, accum, stream); break;
                case 128: err = launch_bwd_fast<VT, 128>(a, accum, stream); break;     // one (query, head) per warp
                default: done = false;
            }
        }
        if (!done) err = launch_bwd_generic<VT>(a, accum, stream);
        if (err != cudaSuccess) return err;
    }
    if (k16 && count > 0) {
        long long blocks = (long long)((count / 8 + 255) / 256);
        if (blocks < 1) blocks = 1;
        if (blocks > 148 * 16) blocks = 148 * 16;
        msda_cast_accum_kernel<VT><<<(unsigned)blocks, 256, 0, stream>>>(accum, (VT*)a.grad_value, (long long)count);
        err = cudaGetLastError();
    }
    return err;
}

template <typename VT, typename RT>
static cudaError_t run_bwd_fused(const FusedArgs& a, cudaStream_t stream)
{
    constexpr bool k16 = sizeof(VT) == 2;
    const size_t count = (size_t)a.N * a.S * a.M * a.D;
    float* accum = k16 ? a.grad_value_accum : (float*)a.grad_value;
    if (k16 && accum == nullptr) return cudaErrorInvalidValue;
    cudaError_t err = cudaMemsetAsync(accum, 0, count * sizeof(float), stream);
    if (err != cudaSuccess) return err;
    if ((long long)a.N * a.Lq * a.M > 0) {
        switch (a.D) {
            case 16: err = launch_bwd_fused<VT, 16, RT>(a, accum, stream); break;
            case 32: err = launch_bwd_fused<VT, 32, RT>(a, accum, stream); break;
            case 64: err = launch_bwd_fused<VT, 64, RT>(a, accum, stream); break;
            default: err = cudaErrorInvalidValue;
        }
        if (err != cudaSuccess) return err;
    }
    if (k16 && count > 0) {
        long long blocks = (long long)((count / 8 + 255) / 256);
        if (blocks < 1) blocks = 1;
        if (blocks > 148 * 16) blocks = 148 * 16;
        msda_cast_accum_kernel<VT><<<(unsigned)blocks, 256, 0, stream>>>(accum, (VT*)a.grad_value, (long long)count);
        err = cudaGetLastError();
    }
    return err;
}

cudaError_t fused_backward(const FusedArgs& a, cudaStream_t stream)
{
    if (!fused_supported(a) || !fused_raw_layout_ok(a)) return cudaErrorInvalidValue;
    if (a.dtype == kF32) return run_bwd_fused<float, float>(a, stream);
    if (a.raw_dtype == kF32) return run_bwd_fused<__nv_bfloat16, float>(a, stream);
    return run_bwd_fused<__nv_bfloat16, __nv_bfloat16>(a, stream);
}

cudaError_t backward(const BwdArgs& a, cudaStream_t stream)
{
    if (a.L == 0 || a.P == 0) {
        // empty sum (the forward writes zeros, msda_forward): grad_value is all zeros, grad_loc / grad_attn have no
        // elements.  Also keeps P = 0 away from the launchers' 65536 / P.
        if (a.dtype < kF32 || a.dtype > kF16) return cudaErrorInvalidValue;
        const size_t esz = a.dtype == kF64 ? 8 : (a.dtype == kF32 ? 4 : 2);
        return cudaMemsetAsync(a.grad_value, 0, (size_t)a.N * a.S * a.M * a.D * esz, stream);
    }
    switch (a.dtype) {
        case kF32:  return run_bwd_16or32<float>(a, stream);
        case kBF16: return run_bwd_16or32<__nv_bfloat16>(a, stream);
        case kF16:  return run_bwd_16or32<__half>(a, stream);
        case kF64: {
            const size_t count = (size_t)a.N * a.S * a.M * a.D;
            cudaError_t err = cudaMemsetAsync(a.grad_value, 0, count * sizeof(double), stream);
            if (err != cudaSuccess) return err;
            if ((long long)a.N * a.Lq * a.M == 0 || a.D == 0) return cudaSuccess;
            return launch_bwd_generic<double>(a, (double*)a.grad_value, stream);
        }
    }
    return cudaErrorInvalidValue;
}

}  // namespace msda
