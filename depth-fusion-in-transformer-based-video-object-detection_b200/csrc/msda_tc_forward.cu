// Forward multi-scale deformable attention on the tensor cores (tcgen05 + TMEM + TMA), bf16 values, 32 channels / head.
//
//   out[n,q,m,:] = sum_{l,p} A[n,q,m,l,p] * bilinear(value_l[n,:,m,:], loc[n,q,m,l,p])      (reference cuh:237-299)
//
// restated per tile of 128 neighbouring queries of one (frame, head) as  Out[128, 32] = sum_levels C_l . V_window_l
// (msda_tc.cuh).  The lane-group gather of msda_forward.cu moves every corner row through the SM's L1 data pipe once
// per (query, corner) and is bound by that pipe (profiles/ncu_summary.json: 96 % for bf16); here a value window enters
// the SM once per tile by TMA and the products run on tcgen05.mma.
//
// One persistent CTA per SM, 18 warps:
//   warps 0-15  512 "build" threads: thread t owns query slot t / 4 and point t % 4.  Per tile: load the sample
//               (location, attention weight), bilinear footprint, bounding boxes of the tile's corner pixels per level
//               (warp redux + shared atomics) -> windows and segments.  Per segment: clear the entries this thread
//               wrote into the C block two segments ago, write the new ones, fence, arrive.  The 4 lanes of a query
//               split the CORNERS BY PIXEL PARITY (x & 1, y & 1): every sample has exactly one corner of each parity,
//               so two lanes never write the same C element and no atomics are needed; the entries of one lane that
//               fall on the same pixel are merged in registers.  Warps 0-3 also read the finished accumulator rows
//               back (tcgen05.ld) and store the output.
//   warp 16     TMA producer: one box {32 channels of the head, BW pixels} per window row into the V block.
//   warp 17     MMA issuer: per segment rows * BW / 16 tcgen05.mma (M 128, N 32, K 16) into the tile's accumulator in
//               tensor memory; tcgen05.commit frees the C block and the V block.
// Tiles whose windows do not fit are gathered by the build threads with plain loads (4 lanes x 16 bytes per corner row,
// the mapping of msda_forward.cu).
#include "msda_launch.h"
#include "msda_tc.cuh"

namespace msda {
namespace tc {

using namespace umma;

constexpr int kFwdVStages = 4;
constexpr int kFwdThreads = kBuildThreads + 64;
constexpr int kFwdSmem = 2 * kCBytes + kFwdVStages * kVBytes + 1024;

struct FwdPlan {
    int bad, nseg_total;
    int bw[kMaxL], rows[kMaxL], rshift[kMaxL], nseg[kMaxL];
    int pix0[kMaxL];     // TMA pixel coordinate of the window's first pixel: n*S + start + y0*W + x0
    int W[kMaxL];
};

struct FwdBars {
    unsigned long long c_full[2], mma_done[2], v_full[kFwdVStages], v_free[kFwdVStages];
    unsigned long long plan_ready[2], plan_free[2], out_ready[2], out_free[2];
};

__global__ void __launch_bounds__(kFwdThreads, 1)
msda_tc_fwd_kernel(const __grid_constant__ Maps maps, const __nv_bfloat16* __restrict__ value,
                   const int64_t* __restrict__ shapes, const int64_t* __restrict__ lsi,
                   const float* __restrict__ loc, const float* __restrict__ attn, __nv_bfloat16* __restrict__ out,
                   int N, int S, int M, int L, int Lq, int P, int value_ld, int want_pyramid)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* sC = smem;                        // 2 x 32 KB
    unsigned char* sV = smem + 2 * kCBytes;          // kFwdVStages x 8 KB
    __shared__ LevelMeta lm;
    __shared__ FwdPlan s_plan[2];
    __shared__ int s_bb[2][kMaxL][4];
    __shared__ __align__(8) FwdBars bars;
    __shared__ unsigned tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        level_meta_init(&lm, shapes, lsi, L, Lq, want_pyramid);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars.c_full[i], kBuildWarps);
            mbar_init(&bars.mma_done[i], 1);
            mbar_init(&bars.plan_ready[i], 1);
            mbar_init(&bars.plan_free[i], 2);
            mbar_init(&bars.out_ready[i], 1);
            mbar_init(&bars.out_free[i], 4);
        }
        for (int i = 0; i < kFwdVStages; ++i) { mbar_init(&bars.v_full[i], 1); mbar_init(&bars.v_free[i], 1); }
        fence_mbar_init();
        for (int i = 0; i < 2 * kMaxL; ++i) {
            s_bb[0][0][4 * i + 0] = 0x7fffffff; s_bb[0][0][4 * i + 1] = 0x7fffffff;
            s_bb[0][0][4 * i + 2] = -1;         s_bb[0][0][4 * i + 3] = -1;
        }
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, 64);
    for (int i = tid; i < 2 * kCBytes / 16; i += kFwdThreads) reinterpret_cast<uint4*>(sC)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (warp == kBuildWarps && lane < kMaxBW / 8) tma_prefetch_desc(&maps.m[lane]);
    fence_proxy_async();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const unsigned tmem = tmem_base_s;
    const int tiles = lm.tiles;
    const long long total_items = (long long)N * tiles * M;

    if (warp == kBuildWarps + 1) {
        // ================================ MMA issuer ================================
        const unsigned idesc = make_idesc(128, kD, 0, 1);
        unsigned long long descA[2], descB[kFwdVStages];
        for (int i = 0; i < 2; ++i) descA[i] = make_desc_sw128(sC + i * kCBytes);
        for (int i = 0; i < kFwdVStages; ++i) descB[i] = make_desc(sV + i * kVBytes, 0, 512, 4);
        unsigned g = 0, good = 0, it = 0;
        for (long long item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
            mbar_wait(&bars.plan_ready[it & 1], (it >> 1) & 1);
            const FwdPlan& pl = s_plan[it & 1];
            if (pl.bad || pl.nseg_total == 0) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars.plan_free[it & 1]);
                continue;
            }
            const unsigned tb = good & 1;
            if (good >= 2) mbar_wait(&bars.out_free[tb], ((good >> 1) - 1) & 1);
            tcgen05_fence_after();
            bool first = true;
            for (int l = 0; l < L; ++l) {
                const int nseg = pl.nseg[l], bw = pl.bw[l], rows = pl.rows[l], rshift = pl.rshift[l];
                for (int sidx = 0; sidx < nseg; ++sidx, ++g) {
                    const unsigned b = g & 1, vs = g % kFwdVStages;
                    int r = min(1 << rshift, rows - (sidx << rshift));
                    if ((bw & 15) && (r & 1)) ++r;
                    const int ksteps = (r * bw) >> 4;
                    mbar_wait(&bars.c_full[b], (g >> 1) & 1);
                    mbar_wait(&bars.v_full[vs], (g / kFwdVStages) & 1);
                    tcgen05_fence_after();
                    if (elect_one()) {
                        for (int ks = 0; ks < ksteps; ++ks)
                            mma_bf16(tmem + tb * kD, desc_advance(descA[b], (unsigned)((ks >> 2) * 16384 + (ks & 3) * 32)),
                                     desc_advance(descB[vs], (unsigned)(ks * 1024)), idesc, !(first && ks == 0));
                        mma_commit(&bars.mma_done[b]);
                        mma_commit(&bars.v_free[vs]);
                    }
                    __syncwarp();
                    first = false;
                }
            }
            if (elect_one()) mma_commit(&bars.out_ready[tb]);
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars.plan_free[it & 1]);
            ++good;
        }
    } else if (warp == kBuildWarps) {
        // ================================ TMA producer ================================
        unsigned g = 0, it = 0;
        for (long long item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
            mbar_wait(&bars.plan_ready[it & 1], (it >> 1) & 1);
            const FwdPlan& pl = s_plan[it & 1];
            if (pl.bad || pl.nseg_total == 0) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars.plan_free[it & 1]);
                continue;
            }
            const int h = (int)(item % M);
            for (int l = 0; l < L; ++l) {
                const int nseg = pl.nseg[l], bw = pl.bw[l], rows = pl.rows[l], rshift = pl.rshift[l];
                const int W = pl.W[l], pix0 = pl.pix0[l];
                const CUtensorMap* map = &maps.m[(bw >> 3) - 1];
                for (int sidx = 0; sidx < nseg; ++sidx, ++g) {
                    const unsigned vs = g % kFwdVStages;
                    int r = min(1 << rshift, rows - (sidx << rshift));
                    if ((bw & 15) && (r & 1)) ++r;
                    if (g >= kFwdVStages) mbar_wait(&bars.v_free[vs], ((g / kFwdVStages) - 1) & 1);
                    if (elect_one()) {
                        mbar_expect_tx(&bars.v_full[vs], (unsigned)(r * bw * 64));
                        for (int rr = 0; rr < r; ++rr)
                            tma_load_2d(sV + vs * kVBytes + rr * bw * 64, map, h * kD,
                                        pix0 + ((sidx << rshift) + rr) * W, &bars.v_full[vs]);
                    }
                    __syncwarp();
                }
            }
            if (lane == 0) mbar_arrive(&bars.plan_free[it & 1]);
        }
    } else {
        // ================================ build threads ================================
        const int q = tid >> 2, j = tid & 3;          // query slot, point (= corner parity class)
        const int px = j & 1, py = j >> 1;
        const unsigned row_base = c_row_base(q);
        const int q7 = q & 7;
        const unsigned sC_u32 = smem_u32(sC);
        unsigned pend_cur0 = 0xffffffffu, pend_cur1 = 0xffffffffu;     // offsets written into the C block of parity g & 1 ...
        unsigned pend_oth0 = 0xffffffffu, pend_oth1 = 0xffffffffu;     // ... and of the other parity
        unsigned g = 0, good = 0, it = 0;
        const int LP = L * P;
        for (long long item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
            const int h = (int)(item % M);
            const long long rest = item / M;
            const int t = (int)(rest % tiles);
            const int n = (int)(rest / tiles);
            const Tile tl = tile_decode(lm, t, L, Lq);
            const int qi = tile_query(tl, q);
            const bool have = qi >= 0 && j < P;
            const long long pair = ((long long)n * Lq + (qi >= 0 ? qi : 0)) * M + h;
            // ---- this thread's sample on every level ----
            int fx[kMaxL], fy[kMaxL];
            float flw[kMaxL], flh[kMaxL], fa[kMaxL];
            unsigned inside = 0;
            const int pb = it & 1;
#pragma unroll
            for (int l = 0; l < kMaxL; ++l) {
                fx[l] = 0; fy[l] = 0; flw[l] = 0.f; flh[l] = 0.f; fa[l] = 0.f;
                if (l < L && have) {
                    const float2 xy = ldg_stream_f32x2(loc + (pair * LP + l * P + j) * 2);
                    fa[l] = ldg_stream_f32(attn + pair * LP + l * P + j);
                    const int H = lm.H[l], W = lm.W[l];
                    const float w_im = xy.x * (float)W - 0.5f, h_im = xy.y * (float)H - 0.5f;      // cuh:285-286
                    const bool in = (h_im > -1.f) && (w_im > -1.f) && (h_im < (float)H) && (w_im < (float)W);
                    const float hf = floorf(h_im), wf = floorf(w_im);
                    fx[l] = (int)wf; fy[l] = (int)hf;
                    flw[l] = w_im - wf; flh[l] = h_im - hf;
                    if (in) inside |= 1u << l;
                }
            }
            // bounding box of the corner pixels per level: warp reduction, then one shared atomic per warp
#pragma unroll
            for (int l = 0; l < kMaxL; ++l) {
                if (l < L) {
                    const bool in = (inside >> l) & 1u;
                    const int W = lm.W[l], H = lm.H[l];
                    const int xa = in ? max(fx[l], 0) : 0x7fffffff, ya = in ? max(fy[l], 0) : 0x7fffffff;
                    const int xb = in ? min(fx[l] + 1, W - 1) : -1, yb = in ? min(fy[l] + 1, H - 1) : -1;
                    const int mnx = __reduce_min_sync(0xffffffffu, xa), mny = __reduce_min_sync(0xffffffffu, ya);
                    const int mxx = __reduce_max_sync(0xffffffffu, xb), mxy = __reduce_max_sync(0xffffffffu, yb);
                    if (lane == 0 && mxx >= 0) {
                        atomicMin(&s_bb[pb][l][0], mnx); atomicMin(&s_bb[pb][l][1], mny);
                        atomicMax(&s_bb[pb][l][2], mxx); atomicMax(&s_bb[pb][l][3], mxy);
                    }
                }
            }
            named_bar_sync(1, kBuildThreads);
            Window win[kMaxL];
            bool bad = false;
            int nseg_total = 0;
#pragma unroll
            for (int l = 0; l < kMaxL; ++l) {
                win[l].nseg = 0;
                if (l < L) {
                    bad |= !window_from_bbox(s_bb[pb][l][0], s_bb[pb][l][1], s_bb[pb][l][2], s_bb[pb][l][3], &win[l]);
                    nseg_total += win[l].nseg;
                }
            }
            if (tid == 0) {
                // both control warps are done with the plan of the tile before the previous one
                if (it >= 2) mbar_wait(&bars.plan_free[pb], ((it >> 1) - 1) & 1);
                FwdPlan& pl = s_plan[pb];
                pl.bad = bad; pl.nseg_total = nseg_total;
#pragma unroll
                for (int l = 0; l < kMaxL; ++l) {
                    if (l < L) {
                        pl.bw[l] = win[l].bw; pl.rows[l] = win[l].rows; pl.rshift[l] = win[l].rshift; pl.nseg[l] = win[l].nseg;
                        pl.W[l] = lm.W[l];
                        pl.pix0[l] = n * S + lm.start[l] + win[l].y0 * lm.W[l] + win[l].x0;
                    }
                }
                mbar_arrive(&bars.plan_ready[pb]);
            }
            // every thread has read the boxes: reset them for the tile after the next one
            named_bar_sync(2, kBuildThreads);
            if (tid < 4 * kMaxL) s_bb[pb][0][tid] = (tid & 2) ? -1 : 0x7fffffff;

            if (bad) {
                // ---- gather with plain loads: 4 lanes x 16 bytes per corner row ----
                float acc[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[c] = 0.f;
                const __nv_bfloat16* vbase = value + (long long)n * S * value_ld + h * kD + j * 8;
#pragma unroll
                for (int l = 0; l < kMaxL; ++l) {
                    if (l < L) {
                        const int W = lm.W[l], H = lm.H[l], start = lm.start[l];
                        for (int s = 0; s < P; ++s) {
                            const int bx = __shfl_sync(0xffffffffu, fx[l], s, 4), by = __shfl_sync(0xffffffffu, fy[l], s, 4);
                            const float lw = __shfl_sync(0xffffffffu, flw[l], s, 4), lh = __shfl_sync(0xffffffffu, flh[l], s, 4);
                            const float a = __shfl_sync(0xffffffffu, fa[l], s, 4);
                            const unsigned in = __shfl_sync(0xffffffffu, inside, s, 4) & (1u << l);
                            if (!in) continue;
                            const float hw = 1.f - lw, hh = 1.f - lh;
                            const float wk[4] = {hh * hw * a, hh * lw * a, lh * hw * a, lh * lw * a};
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const int x = bx + (k & 1), y = by + (k >> 1);
                                if (x >= 0 && x < W && y >= 0 && y < H) {
                                    float v[8];
                                    unpack<__nv_bfloat16>(ldg_v4(vbase + (long long)(start + y * W + x) * value_ld), v);
#pragma unroll
                                    for (int c = 0; c < 8; ++c) acc[c] = fmaf(wk[k], v[c], acc[c]);
                                }
                            }
                        }
                    }
                }
                if (qi >= 0) stg_stream_v4(out + pair * kD + j * 8, pack<__nv_bfloat16>(acc));
                continue;
            }

            // ---- segments ----
            unsigned gl = g;                  // global index of the level's first segment
#pragma unroll
            for (int l = 0; l < kMaxL; ++l) {
                if (l < L && win[l].nseg > 0) {
                    const Window w = win[l];
                    const int W = lm.W[l], H = lm.H[l];
                    // my corner (the one of pixel parity (px, py)) of each of the query's samples on this level
                    unsigned e_off[kMaxP], e_seg[kMaxP];
                    float e_c[kMaxP];
                    bool e_ok[kMaxP];
#pragma unroll
                    for (int s = 0; s < kMaxP; ++s) {
                        const int bx = __shfl_sync(0xffffffffu, fx[l], s, 4), by = __shfl_sync(0xffffffffu, fy[l], s, 4);
                        const float lw = __shfl_sync(0xffffffffu, flw[l], s, 4), lh = __shfl_sync(0xffffffffu, flh[l], s, 4);
                        const float a = __shfl_sync(0xffffffffu, fa[l], s, 4);
                        const unsigned in = __shfl_sync(0xffffffffu, inside, s, 4) & (1u << l);
                        const int x = bx + ((bx ^ px) & 1), y = by + ((by ^ py) & 1);
                        const float wx = x == bx ? 1.f - lw : lw, wy = y == by ? 1.f - lh : lh;
                        e_ok[s] = in && x >= 0 && x < W && y >= 0 && y < H;
                        e_c[s] = wy * wx * a;
                        const int yrel = y - w.y0;
                        const int k = (yrel & ((1 << w.rshift) - 1)) * w.bw + (x - w.x0);
                        e_seg[s] = gl + (unsigned)(yrel >> w.rshift);
                        e_off[s] = c_offset(row_base, q7, k);
                    }
                    // samples of this lane that share a pixel: one entry with the summed coefficient
#pragma unroll
                    for (int a2 = 0; a2 < kMaxP; ++a2)
#pragma unroll
                        for (int b2 = a2 + 1; b2 < kMaxP; ++b2)
                            if (e_ok[a2] && e_ok[b2] && e_off[a2] == e_off[b2] && e_seg[a2] == e_seg[b2]) {
                                e_c[a2] += e_c[b2];
                                e_ok[b2] = false;
                            }
                    for (int sidx = 0; sidx < w.nseg; ++sidx, ++g) {
                        const unsigned b = g & 1;
                        if (g >= 2) mbar_wait(&bars.mma_done[b], ((g >> 1) - 1) & 1);
                        const unsigned cb = sC_u32 + b * kCBytes;
                        // clear what this thread wrote into this block two segments ago
                        if ((pend_cur0 & 0xffffu) != 0xffffu) sts_u16(cb + (pend_cur0 & 0xffffu), 0);
                        if ((pend_cur0 >> 16) != 0xffffu) sts_u16(cb + (pend_cur0 >> 16), 0);
                        if ((pend_cur1 & 0xffffu) != 0xffffu) sts_u16(cb + (pend_cur1 & 0xffffu), 0);
                        if ((pend_cur1 >> 16) != 0xffffu) sts_u16(cb + (pend_cur1 >> 16), 0);
                        unsigned po[kMaxP];
#pragma unroll
                        for (int s = 0; s < kMaxP; ++s) {
                            const bool wr = e_ok[s] && e_seg[s] == g;
                            if (wr) sts_u16(cb + e_off[s], __bfloat16_as_ushort(__float2bfloat16_rn(e_c[s])));
                            po[s] = wr ? e_off[s] : 0xffffu;
                        }
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&bars.c_full[b]);
                        // the other block is the current one of the next segment
                        pend_cur0 = pend_oth0; pend_cur1 = pend_oth1;
                        pend_oth0 = po[0] | (po[1] << 16); pend_oth1 = po[2] | (po[3] << 16);
                    }
                    gl = g;
                }
            }
            // ---- epilogue: accumulator rows -> output ----
            if (nseg_total > 0) {
                const unsigned tb = good & 1;
                if (warp < 4) {
                    mbar_wait(&bars.out_ready[tb], (good >> 1) & 1);
                    tcgen05_fence_after();
                    float v[32];
                    tmem_ld32(tmem + ((unsigned)(warp * 32) << 16) + tb * kD, v);
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars.out_free[tb]);
                    const int qe = tile_query(tl, tid);
                    if (qe >= 0) {
                        __nv_bfloat16* o = out + (((long long)n * Lq + qe) * M + h) * kD;
#pragma unroll
                        for (int c = 0; c < 4; ++c) stg_stream_v4(o + c * 8, pack<__nv_bfloat16>(v + c * 8));
                    }
                }
                ++good;
            } else if (warp < 4) {
                const int qe = tile_query(tl, tid);
                if (qe >= 0) {
                    __nv_bfloat16* o = out + (((long long)n * Lq + qe) * M + h) * kD;
#pragma unroll
                    for (int c = 0; c < 4; ++c) stg_stream_v4(o + c * 8, make_uint4(0u, 0u, 0u, 0u));
                }
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free(tmem, 64);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tc_encode_fn()
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

// value [pixels, ld] bf16: boxes {32 channels, 8 (i + 1) pixels}, SWIZZLE_64B
bool make_value_maps(Maps* maps, const void* value, long long pixels, long long ld)
{
    EncodeTiledFn fn = tc_encode_fn();
    if (fn == nullptr) return false;
    for (int i = 0; i < kMaxBW / 8; ++i) {
        const cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)pixels};
        const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
        const cuuint32_t box[2] = {(cuuint32_t)kD, (cuuint32_t)(8 * (i + 1))};
        const cuuint32_t estr[2] = {1, 1};
        if (fn(&maps->m[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(value), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
    }
    return true;
}

}  // namespace tc

bool tc_forward_supported(const FwdArgs& a)
{
    return a.dtype == kBF16 && a.D == tc::kD && a.L >= 1 && a.L <= tc::kMaxL && a.P >= 1 && a.P <= tc::kMaxP &&
           !a.force_generic && (long long)a.N * a.S < (1ll << 31) && ((size_t)a.value % 16) == 0 &&
           (long long)a.N * a.Lq >= 2048;
}

cudaError_t tc_forward(const FwdArgs& a, cudaStream_t stream)
{
    if (!tc_forward_supported(a)) return cudaErrorInvalidValue;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    static bool configured[64] = {};
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(tc::msda_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kFwdSmem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    alignas(64) tc::Maps maps;
    if (!tc::make_value_maps(&maps, a.value, (long long)a.N * a.S, (long long)a.M * a.D)) return cudaErrorNotSupported;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    tc::msda_tc_fwd_kernel<<<sms, tc::kFwdThreads, tc::kFwdSmem, stream>>>(
        maps, (const __nv_bfloat16*)a.value, a.shapes, a.lsi, (const float*)a.loc, (const float*)a.attn,
        (__nv_bfloat16*)a.out, a.N, a.S, a.M, a.L, a.Lq, a.P, a.M * a.D, 1);
    return cudaGetLastError();
}

}  // namespace msda
