// Forward multi-scale deformable attention on the tensor cores (tcgen05 + TMEM + TMA), bf16 values, 32 channels / head.
//
//   out[n,q,m,:] = sum_{l,p} A[n,q,m,l,p] * bilinear(value_l[n,:,m,:], loc[n,q,m,l,p])      (reference cuh:237-299)
//
// restated per tile of 128 neighbouring queries of one (frame, head) as  Out[128, 32] = sum_levels C_l . V_window_l
// (msda_tc.cuh).  The lane-group gather of msda_forward.cu moves every corner row through the SM's L1 data pipe once
// per (query, corner) and is bound by that pipe (profiles/ncu_summary.json: 96 % for bf16); here a value window enters
// the SM once per tile by TMA and the products run on tcgen05.mma.
//
// Small persistent CTAs, four per SM (their phases interleave on the SM: while one builds, another's MMAs run):
//   warps 0-3   128 "build" threads, thread t = query slot t = row t of C = lane t of the accumulator in tensor memory.
//               Level by level: the query's samples (location, attention weight) -> bilinear footprints -> bounding
//               box of the tile's corner pixels (warp redux + shared atomics) -> the level's window and its segments,
//               published to the control warps.  Per segment: add this query's coefficients into its row of the C
//               block (the row is private to the thread, so duplicates -- samples sharing a pixel -- are summed with
//               plain shared-memory read-modify-writes), fence, arrive; when the segment's MMAs have completed the same
//               entries are zeroed again.  After the last level: accumulator row -> output.
//   warp 4      TMA producer: one box {32 channels of the head, BW pixels} per window row into a V block.
//   warp 5      MMA issuer: per segment rows * BW / 16 tcgen05.mma (M 128, N 32, K 16) into the tile's accumulator;
//               tcgen05.commit frees the C block and the V block.
// The control warps follow a stream of level plans (shared-memory ring), so they never decode tiles themselves.
// A tile with a level whose window does not fit (scattered sampling locations) is appended to a list and gathered by
// msda_tc_fwd_fallback_kernel with plain loads (4 lanes x 16 bytes per corner row, the mapping of msda_forward.cu), so
// the result is exact for any input.
#include <cstdio>
#include <cstdlib>
#include "msda_launch.h"
#include "msda_tc.cuh"

namespace msda {
namespace tc {

using namespace umma;

constexpr int kFwdBuild = kTileQ;
constexpr int kFwdThreads = kFwdBuild + 64;
constexpr int kFwdVStages = 2;
constexpr int kFwdSmem = kCBytes + kFwdVStages * kVBytes + 1024;
constexpr int kLpRing = 4;
#ifndef MSDA_TC_FWD_CTAS
#define MSDA_TC_FWD_CTAS 3       // resident CTAs per SM the register budget must allow
#endif
enum { kLpTileStart = 1, kLpTileEnd = 2, kLpEnd = 4 };

struct LevelPlan {
    int flags, bw, rows, rshift, nseg;
    int pix0;        // TMA pixel coordinate of the window's first pixel: n*S + start + y0*W + x0
    int W, head;
};

struct FwdBars {
    unsigned long long c_full, mma_done, out_ready;
    unsigned long long v_full[kFwdVStages], v_free[kFwdVStages];
    unsigned long long lp_ready[kLpRing], lp_free[kLpRing];
};

__global__ void __launch_bounds__(kFwdThreads, MSDA_TC_FWD_CTAS)
msda_tc_fwd_kernel(const __grid_constant__ Maps maps, const int64_t* __restrict__ shapes, const int64_t* __restrict__ lsi,
                   const float* __restrict__ loc, const float* __restrict__ attn, __nv_bfloat16* __restrict__ out,
                   int N, int S, int M, int L, int Lq, int P, int want_pyramid, int* __restrict__ bad_list, int bad_cap,
                   unsigned long long* __restrict__ trace)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* sC = smem;                    // 32 KB
    unsigned char* sV = smem + kCBytes;          // kFwdVStages x 8 KB
    __shared__ LevelMeta lm;
    __shared__ LevelPlan s_lp[kLpRing];
    __shared__ int s_bb[3][4];
    __shared__ __align__(8) FwdBars bars;
    __shared__ unsigned tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        level_meta_init(&lm, shapes, lsi, L, Lq, want_pyramid);
        mbar_init(&bars.c_full, kFwdBuild / 32);
        mbar_init(&bars.mma_done, 1);
        mbar_init(&bars.out_ready, 1);
        for (int i = 0; i < kFwdVStages; ++i) { mbar_init(&bars.v_full[i], 1); mbar_init(&bars.v_free[i], 1); }
        for (int i = 0; i < kLpRing; ++i) { mbar_init(&bars.lp_ready[i], 1); mbar_init(&bars.lp_free[i], 2); }
        fence_mbar_init();
        for (int i = 0; i < 3; ++i) { s_bb[i][0] = 0x7fffffff; s_bb[i][1] = 0x7fffffff; s_bb[i][2] = -1; s_bb[i][3] = -1; }
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, 32);
    // zero the C block (its invariant between segments) and the V blocks (rows no TMA box has written yet must be finite)
    for (int i = tid; i < (kCBytes + kFwdVStages * kVBytes) / 16; i += kFwdThreads)
        reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (warp == 4 && lane < kMaxBW / 8) tma_prefetch_desc(&maps.m[lane]);
    fence_proxy_async();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const unsigned tmem = tmem_base_s;
    // optional cycle accounting (compile with -DMSDA_TC_TRACE, run with MSDA_TC_TRACE=1): one representative thread
    // per role accumulates where its time goes
#ifdef MSDA_TC_TRACE
    unsigned long long tr[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long tr_t = 0;
    const bool tracing = trace != nullptr && lane == 0 && (warp == 0 || warp >= 4);
#define TR_START() do { if (tracing) tr_t = clock64(); } while (0)
#define TR_ADD(i) do { if (tracing) { const long long t_ = clock64(); tr[i] += (unsigned long long)(t_ - tr_t); tr_t = t_; } } while (0)
#define TR_COUNT(i, v) do { if (tracing) tr[i] += (v); } while (0)
#else
#define TR_START() do { } while (0)
#define TR_ADD(i) do { } while (0)
#define TR_COUNT(i, v) do { } while (0)
#endif

    if (warp == 5) {
        // ================================ MMA issuer ================================
        const unsigned idesc = make_idesc(128, kD, 0, 1);
        const unsigned long long descA = make_desc_sw128(sC);
        unsigned long long descB[kFwdVStages];
#pragma unroll
        for (int i = 0; i < kFwdVStages; ++i) descB[i] = make_desc(sV + i * kVBytes, 0, 512, 4);
        unsigned gseg = 0;
        bool first = true;
        TR_START();
        for (unsigned u = 0;; ++u) {
            mbar_wait(&bars.lp_ready[u % kLpRing], (u / kLpRing) & 1);
            TR_ADD(0);
            const LevelPlan pl = s_lp[u % kLpRing];
            if (pl.flags & kLpEnd) break;
            if (pl.flags & kLpTileStart) first = true;
            for (int sidx = 0; sidx < pl.nseg; ++sidx, ++gseg) {
                const unsigned vs = gseg % kFwdVStages;
                int r = min(1 << pl.rshift, pl.rows - (sidx << pl.rshift));
                if ((pl.bw & 15) && (r & 1)) ++r;
                const int ksteps = (r * pl.bw) >> 4;
                mbar_wait(&bars.c_full, gseg & 1);
                TR_ADD(1);
                mbar_wait(&bars.v_full[vs], (gseg / kFwdVStages) & 1);
                TR_ADD(2);
                tcgen05_fence_after();
                TR_ADD(8);
                if (elect_one()) {
                    for (int ks = 0; ks < ksteps; ++ks)
                        mma_bf16(tmem, desc_advance(descA, (unsigned)((ks >> 2) * 16384 + (ks & 3) * 32)),
                                 desc_advance(descB[vs], (unsigned)(ks * 1024)), idesc, !(first && ks == 0));
                    TR_ADD(9);
                    mma_commit(&bars.mma_done);
                    mma_commit(&bars.v_free[vs]);
                }
                __syncwarp();
                first = false;
                TR_ADD(3);
                TR_COUNT(4, ksteps); TR_COUNT(5, 1);
            }
            if (pl.flags & kLpTileEnd) {
                if (elect_one()) mma_commit(&bars.out_ready);
                __syncwarp();
            }
            if (lane == 0) mbar_arrive(&bars.lp_free[u % kLpRing]);
        }
    } else if (warp == 4) {
        // ================================ TMA producer ================================
        unsigned gseg = 0;
        TR_START();
        for (unsigned u = 0;; ++u) {
            mbar_wait(&bars.lp_ready[u % kLpRing], (u / kLpRing) & 1);
            TR_ADD(0);
            const LevelPlan pl = s_lp[u % kLpRing];
            if (pl.flags & kLpEnd) break;
            const CUtensorMap* map = &maps.m[(pl.bw >> 3) - 1];
            for (int sidx = 0; sidx < pl.nseg; ++sidx, ++gseg) {
                const unsigned vs = gseg % kFwdVStages;
                int r = min(1 << pl.rshift, pl.rows - (sidx << pl.rshift));
                if ((pl.bw & 15) && (r & 1)) ++r;
                if (gseg >= kFwdVStages) mbar_wait(&bars.v_free[vs], ((gseg / kFwdVStages) - 1) & 1);
                TR_ADD(1);
                if (elect_one()) {
                    mbar_expect_tx(&bars.v_full[vs], (unsigned)(r * pl.bw * 64));
                    for (int rr = 0; rr < r; ++rr)
                        tma_load_2d(sV + vs * kVBytes + rr * pl.bw * 64, map, pl.head * kD,
                                    pl.pix0 + ((sidx << pl.rshift) + rr) * pl.W, &bars.v_full[vs]);
                }
                __syncwarp();
                TR_ADD(2);
            }
            if (lane == 0) mbar_arrive(&bars.lp_free[u % kLpRing]);
        }
    } else {
        // ================================ build threads ================================
        const int q = tid;
        const unsigned row_base = c_row_base(q);
        const int q7 = q & 7;
        const unsigned cb = smem_u32(sC);
        const int tiles = lm.tiles;
        const int total_items = N * tiles * M;
        const int LP = L * P;
        unsigned u = 0;              // level plans published
        unsigned waited = 0;         // segments whose MMAs this thread has waited for
        unsigned items_done = 0;

        auto publish = [&](int flags, const Window& w, int pix0, int Wl, int head) {
            if (tid == 0) {
                if (u >= kLpRing) mbar_wait(&bars.lp_free[u % kLpRing], ((u / kLpRing) - 1) & 1);
                LevelPlan& pl = s_lp[u % kLpRing];
                pl.flags = flags; pl.bw = w.bw; pl.rows = w.rows; pl.rshift = w.rshift; pl.nseg = w.nseg;
                pl.pix0 = pix0; pl.W = Wl; pl.head = head;
                mbar_arrive(&bars.lp_ready[u % kLpRing]);
            }
        };
        auto decode = [&](int it, int& h_, int& n_, int& qi_, int& pair_) {
            h_ = it % M;
            const int rest = it / M;
            const Tile tl = tile_decode(lm, rest % tiles, L, Lq);
            n_ = rest / tiles;
            qi_ = tile_query(tl, q);
            pair_ = (n_ * Lq + (qi_ >= 0 ? qi_ : 0)) * M + h_;
        };

        int item = blockIdx.x;
        int h = 0, n = 0, qi = -1;
        int pair = 0;
        LevelSamples pf;             // prefetched samples of the next (item, level)
        if (item < total_items) {
            decode(item, h, n, qi, pair);
            load_level(pf, loc, attn, pair, LP, 0, P, qi >= 0);
        }
        TR_START();
        while (item < total_items) {
            bool bad = false;
            int nseg_total = 0;
            // next item (its level-0 samples are prefetched during this item's last level)
            const int item_next = item + gridDim.x;
            int h2 = 0, n2 = 0, qi2 = -1;
            int pair2 = 0;
            if (item_next < total_items) decode(item_next, h2, n2, qi2, pair2);

            for (int l = 0; l < L; ++l, ++u) {
                const LevelSamples ls = pf;
                if (l + 1 < L) load_level(pf, loc, attn, pair, LP, l + 1, P, qi >= 0);
                else if (item_next < total_items) load_level(pf, loc, attn, pair2, LP, 0, P, qi2 >= 0);
                const int H = lm.H[l], W = lm.W[l];
                // ---- footprints and the bounding box of the tile's corner pixels on this level ----
                Footprints fp;
                footprints_of(ls, H, W, fp);
                int* bb = s_bb[u % 3];
                bbox_merge(fp, H, W, lane, bb);
                TR_ADD(0);
                named_bar_sync(1, kFwdBuild);
                TR_ADD(1);
                Window w;
                const bool fits = window_from_bbox(bb[0], bb[1], bb[2], bb[3], &w);
                if (tid == 0) {          // the box of the previous level: everybody has read it before this barrier
                    int* pb = s_bb[(u + 2) % 3];
                    pb[0] = 0x7fffffff; pb[1] = 0x7fffffff; pb[2] = -1; pb[3] = -1;
                }
                if (!fits) {
                    w.nseg = 0;
                    publish((l == 0 ? kLpTileStart : 0) | kLpTileEnd, w, 0, W, h);
                    bad = true;
                    ++u;
                    // the prefetch holds the next level of THIS item: replace it by the next item's first level
                    if (l + 1 < L && item_next < total_items) load_level(pf, loc, attn, pair2, LP, 0, P, qi2 >= 0);
                    break;
                }
                publish((l == 0 ? kLpTileStart : 0) | (l == L - 1 ? kLpTileEnd : 0), w,
                        n * S + lm.start[l] + w.y0 * W + w.x0, W, h);
                if (w.nseg == 0) continue;
                nseg_total += w.nseg;
                // ---- this query's entries of the level ----
                SampleEntries en[kMaxP];
#pragma unroll
                for (int s = 0; s < kMaxP; ++s) sample_entries(fp, s, ls.a[s], w, H, W, row_base, q7, en[s]);
                TR_ADD(2);
                for (int sidx = 0; sidx <= w.nseg; ++sidx) {
                    if (sidx > 0) {
                        // the MMAs of segment sidx - 1 have read the block: zero what this thread put there
                        mbar_wait(&bars.mma_done, waited & 1);
                        TR_ADD(3);
                        ++waited;
#pragma unroll
                        for (int s = 0; s < kMaxP; ++s) {      // predicated, not branched: the rows of a warp's queries differ
                            c_row_clear(cb, (en[s].segs & 0xffu) == (unsigned)(sidx - 1) ? en[s].off01 : 0xffffffffu);
                            c_row_clear(cb, (en[s].segs >> 8) == (unsigned)(sidx - 1) ? en[s].off23 : 0xffffffffu);
                        }
                        TR_ADD(8);
                        if (sidx == w.nseg) break;
                    }
                    // read-modify-write: the row is private to this thread; samples that share a pixel follow each other
#pragma unroll
                    for (int s = 0; s < kMaxP; ++s) {
                        c_row_add(cb, (en[s].segs & 0xffu) == (unsigned)sidx ? en[s].off01 : 0xffffffffu, en[s].cf01);
                        c_row_add(cb, (en[s].segs >> 8) == (unsigned)sidx ? en[s].off23 : 0xffffffffu, en[s].cf23);
                    }
                    TR_ADD(9);
                    fence_proxy_async();
                    TR_ADD(10);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars.c_full);
                    TR_ADD(4);
                }
            }
            // ---- accumulator row -> output ----
            TR_ADD(5);
            mbar_wait(&bars.out_ready, items_done & 1);
            TR_ADD(6);
            ++items_done;
            tcgen05_fence_after();
            if (!bad) {
                if (nseg_total > 0) {
                    float v[32];
                    tmem_ld32(tmem + ((unsigned)(warp * 32) << 16), v);
                    if (qi >= 0) {
                        __nv_bfloat16* o = out + (long long)pair * kD;
#pragma unroll
                        for (int c = 0; c < 4; ++c) stg_stream_v4(o + c * 8, pack<__nv_bfloat16>(v + c * 8));
                    }
                } else if (qi >= 0) {
                    __nv_bfloat16* o = out + (long long)pair * kD;
#pragma unroll
                    for (int c = 0; c < 4; ++c) stg_stream_v4(o + c * 8, make_uint4(0u, 0u, 0u, 0u));
                }
            } else if (tid == 0) {
                const int slot = atomicAdd(bad_list, 1);
                if (slot < bad_cap) bad_list[1 + slot] = item;
            }
            tcgen05_fence_before();
            item = item_next; h = h2; n = n2; qi = qi2; pair = pair2;
            TR_ADD(7);
        }
        Window wend;
        wend.bw = 8; wend.rows = 0; wend.rshift = 4; wend.nseg = 0; wend.x0 = 0; wend.y0 = 0;
        publish(kLpEnd, wend, 0, 0, 0);
    }
#ifdef MSDA_TC_TRACE
    if (tracing && blockIdx.x < 64) {
        const int role = warp == 0 ? 0 : warp - 3;          // 0 build, 1 producer, 2 issuer
#pragma unroll
        for (int i = 0; i < 12; ++i) trace[(blockIdx.x * 3 + role) * 12 + i] = tr[i];
    }
#endif
#undef TR_START
#undef TR_ADD
#undef TR_COUNT
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free(tmem, 32);
}

// Tiles the tensor-core kernel could not take: 4 lanes per query (16 bytes of the head's 64-byte row each), 8 queries
// per warp, every lane walks all samples of its query.
__global__ void __launch_bounds__(128)
msda_tc_fwd_fallback_kernel(const __nv_bfloat16* __restrict__ value, const int64_t* __restrict__ shapes,
                            const int64_t* __restrict__ lsi, const float* __restrict__ loc, const float* __restrict__ attn,
                            __nv_bfloat16* __restrict__ out, int N, int S, int M, int L, int Lq, int P, int value_ld,
                            int want_pyramid, const int* __restrict__ bad_list, int bad_cap)
{
    __shared__ LevelMeta lm;
    const int count = min(bad_list[0], bad_cap);
    if (count == 0) return;
    if (threadIdx.x == 0) level_meta_init(&lm, shapes, lsi, L, Lq, want_pyramid);
    __syncthreads();
    const int tiles = lm.tiles;
    const int LP = L * P;
    const int slot_in_cta = threadIdx.x >> 2, j = threadIdx.x & 3;
    for (long long unit = blockIdx.x; unit < (long long)count * 4; unit += gridDim.x) {
        const long long item = bad_list[1 + unit / 4];
        const int h = (int)(item % M);
        const long long rest = item / M;
        const Tile tl = tile_decode(lm, (int)(rest % tiles), L, Lq);
        const int n = (int)(rest / tiles);
        const int qi = tile_query(tl, (int)(unit % 4) * 32 + slot_in_cta);
        if (qi < 0) continue;
        const long long pair = ((long long)n * Lq + qi) * M + h;
        const __nv_bfloat16* vbase = value + (long long)n * S * value_ld + h * kD + j * 8;
        float acc[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = 0.f;
        for (int l = 0; l < L; ++l) {
            const int H = lm.H[l], W = lm.W[l], start = lm.start[l];
            for (int s = 0; s < P; ++s) {
                const float2 xy = *reinterpret_cast<const float2*>(loc + (pair * LP + l * P + s) * 2);
                const float a = attn[pair * LP + l * P + s];
                const Footprint f = footprint<float>(xy.x, xy.y, H, W, start);
                if (!f.ok) continue;
                const float hw = 1.f - f.lw, hh = 1.f - f.lh;
                const float wk[4] = {hh * hw * a, hh * f.lw * a, f.lh * hw * a, f.lh * f.lw * a};
                const int pix[4] = {f.pix00, f.pix00 + 1, f.pix00 + W, f.pix00 + W + 1};
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (f.ok & (1u << k)) {
                        float v[8];
                        unpack<__nv_bfloat16>(ldg_v4(vbase + (long long)pix[k] * value_ld), v);
#pragma unroll
                        for (int c = 0; c < 8; ++c) acc[c] = fmaf(wk[k], v[c], acc[c]);
                    }
            }
        }
        stg_stream_v4(out + pair * kD + j * 8, pack<__nv_bfloat16>(acc));
    }
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

// upper bound of the tiles of one frame for either tiling (host side: the level shapes live on the device)
long long tile_bound(int L, int Lq) { return (long long)Lq / 5 + L + 2; }

}  // namespace tc

bool tc_forward_supported(const FwdArgs& a)
{
    return a.dtype == kBF16 && a.D == tc::kD && a.L >= 1 && a.L <= tc::kMaxL && a.P >= 1 && a.P <= tc::kMaxP &&
           !a.force_generic && (long long)a.N * a.S < (1ll << 31) && ((size_t)a.value % 16) == 0 &&
           (long long)a.N * a.Lq >= 2048 && (long long)a.N * a.M * tc::tile_bound(a.L, a.Lq) < (1ll << 30) &&
           (long long)a.N * a.Lq * a.M < (1ll << 30);
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
    // tiles the kernel hands to the plain-load gather: [0] = count, [1 ...] = items
    const int bad_cap = (int)((long long)a.N * a.M * tc::tile_bound(a.L, a.Lq));
    int* bad_list = nullptr;
    cudaError_t e = cudaMallocAsync((void**)&bad_list, ((size_t)bad_cap + 1) * sizeof(int), stream);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(bad_list, 0, sizeof(int), stream);
    static const bool want_trace = getenv("MSDA_TC_TRACE") != nullptr;       // development aid: per-role cycle accounting
    unsigned long long* trace = nullptr;
    if (want_trace) { cudaMalloc((void**)&trace, 64 * 3 * 12 * 8); cudaMemset(trace, 0, 64 * 3 * 12 * 8); }
    if (e == cudaSuccess) {
        tc::msda_tc_fwd_kernel<<<MSDA_TC_FWD_CTAS * sms, tc::kFwdThreads, tc::kFwdSmem, stream>>>(
            maps, a.shapes, a.lsi, (const float*)a.loc, (const float*)a.attn, (__nv_bfloat16*)a.out,
            a.N, a.S, a.M, a.L, a.Lq, a.P, 1, bad_list, bad_cap, trace);
        e = cudaGetLastError();
    }
    if (want_trace) {
        cudaStreamSynchronize(stream);
        unsigned long long h[64 * 3 * 12];
        cudaMemcpy(h, trace, sizeof(h), cudaMemcpyDeviceToHost);
        cudaFree(trace);
        static const char* names[3][12] = {
            {"load+footprint", "bbox barrier", "window+entries", "wait mma_done", "syncwarp+arrive", "to epilogue",
             "wait out_ready", "epilogue", "clear", "write", "fence.proxy", "-"},
            {"wait plan", "wait v_free", "issue TMA", "-", "-", "-", "-", "-", "-", "-", "-", "-"},
            {"wait plan", "wait c_full", "wait v_full", "commit+syncwarp", "(k-steps)", "(segments)", "-", "-", "fence_after",
             "issue loop", "-", "-"}};
        static const char* roles[3] = {"build", "producer", "issuer"};
        for (int role = 0; role < 3; ++role) {
            fprintf(stderr, "[tc fwd trace] %-8s:", roles[role]);
            for (int i = 0; i < 12; ++i) {
                double sum = 0;
                for (int b = 0; b < 64; ++b) sum += (double)h[(b * 3 + role) * 12 + i];
                if (names[role][i][0] != '-') fprintf(stderr, "  %s %.0f", names[role][i], sum / 64);
            }
            fprintf(stderr, "\n");
        }
    }
    if (e == cudaSuccess) {
        tc::msda_tc_fwd_fallback_kernel<<<8 * sms, 128, 0, stream>>>(
            (const __nv_bfloat16*)a.value, a.shapes, a.lsi, (const float*)a.loc, (const float*)a.attn,
            (__nv_bfloat16*)a.out, a.N, a.S, a.M, a.L, a.Lq, a.P, a.M * a.D, 1, bad_list, bad_cap);
        e = cudaGetLastError();
    }
    const cudaError_t e2 = cudaFreeAsync(bad_list, stream);
    return e != cudaSuccess ? e : e2;
}

}  // namespace msda
