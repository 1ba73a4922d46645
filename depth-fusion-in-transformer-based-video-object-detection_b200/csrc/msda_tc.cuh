// Shared pieces of the tensor-core (tcgen05) deformable-attention kernels: msda_tc_forward.cu, msda_tc_backward.cu.
//
// Formulation.  A TILE is 128 neighbouring queries of one (frame, head).  Per level, all samples of the tile fall into
// a WINDOW of K pixels (the bounding box of the corner pixels the tile's samples touch, computed from the data).  With
//   C[128 q, K]  = sum over the query's samples on that level of (bilinear weight x attention weight) at the corner
//                  pixels (16 non-zeros per row for P = 4),
//   forward   Out[128 q, 32]  += C . V_window[K, 32]                     (reference cuh:237-299 summed per window pixel)
//   backward  dV_window[K, 32] = C^T . G[128 q, 32]                      (reference cuh:113-152: the atomicAdd scatter)
// are dense tcgen05.mma products: the value window is read ONCE per tile by TMA instead of once per (query, corner)
// through L1, and the backward's scatter becomes one reduction per WINDOW pixel instead of one per (query, corner).
// The window is cut into SEGMENTS of <= 128 pixels (whole window rows); a segment's C block is a 128 x 128 bf16
// K-major SWIZZLE_128B operand (32 KB) that the query-owning threads fill and clear entry by entry, its V block is the
// 64-byte-row SWIZZLE_64B image a TMA box {32 channels, BW pixels} writes -- which IS the MN-major B operand layout
// (tools/umma/msda_tc_probe.cu checks both descriptor encodings against a CPU product).
// Tiles whose window does not fit (scattered sampling locations) are processed by the same CTA with the lane-group
// gather of msda_forward.cu / msda_backward.cu, so the kernels are exact for any input.
#pragma once
#include <cuda.h>
#include "msda_common.cuh"
#include "umma.cuh"

namespace msda {
namespace tc {

constexpr int kTileQ = 128;            // queries per tile = MMA M
constexpr int kTileW = 16, kTileH = 8; // 2-D tile of a query grid (pyramid mode)
constexpr int kSegPx = 128;            // window pixels per segment
constexpr int kMaxL = 4;               // levels (register arrays)
constexpr int kMaxP = 4;               // points per level: lane j of a query's 4-lane group owns point j
constexpr int kMaxBW = 64;             // widest window (pixels)
constexpr int kMaxRows = 64;           // tallest window
constexpr int kD = 32;                 // channels per head
constexpr int kBuildThreads = 4 * kTileQ;
constexpr int kBuildWarps = kBuildThreads / 32;
constexpr int kCBytes = kTileQ * kSegPx * 2;      // one C block
constexpr int kVBytes = kSegPx * kD * 2;          // one V block (forward)

// value viewed as [N*S pixels, ld channels] bf16: map i has box {32 channels, 8 (i + 1) pixels}, SWIZZLE_64B
struct Maps { CUtensorMap m[kMaxBW / 8]; };

__device__ __forceinline__ unsigned long long make_desc(const void* p, unsigned lbo_bytes, unsigned sbo_bytes, unsigned layout)
{
    const unsigned addr = umma::smem_u32(p);
    unsigned long long d = 0;
    d |= (unsigned long long)((addr & 0x3FFFF) >> 4);
    d |= (unsigned long long)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (unsigned long long)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (unsigned long long)1 << 46;
    d |= (unsigned long long)layout << 61;      // 2 = SWIZZLE_128B, 4 = SWIZZLE_64B
    return d;
}
// kind::f16 instruction descriptor: D fp32, A / B bf16, M x N tile, a_mn / b_mn = 1: operand is MN-major
__device__ __forceinline__ unsigned make_idesc(int m, int n, int a_mn, int b_mn)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)a_mn << 15) | ((unsigned)b_mn << 16) |
           ((unsigned)(n >> 3) << 17) | ((unsigned)(m >> 4) << 24);
}

// byte offset of C[q, k] inside a C block: K-major SWIZZLE_128B, two 64-column chunks of [128 rows x 128 B]
__device__ __forceinline__ unsigned c_row_base(int q) { return (unsigned)(((q >> 3) << 10) | ((q & 7) << 7)); }
__device__ __forceinline__ unsigned c_offset(unsigned row_base, int q7, int k)
{
    return (unsigned)((k >> 6) << 14) | row_base | (unsigned)((((k >> 3) & 7) ^ q7) << 4) | (unsigned)((k & 7) << 1);
}

// rows of a window of padded width bw (multiple of 8) per segment, as a shift: bw * rows <= 128 and a multiple of 16
__device__ __forceinline__ int seg_row_shift(int bw)
{
    return bw <= 8 ? 4 : (bw <= 16 ? 3 : (bw <= 32 ? 2 : 1));
}

// streaming loads that stay where they are written (asm volatile): the kernels issue a level's loads one level ahead,
// and a plain asm load is free to sink down to its first use
__device__ __forceinline__ uint4 ldg_prefetch_v4(const void* p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float2 ldg_prefetch_f32x2(const float* p)
{
    float2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ float ldg_prefetch_f32(const float* p)
{
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}

__device__ __forceinline__ void sts_u16(unsigned addr, unsigned short v)
{
    asm volatile("st.shared.u16 [%0], %1;" :: "r"(addr), "h"(v) : "memory");
}

// ---- tile geometry ---------------------------------------------------------------------------------------------
// Query slots of a tile: slot s -> (sy, sx) = (s / 16, s % 16) on a query grid (pyramid mode: the queries ARE the
// pixels of the value pyramid, Lq == S, encoder self-attention), or 128 consecutive queries (linear mode).
struct Tile { int q0, stride, tw, th, lin; };

struct LevelMeta {
    int H[kMaxL], W[kMaxL], start[kMaxL];
    int tiles_x[kMaxL], tile_base[kMaxL + 1];   // pyramid mode: tiles per row / first tile of each level
    int tiles;                                  // tiles per frame
    int pyramid;
};

__device__ __forceinline__ void level_meta_init(LevelMeta* lm, const int64_t* shapes, const int64_t* lsi, int L, int Lq,
                                                int want_pyramid)
{
    long long total = 0;
    int base = 0;
    for (int l = 0; l < L; ++l) {
        lm->H[l] = (int)shapes[2 * l];
        lm->W[l] = (int)shapes[2 * l + 1];
        lm->start[l] = (int)lsi[l];
        lm->tiles_x[l] = (lm->W[l] + kTileW - 1) / kTileW;
        lm->tile_base[l] = base;
        base += lm->tiles_x[l] * ((lm->H[l] + kTileH - 1) / kTileH);
        // the queries can only be the pyramid's pixels if the levels tile [0, Lq) back to back
        if (lm->start[l] != total) want_pyramid = 0;
        total += (long long)lm->H[l] * lm->W[l];
    }
    lm->tile_base[L] = base;
    lm->pyramid = want_pyramid && total == Lq;
    lm->tiles = lm->pyramid ? base : (Lq + kTileQ - 1) / kTileQ;
}

__device__ __forceinline__ Tile tile_decode(const LevelMeta& lm, int t, int L, int Lq)
{
    Tile tl;
    if (!lm.pyramid) {
        tl.q0 = t * kTileQ; tl.stride = 0; tl.tw = min(kTileQ, Lq - tl.q0); tl.th = 1; tl.lin = 1;
        return tl;
    }
    int l = 0;
    while (l + 1 < L && t >= lm.tile_base[l + 1]) ++l;
    const int r = t - lm.tile_base[l];
    const int ty = r / lm.tiles_x[l], tx = r - ty * lm.tiles_x[l];
    tl.q0 = lm.start[l] + ty * kTileH * lm.W[l] + tx * kTileW;
    tl.stride = lm.W[l];
    tl.tw = min(kTileW, lm.W[l] - tx * kTileW);
    tl.th = min(kTileH, lm.H[l] - ty * kTileH);
    tl.lin = 0;
    return tl;
}
// query index of slot s, or -1
__device__ __forceinline__ int tile_query(const Tile& tl, int s)
{
    const int sy = tl.lin ? 0 : (s >> 4), sx = tl.lin ? s : (s & 15);
    return (sx < tl.tw && sy < tl.th) ? tl.q0 + sy * tl.stride + sx : -1;
}

// window of one level of one tile, from the bounding box of the corner pixels
struct Window {
    int x0, y0;        // first column / row
    int bw;            // padded width, multiple of 8
    int rows;
    int rshift;        // rows per segment = 1 << rshift
    int nseg;
};
__device__ __forceinline__ bool window_from_bbox(int minx, int miny, int maxx, int maxy, Window* w)
{
    w->x0 = minx; w->y0 = miny; w->bw = 8; w->rows = 0; w->rshift = 4; w->nseg = 0;
    if (maxx < minx) return true;                       // no sample of the tile lands inside this level
    w->bw = (maxx - minx + 1 + 7) & ~7;
    w->rows = maxy - miny + 1;
    if (w->bw > kMaxBW || w->rows > kMaxRows) return false;
    w->rshift = seg_row_shift(w->bw);
    w->nseg = (w->rows + (1 << w->rshift) - 1) >> w->rshift;
    return true;
}
// rows the MMAs of segment `sidx` cover (loaded by TMA): the segment's window rows, rounded so that rows * bw is a
// multiple of the MMA K (16 pixels)
__device__ __forceinline__ int seg_rows(const Window& w, int sidx)
{
    int r = min(1 << w.rshift, w.rows - (sidx << w.rshift));
    if ((w.bw & 15) && (r & 1)) ++r;
    return r;
}

// ---- one level of one query: samples -> footprints -> C entries ----------------------------------------------
// the samples of one (pair, level): P <= 4 locations and attention weights
struct LevelSamples { float x[kMaxP], y[kMaxP], a[kMaxP]; };

__device__ __forceinline__ void load_level(LevelSamples& ls, const float* __restrict__ loc, const float* __restrict__ attn,
                                           int pair, int LP, int l, int P, bool have)
{
#pragma unroll
    for (int s = 0; s < kMaxP; ++s) { ls.x[s] = -4.f; ls.y[s] = -4.f; ls.a[s] = 0.f; }     // far outside: contributes nothing
    if (!have) return;
    const float* lp = loc + ((long long)pair * LP + l * P) * 2;
    const float* ap = attn + (long long)pair * LP + l * P;
    if (P == kMaxP) {
        const uint4 u0 = ldg_prefetch_v4(lp), u1 = ldg_prefetch_v4(lp + 4), ua = ldg_prefetch_v4(ap);
        ls.x[0] = __uint_as_float(u0.x); ls.y[0] = __uint_as_float(u0.y); ls.x[1] = __uint_as_float(u0.z); ls.y[1] = __uint_as_float(u0.w);
        ls.x[2] = __uint_as_float(u1.x); ls.y[2] = __uint_as_float(u1.y); ls.x[3] = __uint_as_float(u1.z); ls.y[3] = __uint_as_float(u1.w);
        ls.a[0] = __uint_as_float(ua.x); ls.a[1] = __uint_as_float(ua.y); ls.a[2] = __uint_as_float(ua.z); ls.a[3] = __uint_as_float(ua.w);
    } else {
#pragma unroll
        for (int s = 0; s < kMaxP; ++s)
            if (s < P) {
                const float2 xy = ldg_prefetch_f32x2(lp + 2 * s);
                ls.x[s] = xy.x; ls.y[s] = xy.y;
                ls.a[s] = ldg_prefetch_f32(ap + s);
            }
    }
}

// bilinear footprints of the level's samples (reference cuh:285-288, :33-84)
struct Footprints {
    int bx[kMaxP], by[kMaxP];        // floor of the pixel coordinates
    float lw[kMaxP], lh[kMaxP];      // fractional parts
    unsigned inside;                 // bit s: sample s passes the reference's range test (cuh:288)
};
__device__ __forceinline__ void footprints_of(const LevelSamples& ls, int H, int W, Footprints& f)
{
    f.inside = 0;
#pragma unroll
    for (int s = 0; s < kMaxP; ++s) {
        const float w_im = ls.x[s] * (float)W - 0.5f, h_im = ls.y[s] * (float)H - 0.5f;
        const bool in = (h_im > -1.f) && (w_im > -1.f) && (h_im < (float)H) && (w_im < (float)W);
        const float hf = floorf(h_im), wf = floorf(w_im);
        f.bx[s] = (int)wf; f.by[s] = (int)hf;
        f.lw[s] = w_im - wf; f.lh[s] = h_im - hf;
        if (in) f.inside |= 1u << s;
    }
}
// bounding box of the corner pixels of this thread's samples, reduced over the warp and merged into bb[4]
// (min x, min y, max x, max y) in shared memory
__device__ __forceinline__ void bbox_merge(const Footprints& f, int H, int W, int lane, int* bb)
{
    int mnx = 0x7fffffff, mny = 0x7fffffff, mxx = -2, mxy = -2;
#pragma unroll
    for (int s = 0; s < kMaxP; ++s)
        if ((f.inside >> s) & 1u) {
            mnx = min(mnx, f.bx[s]); mxx = max(mxx, f.bx[s]);
            mny = min(mny, f.by[s]); mxy = max(mxy, f.by[s]);
        }
    mnx = __reduce_min_sync(0xffffffffu, mnx); mny = __reduce_min_sync(0xffffffffu, mny);
    mxx = __reduce_max_sync(0xffffffffu, mxx); mxy = __reduce_max_sync(0xffffffffu, mxy);
    if (lane == 0 && mxx >= -1) {        // inside samples have bx in [-1, W-1]: clamp the corner range to the map
        atomicMin(&bb[0], max(mnx, 0)); atomicMin(&bb[1], max(mny, 0));
        atomicMax(&bb[2], min(mxx + 1, W - 1)); atomicMax(&bb[3], min(mxy + 1, H - 1));
    }
}

// C entries of one sample: its 4 corners as (byte offset in the C block, 0xffff = corner outside the map), the
// segments of its two pixel rows (0xff = row outside the map) and the 4 coefficients rounded to bf16
struct SampleEntries {
    unsigned off01, off23;       // corners (y0,x0) | (y0,x1) << 16 ;  (y1,x0) | (y1,x1) << 16
    unsigned cf01, cf23;         // bf16 coefficients, same packing
    unsigned segs;               // segment of row y0 | segment of row y1 << 8
};
__device__ __forceinline__ unsigned pack_bf16x2(float lo, float hi)
{
    return (unsigned)__bfloat16_as_ushort(__float2bfloat16_rn(lo)) | ((unsigned)__bfloat16_as_ushort(__float2bfloat16_rn(hi)) << 16);
}
__device__ __forceinline__ void sample_entries(const Footprints& f, int s, float a, const Window& w, int H, int W,
                                               unsigned row_base, int q7, SampleEntries& e)
{
    const bool in = (f.inside >> s) & 1u;
    const int bx = f.bx[s], by = f.by[s];
    const bool x0ok = in && bx >= 0, x1ok = in && bx + 1 < W, y0ok = by >= 0, y1ok = by + 1 < H;
    const int kx = bx - w.x0, y0r = by - w.y0, rmask = (1 << w.rshift) - 1;
    const int k0 = (y0r & rmask) * w.bw + kx, k1 = ((y0r + 1) & rmask) * w.bw + kx;
    const unsigned o00 = (x0ok && y0ok) ? c_offset(row_base, q7, k0) : 0xffffu;
    const unsigned o01 = (x1ok && y0ok) ? c_offset(row_base, q7, k0 + 1) : 0xffffu;
    const unsigned o10 = (x0ok && y1ok) ? c_offset(row_base, q7, k1) : 0xffffu;
    const unsigned o11 = (x1ok && y1ok) ? c_offset(row_base, q7, k1 + 1) : 0xffffu;
    e.off01 = o00 | (o01 << 16);
    e.off23 = o10 | (o11 << 16);
    e.segs = ((in && y0ok) ? (unsigned)(y0r >> w.rshift) : 0xffu) | (((in && y1ok) ? (unsigned)((y0r + 1) >> w.rshift) : 0xffu) << 8);
    const float hw = 1.f - f.lw[s], hh = 1.f - f.lh[s];
    e.cf01 = pack_bf16x2(hh * hw * a, hh * f.lw[s] * a);            // (bilinear weight) x attention weight, cuh:113-116
    e.cf23 = pack_bf16x2(f.lh[s] * hw * a, f.lh[s] * f.lw[s] * a);
}
// add one pixel row of a sample (two corners) into the thread's row of the C block / zero it again
__device__ __forceinline__ void c_row_add(unsigned cb, unsigned offs, unsigned cfs)
{
    const unsigned o0 = offs & 0xffffu, o1 = offs >> 16;
    unsigned short v0 = 0, v1 = 0;
    if (o0 != 0xffffu) asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v0) : "r"(cb + o0));
    if (o1 != 0xffffu) asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v1) : "r"(cb + o1));
    const float s0 = __uint_as_float((unsigned)v0 << 16) + __uint_as_float(cfs << 16);
    const float s1 = __uint_as_float((unsigned)v1 << 16) + __uint_as_float(cfs & 0xffff0000u);
    if (o0 != 0xffffu) asm volatile("st.shared.u16 [%0], %1;" :: "r"(cb + o0), "h"(__bfloat16_as_ushort(__float2bfloat16_rn(s0))));
    if (o1 != 0xffffu) asm volatile("st.shared.u16 [%0], %1;" :: "r"(cb + o1), "h"(__bfloat16_as_ushort(__float2bfloat16_rn(s1))));
}
__device__ __forceinline__ void c_row_clear(unsigned cb, unsigned offs)
{
    const unsigned o0 = offs & 0xffffu, o1 = offs >> 16;
    const unsigned short z = 0;
    if (o0 != 0xffffu) asm volatile("st.shared.u16 [%0], %1;" :: "r"(cb + o0), "h"(z));
    if (o1 != 0xffffu) asm volatile("st.shared.u16 [%0], %1;" :: "r"(cb + o1), "h"(z));
}

// mbarrier wait that lets the hardware park the warp (suspend-time hint) instead of re-polling every few cycles:
// the polling loops of idle roles were a quarter of all executed instructions (profiles/: ncu source page, round 2)
__device__ __forceinline__ void mbar_wait_parked(unsigned long long* bar, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAITP_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONEP_%=;\n\t"
        "bra WAITP_%=;\n\t"
        "DONEP_%=:\n\t}\n"
        :: "r"(umma::smem_u32(bar)), "r"(parity), "r"(4096u) : "memory");
}

}  // namespace tc
}  // namespace msda
