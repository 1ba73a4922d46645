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

}  // namespace tc
}  // namespace msda
