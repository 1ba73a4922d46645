// grad_value of multi-scale deformable attention on the tensor cores (tcgen05 + TMEM + TMA reduce), bf16 values,
// 32 channels / head -- opt-in (MSDA_FLAG_TC), see DESIGN.md section 3.10 / 3.11.
//
//   grad_value[n, pix, m, :] += sum over samples touching pix of (bilinear weight x attention weight) * grad_out[n,q,m,:]
//
// (reference cuh:113-152: four atomicAdd per channel and sample).  Per tile of 128 neighbouring queries of one
// (frame, head) and per <= 128-pixel segment of a level's window (msda_tc.cuh):
//
//   dV_segment[128 px, 32] = C^T[128 px, 128 q] . G[128 q, 32]
//
// A = the C block read MN-major (the bytes the forward reads K-major), B = the tile's grad_output rows, D in tensor
// memory; the finished block goes TMEM -> registers -> shared memory -> ONE bulk tensor reduction (cp.reduce.async.bulk
// .tensor add.f32) per window row into the fp32 accumulation buffer: one reduction per WINDOW pixel instead of one per
// (query, corner).  grad_sampling_loc / grad_attn_weight stay with msda_bwd_fast_kernel, which skips its reductions for
// every (pair, level) this kernel took (a per-pair level mask written here).
//
// What the forward kernel taught (profiles/r2_tc_forward_trace.txt): a builder that waits for its own segment's MMAs
// (700 cycles of commit -> mbarrier -> wake-up alone) leaves the SM idle.  So here
//   * 4 teams of 128 build threads (thread = query slot) work on 4 tiles at once and NEVER wait for an MMA: a segment is
//     built into the next block of a POOL of 5 C blocks (ticket order = service order), handed over, forgotten;
//   * one issuer warp runs the 8 MMAs of each ticket into a ring of 4 accumulators;
//   * 4 service warps (one per TMEM lane quadrant) flush the ticket's accumulator, zero the whole C block with 16-byte
//     stores (the builders keep no record of what they wrote) and return the block to the pool.
#include <cstdio>
#include <cstdlib>
#include "msda_launch.h"
#include "msda_tc.cuh"

namespace msda {
namespace tc {

using namespace umma;

constexpr int kDvTeams = 4;
constexpr int kDvBuild = kDvTeams * kTileQ;            // 512 build threads
constexpr int kDvSvcWarp0 = kDvBuild / 32;             // warps 16..19: service (warp % 4 = TMEM lane quadrant)
constexpr int kDvIssuerWarp = kDvSvcWarp0 + 4;         // warp 20
constexpr int kDvThreads = (kDvIssuerWarp + 1) * 32;   // 672
constexpr int kDvBlocks = 5;                           // C block pool
constexpr int kDvAcc = 4;                              // accumulator ring in tensor memory (32 columns each)
constexpr int kDvGBytes = kTileQ * kD * 2;             // 8 KB: one team's grad_output tile
constexpr int kDvStageBytes = kSegPx * kD * 4;         // 16 KB: one fp32 dV block
constexpr int kDvSmem = kDvBlocks * kCBytes + kDvTeams * kDvGBytes + 2 * kDvStageBytes + 1024;

// grad_value accumulation buffer viewed as [N*S pixels, M*32 channels] fp32: map i has box {32 channels, 8 (i+1) pixels}
struct AccMaps { CUtensorMap m[kMaxBW / 8]; };

struct DvMeta { int end, team, head, pix0, W, bw, rows; };

struct DvBars {
    unsigned long long clean[kDvBlocks], full[kDvBlocks], done[kDvBlocks], acc_free[kDvAcc];
};

__device__ __forceinline__ void tma_reduce_add_2d(const void* tensor_map, const void* smem_src, int c0, int c1)
{
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
                 :: "l"(tensor_map), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}

__global__ void __launch_bounds__(kDvThreads, 1)
msda_tc_dv_kernel(const __grid_constant__ AccMaps maps, const __nv_bfloat16* __restrict__ grad_out,
                  const int64_t* __restrict__ shapes, const int64_t* __restrict__ lsi, const float* __restrict__ loc,
                  const float* __restrict__ attn, unsigned char* __restrict__ red_levels,
                  int N, int S, int M, int L, int Lq, int P, int want_pyramid, unsigned long long* __restrict__ trace)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* sC = smem;                                        // kDvBlocks x 32 KB
    unsigned char* sG = smem + kDvBlocks * kCBytes;                  // kDvTeams x 8 KB
    unsigned char* sStage = sG + kDvTeams * kDvGBytes;               // 2 x 16 KB
    __shared__ LevelMeta lm;
    __shared__ DvMeta s_meta[kDvBlocks];
    __shared__ int s_bb[kDvTeams][3][4];
    __shared__ int s_ticket, s_teams_done;
    __shared__ int s_cur[kDvTeams][2];                               // ticket of the team's current segment (by segment parity)
    __shared__ volatile int s_served[kDvTeams];                      // tickets of each team whose MMAs have completed
    __shared__ __align__(8) DvBars bars;
    __shared__ unsigned tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        level_meta_init(&lm, shapes, lsi, L, Lq, want_pyramid);
        for (int i = 0; i < kDvBlocks; ++i) {
            mbar_init(&bars.clean[i], 1);
            mbar_init(&bars.full[i], kTileQ / 32);
            mbar_init(&bars.done[i], 1);
        }
        for (int i = 0; i < kDvAcc; ++i) mbar_init(&bars.acc_free[i], 1);
        fence_mbar_init();
        for (int t = 0; t < kDvTeams; ++t)
            for (int i = 0; i < 3; ++i) { s_bb[t][i][0] = 0x7fffffff; s_bb[t][i][1] = 0x7fffffff; s_bb[t][i][2] = -1; s_bb[t][i][3] = -1; }
        s_ticket = 0;
        s_teams_done = 0;
        for (int t = 0; t < kDvTeams; ++t) s_served[t] = 0;
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, kDvAcc * kD);
    for (int i = tid; i < kDvBlocks * kCBytes / 16; i += kDvThreads) reinterpret_cast<uint4*>(sC)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (warp == kDvSvcWarp0 && lane < kMaxBW / 8) tma_prefetch_desc(&maps.m[lane]);
    fence_proxy_async();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const unsigned tmem = tmem_base_s;
#ifdef MSDA_TC_TRACE
    unsigned long long tr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tr_t = 0;
    const bool tracing = trace != nullptr && lane == 0 && (warp == 0 || warp == kDvSvcWarp0 || warp == kDvIssuerWarp);
#define TR_START() do { if (tracing) tr_t = clock64(); } while (0)
#define TR_ADD(i) do { if (tracing) { const long long t_ = clock64(); tr[i] += (unsigned long long)(t_ - tr_t); tr_t = t_; } } while (0)
#else
#define TR_START() do { } while (0)
#define TR_ADD(i) do { } while (0)
#endif

    if (warp == kDvIssuerWarp) {
        // ================================ MMA issuer: tickets in order ================================
        const unsigned idesc = make_idesc(128, kD, 1, 1);
        TR_START();
        for (unsigned t = 0;; ++t) {
            const unsigned b = t % kDvBlocks, u = t / kDvBlocks, a = t % kDvAcc;
            mbar_wait(&bars.full[b], u & 1);
            TR_ADD(0);
            const DvMeta m = s_meta[b];
            if (m.end) break;
            if (t >= kDvAcc) mbar_wait(&bars.acc_free[a], ((t / kDvAcc) - 1) & 1);
            TR_ADD(1);
            tcgen05_fence_after();
            if (elect_one()) {
                const unsigned long long dA = make_desc(sC + b * kCBytes, 16384, 1024, 2);      // C^T: M = pixels, K = queries
                const unsigned long long dB = make_desc(sG + m.team * kDvGBytes, 0, 512, 4);
#pragma unroll
                for (int ks = 0; ks < kTileQ / 16; ++ks)
                    mma_bf16(tmem + a * kD, desc_advance(dA, (unsigned)(ks * 2048)), desc_advance(dB, (unsigned)(ks * 1024)),
                             idesc, ks != 0);
                mma_commit(&bars.done[b]);
            }
            __syncwarp();
            TR_ADD(2);
        }
    } else if (warp >= kDvSvcWarp0) {
        // ================================ service warps: flush, zero, recycle ================================
        const int wq = warp - kDvSvcWarp0;                 // TMEM lane quadrant
        const int k = wq * 32 + lane;                      // pixel of the segment = TMEM lane
        const int stid = tid - kDvSvcWarp0 * 32;           // 0..127
        TR_START();
        for (unsigned t = 0;; ++t) {
            const unsigned b = t % kDvBlocks, u = t / kDvBlocks, a = t % kDvAcc;
            mbar_wait(&bars.full[b], u & 1);
            TR_ADD(0);
            const DvMeta m = s_meta[b];
            if (m.end) break;
            mbar_wait(&bars.done[b], u & 1);
            TR_ADD(1);
            if (stid == 0) s_served[m.team] = s_served[m.team] + 1;      // the team's G tile is no longer read by this ticket
            tcgen05_fence_after();
            float v[32];
            tmem_ld32(tmem + ((unsigned)(wq * 32) << 16) + a * kD, v);
            tcgen05_fence_before();
            // the bulk reductions that read this staging buffer two tickets ago have finished reading it
            TR_ADD(2);
            if (stid == 0 && t >= 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            TR_ADD(3);
            named_bar_sync(9, 128);
            TR_ADD(4);
            unsigned char* stage = sStage + (t & 1) * kDvStageBytes;
#pragma unroll
            for (int c = 0; c < 8; ++c)
                *reinterpret_cast<uint4*>(stage + k * 128 + ((c ^ (k & 7)) << 4)) = pack<float>(v + 4 * c);
            // zero the C block (whatever the builders wrote): 32 KB over 128 threads
            uint4* cz = reinterpret_cast<uint4*>(sC + b * kCBytes);
#pragma unroll
            for (int i = 0; i < kCBytes / 16 / 128; ++i) cz[i * 128 + stid] = make_uint4(0u, 0u, 0u, 0u);
            fence_proxy_async();
            TR_ADD(5);
            named_bar_sync(10, 128);
            TR_ADD(6);
            if (stid == 0) {
                const CUtensorMap* map = &maps.m[(m.bw >> 3) - 1];
                for (int r = 0; r < m.rows; ++r)
                    tma_reduce_add_2d(map, stage + r * m.bw * 128, m.head * kD, m.pix0 + r * m.W);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                mbar_arrive(&bars.clean[b]);
                mbar_arrive(&bars.acc_free[a]);
            }
            TR_ADD(7);
        }
        if (stid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    } else {
        // ================================ build teams ================================
        const int team = warp >> 2;
        const int q = tid & (kTileQ - 1);                  // query slot within the team's tile
        const int twarp = warp & 3;
        const unsigned row_base = c_row_base(q);
        const int q7 = q & 7;
        const unsigned sC_u32 = smem_u32(sC);
        const int tiles = lm.tiles;
        const int total_items = N * tiles * M;
        const int LP = L * P;
        const int bar_id = 1 + team;
        unsigned lev = 0;                                  // levels processed by this team (bbox buffer rotation)
        int issued = 0;                                    // segments this team has handed over so far

        auto decode = [&](int it, int& h_, int& n_, int& qi_, int& pair_) {
            h_ = it % M;
            const int rest = it / M;
            const Tile tl = tile_decode(lm, rest % tiles, L, Lq);
            n_ = rest / tiles;
            qi_ = tile_query(tl, q);
            pair_ = (n_ * Lq + (qi_ >= 0 ? qi_ : 0)) * M + h_;
        };
        const int stride = gridDim.x * kDvTeams;
        int item = blockIdx.x * kDvTeams + team;
        int h = 0, n = 0, qi = -1, pair = 0;
        LevelSamples pf;
        if (item < total_items) {
            decode(item, h, n, qi, pair);
            load_level(pf, loc, attn, pair, LP, 0, P, qi >= 0);
        }
        TR_START();
        while (item < total_items) {
            const int item_next = item + stride;
            int h2 = 0, n2 = 0, qi2 = -1, pair2 = 0;
            if (item_next < total_items) decode(item_next, h2, n2, qi2, pair2);
            // ---- the tile's grad_output rows -> the team's G tile (B operand, MN-major SWIZZLE_64B image) ----
            // the MMAs of the team's previous tiles read the same buffer: wait until all of them have completed
            TR_ADD(0);
            while (s_served[team] < issued) { }
            TR_ADD(1);
            {
                unsigned char* grow = sG + team * kDvGBytes + q * 64;
                const __nv_bfloat16* gsrc = grad_out + (long long)pair * kD;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint4 gv = qi >= 0 ? ldg_stream_v4(gsrc + c * 8) : make_uint4(0u, 0u, 0u, 0u);
                    *reinterpret_cast<uint4*>(grow + ((c ^ ((q >> 1) & 3)) << 4)) = gv;
                }
            }
            unsigned taken = 0;                            // bit l: this kernel accumulates level l of the tile
            for (int l = 0; l < L; ++l, ++lev) {
                const LevelSamples ls = pf;
                if (l + 1 < L) load_level(pf, loc, attn, pair, LP, l + 1, P, qi >= 0);
                else if (item_next < total_items) load_level(pf, loc, attn, pair2, LP, 0, P, qi2 >= 0);
                const int H = lm.H[l], W = lm.W[l];
                Footprints fp;
                footprints_of(ls, H, W, fp);
                int* bb = s_bb[team][lev % 3];
                bbox_merge(fp, H, W, lane, bb);
                TR_ADD(2);
                named_bar_sync(bar_id, kTileQ);
                TR_ADD(3);
                Window w;
                const bool fits = window_from_bbox(bb[0], bb[1], bb[2], bb[3], &w);
                if (q == 0) {
                    int* pb = s_bb[team][(lev + 2) % 3];
                    pb[0] = 0x7fffffff; pb[1] = 0x7fffffff; pb[2] = -1; pb[3] = -1;
                }
                if (!fits) continue;                       // this level of the tile stays with the reduction kernel
                taken |= 1u << l;
                if (w.nseg == 0) continue;
                SampleEntries en[kMaxP];
#pragma unroll
                for (int s = 0; s < kMaxP; ++s) sample_entries(fp, s, ls.a[s], w, H, W, row_base, q7, en[s]);
                TR_ADD(4);
                for (int sidx = 0; sidx < w.nseg; ++sidx) {
                    // next block of the pool: the leader takes the ticket, the team learns it at the barrier
                    if (q == 0) s_cur[team][issued & 1] = atomicAdd(&s_ticket, 1);
                    named_bar_sync(bar_id, kTileQ);
                    TR_ADD(5);
                    const unsigned t = (unsigned)s_cur[team][issued & 1];
                    const unsigned b = t % kDvBlocks, u = t / kDvBlocks;
                    if (u >= 1) mbar_wait(&bars.clean[b], (u - 1) & 1);
                    TR_ADD(6);
                    if (q == 0) {
                        DvMeta& m = s_meta[b];
                        m.end = 0; m.team = team; m.head = h; m.W = W; m.bw = w.bw;
                        m.rows = min(1 << w.rshift, w.rows - (sidx << w.rshift));
                        m.pix0 = n * S + lm.start[l] + (w.y0 + (sidx << w.rshift)) * W + w.x0;
                    }
                    const unsigned cb = sC_u32 + b * kCBytes;
#pragma unroll
                    for (int s = 0; s < kMaxP; ++s) {
                        c_row_add(cb, (en[s].segs & 0xffu) == (unsigned)sidx ? en[s].off01 : 0xffffffffu, en[s].cf01);
                        c_row_add(cb, (en[s].segs >> 8) == (unsigned)sidx ? en[s].off23 : 0xffffffffu, en[s].cf23);
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars.full[b]);
                    ++issued;
                    TR_ADD(7);
                }
            }
            // levels this kernel did NOT take are left to msda_bwd_fast_kernel's reductions
            if (qi >= 0) red_levels[pair] = (unsigned char)(~taken & ((1u << L) - 1u));
            item = item_next; h = h2; n = n2; qi = qi2; pair = pair2;
            (void)twarp;
        }
        // the last team to finish closes the ticket stream
        named_bar_sync(bar_id, kTileQ);
        if (q == 0) {
            if (atomicAdd(&s_teams_done, 1) == kDvTeams - 1) {
                const unsigned t = (unsigned)atomicAdd(&s_ticket, 1);
                const unsigned b = t % kDvBlocks, u = t / kDvBlocks;
                if (u >= 1) mbar_wait(&bars.clean[b], (u - 1) & 1);
                s_meta[b].end = 1;
                for (int i = 0; i < kTileQ / 32; ++i) mbar_arrive(&bars.full[b]);
            }
        }
    }
#ifdef MSDA_TC_TRACE
    if (tracing && blockIdx.x < 64) {
        const int role = warp == 0 ? 0 : (warp == kDvSvcWarp0 ? 1 : 2);
        for (int i = 0; i < 8; ++i) trace[(blockIdx.x * 3 + role) * 8 + i] = tr[i];
    }
#endif
#undef TR_START
#undef TR_ADD
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free(tmem, kDvAcc * kD);
}

typedef CUresult (*DvEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static DvEncodeTiledFn dv_encode_fn()
{
    static DvEncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<DvEncodeTiledFn>(p);
    }
    return fn;
}

static bool make_acc_maps(AccMaps* maps, float* accum, long long pixels, long long ld)
{
    DvEncodeTiledFn fn = dv_encode_fn();
    if (fn == nullptr) return false;
    for (int i = 0; i < kMaxBW / 8; ++i) {
        const cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)pixels};
        const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
        const cuuint32_t box[2] = {(cuuint32_t)kD, (cuuint32_t)(8 * (i + 1))};
        const cuuint32_t estr[2] = {1, 1};
        if (fn(&maps->m[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, accum, dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
    }
    return true;
}

}  // namespace tc

bool tc_backward_supported(const BwdArgs& a)
{
    return a.dtype == kBF16 && a.D == tc::kD && a.L >= 1 && a.L <= tc::kMaxL && a.P >= 1 && a.P <= tc::kMaxP &&
           !a.force_generic && (long long)a.N * a.S < (1ll << 31) && (long long)a.N * a.Lq >= 2048 &&
           (long long)a.N * a.Lq * a.M < (1ll << 30) && (long long)a.N * a.M * ((long long)a.Lq / 5 + a.L + 2) < (1ll << 30) &&
           a.grad_value_accum != nullptr &&
           ((size_t)a.grad_value_accum % 16) == 0 && ((size_t)a.grad_out % 16) == 0;
}

// grad_value contributions of every (tile, level) whose window fits, accumulated into a.grad_value_accum (already
// zeroed); red_levels[pair] = mask of the levels left to the reduction kernel.
cudaError_t tc_backward_dv(const BwdArgs& a, unsigned char* red_levels, cudaStream_t stream)
{
    if (!tc_backward_supported(a)) return cudaErrorInvalidValue;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    static bool configured[64] = {};
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(tc::msda_tc_dv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kDvSmem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    alignas(64) tc::AccMaps maps;
    if (!tc::make_acc_maps(&maps, a.grad_value_accum, (long long)a.N * a.S, (long long)a.M * a.D)) return cudaErrorNotSupported;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    static const bool want_trace = getenv("MSDA_TC_TRACE") != nullptr;       // development aid (-DMSDA_TC_TRACE builds)
    unsigned long long* trace = nullptr;
    if (want_trace) { cudaMalloc((void**)&trace, 64 * 3 * 8 * 8); cudaMemset(trace, 0, 64 * 3 * 8 * 8); }
    tc::msda_tc_dv_kernel<<<sms, tc::kDvThreads, tc::kDvSmem, stream>>>(
        maps, (const __nv_bfloat16*)a.grad_out, a.shapes, a.lsi, (const float*)a.loc, (const float*)a.attn, red_levels,
        a.N, a.S, a.M, a.L, a.Lq, a.P, 1, trace);
    const cudaError_t err = cudaGetLastError();
    if (want_trace) {
        cudaStreamSynchronize(stream);
        unsigned long long h[64 * 3 * 8];
        cudaMemcpy(h, trace, sizeof(h), cudaMemcpyDeviceToHost);
        cudaFree(trace);
        static const char* names[3][8] = {
            {"item start", "wait G free", "load+footprint+bbox", "level barrier", "window+entries", "ticket+barrier",
             "wait clean block", "write+arrive"},
            {"wait full", "wait done", "tmem ld", "wait staging", "barrier A", "stage+zero", "barrier B", "reduce issue"},
            {"wait full", "wait acc", "issue", "-", "-", "-", "-", "-"}};
        static const char* roles[3] = {"build", "service", "issuer"};
        for (int role = 0; role < 3; ++role) {
            fprintf(stderr, "[tc dv trace] %-8s:", roles[role]);
            for (int i = 0; i < 8; ++i) {
                double sum = 0;
                for (int b = 0; b < 64; ++b) sum += (double)h[(b * 3 + role) * 8 + i];
                if (names[role][i][0] != '-') fprintf(stderr, "  %s %.0f", names[role][i], sum / 64);
            }
            fprintf(stderr, "\n");
        }
    }
    return err;
}

}  // namespace msda
