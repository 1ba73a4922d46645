// Probe for the tensor-core formulation of the deformable-attention gather (round 2):
//   forward   Out[128 q, 32 d]  = C[128 q, K px] . Vwin[K px, 32 d]        (A K-major,  B MN-major)
//   backward  dV [K px, 32 d]   = C^T[K px, 128 q] . G[128 q, 32 d]        (A MN-major, B MN-major)
// C (bilinear x attention coefficients, 16 non-zeros per row and level) is built by the query-owning threads in the
// K-major SWIZZLE_128B image; the SAME bytes are the MN-major image of C^T.  Vwin lands by TMA (SWIZZLE_64B box of one
// head's 32 channels = 64-byte rows), which is the MN-major SWIZZLE_64B operand layout.
//
// What this program answers on a B200 (nothing here is a product path):
//   1. which (LBO, SBO) encoding the MN-major descriptors want (every variant is tried, errors printed);
//   2. that a 4-D TMA box with SWIZZLE_64B and out-of-map coordinates (zero fill) feeds the MMA directly;
//   3. cycles per tcgen05.mma for M=128, N in {32,64,128,256}, operands in shared memory, and how much of the
//      shared-memory bandwidth is left to LDS traffic of other warps while the tensor core streams A.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o msda_tc_probe msda_tc_probe.cu -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../depth-fusion-in-transformer-based-video-object-detection_b200/csrc/umma.cuh"

using namespace umma;
typedef __nv_bfloat16 bf16;

__device__ __forceinline__ unsigned long long make_desc(const void* p, unsigned lbo_bytes, unsigned sbo_bytes, unsigned layout)
{
    const unsigned addr = smem_u32(p);
    unsigned long long d = 0;
    d |= (unsigned long long)((addr & 0x3FFFF) >> 4);
    d |= (unsigned long long)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (unsigned long long)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (unsigned long long)1 << 46;
    d |= (unsigned long long)layout << 61;      // 2 = SWIZZLE_128B, 4 = SWIZZLE_64B
    return d;
}
__device__ __forceinline__ unsigned make_idesc(int m, int n, int a_mn, int b_mn)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)a_mn << 15) | ((unsigned)b_mn << 16) |
           ((unsigned)(n >> 3) << 17) | ((unsigned)(m >> 4) << 24);
}
// byte offset of element (row r, 16-bit column k) in the K-major SWIZZLE_128B image with 128 rows per 64-column chunk
__host__ __device__ inline unsigned c_offset(int r, int k)
{
    const int chunk = k >> 6, kk = k & 63;
    return (unsigned)(chunk * 16384 + (r >> 3) * 1024 + (r & 7) * 128 + ((((kk >> 3) ^ (r & 7)) << 4) | ((kk & 7) << 1)));
}
// byte offset of element (row p, 16-bit channel c < 32) in the SWIZZLE_64B image of 64-byte rows
__host__ __device__ inline unsigned v_offset(int p, int c)
{
    return (unsigned)(p * 64 + (((c >> 3) ^ ((p >> 1) & 3)) << 4) + ((c & 7) << 1));
}

constexpr int KW = 256;        // window pixels in the correctness tests

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const void* map, int c0, int c1, int c2, int c3, unsigned long long* bar)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        :: "r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

// mode 0: forward form with V written by threads; mode 1: backward form; mode 2: forward form, V by TMA.
// variant bit0: swap LBO/SBO of the B descriptor; bit1: swap LBO/SBO of the (MN-major) A descriptor.
__global__ void __launch_bounds__(128, 1)
check_kernel(const bf16* __restrict__ Cg, const bf16* __restrict__ Vg, const bf16* __restrict__ Gg, float* __restrict__ out,
             int mode, int variant, const __grid_constant__ CUtensorMap tm_v, int x0, int y0, int head, int frame)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* sC = smem;                  // 64 KB
    unsigned char* sV = smem + 65536;          // 16 KB
    unsigned char* sG = smem + 65536 + 16384;  // 8 KB
    __shared__ __align__(8) unsigned long long bar, tbar;
    __shared__ unsigned tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc(&tmem_base_s, 64);
    if (tid == 0) { mbar_init(&bar, 1); mbar_init(&tbar, 1); fence_mbar_init(); }
    for (int i = tid; i < 128 * KW; i += 128) {
        const int r = i / KW, k = i % KW;
        *reinterpret_cast<bf16*>(sC + c_offset(r, k)) = Cg[i];
    }
    if (mode != 2)
        for (int i = tid; i < KW * 32; i += 128) *reinterpret_cast<bf16*>(sV + v_offset(i / 32, i % 32)) = Vg[i];
    for (int i = tid; i < 128 * 32; i += 128) *reinterpret_cast<bf16*>(sG + v_offset(i / 32, i % 32)) = Gg[i];
    fence_proxy_async();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const unsigned tmem = tmem_base_s;
    if (mode == 2) {
        if (tid == 0) {
            mbar_expect_tx(&tbar, 16384);
            tma_load_4d(sV, &tm_v, head * 32, x0, y0, frame, &tbar);      // box {32 ch, 32 px, 8 rows, 1}
        }
        mbar_wait(&tbar, 0);
    }
    if (tid == 0) {
        if (mode == 0 || mode == 2) {
            const unsigned idesc = make_idesc(128, 32, 0, 1);
            for (int ks = 0; ks < KW / 16; ++ks) {
                const unsigned long long da = make_desc_sw128(sC + (ks >> 2) * 16384 + (ks & 3) * 32);
                const unsigned long long db = (variant & 1) ? make_desc(sV + ks * 1024, 512, 0, 4)
                                                            : make_desc(sV + ks * 1024, 0, 512, 4);
                mma_bf16(tmem, da, db, idesc, ks != 0);
            }
        } else {
            const unsigned idesc = make_idesc(128, 32, 1, 1);
            for (int mb = 0; mb < KW / 128; ++mb)
                for (int ks = 0; ks < 8; ++ks) {
                    const void* pa = sC + (2 * mb) * 16384 + ks * 2048;
                    const unsigned long long da = (variant & 2) ? make_desc(pa, 1024, 16384, 2) : make_desc(pa, 16384, 1024, 2);
                    const unsigned long long db = (variant & 1) ? make_desc(sG + ks * 1024, 512, 0, 4)
                                                                : make_desc(sG + ks * 1024, 0, 512, 4);
                    mma_bf16(tmem + mb * 32, da, db, idesc, ks != 0);
                }
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tcgen05_fence_after();
    const int nblk = (mode == 1) ? KW / 128 : 1;
    for (int mb = 0; mb < nblk; ++mb) {
        float v[32];
        tmem_ld32(tmem + ((unsigned)(warp * 32) << 16) + mb * 32, v);
        for (int c = 0; c < 32; ++c) out[(mb * 128 + tid) * 32 + c] = v[c];
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free(tmem, 64);
}

// ---------------------------------------------------------------------------------------------------------------
// rate: one CTA per SM; thread 0 of warp 0 streams MMAs over a 64 KB A image (4 chunks) and a B image; warps 1..4
// optionally hammer shared memory with conflict-free LDS.128 at the same time.
// ---------------------------------------------------------------------------------------------------------------
template <int N, int A_MN>
__global__ void __launch_bounds__(160, 1)
rate_kernel(int segs, int noise, unsigned long long* __restrict__ cycles, unsigned long long* __restrict__ noise_bytes)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* sA = smem;                   // 64 KB
    unsigned char* sB = smem + 65536;           // 256 k x N x 2 B  (<= 128 KB)
    __shared__ __align__(8) unsigned long long bar;
    __shared__ unsigned tmem_base_s;
    __shared__ volatile int stop;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); stop = 0; }
    for (int i = tid; i < (65536 + 256 * N * 2) / 4; i += 160) reinterpret_cast<unsigned*>(smem)[i] = 0x3c003c00u;
    fence_proxy_async();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const unsigned tmem = tmem_base_s;
    if (warp == 0) {
        const unsigned idesc = make_idesc(128, N, A_MN, 1);
        const unsigned layoutB = N == 32 ? 4u : 2u;
        const unsigned rowB = N == 32 ? 64u : 128u;            // bytes of one MN chunk row
        const unsigned chunksB = (N * 2 + rowB - 1) / rowB;    // MN chunks (LBO steps)
        // B image: [chunk][256 k rows][rowB bytes]
        const unsigned long long dA0 = A_MN ? make_desc(sA, 16384, 1024, 2) : make_desc_sw128(sA);
        const unsigned long long dB0 = make_desc(sB, 256 * rowB, 8 * rowB, layoutB);
        (void)chunksB;
        long long t0 = 0, t1 = 0;
        __syncwarp();
        if (elect_one()) t0 = clock64();
        for (int s = 0; s < segs; ++s) {
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 16; ++ks) {
                    // K-major A: chunk ks/4, 32 B per k16 step.  MN-major A (M = k): M block ks/8 (2 chunks), k-step ks%8
                    const unsigned a_off = A_MN ? (unsigned)((ks >> 3) * 32768 + (ks & 7) * 2048)
                                                : (unsigned)((ks >> 2) * 16384 + (ks & 3) * 32);
                    const unsigned b_off = A_MN ? (unsigned)((ks & 7) * 16 * rowB) : (unsigned)(ks * 16 * rowB);
                    mma_bf16(tmem + (A_MN ? (ks >> 3) * N : 0), desc_advance(dA0, a_off), desc_advance(dB0, b_off), idesc, true);
                }
            }
            __syncwarp();
        }
        if (elect_one()) mma_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, 0);
        if (elect_one()) {
            t1 = clock64();
            cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
            stop = 1;
        }
        __syncwarp();
    } else if (noise) {
        unsigned acc = 0;
        unsigned long long n = 0;
        const unsigned base = smem_u32(sA) + (warp - 1) * 16384 + lane * 16;
        while (!stop) {
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                unsigned a, b, c, d;
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(base + u * 512));
                acc ^= a ^ b ^ c ^ d;
            }
            n += 16 * 512;
        }
        if (lane == 0) atomicAdd(&noise_bytes[blockIdx.x], n);
        if (acc == 0x12345678u) cycles[blockIdx.x] = 0;
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free(tmem, 512);
}

template <int N, int A_MN>
static void run_rate(const char* name, int noise)
{
    const int segs = 2000, grid = 148;
    unsigned long long *dc, *dn;
    cudaMalloc(&dc, grid * 8); cudaMalloc(&dn, grid * 8);
    cudaMemset(dc, 0, grid * 8); cudaMemset(dn, 0, grid * 8);
    const int smem = 65536 + 256 * N * 2 + 2048;
    cudaFuncSetAttribute(rate_kernel<N, A_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    rate_kernel<N, A_MN><<<grid, 160, smem>>>(10, noise, dc, dn);
    cudaMemset(dn, 0, grid * 8);
    cudaEventRecord(e0);
    rate_kernel<N, A_MN><<<grid, 160, smem>>>(segs, noise, dc, dn);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<unsigned long long> hc(grid), hn(grid);
    cudaMemcpy(hc.data(), dc, grid * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(hn.data(), dn, grid * 8, cudaMemcpyDeviceToHost);
    double avg = 0, mx = 0, nb = 0;
    for (int i = 0; i < grid; ++i) { avg += hc[i]; mx = fmax(mx, (double)hc[i]); nb += hn[i]; }
    avg /= grid;
    const double n_mma = segs * 16.0;
    printf("rate %-28s noise=%d: %s  %.1f cyc/MMA (max %.1f)  A %.0f B/cyc  kernel %.3f ms  LDS noise %.1f B/cyc/SM\n",
           name, noise, cudaGetErrorString(e), avg / n_mma, mx / n_mma, 4096.0 * n_mma / avg, ms, nb / grid / avg);
    cudaFree(dc); cudaFree(dn);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main()
{
    srand(7);
    auto rnd = []() { return (rand() / (float)RAND_MAX - 0.5f) * 2.f; };
    // C: 128 x 256, 16 non-zeros per row in [0, 0.25)
    std::vector<bf16> hC(128 * KW, __float2bfloat16(0.f)), hV(KW * 32), hG(128 * 32);
    std::vector<float> fC(128 * KW, 0.f), fV(KW * 32), fG(128 * 32);
    for (int r = 0; r < 128; ++r)
        for (int j = 0; j < 16; ++j) {
            const int k = rand() % KW;
            hC[r * KW + k] = __float2bfloat16(fabsf(rnd()) * 0.25f);
        }
    for (int i = 0; i < 128 * KW; ++i) fC[i] = __bfloat162float(hC[i]);
    for (int i = 0; i < KW * 32; ++i) { hV[i] = __float2bfloat16(rnd()); fV[i] = __bfloat162float(hV[i]); }
    for (int i = 0; i < 128 * 32; ++i) { hG[i] = __float2bfloat16(rnd()); fG[i] = __bfloat162float(hG[i]); }
    // an image for the TMA test: [N=2][H=20][W=40][256 ch]
    const int IN = 2, IH = 20, IW = 40, IC = 256;
    std::vector<bf16> hI((size_t)IN * IH * IW * IC);
    for (size_t i = 0; i < hI.size(); ++i) hI[i] = __float2bfloat16(rnd());
    bf16 *dC, *dV, *dG, *dI; float* dO;
    cudaMalloc(&dC, hC.size() * 2); cudaMalloc(&dV, hV.size() * 2); cudaMalloc(&dG, hG.size() * 2);
    cudaMalloc(&dI, hI.size() * 2); cudaMalloc(&dO, KW * 32 * 4);
    cudaMemcpy(dC, hC.data(), hC.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dV, hV.data(), hV.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dG, hG.data(), hG.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dI, hI.data(), hI.size() * 2, cudaMemcpyHostToDevice);

    CUtensorMap tm;
    {
        void* p = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            printf("no cuTensorMapEncodeTiled\n"); return 1;
        }
        const cuuint64_t dims[4] = {IC, IW, IH, IN};
        const cuuint64_t strides[3] = {IC * 2, (cuuint64_t)IW * IC * 2, (cuuint64_t)IH * IW * IC * 2};
        const cuuint32_t box[4] = {32, 32, 8, 1};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = ((EncodeTiledFn)p)(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dI, dims, strides, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("tensor map encode: %d\n", (int)r);
    }
    const int smem = 65536 + 16384 + 8192 + 2048;
    cudaFuncSetAttribute(check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    std::vector<float> hO(KW * 32);
    auto run_checks = [&](int vlo, int vhi) {
    for (int mode = 0; mode < 3; ++mode)
        for (int variant = vlo; variant < (mode == 1 ? vhi : (vhi > 2 ? 2 : vhi)); ++variant) {
            const int x0 = -3, y0 = 15, head = 5, frame = 1;      // window hangs over the left and bottom edges
            cudaMemset(dO, 0, KW * 32 * 4);
            check_kernel<<<1, 128, smem>>>(dC, dV, dG, dO, mode, variant, tm, x0, y0, head, frame);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mode %d variant %d: %s\n", mode, variant, cudaGetErrorString(e)); exit(1); }
            cudaMemcpy(hO.data(), dO, KW * 32 * 4, cudaMemcpyDeviceToHost);
            double maxerr = 0, maxref = 0;
            if (mode == 0 || mode == 2) {
                for (int r = 0; r < 128; ++r)
                    for (int c = 0; c < 32; ++c) {
                        double acc = 0;
                        for (int k = 0; k < KW; ++k) {
                            double v;
                            if (mode == 0) v = fV[k * 32 + c];
                            else {
                                const int y = y0 + k / 32, x = x0 + k % 32;
                                v = (y >= 0 && y < IH && x >= 0 && x < IW)
                                        ? __bfloat162float(hI[(((size_t)frame * IH + y) * IW + x) * IC + head * 32 + c]) : 0.0;
                            }
                            acc += (double)fC[r * KW + k] * v;
                        }
                        maxerr = fmax(maxerr, fabs(acc - hO[r * 32 + c])); maxref = fmax(maxref, fabs(acc));
                    }
            } else {
                for (int k = 0; k < KW; ++k)
                    for (int c = 0; c < 32; ++c) {
                        double acc = 0;
                        for (int r = 0; r < 128; ++r) acc += (double)fC[r * KW + k] * fG[r * 32 + c];
                        maxerr = fmax(maxerr, fabs(acc - hO[k * 32 + c])); maxref = fmax(maxref, fabs(acc));
                    }
            }
            printf("check mode %d (%s) variant %d: max|err| %.3e max|ref| %.3e -> %s\n", mode,
                   mode == 0 ? "fwd, V by threads" : mode == 1 ? "bwd C^T.G" : "fwd, V by TMA SW64 4-D box", variant, maxerr,
                   maxref, maxerr / maxref < 1e-5 ? "EXACT" : "mismatch");
            fflush(stdout);
        }
    };
    run_checks(0, 1);          // the expected encodings first; the alternatives (which may read stray addresses) last

    for (int noise = 0; noise < 2; ++noise) {
        run_rate<32, 0>("N=32  A K-major", noise);
        run_rate<64, 0>("N=64  A K-major", noise);
        run_rate<128, 0>("N=128 A K-major", noise);
        run_rate<256, 0>("N=256 A K-major", noise);
        run_rate<32, 1>("N=32  A MN-major (bwd form)", noise);
        run_rate<64, 1>("N=64  A MN-major (bwd form)", noise);
    }
    run_checks(1, 4);
    return 0;
}
