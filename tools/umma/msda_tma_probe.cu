// How fast does one SM pull value-window rows into shared memory?  (round 2, after the first tensor-core forward ran at
// ~20k cycles per tile no matter how the build threads were organised.)
//   A  TMA box {32 channels (64 B), BW pixels} out of the pixel-major value tensor [pixels, 256 ch]: BW runs of 64 B,
//      512 B apart  -- what msda_tc_forward.cu v1/v2 issue, one box per window row;
//   B  TMA box {32 ch, BW px} out of a HEAD-major copy [heads, pixels, 32 ch]: one contiguous run of BW * 64 B;
//   C  one warp copying the same window row with cp.async (16 B per lane-op) out of the pixel-major tensor.
// Every CTA (4 per SM, like the kernel) loops over window rows scattered over a 180 MB tensor (L2/HBM resident),
// `depth` rows in flight; prints cycles per row and the implied bytes per cycle per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o msda_tma_probe msda_tma_probe.cu -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../depth-fusion-in-transformer-based-video-object-detection_b200/csrc/umma.cuh"

using namespace umma;
typedef __nv_bfloat16 bf16;
constexpr int DEPTH = 4;

__device__ __forceinline__ unsigned hash32(unsigned x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

template <int MODE>   // 0 TMA pixel-major, 1 TMA head-major, 2 cp.async pixel-major
__global__ void __launch_bounds__(64, 4)
row_kernel(const __grid_constant__ CUtensorMap map, const bf16* __restrict__ value, int pixels, int bw, int rows_total,
           unsigned long long* __restrict__ cycles)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) unsigned long long bar[DEPTH];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { for (int i = 0; i < DEPTH; ++i) mbar_init(&bar[i], 1); fence_mbar_init(); }
    __syncthreads();
    if (warp != 0) return;
    const long long t0 = clock64();
    const int row_bytes = bw * 64;
    for (int r = 0; r < rows_total + DEPTH; ++r) {
        const int st = r % DEPTH;
        if (r >= DEPTH) {
            if (MODE == 2) { cp_async_wait<DEPTH - 1>(); }
            else mbar_wait(&bar[st], ((r / DEPTH) - 1) & 1);
        }
        if (r < rows_total) {
            const unsigned hsh = hash32(blockIdx.x * 7919u + r);
            const int pix = (int)(hsh % (unsigned)(pixels - bw));
            const int head = (int)((hsh >> 20) & 7);
            if (MODE == 2) {
                for (int i = lane; i < bw * 4; i += 32) {
                    const int p = i >> 2, c = i & 3;
                    cp_async16(smem + st * 4096 + p * 64 + ((c ^ ((p >> 1) & 3)) << 4),
                               value + (size_t)(pix + p) * 256 + head * 32 + c * 8);
                }
                cp_async_commit();
            } else if (elect_one()) {
                mbar_expect_tx(&bar[st], (unsigned)row_bytes);
                if (MODE == 0) tma_load_2d(smem + st * 4096, &map, head * 32, pix, &bar[st]);
                else tma_load_2d(smem + st * 4096, &map, 0, head * pixels + pix, &bar[st]);
            }
            __syncwarp();
        } else if (MODE == 2) cp_async_commit();
    }
    if (MODE == 2) cp_async_wait_all();
    if (lane == 0) cycles[blockIdx.x] = (unsigned long long)(clock64() - t0);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main()
{
    const int pixels = 8 * 22223;
    bf16* dV;
    cudaMalloc(&dV, (size_t)pixels * 256 * 2);
    cudaMemset(dV, 0, (size_t)pixels * 256 * 2);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)p;
    const int grid = 148 * 4, rows = 2000;
    unsigned long long* dc; cudaMalloc(&dc, grid * 8);
    std::vector<unsigned long long> hc(grid);
    for (int bw = 8; bw <= 64; bw *= 2) {
        for (int mode = 0; mode < 3; ++mode) {
            CUtensorMap map;
            const cuuint32_t estr[2] = {1, 1};
            if (mode == 1) {
                const cuuint64_t dims[2] = {32, (cuuint64_t)pixels * 8};
                const cuuint64_t strides[1] = {64};
                const cuuint32_t box[2] = {32, (cuuint32_t)bw};
                enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dV, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            } else {
                const cuuint64_t dims[2] = {256, (cuuint64_t)pixels};
                const cuuint64_t strides[1] = {512};
                const cuuint32_t box[2] = {32, (cuuint32_t)bw};
                enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dV, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            }
            const int smem = DEPTH * 4096 + 1024;
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            if (mode == 0) row_kernel<0><<<grid, 64, smem>>>(map, dV, pixels, bw, rows, dc);
            if (mode == 1) row_kernel<1><<<grid, 64, smem>>>(map, dV, pixels, bw, rows, dc);
            if (mode == 2) row_kernel<2><<<grid, 64, smem>>>(map, dV, pixels, bw, rows, dc);
            cudaEventRecord(e1);
            cudaError_t e = cudaDeviceSynchronize();
            float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
            cudaMemcpy(hc.data(), dc, grid * 8, cudaMemcpyDeviceToHost);
            double avg = 0; for (int i = 0; i < grid; ++i) avg += hc[i]; avg /= grid;
            printf("bw %2d %-22s: %s  %.0f cyc/row/CTA  -> %.1f B/cyc/SM (4 CTAs)  kernel %.3f ms  %.0f GB/s chip\n", bw,
                   mode == 0 ? "TMA pixel-major 64B runs" : mode == 1 ? "TMA head-major contiguous" : "cp.async warp",
                   cudaGetErrorString(e), avg / rows, 4.0 * bw * 64 * rows / avg, ms, (double)grid * rows * bw * 64 / ms / 1e6);
        }
    }
    return 0;
}
