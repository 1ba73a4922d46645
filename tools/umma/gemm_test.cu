// Stand-alone bring-up test of the tcgen05 building blocks used by csrc/ffn_fused.cu:
//   Y[rows, 256] = X[rows, 256] @ W[256, 256]^T + bias   (bf16 in, fp32 accumulate in TMEM, bf16 out)
// 128-row tile per CTA, operands staged in shared memory in the K-major SWIZZLE_128B layout,
// one thread issues tcgen05.mma, 4 warps read the accumulator back with tcgen05.ld.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o gemm_test gemm_test.cu ; run on a B200.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../depth-fusion-in-transformer-based-video-object-detection_b200/csrc/umma.cuh"

using namespace umma;

constexpr int K = 256, N = 256, TM = 128;

__global__ void __launch_bounds__(256, 1)
gemm_kernel(const __nv_bfloat16* __restrict__ X, const __nv_bfloat16* __restrict__ W,
            const __nv_bfloat16* __restrict__ bias, __nv_bfloat16* __restrict__ Y, int rows)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* sA = smem;                       // 4 k-blocks x [128 rows x 128 B]
    unsigned char* sB = smem + 4 * TM * 128;        // 4 k-blocks x [256 rows x 128 B]
    __shared__ __align__(8) unsigned long long bar;
    __shared__ unsigned tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5;
    const int row0 = blockIdx.x * TM;
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    // A: 128 rows x 32 chunks of 16 B; B: 256 rows x 32 chunks
    for (int i = tid; i < TM * 32; i += 256) {
        const int r = i >> 5, c = i & 31;
        const int gr = min(row0 + r, rows - 1);
        cp_async16(sA + sw128_offset(r, c, TM), X + (size_t)gr * K + c * 8);
    }
    for (int i = tid; i < N * 32; i += 256) {
        const int r = i >> 5, c = i & 31;
        cp_async16(sB + sw128_offset(r, c, N), W + (size_t)r * K + c * 8);
    }
    cp_async_wait_all();
    fence_proxy_async();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const unsigned tmem = tmem_base_s;
    if (tid == 0) {
        const unsigned idesc = make_idesc_bf16(TM, N);
#ifdef A_FROM_TMEM
        // A: smem -> TMEM columns [256, 384) (16 slabs of 128 rows x 16 bf16 = 8 columns each), then A from TMEM
#pragma unroll
        for (int kb = 0; kb < 4; ++kb)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                tmem_cp_128x256b(tmem + 256 + (kb * 4 + j) * 8, make_desc_sw128(sA + kb * TM * 128 + j * 32));
#pragma unroll
        for (int kb = 0; kb < 4; ++kb)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const unsigned long long db = make_desc_sw128(sB + kb * N * 128 + j * 32);
                mma_bf16_ts(tmem, tmem + 256 + (kb * 4 + j) * 8, db, idesc, (kb | j) != 0);
            }
#else
#pragma unroll
        for (int kb = 0; kb < 4; ++kb)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const unsigned long long da = make_desc_sw128(sA + kb * TM * 128 + j * 32);
                const unsigned long long db = make_desc_sw128(sB + kb * N * 128 + j * 32);
                mma_bf16(tmem, da, db, idesc, (kb | j) != 0);
            }
#endif
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tcgen05_fence_after();
    if (warp < 4) {
        const int r = warp * 32 + (tid & 31);
        const int gr = row0 + r;
#pragma unroll 1
        for (int c0 = 0; c0 < N; c0 += 32) {
            float v[32];
            tmem_ld32(tmem + ((unsigned)(warp * 32) << 16) + c0, v);
            if (gr < rows) {
#pragma unroll
                for (int c = 0; c < 32; c += 2) {
                    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(bias + c0 + c));
                    *reinterpret_cast<__nv_bfloat162*>(Y + (size_t)gr * N + c0 + c) =
                        __floats2bfloat162_rn(v[c] + b.x, v[c + 1] + b.y);
                }
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free(tmem, 512);
}

int main()
{
    const int rows = 1000;
    std::vector<__nv_bfloat16> hX((size_t)rows * K), hW((size_t)N * K), hb(N);
    std::vector<float> fX(hX.size()), fW(hW.size()), fb(N);
    srand(1);
    auto rnd = []() { return (rand() / (float)RAND_MAX - 0.5f) * 2.f; };
    for (size_t i = 0; i < hX.size(); ++i) { hX[i] = __float2bfloat16(rnd()); fX[i] = __bfloat162float(hX[i]); }
    for (size_t i = 0; i < hW.size(); ++i) { hW[i] = __float2bfloat16(rnd() * 0.1f); fW[i] = __bfloat162float(hW[i]); }
    for (int i = 0; i < N; ++i) { hb[i] = __float2bfloat16(rnd()); fb[i] = __bfloat162float(hb[i]); }
    __nv_bfloat16 *dX, *dW, *db, *dY;
    cudaMalloc(&dX, hX.size() * 2); cudaMalloc(&dW, hW.size() * 2); cudaMalloc(&db, N * 2); cudaMalloc(&dY, (size_t)rows * N * 2);
    cudaMemcpy(dX, hX.data(), hX.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dW, hW.data(), hW.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, hb.data(), N * 2, cudaMemcpyHostToDevice);
    cudaMemset(dY, 0, (size_t)rows * N * 2);
    const int smem = 4 * TM * 128 + 4 * N * 128 + 1024;
    cudaFuncSetAttribute(gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    gemm_kernel<<<(rows + TM - 1) / TM, 256, smem>>>(dX, dW, db, dY, rows);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<__nv_bfloat16> hY((size_t)rows * N);
    cudaMemcpy(hY.data(), dY, hY.size() * 2, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int r = 0; r < rows; ++r)
        for (int n = 0; n < N; ++n) {
            double acc = fb[n];
            for (int k = 0; k < K; ++k) acc += (double)fX[(size_t)r * K + k] * fW[(size_t)n * K + k];
            maxerr = fmax(maxerr, fabs(acc - __bfloat162float(hY[(size_t)r * N + n])));
            maxref = fmax(maxref, fabs(acc));
        }
    printf("max |err| %.4e  max |ref| %.4e  normalised %.3e  %s\n", maxerr, maxref, maxerr / maxref,
           maxerr / maxref < 8e-3 ? "OK" : "MISMATCH");
    return maxerr / maxref < 8e-3 ? 0 : 2;
}
