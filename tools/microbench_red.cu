// Micro-benchmark: cost of the ways a warp can add one 128-byte fp32 row into global memory on
// sm_100a.  Guides the grad_value scatter of the backward kernel.  Standalone:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_red microbench_red.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ void red_v4(float* p, float a)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" :: "l"(p), "f"(a) : "memory");
}
__device__ __forceinline__ void red_v2(float* p, float a)
{
    asm volatile("red.global.add.v2.f32 [%0], {%1,%1};" :: "l"(p), "f"(a) : "memory");
}
__device__ __forceinline__ void red_bf16x2_v4(float* p, unsigned a)
{
    asm volatile("red.global.add.noftz.v4.bf16x2 [%0], {%1,%1,%1,%1};" :: "l"(p), "r"(a) : "memory");
}
__device__ __forceinline__ void st_v4(float* p, float a)
{
    asm volatile("st.global.v4.f32 [%0], {%1,%1,%1,%1};" :: "l"(p), "f"(a) : "memory");
}
__device__ __forceinline__ void red_s(float* p, float a)
{
    asm volatile("red.global.add.f32 [%0], %1;" :: "l"(p), "f"(a) : "memory");
}

// mode 0: scalar, one row (128 B) per warp instruction
// mode 1: v4, 8 lanes per row, 4 random rows per warp instruction
// mode 2: v4, 4 adjacent rows (512 B contiguous) per warp instruction
// mode 3: v2, 16 lanes per row, 2 random rows per warp instruction
// mode 4: LDG.128 gather, 8 lanes per row (reference point for the load path)
// mode 5: TMA bulk reduce smem->global, 128 B per op, one op per 8-lane group leader
template <int MODE>
__global__ void __launch_bounds__(256) k(float* buf, uint32_t rows_mask, int iters, float* sink, int local)
{
    __shared__ __align__(128) float src[8][4][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t wid = blockIdx.x * 8 + warp;
    const int grp = lane >> 3, sub = lane & 7;
    for (int i = 0; i < 32; ++i) src[warp][grp][i] = 1.0f;
    __syncthreads();
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 4
        for (int j = 0; j < 16; ++j) {
            const uint32_t seed = (wid * 1315423911u) ^ (it * 16 + j) * 2654435761u;
            // local != 0: rows drawn from a 256-row window that slides slowly (cache-friendly)
            uint32_t r0 = local ? ((wid * 64 + (it >> 2)) + (hash32(seed) & 255u)) & rows_mask : hash32(seed) & rows_mask;
            uint32_t rg = local ? ((wid * 64 + (it >> 2)) + (hash32(seed + grp * 7919u) & 255u)) & rows_mask
                                : hash32(seed + grp * 7919u) & rows_mask;
            if (MODE == 0) red_s(buf + (size_t)r0 * 32 + lane, 1.0f);
            if (MODE == 1) red_v4(buf + (size_t)rg * 32 + sub * 4, 1.0f);
            if (MODE == 2) red_v4(buf + (size_t)(r0 & ~3u) * 32 + lane * 4, 1.0f);
            if (MODE == 3) red_v2(buf + (size_t)(hash32(seed + (lane >> 4) * 7919u) & rows_mask) * 32 + (lane & 15) * 2, 1.0f);
            if (MODE == 4) { float4 v = *reinterpret_cast<const float4*>(buf + (size_t)rg * 32 + sub * 4); acc += v.x + v.w; }
            if (MODE == 6) red_bf16x2_v4(buf + (size_t)(hash32(seed + (lane >> 2) * 7919u) & rows_mask) * 32 + (lane & 3) * 4, 0x3f803f80u);
            if (MODE == 7) st_v4(buf + (size_t)rg * 32 + sub * 4, 1.0f);
            if (MODE == 5) {
                if (sub == 0) {
                    uint32_t s = (uint32_t)__cvta_generic_to_shared(&src[warp][grp][0]);
                    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 128;"
                                 :: "l"(buf + (size_t)rg * 32), "r"(s) : "memory");
                }
            }
        }
        if (MODE == 5) {
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    }
    if (MODE == 5) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (acc == 123.456f) sink[0] = acc;
}

// shared-memory integer atomics: mode 0: ATOMS.ADD one 128-B row per warp instr (32 lanes, consecutive words)
//                                mode 1: ATOMS.ADD, 4 random rows per instr (8 lanes x 4 words each -> 4 instrs per 4 rows)
//                                mode 2: ATOMS.CAS by one lane per 8-lane group + shuffle broadcast
//                                mode 3: plain LDS+FADD+STS row read-modify-write (32 lanes, one row)
template <int MODE>
__global__ void __launch_bounds__(256) ks(int iters, float* sink)
{
    extern __shared__ int tab[];      // 512 rows x 32 words = 64 KiB
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t wid = blockIdx.x * 8 + warp;
    const int grp = lane >> 3, sub = lane & 7;
    for (int i = threadIdx.x; i < 512 * 32; i += 256) tab[i] = 0;
    __syncthreads();
    int acc = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 4
        for (int j = 0; j < 16; ++j) {
            const uint32_t seed = (wid * 1315423911u) ^ (it * 16 + j) * 2654435761u;
            const uint32_t r0 = hash32(seed) & 511u, rg = hash32(seed + grp * 7919u) & 511u;
            if (MODE == 0) atomicAdd(&tab[r0 * 32 + lane], (int)(seed >> 20));
            if (MODE == 1) {
#pragma unroll
                for (int c = 0; c < 4; ++c) atomicAdd(&tab[rg * 32 + sub * 4 + c], (int)(seed >> 20));
            }
            if (MODE == 2) {
                int old = 0;
                if (sub == 0) old = atomicCAS(&tab[rg * 32], 0, (int)rg + 1);
                acc += __shfl_sync(0xffffffffu, old, grp * 8);
            }
            if (MODE == 3) {
                float* f = reinterpret_cast<float*>(tab);
                f[r0 * 32 + lane] += 1.0f;
            }
        }
    }
    __syncthreads();
    if (acc == 123456789) sink[0] = (float)acc + tab[threadIdx.x];
}

template <int MODE>
void run_s(const char* name, float* sink)
{
    const int iters = 256, blocks = 148 * 2;
    cudaFuncSetAttribute(ks<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    ks<MODE><<<blocks, 256, 65536>>>(4, sink);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    ks<MODE><<<blocks, 256, 65536>>>(iters, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)blocks * 8 * iters * 16 * ((MODE == 1 || MODE == 2) ? 4 : 1);   // row-level ops
    printf("[smem] %-46s : %8.3f ms  %7.2f G row-ops/s  (~%.2f cyc/row-op/SM @1.92GHz)  err=%s\n", name, ms,
           ops / ms * 1e-6, ms * 1e6 / (ops / 148) * 1.92, cudaGetErrorString(cudaGetLastError()));
}

template <int MODE>
void run(const char* name, float* buf, uint32_t rows, float* sink, int local, int blocks = 148 * 8)
{
    const int iters = 64;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(buf, rows - 1, 4, sink, local);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(buf, rows - 1, iters, sink, local);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double warp_instr = (double)blocks * 8 * iters * 16;
    double rows_per_instr = (MODE == 0) ? 1 : (MODE == 3 ? 2 : (MODE == 6 ? 8 : 4));
    const double row_updates = warp_instr * rows_per_instr;
    printf("[%4d blocks] %-34s rows=%8u %s : %8.3f ms  %7.2f G row-updates/s  %6.2f TB/s  (%.2f ns/row/SM => ~%.1f cyc @1.92GHz)  err=%s\n",
           blocks, name, rows, local ? "local " : "random", ms, row_updates / ms * 1e-6, row_updates * 128 / ms * 1e-9,
           ms * 1e6 / (row_updates / (blocks < 148 ? blocks : 148)), ms * 1e6 / (row_updates / (blocks < 148 ? blocks : 148)) * 1.92, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    float *buf, *sink;
    const uint32_t rows_big = 1u << 21;   // 2 Mi rows * 128 B = 256 MiB (> L2)
    cudaMalloc(&buf, (size_t)rows_big * 128);
    cudaMalloc(&sink, 4);
    cudaMemset(buf, 0, (size_t)rows_big * 128);
    run_s<0>("ATOMS.ADD s32, 1 row (32 lanes) per instr", sink);
    run_s<1>("ATOMS.ADD s32, 4 rows per 4 instrs (8 lanes/row)", sink);
    run_s<2>("ATOMS.CAS tag claim, 1 lane per group + shfl", sink);
    run_s<3>("LDS+FADD+STS row RMW (not atomic)", sink);
    printf("\n");
    // fewer SMs: is the RED limit per SM or chip-wide?
    for (int blocks : {18, 37, 74, 148}) {
        run<1>("v4 red, 4 random rows/instr", buf, 1u << 19, sink, 0, blocks);
        run<4>("LDG.128 gather, 4 rows/instr", buf, 1u << 19, sink, 0, blocks);
    }
    printf("\n");
    // SM-side or L2-side limit?  same work per block, fewer blocks (1 block = 8 warps; 8 blocks/SM at full grid)
    for (int blocks : {37 * 8, 74 * 8, 148 * 8, 148 * 2, 148 * 1}) {
        run<1>("v4 red, 4 random rows/instr", buf, 1u << 19, sink, 0, blocks);
        run<6>("v4 bf16x2 red (64B rows), 8 rows/instr", buf, 1u << 19, sink, 0, blocks);
        run<7>("STG.128, 4 random rows/instr", buf, 1u << 19, sink, 0, blocks);
        run<4>("LDG.128 gather, 4 rows/instr", buf, 1u << 19, sink, 0, blocks);
    }
    printf("\n");
    for (int local = 0; local < 1; ++local)
        for (uint32_t rows : {1u << 14 /* 2 MiB */, 1u << 19 /* 64 MiB, L2 resident */, rows_big}) {
            run<0>("scalar red, 1 row/instr", buf, rows, sink, local);
            run<1>("v4 red, 4 random rows/instr", buf, rows, sink, local);
            run<2>("v4 red, 4 adjacent rows/instr", buf, rows, sink, local);
            run<3>("v2 red, 2 random rows/instr", buf, rows, sink, local);
            run<4>("LDG.128 gather, 4 rows/instr", buf, rows, sink, local);
            run<5>("TMA bulk reduce 128B/op", buf, rows, sink, local);
            printf("\n");
        }
    return 0;
}
