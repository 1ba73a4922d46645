// Cycle trace of the first tiles of the tcgen05 projection + LayerNorm kernel (block 0): where each role waits.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/umma/proj_trace tools/umma/proj_trace.cu -lcuda ; run on a B200.
#define MSDA_PROJ_TRACE 1
#include "../../depth-fusion-in-transformer-based-video-object-detection_b200/csrc/proj_fused.cu"
#include <cstdio>
int main()
{
    const long long rows = 8 * 22223; const int C = 256;
    __nv_bfloat16 *x, *w, *b, *g, *bt, *res, *y;
    cudaMalloc(&x, rows * C * 2); cudaMalloc(&res, rows * C * 2); cudaMalloc(&y, rows * C * 2);
    cudaMalloc(&w, C * C * 2); cudaMalloc(&b, C * 2); cudaMalloc(&g, C * 2); cudaMalloc(&bt, C * 2);
    cudaMemset(x, 0, rows * C * 2); cudaMemset(res, 0, rows * C * 2); cudaMemset(w, 0, C * C * 2);
    cudaMemset(b, 0, C * 2); cudaMemset(g, 0, C * 2); cudaMemset(bt, 0, C * 2);
    msda::ProjArgs a = {};
    a.dtype = msda::kBF16; a.rows = rows; a.C = C; a.eps = 1e-5f;
    a.x = x; a.w = w; a.b = b; a.residual = res; a.gamma = g; a.beta = bt; a.y = y;
    for (int i = 0; i < 3; ++i) msda::proj_layernorm_forward(a, 0);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    static long long t[4][8][8];
    cudaMemcpyFromSymbol(t, msda::g_proj_trace, sizeof(t));
    const long long t0 = t[1][0][0];
    for (int it = 0; it < 8; ++it)
        printf("it=%d | prod: A issue %6lld store_done(it-1) %6lld | mma: top %6lld a_full %6lld yacc_free %6lld issued %6lld | "
               "epi: top %6lld mma_done %6lld staged %6lld | store: top %6lld stage_full %6lld res_full %6lld done %6lld\n",
               it, t[1][it][0] - t0, t[1][it][1] - t0, t[0][it][0] - t0, t[0][it][1] - t0, t[0][it][2] - t0, t[0][it][3] - t0,
               t[2][it][0] - t0, t[2][it][1] - t0, t[2][it][2] - t0, t[3][it][0] - t0, t[3][it][1] - t0, t[3][it][2] - t0,
               t[3][it][3] - t0);
    return 0;
}
