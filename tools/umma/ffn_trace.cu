// Cycle trace of one tile of the tcgen05 feed-forward kernel (block 0, second tile): where the
// producer/issuer lane and the epilogue warps wait.  nvcc -DMSDA_FFN_TRACE ... ; run on a B200.
#define MSDA_FFN_TRACE 1
#include "../../depth-fusion-in-transformer-based-video-object-detection_b200/csrc/ffn_fused.cu"
#include <cstdio>
#include <vector>
int main()
{
    const long long rows = 8 * 22223; const int C = 256, F = 1024;
    __nv_bfloat16 *x, *w1, *b1, *w2, *b2, *g, *bt, *pos, *y, *yp;
    cudaMalloc(&x, rows * C * 2); cudaMalloc(&pos, rows * C * 2); cudaMalloc(&y, rows * C * 2); cudaMalloc(&yp, rows * C * 2);
    cudaMalloc(&w1, F * C * 2); cudaMalloc(&w2, F * C * 2); cudaMalloc(&b1, F * 2); cudaMalloc(&b2, C * 2);
    cudaMalloc(&g, C * 2); cudaMalloc(&bt, C * 2);
    cudaMemset(x, 0, rows * C * 2); cudaMemset(pos, 0, rows * C * 2); cudaMemset(w1, 0, F * C * 2); cudaMemset(w2, 0, F * C * 2);
    cudaMemset(b1, 0, F * 2); cudaMemset(b2, 0, C * 2); cudaMemset(g, 0, C * 2); cudaMemset(bt, 0, C * 2);
    msda::FfnArgs a = {};
    a.dtype = msda::kBF16; a.rows = rows; a.C = C; a.F = F; a.eps = 1e-5f;
    a.x = x; a.w1 = w1; a.b1 = b1; a.w2 = w2; a.b2 = b2; a.gamma = g; a.beta = bt; a.pos = pos; a.y = y; a.y_pos = yp;
    for (int i = 0; i < 3; ++i) msda::ffn_layernorm_forward(a, 0);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    static long long t[2][40][8];
    cudaMemcpyFromSymbol(t, msda::g_ffn_trace, sizeof(t));
    const long long t0 = t[0][0][0];
    printf("control: tile start 0, X+W1(0) landed %lld\n", t[0][32][1] - t0);
    for (int c = 0; c < 16; ++c)
        printf("c=%2d ctl: top %6lld w1ok %6lld g1issued %6lld hfull %6lld w2ok %6lld g2issued %6lld | epi: top %6lld g1done %6lld ld %6lld g2free %6lld arrived %6lld\n",
               c, t[0][c][0] - t0, t[0][c][1] - t0, t[0][c][2] - t0, t[0][c][3] - t0, t[0][c][4] - t0, t[0][c][5] - t0,
               t[1][c][0] - t0, t[1][c][1] - t0, t[1][c][2] - t0, t[1][c][3] - t0, t[1][c][4] - t0);
    printf("epi: all MMAs done %lld, pass 1 done (X/Yacc released) +%lld, tile released +%lld\n", t[1][32][0] - t0,
           t[1][32][2] - t[1][32][0], t[1][32][1] - t[1][32][0]);
    printf("epi: after barrier +%lld, loads issued +%lld\n", t[1][33][0] - t[1][32][0], t[1][33][1] - t[1][32][0]);
    return 0;
}
