/* C-ABI shim around the UNMODIFIED reference CUDA launchers.
 *
 * TEST INFRASTRUCTURE ONLY.  This file contains no kernel code: it #includes the
 * reference header where it lies (/root/reference/models/ops/src/cuda/
 * ms_deform_im2col_cuda.cuh, found through -I in oracle/Makefile) and exposes its two
 * host launchers -- ms_deformable_im2col_cuda (cuh:923-954) and
 * ms_deformable_col2im_cuda (cuh:956-1327) -- with plain-pointer signatures, so the
 * reference's own kernels, recompiled for sm_100a, can be run on the B200 next to ours:
 *   - as the strongest parity oracle at full sizes (tests/test_gpu_vs_ref_cuda.py), and
 *   - as the "reference kernel recompiled for Blackwell" timing baseline.
 * Output goes to oracle/_ref/ (git-ignored; travels to the GPU box with the snapshot).
 * The reference allocates its outputs with at::zeros / zeros_like
 * (cuda/ms_deform_attn_cuda.cu:54,121-123); callers of this shim must zero them.
 */
#include "ms_deform_im2col_cuda.cuh"

#define SHIM(T, SUFFIX)                                                                          \
extern "C" int ref_msda_forward_##SUFFIX(void* stream, const T* value, const int64_t* shapes,    \
        const int64_t* lsi, const T* loc, const T* attn, int N, int S, int M, int D, int L,      \
        int Lq, int P, T* out)                                                                   \
{                                                                                                \
    ms_deformable_im2col_cuda<T>((cudaStream_t)stream, value, shapes, lsi, loc, attn,            \
                                 N, S, M, D, L, Lq, P, out);                                     \
    return (int)cudaPeekAtLastError();                                                           \
}                                                                                                \
extern "C" int ref_msda_backward_##SUFFIX(void* stream, const T* grad_out, const T* value,       \
        const int64_t* shapes, const int64_t* lsi, const T* loc, const T* attn, int N, int S,    \
        int M, int D, int L, int Lq, int P, T* grad_value, T* grad_loc, T* grad_attn)            \
{                                                                                                \
    ms_deformable_col2im_cuda<T>((cudaStream_t)stream, grad_out, value, shapes, lsi, loc, attn,  \
                                 N, S, M, D, L, Lq, P, grad_value, grad_loc, grad_attn);         \
    return (int)cudaPeekAtLastError();                                                           \
}

SHIM(float, f32)
SHIM(double, f64)
