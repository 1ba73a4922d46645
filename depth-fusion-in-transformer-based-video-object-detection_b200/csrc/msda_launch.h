// Internal (C++) launch interface between the C-ABI layer (msda_api.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace msda {

enum DType { kF32 = 0, kF64 = 1, kBF16 = 2, kF16 = 3 };   // == msda_dtype in include/msda_b200.h

struct FwdArgs {
    int dtype;
    const void* value;        // [N,S,M,D] dtype
    const int64_t* shapes;    // [L,2] device
    const int64_t* lsi;       // [L]   device
    const void* loc;          // [N,Lq,M,L,P,2]  fp32 (fp64 when dtype == kF64)
    const void* attn;         // [N,Lq,M,L,P]    same type as loc
    void* out;                // [N,Lq,M,D] dtype
    int N, S, M, D, L, Lq, P;
    int force_generic;        // tests: route through the generic kernel
};

struct BwdArgs {
    int dtype;
    const void* grad_out;     // [N,Lq,M,D] dtype
    const void* value;
    const int64_t* shapes;
    const int64_t* lsi;
    const void* loc;
    const void* attn;
    void* grad_value;         // [N,S,M,D] dtype, written (zero-initialised by the launcher)
    void* grad_loc;           // like loc, every element written
    void* grad_attn;          // like attn, every element written
    float* grad_value_accum;  // 16-bit dtypes only: fp32 [N,S,M,D] scratch the atomics land in
    int N, S, M, D, L, Lq, P;
    int force_generic;
};

cudaError_t forward(const FwdArgs& a, cudaStream_t stream);
cudaError_t backward(const BwdArgs& a, cudaStream_t stream);

}  // namespace msda
