// Shared pieces of the bf16 ProbSparse core kernels (probsparse_core_v3.cuh forward, probsparse_core_bwd_v2.cuh backward;
// reference ProbSparse/attn.py:287-342): ldmatrix / mma.sync.m16n8k16 wrappers and the argument struct.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"
#include "probsparse_core.cuh"

namespace lewin {

namespace pc {
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
    const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
    const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
}  // namespace pc

struct CoreBf16Args {
    const __nv_bfloat16* qkv; __nv_bfloat16* ctx; uint8_t* top;
    const float* rpb_table; const float* rpb_dense; const int32_t* index_sample;   // [64, 25] (attn.py:91)
    const float* mask; int nW_mask;
    int B_, nH, C, use_rpb, shift, H, W, nWw, nWin;
    int y0, Hg;           // row band of a taller image: see CoreFwdArgs
};

}  // namespace lewin
