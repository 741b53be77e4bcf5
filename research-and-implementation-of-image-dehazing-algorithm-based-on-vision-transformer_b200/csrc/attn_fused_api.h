// Entry of the fused attention-half kernel (attn_fused.cuh), compiled as its own translation unit (lewin_attn_fused.cu).
#pragma once
#include <cuda_runtime.h>
#include "../../include/lewin_b200.h"

namespace lewin {
// true if lewin_attn_fwd_bf16 can run these arguments as the single fused kernel
bool attn_fused_supported(const LewinAttnFwdArgs* a);
// launches it on `stream`; returns a cudaError_t as int
int attn_fused_launch(const LewinAttnFwdArgs* a, int num_sms, cudaStream_t stream);
}  // namespace lewin
