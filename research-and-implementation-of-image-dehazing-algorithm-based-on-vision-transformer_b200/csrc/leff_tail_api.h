// Entry of the fused LeFF-tail kernel (leff_tail.cuh), compiled as its own translation unit (lewin_leff_tail.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/lewin_b200.h"

namespace lewin {
// true if lewin_leff_fwd_bf16 can run depthwise conv + GELU + linear2 + residual of these arguments as the single fused kernel
bool leff_tail_supported(const LewinLeffFwdArgs* a);
// launches it on `stream` (gelu_tab2: device address of the wide GELU table owned by lewin_abi.cu); returns a cudaError_t as int
int leff_tail_launch(const LewinLeffFwdArgs* a, const uint16_t* gelu_tab2, int num_sms, cudaStream_t stream);
}  // namespace lewin
