// Backward kernels of the LeWin hot path (placeholder: forward lands first).
#pragma once
#include "../../include/lewin_b200.h"
#include "common.cuh"

namespace lewin {
#define LEWIN_E_UNIMPL (-7)
template <typename T> int attn_bwd(const LewinAttnBwdArgs*, void*, size_t, cudaStream_t) { return LEWIN_E_UNIMPL; }
template <typename T> int leff_bwd(const LewinLeffBwdArgs*, void*, size_t, cudaStream_t) { return LEWIN_E_UNIMPL; }
inline size_t attn_bwd_ws(const LewinAttnBwdArgs*, int) { return 0; }
inline size_t leff_bwd_ws(const LewinLeffBwdArgs*, int) { return 0; }
}  // namespace lewin
