// Arguments of the weight-gradient kernels (backward.cuh: generic wgrad_kernel; wgrad_bf16.cuh: bf16 fast path).
#pragma once
#include "common.cuh"

namespace lewin {

template <typename T>
struct WgradArgs {
    const T* dY; long long lddy;      // [rows, lddy]; columns [0, N) used
    const T* X;  long long ldx;       // [rows, ldx];  columns [0, K) used
    float* dW;                        // [N, K] fp32, accumulated with atomics
    float* db;                        // [N] or null
    long long M;
    int N, K;
    int mapDY, mapX;                  // operand rows are tokens addressed through `map` (row m is window-ordered)
    WinMap map;
    const float* dy_row_scale;        // [B] or null (DropPath factor on dY rows)
    int tokens_per_image;
    const float* mean; const float* rstd; const float* ln_w; const float* ln_b;   // LN prologue on X (null mean => none)
    const T* dy_aux;                  // null, or pre-activation: dY is multiplied by gelu'(aux) (same indexing as dY)
    long long rows_per_split;         // multiple of 32
};


}  // namespace lewin
