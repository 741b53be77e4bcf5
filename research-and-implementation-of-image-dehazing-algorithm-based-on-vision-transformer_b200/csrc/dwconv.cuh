// LeFF depthwise 3x3 convolution + bias + exact GELU on channel-last maps.
//
// Reference: nn.Conv2d(hidden, hidden, groups=hidden, kernel_size=3, stride=1, padding=1) + nn.GELU,
// My_model_1.py:489-491, applied at :517 over the WHOLE H x W map (zero padding at the image border;
// it crosses window borders).  The reference round-trips [B,L,4C] -> NCHW -> conv -> NHWC
// (My_model_1.py:514,520); here the map stays token-major [B,H,W,Ch] and each thread owns 4
// channels of a vertical strip of pixels with a rolling 3x3 register window (3 new 128-bit loads
// per output instead of 9).  HBM-bound: algorithmic traffic = read Ch + write Ch per pixel.
#pragma once
#include "common.cuh"

namespace lewin {

constexpr int DW_TY = 8;   // pixels per thread along y

template <typename T>
__global__ void __launch_bounds__(256) dwconv3x3_gelu_kernel(const T* __restrict__ x, T* __restrict__ out,
                                                             T* __restrict__ preact,
                                                             const float* __restrict__ w, const float* __restrict__ bias,
                                                             int B, int H, int W, int Ch) {
    const int cgs = Ch >> 2;
    const int strips = (H + DW_TY - 1) / DW_TY;
    const long long total = static_cast<long long>(B) * strips * W * cgs;
    const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    const int cg = static_cast<int>(gid % cgs);
    long long rest = gid / cgs;
    const int xx = static_cast<int>(rest % W);
    rest /= W;
    const int strip = static_cast<int>(rest % strips);
    const int b = static_cast<int>(rest / strips);
    const int c = cg * 4;
    const int y0 = strip * DW_TY;

    float wk[4][9];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int t = 0; t < 9; ++t) wk[i][t] = Act<T>::round(__ldg(w + (c + i) * 9 + t));   // autocast: conv weight/bias cast to bf16
    float4 bz = *reinterpret_cast<const float4*>(bias + c);
    bz.x = Act<T>::round(bz.x); bz.y = Act<T>::round(bz.y); bz.z = Act<T>::round(bz.z); bz.w = Act<T>::round(bz.w);

    const T* xb = x + static_cast<long long>(b) * H * W * Ch + c;
    auto ldpix = [&](int y, int xq) -> float4 {
        if (y < 0 || y >= H || xq < 0 || xq >= W) return make_float4(0.f, 0.f, 0.f, 0.f);
        return ld4(xb + (static_cast<long long>(y) * W + xq) * Ch);
    };
    float4 win[3][3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        win[0][j] = ldpix(y0 - 1, xx - 1 + j);
        win[1][j] = ldpix(y0, xx - 1 + j);
    }
#pragma unroll
    for (int dy = 0; dy < DW_TY; ++dy) {
        const int y = y0 + dy;
        if (y >= H) break;
#pragma unroll
        for (int j = 0; j < 3; ++j) win[2][j] = ldpix(y + 1, xx - 1 + j);
        float4 acc = bz;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const float4 v = win[ky][kx];
                acc.x = fmaf(v.x, wk[0][ky * 3 + kx], acc.x);
                acc.y = fmaf(v.y, wk[1][ky * 3 + kx], acc.y);
                acc.z = fmaf(v.z, wk[2][ky * 3 + kx], acc.z);
                acc.w = fmaf(v.w, wk[3][ky * 3 + kx], acc.w);
            }
        if (Act<T>::kIsBf16) {
            acc.x = Act<T>::round(acc.x); acc.y = Act<T>::round(acc.y);
            acc.z = Act<T>::round(acc.z); acc.w = Act<T>::round(acc.w);
        }
        const long long o = (static_cast<long long>(b) * H * W + static_cast<long long>(y) * W + xx) * Ch + c;
        if (preact) st4(preact + o, acc);
        if (Act<T>::kIsBf16)
            st4(out + o, make_float4(gelu_tab(g_gelu_tab, acc.x), gelu_tab(g_gelu_tab, acc.y), gelu_tab(g_gelu_tab, acc.z),
                                     gelu_tab(g_gelu_tab, acc.w)));
        else
            st4(out + o, make_float4(gelu_erf(acc.x), gelu_erf(acc.y), gelu_erf(acc.z), gelu_erf(acc.w)));
#pragma unroll
        for (int j = 0; j < 3; ++j) { win[0][j] = win[1][j]; win[1][j] = win[2][j]; }
    }
}

template <typename T>
cudaError_t launch_dwconv_gelu(const T* x, T* out, T* preact, const float* w, const float* bias,
                               int B, int H, int W, int Ch, cudaStream_t stream) {
    const int strips = (H + DW_TY - 1) / DW_TY;
    const long long total = static_cast<long long>(B) * strips * W * (Ch >> 2);
    const unsigned grid = static_cast<unsigned>((total + 255) / 256);
    dwconv3x3_gelu_kernel<T><<<grid, 256, 0, stream>>>(x, out, preact, w, bias, B, H, W, Ch);
    return cudaGetLastError();
}

}  // namespace lewin
