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

// ------------------------------------------------------------------------------------------------------------------
// Shared-memory tiled variant (used when the map is at least 8x8): a CTA stages a (8+2) x (TX+2) pixel halo tile of
// one 8-chunk channel slab (64 bf16 / 32 fp32 channels) with 16-byte cp.async (zero-filled outside the map == the
// conv's zero padding), then every thread produces 4 (TX=16) or 2 (TX=8) pixels of one 16-byte channel chunk from
// shared memory.  Global traffic is one coalesced read (+ halo from L2) and one coalesced write per element; the
// bf16 GELU is the exact 16-bit table.
namespace lewin {

__device__ __forceinline__ void cp_async16_zfill(void* smem, const void* gmem, bool valid) {
    const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz));
}

template <typename T, int TX>
__global__ void __launch_bounds__(256) dwconv3x3_gelu_tiled_kernel(const T* __restrict__ x, T* __restrict__ out,
                                                                   T* __restrict__ preact, const float* __restrict__ w,
                                                                   const float* __restrict__ bias, int B, int H, int W, int Ch) {
    constexpr int EPC = 16 / sizeof(T);            // channels per 16-byte chunk
    constexpr int SLAB = 8 * EPC;                  // channels per CTA
    constexpr int TY = 8;
    constexpr int HX = TX + 2, HY = TY + 2;
    constexpr int PPT = TY * TX / 32;              // pixels per thread (4 or 2)
    __shared__ __align__(16) unsigned char tile[HY * HX * 8 * 16];
    __shared__ __align__(16) float ws[9 * SLAB];
    __shared__ float bs[SLAB];
    __shared__ __align__(16) uint16_t gtab[Act<T>::kIsBf16 ? kGeluTabSize : 8];

    const int tid = threadIdx.x;
    const int slabs = Ch / SLAB, tiles_x = W / TX, tiles_y = H / TY;
    int t = blockIdx.x;
    const int slab = t % slabs; t /= slabs;
    const int tx = t % tiles_x; t /= tiles_x;
    const int ty = t % tiles_y;
    const int b = t / tiles_y;
    const int c0 = slab * SLAB;
    const int y0 = ty * TY - 1, x0 = tx * TX - 1;

    for (int i = tid; i < HY * HX * 8; i += 256) {
        const int ch = i & 7, p = i >> 3;
        const int hy = p / HX, hx = p - hy * HX;
        const int yy = y0 + hy, xx = x0 + hx;
        const bool ok = yy >= 0 && yy < H && xx >= 0 && xx < W;
        const T* src = x + ((static_cast<long long>(b) * H + (ok ? yy : 0)) * W + (ok ? xx : 0)) * Ch + c0 + ch * EPC;
        cp_async16_zfill(tile + i * 16, src, ok);
    }
    cp_async_commit();
    for (int i = tid; i < 9 * SLAB; i += 256) {
        const int tap = i / SLAB, c = i - tap * SLAB;
        ws[i] = Act<T>::round(w[static_cast<long long>(c0 + c) * 9 + tap]);
    }
    if (tid < SLAB) bs[tid] = Act<T>::round(bias[c0 + tid]);
    if (Act<T>::kIsBf16) gelu_tab_to_smem(gtab, tid, 256);
    cp_async_wait<0>();
    __syncthreads();

    const int ch = tid & 7, lane = tid >> 3;                 // 32 pixel lanes
    const int px = lane % TX, py0 = (lane / TX) * PPT;       // column, first row of this thread's strip
    float acc[PPT][EPC];
#pragma unroll
    for (int p = 0; p < PPT; ++p)
#pragma unroll
        for (int j = 0; j < EPC; ++j) acc[p][j] = bs[ch * EPC + j];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            float wv[EPC];
#pragma unroll
            for (int j = 0; j < EPC; j += 4) {
                const float4 t4 = *reinterpret_cast<const float4*>(ws + (ky * 3 + kx) * SLAB + ch * EPC + j);
                wv[j] = t4.x; wv[j + 1] = t4.y; wv[j + 2] = t4.z; wv[j + 3] = t4.w;
            }
#pragma unroll
            for (int p = 0; p < PPT; ++p) {
                const unsigned char* src = tile + (((py0 + p + ky) * HX + px + kx) * 8 + ch) * 16;
                float f[EPC];
                if (sizeof(T) == 2) {
                    const uint4 u = *reinterpret_cast<const uint4*>(src);
                    f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xFFFF0000u);
                    f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xFFFF0000u);
                    if (EPC == 8) {
                        f[EPC - 4] = __uint_as_float(u.z << 16); f[EPC - 3] = __uint_as_float(u.z & 0xFFFF0000u);
                        f[EPC - 2] = __uint_as_float(u.w << 16); f[EPC - 1] = __uint_as_float(u.w & 0xFFFF0000u);
                    }
                } else {
                    const float4 u = *reinterpret_cast<const float4*>(src);
                    f[0] = u.x; f[1] = u.y; f[2] = u.z; f[3] = u.w;
                }
#pragma unroll
                for (int j = 0; j < EPC; ++j) acc[p][j] = fmaf(f[j], wv[j], acc[p][j]);
            }
        }
#pragma unroll
    for (int p = 0; p < PPT; ++p) {
        const int yy = ty * TY + py0 + p, xx = tx * TX + px;
        const long long o = ((static_cast<long long>(b) * H + yy) * W + xx) * Ch + c0 + ch * EPC;
        if (sizeof(T) == 2) {
            uint32_t in[4], q[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                __nv_bfloat162 h2 = __floats2bfloat162_rn(acc[p][2 * j], acc[p][2 * j + 1]);
                in[j] = *reinterpret_cast<uint32_t*>(&h2);
                q[j] = gelu_bits(gtab, in[j] & 0xFFFFu) | (gelu_bits(gtab, in[j] >> 16) << 16);
            }
            if (preact) *reinterpret_cast<uint4*>(preact + o) = make_uint4(in[0], in[1], in[2], in[3]);
            *reinterpret_cast<uint4*>(out + o) = make_uint4(q[0], q[1], q[2], q[3]);
        } else {
            if (preact) st4(preact + o, make_float4(acc[p][0], acc[p][1], acc[p][2], acc[p][3]));
            st4(out + o, make_float4(gelu_erf(acc[p][0]), gelu_erf(acc[p][1]), gelu_erf(acc[p][2]), gelu_erf(acc[p][3])));
        }
    }
}

template <typename T>
cudaError_t launch_dwconv_gelu_auto(const T* x, T* out, T* preact, const float* w, const float* bias,
                                    int B, int H, int W, int Ch, cudaStream_t stream) {
    constexpr int SLAB = 8 * (16 / sizeof(T));
    if (H % 8 == 0 && W % 8 == 0 && Ch % SLAB == 0) {
        if (W % 16 == 0) {
            const unsigned grid = static_cast<unsigned>(B) * (H / 8) * (W / 16) * (Ch / SLAB);
            dwconv3x3_gelu_tiled_kernel<T, 16><<<grid, 256, 0, stream>>>(x, out, preact, w, bias, B, H, W, Ch);
        } else {
            const unsigned grid = static_cast<unsigned>(B) * (H / 8) * (W / 8) * (Ch / SLAB);
            dwconv3x3_gelu_tiled_kernel<T, 8><<<grid, 256, 0, stream>>>(x, out, preact, w, bias, B, H, W, Ch);
        }
        return cudaGetLastError();
    }
    return launch_dwconv_gelu<T>(x, out, preact, w, bias, B, H, W, Ch, stream);
}

}  // namespace lewin
