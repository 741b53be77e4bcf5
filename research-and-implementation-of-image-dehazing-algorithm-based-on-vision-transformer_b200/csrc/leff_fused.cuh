// Fully fused LeFF half of a LeWin block for the HBM-bound levels (C <= 128), bf16 inference:
//
//   out = y + s_b * Linear2( GELU( dwconv3x3( GELU( Linear1( LN2(y) ) ) ) ) )          (My_model_1.py:873, :496-534)
//
// One CTA owns an 8x8 pixel tile.  It loads the 10x10 halo of y once, applies LN2 in registers, and walks the hidden
// dimension (4C) in chunks of 64 channels:  GEMM1 (112 halo rows x 64) on tensor cores -> +b1 -> GELU -> zero outside the
// image (the reference's conv zero-padding) -> shared memory -> depthwise 3x3 + bias -> GELU -> shared memory ->
// GEMM2 accumulating the 64 x C output tile in registers.  The hidden activations (4C per pixel, written and re-read
// three times by the unfused path and four times by the reference) never touch HBM: algorithmic traffic is read C +
// write C per pixel plus the L2-resident halo.  The price is recomputing GEMM1 on the halo ring (112/64 rows).
// Tensor cores: mma.sync.m16n8k16 bf16 with ldmatrix from padded (conflict-free) shared-memory tiles.
#pragma once
#include "common.cuh"

namespace lewin {

namespace lf {

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
    const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float rbf(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float2 t = __bfloat1622float2(h[j]); f[2 * j] = t.x; f[2 * j + 1] = t.y; }
}

// packed fp32x2 FMA (Blackwell FFMA2): d.{x,y} += a.{x,y} * b.{x,y} in one instruction
__device__ __forceinline__ void ffma2(float2& d, const float2 a, const float2 b) {
    unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d);
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd)
        : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
    d = *reinterpret_cast<float2*>(&dd);
}
__device__ __forceinline__ void unpack8_2(const uint4& u, float2 (&f)[4]) {
    f[0] = make_float2(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xFFFF0000u));
    f[1] = make_float2(__uint_as_float(u.y << 16), __uint_as_float(u.y & 0xFFFF0000u));
    f[2] = make_float2(__uint_as_float(u.z << 16), __uint_as_float(u.z & 0xFFFF0000u));
    f[3] = make_float2(__uint_as_float(u.w << 16), __uint_as_float(u.w & 0xFFFF0000u));
}

}  // namespace lf

struct LeffFusedArgs {
    const __nv_bfloat16* y;      // [B, H, W, C]
    __nv_bfloat16* out;          // [B, H, W, C]
    const float* ln_w; const float* ln_b;
    const float* w1; const float* b1;        // [4C, C], [4C]
    const float* w_dw; const float* b_dw;    // [4C, 9], [4C]
    const float* w2; const float* b2;        // [C, 4C], [C]
    const float* drop_scale;                 // [B] or null
    int B, H, W;
    const unsigned char* wimg;               // per-chunk bf16/fp32 weight images written by leff_prep_weights_kernel
};

// Per-chunk weight image (exactly the shared-memory layout, so staging is plain 16-byte cp.async copies):
//   w1 [64][C+8] bf16 | b1 [64] f32 | w2 [C][72] bf16 | dww [9][64] f32 | dwb [64] f32      (autocast-rounded values)
template <int C> struct LeffImg {
    static constexpr int W1 = 0;
    static constexpr int B1 = 64 * (C + 8) * 2;
    static constexpr int W2 = B1 + 256;
    static constexpr int DWW = W2 + C * 72 * 2;
    static constexpr int DWB = DWW + 9 * 64 * 4;
    static constexpr int BYTES = DWB + 256;
};
inline size_t leff_img_bytes(int C) { return static_cast<size_t>(4 * C / 64) * (64 * (C + 8) * 2 + 256 + C * 144 + 2304 + 256); }

template <int C>
__global__ void leff_prep_weights_kernel(const float* __restrict__ w1, const float* __restrict__ b1,
                                         const float* __restrict__ w2, const float* __restrict__ w_dw,
                                         const float* __restrict__ b_dw, unsigned char* __restrict__ img) {
    using I = LeffImg<C>;
    const int ch = blockIdx.x, h0 = ch * 64, HID = 4 * C;
    unsigned char* base = img + static_cast<size_t>(ch) * I::BYTES;
    __nv_bfloat16* o1 = reinterpret_cast<__nv_bfloat16*>(base + I::W1);
    for (int i = threadIdx.x; i < 64 * (C + 8); i += blockDim.x) {
        const int r = i / (C + 8), c = i - r * (C + 8);
        o1[i] = __float2bfloat16_rn(c < C ? w1[static_cast<long long>(h0 + r) * C + c] : 0.f);
    }
    __nv_bfloat16* o2 = reinterpret_cast<__nv_bfloat16*>(base + I::W2);
    for (int i = threadIdx.x; i < C * 72; i += blockDim.x) {
        const int r = i / 72, c = i - r * 72;
        o2[i] = __float2bfloat16_rn(c < 64 ? w2[static_cast<long long>(r) * HID + h0 + c] : 0.f);
    }
    float* ob1 = reinterpret_cast<float*>(base + I::B1);
    float* odw = reinterpret_cast<float*>(base + I::DWW);
    float* odb = reinterpret_cast<float*>(base + I::DWB);
    for (int i = threadIdx.x; i < 9 * 64; i += blockDim.x) {
        const int tap = i / 64, c = i - tap * 64;
        odw[i] = lf::rbf(w_dw[static_cast<long long>(h0 + c) * 9 + tap]);
    }
    for (int i = threadIdx.x; i < 64; i += blockDim.x) { ob1[i] = lf::rbf(b1[h0 + i]); odb[i] = lf::rbf(b_dw[h0 + i]); }
}

constexpr int LF_THREADS = 256;
constexpr int LF_CH = 64;          // hidden channels per chunk
constexpr int LF_HROWS = 112;      // 10x10 halo pixels padded to 7 m16-tiles
constexpr int LF_HS_LD = LF_CH + 8;

template <int C>
struct LeffFusedSmem {
    static constexpr int XS_LD = C + 8;
    __nv_bfloat16 xs[LF_HROWS * XS_LD];      // LN2(y) on the halo (zeros outside the image / padding rows)
    __nv_bfloat16 w1s[LF_CH * XS_LD];        // W1 chunk  [64 hidden][C]
    __nv_bfloat16 w2s[C * LF_HS_LD];         // W2 chunk  [C][64 hidden]
    __nv_bfloat16 h1s[LF_HROWS * LF_HS_LD];  // GELU(linear1) on the halo
    __nv_bfloat16 h2s[64 * LF_HS_LD];        // GELU(dwconv) on the 8x8 interior
    float dww[9 * LF_CH];
    float dwb[LF_CH];
    float b1s[LF_CH];
    float b2s[C];
    uint16_t gtab[kGeluTabSize];             // exact bf16 GELU table (common.cuh)
    long long pix_off[100];                  // element offset of each halo pixel in y (or -1 outside the image)
};

template <int C>
__global__ void __launch_bounds__(LF_THREADS, (C == 32 ? 4 : C == 64 ? 3 : 2)) leff_fused_kernel(const LeffFusedArgs a) {
    using S = LeffFusedSmem<C>;
    constexpr int XS_LD = S::XS_LD;
    constexpr int HID = 4 * C;
    constexpr int NCHUNK = HID / LF_CH;
    constexpr int NT2 = C / 16;              // n8-tiles per warp in GEMM2 (warp covers C/2 output channels)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    S& s = *reinterpret_cast<S*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gq = lane >> 2, tq = lane & 3;

    const int tiles_x = a.W / 8, tiles_y = a.H / 8;
    int t = blockIdx.x;
    const int tx = t % tiles_x; t /= tiles_x;
    const int ty = t % tiles_y;
    const int b = t / tiles_y;
    const int y0 = ty * 8 - 1, x0 = tx * 8 - 1;      // halo origin

    // ---- halo pixel offsets
    if (tid < 100) {
        const int hy = tid / 10, hx = tid - hy * 10;
        const int yy = y0 + hy, xx = x0 + hx;
        s.pix_off[tid] = (yy >= 0 && yy < a.H && xx >= 0 && xx < a.W)
                             ? ((static_cast<long long>(b) * a.H + yy) * a.W + xx) * C : -1;
    }
    for (int i = tid; i < C; i += LF_THREADS) s.b2s[i] = lf::rbf(a.b2[i]);
    gelu_tab_to_smem(s.gtab, tid, LF_THREADS);
    __syncthreads();

    // ---- load halo of y, LayerNorm (fp32 stats over C), round to bf16 -> xs
    {
        constexpr int G = C / 8;                       // lanes per pixel (16-byte chunk each)
        constexpr int PPW = 32 / G;                    // pixels per warp pass
        const int sub = lane / G, gl = lane % G;
        float gam[8], bet[8];
        {
            const float4 g0 = *reinterpret_cast<const float4*>(a.ln_w + gl * 8);
            const float4 g1 = *reinterpret_cast<const float4*>(a.ln_w + gl * 8 + 4);
            const float4 e0 = *reinterpret_cast<const float4*>(a.ln_b + gl * 8);
            const float4 e1 = *reinterpret_cast<const float4*>(a.ln_b + gl * 8 + 4);
            gam[0] = g0.x; gam[1] = g0.y; gam[2] = g0.z; gam[3] = g0.w; gam[4] = g1.x; gam[5] = g1.y; gam[6] = g1.z; gam[7] = g1.w;
            bet[0] = e0.x; bet[1] = e0.y; bet[2] = e0.z; bet[3] = e0.w; bet[4] = e1.x; bet[5] = e1.y; bet[6] = e1.z; bet[7] = e1.w;
        }
        for (int p0 = warp * PPW; p0 < LF_HROWS; p0 += 8 * PPW) {
            const int p = p0 + sub;
            const long long off = p < 100 ? s.pix_off[p] : -1;
            uint4 raw = make_uint4(0u, 0u, 0u, 0u);
            if (off >= 0) raw = *reinterpret_cast<const uint4*>(a.y + off + gl * 8);
            float f[8];
            lf::unpack8(raw, f);
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) sum += f[j];
            const float mu = group_sum<G>(sum) / C;
            float q = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) { f[j] -= mu; q += f[j] * f[j]; }
            const float rs = rsqrtf(group_sum<G>(q) / C + 1e-5f);
            uint4 o = make_uint4(0u, 0u, 0u, 0u);
            if (off >= 0) {
                o.x = lf::pack2(f[0] * rs * gam[0] + bet[0], f[1] * rs * gam[1] + bet[1]);
                o.y = lf::pack2(f[2] * rs * gam[2] + bet[2], f[3] * rs * gam[3] + bet[3]);
                o.z = lf::pack2(f[4] * rs * gam[4] + bet[4], f[5] * rs * gam[5] + bet[5]);
                o.w = lf::pack2(f[6] * rs * gam[6] + bet[6], f[7] * rs * gam[7] + bet[7]);
            }
            if (p < LF_HROWS) *reinterpret_cast<uint4*>(s.xs + p * XS_LD + gl * 8) = o;
        }
    }

    // GEMM2 accumulators: warp -> m-tile (warp & 3) of the 64 interior pixels, output channels [(warp >> 2) * C/2, +C/2)
    float acc2[NT2][4];
#pragma unroll
    for (int j = 0; j < NT2; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc2[j][c] = 0.f;

    const int mg = warp & 1, ng = warp >> 1;           // GEMM1: m-group (tiles 0-3 / 4-6), n-group (16 hidden channels)
    const int m2 = warp & 3, nh = warp >> 2;           // GEMM2 roles

    using I = LeffImg<C>;
    auto stage = [&](void* dst, const unsigned char* src, int bytes) {     // 16-byte async copies, all threads
        for (int i = tid * 16; i < bytes; i += LF_THREADS * 16)
            cp_async16(static_cast<unsigned char*>(dst) + i, src + i);
    };
    // chunk 0: GEMM1 operands (group A) then the dwconv / GEMM2 operands (group B)
    stage(s.w1s, a.wimg + I::W1, 64 * XS_LD * 2);
    stage(s.b1s, a.wimg + I::B1, 256);
    cp_async_commit();

#pragma unroll 1
    for (int ch = 0; ch < NCHUNK; ++ch) {
        const unsigned char* img = a.wimg + static_cast<size_t>(ch) * I::BYTES;
        cp_async_wait<0>();                            // this chunk's w1s / b1s (issued during the previous chunk) landed
        __syncthreads();                               // ... for every thread; previous chunk's GEMM2 / dwconv are done
        stage(s.w2s, img + I::W2, C * LF_HS_LD * 2);   // consumed after GEMM1 + epilogue 1: latency hidden
        stage(s.dww, img + I::DWW, 9 * LF_CH * 4);
        stage(s.dwb, img + I::DWB, 256);
        cp_async_commit();

        // ---- GEMM1: h1[112 x 64] = xs[112 x C] . w1s^T ; warp: 4 (or 3) m-tiles x 2 n-tiles
        {
            const int mt0 = mg * 4, nmt = mg == 0 ? 4 : 3;
            float acc1[4][2][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc1[i][j][c] = 0.f;
#pragma unroll
            for (int ks = 0; ks < C / 16; ++ks) {
                uint32_t bfr[4];
                lf::ldsm_x4(bfr, s.w1s + (ng * 16 + (lane & 7) + ((lane >> 4) & 1) * 8) * XS_LD + ks * 16 + ((lane >> 3) & 1) * 8);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (i < nmt) {
                        uint32_t afr[4];
                        lf::ldsm_x4(afr, s.xs + ((mt0 + i) * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * XS_LD + ks * 16 + (lane >> 4) * 8);
                        lf::mma_bf16(acc1[i][0], afr, bfr[0], bfr[1]);
                        lf::mma_bf16(acc1[i][1], afr, bfr[2], bfr[3]);
                    }
                }
            }
            // epilogue 1: + b1 -> bf16 -> GELU -> bf16 ; zero outside the image (conv zero padding)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (i < nmt) {
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const int row = (mt0 + i) * 16 + gq + half * 8;
                        if (row < 100) {
                            const bool inside = s.pix_off[row] >= 0;
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                const int col = ng * 16 + j * 8 + 2 * tq;
                                const uint32_t in2 = lf::pack2(acc1[i][j][half * 2] + s.b1s[col], acc1[i][j][half * 2 + 1] + s.b1s[col + 1]);
                                uint32_t o2 = gelu_bits(s.gtab, in2 & 0xFFFFu) | (gelu_bits(s.gtab, in2 >> 16) << 16);
                                if (!inside) o2 = 0u;
                                *reinterpret_cast<uint32_t*>(s.h1s + row * LF_HS_LD + col) = o2;
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();                               // h1s complete; every warp is done reading w1s / b1s
        if (ch + 1 < NCHUNK) {
            stage(s.w1s, img + I::BYTES + I::W1, 64 * XS_LD * 2);
            stage(s.b1s, img + I::BYTES + I::B1, 256);
        }
        cp_async_commit();
        cp_async_wait<1>();                            // w2s / dww / dwb of this chunk landed (newest group may still fly)
        __syncthreads();

        // ---- depthwise 3x3 + bias -> bf16 -> GELU -> bf16 on the 8x8 interior: thread = (8-channel group, 2 pixels)
        {
            const int c8 = (tid & 7) * 8;
            const int p_a = tid >> 3, p_b = p_a + 32;          // interior pixel ids
            float2 o0[4], o1[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) { o0[j] = make_float2(s.dwb[c8 + 2 * j], s.dwb[c8 + 2 * j + 1]); o1[j] = o0[j]; }
            const int ra = (p_a >> 3) * 10 + (p_a & 7), rb = (p_b >> 3) * 10 + (p_b & 7);   // halo row of tap (0,0)
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const int d = (tap / 3) * 10 + (tap % 3);
                const float4 w0 = *reinterpret_cast<const float4*>(s.dww + tap * LF_CH + c8);
                const float4 w1 = *reinterpret_cast<const float4*>(s.dww + tap * LF_CH + c8 + 4);
                const float2 wv[4] = {make_float2(w0.x, w0.y), make_float2(w0.z, w0.w), make_float2(w1.x, w1.y), make_float2(w1.z, w1.w)};
                float2 fa[4], fb[4];
                lf::unpack8_2(*reinterpret_cast<const uint4*>(s.h1s + (ra + d) * LF_HS_LD + c8), fa);
                lf::unpack8_2(*reinterpret_cast<const uint4*>(s.h1s + (rb + d) * LF_HS_LD + c8), fb);
#pragma unroll
                for (int j = 0; j < 4; ++j) { lf::ffma2(o0[j], fa[j], wv[j]); lf::ffma2(o1[j], fb[j], wv[j]); }
            }
            uint32_t qa[4], qb[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t ia = lf::pack2(o0[j].x, o0[j].y), ib = lf::pack2(o1[j].x, o1[j].y);
                qa[j] = gelu_bits(s.gtab, ia & 0xFFFFu) | (gelu_bits(s.gtab, ia >> 16) << 16);
                qb[j] = gelu_bits(s.gtab, ib & 0xFFFFu) | (gelu_bits(s.gtab, ib >> 16) << 16);
            }
            *reinterpret_cast<uint4*>(s.h2s + p_a * LF_HS_LD + c8) = make_uint4(qa[0], qa[1], qa[2], qa[3]);
            *reinterpret_cast<uint4*>(s.h2s + p_b * LF_HS_LD + c8) = make_uint4(qb[0], qb[1], qb[2], qb[3]);
        }
        __syncthreads();

        // ---- GEMM2: out[64 x C] += h2s[64 x 64] . w2s^T
#pragma unroll
        for (int ks = 0; ks < LF_CH / 16; ++ks) {
            uint32_t afr[4];
            lf::ldsm_x4(afr, s.h2s + (m2 * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LF_HS_LD + ks * 16 + (lane >> 4) * 8);
#pragma unroll
            for (int jp = 0; jp < NT2 / 2; ++jp) {
                uint32_t bfr[4];
                const int n0 = nh * (C / 2) + jp * 16;
                lf::ldsm_x4(bfr, s.w2s + (n0 + (lane & 7) + ((lane >> 4) & 1) * 8) * LF_HS_LD + ks * 16 + ((lane >> 3) & 1) * 8);
                lf::mma_bf16(acc2[2 * jp], afr, bfr[0], bfr[1]);
                lf::mma_bf16(acc2[2 * jp + 1], afr, bfr[2], bfr[3]);
            }
        }
    }

    // ---- final epilogue: out = y + s_b * bf16(acc + b2)
    {
        const float sc = a.drop_scale ? a.drop_scale[b] : 1.0f;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int p = m2 * 16 + gq + half * 8;                       // interior pixel
            const long long off = s.pix_off[((p >> 3) + 1) * 10 + (p & 7) + 1];
#pragma unroll
            for (int j = 0; j < NT2; ++j) {
                const int col = nh * (C / 2) + j * 8 + 2 * tq;
                const float2 r = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(a.y + off + col));
                const float v0 = r.x + sc * lf::rbf(acc2[j][half * 2] + s.b2s[col]);
                const float v1 = r.y + sc * lf::rbf(acc2[j][half * 2 + 1] + s.b2s[col + 1]);
                *reinterpret_cast<uint32_t*>(a.out + off + col) = lf::pack2(v0, v1);
            }
        }
    }
}

template <int C>
cudaError_t launch_leff_fused_c(const LeffFusedArgs& a, cudaStream_t stream) {
    leff_prep_weights_kernel<C><<<4 * C / 64, 256, 0, stream>>>(a.w1, a.b1, a.w2, a.w_dw, a.b_dw,
                                                              const_cast<unsigned char*>(a.wimg));
    auto k = leff_fused_kernel<C>;
    const size_t smem = sizeof(LeffFusedSmem<C>);
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    const unsigned grid = static_cast<unsigned>(a.B) * (a.H / 8) * (a.W / 8);
    k<<<grid, LF_THREADS, smem, stream>>>(a);
    return cudaGetLastError();
}

inline bool leff_fused_supported(int C, int hidden) { return hidden == 4 * C && (C == 32 || C == 64 || C == 128); }

inline cudaError_t launch_leff_fused(int C, const LeffFusedArgs& a, cudaStream_t stream) {
    switch (C) {
        case 32: return launch_leff_fused_c<32>(a, stream);
        case 64: return launch_leff_fused_c<64>(a, stream);
        case 128: return launch_leff_fused_c<128>(a, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace lewin
