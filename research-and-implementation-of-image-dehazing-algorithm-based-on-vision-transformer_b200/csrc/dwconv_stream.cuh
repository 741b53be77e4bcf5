// Persistent, double-buffered depthwise 3x3 + bias + exact GELU for bf16 channel-last maps (My_model_1.py:489-491, :517).
//
// The element-wise work, not HBM, bounded the earlier kernels (dwconv.cuh): ~0.66 SM-cycles per hidden element.  This
// variant is built around the instruction count per element:
//   * persistent CTAs (2 per SM) walk 16x16-pixel x 64-channel tiles, slab-major so the 36 weights + 4 biases a thread
//     needs stay in registers until the channel slab changes; the 16 KB GELU table is staged once per CTA;
//   * the (16+2)x(16+2) halo tile of the NEXT tile is fetched with 16-byte cp.async (zero-filled outside the map ==
//     the conv's zero padding) while the current one is computed (two shared-memory buffers);
//   * thread = (4-channel group, pixel column): it walks the 16 rows with a rolling 3x3 register window, so every input
//     value is loaded from shared memory and unpacked once per thread-row (3 LDS.64 + 12 unpacks per 4 outputs) and the
//     36 MACs are 18 packed FFMA2;
//   * GELU is the branch-free pair lookup (common.cuh) with one deferred range test per 4 outputs;
//   * a warp stores 2 pixels x 128 contiguous bytes per instruction.
// Algorithmic HBM traffic: read Ch + write Ch per pixel (the halo re-reads hit L2).
#pragma once
#include <type_traits>
#include "common.cuh"
#include "tma.cuh"

namespace lewin {
namespace dws {

constexpr int TY = 16, TX = 16, SLAB = 64;
constexpr int HY = TY + 2, HX = TX + 2;
constexpr int TILE_BYTES = HY * HX * SLAB * 2;            // 41472
constexpr int THREADS = 256;
constexpr size_t SMEM = 128 /*align*/ + 2 * TILE_BYTES + kGelu2TabSize * 2 + 64 /*mbarriers*/;

__device__ __forceinline__ void ffma2(float2& d, const float2 a, const float2 b) {
    unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d);
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd)
        : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
    d = *reinterpret_cast<float2*>(&dd);
}
__device__ __forceinline__ void cp16z(uint32_t smem_addr, const void* gmem, bool valid) {
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_addr), "l"(gmem), "r"(sz));
}

struct Args {
    const __nv_bfloat16* x;      // [B, H, W, Ch]
    __nv_bfloat16* out;          // [B, H, W, Ch]
    __nv_bfloat16* preact;       // optional pre-GELU copy (training), may be null
    const __nv_bfloat16* aux;    // BWD mode: a1 (pre-GELU linear1 output), da1 = conv^T(da2) * gelu'(a1)
    const float* w;              // [Ch, 9]
    const float* bias;           // [Ch]
    int B, H, W, Ch;
    int tiles_x, tiles_y, spatial_tiles, total_tiles;
};

// USE_TMA: the halo tile is ONE cp.async.bulk.tensor box per tile (issued by thread 0, zero-filled outside the map by
// the TMA unit, completion on an mbarrier) instead of 2592 per-thread 16-byte cp.async with their index arithmetic.
// BWD: the same pipeline computes the data gradient of the depthwise conv: x = da2, taps flipped, no bias, and the
// epilogue multiplies by gelu'(a1) instead of applying GELU (da1 = conv^T(da2) * gelu'(a1), autograd of :508-517).
template <bool USE_TMA, bool BWD = false>
__global__ void __launch_bounds__(THREADS, 2) dwconv_stream_kernel(const Args a, const __grid_constant__ CUtensorMap xmap) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((128u - (tma::smem_u32(smem_raw) & 127u)) & 127u);
    uint16_t* gtab = reinterpret_cast<uint16_t*>(smem + 2 * TILE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * TILE_BYTES + kGelu2TabSize * 2);
    const uint32_t smem_u = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
    const int tid = threadIdx.x;
    const int cq = tid & 15, px = tid >> 4;                // 4-channel group, pixel column

    auto decode = [&](int t, int& slab, int& b, int& ty, int& tx) {
        slab = t / a.spatial_tiles;
        int r = t - slab * a.spatial_tiles;
        tx = r % a.tiles_x; r /= a.tiles_x;
        ty = r % a.tiles_y;
        b = r / a.tiles_y;
    };
    auto fetch = [&](int t, int bufi) {
        int slab, b, ty, tx;
        decode(t, slab, b, ty, tx);
        if constexpr (USE_TMA) {
            if (tid == 0) {
                tma::mbar_expect_tx(&bars[bufi], TILE_BYTES);
                tma::load_4d(smem + bufi * TILE_BYTES, &xmap, &bars[bufi], slab * SLAB, tx * TX - 1, ty * TY - 1, b);
            }
            return;
        }
        const int y0 = ty * TY - 1, x0 = tx * TX - 1;
        const __nv_bfloat16* src0 = a.x + static_cast<long long>(b) * a.H * a.W * a.Ch + slab * SLAB;
        const uint32_t dst0 = smem_u + bufi * TILE_BYTES;
        for (int i = tid; i < HY * HX * 8; i += THREADS) {
            const int ch = i & 7, p = i >> 3;
            const int hy = p / HX, hx = p - hy * HX;
            const int yy = y0 + hy, xx = x0 + hx;
            const bool ok = yy >= 0 && yy < a.H && xx >= 0 && xx < a.W;
            cp16z(dst0 + i * 16, src0 + (static_cast<long long>(ok ? yy : 0) * a.W + (ok ? xx : 0)) * a.Ch + ch * 8, ok);
        }
    };

    if constexpr (USE_TMA) {
        if (tid == 0) {
            tma::prefetch_map(&xmap);
            tma::mbar_init(&bars[0], 1);
            tma::mbar_init(&bars[1], 1);
            tma::fence_barrier_init();
        }
        __syncthreads();
    }
    int t = blockIdx.x;
    if (t < a.total_tiles) fetch(t, 0);
    cp_async_commit();
    if constexpr (BWD) gelu_grad_tab2_to_smem(gtab, tid, THREADS); else gelu_tab2_to_smem(gtab, tid, THREADS);
    uint32_t bphase[2] = {0u, 0u};

    float2 wk[9][2];                                       // taps x channel pairs (autocast: bf16-rounded weights)
    float2 bz[2];
    int cur_slab = -1;
    bool cur_slab_first = true;
    int bufi = 0;
    for (; t < a.total_tiles; t += gridDim.x, bufi ^= 1) {
        const int tn = t + gridDim.x;
        if (tn < a.total_tiles) fetch(tn, bufi ^ 1);
        cp_async_commit();
        int slab, b, ty, tx;
        decode(t, slab, b, ty, tx);
        if (slab != cur_slab) {
            cur_slab = slab;
            const int c = slab * SLAB + cq * 4;
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                wk[k][0] = make_float2(Act<__nv_bfloat16>::round(__ldg(a.w + (c + 0) * 9 + k)), Act<__nv_bfloat16>::round(__ldg(a.w + (c + 1) * 9 + k)));
                wk[k][1] = make_float2(Act<__nv_bfloat16>::round(__ldg(a.w + (c + 2) * 9 + k)), Act<__nv_bfloat16>::round(__ldg(a.w + (c + 3) * 9 + k)));
            }
            if constexpr (BWD) {
                bz[0] = make_float2(0.f, 0.f); bz[1] = bz[0];
            } else {
                const float4 b4 = *reinterpret_cast<const float4*>(a.bias + c);
                bz[0] = make_float2(Act<__nv_bfloat16>::round(b4.x), Act<__nv_bfloat16>::round(b4.y));
                bz[1] = make_float2(Act<__nv_bfloat16>::round(b4.z), Act<__nv_bfloat16>::round(b4.w));
            }
        }
        if constexpr (USE_TMA) {
            tma::mbar_wait(&bars[bufi], bphase[bufi]);         // this tile's box has landed (async proxy -> visible after the wait)
            bphase[bufi] ^= 1u;
            if (cur_slab_first) { __syncthreads(); cur_slab_first = false; }   // GELU table staged by all threads
        } else {
            cp_async_wait<1>();                            // this tile's halo has landed (the next one may still fly)
            __syncthreads();
        }

        const unsigned char* tile = smem + bufi * TILE_BYTES + cq * 8;
        auto ldrow = [&](float2 (&dst)[3][2], int hy) {    // 3 columns x 4 channels of halo row hy
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const uint2 u = *reinterpret_cast<const uint2*>(tile + ((hy * HX + px + j) * SLAB) * 2);
                dst[j][0] = make_float2(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xFFFF0000u));
                dst[j][1] = make_float2(__uint_as_float(u.y << 16), __uint_as_float(u.y & 0xFFFF0000u));
            }
        };
        float2 win[3][3][2];                               // [row][col][channel pair]
        ldrow(win[0], 0);
        ldrow(win[1], 1);
        const long long obase = ((static_cast<long long>(b) * a.H + ty * TY) * a.W + tx * TX + px) * a.Ch + slab * SLAB + cq * 4;
        const long long rstride = static_cast<long long>(a.W) * a.Ch;
        __nv_bfloat16* op = a.out + obase;
        __nv_bfloat16* pp = a.preact ? a.preact + obase : nullptr;
        const __nv_bfloat16* ap = BWD ? a.aux + obase : nullptr;
#pragma unroll
        for (int y = 0; y < TY; ++y) {
            ldrow(win[(y + 2) % 3], y + 2);
            float2 acc0 = bz[0], acc1 = bz[1];
            if constexpr (BWD) {
                // dh1[y,x] = sum_{ky,kx} da2[y - ky + 1, x - kx + 1] * w[ky,kx]  (same summation order as dwconv_bwd_data_kernel)
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        ffma2(acc0, win[(y + 2 - ky) % 3][2 - kx][0], wk[ky * 3 + kx][0]);
                        ffma2(acc1, win[(y + 2 - ky) % 3][2 - kx][1], wk[ky * 3 + kx][1]);
                    }
                const uint2 pa = *reinterpret_cast<const uint2*>(ap);
                ap += rstride;
                uint32_t oor = 0;
                uint32_t g0 = gelu_pair_fast(gtab, pa.x, oor), g1 = gelu_pair_fast(gtab, pa.y, oor);     // gtab holds gelu' here
                if (__builtin_expect(gelu_pair_oor(oor), 0)) { g0 = gelu_grad_pair_exact(gtab, pa.x); g1 = gelu_grad_pair_exact(gtab, pa.y); }
                __nv_bfloat162 h0 = __floats2bfloat162_rn(acc0.x * __uint_as_float(g0 << 16), acc0.y * __uint_as_float(g0 & 0xFFFF0000u));
                __nv_bfloat162 h1 = __floats2bfloat162_rn(acc1.x * __uint_as_float(g1 << 16), acc1.y * __uint_as_float(g1 & 0xFFFF0000u));
                *reinterpret_cast<uint2*>(op) = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
                op += rstride;
            } else {
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        ffma2(acc0, win[(y + ky) % 3][kx][0], wk[ky * 3 + kx][0]);
                        ffma2(acc1, win[(y + ky) % 3][kx][1], wk[ky * 3 + kx][1]);
                    }
                __nv_bfloat162 h0 = __floats2bfloat162_rn(acc0.x, acc0.y), h1 = __floats2bfloat162_rn(acc1.x, acc1.y);
                const uint32_t in0 = *reinterpret_cast<uint32_t*>(&h0), in1 = *reinterpret_cast<uint32_t*>(&h1);
                uint32_t oor = 0;
                uint32_t q0 = gelu_pair_fast(gtab, in0, oor), q1 = gelu_pair_fast(gtab, in1, oor);
                if (__builtin_expect(gelu_pair_oor(oor), 0)) { q0 = gelu_pair_exact(gtab, in0); q1 = gelu_pair_exact(gtab, in1); }
                if (pp) { *reinterpret_cast<uint2*>(pp) = make_uint2(in0, in1); pp += rstride; }
                *reinterpret_cast<uint2*>(op) = make_uint2(q0, q1);
                op += rstride;
            }
        }
        __syncthreads();                                   // every thread is done with this buffer before it is refilled
    }
}


// ---------------------------------------------------------------- tensor-core variant (opt-in: LEWIN_MMA_DWCONV=1)
// MEASURED NEGATIVE RESULT (B200, 16 x 128^2 x 256 channels): 0.43 warp-instructions per output instead of 0.66, same
// results (1 bf16 ulp on < 0.01 % of the elements: tensor-core accumulation rounding), but 1.5x SLOWER than the FFMA2
// kernel above (ncu: LSU data pipe 71 %, issue 35 %).  The fragment layout leaves a lane with 4 channels of ONE pixel
// per 16-byte segment, so a warp's STG.64 touches 16 lines (16 LSU wavefronts instead of 2), and the GELU table lookups
// (4.2 wavefronts per LDS.U16 from bank conflicts) are unchanged.  Both depthwise kernels are bound by the LSU pipe of
// the GELU lookups, not by the 9-tap MACs; kept for the record and for a TMA-store epilogue experiment.
// The streaming kernel above spends 21 thread-instructions per hidden element: 4.5 FFMA2, 3 bf16 -> fp32 unpacks (every
// input is unpacked by the three threads whose 3-column windows cover it) and the GELU lookup.  Here the 9 taps run on
// the tensor pipe instead: for an 8-channel block, out[16 px, 8 ch] += In_tap[16 px, 8 ch] x diag(w_tap[8 ch]) is one
// mma.m16n8k8 (bf16 operands, fp32 accumulate; 1/8 of the MACs are useful, and the pipe is otherwise idle).  The A
// fragment comes straight from the bf16 halo tile by ldmatrix (no unpack), one input row's three column-shifted
// fragments feed the three output rows that use it (rolling accumulators), the bias is the accumulator's initial value,
// and GELU acts on the accumulator fragments.  Per warp and output row (128 outputs): 3 LDSM + 9 HMMA + the GELU pair
// lookups, ~0.38 warp-instructions per output instead of 0.66.
//   * warp = one 8-channel block of the 64-channel slab, all 16 rows of the tile; B fragments (9 registers) and the
//     bias stay in registers while the slab is unchanged;
//   * the halo tile is one SWIZZLE_128B TMA box (128 bytes per pixel, 16-byte chunk index XOR pixel index mod 8), so
//     the 8 row addresses of an ldmatrix (8 consecutive pixels, same channel block) fall into 8 different bank groups;
//     18 * row + column shifts the swizzle phase by a compile-time constant, 8 pre-swizzled lane offsets cover all;
//   * epilogue: lanes of a pair exchange one word so that each lane stores 8 contiguous bytes (4 channels of one pixel).
constexpr int TILE_PAD = 41 * 1024;                      // TILE_BYTES rounded up to the 1024-byte swizzle atom
constexpr size_t SMEM_MMA = 1024 /*align*/ + 2 * TILE_PAD + kGelu2TabSize * 2 + 64 /*mbarriers*/;

__device__ __forceinline__ void ldsm_x2(uint32_t& r0, uint32_t& r1, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void mma_k8(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5}, {%6}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(b0));
}

constexpr int THREADS_MMA = 512;                         // 16 warps: (8-channel block, upper / lower 8 rows of the tile)
constexpr int RY = TY / 2;                               // output rows per warp

// first tap of an output row: the accumulator starts from the bias (c0, c1 repeated for the two pixel rows of the fragment)
__device__ __forceinline__ void mma_k8_init(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0, float c0, float c1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5}, {%6}, {%7, %8, %7, %8};"
        : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3]) : "r"(a0), "r"(a1), "r"(b0), "f"(c0), "f"(c1));
}

template <bool BWD>
__global__ void __launch_bounds__(THREADS_MMA, 2) dwconv_mma_kernel(const Args a, const __grid_constant__ CUtensorMap xmap) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (tma::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint16_t* gtab = reinterpret_cast<uint16_t*>(smem + 2 * TILE_PAD);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * TILE_PAD + kGelu2TabSize * 2);
    const uint32_t smem_u = tma::smem_u32(smem);
    const int tid = threadIdx.x, lane = tid & 31;
    const int cb = (tid >> 5) & 7, half = tid >> 8;        // 8-channel block of the slab; rows [8 half, 8 half + 8)
    const int g = lane >> 2, tq = lane & 3;
    const bool odd = (tq & 1) != 0;

    auto decode = [&](int t, int& slab, int& b, int& ty, int& tx) {
        slab = t / a.spatial_tiles;
        int r = t - slab * a.spatial_tiles;
        tx = r % a.tiles_x; r /= a.tiles_x;
        ty = r % a.tiles_y;
        b = r / a.tiles_y;
    };
    auto fetch = [&](int t, int bufi) {
        if (tid == 0) {
            int slab, b, ty, tx;
            decode(t, slab, b, ty, tx);
            tma::mbar_expect_tx(&bars[bufi], TILE_BYTES);
            tma::load_4d(smem + bufi * TILE_PAD, &xmap, &bars[bufi], slab * SLAB, tx * TX - 1, ty * TY - 1, b);
        }
    };
    if (tid == 0) {
        tma::prefetch_map(&xmap);
        tma::mbar_init(&bars[0], 1);
        tma::mbar_init(&bars[1], 1);
        tma::fence_barrier_init();
    }
    __syncthreads();
    int t = blockIdx.x;
    if (t < a.total_tiles) fetch(t, 0);
    if constexpr (BWD) gelu_grad_tab2_to_smem(gtab, tid, THREADS_MMA); else gelu_tab2_to_smem(gtab, tid, THREADS_MMA);
    __syncthreads();
    uint32_t bphase = 0u;                                  // bit i: parity of buffer i

    // ldmatrix row of this lane: pixel column lane & 15 (lanes 16..31 repeat valid addresses, ignored by .x2).
    // Halo pixel index = 18 * row + column; its swizzle phase (index & 7) = (lp + 2 * row + kx) & 7.
    const int lp = lane & 15;
    uint32_t swz[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) swz[j] = (half * RY * HX + lp) * 128 + ((((lp + j) & 7) ^ cb) << 4);

    uint32_t bw[9];                                        // diag(w_tap) B fragments: k = 2 tq + {0, 1}, n = g
    float bz0 = 0.f, bz1 = 0.f;                            // bias of channels 2 tq, 2 tq + 1 of the block
    int cur_slab = -1;
    int bufi = 0;
    for (; t < a.total_tiles; t += gridDim.x, bufi ^= 1) {
        const int tn = t + gridDim.x;
        if (tn < a.total_tiles) fetch(tn, bufi ^ 1);
        int slab, b, ty, tx;
        decode(t, slab, b, ty, tx);
        if (slab != cur_slab) {
            cur_slab = slab;
            const int c0 = slab * SLAB + cb * 8;
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const int ks = BWD ? 8 - k : k;            // data gradient: taps flipped in both directions
                const uint32_t wb = __bfloat16_as_ushort(__float2bfloat16_rn(__ldg(a.w + (c0 + g) * 9 + ks)));
                bw[k] = ((g >> 1) == tq) ? ((g & 1) ? (wb << 16) : wb) : 0u;
            }
            if constexpr (!BWD) {
                bz0 = Act<__nv_bfloat16>::round(__ldg(a.bias + c0 + 2 * tq));
                bz1 = Act<__nv_bfloat16>::round(__ldg(a.bias + c0 + 2 * tq + 1));
            }
        }
        tma::mbar_wait(&bars[bufi], (bphase >> bufi) & 1u); // this tile's box has landed
        bphase ^= 1u << bufi;

        const uint32_t tb = smem_u + bufi * TILE_PAD;
        // this lane's 8 output bytes: pixel g (even lane of a pair) or g + 8 (odd lane), channels 4 * (tq >> 1) .. + 3
        const long long obase = ((static_cast<long long>(b) * a.H + ty * TY + half * RY) * a.W + tx * TX + g + (odd ? 8 : 0)) * a.Ch +
                                slab * SLAB + cb * 8 + (tq >> 1) * 4;
        const long long rstride = static_cast<long long>(a.W) * a.Ch;

        // One pass over the warp's 8 output rows (10 halo rows), fully unrolled, no control flow inside so that the next
        // row's LDSM / HMMA overlap this row's GELU lookups.  Elements outside the GELU table (|x| < 2^-28 or >= 16,
        // practically never) are detected once per pass and the pass is repeated with the exact lookup.
        auto pass = [&](auto exact_tag, auto pre_tag) -> uint32_t {
            constexpr bool EXACT = decltype(exact_tag)::value, PRE = decltype(pre_tag)::value;
            uint32_t oor = 0;
            __nv_bfloat16* op = a.out + obase;
            // second stream at the same element offset: pre-GELU copy (forward, training) or a1 (data gradient)
            const long long delta = BWD ? (a.aux - a.out) : (PRE ? (a.preact - a.out) : 0);
            float acc[3][4];
#pragma unroll
            for (int r = 0; r < RY + 2; ++r) {
                uint32_t A[3][2];
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) ldsm_x2(A[kx][0], A[kx][1], tb + swz[(2 * r + kx) & 7] + (r * HX + kx) * 128);
#pragma unroll
                for (int ky = 2; ky >= 0; --ky) {          // ky = 2 first: it completes output row r - 2
                    const int y = r - ky;
                    if (y < 0 || y >= RY) continue;
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        if (ky == 0 && kx == 0) mma_k8_init(acc[y % 3], A[kx][0], A[kx][1], bw[0], bz0, bz1);   // row y starts here
                        else mma_k8(acc[y % 3], A[kx][0], A[kx][1], bw[ky * 3 + kx]);
                    }
                }
                if (r < 2) continue;
                float (&c)[4] = acc[(r - 2) % 3];
                if constexpr (BWD) {
                    const float s0 = odd ? c[0] : c[2], s1 = odd ? c[1] : c[3];
                    const float r0 = __shfl_xor_sync(0xFFFFFFFFu, s0, 1), r1 = __shfl_xor_sync(0xFFFFFFFFu, s1, 1);
                    const float v0 = odd ? r0 : c[0], v1 = odd ? r1 : c[1], v2 = odd ? c[2] : r0, v3 = odd ? c[3] : r1;
                    const uint2 pa = *reinterpret_cast<const uint2*>(op + delta);
                    uint32_t g0, g1;                       // gtab holds gelu' here
                    if constexpr (EXACT) { g0 = gelu_grad_pair_exact(gtab, pa.x); g1 = gelu_grad_pair_exact(gtab, pa.y); }
                    else { g0 = gelu_pair_fast(gtab, pa.x, oor); g1 = gelu_pair_fast(gtab, pa.y, oor); }
                    __nv_bfloat162 h0 = __floats2bfloat162_rn(v0 * __uint_as_float(g0 << 16), v1 * __uint_as_float(g0 & 0xFFFF0000u));
                    __nv_bfloat162 h1 = __floats2bfloat162_rn(v2 * __uint_as_float(g1 << 16), v3 * __uint_as_float(g1 & 0xFFFF0000u));
                    *reinterpret_cast<uint2*>(op) = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
                } else {
                    __nv_bfloat162 h0 = __floats2bfloat162_rn(c[0], c[1]), h1 = __floats2bfloat162_rn(c[2], c[3]);
                    const uint32_t in0 = *reinterpret_cast<uint32_t*>(&h0), in1 = *reinterpret_cast<uint32_t*>(&h1);   // pixel g, pixel g + 8
                    uint32_t q0, q1;
                    if constexpr (EXACT) { q0 = gelu_pair_exact(gtab, in0); q1 = gelu_pair_exact(gtab, in1); }
                    else { q0 = gelu_pair_fast(gtab, in0, oor); q1 = gelu_pair_fast(gtab, in1, oor); }
                    const uint32_t rq = __shfl_xor_sync(0xFFFFFFFFu, odd ? q0 : q1, 1);
                    if constexpr (PRE) {
                        const uint32_t ri = __shfl_xor_sync(0xFFFFFFFFu, odd ? in0 : in1, 1);
                        *reinterpret_cast<uint2*>(op + delta) = odd ? make_uint2(ri, in1) : make_uint2(in0, ri);
                    }
                    *reinterpret_cast<uint2*>(op) = odd ? make_uint2(rq, q1) : make_uint2(q0, rq);
                }
                op += rstride;
            }
            return oor;
        };
        if (!BWD && a.preact) {
            const uint32_t oor = pass(std::false_type{}, std::true_type{});
            if (__builtin_expect(__any_sync(0xFFFFFFFFu, gelu_pair_oor(oor)), 0)) pass(std::true_type{}, std::true_type{});
        } else {
            const uint32_t oor = pass(std::false_type{}, std::false_type{});
            if (__builtin_expect(__any_sync(0xFFFFFFFFu, gelu_pair_oor(oor)), 0)) pass(std::true_type{}, std::false_type{});
        }
        __syncthreads();                                   // every warp is done with this buffer before it is refilled
    }
}

inline bool supported(int H, int W, int Ch) {
    static const bool on = [] { const char* e = getenv("LEWIN_NO_STREAM_DWCONV"); return !(e && e[0] == '1'); }();
    return on && H % TY == 0 && W % TX == 0 && Ch % SLAB == 0;
}

template <bool BWD>
inline cudaError_t launch_mode(const __nv_bfloat16* x, __nv_bfloat16* out, __nv_bfloat16* preact, const __nv_bfloat16* aux, const float* w,
                               const float* bias, int B, int H, int W, int Ch, int num_sms, cudaStream_t stream) {
    Args a{};
    a.x = x; a.out = out; a.preact = preact; a.aux = aux; a.w = w; a.bias = bias;
    a.B = B; a.H = H; a.W = W; a.Ch = Ch;
    a.tiles_x = W / TX; a.tiles_y = H / TY;
    a.spatial_tiles = B * a.tiles_x * a.tiles_y;
    a.total_tiles = a.spatial_tiles * (Ch / SLAB);
    int grid = 2 * num_sms;
    if (grid > a.total_tiles) grid = a.total_tiles;
    static const bool tma_on = [] { const char* e = getenv("LEWIN_NO_TMA"); return !(e && e[0] == '1'); }();
    static const bool mma_on = [] { const char* e = getenv("LEWIN_MMA_DWCONV"); return e && e[0] == '1'; }();   // opt-in (measured slower)
    CUtensorMap map{};
    if (tma_on && mma_on && tma::make_nhwc_bf16(&map, x, B, H, W, Ch, HY, HX, SLAB, true)) {
        cudaError_t e = cudaFuncSetAttribute(dwconv_mma_kernel<BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(SMEM_MMA));
        if (e != cudaSuccess) return e;
        dwconv_mma_kernel<BWD><<<grid, THREADS_MMA, SMEM_MMA, stream>>>(a, map);
        return cudaGetLastError();
    }
    if (tma_on && tma::make_nhwc_bf16(&map, x, B, H, W, Ch, HY, HX, SLAB)) {
        cudaError_t e = cudaFuncSetAttribute(dwconv_stream_kernel<true, BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(SMEM));
        if (e != cudaSuccess) return e;
        dwconv_stream_kernel<true, BWD><<<grid, THREADS, SMEM, stream>>>(a, map);
    } else {
        cudaError_t e = cudaFuncSetAttribute(dwconv_stream_kernel<false, BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(SMEM));
        if (e != cudaSuccess) return e;
        dwconv_stream_kernel<false, BWD><<<grid, THREADS, SMEM, stream>>>(a, map);
    }
    return cudaGetLastError();
}

inline cudaError_t launch(const __nv_bfloat16* x, __nv_bfloat16* out, __nv_bfloat16* preact, const float* w, const float* bias,
                          int B, int H, int W, int Ch, int num_sms, cudaStream_t stream) {
    return launch_mode<false>(x, out, preact, nullptr, w, bias, B, H, W, Ch, num_sms, stream);
}

// ---- backward, data path:  da2 = g2 * gelu'(a2) (element-wise, also the weight-gradient kernel's input), then
//      da1 = conv^T(da2) * gelu'(a1) on the streaming kernel above
__global__ void __launch_bounds__(256) dgelu_mul_kernel(const __nv_bfloat16* __restrict__ g2, const __nv_bfloat16* __restrict__ a2,
                                                        __nv_bfloat16* __restrict__ da2, long long n8) {
    __shared__ __align__(16) uint16_t gtab[kGelu2TabSize];
    gelu_grad_tab2_to_smem(gtab, threadIdx.x, 256);
    __syncthreads();
    for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < n8; i += static_cast<long long>(gridDim.x) * 256) {
        const uint4 g = reinterpret_cast<const uint4*>(g2)[i];
        const uint4 p = reinterpret_cast<const uint4*>(a2)[i];
        const uint32_t gw[4] = {g.x, g.y, g.z, g.w}, pw[4] = {p.x, p.y, p.z, p.w};
        uint32_t o[4], oor = 0, d[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) d[j] = gelu_pair_fast(gtab, pw[j], oor);
        if (__builtin_expect(gelu_pair_oor(oor), 0)) {
#pragma unroll
            for (int j = 0; j < 4; ++j) d[j] = gelu_grad_pair_exact(gtab, pw[j]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(gw[j] << 16) * __uint_as_float(d[j] << 16),
                                                     __uint_as_float(gw[j] & 0xFFFF0000u) * __uint_as_float(d[j] & 0xFFFF0000u));
            o[j] = *reinterpret_cast<uint32_t*>(&h);
        }
        reinterpret_cast<uint4*>(da2)[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

inline bool bwd_enabled() {
    static const bool on = [] { const char* e = getenv("LEWIN_NO_BWD2"); return !(e && e[0] == '1'); }();
    return on;
}
inline cudaError_t launch_bwd_data(const __nv_bfloat16* g2, const __nv_bfloat16* a2, const __nv_bfloat16* a1, __nv_bfloat16* da1,
                                   __nv_bfloat16* da2, const float* w, int B, int H, int W, int Ch, int num_sms, cudaStream_t stream) {
    const long long n8 = static_cast<long long>(B) * H * W * Ch / 8;
    long long grid = static_cast<long long>(num_sms) * 8;
    if (grid > (n8 + 255) / 256) grid = (n8 + 255) / 256;
    dgelu_mul_kernel<<<static_cast<unsigned>(grid), 256, 0, stream>>>(g2, a2, da2, n8);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return launch_mode<true>(da2, da1, nullptr, a1, w, nullptr, B, H, W, Ch, num_sms, stream);
}

}  // namespace dws
}  // namespace lewin
