// Implicit-GEMM convolutions of the U-Net's projection layers on tcgen05 (bf16 inference, channel-last token maps):
//
//   Downsample.forward  My_model_1.py:606-630   Conv2d(C, 2C, kernel 4, stride 2, padding 1)  -> tokens [B, H/2 * W/2, 2C]
//   OutputProj.forward  My_model_1.py:696-733   Conv2d(2C, 3, kernel 3, stride 1, padding 1)  -> image  [B, 3, H, W] (+ x)
//
// (round 1 ran both on stock cuDNN plus separate bias-add / residual passes.)
//
// out[pixel, n] = sum over taps t, channels c of  x[pixel * stride + offset(t), c] * w[n, c, t]:  one tcgen05.mma per
// (tap, 64-channel chunk) accumulating into the tile's TMEM accumulator.  The A operand of a tap is ONE TMA box: the 128
// output pixels of a tile are a PY x PX patch, and the input pixels a tap reads for them form a PY x PX box of the 5-D view
//
//   x[b, 2 i' + q, 2 j' + p, c]  ->  dims (c' = p * ld + c,  j',  q,  i',  b)          (stride 2: pixel-pair / row-pair view)
//   x[b, i, j, c]                ->  dims (c, j, 1, i, b)                               (stride 1)
//
// so the stride-2 gather needs no traversal strides, out-of-range coordinates are zero-filled by the TMA unit (== the
// convolution's zero padding, per image), and the box lands in the K-major SWIZZLE_128B / _64B layout the MMA consumes.
// The input may have a row stride larger than its channel count (the right half of a torch.cat([up, skip]) buffer).
// Weights: bf16 image [tap][N][C] made per call by conv_prep_kernel.  Warp roles as ws::gemm_wss_kernel: one TMA thread,
// one MMA thread, 8 epilogue warps (tcgen05.ld -> bf16 -> + bias -> bf16, the two roundings of torch's conv + bias add under
// autocast -> coalesced token rows).  The 3-channel OutputProj has its own kernel below (op::outproj_mma_kernel).
#pragma once
#include "tc_helpers.cuh"
#include "tma.cuh"

namespace lewin {
namespace cv {

constexpr int NEW = 8;                                  // epilogue warps (4 TMEM lane groups x 2 column groups)
constexpr int MMA_WARP = NEW, TMA_WARP = NEW + 1;
constexpr int THREADS = (NEW + 2) * 32;
constexpr int STG_ROW = 80, STG_BUF = 32 * STG_ROW;
constexpr int SMEM_MAX = 227 * 1024;

struct Tap { int dc, dj, dq, di; };

struct Args {
    int B, Hout, Wout;
    int N;                       // GEMM columns (Cout)
    int n_real;                  // == N
    int px_shift;                // tile = (128 >> px_shift) rows x (1 << px_shift) columns of output pixels
    int tiles_x, tiles_y, col_tiles, tiles;
    int ntaps, nkc;
    Tap taps[16];
    const float* bias;           // [n_real]
    int relu;                    // 1: max(0, .) after the bias (conv + ReLU pairs of the VGG19 feature extractor, My_CR.py:56-84)
    __nv_bfloat16* out_tok;      // [B, Hout, Wout, ld_out]
    long long ld_out;
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void load_5d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(tc::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(tc::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
// epilogue: bf16 token rows + bias
template <int KCH, int BN>
__global__ void __launch_bounds__(THREADS, 1) conv_igemm_kernel(const Args a, const __grid_constant__ CUtensorMap amap,
                                                                const __grid_constant__ CUtensorMap wmap, int S) {
    constexpr int A_CHUNK = 128 * KCH * 2, W_CHUNK = BN * KCH * 2, STAGE = A_CHUNK + ((W_CHUNK + 1023) / 1024) * 1024;
    constexpr int ACC = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
    constexpr int TMEM_COLS = 2 * ACC;
    constexpr int NCG = NEW / 4, NCH = BN / 32;
    constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(BN >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* base = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* ring = base;                                      // [S][A_CHUNK | W_CHUNK]
    unsigned char* stg = ring + static_cast<size_t>(S) * STAGE;      // [NEW][STG_BUF]
    float* s_bias = reinterpret_cast<float*>(stg + NEW * STG_BUF);   // [col_tiles * BN]
    uint64_t* full = reinterpret_cast<uint64_t*>(s_bias + a.col_tiles * BN);
    uint64_t* empty = full + 8;
    uint64_t* tfull = empty + 8;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int my_tiles = (a.tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    const int ksteps = a.ntaps * a.nkc;

    for (int i = tid; i < a.col_tiles * BN; i += THREADS) s_bias[i] = (a.bias && i < a.n_real) ? Act<__nv_bfloat16>::round(a.bias[i]) : 0.f;
    if (tid == 0) {
        for (int i = 0; i < S; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { tc::mbar_init(&tfull[i], 1); tc::mbar_init(&tempty[i], NEW * 32); }
        tc::fence_barrier_init();
        tma::prefetch_map(&amap);
        tma::prefetch_map(&wmap);
    }
    if (warp == MMA_WARP) tc::tmem_alloc<TMEM_COLS>(tmem_slot);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;

    // tile t -> (column tile, b, ty, tx); column tiles fastest so the CTAs sharing a pixel patch read it from L2 together
    auto decode = [&](int t, int& ct, int& b, int& ty, int& tx) {
        ct = t % a.col_tiles;
        int r = t / a.col_tiles;
        tx = r % a.tiles_x; r /= a.tiles_x;
        ty = r % a.tiles_y;
        b = r / a.tiles_y;
    };
    const int PX = 1 << a.px_shift, PY = 128 >> a.px_shift;

    if (warp == TMA_WARP) {
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int it = 0; it < my_tiles; ++it) {
                int ct, b, ty, tx;
                decode(static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x), ct, b, ty, tx);
                for (int t = 0; t < a.ntaps; ++t) {
                    const Tap tp = a.taps[t];
                    for (int kc = 0; kc < a.nkc; ++kc) {
                        tc::mbar_wait(&empty[s], ph ^ 1u);
                        tma::mbar_expect_tx(&full[s], A_CHUNK + W_CHUNK);
                        load_5d(ring + s * STAGE, &amap, &full[s], tp.dc + kc * KCH, tx * PX + tp.dj, tp.dq, ty * PY + tp.di, b);
                        tma::load_2d(ring + s * STAGE + A_CHUNK, &wmap, &full[s], kc * KCH, t * a.N + ct * BN);
                        if (++s == S) { s = 0; ph ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == MMA_WARP) {
        if (lane == 0) {
            const uint32_t ring_u = tc::smem_u32(ring);
            int s = 0;
            uint32_t ph = 0;
            for (int it = 0; it < my_tiles; ++it) {
                const int acc = it & 1;
                tc::mbar_wait(&tempty[acc], (static_cast<uint32_t>(it >> 1) & 1u) ^ 1u);
                tc::tc_fence_after();
                const uint32_t d_addr = tmem_d + static_cast<uint32_t>(acc * ACC);
                for (int k = 0; k < ksteps; ++k) {
                    tc::mbar_wait(&full[s], ph);
                    tc::tc_fence_after();
                    const uint64_t da = tc::make_desc<KCH>(ring_u + s * STAGE);
                    const uint64_t db = tc::make_desc<KCH>(ring_u + s * STAGE + A_CHUNK);
#pragma unroll
                    for (int k16 = 0; k16 < KCH / 16; ++k16) tc::mma_bf16(d_addr, da + 2 * k16, db + 2 * k16, IDESC, (k > 0 || k16 > 0) ? 1u : 0u);
                    tc::mma_commit(&empty[s]);
                    if (++s == S) { s = 0; ph ^= 1u; }
                }
                tc::mma_commit(&tfull[acc]);
            }
        }
    } else {
        const int lg = warp & 3, half = warp >> 2;
        unsigned char* my_stg = stg + warp * STG_BUF;
        const int r = lg * 32 + lane, py = r >> a.px_shift, px = r & (PX - 1);
        for (int it = 0; it < my_tiles; ++it) {
            int ct, b, ty, tx;
            decode(static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x), ct, b, ty, tx);
            const int acc = it & 1;
            const int gy = ty * PY + py, gx = tx * PX + px;
            const bool ok = gy < a.Hout && gx < a.Wout;
            tc::mbar_wait(&tfull[acc], static_cast<uint32_t>(it >> 1) & 1u);
            tc::tc_fence_after();
            const uint32_t t_addr = tmem_d + (static_cast<uint32_t>(lg * 32) << 16) + static_cast<uint32_t>(acc * ACC);
            {
                const long long oy = ok ? ((static_cast<long long>(b) * a.Hout + gy) * a.Wout + gx) * a.ld_out + ct * BN : -1;
                const float* bs = s_bias + ct * BN;
                for (int c = half; c < NCH; c += NCG) {
                    float v[32];
                    tc::tmem_ld32(t_addr + c * 32, v);
                    if (c + NCG >= NCH) {
                        tc::tc_fence_before();
                        mbar_arrive(&tempty[acc]);
                    }
                    unsigned char* srow = my_stg + lane * STG_ROW;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint32_t pk[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int i = j * 8 + 2 * e;
                            float o0 = Act<__nv_bfloat16>::round(v[i]) + bs[c * 32 + i], o1 = Act<__nv_bfloat16>::round(v[i + 1]) + bs[c * 32 + i + 1];
                            if (a.relu) { o0 = fmaxf(o0, 0.f); o1 = fmaxf(o1, 0.f); }
                            pk[e] = tc::pack_bf16(o0, o1);
                        }
                        *reinterpret_cast<uint4*>(srow + j * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    }
                    __syncwarp();
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {                 // 8 rows x 64 contiguous bytes per warp instruction
                        const int i = lane + 32 * jj, rl = i >> 2, cc = i & 3;
                        const uint4 val = *reinterpret_cast<const uint4*>(my_stg + rl * STG_ROW + cc * 16);
                        const long long o = __shfl_sync(0xffffffffu, oy, rl);
                        if (o >= 0) *reinterpret_cast<uint4*>(a.out_tok + o + c * 32 + cc * 8) = val;
                    }
                    __syncwarp();
                }
                if (half >= NCH) {                                   // this warp owns no chunk (BN == 32)
                    tc::tc_fence_before();
                    mbar_arrive(&tempty[acc]);
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) tc::tmem_dealloc<TMEM_COLS>(tmem_d);
}

// ---- host side
// 5-D bf16 view; box = (box_c, box_j, 1, box_i, 1); zero fill outside
inline bool make_5d(CUtensorMap* map, const void* base, const unsigned long long dims[5], const unsigned long long strides_bytes[4],
                    int box_c, int box_j, int box_i, bool sw128) {
    tma::EncodeTiledFn fn = tma::encode_fn();
    if (!fn) return false;
    const cuuint64_t d[5] = {dims[0], dims[1], dims[2], dims[3], dims[4]};
    const cuuint64_t st[4] = {strides_bytes[0], strides_bytes[1], strides_bytes[2], strides_bytes[3]};
    const cuuint32_t box[5] = {static_cast<cuuint32_t>(box_c), static_cast<cuuint32_t>(box_j), 1u, static_cast<cuuint32_t>(box_i), 1u};
    const cuuint32_t es[5] = {1u, 1u, 1u, 1u, 1u};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), d, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              sw128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// row-major bf16 matrix [rows, cols]; box = [box_rows x kch columns], swizzle by the row width (128 / 64 bytes)
inline bool make_w2d(CUtensorMap* map, const void* base, long long rows, long long cols, int box_rows, int kch) {
    tma::EncodeTiledFn fn = tma::encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(cols) * 2};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(kch), static_cast<cuuint32_t>(box_rows)};
    const cuuint32_t es[2] = {1u, 1u};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              kch == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// wb[(t * N + n) * C + c] = bf16(w[n][c][t])  (w: [n_real, C, taps]); rows n >= n_real are zero
__global__ void conv_prep_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wb, int n_real, int N, int C, int taps) {
    const long long total = static_cast<long long>(taps) * N * C;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % C);
        const long long r = i / C;
        const int n = static_cast<int>(r % N), t = static_cast<int>(r / N);
        wb[i] = __float2bfloat16_rn(n < n_real ? w[(static_cast<long long>(n) * C + c) * taps + t] : 0.f);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// OutputProj (Cout <= 8 channels): the implicit GEMM above reloads every input pixel once per tap (9 boxes per tile) and
// measured L2-bound (355 us per 1664^2 canvas).  Here the (16+2) x (16+2) x 64-channel halo tile is fetched ONCE by TMA
// (SWIZZLE_128B, zero fill == the conv padding), and the 9 shifted views are gathered from it by ldmatrix row addresses
// (one 128-byte row per pixel, conflict-free through the swizzle) into mma.sync.m16n8k16 A fragments; the n8 B fragments of
// the 3 x Cin x 9 weights live in shared memory in fragment order.  HBM traffic = the map once.
namespace op {
constexpr int TY = 16, TX = 16, HY = TY + 2, HX = TX + 2, SLAB = 64;
constexpr int TILE_BYTES = HY * HX * SLAB * 2;                    // 41472
constexpr int TILE_STRIDE = (TILE_BYTES + 1023) / 1024 * 1024;    // SWIZZLE_128B destinations are 1024-byte aligned
constexpr int THREADS = 256;

struct Args {
    int B, H, W, Cin, Cout, Hout, pad;
    int tiles_x, tiles_y, tiles;
    const float* weight;         // [Cout, Cin, 3, 3]
    const float* bias;
    const float* resid;          // [B, Cout, Hout, W] or null
    float* out;                  // [B, Cout, Hout, W]
};

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(THREADS, 2) outproj_mma_kernel(const Args a, const __grid_constant__ CUtensorMap xmap) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* base = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* tiles = base;                                              // [2][TILE_STRIDE]
    uint2* bfrag = reinterpret_cast<uint2*>(tiles + 2 * TILE_STRIDE);         // [nslab][9 taps][4 kk][32 lanes]
    const int nslab = a.Cin / SLAB;
    uint64_t* bars = reinterpret_cast<uint64_t*>(bfrag + nslab * 9 * 4 * 32);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;

    for (int i = tid; i < nslab * 9 * 4 * 32; i += THREADS) {                 // B fragments: b0 = B[k = 2tq, 2tq+1][n = gq], b1 = B[k + 8][n]
        const int l = i & 31, kk = (i >> 5) & 3, r = i >> 7, tap = r % 9, sl = r / 9;
        const int n = l >> 2, c0 = sl * SLAB + kk * 16 + 2 * (l & 3);
        float w[4] = {0.f, 0.f, 0.f, 0.f};
        if (n < a.Cout) {
            const float* wr = a.weight + static_cast<long long>(n) * a.Cin * 9 + tap;
            w[0] = wr[(c0) * 9]; w[1] = wr[(c0 + 1) * 9]; w[2] = wr[(c0 + 8) * 9]; w[3] = wr[(c0 + 9) * 9];
        }
        bfrag[i] = make_uint2(tc::pack_bf16(w[0], w[1]), tc::pack_bf16(w[2], w[3]));
    }
    if (tid == 0) {
        tma::prefetch_map(&xmap);
        tma::mbar_init(&bars[0], 1);
        tma::mbar_init(&bars[1], 1);
        tma::fence_barrier_init();
    }
    __syncthreads();

    const int units = a.tiles * nslab;
    auto fetch = [&](int u, int bufi) {
        if (tid != 0) return;
        const int t = u / nslab, sl = u - t * nslab;
        const int tx = t % a.tiles_x, r = t / a.tiles_x, ty = r % a.tiles_y, b = r / a.tiles_y;
        tma::mbar_expect_tx(&bars[bufi], TILE_BYTES);
        tma::load_4d(tiles + bufi * TILE_STRIDE, &xmap, &bars[bufi], sl * SLAB, tx * TX - 1, ty * TY - a.pad, b);
    };
    // my unit sequence: tiles blockIdx.x, + gridDim.x, ... each with nslab slabs
    const int my_tiles = (a.tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    const int my_units = my_tiles * nslab;
    auto unit_of = [&](int j) { return (static_cast<int>(blockIdx.x) + (j / nslab) * static_cast<int>(gridDim.x)) * nslab + (j % nslab); };
    (void)units;
    if (my_units > 0) fetch(unit_of(0), 0);
    uint32_t phase[2] = {0u, 0u};
    float acc[2][4];
    const float b0v = (2 * tq < a.Cout) ? Act<__nv_bfloat16>::round(a.bias[2 * tq]) : 0.f;
    const float b1v = (2 * tq + 1 < a.Cout) ? Act<__nv_bfloat16>::round(a.bias[2 * tq + 1]) : 0.f;
    for (int j = 0; j < my_units; ++j) {
        const int bufi = j & 1;
        if (j + 1 < my_units) fetch(unit_of(j + 1), bufi ^ 1);
        const int u = unit_of(j), t = u / nslab, sl = u - t * nslab;
        if (sl == 0) {
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[m][c] = 0.f;
        }
        tma::mbar_wait(&bars[bufi], phase[bufi]);
        phase[bufi] ^= 1u;
        const uint32_t tile_u = tc::smem_u32(tiles + bufi * TILE_STRIDE);
        const uint2* bf = bfrag + sl * (9 * 4 * 32) + lane;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                // my row address of this tap with the swizzle key folded in: chunk = 2 kk + hi, so (chunk ^ key) << 4 ==
                // (kk << 5) ^ ((hi ^ key) << 4), and with 128-byte aligned rows the whole address is one XOR per k-step
                uint32_t rowa[2];
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    const int pix = (2 * warp + m + ky) * HX + (lane & 15) + kx;
                    rowa[m] = tile_u + pix * 128 + ((((lane >> 4) ^ pix) & 7) << 4);
                }
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const uint2 b = bf[((ky * 3 + kx) * 4 + kk) * 32];
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
                        uint32_t af[4];
                        ldsm_x4(af, rowa[m] ^ (kk << 5));
                        mma16816(acc[m], af, b.x, b.y);
                    }
                }
            }
        if (sl == nslab - 1) {
            const int tx = t % a.tiles_x, r = t / a.tiles_x, ty = r % a.tiles_y, b = r / a.tiles_y;
            const long long plane = static_cast<long long>(a.Hout) * a.W;
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                const int gy = ty * TY + 2 * warp + m;
                if (gy >= a.Hout) continue;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int gx = tx * TX + gq + 8 * h;
                    if (gx >= a.W) continue;
                    const long long o = static_cast<long long>(b) * a.Cout * plane + static_cast<long long>(gy) * a.W + gx;
                    if (2 * tq < a.Cout) {
                        float y = Act<__nv_bfloat16>::round(Act<__nv_bfloat16>::round(acc[m][2 * h]) + b0v);       // conv -> bf16, + bias -> bf16
                        if (a.resid) y += a.resid[o + (2 * tq) * plane];
                        a.out[o + (2 * tq) * plane] = y;
                    }
                    if (2 * tq + 1 < a.Cout) {
                        float y = Act<__nv_bfloat16>::round(Act<__nv_bfloat16>::round(acc[m][2 * h + 1]) + b1v);
                        if (a.resid) y += a.resid[o + (2 * tq + 1) * plane];
                        a.out[o + (2 * tq + 1) * plane] = y;
                    }
                }
            }
        }
        __syncthreads();                                   // every warp is done with this buffer before it is refilled
    }
}

inline cudaError_t launch(Args a, const void* x, int ldx, int num_sms, cudaStream_t stream) {
    a.tiles_x = (a.W + TX - 1) / TX;
    a.tiles_y = (a.Hout + TY - 1) / TY;
    a.tiles = a.B * a.tiles_x * a.tiles_y;
    tma::EncodeTiledFn fn = tma::encode_fn();
    if (!fn) return cudaErrorNotSupported;
    CUtensorMap map{};
    const cuuint64_t dims[4] = {static_cast<cuuint64_t>(a.Cin), static_cast<cuuint64_t>(a.W), static_cast<cuuint64_t>(a.H), static_cast<cuuint64_t>(a.B)};
    const cuuint64_t st[3] = {static_cast<cuuint64_t>(ldx) * 2, static_cast<cuuint64_t>(a.W) * ldx * 2, static_cast<cuuint64_t>(a.H) * a.W * ldx * 2};
    const cuuint32_t box[4] = {SLAB, HX, HY, 1u};
    const cuuint32_t es[4] = {1u, 1u, 1u, 1u};
    if (fn(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return cudaErrorNotSupported;
    const size_t smem = 1024 + 2 * TILE_STRIDE + static_cast<size_t>(a.Cin / SLAB) * 9 * 4 * 32 * 8 + 64;
    cudaError_t e = cudaFuncSetAttribute(outproj_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    int grid = 2 * num_sms;
    if (grid > a.tiles) grid = a.tiles;
    outproj_mma_kernel<<<grid, THREADS, smem, stream>>>(a, map);
    return cudaGetLastError();
}
}  // namespace op

template <int KCH, int BN>
inline cudaError_t launch_inst(Args& a, const CUtensorMap& amap, const CUtensorMap& wmap, int num_sms, cudaStream_t stream) {
    constexpr int A_CHUNK = 128 * KCH * 2, W_CHUNK = BN * KCH * 2, STAGE = A_CHUNK + ((W_CHUNK + 1023) / 1024) * 1024;
    const size_t fixed = 1024 + NEW * STG_BUF + static_cast<size_t>(a.col_tiles) * BN * 4 + (2 * 8 + 4) * 8 + 16;
    int S = static_cast<int>((SMEM_MAX - fixed) / STAGE);
    if (S > 8) S = 8;
    if (S < 2) return cudaErrorInvalidConfiguration;
    const size_t smem = fixed + static_cast<size_t>(S) * STAGE;
    auto k = conv_igemm_kernel<KCH, BN>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    int grid = num_sms < a.tiles ? num_sms : a.tiles;
    k<<<grid, THREADS, smem, stream>>>(a, amap, wmap, S);
    return cudaGetLastError();
}

}  // namespace cv
}  // namespace lewin
