// fp32 token GEMM on tcgen05 (kind::tf32) with 3xTF32 error compensation -- the reference's own inference precision
// (test_long_GPU.py:91 runs the model in fp32, no autocast) on the 5th-generation tensor cores:
//
//   Y = epi( LN?(A)[M,K] * W[N,K]^T + bias ),       A, W, Y fp32
//
// Same contract as gemm_fused_kernel<float> (q|k|v projections attn.py:420-422, out projection :456, LeFF linear1
// My_model_1.py:508, linear2 :529; LN from precomputed row statistics, roll + window_partition / window_reverse as row
// addressing, bias, exact erf GELU, DropPath * residual), which it replaces for the forward pass.  An fp32 operand x is
// split into hi = tf32(x) and lo = tf32(x - hi); a product is accumulated as a_lo*b_hi + a_hi*b_lo + a_hi*b_hi (the
// a_lo*b_lo term is below 2^-22 relative), so the result is fp32-grade -- the top-u selection downstream of q|k|v is
// precision-critical (SURVEY finding 9) -- at three tensor-core passes instead of one.
//
// One persistent CTA per SM, 17 warps, every hand-over an mbarrier:
//   * 8 producer warps : raw fp32 k-chunks (32 floats = one 128-byte swizzle row) of A (gathered through roll +
//                        window_partition) and of W arrive by cp.async, S - 1 stages ahead, directly at their swizzled
//                        position in the stage's hi tiles; once landed, the thread that issued a 16-byte piece applies
//                        LayerNorm (A only), splits it and writes hi in place and lo into the stage's lo tile;
//   * 1 MMA thread     : per k-chunk 4 k8-steps x 3 tcgen05.mma kind::tf32 (M = 128, N = BN) into one of two TMEM
//                        accumulator stages; tcgen05.commit frees the stage / publishes the accumulator;
//   * 8 epilogue warps : tcgen05.ld (thread == row) -> bias -> erf GELU | DropPath * residual -> per-warp staging tile ->
//                        row-cooperative coalesced 16-byte stores through window_reverse + un-roll.
// Tiles are walked column-fastest so the CTAs that work on one row band share its A rows in L2.
#pragma once
#include "gemm_fused.cuh"
#include "tc_helpers.cuh"

namespace lewin {
namespace t32 {

constexpr int BM = 128;
constexpr int KC = 32;                        // floats per k-chunk: 128-byte rows (SWIZZLE_128B)
constexpr int NEW = 8, NPW = 8;               // epilogue / producer warps
constexpr int MMA_WARP = NEW;
constexpr int WARPS = NEW + 1 + NPW;
constexpr int THREADS = WARPS * 32;           // 544
constexpr int PTHREADS = NPW * 32;
constexpr int A_TILE = BM * 128;              // bytes of one [128 x 32] fp32 tile
constexpr int STG_ROW = 80;                   // staging row: 16 fp32 columns (64 B) + 16 B pad (conflict-free 16-byte accesses)
constexpr int STG_BUF = 32 * STG_ROW;
constexpr int NTAB = 8;                       // per-tile row tables in flight: issue runs up to 3 stages (tiles) ahead of convert, and a slow thread may still convert one tile behind the barrier
constexpr int SMEM_MAX = 227 * 1024;

inline bool enabled() {
    static const bool on = [] { const char* e = getenv("LEWIN_NO_T32_GEMM"); return !(e && e[0] == '1'); }();
    return on;
}

__device__ __forceinline__ void mma_tf32_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void producer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(PTHREADS) : "memory"); }

// byte offset of 16-byte piece `c` (0..7) of row `r` in a SWIZZLE_128B tile of 128-byte rows
__device__ __forceinline__ uint32_t swz(int r, int c) { return (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4); }

__device__ __forceinline__ void split4(const float4 v, float4& hi, float4& lo) {
    hi.x = __uint_as_float(f2tf32(v.x)); hi.y = __uint_as_float(f2tf32(v.y));
    hi.z = __uint_as_float(f2tf32(v.z)); hi.w = __uint_as_float(f2tf32(v.w));
    lo.x = __uint_as_float(f2tf32(v.x - hi.x)); lo.y = __uint_as_float(f2tf32(v.y - hi.y));
    lo.z = __uint_as_float(f2tf32(v.z - hi.z)); lo.w = __uint_as_float(f2tf32(v.w - hi.w));
}

template <int BN>
constexpr size_t stage_bytes() { return 2 * A_TILE + 2 * static_cast<size_t>(BN) * 128; }
constexpr size_t fixed_smem() {
    return 1024 /*align*/ + NEW * STG_BUF + NTAB * BM * 12 + (2 * 8 + 4) * 8 + 16;
}

template <int BN, int EPI>
__global__ void __launch_bounds__(THREADS, 1) gemm_t32_kernel(const GemmArgs<float> g, int row_tiles, int col_tiles, int nkc, int S) {
    constexpr int W_TILE = BN * 128;
    constexpr int STAGE = 2 * A_TILE + 2 * W_TILE;
    constexpr int ACC = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;     // TMEM columns per accumulator stage
    constexpr int TMEM_COLS = 2 * ACC;
    constexpr int NCH = BN / 32;                   // 32-column epilogue chunks
    constexpr int WP = BN / 32;                    // producer passes over the W rows
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(BN >> 3) << 17) |
                               (static_cast<uint32_t>(BM >> 4) << 24);            // D = f32, A = B = tf32, K-major
    static_assert(BN % 32 == 0 && BN <= 256, "tile width");

    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* base = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* ring = base;                                         // [S][A_hi | A_lo | W_hi | W_lo]
    unsigned char* stg = ring + static_cast<size_t>(S) * STAGE;         // [NEW][STG_BUF]
    uint32_t* t_tok = reinterpret_cast<uint32_t*>(stg + NEW * STG_BUF); // [NTAB][BM] A row (token) index, 0xFFFFFFFF = beyond M
    float* t_mu = reinterpret_cast<float*>(t_tok + NTAB * BM);          // [NTAB][BM]
    float* t_rs = t_mu + NTAB * BM;                                     // [NTAB][BM]
    uint64_t* full = reinterpret_cast<uint64_t*>(t_rs + NTAB * BM);     // [8]
    uint64_t* empty = full + 8;                                         // [8]
    uint64_t* tfull = empty + 8;                                        // [2]
    uint64_t* tempty = tfull + 2;                                       // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int total_tiles = row_tiles * col_tiles;
    const int my_tiles = (total_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    const bool has_ln = g.mean != nullptr;

    if (tid == 0) {
        for (int i = 0; i < S; ++i) { tc::mbar_init(&full[i], PTHREADS); tc::mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { tc::mbar_init(&tfull[i], 1); tc::mbar_init(&tempty[i], NEW * 32); }
        tc::fence_barrier_init();
    }
    if (warp == MMA_WARP) tc::tmem_alloc<TMEM_COLS>(tmem_slot);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;

    if (warp > MMA_WARP) {
        // ============================================================ producers
        const int pt = tid - (MMA_WARP + 1) * 32;      // 0 .. 255
        const int row_l = pt >> 3, ch = pt & 7;        // 8 lanes x 16 B = one 128-byte row piece; 32 rows per pass
        const uint32_t ring_u = tc::smem_u32(ring);
        const int DD = S - 1 < 3 ? S - 1 : 3;          // stages issued ahead
        const int total = my_tiles * nkc;
        // issue side
        int i_it = 0, i_kc = 0, i_s = 0;
        uint32_t i_ph = 0;
        auto issue = [&](int j) {
            if (j < total) {
                tc::mbar_wait(&empty[i_s], i_ph ^ 1u);
                const int t = static_cast<int>(blockIdx.x) + i_it * static_cast<int>(gridDim.x);
                const int rt = t / col_tiles, ct = t - rt * col_tiles;
                const int tb = (i_it & (NTAB - 1)) * BM;
                if (i_kc == 0) {                       // per-tile row table: gathered token index + LayerNorm statistics
                    if (pt < BM) {
                        const uint32_t m = static_cast<uint32_t>(rt) * BM + pt;
                        uint32_t tok = 0xFFFFFFFFu;
                        float mu = 0.f, rs = 1.f;
                        if (m < g.M) {
                            tok = g.mapA ? g.map.token32(m) : m;
                            if (has_ln) { mu = g.mean[tok]; rs = g.rstd[tok]; }
                        }
                        t_tok[tb + pt] = tok; t_mu[tb + pt] = mu; t_rs[tb + pt] = rs;
                    }
                    producer_bar();
                }
                const uint32_t st_u = ring_u + static_cast<uint32_t>(i_s) * STAGE;
                const int k0 = i_kc * KC + ch * 4;
#pragma unroll
                for (int p = 0; p < BM / 32; ++p) {
                    const int r = p * 32 + row_l;
                    const uint32_t tok = t_tok[tb + r];
                    const bool ok = tok != 0xFFFFFFFFu;
                    cp_async16_z(st_u + swz(r, ch), ok ? g.A + static_cast<long long>(tok) * g.lda + k0 : g.A, ok);
                }
                const float* wsrc = g.Wt + static_cast<long long>(ct * BN + row_l) * g.K + k0;
#pragma unroll
                for (int p = 0; p < WP; ++p)
                    cp_async16_z(st_u + 2 * A_TILE + swz(p * 32 + row_l, ch), wsrc + static_cast<long long>(p) * 32 * g.K, true);
                if (++i_kc == nkc) { i_kc = 0; ++i_it; }
                if (++i_s == S) { i_s = 0; i_ph ^= 1u; }
            }
            cp_async_commit();
        };
        // convert side
        int c_it = 0, c_kc = 0, c_s = 0;
        auto convert = [&]() {
            unsigned char* st = ring + static_cast<size_t>(c_s) * STAGE;
            const int tb = (c_it & (NTAB - 1)) * BM;
            float4 gam = make_float4(1.f, 1.f, 1.f, 1.f), bet = make_float4(0.f, 0.f, 0.f, 0.f);
            if (has_ln) {
                gam = *reinterpret_cast<const float4*>(g.ln_w + c_kc * KC + ch * 4);
                bet = *reinterpret_cast<const float4*>(g.ln_b + c_kc * KC + ch * 4);
            }
#pragma unroll
            for (int p = 0; p < BM / 32; ++p) {
                const int r = p * 32 + row_l;
                unsigned char* a = st + swz(r, ch);
                float4 v = *reinterpret_cast<const float4*>(a);
                if (has_ln) {
                    const float mu = t_mu[tb + r], rs = t_rs[tb + r];
                    v.x = (v.x - mu) * rs * gam.x + bet.x;
                    v.y = (v.y - mu) * rs * gam.y + bet.y;
                    v.z = (v.z - mu) * rs * gam.z + bet.z;
                    v.w = (v.w - mu) * rs * gam.w + bet.w;
                }
                float4 hi, lo;
                split4(v, hi, lo);
                *reinterpret_cast<float4*>(a) = hi;
                *reinterpret_cast<float4*>(a + A_TILE) = lo;
            }
#pragma unroll
            for (int p = 0; p < WP; ++p) {
                unsigned char* w = st + 2 * A_TILE + swz(p * 32 + row_l, ch);
                const float4 v = *reinterpret_cast<const float4*>(w);
                float4 hi, lo;
                split4(v, hi, lo);
                *reinterpret_cast<float4*>(w) = hi;
                *reinterpret_cast<float4*>(w + W_TILE) = lo;
            }
            tc::fence_proxy_async();                   // generic-proxy smem writes -> visible to the tensor core
            mbar_arrive(&full[c_s]);
            if (++c_kc == nkc) { c_kc = 0; ++c_it; }
            if (++c_s == S) c_s = 0;
        };
        for (int d = 0; d < DD; ++d) issue(d);
        for (int j = 0; j < total; ++j) {
            issue(j + DD);
            if (DD == 3) cp_async_wait<3>(); else if (DD == 2) cp_async_wait<2>(); else cp_async_wait<1>();
            convert();
        }
    } else if (warp == MMA_WARP) {
        // ============================================================ MMA issuer (one thread)
        if (lane == 0) {
            const uint32_t ring_u = tc::smem_u32(ring);
            int s = 0;
            uint32_t ph = 0;
            for (int it = 0; it < my_tiles; ++it) {
                const int acc = it & 1;
                const uint32_t aph = static_cast<uint32_t>(it >> 1) & 1u;
                tc::mbar_wait(&tempty[acc], aph ^ 1u);             // epilogue drained this accumulator
                tc::tc_fence_after();
                const uint32_t d_addr = tmem_d + static_cast<uint32_t>(acc * ACC);
                for (int kc = 0; kc < nkc; ++kc) {
                    tc::mbar_wait(&full[s], ph);
                    tc::tc_fence_after();
                    const uint32_t st_u = ring_u + static_cast<uint32_t>(s) * STAGE;
                    const uint64_t a_hi = tc::make_desc<64>(st_u), a_lo = tc::make_desc<64>(st_u + A_TILE);
                    const uint64_t w_hi = tc::make_desc<64>(st_u + 2 * A_TILE), w_lo = tc::make_desc<64>(st_u + 2 * A_TILE + W_TILE);
#pragma unroll
                    for (int k8 = 0; k8 < KC / 8; ++k8) {          // small terms first
                        mma_tf32_ss(d_addr, a_lo + 2 * k8, w_hi + 2 * k8, IDESC, (kc > 0 || k8 > 0) ? 1u : 0u);
                        mma_tf32_ss(d_addr, a_hi + 2 * k8, w_lo + 2 * k8, IDESC, 1u);
                        mma_tf32_ss(d_addr, a_hi + 2 * k8, w_hi + 2 * k8, IDESC, 1u);
                    }
                    tc::mma_commit(&empty[s]);                     // stage free once these MMAs have read it
                    if (++s == S) { s = 0; ph ^= 1u; }
                }
                tc::mma_commit(&tfull[acc]);                       // accumulator complete
            }
        }
    } else {
        // ============================================================ epilogue: thread == TMEM lane == tile row
        const int lg = warp & 3, half = warp >> 2;     // TMEM lane group, column group
        unsigned char* sb = stg + warp * STG_BUF;
        unsigned char* srow = sb + lane * STG_ROW;
        for (int it = 0; it < my_tiles; ++it) {
            const int t = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
            const int rt = t / col_tiles, ct = t - rt * col_tiles;
            const int n0 = ct * BN;
            const int acc = it & 1;
            const uint32_t aph = static_cast<uint32_t>(it >> 1) & 1u;
            const uint32_t m = static_cast<uint32_t>(rt) * BM + lg * 32 + lane;
            long long oy = -1, orr = -1;
            float sc = 1.f;
            if (m < g.M) {
                const uint32_t ry = g.mapY ? g.map.token32(m) : m;
                oy = static_cast<long long>(ry) * g.ldy;
                orr = static_cast<long long>(ry) * (g.ldr ? g.ldr : g.ldy);
                if (EPI == EPI_BIAS_RESID && g.drop_scale) sc = g.drop_scale[ry / static_cast<uint32_t>(g.tokens_per_image)];
            }
            // row-cooperative pass over a staged [32 rows x 16 columns] piece: 4 lanes x 16 B per row, 8 rows per instruction
            auto flush16 = [&](float* dst, int col, bool resid) {
                __syncwarp();
                float4 val[4], rv[4];
                long long o[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int i = lane + 32 * jj, rl = i >> 2, cc = i & 3;
                    o[jj] = __shfl_sync(0xffffffffu, oy, rl);
                    const long long orl = __shfl_sync(0xffffffffu, orr, rl);
                    if (resid && o[jj] >= 0) rv[jj] = *reinterpret_cast<const float4*>(g.R + orl + col + cc * 4);
                    val[jj] = *reinterpret_cast<const float4*>(sb + rl * STG_ROW + cc * 16);
                }
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int i = lane + 32 * jj, rl = i >> 2, cc = i & 3;
                    if (resid) {
                        const float s = __shfl_sync(0xffffffffu, sc, rl);
                        if (o[jj] >= 0) {
                            val[jj].x = rv[jj].x + s * val[jj].x; val[jj].y = rv[jj].y + s * val[jj].y;
                            val[jj].z = rv[jj].z + s * val[jj].z; val[jj].w = rv[jj].w + s * val[jj].w;
                        }
                    }
                    if (o[jj] >= 0) *reinterpret_cast<float4*>(dst + o[jj] + col + cc * 4) = val[jj];
                }
                __syncwarp();
            };
            tc::mbar_wait(&tfull[acc], aph);
            tc::tc_fence_after();
            const uint32_t t_addr = tmem_d + (static_cast<uint32_t>(lg * 32) << 16) + static_cast<uint32_t>(acc * ACC);
            for (int c = half; c < NCH; c += NEW / 4) {
                float v[32];
                tc::tmem_ld32(t_addr + c * 32, v);
                if (c + NEW / 4 >= NCH) {              // last chunk of this warp is in registers: hand the accumulator back
                    tc::tc_fence_before();
                    mbar_arrive(&tempty[acc]);
                }
                const int col0 = n0 + c * 32;
                if (g.bias) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(g.bias + col0 + j));
                        v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
                    }
                }
                if (EPI == EPI_BIAS_GELU && g.Y2) {    // training: pre-activation copy
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4)
                            *reinterpret_cast<float4*>(srow + j * 4) = make_float4(v[hh * 16 + j], v[hh * 16 + j + 1], v[hh * 16 + j + 2], v[hh * 16 + j + 3]);
                        flush16(g.Y2, col0 + hh * 16, false);
                    }
                }
                if (EPI == EPI_BIAS_GELU) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
                }
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        *reinterpret_cast<float4*>(srow + j * 4) = make_float4(v[hh * 16 + j], v[hh * 16 + j + 1], v[hh * 16 + j + 2], v[hh * 16 + j + 3]);
                    flush16(g.Y, col0 + hh * 16, EPI == EPI_BIAS_RESID);
                }
            }
            if (half >= NCH) {                         // this warp owns no chunk (BN == 32): still hand back
                tc::tc_fence_before();
                mbar_arrive(&tempty[acc]);
            }
        }
    }

    // ---------------------------------------------------------------- teardown
    tc::tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) tc::tmem_dealloc<TMEM_COLS>(tmem_d);
}

template <int BN, int EPI>
cudaError_t launch_bn(const GemmArgs<float>& g, int num_sms, cudaStream_t stream) {
    constexpr size_t STAGE = stage_bytes<BN>();
    constexpr size_t fixed = fixed_smem();
    static_assert(fixed + 2 * STAGE <= SMEM_MAX, "two stages must fit");
    int S = static_cast<int>((SMEM_MAX - fixed) / STAGE);
    if (S > 8) S = 8;
    const size_t smem = fixed + static_cast<size_t>(S) * STAGE;
    auto k = gemm_t32_kernel<BN, EPI>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    const int row_tiles = static_cast<int>((g.M + BM - 1) / BM);
    const int col_tiles = g.N / BN;
    int grid = num_sms;
    if (grid > row_tiles * col_tiles) grid = row_tiles * col_tiles;
    k<<<grid, THREADS, smem, stream>>>(g, row_tiles, col_tiles, g.K / KC, S);
    return cudaGetLastError();
}

// forward-pass shapes of the LeWin block: K, N multiples of 32, 16-byte aligned rows, no backward-only prologues
template <int EPI>
inline bool supported(const GemmArgs<float>& g) {
    if (!enabled() || g.a_row_scale || g.aux || g.up2) return false;
    if (EPI != EPI_BIAS && EPI != EPI_BIAS_GELU && EPI != EPI_BIAS_RESID) return false;
    if (EPI == EPI_BIAS_RESID && !g.R) return false;
    if (g.K % 32 || g.N % 32 || g.M <= 0 || g.M >= (1ll << 31)) return false;
    if ((g.lda % 4) || (g.ldy % 4) || (g.ldr % 4)) return false;
    if (g.mean && (!g.rstd || !g.ln_w || !g.ln_b)) return false;
    return true;
}

template <int EPI>
cudaError_t launch(const GemmArgs<float>& g, int num_sms, cudaStream_t stream) {
    if (g.N % 128 == 0) return launch_bn<128, EPI>(g, num_sms, stream);
    if (g.N % 96 == 0) return launch_bn<96, EPI>(g, num_sms, stream);
    if (g.N % 64 == 0) return launch_bn<64, EPI>(g, num_sms, stream);
    return launch_bn<32, EPI>(g, num_sms, stream);
}

}  // namespace t32
}  // namespace lewin
