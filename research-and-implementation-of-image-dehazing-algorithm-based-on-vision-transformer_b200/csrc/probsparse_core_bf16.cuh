// ProbSparse window-attention core, forward, bf16 specialisation (same algorithm and rounding points as
// probsparse_core_fwd_kernel<__nv_bfloat16>; reference ProbSparse/attn.py:287-342).
//
// Instruction-lean restatement for the instruction-issue-bound regime measured on B200 (profiles/): q, k, v stay bf16
// in shared memory (staged with raw 16-byte copies), S = Q K^T and P.V run on mma.sync.m16n8k16 bf16 with ldmatrix
// (ldmatrix.trans for V), the sampled-key statistics use a packed half2 {multiplicity, 0/-inf} table, the top-u rank
// count uses all 128 threads, and the double softmax processes two rows per warp (16 lanes x 4 columns).
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"
#include "probsparse_core.cuh"

namespace lewin {

namespace pc {
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
    const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
    const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
}  // namespace pc

// {multiplicity, 0 or -inf} per (query n, key m) as half2, built from index_sample (attn.py:91) next to `cnt`
__global__ void __launch_bounds__(256) build_cw_kernel(const int32_t* __restrict__ idx, __half2* __restrict__ cw) {
    __shared__ int cnt[kTok * kTok];
    for (int i = threadIdx.x; i < kTok * kTok; i += blockDim.x) cnt[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < kTok * kSampleK; i += blockDim.x) atomicAdd(&cnt[(i / kSampleK) * kTok + (idx[i] & 63)], 1);
    __syncthreads();
    for (int i = threadIdx.x; i < kTok * kTok; i += blockDim.x) {
        const int c = cnt[i];
        cw[i] = __halves2half2(__int2half_rn(c), c ? __float2half(0.f) : __ushort_as_half(0xFC00));
    }
}

constexpr int PC_LD = 40;      // bf16 row stride of q/k/v tiles (80 B: conflict-free ldmatrix)
constexpr int PC_PLD = 72;     // bf16 row stride of the P2 tile
constexpr int PC_SLD = 68;     // fp32 row stride of the selected-score tile

struct CoreBf16Smem {
    __nv_bfloat16 q[kTok * PC_LD];
    __nv_bfloat16 k[kTok * PC_LD];
    __nv_bfloat16 v[kTok * PC_LD];
    __nv_bfloat16 p2[32 * PC_PLD];
    float sc[32 * PC_SLD];         // scaled scores of the selected rows
    float M[kTok];
    float vpart[4 * kHeadDim];
    float vmean[kHeadDim];
    float tbl[232];
    __half2 cw[kTok * kTok];
    int rank_part[2 * kTok];
    int slot_of[kTok];
    int tok_of[32];
    int region[kTok];
};

struct CoreBf16Args {
    const __nv_bfloat16* qkv; __nv_bfloat16* ctx; uint8_t* top;
    const float* rpb_table; const float* rpb_dense; const __half2* cw;
    const float* mask; int nW_mask;
    int B_, nH, C, use_rpb, shift, H, W, nWw, nWin;
};

__global__ void __launch_bounds__(CORE_THREADS, 4) probsparse_core_bf16_kernel(const CoreBf16Args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CoreBf16Smem& s = *reinterpret_cast<CoreBf16Smem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gq = lane >> 2, tq = lane & 3;
    const int C3 = 3 * a.C;
    const float scale = rsqrtf(static_cast<float>(kHeadDim));

    for (int i = tid; i < kTok * kTok / 4; i += CORE_THREADS)
        reinterpret_cast<uint4*>(s.cw)[i] = reinterpret_cast<const uint4*>(a.cw)[i];

    const int items = a.B_ * a.nH;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int wg = item / a.nH, h = item - wg * a.nH;
        __syncthreads();
        {   // stage q, k, v: 3 x 64 rows x 4 sixteen-byte chunks, raw copies
            const __nv_bfloat16* base = a.qkv + static_cast<long long>(wg) * kTok * C3 + h * kHeadDim;
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const int c = tid + i * CORE_THREADS;
                const int which = c >> 8, r = (c >> 2) & 63, ch = c & 3;
                const uint4 val = *reinterpret_cast<const uint4*>(base + static_cast<long long>(r) * C3 + which * a.C + ch * 8);
                __nv_bfloat16* dst = (which == 0 ? s.q : which == 1 ? s.k : s.v) + r * PC_LD + ch * 8;
                *reinterpret_cast<uint4*>(dst) = val;
            }
            if (a.use_rpb && a.rpb_table)
                for (int i = tid; i < 225; i += CORE_THREADS) s.tbl[i] = a.rpb_table[i * a.nH + h];
            if (a.shift > 0 && tid < kTok) {
                const int w = wg % a.nWin, wy = w / a.nWw, wx = w - wy * a.nWw;
                const int y = wy * 8 + (tid >> 3), x = wx * 8 + (tid & 7);
                const int rb = y < a.H - 8 ? 0 : (y < a.H - a.shift ? 1 : 2);
                const int cb = x < a.W - 8 ? 0 : (x < a.W - a.shift ? 1 : 2);
                s.region[tid] = rb * 3 + cb;
            }
        }
        __syncthreads();

        // ---- S = Q K^T (rows 16*warp..+15, 64 keys, K = 32)
        float acc[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[j][c] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            uint32_t af[4];
            pc::ldsm_x4(af, s.q + (warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * PC_LD + ks * 16 + (lane >> 4) * 8);
#pragma unroll
            for (int jp = 0; jp < 4; ++jp) {
                uint32_t bf[4];
                pc::ldsm_x4(bf, s.k + (jp * 16 + (lane & 7) + ((lane >> 4) & 1) * 8) * PC_LD + ks * 16 + ((lane >> 3) & 1) * 8);
                pc::mma16816(acc[2 * jp], af, bf[0], bf[1]);
                pc::mma16816(acc[2 * jp + 1], af, bf[2], bf[3]);
            }
        }
        // S~ is a bf16 matmul output under autocast: round once, keep the rounded values (A.4)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t lo = pc::pack2(acc[j][0], acc[j][1]), hi = pc::pack2(acc[j][2], acc[j][3]);
            acc[j][0] = __uint_as_float(lo << 16); acc[j][1] = __uint_as_float(lo & 0xFFFF0000u);
            acc[j][2] = __uint_as_float(hi << 16); acc[j][3] = __uint_as_float(hi & 0xFFFF0000u);
        }
        // ---- sparsity measure M_n = max_sampled S~ - (sum cnt * S~) / 64  (attn.py:117)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int r = warp * 16 + gq + half * 8;
            float mx = -INFINITY, sm = 0.f;
            const __half2* crow = s.cw + r * kTok + 2 * tq;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float2 c0 = __half22float2(crow[j * 8]);          // {cnt, neg} of column j*8 + 2*tq
                const float2 c1 = __half22float2(crow[j * 8 + 1]);
                const float s0 = acc[j][half * 2], s1 = acc[j][half * 2 + 1];
                mx = fmaxf(mx, fmaxf(s0 + c0.y, s1 + c1.y));
                sm = fmaf(c0.x, s0, sm);
                sm = fmaf(c1.x, s1, sm);
            }
            mx = group_max<4>(mx);
            sm = group_sum<4>(sm);
            if (tq == 0) s.M[r] = mx - sm * (1.0f / kTok);
        }
        __syncthreads();

        // ---- top-u by rank counting with all 128 threads (row = tid & 63, half of the comparisons each)
        {
            const int r = tid & 63, hf = tid >> 6;
            const float mine = s.M[r];
            int rank = 0;
#pragma unroll
            for (int m4 = 0; m4 < 8; ++m4) {
                const float4 o = *reinterpret_cast<const float4*>(s.M + hf * 32 + m4 * 4);
                const int m = hf * 32 + m4 * 4;
                rank += (o.x > mine) || (o.x == mine && m < r);
                rank += (o.y > mine) || (o.y == mine && m + 1 < r);
                rank += (o.z > mine) || (o.z == mine && m + 2 < r);
                rank += (o.w > mine) || (o.w == mine && m + 3 < r);
            }
            s.rank_part[hf * kTok + r] = rank;
        }
        __syncthreads();
        if (tid < kTok) {
            const int rank = s.rank_part[tid] + s.rank_part[kTok + tid];
            const int slot = rank < kTopU ? rank : -1;
            s.slot_of[tid] = slot;
            if (slot >= 0) {
                s.tok_of[slot] = tid;
                if (a.top) a.top[static_cast<long long>(item) * kTopU + slot] = static_cast<uint8_t>(tid);
            }
        } else if (tid < kTok + 7) {
            s.tok_of[kTopU + tid - kTok] = -1;
        }
        {   // column sums of V for the mean(V) fill (attn.py:168)
            const int d = tid & 31, part = tid >> 5;
            float sum = 0.f;
#pragma unroll
            for (int r = 0; r < 16; ++r) sum += __bfloat162float(s.v[(part * 16 + r) * PC_LD + d]);
            s.vpart[part * kHeadDim + d] = sum;
        }
        __syncthreads();

        // ---- selected rows: bf16(S) * scale -> bf16 -> fp32 tile (attn.py:150, 327-329)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int slot = s.slot_of[warp * 16 + gq + half * 8];
            if (slot >= 0) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t pk = pc::pack2(acc[j][half * 2] * scale, acc[j][half * 2 + 1] * scale);
                    *reinterpret_cast<float2*>(s.sc + slot * PC_SLD + j * 8 + 2 * tq) =
                        make_float2(__uint_as_float(pk << 16), __uint_as_float(pk & 0xFFFF0000u));
                }
            }
        }
        if (tid < kHeadDim)
            s.vmean[tid] = __bfloat162float(__float2bfloat16_rn(
                (s.vpart[tid] + s.vpart[32 + tid] + s.vpart[64 + tid] + s.vpart[96 + tid]) * (1.0f / kTok)));
        __syncthreads();

        // ---- softmax -> +rpb -> +mask -> softmax, two rows per warp (16 lanes x 4 columns each)
        {
            const int sub = lane >> 4, l16 = lane & 15, c0 = l16 * 4;
#pragma unroll 1
            for (int pass = 0; pass < 4; ++pass) {
                const int slot = pass * 8 + warp * 2 + sub;
                uint2 outp = make_uint2(0u, 0u);
                const bool live = slot < kTopU;              // uniform over the 16-lane group
                float x[4] = {0.f, 0.f, 0.f, 0.f};
                int r = 0;
                if (live) {
                    r = s.tok_of[slot];
                    const float4 t4 = *reinterpret_cast<const float4*>(s.sc + slot * PC_SLD + c0);
                    x[0] = t4.x; x[1] = t4.y; x[2] = t4.z; x[3] = t4.w;
                }
                float mx = group_max<16>(fmaxf(fmaxf(x[0], x[1]), fmaxf(x[2], x[3])));
                float e[4], sum = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) { e[j] = __expf(x[j] - mx); sum += e[j]; }
                float inv = __fdividef(1.0f, group_sum<16>(sum));
                if (live) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) x[j] = e[j] * inv;                 // P1
                    if (a.use_rpb) {
                        if (a.rpb_table) {
                            const int ry = r >> 3, rx = r & 7, cy = c0 >> 3, cx = c0 & 7;
                            const float* tb = s.tbl + (ry - cy + 7) * 15 + (rx - cx + 7);
                            x[0] += tb[0]; x[1] += tb[-1]; x[2] += tb[-2]; x[3] += tb[-3];
                        } else {
                            const float4 b4 = *reinterpret_cast<const float4*>(a.rpb_dense + (static_cast<long long>(h) * kTok + r) * kTok + c0);
                            x[0] += b4.x; x[1] += b4.y; x[2] += b4.z; x[3] += b4.w;
                        }
                    }
                    if (a.mask) {
                        const float4 m4 = *reinterpret_cast<const float4*>(a.mask + (static_cast<long long>(wg % a.nW_mask) * kTok + r) * kTok + c0);
                        x[0] += m4.x; x[1] += m4.y; x[2] += m4.z; x[3] += m4.w;
                    }
                    if (a.shift > 0) {
                        const int rr = s.region[r];
#pragma unroll
                        for (int j = 0; j < 4; ++j) x[j] += (s.region[c0 + j] != rr) ? -100.0f : 0.f;
                    }
                }
                mx = group_max<16>(fmaxf(fmaxf(x[0], x[1]), fmaxf(x[2], x[3])));
                sum = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) { e[j] = __expf(x[j] - mx); sum += e[j]; }
                inv = __fdividef(1.0f, group_sum<16>(sum));
                if (live) outp = make_uint2(pc::pack2(e[0] * inv, e[1] * inv), pc::pack2(e[2] * inv, e[3] * inv));   // P2 (bf16)
                *reinterpret_cast<uint2*>(s.p2 + slot * PC_PLD + c0) = outp;
            }
        }
        __syncthreads();

        // ---- ctx[top] = P2 . V (32 x 32 x 64): warp -> m-tile (warp & 1), d-columns 16 * (warp >> 1)
        {
            const int mt = warp & 1, nb = (warp >> 1) * 16;
            float o[2][4];
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int c = 0; c < 4; ++c) o[j][c] = 0.f;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                uint32_t af[4], bf[4];
                pc::ldsm_x4(af, s.p2 + (mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * PC_PLD + ks * 16 + (lane >> 4) * 8);
                pc::ldsm_x4_t(bf, s.v + (ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * PC_LD + nb + (lane >> 4) * 8);
                pc::mma16816(o[0], af, bf[0], bf[1]);
                pc::mma16816(o[1], af, bf[2], bf[3]);
            }
            __nv_bfloat16* cbase = a.ctx + static_cast<long long>(wg) * kTok * a.C + h * kHeadDim;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int r = s.tok_of[mt * 16 + gq + half * 8];
                if (r >= 0) {
#pragma unroll
                    for (int j = 0; j < 2; ++j)
                        *reinterpret_cast<uint32_t*>(cbase + static_cast<long long>(r) * a.C + nb + j * 8 + 2 * tq) =
                            pc::pack2(o[j][half * 2], o[j][half * 2 + 1]);
                }
            }
            // lazy queries: mean(V) (attn.py:172), 16-byte stores
            for (int c = tid; c < kTok * 4; c += CORE_THREADS) {
                const int r = c >> 2, d8 = (c & 3) * 8;
                if (s.slot_of[r] < 0) {
                    const float* vm = s.vmean + d8;
                    *reinterpret_cast<uint4*>(cbase + static_cast<long long>(r) * a.C + d8) =
                        make_uint4(pc::pack2(vm[0], vm[1]), pc::pack2(vm[2], vm[3]), pc::pack2(vm[4], vm[5]), pc::pack2(vm[6], vm[7]));
                }
            }
        }
    }
}

inline cudaError_t launch_core_bf16(const CoreBf16Args& a, int num_sms, cudaStream_t stream) {
    auto k = probsparse_core_bf16_kernel;
    const size_t smem = sizeof(CoreBf16Smem);
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    const long long items = static_cast<long long>(a.B_) * a.nH;
    const long long cap = static_cast<long long>(num_sms) * 4 * 4;
    k<<<static_cast<unsigned>(items < cap ? items : cap), CORE_THREADS, smem, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace lewin
