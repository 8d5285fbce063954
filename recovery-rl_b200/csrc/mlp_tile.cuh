// mlp_tile.cuh -- the shared-memory tiled fp32 building block of the SIMT MLP kernels (agent.cu, mpc.cu):
// one CTA pushes a tile of BM rows through  layer 1 (K = 2|4)  ->  256x256 GEMM (k-major operand streamed with
// cp.async, double-buffered)  ->  +bias / ReLU  ->  up to four head outputs per row.
#pragma once
#include "agent_common.cuh"

namespace rrl {

constexpr int kThreads = 256;
constexpr int KC = 16;  // k-chunk of the streamed operand

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// ---------------------------------------------------------------------------------------------
// forward tile: BM rows through one head.  Result: S.raw[m][0..n_out) = W3 h2 + b3.
// ---------------------------------------------------------------------------------------------
template <int BM>
struct FwdSmem {
    float As[H][BM];      // h1 tile, k-major (operand A)
    float Bs[2][KC][H];   // streamed W2T chunks (operand B)
    float W1s[H][4];
    float b1s[H], b2s[H];
    float w3s[4][H];
    float b3s[4];
    float xin[4][BM];     // s0, s1, a0, a1
    float raw[BM][4];
    float aux[BM][4];     // kernel-specific per-row stash
    int flag;
};

template <int BM>
__device__ __forceinline__ void load_b_chunk(FwdSmem<BM>& S, const float* __restrict__ W2T, int c, int buf) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int idx = threadIdx.x + kThreads * j;  // 1024 float4 per chunk
        const int row = idx >> 6, c4 = idx & 63;
        cp_async16(&S.Bs[buf][row][c4 * 4], W2T + (size_t)(c * KC + row) * H + c4 * 4);
    }
    cp_async_commit();
}

// acc[r][j] = sum_k As[k][warp*RPW + r] * W[k][col_j],  cols {4*lane .. 4*lane+3, 128 + 4*lane .. +3}.
// Chunk 0 of W must already be in flight (load_b_chunk(S, W, 0, 0)); ends with a __syncthreads (As / Bs reusable).
template <int BM>
__device__ __forceinline__ void tile_gemm(FwdSmem<BM>& S, const float* __restrict__ W, float (&acc)[BM / 8][8]) {
    constexpr int RPW = BM / 8;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[r][j] = 0.f;
    for (int c = 0; c < H / KC; ++c) {
        if (c + 1 < H / KC) {
            load_b_chunk(S, W, c + 1, (c + 1) & 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const int buf = c & 1;
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
            float a[RPW];
#pragma unroll
            for (int r4 = 0; r4 < RPW / 4; ++r4) {
                const float4 av = *reinterpret_cast<const float4*>(&S.As[c * KC + kk][warp * RPW + r4 * 4]);
                a[r4 * 4 + 0] = av.x; a[r4 * 4 + 1] = av.y; a[r4 * 4 + 2] = av.z; a[r4 * 4 + 3] = av.w;
            }
            const float4 b0 = *reinterpret_cast<const float4*>(&S.Bs[buf][kk][lane * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&S.Bs[buf][kk][128 + lane * 4]);
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                acc[r][0] = fmaf(a[r], b0.x, acc[r][0]);
                acc[r][1] = fmaf(a[r], b0.y, acc[r][1]);
                acc[r][2] = fmaf(a[r], b0.z, acc[r][2]);
                acc[r][3] = fmaf(a[r], b0.w, acc[r][3]);
                acc[r][4] = fmaf(a[r], b1.x, acc[r][4]);
                acc[r][5] = fmaf(a[r], b1.y, acc[r][5]);
                acc[r][6] = fmaf(a[r], b1.z, acc[r][6]);
                acc[r][7] = fmaf(a[r], b1.w, acc[r][7]);
            }
        }
        __syncthreads();
    }
}

template <int BM>
__device__ void mlp_tile_forward(FwdSmem<BM>& S, const HeadW& w, float* __restrict__ h1_out, float* __restrict__ h2_out,
                                 int64_t row0, int64_t rows) {
    constexpr int RPW = BM / 8;  // rows per warp
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    load_b_chunk(S, w.W2T, 0, 0);
    {  // stage the small tensors (t == hidden unit)
        const int n_in = w.n_in;
        float4 w1 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n_in == 4) {
            w1 = *reinterpret_cast<const float4*>(w.W1 + t * 4);
        } else {
            const float2 v = *reinterpret_cast<const float2*>(w.W1 + t * 2);
            w1.x = v.x; w1.y = v.y;
        }
        *reinterpret_cast<float4*>(S.W1s[t]) = w1;
        S.b1s[t] = w.b1[t];
        S.b2s[t] = w.b2[t];
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            float v = 0.f;
            if (o < w.na) v = w.W3a[o * H + t];
            else if (o < w.na + w.nb) v = w.W3b[(o - w.na) * H + t];
            S.w3s[o][t] = v;
        }
        if (t < 4) {
            float v = 0.f;
            if (t < w.na) v = w.b3a[t];
            else if (t < w.na + w.nb) v = w.b3b[t - w.na];
            S.b3s[t] = v;
        }
    }
    __syncthreads();  // small tensors + xin (written by the caller) visible
    {  // layer 1: h1 = relu(W1 x + b1)   (model.py:68,191,318,513)
        const int m = t % BM, kb = t / BM;
        constexpr int KSTEP = kThreads / BM;
        const float x0 = S.xin[0][m], x1 = S.xin[1][m], x2 = S.xin[2][m], x3 = S.xin[3][m];
        const bool store = h1_out != nullptr && (row0 + m) < rows;
        const bool four = w.n_in == 4;
#pragma unroll 4
        for (int k = kb; k < H; k += KSTEP) {
            const float4 wv = *reinterpret_cast<const float4*>(S.W1s[k]);
            float h = fmaf(wv.x, x0, S.b1s[k]);
            h = fmaf(wv.y, x1, h);
            if (four) {
                h = fmaf(wv.z, x2, h);
                h = fmaf(wv.w, x3, h);
            }
            h = fmaxf(h, 0.f);
            S.As[k][m] = h;
            if (store) h1_out[(row0 + m) * H + k] = h;
        }
    }
    float acc[RPW][8];
    tile_gemm<BM>(S, w.W2T, acc);
    // epilogue: h2 = relu(acc + b2); heads
    const int n_out = w.na + w.nb;
    const float4 bb0 = *reinterpret_cast<const float4*>(&S.b2s[lane * 4]);
    const float4 bb1 = *reinterpret_cast<const float4*>(&S.b2s[128 + lane * 4]);
    float myraw[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        float h[8];
        h[0] = fmaxf(acc[r][0] + bb0.x, 0.f); h[1] = fmaxf(acc[r][1] + bb0.y, 0.f);
        h[2] = fmaxf(acc[r][2] + bb0.z, 0.f); h[3] = fmaxf(acc[r][3] + bb0.w, 0.f);
        h[4] = fmaxf(acc[r][4] + bb1.x, 0.f); h[5] = fmaxf(acc[r][5] + bb1.y, 0.f);
        h[6] = fmaxf(acc[r][6] + bb1.z, 0.f); h[7] = fmaxf(acc[r][7] + bb1.w, 0.f);
        const int64_t row = row0 + warp * RPW + r;
        if (h2_out != nullptr && row < rows) {
            *reinterpret_cast<float4*>(h2_out + row * H + lane * 4) = make_float4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<float4*>(h2_out + row * H + 128 + lane * 4) = make_float4(h[4], h[5], h[6], h[7]);
        }
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            if (o < n_out) {
                const float4 w0 = *reinterpret_cast<const float4*>(&S.w3s[o][lane * 4]);
                const float4 w1 = *reinterpret_cast<const float4*>(&S.w3s[o][128 + lane * 4]);
                float p = h[0] * w0.x;
                p = fmaf(h[1], w0.y, p); p = fmaf(h[2], w0.z, p); p = fmaf(h[3], w0.w, p);
                p = fmaf(h[4], w1.x, p); p = fmaf(h[5], w1.y, p); p = fmaf(h[6], w1.z, p); p = fmaf(h[7], w1.w, p);
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) p += __shfl_xor_sync(0xffffffffu, p, s);
                if (lane == r) myraw[o] = p + S.b3s[o];
            }
        }
    }
    if (lane < RPW) *reinterpret_cast<float4*>(S.raw[warp * RPW + lane]) = make_float4(myraw[0], myraw[1], myraw[2], myraw[3]);
    __syncthreads();
}


}  // namespace rrl
