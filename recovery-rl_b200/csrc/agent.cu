// agent.cu -- SAC + safety-critic (Q_risk) + recovery-policy networks of Recovery RL on one flat arena.
//
// Replaces the torch modules / autograd / Adam calls of recovery_rl/model.py:49-76,172-199,295-343,
// 489-530, recovery_rl/sac.py:133-277, recovery_rl/qrisk.py:86-213, recovery_rl/utils.py:46-54 and the
// composite action selection of recovery_rl/experiment.py:546-577.
//
// All four MLPs are 2x256-hidden: the first layer (K = 2 or 4) and the heads (N = 1..4) are SIMT
// prologue/epilogue of the one 256x256 contraction, which is a shared-memory tiled fp32 GEMM here
// (agent_tc.cu holds the tcgen05 version of the same contraction for the large-N acting kernel).
#include "mlp_tile.cuh"

using namespace rrl;

namespace {

// ---------------------------------------------------------------------------------------------
// grouped forward kernel (training batches and the stand-alone forward entry points)
// ---------------------------------------------------------------------------------------------
template <int BM>
__global__ void __launch_bounds__(kThreads, (BM == 64) ? 2 : 3) mlp_forward_kernel(const __grid_constant__ FwdArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FwdSmem<BM>& S = *reinterpret_cast<FwdSmem<BM>*>(smem_raw);
    const int64_t rows = A.rows_ptr ? *A.rows_ptr : A.rows_const;
    const int64_t row0 = (int64_t)blockIdx.x * BM;
    if (row0 >= rows) return;
    const FwdPass& P = A.p[blockIdx.y];
    const int t = threadIdx.x;
    if (t < BM) {
        const int64_t row = row0 + t;
        float2 s = make_float2(0.f, 0.f), a = make_float2(0.f, 0.f);
        if (row < rows) {
            s = reinterpret_cast<const float2*>(P.xs)[row];
            if (P.xa) a = reinterpret_cast<const float2*>(P.xa)[row];
        }
        S.xin[0][t] = s.x; S.xin[1][t] = s.y; S.xin[2][t] = a.x; S.xin[3][t] = a.y;
    }
    mlp_tile_forward<BM>(S, P.w, P.h1, P.h2, row0, rows);
    if (t < BM) {
        const int64_t row = row0 + t;
        if (row < rows) {
            const float4 rv = *reinterpret_cast<const float4*>(S.raw[t]);
            const float raw[4] = {rv.x, rv.y, rv.z, rv.w};
            forward_tail(P, A, row, raw);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// fused acting kernel (experiment.py:546-577): policy -> Q_risk threshold -> recovery policy -> select
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 2) act_kernel(const __grid_constant__ ActArgs A) {
    constexpr int BM = 64;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FwdSmem<BM>& S = *reinterpret_cast<FwdSmem<BM>*>(smem_raw);
    const int t = threadIdx.x;
    const uint64_t vstep = A.counters ? (uint64_t)A.counters[RRL_C_VEC_STEP] : 0;
    const bool random_phase = A.counters && !A.eval && (A.start_steps > A.counters[RRL_C_TOTAL_NUMSTEPS]);
    for (int64_t row0 = (int64_t)blockIdx.x * BM; row0 < A.n; row0 += (int64_t)gridDim.x * BM) {
        const int64_t row = row0 + t;
        const bool live = t < BM && row < A.n;
        if (t < BM) {
            float sx = 0.f, sy = 0.f;
            if (live) {  // torch.FloatTensor(state): fp64 -> fp32 (sac.py:137)
                sx = (float)A.state[row];
                sy = (float)A.state[A.n + row];
            }
            S.xin[0][t] = sx; S.xin[1][t] = sy; S.xin[2][t] = 0.f; S.xin[3][t] = 0.f;
        }
        float at[2] = {0.f, 0.f};
        if (!random_phase) {
            mlp_tile_forward<BM>(S, A.pol, nullptr, nullptr, row0, A.n);
            if (live) {
                const float4 rv = *reinterpret_cast<const float4*>(S.raw[t]);
                const float raw[4] = {rv.x, rv.y, rv.z, rv.w};
                float e[2], mean_a[2], lp;
                if (A.eps_task) {
                    const float2 ev = reinterpret_cast<const float2*>(A.eps_task)[row];
                    e[0] = ev.x; e[1] = ev.y;
                } else {
                    philox_eps(A.seed, A.stream_id, (uint64_t)row, vstep, RRL_DRAW_ACT_TASK, e);
                    if (A.det) det_noise(e);
                }
                if (A.det) {  // DeterministicPolicy.sample (model.py:475-481): every env copy is its own sample() call
                    const float zero_ls[2] = {0.f, 0.f};
                    stoch_sample(raw, zero_ls, e, A.sp, at, mean_a, &lp);
                } else {
                    gauss_sample(raw, e, A.sp, at, &lp, mean_a);
                }
                if (A.eval) { at[0] = mean_a[0]; at[1] = mean_a[1]; }  // sac.py:166-167
            }
        } else if (live) {  // env.action_space.sample(): U(low, high)  (experiment.py:559-560)
            float u[2];
            if (A.rand_u) {
                const float2 uv = reinterpret_cast<const float2*>(A.rand_u)[row];
                u[0] = uv.x; u[1] = uv.y;
            } else {
                const Philox4 p = rrl_philox(A.seed, A.stream_id, (uint64_t)row, vstep, RRL_DRAW_ACT_RAND);
                u[0] = rrl_u24(p.x); u[1] = rrl_u24(p.y);
            }
            at[0] = fmaf(2.0f * u[0] - 1.0f, A.sp.scale[0], A.sp.bias[0]);
            at[1] = fmaf(2.0f * u[1] - 1.0f, A.sp.scale[1], A.sp.bias[1]);
        }
        bool rec = false;
        float qmax = 0.f;
        float ar[2] = {at[0], at[1]};
        if (A.use_recovery) {
            __syncthreads();
            if (t < BM) { S.xin[2][t] = at[0]; S.xin[3][t] = at[1]; }
            mlp_tile_forward<BM>(S, A.qr1, nullptr, nullptr, row0, A.n);
            const float q1 = (t < BM) ? sigmoidf_(S.raw[t][0]) : 0.f;
            __syncthreads();
            mlp_tile_forward<BM>(S, A.qr2, nullptr, nullptr, row0, A.n);
            if (live) {
                const float q2 = sigmoidf_(S.raw[t][0]);
                qmax = fmaxf(q1, q2);               // qrisk.py:196
                rec = qmax > A.eps_safe;            // experiment.py:555
            }
            if (t == 0) S.flag = 0;
            __syncthreads();
            if (rec) S.flag = 1;
            __syncthreads();
            if (S.flag) {  // at least one env of this tile recovers (block-uniform branch)
                mlp_tile_forward<BM>(S, A.rec, nullptr, nullptr, row0, A.n);
                if (rec) {
                    const float raw[2] = {S.raw[t][0], S.raw[t][1]};
                    float e[2], mean_a[2], lp;
                    if (A.eps_rec) {
                        const float2 ev = reinterpret_cast<const float2*>(A.eps_rec)[row];
                        e[0] = ev.x; e[1] = ev.y;
                    } else {
                        philox_eps(A.seed, A.stream_id, (uint64_t)row, vstep, RRL_DRAW_ACT_REC, e);
                    }
                    stoch_sample(raw, A.rec.log_std, e, A.sp, ar, mean_a, &lp);  // qrisk.py:207-213
                }
            }
        }
        if (live) {
            reinterpret_cast<float2*>(A.action_task)[row] = make_float2(at[0], at[1]);
            reinterpret_cast<float2*>(A.action_real)[row] = make_float2(ar[0], ar[1]);
            if (A.recovery) A.recovery[row] = rec ? 1 : 0;
            if (A.qrisk_out) A.qrisk_out[row] = qmax;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// backward building blocks (training batches: rows <= max_batch)
// ---------------------------------------------------------------------------------------------
// (1)+(2) head backward fused into the streamed GEMM.  dh2 = (dout W3) * relu'(h2) is never materialised: the
//     CTAs rebuild their A tile from h2 / dout / W3 while staging it.
//     C[m][n] = sum_k A[k][m] * B[k][n], n = 0..255, m tile of 32, both operands k-major.
//     DATA  : A(k = hidden, m = row) = dh2[m][k], B = W2 [H][H], C = dh1 [row][H] masked by h1 > 0
//     WEIGHT: A(k = row, m = out unit) = dh2[k][m], B = h1 [rows][H], C = gW2 [H][H]; the same CTAs also reduce
//             gb2[m] = sum_r dh2[r][m], gW3[o][m] = sum_r dout[r][o] h2[r][m] and (m0 == 0) gb3[o] = sum_r dout[r][o]
__global__ void __launch_bounds__(kThreads) gemm_stream_kernel(const __grid_constant__ GemmArgs G) {
    constexpr int BM = 32;
    const int64_t rows = *G.rows_ptr;
    if (rows <= 0) return;
    const GemmPass& P = G.p[blockIdx.y];
    const int m0 = blockIdx.x * BM;
    const int M = P.k_is_rows ? H : (int)rows;
    if (m0 >= M) return;
    const int K = P.k_is_rows ? (int)rows : H;
    __shared__ __align__(16) float As[2][KC][BM];
    __shared__ __align__(16) float Bs[2][KC][H];
    __shared__ __align__(16) float w3s[4][H];
    __shared__ __align__(16) float ds[BM][4];       // DATA: dout of the CTA's 32 rows
    float (*red)[8][21] = reinterpret_cast<float (*)[8][21]>(&Bs[0][0][0]);  // WEIGHT epilogue: partial sums of the
                                                                             // 16 row-lanes (Bs is free by then)
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const bool weight = P.k_is_rows != 0;
    {   // stage W3 (zero rows beyond n_out) and, for DATA passes, the tile's dout
#pragma unroll
        for (int o = 0; o < 4; ++o)
            w3s[o][t] = o < P.na ? P.W3a[o * H + t] : (o < P.n_out ? P.W3b[(o - P.na) * H + t] : 0.f);
        if (!weight && t < BM * 4) {
            const int m = t >> 2, o = t & 3;
            ds[m][o] = (m0 + m < M && o < P.n_out) ? P.dout[(size_t)(m0 + m) * P.stride + o] : 0.f;
        }
    }
    __syncthreads();
    // per-thread A staging role (threads 0..127):
    //   DATA  : row m = t >> 2, four consecutive k (kq = t & 3)       WEIGHT: row kk = t >> 3, four consecutive m (m4 = t & 7)
    float4 hreg = make_float4(0.f, 0.f, 0.f, 0.f);
    float dreg[4] = {0.f, 0.f, 0.f, 0.f};
    float acc_gb2[4] = {0.f, 0.f, 0.f, 0.f}, acc_gw[4][4], acc_gb3[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc_gw[o][j] = 0.f;
    auto fetch = [&](int c) {   // global loads of chunk c into registers (consumed by stage_a after the FMAs of chunk c-1)
        const int k0 = c * KC;
        if (t < 128) {
            if (weight) {
                const int kk = t >> 3, m4 = t & 7;
                if (k0 + kk < K) {
                    hreg = *reinterpret_cast<const float4*>(P.h2 + (size_t)(k0 + kk) * H + m0 + m4 * 4);
#pragma unroll
                    for (int o = 0; o < 4; ++o) dreg[o] = o < P.n_out ? P.dout[(size_t)(k0 + kk) * P.stride + o] : 0.f;
                } else {
                    hreg = make_float4(0.f, 0.f, 0.f, 0.f);
                    dreg[0] = dreg[1] = dreg[2] = dreg[3] = 0.f;
                }
            } else {
                const int m = t >> 2, kq = t & 3;
                hreg = (m0 + m < M) ? *reinterpret_cast<const float4*>(P.h2 + (size_t)(m0 + m) * H + k0 + kq * 4)
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int idx = t + kThreads * j;
            const int kk = idx >> 6, c4 = idx & 63;
            if (k0 + kk < K) cp_async16(&Bs[c & 1][kk][c4 * 4], P.B + (size_t)(k0 + kk) * H + c4 * 4);
            else *reinterpret_cast<float4*>(&Bs[c & 1][kk][c4 * 4]) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        cp_async_commit();
    };
    auto stage_a = [&](int c) {  // dh2 of the fetched elements -> As[c & 1]
        if (t >= 128) return;
        const int buf = c & 1;
        const float hv[4] = {hreg.x, hreg.y, hreg.z, hreg.w};
        if (weight) {
            const int kk = t >> 3, m4 = t & 7;
            float g[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int col = m0 + m4 * 4 + j;
                float v = dreg[0] * w3s[0][col];
                v = fmaf(dreg[1], w3s[1][col], v); v = fmaf(dreg[2], w3s[2][col], v); v = fmaf(dreg[3], w3s[3][col], v);
                g[j] = hv[j] > 0.f ? v : 0.f;
                acc_gb2[j] += g[j];
#pragma unroll
                for (int o = 0; o < 4; ++o) acc_gw[o][j] = fmaf(dreg[o], hv[j], acc_gw[o][j]);
            }
            if (m4 == 0) {
#pragma unroll
                for (int o = 0; o < 4; ++o) acc_gb3[o] += dreg[o];
            }
            *reinterpret_cast<float4*>(&As[buf][kk][m4 * 4]) = make_float4(g[0], g[1], g[2], g[3]);
        } else {
            const int m = t >> 2, kq = t & 3;
            const int k0 = c * KC;
            const float4 dv = *reinterpret_cast<const float4*>(ds[m]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int col = k0 + kq * 4 + j;
                float v = dv.x * w3s[0][col];
                v = fmaf(dv.y, w3s[1][col], v); v = fmaf(dv.z, w3s[2][col], v); v = fmaf(dv.w, w3s[3][col], v);
                As[buf][kq * 4 + j][m] = hv[j] > 0.f ? v : 0.f;
            }
        }
    };
    float acc[4][8];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[r][j] = 0.f;
    const int nchunks = (K + KC - 1) / KC;
    fetch(0);
    stage_a(0);
    for (int c = 0; c < nchunks; ++c) {
        if (c + 1 < nchunks) {
            fetch(c + 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();   // chunk c staged (A by stage_a, B by cp.async); buffer (c+1)&1 free since the sync below
        const int buf = c & 1;
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
            const float4 av = *reinterpret_cast<const float4*>(&As[buf][kk][warp * 4]);
            const float a[4] = {av.x, av.y, av.z, av.w};
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][lane * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][128 + lane * 4]);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                acc[r][0] = fmaf(a[r], b0.x, acc[r][0]); acc[r][1] = fmaf(a[r], b0.y, acc[r][1]);
                acc[r][2] = fmaf(a[r], b0.z, acc[r][2]); acc[r][3] = fmaf(a[r], b0.w, acc[r][3]);
                acc[r][4] = fmaf(a[r], b1.x, acc[r][4]); acc[r][5] = fmaf(a[r], b1.y, acc[r][5]);
                acc[r][6] = fmaf(a[r], b1.z, acc[r][6]); acc[r][7] = fmaf(a[r], b1.w, acc[r][7]);
            }
        }
        if (c + 1 < nchunks) stage_a(c + 1);   // writes As[(c+1)&1]: last read before the previous end-of-iteration sync
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int m = m0 + warp * 4 + r;
        if (m >= M) continue;
        float4 o0 = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
        float4 o1 = make_float4(acc[r][4], acc[r][5], acc[r][6], acc[r][7]);
        if (P.mask) {
            const float4 k0 = *reinterpret_cast<const float4*>(P.mask + (size_t)m * H + lane * 4);
            const float4 k1 = *reinterpret_cast<const float4*>(P.mask + (size_t)m * H + 128 + lane * 4);
            o0.x = k0.x > 0.f ? o0.x : 0.f; o0.y = k0.y > 0.f ? o0.y : 0.f;
            o0.z = k0.z > 0.f ? o0.z : 0.f; o0.w = k0.w > 0.f ? o0.w : 0.f;
            o1.x = k1.x > 0.f ? o1.x : 0.f; o1.y = k1.y > 0.f ? o1.y : 0.f;
            o1.z = k1.z > 0.f ? o1.z : 0.f; o1.w = k1.w > 0.f ? o1.w : 0.f;
        }
        *reinterpret_cast<float4*>(P.C + (size_t)m * H + lane * 4) = o0;
        *reinterpret_cast<float4*>(P.C + (size_t)m * H + 128 + lane * 4) = o1;
    }
    if (weight && P.gb2) {   // head-layer gradients: deterministic reduction over the 16 row-lanes
        if (t < 128) {
            const int kk = t >> 3, m4 = t & 7;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                red[kk][m4][j] = acc_gb2[j];
#pragma unroll
                for (int o = 0; o < 4; ++o) red[kk][m4][4 + o * 4 + j] = acc_gw[o][j];
            }
            if (m4 == 0) {  // gb3 partials: column 20 of slots 0..3 (one head output each), written by this thread only
#pragma unroll
                for (int o = 0; o < 4; ++o) red[kk][o][20] = acc_gb3[o];
            }
        }
        __syncthreads();
        if (t < 32 * 5) {   // 32 columns x {gb2, gW3[0..3]}
            const int q = t >> 5, mm = t & 31;  // q = 0: gb2, 1..4: gW3 row q-1
            float v = 0.f;
#pragma unroll
            for (int kk = 0; kk < 16; ++kk) v += red[kk][mm >> 2][q == 0 ? (mm & 3) : (4 + (q - 1) * 4 + (mm & 3))];
            const int col = m0 + mm;
            if (q == 0) P.gb2[col] = v;
            else {
                const int o = q - 1;
                if (o < P.na) P.gW3a[o * H + col] = v;
                else if (o < P.n_out) P.gW3b[(o - P.na) * H + col] = v;
            }
        } else if (t >= 192 && t < 196 && m0 == 0) {
            const int o = t - 192;
            float v = 0.f;
#pragma unroll
            for (int kk = 0; kk < 16; ++kk) v += red[kk][o][20];
            if (o < P.na) P.gb3a[o] = v;
            else if (o < P.n_out) P.gb3b[o - P.na] = v;
        }
    }
}

// (3) layer-1 backward: gW1, gb1 (first H/32 blocks of x) and d(action input) (remaining blocks of x).
//     Every warp keeps kL1Unroll independent row loads in flight (the batch is L2-resident: the kernel is a latency
//     chain, not a bandwidth problem).  Optional tail (update_tails.cuh): the policy-sample backward that consumes the
//     d(action) rows of all passes, run by the last CTA.
struct L1BwdPass {
    const float *dh1, *xs, *xa, *W1;
    int n_in;
    float *gW1, *gb1;  // NULL: skip
    float* dxa;        // [rows][2] gradient w.r.t. the action input; NULL: skip
};
struct L1BwdArgs {
    L1BwdPass p[6];
    const int64_t* rows_ptr;
    TailArgs tail;
};

constexpr int kL1ColBlocks = H / 32;  // blockIdx.x <  kL1ColBlocks : weight grads of 32 hidden units
constexpr int kL1RowBlocks = 8;       // blockIdx.x >= kL1ColBlocks : d(action) for a slice of the rows
constexpr int kL1Unroll = 8;

__global__ void __launch_bounds__(kThreads) layer1_backward_kernel(const __grid_constant__ L1BwdArgs A) {
    const int64_t rows = *A.rows_ptr;
    const L1BwdPass& P = A.p[blockIdx.y];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    __shared__ float red[5][8][32];
    if (rows > 0 && blockIdx.x < kL1ColBlocks) {
        if (P.gW1) {   // uniform per block
            const int j = blockIdx.x * 32 + lane;
            const bool four = P.n_in == 4;
            float g[4] = {0.f, 0.f, 0.f, 0.f}, gb = 0.f;
            for (int64_t r0 = warp; r0 < rows; r0 += 8 * kL1Unroll) {
                float d[kL1Unroll];
                float2 sv[kL1Unroll], av[kL1Unroll];
#pragma unroll
                for (int u = 0; u < kL1Unroll; ++u) {
                    const int64_t r = r0 + 8 * u;
                    const bool ok = r < rows;
                    d[u] = ok ? __ldcg(P.dh1 + r * H + j) : 0.f;
                    sv[u] = ok ? reinterpret_cast<const float2*>(P.xs)[r] : make_float2(0.f, 0.f);
                    av[u] = (ok && four) ? reinterpret_cast<const float2*>(P.xa)[r] : make_float2(0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < kL1Unroll; ++u) {
                    g[0] = fmaf(d[u], sv[u].x, g[0]);
                    g[1] = fmaf(d[u], sv[u].y, g[1]);
                    g[2] = fmaf(d[u], av[u].x, g[2]);
                    g[3] = fmaf(d[u], av[u].y, g[3]);
                    gb += d[u];
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) red[i][warp][lane] = g[i];
            red[4][warp][lane] = gb;
            __syncthreads();
            if (warp == 0) {
                float sum[5];
#pragma unroll
                for (int q = 0; q < 5; ++q) {
                    float v = 0.f;
#pragma unroll
                    for (int y = 0; y < 8; ++y) v += red[q][y][lane];
                    sum[q] = v;
                }
                for (int i = 0; i < P.n_in; ++i) P.gW1[j * P.n_in + i] = sum[i];
                P.gb1[j] = sum[4];
            }
        }
    } else if (rows > 0 && P.dxa) {
        float w2[8], w3[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            w2[q] = P.W1[(lane + 32 * q) * 4 + 2];
            w3[q] = P.W1[(lane + 32 * q) * 4 + 3];
        }
        constexpr int RU = 4;   // rows in flight per warp (8 coalesced 128-byte loads each)
        for (int64_t r0 = (int64_t)(blockIdx.x - kL1ColBlocks) * 8 + warp; r0 < rows; r0 += 8 * kL1RowBlocks * RU) {
            float d[RU][8];
#pragma unroll
            for (int u = 0; u < RU; ++u) {
                const int64_t r = r0 + (int64_t)u * 8 * kL1RowBlocks;
#pragma unroll
                for (int q = 0; q < 8; ++q) d[u][q] = r < rows ? __ldcg(P.dh1 + r * H + lane + 32 * q) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < RU; ++u) {
                const int64_t r = r0 + (int64_t)u * 8 * kL1RowBlocks;
                float a0 = 0.f, a1 = 0.f;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    a0 = fmaf(d[u][q], w2[q], a0);
                    a1 = fmaf(d[u][q], w3[q], a1);
                }
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) {
                    a0 += __shfl_xor_sync(0xffffffffu, a0, s);
                    a1 += __shfl_xor_sync(0xffffffffu, a1, s);
                }
                if (lane == 0 && r < rows) reinterpret_cast<float2*>(P.dxa)[r] = make_float2(a0, a1);
            }
        }
    }
    run_tail(A.tail);
}

// (4)-(8) the per-row stages as kernels of their own (SIMT path and the unfused tcgen05 path; the fused path runs the
//     same bodies as tails of the producing kernels, update_tails.cuh)
__global__ void __launch_bounds__(kThreads) sac_loss_kernel(const __grid_constant__ SacLossArgs A) {
    __shared__ float red[4 * 32];
    __shared__ double redd[32];
    sac_loss_body(A, red, redd);
}
__global__ void __launch_bounds__(kThreads) gauss_backward_kernel(const __grid_constant__ GaussBwdArgs A) {
    const int64_t rows = *A.rows_ptr;
    const int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x;
    if (i < rows) gauss_backward_row(A, i, rows);
}
__global__ void __launch_bounds__(kThreads) qrisk_loss_kernel(const __grid_constant__ QrLossArgs A) {
    __shared__ float red[4 * 32];
    qrisk_loss_body(A, red);
}
__global__ void __launch_bounds__(kThreads) recovery_loss_kernel(const __grid_constant__ RecLossArgs A) {
    __shared__ float red[32];
    recovery_loss_body(A, red);
}
__global__ void __launch_bounds__(kThreads) stoch_backward_kernel(const __grid_constant__ StochBwdArgs A) {
    __shared__ float red[32];
    stoch_backward_body(A, red);
}

// (9) optimizer step, fused: Adam (torch.optim.Adam defaults: sac.py:84,114; qrisk.py:58,75) over a flat range,
//     refresh of the k-major W2 images, the soft target update of the net's target copy (utils.py:46-49:
//     target = target*(1-tau) + new_param*tau, sac.py:273-274 / qrisk.py:160-162) and the step bookkeeping
//     (Adam step counts, update counters) done by the last CTA to finish.
struct ImgRef {
    int64_t off;  // offset of a 256x256 tensor inside the arena
    int64_t img;  // offset of its transposed (k-major) image
    int64_t tc;   // offset of its fp16 hi/lo tcgen05 operand image
    int64_t tcT;  // ... and of the image of its transpose
};
__device__ __forceinline__ void refresh_images(float* arena, const ImgRef* img, int n_img, int64_t o, float p) {
    for (int q = 0; q < n_img; ++q) {
        const int64_t d = o - img[q].off;
        if (d >= 0 && d < (int64_t)H * H) {
            arena[img[q].img + (d & (H - 1)) * H + (d >> 8)] = p;
            tc_image_store(reinterpret_cast<__half*>(arena + img[q].tc), (int)(d >> 8), (int)(d & (H - 1)), p);
            tc_image_store(reinterpret_cast<__half*>(arena + img[q].tcT), (int)(d & (H - 1)), (int)(d >> 8), p);
        }
    }
}
struct AdamArgs {
    float* arena;
    int64_t off, count, grad_off, m_off, v_off;
    float lr, b1, b2, eps, grad_scale;
    double lr64;
    int64_t* counters;
    int t_counter, rows_counter;
    ImgRef img[6];
    int n_img;
    // soft target update of [tgt_src_off, tgt_src_off + tgt_count) into tgt_off (tgt_count == 0: none)
    int64_t tgt_src_off, tgt_off, tgt_count;
    float tau;
    int upd_counter, interval;       // polyak only if counters[upd_counter] % interval == 0
    ImgRef timg[2];
    int n_timg;
    // bookkeeping by the last CTA: counters[bump[i]] += 1 (i < n_bump)
    int bump[3];
    int n_bump;
    // multi-GPU peer mode: the gradient is the sum over the ranks' arenas (peer-mapped), in rank order
    const float* peer[8];
    int n_peer;
};
__global__ void __launch_bounds__(kThreads) adam_kernel(const __grid_constant__ AdamArgs A) {
    if (A.counters[A.rows_counter] <= 0) return;
    __shared__ float s_bc[2];
    __shared__ int s_polyak;
    if (threadIdx.x == 0) {  // bias corrections and step size in double (python floats in torch), once per block
        const double tstep = (double)(A.counters[A.t_counter] + 1);
        s_bc[0] = (float)(A.lr64 / (1.0 - pow((double)A.b1, tstep)));
        s_bc[1] = (float)sqrt(1.0 - pow((double)A.b2, tstep));
        s_polyak = A.tgt_count > 0 && (A.interval <= 1 || (A.counters[A.upd_counter] % A.interval) == 0);
    }
    __syncthreads();
    const float step_size = s_bc[0], bc2s = s_bc[1];
    const bool polyak = s_polyak != 0;
    const float omt = (float)(1.0 - (double)A.tau);
    for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < A.count; i += (int64_t)gridDim.x * kThreads) {
        const int64_t o = A.off + i;
        float p = A.arena[o];
        if (A.lr64 > 0.0) {
            float g;
            if (A.n_peer > 0) {
                float gs[8];   // all NVLink peer loads in flight together, then summed in rank order
#pragma unroll
                for (int r = 0; r < 8; ++r) gs[r] = r < A.n_peer ? __ldcv(A.peer[r] + A.grad_off + o) : 0.f;
                g = gs[0];
#pragma unroll
                for (int r = 1; r < 8; ++r) g += gs[r];
                g *= A.grad_scale;
            } else {
                g = A.arena[A.grad_off + o] * A.grad_scale;
            }
            float m = A.arena[A.m_off + o], v = A.arena[A.v_off + o];
            m = m + (g - m) * (1.0f - A.b1);               // exp_avg.lerp_(grad, 1 - beta1)
            v = v * A.b2 + (1.0f - A.b2) * g * g;           // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
            const float denom = sqrtf(v) / bc2s + A.eps;
            p = p - step_size * (m / denom);
            A.arena[A.m_off + o] = m;
            A.arena[A.v_off + o] = v;
            A.arena[o] = p;
            refresh_images(A.arena, A.img, A.n_img, o, p);
        }
        if (polyak) {
            const int64_t d0 = o - A.tgt_src_off;
            if (d0 >= 0 && d0 < A.tgt_count) {
                const int64_t to = A.tgt_off + d0;
                const float tp = A.arena[to] * omt + p * A.tau;
                A.arena[to] = tp;
                refresh_images(A.arena, A.timg, A.n_timg, to, tp);
            }
        }
    }
    // every CTA has read the counters above; the last one to arrive bumps them
    __syncthreads();
    if (threadIdx.x == 0 && A.n_bump > 0) {
        __threadfence();
        const unsigned long long ticket = atomicAdd(reinterpret_cast<unsigned long long*>(A.counters + RRL_C_TICKET), 1ull);
        if (ticket == (unsigned long long)gridDim.x - 1) {
            A.counters[RRL_C_TICKET] = 0;
            for (int i = 0; i < A.n_bump; ++i) A.counters[A.bump[i]] += 1;
        }
    }
}

constexpr int kPadBase = 256;   // first signal-pad word used by our barriers (the first KB is left to torch's own primitives)
// (9') the same optimizer step with COALESCED operand-image refresh (tcgen05 path).  adam_kernel above scatters three
//      image elements per parameter (4-byte stores 1 KB apart for the transposed fp32 image, 2-byte stores for the fp16
//      hi/lo images): ~1.5 M sector writes per SAC step.  Here a CTA owns a 32x32 tile of a hidden matrix: float4 loads
//      of p / g / m / v (peer gradients included), Adam + soft target update in registers, the tile (and the target's
//      tile) transposed through shared memory, then every image is written in 16-byte pieces, contiguous per warp:
//          fp32 W2^T image      thread (k, 4 consecutive n)                      128-byte runs
//          tcgen05 image of W2  thread (n, 8 consecutive k) -> hi + lo uint4     512-byte runs
//          ... and of W2^T      thread (k, 8 consecutive n) -> hi + lo uint4     512-byte runs
//      The remaining tensors (W1, biases, heads: a few thousand floats per net) go through flat CTAs.  Same arithmetic
//      per element as adam_kernel (bit-identical parameters).
struct AdamW2 {
    ImgRef src;      // the stepped hidden matrix and its three images
    ImgRef tgt;      // its target-net counterpart (valid if has_tgt)
    int has_tgt;
};
struct AdamTileArgs {
    float* arena;
    int64_t grad_off, m_off, v_off;
    float b1, b2, eps, grad_scale;
    double lr64;
    int64_t* counters;
    int t_counter, rows_counter;
    AdamW2 w2[4];
    int n_w2;
    int64_t seg_off[24], seg_cnt[24];   // the other tensors of the stepped nets
    int n_seg, n_flat;                  // flat CTAs: one per 256 elements of a segment (a segment is a latency chain)
    unsigned char flat_seg[64], flat_blk[64];
    int64_t tgt_src_off, tgt_off, tgt_count;
    float tau;
    int upd_counter, interval;
    int bump[3];
    int n_bump;
    const float* peer[8];
    int n_peer;
    // fused cross-GPU flag barrier (rrl_peers_t::epoch != 0): signal pads of all ranks, this rank, the generation counter
    uint32_t* signal[8];
    int rank;
    int64_t* epoch;
    const float* mc;   // multicast (NVLS) address of the arena or NULL: gradient sum by multimem.ld_reduce instead of n_peer loads
};
__device__ __forceinline__ float4 multimem_sum4(const float* mc_addr) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc_addr) : "memory");
    return v;
}
__device__ __forceinline__ float multimem_sum1(const float* mc_addr) {
    float v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f32 %0, [%1];" : "=f"(v) : "l"(mc_addr) : "memory");
    return v;
}
__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, float b1, float b2, float eps, float step_size,
                                          float bc2s) {
    m = m + (g - m) * (1.0f - b1);               // exp_avg.lerp_(grad, 1 - beta1)
    v = v * b2 + (1.0f - b2) * g * g;             // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(v) / bc2s + eps;
    p = p - step_size * (m / denom);
}
__device__ __forceinline__ void half_split8(const float* w, uint4* hi, uint4* lo) {   // tc_image_store's split, 8 at once
    __align__(16) __half h[8], l[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const float sv = fmaxf(fminf(w[e] * kTcScaleB, 60000.0f), -60000.0f);
        h[e] = __float2half_rn(sv);
        l[e] = __float2half_rn(sv - __half2float(h[e]));
    }
    *hi = *reinterpret_cast<const uint4*>(h);
    *lo = *reinterpret_cast<const uint4*>(l);
}
// the three images of one 32x32 tile (rows n0.., columns k0..) held in shared memory as tile[n][k]
__device__ __forceinline__ void write_tile_images(float* arena, const ImgRef& R, const float (*tile)[33], int n0, int k0) {
    const int t = threadIdx.x;
    {   // fp32 W2^T image: [k][n]
        const int kk = t >> 3, nq = t & 7;
        const float4 v = make_float4(tile[nq * 4 + 0][kk], tile[nq * 4 + 1][kk], tile[nq * 4 + 2][kk], tile[nq * 4 + 3][kk]);
        *reinterpret_cast<float4*>(arena + R.img + (int64_t)(k0 + kk) * H + n0 + nq * 4) = v;
    }
    constexpr size_t kLo = (size_t)4 * 32 * 64;   // halves between the hi and the lo image of one 32-wide k chunk
    if (t < 128) {   // tcgen05 image of W2: element (n, k)
        const int nl = t & 31, kg = t >> 5;
        float w[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) w[e] = tile[nl][kg * 8 + e];
        uint4 hi, lo;
        half_split8(w, &hi, &lo);
        const int n = n0 + nl, c = k0 >> 5;
        __half* img = reinterpret_cast<__half*>(arena + R.tc);
        const size_t base = ((((size_t)c * 2) * 4 + kg) * 32 + (n >> 3)) * 64 + (n & 7) * 8;
        *reinterpret_cast<uint4*>(img + base) = hi;
        *reinterpret_cast<uint4*>(img + base + kLo) = lo;
    } else {         // tcgen05 image of W2^T: element (n' = k, k' = n)
        const int kl = (t - 128) & 31, ng = (t - 128) >> 5;
        float w[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) w[e] = tile[ng * 8 + e][kl];
        uint4 hi, lo;
        half_split8(w, &hi, &lo);
        const int k = k0 + kl, c = n0 >> 5;
        __half* img = reinterpret_cast<__half*>(arena + R.tcT);
        const size_t base = ((((size_t)c * 2) * 4 + ng) * 32 + (k >> 3)) * 64 + (k & 7) * 8;
        *reinterpret_cast<uint4*>(img + base) = hi;
        *reinterpret_cast<uint4*>(img + base + kLo) = lo;
    }
}
__device__ unsigned long long g_opt_t0;   // diagnostic: start stamp (CTA 0) of the running optimizer-step kernel
// diagnostics of the optimizer-step kernels (rrl_debug_opt_times, include/rrl.h), summed over launches (ns): [0] CTA 0 start ->
// barrier done, [1] -> gradients (own + peers') loaded, [2] -> CTA 0 done, [3] launches, [4] CTA 0's wait for the slowest peer's
// flag, [5] barriers, [6] CTA 0 start -> last CTA end.  Deliberately NOT in the counter block: that one stays deterministic
__device__ unsigned long long g_opt_dbg[8];
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)::"memory");
    return t;
}
__global__ void __launch_bounds__(kThreads) adam_tile_kernel(const __grid_constant__ AdamTileArgs A) {
    pdl_wait();   // programmatic dependent launch (common.cuh): before the first global read, on every path
    if (A.counters[A.rows_counter] <= 0) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        g_opt_t0 = t0;
    }
    __shared__ float s_bc[2];
    __shared__ int s_polyak;
    __shared__ float tile[32][33], ttile[32][33];
    const int t = threadIdx.x;
    // tile CTAs: parameters and moments are loaded BEFORE the block waits for thread 0's double-precision bias corrections (two
    // pow() calls, ~1.5 us) and for the peers' flags: independent of both
    const int n_tile_ctas = A.n_w2 * 64;
    const bool is_tile = (int)blockIdx.x < n_tile_ctas;
    const AdamW2& W = A.w2[is_tile ? (blockIdx.x >> 6) : 0];
    const int tl = blockIdx.x & 63, n0 = (tl >> 3) * 32, k0 = (tl & 7) * 32;
    const int rn = t >> 3, kq = t & 7;
    const int64_t o = W.src.off + (int64_t)(n0 + rn) * H + k0 + kq * 4;
    const int64_t to = W.tgt.off + (int64_t)(n0 + rn) * H + k0 + kq * 4;
    float4 p4 = make_float4(0.f, 0.f, 0.f, 0.f), m4 = p4, v4 = p4, tp4 = p4, g4 = p4;
    if (is_tile) {
        p4 = *reinterpret_cast<const float4*>(A.arena + o);
        m4 = *reinterpret_cast<const float4*>(A.arena + A.m_off + o);
        v4 = *reinterpret_cast<const float4*>(A.arena + A.v_off + o);
        if (W.has_tgt && A.tgt_count > 0) tp4 = *reinterpret_cast<const float4*>(A.arena + to);
        if (A.n_peer == 0) g4 = *reinterpret_cast<const float4*>(A.arena + A.grad_off + o);
    }
    if (t == 0) {  // bias corrections and step size in double (python floats in torch), once per block
        const double tstep = (double)(A.counters[A.t_counter] + 1);
        s_bc[0] = (float)(A.lr64 / (1.0 - pow((double)A.b1, tstep)));
        s_bc[1] = (float)sqrt(1.0 - pow((double)A.b2, tstep));
        s_polyak = A.tgt_count > 0 && (A.interval <= 1 || (A.counters[A.upd_counter] % A.interval) == 0);
    }
    __syncthreads();
    const float step_size = s_bc[0], bc2s = s_bc[1];
    const bool polyak = s_polyak != 0;
    const float omt = (float)(1.0 - (double)A.tau);
    if (A.epoch) {
        // Cross-GPU barrier inside the optimizer step: this rank's gradients were completed by the previous kernels of the
        // stream, so CTA 0 publishes generation g = *epoch + 1 into every peer's pad (release, system scope); EVERY CTA then
        // waits until all peers' generation g has arrived in the local pad before it loads their gradients.  The last CTA
        // stores g (below).  A CTA waits only on OTHER ranks' CTA 0, never on a CTA of its own grid.
        const unsigned g = (unsigned)(*reinterpret_cast<volatile int64_t*>(A.epoch) + 1);
        __shared__ unsigned long long s_wait;
        if (t == 0) s_wait = 0ull;
        __syncthreads();
        if (blockIdx.x == 0 && t < A.n_peer) {
            // the gradients were written by EARLIER kernels of the stream: the grid boundary orders them before this thread, and
            // the release store is cumulative over that order -- no separate fence.sys (it cost 2 us per barrier: 0.3608 ->
            // 0.3538 ms per step on 2 GPUs, profiles/r2/sync_modes.txt)
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(A.signal[t] + kPadBase + A.rank), "r"(g) : "memory");
        }
        if (t < A.n_peer) {
            const uint32_t* src = A.signal[A.rank] + kPadBase + t;
            const bool lost = A.counters[RRL_C_ERROR] == 2;   // a peer was lost earlier: do not stall every later step as well
            unsigned v;
            unsigned long long t0, t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
            do {
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(src) : "memory");
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            } while (!lost && (int)(v - g) < 0 && t1 - t0 < 60000000000ull);   // give up after 60 s instead of hanging the GPU
            if ((int)(v - g) < 0) A.counters[RRL_C_ERROR] = 2;
            if (blockIdx.x == 0) atomicMax(&s_wait, t1 - t0);
        }
        __syncthreads();
        if (blockIdx.x == 0 && t == 0) {   // diagnostic: how long this rank waited for its slowest peer (skew + signal latency)
            g_opt_dbg[4] += s_wait;
            g_opt_dbg[5] += 1;
        }
    }
    const bool dbg = blockIdx.x == 0 && t == 0;
    unsigned long long d1 = 0, d2 = 0;
    if (dbg) d1 = gtime();
    if (is_tile) {
        if (A.mc) {
            g4 = multimem_sum4(A.mc + A.grad_off + o);   // reduced inside the NVSwitch
        } else if (A.n_peer > 0) {
            float4 gs[8];   // all NVLink peer loads in flight together, then summed in rank order
#pragma unroll
            for (int r = 0; r < 8; ++r)
                gs[r] = r < A.n_peer ? __ldcv(reinterpret_cast<const float4*>(A.peer[r] + A.grad_off + o)) : make_float4(0.f, 0.f, 0.f, 0.f);
            g4 = gs[0];
#pragma unroll
            for (int r = 1; r < 8; ++r) { g4.x += gs[r].x; g4.y += gs[r].y; g4.z += gs[r].z; g4.w += gs[r].w; }
        }
        if (dbg) { d2 = gtime(); if (g4.x == 1.2345678e33f) d2 += 1; }   // (the compare pins the stamp behind the loads)
        const bool do_tgt = polyak && W.has_tgt;
        adam_elem(p4.x, g4.x * A.grad_scale, m4.x, v4.x, A.b1, A.b2, A.eps, step_size, bc2s);
        adam_elem(p4.y, g4.y * A.grad_scale, m4.y, v4.y, A.b1, A.b2, A.eps, step_size, bc2s);
        adam_elem(p4.z, g4.z * A.grad_scale, m4.z, v4.z, A.b1, A.b2, A.eps, step_size, bc2s);
        adam_elem(p4.w, g4.w * A.grad_scale, m4.w, v4.w, A.b1, A.b2, A.eps, step_size, bc2s);
        *reinterpret_cast<float4*>(A.arena + o) = p4;
        *reinterpret_cast<float4*>(A.arena + A.m_off + o) = m4;
        *reinterpret_cast<float4*>(A.arena + A.v_off + o) = v4;
        tile[rn][kq * 4 + 0] = p4.x; tile[rn][kq * 4 + 1] = p4.y; tile[rn][kq * 4 + 2] = p4.z; tile[rn][kq * 4 + 3] = p4.w;
        if (do_tgt) {
            tp4.x = tp4.x * omt + p4.x * A.tau; tp4.y = tp4.y * omt + p4.y * A.tau;
            tp4.z = tp4.z * omt + p4.z * A.tau; tp4.w = tp4.w * omt + p4.w * A.tau;
            *reinterpret_cast<float4*>(A.arena + to) = tp4;
            ttile[rn][kq * 4 + 0] = tp4.x; ttile[rn][kq * 4 + 1] = tp4.y; ttile[rn][kq * 4 + 2] = tp4.z; ttile[rn][kq * 4 + 3] = tp4.w;
        }
        __syncthreads();
        write_tile_images(A.arena, W.src, tile, n0, k0);
        if (do_tgt) write_tile_images(A.arena, W.tgt, ttile, n0, k0);
    } else {
        const int fb = blockIdx.x - n_tile_ctas;
        const int sg = A.flat_seg[fb];
        const int64_t i = (int64_t)A.flat_blk[fb] * kThreads + t;
        if (i < A.seg_cnt[sg]) {
            const int64_t o = A.seg_off[sg] + i;
            float p = A.arena[o];
            float g;
            if (A.mc) {
                g = multimem_sum1(A.mc + A.grad_off + o);
            } else if (A.n_peer > 0) {
                float gs[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) gs[r] = r < A.n_peer ? __ldcv(A.peer[r] + A.grad_off + o) : 0.f;
                g = gs[0];
#pragma unroll
                for (int r = 1; r < 8; ++r) g += gs[r];
            } else {
                g = A.arena[A.grad_off + o];
            }
            float m = A.arena[A.m_off + o], v = A.arena[A.v_off + o];
            const int64_t d0 = o - A.tgt_src_off;
            const bool tg = polyak && d0 >= 0 && d0 < A.tgt_count;
            const float tp = tg ? A.arena[A.tgt_off + d0] : 0.f;
            adam_elem(p, g * A.grad_scale, m, v, A.b1, A.b2, A.eps, step_size, bc2s);
            A.arena[A.m_off + o] = m;
            A.arena[A.v_off + o] = v;
            A.arena[o] = p;
            if (tg) A.arena[A.tgt_off + d0] = tp * omt + p * A.tau;
        }
    }
    // every CTA has read the counters above; the last one to arrive bumps them
    __syncthreads();
    if (dbg) {
        const unsigned long long d3 = gtime();
        const unsigned long long t0 = *reinterpret_cast<volatile unsigned long long*>(&g_opt_t0);
        g_opt_dbg[0] += d1 - t0; g_opt_dbg[1] += d2 - d1; g_opt_dbg[2] += d3 - d2; g_opt_dbg[3] += 1;
    }
    if (t == 0 && A.n_bump > 0) {
        __threadfence();
        const unsigned long long ticket = atomicAdd(reinterpret_cast<unsigned long long*>(A.counters + RRL_C_TICKET), 1ull);
        if (ticket == (unsigned long long)gridDim.x - 1) {
            A.counters[RRL_C_TICKET] = 0;
            for (int i = 0; i < A.n_bump; ++i) A.counters[A.bump[i]] += 1;
            if (A.epoch) *A.epoch += 1;       // every CTA has read the old generation (its ticket came after its wait)
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            const unsigned long long t0 = *reinterpret_cast<volatile unsigned long long*>(&g_opt_t0);
            if (t1 > t0) g_opt_dbg[6] += t1 - t0;
        }
    }
}

// (9b) Adam on the scalar multipliers (sac.py:241-271): log_alpha is a float32 tensor with lr; log_nu and
//      log_lambda_RCPO are float64 tensors (np.log of a python float) with lr 0.1*lr.  One thread.
struct ScalarAdamArgs {
    float* scal;
    int64_t* counters;
    double lr, b1, b2, eps, grad_scale;
    int flags, rows_counter;
};
template <typename T>
__device__ void adam_scalar(T& p, T g, T& m, T& v, double lr, double b1, double b2, double eps, double tstep) {
    m = m + (g - m) * (T)(1.0 - b1);
    v = v * (T)b2 + (T)(1.0 - b2) * g * g;
    const double bc1 = 1.0 - pow(b1, tstep), bc2s = sqrt(1.0 - pow(b2, tstep));
    const T denom = (T)sqrt((double)v) / (T)bc2s + (T)eps;
    p = p - (T)(lr / bc1) * (m / denom);
}
__global__ void scalar_adam_kernel(const ScalarAdamArgs A) {
    if (threadIdx.x != 0 || blockIdx.x != 0 || A.counters[A.rows_counter] <= 0) return;
    double* sd = reinterpret_cast<double*>(A.scal + RRL_S_F64_BASE);
    if (A.flags & RRL_ALGO_AUTO_ALPHA) {
        const double t = (double)(A.counters[RRL_C_ADAM_T_ALPHA] + 1);
        adam_scalar<float>(A.scal[RRL_S_LOG_ALPHA], A.scal[RRL_S_G_LOG_ALPHA] * (float)A.grad_scale, A.scal[RRL_S_M_ALPHA],
                           A.scal[RRL_S_V_ALPHA], A.lr, A.b1, A.b2, A.eps, t);
        A.scal[RRL_S_ALPHA] = expf(A.scal[RRL_S_LOG_ALPHA]);  // self.alpha = self.log_alpha.exp()
        A.counters[RRL_C_ADAM_T_ALPHA] += 1;
    }
    if (A.flags & RRL_ALGO_UPDATE_NU) {
        const double t = (double)(A.counters[RRL_C_ADAM_T_NU] + 1);
        adam_scalar<double>(sd[RRL_D_LOG_NU], sd[RRL_D_G_LOG_NU] * A.grad_scale, sd[RRL_D_M_NU], sd[RRL_D_V_NU], 0.1 * A.lr,
                            A.b1, A.b2, A.eps, t);
        sd[RRL_D_NU_LEARNED] = exp(sd[RRL_D_LOG_NU]);
        A.counters[RRL_C_ADAM_T_NU] += 1;
    }
    if (A.flags & RRL_ALGO_RCPO) {
        const double t = (double)(A.counters[RRL_C_ADAM_T_LAMBDA] + 1);
        adam_scalar<double>(sd[RRL_D_LOG_LAMBDA], sd[RRL_D_G_LOG_LAMBDA] * A.grad_scale, sd[RRL_D_M_LAMBDA], sd[RRL_D_V_LAMBDA],
                            0.1 * A.lr, A.b1, A.b2, A.eps, t);
        sd[RRL_D_LAMBDA] = exp(sd[RRL_D_LOG_LAMBDA]);
        A.counters[RRL_C_ADAM_T_LAMBDA] += 1;
    }
}
__global__ void init_scalars_kernel(float* scal, float alpha, double nu, double lambda) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    for (int i = 0; i < 32; ++i) scal[i] = 0.f;
    double* sd = reinterpret_cast<double*>(scal + RRL_S_F64_BASE);
    scal[RRL_S_ALPHA] = alpha;
    scal[RRL_S_NU_ARG] = (float)nu;
    scal[RRL_S_LOG_ALPHA] = 0.f;  // torch.zeros(1) (sac.py:99-101)
    sd[RRL_D_LOG_NU] = log(nu);
    sd[RRL_D_LOG_LAMBDA] = log(lambda);
    sd[RRL_D_LAMBDA] = lambda;
    sd[RRL_D_NU_LEARNED] = nu;
}

// (9c) cross-GPU barrier over peer-mapped signal pads (one thread per peer): publish this rank's generation into
//      every peer's pad (release, system scope), then wait until every peer's generation has arrived in ours.
struct PeerBarrierArgs {
    uint32_t* signal[8];
    int world, rank;
    int64_t* epoch;
    int64_t* counters;
    int exchange;   // also exchange the Q_risk gate counts: EXT_VIOLS = sum over the OTHER ranks of (num_viols + offline_viols)
    int gate_batch;             // batch size and pos_fraction of the online gate (experiment.py:410); 0: never stop exchanging
    double gate_pos_fraction;
};
__global__ void peer_barrier_kernel(const PeerBarrierArgs A) {
    // gate exchange: once the violation count has passed the gate's threshold on every rank (the same total everywhere, so
    // the same step everywhere) it stays passed -- counts only grow -- and the exchange has nothing left to decide
    if (A.exchange && A.counters[RRL_C_GATE_SATISFIED]) return;
    __shared__ unsigned s_epoch;
    if (threadIdx.x == 0) {
        s_epoch = (unsigned)(++(*A.epoch));   // (this rank's gradient writes come from earlier kernels of the stream: ordered
                                              //  before the release stores below by the grid boundary, no separate fence.sys)
    }
    __syncthreads();
    const unsigned epoch = s_epoch;
    const int r = threadIdx.x;
    long long theirs = 0;
    if (r < A.world) {
        uint32_t* dst = A.signal[r] + kPadBase + A.rank;
        if (A.exchange) {   // value first, then the flag (release): slot [rank] of every peer's pad
            const long long mine = A.counters[RRL_C_NUM_VIOLS] + A.counters[RRL_C_OFFLINE_VIOLS];
            long long* vdst = reinterpret_cast<long long*>(A.signal[r] + kPadBase + 16) + A.rank;
            asm volatile("st.relaxed.sys.global.s64 [%0], %1;" ::"l"(vdst), "l"(mine) : "memory");
        }
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(dst), "r"(epoch) : "memory");
        const uint32_t* src = A.signal[A.rank] + kPadBase + r;
        const bool lost = A.counters[RRL_C_ERROR] == 2;   // a peer was lost earlier: do not stall every later step as well
        unsigned v;
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(src) : "memory");
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        } while (!lost && (int)(v - epoch) < 0 && t1 - t0 < 60000000000ull);   // give up after 60 s instead of hanging the GPU
        if ((int)(v - epoch) < 0) A.counters[RRL_C_ERROR] = 2;
        if (A.exchange && r != A.rank) {
            const long long* vsrc = reinterpret_cast<const long long*>(A.signal[A.rank] + kPadBase + 16) + r;
            asm volatile("ld.relaxed.sys.global.s64 %0, [%1];" : "=l"(theirs) : "l"(vsrc) : "memory");
        }
    }
    if (A.exchange) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) theirs += __shfl_xor_sync(0xffffffffu, theirs, o);
        if (threadIdx.x == 0) {
            A.counters[RRL_C_EXT_VIOLS] = theirs;
            const long long total = theirs + A.counters[RRL_C_NUM_VIOLS] + A.counters[RRL_C_OFFLINE_VIOLS];
            if (A.gate_batch > 0 && (double)total / (double)A.gate_batch > A.gate_pos_fraction && A.counters[RRL_C_ERROR] == 0)
                A.counters[RRL_C_GATE_SATISFIED] = 1;
        }
    }
}

// (10) soft_update (utils.py:46-49): target = target*(1-tau) + source*tau over a whole net (+ image refresh)
struct PolyakArgs {
    float* arena;
    int64_t dst_off, src_off, count;
    float tau;
    const int64_t* counters;
    int rows_counter, upd_counter, interval;  // rows_counter < 0: unconditional
    ImgRef img[2];                            // of the destination net
    int n_img;
};
__global__ void __launch_bounds__(kThreads) polyak_kernel(const __grid_constant__ PolyakArgs A) {
    if (A.rows_counter >= 0) {
        if (A.counters[A.rows_counter] <= 0) return;
        if (A.interval > 1 && (A.counters[A.upd_counter] % A.interval) != 0) return;
    }
    const float omt = (float)(1.0 - (double)A.tau);
    for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < A.count; i += (int64_t)gridDim.x * kThreads) {
        const float s = A.arena[A.src_off + i];
        const float p = A.tau >= 1.0f ? s : (A.arena[A.dst_off + i] * omt + s * A.tau);
        A.arena[A.dst_off + i] = p;
        refresh_images(A.arena, A.img, A.n_img, A.dst_off + i, p);
    }
}

// (11) rebuild every W2T image from the parameters (after the host wrote parameters)
struct RefreshArgs {
    float* arena;
    ImgRef img[kNumImages];
};
__global__ void __launch_bounds__(kThreads) refresh_images_kernel(const __grid_constant__ RefreshArgs A) {
    __shared__ float tile[32][33];
    const ImgRef R = A.img[blockIdx.z];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) tile[r][tx] = A.arena[R.off + (int64_t)(by + r) * H + bx + tx];
    __syncthreads();
    for (int r = ty; r < 32; r += 8) A.arena[R.img + (int64_t)(bx + r) * H + by + tx] = tile[tx][r];
}

// (12) bookkeeping after an apply: Adam step counts, update counters (single thread)
__global__ void bump_kernel(int64_t* counters, int rows_counter, int t0, int t1, int upd_counter) {
    if (threadIdx.x == 0 && blockIdx.x == 0 && counters[rows_counter] > 0) {
        if (t0 >= 0) counters[t0] += 1;
        if (t1 >= 0) counters[t1] += 1;
        if (upd_counter >= 0) counters[upd_counter] += 1;
    }
}

// ---------------------------------------------------------------------------------------------
// host helpers
// ---------------------------------------------------------------------------------------------
int check_cfg(const rrl_agent_config_t* cfg) {
    if (!cfg) { rrl_set_error("null config"); return -2; }
    if (cfg->hidden != H) { rrl_set_error("kernels are specialised for hidden_size 256 (got %d)", cfg->hidden); return -2; }
    if (cfg->max_batch < 32 || cfg->max_batch % 32 != 0 || cfg->max_batch > 8192) {
        rrl_set_error("max_batch must be a multiple of 32 in [32, 8192] (got %d)", cfg->max_batch);
        return -2;
    }
    return 0;
}
#define CHECK_CFG(cfg)                      \
    do {                                    \
        int rc_ = check_cfg(cfg);           \
        if (rc_) return rc_;                \
    } while (0)

ActionSpace action_space(const rrl_agent_config_t* cfg) {
    ActionSpace sp;
    sp.scale[0] = cfg->action_scale[0]; sp.scale[1] = cfg->action_scale[1];
    sp.bias[0] = cfg->action_bias[0]; sp.bias[1] = cfg->action_bias[1];
    return sp;
}

template <int BM>
int launch_forward(const FwdArgs& A, int64_t max_rows, cudaStream_t st) {
    if (A.use_tc) return fwd_tc_launch(A, max_rows, st);
    static bool configured = false;
    const size_t smem = sizeof(FwdSmem<BM>);
    if (!configured) {
        RRL_CUDA(cudaFuncSetAttribute(mlp_forward_kernel<BM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    dim3 grid((unsigned)((max_rows + BM - 1) / BM), (unsigned)A.n_pass);
    mlp_forward_kernel<BM><<<grid, kThreads, smem, st>>>(A);
    RRL_CHECK_LAUNCH();
    return 0;
}

int imgs_of_net(const Layout& L, int net, ImgRef* out) {
    const int heads = (net == RRL_NET_POLICY || net == RRL_NET_RECOVERY) ? 1 : 2;
    for (int h = 0; h < heads; ++h)
        out[h] = ImgRef{L.t_off[net][w2_tensor(net, h)], L.img_off[image_index(net, h)], L.tc_img_off[image_index(net, h)],
                        L.tc_imgT_off[image_index(net, h)]};
    return heads;
}

int launch_adam_tiled(const rrl_agent_config_t* cfg, const Layout& L, float* arena, int64_t* counters, int net_a, int net_b,
                      int t_counter, int rows_counter, cudaStream_t st, int polyak_dst, int polyak_src, float tau,
                      int upd_counter, const int* bump, int n_bump, const rrl_peers_t* peers);
// nets a (and b, contiguous after a in storage order) share one launch.  net_a < 0 with polyak_src >= 0: no Adam
// (lr64 = 0), only the soft update of polyak_src into polyak_dst and the bookkeeping.
int launch_adam(const rrl_agent_config_t* cfg, const Layout& L, float* arena, int64_t* counters, int net_a, int net_b,
                int t_counter, int rows_counter, cudaStream_t st, int polyak_dst = -1, int polyak_src = -1, float tau = 0.f,
                int upd_counter = -1, const int* bump = nullptr, int n_bump = 0, const rrl_peers_t* peers = nullptr) {
    if (cfg->use_tensor_cores && net_a >= 0)   // tcgen05 path: coalesced image refresh (bit-identical parameters)
        return launch_adam_tiled(cfg, L, arena, counters, net_a, net_b, t_counter >= 0 ? t_counter : RRL_C_ADAM_T0, rows_counter, st,
                                 polyak_dst, polyak_src, tau, upd_counter, bump, n_bump, peers);
    AdamArgs A;
    memset(&A, 0, sizeof(A));
    if (peers) {
        A.n_peer = peers->world;
        for (int r = 0; r < peers->world && r < 8; ++r) A.peer[r] = reinterpret_cast<const float*>(peers->arena[r]);
    }
    A.arena = arena;
    const int first = net_a >= 0 ? net_a : polyak_src;
    A.off = L.net_off[first];
    A.count = L.net_size[first] + ((net_a >= 0 && net_b >= 0) ? L.net_size[net_b] : 0);
    A.grad_off = L.grad_off; A.m_off = L.m_off; A.v_off = L.v_off;
    A.lr = cfg->lr; A.b1 = cfg->beta1; A.b2 = cfg->beta2; A.eps = cfg->adam_eps;
    A.lr64 = net_a >= 0 ? (cfg->lr64 > 0.0 ? cfg->lr64 : (double)cfg->lr) : 0.0;
    A.grad_scale = cfg->grad_scale;
    A.counters = counters;
    A.t_counter = t_counter >= 0 ? t_counter : RRL_C_ADAM_T0;
    A.rows_counter = rows_counter;
    if (net_a >= 0) {
        A.n_img = imgs_of_net(L, net_a, A.img);
        if (net_b >= 0) A.n_img += imgs_of_net(L, net_b, A.img + A.n_img);
    }
    if (polyak_dst >= 0) {
        A.tgt_src_off = L.net_off[polyak_src]; A.tgt_off = L.net_off[polyak_dst]; A.tgt_count = L.net_size[polyak_dst];
        A.tau = tau; A.upd_counter = upd_counter; A.interval = cfg->target_update_interval;
        A.n_timg = imgs_of_net(L, polyak_dst, A.timg);
    }
    for (int i = 0; i < n_bump && i < 3; ++i) A.bump[i] = bump[i];
    A.n_bump = n_bump < 3 ? n_bump : 3;
    const int blocks = (int)((A.count + kThreads * 4 - 1) / (kThreads * 4));
    adam_kernel<<<blocks, kThreads, 0, st>>>(A);
    RRL_CHECK_LAUNCH();
    return 0;
}
// the tiled variant (adam_tile_kernel): same arguments; nets a (and b) are stepped, polyak_src -> polyak_dst soft-updated
int launch_adam_tiled(const rrl_agent_config_t* cfg, const Layout& L, float* arena, int64_t* counters, int net_a, int net_b,
                      int t_counter, int rows_counter, cudaStream_t st, int polyak_dst, int polyak_src, float tau,
                      int upd_counter, const int* bump, int n_bump, const rrl_peers_t* peers) {
    AdamTileArgs A;
    memset(&A, 0, sizeof(A));
    if (peers) {
        A.n_peer = peers->world;
        for (int r = 0; r < peers->world && r < 8; ++r) {
            A.peer[r] = reinterpret_cast<const float*>(peers->arena[r]);
            A.signal[r] = reinterpret_cast<uint32_t*>(peers->signal[r]);
        }
        A.rank = peers->rank;
        A.epoch = reinterpret_cast<int64_t*>(peers->epoch);
        A.mc = (peers->epoch && peers->mc_arena) ? reinterpret_cast<const float*>(peers->mc_arena) : nullptr;
    }
    A.arena = arena;
    A.grad_off = L.grad_off; A.m_off = L.m_off; A.v_off = L.v_off;
    A.b1 = cfg->beta1; A.b2 = cfg->beta2; A.eps = cfg->adam_eps; A.grad_scale = cfg->grad_scale;
    A.lr64 = cfg->lr64 > 0.0 ? cfg->lr64 : (double)cfg->lr;
    A.counters = counters;
    A.t_counter = t_counter; A.rows_counter = rows_counter;
    const int nets[2] = {net_a, net_b};
    for (int ni = 0; ni < 2; ++ni) {
        const int net = nets[ni];
        if (net < 0) continue;
        ImgRef src[2], tgt[2];
        const int heads = imgs_of_net(L, net, src);
        const bool has_tgt = polyak_dst >= 0 && polyak_src == net;
        if (has_tgt) imgs_of_net(L, polyak_dst, tgt);
        for (int h = 0; h < heads; ++h) {
            AdamW2& W = A.w2[A.n_w2++];
            W.src = src[h];
            W.has_tgt = has_tgt ? 1 : 0;
            if (has_tgt) W.tgt = tgt[h];
        }
        for (int t = 0; t < L.n_tensors[net]; ++t) {
            bool is_w2 = false;
            for (int h = 0; h < heads; ++h) is_w2 = is_w2 || (t == w2_tensor(net, h));
            if (is_w2) continue;
            const TDesc d = L.t_desc[net][t];
            if (A.n_seg >= 24) { rrl_set_error("launch_adam_tiled: too many tensors"); return -2; }
            A.seg_off[A.n_seg] = L.t_off[net][t];
            A.seg_cnt[A.n_seg] = (int64_t)d.rows * (d.cols ? d.cols : 1);
            ++A.n_seg;
        }
    }
    if (polyak_dst >= 0) {
        A.tgt_src_off = L.net_off[polyak_src]; A.tgt_off = L.net_off[polyak_dst]; A.tgt_count = L.net_size[polyak_dst];
        A.tau = tau; A.upd_counter = upd_counter; A.interval = cfg->target_update_interval;
    }
    for (int i = 0; i < n_bump && i < 3; ++i) A.bump[i] = bump[i];
    A.n_bump = n_bump < 3 ? n_bump : 3;
    for (int sg = 0; sg < A.n_seg; ++sg)
        for (int64_t b = 0; b * kThreads < A.seg_cnt[sg]; ++b) {
            if (A.n_flat >= 64) { rrl_set_error("launch_adam_tiled: too many flat blocks"); return -2; }
            A.flat_seg[A.n_flat] = (unsigned char)sg;
            A.flat_blk[A.n_flat] = (unsigned char)b;
            ++A.n_flat;
        }
    RRL_CUDA(rrl_launch_pdl(adam_tile_kernel, dim3(A.n_w2 * 64 + A.n_flat), dim3(kThreads), 0, st, A));
    return 0;
}
int launch_polyak(const Layout& L, float* arena, const int64_t* counters, int dst, int src, float tau, int rows_counter,
                  int upd_counter, int interval, cudaStream_t st) {
    PolyakArgs A;
    memset(&A, 0, sizeof(A));
    A.arena = arena;
    A.dst_off = L.net_off[dst]; A.src_off = L.net_off[src]; A.count = L.net_size[dst];
    A.tau = tau;
    A.counters = counters;
    A.rows_counter = rows_counter; A.upd_counter = upd_counter; A.interval = interval;
    A.n_img = imgs_of_net(L, dst, A.img);
    const int blocks = (int)((A.count + kThreads * 4 - 1) / (kThreads * 4));
    polyak_kernel<<<blocks, kThreads, 0, st>>>(A);
    RRL_CHECK_LAUNCH();
    return 0;
}

int launch_gemm(GemmArgs& G, int n_pass, int mt, int64_t max_rows, int use_tc, cudaStream_t st) {
    G.n_pass = n_pass;
    G.use_tc = use_tc;
    if (use_tc) return bwd_tc_launch(G, max_rows, st);
    gemm_stream_kernel<<<dim3(mt, n_pass), kThreads, 0, st>>>(G);
    RRL_CHECK_LAUNCH();
    return 0;
}

inline uint32_t* slot_bits(const Layout& L, float* arena, int slot) { return reinterpret_cast<uint32_t*>(arena + L.dh2t[slot]); }
inline uint32_t* slot_bits1(const Layout& L, float* arena, int slot) { return slot_bits(L, arena, slot) + L.R * (H / 32); }

// layer 1 of the pass + its inputs + the sign bits of h2 (tcgen05 backward: h1 / relu' are recomputed, never loaded)
inline void bwd_inputs(GemmPass& p, const HeadW& w, const float* xs, const float* xa, const Layout& L, float* arena, int slot) {
    p.W1 = w.W1; p.b1 = w.b1; p.n_in = w.n_in; p.xs = xs; p.xa = xa; p.h2bits = slot_bits(L, arena, slot);
    p.h1bits = slot_bits1(L, arena, slot);
}
// fused layer-1 backward of a DATA pass (tcgen05 path, use_tensor_cores 2): weight gradients (gW1, gb1) or d(action) (dxa);
// dh1 is then not stored.  The slot's dh2 region holds the cross-tile ticket (word 0) and the per-tile partial sums.
inline void fuse_l1(GemmPass& p, float* gW1, float* gb1, float* dxa, const Layout& L, float* arena, int slot) {
    p.gW1 = gW1; p.gb1 = gb1; p.dxa = dxa; p.C = nullptr;
    p.l1ticket = reinterpret_cast<int*>(arena + L.dh2[slot]);
    p.l1part = arena + L.dh2[slot] + 64;
}

// slot >= 0: activations are kept for the backward pass; weight_pass: the backward also has a WEIGHT pass for it
FwdPass q_pass(const Layout& L, float* arena, int net, int head, const float* xs, const float* xa, int slot, float* out_q,
               bool weight_pass = false) {
    FwdPass p;
    memset(&p, 0, sizeof(p));
    p.w = head_w(L, arena, net, head);
    p.tc_img = tc_img_of(L, arena, net, head);
    p.head = (net == RRL_NET_QRISK || net == RRL_NET_QRISK_TARGET) ? HEAD_QRISK : HEAD_Q;
    p.xs = xs; p.xa = xa;
    if (slot >= 0) {
        p.h1 = arena + L.h1[slot]; p.h2 = arena + L.h2[slot]; p.h2bits = slot_bits(L, arena, slot); p.h1bits = slot_bits1(L, arena, slot);
    }
    p.keep_h2 = weight_pass ? 1 : 0;
    p.out_q = out_q;
    return p;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" int64_t rrl_agent_arena_floats(const rrl_agent_config_t* cfg) {
    if (check_cfg(cfg)) return -1;
    return make_layout(cfg).total;
}

extern "C" int rrl_agent_num_tensors(int net) {
    TDesc d[kMaxTensors];
    if (net < 0 || net >= RRL_NUM_NETS) return -1;
    return net_tensors(net, d);
}

extern "C" int rrl_agent_tensor_info(const rrl_agent_config_t* cfg, int net, int tensor, int64_t* offset, int64_t* rows,
                                     int64_t* cols) {
    CHECK_CFG(cfg);
    RRL_CHECK_ARG(net >= 0 && net < RRL_NUM_NETS, "bad net id");
    const Layout L = make_layout(cfg);
    RRL_CHECK_ARG(tensor >= 0 && tensor < L.n_tensors[net], "bad tensor index");
    if (offset) *offset = L.t_off[net][tensor];
    if (rows) *rows = L.t_desc[net][tensor].rows;
    if (cols) *cols = L.t_desc[net][tensor].cols;
    return 0;
}

extern "C" int rrl_agent_grad_range(const rrl_agent_config_t* cfg, int net, int64_t* offset, int64_t* count) {
    CHECK_CFG(cfg);
    const Layout L = make_layout(cfg);
    if (net < 0) {  // whole block
        if (offset) *offset = L.grad_off;
        if (count) *count = L.train_floats;
        return 0;
    }
    RRL_CHECK_ARG(net == RRL_NET_CRITIC || net == RRL_NET_POLICY || net == RRL_NET_QRISK || net == RRL_NET_RECOVERY,
                  "net has no gradients");
    if (offset) *offset = L.grad_off + L.net_off[net];
    if (count) *count = L.net_size[net];
    return 0;
}

extern "C" int rrl_agent_scratch_info(const rrl_agent_config_t* cfg, const char* name, int64_t* offset, int64_t* count) {
    CHECK_CFG(cfg);
    RRL_CHECK_ARG(name, "null name");
    const Layout L = make_layout(cfg);
    for (int i = 0; i < L.n_names; ++i)
        if (strcmp(L.names[i].name, name) == 0) {
            if (offset) *offset = L.names[i].off;
            if (count) *count = L.names[i].count;
            return 0;
        }
    if (strcmp(name, "adam_m") == 0 || strcmp(name, "adam_v") == 0) {
        if (offset) *offset = name[5] == 'm' ? L.m_off : L.v_off;
        if (count) *count = L.train_floats;
        return 0;
    }
    rrl_set_error("rrl_agent_scratch_info: unknown region '%s'", name);
    return -2;
}

extern "C" int rrl_agent_init_scalars(const rrl_agent_config_t* cfg, float* arena, void* stream) {
    CHECK_CFG(cfg);
    RRL_CHECK_ARG(arena, "null arena");
    RRL_CHECK_ARG(cfg->nu > 0.0 && cfg->lambda_rcpo > 0.0, "nu and lambda_rcpo must be positive (their logs are the parameters)");
    const Layout L = make_layout(cfg);
    const float alpha = (cfg->algo_flags & RRL_ALGO_DETERMINISTIC) ? 0.f : cfg->alpha;  // sac.py:116
    init_scalars_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(arena + L.scalars, alpha, cfg->nu, cfg->lambda_rcpo);
    RRL_CHECK_LAUNCH();
    return 0;
}

extern "C" int rrl_agent_refresh(const rrl_agent_config_t* cfg, float* arena, void* stream) {
    CHECK_CFG(cfg);
    RRL_CHECK_ARG(arena, "null arena");
    const Layout L = make_layout(cfg);
    RefreshArgs A;
    A.arena = arena;
    int n = 0;
    static const int nets[6] = {RRL_NET_CRITIC, RRL_NET_CRITIC_TARGET, RRL_NET_POLICY, RRL_NET_QRISK, RRL_NET_QRISK_TARGET,
                                RRL_NET_RECOVERY};
    for (int i = 0; i < 6; ++i) n += imgs_of_net(L, nets[i], A.img + n);
    refresh_images_kernel<<<dim3(H / 32, H / 32, kNumImages), kThreads, 0, (cudaStream_t)stream>>>(A);
    RRL_CHECK_LAUNCH();
    return tc_images_launch(arena, L, (cudaStream_t)stream);
}

extern "C" int rrl_agent_tc_refresh(const rrl_agent_config_t* cfg, float* arena, void* stream) {
    CHECK_CFG(cfg);
    RRL_CHECK_ARG(arena, "null arena");
    const Layout L = make_layout(cfg);
    return tc_images_launch(arena, L, (cudaStream_t)stream);
}

extern "C" int rrl_hard_update(const rrl_agent_config_t* cfg, float* arena, int dst_net, int src_net, void* stream) {
    CHECK_CFG(cfg);
    RRL_CHECK_ARG(arena, "null arena");
    RRL_CHECK_ARG((dst_net == RRL_NET_CRITIC_TARGET && src_net == RRL_NET_CRITIC) ||
                      (dst_net == RRL_NET_QRISK_TARGET && src_net == RRL_NET_QRISK),
                  "hard_update: dst must be the target of src");
    const Layout L = make_layout(cfg);
    return launch_polyak(L, arena, nullptr, dst_net, src_net, 1.0f, -1, -1, 1, (cudaStream_t)stream);
}

// utils.py:46-49 as a stand-alone call (the update kernels fuse it into their optimizer step; this is the reference's
// `soft_update(target, source, tau)` helper for callers that use it directly)
extern "C" int rrl_soft_update(const rrl_agent_config_t* cfg, float* arena, int dst_net, int src_net, float tau, void* stream) {
    CHECK_CFG(cfg);
    RRL_CHECK_ARG(arena, "null arena");
    RRL_CHECK_ARG((dst_net == RRL_NET_CRITIC_TARGET && src_net == RRL_NET_CRITIC) ||
                      (dst_net == RRL_NET_QRISK_TARGET && src_net == RRL_NET_QRISK),
                  "soft_update: dst must be the target of src");
    RRL_CHECK_ARG(tau >= 0.f && tau <= 1.f, "soft_update: tau must be in [0, 1]");
    const Layout L = make_layout(cfg);
    return launch_polyak(L, arena, nullptr, dst_net, src_net, tau, -1, -1, 1, (cudaStream_t)stream);
}

extern "C" int rrl_agent_act(const rrl_agent_config_t* cfg, float* arena, int64_t n, const double* state,
                             const float* eps_task, const float* eps_rec, const float* rand_u, int use_recovery,
                             int eval, int64_t start_steps, uint64_t seed, int32_t stream_id, const int64_t* counters,
                             float* action_task, float* action_real, uint8_t* recovery, float* qrisk_out,
                             void* stream) {
    return rrl_agent_act_stage(cfg, arena, n, state, eps_task, eps_rec, rand_u, use_recovery, eval, start_steps, seed, stream_id,
                               counters, action_task, action_real, recovery, qrisk_out, RRL_ACT_STAGE_ALL, 0, stream);
}

extern "C" int rrl_agent_act_stage(const rrl_agent_config_t* cfg, float* arena, int64_t n, const double* state,
                                   const float* eps_task, const float* eps_rec, const float* rand_u, int use_recovery,
                                   int eval, int64_t start_steps, uint64_t seed, int32_t stream_id, const int64_t* counters,
                                   float* action_task, float* action_real, uint8_t* recovery, float* qrisk_out, int stages,
                                   int max_ctas, void* stream) {
    CHECK_CFG(cfg);
    RRL_CHECK_ARG(arena && state && action_task && action_real, "null argument");
    RRL_CHECK_ARG(n > 0, "n must be positive");
    const Layout L = make_layout(cfg);
    ActArgs A;
    memset(&A, 0, sizeof(A));
    A.pol = head_w(L, arena, RRL_NET_POLICY, 0);
    A.qr1 = head_w(L, arena, RRL_NET_QRISK, 0);
    A.qr2 = head_w(L, arena, RRL_NET_QRISK, 1);
    A.rec = head_w(L, arena, RRL_NET_RECOVERY, 0);
    A.n = n; A.state = state;
    A.eps_task = eps_task; A.eps_rec = eps_rec; A.rand_u = rand_u;
    A.use_recovery = use_recovery; A.eval = eval; A.start_steps = start_steps;
    A.seed = seed; A.stream_id = (uint32_t)stream_id; A.counters = counters;
    A.eps_safe = cfg->eps_safe;
    A.det = (cfg->algo_flags & RRL_ALGO_DETERMINISTIC) ? 1 : 0;
    A.sp = action_space(cfg);
    A.action_task = action_task; A.action_real = action_real; A.qrisk_out = qrisk_out; A.recovery = recovery;
    if (cfg->use_tensor_cores) return act_tc_launch(A, arena, L, stages, max_ctas, (cudaStream_t)stream);
    RRL_CHECK_ARG(stages == RRL_ACT_STAGE_ALL, "staged acting runs on the tcgen05 path only (use_tensor_cores >= 1)");
    static bool configured = false;
    const size_t smem = sizeof(FwdSmem<64>);
    if (!configured) {
        RRL_CUDA(cudaFuncSetAttribute(act_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    int64_t tiles = (n + 63) / 64;
    const int64_t cap = (int64_t)rrl_num_sms() * 2;
    const int grid = (int)(tiles < cap ? tiles : cap);
    act_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(A);
    RRL_CHECK_LAUNCH();
    return 0;
}

extern "C" int rrl_twin_q_forward(const rrl_agent_config_t* cfg, const float* arena, int net, int64_t n, const float* s,
                                  const float* a, float* q1, float* q2, void* stream) {
    CHECK_CFG(cfg);
    RRL_CHECK_ARG(arena && s && a && q1 && q2 && n > 0, "bad argument");
    RRL_CHECK_ARG(net == RRL_NET_CRITIC || net == RRL_NET_CRITIC_TARGET || net == RRL_NET_QRISK || net == RRL_NET_QRISK_TARGET,
                  "net is not a twin critic");
    const Layout L = make_layout(cfg);
    FwdArgs A;
    memset(&A, 0, sizeof(A));
    A.p[0] = q_pass(L, const_cast<float*>(arena), net, 0, s, a, -1, q1);
    A.p[1] = q_pass(L, const_cast<float*>(arena), net, 1, s, a, -1, q2);
    A.n_pass = 2;
    A.rows_const = n;
    A.sp = action_space(cfg);
    A.use_tc = cfg->use_tensor_cores;
    return launch_forward<64>(A, n, (cudaStream_t)stream);
}

extern "C" int rrl_policy_sample(const rrl_agent_config_t* cfg, const float* arena, int net, int64_t n, const float* s,
                                 const float* eps, float* action, float* log_prob, float* mean_action, void* stream) {
    CHECK_CFG(cfg);
    RRL_CHECK_ARG(arena && s && eps && action && n > 0, "bad argument");
    RRL_CHECK_ARG(net == RRL_NET_POLICY || net == RRL_NET_RECOVERY, "net is not a policy");
    const Layout L = make_layout(cfg);
    FwdArgs A;
    memset(&A, 0, sizeof(A));
    FwdPass& p = A.p[0];
    p.w = head_w(L, arena, net, 0);
    p.head = HEAD_STOCH;
    if (net == RRL_NET_POLICY) p.w = task_policy_w(L, arena, cfg, &p.head);
    p.xs = s; p.eps = eps;
    p.out_a = action; p.out_logp = log_prob; p.out_mean = mean_action;
    A.n_pass = 1;
    A.rows_const = n;
    A.sp = action_space(cfg);
    return launch_forward<64>(A, n, (cudaStream_t)stream);
}

// ---- SAC.update_parameters (sac.py:170-277), "Variant B" ordering ------------------------------
extern "C" int rrl_sac_backward(const rrl_agent_config_t* cfg, float* arena, const float* eps_next, const float* eps_cur,
                                uint64_t seed, int32_t stream_id, int64_t* counters, float* losses, void* stream) {
    CHECK_CFG(cfg);
    RRL_CHECK_ARG(arena && counters, "null argument");
    const Layout L = make_layout(cfg);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t R = L.R;
    const int64_t* rows_ptr = counters + RRL_C_SAC_ROWS;
    float* s = arena + L.batch_off[0][0]; float* a = arena + L.batch_off[0][1]; float* r = arena + L.batch_off[0][2];
    float* s2 = arena + L.batch_off[0][3]; float* m = arena + L.batch_off[0][4];
    auto RA = [&](int id) { return arena + L.rows_f[id]; };
    auto R2 = [&](int id) { return arena + L.rows2_f[id]; };
    auto R4 = [&](int id) { return arena + L.rows4_f[id]; };
    if (!losses) losses = arena + L.losses;
    float* scal = arena + L.scalars;
    const ActionSpace sp = action_space(cfg);
    const int flags = cfg->algo_flags;
    const bool det = (flags & RRL_ALGO_DETERMINISTIC) != 0;
    const bool rcpo = (flags & RRL_ALGO_RCPO) != 0;
    const bool dgd = (flags & RRL_ALGO_DGD) != 0;
    const bool sq_pi = dgd || (flags & RRL_ALGO_UPDATE_NU);  // sac.py:221-222 is dead code otherwise
    // (Deterministic policy: eps_next / eps_cur are the caller's noise vectors; NULL draws them from Philox, index 0 for every row)
    const bool fuse = cfg->use_tensor_cores >= 2;   // per-row stages run as tails of the producing kernels
    int pol_head = HEAD_GAUSS;
    const HeadW pw = task_policy_w(L, arena, cfg, &pol_head);
    const HeadG pg = head_g(L, arena, RRL_NET_POLICY, 0);

    {  // policy on s' (no grad, sac.py:192-194) and on s (sac.py:216)
        FwdArgs A;
        memset(&A, 0, sizeof(A));
        A.n_pass = 2; A.rows_ptr = rows_ptr; A.sp = sp; A.use_tc = cfg->use_tensor_cores; A.seed = seed; A.stream_id = (uint32_t)stream_id;
        A.counters = counters; A.step_counter = RRL_C_SAC_UPDATES;
        FwdPass& p0 = A.p[0];
        p0.w = pw; p0.head = pol_head; p0.xs = s2; p0.eps = eps_next; p0.tc_img = tc_img_of(L, arena, RRL_NET_POLICY, 0);
        p0.draw_id = RRL_DRAW_SAC_NEXT; p0.out_a = R2(R2_NEXT_A); p0.out_logp = RA(RA_NEXT_LOGP);
        FwdPass& p1 = A.p[1];
        p1.w = pw; p1.head = pol_head; p1.xs = s; p1.eps = eps_cur; p1.draw_id = RRL_DRAW_SAC_CUR; p1.tc_img = p0.tc_img;
        p1.h1 = arena + L.h1[4]; p1.h2 = arena + L.h2[4]; p1.h2bits = slot_bits(L, arena, 4); p1.h1bits = slot_bits1(L, arena, 4);
        p1.keep_h2 = 1;
        p1.out_a = R2(R2_PI); p1.out_logp = RA(RA_LOGP); p1.out_raw = R4(R4_RAW_POL); p1.out_eps = R2(R2_EPS_CUR);
        int rc = launch_forward<32>(A, R, st);
        if (rc) return rc;
    }
    {  // critic_target(s', a'), critic(s, a), critic(s, pi)   (sac.py:195-196, 206-207, 218)
       // [+ Q_risk(s, pi) (sac.py:221) for DGD / update_nu, + Q_risk(s, a) (sac.py:203-204) for RCPO]
        FwdArgs A;
        memset(&A, 0, sizeof(A));
        A.rows_ptr = rows_ptr; A.sp = sp; A.use_tc = cfg->use_tensor_cores;
        int n = 0;
        A.p[n++] = q_pass(L, arena, RRL_NET_CRITIC_TARGET, 0, s2, R2(R2_NEXT_A), -1, RA(RA_QT1));
        A.p[n++] = q_pass(L, arena, RRL_NET_CRITIC_TARGET, 1, s2, R2(R2_NEXT_A), -1, RA(RA_QT2));
        A.p[n++] = q_pass(L, arena, RRL_NET_CRITIC, 0, s, a, 0, RA(RA_QF1), true);
        A.p[n++] = q_pass(L, arena, RRL_NET_CRITIC, 1, s, a, 1, RA(RA_QF2), true);
        A.p[n++] = q_pass(L, arena, RRL_NET_CRITIC, 0, s, R2(R2_PI), 2, RA(RA_QP1));
        A.p[n++] = q_pass(L, arena, RRL_NET_CRITIC, 1, s, R2(R2_PI), 3, RA(RA_QP2));
        if (sq_pi) {
            A.p[n++] = q_pass(L, arena, RRL_NET_QRISK, 0, s, R2(R2_PI), dgd ? 5 : -1, RA(RA_SQ1));
            A.p[n++] = q_pass(L, arena, RRL_NET_QRISK, 1, s, R2(R2_PI), dgd ? 6 : -1, RA(RA_SQ2));
        }
        if (rcpo) {
            A.p[n++] = q_pass(L, arena, RRL_NET_QRISK, 0, s, a, -1, RA(RA_QS1));
            A.p[n++] = q_pass(L, arena, RRL_NET_QRISK, 1, s, a, -1, RA(RA_QS2));
        }
        A.n_pass = n;
        SacLossArgs& T = A.tail.sac;   // TD target + losses + output gradients: tail of this launch (fused) or its own launch
        T.r = r; T.m = m; T.next_logp = RA(RA_NEXT_LOGP); T.qt1 = RA(RA_QT1); T.qt2 = RA(RA_QT2);
        T.qf1 = RA(RA_QF1); T.qf2 = RA(RA_QF2); T.logp = RA(RA_LOGP); T.qp1 = RA(RA_QP1); T.qp2 = RA(RA_QP2);
        if (sq_pi) { T.sq1 = RA(RA_SQ1); T.sq2 = RA(RA_SQ2); T.dsq1 = RA(RA_DSQ1); T.dsq2 = RA(RA_DSQ2); }
        if (rcpo) { T.qs1 = RA(RA_QS1); T.qs2 = RA(RA_QS2); }
        T.target = RA(RA_TARGET); T.dqf1 = RA(RA_DQF1); T.dqf2 = RA(RA_DQF2); T.dqp1 = RA(RA_DQP1); T.dqp2 = RA(RA_DQP2);
        T.minq = RA(RA_MINQ); T.losses = losses; T.scal = scal; T.gamma = cfg->gamma; T.eps_safe = cfg->eps_safe;
        T.target_entropy = cfg->target_entropy; T.flags = flags; T.rows_ptr = rows_ptr;
        if (fuse) { A.tail.kind = TAIL_SAC_LOSS; A.tail.ticket = counters + RRL_C_TICKET2; }
        int rc = launch_forward<32>(A, R, st);
        if (rc) return rc;
        if (!fuse) {
            sac_loss_kernel<<<1, kThreads, 0, st>>>(T);
            RRL_CHECK_LAUNCH();
        }
    }
    const HeadW c1 = head_w(L, arena, RRL_NET_CRITIC, 0), c2 = head_w(L, arena, RRL_NET_CRITIC, 1);
    const HeadG g1 = head_g(L, arena, RRL_NET_CRITIC, 0), g2 = head_g(L, arena, RRL_NET_CRITIC, 1);
    const HeadW k1 = head_w(L, arena, RRL_NET_QRISK, 0), k2 = head_w(L, arena, RRL_NET_QRISK, 1);
    const int nq = dgd ? 6 : 4;  // passes whose gradient flows back: critic x4 [+ Q_risk(s, pi) x2 into the action]
    const int slot[6] = {0, 1, 2, 3, 5, 6};
    {  // head backward + dh1 for all passes + gW2 / gW3 / gb3 / gb2 for the (s, a) passes, one launch
        GemmArgs G;
        memset(&G, 0, sizeof(G));
        G.rows_ptr = rows_ptr;
        const float* dout[6] = {RA(RA_DQF1), RA(RA_DQF2), RA(RA_DQP1), RA(RA_DQP2), RA(RA_DSQ1), RA(RA_DSQ2)};
        int n = 0;
        for (int q = 0; q < nq; ++q) {
            GemmPass& p = G.p[n++];
            const HeadW& w = q < 4 ? ((q & 1) ? c2 : c1) : ((q & 1) ? k2 : k1);
            p.dout = dout[q]; p.stride = 1; p.n_out = 1; p.na = 1; p.W3a = w.W3a; p.h2 = arena + L.h2[slot[q]];
            p.B = w.W2; p.tc_imgT = w.tc_imgT; p.k_is_rows = 0; p.mask = arena + L.h1[slot[q]]; p.C = arena + L.dh1[slot[q]];
            bwd_inputs(p, w, s, q < 2 ? a : R2(R2_PI), L, arena, slot[q]);
        }
        for (int q = 0; q < 2; ++q) {
            GemmPass& p = G.p[n++];
            const HeadW& w = q ? c2 : c1;
            const HeadG& g = q ? g2 : g1;
            p.dout = dout[q]; p.stride = 1; p.n_out = 1; p.na = 1; p.W3a = w.W3a; p.h2 = arena + L.h2[q];
            p.B = arena + L.h1[q]; p.k_is_rows = 1; p.C = g.W2;
            p.gW3a = g.W3a; p.gb3a = g.b3a; p.gb2 = g.b2;
            bwd_inputs(p, w, s, a, L, arena, q);
        }
        // layer-1 backward: weight grads for (s, a); d/d(pi) for (s, pi) -- then the policy-sample backward (needs d(action)
        // of every pass).  Fused (use_tensor_cores 2): epilogue + tail of the GEMM launch; else launches of their own.
        L1BwdArgs A;
        memset(&A, 0, sizeof(A));
        A.rows_ptr = rows_ptr;
        float* dxa[6] = {nullptr, nullptr, R2(R2_DPI), R2(R2_DPI_B), R2(R2_DPI_S1), R2(R2_DPI_S2)};
        for (int q = 0; q < nq; ++q) {
            L1BwdPass& p = A.p[q];
            const HeadW& w = q < 4 ? ((q & 1) ? c2 : c1) : ((q & 1) ? k2 : k1);
            const HeadG& g = (q & 1) ? g2 : g1;
            p.dh1 = arena + L.dh1[slot[q]]; p.xs = s; p.xa = q < 2 ? a : R2(R2_PI); p.W1 = w.W1; p.n_in = 4;
            if (q < 2) { p.gW1 = g.W1; p.gb1 = g.b1; }
            else p.dxa = dxa[q];
            if (fuse) fuse_l1(G.p[q], p.gW1, p.gb1, p.dxa, L, arena, slot[q]);
        }
        TailArgs T;
        memset(&T, 0, sizeof(T));
        T.ticket = counters + RRL_C_TICKET2;
        if (!det) {
            GaussBwdArgs& B = T.gauss;
            B.raw = R4(R4_RAW_POL); B.eps = R2(R2_EPS_CUR); B.dxa1 = R2(R2_DPI); B.dxa2 = R2(R2_DPI_B);
            if (dgd) { B.dxa3 = R2(R2_DPI_S1); B.dxa4 = R2(R2_DPI_S2); }
            B.draw = R4(R4_DRAW_POL); B.scal = scal; B.sp = sp; B.rows_ptr = rows_ptr;
            T.kind = TAIL_GAUSS_BWD;
        } else {  // DeterministicPolicy: a = tanh(raw)*scale + bias + noise
            StochBwdArgs& B = T.stoch;
            B.raw = R4(R4_RAW_POL); B.eps = R2(R2_EPS_CUR); B.dxa1 = R2(R2_DPI); B.dxa2 = R2(R2_DPI_B);
            if (dgd) { B.dxa3 = R2(R2_DPI_S1); B.dxa4 = R2(R2_DPI_S2); }
            B.log_std = pw.log_std; B.draw = R4(R4_DRAW_POL); B.g_log_std = nullptr; B.sp = sp; B.rows_ptr = rows_ptr;
            T.kind = TAIL_STOCH_BWD;
        }
        if (fuse) G.tail = T;
        const int mt = (int)((R > H ? R : H) / 32);
        { int rc = launch_gemm(G, n, mt, R, cfg->use_tensor_cores, st); if (rc) return rc; }
        if (!fuse) {
            layer1_backward_kernel<<<dim3(kL1ColBlocks + kL1RowBlocks, nq), kThreads, 0, st>>>(A);
            RRL_CHECK_LAUNCH();
            if (!det) gauss_backward_kernel<<<(unsigned)((R + kThreads - 1) / kThreads), kThreads, 0, st>>>(T.gauss);
            else stoch_backward_kernel<<<1, kThreads, 0, st>>>(T.stoch);
            RRL_CHECK_LAUNCH();
        }
    }
    {  // policy: head backward + dh1 + gW2 / gW3 / gb3 / gb2
        GemmArgs G;
        memset(&G, 0, sizeof(G));
        G.rows_ptr = rows_ptr;
        for (int q = 0; q < 2; ++q) {
            GemmPass& p = G.p[q];
            p.dout = R4(R4_DRAW_POL); p.stride = 4; p.n_out = det ? 2 : 4; p.na = 2; p.W3a = pw.W3a; p.W3b = pw.W3b;
            p.h2 = arena + L.h2[4];
            bwd_inputs(p, pw, s, nullptr, L, arena, 4);
        }
        G.p[0].B = pw.W2; G.p[0].tc_imgT = pw.tc_imgT; G.p[0].mask = arena + L.h1[4]; G.p[0].C = arena + L.dh1[4];
        G.p[1].B = arena + L.h1[4]; G.p[1].k_is_rows = 1; G.p[1].C = pg.W2;
        G.p[1].gW3a = pg.W3a; G.p[1].gb3a = pg.b3a; G.p[1].gb2 = pg.b2;
        if (!det) { G.p[1].gW3b = pg.W3b; G.p[1].gb3b = pg.b3b; }
        if (fuse) fuse_l1(G.p[0], pg.W1, pg.b1, nullptr, L, arena, 4);
        const int mt = (int)((R > H ? R : H) / 32);
        { int rc = launch_gemm(G, 2, mt, R, cfg->use_tensor_cores, st); if (rc) return rc; }
    }
    if (!fuse) {
        L1BwdArgs A;
        memset(&A, 0, sizeof(A));
        A.rows_ptr = rows_ptr;
        L1BwdPass& p = A.p[0];
        p.dh1 = arena + L.dh1[4]; p.xs = s; p.W1 = pw.W1; p.n_in = 2; p.gW1 = pg.W1; p.gb1 = pg.b1;
        layer1_backward_kernel<<<dim3(kL1ColBlocks, 1), kThreads, 0, st>>>(A);
        RRL_CHECK_LAUNCH();
    }
    return 0;
}

static int sac_apply_impl(const rrl_agent_config_t* cfg, float* arena, int64_t* counters, const rrl_peers_t* peers, void* stream);
extern "C" int rrl_sac_apply(const rrl_agent_config_t* cfg, float* arena, int64_t* counters, void* stream) {
    return sac_apply_impl(cfg, arena, counters, nullptr, stream);
}
extern "C" int rrl_sac_apply_p2p(const rrl_agent_config_t* cfg, float* arena, int64_t* counters, const rrl_peers_t* peers,
                                 void* stream) {
    RRL_CHECK_ARG(peers && peers->world >= 1 && peers->world <= 8, "bad peers");
    return sac_apply_impl(cfg, arena, counters, peers, stream);
}
static int sac_apply_impl(const rrl_agent_config_t* cfg, float* arena, int64_t* counters, const rrl_peers_t* peers, void* stream) {
    CHECK_CFG(cfg);
    RRL_CHECK_ARG(arena && counters, "null argument");
    const Layout L = make_layout(cfg);
    cudaStream_t st = (cudaStream_t)stream;
    // critic_optim.step(); policy_optim.step()  (sac.py:233-239; both Adams have the same step count), then
    // soft_update(critic_target, critic, tau) if updates % target_update_interval == 0  (sac.py:273-274), then the
    // step bookkeeping -- one launch
    const int bump[3] = {RRL_C_ADAM_T0 + 0, RRL_C_ADAM_T0 + 1, RRL_C_SAC_UPDATES};
    int rc = launch_adam(cfg, L, arena, counters, RRL_NET_CRITIC, RRL_NET_POLICY, RRL_C_ADAM_T0 + 0, RRL_C_SAC_ROWS, st,
                         RRL_NET_CRITIC_TARGET, RRL_NET_CRITIC, cfg->tau, RRL_C_SAC_UPDATES, bump, 3, peers);
    if (rc) return rc;
    if (cfg->algo_flags & (RRL_ALGO_AUTO_ALPHA | RRL_ALGO_UPDATE_NU | RRL_ALGO_RCPO)) {  // sac.py:241-271
        ScalarAdamArgs A;
        A.scal = arena + L.scalars; A.counters = counters;
        A.lr = cfg->lr64 > 0.0 ? cfg->lr64 : (double)cfg->lr;
        A.b1 = (double)cfg->beta1; A.b2 = (double)cfg->beta2; A.eps = (double)cfg->adam_eps;
        A.grad_scale = (double)cfg->grad_scale; A.flags = cfg->algo_flags; A.rows_counter = RRL_C_SAC_ROWS;
        scalar_adam_kernel<<<1, 32, 0, st>>>(A);
        RRL_CHECK_LAUNCH();
    }
    return 0;
}

// ---- QRiskWrapper.update_parameters (qrisk.py:86-182) --------------------------------------------
extern "C" int rrl_qrisk_backward(const rrl_agent_config_t* cfg, float* arena, const float* eps_next, uint64_t seed,
                                  int32_t stream_id, int64_t* counters, float* losses, void* stream) {
    CHECK_CFG(cfg);
    RRL_CHECK_ARG(arena && counters, "null argument");
    const Layout L = make_layout(cfg);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t R = L.R;
    const int64_t* rows_ptr = counters + RRL_C_QRISK_ROWS;
    float* s = arena + L.batch_off[1][0]; float* a = arena + L.batch_off[1][1]; float* c = arena + L.batch_off[1][2];
    float* s2 = arena + L.batch_off[1][3]; float* m = arena + L.batch_off[1][4];
    auto RA = [&](int id) { return arena + L.rows_f[id]; };
    auto R2 = [&](int id) { return arena + L.rows2_f[id]; };
    if (!losses) losses = arena + L.losses + 8;
    const ActionSpace sp = action_space(cfg);
    const bool fuse = cfg->use_tensor_cores >= 2;
    {  // a' ~ TASK policy(s')  (qrisk.py:118-120; policy = agent.policy, experiment.py:413)
        FwdArgs A;
        memset(&A, 0, sizeof(A));
        A.n_pass = 1; A.rows_ptr = rows_ptr; A.sp = sp; A.use_tc = cfg->use_tensor_cores; A.seed = seed; A.stream_id = (uint32_t)stream_id;
        A.counters = counters; A.step_counter = RRL_C_QRISK_UPDATES;
        FwdPass& p0 = A.p[0];
        p0.w = task_policy_w(L, arena, cfg, &p0.head); p0.xs = s2; p0.eps = eps_next;
        p0.tc_img = tc_img_of(L, arena, RRL_NET_POLICY, 0);
        p0.draw_id = RRL_DRAW_QR_NEXT; p0.out_a = R2(R2_QR_NEXT_A); p0.out_logp = RA(RA_QR_NEXT_LOGP);
        int rc = launch_forward<32>(A, R, st);
        if (rc) return rc;
    }
    {
        FwdArgs A;
        memset(&A, 0, sizeof(A));
        A.n_pass = 4; A.rows_ptr = rows_ptr; A.sp = sp; A.use_tc = cfg->use_tensor_cores;
        A.p[0] = q_pass(L, arena, RRL_NET_QRISK_TARGET, 0, s2, R2(R2_QR_NEXT_A), -1, RA(RA_QR_QT1));
        A.p[1] = q_pass(L, arena, RRL_NET_QRISK_TARGET, 1, s2, R2(R2_QR_NEXT_A), -1, RA(RA_QR_QT2));
        A.p[2] = q_pass(L, arena, RRL_NET_QRISK, 0, s, a, 0, RA(RA_QR_Q1), true);
        A.p[3] = q_pass(L, arena, RRL_NET_QRISK, 1, s, a, 1, RA(RA_QR_Q2), true);
        QrLossArgs& T = A.tail.qr;
        T.c = c; T.m = m; T.qt1 = RA(RA_QR_QT1); T.qt2 = RA(RA_QR_QT2); T.q1 = RA(RA_QR_Q1); T.q2 = RA(RA_QR_Q2);
        T.target = RA(RA_QR_TARGET); T.dq1 = RA(RA_QR_DQ1); T.dq2 = RA(RA_QR_DQ2); T.losses = losses;
        T.gamma_safe = cfg->gamma_safe; T.rows_ptr = rows_ptr;
        if (fuse) { A.tail.kind = TAIL_QR_LOSS; A.tail.ticket = counters + RRL_C_TICKET2; }
        int rc = launch_forward<32>(A, R, st);
        if (rc) return rc;
        if (!fuse) {
            qrisk_loss_kernel<<<1, kThreads, 0, st>>>(T);
            RRL_CHECK_LAUNCH();
        }
    }
    const HeadW c1 = head_w(L, arena, RRL_NET_QRISK, 0), c2 = head_w(L, arena, RRL_NET_QRISK, 1);
    const HeadG g1 = head_g(L, arena, RRL_NET_QRISK, 0), g2 = head_g(L, arena, RRL_NET_QRISK, 1);
    {  // head backward + dh1 + gW2 / gW3 / gb3 / gb2 of both heads
        GemmArgs G;
        memset(&G, 0, sizeof(G));
        G.rows_ptr = rows_ptr;
        for (int q = 0; q < 2; ++q) {
            const HeadW& w = q ? c2 : c1;
            const HeadG& g = q ? g2 : g1;
            GemmPass& p = G.p[q];
            p.dout = q ? RA(RA_QR_DQ2) : RA(RA_QR_DQ1); p.stride = 1; p.n_out = 1; p.na = 1; p.W3a = w.W3a;
            p.h2 = arena + L.h2[q]; p.B = w.W2; p.tc_imgT = w.tc_imgT; p.mask = arena + L.h1[q]; p.C = arena + L.dh1[q];
            bwd_inputs(p, w, s, a, L, arena, q);
            GemmPass& ww = G.p[2 + q];
            ww = p;
            ww.B = arena + L.h1[q]; ww.k_is_rows = 1; ww.mask = nullptr; ww.C = g.W2;
            ww.gW3a = g.W3a; ww.gb3a = g.b3a; ww.gb2 = g.b2;
            if (fuse) fuse_l1(p, g.W1, g.b1, nullptr, L, arena, q);
        }
        const int mt = (int)((R > H ? R : H) / 32);
        { int rc = launch_gemm(G, 4, mt, R, cfg->use_tensor_cores, st); if (rc) return rc; }
    }
    if (!fuse) {
        L1BwdArgs A;
        memset(&A, 0, sizeof(A));
        A.rows_ptr = rows_ptr;
        for (int q = 0; q < 2; ++q) {
            L1BwdPass& p = A.p[q];
            p.dh1 = arena + L.dh1[q]; p.xs = s; p.xa = a; p.W1 = (q ? c2 : c1).W1; p.n_in = 4;
            p.gW1 = (q ? g2 : g1).W1; p.gb1 = (q ? g2 : g1).b1;
        }
        layer1_backward_kernel<<<dim3(kL1ColBlocks, 2), kThreads, 0, st>>>(A);
        RRL_CHECK_LAUNCH();
    }
    return 0;
}

static int qrisk_apply_impl(const rrl_agent_config_t* cfg, float* arena, int64_t* counters, const rrl_peers_t* peers, void* stream);
extern "C" int rrl_qrisk_apply(const rrl_agent_config_t* cfg, float* arena, int64_t* counters, void* stream) {
    return qrisk_apply_impl(cfg, arena, counters, nullptr, stream);
}
extern "C" int rrl_qrisk_apply_p2p(const rrl_agent_config_t* cfg, float* arena, int64_t* counters, const rrl_peers_t* peers,
                                   void* stream) {
    RRL_CHECK_ARG(peers && peers->world >= 1 && peers->world <= 8, "bad peers");
    return qrisk_apply_impl(cfg, arena, counters, peers, stream);
}
static int qrisk_apply_impl(const rrl_agent_config_t* cfg, float* arena, int64_t* counters, const rrl_peers_t* peers, void* stream) {
    CHECK_CFG(cfg);
    RRL_CHECK_ARG(arena && counters, "null argument");
    const Layout L = make_layout(cfg);
    cudaStream_t st = (cudaStream_t)stream;
    // safety_critic_optim.step() (qrisk.py:146-148) and, in the same launch, the soft update of the target
    // (qrisk.py:160-162: it reads the post-step critic, which the recovery-policy update in between does not touch)
    const int bump[1] = {RRL_C_ADAM_T0 + 2};
    return launch_adam(cfg, L, arena, counters, RRL_NET_QRISK, -1, RRL_C_ADAM_T0 + 2, RRL_C_QRISK_ROWS, st,
                       RRL_NET_QRISK_TARGET, RRL_NET_QRISK, cfg->tau_safe, RRL_C_QRISK_UPDATES, bump, 1, peers);
}

// recovery policy on the POST-step safety critic (qrisk.py:150-158), then Polyak (qrisk.py:160-163)
// do_forward: the recovery policy's own forward pass (needs only the sampled batch and the recovery policy: the vector
// engine enqueues it next to the SAC update);  do_rest: everything that needs the POST-step safety critic
static int recovery_backward_impl(const rrl_agent_config_t* cfg, float* arena, const float* eps_rec, uint64_t seed,
                                  int32_t stream_id, int64_t* counters, float* losses, void* stream, bool do_forward, bool do_rest) {
    CHECK_CFG(cfg);
    RRL_CHECK_ARG(arena && counters, "null argument");
    const Layout L = make_layout(cfg);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t R = L.R;
    const int64_t* rows_ptr = counters + RRL_C_QRISK_ROWS;
    float* s = arena + L.batch_off[1][0];
    auto RA = [&](int id) { return arena + L.rows_f[id]; };
    auto R2 = [&](int id) { return arena + L.rows2_f[id]; };
    auto R4 = [&](int id) { return arena + L.rows4_f[id]; };
    if (!losses) losses = arena + L.losses + 8;
    const ActionSpace sp = action_space(cfg);
    const bool fuse = cfg->use_tensor_cores >= 2;
    if (!cfg->mf_recovery) return 0;
    if (do_forward) {
        FwdArgs A;
        memset(&A, 0, sizeof(A));
        A.n_pass = 1; A.rows_ptr = rows_ptr; A.sp = sp; A.use_tc = cfg->use_tensor_cores; A.seed = seed; A.stream_id = (uint32_t)stream_id;
        A.counters = counters; A.step_counter = RRL_C_QRISK_UPDATES;
        FwdPass& p0 = A.p[0];
        p0.w = head_w(L, arena, RRL_NET_RECOVERY, 0); p0.head = HEAD_STOCH; p0.xs = s; p0.eps = eps_rec;
        p0.tc_img = tc_img_of(L, arena, RRL_NET_RECOVERY, 0);
        p0.draw_id = RRL_DRAW_QR_REC; p0.h1 = arena + L.h1[kRecSlot]; p0.h2 = arena + L.h2[kRecSlot];
        p0.h2bits = slot_bits(L, arena, kRecSlot); p0.h1bits = slot_bits1(L, arena, kRecSlot); p0.keep_h2 = 1;
        p0.out_a = R2(R2_REC_PI); p0.out_logp = RA(RA_REC_LOGP); p0.out_raw = R4(R4_RAW_REC); p0.out_eps = R2(R2_REC_EPS);
        int rc = launch_forward<32>(A, R, st);
        if (rc) return rc;
    }
    if (!do_rest) return 0;
    {
        FwdArgs A;
        memset(&A, 0, sizeof(A));
        A.n_pass = 2; A.rows_ptr = rows_ptr; A.sp = sp; A.use_tc = cfg->use_tensor_cores;
        A.p[0] = q_pass(L, arena, RRL_NET_QRISK, 0, s, R2(R2_REC_PI), 2, RA(RA_REC_Q1));
        A.p[1] = q_pass(L, arena, RRL_NET_QRISK, 1, s, R2(R2_REC_PI), 3, RA(RA_REC_Q2));
        RecLossArgs& T = A.tail.rec;
        T.q1 = RA(RA_REC_Q1); T.q2 = RA(RA_REC_Q2); T.dq1 = RA(RA_REC_DQ1); T.dq2 = RA(RA_REC_DQ2); T.losses = losses;
        T.rows_ptr = rows_ptr;
        if (fuse) { A.tail.kind = TAIL_REC_LOSS; A.tail.ticket = counters + RRL_C_TICKET2; }
        int rc = launch_forward<32>(A, R, st);
        if (rc) return rc;
        if (!fuse) {
            recovery_loss_kernel<<<1, kThreads, 0, st>>>(T);
            RRL_CHECK_LAUNCH();
        }
    }
    const HeadW c1 = head_w(L, arena, RRL_NET_QRISK, 0), c2 = head_w(L, arena, RRL_NET_QRISK, 1);
    const HeadW pw = head_w(L, arena, RRL_NET_RECOVERY, 0);
    const HeadG pg = head_g(L, arena, RRL_NET_RECOVERY, 0);
    {  // back through the (post-step) safety critic into the action: head backward + dh1, layer-1 backward (d(action)) and
       // the StochasticPolicy.sample backward -- fused into the GEMM launch (use_tensor_cores 2) or launches of their own
        GemmArgs G;
        memset(&G, 0, sizeof(G));
        G.rows_ptr = rows_ptr;
        L1BwdArgs A;
        memset(&A, 0, sizeof(A));
        A.rows_ptr = rows_ptr;
        for (int q = 0; q < 2; ++q) {
            GemmPass& p = G.p[q];
            p.dout = q ? RA(RA_REC_DQ2) : RA(RA_REC_DQ1); p.stride = 1; p.n_out = 1; p.na = 1; p.W3a = (q ? c2 : c1).W3a;
            p.h2 = arena + L.h2[2 + q]; p.B = (q ? c2 : c1).W2; p.tc_imgT = (q ? c2 : c1).tc_imgT;
            p.mask = arena + L.h1[2 + q]; p.C = arena + L.dh1[2 + q];
            bwd_inputs(p, q ? c2 : c1, s, R2(R2_REC_PI), L, arena, 2 + q);
            L1BwdPass& l = A.p[q];
            l.dh1 = arena + L.dh1[2 + q]; l.xs = s; l.xa = R2(R2_REC_PI); l.W1 = (q ? c2 : c1).W1; l.n_in = 4;
            l.dxa = q ? R2(R2_DPI) : R2(R2_REC_DPI);
            if (fuse) fuse_l1(p, nullptr, nullptr, l.dxa, L, arena, 2 + q);
        }
        TailArgs T;
        memset(&T, 0, sizeof(T));
        T.ticket = counters + RRL_C_TICKET2;
        T.kind = TAIL_STOCH_BWD;
        StochBwdArgs& B = T.stoch;
        B.raw = R4(R4_RAW_REC); B.eps = R2(R2_REC_EPS); B.dxa1 = R2(R2_REC_DPI); B.dxa2 = R2(R2_DPI);
        B.log_std = pw.log_std; B.draw = R4(R4_DRAW_REC); B.g_log_std = pg.log_std; B.sp = sp; B.rows_ptr = rows_ptr;
        if (fuse) G.tail = T;
        { int rc = launch_gemm(G, 2, (int)(R / 32), R, cfg->use_tensor_cores, st); if (rc) return rc; }
        if (!fuse) {
            layer1_backward_kernel<<<dim3(kL1ColBlocks + kL1RowBlocks, 2), kThreads, 0, st>>>(A);
            RRL_CHECK_LAUNCH();
            stoch_backward_kernel<<<1, kThreads, 0, st>>>(B);
            RRL_CHECK_LAUNCH();
        }
    }
    {  // recovery policy: head backward + dh1 + gW2 / gW3 / gb3 / gb2 (+ layer-1 weight gradients when fused)
        GemmArgs G;
        memset(&G, 0, sizeof(G));
        G.rows_ptr = rows_ptr;
        for (int q = 0; q < 2; ++q) {
            GemmPass& p = G.p[q];
            p.dout = R4(R4_DRAW_REC); p.stride = 4; p.n_out = 2; p.na = 2; p.W3a = pw.W3a; p.h2 = arena + L.h2[kRecSlot];
            bwd_inputs(p, pw, s, nullptr, L, arena, kRecSlot);
        }
        G.p[0].B = pw.W2; G.p[0].tc_imgT = pw.tc_imgT; G.p[0].mask = arena + L.h1[kRecSlot]; G.p[0].C = arena + L.dh1[kRecSlot];
        G.p[1].B = arena + L.h1[kRecSlot]; G.p[1].k_is_rows = 1; G.p[1].C = pg.W2;
        G.p[1].gW3a = pg.W3a; G.p[1].gb3a = pg.b3a; G.p[1].gb2 = pg.b2;
        if (fuse) fuse_l1(G.p[0], pg.W1, pg.b1, nullptr, L, arena, kRecSlot);
        const int mt = (int)((R > H ? R : H) / 32);
        { int rc = launch_gemm(G, 2, mt, R, cfg->use_tensor_cores, st); if (rc) return rc; }
    }
    if (!fuse) {
        L1BwdArgs A;
        memset(&A, 0, sizeof(A));
        A.rows_ptr = rows_ptr;
        L1BwdPass& p = A.p[0];
        p.dh1 = arena + L.dh1[kRecSlot]; p.xs = s; p.W1 = pw.W1; p.n_in = 2; p.gW1 = pg.W1; p.gb1 = pg.b1;
        layer1_backward_kernel<<<dim3(kL1ColBlocks, 1), kThreads, 0, st>>>(A);
        RRL_CHECK_LAUNCH();
    }
    return 0;
}
extern "C" int rrl_recovery_backward(const rrl_agent_config_t* cfg, float* arena, const float* eps_rec, uint64_t seed,
                                     int32_t stream_id, int64_t* counters, float* losses, void* stream) {
    return recovery_backward_impl(cfg, arena, eps_rec, seed, stream_id, counters, losses, stream, true, true);
}
extern "C" int rrl_recovery_forward(const rrl_agent_config_t* cfg, float* arena, const float* eps_rec, uint64_t seed,
                                    int32_t stream_id, int64_t* counters, void* stream) {
    return recovery_backward_impl(cfg, arena, eps_rec, seed, stream_id, counters, nullptr, stream, true, false);
}
extern "C" int rrl_recovery_backward_rest(const rrl_agent_config_t* cfg, float* arena, int64_t* counters, float* losses, void* stream) {
    return recovery_backward_impl(cfg, arena, nullptr, 0, 0, counters, losses, stream, false, true);
}

static int recovery_apply_impl(const rrl_agent_config_t* cfg, float* arena, int64_t* counters, const rrl_peers_t* peers, void* stream);
extern "C" int rrl_recovery_apply(const rrl_agent_config_t* cfg, float* arena, int64_t* counters, void* stream) {
    return recovery_apply_impl(cfg, arena, counters, nullptr, stream);
}
extern "C" int rrl_recovery_apply_p2p(const rrl_agent_config_t* cfg, float* arena, int64_t* counters, const rrl_peers_t* peers,
                                      void* stream) {
    RRL_CHECK_ARG(peers && peers->world >= 1 && peers->world <= 8, "bad peers");
    return recovery_apply_impl(cfg, arena, counters, peers, stream);
}
static int recovery_apply_impl(const rrl_agent_config_t* cfg, float* arena, int64_t* counters, const rrl_peers_t* peers, void* stream) {
    CHECK_CFG(cfg);
    RRL_CHECK_ARG(arena && counters, "null argument");
    const Layout L = make_layout(cfg);
    cudaStream_t st = (cudaStream_t)stream;
    // policy_optim.step() (qrisk.py:156-158); self.updates += 1 (qrisk.py:163)
    if (cfg->mf_recovery) {
        const int bump[2] = {RRL_C_ADAM_T0 + 3, RRL_C_QRISK_UPDATES};
        return launch_adam(cfg, L, arena, counters, RRL_NET_RECOVERY, -1, RRL_C_ADAM_T0 + 3, RRL_C_QRISK_ROWS, st, -1, -1,
                           0.f, -1, bump, 2, peers);
    }
    bump_kernel<<<1, 32, 0, st>>>(counters, RRL_C_QRISK_ROWS, -1, -1, RRL_C_QRISK_UPDATES);
    RRL_CHECK_LAUNCH();
    return 0;
}

static int peer_barrier_impl(const rrl_peers_t* peers, int64_t* epoch, int64_t* counters, int exchange, int gate_batch,
                             double gate_pos_fraction, void* stream);
extern "C" int rrl_peer_barrier(const rrl_peers_t* peers, int64_t* epoch, int64_t* counters, void* stream) {
    return peer_barrier_impl(peers, epoch, counters, 0, 0, 0.0, stream);
}
extern "C" int rrl_peer_sync_gate_counts(const rrl_peers_t* peers, int64_t* epoch, int64_t* counters, int32_t gate_batch,
                                         double gate_pos_fraction, void* stream) {
    return peer_barrier_impl(peers, epoch, counters, 1, gate_batch, gate_pos_fraction, stream);
}
static int peer_barrier_impl(const rrl_peers_t* peers, int64_t* epoch, int64_t* counters, int exchange, int gate_batch,
                             double gate_pos_fraction, void* stream) {
    RRL_CHECK_ARG(peers && epoch && counters && peers->world >= 1 && peers->world <= 8 && peers->rank >= 0 &&
                      peers->rank < peers->world, "bad argument");
    PeerBarrierArgs A;
    memset(&A, 0, sizeof(A));
    for (int r = 0; r < peers->world; ++r) A.signal[r] = reinterpret_cast<uint32_t*>(peers->signal[r]);
    A.world = peers->world; A.rank = peers->rank; A.epoch = epoch; A.counters = counters; A.exchange = exchange;
    A.gate_batch = gate_batch; A.gate_pos_fraction = gate_pos_fraction;
    peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(A);
    RRL_CHECK_LAUNCH();
    return 0;
}

extern "C" int rrl_debug_opt_times(unsigned long long* out8) {
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    if (cudaMemcpyFromSymbol(out8, g_opt_dbg, sizeof(unsigned long long) * 8) != cudaSuccess) return -1;
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (cudaMemcpyToSymbol(g_opt_dbg, z, sizeof(z)) != cudaSuccess) return -1;
    return 0;
}
