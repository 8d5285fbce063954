// mpc.cu -- model-based recovery policy (BASELINE config 5): the PETS / CEM planner of the reference on the device.
//
// Replaces recovery_rl/MPC.py:322-347 (act), :374-467 (_compile_cost, _predict_next_obs, TS-infinity particle
// bookkeeping), recovery_rl/optimizers.py:73-124 (CEMOptimizer.obtain_solution) and the PtModel forward of
// config/maze.py:71-96 (same class in config/navigation1.py / navigation2.py).
//
// One planning call = max_iters x [ sample candidates -> roll every (candidate, particle) through the learned
// ensemble for plan_hor steps, accumulating max(Q1, Q2)_risk (the ONLY cost: obs_cost_fn / ac_cost_fn are never
// called by the reference, MPC.py:406-412) -> elite statistics ], for E env copies at once.
//
// Layout.  The ensemble (5 nets, 4 -> 200 -> 200 -> 200 -> 4, swish) is packed into a "dyn image": hidden width
// zero-padded to 256 (a padded unit has pre-activation 0 and swish(0) = 0, so the arithmetic is unchanged) with
// every matrix k-major -- the reference stores lin_w as [net][in][out], which IS k-major.  Rows are ordered
// (env, net, candidate, particle-in-net): particle p of a candidate runs on net p / (npart / nets) for the whole
// horizon (TS-infinity, MPC.py:441-467), so a 64-row tile shares one net's weights.
// This file is the fp32 SIMT implementation (4 tile GEMMs per horizon step: Q_risk head 1 / 2, ensemble layers 1 / 2).
#include "mlp_tile.cuh"
#include "mpc_layout.cuh"

using namespace rrl;
using namespace rrl::dyn;

namespace {

struct PlanSmem {
    FwdSmem<PBM> f;
    alignas(16) float b2b[H];  // bias of the ensemble's second 256x256 layer
    float obs[2][PBM];   // the particles' current observation
    float act[2][PBM];
    float cost[PBM];
};

__device__ __forceinline__ float swishf(float x) { return x / (1.0f + expf(-x)); }   // x * sigmoid(x), config/utils.py:6
__device__ __forceinline__ float softplusf(float x) { return x > 20.0f ? x : log1pf(expf(x)); }  // F.softplus (threshold 20)

// pack kernel: reference-layout tensors -> dyn image
struct PackArgs {
    const float *w0, *b0, *w1, *b1, *w2, *b2, *w3, *b3, *mu, *sigma, *maxlv, *minlv;  // [5][4][200], [5][1][200], [5][200][200], ...
    float* img;
    int hid;  // 200
};
__global__ void __launch_bounds__(256) dyn_pack_kernel(const PackArgs A) {
    const int64_t i = blockIdx.x * 256ll + threadIdx.x;
    if (i >= kDynSimtFloats) return;
    const int hid = A.hid;
    float v = 0.f;
    if (i < kB0) {
        const int e = (int)(i / (DYN_IN * H)), r = (int)(i % (DYN_IN * H)), k = r / H, n = r % H;
        if (n < hid) v = A.w0[((int64_t)e * DYN_IN + k) * hid + n];
    } else if (i < kW1) {
        const int e = (int)((i - kB0) / H), n = (int)((i - kB0) % H);
        if (n < hid) v = A.b0[(int64_t)e * hid + n];
    } else if (i < kB1) {
        const int64_t j = i - kW1;
        const int e = (int)(j / (H * H)), k = (int)((j % (H * H)) / H), n = (int)(j % H);
        if (k < hid && n < hid) v = A.w1[((int64_t)e * hid + k) * hid + n];
    } else if (i < kW2) {
        const int e = (int)((i - kB1) / H), n = (int)((i - kB1) % H);
        if (n < hid) v = A.b1[(int64_t)e * hid + n];
    } else if (i < kB2) {
        const int64_t j = i - kW2;
        const int e = (int)(j / (H * H)), k = (int)((j % (H * H)) / H), n = (int)(j % H);
        if (k < hid && n < hid) v = A.w2[((int64_t)e * hid + k) * hid + n];
    } else if (i < kW3) {
        const int e = (int)((i - kB2) / H), n = (int)((i - kB2) % H);
        if (n < hid) v = A.b2[(int64_t)e * hid + n];
    } else if (i < kB3) {
        const int64_t j = i - kW3;
        const int e = (int)(j / (DYN_OUT * H)), o = (int)((j % (DYN_OUT * H)) / H), k = (int)(j % H);
        if (k < hid) v = A.w3[((int64_t)e * hid + k) * DYN_OUT + o];
    } else if (i < kMu) {
        v = A.b3[i - kB3];
    } else if (i < kSigma) {
        v = A.mu[i - kMu];
    } else if (i < kMaxLv) {
        v = A.sigma[i - kSigma];
    } else if (i < kMinLv) {
        v = A.maxlv[i - kMaxLv];
    } else if (i < kMinLv + 2) {
        v = A.minlv[i - kMinLv];
    }
    A.img[i] = v;
}

// ---- the rollout kernel ------------------------------------------------------------------------------
struct RolloutArgs {
    const float* dyn;
    HeadW qr1, qr2;
    const double* state;   // [2][E] fp64 (MPC.py:384 torch.from_numpy(obs).float())
    const float* samples;  // [E][pop][hor*2]
    const float* eps;      // NULL (Philox) or [E][hor][nets][pop*npn][2]
    const int32_t* active; // [E] or NULL
    float* row_cost;       // [E][pop][npart]
    int64_t E;
    int pop, hor, npart, npn, tiles_per_net;
    uint64_t seed;
    uint32_t stream_id;
    const int64_t* counters;
    int iter;
};

// one ensemble net on the tile: xin (already normalised) -> swish(L0) -> swish(L1) -> swish(L2) -> 4 outputs in S.f.raw
__device__ void ens_tile_forward(PlanSmem& S, const float* __restrict__ dyn, int net) {
    constexpr int RPW = PBM / 8;
    FwdSmem<PBM>& F = S.f;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const float* W1 = dyn + kW1 + (int64_t)net * H * H;
    const float* W2 = dyn + kW2 + (int64_t)net * H * H;
    load_b_chunk(F, W1, 0, 0);
    {   // stage the small tensors (t == hidden unit)
        const float* w0 = dyn + kW0 + (int64_t)net * DYN_IN * H;
        *reinterpret_cast<float4*>(F.W1s[t]) = make_float4(w0[t], w0[H + t], w0[2 * H + t], w0[3 * H + t]);
        F.b1s[t] = dyn[kB0 + net * H + t];
        F.b2s[t] = dyn[kB1 + net * H + t];
        S.b2b[t] = dyn[kB2 + net * H + t];
#pragma unroll
        for (int o = 0; o < 4; ++o) F.w3s[o][t] = dyn[kW3 + ((int64_t)net * DYN_OUT + o) * H + t];
        if (t < 4) F.b3s[t] = dyn[kB3 + net * DYN_OUT + t];
    }
    __syncthreads();
    {   // layer 0 (K = 4)
        const int m = t % PBM, kb = t / PBM;
        constexpr int KSTEP = kThreads / PBM;
        const float x0 = F.xin[0][m], x1 = F.xin[1][m], x2 = F.xin[2][m], x3 = F.xin[3][m];
#pragma unroll 4
        for (int k = kb; k < H; k += KSTEP) {
            const float4 wv = *reinterpret_cast<const float4*>(F.W1s[k]);
            float h = fmaf(wv.x, x0, F.b1s[k]);
            h = fmaf(wv.y, x1, h);
            h = fmaf(wv.z, x2, h);
            h = fmaf(wv.w, x3, h);
            F.As[k][m] = swishf(h);
        }
    }
    float acc[RPW][8];
    tile_gemm<PBM>(F, W1, acc);                 // ends with a __syncthreads: As / Bs free
    load_b_chunk(F, W2, 0, 0);
    {   // h = swish(acc + b1) back into the k-major tile: rows warp*8 .. +7 are contiguous in As[col][*]
        const float4 bb0 = *reinterpret_cast<const float4*>(&F.b2s[lane * 4]);
        const float4 bb1 = *reinterpret_cast<const float4*>(&F.b2s[128 + lane * 4]);
        const float bias[8] = {bb0.x, bb0.y, bb0.z, bb0.w, bb1.x, bb1.y, bb1.z, bb1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int col = (j < 4 ? 0 : 128) + lane * 4 + (j & 3);
            float hv[RPW];
#pragma unroll
            for (int r = 0; r < RPW; ++r) hv[r] = swishf(acc[r][j] + bias[j]);
            *reinterpret_cast<float4*>(&F.As[col][warp * RPW]) = make_float4(hv[0], hv[1], hv[2], hv[3]);
            *reinterpret_cast<float4*>(&F.As[col][warp * RPW + 4]) = make_float4(hv[4], hv[5], hv[6], hv[7]);
        }
    }
    tile_gemm<PBM>(F, W2, acc);
    {   // h = swish(acc + b2); the four outputs (mean x2, raw log-variance x2)
        const float4 bb0 = *reinterpret_cast<const float4*>(&S.b2b[lane * 4]);
        const float4 bb1 = *reinterpret_cast<const float4*>(&S.b2b[128 + lane * 4]);
        float myraw[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            float h[8];
            h[0] = swishf(acc[r][0] + bb0.x); h[1] = swishf(acc[r][1] + bb0.y);
            h[2] = swishf(acc[r][2] + bb0.z); h[3] = swishf(acc[r][3] + bb0.w);
            h[4] = swishf(acc[r][4] + bb1.x); h[5] = swishf(acc[r][5] + bb1.y);
            h[6] = swishf(acc[r][6] + bb1.z); h[7] = swishf(acc[r][7] + bb1.w);
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                const float4 w0 = *reinterpret_cast<const float4*>(&F.w3s[o][lane * 4]);
                const float4 w1 = *reinterpret_cast<const float4*>(&F.w3s[o][128 + lane * 4]);
                float p = h[0] * w0.x;
                p = fmaf(h[1], w0.y, p); p = fmaf(h[2], w0.z, p); p = fmaf(h[3], w0.w, p);
                p = fmaf(h[4], w1.x, p); p = fmaf(h[5], w1.y, p); p = fmaf(h[6], w1.z, p); p = fmaf(h[7], w1.w, p);
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) p += __shfl_xor_sync(0xffffffffu, p, s);
                if (lane == r) myraw[o] = p + F.b3s[o];
            }
        }
        if (lane < RPW) *reinterpret_cast<float4*>(F.raw[warp * RPW + lane]) = make_float4(myraw[0], myraw[1], myraw[2], myraw[3]);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kThreads, 2) mpc_rollout_kernel(const __grid_constant__ RolloutArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PlanSmem& S = *reinterpret_cast<PlanSmem*>(smem_raw);
    const int t = threadIdx.x;
    const int64_t tile = blockIdx.x;
    const int jt = (int)(tile % A.tiles_per_net);
    const int net = (int)((tile / A.tiles_per_net) % NETS);
    const int64_t e = tile / ((int64_t)A.tiles_per_net * NETS);
    if (A.active && !A.active[e]) return;
    const int rows_net = A.pop * A.npn;            // rows of one (env, net)
    const int j = jt * PBM + t;                    // row inside (env, net); candidate c, particle pl of this net
    const bool live = t < PBM && j < rows_net;
    const int c = live ? j / A.npn : 0, pl = live ? j % A.npn : 0;
    const float* mu = A.dyn + kMu;
    const float* sg = A.dyn + kSigma;
    float ox = 0.f, oy = 0.f, cost = 0.f;
    if (t < PBM) {
        ox = (float)A.state[e];
        oy = (float)A.state[A.E + e];
    }
    const uint64_t vstep = A.counters ? (uint64_t)A.counters[RRL_C_VEC_STEP] : 0;
    for (int step = 0; step < A.hor; ++step) {
        float ax = 0.f, ay = 0.f;
        if (t < PBM) {
            if (live) {
                const float2 a = *reinterpret_cast<const float2*>(A.samples + ((size_t)e * A.pop + c) * (A.hor * 2) + step * 2);
                ax = a.x; ay = a.y;
            }
            S.f.xin[0][t] = ox; S.f.xin[1][t] = oy; S.f.xin[2][t] = ax; S.f.xin[3][t] = ay;
        }
        // cost of (cur_obs, cur_acs): max(Q1, Q2)_risk (MPC.py:409, qrisk.py:184-196) -- before the transition
        mlp_tile_forward<PBM>(S.f, A.qr1, nullptr, nullptr, 0, PBM);
        float q1 = 0.f;
        if (t < PBM) q1 = sigmoidf_(S.f.raw[t][0]);
        __syncthreads();
        mlp_tile_forward<PBM>(S.f, A.qr2, nullptr, nullptr, 0, PBM);
        if (t < PBM) cost += fmaxf(q1, sigmoidf_(S.f.raw[t][0]));
        __syncthreads();
        // next observation through this tile's bootstrap net (MPC.py:421-439)
        if (t < PBM) {
            S.f.xin[0][t] = (ox - mu[0]) / sg[0]; S.f.xin[1][t] = (oy - mu[1]) / sg[1];
            S.f.xin[2][t] = (ax - mu[2]) / sg[2]; S.f.xin[3][t] = (ay - mu[3]) / sg[3];
        }
        ens_tile_forward(S, A.dyn, net);
        if (t < PBM) {
            const float4 rv = *reinterpret_cast<const float4*>(S.f.raw[t]);
            const float mx = A.dyn[kMaxLv], my = A.dyn[kMaxLv + 1], nx = A.dyn[kMinLv], ny = A.dyn[kMinLv + 1];
            float lvx = mx - softplusf(mx - rv.z), lvy = my - softplusf(my - rv.w);
            lvx = nx + softplusf(lvx - nx);
            lvy = ny + softplusf(lvy - ny);
            float ex = 0.f, ey = 0.f;
            if (live) {
                if (A.eps) {
                    const float2 ev = *reinterpret_cast<const float2*>(
                        A.eps + ((((size_t)e * A.hor + step) * NETS + net) * rows_net + j) * 2);
                    ex = ev.x; ey = ev.y;
                } else {
                    float ee[2];
                    const uint64_t idx = (((uint64_t)e * A.hor + step) * NETS + net) * (uint64_t)rows_net + j;
                    philox_eps(A.seed, A.stream_id, idx, vstep, RRL_DRAW_MPC_EPS + (uint32_t)A.iter * 16u, ee);
                    ex = ee[0]; ey = ee[1];
                }
            }
            ox = ox + fmaf(ex, sqrtf(expf(lvx)), rv.x);     // obs_postproc: obs + (mean + eps * sqrt(var))
            oy = oy + fmaf(ey, sqrtf(expf(lvy)), rv.y);
        }
        __syncthreads();
    }
    if (live) {
        const int p = net * A.npn + pl;
        A.row_cost[((size_t)e * A.pop + c) * A.npart + p] = (cost != cost) ? 1e6f : cost;   // costs[costs != costs] = 1e6
    }
}

// ---- CEM bookkeeping (optimizers.py:73-124), one CTA per env ----------------------------------------------
struct CemArgs {
    int64_t E;
    int pop, sol, hor, num_elites, npart, max_iters, iter;
    double alpha, epsilon;
    float lb[2], ub[2];
    double* mean;      // [E][sol]
    double* var;       // [E][sol]
    const double* z;   // NULL (Philox) or [E][pop][sol] draws in [-2, 2]
    float* samples;    // [E][pop][sol]
    const float* row_cost;  // [E][pop][npart]
    int32_t* active;   // [E]
    uint64_t seed;
    uint32_t stream_id;
    const int64_t* counters;
};

// truncnorm(-2, 2).rvs: inverse-CDF of a uniform, as scipy does
__device__ __forceinline__ double truncnorm_from_uniform(double u) {
    const double lo = normcdf(-2.0), hi = normcdf(2.0);
    double x = normcdfinv(lo + u * (hi - lo));
    return fmin(fmax(x, -2.0), 2.0);
}

__global__ void __launch_bounds__(256) cem_sample_kernel(const CemArgs A) {
    const int64_t e = blockIdx.x;
    __shared__ double s_cvar[64], s_mean[64];
    __shared__ int s_active;
    const int t = threadIdx.x;
    if (t == 0) {   // while (t < iters) and np.max(var) > epsilon
        double mv = -1.0;
        for (int d = 0; d < A.sol; ++d) mv = fmax(mv, A.var[e * A.sol + d]);
        s_active = (A.iter < A.max_iters && mv > A.epsilon) ? 1 : 0;
        if (A.iter > 0 && !A.active[e]) s_active = 0;      // the loop has already stopped for this env
        A.active[e] = s_active;
    }
    if (t < A.sol) {
        const double m = A.mean[e * A.sol + t], v = A.var[e * A.sol + t];
        const double lb = (double)A.lb[t & 1], ub = (double)A.ub[t & 1];
        const double lbd = m - lb, ubd = ub - m;
        s_cvar[t] = sqrt(fmin(fmin((lbd / 2) * (lbd / 2), (ubd / 2) * (ubd / 2)), v));
        s_mean[t] = m;
    }
    __syncthreads();
    if (!s_active) return;
    const uint64_t vstep = A.counters ? (uint64_t)A.counters[RRL_C_VEC_STEP] : 0;
    for (int i = t; i < A.pop * A.sol; i += 256) {
        const int d = i % A.sol;
        double zz;
        if (A.z) {
            zz = A.z[(size_t)e * A.pop * A.sol + i];
        } else {
            const Philox4 p = rrl_philox(A.seed, A.stream_id, (uint64_t)e * A.pop * A.sol + i, vstep,
                                         RRL_DRAW_MPC_Z + (uint32_t)A.iter * 16u);
            zz = truncnorm_from_uniform((rrl_u53(p.x, p.y) + 0x1p-54));
        }
        A.samples[(size_t)e * A.pop * A.sol + i] = (float)(zz * s_cvar[d] + s_mean[d]);   // .astype(np.float32)
    }
}

__global__ void __launch_bounds__(256) cem_update_kernel(const CemArgs A) {
    const int64_t e = blockIdx.x;
    if (!A.active[e]) return;
    extern __shared__ float sm[];
    float* cost = sm;                         // [pop]
    int* elite = reinterpret_cast<int*>(sm + A.pop);   // [num_elites] candidate index by rank
    const int t = threadIdx.x;
    for (int c = t; c < A.pop; c += 256) {    // costs.mean(dim=1) over the particles
        const float* rc = A.row_cost + ((size_t)e * A.pop + c) * A.npart;
        float s = 0.f;
        for (int p = 0; p < A.npart; ++p) s += rc[p];
        cost[c] = s / (float)A.npart;
    }
    __syncthreads();
    for (int c = t; c < A.pop; c += 256) {    // rank = position in np.argsort(costs) (ties by index)
        const float mine = cost[c];
        int rank = 0;
        for (int o = 0; o < A.pop; ++o) {
            const float v = cost[o];
            rank += (v < mine || (v == mine && o < c)) ? 1 : 0;
        }
        if (rank < A.num_elites) elite[rank] = c;
    }
    __syncthreads();
    if (t < A.sol) {   // np.mean / np.var of the elites along axis 0: float32, rows added in rank order
        const float* smp = A.samples + (size_t)e * A.pop * A.sol;
        float s = 0.f;
        for (int r = 0; r < A.num_elites; ++r) s += smp[(size_t)elite[r] * A.sol + t];
        const float mean = s / (float)A.num_elites;
        float q = 0.f;
        for (int r = 0; r < A.num_elites; ++r) {
            const float d = smp[(size_t)elite[r] * A.sol + t] - mean;
            q += d * d;
        }
        const float var = q / (float)A.num_elites;
        double& m = A.mean[e * A.sol + t];
        double& v = A.var[e * A.sol + t];
        m = A.alpha * m + (1.0 - A.alpha) * (double)mean;
        v = A.alpha * v + (1.0 - A.alpha) * (double)var;
    }
}

// MPC.act head / tail (MPC.py:337-347): mean <- prev_sol, var <- init_var ; action <- soln[:dU], prev_sol <- shift
struct ActIoArgs {
    int64_t E;
    int sol;
    float lb[2], ub[2];
    double *mean, *var, *prev_sol, *action;
    const uint8_t* mask;
    int32_t* active;
};
__global__ void mpc_begin_kernel(const ActIoArgs A) {
    const int64_t i = blockIdx.x * 256ll + threadIdx.x;
    if (i >= A.E * A.sol) return;
    const int d = (int)(i % A.sol);
    A.mean[i] = A.prev_sol[i];
    const double w = (double)(A.ub[d & 1] - A.lb[d & 1]);     // np.square(ac_ub - ac_lb) / 16 in float32
    const float wf = A.ub[d & 1] - A.lb[d & 1];
    A.var[i] = (double)((wf * wf) / 16.0f);
    (void)w;
    if (d == 0) A.active[i / A.sol] = 1;
}
__global__ void mpc_finish_kernel(const ActIoArgs A) {
    const int64_t i = blockIdx.x * 256ll + threadIdx.x;
    if (i >= A.E * A.sol) return;
    const int64_t e = i / A.sol;
    const int d = (int)(i % A.sol);
    const double v = A.mean[i];
    if (d < 2) A.action[e * 2 + d] = v;
    if (A.mask && !A.mask[e]) return;          // only the envs that really planned shift their warm start
    // prev_sol = concat(soln[per*dU:], zeros(per*dU)), per = 1
    A.prev_sol[i] = d + 2 < A.sol ? A.mean[i + 2] : 0.0;
}

int check_mpc(const rrl_mpc_config_t* c) {
    if (!c) { rrl_set_error("null mpc config"); return -2; }
    if (c->num_nets != NETS) { rrl_set_error("the planner is specialised for 5 bootstrap nets (got %d)", c->num_nets); return -2; }
    if (c->npart <= 0 || c->npart % NETS != 0) { rrl_set_error("npart must be a positive multiple of 5 (MPC.py:160)"); return -2; }
    if (c->plan_hor <= 0 || c->plan_hor * 2 > 64) { rrl_set_error("plan_hor must be in [1, 32]"); return -2; }
    if (c->popsize <= 0 || c->num_elites <= 0 || c->num_elites > c->popsize) {
        rrl_set_error("need 0 < num_elites <= popsize (optimizers.py:64-66)");
        return -2;
    }
    return 0;
}

CemArgs cem_args(const rrl_mpc_config_t* c, int64_t E, int iter) {
    CemArgs A;
    memset(&A, 0, sizeof(A));
    A.E = E; A.pop = c->popsize; A.sol = c->plan_hor * 2; A.hor = c->plan_hor; A.num_elites = c->num_elites;
    A.npart = c->npart; A.max_iters = c->max_iters; A.iter = iter; A.alpha = c->alpha; A.epsilon = c->epsilon;
    A.lb[0] = c->ac_lb[0]; A.lb[1] = c->ac_lb[1]; A.ub[0] = c->ac_ub[0]; A.ub[1] = c->ac_ub[1];
    A.seed = c->seed; A.stream_id = (uint32_t)c->stream_id;
    return A;
}

}  // namespace

extern "C" int64_t rrl_dyn_image_floats(void) { return kDynFloats; }

extern "C" int rrl_dyn_pack(const float* lin0_w, const float* lin0_b, const float* lin1_w, const float* lin1_b,
                            const float* lin2_w, const float* lin2_b, const float* lin3_w, const float* lin3_b,
                            const float* inputs_mu, const float* inputs_sigma, const float* max_logvar,
                            const float* min_logvar, int hidden, float* image, void* stream) {
    RRL_CHECK_ARG(lin0_w && lin0_b && lin1_w && lin1_b && lin2_w && lin2_b && lin3_w && lin3_b && inputs_mu && inputs_sigma &&
                      max_logvar && min_logvar && image, "null argument");
    RRL_CHECK_ARG(hidden > 0 && hidden <= H, "hidden width must be in [1, 256]");
    PackArgs A = {lin0_w, lin0_b, lin1_w, lin1_b, lin2_w, lin2_b, lin3_w, lin3_b, inputs_mu, inputs_sigma, max_logvar,
                  min_logvar, image, hidden};
    dyn_pack_kernel<<<(unsigned)((kDynSimtFloats + 255) / 256), 256, 0, (cudaStream_t)stream>>>(A);
    RRL_CHECK_LAUNCH();
    return rrl::dyn_tc_images_launch(image, (cudaStream_t)stream);   // fp16 hi/lo operand images of lin1 / lin2 (mpc_tc.cu)
}

extern "C" int rrl_mpc_begin(const rrl_mpc_config_t* cfg, int64_t n_envs, const double* prev_sol, double* mean, double* var,
                             int32_t* active, void* stream) {
    int rc = check_mpc(cfg);
    if (rc) return rc;
    RRL_CHECK_ARG(n_envs > 0 && prev_sol && mean && var && active, "bad argument");
    ActIoArgs A;
    memset(&A, 0, sizeof(A));
    A.E = n_envs; A.sol = cfg->plan_hor * 2;
    A.lb[0] = cfg->ac_lb[0]; A.lb[1] = cfg->ac_lb[1]; A.ub[0] = cfg->ac_ub[0]; A.ub[1] = cfg->ac_ub[1];
    A.mean = mean; A.var = var; A.prev_sol = const_cast<double*>(prev_sol); A.active = active;
    mpc_begin_kernel<<<(unsigned)((n_envs * A.sol + 255) / 256), 256, 0, (cudaStream_t)stream>>>(A);
    RRL_CHECK_LAUNCH();
    return 0;
}

extern "C" int rrl_mpc_sample(const rrl_mpc_config_t* cfg, int64_t n_envs, int iter, const double* mean, const double* var,
                              const double* z, const int64_t* counters, float* samples, int32_t* active, void* stream) {
    int rc = check_mpc(cfg);
    if (rc) return rc;
    RRL_CHECK_ARG(n_envs > 0 && mean && var && samples && active, "bad argument");
    CemArgs A = cem_args(cfg, n_envs, iter);
    A.mean = const_cast<double*>(mean); A.var = const_cast<double*>(var); A.z = z; A.samples = samples; A.active = active;
    A.counters = counters;
    cem_sample_kernel<<<(unsigned)n_envs, 256, 0, (cudaStream_t)stream>>>(A);
    RRL_CHECK_LAUNCH();
    return 0;
}

extern "C" int rrl_mpc_rollout(const rrl_mpc_config_t* cfg, const rrl_agent_config_t* agent_cfg, const float* arena,
                               const float* dyn_image, int64_t n_envs, const double* state, const float* samples,
                               const float* eps, const int32_t* active, int iter, const int64_t* counters, float* row_cost,
                               void* stream) {
    int rc = check_mpc(cfg);
    if (rc) return rc;
    RRL_CHECK_ARG(agent_cfg && arena && dyn_image && n_envs > 0 && state && samples && row_cost, "bad argument");
    RRL_CHECK_ARG(agent_cfg->hidden == H, "kernels are specialised for hidden_size 256");
    const Layout L = make_layout(agent_cfg);
    RolloutArgs A;
    memset(&A, 0, sizeof(A));
    A.dyn = dyn_image;
    A.qr1 = head_w(L, arena, RRL_NET_QRISK, 0);
    A.qr2 = head_w(L, arena, RRL_NET_QRISK, 1);
    A.state = state; A.samples = samples; A.eps = eps; A.active = active; A.row_cost = row_cost;
    A.E = n_envs; A.pop = cfg->popsize; A.hor = cfg->plan_hor; A.npart = cfg->npart; A.npn = cfg->npart / NETS;
    A.tiles_per_net = (A.pop * A.npn + PBM - 1) / PBM;
    A.seed = cfg->seed; A.stream_id = (uint32_t)cfg->stream_id; A.counters = counters; A.iter = iter;
    if (agent_cfg->use_tensor_cores) {   // tcgen05 planner (mpc_tc.cu): same rows, same noise indexing
        rrl::MpcTcArgs T;
        T.dyn = dyn_image; T.qr1 = A.qr1; T.qr2 = A.qr2; T.state = state; T.samples = samples; T.eps = eps; T.active = active;
        T.row_cost = row_cost; T.E = n_envs; T.pop = A.pop; T.hor = A.hor; T.npart = A.npart; T.npn = A.npn;
        T.seed = A.seed; T.stream_id = A.stream_id; T.counters = counters; T.iter = iter;
        return rrl::mpc_rollout_tc_launch(T, (cudaStream_t)stream);
    }
    static bool configured = false;
    const size_t smem = sizeof(PlanSmem);
    if (!configured) {
        RRL_CUDA(cudaFuncSetAttribute(mpc_rollout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const int64_t grid = n_envs * NETS * A.tiles_per_net;
    RRL_CHECK_ARG(grid < (1ll << 31), "too many planner tiles for one launch");
    mpc_rollout_kernel<<<(unsigned)grid, kThreads, smem, (cudaStream_t)stream>>>(A);
    RRL_CHECK_LAUNCH();
    return 0;
}

extern "C" int rrl_mpc_update(const rrl_mpc_config_t* cfg, int64_t n_envs, int iter, const float* samples,
                              const float* row_cost, const int32_t* active, double* mean, double* var, void* stream) {
    int rc = check_mpc(cfg);
    if (rc) return rc;
    RRL_CHECK_ARG(n_envs > 0 && samples && row_cost && active && mean && var, "bad argument");
    CemArgs A = cem_args(cfg, n_envs, iter);
    A.mean = mean; A.var = var; A.samples = const_cast<float*>(samples); A.row_cost = row_cost;
    A.active = const_cast<int32_t*>(active);
    const size_t smem = (size_t)(cfg->popsize + cfg->num_elites) * 4;
    RRL_CHECK_ARG(smem <= 48 * 1024, "popsize too large for the elite-selection kernel");
    cem_update_kernel<<<(unsigned)n_envs, 256, smem, (cudaStream_t)stream>>>(A);
    RRL_CHECK_LAUNCH();
    return 0;
}

extern "C" int rrl_mpc_finish(const rrl_mpc_config_t* cfg, int64_t n_envs, const double* mean, const uint8_t* mask,
                              double* prev_sol, double* action, void* stream) {
    int rc = check_mpc(cfg);
    if (rc) return rc;
    RRL_CHECK_ARG(n_envs > 0 && mean && prev_sol && action, "bad argument");
    ActIoArgs A;
    memset(&A, 0, sizeof(A));
    A.E = n_envs; A.sol = cfg->plan_hor * 2;
    A.mean = const_cast<double*>(mean); A.prev_sol = prev_sol; A.action = action; A.mask = mask;
    mpc_finish_kernel<<<(unsigned)((n_envs * A.sol + 255) / 256), 256, 0, (cudaStream_t)stream>>>(A);
    RRL_CHECK_LAUNCH();
    return 0;
}

// =====================================================================================================================
// Ensemble training step (MPC.train, recovery_rl/MPC.py:268-296; PtModel.forward / compute_decays, config/maze.py:52-96):
// ONE launch per mini-batch does, for all five bootstrap nets, the forward pass on the net's own bootstrap batch, the
// Gaussian NLL + log-variance-bound + weight-decay loss, the backward pass and the torch-Adam update of every parameter.
//   * one CTA per net (the nets share nothing but max/min_logvar, whose gradient partials are summed, in net order, by
//     the last CTA to finish); activations of the <= 32-row batch stay in shared memory;
//   * thread j owns output column j: forward  z[:, j] = sum_k in[:, k] W[k][j]  (W rows are contiguous: coalesced),
//     weight gradient dW[k][j] = sum_r in[r][k] dz[r][j] computed straight into the Adam update (never materialised),
//     input gradient through a maintained transposed copy W^T so that it has the same coalesced access pattern.
// Parameters live in a flat "train arena" in the reference's layout (the torch PtModel parameters are views of it).
// =====================================================================================================================
namespace {

constexpr int TB = 32;        // MPC.py:263 batch_size
constexpr int HID = 200;      // hidden width of the reference ensemble
constexpr int kTrainThreads = 256;

struct DynTrainArgs {
    float* P;            // params:  [w0 5x4x200][b0 5x200][w1 5x200x200][b1][w2][b2][w3 5x200x4][b3 5x4][max_lv 2][min_lv 2]
    float* M;            // Adam exp_avg, same layout
    float* V;            // Adam exp_avg_sq
    float* WT;           // transposed copies: [w1T 5x200x200][w2T 5x200x200]
    float* partial;      // [5][4] gradient partials of (max_lv[2], min_lv[2]) + [5] loss partials at offset 20
    const float* mu;     // [4] input normalisation
    const float* sigma;  // [4]
    const float* inputs; // all training inputs  [n_data][4]
    const float* targets;// all training targets [n_data][2]
    const int64_t* idx;  // bootstrap indices [5][n_idx]
    int64_t n_idx;       // columns of idx
    int64_t col0;        // first column of this batch
    int rows;            // rows of this batch (<= 32)
    float lr, b1, b2, eps;
    int64_t* step;       // device: Adam step count (incremented by the last CTA)
    unsigned int* ticket;
    float* loss_out;     // optional: total loss of this batch
};

constexpr int64_t oW0 = 0, oB0 = oW0 + 5 * 4 * HID, oW1 = oB0 + 5 * HID, oB1 = oW1 + 5 * HID * HID, oW2 = oB1 + 5 * HID,
                  oB2 = oW2 + 5 * HID * HID, oW3 = oB2 + 5 * HID, oB3 = oW3 + 5 * HID * 4, oMax = oB3 + 5 * 4, oMin = oMax + 2,
                  kTrainFloats = oMin + 2;

struct TrainSmem {
    float x[TB][4];
    float z[3][TB][HID];      // pre-activations of the three hidden layers
    float a[TB][HID];         // current activation (input of the layer being processed)
    float d[TB][HID];         // current dz
    float out[TB][4], dout[TB][4];
    float red[8][8];
};

__device__ __forceinline__ float sigm(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float dswish(float z) { const float s = sigm(z); return s * (1.0f + z * (1.0f - s)); }

struct AdamC { float lr_bc1, bc2s, b1, b2, eps; };
__device__ __forceinline__ float adam_apply(float* P, float* M, float* V, int64_t o, float g, const AdamC& c) {
    float m = M[o], v = V[o];
    m = m + (g - m) * (1.0f - c.b1);
    v = v * c.b2 + (1.0f - c.b2) * g * g;
    const float p = P[o] - c.lr_bc1 * (m / (sqrtf(v) / c.bc2s + c.eps));
    M[o] = m; V[o] = v; P[o] = p;
    return p;
}

__global__ void __launch_bounds__(kTrainThreads, 1) dyn_train_kernel(const DynTrainArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TrainSmem& S = *reinterpret_cast<TrainSmem*>(smem_raw);
    const int net = blockIdx.x, t = threadIdx.x, R = A.rows;
    const float invR2 = 1.0f / (2.0f * (float)R);          // .mean(-1).mean(-1): 2 output dims x R rows
    const double tstep = (double)(*A.step + 1);
    AdamC ad;
    ad.lr_bc1 = (float)((double)A.lr / (1.0 - pow((double)A.b1, tstep)));
    ad.bc2s = (float)sqrt(1.0 - pow((double)A.b2, tstep));
    ad.b1 = A.b1; ad.b2 = A.b2; ad.eps = A.eps;
    float* W0 = A.P + oW0 + net * 4 * HID; float* B0 = A.P + oB0 + net * HID;
    float* W1 = A.P + oW1 + (int64_t)net * HID * HID; float* B1 = A.P + oB1 + net * HID;
    float* W2 = A.P + oW2 + (int64_t)net * HID * HID; float* B2 = A.P + oB2 + net * HID;
    float* W3 = A.P + oW3 + net * HID * 4; float* B3 = A.P + oB3 + net * 4;
    float* W1T = A.WT + (int64_t)net * HID * HID; float* W2T = A.WT + (int64_t)(5 + net) * HID * HID;
    const float mxl[2] = {A.P[oMax], A.P[oMax + 1]}, mnl[2] = {A.P[oMin], A.P[oMin + 1]};
    // ---- gather + normalise the batch (MPC.py:283-286, config/maze.py:74) ----
    if (t < TB * 4) {
        const int r = t >> 2, i = t & 3;
        float v = 0.f;
        if (r < R) {
            const int64_t row = A.idx[(int64_t)net * A.n_idx + A.col0 + r];
            v = (A.inputs[row * 4 + i] - A.mu[i]) / A.sigma[i];
        }
        S.x[r][i] = v;
    }
    __syncthreads();
    const int j = t;                 // output column owned by this thread (t < HID)
    // ---- forward ----
    if (j < HID) {
        const float w0 = W0[j], w1 = W0[HID + j], w2 = W0[2 * HID + j], w3 = W0[3 * HID + j], b = B0[j];
        for (int r = 0; r < TB; ++r) {
            const float z = fmaf(S.x[r][3], w3, fmaf(S.x[r][2], w2, fmaf(S.x[r][1], w1, fmaf(S.x[r][0], w0, b))));
            S.z[0][r][j] = z;
            S.a[r][j] = swishf(z);
        }
    }
    __syncthreads();
    for (int l = 1; l <= 2; ++l) {
        const float* W = l == 1 ? W1 : W2;
        const float* Bv = l == 1 ? B1 : B2;
        float acc[TB];
        if (j < HID) {
#pragma unroll
            for (int r = 0; r < TB; ++r) acc[r] = 0.f;
            for (int k = 0; k < HID; k += 4) {
                const float wa = W[(k + 0) * HID + j], wb = W[(k + 1) * HID + j], wc = W[(k + 2) * HID + j], wd = W[(k + 3) * HID + j];
#pragma unroll
                for (int r = 0; r < TB; ++r) {
                    const float4 av = *reinterpret_cast<const float4*>(&S.a[r][k]);
                    acc[r] = fmaf(av.w, wd, fmaf(av.z, wc, fmaf(av.y, wb, fmaf(av.x, wa, acc[r]))));
                }
            }
        }
        __syncthreads();             // everyone has finished reading S.a
        if (j < HID) {
            const float b = Bv[j];
#pragma unroll
            for (int r = 0; r < TB; ++r) {
                const float z = acc[r] + b;
                S.z[l][r][j] = z;
                S.a[r][j] = swishf(z);
            }
        }
        __syncthreads();
    }
    // output layer (4 outputs) + loss gradients; thread (r, o)
    if (t < TB * 4) {
        const int r = t >> 2, o = t & 3;
        float acc = B3[o];
        for (int k = 0; k < HID; ++k) acc = fmaf(S.a[r][k], W3[k * 4 + o], acc);
        S.out[r][o] = acc;
    }
    __syncthreads();
    if (t < TB * 2) {                // thread (r, dim): mean / log-variance pair of one output dimension
        const int r = t >> 1, dm = t & 1;
        float dmean = 0.f, draw = 0.f;
        if (r < R) {
            const int64_t row = A.idx[(int64_t)net * A.n_idx + A.col0 + r];
            const float mean = S.out[r][dm], raw = S.out[r][2 + dm], targ = A.targets[row * 2 + dm];
            const float lv1 = mxl[dm] - softplusf(mxl[dm] - raw);
            const float lv = mnl[dm] + softplusf(lv1 - mnl[dm]);
            const float inv_var = expf(-lv), e = mean - targ;
            dmean = 2.0f * e * inv_var * invR2;
            const float dlv = (1.0f - e * e * inv_var) * invR2;
            const float s_min = sigm(lv1 - mnl[dm]);       // d lv / d lv1
            const float s_max = sigm(mxl[dm] - raw);       // d lv1 / d raw
            draw = dlv * s_min * s_max;
        }
        S.dout[r][dm] = dmean;
        S.dout[r][2 + dm] = draw;
    }
    __syncthreads();
    // per-net partials of the shared log-variance bounds and of the loss (fixed order: deterministic)
    if (t == 0) {
        float g[4] = {0.f, 0.f, 0.f, 0.f}, ls = 0.f;
        for (int r = 0; r < R; ++r)
            for (int dm = 0; dm < 2; ++dm) {
                const int64_t row = A.idx[(int64_t)net * A.n_idx + A.col0 + r];
                const float mean = S.out[r][dm], raw = S.out[r][2 + dm], targ = A.targets[row * 2 + dm];
                const float lv1 = mxl[dm] - softplusf(mxl[dm] - raw);
                const float lv = mnl[dm] + softplusf(lv1 - mnl[dm]);
                const float inv_var = expf(-lv), e = mean - targ;
                const float dlv = (1.0f - e * e * inv_var) * invR2;
                const float s_min = sigm(lv1 - mnl[dm]), s_max = sigm(mxl[dm] - raw);
                g[dm] += dlv * s_min * (1.0f - s_max);      // d loss / d max_logvar[dm]
                g[2 + dm] += dlv * (1.0f - s_min);          // d loss / d min_logvar[dm]
                ls += (e * e * inv_var + lv) * invR2;
            }
        for (int i = 0; i < 4; ++i) A.partial[net * 4 + i] = g[i];
        A.partial[20 + net] = ls;
    }
    // ---- backward: layer 3 ----
    // da2[r][k] = sum_o dout[r][o] W3[k][o];  dz2 = da2 * swish'(z2)   (old W3), then dW3 / db3 -> Adam
    if (j < HID) {
        const float4 w = *reinterpret_cast<const float4*>(W3 + j * 4);
        float g[4] = {0.f, 0.f, 0.f, 0.f};
        for (int r = 0; r < TB; ++r) {
            const float4 dv = *reinterpret_cast<const float4*>(S.dout[r]);
            const float da = dv.x * w.x + dv.y * w.y + dv.z * w.z + dv.w * w.w;
            const float a2 = S.a[r][j];
            g[0] = fmaf(a2, dv.x, g[0]); g[1] = fmaf(a2, dv.y, g[1]); g[2] = fmaf(a2, dv.z, g[2]); g[3] = fmaf(a2, dv.w, g[3]);
            S.d[r][j] = da * dswish(S.z[2][r][j]);
        }
        const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int o = 0; o < 4; ++o) adam_apply(A.P, A.M, A.V, oW3 + net * HID * 4 + j * 4 + o, g[o] + 0.00075f * wv[o], ad);
    }
    if (t >= 224 && t < 228) {       // db3
        const int o = t - 224;
        float g = 0.f;
        for (int r = 0; r < TB; ++r) g += S.dout[r][o];
        adam_apply(A.P, A.M, A.V, oB3 + net * 4 + o, g, ad);
    }
    __syncthreads();
    // ---- backward: layers 2 and 1 ----
    for (int l = 2; l >= 1; --l) {
        float* W = l == 2 ? W2 : W1;
        float* WT = l == 2 ? W2T : W1T;
        const int64_t oW = (l == 2 ? oW2 : oW1) + (int64_t)net * HID * HID, oB = (l == 2 ? oB2 : oB1) + net * HID;
        // input of layer l = swish(z[l-1]); rebuild it in S.a
        if (j < HID)
            for (int r = 0; r < TB; ++r) S.a[r][j] = swishf(S.z[l - 1][r][j]);
        // da[r][k] = sum_j dz[r][j] W[k][j] = sum_j dz[r][j] WT[j][k]: thread k, coalesced over WT rows
        float acc[TB];
        if (j < HID) {
#pragma unroll
            for (int r = 0; r < TB; ++r) acc[r] = 0.f;
            for (int q = 0; q < HID; q += 4) {
                const float wa = WT[(q + 0) * HID + j], wb = WT[(q + 1) * HID + j], wc = WT[(q + 2) * HID + j], wd = WT[(q + 3) * HID + j];
#pragma unroll
                for (int r = 0; r < TB; ++r) {
                    const float4 dv = *reinterpret_cast<const float4*>(&S.d[r][q]);
                    acc[r] = fmaf(dv.w, wd, fmaf(dv.z, wc, fmaf(dv.y, wb, fmaf(dv.x, wa, acc[r]))));
                }
            }
        }
        __syncthreads();             // S.a rebuilt; every thread has its da column in registers
        // dW[k][j] = sum_r a[r][k] dz[r][j] (+ decay) -> Adam on W[k][j] and its transposed copy; db[j] = sum_r dz[r][j]
        if (j < HID) {
            float dz[TB], gb = 0.f;
#pragma unroll
            for (int r = 0; r < TB; ++r) { dz[r] = S.d[r][j]; gb += dz[r]; }
            for (int k = 0; k < HID; ++k) {
                float g = 0.f;
#pragma unroll
                for (int r = 0; r < TB; ++r) g = fmaf(S.a[r][k], dz[r], g);
                const float w = W[k * HID + j];
                const float p = adam_apply(A.P, A.M, A.V, oW + (int64_t)k * HID + j, g + 0.0005f * w, ad);
                WT[(int64_t)j * HID + k] = p;
            }
            adam_apply(A.P, A.M, A.V, oB + j, gb, ad);
        }
        __syncthreads();             // all reads of S.d done
        if (j < HID)
#pragma unroll
            for (int r = 0; r < TB; ++r) S.d[r][j] = acc[r] * dswish(S.z[l - 1][r][j]);
        __syncthreads();
    }
    // ---- backward: layer 0 (K = 4) ----
    if (j < HID) {
        float g[4] = {0.f, 0.f, 0.f, 0.f}, gb = 0.f;
        for (int r = 0; r < TB; ++r) {
            const float dz = S.d[r][j];
            g[0] = fmaf(S.x[r][0], dz, g[0]); g[1] = fmaf(S.x[r][1], dz, g[1]);
            g[2] = fmaf(S.x[r][2], dz, g[2]); g[3] = fmaf(S.x[r][3], dz, g[3]);
            gb += dz;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t o = oW0 + net * 4 * HID + i * HID + j;
            adam_apply(A.P, A.M, A.V, o, g[i] + 0.00025f * A.P[o], ad);
        }
        adam_apply(A.P, A.M, A.V, oB0 + net * HID + j, gb, ad);
    }
    // ---- shared log-variance bounds + step count: last CTA to finish ----
    __syncthreads();
    if (t == 0) {
        __threadfence();
        const unsigned int tk = atomicAdd(A.ticket, 1u);
        if (tk == gridDim.x - 1) {
            __threadfence();
            float g[4] = {0.01f, 0.01f, -0.01f, -0.01f};     // loss = 0.01 * (max_logvar.sum() - min_logvar.sum()) + ...
            float ls = 0.f;
            for (int n = 0; n < 5; ++n) {
                for (int i = 0; i < 4; ++i) g[i] += A.partial[n * 4 + i];
                ls += A.partial[20 + n];
            }
            for (int i = 0; i < 4; ++i) adam_apply(A.P, A.M, A.V, oMax + i, g[i], ad);
            if (A.loss_out) *A.loss_out = ls;
            *A.step += 1;
            *A.ticket = 0;
        }
    }
}

}  // namespace

extern "C" int64_t rrl_dyn_train_floats(void) { return kTrainFloats; }

extern "C" int rrl_dyn_train_step(float* params, float* adam_m, float* adam_v, float* wt, float* partial, const float* mu,
                                  const float* sigma, const float* inputs, const float* targets, const int64_t* idx,
                                  int64_t n_idx, int64_t col0, int rows, float lr, int64_t* step, uint32_t* ticket,
                                  float* loss_out, void* stream) {
    RRL_CHECK_ARG(params && adam_m && adam_v && wt && partial && mu && sigma && inputs && targets && idx && step && ticket,
                  "null argument");
    RRL_CHECK_ARG(rows > 0 && rows <= TB && col0 >= 0 && col0 + rows <= n_idx, "bad batch window");
    DynTrainArgs A;
    A.P = params; A.M = adam_m; A.V = adam_v; A.WT = wt; A.partial = partial; A.mu = mu; A.sigma = sigma;
    A.inputs = inputs; A.targets = targets; A.idx = idx; A.n_idx = n_idx; A.col0 = col0; A.rows = rows;
    A.lr = lr; A.b1 = 0.9f; A.b2 = 0.999f; A.eps = 1e-8f; A.step = step; A.ticket = ticket; A.loss_out = loss_out;
    static bool configured = false;
    const size_t smem = sizeof(TrainSmem);
    if (!configured) {
        RRL_CUDA(cudaFuncSetAttribute(dyn_train_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    dyn_train_kernel<<<5, kTrainThreads, smem, (cudaStream_t)stream>>>(A);
    RRL_CHECK_LAUNCH();
    return 0;
}

// transposed copies of lin1_w / lin2_w (operand of the input-gradient pass); call after the host wrote the parameters
namespace {
__global__ void dyn_transpose_kernel(const float* __restrict__ P, float* __restrict__ WT) {
    const int64_t i = blockIdx.x * 256ll + threadIdx.x;
    if (i >= 2ll * 5 * HID * HID) return;
    const int l = (int)(i / (5ll * HID * HID));
    const int64_t rem = i % (5ll * HID * HID);
    const int net = (int)(rem / (HID * HID)), k = (int)((rem % (HID * HID)) / HID), j = (int)(rem % HID);
    WT[(int64_t)(l * 5 + net) * HID * HID + (int64_t)j * HID + k] = P[(l == 0 ? oW1 : oW2) + (int64_t)net * HID * HID + (int64_t)k * HID + j];
}
}  // namespace
extern "C" int rrl_dyn_train_sync(const float* params, float* wt, void* stream) {
    RRL_CHECK_ARG(params && wt, "null argument");
    dyn_transpose_kernel<<<(unsigned)((2ll * 5 * HID * HID + 255) / 256), 256, 0, (cudaStream_t)stream>>>(params, wt);
    RRL_CHECK_LAUNCH();
    return 0;
}
