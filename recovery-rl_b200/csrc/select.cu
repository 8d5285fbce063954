// select.cu -- the comparison branches of the vectorised acting path and the add_both_transitions push:
//   rrl_sqrl_select_action      SAC.select_action with --use_constraint_sampling      (sac.py:139-161)
//   rrl_qsample_recovery_action QRiskWrapper.select_action with --Q_sampling_recovery (qrisk.py:214-225)
//   rrl_replay_push_both        the second task-buffer push of --add_both_transitions (experiment.py:446-448)
// Every env copy evaluates its `samples` candidate actions through the ordinary batched forward passes
// (rrl_policy_sample / rrl_twin_q_forward over env-major candidate rows); the kernels here only expand the states,
// draw the candidates and pick.  None of this is on the bench path (the scripts' SQRL lines are comparison runs).
#include "common.cuh"

namespace {
constexpr int kSelThreads = 256;

// one candidate row per thread: fp32 state copy (torch.FloatTensor(state).repeat(samples, 1), sac.py:137-142) and the
// candidate's draw: N(0,1) pair (SQRL) or a uniform action in the Box (Q-sampling, gym Box.sample)
struct ExpandArgs {
    int64_t n, env0, envs;  // all env copies / first env of this chunk / envs in this chunk
    int samples;
    const double* state;    // [2][n]
    const float* draws_in;  // [n][samples][2] caller-supplied draws or NULL (Philox)
    int uniform_actions;    // 0: eps ~ N(0,1) -> eps_out;  1: a = (2u - 1) * scale + bias -> eps_out
    float scale[2], bias[2];
    uint64_t seed;
    uint32_t stream_id;
    const int64_t* counters;
    float* s_rep;    // [envs * samples][2]
    float* eps_out;  // [envs * samples][2]
};
__global__ void __launch_bounds__(kSelThreads) expand_kernel(const ExpandArgs A) {
    const int64_t rows = A.envs * (int64_t)A.samples;
    const uint64_t vstep = A.counters ? (uint64_t)A.counters[RRL_C_VEC_STEP] : 0;
    for (int64_t i = (int64_t)blockIdx.x * kSelThreads + threadIdx.x; i < rows; i += (int64_t)gridDim.x * kSelThreads) {
        const int64_t env = A.env0 + i / A.samples;
        const int64_t g = env * A.samples + i % A.samples;  // candidate index over ALL envs (chunking does not change the draws)
        reinterpret_cast<float2*>(A.s_rep)[i] = make_float2((float)A.state[env], (float)A.state[A.n + env]);
        float d0, d1;
        if (A.draws_in) {
            const float2 v = reinterpret_cast<const float2*>(A.draws_in)[g];
            d0 = v.x; d1 = v.y;
        } else {
            const Philox4 p = rrl_philox(A.seed, A.stream_id, (uint64_t)g, vstep, A.uniform_actions ? RRL_DRAW_QSAMPLE : RRL_DRAW_SQRL_EPS);
            if (A.uniform_actions) { d0 = rrl_u24(p.x); d1 = rrl_u24(p.y); }
            else rrl_normal2_f32(p.x, p.y, &d0, &d1);
        }
        if (A.uniform_actions) {
            d0 = fmaf(2.0f * d0 - 1.0f, A.scale[0], A.bias[0]);
            d1 = fmaf(2.0f * d1 - 1.0f, A.scale[1], A.bias[1]);
        }
        reinterpret_cast<float2*>(A.eps_out)[i] = make_float2(d0, d1);
    }
}

struct PickArgs {
    int64_t n, env0, envs;
    int samples;
    int mode;               // 0: SQRL filter + Categorical (sac.py:146-159);  1: argmin over all candidates (qrisk.py:222-224)
    const float *q1, *q2;   // [envs * samples]
    const float* logp;      // [envs * samples]   (mode 0)
    const float* cand;      // [envs * samples][2]
    const float* cat_u;     // [n] U[0,1) or NULL (Philox)   (mode 0)
    const float* rand_u;    // [n][2] U[0,1) or NULL (Philox): actions of the random start phase (mode 0)
    const uint8_t* only;    // mode 1: pick only for envs whose flag is set (the recovery flags), NULL: all
    int64_t start_steps;
    uint64_t seed;
    uint32_t stream_id;
    const int64_t* counters;
    float eps_safe;
    float scale[2], bias[2];
    float *action_task, *action_real, *qrisk_out;  // [n][2], [n][2], [n]; any may be NULL
    uint8_t* recovery;                              // mode 0: cleared
};
// one warp per env copy; lane l owns the contiguous candidates [l * per, (l + 1) * per) so that the order of the filtered
// list is the order of the samples
__global__ void __launch_bounds__(kSelThreads) pick_kernel(const PickArgs A) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * kSelThreads + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * kSelThreads) >> 5;
    const uint64_t vstep = A.counters ? (uint64_t)A.counters[RRL_C_VEC_STEP] : 0;
    const bool random_phase = A.mode == 0 && A.counters && (A.start_steps > A.counters[RRL_C_TOTAL_NUMSTEPS]);
    const int per = (A.samples + 31) / 32;
    for (int64_t e = warp; e < A.envs; e += n_warps) {
        const int64_t env = A.env0 + e;
        if (random_phase) {  // env.action_space.sample() (experiment.py:559-560), the same draw as the acting kernels'
            if (lane == 0) {
                float u0, u1;
                if (A.rand_u) {
                    const float2 uv = reinterpret_cast<const float2*>(A.rand_u)[env];
                    u0 = uv.x; u1 = uv.y;
                } else {
                    const Philox4 p = rrl_philox(A.seed, A.stream_id, (uint64_t)env, vstep, RRL_DRAW_ACT_RAND);
                    u0 = rrl_u24(p.x); u1 = rrl_u24(p.y);
                }
                const float2 a = make_float2(fmaf(2.0f * u0 - 1.0f, A.scale[0], A.bias[0]), fmaf(2.0f * u1 - 1.0f, A.scale[1], A.bias[1]));
                if (A.action_task) reinterpret_cast<float2*>(A.action_task)[env] = a;
                if (A.action_real) reinterpret_cast<float2*>(A.action_real)[env] = a;
                if (A.recovery) A.recovery[env] = 0;
                if (A.qrisk_out) A.qrisk_out[env] = 0.f;
            }
            continue;
        }
        if (A.mode == 1 && A.only && !A.only[env]) continue;  // warp-uniform
        const int64_t base = e * A.samples;
        const int j0 = lane * per, j1 = min(j0 + per, A.samples);
        int cnt = 0, minj = 0x7fffffff;
        float sum = 0.f, minq = __int_as_float(0x7f800000);
        for (int j = j0; j < j1; ++j) {
            const float q = fmaxf(A.q1[base + j], A.q2[base + j]);  // qrisk.py:196
            if (q < minq) { minq = q; minj = j; }
            if (A.mode == 0 && q <= A.eps_safe) {                    // sac.py:146
                cnt += 1;
                sum += expf(A.logp[base + j]);                       // sac.py:149
            }
        }
        // argmin over the warp (first index on ties)
        for (int o = 16; o; o >>= 1) {
            const float oq = __shfl_xor_sync(0xffffffffu, minq, o);
            const int oj = __shfl_xor_sync(0xffffffffu, minj, o);
            if (oq < minq || (oq == minq && oj < minj)) { minq = oq; minj = oj; }
        }
        int chosen = minj;  // sac.py:153-154 (nothing passes the filter) / qrisk.py:223
        if (A.mode == 0) {
            int icnt = cnt;
            float isum = sum;
            for (int o = 1; o < 32; o <<= 1) {  // inclusive scans in lane order
                const int c = __shfl_up_sync(0xffffffffu, icnt, o);
                const float s = __shfl_up_sync(0xffffffffu, isum, o);
                if (lane >= o) { icnt += c; isum += s; }
            }
            const int total = __shfl_sync(0xffffffffu, icnt, 31);
            const float total_p = __shfl_sync(0xffffffffu, isum, 31);
            if (total > 0) {
                // Categorical(probs).sample() (sac.py:156-157) as an inverse-CDF draw over the filtered list
                float u;
                if (A.cat_u) u = A.cat_u[env];
                else u = rrl_u24(rrl_philox(A.seed, A.stream_id, (uint64_t)env, vstep, RRL_DRAW_SQRL_CAT).x);
                const float target = u * total_p;
                float run = isum - sum;
                int rank = icnt - cnt, found = -1;
                for (int j = j0; j < j1; ++j) {
                    const float q = fmaxf(A.q1[base + j], A.q2[base + j]);
                    if (q <= A.eps_safe) {
                        run += expf(A.logp[base + j]);
                        if (found < 0 && run > target) found = rank;
                        rank += 1;
                    }
                }
                const unsigned hit = __ballot_sync(0xffffffffu, found >= 0);
                int r = total - 1;
                if (hit) r = __shfl_sync(0xffffffffu, found, __ffs(hit) - 1);
                // sac.py:158 indexes `pi` with the index INTO THE FILTERED SET (not thresh_idxs[sampled_idx]): as written
                chosen = r;
            }
        }
        if (lane == 0) {
            const float2 a = reinterpret_cast<const float2*>(A.cand)[base + chosen];
            if (A.action_task) reinterpret_cast<float2*>(A.action_task)[env] = a;
            if (A.action_real) reinterpret_cast<float2*>(A.action_real)[env] = a;
            if (A.mode == 0 && A.recovery) A.recovery[env] = 0;
            if (A.qrisk_out) A.qrisk_out[env] = fmaxf(A.q1[base + chosen], A.q2[base + chosen]);
        }
    }
}

// experiment.py:446-448: every env copy whose recovery policy acted pushes (state, real_action, reward, next_state, mask)
// into the task buffer as well.  One CTA: flags -> exclusive scan in env order -> rows appended after this step's n rows.
constexpr int kBothThreads = 1024;
__global__ void __launch_bounds__(kBothThreads) push_both_kernel(float* ring, int64_t cap, int64_t n, const uint8_t* recovery,
                                                                 const float* action_real, int64_t* counters) {
    __shared__ int warp_sums[kBothThreads / 32];
    __shared__ int total_s;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int64_t pos = counters[RRL_C_TASK_POS];   // already advanced past this step's n rows (rrl_counters_advance)
    const int64_t per = (n + kBothThreads - 1) / kBothThreads;
    const int64_t i0 = (int64_t)t * per, i1 = i0 + per < n ? i0 + per : n;
    int cnt = 0;
    for (int64_t i = i0; i < i1; ++i) cnt += recovery[i] ? 1 : 0;
    int inc = cnt;
    for (int o = 1; o < 32; o <<= 1) {
        const int c = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += c;
    }
    if (lane == 31) warp_sums[w] = inc;
    __syncthreads();
    if (w == 0) {
        int v = warp_sums[lane];
        int iv = v;
        for (int o = 1; o < 32; o <<= 1) {
            const int c = __shfl_up_sync(0xffffffffu, iv, o);
            if (lane >= o) iv += c;
        }
        warp_sums[lane] = iv - v;  // exclusive
        if (lane == 31) total_s = iv;
    }
    __syncthreads();
    int64_t rank = warp_sums[w] + inc - cnt;
    for (int64_t i = i0; i < i1; ++i) {
        if (!recovery[i]) continue;
        const int64_t src = ((pos - n + i) % cap + cap) % cap;
        const int64_t dst = (pos + rank) % cap;
        const float4 lo = *reinterpret_cast<const float4*>(ring + src * 8);
        const float4 hi = *reinterpret_cast<const float4*>(ring + src * 8 + 4);
        const float2 a = reinterpret_cast<const float2*>(action_real)[i];
        *reinterpret_cast<float4*>(ring + dst * 8) = make_float4(lo.x, lo.y, a.x, a.y);
        *reinterpret_cast<float4*>(ring + dst * 8 + 4) = hi;
        rank += 1;
    }
    if (t == 0) {
        const int64_t total = total_s;
        const int64_t len = counters[RRL_C_TASK_LEN] + total;
        counters[RRL_C_TASK_POS] = (pos + total) % cap;
        counters[RRL_C_TASK_LEN] = len < cap ? len : cap;
    }
}

inline int grid_for(int64_t items, int per_block) {
    int64_t b = (items + per_block - 1) / per_block;
    const int64_t cap = (int64_t)rrl_num_sms() * 8;
    return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}
}  // namespace

// floats per (env, candidate): state 2 + draw 2 + action 2 + log-prob 1 + q1 1 + q2 1
static const int64_t kFloatsPerCandidate = 9;

extern "C" int64_t rrl_select_workspace_floats(int64_t chunk_envs, int32_t samples) {
    if (chunk_envs <= 0 || samples <= 0) return 0;
    return chunk_envs * (int64_t)samples * kFloatsPerCandidate;
}

static int candidates_select(const rrl_agent_config_t* cfg, float* arena, int mode, int64_t n, int32_t samples, const double* state,
                             const float* draws, const float* cat_u, const float* rand_u, const uint8_t* only, int64_t start_steps,
                             uint64_t seed, int32_t stream_id, const int64_t* counters, float* workspace, int64_t workspace_floats,
                             float* action_task, float* action_real, uint8_t* recovery, float* qrisk_out, void* stream) {
    const int64_t chunk = workspace_floats / ((int64_t)samples * kFloatsPerCandidate);
    if (chunk <= 0) { rrl_set_error("candidates_select: workspace holds no env copy (need samples * 9 floats per env)"); return -2; }
    cudaStream_t st = (cudaStream_t)stream;
    for (int64_t env0 = 0; env0 < n; env0 += chunk) {
        const int64_t envs = n - env0 < chunk ? n - env0 : chunk;
        const int64_t rows = envs * samples;
        float* s_rep = workspace;
        float* eps = s_rep + rows * 2;
        float* act = eps + rows * 2;
        float* logp = act + rows * 2;
        float* q1 = logp + rows;
        float* q2 = q1 + rows;
        ExpandArgs E;
        memset(&E, 0, sizeof(E));
        E.n = n; E.env0 = env0; E.envs = envs; E.samples = samples; E.state = state; E.draws_in = draws;
        E.uniform_actions = mode;
        for (int i = 0; i < 2; ++i) { E.scale[i] = cfg->action_scale[i]; E.bias[i] = cfg->action_bias[i]; }
        E.seed = seed; E.stream_id = (uint32_t)stream_id; E.counters = counters; E.s_rep = s_rep; E.eps_out = eps;
        expand_kernel<<<grid_for(rows, kSelThreads), kSelThreads, 0, st>>>(E);
        RRL_CHECK_LAUNCH();
        const float* cand = eps;  // Q-sampling: the uniform actions themselves
        if (mode == 0) {          // pi, log_pi = policy.sample(state_batch)  (sac.py:143)
            int rc = rrl_policy_sample(cfg, arena, RRL_NET_POLICY, rows, s_rep, eps, act, logp, nullptr, stream);
            if (rc) return rc;
            cand = act;
        }
        int rc = rrl_twin_q_forward(cfg, arena, RRL_NET_QRISK, rows, s_rep, cand, q1, q2, stream);  // qrisk.py:184-196
        if (rc) return rc;
        PickArgs P;
        memset(&P, 0, sizeof(P));
        P.n = n; P.env0 = env0; P.envs = envs; P.samples = samples; P.mode = mode; P.q1 = q1; P.q2 = q2; P.logp = logp; P.cand = cand;
        P.cat_u = cat_u; P.rand_u = rand_u; P.only = only; P.start_steps = start_steps; P.seed = seed; P.stream_id = (uint32_t)stream_id;
        P.counters = counters; P.eps_safe = cfg->eps_safe;
        for (int i = 0; i < 2; ++i) { P.scale[i] = cfg->action_scale[i]; P.bias[i] = cfg->action_bias[i]; }
        P.action_task = action_task; P.action_real = action_real; P.qrisk_out = qrisk_out; P.recovery = recovery;
        pick_kernel<<<grid_for(envs * 32, kSelThreads), kSelThreads, 0, st>>>(P);
        RRL_CHECK_LAUNCH();
    }
    return 0;
}

extern "C" int rrl_sqrl_select_action(const rrl_agent_config_t* cfg, float* arena, int64_t n, int32_t samples, const double* state,
                                      const float* eps_cand, const float* cat_u, const float* rand_u, int64_t start_steps,
                                      uint64_t seed, int32_t stream_id, const int64_t* counters, float* workspace,
                                      int64_t workspace_floats, float* action_task, float* action_real, uint8_t* recovery,
                                      float* qrisk_out, void* stream) {
    RRL_CHECK_ARG(cfg && arena && state && workspace && action_task, "null argument");
    RRL_CHECK_ARG(n > 0 && samples > 0 && samples <= 1024, "n must be positive and samples in [1, 1024]");
    RRL_CHECK_ARG(!(cfg->algo_flags & RRL_ALGO_DETERMINISTIC), "the SQRL filter weighs candidates by the Gaussian policy's densities");
    return candidates_select(cfg, arena, 0, n, samples, state, eps_cand, cat_u, rand_u, nullptr, start_steps, seed, stream_id, counters,
                             workspace, workspace_floats, action_task, action_real, recovery, qrisk_out, stream);
}

extern "C" int rrl_qsample_recovery_action(const rrl_agent_config_t* cfg, float* arena, int64_t n, int32_t samples,
                                           const double* state, const float* cand_u, const uint8_t* recovery, uint64_t seed,
                                           int32_t stream_id, const int64_t* counters, float* workspace, int64_t workspace_floats,
                                           float* action_real, void* stream) {
    RRL_CHECK_ARG(cfg && arena && state && workspace && action_real, "null argument");
    RRL_CHECK_ARG(n > 0 && samples > 0 && samples <= 4096, "n must be positive and samples in [1, 4096]");
    return candidates_select(cfg, arena, 1, n, samples, state, cand_u, nullptr, nullptr, recovery, 0, seed, stream_id, counters,
                             workspace, workspace_floats, nullptr, action_real, nullptr, nullptr, stream);
}

extern "C" int rrl_replay_push_both(float* task_ring, int64_t task_capacity, int64_t n, const uint8_t* recovery,
                                    const float* action_real, int64_t* counters, void* stream) {
    RRL_CHECK_ARG(task_ring && recovery && action_real && counters, "null argument");
    RRL_CHECK_ARG(n > 0 && task_capacity >= 2 * n, "the task ring must hold two vector steps (capacity >= 2 n)");
    push_both_kernel<<<1, kBothThreads, 0, (cudaStream_t)stream>>>(task_ring, task_capacity, n, recovery, action_real, counters);
    RRL_CHECK_LAUNCH();
    return 0;
}
