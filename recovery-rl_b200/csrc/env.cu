// env.cu -- vectorised point environments (Navigation1 / Navigation2 / Maze) fused with the
// replay push, episode statistics and auto-reset.  One thread per env copy; fp64 SoA state.
//
// Compiled with -fmad=false AND written with explicit __dadd_rn/__dmul_rn/__fma_rn so that the
// arithmetic is bit-identical to the numpy expressions of the reference:
//   navigation  s' = (s + (double)clip(a)) + (0.05 * n)     env/navigation1.py:99-104
//               cost = -sqrt(fma(s.y, s.y, s.x*s.x))        env/navigation1.py:106-110 (OpenBLAS ddot)
//   maze        restated physics, see oracle/envs.py (MuJoCo is a closed third-party binary)
#include "common.cuh"
#include <math.h>

namespace {

struct EnvParams {
    rrl_env_config_t cfg;
    // maze constants (host-computed in double, identical expressions in oracle/envs.py)
    double c_a, c_b, h, gear;
    double plane_lo, plane_hi;  // x <= plane_lo  <=>  fl(x - R) <= -0.3 ;  x >= plane_hi  <=>  fl(x + R) >= 0.3
    double wx0[4], wx1[4], wy0[4], wy1[4];  // 1A, 1B, 2A, 2B rectangles
    float f_wx0[4], f_wx1[4], f_wy0[4], f_wy1[4];  // fp32 copies for the conservative pre-test (maze_may_touch)
    float f_plane_lo, f_plane_hi, f_rm, f_rm2;
};

// ---- obstacle.py:13-15, 44-45 : closed-interval rectangles ---------------------------------
__device__ __forceinline__ bool in_rect(double x, double y, double x0, double x1, double y0, double y1) {
    return (x0 <= x) && (x <= x1) && (y0 <= y) && (y <= y1);
}
__device__ __forceinline__ bool nav_obstacle(int kind, double x, double y) {
    if (kind == RRL_ENV_NAV1) {  // navigation1.py:41-42
        return in_rect(x, y, -100.0, 150.0, 5.0, 10.0) || in_rect(x, y, -100.0, -80.0, -10.0, 10.0) ||
               in_rect(x, y, -100.0, 150.0, -10.0, -5.0);
    }
    return in_rect(x, y, -30.0, -20.0, -7.5, 7.5);  // navigation2.py:41
}

// ---- maze geometry: simple_maze.xml:16-25 + maze.py:201-206 ----------------------------------
// disc radius r against the 4 outer planes (closed) and 4 axis-aligned rectangles (strict).
// The oracle's tests are   x - R <= -0.3 | x + R >= 0.3 | ...   and   dx*dx + dy*dy < R*R  with
// dx = max(max(x0 - x, 0), x - x1).  Both are evaluated here in exactly equivalent, cheaper forms:
//   * fl(x - R) <= -0.3 is a down-set in x (rounding is monotonic), i.e. x <= plane_lo with plane_lo the
//     largest double that satisfies it (found on the host with the same IEEE arithmetic); same for the
//     other three planes -> four compares, no arithmetic;
//   * dx >= R  =>  fl(dx*dx) >= fl(R*R)  =>  fl(dx*dx + dy*dy) >= fl(R*R): no touch, so dy is only needed
//     inside the wall's x band; walls 1A/1B and 2A/2B share their x range;
//   * tests that provably fail are skipped altogether (speculative chunks, see maze_substeps_warp).
#define MAZE_R 0.025
__device__ __forceinline__ bool maze_rect_touch_y(double dx, double y, double y0, double y1) {
    double dy = fmax(fmax(__dsub_rn(y0, y), 0.0), __dsub_rn(y, y1));
    double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
    return d2 < (MAZE_R * MAZE_R);
}
__device__ __forceinline__ bool maze_touch(const EnvParams& P, double x, double y) {
    if ((x <= P.plane_lo) || (x >= P.plane_hi) || (y <= P.plane_lo) || (y >= P.plane_hi)) return true;
    const double dx1 = fmax(fmax(__dsub_rn(P.wx0[0], x), 0.0), __dsub_rn(x, P.wx1[0]));
    const double dx2 = fmax(fmax(__dsub_rn(P.wx0[2], x), 0.0), __dsub_rn(x, P.wx1[2]));
    bool hit = false;
    if (dx1 < MAZE_R) hit = maze_rect_touch_y(dx1, y, P.wy0[0], P.wy1[0]) || maze_rect_touch_y(dx1, y, P.wy0[1], P.wy1[1]);
    if (dx2 < MAZE_R) hit = hit || maze_rect_touch_y(dx2, y, P.wy0[2], P.wy1[2]) || maze_rect_touch_y(dx2, y, P.wy0[3], P.wy1[3]);
    return hit;
}
// 500 substeps for one warp of envs.  Per-substep semantics (oracle/envs.py): contact test on the
// pre-integration position, then one semi-implicit Euler substep; contact freezes the disc.
//
// Speculative chunks: v starts at 0 and v <- c_a*v + fb keeps the sign of fb, so every coordinate moves
// MONOTONICALLY during one env step (rounding is monotone).  The positions tested inside a chunk of MAZE_CHUNK
// substeps therefore lie in the box spanned by the chunk's first and last position.  A chunk is integrated
// test-free; only if that box comes within R + 2e-6 of a solid (maze_may_touch) the lane rewinds and replays the
// chunk with the exact per-substep test.  A lane replays the chunk in which it touches (plus, rarely, chunks
// spent within 2e-6 of a solid without touching), so all 32 envs of a warp run the same 500-substep schedule
// instead of serialising divergent test paths.
constexpr int MAZE_CHUNK = 10;
// Conservative pre-test in fp32 (FMA/ALU pipes, the fp64 pipe is the busy one): can ANY point of the box
// [xl,xh]x[yl,yh] touch a solid?  Planes: threshold + 1e-6.  Walls: the distance from the box to the rectangle
// (exact rounded corners, not the inflated bounding box, so that a disc resting in a corner pocket does not keep
// replaying chunks) against R + 2e-6.  fp32 conversion and arithmetic errors are < 1e-7 for |x| <= 0.3, far inside
// the margins; the exact fp64 tests need distance < R(1 + 1e-15).
__device__ __forceinline__ bool maze_may_touch(const EnvParams& P, float xl, float xh, float yl, float yh) {
    if ((xl <= P.f_plane_lo) || (xh >= P.f_plane_hi) || (yl <= P.f_plane_lo) || (yh >= P.f_plane_hi)) return true;
    const float dx1 = fmaxf(fmaxf(P.f_wx0[0] - xh, xl - P.f_wx1[0]), 0.f);  // walls 1A / 1B share their x range
    const float dx2 = fmaxf(fmaxf(P.f_wx0[2] - xh, xl - P.f_wx1[2]), 0.f);  // walls 2A / 2B
    bool hit = false;
    if (dx1 < P.f_rm) {
        const float da = fmaxf(fmaxf(P.f_wy0[0] - yh, yl - P.f_wy1[0]), 0.f);
        const float db = fmaxf(fmaxf(P.f_wy0[1] - yh, yl - P.f_wy1[1]), 0.f);
        hit = (fmaf(dx1, dx1, da * da) < P.f_rm2) || (fmaf(dx1, dx1, db * db) < P.f_rm2);
    }
    if (dx2 < P.f_rm) {
        const float da = fmaxf(fmaxf(P.f_wy0[2] - yh, yl - P.f_wy1[2]), 0.f);
        const float db = fmaxf(fmaxf(P.f_wy0[3] - yh, yl - P.f_wy1[3]), 0.f);
        hit = hit || (fmaf(dx2, dx2, da * da) < P.f_rm2) || (fmaf(dx2, dx2, db * db) < P.f_rm2);
    }
    return hit;
}
__device__ __forceinline__ void maze_integrate(const EnvParams& P, double fbx, double fby, double& x, double& y, double& vx,
                                               double& vy) {
    vx = __dadd_rn(__dmul_rn(P.c_a, vx), fbx);
    vy = __dadd_rn(__dmul_rn(P.c_a, vy), fby);
    x = __dadd_rn(x, __dmul_rn(P.h, vx));
    y = __dadd_rn(y, __dmul_rn(P.h, vy));
}
__device__ __forceinline__ void maze_substeps_warp(const EnvParams& P, bool idle, double fbx, double fby, double& x,
                                                   double& y, bool& contact) {
    const int nsub = P.cfg.maze_substeps;
    double vx = 0.0, vy = 0.0;
    bool frozen = idle;  // idle lanes (beyond n) and lanes in contact do not move
    for (int k = 0; k < nsub; k += MAZE_CHUNK) {
        const int c = min(MAZE_CHUNK, nsub - k);
        const double sx = x, sy = y, svx = vx, svy = vy;
        bool may = false;
        if (!frozen) {
            if (c == MAZE_CHUNK) {
#pragma unroll
                for (int j = 0; j < MAZE_CHUNK; ++j) maze_integrate(P, fbx, fby, x, y, vx, vy);
            } else {
                for (int j = 0; j < c; ++j) maze_integrate(P, fbx, fby, x, y, vx, vy);
            }
            const float fx0 = (float)sx, fx1 = (float)x, fy0 = (float)sy, fy1 = (float)y;
            may = maze_may_touch(P, fminf(fx0, fx1), fmaxf(fx0, fx1), fminf(fy0, fy1), fmaxf(fy0, fy1));
        }
        if (__any_sync(0xffffffffu, may)) {
            if (may) {  // rewind, replay the chunk with the exact test before every substep
                x = sx; y = sy; vx = svx; vy = svy;
                for (int j = 0; j < c; ++j) {
                    if (maze_touch(P, x, y)) {
                        frozen = true;
                        contact = true;
                        break;
                    }
                    maze_integrate(P, fbx, fby, x, y, vx, vy);
                }
            }
            if (__all_sync(0xffffffffu, frozen)) break;
        }
    }
}

__device__ __forceinline__ void reset_state(const EnvParams& P, int64_t i, const double* draws, int64_t n,
                                            uint64_t vec_step, uint32_t draw_id, double* x, double* y) {
    double d0, d1;
    if (draws) {
        d0 = draws[i];
        d1 = draws[n + i];
    } else {
        Philox4 p = rrl_philox(P.cfg.seed, (uint32_t)P.cfg.stream_id, (uint64_t)i, vec_step, draw_id);
        if (P.cfg.kind == RRL_ENV_MAZE) {
            d0 = rrl_u53(p.x, p.y);
            d1 = rrl_u53(p.z, p.w);
        } else {
            rrl_normal2_f64(p, &d0, &d1);
        }
    }
    if (P.cfg.kind == RRL_ENV_MAZE) {
        // maze.py:195-197 mode 'h': np.random.uniform(lo, hi) = lo + (hi - lo) * u
        *x = __dadd_rn(-0.22, __dmul_rn((-0.13) - (-0.22), d0));
        *y = __dadd_rn(-0.22, __dmul_rn(0.22 - (-0.22), d1));
    } else {
        // navigation1.py:92  START_STATE + np.random.randn(2)
        *x = __dadd_rn(-50.0, d0);
        *y = __dadd_rn(0.0, d1);
    }
}

__global__ void __launch_bounds__(256) env_reset_kernel(EnvParams P, const uint8_t* __restrict__ mask,
                                                        const double* __restrict__ draws, double* __restrict__ state,
                                                        int32_t* __restrict__ ep_steps, double* __restrict__ ep_return,
                                                        const int64_t* __restrict__ counters) {
    const int64_t n = P.cfg.n_envs;
    const uint64_t vstep = counters ? (uint64_t)counters[RRL_C_VEC_STEP] : 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (mask && !mask[i]) continue;
        double x, y;
        reset_state(P, i, draws, n, vstep, RRL_DRAW_INIT_RESET, &x, &y);
        state[i] = x;
        state[n + i] = y;
        if (ep_steps) ep_steps[i] = 0;
        if (ep_return) ep_return[i] = 0.0;
    }
}

__device__ __forceinline__ void warp_count_add(bool pred, int64_t* addr) {
    unsigned m = __ballot_sync(__activemask(), pred);
    if (pred && (threadIdx.x & 31) == (__ffs(m) - 1)) atomicAdd((unsigned long long*)addr, (unsigned long long)__popc(m));
}

__global__ void __launch_bounds__(256)
env_step_kernel(EnvParams P, const float* __restrict__ a_task, const float* __restrict__ a_real,
                const uint8_t* __restrict__ recovery, const double* __restrict__ noise,
                const double* __restrict__ reset_draws, double* __restrict__ state, int32_t* __restrict__ ep_steps,
                double* __restrict__ ep_return, float* __restrict__ task_ring, int64_t task_cap,
                float* __restrict__ cons_ring, uint8_t* __restrict__ cons_flags, int64_t cons_cap,
                int64_t* __restrict__ counters, double* __restrict__ o_next, double* __restrict__ o_reward,
                uint8_t* __restrict__ o_done, uint8_t* __restrict__ o_cons, uint8_t* __restrict__ o_succ,
                const double* __restrict__ a_f64) {
    pdl_wait();   // programmatic dependent launch (common.cuh): the acting kernel has completed past here
    const int64_t n = P.cfg.n_envs;
    const int kind = P.cfg.kind;
    const uint64_t vstep = (uint64_t)counters[RRL_C_VEC_STEP];
    const int64_t task_pos = counters[RRL_C_TASK_POS];
    const int64_t cons_pos = counters[RRL_C_CONS_POS];
    for (int64_t i0 = blockIdx.x * (int64_t)blockDim.x; i0 < n; i0 += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = i0 + threadIdx.x;
        const bool live = i < n;
        bool ep_end = false, viol = false, succ_end = false, rec_used = false;
        double ep_ret_final = 0.0;
        double mz_x = 0.0, mz_y = 0.0;
        bool mz_contact = false;
        if (kind == RRL_ENV_MAZE) {  // all 32 lanes take part (warp-uniform substep loop)
            double fbx = 0.0, fby = 0.0;
            if (live) {
                mz_x = state[i];
                mz_y = state[n + i];
                // E8: maze.  clip in fp32 to float32(0.1), then ctrl is fp64
                const float2 ar = reinterpret_cast<const float2*>(a_real)[i];
                const float cx = fminf(fmaxf(ar.x, -0.1f), 0.1f), cy = fminf(fmaxf(ar.y, -0.1f), 0.1f);
                const double dax = a_f64 ? fmin(fmax(a_f64[2 * i], -0.1), 0.1) : (double)cx;
                const double day = a_f64 ? fmin(fmax(a_f64[2 * i + 1], -0.1), 0.1) : (double)cy;
                fbx = __dmul_rn(P.c_b, __dmul_rn(P.gear, dax));
                fby = __dmul_rn(P.c_b, __dmul_rn(P.gear, day));
            }
            maze_substeps_warp(P, !live, fbx, fby, mz_x, mz_y, mz_contact);
        }
        if (live) {
            const double sx = state[i], sy = state[n + i];
            const float2 at = reinterpret_cast<const float2*>(a_task)[i];
            const float2 ar = reinterpret_cast<const float2*>(a_real)[i];
            double nx, ny, reward;
            bool constraint, done, success;
            if (kind != RRL_ENV_MAZE) {
                // E1: process_action  np.clip(a, -1, 1) stays fp32
                const float cx = fminf(fmaxf(ar.x, -1.0f), 1.0f), cy = fminf(fmaxf(ar.y, -1.0f), 1.0f);
                // fp64 actions (offline-data generators: np.clip(np.random.randn(2), -1, 1)) stay fp64
                const double dax = a_f64 ? fmin(fmax(a_f64[2 * i], -1.0), 1.0) : (double)cx;
                const double day = a_f64 ? fmin(fmax(a_f64[2 * i + 1], -1.0), 1.0) : (double)cy;
                if (nav_obstacle(kind, sx, sy)) {  // E2: stuck inside an obstacle
                    nx = sx;
                    ny = sy;
                } else {
                    double e0, e1;
                    if (noise) {
                        e0 = noise[i];
                        e1 = noise[n + i];
                    } else {
                        Philox4 p = rrl_philox(P.cfg.seed, (uint32_t)P.cfg.stream_id, (uint64_t)i, vstep, RRL_DRAW_ENV_NOISE);
                        rrl_normal2_f64(p, &e0, &e1);
                    }
                    nx = __dadd_rn(__dadd_rn(sx, dax), __dmul_rn(0.05, e0));
                    ny = __dadd_rn(__dadd_rn(sy, day), __dmul_rn(0.05, e1));
                }
                // E3: -||GOAL - s|| on the PRE-step state; ddot accumulates with one FMA
                const double cost = -sqrt(__fma_rn(sy, sy, __dmul_rn(sx, sx)));
                constraint = nav_obstacle(kind, nx, ny);
                success = cost > -4.0;
                done = success || constraint;
                reward = cost;
            } else {
                // E8: maze.  clip in fp32 to float32(0.1), then ctrl is fp64
                const double x = mz_x, y = mz_y;
                const bool contact = mz_contact;
                nx = x;
                ny = y;
                constraint = contact;
                // E10: sqrt(mean((goal - qpos)^2)), goal = (0.25, 0), unfused
                const double d0 = __dsub_rn(0.25, nx), d1 = __dsub_rn(0.0, ny);
                const double dist = sqrt(__dmul_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), 0.5));
                reward = -dist;
                done = (ep_steps[i] + 1 >= P.cfg.horizon) || constraint || (dist < 0.03);
                success = reward > -0.03;
            }
            const int steps = ep_steps[i] + 1;
            // experiment.py:431-435: penalty, mask BEFORE the horizon truncation
            const double r_pen = constraint ? __dsub_rn(reward, P.cfg.reward_penalty) : reward;
            const float mask = done ? 0.0f : 1.0f;
            const bool done_h = done || (steps == P.cfg.horizon);
            const double ret = __dadd_rn(ep_return[i], reward);

            if (o_next) {
                o_next[i] = nx;
                o_next[n + i] = ny;
            }
            if (o_reward) o_reward[i] = reward;
            if (o_done) o_done[i] = done_h;
            if (o_cons) o_cons[i] = constraint;
            if (o_succ) o_succ[i] = success;

            const float fsx = (float)sx, fsy = (float)sy, fnx = (float)nx, fny = (float)ny;
            if (task_ring) {
                float4* rec = reinterpret_cast<float4*>(task_ring + ((task_pos + i) % task_cap) * 8);
                rec[0] = make_float4(fsx, fsy, at.x, at.y);
                rec[1] = make_float4((float)r_pen, fnx, fny, mask);
            }
            if (cons_ring) {
                const int64_t slot = (cons_pos + i) % cons_cap;
                float4* rec = reinterpret_cast<float4*>(cons_ring + slot * 8);
                rec[0] = make_float4(fsx, fsy, ar.x, ar.y);
                rec[1] = make_float4(constraint ? 1.0f : 0.0f, fnx, fny, mask);
                cons_flags[slot] = constraint ? 1 : 2;  // bit0: pos_idx != 0, bit1: (1 - pos_idx) != 0
            }
            if (done_h && !(P.cfg.flags & RRL_ENV_NO_AUTO_RESET)) {
                ep_end = true;
                viol = constraint;
                succ_end = success;
                rec_used = recovery ? (recovery[i] != 0) : false;
                ep_ret_final = ret;
                double rx, ry;
                reset_state(P, i, reset_draws, n, vstep, RRL_DRAW_ENV_RESET, &rx, &ry);
                state[i] = rx;
                state[n + i] = ry;
                ep_steps[i] = 0;
                ep_return[i] = 0.0;
            } else {
                state[i] = nx;
                state[n + i] = ny;
                ep_steps[i] = steps;
                ep_return[i] = ret;
            }
        }
        // experiment.py:455-461 episode statistics from the LAST step's info
        warp_count_add(ep_end, counters + RRL_C_EPISODES);
        warp_count_add(ep_end && viol, counters + RRL_C_NUM_VIOLS);
        warp_count_add(ep_end && succ_end, counters + RRL_C_NUM_SUCCESSES);
        warp_count_add(ep_end && viol && rec_used, counters + RRL_C_VIOL_RECOVERY);
        warp_count_add(ep_end && viol && !rec_used, counters + RRL_C_VIOL_NO_RECOV);
        if (ep_end) atomicAdd(reinterpret_cast<double*>(counters + RRL_C_RETURN_SUM_BITS), ep_ret_final);
    }
}

__global__ void counters_advance_kernel(int64_t* counters, int64_t n, int64_t task_cap, int64_t cons_cap,
                                        int push_task, int push_cons) {
    pdl_wait();
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        counters[RRL_C_TOTAL_NUMSTEPS] += n;
        counters[RRL_C_VEC_STEP] += 1;
        if (push_task) {
            counters[RRL_C_TASK_POS] = (counters[RRL_C_TASK_POS] + n) % task_cap;
            int64_t l = counters[RRL_C_TASK_LEN] + n;
            counters[RRL_C_TASK_LEN] = l < task_cap ? l : task_cap;
        }
        if (push_cons) {
            counters[RRL_C_CONS_POS] = (counters[RRL_C_CONS_POS] + n) % cons_cap;
            int64_t l = counters[RRL_C_CONS_LEN] + n;
            counters[RRL_C_CONS_LEN] = l < cons_cap ? l : cons_cap;
        }
    }
}

EnvParams make_params(const rrl_env_config_t* cfg) {
    EnvParams P;
    P.cfg = *cfg;
    // simple_maze.xml: cylinder r = 0.025, half-height 0.025 (:28), default density 1000,
    // slide joints damping 0.01 (:29-30), motor gear 0.05 (:8), timestep 0.002 (MuJoCo default)
    const double mass = 1000.0 * 3.141592653589793 * 0.025 * 0.025 * 0.05;
    const double h = 0.002, d = 0.01;
    P.h = h;
    P.gear = 0.05;
    P.c_a = mass / (mass + h * d);
    P.c_b = h / (mass + h * d);
    // wall boxes: half-sizes (.02,.2,.005) with the thin axis along world x at x = -/+0.1
    // (simple_maze.xml:22-25); y centres after reset(): maze.py:199-206
    const double w1 = -0.08, w2 = 0.08;
    const double cy[4] = {0.5 + w1, -0.25 + w1, 0.4 + w2, -0.25 + w2};
    const double cx[4] = {-0.1, -0.1, 0.1, 0.1};
    for (int w = 0; w < 4; ++w) {
        P.wx0[w] = cx[w] - 0.005;
        P.wx1[w] = cx[w] + 0.005;
        P.wy0[w] = cy[w] - 0.2;
        P.wy1[w] = cy[w] + 0.2;
        P.f_wx0[w] = (float)P.wx0[w]; P.f_wx1[w] = (float)P.wx1[w];
        P.f_wy0[w] = (float)P.wy0[w]; P.f_wy1[w] = (float)P.wy1[w];
    }
    P.f_rm = (float)(MAZE_R + 2e-6);
    P.f_rm2 = P.f_rm * P.f_rm;
    // exact thresholds of the plane tests (see maze_touch): walk to the boundary of the monotone predicate
    {
        double c = -0.3 + MAZE_R;
        while (c - MAZE_R <= -0.3) c = nextafter(c, 1.0);
        while (!(c - MAZE_R <= -0.3)) c = nextafter(c, -1.0);
        P.plane_lo = c;  // largest x with fl(x - R) <= -0.3
        c = 0.3 - MAZE_R;
        while (c + MAZE_R >= 0.3) c = nextafter(c, -1.0);
        while (!(c + MAZE_R >= 0.3)) c = nextafter(c, 1.0);
        P.plane_hi = c;  // smallest x with fl(x + R) >= 0.3
        P.f_plane_lo = (float)(P.plane_lo + 1e-6);
        P.f_plane_hi = (float)(P.plane_hi - 1e-6);
    }
    return P;
}

int grid_for(int64_t n, int threads = 256) {
    int64_t blocks = (n + threads - 1) / threads;
    int64_t cap = (int64_t)rrl_num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace

extern "C" int rrl_env_reset(const rrl_env_config_t* cfg, const uint8_t* mask, const double* draws, double* state,
                             int32_t* ep_steps, double* ep_return, const int64_t* counters, void* stream) {
    RRL_CHECK_ARG(cfg && state, "null argument");
    RRL_CHECK_ARG(cfg->kind >= RRL_ENV_NAV1 && cfg->kind <= RRL_ENV_MAZE, "unknown env kind");
    RRL_CHECK_ARG(cfg->n_envs > 0, "n_envs must be positive");
    EnvParams P = make_params(cfg);
    env_reset_kernel<<<grid_for(cfg->n_envs), 256, 0, (cudaStream_t)stream>>>(P, mask, draws, state, ep_steps,
                                                                              ep_return, counters);
    RRL_CHECK_LAUNCH();
    return 0;
}

extern "C" int rrl_env_step(const rrl_env_config_t* cfg, const float* action_task, const float* action_real,
                            const uint8_t* recovery, const double* noise, const double* reset_draws, double* state,
                            int32_t* ep_steps, double* ep_return, float* task_ring, int64_t task_capacity,
                            float* cons_ring, uint8_t* cons_flags, int64_t cons_capacity, int64_t* counters,
                            double* out_next_state, double* out_reward, uint8_t* out_done, uint8_t* out_constraint,
                            uint8_t* out_success, const double* action_real_f64, void* stream) {
    RRL_CHECK_ARG(cfg && action_task && action_real && state && ep_steps && ep_return && counters, "null argument");
    RRL_CHECK_ARG(cfg->kind >= RRL_ENV_NAV1 && cfg->kind <= RRL_ENV_MAZE, "unknown env kind");
    RRL_CHECK_ARG(cfg->n_envs > 0, "n_envs must be positive");
    RRL_CHECK_ARG(!task_ring || task_capacity >= cfg->n_envs, "task ring smaller than one vector step");
    RRL_CHECK_ARG(!cons_ring || (cons_flags && cons_capacity >= cfg->n_envs), "constraint ring too small / flags missing");
    EnvParams P = make_params(cfg);
    // maze: 500 dependent substeps per env and divergent contact paths -> small CTAs balance the 148 SMs better
    const int threads = cfg->kind == RRL_ENV_MAZE ? 64 : 256;
    RRL_CUDA(rrl_launch_pdl(env_step_kernel, dim3(grid_for(cfg->n_envs, threads)), dim3(threads), 0, (cudaStream_t)stream,
                            P, action_task, action_real, recovery, noise, reset_draws, state, ep_steps, ep_return, task_ring,
                            task_capacity > 0 ? task_capacity : (int64_t)1, cons_ring, cons_flags,
                            cons_capacity > 0 ? cons_capacity : (int64_t)1, counters, out_next_state, out_reward, out_done,
                            out_constraint, out_success, action_real_f64));
    return 0;
}

extern "C" int rrl_counters_advance(int64_t* counters, int64_t n, int64_t task_capacity, int64_t cons_capacity,
                                    int push_task, int push_cons, void* stream) {
    RRL_CHECK_ARG(counters, "null counters");
    RRL_CUDA(rrl_launch_pdl(counters_advance_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, counters, n,
                            task_capacity > 0 ? task_capacity : (int64_t)1, cons_capacity > 0 ? cons_capacity : (int64_t)1,
                            push_task, push_cons));
    return 0;
}
