// update_tails.cuh -- the small per-row stages of the three updates (losses + output gradients, policy-sample
// backward) as device functions, and the "tail" mechanism that runs one of them inside the LAST CTA of the kernel
// that produces its inputs instead of as a launch of its own:
//
//     fwd kernel (critic / Q_risk forward passes)  --tail-->  TD target + MSE / entropy / recovery loss + d(out)
//     layer-1 backward kernel (d(action) of the critics)  --tail-->  GaussianPolicy / StochasticPolicy sample backward
//
// (sac.py:192-231, qrisk.py:118-155, model.py:325-338,512-525).  Every CTA of the producing kernel takes a ticket
// after its last global store (release: __threadfence + atomicAdd); the CTA that draws the last ticket reads the other
// CTAs' outputs through L2 (__ldcg) and runs the stage with all of its threads.  The stand-alone kernels of the SIMT
// path (agent.cu) call the same bodies.  All reductions are fixed-order (deterministic).
#pragma once
#include "agent_layout.cuh"

namespace rrl {

#define LOG_SIG_MAX 2.0f
#define LOG_SIG_MIN (-20.0f)
#define MIN_LOG_STD (-13.815510557964274f) /* np.log(1e-6), model.py:499 */
#define HALF_LOG_2PI 0.9189385332046727f   /* math.log(math.sqrt(2*math.pi)) */

struct ActionSpace {
    float scale[2], bias[2];
};

// deterministic block sum of one value per thread (any blockDim <= 1024, all threads must call); valid on thread 0
template <typename T>
__device__ __forceinline__ T block_sum_any(T v, T* red /*[32]*/) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    T r = (T)0;
    if (threadIdx.x == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        for (int w = 0; w < nw; ++w) r += red[w];
    }
    return r;
}

// four sums at once (one shuffle tree, one shared-memory exchange): valid on thread 0.  red: [4][32]
__device__ __forceinline__ void block_sum4(float (&v)[4], float* red) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1)
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], s);
    __syncthreads();
    if ((threadIdx.x & 31) == 0)
#pragma unroll
        for (int i = 0; i < 4; ++i) red[i * 32 + (threadIdx.x >> 5)] = v[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nw = (blockDim.x + 31) >> 5;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float r = 0.f;
            for (int w = 0; w < nw; ++w) r += red[i * 32 + w];
            v[i] = r;
        }
    }
}

// ---- SAC losses (sac.py:192-231): TD target, critic MSE, policy loss, and the output gradients; plus the comparison
//      branches: RCPO target penalty (:202-205), DGD policy penalty (:224-228) and the gradients of the three scalar
//      multipliers log_alpha (:241-243), log_nu (:257-258), log_lambda (:266-267)
struct SacLossArgs {
    const float *r, *m, *next_logp, *qt1, *qt2, *qf1, *qf2, *logp, *qp1, *qp2;
    const float *sq1, *sq2;  // Q_risk(s, pi)  (DGD / update_nu) or NULL
    const float *qs1, *qs2;  // Q_risk(s, a)   (RCPO) or NULL
    float *target, *dqf1, *dqf2, *dqp1, *dqp2, *minq, *dsq1, *dsq2, *losses;
    float* scal;             // scalar block (RRL_S_* / RRL_D_*)
    float gamma, eps_safe, target_entropy;
    int flags;
    const int64_t* rows_ptr;
};
__device__ __forceinline__ void sac_loss_body(const SacLossArgs& A, float* red, double* redd) {
    const int64_t rows = *A.rows_ptr;
    if (rows <= 0) return;
    const float inv = 1.0f / (float)rows;
    const float alpha = A.scal[RRL_S_ALPHA];
    const float nu = A.scal[RRL_S_NU_ARG];
    double* sd = reinterpret_cast<double*>(A.scal + RRL_S_F64_BASE);
    const float lambda = (float)sd[RRL_D_LAMBDA];  // 0-dim float64 tensor times a float32 tensor: computed in float32
    const bool dgd = (A.flags & RRL_ALGO_DGD) != 0, rcpo = (A.flags & RRL_ALGO_RCPO) != 0;
    float l1 = 0.f, l2 = 0.f, lp = 0.f, la = 0.f;
    double gnu = 0.0, glam = 0.0;
    for (int64_t i = threadIdx.x; i < rows; i += blockDim.x) {
        const float minq_next = fminf(__ldcg(A.qt1 + i), __ldcg(A.qt2 + i)) - alpha * __ldcg(A.next_logp + i);
        float y = __ldcg(A.r + i) + __ldcg(A.m + i) * A.gamma * minq_next;
        if (rcpo) {
            const float qsafe = fmaxf(__ldcg(A.qs1 + i), __ldcg(A.qs2 + i));
            y -= lambda * qsafe;
            glam += (double)(A.eps_safe - qsafe);
        }
        A.target[i] = y;
        const float e1 = __ldcg(A.qf1 + i) - y, e2 = __ldcg(A.qf2 + i) - y;
        l1 = fmaf(e1, e1, l1);
        l2 = fmaf(e2, e2, l2);
        A.dqf1[i] = 2.0f * e1 * inv;
        A.dqf2[i] = 2.0f * e2 * inv;
        const float p1 = __ldcg(A.qp1 + i), p2 = __ldcg(A.qp2 + i);
        const float mq = fminf(p1, p2);
        A.minq[i] = mq;
        const float lg = __ldcg(A.logp + i);
        float row = alpha * lg;
        if (A.sq1) {
            const float s1 = __ldcg(A.sq1 + i), s2 = __ldcg(A.sq2 + i);
            const float ms = fmaxf(s1, s2);
            gnu += (double)(A.eps_safe - ms);
            if (dgd) {
                row += nu * (ms - A.eps_safe);
                // d(nu*max)/d(raw): torch.max routes to the larger input (ties split evenly), through the sigmoid
                const float g1 = s1 > s2 ? nu * inv : (s1 == s2 ? 0.5f * nu * inv : 0.f);
                const float g2 = s2 > s1 ? nu * inv : (s1 == s2 ? 0.5f * nu * inv : 0.f);
                A.dsq1[i] = g1 * s1 * (1.0f - s1);
                A.dsq2[i] = g2 * s2 * (1.0f - s2);
            }
        }
        lp += row - mq;
        la += lg + A.target_entropy;
        // d(-min)/dq: torch.min(a, b) routes the gradient to the smaller input (ties split evenly)
        A.dqp1[i] = p1 < p2 ? -inv : (p1 == p2 ? -0.5f * inv : 0.f);
        A.dqp2[i] = p2 < p1 ? -inv : (p1 == p2 ? -0.5f * inv : 0.f);
    }
    float s4v[4] = {l1, l2, lp, la};
    block_sum4(s4v, red);
    const float s1 = s4v[0], s2 = s4v[1], s3 = s4v[2], s4 = s4v[3];
    const bool need_d = (A.flags & (RRL_ALGO_UPDATE_NU | RRL_ALGO_RCPO)) != 0;   // uniform
    const double d1 = need_d ? block_sum_any(gnu, redd) : 0.0;
    const double d2 = need_d ? block_sum_any(glam, redd) : 0.0;
    if (threadIdx.x == 0) {
        A.losses[0] = s1 * inv;
        A.losses[1] = s2 * inv;
        A.losses[2] = s3 * inv;
        A.losses[4] = alpha;
        float alpha_loss = 0.f;
        if (A.flags & RRL_ALGO_AUTO_ALPHA) {  // alpha_loss = -(log_alpha * (log_pi + target_entropy)).mean()
            const float mean_t = s4 * inv;
            alpha_loss = -(A.scal[RRL_S_LOG_ALPHA] * mean_t);
            A.scal[RRL_S_G_LOG_ALPHA] = -mean_t;
        }
        A.losses[3] = alpha_loss;
        A.scal[RRL_S_ALPHA_LOSS] = alpha_loss;
        if (A.flags & RRL_ALGO_UPDATE_NU) sd[RRL_D_G_LOG_NU] = d1 / (double)rows;
        if (rcpo) sd[RRL_D_G_LOG_LAMBDA] = d2 / (double)rows;
    }
}

// ---- Q_risk losses (qrisk.py:118-148): target = c + m*gamma_safe*max(q1', q2'); MSE through the sigmoid
struct QrLossArgs {
    const float *c, *m, *qt1, *qt2, *q1, *q2;
    float *target, *dq1, *dq2, *losses;
    float gamma_safe;
    const int64_t* rows_ptr;
};
__device__ __forceinline__ void qrisk_loss_body(const QrLossArgs& A, float* red) {
    const int64_t rows = *A.rows_ptr;
    if (rows <= 0) return;
    const float inv = 1.0f / (float)rows;
    float l1 = 0.f, l2 = 0.f;
    for (int64_t i = threadIdx.x; i < rows; i += blockDim.x) {
        const float y = __ldcg(A.c + i) + __ldcg(A.m + i) * A.gamma_safe * fmaxf(__ldcg(A.qt1 + i), __ldcg(A.qt2 + i));
        A.target[i] = y;
        const float q1 = __ldcg(A.q1 + i), q2 = __ldcg(A.q2 + i);
        const float e1 = q1 - y, e2 = q2 - y;
        l1 = fmaf(e1, e1, l1);
        l2 = fmaf(e2, e2, l2);
        A.dq1[i] = 2.0f * e1 * inv * q1 * (1.0f - q1);  // d/d(raw) through sigmoid
        A.dq2[i] = 2.0f * e2 * inv * q2 * (1.0f - q2);
    }
    float s4v[4] = {l1, l2, 0.f, 0.f};
    block_sum4(s4v, red);
    if (threadIdx.x == 0) {
        A.losses[0] = s4v[0] * inv;
        A.losses[1] = s4v[1] * inv;
    }
}

// ---- recovery-policy loss (qrisk.py:150-155): mean max(Q1, Q2)(s, pi_rec(s))
struct RecLossArgs {
    const float *q1, *q2;
    float *dq1, *dq2, *losses;
    const int64_t* rows_ptr;
};
__device__ __forceinline__ void recovery_loss_body(const RecLossArgs& A, float* red) {
    const int64_t rows = *A.rows_ptr;
    if (rows <= 0) return;
    const float inv = 1.0f / (float)rows;
    float l = 0.f;
    for (int64_t i = threadIdx.x; i < rows; i += blockDim.x) {
        const float q1 = __ldcg(A.q1 + i), q2 = __ldcg(A.q2 + i);
        l += fmaxf(q1, q2);
        const float g1 = q1 > q2 ? inv : (q1 == q2 ? 0.5f * inv : 0.f);
        const float g2 = q2 > q1 ? inv : (q1 == q2 ? 0.5f * inv : 0.f);
        A.dq1[i] = g1 * q1 * (1.0f - q1);
        A.dq2[i] = g2 * q2 * (1.0f - q2);
    }
    const float s = block_sum_any(l, red);
    if (threadIdx.x == 0) A.losses[2] = s * inv;
}

// ---- GaussianPolicy.sample backward: d raw(mean, log_std) from dL/da (through the critic) and alpha*logp
struct GaussBwdArgs {
    const float *raw, *eps, *dxa1, *dxa2, *dxa3, *dxa4;  // dxa3/4: through Q_risk(s, pi) (DGD) or NULL
    float* draw;
    const float* scal;
    ActionSpace sp;
    const int64_t* rows_ptr;
};
__device__ __forceinline__ void gauss_backward_row(const GaussBwdArgs& A, int64_t i, int64_t rows) {
    const float inv = 1.0f / (float)rows;
    const float4 rv = __ldcg(reinterpret_cast<const float4*>(A.raw) + i);
    const float raw[4] = {rv.x, rv.y, rv.z, rv.w};
    const float2 e = __ldcg(reinterpret_cast<const float2*>(A.eps) + i);
    const float2 d1 = __ldcg(reinterpret_cast<const float2*>(A.dxa1) + i), d2 = __ldcg(reinterpret_cast<const float2*>(A.dxa2) + i);
    float da[2] = {d1.x + d2.x, d1.y + d2.y};
    if (A.dxa3) {
        const float2 d3 = __ldcg(reinterpret_cast<const float2*>(A.dxa3) + i), d4 = __ldcg(reinterpret_cast<const float2*>(A.dxa4) + i);
        da[0] += d3.x + d4.x;
        da[1] += d3.y + d4.y;
    }
    const float alpha = A.scal[RRL_S_ALPHA];
    const float ev[2] = {e.x, e.y};
    float out[4];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const float lsr = raw[2 + k];
        const float ls = fminf(fmaxf(lsr, LOG_SIG_MIN), LOG_SIG_MAX);
        const float sd = expf(ls);
        const float x = fmaf(sd, ev[k], raw[k]);
        const float y = tanhf(x);
        const float om = 1.0f - y * y;
        const float den = A.sp.scale[k] * om + 1e-6f;
        // dL/dy = dL/da * scale + (alpha/B) * d(-log(scale*(1-y^2)+1e-6))/dy
        const float dy = da[k] * A.sp.scale[k] + alpha * inv * (2.0f * A.sp.scale[k] * y) / den;
        const float dx = dy * om;
        out[k] = dx;                                               // d mean
        const float dls = dx * sd * ev[k] - alpha * inv;           // through x and through -log(std)
        out[2 + k] = (lsr >= LOG_SIG_MIN && lsr <= LOG_SIG_MAX) ? dls : 0.f;  // clamp backward
    }
    reinterpret_cast<float4*>(A.draw)[i] = make_float4(out[0], out[1], out[2], out[3]);
}
__device__ __forceinline__ void gauss_backward_body(const GaussBwdArgs& A) {   // one CTA over all rows
    const int64_t rows = *A.rows_ptr;
    for (int64_t i = threadIdx.x; i < rows; i += blockDim.x) gauss_backward_row(A, i, rows);
}

// ---- StochasticPolicy.sample backward (single CTA): d raw mean, d log_std
struct StochBwdArgs {
    const float *raw, *eps, *dxa1, *dxa2, *dxa3, *dxa4, *log_std;  // dxa3/4 optional; g_log_std NULL: Deterministic policy
    float *draw, *g_log_std;
    ActionSpace sp;
    const int64_t* rows_ptr;
};
__device__ __forceinline__ void stoch_backward_body(const StochBwdArgs& A, float* red) {
    const int64_t rows = *A.rows_ptr;
    if (rows <= 0) return;
    float gl[2] = {0.f, 0.f};
    float sd[2], pass[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const float ls = A.log_std[k];
        sd[k] = expf(fmaxf(ls, MIN_LOG_STD));
        pass[k] = ls >= MIN_LOG_STD ? 1.f : 0.f;
    }
    for (int64_t i = threadIdx.x; i < rows; i += blockDim.x) {
        const float4 rv = __ldcg(reinterpret_cast<const float4*>(A.raw) + i);
        const float2 e = __ldcg(reinterpret_cast<const float2*>(A.eps) + i);
        const float2 d1 = __ldcg(reinterpret_cast<const float2*>(A.dxa1) + i), d2 = __ldcg(reinterpret_cast<const float2*>(A.dxa2) + i);
        float da0 = d1.x + d2.x, da1 = d1.y + d2.y;
        if (A.dxa3) {
            const float2 d3 = __ldcg(reinterpret_cast<const float2*>(A.dxa3) + i), d4 = __ldcg(reinterpret_cast<const float2*>(A.dxa4) + i);
            da0 += d3.x + d4.x;
            da1 += d3.y + d4.y;
        }
        const float t0 = tanhf(rv.x), t1 = tanhf(rv.y);
        reinterpret_cast<float4*>(A.draw)[i] =
            make_float4(da0 * A.sp.scale[0] * (1.0f - t0 * t0), da1 * A.sp.scale[1] * (1.0f - t1 * t1), 0.f, 0.f);
        gl[0] = fmaf(da0 * sd[0], e.x, gl[0]);
        gl[1] = fmaf(da1 * sd[1], e.y, gl[1]);
    }
    const float s0 = block_sum_any(gl[0], red);
    const float s1 = block_sum_any(gl[1], red);
    if (threadIdx.x == 0 && A.g_log_std) {
        A.g_log_std[0] = s0 * pass[0];
        A.g_log_std[1] = s1 * pass[1];
    }
}

// ---- the tail ------------------------------------------------------------------------------------------------
enum TailKind { TAIL_NONE = 0, TAIL_SAC_LOSS = 1, TAIL_QR_LOSS = 2, TAIL_REC_LOSS = 3, TAIL_GAUSS_BWD = 4, TAIL_STOCH_BWD = 5 };
struct TailArgs {
    int kind;
    int64_t* ticket;   // device counter, 0 between launches (RRL_C_TICKET2)
    union {
        SacLossArgs sac;
        QrLossArgs qr;
        RecLossArgs rec;
        GaussBwdArgs gauss;
        StochBwdArgs stoch;
    };
};

// Called by ALL threads of EVERY CTA of the grid after the CTA's last global store (also by CTAs that have nothing to
// do).  The last CTA to arrive runs the stage; the ticket is left at 0 for the next launch.
// NOT inlined: the stage bodies are large and a kernel calls this from its early-exit path as well as from its end; one
// out-of-line copy keeps the kernels' instruction footprint (the tensor-core kernels run three warp roles through
// different code at once) within the instruction cache.
static __device__ __noinline__ void run_tail_stage(const TailArgs& T, float* red, double* redd) {
    switch (T.kind) {
        case TAIL_SAC_LOSS: sac_loss_body(T.sac, red, redd); break;
        case TAIL_QR_LOSS: qrisk_loss_body(T.qr, red); break;
        case TAIL_REC_LOSS: recovery_loss_body(T.rec, red); break;
        case TAIL_GAUSS_BWD: gauss_backward_body(T.gauss); break;
        case TAIL_STOCH_BWD: stoch_backward_body(T.stoch, red); break;
        default: break;
    }
}
__device__ __forceinline__ void run_tail(const TailArgs& T) {
    if (T.kind == TAIL_NONE) return;
    __shared__ int s_tail_last;
    __shared__ float s_tail_red[4 * 32];
    __shared__ double s_tail_redd[32];
    __syncthreads();                       // every thread's stores precede thread 0's fence + ticket
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned long long total = (unsigned long long)gridDim.x * gridDim.y * gridDim.z;
        const unsigned long long t = atomicAdd(reinterpret_cast<unsigned long long*>(T.ticket), 1ull);
        s_tail_last = (t == total - 1);
        if (s_tail_last) {
            *reinterpret_cast<volatile unsigned long long*>(T.ticket) = 0ull;
            __threadfence();
        }
    }
    __syncthreads();
    if (!s_tail_last) return;
    run_tail_stage(T, s_tail_red, s_tail_redd);
}

}  // namespace rrl
