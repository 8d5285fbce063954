// agent_common.cuh -- pieces shared by the SIMT (agent.cu) and tcgen05 (agent_tc.cu) agent kernels:
// weight views into the arena, the head post-processing of the four network families, acting arguments.
#pragma once
#include "agent_layout.cuh"
#include "update_tails.cuh"
#include <cuda_fp16.h>

namespace rrl {

enum Head { HEAD_Q = 0, HEAD_QRISK = 1, HEAD_GAUSS = 2, HEAD_STOCH = 3, HEAD_DET = 4 };


static __device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// ---------------------------------------------------------------------------------------------
// weights of one single-head MLP (pointers into the arena)
// ---------------------------------------------------------------------------------------------
struct HeadW {
    const float *W1, *b1, *W2, *W2T, *b2, *W3a, *b3a, *W3b, *b3b, *log_std;
    int n_in, na, nb;  // inputs (2|4); rows of W3a / W3b
    const __half *tc_img, *tc_imgT;  // fp16 hi/lo tcgen05 operand images of W2 and W2^T
};
struct HeadG {  // gradient pointers (same shapes); NULL = not needed
    float *W1, *b1, *W2, *b2, *W3a, *b3a, *W3b, *b3b, *log_std;
};

inline HeadW head_w(const Layout& L, const float* arena, int net, int head) {
    HeadW w;
    memset(&w, 0, sizeof(w));
    const int64_t* t = L.t_off[net];
    if (net == RRL_NET_POLICY) {
        w.W1 = arena + t[0]; w.b1 = arena + t[1]; w.W2 = arena + t[2]; w.b2 = arena + t[3];
        w.W3a = arena + t[4]; w.b3a = arena + t[5]; w.W3b = arena + t[6]; w.b3b = arena + t[7];
        w.n_in = 2; w.na = 2; w.nb = 2;
    } else if (net == RRL_NET_RECOVERY) {
        w.log_std = arena + t[0];
        w.W1 = arena + t[1]; w.b1 = arena + t[2]; w.W2 = arena + t[3]; w.b2 = arena + t[4];
        w.W3a = arena + t[5]; w.b3a = arena + t[6];
        w.n_in = 2; w.na = 2; w.nb = 0;
    } else {
        const int b = ((net == RRL_NET_QRISK || net == RRL_NET_QRISK_TARGET) ? 2 : 0) + 6 * head;
        w.W1 = arena + t[b]; w.b1 = arena + t[b + 1]; w.W2 = arena + t[b + 2]; w.b2 = arena + t[b + 3];
        w.W3a = arena + t[b + 4]; w.b3a = arena + t[b + 5];
        w.n_in = 4; w.na = 1; w.nb = 0;
    }
    w.W2T = arena + L.img_off[image_index(net, head)];
    w.tc_img = reinterpret_cast<const __half*>(arena + L.tc_img_off[image_index(net, head)]);
    w.tc_imgT = reinterpret_cast<const __half*>(arena + L.tc_imgT_off[image_index(net, head)]);
    return w;
}
// the task policy as configured: GaussianPolicy (two heads) or, with RRL_ALGO_DETERMINISTIC, DeterministicPolicy
// (model.py:447-485): the mean head only (the arena keeps the Gaussian layout; the log_std head stays unused and
// its gradients stay zero), action = tanh(mean)*scale + bias + noise, i.e. the StochasticPolicy arithmetic with a
// constant log_std of 0 and the caller's pre-scaled noise.
inline HeadW task_policy_w(const Layout& L, const float* arena, const rrl_agent_config_t* cfg, int* head_kind) {
    HeadW w = head_w(L, arena, RRL_NET_POLICY, 0);
    *head_kind = HEAD_GAUSS;
    if (cfg->algo_flags & RRL_ALGO_DETERMINISTIC) {
        w.nb = 0; w.W3b = nullptr; w.b3b = nullptr;
        w.log_std = arena + L.scalars + RRL_S_F64_BASE + 2 * RRL_D_ZERO;
        *head_kind = HEAD_DET;
    }
    return w;
}
inline HeadG head_g(const Layout& L, float* arena, int net, int head) {
    HeadW w = head_w(L, arena, net, head);
    const int64_t d = L.grad_off;  // trainable params start at offset 0 of the arena
    HeadG g;
    g.W1 = const_cast<float*>(w.W1) + d; g.b1 = const_cast<float*>(w.b1) + d;
    g.W2 = const_cast<float*>(w.W2) + d; g.b2 = const_cast<float*>(w.b2) + d;
    g.W3a = const_cast<float*>(w.W3a) + d; g.b3a = const_cast<float*>(w.b3a) + d;
    g.W3b = w.W3b ? const_cast<float*>(w.W3b) + d : nullptr;
    g.b3b = w.b3b ? const_cast<float*>(w.b3b) + d : nullptr;
    g.log_std = w.log_std ? const_cast<float*>(w.log_std) + d : nullptr;
    return g;
}

// ---------------------------------------------------------------------------------------------
// head post-processing
// ---------------------------------------------------------------------------------------------

// GaussianPolicy.sample (model.py:325-338)
static __device__ __forceinline__ void gauss_sample(const float raw[4], const float eps[2], const ActionSpace& sp, float a[2],
                                             float* logp, float mean_a[2]) {
    float lp = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float mean = raw[i];
        const float ls = fminf(fmaxf(raw[2 + i], LOG_SIG_MIN), LOG_SIG_MAX);
        const float sd = expf(ls);
        const float x = fmaf(sd, eps[i], mean);  // rsample: loc + eps * scale
        const float y = tanhf(x);
        a[i] = fmaf(y, sp.scale[i], sp.bias[i]);
        const float d = x - mean;
        float l = -(d * d) / (2.0f * (sd * sd)) - logf(sd) - HALF_LOG_2PI;  // Normal.log_prob
        l -= logf(sp.scale[i] * (1.0f - y * y) + 1e-6f);
        lp += l;
        mean_a[i] = fmaf(tanhf(mean), sp.scale[i], sp.bias[i]);
    }
    *logp = lp;
}

// StochasticPolicy.sample (model.py:512-525)
static __device__ __forceinline__ void stoch_sample(const float raw[2], const float* log_std, const float eps[2],
                                             const ActionSpace& sp, float a[2], float mean_a[2], float* logp) {
    float lp = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float mean = fmaf(tanhf(raw[i]), sp.scale[i], sp.bias[i]);
        const float ls = fmaxf(log_std[i], MIN_LOG_STD);
        const float sd = expf(ls);
        a[i] = fmaf(sd, eps[i], mean);
        mean_a[i] = mean;
        const float d = a[i] - mean;
        lp += -(d * d) / (2.0f * (sd * sd)) - logf(sd) - HALF_LOG_2PI;
    }
    *logp = lp;
}

// DeterministicPolicy.sample (model.py:478-479): noise ~ N(0, 0.1) clamped to +-0.25, from a standard normal draw
static __device__ __forceinline__ void det_noise(float e[2]) {
    e[0] = fminf(fmaxf(0.1f * e[0], -0.25f), 0.25f);
    e[1] = fminf(fmaxf(0.1f * e[1], -0.25f), 0.25f);
}

static __device__ __forceinline__ void philox_eps(uint64_t seed, uint32_t stream_id, uint64_t row, uint64_t step, uint32_t draw,
                                           float e[2]) {
    const Philox4 p = rrl_philox(seed, stream_id, row, step, draw);
    rrl_normal2_f32(p.x, p.y, &e[0], &e[1]);
}

// ---------------------------------------------------------------------------------------------
// tcgen05 operand images: element (n, k) of W2[n][k] * 64 as fp16 hi / lo in the UMMA canonical K-major layout
//   halves index (((c*2 + hl)*4 + kc)*32 + g)*64 + r*8 + e,  c = k/32, kc = (k%32)/8, e = k%8, g = n/8, r = n%8
// ---------------------------------------------------------------------------------------------
constexpr float kTcScaleB = 64.0f;
static __device__ __forceinline__ void tc_image_store(__half* __restrict__ img, int n, int k, float w) {
    const float s = fmaxf(fminf(w * kTcScaleB, 60000.0f), -60000.0f);
    const __half h = __float2half_rn(s);
    const __half l = __float2half_rn(s - __half2float(h));
    const int c = k >> 5, kc = (k & 31) >> 3, e = k & 7, g = n >> 3, r = n & 7;
    const size_t base = ((((size_t)c * 2) * 4 + kc) * 32 + g) * 64 + r * 8 + e;
    img[base] = h;
    img[base + (size_t)4 * 32 * 64] = l;   // hl = 1
}

// ---------------------------------------------------------------------------------------------
// grouped forward passes (training batches and the stand-alone forward entry points)
// ---------------------------------------------------------------------------------------------
struct FwdPass {
    HeadW w;
    int head;
    const float* xs;  // [rows][2]
    const float* xa;  // [rows][2] (n_in == 4)
    float *h1, *h2;    // activations kept for the backward pass (SIMT path: both; tcgen05 path: h2 only where a WEIGHT pass
                       // needs its values -- h1 is recomputed from the inputs, relu'(h2) comes from h2bits)
    uint32_t* h2bits;  // tcgen05 path: sign bits of h2, [rows][H / 32] words (bit j of word w: h2[row][32 w + j] > 0)
    uint32_t* h1bits;  // tcgen05 path: sign bits of h1, [rows][32] BYTES: byte (k / 8 % 4) * 8 + k / 32, bit k % 8 (the
                       // order the forward producers hold them in: thread (row, quarter) owns 8 contiguous bytes)
    int keep_h2;       // tcgen05 path: the backward has a WEIGHT pass for this forward pass (h2 values needed, not only signs)
    const float* eps;  // [rows][2] or NULL (Philox)
    uint32_t draw_id;
    float *out_q, *out_a, *out_logp, *out_mean, *out_raw, *out_eps;
    const __half* tc_img;  // fp16 hi/lo image of w.W2 (tcgen05 path)
};
struct FwdArgs {
    FwdPass p[10];
    int n_pass;
    const int64_t* rows_ptr;
    int64_t rows_const;
    ActionSpace sp;
    uint64_t seed;
    uint32_t stream_id;
    const int64_t* counters;
    int step_counter;  // which counter supplies the Philox step
    int use_tc;        // run the 256x256 contraction on tcgen05 (agent_tc.cu) instead of the fp32 SIMT tile
    TailArgs tail;     // stage run by the last CTA of the launch (update_tails.cuh); kind 0: none
};

// head-specific tail of one row: raw[0..3] = W3 h2 + b3 -> Q value / sigmoid / sampled action + log-prob
static __device__ __forceinline__ void forward_tail(const FwdPass& P, const FwdArgs& A, int64_t row, const float raw[4]) {
    if (P.out_raw) *reinterpret_cast<float4*>(P.out_raw + row * 4) = make_float4(raw[0], raw[1], raw[2], raw[3]);
    if (P.head == HEAD_Q) {
        P.out_q[row] = raw[0];
    } else if (P.head == HEAD_QRISK) {
        P.out_q[row] = sigmoidf_(raw[0]);
    } else {
        float e[2];
        if (P.eps) {
            const float2 ev = reinterpret_cast<const float2*>(P.eps)[row];
            e[0] = ev.x; e[1] = ev.y;
        } else {
            // DeterministicPolicy: ONE noise vector per sample() call, broadcast over the batch (model.py:478-480)
            philox_eps(A.seed, A.stream_id, P.head == HEAD_DET ? 0 : (uint64_t)row, (uint64_t)A.counters[A.step_counter], P.draw_id, e);
            if (P.head == HEAD_DET) det_noise(e);
        }
        float a[2], mean_a[2], lp;
        if (P.head == HEAD_GAUSS) gauss_sample(raw, e, A.sp, a, &lp, mean_a);
        else stoch_sample(raw, P.w.log_std, e, A.sp, a, mean_a, &lp);
        if (P.head == HEAD_DET) lp = 0.f;  // DeterministicPolicy.sample returns torch.tensor(0.) (model.py:481)
        if (P.out_a) reinterpret_cast<float2*>(P.out_a)[row] = make_float2(a[0], a[1]);
        if (P.out_logp) P.out_logp[row] = lp;
        if (P.out_mean) reinterpret_cast<float2*>(P.out_mean)[row] = make_float2(mean_a[0], mean_a[1]);
        if (P.out_eps) reinterpret_cast<float2*>(P.out_eps)[row] = make_float2(e[0], e[1]);
    }
}

struct ActArgs {
    HeadW pol, qr1, qr2, rec;
    int64_t n;
    const double* state;  // [2][n]
    const float *eps_task, *eps_rec, *rand_u;
    int use_recovery, eval;
    int det;  // task policy = DeterministicPolicy (model.py:447-485): a = tanh(mean)*scale + bias + noise, eps_task = the noise
    int64_t start_steps;
    uint64_t seed;
    uint32_t stream_id;
    const int64_t* counters;
    float eps_safe;
    ActionSpace sp;
    float *action_task, *action_real, *qrisk_out;
    uint8_t* recovery;
};


// mpc_tc.cu: the model-based planner's rollout on tcgen05
struct MpcTcArgs {
    const float* dyn;      // packed ensemble incl. its tcgen05 images (mpc_layout.cuh)
    HeadW qr1, qr2;
    const double* state;   // [2][E]
    const float* samples;  // [E][pop][hor*2]
    const float* eps;      // NULL (Philox) or [E][hor][nets][pop*npn][2]
    const int32_t* active; // [E] or NULL
    float* row_cost;       // [E][pop][npart]
    int64_t E;
    int pop, hor, npart, npn;
    uint64_t seed;
    uint32_t stream_id;
    const int64_t* counters;
    int iter;
};
int mpc_rollout_tc_launch(const MpcTcArgs& T, cudaStream_t st);
int dyn_tc_images_launch(float* dyn_image, cudaStream_t st);

// agent_tc.cu
int act_tc_launch(const ActArgs& A, const float* arena, const Layout& L, int stages, int max_ctas, cudaStream_t st);
int tc_images_launch(float* arena, const Layout& L, cudaStream_t st);
int fwd_tc_launch(const FwdArgs& A, int64_t max_rows, cudaStream_t st);
inline const __half* tc_img_of(const Layout& L, const float* arena, int net, int head) {
    return reinterpret_cast<const __half*>(arena + L.tc_img_off[image_index(net, head)]);
}
inline const __half* tc_imgT_of(const Layout& L, const float* arena, int net, int head) {
    return reinterpret_cast<const __half*>(arena + L.tc_imgT_off[image_index(net, head)]);
}

// ---------------------------------------------------------------------------------------------
// backward GEMM passes (head backward fused): see gemm_stream_kernel (agent.cu) / bwd_tc_kernel (agent_tc.cu)
//     DATA  : dh1[row][n] = (sum_k dh2[row][k] W2[k][n]) * relu'(h1[row][n])
//     WEIGHT: gW2[m][n] = sum_row dh2[row][m] h1[row][n];  gb2, gW3, gb3 from the same staged dh2 / h2 / dout
// with dh2[row][k] = (sum_o dout[row][o] W3[o][k]) * relu'(h2[row][k]) rebuilt on the fly.
// ---------------------------------------------------------------------------------------------
struct GemmPass {
    const float* dout;  // [rows][stride]
    int stride, n_out, na;
    const float *W3a, *W3b, *h2;
    const float* B;     // SIMT: W2 [H][H] (DATA) or h1 [rows][H] (WEIGHT)
    int k_is_rows;      // K = rows (WEIGHT) else K = H (DATA)
    const float* mask;  // DATA: h1 (SIMT path)
    float* C;
    float *gW3a, *gW3b, *gb3a, *gb3b, *gb2;  // WEIGHT only
    const __half* tc_imgT;  // DATA on tcgen05: fp16 hi/lo image of W2^T
    // tcgen05 path: layer 1 of the pass and its inputs (h1 and relu'(h1) are recomputed, never loaded) and the sign bits
    // of h2 written by the forward kernel (relu'(h2) of the DATA producers)
    const float *W1, *b1, *xs, *xa;
    int n_in;
    const uint32_t* h2bits;
    const uint32_t* h1bits;   // relu'(h1) of the DATA epilogue (layout: FwdPass::h1bits)
    // tcgen05 path, fused layer-1 backward (DATA passes; C may then be NULL: dh1 is consumed in the epilogue, never stored):
    //   gW1 / gb1 (weight gradients of layer 1; column sums over the rows, reduced across the row tiles through l1part by
    //   the last tile to arrive at l1ticket) or dxa ([rows][2], gradient w.r.t. the action inputs); never both
    float *gW1, *gb1, *dxa, *l1part;
    int* l1ticket;
};
struct GemmArgs {
    GemmPass p[8];
    const int64_t* rows_ptr;
    int n_pass;
    int use_tc;
    TailArgs tail;      // stage run by the last CTA of the launch (update_tails.cuh); kind 0: none
};
int bwd_tc_launch(const GemmArgs& G, int64_t max_rows, cudaStream_t st);

}  // namespace rrl
