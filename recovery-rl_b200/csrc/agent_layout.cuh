// agent_layout.cuh -- layout of the agent arena (one flat fp32 device buffer, caller-owned).
//
//   [ params: critic | policy | qrisk | recovery | critic_target | qrisk_target ]
//   [ grads : critic | policy | qrisk | recovery ]      <- ONE contiguous block = the NCCL all-reduce payload
//   [ adam m: same 4 nets ] [ adam v: same 4 nets ]
//   [ W2T images: the ten 256x256 hidden matrices transposed to k-major (operand B of the forward GEMM) ]
//   [ tcgen05 images: fp16 hi/lo split of the ten hidden matrices in UMMA canonical K-major layout ]
//   [ scratch: sampled batches, activations, per-row outputs, losses ]
//
// Within a net the tensors follow torch's parameters() order of the reference module
// (recovery_rl/model.py:49-76 QNetwork, :172-199 QNetworkConstraint incl. its dead BatchNorm1d,
// :295-343 GaussianPolicy, :489-530 StochasticPolicy), each padded to a multiple of 4 floats.
#pragma once
#include "common.cuh"

namespace rrl {

constexpr int H = 256;  // hidden width (arg_utils.py:89-92); all kernels are specialised for it
constexpr int kMaxTensors = 14;
constexpr int kNumImages = 10;
constexpr int kTcHeads = kNumImages;  // one fp16 hi/lo tcgen05 operand image per 256x256 hidden matrix (index = image_index)
constexpr int kPassSlots = 8;  // activation slots shared by the SAC and Q_risk updates (5, 6: Q_risk(s, pi) of the DGD branch;
                               // 7: the recovery policy, whose forward pass may run next to the SAC update)
constexpr int kRecSlot = 7;

struct TDesc {
    int rows, cols;  // cols == 0: 1-D tensor of `rows` elements
};

// storage order of the nets inside the param block (trainable nets first, in grad-block order)
__host__ __device__ inline int storage_rank(int net) {
    switch (net) {
        case RRL_NET_CRITIC: return 0;
        case RRL_NET_POLICY: return 1;
        case RRL_NET_QRISK: return 2;
        case RRL_NET_RECOVERY: return 3;
        case RRL_NET_CRITIC_TARGET: return 4;
        default: return 5;
    }
}

inline int net_tensors(int net, TDesc* out) {
    static const TDesc q[12] = {{H, 4}, {H, 0}, {H, H}, {H, 0}, {1, H}, {1, 0}, {H, 4}, {H, 0}, {H, H}, {H, 0}, {1, H}, {1, 0}};
    static const TDesc pol[8] = {{H, 2}, {H, 0}, {H, H}, {H, 0}, {2, H}, {2, 0}, {2, H}, {2, 0}};
    static const TDesc rec[7] = {{2, 0}, {H, 2}, {H, 0}, {H, H}, {H, 0}, {2, H}, {2, 0}};
    int n = 0;
    switch (net) {
        case RRL_NET_CRITIC:
        case RRL_NET_CRITIC_TARGET:
            for (int i = 0; i < 12; ++i) out[n++] = q[i];
            break;
        case RRL_NET_QRISK:
        case RRL_NET_QRISK_TARGET:
            out[n++] = TDesc{4, 0};  // bn1.weight (dead: model.py:175)
            out[n++] = TDesc{4, 0};  // bn1.bias
            for (int i = 0; i < 12; ++i) out[n++] = q[i];
            break;
        case RRL_NET_POLICY:
            for (int i = 0; i < 8; ++i) out[n++] = pol[i];
            break;
        case RRL_NET_RECOVERY:
            for (int i = 0; i < 7; ++i) out[n++] = rec[i];
            break;
        default: break;
    }
    return n;
}

inline int64_t pad4(int64_t x) { return (x + 3) & ~(int64_t)3; }

struct Scratch {
    const char* name;
    int64_t off, count;
};

struct Layout {
    int64_t net_off[RRL_NUM_NETS];
    int64_t net_size[RRL_NUM_NETS];
    int n_tensors[RRL_NUM_NETS];
    int64_t t_off[RRL_NUM_NETS][kMaxTensors];
    TDesc t_desc[RRL_NUM_NETS][kMaxTensors];
    int64_t train_floats;  // size of the grad / m / v blocks
    int64_t grad_off, m_off, v_off;
    int64_t img_off[kNumImages];
    int64_t tc_img_off[kTcHeads];          // fp16 hi/lo operand images for tcgen05 (agent_tc.cu), 65536 floats each
    int64_t tc_imgT_off[kTcHeads];         // the same for W2^T (operand B of the backward dh1 = dh2 W2)
    int64_t R;  // max_batch
    // scratch
    int64_t batch_off[2][5];               // [sac|qr][s,a,r,s2,m]
    int64_t h1[kPassSlots], h2[kPassSlots], dh2[kPassSlots], dh2t[kPassSlots], dh1[kPassSlots];
    int64_t rows_f[40];                    // per-row float arrays of R (see names in agent.cu)
    int64_t rows2_f[12];                   // per-row [R][2] arrays
    int64_t rows4_f[4];                    // per-row [R][4] arrays
    int64_t losses;                        // 16 floats
    int64_t scalars;                       // 32 floats, 8-byte aligned (RRL_S_* / RRL_D_* of rrl.h)
    int64_t total;
    Scratch names[96];
    int n_names;
};

// image index of (net, head)
inline int image_index(int net, int head) {
    switch (net) {
        case RRL_NET_CRITIC: return 0 + head;
        case RRL_NET_CRITIC_TARGET: return 2 + head;
        case RRL_NET_POLICY: return 4;
        case RRL_NET_QRISK: return 5 + head;
        case RRL_NET_QRISK_TARGET: return 7 + head;
        default: return 9;
    }
}
// tensor index of the hidden (256x256) matrix of (net, head)
inline int w2_tensor(int net, int head) {
    switch (net) {
        case RRL_NET_CRITIC:
        case RRL_NET_CRITIC_TARGET: return 2 + 6 * head;
        case RRL_NET_QRISK:
        case RRL_NET_QRISK_TARGET: return 4 + 6 * head;
        case RRL_NET_POLICY: return 2;
        default: return 3;
    }
}

// per-row scratch array ids
enum RowArr {
    RA_NEXT_LOGP = 0, RA_LOGP, RA_QT1, RA_QT2, RA_QF1, RA_QF2, RA_QP1, RA_QP2, RA_TARGET, RA_DQF1, RA_DQF2, RA_DQP1,
    RA_DQP2, RA_MINQ, RA_QR_QT1, RA_QR_QT2, RA_QR_Q1, RA_QR_Q2, RA_QR_TARGET, RA_QR_DQ1, RA_QR_DQ2, RA_QR_NEXT_LOGP,
    RA_REC_Q1, RA_REC_Q2, RA_REC_DQ1, RA_REC_DQ2, RA_REC_LOGP,
    RA_SQ1, RA_SQ2, RA_DSQ1, RA_DSQ2,  // Q_risk(s, pi) of the DGD / update_nu branch (sac.py:221-228) and d/d(raw)
    RA_QS1, RA_QS2,                    // Q_risk(s, a) of the RCPO branch (sac.py:202-205)
    RA_COUNT
};
static_assert(RA_COUNT <= 40, "rows_f too small");
enum Row2Arr { R2_NEXT_A = 0, R2_PI, R2_EPS_CUR, R2_DPI, R2_QR_NEXT_A, R2_REC_PI, R2_REC_EPS, R2_REC_DPI, R2_DPI_B, R2_DPI_S1,
               R2_DPI_S2, R2_COUNT };
static_assert(R2_COUNT <= 12, "rows2_f too small");
enum Row4Arr { R4_RAW_POL = 0, R4_DRAW_POL, R4_RAW_REC, R4_DRAW_REC, R4_COUNT };

inline void add_name(Layout& L, const char* name, int64_t off, int64_t count) {
    if (L.n_names < 96) L.names[L.n_names++] = Scratch{name, off, count};
}

inline Layout make_layout(const rrl_agent_config_t* cfg) {
    Layout L;
    memset(&L, 0, sizeof(L));
    L.R = cfg->max_batch;
    static const int order[RRL_NUM_NETS] = {RRL_NET_CRITIC, RRL_NET_POLICY, RRL_NET_QRISK, RRL_NET_RECOVERY,
                                            RRL_NET_CRITIC_TARGET, RRL_NET_QRISK_TARGET};
    int64_t off = 0;
    for (int oi = 0; oi < RRL_NUM_NETS; ++oi) {
        const int net = order[oi];
        L.net_off[net] = off;
        L.n_tensors[net] = net_tensors(net, L.t_desc[net]);
        for (int t = 0; t < L.n_tensors[net]; ++t) {
            L.t_off[net][t] = off;
            const TDesc d = L.t_desc[net][t];
            off += pad4((int64_t)d.rows * (d.cols ? d.cols : 1));
        }
        L.net_size[net] = off - L.net_off[net];
        if (oi == 3) L.train_floats = off;
    }
    L.grad_off = off;
    off += L.train_floats;
    L.m_off = off;
    off += L.train_floats;
    L.v_off = off;
    off += L.train_floats;
    for (int i = 0; i < kNumImages; ++i) {
        L.img_off[i] = off;
        off += (int64_t)H * H;
    }
    for (int i = 0; i < kTcHeads; ++i) {
        L.tc_img_off[i] = off;
        off += (int64_t)H * H;  // 2 images (hi, lo) x 65536 halves = 65536 floats
    }
    for (int i = 0; i < kTcHeads; ++i) {
        L.tc_imgT_off[i] = off;
        off += (int64_t)H * H;
    }
    const int64_t R = L.R;
    static const char* bn[2][5] = {{"sac_s", "sac_a", "sac_r", "sac_s2", "sac_m"}, {"qr_s", "qr_a", "qr_c", "qr_s2", "qr_m"}};
    static const int bw[5] = {2, 2, 1, 2, 1};
    for (int u = 0; u < 2; ++u)
        for (int f = 0; f < 5; ++f) {
            L.batch_off[u][f] = off;
            add_name(L, bn[u][f], off, R * bw[f]);
            off += pad4(R * bw[f]);
        }
    for (int p = 0; p < kPassSlots; ++p) {
        L.h1[p] = off; off += R * H;
        L.h2[p] = off; off += R * H;
        L.dh2[p] = off; off += R * H;
        L.dh2t[p] = off; off += R * H;
        L.dh1[p] = off; off += R * H;
    }
    static const char* ra[RA_COUNT] = {"next_logp", "logp", "qt1", "qt2", "qf1", "qf2", "qp1", "qp2", "target", "dqf1", "dqf2",
                                       "dqp1", "dqp2", "minq", "qr_qt1", "qr_qt2", "qr_q1", "qr_q2", "qr_target", "qr_dq1",
                                       "qr_dq2", "qr_next_logp", "rec_q1", "rec_q2", "rec_dq1", "rec_dq2", "rec_logp",
                                       "sq1", "sq2", "dsq1", "dsq2", "qs1", "qs2"};
    for (int i = 0; i < RA_COUNT; ++i) {
        L.rows_f[i] = off;
        add_name(L, ra[i], off, R);
        off += pad4(R);
    }
    static const char* r2[R2_COUNT] = {"next_a", "pi", "eps_cur", "dpi", "qr_next_a", "rec_pi", "rec_eps", "rec_dpi", "dpi_b",
                                       "dpi_s1", "dpi_s2"};
    for (int i = 0; i < R2_COUNT; ++i) {
        L.rows2_f[i] = off;
        add_name(L, r2[i], off, 2 * R);
        off += pad4(2 * R);
    }
    static const char* r4[R4_COUNT] = {"raw_pol", "draw_pol", "raw_rec", "draw_rec"};
    for (int i = 0; i < R4_COUNT; ++i) {
        L.rows4_f[i] = off;
        add_name(L, r4[i], off, 4 * R);
        off += 4 * R;
    }
    L.losses = off;
    add_name(L, "losses", off, 16);
    off += 16;
    off = (off + 1) & ~(int64_t)1;  // 8-byte alignment for the float64 slots
    L.scalars = off;
    add_name(L, "scalars", off, 32);
    off += 32;
    L.total = off;
    return L;
}

}  // namespace rrl
