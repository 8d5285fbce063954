// mpc_tc.cu -- the model-based planner's rollout (recovery_rl/MPC.py:374-439, see mpc.cu) on the 5th-gen tensor cores.
//
// Same rows, same noise indexing and the same arithmetic as mpc_rollout_kernel (mpc.cu), with the four 256x256
// contractions of every horizon step -- Q_risk head 1, Q_risk head 2, ensemble layer 1, ensemble layer 2 -- issued as
// tcgen05.mma (fp16 hi/lo split x3, fp32 accumulators in TMEM) by the acting kernel's pipeline (agent_tc.cu):
// 16 producer/epilogue warps (4 threads per row), one MMA thread, one weight-image loader thread, a 3-stage operand
// ring and two 256-column accumulators.  A persistent CTA owns a 128-row tile for the WHOLE horizon (the particles'
// observations stay in registers); rows are ordered (net, env, candidate, particle-in-net) so that a tile shares one
// bootstrap net (TS-infinity).  What is new against the acting kernel: the epilogue of ensemble layer 1 IS the
// producer of layer 2 -- thread (row, q) reads its 64 accumulator columns from TMEM, applies bias + swish and writes
// them straight into the fp16 hi/lo A operand chunks 2q, 2q+1 of the next MMA.
#include "tc_common.cuh"
#include "mpc_layout.cuh"

using namespace rrl;
using namespace rrl::tc;
using namespace rrl::dyn;

namespace {

enum { P_QR1 = 0, P_QR2 = 1, P_E1 = 2, P_E2 = 3 };

struct PlanSmall {
    float W1[3][H][4];   // layer 1 of Q_risk head 1 / 2, layer 0 of the ensemble net (x SA)
    float b1[3][H];
    float b2[4][H];      // bias after the 256x256 contraction of passes QR1, QR2, E1, E2
    float w3[6][H];      // output rows: QR1, QR2, ensemble x4
    float b3[6];
    float norm[12];      // mu[4], sigma[4], max_logvar[2], min_logvar[2]
};
struct PlanTcSmem {
    unsigned char stage[NSTAGE][STAGE_BYTES];
    PlanSmall sm;
    float4 part[2][4][TM];
    unsigned long long full[NSTAGE], empty[NSTAGE], acc_full[2], acc_empty[2];
    uint32_t tmem_base;
};

__device__ __forceinline__ float swishf(float x) { return x / (1.0f + expf(-x)); }
__device__ __forceinline__ float softplusf(float x) { return x > 20.0f ? x : log1pf(expf(x)); }

// fp16 hi/lo tcgen05 images of lin1 / lin2 of every net from the k-major fp32 image (element (n, k) = W[k][n])
__global__ void __launch_bounds__(256) dyn_tc_images_kernel(float* __restrict__ dynimg) {
    const int img = blockIdx.y;                       // layer * NETS + net
    const int layer = img / NETS, net = img % NETS;
    const float* __restrict__ W = dynimg + (layer == 0 ? kW1 : kW2) + (int64_t)net * H * H;
    __half* __restrict__ dst = reinterpret_cast<__half*>(dynimg + tc_img_off(layer, net));
    const int idx = blockIdx.x * 256 + threadIdx.x;   // 0 .. 65535: k = idx / 256, n = idx % 256 (coalesced reads)
    const int k = idx >> 8, n = idx & 255;
    tc_image_store(dst, n, k, W[(int64_t)k * H + n]);
}

__global__ void __launch_bounds__(kTcThreads, 1) mpc_rollout_tc_kernel(const __grid_constant__ MpcTcArgs A) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    PlanTcSmem& S = *reinterpret_cast<PlanTcSmem*>(smem_raw);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int64_t rows_env = (int64_t)A.pop * A.npn;            // rows of one (env, net)
    const int64_t rows_net = A.E * rows_env;                     // rows of one net
    const int64_t tiles_net = (rows_net + TM - 1) / TM;
    const int64_t n_tiles = tiles_net * NETS;

    // ---- one-time setup: the two Q_risk heads, barriers, TMEM ----
    for (int p = 0; p < 2; ++p) {
        const HeadW& w = p == 0 ? A.qr1 : A.qr2;
        for (int k = t; k < H; k += kTcThreads) {
            float4 w1 = *reinterpret_cast<const float4*>(w.W1 + k * 4);
            w1.x *= SA; w1.y *= SA; w1.z *= SA; w1.w *= SA;
            *reinterpret_cast<float4*>(S.sm.W1[p][k]) = w1;
            S.sm.b1[p][k] = w.b1[k] * SA;
            S.sm.b2[p][k] = w.b2[k];
            S.sm.w3[p][k] = w.W3a[k];
        }
        if (t == 0) S.sm.b3[p] = w.b3a[0];
    }
    if (t < 12) S.sm.norm[t] = A.dyn[kMu + t];
    if (t == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(smem_u32(&S.full[s]), kProdWarps + 1);
            mbar_init(smem_u32(&S.empty[s]), 1);
        }
        for (int d = 0; d < 2; ++d) {
            mbar_init(smem_u32(&S.acc_full[d]), 1);
            mbar_init(smem_u32(&S.acc_empty[d]), kProdWarps);
        }
        fence_barrier_init();
    }
    if (warp == kProd / 32) {
        tmem_alloc(smem_u32(&S.tmem_base), 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = S.tmem_base;
    const uint64_t vstep = A.counters ? (uint64_t)A.counters[RRL_C_VEC_STEP] : 0;

    if (warp < kProd / 32) {
        // ========== producer + epilogue: 4 threads (q) per row r ==========
        uint32_t it = 0, acc_use[2] = {0, 0}, n_epi = 0;
        const int q = warp >> 2, r = (warp & 3) * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        int cur_net = -1;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int net = (int)(tile / tiles_net);
            const int64_t g = (tile % tiles_net) * TM + r;       // row inside the net
            const bool in_range = g < rows_net;
            const int64_t e = in_range ? g / rows_env : 0;
            const int64_t j = in_range ? g % rows_env : 0;        // row inside (env, net): candidate c, particle pl
            const int c = (int)(j / A.npn), pl = (int)(j % A.npn);
            const bool live = in_range && (!A.active || A.active[e] != 0);
            if (net != cur_net) {   // (re)stage the ensemble net's small tensors; all MMAs of the previous tile are done
                asm volatile("bar.sync 1, %0;" ::"n"(kProd) : "memory");
                const float* w0 = A.dyn + kW0 + (int64_t)net * DYN_IN * H;
                for (int k = t; k < H; k += kProd) {
                    *reinterpret_cast<float4*>(S.sm.W1[2][k]) = make_float4(w0[k] * SA, w0[H + k] * SA, w0[2 * H + k] * SA, w0[3 * H + k] * SA);
                    S.sm.b1[2][k] = A.dyn[kB0 + net * H + k] * SA;
                    S.sm.b2[2][k] = A.dyn[kB1 + net * H + k];
                    S.sm.b2[3][k] = A.dyn[kB2 + net * H + k];
#pragma unroll
                    for (int o = 0; o < 4; ++o) S.sm.w3[2 + o][k] = A.dyn[kW3 + ((int64_t)net * DYN_OUT + o) * H + k];
                }
                if (t < 4) S.sm.b3[2 + t] = A.dyn[kB3 + net * DYN_OUT + t];
                asm volatile("bar.sync 1, %0;" ::"n"(kProd) : "memory");
                cur_net = net;
            }
            float ox = 0.f, oy = 0.f, cost = 0.f;
            if (in_range) {
                ox = (float)A.state[e];
                oy = (float)A.state[A.E + e];
            }
            // layer-1 style producer: relu (Q_risk) or swish (ensemble layer 0) of W1 x + b1, 8 k-chunks
            auto produce_l1 = [&](int slot, bool swish, float x0, float x1, float x2, float x3) {
                for (int cc = 0; cc < NCHUNK; ++cc, ++it) {
                    const int stage = it % NSTAGE;
                    float hv[8];
#pragma unroll
                    for (int k8 = 0; k8 < 8; ++k8) {
                        const int k = cc * KCH + q * 8 + k8;
                        const float4 wv = *reinterpret_cast<const float4*>(S.sm.W1[slot][k]);
                        float h = fmaf(wv.x, x0, S.sm.b1[slot][k]);
                        h = fmaf(wv.y, x1, h); h = fmaf(wv.z, x2, h); h = fmaf(wv.w, x3, h);
                        // h carries the factor SA: swish(h / SA) * SA keeps the operand scale
                        h = swish ? swishf(h * (1.0f / SA)) * SA : fmaxf(h, 0.f);
                        hv[k8] = fminf(fmaxf(h, -60000.0f), 60000.0f);
                    }
                    uint4 hi, lo;
                    split8(hv, &hi, &lo);
                    mbar_wait(smem_u32(&S.empty[stage]), ((it / NSTAGE) & 1) ^ 1);
                    unsigned char* a_hi = S.stage[stage];
                    *reinterpret_cast<uint4*>(a_hi + q * LBO_A + r * 16) = hi;
                    *reinterpret_cast<uint4*>(a_hi + A_IMG + q * LBO_A + r * 16) = lo;
                    fence_proxy_async();
                    mbar_arrive_warp(smem_u32(&S.full[stage]));
                }
            };
            // epilogue with head outputs: columns [64 q, 64 q + 64) of accumulator d, activation relu / swish
            auto epilogue_heads = [&](int d, int bslot, int w3row, int n_out, bool swish, float raw[4]) {
                mbar_wait(smem_u32(&S.acc_full[d]), acc_use[d] & 1);
                tc_fence_after();
                float out[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
                for (int cc = 0; cc < 2; ++cc) {
                    float v[32];
                    const int col0 = q * 64 + cc * 32;
                    tmem_ld32(lane_addr + d * H + col0, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj) {
                        const float z = fmaf(v[jj], INV_SCALE, S.sm.b2[bslot][col0 + jj]);
                        const float h = swish ? swishf(z) : fmaxf(z, 0.f);
                        out[0] = fmaf(h, S.sm.w3[w3row][col0 + jj], out[0]);
                        if (n_out > 1) {
                            out[1] = fmaf(h, S.sm.w3[w3row + 1][col0 + jj], out[1]);
                            out[2] = fmaf(h, S.sm.w3[w3row + 2][col0 + jj], out[2]);
                            out[3] = fmaf(h, S.sm.w3[w3row + 3][col0 + jj], out[3]);
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive_warp(smem_u32(&S.acc_empty[d]));
                ++acc_use[d];
                const int pb = n_epi & 1;
                ++n_epi;
                S.part[pb][q][r] = make_float4(out[0], out[1], out[2], out[3]);
                asm volatile("bar.sync 1, %0;" ::"n"(kProd) : "memory");
                const float4 p0 = S.part[pb][0][r], p1 = S.part[pb][1][r], p2 = S.part[pb][2][r], p3 = S.part[pb][3][r];
                raw[0] = ((p0.x + p1.x) + (p2.x + p3.x)) + S.sm.b3[w3row];
                if (n_out > 1) {
                    raw[1] = ((p0.y + p1.y) + (p2.y + p3.y)) + S.sm.b3[w3row + 1];
                    raw[2] = ((p0.z + p1.z) + (p2.z + p3.z)) + S.sm.b3[w3row + 2];
                    raw[3] = ((p0.w + p1.w) + (p2.w + p3.w)) + S.sm.b3[w3row + 3];
                }
            };
            for (int step = 0; step < A.hor; ++step) {
                float ax = 0.f, ay = 0.f;
                if (in_range) {
                    const float2 a = *reinterpret_cast<const float2*>(A.samples + ((size_t)e * A.pop + c) * (A.hor * 2) + step * 2);
                    ax = a.x; ay = a.y;
                }
                float raw[4];
                produce_l1(0, false, ox, oy, ax, ay);                               // Q_risk head 1 -> acc 0
                produce_l1(1, false, ox, oy, ax, ay);                               // Q_risk head 2 -> acc 1
                epilogue_heads(0, 0, 0, 1, false, raw);
                const float q1 = sigmoidf_(raw[0]);
                produce_l1(2, true, (ox - S.sm.norm[0]) / S.sm.norm[4], (oy - S.sm.norm[1]) / S.sm.norm[5],
                           (ax - S.sm.norm[2]) / S.sm.norm[6], (ay - S.sm.norm[3]) / S.sm.norm[7]);   // ensemble layer 1 -> acc 0
                epilogue_heads(1, 1, 1, 1, false, raw);
                cost += fmaxf(q1, sigmoidf_(raw[0]));                               // MPC.py:409 on (cur_obs, cur_acs)
                // ---- ensemble layer 1's epilogue produces layer 2's A operand: acc 0 -> chunks 2q, 2q+1 ----
                mbar_wait(smem_u32(&S.acc_full[0]), acc_use[0] & 1);
                tc_fence_after();
                for (int cc = 0; cc < NCHUNK; ++cc, ++it) {
                    const int stage = it % NSTAGE;
                    uint4 hi[4], lo[4];
                    const bool mine = (cc >> 1) == q;
                    if (mine) {
                        float v[32];
                        tmem_ld32(lane_addr + cc * KCH, v);     // acc 0, columns [32 cc, 32 cc + 32)
                        tmem_ld_wait();
#pragma unroll
                        for (int kc = 0; kc < 4; ++kc) {
                            float hv[8];
#pragma unroll
                            for (int k8 = 0; k8 < 8; ++k8) {
                                const int col = cc * KCH + kc * 8 + k8;
                                const float z = fmaf(v[kc * 8 + k8], INV_SCALE, S.sm.b2[2][col]);
                                hv[k8] = fminf(fmaxf(swishf(z) * SA, -60000.0f), 60000.0f);
                            }
                            split8(hv, &hi[kc], &lo[kc]);
                        }
                    }
                    mbar_wait(smem_u32(&S.empty[stage]), ((it / NSTAGE) & 1) ^ 1);
                    if (mine) {
                        unsigned char* a_hi = S.stage[stage];
#pragma unroll
                        for (int kc = 0; kc < 4; ++kc) {
                            *reinterpret_cast<uint4*>(a_hi + kc * LBO_A + r * 16) = hi[kc];
                            *reinterpret_cast<uint4*>(a_hi + A_IMG + kc * LBO_A + r * 16) = lo[kc];
                        }
                    }
                    fence_proxy_async();
                    mbar_arrive_warp(smem_u32(&S.full[stage]));
                }
                tc_fence_before();
                mbar_arrive_warp(smem_u32(&S.acc_empty[0]));
                ++acc_use[0];
                // ---- ensemble layer 2 -> mean, log-variance -> next observation (MPC.py:421-439) ----
                epilogue_heads(1, 3, 2, 4, true, raw);
                {
                    const float mx = S.sm.norm[8], my = S.sm.norm[9], nx = S.sm.norm[10], ny = S.sm.norm[11];
                    float lvx = mx - softplusf(mx - raw[2]), lvy = my - softplusf(my - raw[3]);
                    lvx = nx + softplusf(lvx - nx);
                    lvy = ny + softplusf(lvy - ny);
                    float ex = 0.f, ey = 0.f;
                    if (in_range) {
                        if (A.eps) {
                            const float2 ev = *reinterpret_cast<const float2*>(
                                A.eps + ((((size_t)e * A.hor + step) * NETS + net) * rows_env + j) * 2);
                            ex = ev.x; ey = ev.y;
                        } else {
                            float ee[2];
                            const uint64_t idx = (((uint64_t)e * A.hor + step) * NETS + net) * (uint64_t)rows_env + j;
                            philox_eps(A.seed, A.stream_id, idx, vstep, RRL_DRAW_MPC_EPS + (uint32_t)A.iter * 16u, ee);
                            ex = ee[0]; ey = ee[1];
                        }
                    }
                    ox = ox + fmaf(ex, sqrtf(expf(lvx)), raw[0]);
                    oy = oy + fmaf(ey, sqrtf(expf(lvy)), raw[1]);
                }
            }
            if (live && q == 0) {
                const int p = net * A.npn + pl;
                A.row_cost[((size_t)e * A.pop + c) * A.npart + p] = (cost != cost) ? 1e6f : cost;
            }
        }
    } else if (warp == kProd / 32) {
        // ================= MMA issuer =================
        if (lane == 0) {
            uint32_t it = 0, acc_use[2] = {0, 0};
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int step = 0; step < A.hor; ++step) {
                    for (int p = 0; p < 4; ++p) {
                        const int d = p & 1;       // QR1 -> 0, QR2 -> 1, E1 -> 0, E2 -> 1
                        mbar_wait(smem_u32(&S.acc_empty[d]), (acc_use[d] & 1) ^ 1);
                        tc_fence_after();
                        const uint32_t d_tmem = tmem_base + d * H;
                        for (int cc = 0; cc < NCHUNK; ++cc, ++it) {
                            const int stage = it % NSTAGE;
                            mbar_wait(smem_u32(&S.full[stage]), (it / NSTAGE) & 1);
                            tc_fence_after();
                            const uint32_t a_hi = smem_u32(S.stage[stage]);
                            const uint32_t a_lo = a_hi + A_IMG, b_hi = a_hi + 2 * A_IMG, b_lo = b_hi + B_IMG;
#pragma unroll
                            for (int jj = 0; jj < KCH / 16; ++jj) {
                                const uint64_t dah = make_desc(a_hi + jj * 2 * LBO_A, LBO_A, SBO);
                                const uint64_t dal = make_desc(a_lo + jj * 2 * LBO_A, LBO_A, SBO);
                                const uint64_t dbh = make_desc(b_hi + jj * 2 * LBO_B, LBO_B, SBO);
                                const uint64_t dbl = make_desc(b_lo + jj * 2 * LBO_B, LBO_B, SBO);
                                umma_f16(d_tmem, dah, dbh, (cc | jj) ? 1u : 0u);
                                umma_f16(d_tmem, dah, dbl, 1u);
                                umma_f16(d_tmem, dal, dbh, 1u);
                            }
                            umma_commit(smem_u32(&S.empty[stage]));
                        }
                        umma_commit(smem_u32(&S.acc_full[d]));
                        ++acc_use[d];
                    }
                }
            }
        }
    } else {
        // ================= weight-image loader =================
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int net = (int)(tile / tiles_net);
                const unsigned char* img[4] = {reinterpret_cast<const unsigned char*>(A.qr1.tc_img),
                                               reinterpret_cast<const unsigned char*>(A.qr2.tc_img),
                                               reinterpret_cast<const unsigned char*>(A.dyn + tc_img_off(0, net)),
                                               reinterpret_cast<const unsigned char*>(A.dyn + tc_img_off(1, net))};
                for (int step = 0; step < A.hor; ++step)
                    for (int p = 0; p < 4; ++p)
                        for (int cc = 0; cc < NCHUNK; ++cc, ++it) {
                            const int stage = it % NSTAGE;
                            mbar_wait(smem_u32(&S.empty[stage]), ((it / NSTAGE) & 1) ^ 1);
                            const uint32_t bar = smem_u32(&S.full[stage]);
                            const uint32_t dst = smem_u32(S.stage[stage]) + 2 * A_IMG;
                            mbar_arrive_expect_tx(bar, 2 * B_IMG);
                            bulk_g2s(dst, img[p] + (size_t)cc * 2 * B_IMG, B_IMG, bar);
                            bulk_g2s(dst + B_IMG, img[p] + (size_t)cc * 2 * B_IMG + B_IMG, B_IMG, bar);
                        }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kProd / 32) tmem_dealloc(tmem_base, 512);
}

}  // namespace

namespace rrl {

int dyn_tc_images_launch(float* dyn_image, cudaStream_t st) {
    dyn_tc_images_kernel<<<dim3(H * H / 256, 2 * NETS), 256, 0, st>>>(dyn_image);
    RRL_CHECK_LAUNCH();
    return 0;
}

int mpc_rollout_tc_launch(const MpcTcArgs& T, cudaStream_t st) {
    static bool configured = false;
    const size_t smem = sizeof(PlanTcSmem);
    if (!configured) {
        RRL_CUDA(cudaFuncSetAttribute(mpc_rollout_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const int64_t rows_net = T.E * (int64_t)T.pop * T.npn;
    const int64_t tiles = ((rows_net + TM - 1) / TM) * NETS;
    const int64_t sms = rrl_num_sms();
    const int grid = (int)(tiles < sms ? tiles : sms);
    mpc_rollout_tc_kernel<<<grid, kTcThreads, smem, st>>>(T);
    RRL_CHECK_LAUNCH();
    return 0;
}

}  // namespace rrl
