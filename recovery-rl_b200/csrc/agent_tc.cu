// agent_tc.cu -- the acting kernel (experiment.py:546-577) with the 256x256 contractions on the 5th-gen
// tensor cores: tcgen05.mma (kind::f16, M128 N256 K16) issued by one thread, operands staged in shared
// memory, accumulators in TMEM, epilogue through tcgen05.ld.
//
// Precision: the reference computes these layers in fp32 and the parity bar is 1e-4, so a single fp16/bf16
// MMA is not enough.  Both operands are split into two fp16 terms (x*S = hi + lo, S a power of two that
// keeps lo out of the subnormal range) and three MMAs are accumulated in fp32:
//     D += Ahi*Bhi ;  D += Ahi*Blo ;  D += Alo*Bhi          (the dropped Alo*Blo term is ~2^-22 relative)
// which reproduces the fp32 product to ~1e-6 relative -- measured against the SIMT path in the tests.
//
// Per CTA (one per SM, persistent over 128-row tiles), 18 warps:
//   warps 0-15 four threads per env row (TMEM lane quadrant = warp % 4, quarter q = warp / 4): layer 1
//              (K = 2|4, SIMT) -> fp16 hi/lo A tiles written straight into the UMMA canonical K-major layout
//              (thread q writes core column q of each 32-wide k-chunk); later the epilogue of the same row:
//              columns [64q, 64q+64) through tcgen05.ld, +b2, ReLU, partial head sums exchanged through smem,
//              then the tanh-Gaussian / sigmoid / recovery maths and the action select
//   warp 16    TMEM allocation + the single MMA-issuing thread
//   warp 17    one thread streams the pre-split weight images (32 KB per k-chunk) with cp.async.bulk
// Pipelines: a 3-stage smem ring (full/empty mbarriers; tcgen05.commit frees a stage) and two 256-column
// TMEM accumulators (acc_full / acc_empty) so the epilogue of one network overlaps the MMAs of the next.
// Pass order per tile: policy, recovery, Q_risk head 1, Q_risk head 2 (the two state-only networks first so
// the tensor pipe has work while the policy epilogue produces the action the Q_risk passes need).
#include <type_traits>
#include "tc_common.cuh"

using namespace rrl;
using namespace rrl::tc;

namespace {

#ifdef RRL_TC_TIMING
// profiling build (-DRRL_TC_TIMING): per-launch, per-CTA stage time stamps of the update kernels (globaltimer, ns)
//   g_tc_log[launch % 256][cta % 16][0..7]: 0 start, 1 setup done, 2 producers done, 3 accumulator ready, 4 epilogue
//   done, 5 TMEM released, 6 tail done;  g_tc_kind[launch % 256]: 1 fwd, 2 bwd (+ grid dims)
__device__ unsigned long long g_tc_log[256][16][8];
__device__ unsigned int g_tc_launch;
__device__ int g_tc_kind[256][4];
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define TSTAMP(i) do { if (threadIdx.x == 0) g_tc_log[*(volatile unsigned int*)&g_tc_launch & 255][(blockIdx.y * gridDim.x + blockIdx.x) & 15][i] = gtime(); } while (0)
#define TLAUNCH_END(kind) do { __syncthreads(); if (threadIdx.x == 0) { const unsigned l = *(volatile unsigned int*)&g_tc_launch & 255; \
        if (blockIdx.x == 0 && blockIdx.y == 0) { g_tc_kind[l][0] = kind; g_tc_kind[l][1] = gridDim.x; g_tc_kind[l][2] = gridDim.y; } \
        __threadfence(); const unsigned done = atomicAdd(&g_tc_done, 1u); \
        if (done == gridDim.x * gridDim.y - 1) { g_tc_done = 0; __threadfence(); atomicAdd(&g_tc_launch, 1u); } } } while (0)
__device__ unsigned int g_tc_done;
#else
#define TSTAMP(i)
#define TLAUNCH_END(kind)
#endif

enum { PASS_POL = 0, PASS_REC = 1, PASS_QR1 = 2, PASS_QR2 = 3 };
constexpr int kActProdWarps = 8, kActEpiWarps = 8;   // acting kernel: warps 0-7 stage operands, warps 8-15 drain accumulators

struct TcSmall {
    float W1[4][H][4];
    float b1[4][H];
    float b2[4][H];
    float w3[4][4][H];
    float b3[4][4];
    float log_std[2];
};

struct TcSmem {
    unsigned char stage[NSTAGE][STAGE_BYTES];
    TcSmall sm;
    float4 part[2][2][TM];   // per-row partial head sums of the 2 column halves (double-buffered)
    float2 act[TM];          // the tile's task actions: epilogue role (lane = row) -> producer role (two other rows)
    unsigned long long full[NSTAGE], empty[NSTAGE], acc_full[2], acc_empty[2];
    uint32_t tmem_base;
};

// ---- weight images -----------------------------------------------------------------------------------
// element (n, k) of W2[n][k] * SB -> fp16 hi / lo at halves index
//   (((c * 2 + hl) * 4 + kc) * 32 + g) * 64 + r * 8 + e,   c = k / 32, kc = (k % 32) / 8, e = k % 8, g = n / 8, r = n % 8
struct TcImgArgs {
    const float* W2[kTcHeads];
    __half* img[kTcHeads];
};
__global__ void __launch_bounds__(256) tc_images_kernel(const __grid_constant__ TcImgArgs A) {
    const int head = blockIdx.y;
    const float* __restrict__ W = A.W2[head];
    __half* __restrict__ img = A.img[head];
    // one thread per (n, 8 consecutive k): coalesced 32 B reads, 16 B writes
    const int idx = blockIdx.x * 256 + threadIdx.x;  // 0 .. 256*32
    const int n = idx >> 5, k8 = idx & 31;
    const float4 v0 = *reinterpret_cast<const float4*>(W + n * H + k8 * 8);
    const float4 v1 = *reinterpret_cast<const float4*>(W + n * H + k8 * 8 + 4);
    const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    __align__(16) __half hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const float s = fmaxf(fminf(v[e] * SB, 60000.0f), -60000.0f);
        const __half h = __float2half_rn(s);
        hi[e] = h;
        lo[e] = __float2half_rn(s - __half2float(h));
    }
    const int c = k8 >> 2, kc = k8 & 3, g = n >> 3, r = n & 7;
    const size_t base_hi = ((((size_t)c * 2 + 0) * 4 + kc) * 32 + g) * 64 + r * 8;
    const size_t base_lo = ((((size_t)c * 2 + 1) * 4 + kc) * 32 + g) * 64 + r * 8;
    *reinterpret_cast<uint4*>(img + base_hi) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(img + base_lo) = *reinterpret_cast<const uint4*>(lo);
}

// ---- the acting kernel -----------------------------------------------------------------------------
// Stages of the composite action (experiment.py:546-577).  All three in one launch reproduce the fused kernel; the vector
// engine launches them separately so that each runs as soon as ITS networks have been stepped, next to the remaining updates
// (task policy after the SAC step, Q_risk after the safety-critic step, recovery policy + select after the recovery step):
//   ACT_STAGE_POLICY    task action (policy pass, or the uniform random action of the start phase)  -> action_task
//   ACT_STAGE_QRISK     Q_risk(s, action_task) twin pass -> qrisk_out, recovery flag          (reads action_task)
//   ACT_STAGE_RECOVERY  recovery policy pass, select  -> action_real          (reads action_task, recovery flag)
struct TcActArgs {
    ActArgs a;
    const __half* img[4];  // by pass: POL, REC, QR1, QR2
    int stages;            // RRL_ACT_STAGE_* bits
};

__device__ __forceinline__ const HeadW& pass_head(const ActArgs& a, int p) {
    return p == PASS_POL ? a.pol : (p == PASS_REC ? a.rec : (p == PASS_QR1 ? a.qr1 : a.qr2));
}

// producer: layer 1 of `pass`, the 8 hidden units of core column q of k-chunk c for TWO rows, as fp16 hi/lo pieces of the
// canonical A layout (registers).  W1 / b1 are staged pre-multiplied by SA (a power of two: exact) and read once for both rows.
template <bool FOUR>
__device__ __forceinline__ void layer1_chunk2(const TcSmem& S, int pass, int c, int q, const float (&xa)[4], const float (&xb)[4],
                                              uint4* hia, uint4* loa, uint4* hib, uint4* lob) {
    float ha[8], hb[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int k = c * KCH + q * 8 + e;
        const float4 wv = *reinterpret_cast<const float4*>(S.sm.W1[pass][k]);
        const float b = S.sm.b1[pass][k];
        float u = fmaf(wv.x, xa[0], b), v = fmaf(wv.x, xb[0], b);
        u = fmaf(wv.y, xa[1], u); v = fmaf(wv.y, xb[1], v);
        if (FOUR) {   // compile-time: a predicated FFMA still takes its issue slot
            u = fmaf(wv.z, xa[2], u); v = fmaf(wv.z, xb[2], v);
            u = fmaf(wv.w, xa[3], u); v = fmaf(wv.w, xb[3], v);
        }
        ha[e] = fminf(fmaxf(u, 0.f), 60000.0f);
        hb[e] = fminf(fmaxf(v, 0.f), 60000.0f);
    }
    split8(ha, hia, loa);
    split8(hb, hib, lob);
}

// epilogue: N_BLK 32-column blocks starting at column col_begin of this thread's row of an accumulator -> partial head sums
// (bias + ReLU of layer 2, dot products with the N_OUT head rows)
template <int N_OUT, int N_BLK>
__device__ __forceinline__ float4 epilogue_cols(TcSmem& S, int pass, uint32_t taddr, int col_begin) {
    float out[4] = {0.f, 0.f, 0.f, 0.f};
    const float4* b2v = reinterpret_cast<const float4*>(S.sm.b2[pass]);
    const float4* w0v = reinterpret_cast<const float4*>(S.sm.w3[pass][0]);
    const float4* w1v = reinterpret_cast<const float4*>(S.sm.w3[pass][1]);
    const float4* w2v = reinterpret_cast<const float4*>(S.sm.w3[pass][2]);
    const float4* w3v = reinterpret_cast<const float4*>(S.sm.w3[pass][3]);
#pragma unroll 1
    for (int cc = 0; cc < N_BLK; ++cc) {
        float v[32];
        const int col0 = col_begin + cc * 32;
        tmem_ld32(taddr + col0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
            const int i4 = (col0 >> 2) + j4;
            const float4 b = b2v[i4];
            const float h0 = fmaxf(fmaf(v[4 * j4 + 0], INV_SCALE, b.x), 0.f);
            const float h1 = fmaxf(fmaf(v[4 * j4 + 1], INV_SCALE, b.y), 0.f);
            const float h2 = fmaxf(fmaf(v[4 * j4 + 2], INV_SCALE, b.z), 0.f);
            const float h3 = fmaxf(fmaf(v[4 * j4 + 3], INV_SCALE, b.w), 0.f);
            const float4 wa = w0v[i4];
            out[0] = fmaf(h3, wa.w, fmaf(h2, wa.z, fmaf(h1, wa.y, fmaf(h0, wa.x, out[0]))));
            if (N_OUT > 1) {
                const float4 wb = w1v[i4];
                out[1] = fmaf(h3, wb.w, fmaf(h2, wb.z, fmaf(h1, wb.y, fmaf(h0, wb.x, out[1]))));
            }
            if (N_OUT > 2) {
                const float4 wc = w2v[i4], wd = w3v[i4];
                out[2] = fmaf(h3, wc.w, fmaf(h2, wc.z, fmaf(h1, wc.y, fmaf(h0, wc.x, out[2]))));
                out[3] = fmaf(h3, wd.w, fmaf(h2, wd.z, fmaf(h1, wd.y, fmaf(h0, wd.x, out[3]))));
            }
        }
    }
    return make_float4(out[0], out[1], out[2], out[3]);
}

__global__ void __launch_bounds__(kTcThreads, 1) act_tc_kernel(const __grid_constant__ TcActArgs T) {
    // declared aligned and used WITHOUT pointer arithmetic: rounding the address up through uintptr_t makes the
    // compiler lose the shared address space and emit generic LD / ST (long-scoreboard) instead of LDS / STS
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    TcSmem& S = *reinterpret_cast<TcSmem*>(smem_raw);
    const ActArgs& A = T.a;
    pdl_wait();   // programmatic dependent launch (common.cuh)
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int64_t n_tiles = (A.n + TM - 1) / TM;
    // passes of this launch, in issue order (POL, REC, QR1, QR2); accumulator of a pass = its position & 1
    const bool sP = (T.stages & RRL_ACT_STAGE_POLICY) != 0;
    const bool sR = A.use_recovery && (T.stages & RRL_ACT_STAGE_RECOVERY) != 0;
    const bool sQ = A.use_recovery && (T.stages & RRL_ACT_STAGE_QRISK) != 0;
    int plist[4], n_pass = 0;
    if (sP) plist[n_pass++] = PASS_POL;
    if (sR) plist[n_pass++] = PASS_REC;
    if (sQ) { plist[n_pass++] = PASS_QR1; plist[n_pass++] = PASS_QR2; }

    // ---- one-time setup: small tensors, barriers, TMEM ----
    for (int pi = 0; pi < n_pass; ++pi) {
        const int p = plist[pi];
        const HeadW& w = pass_head(A, p);
        for (int k = t; k < H; k += kTcThreads) {
            float4 w1 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (w.n_in == 4) {
                w1 = *reinterpret_cast<const float4*>(w.W1 + k * 4);
            } else {
                const float2 v = *reinterpret_cast<const float2*>(w.W1 + k * 2);
                w1.x = v.x; w1.y = v.y;
            }
            w1.x *= SA; w1.y *= SA; w1.z *= SA; w1.w *= SA;   // exact (power of two): h1 * SA comes out of layer 1
            *reinterpret_cast<float4*>(S.sm.W1[p][k]) = w1;
            S.sm.b1[p][k] = w.b1[k] * SA;
            S.sm.b2[p][k] = w.b2[k];
            for (int o = 0; o < 4; ++o) {
                float v = 0.f;
                if (o < w.na) v = w.W3a[o * H + k];
                else if (o < w.na + w.nb) v = w.W3b[(o - w.na) * H + k];
                S.sm.w3[p][o][k] = v;
            }
        }
        if (t < 4) {
            float v = 0.f;
            if (t < w.na) v = w.b3a[t];
            else if (t < w.na + w.nb) v = w.b3b[t - w.na];
            S.sm.b3[p][t] = v;
        }
    }
    if (t < 2 && sR) S.sm.log_std[t] = A.rec.log_std[t];
    if (t == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(smem_u32(&S.full[s]), kActProdWarps + 1);   // the producer warps + the loader's expect_tx arrival
            mbar_init(smem_u32(&S.empty[s]), 1);      // tcgen05.commit
        }
        for (int d = 0; d < 2; ++d) {
            mbar_init(smem_u32(&S.acc_full[d]), 1);   // tcgen05.commit
            mbar_init(smem_u32(&S.acc_empty[d]), kActEpiWarps);   // one arrival per epilogue warp
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
    const bool random_phase = A.counters && !A.eval && (A.start_steps > A.counters[RRL_C_TOTAL_NUMSTEPS]);

    if (warp < kProd / 32) {
        // ========== producer warps (0-7) and epilogue warps (8-15) ==========
        // The CTA's work is a flat list of items k = (tile, pass) in issue order; item k accumulates in TMEM buffer k & 1.
        //   producer warps: a thread stages core column pq of TWO rows (pa, pa + 64) of every k-chunk (W1 / b1 are read from
        //                   shared memory once for both rows); they never touch TMEM and run ahead of the MMAs by the depth of
        //                   the operand ring, across passes and tiles
        //   epilogue warps: two per TMEM lane quadrant (lane = row r), each draining one 128-column half of the accumulator:
        //                   bias + ReLU + head dot products, the halves exchanged through shared memory, then the head maths
        // The only hand-off from the epilogue to the producer role is the tile's task action (the Q_risk passes' input), through
        // S.act and named barrier 5 (all 512 threads): the epilogue warps arrive after the policy head of tile t, the producers
        // before staging its first Q_risk chunk.
        const int64_t my_tiles = (int64_t)blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
        const int64_t n_items = my_tiles * n_pass;
        if (warp < kActProdWarps) {
            const int pq = t >> 6, pa = t & 63, pb2 = pa + 64;
            float xa[4] = {0.f, 0.f, 0.f, 0.f}, xb[4] = {0.f, 0.f, 0.f, 0.f};   // (s, a) of the thread's two rows, current tile
            // the states of a tile are fetched one tile ahead (fp64 loads from HBM / L2: ~1 us if waited for at the tile's start)
            double nsa0 = 0.0, nsa1 = 0.0, nsb0 = 0.0, nsb1 = 0.0;
            auto fetch_state = [&](int64_t tile) {
                const int64_t ra = tile * TM + pa, rb = tile * TM + pb2;
                nsa0 = nsa1 = nsb0 = nsb1 = 0.0;
                if (tile < n_tiles) {
                    if (ra < A.n) { nsa0 = A.state[ra]; nsa1 = A.state[A.n + ra]; }
                    if (rb < A.n) { nsb0 = A.state[rb]; nsb1 = A.state[A.n + rb]; }
                }
            };
            fetch_state(blockIdx.x);
            for (int64_t k = 0; k < n_items; ++k) {
                const int pos = (int)(k % n_pass);
                const int64_t tile = blockIdx.x + (k / n_pass) * gridDim.x;
                const int pass = plist[pos];
                const int64_t ra = tile * TM + pa, rb = tile * TM + pb2;
                if (pos == 0) {   // torch.FloatTensor(state): fp64 -> fp32 (sac.py:137)
                    xa[0] = (float)nsa0; xa[1] = (float)nsa1; xb[0] = (float)nsb0; xb[1] = (float)nsb1;
                    fetch_state(tile + gridDim.x);
                }
                if (pass == PASS_QR1) {   // the task actions of the tile
                    xa[2] = xa[3] = xb[2] = xb[3] = 0.f;
                    if (sP) {
                        asm volatile("bar.sync 5, %0;" ::"n"(kProd) : "memory");     // written by the epilogue warps (policy head)
                        const float2 aa = S.act[pa], ab = S.act[pb2];
                        xa[2] = aa.x; xa[3] = aa.y; xb[2] = ab.x; xb[3] = ab.y;
                    } else {              // a later stage: from the policy stage's launch
                        if (ra < A.n) { const float2 aa = reinterpret_cast<const float2*>(A.action_task)[ra]; xa[2] = aa.x; xa[3] = aa.y; }
                        if (rb < A.n) { const float2 ab = reinterpret_cast<const float2*>(A.action_task)[rb]; xb[2] = ab.x; xb[3] = ab.y; }
                    }
                }
                const bool four = pass >= PASS_QR1;
                const uint32_t it = (uint32_t)k * NCHUNK;    // running chunk index: stage = (it + c) % NSTAGE
                // software pipeline: the next chunk is computed between the stores of a chunk and their proxy fence
                auto produce = [&](auto four_c) {
                    constexpr bool FOUR = decltype(four_c)::value;
                    uint4 hia, loa, hib, lob;
                    layer1_chunk2<FOUR>(S, pass, 0, pq, xa, xb, &hia, &loa, &hib, &lob);
                    for (int c = 0; c < NCHUNK; ++c) {
                        const uint32_t ic = it + c;
                        const int stage = ic % NSTAGE;
                        mbar_wait(smem_u32(&S.empty[stage]), ((ic / NSTAGE) & 1) ^ 1);
                        unsigned char* a_hi = S.stage[stage];
                        *reinterpret_cast<uint4*>(a_hi + pq * LBO_A + pa * 16) = hia;
                        *reinterpret_cast<uint4*>(a_hi + A_IMG + pq * LBO_A + pa * 16) = loa;
                        *reinterpret_cast<uint4*>(a_hi + pq * LBO_A + pb2 * 16) = hib;
                        *reinterpret_cast<uint4*>(a_hi + A_IMG + pq * LBO_A + pb2 * 16) = lob;
                        if (c + 1 < NCHUNK) layer1_chunk2<FOUR>(S, pass, c + 1, pq, xa, xb, &hia, &loa, &hib, &lob);
                        fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
                        mbar_arrive_warp(smem_u32(&S.full[stage]));
                    }
                };
                if (four) produce(std::true_type{});
                else produce(std::false_type{});
            }
        } else {
            const int ew = warp - kActProdWarps;                 // 0..7
            const int hq = ew >> 2, r = (ew & 3) * 32 + lane;    // column half, row (TMEM lane quadrant = warp % 4 = ew % 4)
            const uint32_t lane_addr = tmem_base + ((uint32_t)((ew & 3) * 32) << 16);
            uint32_t acc_use[2] = {0, 0};
            uint32_t n_epi = 0;         // epilogues done (selects the partial-sum buffer)
            float at[2] = {0.f, 0.f}, arec[2] = {0.f, 0.f};      // state of the tile whose accumulators are being drained
            float q1 = 0.f, qmax = 0.f;
            for (int64_t k = 0; k < n_items; ++k) {
                const int pos = (int)(k % n_pass);
                const int64_t tile = blockIdx.x + (k / n_pass) * gridDim.x;
                const int pass = plist[pos];
                const int d = (int)(k & 1);
                const int64_t row = tile * TM + r;
                const bool live = row < A.n;
                float raw[4];
                {
                    mbar_wait(smem_u32(&S.acc_full[d]), acc_use[d] & 1);
                    tc_fence_after();
                    const uint32_t taddr = lane_addr + d * H;
                    float4 mine;
                    if (pass == PASS_POL) mine = epilogue_cols<4, 4>(S, PASS_POL, taddr, hq * 128);
                    else if (pass == PASS_REC) mine = epilogue_cols<2, 4>(S, PASS_REC, taddr, hq * 128);
                    else mine = epilogue_cols<1, 4>(S, pass, taddr, hq * 128);
                    tc_fence_before();
                    mbar_arrive_warp(smem_u32(&S.acc_empty[d]));
                    ++acc_use[d];
                    // exchange the two column halves of the row; both threads of the row then hold the full sums
                    const int pb = n_epi & 1;
                    ++n_epi;
                    S.part[pb][hq][r] = mine;
                    // only the two warps of this TMEM lane quadrant exchange (rows r of quadrant ew & 3)
                    asm volatile("bar.sync %0, 64;" ::"r"(1 + (ew & 3)) : "memory");
                    const float4 p0 = S.part[pb][0][r], p1 = S.part[pb][1][r];
                    raw[0] = (p0.x + p1.x) + S.sm.b3[pass][0];
                    raw[1] = (p0.y + p1.y) + S.sm.b3[pass][1];
                    raw[2] = (p0.z + p1.z) + S.sm.b3[pass][2];
                    raw[3] = (p0.w + p1.w) + S.sm.b3[pass][3];
                }
                if (pos == 0 && !sP) {   // a later stage: the task action of an earlier launch
                    at[0] = at[1] = 0.f;
                    if (live) {
                        const float2 av = reinterpret_cast<const float2*>(A.action_task)[row];
                        at[0] = av.x; at[1] = av.y;
                    }
                }
                if (pass == PASS_POL) {
                    // ---- policy head (model.py:325-338) ----
                    if (!random_phase) {
                        float e[2], mean_a[2], lp;
                        if (A.eps_task) {
                            const float2 ev = live ? reinterpret_cast<const float2*>(A.eps_task)[row] : make_float2(0.f, 0.f);
                            e[0] = ev.x; e[1] = ev.y;
                        } else {
                            philox_eps(A.seed, A.stream_id, (uint64_t)row, vstep, RRL_DRAW_ACT_TASK, e);
                            if (A.det) det_noise(e);
                        }
                        if (A.det) {  // DeterministicPolicy.sample (model.py:475-481)
                            const float zero_ls[2] = {0.f, 0.f};
                            stoch_sample(raw, zero_ls, e, A.sp, at, mean_a, &lp);
                        } else {
                            gauss_sample(raw, e, A.sp, at, &lp, mean_a);
                        }
                        if (A.eval) { at[0] = mean_a[0]; at[1] = mean_a[1]; }
                    } else {  // env.action_space.sample() (experiment.py:559-560)
                        float u[2] = {0.f, 0.f};
                        if (A.rand_u) {
                            if (live) {
                                const float2 uv = reinterpret_cast<const float2*>(A.rand_u)[row];
                                u[0] = uv.x; u[1] = uv.y;
                            }
                        } else {
                            const Philox4 p = rrl_philox(A.seed, A.stream_id, (uint64_t)row, vstep, RRL_DRAW_ACT_RAND);
                            u[0] = rrl_u24(p.x); u[1] = rrl_u24(p.y);
                        }
                        at[0] = fmaf(2.0f * u[0] - 1.0f, A.sp.scale[0], A.sp.bias[0]);
                        at[1] = fmaf(2.0f * u[1] - 1.0f, A.sp.scale[1], A.sp.bias[1]);
                    }
                    if (live && hq == 0) reinterpret_cast<float2*>(A.action_task)[row] = make_float2(at[0], at[1]);
                    if (sQ) {   // -> the producer warps (first Q_risk chunk of this tile).  S.act is rewritten by the NEXT tile's
                                // policy head, which needs that tile's policy MMAs, i.e. the producers past their read of this one
                        if (hq == 0) S.act[r] = make_float2(at[0], at[1]);
                        asm volatile("bar.sync 5, %0;" ::"n"(kProd) : "memory");
                    }
                } else if (pass == PASS_REC) {
                    // ---- recovery policy head (model.py:512-525) ----
                    float e[2], mean_a[2], lp;
                    if (A.eps_rec) {
                        const float2 ev = live ? reinterpret_cast<const float2*>(A.eps_rec)[row] : make_float2(0.f, 0.f);
                        e[0] = ev.x; e[1] = ev.y;
                    } else {
                        philox_eps(A.seed, A.stream_id, (uint64_t)row, vstep, RRL_DRAW_ACT_REC, e);
                    }
                    stoch_sample(raw, S.sm.log_std, e, A.sp, arec, mean_a, &lp);
                } else if (pass == PASS_QR1) {
                    q1 = sigmoidf_(raw[0]);
                } else {
                    qmax = fmaxf(q1, sigmoidf_(raw[0]));   // qrisk.py:196
                }
                if (pos == n_pass - 1 && live && hq == 0) {   // last pass of the tile: decide and write
                    bool rec = false;
                    if (sQ) {
                        rec = qmax > A.eps_safe;               // experiment.py:555
                        if (A.recovery) A.recovery[row] = rec ? 1 : 0;
                        if (A.qrisk_out) A.qrisk_out[row] = qmax;
                    } else if (sR) {
                        rec = A.recovery[row] != 0;            // decided by the Q_risk stage of an earlier launch
                    }
                    // the executed action: decided once the recovery stage has run (or at once without a recovery policy)
                    if (sR || !A.use_recovery)
                        reinterpret_cast<float2*>(A.action_real)[row] = rec ? make_float2(arec[0], arec[1]) : make_float2(at[0], at[1]);
                }
            }
        }
    } else if (warp == kProd / 32) {
        // ================= MMA issuer (one thread) =================
        if (lane == 0) {
            uint32_t it = 0, item = 0;
            uint32_t acc_use[2] = {0, 0};
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int pi = 0; pi < n_pass; ++pi, ++item) {
                    const int d = item & 1;  // accumulator = item parity (four passes per tile: POL -> 0, REC -> 1, QR1 -> 0, QR2 -> 1)
                    mbar_wait(smem_u32(&S.acc_empty[d]), (acc_use[d] & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + d * H;
                    for (int c = 0; c < NCHUNK; ++c, ++it) {
                        const int stage = it % NSTAGE;
                        mbar_wait(smem_u32(&S.full[stage]), (it / NSTAGE) & 1);
                        tc_fence_after();
                        const uint32_t a_hi = smem_u32(S.stage[stage]);
                        const uint32_t a_lo = a_hi + A_IMG, b_hi = a_hi + 2 * A_IMG, b_lo = b_hi + B_IMG;
#pragma unroll
                        for (int j = 0; j < KCH / 16; ++j) {
                            const uint64_t dah = make_desc(a_hi + j * 2 * LBO_A, LBO_A, SBO);
                            const uint64_t dal = make_desc(a_lo + j * 2 * LBO_A, LBO_A, SBO);
                            const uint64_t dbh = make_desc(b_hi + j * 2 * LBO_B, LBO_B, SBO);
                            const uint64_t dbl = make_desc(b_lo + j * 2 * LBO_B, LBO_B, SBO);
                            umma_f16(d_tmem, dah, dbh, (c | j) ? 1u : 0u);
                            umma_f16(d_tmem, dah, dbl, 1u);
                            umma_f16(d_tmem, dal, dbh, 1u);
                        }
                        umma_commit(smem_u32(&S.empty[stage]));  // stage free once these MMAs have read it
                    }
                    umma_commit(smem_u32(&S.acc_full[d]));       // accumulator complete
                    ++acc_use[d];
                }
            }
        }
    } else {
        // ================= weight-image loader (one thread) =================
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int pi = 0; pi < n_pass; ++pi) {
                    const unsigned char* img = reinterpret_cast<const unsigned char*>(T.img[plist[pi]]);
                    for (int c = 0; c < NCHUNK; ++c, ++it) {
                        const int stage = it % NSTAGE;
                        mbar_wait(smem_u32(&S.empty[stage]), ((it / NSTAGE) & 1) ^ 1);
                        const uint32_t bar = smem_u32(&S.full[stage]);
                        const uint32_t dst = smem_u32(S.stage[stage]) + 2 * A_IMG;
                        mbar_arrive_expect_tx(bar, 2 * B_IMG);
                        bulk_g2s(dst, img + (size_t)c * 2 * B_IMG, B_IMG, bar);
                        bulk_g2s(dst + B_IMG, img + (size_t)c * 2 * B_IMG + B_IMG, B_IMG, bar);
                    }
                }
            }
        }
    }
    // ---- teardown ----
    tc_fence_before();
    __syncthreads();
    if (warp == kProd / 32) tmem_dealloc(tmem_base, 512);
}


// ---- grouped forward passes of the updates on tcgen05 -------------------------------------------------
// One CTA = one (pass, 128-row tile): the same producer / MMA / loader / epilogue pipeline as the acting kernel,
// for the batched forward passes of SAC.update_parameters / QRiskWrapper.update_parameters (rows = batch size).
// Producers also store h1, the epilogue h2 (fp32, [row][256]) when the backward pass needs them.
struct FwdTcSmem {
    unsigned char stage[NSTAGE][STAGE_BYTES];
    float W1[H][4];
    float b1[H], b2[H];
    float w3[4][H];
    float b3[4];
    float4 part[4][TM];
    unsigned long long full[NSTAGE], empty[NSTAGE], acc_full;
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(kTcThreads, 1) fwd_tc_kernel(const __grid_constant__ FwdArgs A) {
    // declared aligned and used WITHOUT pointer arithmetic: rounding the address up through uintptr_t makes the
    // compiler lose the shared address space and emit generic LD / ST (long-scoreboard) instead of LDS / STS
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    FwdTcSmem& S = *reinterpret_cast<FwdTcSmem*>(smem_raw);
    pdl_wait();   // programmatic dependent launch (common.cuh): the previous kernel of the stream has completed past here
    const int64_t rows = A.rows_ptr ? *A.rows_ptr : A.rows_const;
    const int64_t row0 = (int64_t)blockIdx.x * TM;
    TSTAMP(0);
    if (row0 >= rows) {                             // uniform: before any barrier / TMEM allocation
        run_tail(A.tail);
        TLAUNCH_END(1);
        return;
    }
    const FwdPass& P = A.p[blockIdx.y];
    const HeadW& w = P.w;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    for (int k = t; k < H; k += kTcThreads) {
        float4 w1 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (w.n_in == 4) {
            w1 = *reinterpret_cast<const float4*>(w.W1 + k * 4);
        } else {
            const float2 v = *reinterpret_cast<const float2*>(w.W1 + k * 2);
            w1.x = v.x; w1.y = v.y;
        }
        w1.x *= SA; w1.y *= SA; w1.z *= SA; w1.w *= SA;
        *reinterpret_cast<float4*>(S.W1[k]) = w1;
        S.b1[k] = w.b1[k] * SA;
        S.b2[k] = w.b2[k];
        for (int o = 0; o < 4; ++o) {
            float v = 0.f;
            if (o < w.na) v = w.W3a[o * H + k];
            else if (o < w.na + w.nb) v = w.W3b[(o - w.na) * H + k];
            S.w3[o][k] = v;
        }
    }
    if (t < 4) {
        float v = 0.f;
        if (t < w.na) v = w.b3a[t];
        else if (t < w.na + w.nb) v = w.b3b[t - w.na];
        S.b3[t] = v;
    }
    if (t == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(smem_u32(&S.full[s]), kProdWarps + 1);
            mbar_init(smem_u32(&S.empty[s]), 1);
        }
        mbar_init(smem_u32(&S.acc_full), 1);
        fence_barrier_init();
    }
    if (warp == kProd / 32) {
        tmem_alloc(smem_u32(&S.tmem_base), 256);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = S.tmem_base;
    TSTAMP(1);

    if (warp < kProd / 32) {
        const int q = warp >> 2, r = (warp & 3) * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        const int64_t row = row0 + r;
        const bool live = row < rows;
        float x0 = 0.f, x1 = 0.f, x2 = 0.f, x3 = 0.f;
        if (live) {
            const float2 sv = reinterpret_cast<const float2*>(P.xs)[row];
            x0 = sv.x; x1 = sv.y;
            if (P.xa) {
                const float2 av = reinterpret_cast<const float2*>(P.xa)[row];
                x2 = av.x; x3 = av.y;
            }
        }
        const bool four = w.n_in == 4;
        // h1 image of this warp's 32-row chunk (tc_common.cuh): [row chunk][hi | lo][n-group][row] x 16 B.  Written for every
        // row of the tile (rows beyond the batch meet a zero A operand in the WEIGHT pass)
        uint4* h1img = (P.keep_h2 && P.h1) ? reinterpret_cast<uint4*>(P.h1) + (size_t)(row0 / 32 + (warp & 3)) * (2 * 32 * 32) : nullptr;
        unsigned long long h1bytes = 0ull;   // byte c: relu'(h1) of this thread's 8 hidden units of k-chunk c
        for (int c = 0; c < NCHUNK; ++c) {
            const int stage = c % NSTAGE;
            unsigned char* a_hi = S.stage[stage];
            unsigned char* a_lo = a_hi + A_IMG;
            float hs[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int k = c * KCH + q * 8 + e;
                const float4 wv = *reinterpret_cast<const float4*>(S.W1[k]);
                float h = fmaf(wv.x, x0, S.b1[k]);
                h = fmaf(wv.y, x1, h);
                if (four) {
                    h = fmaf(wv.z, x2, h);
                    h = fmaf(wv.w, x3, h);
                }
                hs[e] = fminf(fmaxf(h, 0.f), 60000.0f);    // h1 * SA (fp32 h1 itself is never stored)
            }
            if (P.h1bits) {
                uint32_t by = 0u;
#pragma unroll
                for (int e = 0; e < 8; ++e) by |= hs[e] > 0.f ? (1u << e) : 0u;
                h1bytes |= (unsigned long long)by << (8 * c);
            }
            uint4 hi, lo;
            split8(hs, &hi, &lo);
            if (h1img) {   // the same 16-byte pieces, row-major, for the WEIGHT pass of the backward (512 B contiguous per warp)
                const size_t u = (size_t)(c * 4 + q) * 32 + lane;
                h1img[u] = hi;
                h1img[u + 32 * 32] = lo;
            }
            mbar_wait(smem_u32(&S.empty[stage]), ((c / NSTAGE) & 1) ^ 1);   // after the arithmetic: the wait overlaps it
            *reinterpret_cast<uint4*>(a_hi + q * LBO_A + r * 16) = hi;
            *reinterpret_cast<uint4*>(a_lo + q * LBO_A + r * 16) = lo;
            fence_proxy_async();
            mbar_arrive_warp(smem_u32(&S.full[stage]));
        }
        if (P.h1bits && live) reinterpret_cast<unsigned long long*>(P.h1bits)[row * 4 + q] = h1bytes;
        // ---- epilogue: columns [64 q, 64 q + 64) of this row ----
        TSTAMP(2);
        mbar_wait(smem_u32(&S.acc_full), 0);
        TSTAMP(3);
        tc_fence_after();
        float out[4] = {0.f, 0.f, 0.f, 0.f};
        const int n_out = w.na + w.nb;
        // all MMAs have completed (acc_full): the operand stages are free and serve as per-warp transpose scratch
        float* scratch = reinterpret_cast<float*>(S.stage) + warp * kXposeFloats;
        const int64_t wrow0 = row0 + (warp & 3) * 32;     // first row of this warp's TMEM lane quadrant
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
            float v[32];
            const int col0 = q * 64 + cc * 32;
            tmem_ld32(lane_addr + col0, v);
            tmem_ld_wait();
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
                const int i4 = (col0 >> 2) + j4;
                const float4 b = reinterpret_cast<const float4*>(S.b2)[i4];
                const float h0 = fmaxf(fmaf(v[4 * j4 + 0], INV_SCALE, b.x), 0.f);
                const float h1 = fmaxf(fmaf(v[4 * j4 + 1], INV_SCALE, b.y), 0.f);
                const float h2 = fmaxf(fmaf(v[4 * j4 + 2], INV_SCALE, b.z), 0.f);
                const float h3 = fmaxf(fmaf(v[4 * j4 + 3], INV_SCALE, b.w), 0.f);
                v[4 * j4 + 0] = h0; v[4 * j4 + 1] = h1; v[4 * j4 + 2] = h2; v[4 * j4 + 3] = h3;
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    if (o < n_out) {
                        const float4 wv = reinterpret_cast<const float4*>(S.w3[o])[i4];
                        out[o] = fmaf(h3, wv.w, fmaf(h2, wv.z, fmaf(h1, wv.y, fmaf(h0, wv.x, out[o]))));
                    }
                }
            }
            if (P.h2bits && live) {   // relu'(h2) for the backward DATA producers: one word per 32 columns
                uint32_t mb = 0u;   // v >= +0 after the ReLU: non-zero bit pattern <=> h2 > 0
#pragma unroll
                for (int j = 31; j >= 0; --j) mb = (mb << 1) + min(__float_as_uint(v[j]), 1u);
                P.h2bits[row * (H / 32) + q * 2 + cc] = mb;
            }
            if (P.h2 && P.keep_h2)
                warp_block_rows(scratch, v, lane, [&](int rl, int c4, float4 hv) {
                    if (wrow0 + rl < rows) *reinterpret_cast<float4*>(P.h2 + (wrow0 + rl) * H + col0 + 4 * c4) = hv;
                });
        }
        tc_fence_before();
        S.part[q][r] = make_float4(out[0], out[1], out[2], out[3]);
        asm volatile("bar.sync 1, %0;" ::"n"(kProd) : "memory");
        if (q == 0 && live) {
            const float4 p0 = S.part[0][r], p1 = S.part[1][r], p2 = S.part[2][r], p3 = S.part[3][r];
            float raw[4];
            raw[0] = ((p0.x + p1.x) + (p2.x + p3.x)) + S.b3[0];
            raw[1] = ((p0.y + p1.y) + (p2.y + p3.y)) + S.b3[1];
            raw[2] = ((p0.z + p1.z) + (p2.z + p3.z)) + S.b3[2];
            raw[3] = ((p0.w + p1.w) + (p2.w + p3.w)) + S.b3[3];
            forward_tail(P, A, row, raw);
        }
    } else if (warp == kProd / 32) {
        if (lane == 0) {
            for (int c = 0; c < NCHUNK; ++c) {
                const int stage = c % NSTAGE;
                mbar_wait(smem_u32(&S.full[stage]), (c / NSTAGE) & 1);
                tc_fence_after();
                const uint32_t a_hi = smem_u32(S.stage[stage]);
                const uint32_t a_lo = a_hi + A_IMG, b_hi = a_hi + 2 * A_IMG, b_lo = b_hi + B_IMG;
#pragma unroll
                for (int j = 0; j < KCH / 16; ++j) {
                    const uint64_t dah = make_desc(a_hi + j * 2 * LBO_A, LBO_A, SBO);
                    const uint64_t dal = make_desc(a_lo + j * 2 * LBO_A, LBO_A, SBO);
                    const uint64_t dbh = make_desc(b_hi + j * 2 * LBO_B, LBO_B, SBO);
                    const uint64_t dbl = make_desc(b_lo + j * 2 * LBO_B, LBO_B, SBO);
                    umma_f16(tmem_base, dah, dbh, (c | j) ? 1u : 0u);
                    umma_f16(tmem_base, dah, dbl, 1u);
                    umma_f16(tmem_base, dal, dbh, 1u);
                }
                umma_commit(smem_u32(&S.empty[stage]));
            }
            umma_commit(smem_u32(&S.acc_full));
        }
    } else {
        if (lane == 0) {
            const unsigned char* img = reinterpret_cast<const unsigned char*>(P.tc_img);
            for (int c = 0; c < NCHUNK; ++c) {
                const int stage = c % NSTAGE;
                mbar_wait(smem_u32(&S.empty[stage]), ((c / NSTAGE) & 1) ^ 1);
                const uint32_t bar = smem_u32(&S.full[stage]);
                const uint32_t dst = smem_u32(S.stage[stage]) + 2 * A_IMG;
                mbar_arrive_expect_tx(bar, 2 * B_IMG);
                bulk_g2s(dst, img + (size_t)c * 2 * B_IMG, B_IMG, bar);
                bulk_g2s(dst + B_IMG, img + (size_t)c * 2 * B_IMG + B_IMG, B_IMG, bar);
            }
        }
    }
    TSTAMP(4);
    tc_fence_before();
    __syncthreads();
    if (warp == kProd / 32) tmem_dealloc(tmem_base, 256);
    TSTAMP(5);
    run_tail(A.tail);
    TSTAMP(6);
    TLAUNCH_END(1);
}


// ---- backward GEMMs of the updates on tcgen05 (head backward fused) -------------------------------------
// DATA   CTA (128-row tile): A = dh2 tile (rebuilt by the producers from h2 / dout / W3, scaled per ROW by a power
//        of two so that fp16 hi/lo covers it), B = image of W2^T streamed with cp.async.bulk, D -> dh1 (masked by h1).
// WEIGHT CTA (128 out-units m): A[m][k = row] = dh2[row][m] (scaled per COLUMN m), B[n][k = row] = h1[row][n] * 16,
//        both written transposed into the canonical K-major layout by the producers, D -> gW2; the A producers also
//        reduce gb2 / gW3 (and the m-tile 0 CTA gb3) deterministically.
constexpr int kDoutRows = 1024;
struct BwdTcSmem {
    unsigned char stage[NSTAGE][STAGE_BYTES];
    float W1[H][4];                 // layer 1 of the pass, pre-multiplied by SA like the forward kernel (h1 is recomputed)
    float b1[H];
    float4 xin[TM];                 // DATA: the inputs (s, a) of the tile's rows
    float W1z[H], W1w[H];           // DATA: the action columns of W1 (* SA), component-major (conflict-free epilogue reads)
    uint32_t h1b[TM][H / 32];       // DATA: sign bits of the tile's h1 (relu' mask of the epilogue; FwdPass::h1bits layout)
    float w3[4][H];
    float red[kProd][6];            // WEIGHT: per-thread partial sums (gb2, gW3[0..3])
    float4 douts[kDoutRows];        // WEIGHT: dout of the whole batch (rows <= kDoutRows), zero-padded to 4 outputs
    float wmax[4], dmax[4], gb3[4];
    float red4[kProd / 32][8];
    unsigned long long full[NSTAGE], empty[NSTAGE], acc_full;
    uint32_t tmem_base;
};

__device__ __forceinline__ float pow2_scale(float bound) {
    // largest power of two s with s * bound <= 8192 (fp16 max 65504; hi/lo split keeps ~22 bits below that)
    if (!(bound > 0.f) || !isfinite(bound)) return 1.0f;
    int e;
    frexpf(bound, &e);                 // bound = f * 2^e, f in [0.5, 1)
    e = 13 - e;
    e = e > 100 ? 100 : (e < -100 ? -100 : e);
    return ldexpf(1.0f, e);
}

__global__ void __launch_bounds__(kTcThreads, 1) bwd_tc_kernel(const __grid_constant__ GemmArgs G) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];   // see act_tc_kernel: no pointer arithmetic on the base
    pdl_wait();
    TSTAMP(0);
    BwdTcSmem& S = *reinterpret_cast<BwdTcSmem*>(smem_raw);
    const int64_t rows = *G.rows_ptr;
    const GemmPass& P = G.p[blockIdx.y];
    const bool weight = P.k_is_rows != 0;
    const int tile = blockIdx.x;
    if (rows <= 0 || (weight ? (tile >= H / TM) : ((int64_t)tile * TM >= rows))) {   // uniform exit
        run_tail(G.tail);
        return;
    }
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int n_chunks = weight ? (int)((rows + KCH - 1) / KCH) : NCHUNK;

    // ---- setup ----
    const bool four = P.n_in == 4;
    for (int k = t; k < H; k += kTcThreads) {
#pragma unroll
        for (int o = 0; o < 4; ++o)
            S.w3[o][k] = o < P.na ? P.W3a[o * H + k] : (o < P.n_out ? P.W3b[(o - P.na) * H + k] : 0.f);
        if (weight) continue;                                  // layer 1 is only recomputed by the DATA tiles
        float4 w1 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (four) {
            w1 = *reinterpret_cast<const float4*>(P.W1 + k * 4);
        } else {
            const float2 v = *reinterpret_cast<const float2*>(P.W1 + k * 2);
            w1.x = v.x; w1.y = v.y;
        }
        w1.x *= SA; w1.y *= SA; w1.z *= SA; w1.w *= SA;       // exact (power of two): same values as fwd_tc_kernel
        *reinterpret_cast<float4*>(S.W1[k]) = w1;
        S.W1z[k] = w1.z; S.W1w[k] = w1.w;
        S.b1[k] = P.b1[k] * SA;
    }
    if (!weight) {   // inputs (s, a) of the 128 rows of this DATA tile
        const int64_t xbase = (int64_t)tile * TM;
        const int nx = TM;
        for (int i = t; i < nx; i += kTcThreads) {
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            if (xbase + i < rows) {
                const float2 sv = reinterpret_cast<const float2*>(P.xs)[xbase + i];
                x.x = sv.x; x.y = sv.y;
                if (four) {
                    const float2 av = reinterpret_cast<const float2*>(P.xa)[xbase + i];
                    x.z = av.x; x.w = av.y;
                }
            }
            S.xin[i] = x;
        }
        for (int i = t; i < TM * (H / 32); i += kTcThreads)
            (&S.h1b[0][0])[i] = (xbase + i / (H / 32) < rows) ? P.h1bits[xbase * (H / 32) + i] : 0u;
    }
    if (t == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(smem_u32(&S.full[s]), kProdWarps + 1);   // producer warps + the loader's expect_tx arrival
            mbar_init(smem_u32(&S.empty[s]), 1);
        }
        mbar_init(smem_u32(&S.acc_full), 1);
        fence_barrier_init();
    }
    if (warp == kProd / 32) {
        tmem_alloc(smem_u32(&S.tmem_base), 256);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = S.tmem_base;
    const bool dout_smem = weight && rows <= kDoutRows;
    if (dout_smem)
        for (int r = t; r < (int)rows; r += kTcThreads) {
            float4 d4 = make_float4(0.f, 0.f, 0.f, 0.f);
            d4.x = P.dout[(size_t)r * P.stride];
            if (P.n_out > 1) d4.y = P.dout[(size_t)r * P.stride + 1];
            if (P.n_out > 2) d4.z = P.dout[(size_t)r * P.stride + 2];
            if (P.n_out > 3) d4.w = P.dout[(size_t)r * P.stride + 3];
            S.douts[r] = d4;
        }
    // column maxima of |W3| (DATA: per-row bound) / maxima and sums of dout over the batch (WEIGHT: per-column bound, gb3)
    if (warp < 4) {
        const int o = warp;
        float m = 0.f, dm = 0.f, ds = 0.f;
        if (o < P.n_out) {
            for (int k = lane; k < H; k += 32) m = fmaxf(m, fabsf(S.w3[o][k]));
            if (weight)
                for (int64_t r = lane; r < rows; r += 32) {
                    const float d = P.dout[r * P.stride + o];
                    dm = fmaxf(dm, fabsf(d));
                    ds += d;
                }
        }
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) {
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, sft));
            dm = fmaxf(dm, __shfl_xor_sync(0xffffffffu, dm, sft));
            ds += __shfl_xor_sync(0xffffffffu, ds, sft);
        }
        if (lane == 0) { S.wmax[o] = m; S.dmax[o] = dm; S.gb3[o] = ds; }
    }
    __syncthreads();

    TSTAMP(1);
    if (warp < kProd / 32) {
        const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        if (!weight) {
            // ================= DATA: 4 threads per row (q = quarter of the 32-wide k chunk) =================
            const int q = warp >> 2, r = (warp & 3) * 32 + lane;
            const int64_t row = (int64_t)tile * TM + r;
            const bool live = row < rows;
            float d[4] = {0.f, 0.f, 0.f, 0.f};
            float bound = 0.f;
            if (live)
#pragma unroll
                for (int o = 0; o < 4; ++o)
                    if (o < P.n_out) {
                        d[o] = P.dout[row * P.stride + o];
                        bound = fmaf(fabsf(d[o]), S.wmax[o], bound);
                    }
            const float sc = pow2_scale(bound);
            // relu'(h2) of the row: 256 sign bits written by the forward kernel (no activation loads in this loop)
            unsigned long long hbytes;   // byte c = the sign bits of this thread's 8 columns of k-chunk c
            {
                uint4 b0 = make_uint4(0u, 0u, 0u, 0u), b1v = b0;
                if (live) {
                    b0 = *reinterpret_cast<const uint4*>(P.h2bits + row * (H / 32));
                    b1v = *reinterpret_cast<const uint4*>(P.h2bits + row * (H / 32) + 4);
                }
                const int sh = q * 8;
                const uint32_t lo4 = ((b0.x >> sh) & 0xffu) | (((b0.y >> sh) & 0xffu) << 8) | (((b0.z >> sh) & 0xffu) << 16) |
                                     (((b0.w >> sh) & 0xffu) << 24);
                const uint32_t hi4 = ((b1v.x >> sh) & 0xffu) | (((b1v.y >> sh) & 0xffu) << 8) | (((b1v.z >> sh) & 0xffu) << 16) |
                                     (((b1v.w >> sh) & 0xffu) << 24);
                hbytes = ((unsigned long long)hi4 << 32) | lo4;
            }
#pragma unroll 1
            for (int c = 0; c < NCHUNK; ++c) {
                const int stage = c % NSTAGE;
                const int k0 = c * KCH + q * 8;
                const uint32_t bits = (uint32_t)(hbytes >> (8 * c));
                float gv[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    float v = d[0] * S.w3[0][k0 + e];
                    v = fmaf(d[1], S.w3[1][k0 + e], v); v = fmaf(d[2], S.w3[2][k0 + e], v); v = fmaf(d[3], S.w3[3][k0 + e], v);
                    gv[e] = ((bits >> e) & 1u) ? v * sc : 0.f;
                }
                uint4 hi, lo;
                split8(gv, &hi, &lo);
                mbar_wait(smem_u32(&S.empty[stage]), ((c / NSTAGE) & 1) ^ 1);
                unsigned char* a_hi = S.stage[stage];
                *reinterpret_cast<uint4*>(a_hi + q * LBO_A + r * 16) = hi;
                *reinterpret_cast<uint4*>(a_hi + A_IMG + q * LBO_A + r * 16) = lo;
                fence_proxy_async();
                mbar_arrive_warp(smem_u32(&S.full[stage]));
            }
            TSTAMP(2);
            mbar_wait(smem_u32(&S.acc_full), 0);
            TSTAMP(3);
            tc_fence_after();
            const float inv = 1.0f / (sc * SB);
            float* scratch = reinterpret_cast<float*>(S.stage) + warp * kXposeFloats;   // stages are free after acc_full
            // fused layer-1 backward (layer1_backward_kernel's arithmetic on the tile, dh1 never leaves the SM):
            //   colpart [4 row quadrants][H][5]  per-quadrant column sums  sum_r dh1[r][k] * {x0, x1, x2, x3, 1}
            //   dxapart [4 column quarters][TM]  per-quarter row sums      sum_k dh1[r][k] * W1[k][2 | 3]
            float* colpart = reinterpret_cast<float*>(S.stage) + kProdWarps * kXposeFloats;
            float2* dxapart = reinterpret_cast<float2*>(colpart + 4 * H * 5);
            const bool want_gw = P.gW1 != nullptr, want_dxa = P.dxa != nullptr;
            float acc[20];   // want_dxa: [i][2] over the thread's 8 rows; want_gw: [column j][5], restarted per 32-column block
#pragma unroll
            for (int u = 0; u < 20; ++u) acc[u] = 0.f;
            const int64_t wrow0 = (int64_t)tile * TM + (warp & 3) * 32;
#pragma unroll 1
            for (int cc = 0; cc < 2; ++cc) {
                float v[32];
                const int col0 = q * 64 + cc * 32;
                tmem_ld32(lane_addr + col0, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] *= inv;      // this row's scale
                warp_block_rows_i(scratch, v, lane, [&](int i, int rl, int c4, float4 g) {
                    const int64_t rr = wrow0 + rl;
                    const bool ok = rr < rows;
                    const int lr = (warp & 3) * 32 + rl;
                    // relu'(h1) of the 4 columns: byte (k / 8 % 4) * 8 + k / 32 of the row, bits k % 8 .. + 3
                    const uint32_t hbits = ok ? (uint32_t)(reinterpret_cast<const unsigned char*>(S.h1b[lr])[(c4 >> 1) * 8 + (col0 >> 5)] >> ((c4 & 1) * 4)) : 0u;
                    const float4 x = S.xin[lr];
                    float gg[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int k = col0 + 4 * c4 + j;
                        gg[j] = ((hbits >> j) & 1u) ? gg[j] : 0.f;
                        if (want_dxa) {   // W1 is staged pre-multiplied by SA: undone once per row below
                            acc[2 * i] = fmaf(gg[j], S.W1z[k], acc[2 * i]);
                            acc[2 * i + 1] = fmaf(gg[j], S.W1w[k], acc[2 * i + 1]);
                        }
                        if (want_gw) {
                            acc[5 * j + 0] = fmaf(gg[j], x.x, acc[5 * j + 0]);
                            acc[5 * j + 1] = fmaf(gg[j], x.y, acc[5 * j + 1]);
                            acc[5 * j + 2] = fmaf(gg[j], x.z, acc[5 * j + 2]);
                            acc[5 * j + 3] = fmaf(gg[j], x.w, acc[5 * j + 3]);
                            acc[5 * j + 4] += gg[j];
                        }
                    }
                    if (P.C && ok) *reinterpret_cast<float4*>(P.C + rr * H + col0 + 4 * c4) = make_float4(gg[0], gg[1], gg[2], gg[3]);
                });
                if (want_gw) {   // the four row groups of the warp (lane >> 3), then one writer per column
#pragma unroll
                    for (int u = 0; u < 20; ++u) {
                        float a = acc[u];
                        a += __shfl_xor_sync(0xffffffffu, a, 8);
                        a += __shfl_xor_sync(0xffffffffu, a, 16);
                        if (lane < 8) colpart[((warp & 3) * H + col0 + 4 * lane + u / 5) * 5 + u % 5] = a;
                        acc[u] = 0.f;
                    }
                }
            }
            if (want_dxa) {      // the eight column groups of the warp (lane & 7), then one writer per row
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    float a = acc[u];
                    a += __shfl_xor_sync(0xffffffffu, a, 1);
                    a += __shfl_xor_sync(0xffffffffu, a, 2);
                    a += __shfl_xor_sync(0xffffffffu, a, 4);
                    acc[u] = a;
                }
                if ((lane & 7) == 0) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) dxapart[q * TM + (warp & 3) * 32 + i * 4 + (lane >> 3)] = make_float2(acc[2 * i], acc[2 * i + 1]);
                }
            }
            tc_fence_before();
            if (want_gw || want_dxa) {
                asm volatile("bar.sync 1, %0;" ::"n"(kProd) : "memory");
                if (want_dxa && t < TM) {
                    const int64_t rr = (int64_t)tile * TM + t;
                    const float2 p0 = dxapart[t], p1 = dxapart[TM + t], p2 = dxapart[2 * TM + t], p3 = dxapart[3 * TM + t];
                    if (rr < rows)
                        reinterpret_cast<float2*>(P.dxa)[rr] = make_float2(((p0.x + p1.x) + (p2.x + p3.x)) * (1.0f / SA),
                                                                           ((p0.y + p1.y) + (p2.y + p3.y)) * (1.0f / SA));
                }
                if (want_gw) {
                    const int n_tiles = (int)((rows + TM - 1) / TM);
                    float s5[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
                    if (t < H) {
#pragma unroll
                        for (int u = 0; u < 5; ++u)
                            s5[u] = (colpart[(0 * H + t) * 5 + u] + colpart[(1 * H + t) * 5 + u]) +
                                    (colpart[(2 * H + t) * 5 + u] + colpart[(3 * H + t) * 5 + u]);
                    }
                    bool writer = n_tiles == 1;
                    if (n_tiles > 1) {   // partial sums of this tile -> global; the last tile to arrive adds them up in tile order
                        if (t < H) {
#pragma unroll
                            for (int u = 0; u < 5; ++u) P.l1part[((size_t)tile * H + t) * 5 + u] = s5[u];
                        }
                        __threadfence();
                        asm volatile("bar.sync 1, %0;" ::"n"(kProd) : "memory");
                        if (t == 0) {
                            const int arrived = atomicAdd(P.l1ticket, 1);
                            S.red4[0][0] = (arrived == n_tiles - 1) ? 1.f : 0.f;
                            if (arrived == n_tiles - 1) { *reinterpret_cast<volatile int*>(P.l1ticket) = 0; __threadfence(); }
                        }
                        asm volatile("bar.sync 1, %0;" ::"n"(kProd) : "memory");
                        writer = S.red4[0][0] != 0.f;
                        if (writer && t < H) {
#pragma unroll
                            for (int u = 0; u < 5; ++u) s5[u] = 0.f;
                            for (int tl = 0; tl < n_tiles; ++tl)
#pragma unroll
                                for (int u = 0; u < 5; ++u) s5[u] += __ldcg(P.l1part + ((size_t)tl * H + t) * 5 + u);
                        }
                    }
                    if (writer && t < H) {
                        for (int u = 0; u < P.n_in; ++u) P.gW1[t * P.n_in + u] = s5[u];
                        P.gb1[t] = s5[4];
                    }
                }
            }
        } else {
            // ================= WEIGHT: thread (m, kc) stages A; B arrives by bulk copy =================
            const int m = t & (TM - 1), kc = t >> 7;             // A role: out unit m0 + m, core column kc (8 rows)
            const int mcol = tile * TM + m;
            float bound = 0.f;
#pragma unroll
            for (int o = 0; o < 4; ++o) bound = fmaf(S.dmax[o], fabsf(S.w3[o][mcol]), bound);
            const float sc = pow2_scale(bound);
            float w3m[4];
#pragma unroll
            for (int o = 0; o < 4; ++o) w3m[o] = S.w3[o][mcol];
            float gb2 = 0.f, gw[4] = {0.f, 0.f, 0.f, 0.f};
            // B operand = h1 (scaled by SA): the fp16 hi / lo image the forward kernel wrote, row-major == MN-major for this
            // contraction (K = batch rows): the loader thread bulk-copies one 32-row chunk per stage, no transpose, no
            // producer work
            float h2n[8];                                                // prefetched h2 column values of the NEXT chunk
            auto prefetch = [&](int c) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int64_t ra = (int64_t)c * KCH + kc * 8 + e;
                    h2n[e] = ra < rows ? P.h2[ra * H + mcol] : 0.f;
                }
            };
            prefetch(0);
            for (int c = 0; c < n_chunks; ++c) {
                const int stage = c % NSTAGE;
                const int64_t r0 = (int64_t)c * KCH + kc * 8;
                float h2c[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) h2c[e] = h2n[e];
                if (c + 1 < n_chunks) prefetch(c + 1);                   // in flight while this chunk is converted
                uint4 ahi, alo;
                float av[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int64_t r = r0 + e;
                    float v = 0.f;
                    if (r < rows) {
                        const float h = h2c[e];
                        float dd[4];
                        if (dout_smem) {
                            const float4 d4 = S.douts[r];
                            dd[0] = d4.x; dd[1] = d4.y; dd[2] = d4.z; dd[3] = d4.w;
                        } else {
#pragma unroll
                            for (int o = 0; o < 4; ++o) dd[o] = o < P.n_out ? P.dout[r * P.stride + o] : 0.f;
                        }
                        float g = dd[0] * w3m[0];
                        g = fmaf(dd[1], w3m[1], g); g = fmaf(dd[2], w3m[2], g); g = fmaf(dd[3], w3m[3], g);
                        g = h > 0.f ? g : 0.f;
                        gb2 += g;
#pragma unroll
                        for (int o = 0; o < 4; ++o) gw[o] = fmaf(dd[o], h, gw[o]);
                        v = g * sc;
                    }
                    av[e] = v;
                }
                split8(av, &ahi, &alo);
                mbar_wait(smem_u32(&S.empty[stage]), ((c / NSTAGE) & 1) ^ 1);
                unsigned char* a_hi = S.stage[stage];
                *reinterpret_cast<uint4*>(a_hi + kc * LBO_A + m * 16) = ahi;
                *reinterpret_cast<uint4*>(a_hi + A_IMG + kc * LBO_A + m * 16) = alo;
                fence_proxy_async();
                mbar_arrive_warp(smem_u32(&S.full[stage]));
            }
            // head-layer gradients: reduce the four kc partials of each out unit
            if (P.gb2) {
                S.red[t][0] = gb2;
#pragma unroll
                for (int o = 0; o < 4; ++o) S.red[t][1 + o] = gw[o];
            }
            // epilogue: D[m][64 q .. 64 q + 63] -> gW2   (TMEM lane = m: warp & 3 selects the lane quadrant)
            const int q = warp >> 2, mr = (warp & 3) * 32 + lane;
            const int mrow = tile * TM + mr;
            float bound2 = 0.f;
#pragma unroll
            for (int o = 0; o < 4; ++o) bound2 = fmaf(S.dmax[o], fabsf(S.w3[o][mrow]), bound2);
            const float inv = 1.0f / (pow2_scale(bound2) * SA);
            TSTAMP(2);
            mbar_wait(smem_u32(&S.acc_full), 0);
            TSTAMP(3);
            tc_fence_after();
            float* scratch = reinterpret_cast<float*>(S.stage) + warp * kXposeFloats;   // stages are free after acc_full
            const int wm0 = tile * TM + (warp & 3) * 32;
#pragma unroll 1
            for (int cc = 0; cc < 2; ++cc) {
                float v[32];
                const int col0 = q * 64 + cc * 32;
                tmem_ld32(lane_addr + col0, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] *= inv;      // this out unit's scale
                warp_block_rows(scratch, v, lane, [&](int rl, int c4, float4 g) {
                    *reinterpret_cast<float4*>(P.C + (size_t)(wm0 + rl) * H + col0 + 4 * c4) = g;
                });
            }
            tc_fence_before();
            asm volatile("bar.sync 1, %0;" ::"n"(kProd) : "memory");
            if (P.gb2 && t < TM) {
                float s5[5];
#pragma unroll
                for (int j = 0; j < 5; ++j) s5[j] = (S.red[t][j] + S.red[t + TM][j]) + (S.red[t + 2 * TM][j] + S.red[t + 3 * TM][j]);
                const int col = tile * TM + t;
                P.gb2[col] = s5[0];
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    if (o < P.na) P.gW3a[o * H + col] = s5[1 + o];
                    else if (o < P.n_out) P.gW3b[(o - P.na) * H + col] = s5[1 + o];
                }
                if (tile == 0 && t < P.n_out) {
                    if (t < P.na) P.gb3a[t] = S.gb3[t];
                    else P.gb3b[t - P.na] = S.gb3[t];
                }
            }
        }
    } else if (warp == kProd / 32) {
        if (lane == 0) {
            // DATA: B = W2^T image, K-major.  WEIGHT: B = h1 image, MN-major (16 k-rows = two 128-byte k-groups per MMA)
            const uint32_t lbo_b = weight ? LBO_BMN : LBO_B, sbo_b = weight ? SBO_BMN : SBO, kstep_b = weight ? 2 * LBO_BMN : 2 * LBO_B;
            const uint32_t idesc = weight ? IDESC_BMN : IDESC;
            for (int c = 0; c < n_chunks; ++c) {
                const int stage = c % NSTAGE;
                mbar_wait(smem_u32(&S.full[stage]), (c / NSTAGE) & 1);
                tc_fence_after();
                const uint32_t a_hi = smem_u32(S.stage[stage]);
                const uint32_t a_lo = a_hi + A_IMG, b_hi = a_hi + 2 * A_IMG, b_lo = b_hi + B_IMG;
#pragma unroll
                for (int j = 0; j < KCH / 16; ++j) {
                    const uint64_t dah = make_desc(a_hi + j * 2 * LBO_A, LBO_A, SBO);
                    const uint64_t dal = make_desc(a_lo + j * 2 * LBO_A, LBO_A, SBO);
                    const uint64_t dbh = make_desc(b_hi + j * kstep_b, lbo_b, sbo_b);
                    const uint64_t dbl = make_desc(b_lo + j * kstep_b, lbo_b, sbo_b);
                    umma_f16_idesc(tmem_base, dah, dbh, (c | j) ? 1u : 0u, idesc);
                    umma_f16_idesc(tmem_base, dah, dbl, 1u, idesc);
                    umma_f16_idesc(tmem_base, dal, dbh, 1u, idesc);
                }
                umma_commit(smem_u32(&S.empty[stage]));
            }
            umma_commit(smem_u32(&S.acc_full));
        }
    } else {
        if (lane == 0) {
            // DATA: the W2^T image, chunk c of 8.  WEIGHT: the h1 image, row chunk c (32 KB: hi then lo, the stage's B layout)
            const unsigned char* img = reinterpret_cast<const unsigned char*>(weight ? static_cast<const void*>(P.B) : static_cast<const void*>(P.tc_imgT));
            for (int c = 0; c < n_chunks; ++c) {
                const int stage = c % NSTAGE;
                mbar_wait(smem_u32(&S.empty[stage]), ((c / NSTAGE) & 1) ^ 1);
                const uint32_t bar = smem_u32(&S.full[stage]);
                const uint32_t dst = smem_u32(S.stage[stage]) + 2 * A_IMG;
                mbar_arrive_expect_tx(bar, 2 * B_IMG);
                bulk_g2s(dst, img + (size_t)c * 2 * B_IMG, B_IMG, bar);
                bulk_g2s(dst + B_IMG, img + (size_t)c * 2 * B_IMG + B_IMG, B_IMG, bar);
            }
        }
    }
    TSTAMP(4);
    tc_fence_before();
    __syncthreads();
    if (warp == kProd / 32) tmem_dealloc(tmem_base, 256);
    TSTAMP(5);
    run_tail(G.tail);
    TLAUNCH_END(2);
}

}  // namespace

namespace rrl {

int fwd_tc_launch(const FwdArgs& A, int64_t max_rows, cudaStream_t st) {
    static bool configured = false;
    const size_t smem = sizeof(FwdTcSmem) + 128;
    if (!configured) {
        RRL_CUDA(cudaFuncSetAttribute(fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    for (int i = 0; i < A.n_pass; ++i)
        if (!A.p[i].tc_img) { rrl_set_error("fwd_tc_launch: pass %d has no tcgen05 weight image", i); return -2; }
    dim3 grid((unsigned)((max_rows + TM - 1) / TM), (unsigned)A.n_pass);
    RRL_CUDA(rrl_launch_pdl(fwd_tc_kernel, grid, dim3(kTcThreads), smem, st, A));
    return 0;
}

#ifdef RRL_TC_TIMING
// out_log: 256*16*8 uint64, out_kind: 256*4 int32, returns the number of launches logged so far (negative: CUDA error)
extern "C" int rrl_debug_tc_times(unsigned long long* out_log, int* out_kind) {
    unsigned int n = 0;
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    if (cudaMemcpyFromSymbol(out_log, g_tc_log, sizeof(unsigned long long) * 256 * 16 * 8) != cudaSuccess) return -1;
    if (cudaMemcpyFromSymbol(out_kind, g_tc_kind, sizeof(int) * 256 * 4) != cudaSuccess) return -1;
    if (cudaMemcpyFromSymbol(&n, g_tc_launch, sizeof(n)) != cudaSuccess) return -1;
    return (int)n;
}
#endif
int bwd_tc_launch(const GemmArgs& G, int64_t max_rows, cudaStream_t st) {
    static bool configured = false;
    const size_t smem = sizeof(BwdTcSmem) + 128;
    if (!configured) {
        RRL_CUDA(cudaFuncSetAttribute(bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    for (int i = 0; i < G.n_pass; ++i)
        if (!G.p[i].k_is_rows && !G.p[i].tc_imgT) { rrl_set_error("bwd_tc_launch: pass %d has no W2^T image", i); return -2; }
    int64_t tiles = (max_rows + TM - 1) / TM;
    if (tiles < H / TM) tiles = H / TM;
    RRL_CUDA(rrl_launch_pdl(bwd_tc_kernel, dim3((unsigned)tiles, (unsigned)G.n_pass), dim3(kTcThreads), smem, st, G));
    return 0;
}

int tc_images_launch(float* arena, const Layout& L, cudaStream_t st) {
    TcImgArgs A, T;
    static const int nets[6] = {RRL_NET_CRITIC, RRL_NET_CRITIC_TARGET, RRL_NET_POLICY, RRL_NET_QRISK, RRL_NET_QRISK_TARGET,
                                RRL_NET_RECOVERY};
    for (int ni = 0; ni < 6; ++ni) {
        const int net = nets[ni];
        const int heads = (net == RRL_NET_POLICY || net == RRL_NET_RECOVERY) ? 1 : 2;
        for (int h = 0; h < heads; ++h) {
            const int i = image_index(net, h);
            A.W2[i] = arena + L.t_off[net][w2_tensor(net, h)];
            A.img[i] = reinterpret_cast<__half*>(arena + L.tc_img_off[i]);
            T.W2[i] = arena + L.img_off[i];          // W2^T row-major == the k-major image (must be fresh)
            T.img[i] = reinterpret_cast<__half*>(arena + L.tc_imgT_off[i]);
        }
    }
    tc_images_kernel<<<dim3(H * 32 / 256, kTcHeads), 256, 0, st>>>(A);
    RRL_CHECK_LAUNCH();
    tc_images_kernel<<<dim3(H * 32 / 256, kTcHeads), 256, 0, st>>>(T);
    RRL_CHECK_LAUNCH();
    return 0;
}

int act_tc_launch(const ActArgs& A, const float* arena, const Layout& L, int stages, int max_ctas, cudaStream_t st) {
    TcActArgs T;
    T.a = A;
    T.img[PASS_POL] = tc_img_of(L, arena, RRL_NET_POLICY, 0);   // pass order: POL, REC, QR1, QR2
    T.img[PASS_REC] = tc_img_of(L, arena, RRL_NET_RECOVERY, 0);
    T.img[PASS_QR1] = tc_img_of(L, arena, RRL_NET_QRISK, 0);
    T.img[PASS_QR2] = tc_img_of(L, arena, RRL_NET_QRISK, 1);
    T.stages = stages;
    // one stage, or all of them: the kernel's schedule (epilogue two items behind production) relies on the recovery pass
    // sitting between the policy pass and the Q_risk passes that consume its action
    if (stages != RRL_ACT_STAGE_POLICY && stages != RRL_ACT_STAGE_QRISK && stages != RRL_ACT_STAGE_RECOVERY && stages != RRL_ACT_STAGE_ALL) {
        rrl_set_error("act_tc_launch: stages must be one RRL_ACT_STAGE_* or RRL_ACT_STAGE_ALL (got %d)", stages);
        return -2;
    }
    if (!A.use_recovery && !(stages & RRL_ACT_STAGE_POLICY)) return 0;   // nothing to do: no Q_risk / recovery stages without a recovery policy
    if ((stages & (RRL_ACT_STAGE_QRISK | RRL_ACT_STAGE_RECOVERY)) != (RRL_ACT_STAGE_QRISK | RRL_ACT_STAGE_RECOVERY) && A.use_recovery &&
        !A.recovery) { rrl_set_error("act_tc_launch: staged acting needs the recovery flag array"); return -2; }
    static bool configured = false;
    const size_t smem = sizeof(TcSmem) + 128;
    if (!configured) {
        RRL_CUDA(cudaFuncSetAttribute(act_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const int64_t tiles = (A.n + TM - 1) / TM;
    int64_t sms = rrl_num_sms();
    if (max_ctas > 0 && max_ctas < sms) sms = max_ctas;
    const int grid = (int)(tiles < sms ? tiles : sms);
    // a bounded (side-stream) stage is NOT made resident early: its one-CTA-per-SM footprint, blocked in pdl_wait(), would keep
    // the SMs from the update kernels it is meant to run next to
    RRL_CUDA(rrl_launch_pdl_if(max_ctas <= 0, act_tc_kernel, dim3(grid), dim3(kTcThreads), smem, st, T));
    return 0;
}

}  // namespace rrl
