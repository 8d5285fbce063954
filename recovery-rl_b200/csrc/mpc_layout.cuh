// mpc_layout.cuh -- layout of the packed ensemble ("dyn image") shared by the SIMT (mpc.cu) and tcgen05 (mpc_tc.cu)
// planner kernels.
#pragma once
#include "common.cuh"
#include "agent_layout.cuh"

namespace rrl {
namespace dyn {

constexpr int PBM = 64;   // rows per tile
constexpr int NETS = 5;   // config/default.py:91
constexpr int DYN_IN = 4, DYN_OUT = 4;

// ---- dyn image offsets (floats) ----
constexpr int64_t kW0 = 0;                                  // [5][4][256]   k-major layer 0
constexpr int64_t kB0 = kW0 + (int64_t)NETS * DYN_IN * H;   // [5][256]
constexpr int64_t kW1 = kB0 + (int64_t)NETS * H;            // [5][256][256] k-major
constexpr int64_t kB1 = kW1 + (int64_t)NETS * H * H;
constexpr int64_t kW2 = kB1 + (int64_t)NETS * H;
constexpr int64_t kB2 = kW2 + (int64_t)NETS * H * H;
constexpr int64_t kW3 = kB2 + (int64_t)NETS * H;            // [5][4][256]   w3[o][k] = lin3_w[k][o]
constexpr int64_t kB3 = kW3 + (int64_t)NETS * DYN_OUT * H;  // [5][4]
constexpr int64_t kMu = kB3 + NETS * DYN_OUT;               // [4]
constexpr int64_t kSigma = kMu + 4;                         // [4]
constexpr int64_t kMaxLv = kSigma + 4;                      // [2]
constexpr int64_t kMinLv = kMaxLv + 2;                      // [2]
constexpr int64_t kDynSimtFloats = ((kMinLv + 2 + 3) / 4) * 4;
// tcgen05 operand images (fp16 hi/lo, UMMA canonical K-major; 65536 floats each): layer l (0: lin1, 1: lin2) of net e
constexpr int64_t kTcImg = kDynSimtFloats;
constexpr int64_t kDynFloats = kTcImg + 2ll * NETS * H * H;
inline __host__ __device__ int64_t tc_img_off(int layer, int net) { return kTcImg + (int64_t)(layer * NETS + net) * H * H; }


}  // namespace dyn
}  // namespace rrl
