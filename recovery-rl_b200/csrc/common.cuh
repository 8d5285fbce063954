// common.cuh -- shared helpers for librrl.so (sm_100a only)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/rrl.h"

#ifndef __CUDA_ARCH__
#define RRL_HOST_ONLY 1
#endif

void rrl_set_error(const char* fmt, ...);

#define RRL_CHECK_ARG(cond, msg)                                   \
    do {                                                           \
        if (!(cond)) {                                             \
            rrl_set_error("%s: %s", __func__, msg);                \
            return -2;                                             \
        }                                                          \
    } while (0)

#define RRL_CHECK_LAUNCH()                                                              \
    do {                                                                                \
        cudaError_t e_ = cudaPeekAtLastError();                                         \
        if (e_ != cudaSuccess) {                                                        \
            rrl_set_error("%s: CUDA launch failed: %s", __func__, cudaGetErrorString(e_)); \
            (void)cudaGetLastError();                                                   \
            return -1;                                                                  \
        }                                                                               \
    } while (0)

#define RRL_CUDA(call)                                                                  \
    do {                                                                                \
        cudaError_t e_ = (call);                                                        \
        if (e_ != cudaSuccess) {                                                        \
            rrl_set_error("%s: %s failed: %s", __func__, #call, cudaGetErrorString(e_)); \
            return -1;                                                                  \
        }                                                                               \
    } while (0)

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch: the kernels of the vector step are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so that the NEXT kernel of the stream is scheduled (its CTAs resident,
// parameters loaded) while the current one still runs; it blocks in pdl_wait() -- the first thing every such kernel does,
// on every path, before it touches global memory -- until the previous kernel has completed and flushed.  A kernel releases
// its successor right after its own wait, so at most one future kernel is resident at a time.  OFF by default: on B200 the
// captured step got SLOWER with it (0.366 vs 0.339 ms, profiles/r2/pdl_ab.txt); RRL_PDL=1 in the environment (or
// rrl_set_pdl(1)) turns it on.  Without the launch attribute the waits are no-ops and the kernels are stream-ordered.
// ---------------------------------------------------------------------------------------------
int rrl_pdl_enabled();
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
template <typename... KArgs, typename... Args>
static inline cudaError_t rrl_launch_pdl_if(bool allow, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                           Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = (allow && rrl_pdl_enabled()) ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
static inline cudaError_t rrl_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    return rrl_launch_pdl_if(true, kernel, grid, block, smem, st, static_cast<Args&&>(args)...);
}
#endif

static inline int rrl_num_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
            sms = 148;  // B200
    }
    return sms;
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (production RNG mode).  key = (seed_lo ^ stream tag, seed_hi ^ stream_id),
// counter = (index_lo, index_hi, vec_step_lo, draw_id | vec_step_hi<<8).
// ---------------------------------------------------------------------------------------------
enum {
    RRL_DRAW_ENV_NOISE = 1,
    RRL_DRAW_ENV_RESET = 2,
    RRL_DRAW_ACT_TASK = 3,
    RRL_DRAW_ACT_REC = 4,
    RRL_DRAW_ACT_RAND = 5,
    RRL_DRAW_SAC_NEXT = 6,
    RRL_DRAW_SAC_CUR = 7,
    RRL_DRAW_QR_NEXT = 8,
    RRL_DRAW_QR_REC = 9,
    RRL_DRAW_INIT_RESET = 10,
    RRL_DRAW_MPC_EPS = 11,  // + 16 * CEM iteration
    RRL_DRAW_MPC_Z = 12,    // + 16 * CEM iteration
    RRL_DRAW_SQRL_EPS = 13, // SQRL action filter: the candidates' Gaussian noise (index = env * samples + j)
    RRL_DRAW_SQRL_CAT = 14, // SQRL action filter: the Categorical draw (index = env)
    RRL_DRAW_QSAMPLE = 15   // Q-sampling recovery: the uniform candidate actions (index = env * samples + j)
};

struct Philox4 {
    uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ uint32_t rrl_mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

__host__ __device__ __forceinline__ Philox4 rrl_philox(uint64_t seed, uint32_t stream_id, uint64_t index,
                                                       uint64_t vec_step, uint32_t draw_id) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32) ^ (stream_id * 0x9E3779B9u);
    uint32_t c0 = (uint32_t)index, c1 = (uint32_t)(index >> 32), c2 = (uint32_t)vec_step,
             c3 = draw_id | ((uint32_t)(vec_step >> 32) << 8);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = rrl_mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = rrl_mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    Philox4 o = {c0, c1, c2, c3};
    return o;
}

// 53-bit uniform in [0,1) from two words (same construction as numpy's rk_double)
__host__ __device__ __forceinline__ double rrl_u53(uint32_t a, uint32_t b) {
    return (double)(((uint64_t)(a >> 5) << 26) | (uint64_t)(b >> 6)) * (1.0 / 9007199254740992.0);
}
// 24-bit uniform in [0,1)
__host__ __device__ __forceinline__ float rrl_u24(uint32_t a) { return (float)(a >> 8) * (1.0f / 16777216.0f); }

#ifdef __CUDACC__
// two fp64 standard normals (Box-Muller) from one Philox block
__device__ __forceinline__ void rrl_normal2_f64(const Philox4& p, double* n0, double* n1) {
    double u1 = 1.0 - rrl_u53(p.x, p.y);  // (0,1]
    double u2 = rrl_u53(p.z, p.w);
    double r = sqrt(-2.0 * log(u1));
    double s, c;
    sincospi(2.0 * u2, &s, &c);
    *n0 = r * c;
    *n1 = r * s;
}
// two fp32 standard normals
__device__ __forceinline__ void rrl_normal2_f32(uint32_t a, uint32_t b, float* n0, float* n1) {
    float u1 = 1.0f - rrl_u24(a);  // (0,1]
    float u2 = rrl_u24(b);
    float r = sqrtf(-2.0f * logf(u1));
    float s, c;
    sincospif(2.0f * u2, &s, &c);
    *n0 = r * c;
    *n1 = r * s;
}
#endif
