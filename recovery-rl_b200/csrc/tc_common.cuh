// tc_common.cuh -- the tcgen05 / TMEM / mbarrier / bulk-copy building blocks shared by the tensor-core kernels
// (agent_tc.cu: acting + update GEMMs; mpc_tc.cu: the model-based planner): tile geometry, PTX wrappers, the
// shared-memory matrix descriptor, the fp16 hi/lo operand split and the per-warp transpose used by the epilogues.
#pragma once
#include "agent_common.cuh"
#include <cuda_fp16.h>

namespace rrl {
namespace tc {

constexpr int TM = 128;                  // rows per tile == TMEM lanes
constexpr int KCH = 32;                  // k per stage
constexpr int NSTAGE = 3;
constexpr int NCHUNK = H / KCH;          // 8
constexpr int A_IMG = TM * KCH * 2;      // 8 KB   (one fp16 image of the A chunk)
constexpr int B_IMG = H * KCH * 2;       // 16 KB
constexpr int STAGE_BYTES = 2 * A_IMG + 2 * B_IMG;  // 48 KB: [A hi][A lo][B hi][B lo]
constexpr int kProd = 512;                // producer/epilogue threads: 4 per row (16 warps)
constexpr int kTcThreads = kProd + 64;    // + MMA warp + loader warp
constexpr float SA = 16.0f, SB = kTcScaleB;  // power-of-two operand scales (SB: see tc_image_store)
constexpr float INV_SCALE = 1.0f / (SA * SB);
// canonical K-major, no swizzle: core matrix = 8 rows x 16 B (128 B contiguous)
//   A chunk [128 x 32]: core (kc, g) at (kc * 16 + g) * 128  -> LBO (next core along K) = 2048, SBO (next 8 rows) = 128
//   B chunk [256 x 32]: core (kc, g) at (kc * 32 + g) * 128  -> LBO = 4096, SBO = 128
constexpr uint32_t LBO_A = 16 * 128, LBO_B = 32 * 128, SBO = 128;
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (bits 4-5 = 1), a/b format F16 (0),
// K-major A and B, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(H >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
// the same with operand B MN-major (b_major, bit 16): B[n][k] stored with n contiguous -- the WEIGHT pass of the backward
// (K = batch rows) reads the row-major h1 image the forward kernel wrote, without a transpose.  Canonical MN-major layout,
// no swizzle (cute::UMMA make_umma_desc<Major::MN>, INTERLEAVE): 16-byte units of 8 consecutive n; consecutive k 16 B
// apart inside a group of 8 k; LBO = byte stride between k-groups, SBO = byte stride between n-groups.
constexpr uint32_t IDESC_BMN = IDESC | (1u << 16);
// h1 image (global == shared layout of one 32-row k-chunk): [hi | lo][n-group g = n / 8 (32)][row in chunk (32)] x 16 B
constexpr uint32_t LBO_BMN = 8 * 16, SBO_BMN = 32 * 16;


// ---- PTX wrappers -----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// one arrival per WARP: every lane has fenced its own shared-memory writes, __syncwarp orders them before lane 0's
// arrive (512 per-thread arrivals on one mbarrier per k-chunk serialise in the barrier unit)
constexpr int kProdWarps = 16;
__device__ __forceinline__ void mbar_arrive(uint32_t bar);
__device__ __forceinline__ void mbar_arrive_warp(uint32_t bar) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_f16_idesc(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    umma_f16_idesc(d_tmem, adesc, bdesc, accumulate, IDESC);
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 |
// version 1 << 46 | layout SWIZZLE_NONE (0) << 61
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
           ((uint64_t)1 << 46);
}

// x = hi + lo with both terms fp16, for 8 values at once: hi = x with the mantissa TRUNCATED to 10 bits (a bit mask, exactly
// representable in fp16, so no fp16 -> fp32 round trip is needed for the residual), lo = fp16(x - hi); packed
// cvt.rn.f16x2.f32 conversions.  hi + lo carries ~21 significand bits.  Inputs must be within the fp16 range.
__device__ __forceinline__ void split8(const float (&v)[8], uint4* hi, uint4* lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float a = v[2 * i], b = v[2 * i + 1];
        const float ah = __uint_as_float(__float_as_uint(a) & 0xFFFFE000u);
        const float bh = __uint_as_float(__float_as_uint(b) & 0xFFFFE000u);
        const __half2 hh = __floats2half2_rn(ah, bh);
        const __half2 ll = __floats2half2_rn(a - ah, b - bh);
        h[i] = *reinterpret_cast<const uint32_t*>(&hh);
        l[i] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    *hi = make_uint4(h[0], h[1], h[2], h[3]);
    *lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// The 32 x 32 block a warp holds after tcgen05.ld (lane = row, v[j] = column j) -> row-major float4 pieces so that the
// global accesses that follow are coalesced (4 rows x 128 B per warp instruction): f(row_in_block, c4, value).
// scratch: 32 x 36 floats private to the warp (stride 36: both phases are bank-conflict free).
constexpr int kXposeFloats = 32 * 36;
template <typename F>
__device__ __forceinline__ void warp_block_rows(float* scratch, const float (&v)[32], int lane, F&& f) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(scratch + lane * 36 + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int rl = i * 4 + (lane >> 3), c4 = lane & 7;
        f(rl, c4, *reinterpret_cast<const float4*>(scratch + rl * 36 + 4 * c4));
    }
    __syncwarp();
}

// the same, also handing the callback its compile-time step i (rows i * 4 + (lane >> 3)): f(i, row_in_block, c4, value)
template <typename F>
__device__ __forceinline__ void warp_block_rows_i(float* scratch, const float (&v)[32], int lane, F&& f) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(scratch + lane * 36 + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int rl = i * 4 + (lane >> 3), c4 = lane & 7;
        f(i, rl, c4, *reinterpret_cast<const float4*>(scratch + rl * 36 + 4 * c4));
    }
    __syncwarp();
}

}  // namespace tc
}  // namespace rrl
