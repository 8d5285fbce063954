// replay.cu -- HBM-resident replay rings + a CPython-`random`-compatible index sampler fused with
// the gather.  Replaces recovery_rl/replay_memory.py:11-75 and the stdlib calls it makes
// (random.seed / random.sample -> _randbelow_with_getrandbits -> MT19937 genrand_uint32).
//
// Layout: ring of 32-byte records {s.x, s.y, a.x, a.y, r|c, s2.x, s2.y, mask} (fp32): one DRAM
// sector per transition, so push is a coalesced stream and the random gather costs exactly one
// sector per sampled row.  The constraint ring adds one flag byte per slot
// (bit0: pos_idx != 0, bit1: (1 - pos_idx) != 0; replay_memory.py:45,51,58-66).
//
// The sampler is one CTA: MT19937 regeneration is done block-parallel (4 dependency phases), the
// common "set" path of random.sample (n > setsize) is evaluated block-parallel as "first k distinct
// accepted draws of the stream" (hash table + prefix scan), the "pool" path (n <= setsize, only at
// the very start of a run) is the sequential partial Fisher-Yates on one thread.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMT = 624;
constexpr uint32_t kUpper = 0x80000000u, kLower = 0x7fffffffu, kMatrixA = 0x9908b0dfu;

__host__ __device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

__device__ __forceinline__ uint32_t mt_mix(uint32_t cur, uint32_t nxt, uint32_t far) {
    uint32_t y = (cur & kUpper) | (nxt & kLower);
    return far ^ (y >> 1) ^ ((y & 1u) ? kMatrixA : 0u);
}

struct SamplerSmem {
    uint32_t mt[kMT];      // raw state
    uint32_t mt_new[kMT];  // scratch for the parallel twist
    uint32_t tb[kMT];      // tempered outputs of mt[]
    int idx;               // CPython's `index` (next word to consume)
    int scan[kThreads + 8];
    int i_done;            // pool path progress
    int kept;              // set path progress
    int cutoff;
    int flag;
};

// Regenerate the 624-word state (CPython genrand_uint32's `if (self->index >= N)` block), block-parallel.
__device__ void mt_refill(SamplerSmem& S) {
    const int t = threadIdx.x;
    for (int k = t; k < 227; k += kThreads) S.mt_new[k] = mt_mix(S.mt[k], S.mt[k + 1], S.mt[k + 397]);
    __syncthreads();
    for (int k = 227 + t; k < 454; k += kThreads) S.mt_new[k] = mt_mix(S.mt[k], S.mt[k + 1], S.mt_new[k - 227]);
    __syncthreads();
    for (int k = 454 + t; k < 623; k += kThreads) S.mt_new[k] = mt_mix(S.mt[k], S.mt[k + 1], S.mt_new[k - 227]);
    __syncthreads();
    if (t == 0) S.mt_new[623] = mt_mix(S.mt[623], S.mt_new[0], S.mt_new[396]);
    __syncthreads();
    for (int k = t; k < kMT; k += kThreads) {
        uint32_t v = S.mt_new[k];
        S.mt[k] = v;
        S.tb[k] = mt_temper(v);
    }
    if (t == 0) S.idx = 0;
    __syncthreads();
}

__device__ __forceinline__ int bit_length(uint32_t n) { return 32 - __clz(n); }

// setsize of random.sample: 21 + 4 ** ceil(log(3k, 4)) for k > 5 (3k is never a power of 4)
__device__ __forceinline__ int64_t sample_setsize(int k) {
    int64_t setsize = 21;
    if (k > 5) {
        int64_t p = 1;
        while (p < 3 * (int64_t)k) p *= 4;
        setsize += p;
    }
    return setsize;
}

__device__ __forceinline__ uint32_t hash_slot(uint32_t key, int log_t) { return (key * 2654435761u) >> (32 - log_t); }

// block exclusive scan of one int per thread; returns exclusive prefix, *total = sum
__device__ int block_exscan(SamplerSmem& S, int v, int* total) {
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    __syncthreads();
    if (lane == 31) S.scan[w] = inc;
    __syncthreads();
    if (w == 0) {
        int x = lane < (kThreads / 32) ? S.scan[lane] : 0;
        int xi = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int u = __shfl_up_sync(0xffffffffu, xi, o);
            if (lane >= o) xi += u;
        }
        if (lane < (kThreads / 32)) S.scan[16 + lane] = xi - x;
        if (lane == 31) S.scan[32] = xi;
    }
    __syncthreads();
    *total = S.scan[32];
    return inc - v + S.scan[16 + w];
}

// random.sample(range(n), k) -> result[0..k).  work: >= max(2 * tab_size, setsize) ints.
__device__ void sample_range(SamplerSmem& S, int64_t n64, int k, int* result, int* work, int tab_size, int log_t) {
    const int t = threadIdx.x;
    if (k <= 0) return;
    const uint32_t n = (uint32_t)n64;
    if (n64 <= sample_setsize(k)) {
        // ---- pool path: partial Fisher-Yates, sequential on thread 0 --------------------------
        int* pool = work;
        for (int i = t; i < (int)n; i += kThreads) pool[i] = i;
        if (t == 0) S.i_done = 0;
        __syncthreads();
        while (true) {
            if (S.idx >= kMT) mt_refill(S);
            __syncthreads();
            if (t == 0) {
                int i = S.i_done, idx = S.idx;
                while (i < k && idx < kMT) {
                    const uint32_t nn = n - (uint32_t)i;
                    const uint32_t r = S.tb[idx++] >> (32 - bit_length(nn));
                    if (r < nn) {
                        result[i] = pool[r];
                        pool[r] = pool[nn - 1];
                        ++i;
                    }
                }
                S.i_done = i;
                S.idx = idx;
            }
            __syncthreads();
            if (S.i_done >= k) break;
        }
        return;
    }
    // ---- set path: first k distinct accepted draws, block-parallel ------------------------------
    int* tab_key = work;
    int* tab_pos = work + tab_size;
    for (int i = t; i < tab_size; i += kThreads) {
        tab_key[i] = -1;
        tab_pos[i] = 0x7fffffff;
    }
    if (t == 0) S.kept = 0;
    __syncthreads();
    const int shift = 32 - bit_length(n);
    int base = 0;  // stream position of tb[0] of the current block
    while (true) {
        if (S.idx >= kMT) mt_refill(S);
        __syncthreads();
        const int idx0 = S.idx, kept0 = S.kept;
        // each thread owns 3 consecutive stream elements
        uint32_t r[3];
        int slot[3];
        bool acc[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int i = idx0 + 3 * t + j;
            acc[j] = false;
            slot[j] = 0;
            r[j] = 0;
            if (i < kMT) {
                r[j] = S.tb[i] >> shift;
                acc[j] = r[j] < n;
                if (acc[j]) {
                    uint32_t h = hash_slot(r[j], log_t);
                    while (true) {
                        int prev = atomicCAS(&tab_key[h], -1, (int)r[j]);
                        if (prev == -1 || prev == (int)r[j]) break;
                        h = (h + 1) & (tab_size - 1);
                    }
                    slot[j] = (int)h;
                    atomicMin(&tab_pos[h], base + i);
                }
            }
        }
        __syncthreads();
        int first[3], cnt = 0;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int i = idx0 + 3 * t + j;
            first[j] = (acc[j] && tab_pos[slot[j]] == base + i) ? 1 : 0;
            cnt += first[j];
        }
        int total;
        int ex = block_exscan(S, cnt, &total);
        const int need = k - kept0;
        if (t == 0) S.cutoff = -1;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            if (first[j]) {
                if (ex < need) result[kept0 + ex] = (int)r[j];
                if (ex == need - 1) S.cutoff = idx0 + 3 * t + j + 1;  // last consumed word
                ++ex;
            }
        }
        __syncthreads();
        if (total >= need) {
            if (t == 0) {
                S.idx = S.cutoff;
                S.kept = k;
            }
            __syncthreads();
            break;
        }
        if (t == 0) {
            S.kept = kept0 + total;
            S.idx = kMT;
        }
        base += kMT;
        __syncthreads();
    }
}

// rank -> slot of the rank-th flagged entry (ascending).  One THREAD per query: all k binary searches over the chunk
// prefix run side by side, then every thread walks its own chunk in 64-byte steps (4 independent 16-byte loads in
// flight, popcount per group) and resolves the byte inside the group from registers.  The flag bytes were read by
// flag_count_kernel just before, so the walk is served by L2; k x chunk bytes in total.
__device__ void select_ranks(const uint8_t* __restrict__ flags, int64_t capacity, int chunk, int n_chunks,
                             const int* __restrict__ prefix /*exclusive, n_chunks+1*/, uint8_t bit, int* result, int k) {
    const uint32_t m = 0x01010101u * bit;
    for (int q = threadIdx.x; q < k; q += kThreads) {
        const int rank = result[q];
        int lo = 0, hi = n_chunks;  // find c: prefix[c] <= rank < prefix[c+1]
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (prefix[mid] <= rank) lo = mid; else hi = mid;
        }
        int rel = rank - prefix[lo];
        const int64_t cbase = (int64_t)lo * chunk;
        int64_t found = -1;
        for (int j = 0; j < chunk && found < 0; j += 64) {   // chunk is a multiple of 512, capacity of 16
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t b = cbase + j + 16 * u;
                v[u] = b < capacity ? *reinterpret_cast<const uint4*>(flags + b) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                for (int wi = 0; wi < 4; ++wi) {
                    const int c = __popc(w[wi] & m);
                    if (found < 0 && rel < c) {
#pragma unroll
                        for (int by = 0; by < 4; ++by) {
                            if (found < 0 && ((w[wi] >> (8 * by)) & bit)) {
                                if (rel == 0) found = cbase + j + 16 * u + 4 * wi + by;
                                --rel;
                            }
                        }
                    } else if (found < 0) {
                        rel -= c;
                    }
                }
            }
        }
        result[q] = (int)found;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kThreads)
replay_sample_kernel(rrl_sample_config_t cfg, const float* __restrict__ ring, const uint8_t* __restrict__ flags,
                     const int32_t* __restrict__ chunk_counts, int n_chunks, uint32_t* __restrict__ mt_state,
                     int64_t* __restrict__ counters, int rows_counter, int64_t* __restrict__ out_idx,
                     float* __restrict__ out_s, float* __restrict__ out_a, float* __restrict__ out_r,
                     float* __restrict__ out_s2, float* __restrict__ out_m, int tab_size, int log_t, int work_ints) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SamplerSmem& S = *reinterpret_cast<SamplerSmem*>(smem_raw);
    int* result = reinterpret_cast<int*>(smem_raw + ((sizeof(SamplerSmem) + 15) / 16) * 16);
    int* work = result + ((cfg.batch_size + 3) / 4) * 4;
    int* prefix_pos = work + work_ints;
    int* prefix_neg = prefix_pos + (n_chunks + 1);
    const int t = threadIdx.x;
    const int B = cfg.batch_size;
    const bool strat = cfg.is_constraint && cfg.pos_fraction >= 0.0;
    pdl_wait();   // programmatic dependent launch (common.cuh)

    // ---- gates and effective batch size (experiment.py:397,407-410; qrisk.py:100-104) ---------
    const int64_t len = counters[cfg.is_constraint ? RRL_C_CONS_LEN : RRL_C_TASK_LEN];
    bool open = true;
    if (cfg.gate_mode == 1) open = len > B;
    if (cfg.gate_mode == 2) {
        const int64_t viols = counters[RRL_C_NUM_VIOLS] + counters[RRL_C_OFFLINE_VIOLS] + counters[RRL_C_EXT_VIOLS];
        // nested inside `if len(self.memory) > batch_size` (experiment.py:397): needs the task ring gate too
        open = (counters[RRL_C_TASK_LEN] > B) && (len > B) && ((double)viols / (double)B > cfg.gate_pos_fraction);
    }
    int rows = 0;
    if (open) {
        int64_t r = len < B ? len : B;
        if (cfg.is_constraint && cfg.pos_fraction > 0.0) {
            int64_t lim = (int64_t)((1.0 - cfg.pos_fraction) * (double)len);
            r = lim < B ? lim : B;
        }
        rows = (int)r;
    }
    if (rows <= 0) {
        if (t == 0) counters[rows_counter] = 0;
        return;
    }
    for (int i = t; i < kMT; i += kThreads) {
        uint32_t v = mt_state[i];
        S.mt[i] = v;
        S.tb[i] = mt_temper(v);
    }
    if (t == 0) S.idx = (int)mt_state[kMT];
    __syncthreads();

    int err = 0;
    if (!strat) {
        sample_range(S, len, rows, result, work, tab_size, log_t);
    } else {
        // exclusive prefix over the per-chunk counts: per-thread stripe + block scan
        const int stripe = (n_chunks + kThreads - 1) / kThreads;
        for (int which = 0; which < 2; ++which) {
            const int32_t* c = chunk_counts + (int64_t)which * n_chunks;
            int* p = which == 0 ? prefix_pos : prefix_neg;
            const int lo = t * stripe, hi = min(n_chunks, lo + stripe);
            int local = 0;
            for (int i = lo; i < hi; ++i) local += c[i];
            int total;
            int acc = block_exscan(S, local, &total);
            for (int i = lo; i < hi; ++i) {
                p[i] = acc;
                acc += c[i];
            }
            if (t == 0) p[n_chunks] = total;
            __syncthreads();
        }
        const int n_pos = prefix_pos[n_chunks], n_neg = prefix_neg[n_chunks];
        const int pos_size = (int)((double)rows * cfg.pos_fraction);  // int(batch_size * pos_fraction)
        const int neg_size = rows - pos_size;
        if (pos_size > n_pos || neg_size > n_neg) {
            err = 1;  // random.sample: "Sample larger than population"
        } else {
            sample_range(S, n_pos, pos_size, result, work, tab_size, log_t);
            __syncthreads();
            sample_range(S, n_neg, neg_size, result + pos_size, work, tab_size, log_t);
            __syncthreads();
            select_ranks(flags, cfg.capacity, cfg.chunk, n_chunks, prefix_pos, 1, result, pos_size);
            select_ranks(flags, cfg.capacity, cfg.chunk, n_chunks, prefix_neg, 2, result + pos_size, neg_size);
        }
    }
    __syncthreads();
    if (err) {
        if (t == 0) {
            counters[rows_counter] = 0;
            counters[RRL_C_ERROR] = 1;
        }
        return;
    }
    // ---- gather (np.stack over the sampled tuples, replay_memory.py:29,71) ----------------------
    for (int i = t; i < rows; i += kThreads) {
        const int64_t slot = result[i];
        const float4* rec = reinterpret_cast<const float4*>(ring + slot * 8);
        const float4 lo = __ldg(rec), hi = __ldg(rec + 1);
        if (out_idx) out_idx[i] = slot;
        reinterpret_cast<float2*>(out_s)[i] = make_float2(lo.x, lo.y);
        reinterpret_cast<float2*>(out_a)[i] = make_float2(lo.z, lo.w);
        out_r[i] = hi.x;
        reinterpret_cast<float2*>(out_s2)[i] = make_float2(hi.y, hi.z);
        out_m[i] = hi.w;
    }
    for (int i = t; i < kMT; i += kThreads) mt_state[i] = S.mt[i];
    if (t == 0) {
        mt_state[kMT] = (uint32_t)S.idx;
        counters[rows_counter] = rows;
    }
}

// per-chunk counts of positive (bit 1) and negative (bit 2) flag bytes: one WARP per chunk, no block-level sync, every
// chunk's loads in flight at once (the ring is scanned at HBM speed instead of one DRAM round trip per chunk and block)
__global__ void __launch_bounds__(256)
flag_count_kernel(const uint8_t* __restrict__ flags, int64_t capacity, int chunk, int n_chunks,
                  int32_t* __restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const int n_warps = gridDim.x * 8;
    for (int c = blockIdx.x * 8 + (threadIdx.x >> 5); c < n_chunks; c += n_warps) {
        const int64_t base = (int64_t)c * chunk;
        int p = 0, q = 0;
        for (int j = lane * 16; j < chunk; j += 32 * 16) {
            if (base + j < capacity) {  // capacity is padded to a multiple of 16 by the caller
                uint4 v = *reinterpret_cast<const uint4*>(flags + base + j);
                p += __popc(v.x & 0x01010101u) + __popc(v.y & 0x01010101u) + __popc(v.z & 0x01010101u) + __popc(v.w & 0x01010101u);
                q += __popc(v.x & 0x02020202u) + __popc(v.y & 0x02020202u) + __popc(v.z & 0x02020202u) + __popc(v.w & 0x02020202u);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            p += __shfl_xor_sync(0xffffffffu, p, o);
            q += __shfl_xor_sync(0xffffffffu, q, o);
        }
        if (lane == 0) {
            counts[c] = p;
            counts[n_chunks + c] = q;
        }
    }
}

__global__ void __launch_bounds__(256)
replay_push_kernel(float* __restrict__ ring, uint8_t* __restrict__ flags, int64_t capacity,
                   const float* __restrict__ rec, int64_t n, int64_t* __restrict__ counters, int is_cons) {
    const int64_t pos = counters[is_cons ? RRL_C_CONS_POS : RRL_C_TASK_POS];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t slot = (pos + i) % capacity;
        const float4 lo = reinterpret_cast<const float4*>(rec)[2 * i], hi = reinterpret_cast<const float4*>(rec)[2 * i + 1];
        reinterpret_cast<float4*>(ring)[2 * slot] = lo;
        reinterpret_cast<float4*>(ring)[2 * slot + 1] = hi;
        if (flags) flags[slot] = (uint8_t)((hi.x != 0.0f ? 1 : 0) | ((1.0f - hi.x) != 0.0f ? 2 : 0));
    }
}

__global__ void replay_push_commit_kernel(int64_t* counters, int64_t n, int64_t capacity, int is_cons) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const int p = is_cons ? RRL_C_CONS_POS : RRL_C_TASK_POS, l = is_cons ? RRL_C_CONS_LEN : RRL_C_TASK_LEN;
        counters[p] = (counters[p] + n) % capacity;
        int64_t len = counters[l] + n;
        counters[l] = len < capacity ? len : capacity;
    }
}

}  // namespace

// random.seed(int): Python/random.py seed() -> _random.Random.seed -> init_by_array(key) (CPython
// Modules/_randommodule.c).  key_limbs = little-endian 32-bit limbs of |seed| (>= 1 limb).
extern "C" int rrl_mt19937_seed_host(const uint32_t* key, int n_limbs, uint32_t* st) {
    RRL_CHECK_ARG(key && st && n_limbs >= 1, "bad key");
    uint32_t* mt = st;
    mt[0] = 19650218u;
    for (int i = 1; i < kMT; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
    int i = 1, j = 0;
    int k = kMT > n_limbs ? kMT : n_limbs;
    for (; k; --k) {
        mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
        ++i;
        ++j;
        if (i >= kMT) { mt[0] = mt[kMT - 1]; i = 1; }
        if (j >= n_limbs) j = 0;
    }
    for (k = kMT - 1; k; --k) {
        mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
        ++i;
        if (i >= kMT) { mt[0] = mt[kMT - 1]; i = 1; }
    }
    mt[0] = 0x80000000u;
    st[kMT] = kMT;  // index: regenerate on first draw
    return 0;
}

extern "C" int rrl_replay_push(float* ring, uint8_t* cons_flags, int64_t capacity, const float* rec, int64_t n,
                               int64_t* counters, int is_constraint_buffer, void* stream) {
    RRL_CHECK_ARG(ring && rec && counters, "null argument");
    RRL_CHECK_ARG(n >= 0 && n <= capacity, "push larger than the ring capacity");
    RRL_CHECK_ARG(!is_constraint_buffer || cons_flags, "constraint buffer needs a flag array");
    if (n == 0) return 0;
    int blocks = (int)((n + 255) / 256);
    if (blocks > rrl_num_sms() * 8) blocks = rrl_num_sms() * 8;
    replay_push_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(ring, is_constraint_buffer ? cons_flags : nullptr,
                                                                 capacity, rec, n, counters, is_constraint_buffer);
    RRL_CHECK_LAUNCH();
    replay_push_commit_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(counters, n, capacity, is_constraint_buffer);
    RRL_CHECK_LAUNCH();
    return 0;
}

static int n_chunks_for(int64_t capacity, int chunk) { return (int)((capacity + chunk - 1) / chunk); }

extern "C" int rrl_replay_flag_count(const uint8_t* cons_flags, int64_t capacity, int32_t chunk,
                                     int32_t* chunk_counts, void* stream) {
    RRL_CHECK_ARG(cons_flags && chunk_counts, "null argument");
    RRL_CHECK_ARG(chunk >= 512 && (chunk % 512) == 0, "chunk must be a multiple of 512");
    RRL_CHECK_ARG(capacity % 16 == 0, "flag array capacity must be a multiple of 16");
    const int n_chunks = n_chunks_for(capacity, chunk);
    const int want = (n_chunks + 7) / 8;
    int blocks = want < rrl_num_sms() * 8 ? want : rrl_num_sms() * 8;
    flag_count_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(cons_flags, capacity, chunk, n_chunks, chunk_counts);
    RRL_CHECK_LAUNCH();
    return 0;
}

extern "C" int rrl_replay_sample(const rrl_sample_config_t* cfg, const float* ring, const uint8_t* cons_flags,
                                 const int32_t* chunk_counts, uint32_t* mt_state, int64_t* counters,
                                 int rows_counter, int64_t* out_idx, float* out_s, float* out_a, float* out_r,
                                 float* out_s2, float* out_m, void* stream) {
    RRL_CHECK_ARG(cfg && ring && mt_state && counters && out_s && out_a && out_r && out_s2 && out_m, "null argument");
    RRL_CHECK_ARG(cfg->batch_size >= 1 && cfg->batch_size <= 4096, "batch_size must be in [1, 4096]");
    RRL_CHECK_ARG(cfg->capacity >= 1 && cfg->capacity < (int64_t)0x7fffffff, "capacity must be < 2^31");
    RRL_CHECK_ARG(rows_counter >= 0 && rows_counter < RRL_NUM_COUNTERS, "bad rows counter");
    const bool strat = cfg->is_constraint && cfg->pos_fraction >= 0.0;
    int n_chunks = 0;
    if (strat) {
        RRL_CHECK_ARG(cons_flags && chunk_counts, "stratified sampling needs flags and chunk counts");
        RRL_CHECK_ARG(cfg->chunk >= 512 && (cfg->chunk % 512) == 0, "chunk must be a multiple of 512");
        n_chunks = n_chunks_for(cfg->capacity, cfg->chunk);
        RRL_CHECK_ARG(n_chunks <= 8192, "too many flag chunks; use a larger chunk");
        RRL_CHECK_ARG(cfg->capacity % 16 == 0, "stratified sampling needs a capacity that is a multiple of 16");
    }
    const int B = cfg->batch_size;
    int tab = 1, log_t = 0;
    while (tab < 2 * (B + kMT)) { tab <<= 1; ++log_t; }
    int64_t setsize = 21;
    if (B > 5) { int64_t p = 1; while (p < 3 * (int64_t)B) p *= 4; setsize += p; }
    int work_ints = 2 * tab;
    if (work_ints < setsize) work_ints = (int)setsize;
    work_ints = (work_ints + 3) & ~3;
    size_t smem = ((sizeof(SamplerSmem) + 15) / 16) * 16 + (size_t)((B + 3) / 4 * 4) * 4 + (size_t)work_ints * 4 +
                  (size_t)2 * (n_chunks + 1) * 4;
    RRL_CHECK_ARG(smem <= 200 * 1024, "sampler shared memory exceeds 200 KB");
    static size_t configured = 0;
    if (smem > configured) {
        RRL_CUDA(cudaFuncSetAttribute(replay_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    RRL_CUDA(rrl_launch_pdl(replay_sample_kernel, dim3(1), dim3(kThreads), smem, (cudaStream_t)stream, *cfg, ring, cons_flags,
                            chunk_counts, n_chunks, mt_state, counters, rows_counter, out_idx, out_s, out_a, out_r, out_s2, out_m,
                            tab, log_t, work_ints));
    return 0;
}
