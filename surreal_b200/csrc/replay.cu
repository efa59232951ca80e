// HBM-resident replay: FIFO queue control, gather of sampled records into the learner's batch buffers,
// the DDPG n-step SSAR staging + UniformReplay ring insert, and a host-side MT19937 that reproduces
// CPython's random.randint index stream bit for bit.
//
// Replaces surreal/replay/fifo_replay.py:6-48, uniform_replay.py:6-74, the aggregators' np.stack loops
// (aggregator.py:52-103,151-262) and ExpSenderWrapperSSARNStepBootstrap (exp_sender_wrapper.py:72-112).
// Records are stored SoA (one array per field, record-major) so that every field of a record is one
// contiguous, 16-byte aligned run: gathers move whole sectors.
#include "common.cuh"
#include <string.h>

namespace {

struct FifoState {          // must match rollout.cu
    int head;
    int count;
    int capacity;
    int dropped;
    long long total_in;
    long long total_out;
    unsigned int ticket;     // used by the fused sampling kernel (rollout.cu)
    int pad_;
};

// pop `batch` oldest windows: idx[k] = physical slot of the k-th oldest (fifo_replay.py:37-39).
__global__ void fifo_pop_kernel(FifoState* fifo, int batch, int* __restrict__ idx, int* __restrict__ status) {
    const int tid = threadIdx.x;
    const int head = fifo->head, count = fifo->count, cap = fifo->capacity;
    const bool ok = count >= batch;
    for (int k = tid; k < batch; k += blockDim.x) idx[k] = ok ? (int)(((long long)head + k) % cap) : 0;
    __syncthreads();
    if (tid == 0) {
        if (ok) {
            fifo->head = (int)(((long long)head + batch) % cap);
            fifo->count = count - batch;
            fifo->total_out += batch;
            *status = 0;
        } else {
            *status = 1;               // popleft from a too-short deque would raise IndexError in the reference
        }
    }
}

// host-driven push of `k` windows that already sit in slots chosen by this kernel (used by Replay.insert()).
__global__ void fifo_push_kernel(FifoState* fifo, int k, int* __restrict__ slots) {
    if (threadIdx.x != 0) return;
    int head = fifo->head, count = fifo->count;
    const int cap = fifo->capacity;
    for (int j = 0; j < k; ++j) {
        slots[j] = (int)(((long long)head + count) % cap);
        if (count == cap) {            // deque(maxlen) drops the oldest
            head = (head + 1) % cap;
            fifo->dropped += 1;
        } else {
            count += 1;
        }
    }
    fifo->head = head;
    fifo->count = count;
    fifo->total_in += k;
}

// out[b] = src[idx[b]] for records of `rec` floats (rec*4 bytes contiguous per record).
// One warp per record, float4 lanes when aligned: full 32-byte sectors on both sides.
__global__ void __launch_bounds__(256) gather_kernel(const float* __restrict__ src, long long rec,
                                                     const int* __restrict__ idx, const long long* __restrict__ idx64,
                                                     int batch, float* __restrict__ out, int vec) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int b = warp; b < batch; b += nwarps) {
        const long long s = (idx64 != nullptr) ? idx64[b] : (long long)idx[b];
        const float* sp = src + s * rec;
        float* dp = out + (long long)b * rec;
        if (vec) {
            const float4* s4 = reinterpret_cast<const float4*>(sp);
            float4* d4 = reinterpret_cast<float4*>(dp);
            for (long long k = lane; k < rec / 4; k += 32) d4[k] = s4[k];
        } else {
            for (long long k = lane; k < rec; k += 32) dp[k] = sp[k];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// ExpSenderWrapperSSARNStepBootstrap._step (exp_sender_wrapper.py:96-112), batched, + UniformReplay.insert.
// Per actor a deque of < n_step pending transitions [obs, action, reward(fp64), done]; every step each
// pending entry gets obs_next/done overwritten and reward += gamma^(n_step-i-1) * r (exponent by deque
// POSITION, the reference's warm-up quirk); the new transition is appended; when the deque holds n_step
// entries the oldest is emitted.  Emitted records go to ring slot (next_idx + rank) % capacity in actor
// order (uniform_replay.py:36-41).
struct UniformState {
    long long next_idx;
    long long size;
    long long capacity;
    long long total_in;
};

__global__ void __launch_bounds__(1024) ssar_flags_kernel(const float* __restrict__ reward,
                                                          const float* __restrict__ done, int N, int n_step,
                                                          double gamma, int* __restrict__ dq_len,
                                                          double* __restrict__ dq_rew, int* __restrict__ dest,
                                                          float* __restrict__ emit_rew, UniformState* us,
                                                          unsigned long long* step_ctr) {
    __shared__ int warp_tot[32];
    __shared__ int warp_excl[32];
    __shared__ int chunk_total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int base = 0;
    for (int i0 = 0; i0 < N; i0 += 1024) {
        const int i = i0 + tid;
        int flag = 0;
        if (i < N) {
            const int len = dq_len[i];
            const double r = (double)reward[i];
            double* q = dq_rew + (long long)i * n_step;
            for (int e = 0; e < len; ++e) q[e] += pow(gamma, (double)(n_step - e - 1)) * r;
            q[len] = r;                                   // append (position len)
            flag = (len + 1 == n_step) ? 1 : 0;
            if (flag) emit_rew[i] = (float)q[0];          // popleft's accumulated reward -> fp32 at the learner
        }
        int incl = flag;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int w = warp_tot[lane];
            int wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            warp_excl[lane] = wi - w;
            if (lane == 31) chunk_total = wi;
        }
        __syncthreads();
        if (i < N) dest[i] = flag ? (base + warp_excl[warp] + incl - flag) : -1;
        base += chunk_total;
        __syncthreads();
    }
    const long long K = base, cap = us->capacity, nxt = us->next_idx;
    for (int i = tid; i < N; i += 1024) {
        const int r = dest[i];
        if (r >= 0) dest[i] = (K > cap && r < K - cap) ? -1 : (int)((nxt + r) % cap);
    }
    __syncthreads();
    if (tid == 0) {
        us->next_idx = (nxt + K) % cap;
        us->size = (us->size + K > cap) ? cap : us->size + K;
        us->total_in += K;
        if (step_ctr != nullptr) *step_ctr += 1ull;
    }
}

// one block per actor: maintain the obs/action deque, write the emitted record into its replay slot.
__global__ void __launch_bounds__(128) ssar_commit_kernel(const float* __restrict__ obs, const float* __restrict__ action,
                                                          const float* __restrict__ obs_next,
                                                          const float* __restrict__ done, int N, int n_step, int D, int A,
                                                          int* __restrict__ dq_len, float* __restrict__ dq_obs,
                                                          float* __restrict__ dq_act, double* __restrict__ dq_rew,
                                                          const int* __restrict__ dest, const float* __restrict__ emit_rew,
                                                          float* __restrict__ r_obs, float* __restrict__ r_obs_next,
                                                          float* __restrict__ r_act, float* __restrict__ r_rew,
                                                          float* __restrict__ r_done) {
    const int i = blockIdx.x, tid = threadIdx.x;
    const int len = dq_len[i];
    float* qo = dq_obs + (long long)i * n_step * D;
    float* qa = dq_act + (long long)i * n_step * A;
    double* qr = dq_rew + (long long)i * n_step;
    // append the new transition's obs/action at position len
    for (int d = tid; d < D; d += blockDim.x) qo[(long long)len * D + d] = obs[(long long)i * D + d];
    for (int j = tid; j < A; j += blockDim.x) qa[(long long)len * A + j] = action[(long long)i * A + j];
    __syncthreads();
    int nlen = len + 1;
    const int slot = dest[i];
    const bool emit = (nlen == n_step);
    if (emit) {
        if (slot >= 0) {
            for (int d = tid; d < D; d += blockDim.x) {
                r_obs[(long long)slot * D + d] = qo[d];
                r_obs_next[(long long)slot * D + d] = obs_next[(long long)i * D + d];
            }
            for (int j = tid; j < A; j += blockDim.x) r_act[(long long)slot * A + j] = qa[j];
            if (tid == 0) {
                r_rew[slot] = emit_rew[i];
                r_done[slot] = done[i];
            }
        }
        __syncthreads();
        // popleft: shift the remaining n_step-1 entries down by one
        for (int e = 0; e + 1 < nlen; ++e) {
            for (int d = tid; d < D; d += blockDim.x) qo[(long long)e * D + d] = qo[(long long)(e + 1) * D + d];
            for (int j = tid; j < A; j += blockDim.x) qa[(long long)e * A + j] = qa[(long long)(e + 1) * A + j];
            __syncthreads();
        }
        if (tid == 0)
            for (int e = 0; e + 1 < nlen; ++e) qr[e] = qr[e + 1];
        nlen -= 1;
    }
    __syncthreads();
    if (done[i] > 0.5f) nlen = 0;                          // _reset clears the deque (exp_sender_wrapper.py:91-94)
    if (tid == 0) dq_len[i] = nlen;
}

// ---- host: CPython-compatible MT19937 ----------------------------------------------------------
struct MT {
    uint32_t mt[624];
    int idx;
};

void mt_init_genrand(MT* g, uint32_t s) {
    g->mt[0] = s;
    for (int i = 1; i < 624; ++i) g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
    g->idx = 624;
}

void mt_init_by_array(MT* g, const uint32_t* key, int klen) {
    mt_init_genrand(g, 19650218u);
    int i = 1, j = 0;
    uint32_t* mt = g->mt;
    for (int k = (624 > klen ? 624 : klen); k; --k) {
        mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
        ++i; ++j;
        if (i >= 624) { mt[0] = mt[623]; i = 1; }
        if (j >= klen) j = 0;
    }
    for (int k = 623; k; --k) {
        mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
        ++i;
        if (i >= 624) { mt[0] = mt[623]; i = 1; }
    }
    mt[0] = 0x80000000u;
    g->idx = 624;
}

inline uint32_t mt_next(MT* g) {
    uint32_t* mt = g->mt;
    if (g->idx >= 624) {
        for (int k = 0; k < 624; ++k) {
            const uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
            mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        g->idx = 0;
    }
    uint32_t y = mt[g->idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

// All fields of a record in ONE launch: out_f[b] = src_f[idx[b]] for f < nf.  A warp owns a record; its lanes walk the
// record's chunks across fields (16-byte chunks for fields whose record size / alignment allow, 4-byte otherwise), so a
// 552-byte SSAR record (obs 64, obs_next 64, action 8, reward 1, done 1 floats) is two loads + two stores per lane.
constexpr int GATHER_MAX_FIELDS = 8;
struct GatherFields {
    const float* src[GATHER_MAX_FIELDS];
    float* out[GATHER_MAX_FIELDS];
    long long rec[GATHER_MAX_FIELDS];     // floats per record
    int unit[GATHER_MAX_FIELDS];          // 4 (float4 chunks) or 1
    int chunk_end[GATHER_MAX_FIELDS];     // prefix sums of chunks per record
    int nf;
};

__global__ void __launch_bounds__(256) gather_multi_kernel(const __grid_constant__ GatherFields f, const int* __restrict__ idx,
                                                           const long long* __restrict__ idx64, int batch) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int total = f.chunk_end[f.nf - 1];
    for (int b = warp; b < batch; b += nwarps) {
        const long long s = (idx64 != nullptr) ? idx64[b] : (long long)idx[b];
        for (int c = lane; c < total; c += 32) {
            int q = 0;
#pragma unroll
            for (int k = 0; k < GATHER_MAX_FIELDS - 1; ++k) q += (k < f.nf - 1 && c >= f.chunk_end[k]) ? 1 : 0;
            const int local = c - (q > 0 ? f.chunk_end[q - 1] : 0);
            if (f.unit[q] == 4) {
                reinterpret_cast<float4*>(f.out[q] + (long long)b * f.rec[q])[local] =
                    ld_stream4(f.src[q] + s * f.rec[q] + 4 * local);
            } else {
                f.out[q][(long long)b * f.rec[q] + local] = __ldg(f.src[q] + s * f.rec[q] + local);
            }
        }
    }
}

}  // namespace

extern "C" int sb200_fifo_pop(void* fifo_state, int batch, int* idx, int* status, void* stream) {
    SB200_REQUIRE(fifo_state && idx && status && batch >= 1);
    fifo_pop_kernel<<<1, 256, 0, (cudaStream_t)stream>>>((FifoState*)fifo_state, batch, idx, status);
    return sb200_launch_status();
}

extern "C" int sb200_fifo_push(void* fifo_state, int k, int* slots, void* stream) {
    SB200_REQUIRE(fifo_state && slots && k >= 1);
    fifo_push_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((FifoState*)fifo_state, k, slots);
    return sb200_launch_status();
}

extern "C" int sb200_replay_gather_f32(const float* src, int64_t record_floats, const int* idx32, const int64_t* idx64,
                                       int batch, float* out, void* stream) {
    SB200_REQUIRE(src && out && (idx32 || idx64) && batch >= 1 && record_floats >= 1);
    const int vec = (record_floats % 4 == 0) && ((((uintptr_t)src) | ((uintptr_t)out)) & 15) == 0;
    int blocks = (batch + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    gather_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, record_floats, idx32, (const long long*)idx64, batch, out,
                                                            vec);
    return sb200_launch_status();
}

extern "C" int sb200_replay_gather_multi_f32(const float* const* srcs, float* const* outs, const int64_t* record_floats,
                                             int nfields, const int* idx32, const int64_t* idx64, int batch, void* stream) {
    SB200_REQUIRE(srcs && outs && record_floats && (idx32 || idx64) && batch >= 1);
    SB200_REQUIRE(nfields >= 1 && nfields <= GATHER_MAX_FIELDS);
    GatherFields f;
    int end = 0;
    for (int k = 0; k < nfields; ++k) {
        SB200_REQUIRE(srcs[k] && outs[k] && record_floats[k] >= 1 && record_floats[k] < (1ll << 28));
        f.src[k] = srcs[k];
        f.out[k] = outs[k];
        f.rec[k] = record_floats[k];
        const bool vec = (record_floats[k] % 4 == 0) && ((((uintptr_t)srcs[k]) | ((uintptr_t)outs[k])) & 15) == 0;
        f.unit[k] = vec ? 4 : 1;
        end += (int)(record_floats[k] / f.unit[k]);
        f.chunk_end[k] = end;
    }
    for (int k = nfields; k < GATHER_MAX_FIELDS; ++k) { f.src[k] = nullptr; f.out[k] = nullptr; f.rec[k] = 0; f.unit[k] = 1; f.chunk_end[k] = end; }
    f.nf = nfields;
    int blocks = (batch + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    gather_multi_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(f, idx32, (const long long*)idx64, batch);
    return sb200_launch_status();
}

extern "C" size_t sb200_uniform_state_bytes(void) { return sizeof(UniformState); }

extern "C" int sb200_ssar_step_f32(const float* obs, const float* action, const float* obs_next, const float* reward,
                                   const float* done, int N, int n_step, double gamma, int D, int A, int* dq_len,
                                   float* dq_obs, float* dq_act, double* dq_rew, int* dest_scratch, float* emit_scratch,
                                   void* uniform_state, float* r_obs, float* r_obs_next, float* r_act, float* r_rew,
                                   float* r_done, uint64_t* step_counter, void* stream) {
    SB200_REQUIRE(obs && action && obs_next && reward && done && dq_len && dq_obs && dq_act && dq_rew);
    SB200_REQUIRE(dest_scratch && emit_scratch && uniform_state && r_obs && r_obs_next && r_act && r_rew && r_done);
    SB200_REQUIRE(N >= 1 && n_step >= 1 && D >= 1 && A >= 1);
    cudaStream_t st = (cudaStream_t)stream;
    ssar_flags_kernel<<<1, 1024, 0, st>>>(reward, done, N, n_step, gamma, dq_len, dq_rew, dest_scratch, emit_scratch,
                                          (UniformState*)uniform_state, (unsigned long long*)step_counter);
    ssar_commit_kernel<<<N, 128, 0, st>>>(obs, action, obs_next, done, N, n_step, D, A, dq_len, dq_obs, dq_act, dq_rew,
                                          dest_scratch, emit_scratch, r_obs, r_obs_next, r_act, r_rew, r_done);
    return sb200_launch_status(2);
}

// ---- CPython random.Random index stream (host) -------------------------------------------------
extern "C" size_t sb200_mt19937_state_bytes(void) { return sizeof(MT); }

extern "C" int sb200_mt19937_seed_h(void* state_h, const uint32_t* key_h, int key_len) {
    SB200_REQUIRE(state_h && key_h && key_len >= 1);
    mt_init_by_array((MT*)state_h, key_h, key_len);
    return SB200_OK;
}

extern "C" int sb200_mt19937_set_state_h(void* state_h, const uint32_t* words624_h, int index) {
    SB200_REQUIRE(state_h && words624_h && index >= 0 && index <= 624);
    memcpy(((MT*)state_h)->mt, words624_h, 624 * sizeof(uint32_t));
    ((MT*)state_h)->idx = index;
    return SB200_OK;
}

extern "C" int sb200_mt19937_get_state_h(const void* state_h, uint32_t* words624_h, int* index) {
    SB200_REQUIRE(state_h && words624_h && index);
    memcpy(words624_h, ((const MT*)state_h)->mt, 624 * sizeof(uint32_t));
    *index = ((const MT*)state_h)->idx;
    return SB200_OK;
}

extern "C" int sb200_mt19937_randint_fill_h(void* state_h, int64_t population, int64_t count, int64_t* out_h) {
    // random.randint(0, population-1) == _randbelow_with_getrandbits(population), population < 2^32:
    // k = population.bit_length(); r = getrandbits(k) = genrand_uint32() >> (32-k); repeat until r < population.
    SB200_REQUIRE(state_h && out_h && population >= 1 && population <= 0xffffffffLL && count >= 0);
    MT* g = (MT*)state_h;
    int k = 0;
    for (uint64_t v = (uint64_t)population; v; v >>= 1) ++k;
    for (int64_t c = 0; c < count; ++c) {
        uint32_t r;
        do { r = mt_next(g) >> (32 - k); } while ((int64_t)r >= population);
        out_h[c] = (int64_t)r;
    }
    return SB200_OK;
}
