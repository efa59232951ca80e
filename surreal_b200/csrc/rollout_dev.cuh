// Device-side building blocks shared by the per-step rollout kernels (rollout.cu) and the persistent rollout
// kernel (rollout_fused.cu): FIFO control block, replay-slot assignment, per-actor window commit.
#pragma once
#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// ExpSenderWrapperMultiStepMovingWindowWithInfo._step (exp_sender_wrapper.py:209-228), batched.
// Slot assignment (one block, ordered): detect the deques that reach n_step with this step's append and give
// them FIFO slots in (step, actor) order with drop-oldest at `capacity` (fifo_replay.py:27).  It depends only
// on the deque lengths, so it can run before the environment has stepped (fused into the sampling kernel).
struct FifoState {
    int head;        // physical index of the oldest window
    int count;       // windows currently queued
    int capacity;    // memory_size + 3
    int dropped;     // windows silently dropped so far (diagnostic)
    long long total_in;
    long long total_out;
    unsigned int ticket;   // last-block ticket of the fused sampling kernel (self-resetting)
    int pad_;
};

constexpr int SLOT_NONE = -1;      // deque not full after this step
constexpr int SLOT_DROPPED = -2;   // window completes but falls straight out of the deque(maxlen)

// dest[i] <- physical slot | SLOT_NONE | SLOT_DROPPED; advances the queue.  Called by all threads of ONE block.
__device__ void assign_window_slots(int N, int n_step, const int* __restrict__ stage_pos, int* __restrict__ dest,
                                    FifoState* fifo) {
    __shared__ int warp_tot[32];
    __shared__ int warp_excl[32];
    __shared__ int chunk_total;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    int base = 0;                                  // windows completed by actors before this chunk
    for (int i0 = 0; i0 < N; i0 += nt) {
        const int i = i0 + tid;
        const int flag = (i < N && stage_pos[i] + 1 == n_step) ? 1 : 0;
        int incl = flag;                           // inclusive scan inside the warp (actor order)
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int w = (lane < nw) ? warp_tot[lane] : 0;
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
        if (i < N) dest[i] = flag ? (base + warp_excl[warp] + incl - flag) : SLOT_NONE;
        base += chunk_total;
        __syncthreads();
    }
    // translate ranks into physical slots; advance the queue
    const int K = base;
    const int cap = fifo->capacity;
    const int head = fifo->head, count = fifo->count;
    __syncthreads();
    for (int i = tid; i < N; i += nt) {
        const int r = dest[i];
        if (r >= 0) {
            // more arrivals than the deque holds: the earliest of THIS step fall out immediately
            dest[i] = (K > cap && r < K - cap) ? SLOT_DROPPED : (int)(((long long)head + count + r) % cap);
        }
    }
    if (tid == 0) {
        int nc = count + K, nh = head, dr = 0;
        if (nc > cap) {                       // deque(maxlen): the oldest entries fall out
            dr = nc - cap;
            nh = (int)(((long long)head + dr) % cap);
            nc = cap;
        }
        fifo->head = nh;
        fifo->count = nc;
        fifo->dropped += dr;
        fifo->total_in += K;
    }
}

__device__ unsigned int* fifo_ticket(FifoState* fifo) { return &fifo->ticket; }

__device__ __forceinline__ void copy_floats(float* __restrict__ dst, const float* __restrict__ src, int count, int g,
                                            int G) {
    if (((count & 3) == 0) && ((((uintptr_t)dst) | ((uintptr_t)src)) & 15) == 0) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(dst);
        for (int k = g; k < (count >> 2); k += G) d4[k] = s4[k];
    } else {
        for (int k = g; k < count; k += G) dst[k] = src[k];
    }
}

// Commit of one actor's step by a group of G threads (g = index inside the group).  `pos` points at the actor's
// deque length (global or shared memory), `slot` is the record slot its window ships to when it completes with this
// step (SLOT_NONE: not complete, SLOT_DROPPED: complete but not stored).  Appends (reward, done) at
// deque position p and obs_next at p+1; if the window completed, copies it into its replay slot and pops
// `stride` items; if the episode ended, clears the deque (exp_sender_wrapper.py:204-207) and seeds position 0
// with the reset observation.  Barriers are block-wide and executed by EVERY thread (valid or not).
__device__ void commit_actor(bool valid, int i, int g, int G, const float* __restrict__ next_row,
                             const float* __restrict__ reset_row, float rew, float dn, int n_step, int stride, int D,
                             int A, int* __restrict__ pos, int slot, float* __restrict__ stage_obs,
                             float* __restrict__ stage_act, float* __restrict__ stage_pd,
                             float* __restrict__ stage_rew, float* __restrict__ stage_done,
                             float* __restrict__ r_obs, float* __restrict__ r_act,
                             float* __restrict__ r_pd, float* __restrict__ r_rew, float* __restrict__ r_done) {
    const long long ii = valid ? i : 0;
    const int p = valid ? *pos : 0;
    float* so = stage_obs + ii * (n_step + 1) * D;
    float* sa = stage_act + ii * n_step * A;
    float* sp = stage_pd + ii * n_step * 2 * A;
    float* sr = stage_rew + ii * n_step;
    float* sd = stage_done + ii * n_step;
    if (valid) {
        for (int d = g; d < D; d += G) so[(long long)(p + 1) * D + d] = next_row[d];
        if (g == 0) {
            sr[p] = rew;
            sd[p] = dn;
        }
    }
    __syncthreads();
    int len = p + 1;
    if (!valid) slot = SLOT_NONE;
    const bool complete = (slot != SLOT_NONE);               // len == n_step
    if (slot >= 0) {                                         // ship the window
        copy_floats(r_obs + (long long)slot * (n_step + 1) * D, so, (n_step + 1) * D, g, G);
        copy_floats(r_act + (long long)slot * n_step * A, sa, n_step * A, g, G);
        copy_floats(r_pd + (long long)slot * n_step * 2 * A, sp, n_step * 2 * A, g, G);
        copy_floats(r_rew + (long long)slot * n_step, sr, n_step, g, G);
        copy_floats(r_done + (long long)slot * n_step, sd, n_step, g, G);
    }
    __syncthreads();
    const int pop = min(stride, n_step);                     // uniform: every completing deque holds n_step items
    const int keep = n_step - pop;
    if (keep > 0) {                                          // overlapping windows: slide the deque down
        for (int k0 = 0; k0 < (keep + 1) * D; k0 += G) {
            const int k = k0 + g;
            const bool on = complete && k < (keep + 1) * D;
            float v = 0.f;
            if (on) v = so[(long long)pop * D + k];
            __syncthreads();
            if (on) so[k] = v;
            __syncthreads();
        }
        for (int k0 = 0; k0 < keep * 2 * A; k0 += G) {
            const int k = k0 + g;
            float va = 0.f, vp = 0.f;
            if (complete && k < keep * A) va = sa[(long long)pop * A + k];
            if (complete && k < keep * 2 * A) vp = sp[(long long)pop * 2 * A + k];
            __syncthreads();
            if (complete && k < keep * A) sa[k] = va;
            if (complete && k < keep * 2 * A) sp[k] = vp;
            __syncthreads();
        }
        for (int k0 = 0; k0 < keep; k0 += G) {
            const int k = k0 + g;
            float vr = 0.f, vd = 0.f;
            if (complete && k < keep) { vr = sr[pop + k]; vd = sd[pop + k]; }
            __syncthreads();
            if (complete && k < keep) { sr[k] = vr; sd[k] = vd; }
            __syncthreads();
        }
    } else if (complete) {
        for (int d = g; d < D; d += G) so[d] = next_row[d];                 // obs_next -> next window's first obs
    }
    if (complete) len = keep;
    __syncthreads();
    if (valid && dn > 0.5f) {                                // episode over: deque cleared, new episode's first obs
        for (int d = g; d < D; d += G) so[d] = reset_row[d];
        len = 0;
    }
    if (valid && g == 0) *pos = len;
}

}  // namespace
