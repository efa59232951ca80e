// DDPG-specific element-wise kernels around the shared MLP forward / backward / optimiser kernels:
// Bellman target, the actor-loss seed gradient and tanh backward.
// Replaces surreal/learner/ddpg.py:261-262,279,324-331,335-341 (and torch autograd through them).
#include "common.cuh"
#include <stddef.h>

namespace {

constexpr int DT = 256;

struct DdpgWs {
    unsigned int counter;
    int bad_action;          // 1 when max|a| > 1 (ddpg.py:261-262): optimiser / target kernels of this learn() no-op
    double partial[1];       // [blocks][4]
};

// y = r + gamma^n * Q'(s', pi'(s')) * (1 - done)   (ddpg.py:279); also mean(r), mean(|a|_2), max|a| for the
// reference's sanity asserts (ddpg.py:261-262) without host syncs.
__global__ void __launch_bounds__(DT) ddpg_target_kernel(const float* __restrict__ rewards,
                                                         const float* __restrict__ q_next, long long ldq,
                                                         const float* __restrict__ dones,
                                                         const float* __restrict__ actions, long long lda, int B, int A,
                                                         float discount, float* __restrict__ y,
                                                         float* __restrict__ stats, DdpgWs* ws,
                                                         const float* __restrict__ q_next2, long long ldq2) {
    __shared__ double sh[32];
    const int tid = threadIdx.x, b = blockIdx.x * DT + tid;
    double r = 0.0, an = 0.0, amax = 0.0, yy = 0.0;
    if (b < B) {
        const float rb = rewards[b];
        const float t = __fmul_rn(__fmul_rn(discount, q_next[(long long)b * ldq]), __fsub_rn(1.0f, dones[b]));
        float yb = __fadd_rn(rb, t);
        if (q_next2 != nullptr) {                            // TD3 double critic: y = min(y, y2) (ddpg.py:280-283)
            const float t2 = __fmul_rn(__fmul_rn(discount, q_next2[(long long)b * ldq2]), __fsub_rn(1.0f, dones[b]));
            yb = fminf(yb, __fadd_rn(rb, t2));
        }
        y[b] = yb;
        yy = (double)yb;
        r = (double)rb;
        float s = 0.0f, mx = 0.0f;
        for (int j = 0; j < A; ++j) {
            const float a = actions[(long long)b * lda + j];
            s += a * a;
            mx = fmaxf(mx, fabsf(a));
        }
        an = (double)sqrtf(s);
        amax = (double)mx;
    }
    double* part = ws->partial + (size_t)blockIdx.x * 4;
    double t;
    t = block_sum(r, sh);   if (tid == 0) part[0] = t;
    t = block_sum(an, sh);  if (tid == 0) part[1] = t;
    t = block_sum(yy, sh);  if (tid == 0) part[2] = t;
    // max via sum-free reduction: reuse block_sum on an indicator is wrong; do a proper max
    __syncthreads();
    double m = amax;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((tid & 31) == 0) sh[tid >> 5] = m;
    __syncthreads();
    if (tid == 0) {
        double mm = 0.0;
        for (int w = 0; w < DT / 32; ++w) mm = fmax(mm, sh[w]);
        part[3] = mm;
    }
    if (last_block_ticket(&ws->counter, gridDim.x)) {
        if (tid == 0) {
            double s0 = 0, s1 = 0, s2 = 0, mx = 0;
            for (unsigned int k = 0; k < gridDim.x; ++k) {
                s0 += ws->partial[(size_t)k * 4 + 0];
                s1 += ws->partial[(size_t)k * 4 + 1];
                s2 += ws->partial[(size_t)k * 4 + 2];
                mx = fmax(mx, ws->partial[(size_t)k * 4 + 3]);
            }
            stats[SB200_DSTAT_REWARDS] = (float)(s0 / B);
            stats[SB200_DSTAT_ACTION_NORM] = (float)(s1 / B);
            stats[SB200_DSTAT_Q_TARGET] = (float)(s2 / B);
            stats[SB200_DSTAT_ACTION_ABSMAX] = (float)mx;
            ws->bad_action = (mx > 1.0) ? 1 : 0;
        }
    }
}

// critic loss: MSE(q, y) (nn.MSELoss, ddpg.py:305) -> dq = 2(q-y)/B; stats: critic_loss, Q_policy.
__global__ void __launch_bounds__(DT) ddpg_critic_loss_kernel(const float* __restrict__ q, long long ldq,
                                                              const float* __restrict__ y, int B,
                                                              float* __restrict__ dq, long long ldd,
                                                              float* __restrict__ stats, DdpgWs* ws) {
    __shared__ double sh[32];
    const int tid = threadIdx.x, b = blockIdx.x * DT + tid;
    double se = 0.0, qs = 0.0;
    if (b < B) {
        const float qq = q[(long long)b * ldq];
        const float d = qq - y[b];
        se = (double)d * (double)d;
        qs = (double)qq;
        dq[(long long)b * ldd] = (float)(2.0 * (double)d / (double)B);
        for (int j = 1; j < (int)ldd; ++j) dq[(long long)b * ldd + j] = 0.0f;
    }
    double* part = ws->partial + (size_t)blockIdx.x * 4;
    double t;
    t = block_sum(se, sh);  if (tid == 0) part[0] = t;
    t = block_sum(qs, sh);  if (tid == 0) part[1] = t;
    if (last_block_ticket(&ws->counter, gridDim.x)) {
        if (tid == 0) {
            double s0 = 0, s1 = 0;
            for (unsigned int k = 0; k < gridDim.x; ++k) {
                s0 += ws->partial[(size_t)k * 4 + 0];
                s1 += ws->partial[(size_t)k * 4 + 1];
            }
            stats[SB200_DSTAT_CRITIC_LOSS] = (float)(s0 / B);
            stats[SB200_DSTAT_Q_POLICY] = (float)(s1 / B);
        }
    }
}

// actor loss = -mean Q(s, pi(s)) (ddpg.py:324-329): seed gradient dQ = -1/B for the critic's output; stat.
__global__ void __launch_bounds__(DT) ddpg_actor_seed_kernel(const float* __restrict__ q_pi, long long ldq, int B,
                                                             float* __restrict__ dq, long long ldd,
                                                             float* __restrict__ stats, DdpgWs* ws) {
    __shared__ double sh[32];
    const int tid = threadIdx.x, b = blockIdx.x * DT + tid;
    double qs = 0.0;
    if (b < B) {
        qs = (double)q_pi[(long long)b * ldq];
        dq[(long long)b * ldd] = (float)(-1.0 / (double)B);
        for (int j = 1; j < (int)ldd; ++j) dq[(long long)b * ldd + j] = 0.0f;
    }
    const double t = block_sum(qs, sh);
    if (tid == 0) ws->partial[blockIdx.x] = t;
    if (last_block_ticket(&ws->counter, gridDim.x)) {
        if (tid == 0) {
            double s = 0;
            for (unsigned int k = 0; k < gridDim.x; ++k) s += ws->partial[k];
            stats[SB200_DSTAT_ACTOR_LOSS] = (float)(-s / B);
        }
    }
}

// dpre = dout * (1 - out^2): backward through the actor's tanh (builders.py:50)
__global__ void tanh_bwd_kernel(const float* __restrict__ dout, long long ldo, const float* __restrict__ out,
                                long long ldy, int B, int A, float* __restrict__ dpre, long long ldd) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * ldd) return;
    const int b = (int)(i / ldd), j = (int)(i - (long long)b * ldd);
    float v = 0.0f;
    if (j < A) {
        const float o = out[(long long)b * ldy + j];
        v = dout[(long long)b * ldo + j] * (1.0f - o * o);
    }
    dpre[(long long)b * ldd + j] = v;
}

inline int nb(int B) { return (B + DT - 1) / DT; }


// TD3 target-policy smoothing (ddpg.py:267-278): out = clamp(pi + clip(N(0, policy_noise), -c, c), -1, 1).
// unit_noise [B][A] holds injected N(0,1) draws (tests) or NULL for Philox keyed by (seed, *step_counter, row).
__global__ void __launch_bounds__(256) ddpg_smooth_action_kernel(const float* __restrict__ pi, long long ldp,
                                                                 const float* __restrict__ unit_noise, int B, int A,
                                                                 float policy_noise, float noise_clip,
                                                                 unsigned long long seed,
                                                                 const unsigned long long* __restrict__ step_ctr,
                                                                 float* __restrict__ out, long long ldo) {
    const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
    if (idx >= (long long)B * A) return;
    const int b = (int)(idx / A), j = (int)(idx - (long long)b * A);
    float e;
    if (unit_noise != nullptr) {
        e = unit_noise[idx];
    } else {
        const unsigned long long ctr = (step_ctr != nullptr) ? *step_ctr : 0ull;
        const Philox4 r = philox4x32_10(seed, ctr, ((unsigned long long)b << 16) | (unsigned long long)(j >> 2));
        const float2 z01 = box_muller(r.x, r.y), z23 = box_muller(r.z, r.w);
        const int c = j & 3;
        e = (c == 0) ? z01.x : (c == 1) ? z01.y : (c == 2) ? z23.x : z23.y;
    }
    const float nz = fminf(fmaxf(__fmul_rn(e, policy_noise), -noise_clip), noise_clip);
    out[(long long)b * ldo + j] = fminf(fmaxf(__fadd_rn(pi[(long long)b * ldp + j], nz), -1.0f), 1.0f);
}

}  // namespace

extern "C" size_t sb200_ddpg_bad_action_offset(void) { return offsetof(DdpgWs, bad_action); }

extern "C" size_t sb200_ddpg_workspace_bytes(int B) { return sizeof(DdpgWs) + (size_t)nb(B) * 4 * sizeof(double); }

extern "C" int sb200_ddpg_target_f32(const float* rewards, const float* q_next, int64_t ldq, const float* dones,
                                     const float* actions, int64_t lda, int B, int A, double discount, float* y,
                                     float* stats, void* workspace, void* stream) {
    SB200_REQUIRE(rewards && q_next && dones && actions && y && stats && workspace && B >= 1 && A >= 1);
    ddpg_target_kernel<<<nb(B), DT, 0, (cudaStream_t)stream>>>(rewards, q_next, ldq, dones, actions, lda, B, A,
                                                              (float)discount, y, stats, (DdpgWs*)workspace, nullptr, 0);
    return sb200_launch_status();
}

extern "C" int sb200_ddpg_target2_f32(const float* rewards, const float* q_next, int64_t ldq, const float* q_next2,
                                      int64_t ldq2, const float* dones, const float* actions, int64_t lda, int B, int A,
                                      double discount, float* y, float* stats, void* workspace, void* stream) {
    SB200_REQUIRE(rewards && q_next && q_next2 && dones && actions && y && stats && workspace && B >= 1 && A >= 1);
    ddpg_target_kernel<<<nb(B), DT, 0, (cudaStream_t)stream>>>(rewards, q_next, ldq, dones, actions, lda, B, A,
                                                              (float)discount, y, stats, (DdpgWs*)workspace, q_next2,
                                                              ldq2);
    return sb200_launch_status();
}

extern "C" int sb200_ddpg_smooth_action_f32(const float* pi, int64_t ldp, const float* unit_noise, int B, int A,
                                            double policy_noise, double noise_clip, uint64_t seed,
                                            const uint64_t* step_counter, float* out, int64_t ldo, void* stream) {
    SB200_REQUIRE(pi && out && B >= 1 && A >= 1 && ldp >= A && ldo >= A);
    const long long n = (long long)B * A;
    ddpg_smooth_action_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        pi, ldp, unit_noise, B, A, (float)policy_noise, (float)noise_clip, (unsigned long long)seed,
        (const unsigned long long*)step_counter, out, ldo);
    return sb200_launch_status();
}

extern "C" int sb200_ddpg_critic_loss_f32(const float* q, int64_t ldq, const float* y, int B, float* dq, int64_t ldd,
                                          float* stats, void* workspace, void* stream) {
    SB200_REQUIRE(q && y && dq && stats && workspace && B >= 1 && ldd >= 1);
    ddpg_critic_loss_kernel<<<nb(B), DT, 0, (cudaStream_t)stream>>>(q, ldq, y, B, dq, ldd, stats, (DdpgWs*)workspace);
    return sb200_launch_status();
}

extern "C" int sb200_ddpg_actor_seed_f32(const float* q_pi, int64_t ldq, int B, float* dq, int64_t ldd, float* stats,
                                         void* workspace, void* stream) {
    SB200_REQUIRE(q_pi && dq && stats && workspace && B >= 1 && ldd >= 1);
    ddpg_actor_seed_kernel<<<nb(B), DT, 0, (cudaStream_t)stream>>>(q_pi, ldq, B, dq, ldd, stats, (DdpgWs*)workspace);
    return sb200_launch_status();
}

extern "C" int sb200_tanh_bwd_f32(const float* dout, int64_t ldo, const float* out, int64_t ldy, int B, int A,
                                  float* dpre, int64_t ldd, void* stream) {
    SB200_REQUIRE(dout && out && dpre && B >= 1 && A >= 1 && ldd >= A);
    const long long n = (long long)B * ldd;
    tanh_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dout, ldo, out, ldy, B, A, dpre, ldd);
    return sb200_launch_status();
}
