// Persistent learner kernel: ALL minibatch epochs of one optimiser of PPOLearner._optimize in ONE launch.
//
// Replaces, per epoch, the launch chain  forward -> loss -> (dX, dW) x 3 layers -> slab reduce + grad norm -> clip + Adam
// -> post-step forward -> KL  (surreal/learner/ppo.py:194-353 losses and updates, 541-557 the epoch loops and the KL
// early stop; torch autograd + torch.optim.Adam in the reference).  At 1024-row minibatches every one of those ~13
// launches is a few microseconds of arithmetic behind launch latency, cold weights and a drain (profiles/r02a_launches.md:
// 54 % of the step's kernel time).  Here the grid stays resident, phases are separated by grid barriers (one atomic
// counter in L2), and the work of a phase is a static list of 64x64 GEMM tiles / row blocks dealt round-robin to the CTAs:
//
//   F0  x_in = zf(x)            (once per launch: the batch and the z-filter statistics are fixed during learn())
//   F1  h1 = relu(x_in W1 + b1)                 tiles (M/64) x (H1/64)
//   F2  h2 = relu(h1 W2 + b2)                   tiles (M/64) x (H2/64)
//   L   head + loss, one warp per row:          out = act3(h2 W3 + b3), KL(ref || cur) partial, loss row, dpre,
//                                               d2 = (dpre W3^T) * relu'(h2), loss partials per CTA
//   -- gate: mean KL of the forward just done = post-step KL of the previous epoch (ppo.py:553-556): stats, early stop
//   B2  d1 = (d2 W2^T) * relu'(h1)   tiles;     dW2 = h1^T d2, dW3 = h2^T dpre      (split over M into `splits` slabs)
//   B1  dW1 = x_in^T d1
//   R   grad = sum of slabs (fixed order), squared-norm partial per CTA      [data-parallel: + peer all-reduce, below]
//   A   clip by global norm + Adam (torch's arithmetic, optim_dev.cuh), step count
//
// Value mode (critic, MSE on returns) runs the same phases without the gate and without the trailing forward.
// Everything is deterministic: static tile assignment, fixed-order sums.  Buffers that other CTAs rewrite between
// barriers are read with ld.global.cg only (gemm_tiles.cuh).
//
// Data-parallel learner (N > 1): the two exchanges of an epoch happen INSIDE the kernel over NVLink peer memory, with
// the protocol of peer_allreduce.cu on the same symmetric buffer: the KL scalar (CTA 0) and the flat gradient (CTAs
// 0..15 own a slice each: publish to the rank's slot -> flag -> wait for the peers' flags -> sum the W slots in rank
// order -> mean), so an epoch costs no launch and no NCCL call.  Grid barriers on both sides of an exchange give the
// slot-reuse guarantee that stream order gives the stand-alone kernel.
#include <math.h>
#include <stddef.h>

#include "common.cuh"
#include "gemm_tiles.cuh"
#include "optim_dev.cuh"
#include "ppo_loss_dev.cuh"

namespace {

using optim_dev::OptWs;
using ppo_dev::MAX_A;

constexpr int ET = 256;               // threads per CTA
constexpr int EW = ET / 32;           // warps per CTA
constexpr int EP_MAX_G = 296;
constexpr int EP_MAX_OUT = 32;
constexpr int EP_SLOTS = 4 + EP_MAX_OUT;

struct EpWs {                         // global scratch of one launch (caller-owned, sb200_ppo_epochs_workspace_bytes)
    unsigned int bar;                 // grid-barrier arrival counter; zeroed by the wrapper before every launch
    unsigned int pad[3];
    double kl_global;                 // data-parallel: the rank-averaged KL, published by CTA 0
    double pad2;
    double kl_part[EP_MAX_G];
    double sq_part[EP_MAX_G];
    double loss_part[EP_MAX_G * EP_SLOTS];
};

// ---- peer exchange (same layout as peer_allreduce.cu)
constexpr int PAR_MAX_WORLD = 8;
constexpr int PAR_MAX_CTAS = 16;
struct ParHeader {
    unsigned int flags[PAR_MAX_WORLD * PAR_MAX_CTAS];
    unsigned int counter;
    unsigned int ticket;
    unsigned int pad[2];
    double partial[PAR_MAX_CTAS];
};
struct ParCtx {
    void* peers[PAR_MAX_WORLD];
    int world;
    int rank;
    long long max_floats;
};
__device__ __forceinline__ float* slot_of(void* base, long long max_floats, unsigned int parity) {
    return reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(base) + sizeof(ParHeader)) + (size_t)parity * (size_t)max_floats;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_sys1(const float* p) {
    float r;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(r) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ double ld_sys_f64(const double* p) {
    double r;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(r) : "l"(p) : "memory");
    return r;
}

struct EpParams {
    // network (3 layers) inside the flat parameter buffer
    float* params;
    long long n_params;
    int w_off[3], b_off[3], ldw[3];
    int D, H1, H2, NO;
    int act_out;                      // activation of the head (tanh for the policy mean, none for the value)
    int extra_off;                    // policy: offset of log_var[NO] in the flat buffer
    // input rows
    const float* x;
    long long ldx;
    int M;
    const float* zf;
    float zf_eps;
    // activations / gradients (trainer buffers)
    float *x_in, *h1, *h2, *out, *d1, *d2, *dpre;
    int ld_x, ld_h1, ld_h2, ld_out, ld_dpre;
    // optimiser
    float *slabs, *grad, *m, *v;
    int splits, rows_per_split;
    const double* lr;
    double weight_decay, clip_value;
    int clip_mode;
    OptWs* opt;
    float* norm_out;
    // loss
    int mode;                         // 0 clip, 1 adapt, 2 value
    const float* actions; long long lda;
    const float* adv;
    const float* behave; long long ldb;
    const float* ref; long long ldr;
    const float* returns;
    const double* hyper;
    double eta, kl_target, stop_threshold;
    float* stats;
    int* stop;
    int epochs;
    EpWs* ws;
    int shift;                        // rotation of the item -> CTA map (keeps two concurrent launches off the same SMs)
    ParCtx par;                       // world == 1: single process
};

__device__ __forceinline__ void grid_bar(unsigned int* ctr, unsigned int& target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();
        atomicAdd(ctr, 1u);
        for (;;) {
            unsigned int v;
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
            if ((int)(v - target) >= 0) break;
        }
        __threadfence();
    }
    __syncthreads();
}

// first item of this CTA in a phase of n items (then stride gridDim.x)
__device__ __forceinline__ int first_item(int shift) {
    const int G = (int)gridDim.x;
    return (int)(((long long)blockIdx.x - shift % G + G) % G);
}

__global__ void __launch_bounds__(ET, 2) ppo_epochs_kernel(const __grid_constant__ EpParams p) {
    __shared__ __align__(16) gt::Smem sm;
    __shared__ float s_sig[MAX_A];
    __shared__ double s_red[32];
    __shared__ double s_bc[4];
    __shared__ optim_dev::AdamCoef s_adam;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = (int)gridDim.x, c = (int)blockIdx.x;
    if (p.stop != nullptr && *p.stop) return;              // raised only behind a barrier every CTA has passed: uniform

    const int M = p.M, D = p.D, H1 = p.H1, H2 = p.H2, NO = p.NO;
    const float* W1 = p.params + p.w_off[0];
    const float* W2 = p.params + p.w_off[1];
    const float* W3 = p.params + p.w_off[2];
    const float* b1 = p.params + p.b_off[0];
    const float* b2 = p.params + p.b_off[1];
    const float* b3 = p.params + p.b_off[2];
    const int ldw3 = p.ldw[2];
    unsigned int bar_target = 0;
    // ---------------- F0: z-filtered input rows (z_filter.py:59-79), zero padded to ld_x
    {
        const float cnt = (p.zf != nullptr) ? p.zf[2 * D] : 1.0f;
        for (long long i = (long long)c * ET + tid; i < (long long)M * D; i += (long long)G * ET) {
            const int k = (int)(i % D);
            const long long m = i / D;
            float v = p.x[m * p.ldx + k];
            if (p.zf != nullptr) {
                const float mean = p.zf[k] / cnt;
                const float var = p.zf[D + k] / cnt - mean * mean;
                v = gt::zf1(v, mean, fmaxf(sqrtf(var), p.zf_eps));
            }
            p.x_in[m * p.ld_x + k] = v;
        }
    }
    grid_bar(&p.ws->bar, bar_target);
    const int tm = (M + gt::TB - 1) / gt::TB;
    const int tn1 = (H1 + gt::TB - 1) / gt::TB, tn2 = (H2 + gt::TB - 1) / gt::TB;
    const int tkD = (D + gt::TB - 1) / gt::TB, tnO = (NO + gt::TB - 1) / gt::TB;
    const double invM = 1.0 / (double)M;
    const float c0 = (float)(0.5 * 1.8378770664093453 * (double)NO);
    const int step0 = p.opt->step;
    int steps_done = 0;
    const bool dp = p.par.world > 1;
    ParHeader* par_me = dp ? reinterpret_cast<ParHeader*>(p.par.peers[p.par.rank]) : nullptr;
    unsigned int par_k = dp ? par_me->counter : 0u;          // exchanges completed on this channel (same on every rank)
    const bool policy = p.mode != 2;
    const int last_e = policy ? p.epochs : p.epochs - 1;

    for (int e = 0; e <= last_e; ++e) {
        const bool train = e < p.epochs;                    // e == epochs: only the post-step forward + KL of the last epoch
        // ---------------- F1
        for (int it = first_item(p.shift); it < tm * tn1; it += G)
            gt::gt_tile_nn(sm, (it / tn1) * gt::TB, (it % tn1) * gt::TB, p.x_in, p.ld_x, nullptr, nullptr, W1, p.ldw[0], b1,
                           SB200_ACT_RELU, p.h1, p.ld_h1, M, H1, D);
        grid_bar(&p.ws->bar, bar_target);
        // ---------------- F2
        for (int it = first_item(p.shift); it < tm * tn2; it += G)
            gt::gt_tile_nn(sm, (it / tn2) * gt::TB, (it % tn2) * gt::TB, p.h1, p.ld_h1, nullptr, nullptr, W2, p.ldw[1], b2,
                           SB200_ACT_RELU, p.h2, p.ld_h2, M, H2, H1);
        grid_bar(&p.ws->bar, bar_target);
        // ---------------- L: head, KL partial, loss rows (clip / value: now; adapt: after the KL mean is known)
        if (policy && tid < NO) s_sig[tid] = expf(__ldcg(p.params + p.extra_off + tid));     // builders.py:127
        __syncthreads();
        double acc_slot[4] = {0.0, 0.0, 0.0, 0.0};          // per-warp partial sums (lane 0 only)
        double acc_dlv[EP_MAX_OUT];
        for (int j = 0; j < NO; ++j) acc_dlv[j] = 0.0;
        double klm = 0.0;                                   // mean KL(ref || current) of this epoch's forward
        bool leave = false;
        for (int pass = 0; pass < ((p.mode == 1 && train) ? 2 : 1); ++pass) {
            const bool do_head = pass == 0;
            const bool do_loss = train && (p.mode != 1 || pass == 1);
            const double kl_mean_now = klm;                 // adapt, pass 1: the loss needs it (ppo.py:262-270)
            for (int row = first_item(p.shift) * EW + warp; row < M; row += G * EW) {
                float mu[EP_MAX_OUT];
                const float* h2r = p.h2 + (long long)row * p.ld_h2;
                if (do_head) {
                    for (int n8 = 0; n8 < NO; n8 += 8) {
                        float s8[8];
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj) s8[jj] = 0.0f;
                        for (int k = lane; k < H2; k += 32) {
                            const float hv = __ldcg(h2r + k);
                            const float* wr = W3 + (long long)k * ldw3 + n8;
                            const float4 w0 = gt::ldcg4(wr);
                            s8[0] = fmaf(hv, w0.x, s8[0]); s8[1] = fmaf(hv, w0.y, s8[1]);
                            s8[2] = fmaf(hv, w0.z, s8[2]); s8[3] = fmaf(hv, w0.w, s8[3]);
                            if (n8 + 4 < ldw3) {
                                const float4 w1 = gt::ldcg4(wr + 4);
                                s8[4] = fmaf(hv, w1.x, s8[4]); s8[5] = fmaf(hv, w1.y, s8[5]);
                                s8[6] = fmaf(hv, w1.z, s8[6]); s8[7] = fmaf(hv, w1.w, s8[7]);
                            }
                        }
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj) {
                            const float tt = warp_sum(s8[jj]);
                            if (n8 + jj < NO) {
                                float v = tt + __ldcg(b3 + n8 + jj);
                                if (p.act_out == SB200_ACT_TANH) v = tanhf(v);
                                else if (p.act_out == SB200_ACT_RELU) v = fmaxf(v, 0.0f);
                                mu[n8 + jj] = v;
                            }
                        }
                    }
                    if (lane == 0)
                        for (int j = 0; j < NO; ++j) p.out[(long long)row * p.ld_out + j] = mu[j];
                } else {
                    for (int j = 0; j < NO; ++j) mu[j] = __ldcg(p.out + (long long)row * p.ld_out + j);
                }
                float dp_row[EP_MAX_OUT];
                for (int j = 0; j < NO; ++j) dp_row[j] = 0.0f;
                if (lane == 0) {
                    if (policy) {
                        const float* rp = p.ref + (long long)row * p.ldr;
                        if (do_head) acc_slot[3] += (double)ppo_dev::row_kl(rp, rp + NO, mu, s_sig, NO);     // KL(ref || current)
                        if (do_loss) {
                            float act[EP_MAX_OUT];
                            for (int j = 0; j < NO; ++j) act[j] = p.actions[(long long)row * p.lda + j];
                            const ppo_dev::PolicyRow pr = ppo_dev::policy_row(p.mode, mu, act, s_sig, p.behave + (long long)row * p.ldb, rp,
                                                                              p.adv[row], NO, c0, invM, p.hyper, p.eta, p.kl_target,
                                                                              kl_mean_now, dp_row, NO);
                            acc_slot[0] += (double)pr.surr;
                            acc_slot[1] += (double)pr.rowloss;
                            acc_slot[2] += (double)pr.klrow;
                            for (int j = 0; j < NO; ++j) acc_dlv[j] += (double)ppo_dev::policy_dlogvar(p.mode, j, mu, act, s_sig, rp, pr, NO);
                        }
                    } else if (do_loss) {                   // value_loss_kernel's arithmetic (ppo.py:311-332)
                        const float vv = mu[0], rr = p.returns[row];
                        const float df = vv - rr;
                        const double d = (double)(rr - vv), r = (double)rr;
                        dp_row[0] = (float)(2.0 * (double)df / (double)M);
                        acc_slot[0] += d; acc_slot[1] += d * d; acc_slot[2] += r; acc_slot[3] += r * r;
                    }
                }
                if (do_loss) {
                    for (int j = 0; j < NO; ++j) dp_row[j] = __shfl_sync(0xffffffffu, dp_row[j], 0);
                    if (lane == 0) {
                        for (int j = 0; j < NO; ++j) p.dpre[(long long)row * p.ld_dpre + j] = dp_row[j];
                        for (int j = NO; j < p.ld_dpre; ++j) p.dpre[(long long)row * p.ld_dpre + j] = 0.0f;
                    }
                    // d2 = (dpre W3^T) * relu'(h2)
                    for (int k = lane; k < H2; k += 32) {
                        const float* wr = W3 + (long long)k * ldw3;
                        float g = 0.0f;
                        for (int j = 0; j < NO; ++j) g = fmaf(dp_row[j], __ldcg(wr + j), g);
                        if (!(__ldcg(h2r + k) > 0.0f)) g = 0.0f;
                        p.d2[(long long)row * p.ld_h2 + k] = g;
                    }
                }
            }
            // per-CTA partials of this pass
            if (do_head && policy) {
                const double t = block_sum(lane == 0 ? acc_slot[3] : 0.0, s_red);
                if (tid == 0) p.ws->kl_part[c] = t;
            }
            if (do_loss) {
                double* part = p.ws->loss_part + (size_t)c * EP_SLOTS;
                const int ns = policy ? 3 : 4;
                for (int s = 0; s < ns; ++s) {
                    const double t = block_sum(lane == 0 ? acc_slot[s] : 0.0, s_red);
                    if (tid == 0) part[s] = t;
                }
                if (policy)
                    for (int j = 0; j < NO; ++j) {
                        const double t = block_sum(lane == 0 ? acc_dlv[j] : 0.0, s_red);
                        if (tid == 0) part[4 + j] = t;
                    }
            }
            grid_bar(&p.ws->bar, bar_target);
            if (pass == 0 && policy) {
                // ---- gate: mean KL(ref || current) of the forward just done
                if (tid == 0) {
                    double acc = 0.0;
                    for (int k = 0; k < G; ++k) acc += __ldcg(&p.ws->kl_part[k]);
                    s_bc[0] = (double)(float)(acc * invM);          // .mean() in fp32 (kl_kernel)
                }
                __syncthreads();
                klm = s_bc[0];
                if (dp) {                                            // average the scalar over ranks (CTA 0), publish, barrier
                    if (c == 0) {
                        const unsigned int k = par_k + 1u;
                        double* mine = reinterpret_cast<double*>(slot_of(p.par.peers[p.par.rank], p.par.max_floats, k & 1u));
                        if (tid == 0) {
                            *mine = klm;
                            __threadfence_system();
                        }
                        __syncthreads();
                        if (tid < p.par.world && tid != p.par.rank) {
                            st_release_sys(&reinterpret_cast<ParHeader*>(p.par.peers[tid])->flags[p.par.rank * PAR_MAX_CTAS], k);
                            const unsigned int* f = &par_me->flags[tid * PAR_MAX_CTAS];
                            while ((int)(ld_acquire_sys(f) - k) < 0) { }
                        }
                        __syncthreads();
                        if (tid == 0) {
                            double a = 0.0;
                            for (int q = 0; q < p.par.world; ++q)
                                a += ld_sys_f64(reinterpret_cast<const double*>(slot_of(p.par.peers[q], p.par.max_floats, k & 1u)));
                            p.ws->kl_global = (double)(float)(a * (1.0 / (double)p.par.world));
                        }
                    }
                    par_k += 1u;
                    grid_bar(&p.ws->bar, bar_target);
                    klm = __ldcg(&p.ws->kl_global);
                }
                bool stop_now = false;
                if (e > 0) {                                         // post-step KL of epoch e-1 (ppo.py:553-556)
                    stop_now = p.stop_threshold > 0.0 && klm > p.stop_threshold;
                    if (c == 0 && tid == 0) {
                        p.stats[SB200_STAT_KL_POST] = (float)klm;
                        p.stats[SB200_STAT_EPOCHS] += 1.0f;
                        if (stop_now && p.stop != nullptr) *p.stop = 1;
                    }
                }
                if (stop_now || !train) {
                    leave = true;
                    break;
                }
                if (p.mode == 1 && c == 0 && tid == 0) p.stats[SB200_STAT_KL_PRE] = (float)klm;
            }
        }
        if (leave) break;
        // ---- loss statistics + dlog_var (CTA 0); the slab it writes is consumed behind the next barriers
        if (c == 0) {
            if (policy) {
                for (int s = tid; s < 4 + NO; s += ET) {
                    if (s == 3) continue;
                    double acc = 0.0;
                    for (int k = 0; k < G; ++k) acc += __ldcg(&p.ws->loss_part[(size_t)k * EP_SLOTS + s]);
                    if (s >= 4) p.slabs[p.extra_off + (s - 4)] = (float)acc;
                    if (s == 0) p.stats[SB200_STAT_SURR] = (float)(acc * invM);
                    if (s == 1 && p.mode == 0) p.stats[SB200_STAT_LOSS] = (float)(acc * invM);
                    if (s == 0 && p.mode == 1) {
                        const double kl = klm;
                        double loss = acc * invM + p.hyper[1] * kl;
                        if (kl - 2.0 * p.kl_target > 0.0) loss += p.eta * (kl - 2.0 * p.kl_target) * (kl - 2.0 * p.kl_target);
                        p.stats[SB200_STAT_LOSS] = (float)loss;
                    }
                }
                if (tid == 0) {
                    float slog = 0.0f;
                    for (int j = 0; j < NO; ++j) slog += logf(s_sig[j]);
                    p.stats[SB200_STAT_ENTROPY] = 0.5f * slog + (float)(0.5 * 2.8378770664093453 * (double)NO);   // ppo_net.py:72
                }
            } else if (tid == 0) {
                double s[4] = {0, 0, 0, 0};
                for (int k = 0; k < G; ++k)
                    for (int q = 0; q < 4; ++q) s[q] += __ldcg(&p.ws->loss_part[(size_t)k * EP_SLOTS + q]);
                const double n = (double)M;
                const double var_d = (s[1] - s[0] * s[0] / n) / (n - 1.0);
                const double var_r = (s[3] - s[2] * s[2] / n) / (n - 1.0);
                p.stats[SB200_STAT_VAL_LOSS] = (float)(s[1] / n);
                p.stats[SB200_STAT_EXPLAINED_VAR] = (float)(1.0 - var_d / var_r);
                p.stats[SB200_STAT_RETURN_MEAN] = (float)(s[2] / n);
                for (int q = 0; q < 4; ++q) {
                    const double mq = s[q] / n;
                    const float hi = (float)mq;
                    p.stats[SB200_STAT_VAL_MOMENTS + 2 * q] = hi;
                    p.stats[SB200_STAT_VAL_MOMENTS + 2 * q + 1] = (float)(mq - (double)hi);
                }
            }
        }
        // ---------------- B2: d1 tiles, dW2 and dW3 split tiles
        {
            const int n_dx = tm * tn1;                       // d1 [M][H1]
            const int n_w2 = tn1 * tn2 * p.splits;           // dW2 [H1][H2]
            const int n_w3 = tn2 * tnO * p.splits;           // dW3 [H2][NO]
            for (int it = first_item(p.shift); it < n_dx + n_w2 + n_w3; it += G) {
                if (it < n_dx) {
                    gt::gt_tile_nt(sm, (it / tn1) * gt::TB, (it % tn1) * gt::TB, p.d2, p.ld_h2, W2, p.ldw[1], p.h1, p.ld_h1, p.d1,
                                   p.ld_h1, M, H2, H1);
                } else if (it < n_dx + n_w2) {
                    const int q = it - n_dx, z = q / (tn1 * tn2), r = q % (tn1 * tn2);
                    const int mb = z * p.rows_per_split, me = min(M, mb + p.rows_per_split);
                    float* slab = p.slabs + (long long)z * p.n_params;
                    gt::gt_tile_tn(sm, (r / tn2) * gt::TB, (r % tn2) * gt::TB, mb, me, p.h1, p.ld_h1, nullptr, nullptr, p.d2, p.ld_h2,
                                   slab + p.w_off[1], slab + p.b_off[1], p.ldw[1], H1, H2);
                } else {
                    const int q = it - n_dx - n_w2, z = q / (tn2 * tnO), r = q % (tn2 * tnO);
                    const int mb = z * p.rows_per_split, me = min(M, mb + p.rows_per_split);
                    float* slab = p.slabs + (long long)z * p.n_params;
                    gt::gt_tile_tn(sm, (r / tnO) * gt::TB, (r % tnO) * gt::TB, mb, me, p.h2, p.ld_h2, nullptr, nullptr, p.dpre, p.ld_dpre,
                                   slab + p.w_off[2], slab + p.b_off[2], p.ldw[2], H2, NO);
                }
            }
        }
        grid_bar(&p.ws->bar, bar_target);
        // ---------------- B1: dW1 split tiles
        {
            const int n_w1 = tkD * tn1 * p.splits;
            for (int it = first_item(p.shift); it < n_w1; it += G) {
                const int z = it / (tkD * tn1), r = it % (tkD * tn1);
                const int mb = z * p.rows_per_split, me = min(M, mb + p.rows_per_split);
                float* slab = p.slabs + (long long)z * p.n_params;
                gt::gt_tile_tn(sm, (r / tn1) * gt::TB, (r % tn1) * gt::TB, mb, me, p.x_in, p.ld_x, nullptr, nullptr, p.d1, p.ld_h1,
                               slab + p.w_off[0], slab + p.b_off[0], p.ldw[0], D, H1);
            }
        }
        grid_bar(&p.ws->bar, bar_target);
        // ---------------- R: slabs -> grad (fixed order) and the squared-norm partial
        {
            double sq = 0.0;
            for (long long i = (long long)c * ET + tid; i < p.n_params; i += (long long)G * ET) {
                float g = __ldcg(p.slabs + i);
                for (int z = 1; z < p.splits; ++z) g += __ldcg(p.slabs + (long long)z * p.n_params + i);
                p.grad[i] = g;
                sq += (double)g * (double)g;
            }
            if (!dp) {
                const double t = block_sum(sq, s_red);
                if (tid == 0) p.ws->sq_part[c] = t;
            }
        }
        grid_bar(&p.ws->bar, bar_target);
        if (dp) {
            // ---------------- X: all-reduce (mean) of the flat gradient over NVLink peer memory, CTAs 0..PG-1 own a slice each
            const int PG = min(G, PAR_MAX_CTAS);
            const unsigned int k = par_k + 1u;
            double sq = 0.0;
            if (c < PG) {
                long long per = (p.n_params + PG - 1) / PG;
                per = (per + 3) & ~3ll;
                const long long lo = (long long)c * per;
                const long long hi = (lo + per < p.n_params) ? lo + per : p.n_params;
                float* mine = slot_of(p.par.peers[p.par.rank], p.par.max_floats, k & 1u);
                for (long long i = lo + tid; i < hi; i += ET) mine[i] = __ldcg(p.grad + i);
                __threadfence_system();
                __syncthreads();
                if (tid < p.par.world && tid != p.par.rank) {
                    st_release_sys(&reinterpret_cast<ParHeader*>(p.par.peers[tid])->flags[p.par.rank * PAR_MAX_CTAS + c], k);
                    const unsigned int* f = &par_me->flags[tid * PAR_MAX_CTAS + c];
                    while ((int)(ld_acquire_sys(f) - k) < 0) { }
                }
                __syncthreads();
                const float scale = 1.0f / (float)p.par.world;
                for (long long i = lo + tid; i < hi; i += ET) {
                    float a = 0.f;
                    for (int q = 0; q < p.par.world; ++q) a += ld_sys1(slot_of(p.par.peers[q], p.par.max_floats, k & 1u) + i);
                    a = __fmul_rn(a, scale);
                    p.grad[i] = a;
                    sq += (double)a * (double)a;
                }
            }
            const double t = block_sum(sq, s_red);
            if (tid == 0) p.ws->sq_part[c] = t;
            par_k += 1u;
            grid_bar(&p.ws->bar, bar_target);
        }
        // ---------------- A: global norm, clip, Adam
        steps_done += 1;
        if (tid == 0) {
            double acc = 0.0;
            for (int k = 0; k < G; ++k) acc += __ldcg(&p.ws->sq_part[k]);
            const float total_norm = (float)sqrt(acc);
            s_adam = optim_dev::adam_coef(step0 + steps_done, p.lr[0], 0.9, 0.999, 1e-8, p.weight_decay, p.clip_mode, p.clip_value,
                                          total_norm);
            if (c == 0) {
                p.opt->step = step0 + steps_done;
                p.opt->total_norm = total_norm;
                if (p.norm_out != nullptr) *p.norm_out = total_norm;
            }
        }
        __syncthreads();
        {
            const optim_dev::AdamCoef ac = s_adam;
            for (long long i = (long long)c * ET + tid; i < p.n_params; i += (long long)G * ET) {
                float pi = __ldcg(p.params + i), mi = p.m[i], vi = p.v[i];
                optim_dev::adam_apply(ac, __ldcg(p.grad + i), pi, mi, vi);
                p.params[i] = pi;
                p.m[i] = mi;
                p.v[i] = vi;
            }
        }
        grid_bar(&p.ws->bar, bar_target);
    }
    if (dp && c == 0 && tid == 0) par_me->counter = par_k;
}

bool ep_supported(const sb200_mlp* net) {
    if (net == nullptr || net->n_layers != 3 || net->aux_layer >= 0) return false;
    if (net->act[0] != SB200_ACT_RELU || net->act[1] != SB200_ACT_RELU) return false;
    const int D = net->dims[0], H1 = net->dims[1], H2 = net->dims[2], NO = net->dims[3];
    if (D < 1 || H1 < 1 || H2 < 1 || NO < 1 || NO > EP_MAX_OUT) return false;
    return true;
}

}  // namespace

extern "C" int sb200_ppo_epochs_supported(const sb200_mlp* net) { return ep_supported(net) ? 1 : 0; }

extern "C" size_t sb200_ppo_epochs_workspace_bytes(void) { return sizeof(EpWs); }

extern "C" int sb200_ppo_epochs_f32(const sb200_epochs* a, void* stream) {
    SB200_REQUIRE(a != nullptr && a->net != nullptr);
    if (!ep_supported(a->net)) return SB200_ERR_UNSUPPORTED;
    const sb200_mlp* net = a->net;
    SB200_REQUIRE(a->params && a->n_params >= 1 && a->x && a->M >= 2 && a->ldx >= net->dims[0]);
    SB200_REQUIRE(a->x_in && a->h1 && a->h2 && a->out && a->d1 && a->d2 && a->dpre && a->slabs && a->grad && a->exp_avg && a->exp_avg_sq);
    SB200_REQUIRE(a->lr && a->opt_workspace && a->workspace && a->stats && a->splits >= 1 && a->epochs >= 1);
    SB200_REQUIRE(a->mode >= 0 && a->mode <= 2 && a->clip_mode >= 0 && a->clip_mode <= 2);
    SB200_REQUIRE(a->grid >= 1 && a->grid <= EP_MAX_G);
    SB200_REQUIRE((((uintptr_t)a->x_in) & 15) == 0 && (((uintptr_t)a->params) & 15) == 0);
    EpParams p;
    p.params = a->params;
    p.n_params = a->n_params;
    for (int l = 0; l < 3; ++l) {
        SB200_REQUIRE(net->W[l] >= a->params && net->b[l] >= a->params && net->W[l] < a->params + a->n_params);
        p.w_off[l] = (int)(net->W[l] - a->params);
        p.b_off[l] = (int)(net->b[l] - a->params);
        p.ldw[l] = net->ldw[l];
        SB200_REQUIRE(p.ldw[l] % 4 == 0 && p.w_off[l] % 4 == 0);
    }
    p.D = net->dims[0]; p.H1 = net->dims[1]; p.H2 = net->dims[2]; p.NO = net->dims[3];
    p.act_out = net->act[2];
    p.extra_off = a->extra_off;
    p.x = a->x; p.ldx = a->ldx; p.M = a->M;
    p.zf = a->zf_stats; p.zf_eps = (float)a->zf_eps;
    p.x_in = a->x_in; p.ld_x = (p.D + 3) / 4 * 4;
    p.h1 = a->h1; p.h2 = a->h2; p.out = a->out; p.d1 = a->d1; p.d2 = a->d2; p.dpre = a->dpre;
    p.ld_h1 = (p.H1 + 3) / 4 * 4; p.ld_h2 = (p.H2 + 3) / 4 * 4; p.ld_out = (p.NO + 3) / 4 * 4; p.ld_dpre = p.ld_out;
    p.slabs = a->slabs; p.grad = a->grad; p.m = a->exp_avg; p.v = a->exp_avg_sq;
    p.splits = a->splits;
    p.rows_per_split = ((a->M + a->splits - 1) / a->splits + gt::RK - 1) / gt::RK * gt::RK;
    p.lr = a->lr; p.weight_decay = a->weight_decay; p.clip_value = a->clip_value; p.clip_mode = a->clip_mode;
    p.opt = (OptWs*)a->opt_workspace;
    p.norm_out = a->norm_out;
    p.mode = a->mode;
    p.actions = a->actions; p.lda = a->lda; p.adv = a->adv; p.behave = a->behave_pd; p.ldb = a->ldb;
    p.ref = a->ref_pd; p.ldr = a->ldr; p.returns = a->returns;
    p.hyper = a->hyper; p.eta = a->eta; p.kl_target = a->kl_target; p.stop_threshold = a->stop_threshold;
    p.stats = a->stats; p.stop = a->stop_flag; p.epochs = a->epochs;
    p.ws = (EpWs*)a->workspace;
    p.shift = a->cta_shift;
    if (a->mode == 2) {
        SB200_REQUIRE(a->returns != nullptr && p.NO == 1);
    } else {
        SB200_REQUIRE(a->actions && a->adv && a->behave_pd && a->ref_pd && a->hyper);
        SB200_REQUIRE(a->lda >= p.NO && a->ldb >= 2 * p.NO && a->ldr >= 2 * p.NO);
        SB200_REQUIRE(a->extra_off >= 0 && a->extra_off + p.NO <= a->n_params);
    }
    for (int q = 0; q < PAR_MAX_WORLD; ++q) p.par.peers[q] = nullptr;
    p.par.world = 1; p.par.rank = 0; p.par.max_floats = 0;
    if (a->par != nullptr && a->par->world > 1) {
        SB200_REQUIRE(a->par->world <= PAR_MAX_WORLD && a->par->rank >= 0 && a->par->rank < a->par->world);
        SB200_REQUIRE(a->n_params <= a->par->max_floats);
        for (int q = 0; q < a->par->world; ++q) {
            SB200_REQUIRE(a->par->peers[q] != nullptr);
            p.par.peers[q] = a->par->peers[q];
        }
        p.par.world = a->par->world; p.par.rank = a->par->rank; p.par.max_floats = a->par->max_floats;
    }
    cudaStream_t st = (cudaStream_t)stream;
    SB200_CUDA(cudaMemsetAsync(&p.ws->bar, 0, sizeof(unsigned int), st));
    ppo_epochs_kernel<<<a->grid, ET, 0, st>>>(p);
    return sb200_launch_status();
}
