// Optimiser step over ONE flat parameter buffer: deterministic slab reduction + global grad norm,
// then clip (by global norm, ppo.py:244-246,349-351; or by value, ddpg.py:309,332) + Adam.
//
// Replaces nn.utils.clip_grad_norm_ / clip_grad_value_ and torch.optim.Adam.step() (default betas/eps,
// L2 weight decay; surreal/learner/ppo.py:159-168, ddpg.py:145-165).  The arithmetic follows torch's
// single-tensor Adam (lerp_ / mul_.addcmul_ / sqrt./.add_ / addcdiv_) operation by operation so that
// fp32 results agree with the reference to rounding.  The step counter lives on the device so that the
// whole update can be replayed from a CUDA graph.
#include "common.cuh"

namespace {

constexpr int OT = 256;
constexpr int MAX_BLOCKS = 1024;

struct OptWs {
    unsigned int counter;
    int step;
    float total_norm;
    float pad;
    double partial[MAX_BLOCKS];
};

__global__ void __launch_bounds__(OT) grad_reduce_norm_kernel(const float* __restrict__ slabs, long long slab_stride,
                                                              int splits, float* __restrict__ grad, long long n,
                                                              float scale, int bump_step, OptWs* ws,
                                                              const int* __restrict__ stop) {
    if (stop != nullptr && *stop) return;
    __shared__ double sh[32];
    double sq = 0.0;
    for (long long i = (long long)blockIdx.x * OT + threadIdx.x; i < n; i += (long long)gridDim.x * OT) {
        float g = slabs[i];
        for (int z = 1; z < splits; ++z) g += slabs[(long long)z * slab_stride + i];   // fixed order
        if (scale != 1.0f) g = __fmul_rn(g, scale);
        grad[i] = g;
        sq += (double)g * (double)g;
    }
    const double t = block_sum(sq, sh);
    if (threadIdx.x == 0) ws->partial[blockIdx.x] = t;
    if (last_block_ticket(&ws->counter, gridDim.x)) {
        if (threadIdx.x == 0) {
            double acc = 0.0;
            for (unsigned int k = 0; k < gridDim.x; ++k) acc += ws->partial[k];
            ws->total_norm = (float)sqrt(acc);
            if (bump_step) ws->step += 1;
        }
    }
}

__global__ void __launch_bounds__(OT) clip_adam_kernel(float* __restrict__ p, const float* __restrict__ grad,
                                                       float* __restrict__ m, float* __restrict__ v, long long n,
                                                       const double* __restrict__ lr_ptr, double beta1, double beta2,
                                                       double eps, double weight_decay, int clip_mode,
                                                       double clip_value, const OptWs* __restrict__ ws,
                                                       float* __restrict__ norm_out, const int* __restrict__ stop) {
    if (stop != nullptr && *stop) return;
    const int t = ws->step;
    const double bc1 = 1.0 - pow(beta1, (double)t);
    const double bc2 = 1.0 - pow(beta2, (double)t);
    const float neg_step = (float)(-(lr_ptr[0] / bc1));
    const float bc2_sqrt = (float)sqrt(bc2);
    const float w1 = (float)(1.0 - beta1), b2 = (float)beta2, a2 = (float)(1.0 - beta2), epsf = (float)eps;
    float coef = 1.0f;
    if (clip_mode == 1) {
        const float c = (float)clip_value / (ws->total_norm + 1e-6f);      // clip_grad_norm_
        coef = fminf(c, 1.0f);
    }
    if (norm_out != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *norm_out = ws->total_norm;
    const float cv = (float)clip_value, wd = (float)weight_decay;
    for (long long i = (long long)blockIdx.x * OT + threadIdx.x; i < n; i += (long long)gridDim.x * OT) {
        float g = grad[i];
        if (clip_mode == 1) g = __fmul_rn(g, coef);
        else if (clip_mode == 2) g = fminf(fmaxf(g, -cv), cv);
        float pi = p[i];
        if (wd != 0.0f) g = __fadd_rn(g, __fmul_rn(wd, pi));
        float mi = m[i], vi = v[i];
        mi = fmaf(w1, __fsub_rn(g, mi), mi);                                  // exp_avg.lerp_(grad, 1-beta1)
        vi = __fadd_rn(__fmul_rn(vi, b2), __fmul_rn(__fmul_rn(a2, g), g));   // mul_(beta2).addcmul_(g, g, 1-beta2)
        const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vi), bc2_sqrt), epsf);
        pi = __fadd_rn(pi, __fdiv_rn(__fmul_rn(neg_step, mi), denom));      // addcdiv_(exp_avg, denom, -step_size)
        p[i] = pi;
        m[i] = mi;
        v[i] = vi;
    }
}

__global__ void __launch_bounds__(OT) soft_update_kernel(float* __restrict__ target, const float* __restrict__ src,
                                                         long long n, float tau, const int* __restrict__ stop) {
    if (stop != nullptr && *stop) return;
    const float keep = 1.0f - tau;
    for (long long i = (long long)blockIdx.x * OT + threadIdx.x; i < n; i += (long long)gridDim.x * OT)
        target[i] = __fadd_rn(__fmul_rn(target[i], keep), __fmul_rn(tau, src[i]));
}

inline int grid_for(long long n) {
    long long b = (n + 4 * OT - 1) / (4 * OT);   // >= 4 elements per thread: small nets finish in one short wave
    if (b > 296) b = 296;          // 2 x 148 SMs, grid-stride
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace

extern "C" size_t sb200_optim_workspace_bytes(void) { return sizeof(OptWs); }

extern "C" int sb200_grad_reduce_norm_f32(const float* slabs, int64_t slab_stride, int splits, float* grad, int64_t n,
                                          double scale, int bump_step, void* workspace, const int* stop_flag,
                                          void* stream) {
    SB200_REQUIRE(slabs && grad && workspace && splits >= 1 && n >= 1);
    grad_reduce_norm_kernel<<<grid_for(n), OT, 0, (cudaStream_t)stream>>>(slabs, slab_stride, splits, grad, n,
                                                                          (float)scale, bump_step, (OptWs*)workspace,
                                                                          stop_flag);
    return sb200_launch_status();
}

extern "C" int sb200_clip_adam_f32(float* params, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                                   const double* lr, double beta1, double beta2, double eps, double weight_decay,
                                   int clip_mode, double clip_value, void* workspace, float* norm_out,
                                   const int* stop_flag, void* stream) {
    SB200_REQUIRE(params && grad && exp_avg && exp_avg_sq && lr && workspace && n >= 1);
    SB200_REQUIRE(clip_mode >= 0 && clip_mode <= 2);
    clip_adam_kernel<<<grid_for(n), OT, 0, (cudaStream_t)stream>>>(params, grad, exp_avg, exp_avg_sq, n, lr, beta1,
                                                                   beta2, eps, weight_decay, clip_mode, clip_value,
                                                                   (const OptWs*)workspace, norm_out, stop_flag);
    return sb200_launch_status();
}

extern "C" int sb200_soft_update_f32(float* target, const float* src, int64_t n, double tau, const int* stop_flag,
                                     void* stream) {
    SB200_REQUIRE(target && src && n >= 1);
    soft_update_kernel<<<grid_for(n), OT, 0, (cudaStream_t)stream>>>(target, src, n, (float)tau, stop_flag);
    return sb200_launch_status();
}
