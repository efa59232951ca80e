// Small row-wise kernels around the networks: policy-distribution assembly, z-filter statistics.
#include "common.cuh"

namespace {

// pd[b] = [mean(A) | exp(log_var)(A) * exp(log_noise[b])]
// builders.py:127-129 (std = exp(log_var) broadcast, cat) + ppo_agent.py:139 (per-actor noise scaling).
__global__ void make_pd_kernel(const float* __restrict__ mean, long long ldm, const float* __restrict__ log_var,
                               const float* __restrict__ log_noise, int B, int A, float* __restrict__ pd,
                               long long ldp) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * A) return;
    const int b = (int)(i / A), j = (int)(i - (long long)b * A);
    float s = expf(log_var[j]);
    if (log_noise != nullptr) s = __fmul_rn(s, expf(log_noise[b]));
    pd[(long long)b * ldp + j] = mean[(long long)b * ldm + j];
    pd[(long long)b * ldp + A + j] = s;
}

// ZFilter.z_update (z_filter.py:44-57): stats = sum[D] | sumsq[D] | count[1]; one warp per column,
// lanes stride over rows, fp64 partials (fixed order -> deterministic).
__global__ void zfilter_update_kernel(const float* __restrict__ x, long long ldx, long long rows, int D,
                                      float* __restrict__ stats) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= D) return;
    double s = 0.0, q = 0.0;
    for (long long r = lane; r < rows; r += 32) {
        const float v = x[r * ldx + warp];
        s += (double)v;
        q += (double)__fmul_rn(v, v);
    }
    s = warp_sum(s);
    q = warp_sum(q);
    if (lane == 0) {
        stats[warp] = __fadd_rn(stats[warp], (float)s);
        stats[D + warp] = __fadd_rn(stats[D + warp], (float)q);
        if (warp == 0) stats[2 * D] = __fadd_rn(stats[2 * D], (float)rows);
    }
}

// RewardFilter (reward_filter.py:34-57) as used by ppo.py:452-456: out = filter.forward(r*scale) with the
// CURRENT statistics, THEN update: count += numel, running_sum += sum, running_sumsq = sum of squares
// (overwritten, not accumulated -- the reference's quirk, reward_filter.py:42).  stats = count|sum|sumsq.
__global__ void __launch_bounds__(1024) reward_filter_kernel(const float* __restrict__ r, long long n, float scale,
                                                              float eps, float* __restrict__ stats,
                                                              float* __restrict__ out) {
    __shared__ double sh[32];
    const float cnt = stats[0], sum = stats[1], sumsq = stats[2];
    const float mean = sum / cnt;
    const float sd = fmaxf(sqrtf(sumsq / cnt - mean * mean), eps);
    double s = 0.0, q = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const float x = __fmul_rn(r[i], scale);
        out[i] = fminf(fmaxf((x - mean) / sd, -5.0f), 5.0f);
        s += (double)x;
        q += (double)__fmul_rn(x, x);
    }
    s = block_sum(s, sh);
    q = block_sum(q, sh);
    if (threadIdx.x == 0) {
        stats[0] = __fadd_rn(cnt, (float)n);
        stats[1] = __fadd_rn(sum, (float)s);
        stats[2] = (float)q;
    }
}

// moments[0..1] = (sum, sum of squares) of x in fp64 -- the local half of a GLOBAL-batch advantage normalisation
__global__ void __launch_bounds__(1024) moments_kernel(const float* __restrict__ x, long long n,
                                                       double* __restrict__ moments) {
    __shared__ double sh[32];
    double s = 0.0, q = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const double v = (double)x[i];
        s += v;
        q += v * v;
    }
    s = block_sum(s, sh);
    q = block_sum(q, sh);
    if (threadIdx.x == 0) {
        moments[0] = s;
        moments[1] = q;
        moments[2] = (double)n;
    }
}

// x = (x - mean) / max(unbiased_std, floor) from (sum, sumsq, count) moments (ppo.py:413-416 over the global batch)
__global__ void normalize_kernel(float* __restrict__ x, long long n, const double* __restrict__ moments, float floor_) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double cnt = moments[2];
    const double mean = moments[0] / cnt;
    const double var = (moments[1] - moments[0] * moments[0] / cnt) / (cnt - 1.0);
    const float mean_f = (float)mean, std_f = (float)sqrt(var > 0.0 ? var : 0.0);
    x[i] = (x[i] - mean_f) / ((std_f > floor_) ? std_f : floor_);
}

__global__ void add_kernel(float* __restrict__ dst, const float* __restrict__ src, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = __fadd_rn(dst[i], src[i]);
}

}  // namespace

extern "C" int sb200_moments_f32(const float* x, int64_t n, double* moments3, void* stream) {
    SB200_REQUIRE(x && moments3 && n >= 1);
    moments_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(x, n, moments3);
    return sb200_launch_status();
}

extern "C" int sb200_normalize_f32(float* x, int64_t n, const double* moments3, double floor_value, void* stream) {
    SB200_REQUIRE(x && moments3 && n >= 1);
    normalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, n, moments3, (float)floor_value);
    return sb200_launch_status();
}

extern "C" int sb200_add_f32(float* dst, const float* src, int64_t n, void* stream) {
    SB200_REQUIRE(dst && src && n >= 1);
    add_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dst, src, n);
    return sb200_launch_status();
}

extern "C" int sb200_reward_filter_f32(const float* rewards, int64_t n, double reward_scale, double eps, float* stats,
                                       float* out, void* stream) {
    SB200_REQUIRE(rewards && stats && out && n >= 1);
    reward_filter_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(rewards, n, (float)reward_scale, (float)eps, stats, out);
    return sb200_launch_status();
}

extern "C" int sb200_make_pd_f32(const float* mean, int64_t ldm, const float* log_var, const float* log_noise, int B,
                                 int A, float* pd, int64_t ldp, void* stream) {
    SB200_REQUIRE(mean && log_var && pd && B >= 1 && A >= 1 && ldp >= 2 * A && ldm >= A);
    const long long n = (long long)B * A;
    make_pd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(mean, ldm, log_var, log_noise, B, A,
                                                                                pd, ldp);
    return sb200_launch_status();
}

extern "C" int sb200_zfilter_update_f32(const float* x, int64_t ldx, int64_t rows, int D, float* stats, void* stream) {
    SB200_REQUIRE(x && stats && rows >= 1 && D >= 1 && ldx >= D);
    const int warps_per_block = 4;
    zfilter_update_kernel<<<(D + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0,
                            (cudaStream_t)stream>>>(x, ldx, rows, D, stats);
    return sb200_launch_status();
}
