// Backward GEMMs of a Linear layer in the kernel layout W[K][ldw] (fp32 SIMT, 64x64 tiles, 4x4 per thread).
//
//   dX[M,K]  = (dY[M,N] . W^T) * relu'(Xact)        -- sb200_linear_bwd_dx_f32   ("NT": reduce over N)
//   dW[K,N] += X[M,K]^T . dY[M,N],  db[N] = colsum  -- sb200_linear_bwd_dw_f32   ("TN": reduce over M)
//
// dW is produced as `splits` partial slabs (one per M-range, fixed order) that the optimiser's
// reduce kernel sums deterministically -- no atomics, bit-reproducible run to run.
// Replaces torch autograd's addmm backward for builders.py's Linear layers (ppo.py:242,347; ddpg.py:306,330).
#include "common.cuh"

namespace {

constexpr int TB = 64;     // tile edge
constexpr int RK = 16;     // reduction chunk
constexpr int LDT = TB + 4;

// ------------------------------------------------------------------------------------------------
// dX tile: As[r][m] (dY transposed), Bs[r][k] (W transposed); reduce r over N.
__global__ void __launch_bounds__(SB200_THREADS) bwd_dx_kernel(const float* __restrict__ dY, long long ldy,
                                                                const float* __restrict__ W, int ldw,
                                                                const float* __restrict__ Xact, long long ldxa,
                                                                float* __restrict__ dX, long long lddx, int M, int N,
                                                                int K) {
    __shared__ __align__(16) float As[RK][LDT];
    __shared__ __align__(16) float Bs[RK][LDT];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * TB, k0 = blockIdx.x * TB;
    const int lr = tid >> 2, lc = (tid & 3) * 4;     // loader: row 0..63, reduction offset 0,4,8,12
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

    // global -> registers for chunk c+1 overlaps the FFMA work on chunk c (latency-bound at these sizes)
    auto fetch = [&](int n0, float4& a, float4& b) {
        a = make_float4(0.f, 0.f, 0.f, 0.f);
        b = make_float4(0.f, 0.f, 0.f, 0.f);
        const int n = n0 + lc;
        if (m0 + lr < M && n < N) {
            const float* src = dY + (long long)(m0 + lr) * ldy + n;
            if (n + 3 < N) {
                a = *reinterpret_cast<const float4*>(src);
            } else {
                a.x = src[0];
                if (n + 1 < N) a.y = src[1];
                if (n + 2 < N) a.z = src[2];
            }
        }
        if (k0 + lr < K && n < ldw) b = *reinterpret_cast<const float4*>(W + (long long)(k0 + lr) * ldw + n);
    };
    float4 a, b;
    fetch(0, a, b);
    for (int n0 = 0; n0 < N; n0 += RK) {
        As[lc + 0][lr] = a.x; As[lc + 1][lr] = a.y; As[lc + 2][lr] = a.z; As[lc + 3][lr] = a.w;
        Bs[lc + 0][lr] = b.x; Bs[lc + 1][lr] = b.y; Bs[lc + 2][lr] = b.z; Bs[lc + 3][lr] = b.w;
        __syncthreads();
        if (n0 + RK < N) fetch(n0 + RK, a, b);
#pragma unroll
        for (int r = 0; r < RK; ++r) {
            const float4 av = *reinterpret_cast<const float4*>(&As[r][ty * 4]);
            const float4 bv = *reinterpret_cast<const float4*>(&Bs[r][tx * 4]);
            const float aa[4] = {av.x, av.y, av.z, av.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + tx * 4 + j;
            if (k >= K) continue;
            float v = acc[i][j];
            if (Xact != nullptr && !(Xact[(long long)m * ldxa + k] > 0.0f)) v = 0.0f;
            dX[(long long)m * lddx + k] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// dW tile: As[r][k] = X rows, Bs[r][n] = dY rows; reduce r over this split's M-range.
__global__ void __launch_bounds__(SB200_THREADS) bwd_dw_kernel(const float* __restrict__ X, long long ldx,
                                                                const float* __restrict__ dY, long long ldy,
                                                                float* __restrict__ dW_slabs, float* __restrict__ db_slabs,
                                                                long long slab_stride, int ldw, int M, int K, int N,
                                                                int rows_per_split) {
    __shared__ __align__(16) float As[RK][LDT];
    __shared__ __align__(16) float Bs[RK][LDT];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int k0 = blockIdx.y * TB, n0 = blockIdx.x * TB;
    const int z = blockIdx.z;
    const int mb = z * rows_per_split;
    const int me = min(M, mb + rows_per_split);
    const int lr = tid >> 4, lc = (tid & 15) * 4;    // loader: reduction row 0..15, column offset 0..60
    float acc[4][4];
    float colsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

    auto fetch = [&](int mm, float4& a, float4& b) {
        a = make_float4(0.f, 0.f, 0.f, 0.f);
        b = make_float4(0.f, 0.f, 0.f, 0.f);
        const int m = mm + lr;
        if (m < me) {
            const int k = k0 + lc;
            if (k < K) {
                const float* src = X + (long long)m * ldx + k;
                if (k + 3 < K) {
                    a = *reinterpret_cast<const float4*>(src);
                } else {
                    a.x = src[0];
                    if (k + 1 < K) a.y = src[1];
                    if (k + 2 < K) a.z = src[2];
                }
            }
            const int n = n0 + lc;
            if (n < N) {
                const float* src = dY + (long long)m * ldy + n;
                if (n + 3 < N) {
                    b = *reinterpret_cast<const float4*>(src);
                } else {
                    b.x = src[0];
                    if (n + 1 < N) b.y = src[1];
                    if (n + 2 < N) b.z = src[2];
                }
            }
        }
    };
    float4 a, b;
    fetch(mb, a, b);
    for (int mm = mb; mm < me; mm += RK) {
        *reinterpret_cast<float4*>(&As[lr][lc]) = a;
        *reinterpret_cast<float4*>(&Bs[lr][lc]) = b;
        __syncthreads();
        if (mm + RK < me) fetch(mm + RK, a, b);
#pragma unroll
        for (int r = 0; r < RK; ++r) {
            const float4 av = *reinterpret_cast<const float4*>(&As[r][ty * 4]);
            const float4 bv = *reinterpret_cast<const float4*>(&Bs[r][tx * 4]);
            const float aa[4] = {av.x, av.y, av.z, av.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
            if (blockIdx.y == 0 && ty == 0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) colsum[j] += bb[j];
            }
        }
        __syncthreads();
    }
    float* dWz = dW_slabs + (long long)z * slab_stride;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int k = k0 + ty * 4 + i;
        if (k >= K) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < N) dWz[(long long)k * ldw + n] = acc[i][j];
        }
    }
    if (blockIdx.y == 0 && ty == 0 && db_slabs != nullptr) {
        float* dbz = db_slabs + (long long)z * slab_stride;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < N) dbz[n] = colsum[j];
        }
    }
}

}  // namespace

extern "C" int sb200_linear_bwd_dx_f32(const float* dY, int64_t ldy, const float* W, int ldw, const float* x_act,
                                       int64_t ld_xact, float* dX, int64_t lddx, int M, int N, int K, void* stream) {
    SB200_REQUIRE(dY && W && dX && M >= 1 && N >= 1 && K >= 1);
    SB200_REQUIRE(ldw % 4 == 0 && ldw >= N && ldy >= N && lddx >= K);
    SB200_REQUIRE(ldy % 4 == 0 && ((((uintptr_t)dY) | ((uintptr_t)W)) & 15) == 0);
    dim3 grid((K + TB - 1) / TB, (M + TB - 1) / TB);
    bwd_dx_kernel<<<grid, SB200_THREADS, 0, (cudaStream_t)stream>>>(dY, ldy, W, ldw, x_act, ld_xact, dX, lddx, M, N, K);
    return sb200_launch_status();
}

extern "C" int sb200_linear_bwd_dw_f32(const float* X, int64_t ldx, const float* dY, int64_t ldy, float* dW_slabs,
                                       float* db_slabs, int64_t slab_stride, int splits, int ldw, int M, int K, int N,
                                       void* stream) {
    SB200_REQUIRE(X && dY && dW_slabs && M >= 1 && N >= 1 && K >= 1 && splits >= 1);
    SB200_REQUIRE(ldw % 4 == 0 && ldw >= N && ldy >= N && ldx >= K);
    SB200_REQUIRE(ldx % 4 == 0 && ldy % 4 == 0 && ((((uintptr_t)X) | ((uintptr_t)dY)) & 15) == 0);
    const int rps = ((M + splits - 1) / splits + RK - 1) / RK * RK;
    dim3 grid((N + TB - 1) / TB, (K + TB - 1) / TB, splits);
    bwd_dw_kernel<<<grid, SB200_THREADS, 0, (cudaStream_t)stream>>>(X, ldx, dY, ldy, dW_slabs, db_slabs, slab_stride,
                                                                    ldw, M, K, N, rps);
    return sb200_launch_status();
}
