// 64x64 fp32 SIMT GEMM tiles as DEVICE functions (256 threads, 4x4 per thread, reduction chunk 16, register prefetch of
// the next chunk) -- the bodies of mlp_bwd.cu's kernels plus a forward tile, callable with an explicit tile index so
// that a persistent kernel can schedule them itself (the first-generation learner kernel did; epoch2.cu borrows the helpers).
//
//   gt_tile_nn : C[m][n]  = act(zf(X)[m][:] . W[:][n] + b[n])                  forward          (reduce over K)
//   gt_tile_nt : dX[m][k] = (dY[m][:] . W[k][:]) * relu'(Xact[m][k])           input gradient   (reduce over N)
//   gt_tile_tn : dW[k][n] = sum_{m in [mb, me)} zf(X)[m][k] * dY[m][n], db[n]  weight gradient  (reduce over M)
//
// Operands are read with ld.global.cg (L2 is the point of coherence): inside the persistent kernel the same buffers are
// rewritten by other CTAs between grid barriers, so neither L1 nor the non-coherent path may serve them.
#pragma once
#include "common.cuh"

namespace gt {

constexpr int TB = 64;     // tile edge
constexpr int RK = 16;     // reduction chunk
constexpr int LDT = TB + 4;

struct Smem {
    float As[RK][LDT];
    float Bs[RK][LDT];
};

__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

// up to 4 floats of a row that ends at `end` (exclusive index relative to p), zero padded
__device__ __forceinline__ float4 ldcg4_guard(const float* p, int remaining) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (remaining >= 4) return ldcg4(p);
    if (remaining > 0) v.x = __ldcg(p);
    if (remaining > 1) v.y = __ldcg(p + 1);
    if (remaining > 2) v.z = __ldcg(p + 2);
    return v;
}

__device__ __forceinline__ float zf1(float v, float m, float s) { return fminf(fmaxf((v - m) / s, -5.0f), 5.0f); }

__device__ __forceinline__ void mac_chunk(const Smem& s, int ty, int tx, float (&acc)[4][4]) {
#pragma unroll
    for (int r = 0; r < RK; ++r) {
        const float4 av = *reinterpret_cast<const float4*>(&s.As[r][ty * 4]);
        const float4 bv = *reinterpret_cast<const float4*>(&s.Bs[r][tx * 4]);
        const float aa[4] = {av.x, av.y, av.z, av.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
}

// ---- forward tile.  X [M][ldx] (K valid columns; z-filtered on load when zm != nullptr: zm / zs = per-column mean and
// clamped std), W [K][ldw] (k-major), C [M][ldc].  act: SB200_ACT_*.
__device__ __forceinline__ void gt_tile_nn(Smem& s, int m0, int n0, const float* X, long long ldx, const float* zm,
                                           const float* zs, const float* W, int ldw, const float* bias, int act, float* C,
                                           long long ldc, int M, int N, int K) {
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int ar = tid >> 2, ac = (tid & 3) * 4;       // A loader: tile row 0..63, k offset 0,4,8,12  (transposing store)
    const int br = tid >> 4, bc = (tid & 15) * 4;      // B loader: k row 0..15, column offset 0..60
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
    auto fetch = [&](int k0, float4& a, float4& b) {
        a = make_float4(0.f, 0.f, 0.f, 0.f);
        b = make_float4(0.f, 0.f, 0.f, 0.f);
        const int k = k0 + ac;
        if (m0 + ar < M && k < K) {
            a = ldcg4_guard(X + (long long)(m0 + ar) * ldx + k, K - k);
            if (zm != nullptr) {
                a.x = zf1(a.x, zm[k], zs[k]);
                if (k + 1 < K) a.y = zf1(a.y, zm[k + 1], zs[k + 1]);
                if (k + 2 < K) a.z = zf1(a.z, zm[k + 2], zs[k + 2]);
                if (k + 3 < K) a.w = zf1(a.w, zm[k + 3], zs[k + 3]);
            }
        }
        const int kb = k0 + br, n = n0 + bc;
        if (kb < K && n < ldw) b = ldcg4(W + (long long)kb * ldw + n);       // padding columns of W are zero
    };
    float4 a, b;
    fetch(0, a, b);
    for (int k0 = 0; k0 < K; k0 += RK) {
        s.As[ac + 0][ar] = a.x; s.As[ac + 1][ar] = a.y; s.As[ac + 2][ar] = a.z; s.As[ac + 3][ar] = a.w;
        *reinterpret_cast<float4*>(&s.Bs[br][bc]) = b;
        __syncthreads();
        if (k0 + RK < K) fetch(k0 + RK, a, b);
        mac_chunk(s, ty, tx, acc);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j] + __ldcg(bias + n);
            if (act == SB200_ACT_RELU) v = fmaxf(v, 0.0f);
            else if (act == SB200_ACT_TANH) v = tanhf(v);
            C[(long long)m * ldc + n] = v;
        }
    }
}

// ---- input-gradient tile: dX[m][k] = sum_n dY[m][n] W[k][n]; masked by Xact[m][k] > 0 when Xact != nullptr.
__device__ __forceinline__ void gt_tile_nt(Smem& s, int m0, int k0, const float* dY, long long ldy, const float* W, int ldw,
                                           const float* Xact, long long ldxa, float* dX, long long lddx, int M, int N,
                                           int K) {
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int lr = tid >> 2, lc = (tid & 3) * 4;     // loader: row 0..63, reduction offset 0,4,8,12
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
    auto fetch = [&](int n0, float4& a, float4& b) {
        a = make_float4(0.f, 0.f, 0.f, 0.f);
        b = make_float4(0.f, 0.f, 0.f, 0.f);
        const int n = n0 + lc;
        if (m0 + lr < M && n < N) a = ldcg4_guard(dY + (long long)(m0 + lr) * ldy + n, N - n);
        if (k0 + lr < K && n < ldw) b = ldcg4(W + (long long)(k0 + lr) * ldw + n);
    };
    float4 a, b;
    fetch(0, a, b);
    for (int n0 = 0; n0 < N; n0 += RK) {
        s.As[lc + 0][lr] = a.x; s.As[lc + 1][lr] = a.y; s.As[lc + 2][lr] = a.z; s.As[lc + 3][lr] = a.w;
        s.Bs[lc + 0][lr] = b.x; s.Bs[lc + 1][lr] = b.y; s.Bs[lc + 2][lr] = b.z; s.Bs[lc + 3][lr] = b.w;
        __syncthreads();
        if (n0 + RK < N) fetch(n0 + RK, a, b);
        mac_chunk(s, ty, tx, acc);
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
            if (Xact != nullptr && !(__ldcg(Xact + (long long)m * ldxa + k) > 0.0f)) v = 0.0f;
            dX[(long long)m * lddx + k] = v;
        }
    }
}

// ---- weight-gradient tile over the row range [mb, me): dW[k][n] (ldw), db[n] (written by the k0 == 0 tile when db != nullptr).
__device__ __forceinline__ void gt_tile_tn(Smem& s, int k0, int n0, int mb, int me, const float* X, long long ldx,
                                           const float* zm, const float* zs, const float* dY, long long ldy, float* dW,
                                           float* db, int ldw, int K, int N) {
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int lr = tid >> 4, lc = (tid & 15) * 4;    // loader: reduction row 0..15, column offset 0..60
    float acc[4][4];
    float colsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
    const bool do_bias = (k0 == 0) && (ty == 0) && (db != nullptr);
    auto fetch = [&](int mm, float4& a, float4& b) {
        a = make_float4(0.f, 0.f, 0.f, 0.f);
        b = make_float4(0.f, 0.f, 0.f, 0.f);
        const int m = mm + lr;
        if (m < me) {
            const int k = k0 + lc;
            if (k < K) {
                a = ldcg4_guard(X + (long long)m * ldx + k, K - k);
                if (zm != nullptr) {
                    a.x = zf1(a.x, zm[k], zs[k]);
                    if (k + 1 < K) a.y = zf1(a.y, zm[k + 1], zs[k + 1]);
                    if (k + 2 < K) a.z = zf1(a.z, zm[k + 2], zs[k + 2]);
                    if (k + 3 < K) a.w = zf1(a.w, zm[k + 3], zs[k + 3]);
                }
            }
            const int n = n0 + lc;
            if (n < N) b = ldcg4_guard(dY + (long long)m * ldy + n, N - n);
        }
    };
    float4 a, b;
    fetch(mb, a, b);
    for (int mm = mb; mm < me; mm += RK) {
        *reinterpret_cast<float4*>(&s.As[lr][lc]) = a;
        *reinterpret_cast<float4*>(&s.Bs[lr][lc]) = b;
        __syncthreads();
        if (mm + RK < me) fetch(mm + RK, a, b);
        mac_chunk(s, ty, tx, acc);
        if (do_bias) {
#pragma unroll
            for (int r = 0; r < RK; ++r) {
                const float4 bv = *reinterpret_cast<const float4*>(&s.Bs[r][tx * 4]);
                colsum[0] += bv.x; colsum[1] += bv.y; colsum[2] += bv.z; colsum[3] += bv.w;
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int k = k0 + ty * 4 + i;
        if (k >= K) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < N) dW[(long long)k * ldw + n] = acc[i][j];
        }
    }
    if (do_bias) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < N) db[n] = colsum[j];
        }
    }
}

}  // namespace gt
