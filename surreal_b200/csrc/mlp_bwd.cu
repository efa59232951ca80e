// Backward GEMMs of a Linear layer in the kernel layout W[K][ldw] (fp32 SIMT, 64x64 tiles, 4x4 per thread).
//
//   dX[M,K]  = (dY[M,N] . W^T) * relu'(Xact)        -- sb200_linear_bwd_dx_f32   ("NT": reduce over N)
//   dW[K,N] += X[M,K]^T . dY[M,N],  db[N] = colsum  -- sb200_linear_bwd_dw_f32   ("TN": reduce over M)
//
// dW is produced as `splits` partial slabs (one per M-range, fixed order) that the optimiser's
// reduce kernel sums deterministically -- no atomics, bit-reproducible run to run.
// Replaces torch autograd's addmm backward for builders.py's Linear layers (ppo.py:242,347; ddpg.py:306,330).
#include "common.cuh"
#include "gemm_tiles.cuh"

namespace {

using gt::TB;
using gt::RK;

// The tile bodies live in gemm_tiles.cuh.
__global__ void __launch_bounds__(SB200_THREADS) bwd_dx_kernel(const float* dY, long long ldy, const float* W, int ldw,
                                                                const float* Xact, long long ldxa, float* dX,
                                                                long long lddx, int M, int N, int K) {
    __shared__ __align__(16) gt::Smem s;
    gt::gt_tile_nt(s, blockIdx.y * TB, blockIdx.x * TB, dY, ldy, W, ldw, Xact, ldxa, dX, lddx, M, N, K);
}

__global__ void __launch_bounds__(SB200_THREADS) bwd_dw_kernel(const float* X, long long ldx, const float* dY, long long ldy,
                                                                float* dW_slabs, float* db_slabs, long long slab_stride,
                                                                int ldw, int M, int K, int N, int rows_per_split) {
    __shared__ __align__(16) gt::Smem s;
    const int z = blockIdx.z;
    const int mb = z * rows_per_split;
    const int me = min(M, mb + rows_per_split);
    gt::gt_tile_tn(s, blockIdx.y * TB, blockIdx.x * TB, mb, me, X, ldx, nullptr, nullptr, dY, ldy,
                   dW_slabs + (long long)z * slab_stride, db_slabs != nullptr ? db_slabs + (long long)z * slab_stride : nullptr,
                   ldw, K, N);
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
