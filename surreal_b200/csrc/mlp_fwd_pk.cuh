// Inference forward for SMALL batches (one env step of all co-located actors: ~1K rows) on pre-packed weights.
//
// At this size the forward is a chain of dependent latencies, not a throughput problem; ncu on the two earlier
// variants (profiles/r01c_small_forward.md) showed 2 warps per scheduler each retiring one instruction every
// ~5 cycles, 70 % of them address arithmetic, operand splitting and shared-memory reads.  This kernel shortens
// the per-warp chain ~8x:
//   * weights are packed ONCE per parameter version (sb200_mlp_pack_tf32) into mma.m16n8k8 B-fragment order,
//     already split into the 3xTF32 pair (hi, lo): one coalesced 16-byte load per lane per (k-step, n-tile)
//     replaces 2 scalar loads + 6 conversion instructions;
//   * activations are kept in shared memory in A-fragment order, already split: one LDS.128 for hi, one for lo;
//   * a thread-block CLUSTER of 2 CTAs owns 16 rows; each CTA computes half of a layer's output columns and
//     stores them into BOTH CTAs' shared memory (distributed shared memory), so 1024 rows occupy 128 SMs with
//     no duplicated tensor work, one cluster barrier per layer.
// Numerics: the same 3xTF32 error-compensated products as mlp_fwd_mma.cuh (fp32-level accuracy); the main
// (hi*hi) and correction terms accumulate in separate registers to halve the dependent mma chain.
#pragma once
#include <cooperative_groups.h>

namespace {

namespace cg = cooperative_groups;

struct PkParams {
    FwdParams f;
    const float* P[SB200_MAX_LAYERS];   // packed weights of layer l (NULL for narrow layers)
    float* out;
    long long ld_out;
    int nst_max;                        // k-steps of the widest activation
    int ldp;                            // row stride of the plain (row-major) activation copy
};

constexpr int PK_RING = 4;              // k-steps of B fragments in flight per warp

__device__ __forceinline__ void mma_tf32u(float (&d)[4], const uint4& a, unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}

// position of element (row m of 16, column k) inside an A-fragment plane: [k/8][lane = (m%8)*4 + k%4][slot]
__device__ __forceinline__ int afrag_index(int m, int k) {
    const int kk = k & 7;
    return ((k >> 3) * 32 + (m & 7) * 4 + (kk & 3)) * 4 + ((m >> 3) & 1) + ((kk >> 2) << 1);
}

// Pack kernel: P[((s*NT + nt)*32 + lane)] = {hi(b0), hi(b1), lo(b0), lo(b1)},  b0 = W[8s + t][8nt + g],
// b1 = W[8s + t + 4][8nt + g]  (g = lane / 4, t = lane % 4), zero outside [K) x [N).
__global__ void __launch_bounds__(256) mlp_pack_tf32_kernel(const float* __restrict__ W, int K, int N, int ldw,
                                                            float4* __restrict__ P) {
    const int NT = (N + 7) >> 3, nst = (K + 7) >> 3;
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= nst * NT * 32) return;
    const int lane = idx & 31, nt = (idx >> 5) % NT, s = (idx >> 5) / NT;
    const int g = lane >> 2, t = lane & 3;
    const int n = nt * 8 + g, k0 = s * 8 + t, k1 = k0 + 4;
    const float b0 = (n < N && k0 < K) ? W[(long long)k0 * ldw + n] : 0.0f;
    const float b1 = (n < N && k1 < K) ? W[(long long)k1 * ldw + n] : 0.0f;
    unsigned h0, l0, h1, l1;
    split_tf32(b0, h0, l0);
    split_tf32(b1, h1, l1);
    P[idx] = make_float4(__uint_as_float(h0), __uint_as_float(h1), __uint_as_float(l0), __uint_as_float(l1));
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(SB200_THREADS, 2)
    mlp_fwd_pk_kernel(const __grid_constant__ PkParams pp) {
    const FwdParams& p = pp.f;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned crank = cluster.block_rank();
    extern __shared__ __align__(16) float smem[];
    const int plane = pp.nst_max * 128;                 // floats per A-fragment plane
    // [buf 0: hi | lo][buf 1: hi | lo][plain rows 16 x ldp][scratch]
    float* Hp = smem + 4 * plane;
    float* Wsc = Hp + 16 * pp.ldp;
    float* rsmem = cluster.map_shared_rank(smem, crank ^ 1u);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const long long row0 = (long long)(blockIdx.x >> 1) * 16;
    const int K0 = p.dims[0];

    // prefetch the narrow head's weights (used last) while the wide layers run
    int head_l = -1;
    for (int l = 0; l < p.n_layers; ++l)
        if (p.dims[l + 1] <= 32) head_l = l;
    const int zf_floats = (p.zf != nullptr) ? 2 * K0 : 0;
    float* Whead = Wsc + zf_floats;
    bool head_staged = false;
    if (head_l >= 0) {
        const int kl = (p.dims[head_l] + (p.aux_layer == head_l ? p.aux_dim : 0)) * p.ldw[head_l];
        if (zf_floats + kl <= p.scratch_floats) {
            for (int f = tid; f < kl / 4; f += SB200_THREADS) cp_async16(Whead + f * 4, p.W[head_l] + f * 4, 16);
            head_staged = true;
        }
        cp_async_commit();
    }
    if (p.zf != nullptr) {
        const float cnt = p.zf[2 * K0];
        for (int k = tid; k < K0; k += SB200_THREADS) {
            const float mean = p.zf[k] / cnt;
            const float var = p.zf[K0 + k] / cnt - mean * mean;
            Wsc[k] = mean;
            Wsc[K0 + k] = fmaxf(sqrtf(var), p.zf_eps);
        }
        __syncthreads();
    }
    int cur = 0;
    {   // input rows -> fragment order (both CTAs of the cluster stage all 16 rows)
        const int in_w = K0 + (p.aux_layer == 0 ? p.aux_dim : 0);
        const int in_wp = round_up(in_w, 8);
        const bool first_narrow = (p.dims[1] <= 32);
        float* Ahi = smem;
        float* Alo = smem + plane;
        for (int idx = tid; idx < 16 * in_wp; idx += SB200_THREADS) {
            const int m = idx / in_wp, k = idx - m * in_wp;
            const long long r = row0 + m;
            float v = 0.0f;
            if (r < p.rows) {
                if (k < K0) {
                    const float* src;
                    if (p.win_n > 0) {
                        const long long b = r / (p.win_n + 1);
                        const int kk = (int)(r - b * (p.win_n + 1));
                        src = (kk < p.win_n) ? p.x + (b * p.win_n + kk) * p.ldx : p.x_next + b * p.ldx;
                    } else {
                        src = p.x + r * p.ldx;
                    }
                    v = src[k];
                    if (p.zf != nullptr) v = fminf(fmaxf((v - Wsc[k]) / Wsc[K0 + k], -5.0f), 5.0f);
                } else if (k < in_w) {
                    v = p.aux[r * p.aux_ld + (k - K0)];
                }
            }
            if (first_narrow) {
                Hp[m * pp.ldp + k] = v;
            } else {
                unsigned hi, lo;
                split_tf32(v, hi, lo);
                const int a = afrag_index(m, k);
                Ahi[a] = __uint_as_float(hi);
                Alo[a] = __uint_as_float(lo);
            }
        }
    }
    cluster.sync();                                    // also: the peer CTA has started (its smem is addressable)

    for (int l = 0; l < p.n_layers; ++l) {
        const int K = p.dims[l] + (p.aux_layer == l ? p.aux_dim : 0);
        const int N = p.dims[l + 1];
        const float* __restrict__ bias = p.b[l];
        const int act = p.act[l];
        const bool last = (l == p.n_layers - 1);
        if (N > 32) {
            const int nst = (K + 7) >> 3;
            const int NT = (N + 7) >> 3, half = (NT + 1) >> 1;
            const int nt_lo = (int)crank * half, nt_hi = min(NT, nt_lo + half);
            const uint4* Ahi = reinterpret_cast<const uint4*>(smem + cur * 2 * plane);
            const uint4* Alo = reinterpret_cast<const uint4*>(smem + cur * 2 * plane + plane);
            float* Ohi = smem + (cur ^ 1) * 2 * plane;
            float* Olo = Ohi + plane;
            const bool next_narrow = !last && (p.dims[l + 2] <= 32);
            const int auxd = (!last && p.aux_layer == l + 1) ? p.aux_dim : 0;
            const long long roff = rsmem - smem;       // the same offset addresses the peer CTA's copy
            for (int pass0 = nt_lo; pass0 < nt_hi; pass0 += 16) {
                const int nt0 = pass0 + warp * 2;
                if (nt0 >= nt_hi) continue;            // warp-uniform; no block barrier inside the pass loop
                const bool two = (nt0 + 1 < nt_hi);
                float accM[2][4], accC[2][4];
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int c = 0; c < 4; ++c) accM[j][c] = accC[j][c] = 0.0f;
                const float4* bp = reinterpret_cast<const float4*>(pp.P[l]) + ((long long)nt0 * 32 + lane);
                const long long sstride = (long long)NT * 32;
                float4 ring[PK_RING][2];
                auto load_stage = [&](float4 (&dst)[2], int s) {
                    if (s < nst) {
                        dst[0] = __ldg(bp + s * sstride);
                        dst[1] = two ? __ldg(bp + s * sstride + 32) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                };
#pragma unroll
                for (int d = 0; d < PK_RING - 1; ++d) load_stage(ring[d], d);
                for (int s0 = 0; s0 < nst; s0 += PK_RING) {
#pragma unroll
                    for (int d = 0; d < PK_RING; ++d) {
                        const int s = s0 + d;
                        load_stage(ring[(d + PK_RING - 1) % PK_RING], s + PK_RING - 1);
                        if (s < nst) {
                            const uint4 ah = Ahi[s * 32 + lane];
                            const uint4 al = Alo[s * 32 + lane];
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                const float4 b = ring[d][j];
                                mma_tf32u(accC[j], al, __float_as_uint(b.x), __float_as_uint(b.y));
                                mma_tf32u(accC[j], ah, __float_as_uint(b.z), __float_as_uint(b.w));
                                mma_tf32u(accM[j], ah, __float_as_uint(b.x), __float_as_uint(b.y));
                            }
                        }
                    }
                }
                // epilogue: c0,c1 -> (row g, cols 2t, 2t+1); c2,c3 -> (row g+8, same cols)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    if (j == 1 && !two) break;
                    const int nt = nt0 + j;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int m = g + 8 * (c >> 1);
                        const int n = nt * 8 + 2 * t + (c & 1);
                        const long long r = row0 + m;
                        float v = 0.0f;
                        if (n < N) v = apply_act(accM[j][c] + accC[j][c] + bias[n], act);
                        else if (n < N + auxd && r < p.rows) v = p.aux[r * p.aux_ld + (n - N)];
                        if (last) {
                            if (n < N && r < p.rows) pp.out[r * pp.ld_out + n] = v;
                        } else if (next_narrow) {
                            float* q = Hp + m * pp.ldp + n;
                            *q = v;
                            *(q + roff) = v;
                        } else {
                            unsigned hi, lo;
                            split_tf32(v, hi, lo);
                            const int a = afrag_index(m, n);
                            Ohi[a] = __uint_as_float(hi);
                            Olo[a] = __uint_as_float(lo);
                            *(Ohi + a + roff) = __uint_as_float(hi);
                            *(Olo + a + roff) = __uint_as_float(lo);
                        }
                    }
                }
            }
            if (!last) {
                // columns past this layer's last n-tile (aux concatenation / padding of the next layer's K): local copy
                const int kp_next = round_up(N + auxd, 8);
                const int span = kp_next - NT * 8;
                for (int idx = tid; idx < 16 * span; idx += SB200_THREADS) {
                    const int m = idx / span, k = NT * 8 + (idx - m * span);
                    const long long r = row0 + m;
                    float v = 0.0f;
                    if (k < N + auxd && r < p.rows) v = p.aux[r * p.aux_ld + (k - N)];
                    if (next_narrow) {
                        Hp[m * pp.ldp + k] = v;
                    } else {
                        unsigned hi, lo;
                        split_tf32(v, hi, lo);
                        const int a = afrag_index(m, k);
                        Ohi[a] = __uint_as_float(hi);
                        Olo[a] = __uint_as_float(lo);
                    }
                }
                cluster.sync();
                cur ^= 1;
            }
        } else {
            // narrow head (<= 32 outputs): fp32 dot products on the plain activation copy; CTA c owns rows 8c..8c+7
            cp_async_wait<0>();
            __syncthreads();
            const float* Wn = head_staged ? Whead : p.W[l];
            const int ldw = p.ldw[l];
            const int m = (int)crank * 8 + warp;
            const long long r = row0 + m;
            const float* hrow = Hp + m * pp.ldp;
            for (int n8 = 0; n8 < N; n8 += 8) {
                float s8[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) s8[j] = 0.0f;
                const bool second = (n8 + 4 < ldw);
                for (int k = lane; k < K; k += 32) {
                    const float hv = hrow[k];
                    const float* wr = Wn + (long long)k * ldw + n8;
                    const float4 w0 = *reinterpret_cast<const float4*>(wr);
                    s8[0] = fmaf(hv, w0.x, s8[0]); s8[1] = fmaf(hv, w0.y, s8[1]);
                    s8[2] = fmaf(hv, w0.z, s8[2]); s8[3] = fmaf(hv, w0.w, s8[3]);
                    if (second) {
                        const float4 w1 = *reinterpret_cast<const float4*>(wr + 4);
                        s8[4] = fmaf(hv, w1.x, s8[4]); s8[5] = fmaf(hv, w1.y, s8[5]);
                        s8[6] = fmaf(hv, w1.z, s8[6]); s8[7] = fmaf(hv, w1.w, s8[7]);
                    }
                }
                float mine = 0.0f;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float tt = warp_sum(s8[j]);
                    if (lane == j) mine = tt;
                }
                const int n = n8 + lane;
                if (lane < 8 && n < N && r < p.rows) pp.out[r * pp.ld_out + n] = apply_act(mine + bias[n], act);
            }
        }
    }
    cluster.sync();                                    // no CTA may exit while its peer can still write into it
}

}  // namespace
