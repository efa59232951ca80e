// Stand-alone GPU harness for the tcgen05 kernels (no Python): build with
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -lineinfo -o tools/tc5_harness tools/tc5_harness.cu
// and run on a B200.  It (1) checks ONE UMMA chunk (descriptors, swizzled operand layout, tcgen05.ld) against the CPU,
// dumping what the hardware read when that fails, (2) checks sb200_mlp_forward_tc5_f32 against a float64 CPU evaluation
// of the same network on ragged and full-size inputs, (3) times it with CUDA events.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../surreal_b200/csrc/mlp_fwd_tc5.cu"

unsigned long long g_sb200_launches = 0;

#define CK(x)                                                                                         \
    do {                                                                                              \
        cudaError_t e_ = (x);                                                                         \
        if (e_ != cudaSuccess) {                                                                      \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);           \
            exit(2);                                                                                  \
        }                                                                                             \
    } while (0)

// ---------------------------------------------------------------------------------------------------------------------
// unit: D[128][N] = A[128][32] . B[N][32]^T with ONE chunk (4 k-slices), operands written with sw128_off()
__global__ void __launch_bounds__(128, 1) unit_kernel(const float* __restrict__ A, const float* __restrict__ B, int N,
                                                      float* __restrict__ D) {
    extern __shared__ __align__(1024) unsigned char sm[];
    unsigned char* sa = sm;
    unsigned char* sb = sm + 128 * 128;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 128 * 128 + 256 * 128);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { tc5::mbar_init(bar, 1); tc5::mbar_init_fence(); }
    if (warp == 0) tc5::tmem_alloc(slot, 256);
    for (int i = tid; i < 128 * 32; i += 128) {
        const int r = i / 32, k = i % 32;
        *reinterpret_cast<float*>(sa + tc5::sw128_off(r, k)) = A[r * 32 + k];
    }
    for (int i = tid; i < N * 32; i += 128) {
        const int r = i / 32, k = i % 32;
        *reinterpret_cast<float*>(sb + tc5::sw128_off(r, k)) = B[r * 32 + k];
    }
    tc5::fence_async_smem();
    tc5::tc_fence_before();
    __syncthreads();
    tc5::tc_fence_after();
    const uint32_t tmem = *slot;
    if (tid == 0) {
        const uint32_t idesc = tc5::idesc_tf32(128, N);
        for (int ks = 0; ks < 4; ++ks)
            tc5::mma_tf32_ss(tmem, tc5::smem_desc_sw128(tc5::smem_u32(sa) + ks * 32), tc5::smem_desc_sw128(tc5::smem_u32(sb) + ks * 32),
                             idesc, ks > 0 ? 1u : 0u);
        tc5::mma_commit(bar);
    }
    __syncwarp();
    tc5::mbar_wait(bar, 0);
    tc5::tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 32) {
        float v[32];
        tc5::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        for (int i = 0; i < 32; ++i) D[(warp * 32 + lane) * N + c0 + i] = v[i];
    }
    tc5::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc5::tmem_dealloc(tmem, 256);
}

static float tf32_round(float x) {          // round-to-nearest (ties away) to 10 mantissa bits, like cvt.rna.tf32.f32
    uint32_t u;
    memcpy(&u, &x, 4);
    u += 0x1000u;
    u &= 0xFFFFE000u;
    float y;
    memcpy(&y, &u, 4);
    return y;
}

static int run_unit(int N, int pattern) {
    std::vector<float> A(128 * 32), B(N * 32), D(128 * N, 0.f);
    srand(1234 + N + pattern);
    for (auto& v : A) v = tf32_round((rand() / (float)RAND_MAX - 0.5f) * 2.f);
    for (auto& v : B) v = tf32_round((rand() / (float)RAND_MAX - 0.5f) * 2.f);
    if (pattern == 1) {     // A = [I_32; 0]: D[m][n] = B[n][m] for m < 32  -> shows which B element the hardware pairs with A's k
        for (auto& v : A) v = 0.f;
        for (int m = 0; m < 32; ++m) A[m * 32 + m] = 1.f;
        for (int n = 0; n < N; ++n)
            for (int k = 0; k < 32; ++k) B[n * 32 + k] = (float)(n * 32 + k + 1) / 64.f;     // exact in tf32 for N*32 <= 2048... approx otherwise
    }
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0, D.size() * 4));
    const size_t smem = 128 * 128 + 256 * 128 + 64;
    CK(cudaFuncSetAttribute(unit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    unit_kernel<<<1, 128, smem>>>(dA, dB, N, dD);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0;
    int bad = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < 32; ++k) ref += (double)A[m * 32 + k] * (double)B[n * 32 + k];
            const double e = fabs(ref - (double)D[m * N + n]);
            if (e > maxerr) maxerr = e;
            if (fabs(ref) > maxref) maxref = fabs(ref);
            if (e > 1e-4) ++bad;
        }
    printf("[unit N=%d pattern=%d] max|err| = %.3e (max|ref| %.3f), %d of %d entries off\n", N, pattern, maxerr, maxref, bad, 128 * N);
    if (bad && pattern == 1) {
        printf("  D[m][n] for m<8, n<8 (expect B[n][m] = (32n+m+1)/64):\n");
        for (int m = 0; m < 8; ++m) {
            printf("   m=%d:", m);
            for (int n = 0; n < 8; ++n) printf(" %8.4f", D[m * N + n] * 64.f);
            printf("\n");
        }
    }
    if (bad && pattern == 0) {
        printf("  first rows of D vs ref:\n");
        for (int m = 0; m < 4; ++m) {
            printf("   m=%d:", m);
            for (int n = 0; n < 6; ++n) {
                double ref = 0;
                for (int k = 0; k < 32; ++k) ref += (double)A[m * 32 + k] * (double)B[n * 32 + k];
                printf(" %8.4f/%8.4f", D[m * N + n], ref);
            }
            printf("\n");
        }
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return bad;
}

// ---------------------------------------------------------------------------------------------------------------------
struct Net {
    int K0, N1, N2, NO;
    std::vector<float> W1, b1, W2, b2, W3, b3;      // kernel layout W[k][ldw]
    int ldw3;
};

static void cpu_forward(const Net& n, const float* x, const float* zf, float eps, double* out) {
    std::vector<double> h0(n.K0), h1(n.N1), h2(n.N2);
    for (int k = 0; k < n.K0; ++k) {
        float v = x[k];
        if (zf) {
            const float cnt = zf[2 * n.K0];
            const float mean = zf[k] / cnt;
            const float var = zf[n.K0 + k] / cnt - mean * mean;
            const float sd = fmaxf(sqrtf(var), eps);
            v = fminf(fmaxf((v - mean) / sd, -5.f), 5.f);            // the reference divides; the kernel multiplies by 1/sd
        }
        h0[k] = v;
    }
    for (int j = 0; j < n.N1; ++j) {
        double s = n.b1[j];
        for (int k = 0; k < n.K0; ++k) s += h0[k] * (double)n.W1[k * n.N1 + j];
        h1[j] = s > 0 ? s : 0;
    }
    for (int j = 0; j < n.N2; ++j) {
        double s = n.b2[j];
        for (int k = 0; k < n.N1; ++k) s += h1[k] * (double)n.W2[k * n.N2 + j];
        h2[j] = s > 0 ? s : 0;
    }
    for (int o = 0; o < n.NO; ++o) {
        double s = n.b3[o];
        for (int k = 0; k < n.N2; ++k) s += h2[k] * (double)n.W3[k * n.ldw3 + o];
        out[o] = s;
    }
}

static int run_full(long long rows, int K0, int N1, int N2, int NO, bool use_zf, int time_iters) {
    Net n;
    n.K0 = K0; n.N1 = N1; n.N2 = N2; n.NO = NO; n.ldw3 = (NO + 3) / 4 * 4;
    srand(77 + (int)rows + NO);
    auto rnd = [](float sc) { return (rand() / (float)RAND_MAX * 2.f - 1.f) * sc; };
    n.W1.resize(K0 * N1); n.b1.resize(N1); n.W2.resize(N1 * N2); n.b2.resize(N2); n.W3.assign(N2 * n.ldw3, 0.f); n.b3.assign(n.ldw3, 0.f);
    for (auto& v : n.W1) v = rnd(1.f / sqrtf((float)K0));
    for (auto& v : n.b1) v = rnd(0.1f);
    for (auto& v : n.W2) v = rnd(1.f / sqrtf((float)N1));
    for (auto& v : n.b2) v = rnd(0.1f);
    for (int k = 0; k < N2; ++k)
        for (int o = 0; o < NO; ++o) n.W3[k * n.ldw3 + o] = rnd(1.f / sqrtf((float)N2));
    for (int o = 0; o < NO; ++o) n.b3[o] = rnd(0.1f);
    std::vector<float> x((size_t)rows * K0), zf(2 * K0 + 1);
    for (auto& v : x) v = rnd(2.5f);
    for (int k = 0; k < K0; ++k) { zf[k] = rnd(30.f); zf[K0 + k] = 200.f + rnd(50.f); }
    zf[2 * K0] = 100.f;
    float *dW1, *db1, *dW2, *db2, *dW3, *db3, *dx, *dzf, *dout;
    void* dws;
    CK(cudaMalloc(&dW1, n.W1.size() * 4)); CK(cudaMalloc(&db1, n.b1.size() * 4)); CK(cudaMalloc(&dW2, n.W2.size() * 4));
    CK(cudaMalloc(&db2, n.b2.size() * 4)); CK(cudaMalloc(&dW3, n.W3.size() * 4)); CK(cudaMalloc(&db3, n.b3.size() * 4));
    CK(cudaMalloc(&dx, x.size() * 4)); CK(cudaMalloc(&dzf, zf.size() * 4)); CK(cudaMalloc(&dout, (size_t)rows * NO * 4));
    CK(cudaMemcpy(dW1, n.W1.data(), n.W1.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(db1, n.b1.data(), n.b1.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dW2, n.W2.data(), n.W2.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(db2, n.b2.data(), n.b2.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dW3, n.W3.data(), n.W3.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(db3, n.b3.data(), n.b3.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dzf, zf.data(), zf.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dout, 0xFF, (size_t)rows * NO * 4));
    sb200_mlp net;
    memset(&net, 0, sizeof(net));
    net.n_layers = 3;
    net.dims[0] = K0; net.dims[1] = N1; net.dims[2] = N2; net.dims[3] = NO;
    net.act[0] = SB200_ACT_RELU; net.act[1] = SB200_ACT_RELU; net.act[2] = SB200_ACT_NONE;
    net.W[0] = dW1; net.W[1] = dW2; net.W[2] = dW3; net.b[0] = db1; net.b[1] = db2; net.b[2] = db3;
    net.ldw[0] = N1; net.ldw[1] = N2; net.ldw[2] = n.ldw3;
    net.aux_layer = -1;
    sb200_zfilter z;
    z.stats = use_zf ? dzf : nullptr;
    z.eps = 1e-5f;
    sb200_rows in;
    memset(&in, 0, sizeof(in));
    in.x = dx; in.ldx = K0; in.rows = rows;
    CK(cudaMalloc(&dws, sb200_mlp_tc5_workspace_bytes(&net)));
    if (!sb200_mlp_tc5_supported(&net, rows)) { printf("[full] net not supported\n"); return 1; }
    int rc = sb200_mlp_forward_tc5_f32(&net, &z, &in, dout, NO, dws, 0);
    if (rc != 0) { printf("[full] launch rc %d\n", rc); return 1; }
    CK(cudaDeviceSynchronize());
    std::vector<float> out((size_t)rows * NO);
    CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
    // reference on a sample of rows: head, tail, and a stride through the middle
    std::vector<long long> pick;
    for (long long r = 0; r < rows && r < 600; ++r) pick.push_back(r);
    for (long long r = rows > 400 ? rows - 400 : 0; r < rows; ++r) pick.push_back(r);
    for (long long r = 600; r < rows - 400; r += (rows / 1500 + 1)) pick.push_back(r);
    double maxerr = 0, sq = 0;
    long long cnt = 0, nan = 0;
    for (long long r : pick) {
        double ref[8];
        cpu_forward(n, &x[(size_t)r * K0], use_zf ? zf.data() : nullptr, 1e-5f, ref);
        for (int o = 0; o < NO; ++o) {
            const float g = out[(size_t)r * NO + o];
            if (!(g == g)) { ++nan; continue; }
            const double e = fabs((double)g - ref[o]);
            if (e > maxerr) maxerr = e;
            sq += ref[o] * ref[o];
            ++cnt;
        }
    }
    const double rms = sqrt(sq / (cnt > 0 ? cnt : 1));
    const double bar = 1e-5 * (rms > 1 ? rms : 1);
    const int ok = (maxerr <= bar && nan == 0);
    printf("[full rows=%lld %d-%d-%d-%d zf=%d] max|err| %.3e vs bar %.3e (rms %.3f), NaN %lld, checked %lld  -> %s\n", rows, K0, N1, N2,
           NO, (int)use_zf, maxerr, bar, rms, nan, (long long)pick.size(), ok ? "OK" : "FAIL");
    if (time_iters > 0) {
        float* flush;
        const size_t fl = 192u << 20;
        CK(cudaMalloc(&flush, fl));
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        float tot = 0, best = 1e9;
        for (int i = 0; i < time_iters + 2; ++i) {
            CK(cudaMemsetAsync(flush, i, fl, 0));
            cudaEventRecord(e0, 0);
            sb200_mlp_forward_tc5_f32(&net, &z, &in, dout, NO, dws, 0);
            cudaEventRecord(e1, 0);
            CK(cudaEventSynchronize(e1));
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (i >= 2) { tot += ms; if (ms < best) best = ms; }
        }
        const double flop = 2.0 * rows * ((double)K0 * N1 + (double)N1 * N2 + (double)N2 * NO);
        printf("[time rows=%lld] avg %.1f us, best %.1f us (prep + main, L2 flushed)  -> %.1f TFLOP/s effective fp32-accurate\n", rows,
               tot / time_iters * 1e3, best * 1e3, flop / (tot / time_iters * 1e-3) / 1e12);
        cudaFree(flush);
    }
    cudaFree(dW1); cudaFree(db1); cudaFree(dW2); cudaFree(db2); cudaFree(dW3); cudaFree(db3); cudaFree(dx); cudaFree(dzf); cudaFree(dout); cudaFree(dws);
    return ok ? 0 : 1;
}

int main(int argc, char** argv) {
    int dev = 0;
    cudaDeviceProp pr;
    CK(cudaGetDeviceProperties(&pr, dev));
    printf("device: %s  sm_%d%d  %d SMs\n", pr.name, pr.major, pr.minor, pr.multiProcessorCount);
    int fails = 0;
    fails += run_unit(256, 0) ? 1 : 0;
    fails += run_unit(64, 0) ? 1 : 0;
    fails += run_unit(32, 1) ? 1 : 0;
    if (argc > 1 && !strcmp(argv[1], "unit")) return fails;
    CK(sb200_mlp_tc5_init() == 0 ? cudaSuccess : cudaErrorUnknown);
    if (argc > 1 && !strcmp(argv[1], "prof")) return run_full(132096, 64, 256, 256, 1, true, 3);
#ifdef TC5_TRACE
    if (argc > 1 && !strcmp(argv[1], "trace")) {
        run_full(132096, 64, 256, 256, 1, true, 0);
        static long long tr[4096];
        CK(cudaMemcpyFromSymbol(tr, g_tc5_trace, sizeof(tr)));
        long long t0 = tr[1 * 1024 + 0];
        const char* roles[4] = {"loader", "mma", "prod", "head(tile q)"};
        printf("timeline of CTA 0 (cycles since the MMA thread first waited); chunks q = tile*10 + c\n");
        for (int q = 0; q < 30; ++q) {
            printf("q=%2d |", q);
            for (int r = 0; r < 4; ++r) {
                printf(" %s:", roles[r]);
                for (int e = 0; e < 4; ++e) {
                    long long v = tr[r * 1024 + q * 4 + e];
                    if (v) printf(" %7lld", v - t0); else printf("       -");
                }
                printf(" |");
            }
            printf("\n");
        }
        return 0;
    }
#endif
    fails += run_full(128, 64, 256, 256, 1, false, 0);
    fails += run_full(128 * 3 + 37, 64, 256, 256, 1, true, 0);
    fails += run_full(128 * 150 + 5, 32, 128, 64, 3, true, 0);
    fails += run_full(40000, 128, 256, 256, 8, true, 0);
    fails += run_full(132096, 64, 256, 256, 1, true, 10);
    printf("%s (%d failing groups)\n", fails ? "HARNESS FAIL" : "HARNESS OK", fails);
    return fails;
}
