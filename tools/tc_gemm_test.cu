// Stand-alone check of the tcgen05 GEMM core (csrc/gemm_tc.cuh) against a CPU reference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../magpie_tts_cpp_b200/csrc -o tc_gemm_test tc_gemm_test.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "gemm_tc.cuh"
using namespace mgb;
using bf = __nv_bfloat16;

template <int MT>
__global__ void __launch_bounds__(tc::kThreads, 1) tc_test_kernel(const bf * Wt, const bf * Xhi, const bf * Xlo, int KT, float * Y, int N, int M) {
    extern __shared__ unsigned char smem_raw[];
    const int nt = blockIdx.x, mt = blockIdx.y;
    const uint32_t tmem = tc::mainloop<MT>(smem_raw, Wt, Xhi, Xlo, KT, nt, mt);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp >= 2) {
        const int q = warp & 3;                      // TMEM lane quarter this warp may access
        const int n = nt * tc::BM + q * 32 + lane;
        for (int c = 0; c < MT; c += 32) {
            uint32_t v[32];
            tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + c, v);
            for (int j = 0; j < 32; j++) {
                const int m = mt * MT + c + j;
                if (n < N && m < M) Y[(size_t)m * N + n] = __uint_as_float(v[j]);
            }
        }
    }
    tc::finish<MT>(tmem);
}

static void make_image(const std::vector<float> & A, int rows, int K, int RT, std::vector<bf> & hi, std::vector<bf> * lo) {
    const int RTn = (rows + RT - 1) / RT, KT = K / 64;
    hi.assign((size_t)RTn * KT * RT * 64, __float2bfloat16(0.f));
    if (lo) lo->assign(hi.size(), __float2bfloat16(0.f));
    for (int r = 0; r < rows; r++)
        for (int k = 0; k < K; k++) {
            const size_t tile = (size_t)(r / RT) * KT + k / 64;
            const size_t off = tile * RT * 64 + tc::swz_offset(r % RT, k % 64) / 2;
            const float x = A[(size_t)r * K + k];
            const bf h = __float2bfloat16(x);
            hi[off] = h;
            if (lo) (*lo)[off] = __float2bfloat16(x - __bfloat162float(h));
        }
}

template <int MT> static int run(int N, int K, int M) {
    std::vector<float> W((size_t)N * K), X((size_t)M * K);
    srand(1234 + N + K + M);
    for (auto & w : W) w = (rand() / (float)RAND_MAX - 0.5f) * 0.1f;
    for (auto & x : X) x = (rand() / (float)RAND_MAX - 0.5f) * 2.0f;
    std::vector<bf> Wt, Xhi, Xlo;
    make_image(W, N, K, tc::BM, Wt, nullptr);
    make_image(X, M, K, MT, Xhi, &Xlo);
    bf * dW, * dH, * dL; float * dY;
    cudaMalloc(&dW, Wt.size() * 2); cudaMalloc(&dH, Xhi.size() * 2); cudaMalloc(&dL, Xlo.size() * 2); cudaMalloc(&dY, (size_t)M * N * 4);
    cudaMemcpy(dW, Wt.data(), Wt.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dH, Xhi.data(), Xhi.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dL, Xlo.data(), Xlo.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dY, 0xff, (size_t)M * N * 4);
    const int smem = tc::Smem<MT>::kBytes;
    cudaFuncSetAttribute(tc_test_kernel<MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    dim3 grid((N + tc::BM - 1) / tc::BM, (M + MT - 1) / MT);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    tc_test_kernel<MT><<<grid, tc::kThreads, smem>>>(dW, dH, dL, K / 64, dY, N, M);
    cudaEventRecord(e0);
    for (int i = 0; i < 20; i++) tc_test_kernel<MT><<<grid, tc::kThreads, smem>>>(dW, dH, dL, K / 64, dY, N, M);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d K=%d M=%d MT=%d: CUDA error %s\n", N, K, M, MT, cudaGetErrorString(e)); return 1; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<float> Y((size_t)M * N);
    cudaMemcpy(Y.data(), dY, Y.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int m = 0; m < M; m += (M > 64 ? 7 : 1))
        for (int n = 0; n < N; n += (N > 512 ? 13 : 1)) {
            double ref = 0;
            for (int k = 0; k < K; k++) ref += (double)__bfloat162float(__float2bfloat16(W[(size_t)n * K + k])) * (double)X[(size_t)m * K + k];
            maxerr = fmax(maxerr, fabs(ref - Y[(size_t)m * N + n])); maxref = fmax(maxref, fabs(ref));
        }
    const double us = ms * 1e3 / 20;
    printf("N=%4d K=%4d M=%3d MT=%3d: max err %.3e (max |ref| %.3e)  %.2f us  W stream %.0f GB/s  %s\n", N, K, M, MT, maxerr, maxref, us,
           (double)N * K * 2 / us / 1e3, maxerr < 1e-3 * fmax(maxref, 1.0) ? "OK" : "MISMATCH");
    cudaFree(dW); cudaFree(dH); cudaFree(dL); cudaFree(dY);
    return maxerr < 1e-3 * fmax(maxref, 1.0) ? 0 : 1;
}

int main() {
    int bad = 0;
    bad += run<64>(128, 64, 64);
    bad += run<64>(128, 768, 64);
    bad += run<64>(2304, 768, 64);
    bad += run<64>(768, 3072, 64);
    bad += run<64>(3072, 768, 50);
    bad += run<128>(2304, 768, 110);
    bad += run<128>(768, 768, 300);
    printf(bad ? "FAILED\n" : "ALL OK\n");
    return bad;
}
