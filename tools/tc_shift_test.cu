// Checks the shifted-row shared-memory descriptor (start address + s*128 B, base_offset = s & 7) that the implicit-GEMM
// conv uses for its taps:  D[n][m] = sum_k W[n][k] X[m + s][k].
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "gemm_tc.cuh"
using namespace mgb;
using bf = __nv_bfloat16;

__device__ __forceinline__ uint64_t desc_shift(uint32_t addr, int mode) {
    uint64_t d = tc::umma_desc_sw128(addr);
    if (mode == 1) d |= (uint64_t)((addr >> 7) & 7) << 49;      // base_offset
    return d;
}

__global__ void __launch_bounds__(128, 1) shift_kernel(const bf * Wimg, const bf * Ximg, int s, int mode, float * Y) {
    extern __shared__ unsigned char raw[];
    unsigned char * tiles = reinterpret_cast<unsigned char *>(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar_full, bar_acc;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { tc::mbar_init(&bar_full, 1); tc::mbar_init(&bar_acc, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(&tmem_slot)), "n"(64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        tc::mbar_expect_tx(&bar_full, 16384 + 32768);
        tc::bulk_g2s(tiles, Wimg, 16384, &bar_full);
        tc::bulk_g2s(tiles + 16384, Ximg, 32768, &bar_full);      // 256 rows x 128 B
        tc::mbar_wait(&bar_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a0 = tc::smem_u32(tiles), b0 = a0 + 16384 + s * 128;
        for (int j = 0; j < 4; j++) tc::umma_bf16(tmem, tc::umma_desc_sw128(a0 + j * 32), desc_shift(b0 + j * 32, mode), tc::umma_idesc_bf16(128, 64), j != 0);
        tc::umma_commit(&bar_acc);
    }
    tc::mbar_wait(&bar_acc, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[32];
    for (int c = 0; c < 64; c += 32) {
        tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
        for (int j = 0; j < 32; j++) Y[(size_t)(c + j) * 128 + warp * 32 + lane] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(64));
}

int main() {
    const int N = 128, K = 64, R = 256, MT = 64;
    std::vector<float> W(N * K), X(R * K);
    srand(7);
    for (auto & w : W) w = (rand() % 17 - 8) / 8.0f;
    for (auto & x : X) x = (rand() % 13 - 6) / 4.0f;
    std::vector<bf> Wi(N * K), Xi(R * K);
    for (int r = 0; r < N; r++) for (int k = 0; k < K; k++) Wi[tc::swz_offset(r, k) / 2] = __float2bfloat16(W[r * K + k]);
    for (int r = 0; r < R; r++) for (int k = 0; k < K; k++) Xi[tc::swz_offset(r, k) / 2] = __float2bfloat16(X[r * K + k]);
    bf * dW, * dX; float * dY;
    cudaMalloc(&dW, Wi.size() * 2); cudaMalloc(&dX, Xi.size() * 2); cudaMalloc(&dY, MT * N * 4);
    cudaMemcpy(dW, Wi.data(), Wi.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dX, Xi.data(), Xi.size() * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int mode = 0; mode < 2; mode++)
        for (int s : {0, 1, 3, 8, 13, 50, 100}) {
            shift_kernel<<<1, 128, 64 * 1024>>>(dW, dX, s, mode, dY);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
            std::vector<float> Y(MT * N);
            cudaMemcpy(Y.data(), dY, Y.size() * 4, cudaMemcpyDeviceToHost);
            double maxerr = 0;
            for (int m = 0; m < MT; m++) for (int n = 0; n < N; n++) {
                double ref = 0;
                for (int k = 0; k < K; k++) ref += (double)W[n * K + k] * X[(m + s) * K + k];
                maxerr = fmax(maxerr, fabs(ref - Y[m * N + n]));
            }
            printf("mode %d (base_offset %s) shift %3d: max err %.3e %s\n", mode, mode ? "set" : "0", s, maxerr, maxerr < 1e-3 ? "OK" : "WRONG");
        }
    return 0;
}
