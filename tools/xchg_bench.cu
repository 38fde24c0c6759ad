// Micro-benchmark of the all-to-all "flag packet" exchange used by frame_loop.cu: G CTAs, each produces its share of an
// N-packet vector (16-byte packets {v0,v1,v2,flag}, R replicas) and then polls all N packets.  Reports the time per
// exchange for several (N, R, variant) combinations.   nvcc -arch=sm_100a -O3 -o xchg_bench xchg_bench.cu && ./xchg_bench
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint4 ld_pkt(const uint4 * p) {
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_pkt(uint4 * p, unsigned a, unsigned b, unsigned c, unsigned flag) {
    asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(flag) : "memory");
}

// variant: 0 = every thread polls its packets (spin), 1 = spin with __nanosleep(64) between rounds,
//          2 = one flag per producer CTA: data written first, then __threadfence + flag; consumers poll G flags, then load data
template <int VARIANT>
__global__ void __launch_bounds__(512, 1) xchg_kernel(uint4 * buf, int N, int R, int iters, unsigned * flags, long long * out) {
    const int G = gridDim.x, b = blockIdx.x, tid = threadIdx.x;
    const int per = (N + G - 1) / G;                 // packets produced per CTA
    const int p0 = b * per, np = max(0, min(per, N - p0));
    __shared__ unsigned sink[512];
    unsigned acc = 0;
    long long t0 = 0;
    for (int it = 1; it <= iters; it++) {
        if (it == 11 && tid == 0) t0 = clock64();
        const unsigned flag = (unsigned)it;
        uint4 * wb = buf + (size_t)(it & 1) * N * R;        // two buffers: a CTA can be at most one exchange ahead
        const uint4 * in = wb + (size_t)(b % R) * N;
        unsigned * fl = flags + (size_t)(it & 1) * G * 8;
        if (VARIANT == 2) {
            if (tid < np * R) { const int pk = tid % np, r = tid / np; st_pkt(wb + (size_t)r * N + p0 + pk, acc, it, tid, 0u); }
            __syncthreads();
            if (tid < R) { __threadfence(); asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(fl + (size_t)tid * G + b), "r"(flag) : "memory"); }
            if (tid < G) {
                const unsigned * f = fl + (size_t)(b % R) * G + tid;
                unsigned v;
                int spins = 0;
                do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory"); } while (v != flag && ++spins < 2000000);
            }
            __syncthreads();
            for (int i = tid; i < N; i += blockDim.x) acc += ld_pkt(in + i).x;
        } else {
            if (tid < np * R) { const int pk = tid % np, r = tid / np; st_pkt(wb + (size_t)r * N + p0 + pk, acc, it, tid, flag); }
            for (int i = tid; i < N; i += blockDim.x) {
                uint4 v = ld_pkt(in + i);
                int spins = 0;
                while (v.w != flag && ++spins < 2000000) { if (VARIANT == 1) __nanosleep(64); v = ld_pkt(in + i); }
                acc += v.x;
            }
        }
        sink[tid] = acc;
        __syncthreads();
        acc = sink[(tid + 1) & 511] & 1u;            // a little dependent work so that iterations cannot overlap
    }
    if (tid == 0) out[b] = clock64() - t0;
    if (acc == 12345u) out[0] = 0;
}

template <int V> static double run(int G, int N, int R, int iters, uint4 * buf, unsigned * flags, long long * out) {
    cudaMemset(buf, 0, (size_t)N * R * 16 * 2);
    cudaMemset(flags, 0, (size_t)G * 8 * 4 * 2);
    void * args[] = {&buf, &N, &R, &iters, &flags, &out};
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    cudaError_t e = cudaLaunchCooperativeKernel((void *)xchg_kernel<V>, dim3(G), dim3(512), args, 0, 0);
    cudaEventRecord(e1);
    e = e == cudaSuccess ? cudaDeviceSynchronize() : e;
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); exit(1); }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms * 1e3 / iters;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int G = sms;
    uint4 * buf; unsigned * flags; long long * out;
    cudaMalloc(&buf, (size_t)4096 * 8 * 16 * 2); cudaMalloc(&flags, (size_t)G * 8 * 4 * 2); cudaMalloc(&out, G * 8);
    const int iters = 2000;
    printf("G=%d CTAs x 512 threads, %d exchanges each; us per exchange\n", G, iters);
    for (int N : {148, 256, 1024, 3168}) {
        for (int R : {1, 2, 4, 8}) {
            if ((N + G - 1) / G * R > 512) continue;
            printf("N=%4d R=%d   spin %.3f   sleep64 %.3f   cta-flags+fence %.3f\n", N, R, run<0>(G, N, R, iters, buf, flags, out),
                   run<1>(G, N, R, iters, buf, flags, out), run<2>(G, N, R, iters, buf, flags, out));
            fflush(stdout);
        }
    }
    return 0;
}
