// Shared device/host helpers for the sm_100a kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <string>

namespace mgb {

void set_error(const std::string & msg);           // thread-local, read by mgb_last_error()
extern thread_local int64_t g_launch_counter;      // kernels launched by this thread (bench evidence)

#define MGB_CUDA_TRY(expr)                                                                      \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            ::mgb::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));               \
            return false;                                                                       \
        }                                                                                       \
    } while (0)

#define MGB_LAUNCH_CHECK()                                                                      \
    do {                                                                                        \
        ::mgb::g_launch_counter++;                                                              \
        cudaError_t _e = cudaGetLastError();                                                    \
        if (_e != cudaSuccess) {                                                                \
            ::mgb::set_error(std::string("kernel launch failed: ") + cudaGetErrorString(_e) +   \
                             " at " + __FILE__ + ":" + std::to_string(__LINE__));               \
            return false;                                                                       \
        }                                                                                       \
    } while (0)

// Per-device "done once" flags for lazily set function attributes / uploaded constants.  Several host threads (one per GPU,
// sharding.py / mgb_pool) go through the launch wrappers concurrently: the bits are atomic so no device's flag is lost; two
// threads racing on the same device merely repeat an idempotent call.
struct DeviceOnce {
    std::atomic<uint64_t> bits{0};
    bool done(int dev) const { return (bits.load(std::memory_order_acquire) >> (dev & 63)) & 1u; }
    void set(int dev) { bits.fetch_or(1ull << (dev & 63), std::memory_order_release); }
};

constexpr int kWarp = 32;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum; `red` is >= 32 floats of shared scratch. All threads get the result.
__device__ __forceinline__ float block_sum(float v, float * red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    float t = (lane < nw) ? red[lane] : 0.0f;
    t = warp_sum(t);
    return t;
}

// ---- streaming 16-byte weight loads (read once per step: bypass L1 allocation) ----------------
__device__ __forceinline__ uint4 ldg_stream(const void * p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ float bf16lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

// Element-type traits: VEC = elements per 16-byte load.
template <typename T> struct WT;
template <> struct WT<float> {
    static constexpr int VEC = 4;
    __device__ static __forceinline__ void load(const float * p, float (&w)[4]) {
        uint4 u = ldg_stream(p);
        w[0] = __uint_as_float(u.x); w[1] = __uint_as_float(u.y);
        w[2] = __uint_as_float(u.z); w[3] = __uint_as_float(u.w);
    }
    __device__ static __forceinline__ void unpack(const uint4 & u, float (&w)[4]) {
        w[0] = __uint_as_float(u.x); w[1] = __uint_as_float(u.y); w[2] = __uint_as_float(u.z); w[3] = __uint_as_float(u.w);
    }
    __device__ static __forceinline__ float get(const float * p) { return *p; }
    __device__ static __forceinline__ void put(float * p, float v) { *p = v; }
};
template <> struct WT<__nv_bfloat16> {
    static constexpr int VEC = 8;
    __device__ static __forceinline__ void load(const __nv_bfloat16 * p, float (&w)[8]) {
        uint4 u = ldg_stream(p);
        w[0] = bf16lo(u.x); w[1] = bf16hi(u.x); w[2] = bf16lo(u.y); w[3] = bf16hi(u.y);
        w[4] = bf16lo(u.z); w[5] = bf16hi(u.z); w[6] = bf16lo(u.w); w[7] = bf16hi(u.w);
    }
    __device__ static __forceinline__ void unpack(const uint4 & u, float (&w)[8]) {
        w[0] = bf16lo(u.x); w[1] = bf16hi(u.x); w[2] = bf16lo(u.y); w[3] = bf16hi(u.y);
        w[4] = bf16lo(u.z); w[5] = bf16hi(u.z); w[6] = bf16lo(u.w); w[7] = bf16hi(u.w);
    }
    __device__ static __forceinline__ float get(const __nv_bfloat16 * p) { return __bfloat162float(*p); }
    __device__ static __forceinline__ void put(__nv_bfloat16 * p, float v) { *p = __float2bfloat16_rn(v); }
};

// ggml_gelu_f32 (tanh form).  f16_table != 0 reproduces ggml-CPU's 64K-entry f16 lookup:
// y = f32(f16(gelu(f32(f16(x))))), 0 for x <= -10, x for x >= 10  (SURVEY.md 8c item 2).
__device__ __forceinline__ float gelu_tanh(float x) {
    const float a = 0.044715f, s = 0.79788456080286535587989211986876f;
    return 0.5f * x * (1.0f + tanhf(s * x * (1.0f + a * x * x)));
}
__device__ __forceinline__ float gelu_ggml(float x, int f16_table) {
    if (!f16_table) return gelu_tanh(x);
    if (x <= -10.0f) return 0.0f;
    if (x >= 10.0f) return x;
    float xh = __half2float(__float2half_rn(x));
    return __half2float(__float2half_rn(gelu_tanh(xh)));
}

// Same function with tanh evaluated as 1 - 2 / (1 + e^{2z}) on the SFU (abs error ~1e-6, well below the f16 rounding of the
// table emulation): used by the latency-bound tensor-core epilogues, where libdevice's tanhf is most of the code size.
__device__ __forceinline__ float gelu_ggml_fast(float x, int f16_table) {
    const float a = 0.044715f, s = 0.79788456080286535587989211986876f;
    if (f16_table) {
        if (x <= -10.0f) return 0.0f;
        if (x >= 10.0f) return x;
        x = __half2float(__float2half_rn(x));
    }
    const float z = s * x * (1.0f + a * x * x);
    const float t = 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * z));
    const float y = 0.5f * x * (1.0f + t);
    return f16_table ? __half2float(__float2half_rn(y)) : y;
}

}  // namespace mgb
