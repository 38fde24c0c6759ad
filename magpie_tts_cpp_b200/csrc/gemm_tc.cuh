// tcgen05 / TMEM GEMM core for the batched (dense-contraction) paths: Y[m][n] = sum_k X[m][k] W[n][k].
//
// The weight matrix is the MMA "A" operand (M = 128 output features per CTA), the activations are the "B" operand
// (N = MT tokens per CTA), the accumulator D[128 x MT] lives in TMEM (fp32).  Both operands are K-major bf16 tiles
// of 64 k-elements (= one 128-byte swizzle row) stored in global memory as ready-made shared-memory IMAGES
// (SWIZZLE_128B canonical layout), so a tile is staged with ONE cp.async.bulk (no tensor map) and fed to
// tcgen05.mma through a shared-memory matrix descriptor.  Activations are split into bf16 hi + lo parts
// (x = hi + lo to 2^-17), two MMAs per k-slice, so the contraction keeps fp32-accurate activations like the
// batch-1 GEMV path; the tensor pipe is nowhere near the bottleneck at these shapes (HBM-bound on W).
//
// Warp roles (192 threads): warp 0 = bulk-copy producer, warp 1 = TMEM allocator + MMA issuer (one elected lane),
// warps 2..5 = epilogue (tcgen05.ld, 32 TMEM lanes each).  Pipelines: full/empty mbarriers per smem stage,
// tcgen05.commit releases a stage / publishes the accumulator.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace mgb {
namespace tc {

constexpr int BM = 128;          // weight rows per tile (MMA M)
constexpr int BK = 64;           // k per tile: 64 bf16 = 128 bytes = one swizzle row
constexpr int kThreads = 192;

// byte offset of element (r, k) inside a [rows x 64] bf16 K-major SWIZZLE_128B tile image
__host__ __device__ inline int swz_offset(int r, int k) { return r * 128 + ((((k >> 3) ^ (r & 7)) & 7) << 4) + ((k & 7) << 1); }

__device__ __forceinline__ uint32_t smem_u32(const void * p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t * bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t * bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t * bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void * dst, const void * src, uint32_t bytes, uint64_t * bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Two f32 activations -> one 32-bit word of an operand image.  Default: bf16 hi word + bf16 lo word (x = hi + lo to 2^-17, two MMAs per
// k slice).  f16 = true (MGB_ACT_F16, batched decoder step): ONE f16 image (11-bit mantissa: the rounding ggml's CPU path applies to the
// activations of f16-weight matmuls; 1/8 of the bf16 weights' own rounding step), half the operand bytes every CTA of a GEMM pulls from
// L2; values are clamped to the f16 range.  tcgen05 kind::f16 does NOT mix an f16 A with a bf16 B (illegal instruction on sm_100a), so
// the weights get an f16 twin of their images (model.cu: bf16 -> f16 is exact for |w| >= 6.1e-5, absolute error <= 3e-8 below).
__device__ __forceinline__ void pack_act2(float a, float b, bool f16, uint32_t & h, uint32_t & l) {
    if (f16) {
        const __half2 v = __floats2half2_rn(fminf(fmaxf(a, -65504.0f), 65504.0f), fminf(fmaxf(b, -65504.0f), 65504.0f));
        h = *reinterpret_cast<const uint32_t *>(&v); l = 0u;
    } else {
        const __nv_bfloat16 ha = __float2bfloat16_rn(a), hb = __float2bfloat16_rn(b);
        const __nv_bfloat16 la = __float2bfloat16_rn(a - __bfloat162float(ha)), lb = __float2bfloat16_rn(b - __bfloat162float(hb));
        h = (uint32_t)__bfloat16_as_ushort(ha) | ((uint32_t)__bfloat16_as_ushort(hb) << 16);
        l = (uint32_t)__bfloat16_as_ushort(la) | ((uint32_t)__bfloat16_as_ushort(lb) << 16);
    }
}

// shared-memory matrix descriptor, K-major, SWIZZLE_128B: start address (>>4), LBO = 1 (unused for swizzled K-major),
// SBO = 1024 B (8 rows x 128 B) >> 4, version = 1 (sm_100), layout type 2 = SWIZZLE_128B.  (cute/arch/mma_sm100_desc.hpp)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// instruction descriptor, kind::f16: D = f32 (bits 4-5 = 1), A/B = bf16 (1) at bits 7-9 / 10-12, both K-major,
// N >> 3 at bits 17-22, M >> 4 at bits 24-28
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// same with f16 operands (format 0)
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// Warp-uniform issue (the role's whole warp runs the loop, one elected lane issues): the low descriptor word is advanced by integer adds
// in 16-byte units and the constant high word (SBO = 1024 B >> 4 at bits 32-45, version 1 at bit 46, SWIZZLE_128B = 2 at bits 61-63) is
// attached inside the asm block, so nothing has to be moved from per-thread to uniform registers per MMA.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16); }
constexpr uint32_t kDescHiSw128 = 64u | (1u << 14) | (2u << 29);
__device__ __forceinline__ void umma_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "mov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
                 ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHiSw128) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t * bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                   "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                   "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int MT, int NB = 2, int SCAP = 8> struct Smem {
    static constexpr int kStageBytes = BM * 128 + NB * MT * 128;
    static constexpr int kStages = (200 * 1024) / kStageBytes < SCAP ? (200 * 1024) / kStageBytes : SCAP;
    static constexpr int kBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};

// Mainloop: accumulates D[128 x MT] (TMEM, fp32) for weight tile `nt`, token tile `mt`; returns the TMEM base address.
// Wt: [NT][KT] tiles of 16 KB; Xhi/Xlo: [MTiles][KT] tiles of MT*128 bytes.  Must be called by all 192 threads.
// After the call, epilogue warps (warp >= 2) own TMEM lanes 32*(warp%4) .. +31; call tc::finish() when done.
// NB = 2: activations as hi + lo tiles (bf16, two MMAs per k slice); NB = 1: one tile (Xlo unused).  F16: operands are f16.
// Only the k tiles [kt_begin, kt_begin + kt_count) are accumulated (split-K over a cluster; default: all KT).
template <int MT, int NB = 2, bool F16 = false, bool PDL = false, int SCAP = 8>
__device__ __forceinline__ uint32_t mainloop(unsigned char * smem_raw, const void * Wt, const void * Xhi,
                                             const void * Xlo, int KT, int nt, int mt, int kt_begin = 0, int kt_count = -1) {
    constexpr int S = Smem<MT, NB, SCAP>::kStages, SB = Smem<MT, NB, SCAP>::kStageBytes;
    const int NK = kt_count < 0 ? KT : kt_count;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char * tiles = reinterpret_cast<unsigned char *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t * bars = reinterpret_cast<uint64_t *>(tiles + S * SB);
    uint64_t * full = bars, * empty = bars + S, * acc_full = bars + 2 * S;
    uint32_t * tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * S + 1);
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {        // TMEM: MT fp32 columns (power of two >= 32)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(MT));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    // both roles run their loops warp-uniformly, one elected lane issues (see umma_lo above)
    if (warp == 0) {
        const bool leader = elect_one();
        const unsigned char * wsrc = reinterpret_cast<const unsigned char *>(Wt) + ((size_t)nt * KT + kt_begin) * (BM * 128);
        const unsigned char * hsrc = reinterpret_cast<const unsigned char *>(Xhi) + ((size_t)mt * KT + kt_begin) * (MT * 128);
        const unsigned char * lsrc = reinterpret_cast<const unsigned char *>(Xlo) + ((size_t)mt * KT + kt_begin) * (MT * 128);
        // Programmatic dependent launch: the weight tiles do not depend on the preceding kernel (which packs the
        // activations), so the first ring-full of them is requested BEFORE waiting for that kernel to complete.
        const int npre = PDL ? (NK < S ? NK : S) : 0;
        if (leader)
            for (int kt = 0; kt < npre; kt++) {
                mbar_expect_tx(&full[kt], SB);
                bulk_g2s(tiles + kt * SB, wsrc + (size_t)kt * (BM * 128), BM * 128, &full[kt]);
            }
        if (PDL) asm volatile("griddepcontrol.wait;" ::: "memory");
        for (int kt = 0; kt < NK; kt++) {
            const int s = kt % S;
            unsigned char * st = tiles + s * SB;
            if (kt >= npre) mbar_wait(&empty[s], ((kt / S) & 1) ^ 1);
            if (leader) {
                if (kt >= npre) {
                    mbar_expect_tx(&full[s], SB);
                    bulk_g2s(st, wsrc + (size_t)kt * (BM * 128), BM * 128, &full[s]);
                }
                bulk_g2s(st + BM * 128, hsrc + (size_t)kt * (MT * 128), MT * 128, &full[s]);
                if (NB == 2) bulk_g2s(st + BM * 128 + MT * 128, lsrc + (size_t)kt * (MT * 128), MT * 128, &full[s]);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        constexpr uint32_t idesc = F16 ? umma_idesc_f16(BM, MT) : umma_idesc_bf16(BM, MT);
        const bool leader = elect_one();
        const uint32_t tiles_d = desc_lo(smem_u32(tiles));
        for (int kt = 0; kt < NK; kt++) {
            const int s = kt % S;
            mbar_wait(&full[s], (kt / S) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a0 = tiles_d + (uint32_t)s * (SB >> 4), h0 = a0 + ((BM * 128) >> 4), l0 = h0 + ((MT * 128) >> 4);
            if (leader) {
#pragma unroll
                for (int j = 0; j < BK / 16; j++) {      // 16 k-elements = 32 bytes = 2 descriptor units inside the swizzle row
                    umma_lo(tmem_base, a0 + 2 * j, h0 + 2 * j, idesc, (kt | j) != 0);
                    if (NB == 2) umma_lo(tmem_base, a0 + 2 * j, l0 + 2 * j, idesc, 1u);
                }
                umma_commit(&empty[s]);                  // frees the stage when the MMAs above have read it
            }
        }
        if (leader) umma_commit(acc_full);               // accumulator complete
        __syncwarp();
    }
    if (warp >= 2) {
        if (PDL) asm volatile("griddepcontrol.wait;" ::: "memory");
        mbar_wait(acc_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    return tmem_base;
}

template <int MT> __device__ __forceinline__ void finish(uint32_t tmem_base) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if ((threadIdx.x >> 5) == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(MT));
}

}  // namespace tc
}  // namespace mgb
