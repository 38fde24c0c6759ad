// Local transformer, batch-1 / small-batch variant with SHARED-MEMORY-RESIDENT weights (bf16).
//
// Same mathematics as lt_kernel.cu (reference src/magpie.cpp:946-1048, 1072-1317).  One 16-CTA cluster per
// utterance; each CTA keeps its row slice of the layer matrices (qkv, o, ffn) in shared memory for all 8
// codebook passes and streams only its slice of the per-codebook output projection, prefetched one
// codebook ahead with cp.async.bulk while the next pass computes.  The feedback in-projection of the sampled
// code is a row gather from a table precomputed at load time (P_cb = E_cb . Win^T + b), which removes one
// GEMV and one cluster barrier per codebook.  Per codebook: 5 cluster barriers, all GEMVs from smem.
#include <cstdlib>

#include "lt_common.cuh"

namespace mgb {

namespace {

using namespace lt;

constexpr int kCS = 16;
constexpr int kQkvRows = 48, kORows = 16, kF1Rows = 64, kF2Rows = 16, kOutRows = 128;

struct alignas(128) LtResSmem {
    __nv_bfloat16 w_qkv[kQkvRows * kL];
    __nv_bfloat16 w_o[kORows * kL];
    __nv_bfloat16 w_ff1[kF1Rows * kL];
    __nv_bfloat16 w_ff2[kF2Rows * kF];
    __nv_bfloat16 w_out[kOutRows * kL];
    float hid_ffh[kF];                 // decoder hidden (prologue), then the FFN hidden vector
    float seq[kL], x[kL], nrm[kL], q[kL], att[kL], x1[kL], hout[kL];
    float kc[8][kL], vc[8][kL];
    float logits[kV];
    float sel_v[kV], srt_v[kV];
    uint16_t sel_i[kV], srt_i[kV];
    float red[32]; int redi[32];
    unsigned hist[256];
    float scores[8];
    int misc[8];
    uint64_t mbar[2];
};

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

struct RowSlice { int r0, nr; };
__device__ __forceinline__ RowSlice row_slice(int N, int rank) {
    const int rpc = (N + kCS - 1) / kCS;
    RowSlice s; s.r0 = rank * rpc; s.nr = max(0, min(rpc, N - s.r0));
    return s;
}

// GEMV over this CTA's smem-resident rows; every result is stored into all 16 CTAs of the cluster.
// route(n) -> local smem address of output row n; epi(n, acc) -> value.
template <typename Route, typename Epi>
__device__ __forceinline__ void res_gemv(cg::cluster_group & cluster, const __nv_bfloat16 * w, RowSlice sl, int K,
                                         const float * x, Route route, Epi epi) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int rb = warp * 4; rb < sl.nr; rb += kLtWarps * 4) {
        float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        for (int k = lane * 8; k < K; k += 256) {
            const float4 xa = *reinterpret_cast<const float4 *>(x + k), xb = *reinterpret_cast<const float4 *>(x + k + 4);
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const int rr = min(rb + r, sl.nr - 1);
                const uint4 u = *reinterpret_cast<const uint4 *>(w + (size_t)rr * K + k);
                acc[r] = fmaf(bf16lo(u.x), xa.x, acc[r]); acc[r] = fmaf(bf16hi(u.x), xa.y, acc[r]);
                acc[r] = fmaf(bf16lo(u.y), xa.z, acc[r]); acc[r] = fmaf(bf16hi(u.y), xa.w, acc[r]);
                acc[r] = fmaf(bf16lo(u.z), xb.x, acc[r]); acc[r] = fmaf(bf16hi(u.z), xb.y, acc[r]);
                acc[r] = fmaf(bf16lo(u.w), xb.z, acc[r]); acc[r] = fmaf(bf16hi(u.w), xb.w, acc[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < 4; r++) acc[r] = warp_sum(acc[r]);
#pragma unroll
        for (int i = 0; i < 2; i++) {          // 4 rows x 16 destination CTAs = 64 stores over 32 lanes
            const int idx = lane + 32 * i, r = idx / kCS, dr = idx % kCS;
            const float mine = r == 0 ? acc[0] : (r == 1 ? acc[1] : (r == 2 ? acc[2] : acc[3]));
            if (rb + r < sl.nr) {
                const int n = sl.r0 + rb + r;
                cluster.map_shared_rank(route(n), dr)[0] = epi(n, mine);
            }
        }
    }
}

// same, weights streamed from global (used once per frame for the in-projection of the decoder hidden state)
template <typename Epi>
__device__ __forceinline__ void glob_gemv(cg::cluster_group & cluster, const __nv_bfloat16 * W, RowSlice sl, int K,
                                          const float * x, float * dst, Epi epi) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = warp; r < sl.nr; r += kLtWarps) {
        const __nv_bfloat16 * wr = W + (size_t)(sl.r0 + r) * K;
        float acc = 0.0f;
        for (int k = lane * 8; k < K; k += 256) {
            float w[8];
            WT<__nv_bfloat16>::load(wr + k, w);
            const float4 xa = *reinterpret_cast<const float4 *>(x + k), xb = *reinterpret_cast<const float4 *>(x + k + 4);
            acc = fmaf(w[0], xa.x, acc); acc = fmaf(w[1], xa.y, acc); acc = fmaf(w[2], xa.z, acc); acc = fmaf(w[3], xa.w, acc);
            acc = fmaf(w[4], xb.x, acc); acc = fmaf(w[5], xb.y, acc); acc = fmaf(w[6], xb.z, acc); acc = fmaf(w[7], xb.w, acc);
        }
        acc = warp_sum(acc);
        if (lane < kCS) { const int n = sl.r0 + r; cluster.map_shared_rank(dst + n, lane)[0] = epi(n, acc); }
    }
}

__global__ void __cluster_dims__(kCS, 1, 1) __launch_bounds__(kLtThreads, 1) lt_resident_kernel(const LtParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    LtResSmem & S = *reinterpret_cast<LtResSmem *>(smem_raw);
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int utt = blockIdx.x / kCS;
    const int tid = threadIdx.x;
    const int d = p.d, L = p.L, F = p.F, V = p.V;
    const float att_scale = 1.0f / sqrtf((float)L);
    using bf = __nv_bfloat16;

    const RowSlice s_qkv = row_slice(3 * L, rank), s_o = row_slice(L, rank), s_f1 = row_slice(F, rank),
                   s_f2 = row_slice(L, rank), s_out = row_slice(V, rank), s_in = row_slice(L, rank);

    if (tid == 0) {
        mbar_init(&S.mbar[0], 1); mbar_init(&S.mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t b_qkv = s_qkv.nr * L * 2, b_o = s_o.nr * L * 2, b_f1 = s_f1.nr * L * 2, b_f2 = s_f2.nr * F * 2;
        mbar_expect_tx(&S.mbar[0], b_qkv + b_o + b_f1 + b_f2);
        if (b_qkv) bulk_g2s(S.w_qkv, (const bf *)p.qkv_w + (size_t)s_qkv.r0 * L, b_qkv, &S.mbar[0]);
        if (b_o) bulk_g2s(S.w_o, (const bf *)p.o_w + (size_t)s_o.r0 * L, b_o, &S.mbar[0]);
        if (b_f1) bulk_g2s(S.w_ff1, (const bf *)p.ff1_w + (size_t)s_f1.r0 * L, b_f1, &S.mbar[0]);
        if (b_f2) bulk_g2s(S.w_ff2, (const bf *)p.ff2_w + (size_t)s_f2.r0 * F, b_f2, &S.mbar[0]);
        const uint32_t b_out = s_out.nr * L * 2;
        mbar_expect_tx(&S.mbar[1], b_out);
        if (b_out) bulk_g2s(S.w_out, (const bf *)p.out_w[0] + (size_t)s_out.r0 * L, b_out, &S.mbar[1]);
    }

    const bool loop = p.d_step != nullptr || p.utt_step != nullptr;
    const int step = p.utt_step ? p.utt_step[utt] : (loop ? *p.d_step : (int)p.step);
    const size_t row = loop ? (size_t)utt * p.T_total + step : (size_t)utt;
    const int32_t * forced = p.forced ? p.forced + row * 8 : nullptr;
    const float * uniforms = p.uniforms ? p.uniforms + row * 8 : nullptr;
    float * logits_out = p.logits ? p.logits + row * 8 * V : nullptr;
    int32_t * sampled_out = p.sampled + row * 8;
    int32_t * argmax_out = p.argmax + row * 8;

    for (int i = tid; i < d; i += kLtThreads) {
        const float h = p.hidden[(size_t)utt * d + i];
        S.hid_ffh[i] = h;
        if (p.hidden_hist && rank == 0) p.hidden_hist[row * d + i] = h;
    }
    __syncthreads();
    cluster.sync();        // every CTA of the cluster is running before anyone writes into a peer's smem
    glob_gemv(cluster, (const bf *)p.in_w, s_in, d, S.hid_ffh, S.seq, [&](int n, float v) { return v + p.in_b[n]; });
    cluster.sync();
    mbar_wait(&S.mbar[0], 0);

    const bool forbid_eos = p.forbid_eos_all || (p.forbid_eos && p.forbid_eos[utt]) || (loop && step < p.min_frames);
    bool hit_eos = false;

    for (int cb = 0; cb < 8; cb++) {
        if (tid < L) S.x[tid] = S.seq[tid] + p.pos[cb * L + tid];
        __syncthreads();
        block_layer_norm(S.x, p.norm_self, S.nrm, L, p.eps, S.red);
        res_gemv(cluster, S.w_qkv, s_qkv, L, S.nrm,
                 [&](int n) { return n < L ? &S.q[n] : (n < 2 * L ? &S.kc[cb][n - L] : &S.vc[cb][n - 2 * L]); },
                 [](int, float v) { return v; });
        cluster.sync();
        {
            const int warp = tid >> 5, lane = tid & 31;
            if (warp <= cb) {
                float s = 0.0f;
                for (int i = lane; i < L; i += 32) s = fmaf(S.kc[warp][i], S.q[i], s);
                s = warp_sum(s);
                if (lane == 0) S.scores[warp] = s * att_scale;
            }
            __syncthreads();
            if (tid < L) {
                float mxs = S.scores[0];
                for (int j = 1; j <= cb; j++) mxs = fmaxf(mxs, S.scores[j]);
                float e[8], sum = 0.0f;
#pragma unroll
                for (int j = 0; j < 8; j++) { e[j] = (j <= cb) ? expf(S.scores[j] - mxs) : 0.0f; sum += e[j]; }
                const float inv = 1.0f / sum;
                float o = 0.0f;
#pragma unroll
                for (int j = 0; j < 8; j++) if (j <= cb) o = fmaf(e[j] * inv, S.vc[j][tid], o);
                S.att[tid] = o;
            }
            __syncthreads();
        }
        res_gemv(cluster, S.w_o, s_o, L, S.att, [&](int n) { return &S.x1[n]; }, [&](int n, float v) { return v + S.x[n]; });
        cluster.sync();
        block_layer_norm(S.x1, p.norm_ff, S.nrm, L, p.eps, S.red);
        res_gemv(cluster, S.w_ff1, s_f1, L, S.nrm, [&](int n) { return &S.hid_ffh[n]; },
                 [&](int, float v) { return gelu_ggml(v, p.gelu_f16); });
        cluster.sync();
        res_gemv(cluster, S.w_ff2, s_f2, F, S.hid_ffh, [&](int n) { return &S.hout[n]; }, [&](int n, float v) { return v + S.x1[n]; });
        cluster.sync();
        mbar_wait(&S.mbar[1], (uint32_t)(cb & 1));
        const float * ob = p.out_b[cb];
        res_gemv(cluster, S.w_out, s_out, L, S.hout, [&](int n) { return &S.logits[n]; }, [&](int n, float v) { return v + ob[n]; });
        __syncthreads();                         // this CTA is done reading w_out: prefetch the next codebook's slice
        if (tid == 0 && cb < 7) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            const uint32_t b_out = s_out.nr * L * 2;
            mbar_expect_tx(&S.mbar[1], b_out);
            if (b_out) bulk_g2s(S.w_out, (const bf *)p.out_w[cb + 1] + (size_t)s_out.r0 * L, b_out, &S.mbar[1]);
        }
        cluster.sync();
        if (tid < 8) {
            int id = tid == 0 ? p.bos_id : (tid < 7 ? p.bos_id + 1 + tid : (forbid_eos ? p.eos_id : -1));
            if (id >= 0 && id < V) S.logits[id] = -INFINITY;
        }
        __syncthreads();
        if (logits_out && rank == 0)
            for (int i = tid; i < V; i += kLtThreads) logits_out[(size_t)cb * V + i] = S.logits[i];
        const int am = block_argmax(S.logits, V, S.red, S.redi);
        int pick = am;
        if (p.temperature >= 0.01f) {
            float u;
            if (uniforms) u = uniforms[cb];
            else {
                uint32_t r[4];
                philox4x32_10((uint32_t)step, (uint32_t)utt, (uint32_t)cb, 0u, (uint32_t)p.seed, (uint32_t)(p.seed >> 32), r);
                u = (float)(r[0] >> 8) * (1.0f / 16777216.0f);
            }
            pick = block_sample_top_k(S, V, p.temperature, p.top_k, u);
        }
        hit_eos = hit_eos || pick == p.eos_id || am == p.eos_id;
        if (rank == 0 && tid == 0) {
            argmax_out[cb] = am;
            sampled_out[cb] = pick;
            if (p.next_codes) p.next_codes[utt * 8 + cb] = forced ? forced[cb] : pick;
            if (cb == 7) {
                if (p.eos_flag) p.eos_flag[utt] = hit_eos ? 1 : 0;
                if (p.done_step && hit_eos && p.done_step[utt] < 0) p.done_step[utt] = step;
            }
        }
        if (cb < 7) {
            // seq[cb+1] = in_proj . E_cb[code] + b  ==  row `code` of the precomputed table (no 1/8 scale, magpie.cpp:1285-1291)
            const int fed = forced ? forced[cb] : pick;
            if (tid < L) S.seq[tid] = p.in_table[cb][(size_t)fed * L + tid];
            __syncthreads();
        }
    }
    cluster.sync();        // no CTA may exit while peers can still write into its smem
}

}  // namespace

bool lt_resident_supported(const Model & m, int B) {
    const mgb_hparams & hp = m.hp;
    if (m.precision != MGB_PREC_BF16 || !m.lt_in_table[0] || getenv("MGB_LT_STREAM") != nullptr) return false;
    if (B > 9) return false;                      // 16 CTAs per utterance: more utterances are served by the streaming kernel
    if (hp.lt_dim > kL || hp.lt_dim % 8 || hp.lt_ffn_dim > kF || hp.vocab_per_cb > kV || hp.d_model > kF) return false;
    if ((3 * hp.lt_dim + kCS - 1) / kCS > kQkvRows || (hp.lt_dim + kCS - 1) / kCS > kORows ||
        (hp.lt_ffn_dim + kCS - 1) / kCS > kF1Rows || (hp.vocab_per_cb + kCS - 1) / kCS > kOutRows) return false;
    static std::atomic<int> ok_dev[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    if (ok_dev[dev & 63] == 0) {        // probed once per device; the result is published only when the probe is complete
        int ncl = 0;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(kCS); cfg.blockDim = dim3(kLtThreads); cfg.dynamicSmemBytes = sizeof(LtResSmem);
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = kCS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        const bool ok = cudaFuncSetAttribute(lt_resident_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LtResSmem)) == cudaSuccess &&
                        cudaFuncSetAttribute(lt_resident_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
                        cudaOccupancyMaxActiveClusters(&ncl, lt_resident_kernel, &cfg) == cudaSuccess && ncl >= 1;
        cudaGetLastError();
        ok_dev[dev & 63] = ok ? 1 : -1;
    }
    return ok_dev[dev & 63] == 1;
}

bool launch_lt_resident(const lt::LtParams & p, cudaStream_t stream) {
    lt_resident_kernel<<<p.B * kCS, kLtThreads, sizeof(LtResSmem), stream>>>(p);
    MGB_LAUNCH_CHECK();
    return true;
}

}  // namespace mgb
