// Transformer building blocks of the synthesis path (text encoder, context prefill, decoder step):
// fused LayerNorm -> skinny GEMV/GEMM -> bias/GELU/residual/KV-store epilogue, chunked online-softmax
// attention over the resident KV cache, and the embedding gathers.
//
// Reference semantics restated (paths in the reference repo):
//   LayerNorm           src/magpie.cpp:2237-2259      self-attention   src/magpie.cpp:1477-1575, 3395-3480
//   cross-attention     src/magpie.cpp:1663-1767      conv-FFN         src/magpie.cpp:1769-1918
//   embeddings          src/magpie.cpp:1319-1465, 2746-2787
#include <algorithm>

#include "common.cuh"
#include "gemm_tc.cuh"
#include "kernels.cuh"

namespace mgb {

namespace {

constexpr int kLinThreads = 256;
constexpr int kLinWarps = kLinThreads / 32;

struct LinParams {
    const void * W; int N, K, taps;
    const float * X; int ldx;
    const float * ln_w; float eps;
    const float * bias;
    const float * res; int ldr;
    float * Y; int ldy;
    int act, gelu_f16;
    const int32_t * tok_pos;
    int n_q, dkv;
    void * kdst; void * vdst;
    const int32_t * tok_slot;
    int M;
};

// One CTA: MB input rows (staged in smem, LayerNorm applied) x (8 warps * RW) weight rows.
// Each weight row is streamed exactly once per M-tile with 16-byte loads; fp32 accumulation.
template <typename T, int MB, int RW>
__global__ void __launch_bounds__(kLinThreads) linear_kernel(const LinParams p) {
    extern __shared__ __align__(16) float xs[];       // [MB][K]
    __shared__ float red[32];
    constexpr int VEC = WT<T>::VEC;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.y * MB;
    const int row0 = (blockIdx.x * kLinWarps + warp) * RW;
    const int K = p.K, N = p.N;

    float acc[RW][MB];
#pragma unroll
    for (int r = 0; r < RW; r++)
#pragma unroll
        for (int m = 0; m < MB; m++) acc[r][m] = 0.0f;

    for (int tap = 0; tap < p.taps; tap++) {
        const int shift = p.taps - 1 - tap;
        if (tap > 0) __syncthreads();
        for (int m = 0; m < MB; m++) {
            const int t = m0 + m;
            const bool valid = t < p.M && (shift == 0 || p.tok_pos[t] >= shift);
            const float * xr = p.X + (size_t)(valid ? t - shift : 0) * p.ldx;
            float * xm = xs + m * K;
            if (p.ln_w) {
                float s = 0.0f;
                for (int k = tid; k < K; k += kLinThreads) { float v = valid ? xr[k] : 0.0f; xm[k] = v; s += v; }
                const float mean = block_sum(s, red) / (float)K;
                float s2 = 0.0f;
                for (int k = tid; k < K; k += kLinThreads) { float v = xm[k] - mean; xm[k] = v; s2 += v * v; }
                const float var = block_sum(s2, red) / (float)K;
                const float scale = 1.0f / sqrtf(var + p.eps);
                for (int k = tid; k < K; k += kLinThreads) xm[k] = valid ? (xm[k] * scale) * p.ln_w[k] : 0.0f;
            } else {
                for (int k = tid; k < K; k += kLinThreads) xm[k] = valid ? xr[k] : 0.0f;
            }
        }
        __syncthreads();
        if (row0 < N) {
            const T * wbase = (const T *)p.W + ((size_t)tap * N + row0) * K;
            for (int k = lane * VEC; k < K; k += 32 * VEC) {
                float w[RW][VEC];
#pragma unroll
                for (int r = 0; r < RW; r++) {
                    if (row0 + r < N) WT<T>::load(wbase + (size_t)r * K + k, w[r]);
                    else {
#pragma unroll
                        for (int v = 0; v < VEC; v++) w[r][v] = 0.0f;
                    }
                }
#pragma unroll
                for (int m = 0; m < MB; m++) {
                    float x[VEC];
#pragma unroll
                    for (int v4 = 0; v4 < VEC; v4 += 4) {
                        float4 xv = *reinterpret_cast<const float4 *>(xs + m * K + k + v4);
                        x[v4] = xv.x; x[v4 + 1] = xv.y; x[v4 + 2] = xv.z; x[v4 + 3] = xv.w;
                    }
#pragma unroll
                    for (int r = 0; r < RW; r++)
#pragma unroll
                        for (int v = 0; v < VEC; v++) acc[r][m] = fmaf(w[r][v], x[v], acc[r][m]);
                }
            }
        }
    }

    float outv = 0.0f;
#pragma unroll
    for (int r = 0; r < RW; r++)
#pragma unroll
        for (int m = 0; m < MB; m++) {
            float v = warp_sum(acc[r][m]);
            if (lane == r * MB + m) outv = v;
        }
    if (lane < RW * MB) {
        const int r = lane / MB, m = lane % MB;
        const int n = row0 + r, t = m0 + m;
        if (n < N && t < p.M) {
            float v = outv;
            if (p.bias) v += p.bias[n];
            if (p.act == ACT_GELU) v = gelu_ggml(v, p.gelu_f16);
            if (p.res) v += p.res[(size_t)t * p.ldr + n];
            if (p.n_q < 0 || n < p.n_q) p.Y[(size_t)t * p.ldy + n] = v;
            else {
                const size_t slot = (size_t)p.tok_slot[t] * p.dkv;
                const int c = n - p.n_q;
                if (c < p.dkv) WT<T>::put((T *)p.kdst + slot + c, v);
                else WT<T>::put((T *)p.vdst + slot + (c - p.dkv), v);
            }
        }
    }
}

template <typename T, int MB, int RW>
bool launch_linear_t(const LinParams & p, cudaStream_t stream) {
    const size_t smem = (size_t)MB * p.K * sizeof(float);
    static DeviceOnce attr_done;                     // per-device bit (function attributes are per device)
    int dev = 0;
    MGB_CUDA_TRY(cudaGetDevice(&dev));
    if (!attr_done.done(dev)) {
        MGB_CUDA_TRY(cudaFuncSetAttribute(linear_kernel<T, MB, RW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_done.set(dev);
    }
    if (smem > 200 * 1024) { set_error("linear: K too large for the staged input tile"); return false; }
    dim3 grid((p.N + kLinWarps * RW - 1) / (kLinWarps * RW), (p.M + MB - 1) / MB);
    linear_kernel<T, MB, RW><<<grid, kLinThreads, smem, stream>>>(p);
    MGB_LAUNCH_CHECK();
    return true;
}

template <typename T>
bool launch_linear_p(const LinParams & p, cudaStream_t stream) {
    // rows per warp: keep >= ~2 waves of CTAs for the small matrices
    const int M = p.M;
    if (M == 1) {
        if (p.N >= 2048) return launch_linear_t<T, 1, 2>(p, stream);
        return launch_linear_t<T, 1, 1>(p, stream);
    }
    if (M == 2) return launch_linear_t<T, 2, 2>(p, stream);
    if (M <= 4) return launch_linear_t<T, 4, 2>(p, stream);
    if (p.K > 4096) return launch_linear_t<T, 4, 4>(p, stream);      // smem budget for very wide inputs
    return launch_linear_t<T, 8, 4>(p, stream);
}

// ---- attention ------------------------------------------------------------------------------------
struct AttnParams {
    const float * q; int ldq;
    const void * K; const void * V;
    int rows_per_utt, H, causal;
    const int32_t * n_ctx;
    const int32_t * utt; const int32_t * pos;
    float * out; int ldo;
    __nv_bfloat16 * pk_hi; __nv_bfloat16 * pk_lo;      // optional (DH == 64, <= 64 tokens): output as hi | lo tile images for the next GEMM
    int pdl;                                           // launched with programmatic stream serialization (see kernel)
    int split_min_keys;                                // SPLIT kernels: a token's keys are divided over the cluster from this many keys per CTA on
    int pf_keys;                                       // pdl: keys per CTA whose K / V rows are prefetched into L2 ahead of the wait (the launch's share of L2)
    // paged K / V (decoder self-attention cache): key j of utterance u lives in row page_table[u * max_pages + j / 128] * 128 + j % 128
    // of the layer's page pool; null = contiguous rows u * rows_per_utt + j (cross-attention K / V, encoder scratch)
    const int32_t * page_table; int max_pages;
    int dbg;
    int pk_f16;                                        // packed output as one f16 image (gemm_tc.cuh pack_act2)
};
constexpr int kPageShift = 7, kPageRows = 1 << kPageShift;     // = kKvPageRows (kernels.cuh)

constexpr int kAttnWarps = 8;

// grid (H, M), 128 threads.  A key/value row of one head (DH elements) is read by LPK = DH / VEC adjacent lanes with one
// 16-byte load each (fully used sectors), so a warp instruction covers 32 / LPK keys; each lane group keeps its own online
// softmax state (m, l, acc over the lane's VEC output dims) for the keys it owns, 4 key slots are in flight per lane, and
// the groups / warps are merged at the end.
//
// SPLIT (long KV, few tokens: BASELINE configs[4]): grid (H, M, S) launched as clusters (1, 1, S).  The S CTAs of a cluster divide
// the keys of their (head, token) into contiguous 128-aligned ranges (flash-decoding), so H x M x S work items fill the SMs in
// balanced waves and S times as many loads are in flight per KV row; ranks > 0 hand their (max, sum, unnormalised output) to
// rank 0 through DSMEM, which merges them in rank order (deterministic) and writes the output.
__device__ __forceinline__ void attn_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void attn_dsmem_store(float * local, unsigned rank, float v) {
    uint32_t la = (uint32_t)__cvta_generic_to_shared(local), ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(v) : "memory");
}
constexpr int kAttnMaxSplit = 8;

// MGB_ATTN_DBG: globaltimer stamps of CTA (0, 0, 0) of every pdl launch (entry | dependency wait over | q loaded | key scan done |
// warps merged | output stored), a ring of 1024 launches; dumped at the MGB_ATTN_DBG_DUMP-th launch_attention call (tools/ts_timeline.py)
__device__ unsigned long long g_attn_dbg[1024 * 8];
__device__ unsigned g_attn_dbg_n;

template <typename T, int DH, int SPLIT = 0, int NW = kAttnWarps>      // NW warps per CTA; SPLIT: 0 = one CTA per (head, token); 1 / 2 = pipelined scan + cluster key split compiled for 2 / 3 resident CTAs per SM
__global__ void __launch_bounds__(NW * 32, SPLIT == 2 ? (NW == 4 ? 6 : 3) : 0) attention_kernel(const AttnParams p) {
    // (the bf16 decoder-step variants are instruction-issue bound: softmax exponentials on the SFU, 2 instructions instead of ~8)
    auto fexp = [](float x) -> float { if constexpr (SPLIT >= 1) return __expf(x); else return expf(x); };
    constexpr int VEC = WT<T>::VEC;
    constexpr int LPK = DH / VEC;            // lanes per key row
    constexpr int KPI = 32 / LPK;            // keys per warp instruction
    constexpr int U = 4;                     // key slots in flight per lane
    static_assert(LPK >= 1 && LPK <= 32 && (LPK & (LPK - 1)) == 0, "head dim / vector width must be a power of two <= 32");
    __shared__ float s_m[NW], s_l[NW];
    __shared__ float s_acc[NW][DH];
    __shared__ float x_m[SPLIT ? kAttnMaxSplit : 1], x_l[SPLIT ? kAttnMaxSplit : 1], x_o[SPLIT ? kAttnMaxSplit : 1][DH];   // rank 0: the peers' partials
    const int t = blockIdx.y, h = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sub = lane % LPK, grp = lane / LPK;
    const int utt = p.utt[t];
    const int nk = p.causal ? p.pos[t] + 1 : p.n_ctx[utt];
    const float scale = 1.0f / sqrtf((float)DH);
    const int ld = p.H * DH;
    const int32_t * pt = p.page_table ? p.page_table + (size_t)utt * p.max_pages : nullptr;
    auto krow = [&](int j) -> size_t {      // storage row of key j of this token's utterance
        return pt ? (size_t)__ldg(pt + (j >> kPageShift)) * kPageRows + (size_t)(j & (kPageRows - 1)) : (size_t)utt * p.rows_per_utt + (size_t)j;
    };
    int k0 = 0, k1 = nk;                   // this CTA's key range
    unsigned crank = 0, csize = 1;
    if constexpr (SPLIT) {
        crank = blockIdx.z; csize = gridDim.z;             // cluster (1, 1, S) spans the whole z extent
        const int se = nk >= p.split_min_keys * (int)csize ? (int)csize : 1;
        const int chunk = (((nk + se - 1) / se) + 127) & ~127;
        k0 = min(nk, (int)crank * chunk); k1 = (int)crank < se ? min(nk, k0 + chunk) : k0;
    }
    __shared__ unsigned dbg_slot;
    const bool dbg = p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0;
    auto stamp = [&](int i) { unsigned long long tt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt)); g_attn_dbg[(size_t)dbg_slot * 8 + i] = tt; };
    if (dbg) { dbg_slot = atomicAdd(&g_attn_dbg_n, 1u) % 1024u; stamp(0); }
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");      // the following GEMM may start prefetching its weights
    if (p.pdl) {
        // launched as a programmatic dependent of the QKV GEMM: the old keys' K / V rows do not depend on it, so they are
        // pulled into L2 while that kernel is still running; q and the new key's row are read after the wait
        const char * Kp = (const char *)p.K + (size_t)(h * DH) * sizeof(T);
        const char * Vp = (const char *)p.V + (size_t)(h * DH) * sizeof(T);
        for (int j = k0 + tid; j < min(min(k1, nk - 1), k0 + p.pf_keys); j += NW * 32) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(Kp + krow(j) * ld * sizeof(T)));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(Vp + krow(j) * ld * sizeof(T)));
        }
        asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    if (dbg) stamp(1);
    float qv[VEC];
    {
        const float * qp = p.q + (size_t)t * p.ldq + h * DH + sub * VEC;
        if constexpr (VEC == 8) {
            const float4 a = *reinterpret_cast<const float4 *>(qp), b = *reinterpret_cast<const float4 *>(qp + 4);
            qv[0] = a.x * scale; qv[1] = a.y * scale; qv[2] = a.z * scale; qv[3] = a.w * scale;
            qv[4] = b.x * scale; qv[5] = b.y * scale; qv[6] = b.z * scale; qv[7] = b.w * scale;
        } else {
#pragma unroll
            for (int v = 0; v < VEC; v++) qv[v] = qp[v] * scale;
        }
    }
    const T * Kb = (const T *)p.K + h * DH + sub * VEC;
    const T * Vb = (const T *)p.V + h * DH + sub * VEC;
    if (dbg) { if (qv[0] == 123456.0f) stamp(7); stamp(2); }      // (the comparison makes the stamp wait for q)

    float mx = -INFINITY, l = 0.0f, acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; v++) acc[v] = 0.0f;

    if constexpr (SPLIT >= 1) {
        // long KV: software-pipelined scan.  The raw 16-byte K / V words of the NEXT 16 keys of the warp are requested before the
        // current ones are used (twice the bytes in flight per warp, loads overlap the softmax arithmetic); same arithmetic and
        // order per key as the loop below.
        constexpr int STEP = NW * KPI * U;
        static_assert((KPI * U) <= kPageRows && kPageRows % (KPI * U) == 0, "a warp's key group must not straddle a page");
        uint4 kr[U], vr[U];
        int base = k0 + warp * (KPI * U);
        // One address computation per GROUP of KPI * U = 16 keys (k0 and the groups are 16-aligned, so a group lies inside one 128-row page):
        // the kernel is instruction-issue bound at short KV (2 200 instructions per warp for ~130 keys, 24 warps per SM), and the per-key
        // page-table lookup + 64-bit row arithmetic was a third of them.
        auto request = [&](int b0) {
            if (b0 < k1) {                                               // (warp-uniform)
                const size_t r0 = krow(b0);
                if (b0 + KPI * U <= k1) {                                // (warp-uniform) whole group: one 64-bit base, constant strides
                    const T * kp = Kb + (r0 + (size_t)grp) * ld, * vp = Vb + (r0 + (size_t)grp) * ld;
                    const size_t st = (size_t)KPI * ld;
#pragma unroll
                    for (int u = 0; u < U; u++) { kr[u] = ldg_stream(kp); vr[u] = ldg_stream(vp); kp += st; vp += st; }
                } else {
                    // last, partial group: keys past the end are CLAMPED to the last one (no predicated loads, no zero fill); their scores
                    // are masked to -inf below, so their probability is exactly 0 and the finite duplicate V row contributes nothing
                    const unsigned last = (unsigned)(k1 - 1 - b0);
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        const size_t eo = (r0 + min((unsigned)(u * KPI + grp), last)) * (size_t)ld;
                        kr[u] = ldg_stream(Kb + eo); vr[u] = ldg_stream(Vb + eo);
                    }
                }
            }
        };
#pragma unroll
        for (int u = 0; u < U; u++) { kr[u] = make_uint4(0u, 0u, 0u, 0u); vr[u] = make_uint4(0u, 0u, 0u, 0u); }
        // NW == 4 (short KV, 6 resident CTAs per SM = 24 warps): no software pipelining -- the other warps hide the latency, and the copy of
        // the arrived words + the second register set cost instructions and spills the issue-bound kernel cannot afford
        constexpr bool PIPE = NW != 4;
        if (PIPE) request(base);
        for (; base < k1; base += STEP) {
            uint4 kc[U], vc[U];
            if (PIPE) {
#pragma unroll
                for (int u = 0; u < U; u++) { kc[u] = kr[u]; vc[u] = vr[u]; }
                request(base + STEP);
            } else {
                request(base);
#pragma unroll
                for (int u = 0; u < U; u++) { kc[u] = kr[u]; vc[u] = vr[u]; }
            }
            float sc[U], mnew = mx;
#pragma unroll
            for (int u = 0; u < U; u++) {
                float kk[VEC];
                WT<T>::unpack(kc[u], kk);
                float d = 0.0f;
#pragma unroll
                for (int v = 0; v < VEC; v++) d = fmaf(kk[v], qv[v], d);
#pragma unroll
                for (int o = LPK / 2; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
                sc[u] = (base + u * KPI + grp < k1) ? d : -INFINITY;
                mnew = fmaxf(mnew, sc[u]);
            }
            if (mnew != -INFINITY) {
                const float corr = fexp(mx - mnew);
                float ps = 0.0f;
#pragma unroll
                for (int v = 0; v < VEC; v++) acc[v] *= corr;
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const float pj = fexp(sc[u] - mnew);
                    ps += pj;
                    float vv[VEC];
                    WT<T>::unpack(vc[u], vv);
#pragma unroll
                    for (int v = 0; v < VEC; v++) acc[v] = fmaf(pj, vv[v], acc[v]);
                }
                l = l * corr + ps;
                mx = mnew;
            }
        }
    } else {
    for (int base = k0 + warp * (KPI * U); base < k1; base += NW * KPI * U) {
        float kk[U][VEC], vv[U][VEC], sc[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int j = base + u * KPI + grp;
            if (j < k1) {
                WT<T>::load(Kb + krow(j) * ld, kk[u]);
                WT<T>::load(Vb + krow(j) * ld, vv[u]);
            } else {
#pragma unroll
                for (int v = 0; v < VEC; v++) { kk[u][v] = 0.0f; vv[u][v] = 0.0f; }
            }
        }
        float mnew = mx;
#pragma unroll
        for (int u = 0; u < U; u++) {
            float d = 0.0f;
#pragma unroll
            for (int v = 0; v < VEC; v++) d = fmaf(kk[u][v], qv[v], d);
#pragma unroll
            for (int o = LPK / 2; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
            sc[u] = (base + u * KPI + grp < k1) ? d : -INFINITY;
            mnew = fmaxf(mnew, sc[u]);
        }
        if (mnew != -INFINITY) {
            const float corr = fexp(mx - mnew);            // mx = -inf -> 0
            float ps = 0.0f;
#pragma unroll
            for (int v = 0; v < VEC; v++) acc[v] *= corr;
#pragma unroll
            for (int u = 0; u < U; u++) {
                const float pj = fexp(sc[u] - mnew);       // -inf -> 0
                ps += pj;
#pragma unroll
                for (int v = 0; v < VEC; v++) acc[v] = fmaf(pj, vv[u][v], acc[v]);
            }
            l = l * corr + ps;
            mx = mnew;
        }
    }
    }
    if (dbg) { if (l == -1.0f) stamp(7); stamp(3); }
    // merge the key groups of the warp
#pragma unroll
    for (int o = LPK; o < 32; o <<= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, mx, o), ol = __shfl_xor_sync(0xffffffffu, l, o);
        const float mn = fmaxf(mx, om);
        const float fa = mx == -INFINITY ? 0.0f : fexp(mx - mn), fb = om == -INFINITY ? 0.0f : fexp(om - mn);
        l = l * fa + ol * fb;
#pragma unroll
        for (int v = 0; v < VEC; v++) {
            const float oa = __shfl_xor_sync(0xffffffffu, acc[v], o);
            acc[v] = acc[v] * fa + oa * fb;
        }
        mx = mn;
    }
    if (lane == 0) { s_m[warp] = mx; s_l[warp] = l; }
    if (grp == 0) {
#pragma unroll
        for (int v = 0; v < VEC; v++) s_acc[warp][sub * VEC + v] = acc[v];
    }
    __syncthreads();
    if (dbg) stamp(4);
    float M = -INFINITY, L = 0.0f, o = 0.0f;
    if (tid < DH) {
        M = s_m[0];
#pragma unroll
        for (int w = 1; w < NW; w++) M = fmaxf(M, s_m[w]);
#pragma unroll
        for (int w = 0; w < NW; w++) {
            const float f = s_m[w] == -INFINITY ? 0.0f : fexp(s_m[w] - M);
            L += f * s_l[w];
            o += f * s_acc[w][tid];
        }
    }
    if constexpr (SPLIT) {
        if (crank != 0 && tid < DH) {
            if (tid == 0) { attn_dsmem_store(&x_m[crank], 0, M); attn_dsmem_store(&x_l[crank], 0, L); }
            attn_dsmem_store(&x_o[crank][tid], 0, o);
        }
        if (csize > 1) attn_cluster_sync();   // every thread of every CTA of the cluster; rank 0's shared memory is written before it
        if (crank != 0) return;
        if (tid < DH) {
            for (unsigned r = 1; r < csize; r++) {          // fixed rank order: deterministic
                const float om = x_m[r];
                if (om == -INFINITY) continue;              // that rank had no keys
                const float mn = fmaxf(M, om);
                const float fa = M == -INFINITY ? 0.0f : fexp(M - mn), fb = fexp(om - mn);
                L = L * fa + x_l[r] * fb;
                o = o * fa + x_o[r][tid] * fb;
                M = mn;
            }
        }
    }
    if (tid < DH) {
        const float y = o * (1.0f / L);
        if (p.pk_hi) {
            // head h is exactly k tile h of the following GEMM (64 columns = one 128-byte swizzle row): gemm_tc.cu layout
            const size_t off = (size_t)h * (64 * 128) + (size_t)t * 128 + (((((tid >> 3) ^ (t & 7)) & 7) << 4) + ((tid & 7) << 1));
            if (p.pk_f16) {
                *reinterpret_cast<__half *>(reinterpret_cast<unsigned char *>(p.pk_hi) + off) = __float2half_rn(fminf(fmaxf(y, -65504.0f), 65504.0f));
            } else {
                const __nv_bfloat16 hv = __float2bfloat16_rn(y);
                *reinterpret_cast<__nv_bfloat16 *>(reinterpret_cast<unsigned char *>(p.pk_hi) + off) = hv;
                *reinterpret_cast<__nv_bfloat16 *>(reinterpret_cast<unsigned char *>(p.pk_lo) + off) = __float2bfloat16_rn(y - __bfloat162float(hv));
            }
        } else {
            p.out[(size_t)t * p.ldo + h * DH + tid] = y;
        }
    }
    if (dbg) stamp(5);
}

// ---- embeddings / LayerNorm -----------------------------------------------------------------------
struct AudioEmbParams { const float * emb[8]; const float * pos; const int32_t * codes; const int32_t * p; float * x; int d; };

__global__ void audio_embed_kernel(const AudioEmbParams a) {
    const int b = blockIdx.x;
    const int32_t * c = a.codes + b * 8;
    const float * pr = a.pos + (size_t)a.p[b] * a.d;
    for (int i = threadIdx.x; i < a.d; i += blockDim.x) {
        float s = a.emb[0][(size_t)c[0] * a.d + i];
#pragma unroll
        for (int cb = 1; cb < 8; cb++) s = s + a.emb[cb][(size_t)c[cb] * a.d + i];
        a.x[(size_t)b * a.d + i] = s * 0.125f + pr[i];
    }
}

__global__ void context_embed_kernel(const float * ctx, const float * pos, const int32_t * speakers, const int32_t * utt,
                                     const int32_t * tpos, float * x, int d, int C) {
    const int t = blockIdx.x;
    const int c = tpos[t];
    const float * src = ctx + ((size_t)speakers[utt[t]] * C + c) * d;
    const float * pr = pos + (size_t)c * d;
    for (int i = threadIdx.x; i < d; i += blockDim.x) x[(size_t)t * d + i] = src[i] + pr[i];
}

__global__ void text_embed_kernel(const float * emb, const float * pos, const int32_t * tokens, const int32_t * tpos,
                                  float * x, int d) {
    const int t = blockIdx.x;
    const float * src = emb + (size_t)tokens[t] * d;
    const float * pr = pos + (size_t)tpos[t] * d;
    for (int i = threadIdx.x; i < d; i += blockDim.x) x[(size_t)t * d + i] = src[i] + pr[i];
}

__global__ void __launch_bounds__(256) layer_norm_kernel(const float * x, const float * w, float eps, int d, float * y) {
    __shared__ float red[32];
    __shared__ float xs[1024];
    const int t = blockIdx.x, tid = threadIdx.x;
    const float * xr = x + (size_t)t * d;
    float s = 0.0f;
    for (int k = tid; k < d; k += 256) { float v = xr[k]; xs[k] = v; s += v; }
    const float mean = block_sum(s, red) / (float)d;
    float s2 = 0.0f;
    for (int k = tid; k < d; k += 256) { float v = xs[k] - mean; xs[k] = v; s2 += v * v; }
    const float var = block_sum(s2, red) / (float)d;
    const float scale = 1.0f / sqrtf(var + eps);
    for (int k = tid; k < d; k += 256) y[(size_t)t * d + k] = (xs[k] * scale) * w[k];
}

__global__ void add_one_kernel(int32_t * v, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] += 1;
}

}  // namespace

bool launch_linear(const LinearArgs & a, cudaStream_t stream) {
    if (a.M <= 0) return true;
    if (tc_linear_supported(a)) return launch_linear_tc(a, stream);
    if (a.ln_fold_stats || a.next_ln_w) { set_error("linear: a folded LayerNorm needs the tensor-core path"); return false; }
    LinParams p;
    p.W = a.W.w; p.N = a.W.N; p.K = a.W.K; p.taps = a.W.taps;
    p.X = a.X; p.ldx = a.ldx; p.ln_w = a.ln_w; p.eps = a.eps; p.bias = a.bias;
    p.res = a.res; p.ldr = a.ldr; p.Y = a.Y; p.ldy = a.ldy; p.act = a.act; p.gelu_f16 = a.gelu_f16;
    p.tok_pos = a.tok_pos; p.n_q = a.n_q; p.dkv = a.dkv; p.kdst = a.kdst; p.vdst = a.vdst;
    p.tok_slot = a.tok_slot; p.M = a.M;
    if (p.taps > 1 && !p.tok_pos) { set_error("linear: conv taps need token positions"); return false; }
    if (a.precision == MGB_PREC_F32) return launch_linear_p<float>(p, stream);
    return launch_linear_p<__nv_bfloat16>(p, stream);
}

// ---- context prefill: causal self-attention of C consecutive positions 0..C-1 of one utterance (magpie.cpp:3911-3988) -------
// grid (H, B), 256 threads: the head's K / V rows of the utterance are staged ONCE in shared memory (f32, K rows padded to
// 65 floats) and shared by all C queries; warp w handles queries w, w + 8, ...: lane-per-key dot products, warp softmax,
// lane-per-dimension P.V.  (The token-parallel attention_kernel re-reads the K / V rows for every query.)
template <typename T>
__global__ void __launch_bounds__(256) prefill_attention_kernel(const AttnParams p, int C) {
    constexpr int DH = 64, KS = DH + 1;
    extern __shared__ float pa_smem[];
    float * Ks = pa_smem;                    // [C][65]
    float * Vs = Ks + (size_t)C * KS;        // [C][64]
    float * sq = Vs + (size_t)C * DH;        // [8][64]
    float * sp = sq + 8 * DH;                // [8][128]
    const int h = blockIdx.x, u = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ld = p.H * DH;
    // (paged cache: the C <= 128 context rows are page 0 of the utterance)
    const int uu = p.utt[(size_t)u * C];         // utterance (session slot) of this run of C tokens: a prefill may cover a subset of the slots
    const size_t row0 = p.page_table ? (size_t)p.page_table[(size_t)uu * p.max_pages] * kPageRows : (size_t)uu * p.rows_per_utt;
    const T * Kb = (const T *)p.K + row0 * ld + h * DH;
    const T * Vb = (const T *)p.V + row0 * ld + h * DH;
    for (int i = tid; i < C * DH; i += 256) {
        const int j = i / DH, dd = i % DH;
        Ks[j * KS + dd] = WT<T>::get(Kb + (size_t)j * ld + dd);
        Vs[j * DH + dd] = WT<T>::get(Vb + (size_t)j * ld + dd);
    }
    __syncthreads();
    const float scale = 0.125f;              // 1 / sqrt(64)
    // queries are interleaved over gridDim.z CTAs (small batches would otherwise leave most SMs idle)
    for (int t = warp * gridDim.z + blockIdx.z; t < C; t += 8 * gridDim.z) {
        const size_t row = (size_t)u * C + t;
        sq[warp * DH + lane] = p.q[row * p.ldq + h * DH + lane] * scale;
        sq[warp * DH + lane + 32] = p.q[row * p.ldq + h * DH + lane + 32] * scale;
        __syncwarp();
        float sc[4], mx = -INFINITY;
#pragma unroll
        for (int k4 = 0; k4 < 4; k4++) {
            const int j = lane + 32 * k4;
            float dsum = -INFINITY;
            if (j <= t) {
                dsum = 0.0f;
                const float * kr = Ks + j * KS, * qr = sq + warp * DH;
#pragma unroll 16
                for (int dd = 0; dd < DH; dd++) dsum = fmaf(kr[dd], qr[dd], dsum);
            }
            sc[k4] = dsum;
            mx = fmaxf(mx, dsum);
        }
        mx = warp_max(mx);
        float sum = 0.0f;
#pragma unroll
        for (int k4 = 0; k4 < 4; k4++) {
            const float e = sc[k4] == -INFINITY ? 0.0f : expf(sc[k4] - mx);
            sp[warp * 128 + lane + 32 * k4] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        __syncwarp();
        float o0 = 0.0f, o1 = 0.0f;
        for (int j = 0; j <= t; j++) {
            const float pj = sp[warp * 128 + j];
            o0 = fmaf(pj, Vs[j * DH + lane], o0);
            o1 = fmaf(pj, Vs[j * DH + lane + 32], o1);
        }
        const float inv = 1.0f / sum;
        p.out[row * p.ldo + h * DH + lane] = o0 * inv;
        p.out[row * p.ldo + h * DH + lane + 32] = o1 * inv;
        __syncwarp();
    }
}

// Self-attention plan for a decoder step of `items` = heads x tokens (head, token) pairs that will reach `max_keys` keys.
// 0 = the one-CTA kernel; >= 1 = the software-pipelined kernel with a cluster of that many CTAs per item (the default for every
// batched decoder step: it is also 4 % faster per step at 64 utterances x 110..325 keys).  Measured at 32 utterances x 2512 keys (profiles/r1_long_kv_step_launches_summary.txt): the
// software-pipelined scan is what matters (2034 -> 1454 us per step); splitting the keys pays only while there are fewer items than
// resident CTA slots (S = 2: 1421, S = 3: 1521 us), so S just fills the machine.
int attention_plan_kv_split(int items, int max_keys) {
    if (items <= 0 || getenv("MGB_NO_ATTN_SPLIT")) return 0;
    if (getenv("MGB_ATTN_SPLIT")) return std::max(0, std::min(atoi(getenv("MGB_ATTN_SPLIT")), kAttnMaxSplit));
    const int long_min = getenv("MGB_ATTN_LONG_MIN") ? atoi(getenv("MGB_ATTN_LONG_MIN")) : 0;
    if (max_keys < long_min) return 0;
    static std::atomic<int> cap_dev[64];               // resident CTA slots, per device
    int dev = 0;
    cudaGetDevice(&dev);
    int cap = cap_dev[dev & 63];
    if (cap == 0) {
        int sms = 148, per_sm = 2;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, attention_kernel<__nv_bfloat16, 64, 1>, kAttnWarps * 32, 0) != cudaSuccess || per_sm < 1) {
            cudaGetLastError(); per_sm = 2;
        }
        cap = sms * per_sm;
        cap_dev[dev & 63] = cap;
    }
    int S = std::max(1, std::min((cap + items / 2) / items, kAttnMaxSplit));      // nearest: 16 utterances -> 2 (1169 -> 1140 us at 2512 keys)
    while (S > 1 && max_keys / S < 256) S--;
    return S;
}

bool launch_attention(const AttnArgs & a, cudaStream_t stream) {
    if (a.tok.M <= 0) return true;
    AttnParams p;
    p.q = a.q; p.ldq = a.ldq; p.K = a.K; p.V = a.V; p.rows_per_utt = a.rows_per_utt; p.H = a.H;
    p.causal = a.causal; p.n_ctx = a.n_ctx; p.utt = a.tok.utt; p.pos = a.tok.pos; p.out = a.out; p.ldo = a.ldo;
    p.pk_hi = nullptr; p.pk_lo = nullptr; p.pk_f16 = a.pack_f16 ? 1 : 0;
    p.page_table = a.page_table; p.max_pages = a.max_pages;
    if (a.pack_out) {
        if (a.dh != 64 || a.tok.M > 64) { set_error("attention: packed output needs head dim 64 and one token tile"); return false; }
        p.pk_hi = (__nv_bfloat16 *)a.pack_out; p.pk_lo = p.pk_hi + (size_t)64 * a.H * a.dh;
    }
    const bool f32 = a.precision == MGB_PREC_F32;
    if (a.prefill_len > 0 && a.causal && a.dh == 64 && a.prefill_len <= 128 && a.tok.M % a.prefill_len == 0 && !a.pack_out) {
        // tokens are (utterance-major, positions 0..C-1): one CTA per (head, utterance) with the K / V rows staged once
        const int C = a.prefill_len;
        const size_t smem = ((size_t)C * 65 + (size_t)C * 64 + 8 * 64 + 8 * 128) * sizeof(float);
        static DeviceOnce attr_done;
        int dev = 0;
        MGB_CUDA_TRY(cudaGetDevice(&dev));
        if (!attr_done.done(dev)) {
            MGB_CUDA_TRY(cudaFuncSetAttribute(prefill_attention_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
            MGB_CUDA_TRY(cudaFuncSetAttribute(prefill_attention_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
            attr_done.set(dev);
        }
        p.pdl = 0;
        const int nb = a.tok.M / C;
        const int qs = std::max(1, std::min(8, 296 / std::max(1, a.H * nb)));
        dim3 pg(a.H, nb, qs);
        if (f32) prefill_attention_kernel<float><<<pg, 256, smem, stream>>>(p, C);
        else prefill_attention_kernel<__nv_bfloat16><<<pg, 256, smem, stream>>>(p, C);
        MGB_LAUNCH_CHECK();
        return true;
    }
    dim3 grid(a.H, a.tok.M);
    p.pdl = (a.pack_out && !f32 && a.dh == 64) ? 1 : 0;       // decoder-step chain only
    static const bool dbg_on = getenv("MGB_ATTN_DBG") != nullptr;
    p.dbg = dbg_on && p.pdl;
    if (dbg_on && getenv("MGB_ATTN_DBG_DUMP")) {
        static int calls = 0;
        cudaStreamCaptureStatus cst = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(stream, &cst);
        static bool dumped = false;
        if (++calls >= atoi(getenv("MGB_ATTN_DBG_DUMP")) && cst == cudaStreamCaptureStatusNone && !dumped) {
            dumped = true;
            static unsigned long long h[1024 * 8];
            unsigned n = 0;
            cudaDeviceSynchronize();
            cudaMemcpyFromSymbol(h, g_attn_dbg, sizeof(h));
            cudaMemcpyFromSymbol(&n, g_attn_dbg_n, sizeof(n));
            for (unsigned i = n > 36 ? n - 36 : 0; i < n; i++) {
                const unsigned long long * r = h + (size_t)(i % 1024) * 8;
                fprintf(stderr, "attn %4u  entry->dep %6lld | q loaded %6lld | key scan %6lld | merge %6lld | finish %6lld | after dep %6lld ns  (t0 %llu)\n", i,
                        (long long)(r[1] - r[0]), (long long)(r[2] - r[1]), (long long)(r[3] - r[2]), (long long)(r[4] - r[3]), (long long)(r[5] - r[4]),
                        (long long)(r[5] - r[1]), r[0]);
            }
        }
    }
    p.split_min_keys = 256; p.pf_keys = 0;
    if (p.pdl) {
        const int S = std::max(0, std::min(a.kv_split, kAttnMaxSplit));       // 0: one-CTA kernel; >= 1: long-KV kernels, cluster of S
        // the prefetch ahead of the dependency wait is capped at about half of the 126 MB L2 for the whole launch: rows prefetched
        // beyond what L2 holds are evicted before they are used and cross HBM twice (long KV: 247 MB of K / V rows per layer)
        const double pf_mb = getenv("MGB_ATTN_PF_MB") ? atof(getenv("MGB_ATTN_PF_MB")) : 64.0;
        p.pf_keys = (int)std::min(1e9, pf_mb * 1048576.0 / ((double)a.H * a.tok.M * std::max(S, 1) * 2.0 * a.dh * sizeof(__nv_bfloat16)));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(a.H, a.tok.M, std::max(S, 1)); cfg.blockDim = dim3(kAttnWarps * 32); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        at[1].id = cudaLaunchAttributeClusterDimension;
        at[1].val.clusterDim.x = 1; at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = std::max(S, 1);
        cfg.attrs = at; cfg.numAttrs = S > 1 ? 2 : 1;              // (one-CTA "clusters": a plain launch; the kernel skips its cluster barrier)
        const bool occ3 = getenv("MGB_ATTN_OCC3") != nullptr;
        const bool nw4 = getenv("MGB_ATTN_NW8") == nullptr && a.H * a.tok.M >= 512;      // (1040 -> 1012 us per step at 64 utterances)
        if (S >= 1 && occ3) MGB_CUDA_TRY(cudaLaunchKernelEx(&cfg, attention_kernel<__nv_bfloat16, 64, 2>, p));
        else if (S == 1 && nw4) {
            // short KV, many (head, utterance) items: 4-warp CTAs -- 768 items fit the resident slots in about one wave instead of 2.6
            cfg.blockDim = dim3(4 * 32);
            // ... and compiled for 6 resident CTAs per SM (<= 80 registers) when the items would otherwise need a second, partial wave
            // (111 registers = 4 per SM = 592 slots < 768 items at 64 utterances)
            static const bool occ6 = getenv("MGB_ATTN_NO_OCC6") == nullptr;
            if (occ6 && a.H * a.tok.M > 592) MGB_CUDA_TRY(cudaLaunchKernelEx(&cfg, attention_kernel<__nv_bfloat16, 64, 2, 4>, p));
            else MGB_CUDA_TRY(cudaLaunchKernelEx(&cfg, attention_kernel<__nv_bfloat16, 64, 1, 4>, p));
        }
        else if (S >= 1) MGB_CUDA_TRY(cudaLaunchKernelEx(&cfg, attention_kernel<__nv_bfloat16, 64, 1>, p));
        else MGB_CUDA_TRY(cudaLaunchKernelEx(&cfg, attention_kernel<__nv_bfloat16, 64>, p));
        MGB_LAUNCH_CHECK();
        return true;
    }
    if (a.dh == 64) {
        if (f32) attention_kernel<float, 64><<<grid, kAttnWarps * 32, 0, stream>>>(p);
        else attention_kernel<__nv_bfloat16, 64><<<grid, kAttnWarps * 32, 0, stream>>>(p);
    } else if (a.dh == 128) {
        if (f32) attention_kernel<float, 128><<<grid, kAttnWarps * 32, 0, stream>>>(p);
        else attention_kernel<__nv_bfloat16, 128><<<grid, kAttnWarps * 32, 0, stream>>>(p);
    } else { set_error("attention: unsupported head dim"); return false; }
    MGB_LAUNCH_CHECK();
    return true;
}

// ---- folded cross-attention for a batched decoder step -----------------------------------------------------------------
// x_u += softmax(M_u LN(x_u; w)) N_u with the per-utterance tables M = scale K Wq (E x d) and N = V Wo^T (E x d) built at
// prefill (frame_loop.cu xattn_fold_kernel): the reference's q_net GEMV, 1-head attention over E text tokens and o_net GEMV
// (magpie.cpp:1713-1767, 3513) in ONE launch instead of five.  One CTA per utterance, 512 threads.
struct XFoldParams { float * x; const float * ln_w; float eps; const float * xm; const float * xn; const int32_t * n_ctx; int d, max_text;
                     const float * pack_ln_w; __nv_bfloat16 * pk_hi; __nv_bfloat16 * pk_lo; int pk_f16;
                     float * fold_stats; };       // non-null: emit (x_new .* pack_ln_w) + this rank's (sum, sum of squares) [kXC][64][2] instead of LN(x_new) (kernels.cuh)
// One CLUSTER of kXC CTAs per utterance; CTA rank r owns the d / kXC columns [r * dc, (r + 1) * dc) of the row, i.e. a
// quarter of both tables: partial scores and the LayerNorm statistics of the updated row are exchanged through DSMEM.
constexpr int kXC = 4, kXT = 256;
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void dsmem_store(float * local, int rank, float v) {
    uint32_t la = (uint32_t)__cvta_generic_to_shared(local), ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(v) : "memory");
}
// Written for LATENCY (the kernel sits in the dependent chain of the step, 12 times): every pass over a table issues all of a
// thread's loads before it uses any (one L2 round trip per pass instead of one per row), LayerNorm statistics are taken in one
// pass (sum and sum of squares together), and the second LayerNorm's two statistics cross the cluster in ONE exchange.
constexpr int kXRowsPerWarp = 16;      // score rows a warp keeps in flight (E <= 8 warps x 16 = 128 per round)
constexpr int kXRG = 5;                // row groups of the N pass: 5 x 48 float4 column groups = 240 threads
__global__ void __launch_bounds__(kXT) xattn_folded_kernel(const XFoldParams p) {
    __shared__ __align__(16) float xl[256];         // this CTA's columns: LN(x), later LN2(updated row)
    __shared__ __align__(16) float part[kXC][512];  // partial scores of every rank (written through DSMEM)
    __shared__ float prob[512];
    __shared__ __align__(16) float opart[kXRG][256]; // N pass: partial outputs of the row groups
    __shared__ float stat[2][kXC];                  // partial (sum, sum of squares) of the updated row, per rank
    __shared__ float red[2][8];
    const int u = blockIdx.x / kXC, rank = blockIdx.x % kXC;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int d = p.d, dc = d / kXC, c0 = rank * dc, E = p.n_ctx[u];
    float * xr = p.x + (size_t)u * d;
    const float * xm = p.xm + (size_t)u * p.max_text * d + c0, * xn = p.xn + (size_t)u * p.max_text * d + c0;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");      // the following GEMM may start prefetching its weights
    {
        // programmatic dependent of the O-projection GEMM: this CTA's quarter of the (static) tables is pulled into L2 while
        // that kernel is still running
        const char * m0 = (const char *)xm, * n0 = (const char *)xn;
        const int lines = dc * 4 / 128;                 // 128-byte lines per table row quarter
        for (int i = tid; i < E * lines; i += kXT) {
            const size_t off = (size_t)(i / lines) * d * 4 + (size_t)(i % lines) * 128;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(m0 + off));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(n0 + off));
        }
        asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    // LayerNorm statistics over the whole row in one pass, redundantly in every CTA (magpie.cpp:2237-2259)
    float xv[4], s = 0.0f, ss = 0.0f;
#pragma unroll
    for (int q = 0; q < 4; q++) { const int i = tid + q * kXT; xv[q] = i < d ? xr[i] : 0.0f; s += xv[q]; ss = fmaf(xv[q], xv[q], ss); }
    const float xraw = tid < dc ? xr[c0 + tid] : 0.0f;
    const float lw = tid < dc ? p.ln_w[c0 + tid] : 0.0f;
    s = warp_sum(s); ss = warp_sum(ss);
    if (lane == 0) { red[0][warp] = s; red[1][warp] = ss; }
    __syncthreads();
    s = ((red[0][0] + red[0][1]) + (red[0][2] + red[0][3])) + ((red[0][4] + red[0][5]) + (red[0][6] + red[0][7]));
    ss = ((red[1][0] + red[1][1]) + (red[1][2] + red[1][3])) + ((red[1][4] + red[1][5]) + (red[1][6] + red[1][7]));
    const float mean = s / (float)d;
    const float scale = 1.0f / sqrtf(fmaxf(ss / (float)d - mean * mean, 0.0f) + p.eps);
    if (tid < dc) xl[tid] = ((xraw - mean) * scale) * lw;
    __syncthreads();
    // partial scores over this CTA's columns, broadcast to the cluster: warp w owns rows w, w + 8, ...; the loads of up to 16 of
    // its rows are all issued before the first is used
    const int nf2 = dc / 64;                            // float2 per lane and row (dc = 192: 3)
    for (int jb = 0; jb < E; jb += 8 * kXRowsPerWarp) {
        float2 mv[kXRowsPerWarp][3];
#pragma unroll
        for (int r = 0; r < kXRowsPerWarp; r++) {
            const int j = jb + warp + 8 * r;
#pragma unroll
            for (int q = 0; q < 3; q++)
                mv[r][q] = (j < E && q < nf2) ? *reinterpret_cast<const float2 *>(xm + (size_t)j * d + lane * 2 + q * 64) : make_float2(0.0f, 0.0f);
        }
#pragma unroll
        for (int r = 0; r < kXRowsPerWarp; r++) {
            const int j = jb + warp + 8 * r;
            if (j >= E) break;                          // warp-uniform
            float a = 0.0f;
#pragma unroll
            for (int q = 0; q < 3; q++) { a = fmaf(mv[r][q].x, xl[lane * 2 + q * 64], a); a = fmaf(mv[r][q].y, xl[lane * 2 + q * 64 + 1], a); }
            a = warp_sum(a);
            if (lane == 0) part[rank][j] = a;          // (own copy; pushed to the peers in 16-byte pieces below)
        }
    }
    __syncthreads();
    // this rank's partial scores to the three peers: E / 4 sixteen-byte DSMEM stores per peer instead of one 4-byte store per (row, peer)
    {
        const int n4 = (E + 3) / 4;
        for (int i = tid; i < n4 * (kXC - 1); i += kXT) {
            const int peer = (rank + 1 + i / n4) % kXC, c = i % n4;
            const float4 v4 = *reinterpret_cast<const float4 *>(&part[rank][4 * c]);
            uint32_t la = (uint32_t)__cvta_generic_to_shared(&part[rank][4 * c]), ra;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(peer));
            asm volatile("st.shared::cluster.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(ra), "f"(v4.x), "f"(v4.y), "f"(v4.z), "f"(v4.w) : "memory");
        }
    }
    cluster_sync_all();
    // softmax over the E scores (each the sum of the four ranks' partials, in rank order): every warp redundantly, shuffles only
    float mx = -INFINITY;
    for (int j = lane; j < E; j += 32) mx = fmaxf(mx, ((part[0][j] + part[1][j]) + part[2][j]) + part[3][j]);
    mx = warp_max(mx);
    float se = 0.0f;
    for (int j = lane; j < E; j += 32) {
        const float e = expf((((part[0][j] + part[1][j]) + part[2][j]) + part[3][j]) - mx);
        if (warp == 0) prob[j] = e;
        se += e;
    }
    const float inv = 1.0f / warp_sum(se);
    __syncthreads();
    // out[c] = sum_j p_j N[j][c]: thread = (float4 column group cg, row group rg); rows rg, rg + 5, ... of the group are all
    // requested before the first is used; the 5 row groups are then added in a fixed order
    {
        const int ncg = dc / 4;                         // 48
        const int cg = tid % ncg, rg = tid / ncg;
        float4 o = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (rg < kXRG) {
            for (int jb = rg; jb < E; jb += kXRG * 8) {
                float4 nv[8];
#pragma unroll
                for (int r = 0; r < 8; r++) {
                    const int j = jb + r * kXRG;
                    nv[r] = j < E ? *reinterpret_cast<const float4 *>(xn + (size_t)j * d + cg * 4) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                }
#pragma unroll
                for (int r = 0; r < 8; r++) {
                    const int j = jb + r * kXRG;
                    const float pj = j < E ? prob[j] * inv : 0.0f;
                    o.x = fmaf(pj, nv[r].x, o.x); o.y = fmaf(pj, nv[r].y, o.y); o.z = fmaf(pj, nv[r].z, o.z); o.w = fmaf(pj, nv[r].w, o.w);
                }
            }
            *reinterpret_cast<float4 *>(&opart[rg][cg * 4]) = o;
        }
    }
    __syncthreads();
    float v = 0.0f;
    if (tid < dc) {
        v = xraw + ((((opart[0][tid] + opart[1][tid]) + opart[2][tid]) + opart[3][tid]) + opart[4][tid]);
        xr[c0 + tid] = v;
    }
    if (!p.pk_hi) { cluster_sync_all(); return; }         // (no CTA may exit while a peer can still write into its shared memory)
    if (p.fold_stats) {
        // LayerNorm folded through the following GEMM: no statistics exchange, no second cluster barrier -- this rank's columns of
        // (x_new .* w) go out as they are, with the rank's partial (sum, sum of squares) of the row for the GEMM's epilogue
        const float pwf = tid < dc ? p.pack_ln_w[c0 + tid] : 0.0f;
        float ps = warp_sum(tid < dc ? v : 0.0f), pq = warp_sum(tid < dc ? v * v : 0.0f);
        __syncthreads();                                  // (red is reused)
        if (lane == 0) { red[0][warp] = ps; red[1][warp] = pq; }
        if (tid < dc) xl[tid] = v * pwf;
        __syncthreads();
        if (tid == 0) {
            const float * a = red[0], * b = red[1];
            *reinterpret_cast<float2 *>(p.fold_stats + ((size_t)rank * 64 + u) * 2) =
                make_float2(((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7])), ((b[0] + b[1]) + (b[2] + b[3])) + ((b[4] + b[5]) + (b[6] + b[7])));
        }
        for (int kc = tid; kc < dc / 8; kc += kXT) {
            uint32_t h[4], l[4];
#pragma unroll
            for (int q = 0; q < 4; q++) tc::pack_act2(xl[kc * 8 + 2 * q], xl[kc * 8 + 2 * q + 1], p.pk_f16 != 0, h[q], l[q]);
            const int kg = (c0 >> 3) + kc;
            const size_t off = (size_t)(kg / 8) * (64 * 128) + (size_t)u * 128 + ((((kg % 8) ^ (u & 7)) & 7) << 4);
            *reinterpret_cast<uint4 *>(reinterpret_cast<unsigned char *>(p.pk_hi) + off) = make_uint4(h[0], h[1], h[2], h[3]);
            if (!p.pk_f16) *reinterpret_cast<uint4 *>(reinterpret_cast<unsigned char *>(p.pk_lo) + off) = make_uint4(l[0], l[1], l[2], l[3]);
        }
        return;
    }
    // LayerNorm of the updated row with the NEXT sub-block's weight -> bf16 hi | lo tile images (gemm_tc.cu layout, one
    // 64-token tile): the packing kernel in front of the FFN's first GEMM is not needed
    const float pw = tid < dc ? p.pack_ln_w[c0 + tid] : 0.0f;
    float ps = warp_sum(tid < dc ? v : 0.0f), pq = warp_sum(tid < dc ? v * v : 0.0f);
    __syncthreads();                                      // (red is reused)
    if (lane == 0) { red[0][warp] = ps; red[1][warp] = pq; }
    __syncthreads();
    if (tid < 2 * kXC) {
        const int which = tid / kXC, dst = tid % kXC;
        const float * r8 = red[which];
        dsmem_store(&stat[which][rank], dst, ((r8[0] + r8[1]) + (r8[2] + r8[3])) + ((r8[4] + r8[5]) + (r8[6] + r8[7])));
    }
    cluster_sync_all();
    const float mean2 = (((stat[0][0] + stat[0][1]) + stat[0][2]) + stat[0][3]) / (float)d;
    const float ex2 = (((stat[1][0] + stat[1][1]) + stat[1][2]) + stat[1][3]) / (float)d;
    const float scale2 = 1.0f / sqrtf(fmaxf(ex2 - mean2 * mean2, 0.0f) + p.eps);
    if (tid < dc) xl[tid] = ((v - mean2) * scale2) * pw;
    __syncthreads();
    for (int kc = tid; kc < dc / 8; kc += kXT) {
        uint32_t h[4], l[4];
#pragma unroll
        for (int q = 0; q < 4; q++) tc::pack_act2(xl[kc * 8 + 2 * q], xl[kc * 8 + 2 * q + 1], p.pk_f16 != 0, h[q], l[q]);
        const int kg = (c0 >> 3) + kc;                       // 8-column group index in the row
        const size_t off = (size_t)(kg / 8) * (64 * 128) + (size_t)u * 128 + ((((kg % 8) ^ (u & 7)) & 7) << 4);
        *reinterpret_cast<uint4 *>(reinterpret_cast<unsigned char *>(p.pk_hi) + off) = make_uint4(h[0], h[1], h[2], h[3]);
        if (!p.pk_f16) *reinterpret_cast<uint4 *>(reinterpret_cast<unsigned char *>(p.pk_lo) + off) = make_uint4(l[0], l[1], l[2], l[3]);
    }
}

bool launch_xattn_folded(float * x, const float * ln_w, float eps, const float * xm, const float * xn, const int32_t * n_ctx, int B, int d,
                         int max_text, const float * pack_ln_w, void * pack_out, cudaStream_t stream, bool pack_f16, float * fold_stats) {
    // (columns per rank: <= 256 with whole float4 / 64-column groups, at most 3 float2 per lane and row, 5 x dc / 4 <= 256 threads)
    if (d > 768 || max_text > 512 || d % (kXC * 64) != 0 || kXRG * (d / kXC / 4) > kXT) { set_error("xattn_folded: shape not supported"); return false; }
    if (pack_out && B > 64) { set_error("xattn_folded: packed output needs one token tile"); return false; }
    XFoldParams p{x, ln_w, eps, xm, xn, n_ctx, d, max_text, pack_ln_w, (__nv_bfloat16 *)pack_out,
                  pack_out ? (__nv_bfloat16 *)pack_out + (size_t)64 * d : nullptr, pack_f16 ? 1 : 0, pack_out ? fold_stats : nullptr};
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(B * kXC); cfg.blockDim = dim3(kXT); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = kXC; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 2;
    MGB_CUDA_TRY(cudaLaunchKernelEx(&cfg, xattn_folded_kernel, p));
    MGB_LAUNCH_CHECK();
    return true;
}

bool launch_audio_embed(const Model & m, const int32_t * codes, const int32_t * pos, int B, float * x, cudaStream_t stream) {
    AudioEmbParams a;
    for (int cb = 0; cb < 8; cb++) a.emb[cb] = m.audio_emb[cb];
    a.pos = m.dec_pos; a.codes = codes; a.p = pos; a.x = x; a.d = m.hp.d_model;
    audio_embed_kernel<<<B, 256, 0, stream>>>(a);
    MGB_LAUNCH_CHECK();
    return true;
}

bool launch_context_embed(const Model & m, const int32_t * speakers, Tokens tok, float * x, cudaStream_t stream) {
    context_embed_kernel<<<tok.M, 256, 0, stream>>>(m.baked_ctx, m.dec_pos, speakers, tok.utt, tok.pos, x, m.hp.d_model,
                                                    m.hp.context_frames);
    MGB_LAUNCH_CHECK();
    return true;
}

bool launch_text_embed(const Model & m, const int32_t * tokens, Tokens tok, float * x, cudaStream_t stream) {
    text_embed_kernel<<<tok.M, 256, 0, stream>>>(m.text_emb, m.enc_pos, tokens, tok.pos, x, m.hp.d_model);
    MGB_LAUNCH_CHECK();
    return true;
}

bool launch_layer_norm(const float * x, const float * w, float eps, int M, int d, float * y, cudaStream_t stream) {
    if (d > 1024) { set_error("layer_norm: d > 1024"); return false; }
    layer_norm_kernel<<<M, 256, 0, stream>>>(x, w, eps, d, y);
    MGB_LAUNCH_CHECK();
    return true;
}

__global__ void lt_fold_ov_kernel(const __nv_bfloat16 * qkv, const __nv_bfloat16 * o, int L, __nv_bfloat16 * out) {
    const int i = blockIdx.x, k = threadIdx.x;           // output row i, column k
    if (k >= L) return;
    out[(size_t)i * L + k] = qkv[(size_t)i * L + k];                              // Wq
    out[(size_t)(L + i) * L + k] = qkv[(size_t)(L + i) * L + k];                  // Wk
    float acc = 0.0f;
    for (int j = 0; j < L; j++) acc = fmaf(__bfloat162float(o[(size_t)i * L + j]), __bfloat162float(qkv[(size_t)(2 * L + j) * L + k]), acc);
    const __nv_bfloat16 h = __float2bfloat16_rn(acc);
    out[(size_t)(2 * L + i) * L + k] = h;
    out[(size_t)(3 * L + i) * L + k] = __float2bfloat16_rn(acc - __bfloat162float(h));
}
bool launch_lt_fold_ov(const void * qkv, const void * o, int L, void * out, cudaStream_t stream) {
    if (L > 1024) { set_error("lt_fold_ov: lt_dim too large"); return false; }
    lt_fold_ov_kernel<<<L, L, 0, stream>>>((const __nv_bfloat16 *)qkv, (const __nv_bfloat16 *)o, L, (__nv_bfloat16 *)out);
    MGB_LAUNCH_CHECK();
    return true;
}

__global__ void __launch_bounds__(256) lt_qkv_table_kernel(const float * in_table, const float * pos, const float * ln_w, float eps,
                                                           const __nv_bfloat16 * qkvo, int L, float * out) {
    __shared__ float xn[1024];
    __shared__ float red[32];
    const int code = blockIdx.x, tid = threadIdx.x;
    float v = 0.0f;
    for (int i = tid; i < L; i += 256) v += in_table[(size_t)code * L + i] + pos[i];
    const float mean = block_sum(v, red) / (float)L;
    float s2 = 0.0f;
    for (int i = tid; i < L; i += 256) { const float c = in_table[(size_t)code * L + i] + pos[i] - mean; s2 += c * c; }
    const float scale = 1.0f / sqrtf(block_sum(s2, red) / (float)L + eps);
    for (int i = tid; i < L; i += 256) xn[i] = ((in_table[(size_t)code * L + i] + pos[i] - mean) * scale) * ln_w[i];
    __syncthreads();
    // one warp per output row: coalesced 16-byte weight loads; rows [0, 2L) = q, k; rows [2L, 3L) = hi + lo of Wo Wv
    const int warp = tid >> 5, lane = tid & 31;
    for (int r = warp; r < 3 * L; r += 8) {
        float acc = 0.0f;
        for (int k = lane * 8; k < L; k += 256) {
            float w[8];
            WT<__nv_bfloat16>::load(qkvo + (size_t)r * L + k, w);
            if (r >= 2 * L) {
                float w2[8];
                WT<__nv_bfloat16>::load(qkvo + (size_t)(r + L) * L + k, w2);
#pragma unroll
                for (int q = 0; q < 8; q++) w[q] += w2[q];
            }
#pragma unroll
            for (int q = 0; q < 8; q++) acc = fmaf(w[q], xn[k + q], acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) out[(size_t)code * 3 * L + r] = acc;
    }
}
bool launch_lt_qkv_table(const float * in_table, const float * pos, const float * ln_w, float eps, const void * qkvo, int V, int L,
                         float * out, cudaStream_t stream) {
    if (L > 1024 || L % 8 != 0) { set_error("lt_qkv_table: lt_dim not supported"); return false; }
    lt_qkv_table_kernel<<<V, 256, 0, stream>>>(in_table, pos, ln_w, eps, (const __nv_bfloat16 *)qkvo, L, out);
    MGB_LAUNCH_CHECK();
    return true;
}

bool launch_add_one(int32_t * v, int n, cudaStream_t stream) {
    add_one_kernel<<<(n + 127) / 128, 128, 0, stream>>>(v, n);
    MGB_LAUNCH_CHECK();
    return true;
}

}  // namespace mgb
