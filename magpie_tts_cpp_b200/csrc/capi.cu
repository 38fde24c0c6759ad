// C-ABI (include/magpie_b200.h): sessions, the encoder / prefill / decoder-step drivers and the
// device-resident generation loops.  Every entry point only uploads inputs, launches sm_100a kernels
// on the session's stream and downloads results; there is no host arithmetic on the data path.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <functional>
#include <memory>
#include <string>
#include <vector>

#include "../../include/magpie_b200.h"
#include "common.cuh"
#include "kernels.cuh"
#include "model.h"
#include "mega.h"
#include "frame_loop.h"

namespace mgb {
const std::string & get_error();

struct Session {
    Model * m = nullptr;
    int B = 0, max_text = 0, max_seq = 0;
    int Mcap = 0;                      // token capacity of the activation buffers
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<void *> allocs;
    // activations
    float * x = nullptr, * qbuf = nullptr, * attn = nullptr, * xq = nullptr, * xatt = nullptr, * ffh = nullptr;
    float * hidden = nullptr;          // [B][d]
    float * enc_out = nullptr;         // [M_enc][d] compact
    // KV storage (model weight dtype)
    void * kc = nullptr, * vc = nullptr;       // [L][B*max_seq][d]
    void * xk = nullptr, * xv = nullptr;       // [L][B*max_text][dxa]
    void * ek = nullptr, * ev = nullptr;       // encoder scratch [B*max_text][d]
    // token maps
    int32_t * tok_utt = nullptr, * tok_pos = nullptr, * tok_slot = nullptr;   // [Mcap] scratch (encoder / prefill)
    int32_t * dec_utt = nullptr, * dec_pos = nullptr, * dec_slot = nullptr;   // [B] decode step
    int32_t * d_tokens = nullptr;      // [Mcap]
    int32_t * d_ntext = nullptr;       // [B]
    int32_t * d_speakers = nullptr;    // [B]
    int32_t * d_codes = nullptr, * d_sampled = nullptr, * d_argmax = nullptr;   // [B][8]
    int32_t * d_step = nullptr;        // [1]
    int32_t * d_done = nullptr;        // [B]
    uint8_t * d_forbid = nullptr;      // [B]
    int32_t * d_forced = nullptr; float * d_uniforms = nullptr;              // [B][8] single-shot
    float * d_logits1 = nullptr;       // [B][8][V] single-shot
    std::vector<int32_t> h_ntext, h_enc_off;
    std::vector<int32_t> h_speakers;
    std::vector<int> enc_slots;        // slots the resident encoder output belongs to (list order = compact row order)
    int M_enc = 0;
    int pos = 0;                       // host mirror of the (uniform) decode position
    int attn_split = 0;                // batched decoder step: CTAs per (head, utterance) in the self-attention (long KV, few utterances)
    bool encoded = false, prefilled = false;
    // loop buffers (grown on demand)
    int32_t * l_forced = nullptr, * l_sampled = nullptr, * l_argmax = nullptr; float * l_uniforms = nullptr;
    float * l_logits = nullptr, * l_hidden = nullptr;
    size_t l_cap_forced = 0, l_cap_sampled = 0, l_cap_argmax = 0, l_cap_uni = 0, l_cap_logits = 0, l_cap_hidden = 0;
    float last_ms = 0.0f; int64_t last_launches = 0;
    // batch-1 megakernel state
    int mega_grid = 0; float * attn_part = nullptr; unsigned * d_barrier = nullptr; int n_split = 1;
    unsigned long long * d_dbg = nullptr;
    // batch-1 bf16 persistent frame-loop kernel (frame_loop.cu)
    int loop_grid = 0; uint4 * d_xbuf = nullptr; int xoff[X_COUNT + 1] = {}; unsigned * d_seq = nullptr;
    int32_t * d_result = nullptr; float * d_xm = nullptr, * d_xn = nullptr; int loop_E = 0, loop_cap = 0; bool loop_tables = false;
    unsigned long long * d_loop_dbg = nullptr;
    void * tc_scratch = nullptr; size_t tc_scratch_bytes = 0;     // tensor-core path: activation tile images
    void * tc_scratch2 = nullptr;                                  // second image buffer (FF1 epilogue -> FF2 input)
    float * ln_stats = nullptr; bool ln_fold = false, act_f16 = false;              // LayerNorm folded through the QKV GEMM: row statistics [<= 96 slices][64][2] (kernels.cuh)
    void * lt_scratch = nullptr; size_t lt_scratch_bytes = 0;     // batched local transformer: activation scratch
    int prefill_len = 0;                                            // > 0 while mgb_prefill runs decoder_layers on the context frames
    float * fold_xm = nullptr, * fold_xn = nullptr; bool fold_ready = false;     // batched decode: folded cross-attention tables [L][B][max_text][d]
    // paged decoder self-attention cache: per layer a pool of n_pages pages of kKvPageRows rows; utterance b's cache position j
    // lives in row page_table[b][j / 128] * 128 + j % 128.  Pages are handed out lowest id first (a one-utterance session therefore
    // always holds pages 0, 1, 2, ... = contiguous rows, which the batch-1 persistent kernels rely on) and returned when an
    // utterance is retired (mgb_generate_queue).
    int max_pages = 0, n_pages = 0; size_t kv_rows = 0;
    int32_t * d_page_table = nullptr;          // [B][max_pages]
    std::vector<int32_t> h_pt;                 // host mirror
    std::vector<int> utt_pages;                // pages held by each utterance
    std::vector<int32_t> free_pages;           // min-heap of free page ids
    bool pt_dirty = false; bool pages_on_demand = false;
    // per-utterance loop state (continuous batching): steps emitted so far / 1 while the slot is generating
    int32_t * d_utt_step = nullptr, * d_active = nullptr;

    ~Session() {
        if (m) cudaSetDevice(m->device);
        for (void * p : allocs) cudaFree(p);
        for (void * p : {(void *)l_forced, (void *)l_sampled, (void *)l_argmax, (void *)l_uniforms, (void *)l_logits, (void *)l_hidden})
            if (p) cudaFree(p);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (stream) cudaStreamDestroy(stream);
    }
    template <typename T> bool alloc(T *& p, size_t n) {
        void * d = nullptr;
        if (cudaMalloc(&d, std::max<size_t>(n * sizeof(T), 16)) != cudaSuccess) { set_error("cudaMalloc failed (session)"); return false; }
        allocs.push_back(d); p = (T *)d;
        return true;
    }
};

static bool grow(void ** p, size_t * cap, size_t bytes) {
    if (bytes <= *cap) return true;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    MGB_CUDA_TRY(cudaMalloc(p, bytes));
    *cap = bytes;
    return true;
}

// ---- paged K / V cache -------------------------------------------------------------------------------
static bool ensure_pages(Session & s, int b, int rows) {          // utterance b must be able to hold cache positions [0, rows)
    const int need = (rows + kKvPageRows - 1) / kKvPageRows;
    if (need > s.max_pages) { set_error("KV cache: utterance exceeds the session's max_seq"); return false; }
    while (s.utt_pages[b] < need) {
        if (s.free_pages.empty()) { set_error("KV cache: page pool exhausted (create the session with more kv_pages)"); return false; }
        std::pop_heap(s.free_pages.begin(), s.free_pages.end(), std::greater<int32_t>());
        const int32_t pg = s.free_pages.back(); s.free_pages.pop_back();
        s.h_pt[(size_t)b * s.max_pages + s.utt_pages[b]++] = pg;
        s.pt_dirty = true;
    }
    return true;
}
static void release_pages(Session & s, int b) {
    for (int i = 0; i < s.utt_pages[b]; i++) {
        s.free_pages.push_back(s.h_pt[(size_t)b * s.max_pages + i]);
        std::push_heap(s.free_pages.begin(), s.free_pages.end(), std::greater<int32_t>());
        s.h_pt[(size_t)b * s.max_pages + i] = 0;
    }
    s.utt_pages[b] = 0; s.pt_dirty = true;
}
static bool flush_page_table(Session & s) {
    if (!s.pt_dirty) return true;
    MGB_CUDA_TRY(cudaMemcpyAsync(s.d_page_table, s.h_pt.data(), s.h_pt.size() * 4, cudaMemcpyHostToDevice, s.stream));
    s.pt_dirty = false;
    return true;
}
static bool ensure_pages_all(Session & s, int rows) {
    for (int b = 0; b < s.B; b++) if (!ensure_pages(s, b, rows)) return false;
    return flush_page_table(s);
}
// the batch-1 persistent kernels address the cache as contiguous rows: true if utterance 0 holds pages 0, 1, 2, ...
static bool pages_contiguous(const Session & s, int rows) {
    const int need = (rows + kKvPageRows - 1) / kKvPageRows;
    if (s.utt_pages[0] < need) return false;
    for (int i = 0; i < need; i++) if (s.h_pt[i] != i) return false;
    return true;
}
static inline int32_t cache_row(const Session & s, int b, int j) { return s.h_pt[(size_t)b * s.max_pages + j / kKvPageRows] * kKvPageRows + j % kKvPageRows; }

// ---- shared layer drivers -------------------------------------------------------------------------
static bool decoder_layers(Session & s, Tokens tok, bool want_hidden) {
    Model & m = *s.m; const mgb_hparams & hp = m.hp;
    // timing experiments only (results are garbage): bit 0 skips the cross-attention launches, bit 1 the self-attention, bit 2 the
    // QKV GEMM, bit 3 the O GEMM, bit 4 FF1, bit 5 FF2 -- the step-time difference is that kernel's LIVE cost inside the chain
    static const int skip = getenv("MGB_STEP_SKIP") ? atoi(getenv("MGB_STEP_SKIP")) : 0;
    const int d = hp.d_model, dxa = hp.dec_xa_heads * hp.dec_xa_d_head, M = tok.M;
    const size_t kv_layer = s.kv_rows * d * m.wsize, xkv_layer = (size_t)s.B * s.max_text * dxa * m.wsize;
    bool ln_folded = false;              // the previous layer's FF2 has written this layer's QKV operand (x .* norm_self) + row statistics
    // f16 activation images through the whole chain of a batched step (gemm_tc.cuh pack_act2): every producer / consumer pair below agrees
    const bool f16a = s.act_f16 && M == s.B && tok.utt == s.dec_utt && M <= 64 && M >= 4 && s.tc_scratch2 && s.fold_ready && m.precision == MGB_PREC_BF16 &&
                      hp.d_ffn % 64 == 0 && d / hp.dec_sa_heads == 64 && getenv("MGB_NO_CHAIN") == nullptr && getenv("MGB_NO_TS") == nullptr && !skip;
    for (int l = 0; l < hp.dec_layers; l++) {
        const DecLayer & L = m.dec[l];
        char * kl = (char *)s.kc + l * kv_layer, * vl = (char *)s.vc + l * kv_layer;
        LinearArgs a;
        a.tc_scratch = s.tc_scratch; a.tc_scratch_bytes = s.tc_scratch_bytes;
        a.precision = m.precision; a.eps = hp.eps; a.gelu_f16 = m.gelu_f16; a.M = M;
        // self-attention: LN -> QKV (K,V written straight into the cache) -> attention -> O + residual
        a.W = L.qkv; a.X = s.x; a.ldx = d; a.ln_w = L.norm_self; a.Y = s.qbuf; a.ldy = d;
        a.n_q = d; a.dkv = d; a.kdst = kl; a.vdst = vl; a.tok_slot = tok.slot;
        a.act_f16 = f16a; 
        if (ln_folded) {
            a.x_prepacked = true; a.ln_fold_stats = s.ln_stats; a.ln_fold_slices = d / ts_resid_nc(hp.d_ffn); a.ln_fold_csum = L.qkv_csum;
        }
        if (!(skip & 4) && !launch_linear(a, s.stream)) return false;
        LinearArgs o;
        o.tc_scratch = s.tc_scratch; o.tc_scratch_bytes = s.tc_scratch_bytes;
        o.precision = m.precision; o.M = M; o.W = L.o; o.X = s.attn; o.ldx = d; o.res = s.x; o.ldr = d; o.Y = s.x; o.ldy = d;
        AttnArgs at;
        at.precision = m.precision; at.q = s.qbuf; at.ldq = d; at.K = kl; at.V = vl; at.rows_per_utt = 0;
        at.page_table = s.d_page_table; at.max_pages = s.max_pages;
        at.H = hp.dec_sa_heads; at.dh = d / hp.dec_sa_heads; at.causal = 1; at.tok = tok; at.out = s.attn; at.ldo = d;
        at.prefill_len = s.prefill_len;
        // batched decoder step on the tensor-core path: the attention kernel writes the O-projection's packed input itself
        if (M == s.B && tok.utt == s.dec_utt && M <= 64 && at.dh == 64 && tc_linear_supported(o) && getenv("MGB_NO_CHAIN") == nullptr) {
            at.pack_out = s.tc_scratch; o.x_prepacked = true; at.kv_split = s.attn_split;
            at.pack_f16 = f16a; o.act_f16 = f16a;
        }
        if (!(skip & 2) && !launch_attention(at, s.stream)) return false;
        if (!(skip & 8) && !launch_linear(o, s.stream)) return false;
        // cross-attention over the cached encoder K/V (no mask); a batched decoder step (one token per utterance) uses
        // the folded tables: one launch instead of LN+pack, q GEMM, attention, pack, o GEMM
        // a batched decoder step on the tensor-core path chains the packed activations through the kernels: the folded
        // cross-attention emits LN(x) for FF1, FF1's epilogue emits GELU(h) for FF2 (no separate packing kernels)
        LinearArgs f1;
        f1.tc_scratch = s.tc_scratch; f1.tc_scratch_bytes = s.tc_scratch_bytes;
        f1.precision = m.precision; f1.eps = hp.eps; f1.gelu_f16 = m.gelu_f16; f1.M = M; f1.W = L.ff1; f1.X = s.x; f1.ldx = d;
        f1.ln_w = L.norm_ff; f1.act = ACT_GELU; f1.Y = s.ffh; f1.ldy = hp.d_ffn;
        LinearArgs f2;
        f2.tc_scratch = s.tc_scratch; f2.tc_scratch_bytes = s.tc_scratch_bytes;
        f2.precision = m.precision; f2.M = M; f2.W = L.ff2; f2.X = s.ffh; f2.ldx = hp.d_ffn; f2.res = s.x; f2.ldr = d; f2.Y = s.x; f2.ldy = d;
        const bool step = M == s.B && tok.utt == s.dec_utt;
        const bool chain = step && M <= 64 && s.tc_scratch2 && tc_linear_supported(f1) && tc_linear_supported(f2) && hp.d_ffn % 64 == 0 &&
                           getenv("MGB_NO_CHAIN") == nullptr;
        // (the folded cross-attention then emits x .* norm_ff + 4 statistics slices and FF1's epilogue applies the LayerNorm: no second cluster barrier)
        const bool ff1_fold = chain && !skip && s.ln_fold && s.ln_stats && L.ff1_csum && getenv("MGB_NO_TS") == nullptr;
        if (s.fold_ready && step) {
            const size_t tab = (size_t)s.B * s.max_text * d;
            const bool pk = chain;
            if (!(skip & 1) && !launch_xattn_folded(s.x, L.norm_xa_q, hp.eps, s.fold_xm + l * tab, s.fold_xn + l * tab, s.d_ntext, s.B, d, s.max_text,
                                                    pk ? L.norm_ff : nullptr, pk ? s.tc_scratch : nullptr, s.stream, pk && f16a,
                                                    (pk && ff1_fold) ? s.ln_stats + (size_t)96 * 64 * 2 : nullptr)) return false;
            if (pk) { f1.x_prepacked = true; f1.act_f16 = f16a; }
            if (pk && ff1_fold) { f1.ln_fold_stats = s.ln_stats + (size_t)96 * 64 * 2; f1.ln_fold_slices = 4; f1.ln_fold_csum = L.ff1_csum; }
        } else {
        LinearArgs q;
        q.tc_scratch = s.tc_scratch; q.tc_scratch_bytes = s.tc_scratch_bytes;
        q.precision = m.precision; q.eps = hp.eps; q.M = M; q.W = L.xq; q.X = s.x; q.ldx = d; q.ln_w = L.norm_xa_q; q.Y = s.xq; q.ldy = dxa;
        if (!launch_linear(q, s.stream)) return false;
        AttnArgs xt;
        xt.precision = m.precision; xt.q = s.xq; xt.ldq = dxa; xt.K = (char *)s.xk + l * xkv_layer; xt.V = (char *)s.xv + l * xkv_layer;
        xt.rows_per_utt = s.max_text; xt.H = hp.dec_xa_heads; xt.dh = hp.dec_xa_d_head; xt.causal = 0; xt.n_ctx = s.d_ntext;
        xt.tok = tok; xt.out = s.xatt; xt.ldo = dxa;
        if (!launch_attention(xt, s.stream)) return false;
        LinearArgs xo;
        xo.tc_scratch = s.tc_scratch; xo.tc_scratch_bytes = s.tc_scratch_bytes;
        xo.precision = m.precision; xo.M = M; xo.W = L.xo; xo.X = s.xatt; xo.ldx = dxa; xo.res = s.x; xo.ldr = d; xo.Y = s.x; xo.ldy = d;
        if (!launch_linear(xo, s.stream)) return false;
        }
        // conv-FFN (kernel 1): LN -> W1 -> GELU -> W2 + residual
        if (chain) { f1.pack_out = s.tc_scratch2; f1.Y = nullptr; f2.tc_scratch = s.tc_scratch2; f2.x_prepacked = true; f1.pack_f16 = f16a; f2.act_f16 = f16a; }
        else if (f16a) { set_error("decoder step: f16 activation images need the chained tensor-core path"); return false; }
        // ... and FF2's epilogue emits the NEXT layer's QKV operand with the LayerNorm folded through that GEMM (kernels.cuh)
        ln_folded = false;
        if (chain && !skip && s.ln_stats && l + 1 < hp.dec_layers && m.dec[l + 1].qkv_csum && s.ln_fold) {
            f2.pack_out = s.tc_scratch; f2.next_ln_w = m.dec[l + 1].norm_self; f2.stats_out = s.ln_stats; f2.pack_f16 = f16a;
            if (ts_linear_supported(f2)) ln_folded = true;
            else { f2.pack_out = nullptr; f2.next_ln_w = nullptr; f2.stats_out = nullptr; f2.pack_f16 = false; }
        }
        if (!(skip & 16) && !launch_linear(f1, s.stream)) return false;
        if (!(skip & 32) && !launch_linear(f2, s.stream)) return false;
    }
    if (want_hidden) return launch_layer_norm(s.x, m.dec_norm_out, hp.eps, M, d, s.hidden, s.stream);
    return true;
}

// after a step: position + 1 and the cache row of the NEXT position from the page table; with an `active` mask (continuous
// batching) retired / finished slots stay where they are (they keep being computed, their rows are rewritten in place)
__global__ void advance_kernel(int32_t * pos, int32_t * slot, int B, int32_t * step, const int32_t * page_table, int max_pages,
                               const int32_t * active, int32_t * utt_step) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B && (!active || active[i])) {
        const int np = pos[i] + 1;
        pos[i] = np;
        const int pg = np / kKvPageRows;        // (one past the last position of a full cache: no page, the row is never used)
        slot[i] = pg < max_pages ? page_table[(size_t)i * max_pages + pg] * kKvPageRows + np % kKvPageRows : 0;
        if (utt_step) utt_step[i] += 1;
    }
    if (i == 0 && step) *step += 1;
}

__global__ void step_inc_kernel(int32_t * step) { *step += 1; }

static bool launch_advance(Session & s, bool with_step, bool per_utt = false) {
    if (s.mega_grid > 0) {           // the megakernel advances pos/slot itself
        if (with_step) { step_inc_kernel<<<1, 1, 0, s.stream>>>(s.d_step); MGB_LAUNCH_CHECK(); }
        return true;
    }
    advance_kernel<<<(s.B + 127) / 128, 128, 0, s.stream>>>(s.dec_pos, s.dec_slot, s.B, with_step ? s.d_step : nullptr, s.d_page_table,
                                                            s.max_pages, per_utt ? s.d_active : nullptr, per_utt ? s.d_utt_step : nullptr);
    MGB_LAUNCH_CHECK();
    return true;
}

// batch 1: the whole step (embedding, 12 layers, final LayerNorm, position advance) is one cooperative kernel
static bool decoder_step_mega(Session & s) {
    Model & m = *s.m; const mgb_hparams & hp = m.hp;
    MegaParams p = {};
    for (int l = 0; l < hp.dec_layers; l++) {
        const DecLayer & L = m.dec[l];
        p.layer[l] = MegaLayer{L.qkv.w, L.o.w, L.xq.w, L.xo.w, L.ff1.w, L.ff2.w, L.norm_self, L.norm_xa_q, L.norm_ff};
    }
    p.L = hp.dec_layers; p.d = hp.d_model; p.f = hp.d_ffn; p.dxa = hp.dec_xa_heads * hp.dec_xa_d_head; p.H = hp.dec_sa_heads;
    p.n_split = s.n_split; p.eps = hp.eps; p.gelu_f16 = m.gelu_f16;
    for (int cb = 0; cb < 8; cb++) p.audio_emb[cb] = m.audio_emb[cb];
    p.dec_pos = m.dec_pos; p.norm_out = m.dec_norm_out; p.codes = s.d_codes;
    p.pos = s.dec_pos; p.pos_rw = s.dec_pos; p.slot_rw = s.dec_slot;
    p.kcache = s.kc; p.vcache = s.vc; p.kv_layer_stride = s.kv_rows * hp.d_model;
    p.xk = s.xk; p.xv = s.xv; p.xkv_layer_stride = (size_t)s.B * s.max_text * p.dxa; p.n_ctx = s.d_ntext;
    p.x = s.x; p.q = s.qbuf; p.attn_part = s.attn_part; p.xq = s.xq; p.ffh = s.ffh; p.hidden = s.hidden;
    p.barrier = s.d_barrier; p.dbg = s.d_dbg;
    return launch_decoder_mega(p, m.precision, s.mega_grid, s.stream);
}

// one decoder step on the codes in s.d_codes; leaves hidden in s.hidden
static bool decoder_step_device(Session & s) {
    if (s.mega_grid > 0) return decoder_step_mega(s);
    if (!launch_audio_embed(*s.m, s.d_codes, s.dec_pos, s.B, s.x, s.stream)) return false;
    Tokens tok; tok.M = s.B; tok.utt = s.dec_utt; tok.pos = s.dec_pos; tok.slot = s.dec_slot;
    return decoder_layers(s, tok, true);
}

static bool check_ready(Session * s, bool need_prefill) {
    if (!s) { set_error("null session"); return false; }
    if (cudaSetDevice(s->m->device) != cudaSuccess) { set_error("cudaSetDevice failed"); return false; }
    if (need_prefill && !s->prefilled) { set_error("session: call mgb_encode_text and mgb_prefill first"); return false; }
    return true;
}

}  // namespace mgb

using namespace mgb;

extern "C" {

const char * mgb_last_error(void) { return get_error().c_str(); }
const char * mgb_version(void) { return "magpie-b200 0.1 (sm_100a)"; }
int mgb_shard_device(int64_t utterance_index, int n_devices) {
    if (n_devices <= 0 || utterance_index < 0) return MGB_EINVAL;
    return (int)(utterance_index % n_devices);
}

int mgb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

mgb_model * mgb_model_load(const char * path, int device, int precision) {
    if (!path) { set_error("null path"); return nullptr; }
    try { return reinterpret_cast<mgb_model *>(load_model(path, device, precision)); }      // no exception may cross the C ABI
    catch (const std::exception & e) { set_error(std::string("magpie_init: ") + e.what()); return nullptr; }
    catch (...) { set_error("magpie_init: unknown failure"); return nullptr; }
}
void mgb_model_free(mgb_model * m) { delete reinterpret_cast<Model *>(m); }
int mgb_model_get_hparams(const mgb_model * m, mgb_hparams * out) {
    if (!m || !out) return MGB_EINVAL;
    *out = reinterpret_cast<const Model *>(m)->hp; return MGB_OK;
}
int mgb_model_set_max_dec_steps(mgb_model * m, int32_t n) {
    if (!m || n <= 0) return MGB_EINVAL;
    reinterpret_cast<Model *>(m)->hp.max_dec_steps = n; return MGB_OK;
}
int mgb_model_set_gelu_f16(mgb_model * m, int on) {
    if (!m) return MGB_EINVAL;
    reinterpret_cast<Model *>(m)->gelu_f16 = on ? 1 : 0; return MGB_OK;
}
int mgb_model_precision(const mgb_model * m) { return m ? reinterpret_cast<const Model *>(m)->precision : MGB_EINVAL; }
int mgb_model_device(const mgb_model * m) { return m ? reinterpret_cast<const Model *>(m)->device : MGB_EINVAL; }
int64_t mgb_model_step_weight_bytes(const mgb_model * m) { return m ? reinterpret_cast<const Model *>(m)->step_weight_bytes : 0; }
const char * mgb_model_meta_str(const mgb_model * m, const char * key) {
    if (!m || !key) return nullptr;
    auto & ms = reinterpret_cast<const Model *>(m)->meta_str;
    auto it = ms.find(key);
    return it == ms.end() ? nullptr : it->second.c_str();
}
int32_t mgb_model_meta_u32(const mgb_model * m, const char * key, int32_t def) {
    if (!m || !key) return def;
    auto & mu = reinterpret_cast<const Model *>(m)->meta_u32;
    auto it = mu.find(key);
    return it == mu.end() ? def : it->second;
}

mgb_session * mgb_session_new(mgb_model * mm, int batch, int max_text, int max_seq) {
    return mgb_session_new_paged(mm, batch, max_text, max_seq, 0);
}

mgb_session * mgb_session_new_paged(mgb_model * mm, int batch, int max_text, int max_seq, int kv_pages) {
    Model * m = reinterpret_cast<Model *>(mm);
    if (!m || batch <= 0 || max_text <= 0) { set_error("mgb_session_new: invalid arguments"); return nullptr; }
    const mgb_hparams & hp = m->hp;
    if (max_seq <= 0) max_seq = hp.context_frames + hp.max_dec_steps + 16;       // magpie.cpp:4077
    if (max_seq < hp.context_frames + 2) { set_error("mgb_session_new: max_seq too small"); return nullptr; }
    if (max_seq > m->dec_pos_rows) { set_error("mgb_session_new: max_seq exceeds the decoder position table"); return nullptr; }
    if (max_text > m->enc_pos_rows) { set_error("mgb_session_new: max_text exceeds the encoder position table"); return nullptr; }
    if (cudaSetDevice(m->device) != cudaSuccess) { set_error("cudaSetDevice failed"); return nullptr; }
    std::unique_ptr<Session> s(new Session());
    s->m = m; s->B = batch; s->max_text = max_text; s->max_seq = max_seq;
    const int d = hp.d_model, dxa = hp.dec_xa_heads * hp.dec_xa_d_head, L = hp.dec_layers, V = hp.vocab_per_cb;
    s->Mcap = batch * std::max(std::max(hp.context_frames, max_text), 1);
    const size_t M = (size_t)s->Mcap;
    if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&s->ev0) != cudaSuccess || cudaEventCreate(&s->ev1) != cudaSuccess) {
        set_error("mgb_session_new: stream/event creation failed"); return nullptr;
    }
    bool ok = s->alloc(s->x, M * d) && s->alloc(s->qbuf, M * d) && s->alloc(s->attn, M * d) && s->alloc(s->xq, M * dxa) &&
              s->alloc(s->xatt, M * dxa) && s->alloc(s->ffh, M * hp.d_ffn) && s->alloc(s->hidden, (size_t)batch * d) &&
              s->alloc(s->enc_out, (size_t)batch * max_text * d);
    char * p = nullptr;
    // paged self-attention cache: kv_pages = 0 reserves max_seq rows for every utterance (all pages assigned up front, utterance b
    // holding the contiguous pages b * max_pages ...); kv_pages > 0 sizes the pool explicitly and pages are assigned on demand
    s->max_pages = (max_seq + kKvPageRows - 1) / kKvPageRows;
    s->pages_on_demand = kv_pages > 0;
    s->n_pages = kv_pages > 0 ? kv_pages : batch * s->max_pages;
    if (s->n_pages < batch * ((hp.context_frames + 2 + kKvPageRows - 1) / kKvPageRows)) { set_error("mgb_session_new_paged: kv_pages cannot hold the context frames of every utterance"); return nullptr; }
    s->kv_rows = (size_t)s->n_pages * kKvPageRows;
    s->h_pt.assign((size_t)batch * s->max_pages, 0);
    s->utt_pages.assign(batch, 0);
    for (int pg = 0; pg < s->n_pages; pg++) s->free_pages.push_back(pg);
    std::make_heap(s->free_pages.begin(), s->free_pages.end(), std::greater<int32_t>());
    const size_t kvb = (size_t)L * s->kv_rows * d * m->wsize, xkvb = (size_t)L * batch * max_text * dxa * m->wsize;
    const size_t ekb = (size_t)batch * max_text * d * m->wsize;
    ok = ok && s->alloc(p, kvb); s->kc = p; ok = ok && s->alloc(p, kvb); s->vc = p;
    ok = ok && s->alloc(p, xkvb); s->xk = p; ok = ok && s->alloc(p, xkvb); s->xv = p;
    ok = ok && s->alloc(p, ekb); s->ek = p; ok = ok && s->alloc(p, ekb); s->ev = p;
    ok = ok && s->alloc(s->tok_utt, M) && s->alloc(s->tok_pos, M) && s->alloc(s->tok_slot, M) && s->alloc(s->d_tokens, M) &&
         s->alloc(s->dec_utt, batch) && s->alloc(s->dec_pos, batch) && s->alloc(s->dec_slot, batch) &&
         s->alloc(s->d_ntext, batch) && s->alloc(s->d_speakers, batch) && s->alloc(s->d_codes, (size_t)batch * 8) &&
         s->alloc(s->d_sampled, (size_t)batch * 8) && s->alloc(s->d_argmax, (size_t)batch * 8) && s->alloc(s->d_step, 1) &&
         s->alloc(s->d_done, batch) && s->alloc(s->d_forbid, batch) && s->alloc(s->d_forced, (size_t)batch * 8) &&
         s->alloc(s->d_uniforms, (size_t)batch * 8) && s->alloc(s->d_logits1, (size_t)batch * 8 * V) &&
         s->alloc(s->d_page_table, (size_t)batch * s->max_pages) && s->alloc(s->d_utt_step, batch) && s->alloc(s->d_active, batch);
    if (!ok) return nullptr;
    if (!s->pages_on_demand)
        for (int b = 0; b < batch; b++) if (!ensure_pages(*s, b, max_seq)) return nullptr;
    s->pt_dirty = true;
    if (!flush_page_table(*s) || cudaStreamSynchronize(s->stream) != cudaSuccess) { set_error("mgb_session_new: page table upload failed"); return nullptr; }
    if (m->precision == MGB_PREC_BF16 && s->Mcap >= 2 && m->dec.size() && m->dec[0].qkv.tiles) {
        const size_t tb = tc_scratch_bytes(s->Mcap, std::max(hp.d_ffn, hp.d_model));
        char * tp = nullptr;
        if (!s->alloc(tp, tb)) return nullptr;
        s->tc_scratch = tp; s->tc_scratch_bytes = tb;
        char * tp2 = nullptr;
        if (!s->alloc(tp2, tb)) return nullptr;
        s->tc_scratch2 = tp2;
        if (!s->alloc(s->ln_stats, (size_t)2 * 96 * 64 * 2)) return nullptr;      // [FF2 -> QKV: <= 96 slices | cross-attention -> FF1: 4 slices]
        s->ln_fold = getenv("MGB_NO_LNFOLD") == nullptr;
        s->act_f16 = m->dec[0].qkv.tiles16 != nullptr;             // (MGB_ACT_F16 at model load)
    }
    // (launch_xattn_folded holds one score per text position in shared memory: up to 512 positions; longer text capacities keep
    //  the q GEMM + attention + o GEMM kernels)
    if (m->precision == MGB_PREC_BF16 && batch >= 2 && dxa == 128 && max_text <= 512 && d % 256 == 0 && d <= 768 && getenv("MGB_NO_XFOLD") == nullptr &&
        (size_t)L * batch * max_text * d * 8 <= ((size_t)8 << 30)) {
        const size_t tab = (size_t)L * batch * max_text * d;
        if (!s->alloc(s->fold_xm, tab) || !s->alloc(s->fold_xn, tab)) return nullptr;
    }
    if (m->precision == MGB_PREC_BF16 && batch >= 16) {
        const size_t lb = lt_batch_scratch_bytes(*m, batch);
        char * lp = nullptr;
        if (!s->alloc(lp, lb)) return nullptr;
        s->lt_scratch = lp; s->lt_scratch_bytes = lb;
    }
    // batch 1: persistent cooperative megakernel (MGB_NO_MEGA=1 keeps the per-op kernels, e.g. for A/B tests)
    if (batch == 1 && getenv("MGB_NO_MEGA") == nullptr && hp.dec_layers <= kMegaMaxLayers && hp.d_model <= 1024 &&
        hp.d_ffn <= 11 * 1024 && max_text <= 4096) {
        const int g = mega_max_grid(m->precision);
        if (g > 0) {
            s->n_split = std::max(1, std::min(16, g / hp.dec_sa_heads));
            if (s->alloc(s->attn_part, (size_t)hp.dec_sa_heads * s->n_split * 66) && s->alloc(s->d_barrier, 1) &&
                (getenv("MGB_MEGA_DBG") == nullptr || s->alloc(s->d_dbg, 1024 + 160 * 100))) s->mega_grid = g;
            else return nullptr;
        }
    }
    // batch 1, bf16, Magpie-357M shapes: the whole frame loop is one persistent kernel (MGB_NO_LOOPK=1 disables it)
    if (batch == 1 && m->precision == MGB_PREC_BF16 && getenv("MGB_NO_LOOPK") == nullptr &&
        frame_loop_shape_ok(hp.d_model, hp.d_ffn, hp.dec_sa_heads, hp.lt_dim, hp.lt_ffn_dim, hp.vocab_per_cb, hp.dec_layers) &&
        dxa == 128 && m->lt_in_table[0] && m->lt_qkvo && m->lt_qkv_tab) {
        const int g = frame_loop_max_grid();
        if (g > 0) {
            frame_loop_xchg_layout(hp.vocab_per_cb, s->xoff);
            const size_t xb = frame_loop_xchg_bytes(s->xoff);
            char * xp = nullptr;
            s->loop_cap = std::min(max_text, kLoopMaxCtx);
            const size_t tab = (size_t)L * s->loop_cap * d;
            if (!s->alloc(xp, xb) || !s->alloc(s->d_seq, 1) || !s->alloc(s->d_result, 16) || !s->alloc(s->d_xm, tab) || !s->alloc(s->d_xn, tab) ||
                (getenv("MGB_LOOP_DBG") != nullptr && !s->alloc(s->d_loop_dbg, kLoopDbgStamps))) return nullptr;
            s->d_xbuf = (uint4 *)xp;
            const unsigned one = 1;
            if (cudaMemset(xp, 0, xb) != cudaSuccess || cudaMemcpy(s->d_seq, &one, 4, cudaMemcpyHostToDevice) != cudaSuccess ||
                (s->d_loop_dbg && cudaMemset(s->d_loop_dbg, 0, kLoopDbgStamps * 8) != cudaSuccess)) { set_error("loop kernel state init failed"); return nullptr; }
            s->loop_grid = g;
        }
    }
    std::vector<int32_t> ids(batch);
    for (int b = 0; b < batch; b++) ids[b] = b;
    if (cudaMemcpy(s->dec_utt, ids.data(), batch * 4, cudaMemcpyHostToDevice) != cudaSuccess) { set_error("H2D failed"); return nullptr; }
    return reinterpret_cast<mgb_session *>(s.release());
}

void mgb_session_free(mgb_session * s) { delete reinterpret_cast<Session *>(s); }
int mgb_session_kv_pages(const mgb_session * ss, int32_t * total, int32_t * in_use) {
    const Session * s = reinterpret_cast<const Session *>(ss);
    if (!s) return MGB_EINVAL;
    if (total) *total = s->n_pages;
    if (in_use) *in_use = s->n_pages - (int)s->free_pages.size();
    return MGB_OK;
}
int mgb_session_batch(const mgb_session * s) { return s ? reinterpret_cast<const Session *>(s)->B : MGB_EINVAL; }
int mgb_session_max_seq(const mgb_session * s) { return s ? reinterpret_cast<const Session *>(s)->max_seq : MGB_EINVAL; }
int mgb_session_positions(const mgb_session * ss, int32_t * pos_out) {
    const Session * s = reinterpret_cast<const Session *>(ss);
    if (!s || !pos_out) return MGB_EINVAL;
    for (int b = 0; b < s->B; b++) pos_out[b] = s->pos;
    return MGB_OK;
}
int mgb_session_debug_stamps(mgb_session * ss, uint64_t * out, int n) {
    Session * s = reinterpret_cast<Session *>(ss);
    if (!s || !out || n <= 0) return MGB_EINVAL;
    const unsigned long long * src = s->d_loop_dbg ? s->d_loop_dbg : s->d_dbg;      // MGB_LOOP_DBG takes precedence
    if (!src || n > (s->d_loop_dbg ? kLoopDbgStamps : 1024 + 160 * 100)) return MGB_EINVAL;
    if (cudaSetDevice(s->m->device) != cudaSuccess || cudaStreamSynchronize(s->stream) != cudaSuccess ||
        cudaMemcpy(out, src, (size_t)n * 8, cudaMemcpyDeviceToHost) != cudaSuccess) return MGB_ECUDA;
    return MGB_OK;
}
float mgb_session_last_loop_ms(const mgb_session * s) { return s ? reinterpret_cast<const Session *>(s)->last_ms : 0.0f; }
int64_t mgb_session_last_loop_launches(const mgb_session * s) { return s ? reinterpret_cast<const Session *>(s)->last_launches : 0; }

}  // extern "C"

namespace mgb {

// Text encoder for the utterances in `slots` (session slot ids, ascending); tokens[i] / n_tokens[i] belong to slots[i].  The
// compact encoder output rows of exactly these utterances are left in s->enc_out (offsets h_enc_off by list position) and the
// tok_* maps describe them, for the prefill that follows.  enc_out (optional, host): [len(slots)][max_text][d].
static int encode_slots(Session * s, const std::vector<int> & slots, const int32_t * const * tokens, const int32_t * n_tokens, float * enc_out) {
    Model & m = *s->m; const mgb_hparams & hp = m.hp; const int d = hp.d_model;
    std::vector<int32_t> utt, pos, slot, tok;
    const int ns = (int)slots.size();
    if ((int)s->h_ntext.size() != s->B) s->h_ntext.assign(s->B, 1);
    s->h_enc_off.assign(ns + 1, 0);
    s->enc_slots = slots;
    for (int i = 0; i < ns; i++) {
        const int b = slots[i], n = n_tokens[i];
        if (n <= 0 || n > s->max_text) { set_error("magpie_encode_text: token count out of range"); return MGB_EINVAL; }
        s->h_ntext[b] = n;
        s->h_enc_off[i + 1] = s->h_enc_off[i] + n;
        for (int t = 0; t < n; t++) {
            const int32_t id = tokens[i][t];
            if (id < 0 || id >= hp.text_vocab_size) { set_error("magpie_encode_text: token id out of range"); return MGB_EINVAL; }
            utt.push_back(b); pos.push_back(t); slot.push_back(b * s->max_text + t); tok.push_back(id);
        }
    }
    const int M = (int)tok.size();
    s->M_enc = M;
    cudaStream_t st = s->stream;
    auto up = [&](int32_t * dst, const std::vector<int32_t> & v) {
        return cudaMemcpyAsync(dst, v.data(), v.size() * 4, cudaMemcpyHostToDevice, st) == cudaSuccess;
    };
    if (!up(s->tok_utt, utt) || !up(s->tok_pos, pos) || !up(s->tok_slot, slot) || !up(s->d_tokens, tok) || !up(s->d_ntext, s->h_ntext)) {
        set_error("magpie_encode_text: H2D failed"); return MGB_ECUDA;
    }
    Tokens T; T.M = M; T.utt = s->tok_utt; T.pos = s->tok_pos; T.slot = s->tok_slot;
    if (!launch_text_embed(m, s->d_tokens, T, s->x, st)) return MGB_ECUDA;
    for (int l = 0; l < hp.enc_layers; l++) {
        const EncLayer & L = m.enc[l];
        LinearArgs a;
        a.tc_scratch = s->tc_scratch; a.tc_scratch_bytes = s->tc_scratch_bytes;
        a.precision = m.precision; a.eps = hp.eps; a.M = M; a.W = L.qkv; a.X = s->x; a.ldx = d; a.ln_w = L.norm_self;
        a.Y = s->qbuf; a.ldy = d; a.n_q = d; a.dkv = d; a.kdst = s->ek; a.vdst = s->ev; a.tok_slot = T.slot;
        if (!launch_linear(a, st)) return MGB_ECUDA;
        AttnArgs at;     // the NeMo text encoder is causal (magpie.cpp:1948)
        at.precision = m.precision; at.q = s->qbuf; at.ldq = d; at.K = s->ek; at.V = s->ev; at.rows_per_utt = s->max_text;
        at.H = hp.enc_heads; at.dh = d / hp.enc_heads; at.causal = 1; at.tok = T; at.out = s->attn; at.ldo = d;
        if (!launch_attention(at, st)) return MGB_ECUDA;
        LinearArgs o;
        o.tc_scratch = s->tc_scratch; o.tc_scratch_bytes = s->tc_scratch_bytes;
        o.precision = m.precision; o.M = M; o.W = L.o; o.X = s->attn; o.ldx = d; o.res = s->x; o.ldr = d; o.Y = s->x; o.ldy = d;
        if (!launch_linear(o, st)) return MGB_ECUDA;
        LinearArgs f1;   // causal conv k=3 as 3 shifted taps (magpie.cpp:1825-1914)
        f1.tc_scratch = s->tc_scratch; f1.tc_scratch_bytes = s->tc_scratch_bytes;
        f1.precision = m.precision; f1.eps = hp.eps; f1.gelu_f16 = m.gelu_f16; f1.M = M; f1.W = L.ff1; f1.X = s->x; f1.ldx = d;
        f1.ln_w = L.norm_ff; f1.act = ACT_GELU; f1.Y = s->ffh; f1.ldy = hp.d_ffn; f1.tok_pos = T.pos;
        if (!launch_linear(f1, st)) return MGB_ECUDA;
        LinearArgs f2;
        f2.tc_scratch = s->tc_scratch; f2.tc_scratch_bytes = s->tc_scratch_bytes;
        f2.precision = m.precision; f2.M = M; f2.W = L.ff2; f2.X = s->ffh; f2.ldx = hp.d_ffn; f2.res = s->x; f2.ldr = d;
        f2.Y = s->x; f2.ldy = d; f2.tok_pos = T.pos;
        if (!launch_linear(f2, st)) return MGB_ECUDA;
    }
    if (!launch_layer_norm(s->x, m.enc_norm_out, hp.eps, M, d, s->enc_out, st)) return MGB_ECUDA;
    if (enc_out) {
        memset(enc_out, 0, (size_t)ns * s->max_text * d * sizeof(float));
        for (int i = 0; i < ns; i++)
            if (cudaMemcpyAsync(enc_out + (size_t)i * s->max_text * d, s->enc_out + (size_t)s->h_enc_off[i] * d,
                                (size_t)n_tokens[i] * d * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess) {
                set_error("magpie_encode_text: D2H failed"); return MGB_ECUDA;
            }
    }
    if (cudaStreamSynchronize(st) != cudaSuccess) { set_error(std::string("magpie_encode_text: ") + cudaGetErrorString(cudaGetLastError())); return MGB_ECUDA; }
    return MGB_OK;
}

}  // namespace mgb

extern "C" {

int mgb_encode_text(mgb_session * ss, const int32_t * tokens, const int32_t * n_tokens, float * enc_out) {
    Session * s = reinterpret_cast<Session *>(ss);
    if (!check_ready(s, false)) return MGB_EINVAL;
    if (!tokens || !n_tokens) { set_error("magpie_encode_text: null tokens"); return MGB_EINVAL; }
    std::vector<int> slots(s->B);
    std::vector<const int32_t *> rows(s->B);
    for (int b = 0; b < s->B; b++) { slots[b] = b; rows[b] = tokens + (size_t)b * s->max_text; }
    const int rc = encode_slots(s, slots, rows.data(), n_tokens, enc_out);
    if (rc != MGB_OK) return rc;
    s->encoded = true; s->prefilled = false; s->fold_ready = false;
    return MGB_OK;
}

}  // extern "C"

namespace mgb {

// Steps 2-5 of magpie_synthesize_codes_graph_reuse for the utterances whose encoder output is resident (s->enc_slots, left by
// encode_slots): cross-attention K/V, folded tables, fresh cache pages, 110-frame context prefill, decode position = C.  The
// other slots of the session are not touched (continuous batching refills single slots while the rest keep their state).
static int prefill_slots(Session * s, const int32_t * speakers /* by list position, or null */, bool whole_session) {
    Model & m = *s->m; const mgb_hparams & hp = m.hp;
    const int d = hp.d_model, dxa = hp.dec_xa_heads * hp.dec_xa_d_head, C = hp.context_frames;
    const std::vector<int> & slots = s->enc_slots;
    const int ns = (int)slots.size();
    if ((int)s->h_speakers.size() != s->B) s->h_speakers.assign(s->B, 0);
    for (int i = 0; i < ns; i++) {
        const int v = speakers ? speakers[i] : 0;
        if (v < 0 || v >= hp.num_speakers) { set_error("mgb_prefill: speaker id out of range"); return MGB_EINVAL; }
        s->h_speakers[slots[i]] = v;
    }
    cudaStream_t st = s->stream;
    if (cudaMemcpyAsync(s->d_speakers, s->h_speakers.data(), s->B * 4, cudaMemcpyHostToDevice, st) != cudaSuccess) { set_error("H2D failed"); return MGB_ECUDA; }
    // per-layer cross-attention K/V from the (still resident) encoder output; tok_* hold the encoder map
    const size_t xkv_layer = (size_t)s->B * s->max_text * dxa * m.wsize;
    for (int l = 0; l < hp.dec_layers; l++) {
        const DecLayer & L = m.dec[l];
        LinearArgs a;
        a.tc_scratch = s->tc_scratch; a.tc_scratch_bytes = s->tc_scratch_bytes;
        a.precision = m.precision; a.eps = hp.eps; a.M = s->M_enc; a.W = L.xkv; a.X = s->enc_out; a.ldx = d; a.ln_w = L.norm_xa_mem;
        a.n_q = 0; a.dkv = dxa; a.kdst = (char *)s->xk + l * xkv_layer; a.vdst = (char *)s->xv + l * xkv_layer; a.tok_slot = s->tok_slot;
        a.Y = s->qbuf; a.ldy = d;
        if (!launch_linear(a, st)) return MGB_ECUDA;
    }
    // persistent loop kernel: fold q_net / o_net into per-token tables (frame_loop.cu)
    if (whole_session) {
        s->loop_tables = false;
        if (s->loop_grid > 0 && s->h_ntext[0] <= s->loop_cap) {
            s->loop_E = s->h_ntext[0];
            for (int l = 0; l < hp.dec_layers; l++) {
                const DecLayer & L = m.dec[l];
                if (!launch_xattn_fold((char *)s->xk + l * xkv_layer, (char *)s->xv + l * xkv_layer, L.xq.w, L.xo.w, s->loop_E, d, dxa,
                                       1.0f / sqrtf((float)dxa), s->d_xm + (size_t)l * s->loop_cap * d, s->d_xn + (size_t)l * s->loop_cap * d, st)) return MGB_ECUDA;
            }
            s->loop_tables = true;
        }
    }
    // batched decode: the same fold for every (utterance, text position) row of the cross K/V
    // (the folded kernel streams 2 x E x d f32 table rows per utterance and layer, a quarter per CTA of a 4-CTA cluster:
    //  measured faster than the q GEMM + attention + o GEMM launches up to the 80-token texts of config 4; very long texts
    //  keep the unfolded kernels)
    if (whole_session) {
        s->fold_ready = false;
        long long sum_e = 0;
        for (int b = 0; b < s->B; b++) sum_e += s->h_ntext[b];
        s->fold_ready = s->fold_xm && (sum_e <= (long long)(getenv("MGB_XFOLD_MAXE") ? atoi(getenv("MGB_XFOLD_MAXE")) : 128) * s->B);
    }
    if (s->fold_ready) {
        const size_t tab = (size_t)s->B * s->max_text * d;
        for (int l = 0; l < hp.dec_layers; l++) {
            const DecLayer & L = m.dec[l];
            if (whole_session) {
                if (!launch_xattn_fold((char *)s->xk + l * xkv_layer, (char *)s->xv + l * xkv_layer, L.xq.w, L.xo.w, s->B * s->max_text, d, dxa,
                                       1.0f / sqrtf((float)dxa), s->fold_xm + l * tab, s->fold_xn + l * tab, st, s->d_ntext, s->max_text)) return MGB_ECUDA;
            } else {
                for (int b : slots) {          // only the refilled utterances' rows
                    const size_t ro = (size_t)b * s->max_text;
                    if (!launch_xattn_fold((char *)s->xk + l * xkv_layer + ro * dxa * m.wsize, (char *)s->xv + l * xkv_layer + ro * dxa * m.wsize,
                                           L.xq.w, L.xo.w, s->max_text, d, dxa, 1.0f / sqrtf((float)dxa), s->fold_xm + l * tab + ro * d,
                                           s->fold_xn + l * tab + ro * d, st, s->d_ntext + b, s->max_text)) return MGB_ECUDA;
                }
            }
        }
    }
    // context prefill: ns*C tokens, one batched causal pass (magpie.cpp:4170-4238)
    const int M = ns * C;
    std::vector<int32_t> utt(M), pos(M), slot(M);
    for (int b : slots) {
        if (s->pages_on_demand) release_pages(*s, b);      // a new utterance starts with an empty cache
        if (!ensure_pages(*s, b, C + 1)) return MGB_ERANGE;
    }
    if (!flush_page_table(*s)) return MGB_ECUDA;
    for (int i = 0; i < ns; i++)
        for (int c = 0; c < C; c++) { utt[i * C + c] = slots[i]; pos[i * C + c] = c; slot[i * C + c] = cache_row(*s, slots[i], c); }
    if (cudaStreamSynchronize(st) != cudaSuccess) { set_error("mgb_prefill: xattn K/V failed"); return MGB_ECUDA; }
    if (cudaMemcpyAsync(s->tok_utt, utt.data(), M * 4, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(s->tok_pos, pos.data(), M * 4, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(s->tok_slot, slot.data(), M * 4, cudaMemcpyHostToDevice, st) != cudaSuccess) { set_error("H2D failed"); return MGB_ECUDA; }
    Tokens T; T.M = M; T.utt = s->tok_utt; T.pos = s->tok_pos; T.slot = s->tok_slot;
    if (!launch_context_embed(m, s->d_speakers, T, s->x, st)) return MGB_ECUDA;
    s->prefill_len = C;
    const bool pf_ok = decoder_layers(*s, T, false);
    s->prefill_len = 0;
    if (!pf_ok) return MGB_ECUDA;
    for (int b : slots) {
        const int32_t p0 = C, s0 = cache_row(*s, b, C);
        if (cudaMemcpyAsync(s->dec_pos + b, &p0, 4, cudaMemcpyHostToDevice, st) != cudaSuccess ||
            cudaMemcpyAsync(s->dec_slot + b, &s0, 4, cudaMemcpyHostToDevice, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess) { set_error("H2D failed"); return MGB_ECUDA; }      // (p0 / s0 are stack values: sync before they go out of scope)
    }
    if (cudaStreamSynchronize(st) != cudaSuccess) { set_error(std::string("mgb_prefill: ") + cudaGetErrorString(cudaGetLastError())); return MGB_ECUDA; }
    return MGB_OK;
}

}  // namespace mgb

extern "C" {

int mgb_prefill(mgb_session * ss, const int32_t * speakers) {
    Session * s = reinterpret_cast<Session *>(ss);
    if (!check_ready(s, false)) return MGB_EINVAL;
    if (!s->encoded) { set_error("mgb_prefill: call mgb_encode_text first"); return MGB_EINVAL; }
    const int rc = prefill_slots(s, speakers, true);
    if (rc != MGB_OK) return rc;
    s->pos = s->m->hp.context_frames; s->prefilled = true; s->encoded = false;      // tok_* now hold the prefill map
    return MGB_OK;
}

int mgb_decoder_step(mgb_session * ss, const int32_t * codes, float * hidden_out) {
    Session * s = reinterpret_cast<Session *>(ss);
    if (!check_ready(s, true)) return MGB_EINVAL;
    if (s->pos + 1 > s->max_seq) { set_error("mgb_decoder_step: KV cache full"); return MGB_ERANGE; }
    const mgb_hparams & hp = s->m->hp;
    cudaStream_t st = s->stream;
    if (codes) {
        for (int i = 0; i < s->B * 8; i++)
            if (codes[i] < 0 || codes[i] >= hp.vocab_per_cb) { set_error("mgb_decoder_step: code out of range"); return MGB_EINVAL; }
        if (cudaMemcpyAsync(s->d_codes, codes, (size_t)s->B * 8 * 4, cudaMemcpyHostToDevice, st) != cudaSuccess) { set_error("H2D failed"); return MGB_ECUDA; }
    }
    s->attn_split = attention_plan_kv_split(s->m->hp.dec_sa_heads * s->B, s->pos + 1);
    if (!ensure_pages_all(*s, std::min(s->pos + 2, s->max_seq))) return MGB_ERANGE;       // this position and the next (advance reads its page)
    if (s->mega_grid > 0 && !pages_contiguous(*s, s->pos + 1)) { set_error("mgb_decoder_step: the batch-1 kernel needs contiguous cache pages"); return MGB_ERANGE; }
    if (!decoder_step_device(*s) || !launch_advance(*s, false)) return MGB_ECUDA;
    if (hidden_out && cudaMemcpyAsync(hidden_out, s->hidden, (size_t)s->B * hp.d_model * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess) {
        set_error("D2H failed"); return MGB_ECUDA;
    }
    if (cudaStreamSynchronize(st) != cudaSuccess) { set_error(std::string("mgb_decoder_step: ") + cudaGetErrorString(cudaGetLastError())); return MGB_ECUDA; }
    s->pos++;
    return MGB_OK;
}

int mgb_final_proj(mgb_session * ss, const float * hidden, float * logits_out) {
    Session * s = reinterpret_cast<Session *>(ss);
    if (!check_ready(s, false) || !logits_out) return MGB_EINVAL;
    Model & m = *s->m; const mgb_hparams & hp = m.hp;
    const int N = hp.num_codebooks * hp.vocab_per_cb, d = hp.d_model;
    if (!m.final_w.w || !m.final_b) { set_error("mgb_final_proj: the model file has no final_proj.weight / final_proj.bias"); return MGB_EINVAL; }
    cudaStream_t st = s->stream;
    const float * hsrc = s->hidden;
    if (hidden) {
        if (cudaMemcpyAsync(s->attn, hidden, (size_t)s->B * d * 4, cudaMemcpyHostToDevice, st) != cudaSuccess) { set_error("H2D failed"); return MGB_ECUDA; }
        hsrc = s->attn;
    }
    // output [B][N] does not fit the step buffers: use the loop logits buffer
    if (!grow((void **)&s->l_logits, &s->l_cap_logits, (size_t)s->B * N * 4)) return MGB_ECUDA;
    LinearArgs a;
    a.tc_scratch = s->tc_scratch; a.tc_scratch_bytes = s->tc_scratch_bytes;
    a.precision = m.precision; a.M = s->B; a.W = m.final_w; a.X = hsrc; a.ldx = d; a.bias = m.final_b; a.Y = s->l_logits; a.ldy = N;
    if (!launch_linear(a, st)) return MGB_ECUDA;
    if (cudaMemcpyAsync(logits_out, s->l_logits, (size_t)s->B * N * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) { set_error("mgb_final_proj: copy/sync failed"); return MGB_ECUDA; }
    return MGB_OK;
}

int mgb_lt_sample(mgb_session * ss, const float * hidden, float temperature, int top_k, const uint8_t * forbid_eos,
                  const int32_t * forced_codes, const float * uniforms, uint64_t seed,
                  int32_t * sampled, int32_t * argmax, float * logits_out) {
    Session * s = reinterpret_cast<Session *>(ss);
    if (!check_ready(s, false)) return MGB_EINVAL;
    Model & m = *s->m; const mgb_hparams & hp = m.hp;
    const int d = hp.d_model, V = hp.vocab_per_cb, B = s->B;
    cudaStream_t st = s->stream;
    LtArgs a;
    a.B = B; a.hidden = s->hidden; a.temperature = temperature; a.top_k = top_k; a.seed = seed; a.step = (uint32_t)s->pos;
    bool ok = true;
    if (hidden) { ok = ok && cudaMemcpyAsync(s->attn, hidden, (size_t)B * d * 4, cudaMemcpyHostToDevice, st) == cudaSuccess; a.hidden = s->attn; }
    if (forbid_eos) { ok = ok && cudaMemcpyAsync(s->d_forbid, forbid_eos, B, cudaMemcpyHostToDevice, st) == cudaSuccess; a.forbid_eos = s->d_forbid; }
    if (forced_codes) {
        for (int i = 0; i < B * 8; i++)
            if (forced_codes[i] < 0 || forced_codes[i] >= V) { set_error("mgb_lt_sample: forced code out of range"); return MGB_EINVAL; }
        ok = ok && cudaMemcpyAsync(s->d_forced, forced_codes, (size_t)B * 8 * 4, cudaMemcpyHostToDevice, st) == cudaSuccess; a.forced = s->d_forced;
    }
    if (uniforms) { ok = ok && cudaMemcpyAsync(s->d_uniforms, uniforms, (size_t)B * 8 * 4, cudaMemcpyHostToDevice, st) == cudaSuccess; a.uniforms = s->d_uniforms; }
    if (!ok) { set_error("mgb_lt_sample: H2D failed"); return MGB_ECUDA; }
    a.sampled = s->d_sampled; a.argmax = s->d_argmax; a.next_codes = s->d_codes; a.logits = logits_out ? s->d_logits1 : nullptr;
    a.lt_scratch = s->lt_scratch; a.lt_scratch_bytes = s->lt_scratch_bytes;
    if (!launch_local_transformer(m, a, st)) return MGB_ECUDA;
    if (sampled) ok = ok && cudaMemcpyAsync(sampled, s->d_sampled, (size_t)B * 8 * 4, cudaMemcpyDeviceToHost, st) == cudaSuccess;
    if (argmax) ok = ok && cudaMemcpyAsync(argmax, s->d_argmax, (size_t)B * 8 * 4, cudaMemcpyDeviceToHost, st) == cudaSuccess;
    if (logits_out) ok = ok && cudaMemcpyAsync(logits_out, s->d_logits1, (size_t)B * 8 * V * 4, cudaMemcpyDeviceToHost, st) == cudaSuccess;
    if (!ok || cudaStreamSynchronize(st) != cudaSuccess) { set_error(std::string("mgb_lt_sample: ") + cudaGetErrorString(cudaGetLastError())); return MGB_ECUDA; }
    return MGB_OK;
}

}  // extern "C"

// ---- device-resident loops ---------------------------------------------------------------------------
namespace mgb {

struct LoopCfg {
    int T = 0;                 // steps (array stride)
    float temperature = 0.0f; int top_k = 80; uint64_t seed = 0;
    bool teacher = false, want_logits = false, want_hidden = false, ignore_eos = false;
    bool have_uniforms = false;
    int n = -1, step0 = 0;     // persistent batch-1 kernel only: run n steps (default T) starting at step / output row step0 (streaming chunks)
};

// One loop iteration = decoder step on d_codes -> LT (+sampling) -> advance; all per-step indexing is
// done on the device from d_step, so the iteration is captured once as a CUDA graph and replayed.
static bool enqueue_iteration(Session & s, const LoopCfg & c) {
    if (!decoder_step_device(s)) return false;
    LtArgs a;
    a.B = s.B; a.hidden = s.hidden; a.temperature = c.temperature; a.top_k = c.top_k; a.seed = c.seed;
    a.forced = c.teacher ? s.l_forced : nullptr;
    a.uniforms = c.have_uniforms ? s.l_uniforms : nullptr;
    a.sampled = s.l_sampled; a.argmax = s.l_argmax; a.next_codes = s.d_codes;
    a.logits = c.want_logits ? s.l_logits : nullptr;
    a.d_step = s.d_step; a.T_total = c.T; a.min_frames = c.teacher ? 0 : 4;      // magpie.cpp:4267, 4325
    a.done_step = s.d_done; a.hidden_hist = c.want_hidden ? s.l_hidden : nullptr;
    a.lt_scratch = s.lt_scratch; a.lt_scratch_bytes = s.lt_scratch_bytes;
    if (!launch_local_transformer(*s.m, a, s.stream)) return false;
    return launch_advance(s, true);
}

// batch 1 / bf16: decoder step + local transformer + sampling + EOS for all frames in ONE persistent kernel
static int run_loop_persistent(Session & s, const LoopCfg & c, int * steps_run) {
    Model & m = *s.m; const mgb_hparams & hp = m.hp;
    cudaStream_t st = s.stream;
    FrameLoopParams p = {};
    const size_t tab = (size_t)s.loop_cap * hp.d_model;
    for (int l = 0; l < hp.dec_layers; l++) {
        const DecLayer & L = m.dec[l];
        p.layer[l] = LoopLayer{L.qkv.w, L.o.w, L.ff1.w, L.ff2.w, L.norm_self, L.norm_xa_q, L.norm_ff, s.d_xm + l * tab, s.d_xn + l * tab};
    }
    p.L = hp.dec_layers; p.E = s.loop_E; p.eps = hp.eps; p.gelu_f16 = m.gelu_f16;
    for (int cb = 0; cb < 8; cb++) {
        p.audio_emb[cb] = m.audio_emb[cb]; p.lt_out_w[cb] = m.lt_out_w[cb].w; p.lt_out_b[cb] = m.lt_out_b[cb]; p.lt_in_table[cb] = m.lt_in_table[cb];
    }
    p.dec_pos = m.dec_pos; p.norm_out = m.dec_norm_out;
    p.kcache = s.kc; p.vcache = s.vc; p.kv_layer_stride = s.kv_rows * hp.d_model;
    p.V = hp.vocab_per_cb;
    p.lt_in_w = m.lt_in_w.w; p.lt_in_b = m.lt_in_b; p.lt_pos = m.lt_pos; p.lt_norm_self = m.lt_norm_self; p.lt_norm_ff = m.lt_norm_ff;
    p.lt_qkvo = m.lt_qkvo; p.lt_qkv_tab = m.lt_qkv_tab; p.lt_ff1 = m.lt_ff1.w; p.lt_ff2 = m.lt_ff2.w;
    p.n_steps = c.n >= 0 ? c.n : c.T; p.pos0 = s.pos; p.step0 = c.step0; p.row0 = c.step0; p.min_frames = c.teacher ? 0 : 4;       // magpie.cpp:4267, 4325
    p.teacher = c.teacher ? 1 : 0; p.ignore_eos = c.ignore_eos ? 1 : 0;
    p.temperature = c.temperature; p.top_k = c.top_k; p.seed = c.seed;
    p.uniforms = c.have_uniforms ? s.l_uniforms : nullptr;
    p.forced = c.teacher ? s.l_forced : nullptr;
    p.codes_io = s.d_codes; p.bos_id = hp.audio_bos_id; p.eos_id = hp.audio_eos_id;
    p.sampled = s.l_sampled; p.argmax = s.l_argmax; p.logits = c.want_logits ? s.l_logits : nullptr;
    p.hidden_hist = c.want_hidden ? s.l_hidden : nullptr; p.hidden_last = s.hidden;
    p.result = s.d_result; p.xbuf = s.d_xbuf; memcpy(p.xoff, s.xoff, sizeof(p.xoff)); p.seq = s.d_seq; p.dbg = s.d_loop_dbg;
    p.dbg_flags = getenv("MGB_LOOP_FLAGS") ? atoi(getenv("MGB_LOOP_FLAGS")) : 0;
    p.max_split = getenv("MGB_LOOP_MAXSPLIT") ? std::max(1, std::min(6, atoi(getenv("MGB_LOOP_MAXSPLIT")))) : 6;
    p.no_defer_amax = getenv("MGB_LOOP_NO_DEFER") != nullptr ? 1 : 0;
    const int64_t l0 = g_launch_counter;
    cudaEventRecord(s.ev0, st);
    if (!launch_frame_loop(p, s.loop_grid, st)) return MGB_ECUDA;
    cudaEventRecord(s.ev1, st);
    int32_t res[10] = {};
    if (cudaMemcpyAsync(res, s.d_result, sizeof(res), cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) {
        set_error(std::string("generation loop (persistent kernel): ") + cudaGetErrorString(cudaGetLastError())); return MGB_ECUDA;
    }
    cudaEventElapsedTime(&s.last_ms, s.ev0, s.ev1);
    s.last_launches = g_launch_counter - l0;
    const int t = res[0];
    s.pos += t;
    // keep the per-op step path (mgb_decoder_step) usable afterwards: position, slot, next codes, EOS record
    const int32_t pos = s.pos, slot = s.pos;
    if (cudaMemcpyAsync(s.dec_pos, &pos, 4, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(s.dec_slot, &slot, 4, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(s.d_codes, res + 2, 32, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(s.d_done, res + 1, 4, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) { set_error("generation loop: state write-back failed"); return MGB_ECUDA; }
    *steps_run = t;
    return MGB_OK;
}

static int run_loop(Session & s, const LoopCfg & c, int * steps_run) {
    cudaStream_t st = s.stream;
    const int B = s.B;
    if (s.pos + c.T > s.max_seq) { set_error("generation loop: KV cache too small for the requested steps"); return MGB_ERANGE; }
    // (the persistent kernel divides the cached keys over up to 6 CTAs per head; each warp scans 32-key chunks, the first
    //  480 keys of an item before the wait for q, longer items in further rounds)
    if (s.loop_grid > 0 && s.loop_tables) {
        if (!ensure_pages_all(s, std::min(s.pos + c.T + 1, s.max_seq))) return MGB_ERANGE;
        if (!pages_contiguous(s, s.pos + c.T)) { set_error("generation loop: the batch-1 kernel needs contiguous cache pages"); return MGB_ERANGE; }
        return run_loop_persistent(s, c, steps_run);
    }
    if (s.mega_grid > 0 && (!ensure_pages_all(s, std::min(s.pos + c.T + 1, s.max_seq)) || !pages_contiguous(s, s.pos + c.T))) {
        set_error("generation loop: the batch-1 kernel needs contiguous cache pages"); return MGB_ERANGE;
    }
    // the iteration is captured once, so the key split of the self-attention is planned for the KV length the loop will reach
    s.attn_split = attention_plan_kv_split(s.m->hp.dec_sa_heads * B, s.pos + c.T);
    std::vector<int32_t> neg(B, -1);
    if (cudaMemsetAsync(s.d_step, 0, 4, st) != cudaSuccess ||
        cudaMemcpyAsync(s.d_done, neg.data(), B * 4, cudaMemcpyHostToDevice, st) != cudaSuccess) { set_error("loop init failed"); return MGB_ECUDA; }
    const bool use_graph = getenv("MGB_NO_GRAPH") == nullptr;
    cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
    const int64_t launches0 = g_launch_counter;
    int64_t per_iter = 0;
    if (use_graph) {
        if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { set_error("graph capture begin failed"); return MGB_ECUDA; }
        const bool okq = enqueue_iteration(s, c);
        const cudaError_t e = cudaStreamEndCapture(st, &graph);
        if (!okq || e != cudaSuccess || !graph) { if (graph) cudaGraphDestroy(graph); if (okq) set_error("graph capture failed"); return MGB_ECUDA; }
        if (cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) { cudaGraphDestroy(graph); set_error("graph instantiate failed"); return MGB_ECUDA; }
        per_iter = g_launch_counter - launches0;
    }
    int rc = MGB_OK, t = 0;
    const int check_every = 8;
    std::vector<int32_t> done(B);
    cudaEventRecord(s.ev0, st);
    while (t < c.T) {
        const int n = std::min(check_every, c.T - t);
        // cache pages for the positions this chunk writes and the one after it (pools sized on demand grow here)
        if (!ensure_pages_all(s, std::min(s.pos + t + n + 1, s.max_seq))) { rc = MGB_ERANGE; break; }
        for (int i = 0; i < n && rc == MGB_OK; i++) {
            if (use_graph) { if (cudaGraphLaunch(exec, st) != cudaSuccess) { set_error("graph launch failed"); rc = MGB_ECUDA; } }
            else if (!enqueue_iteration(s, c)) rc = MGB_ECUDA;
        }
        if (rc != MGB_OK) break;
        t += n;
        if (!c.teacher && !c.ignore_eos) {
            if (cudaMemcpyAsync(done.data(), s.d_done, B * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
                cudaStreamSynchronize(st) != cudaSuccess) { set_error("loop: EOS poll failed"); rc = MGB_ECUDA; break; }
            bool all = true;
            for (int b = 0; b < B; b++) all = all && done[b] >= 0;
            if (all) break;
        }
    }
    cudaEventRecord(s.ev1, st);
    if (rc == MGB_OK && cudaStreamSynchronize(st) != cudaSuccess) {
        set_error(std::string("generation loop: ") + cudaGetErrorString(cudaGetLastError())); rc = MGB_ECUDA;
    }
    if (rc == MGB_OK) cudaEventElapsedTime(&s.last_ms, s.ev0, s.ev1);
    s.last_launches = use_graph ? per_iter * t : g_launch_counter - launches0;
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
    s.pos += t;
    *steps_run = t;
    return rc;
}

static bool prepare_loop_buffers(Session & s, const LoopCfg & c) {
    const size_t n = (size_t)s.B * c.T;
    const mgb_hparams & hp = s.m->hp;
    if (!grow((void **)&s.l_forced, &s.l_cap_forced, n * 32) || !grow((void **)&s.l_sampled, &s.l_cap_sampled, n * 32) ||
        !grow((void **)&s.l_argmax, &s.l_cap_argmax, n * 32)) return false;
    if (c.have_uniforms && !grow((void **)&s.l_uniforms, &s.l_cap_uni, n * 32)) return false;
    if (c.want_logits && !grow((void **)&s.l_logits, &s.l_cap_logits, n * 8 * hp.vocab_per_cb * 4)) return false;
    if (c.want_hidden && !grow((void **)&s.l_hidden, &s.l_cap_hidden, n * hp.d_model * 4)) return false;
    return true;
}

}  // namespace mgb

extern "C" {

int mgb_generate(mgb_session * ss, int max_steps, float temperature, int top_k, const float * uniforms, uint64_t seed,
                 int ignore_eos, int32_t * codes_out, int32_t * n_frames_out, float * hidden_out) {
    Session * s = reinterpret_cast<Session *>(ss);
    if (!check_ready(s, true)) return MGB_EINVAL;
    const mgb_hparams & hp = s->m->hp;
    if (max_steps <= 0) max_steps = hp.max_dec_steps;
    if (!codes_out || !n_frames_out) { set_error("mgb_generate: null outputs"); return MGB_EINVAL; }
    if (temperature >= 0.01f && top_k < 1) { set_error("mgb_generate: top_k must be >= 1 when sampling"); return MGB_EINVAL; }
    LoopCfg c;
    c.T = max_steps; c.temperature = temperature; c.top_k = top_k; c.seed = seed; c.ignore_eos = ignore_eos != 0;
    c.want_hidden = hidden_out != nullptr; c.have_uniforms = uniforms != nullptr;
    if (!prepare_loop_buffers(*s, c)) return MGB_ECUDA;
    cudaStream_t st = s->stream;
    const int B = s->B; const size_t n = (size_t)B * c.T;
    std::vector<int32_t> bos((size_t)B * 8, hp.audio_bos_id);           // BOS frame (magpie.cpp:4271-4275)
    bool ok = cudaMemcpyAsync(s->d_codes, bos.data(), bos.size() * 4, cudaMemcpyHostToDevice, st) == cudaSuccess;
    if (uniforms) ok = ok && cudaMemcpyAsync(s->l_uniforms, uniforms, n * 32, cudaMemcpyHostToDevice, st) == cudaSuccess;
    if (!ok) { set_error("mgb_generate: H2D failed"); return MGB_ECUDA; }
    int steps = 0;
    const int rc = run_loop(*s, c, &steps);
    if (rc != MGB_OK) return rc;
    std::vector<int32_t> done(B);
    ok = cudaMemcpyAsync(codes_out, s->l_sampled, n * 32, cudaMemcpyDeviceToHost, st) == cudaSuccess &&
         cudaMemcpyAsync(done.data(), s->d_done, B * 4, cudaMemcpyDeviceToHost, st) == cudaSuccess;
    if (hidden_out) ok = ok && cudaMemcpyAsync(hidden_out, s->l_hidden, n * hp.d_model * 4, cudaMemcpyDeviceToHost, st) == cudaSuccess;
    if (!ok || cudaStreamSynchronize(st) != cudaSuccess) { set_error("mgb_generate: D2H failed"); return MGB_ECUDA; }
    for (int b = 0; b < B; b++) n_frames_out[b] = (ignore_eos || done[b] < 0) ? steps : std::min(done[b], steps);
    return MGB_OK;
}

int mgb_teacher_forced(mgb_session * ss, const int32_t * codes_in, int T, float * hidden_out, float * lt_logits_out,
                       int32_t * greedy_out) {
    Session * s = reinterpret_cast<Session *>(ss);
    if (!check_ready(s, true)) return MGB_EINVAL;
    const mgb_hparams & hp = s->m->hp;
    if (!codes_in || T <= 0) { set_error("mgb_teacher_forced: invalid arguments"); return MGB_EINVAL; }
    const int B = s->B; const size_t n = (size_t)B * T;
    for (size_t i = 0; i < n * 8; i++)
        if (codes_in[i] < 0 || codes_in[i] >= hp.vocab_per_cb) { set_error("mgb_teacher_forced: code out of range"); return MGB_EINVAL; }
    LoopCfg c;
    c.T = T; c.teacher = true; c.want_logits = lt_logits_out != nullptr; c.want_hidden = hidden_out != nullptr;
    if (!prepare_loop_buffers(*s, c)) return MGB_ECUDA;
    cudaStream_t st = s->stream;
    std::vector<int32_t> bos((size_t)B * 8, hp.audio_bos_id);
    if (cudaMemcpyAsync(s->d_codes, bos.data(), bos.size() * 4, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(s->l_forced, codes_in, n * 32, cudaMemcpyHostToDevice, st) != cudaSuccess) { set_error("mgb_teacher_forced: H2D failed"); return MGB_ECUDA; }
    int steps = 0;
    const int rc = run_loop(*s, c, &steps);
    if (rc != MGB_OK) return rc;
    bool ok = true;
    if (greedy_out) ok = ok && cudaMemcpyAsync(greedy_out, s->l_argmax, n * 32, cudaMemcpyDeviceToHost, st) == cudaSuccess;
    if (hidden_out) ok = ok && cudaMemcpyAsync(hidden_out, s->l_hidden, n * hp.d_model * 4, cudaMemcpyDeviceToHost, st) == cudaSuccess;
    if (lt_logits_out) ok = ok && cudaMemcpyAsync(lt_logits_out, s->l_logits, n * 8 * hp.vocab_per_cb * 4, cudaMemcpyDeviceToHost, st) == cudaSuccess;
    if (!ok || cudaStreamSynchronize(st) != cudaSuccess) { set_error("mgb_teacher_forced: D2H failed"); return MGB_ECUDA; }
    return MGB_OK;
}

}  // extern "C"

// ---- continuous batching ------------------------------------------------------------------------------
namespace mgb {

// after a queue-mode step: an active slot whose frame hit EOS (done_step set by the LT kernel) or that reached its frame limit
// stops; the others move to the next position / cache row and count one more emitted frame
__global__ void advance_queue_kernel(int32_t * pos, int32_t * slot, int B, const int32_t * page_table, int max_pages, int32_t * active,
                                     int32_t * utt_step, const int32_t * done_step, const int32_t * limit) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B || !active[i]) return;
    if (done_step[i] >= 0) { active[i] = 0; return; }                      // EOS frame: not emitted (magpie.cpp:4341-4352)
    const int np = pos[i] + 1, pg = np / kKvPageRows;
    pos[i] = np;
    slot[i] = pg < max_pages ? page_table[(size_t)i * max_pages + pg] * kKvPageRows + np % kKvPageRows : 0;
    utt_step[i] += 1;
    if (utt_step[i] >= limit[i]) active[i] = 0;
}

}  // namespace mgb

extern "C" {

int mgb_generate_queue(mgb_session * ss, int n_utt, const int32_t * tokens, const int32_t * n_tokens, int max_text,
                       const int32_t * speakers, const int32_t * max_steps_per_utt, int max_steps, float temperature, int top_k,
                       uint64_t seed, int32_t * codes_out, int32_t * n_frames_out, int64_t * steps_run_out) {
    Session * s = reinterpret_cast<Session *>(ss);
    if (!check_ready(s, false)) return MGB_EINVAL;
    Model & m = *s->m; const mgb_hparams & hp = m.hp;
    if (n_utt <= 0 || !tokens || !n_tokens || !codes_out || !n_frames_out || max_steps <= 0 || max_text <= 0 || max_text > s->max_text) {
        set_error("mgb_generate_queue: invalid arguments"); return MGB_EINVAL;
    }
    if (temperature >= 0.01f && top_k < 1) { set_error("mgb_generate_queue: top_k must be >= 1 when sampling"); return MGB_EINVAL; }
    const int B = s->B, C = hp.context_frames;
    if (C + max_steps + 1 > s->max_seq) { set_error("mgb_generate_queue: max_steps exceeds the session's max_seq"); return MGB_ERANGE; }
    if (n_utt < B && s->mega_grid == 0) { set_error("mgb_generate_queue: fewer utterances than session slots (use a smaller session)"); return MGB_EINVAL; }
    for (int q = 0; q < n_utt; q++) {
        if (n_tokens[q] <= 0 || n_tokens[q] > max_text) { set_error("mgb_generate_queue: token count out of range"); return MGB_EINVAL; }
        if (max_steps_per_utt && (max_steps_per_utt[q] <= 0 || max_steps_per_utt[q] > max_steps)) { set_error("mgb_generate_queue: per-utterance step limit out of range"); return MGB_EINVAL; }
    }
    cudaStream_t st = s->stream;
    if (s->mega_grid > 0) {
        // one-slot session: the batch-1 kernels advance their own position; utterances simply run one after the other
        int64_t steps = 0;
        cudaEventRecord(s->ev0, st);
        float ms = 0.0f;
        for (int q = 0; q < n_utt; q++) {
            std::vector<int32_t> row((size_t)s->max_text, 0);
            memcpy(row.data(), tokens + (size_t)q * max_text, (size_t)n_tokens[q] * 4);
            const int32_t spk = speakers ? speakers[q] : 0, lim = max_steps_per_utt ? max_steps_per_utt[q] : max_steps;
            int rc = mgb_encode_text(ss, row.data(), n_tokens + q, nullptr);
            if (rc == MGB_OK) rc = mgb_prefill(ss, &spk);
            if (rc == MGB_OK) rc = mgb_generate(ss, lim, temperature, top_k, nullptr, seed + (uint64_t)q, 0, codes_out + (size_t)q * max_steps * 8, n_frames_out + q, nullptr);
            if (rc != MGB_OK) return rc;
            steps += n_frames_out[q]; ms += s->last_ms;
        }
        s->last_ms = ms;
        if (steps_run_out) *steps_run_out = steps;
        return MGB_OK;
    }
    const int T_total = max_steps + 1;                 // one scratch row per slot: finished slots keep being computed
    LoopCfg c;
    c.T = T_total; c.temperature = temperature; c.top_k = top_k; c.seed = seed;
    if (!prepare_loop_buffers(*s, c)) return MGB_ECUDA;
    std::vector<int> slot_utt(B, -1), slot_steps(B, 0);          // utterance in each slot; host's upper bound of its emitted frames
    std::vector<int32_t> h_limit(B, 1), h_active(B, 0), h_done(B, -1), h_ustep(B, 0);
    int32_t * d_limit = s->d_forced;                   // [B][8] scratch: the first B entries hold the frame limits
    int next = 0, finished = 0;
    int64_t steps = 0, launches0 = g_launch_counter;
    s->attn_split = attention_plan_kv_split(hp.dec_sa_heads * B, C + 1 + max_steps);
    const bool use_graph = getenv("MGB_NO_GRAPH") == nullptr;
    cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
    auto cleanup = [&]() { if (exec) cudaGraphExecDestroy(exec); if (graph) cudaGraphDestroy(graph); };
    // one iteration: decoder step on d_codes -> LT (+sampling) with per-utterance step counters -> queue advance
    auto enqueue = [&]() -> bool {
        if (!launch_audio_embed(m, s->d_codes, s->dec_pos, B, s->x, st)) return false;
        Tokens tok; tok.M = B; tok.utt = s->dec_utt; tok.pos = s->dec_pos; tok.slot = s->dec_slot;
        if (!decoder_layers(*s, tok, true)) return false;
        LtArgs a;
        a.B = B; a.hidden = s->hidden; a.temperature = temperature; a.top_k = top_k; a.seed = seed;
        a.sampled = s->l_sampled; a.argmax = s->l_argmax; a.next_codes = s->d_codes;
        a.utt_step = s->d_utt_step; a.T_total = T_total; a.min_frames = 4; a.done_step = s->d_done;
        a.lt_scratch = s->lt_scratch; a.lt_scratch_bytes = s->lt_scratch_bytes;
        if (!launch_local_transformer(m, a, st)) return false;
        advance_queue_kernel<<<(B + 127) / 128, 128, 0, st>>>(s->dec_pos, s->dec_slot, B, s->d_page_table, s->max_pages, s->d_active,
                                                              s->d_utt_step, s->d_done, d_limit);
        MGB_LAUNCH_CHECK();
        return true;
    };
    // (re)fill the free slots from the queue: encoder + cross K/V + context prefill of just those slots
    auto refill = [&](const std::vector<int> & free_slots) -> int {
        std::vector<int> slots; std::vector<const int32_t *> rows; std::vector<int32_t> nt, spk;
        for (int b : free_slots) {
            if (next >= n_utt) break;
            slots.push_back(b); rows.push_back(tokens + (size_t)next * max_text); nt.push_back(n_tokens[next]);
            spk.push_back(speakers ? speakers[next] : 0);
            slot_utt[b] = next; slot_steps[b] = 0;
            h_limit[b] = max_steps_per_utt ? max_steps_per_utt[next] : max_steps;
            next++;
        }
        if (slots.empty()) return MGB_OK;
        int rc = encode_slots(s, slots, rows.data(), nt.data(), nullptr);
        if (rc == MGB_OK) rc = prefill_slots(s, spk.data(), false);
        if (rc != MGB_OK) return rc;
        std::vector<int32_t> bos(8, hp.audio_bos_id);
        for (int b : slots) {
            h_active[b] = 1; h_done[b] = -1; h_ustep[b] = 0;
            if (cudaMemcpyAsync(s->d_codes + (size_t)b * 8, bos.data(), 32, cudaMemcpyHostToDevice, st) != cudaSuccess) { set_error("H2D failed"); return MGB_ECUDA; }
        }
        if (cudaMemcpyAsync(s->d_active, h_active.data(), B * 4, cudaMemcpyHostToDevice, st) != cudaSuccess ||
            cudaMemcpyAsync(s->d_done, h_done.data(), B * 4, cudaMemcpyHostToDevice, st) != cudaSuccess ||
            cudaMemcpyAsync(s->d_utt_step, h_ustep.data(), B * 4, cudaMemcpyHostToDevice, st) != cudaSuccess ||
            cudaMemcpyAsync(d_limit, h_limit.data(), B * 4, cudaMemcpyHostToDevice, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess) { set_error("mgb_generate_queue: slot state upload failed"); return MGB_ECUDA; }
        return MGB_OK;
    };
    // the folded cross-attention tables serve every slot or none: decided once for the whole queue (texts up to 128 tokens)
    s->fold_ready = s->fold_xm != nullptr && max_text <= 128;
    s->loop_tables = false; s->prefilled = false; s->encoded = false;
    cudaEventRecord(s->ev0, st);
    std::vector<int> all(B);
    for (int b = 0; b < B; b++) all[b] = b;
    int rc = refill(all);
    if (rc != MGB_OK) return rc;
    if (use_graph) {
        if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { set_error("graph capture begin failed"); return MGB_ECUDA; }
        const bool okq = enqueue();
        const cudaError_t e = cudaStreamEndCapture(st, &graph);
        if (!okq || e != cudaSuccess || !graph) { cleanup(); if (okq) set_error("graph capture failed"); return MGB_ECUDA; }
        if (cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) { cleanup(); set_error("graph instantiate failed"); return MGB_ECUDA; }
    }
    const int check_every = 8;
    while (finished < n_utt) {
        // cache pages for the positions the next chunk writes (and the one after): on-demand pools grow here
        for (int b = 0; b < B; b++)
            if (h_active[b] && !ensure_pages(*s, b, std::min(C + 1 + slot_steps[b] + check_every + 1, s->max_seq))) { cleanup(); return MGB_ERANGE; }
        if (!flush_page_table(*s)) { cleanup(); return MGB_ECUDA; }
        for (int i = 0; i < check_every; i++) {
            if (use_graph) { if (cudaGraphLaunch(exec, st) != cudaSuccess) { cleanup(); set_error("graph launch failed"); return MGB_ECUDA; } }
            else if (!enqueue()) { cleanup(); return MGB_ECUDA; }
        }
        steps += check_every;
        for (int b = 0; b < B; b++) if (h_active[b]) slot_steps[b] += check_every;
        if (cudaMemcpyAsync(h_active.data(), s->d_active, B * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaMemcpyAsync(h_done.data(), s->d_done, B * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaMemcpyAsync(h_ustep.data(), s->d_utt_step, B * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess) { cleanup(); set_error(std::string("mgb_generate_queue: ") + cudaGetErrorString(cudaGetLastError())); return MGB_ECUDA; }
        std::vector<int> free_slots;
        for (int b = 0; b < B; b++) {
            if (h_active[b] || slot_utt[b] < 0) continue;
            const int q = slot_utt[b];
            const int nf = h_done[b] >= 0 ? std::min(h_done[b], h_limit[b]) : h_limit[b];
            n_frames_out[q] = nf;
            if (nf > 0 && cudaMemcpyAsync(codes_out + (size_t)q * max_steps * 8, s->l_sampled + (size_t)b * T_total * 8, (size_t)nf * 32,
                                          cudaMemcpyDeviceToHost, st) != cudaSuccess) { cleanup(); set_error("mgb_generate_queue: D2H failed"); return MGB_ECUDA; }
            slot_utt[b] = -1; finished++;
            release_pages(*s, b);
            free_slots.push_back(b);
        }
        if (!free_slots.empty()) {
            if (cudaStreamSynchronize(st) != cudaSuccess) { cleanup(); set_error("mgb_generate_queue: D2H failed"); return MGB_ECUDA; }
            rc = refill(free_slots);
            if (rc != MGB_OK) { cleanup(); return rc; }
            // slots left without an utterance (the queue is drained) keep being computed: park them on position 0 of a page of
            // their own, so the row they rewrite every step can never belong to another utterance
            for (int b : free_slots) {
                if (slot_utt[b] >= 0) continue;
                if (!ensure_pages(*s, b, 1) || !flush_page_table(*s)) { cleanup(); return MGB_ERANGE; }
                const int32_t zero = 0, row = cache_row(*s, b, 0);
                if (cudaMemcpyAsync(s->dec_pos + b, &zero, 4, cudaMemcpyHostToDevice, st) != cudaSuccess ||
                    cudaMemcpyAsync(s->dec_slot + b, &row, 4, cudaMemcpyHostToDevice, st) != cudaSuccess ||
                    cudaStreamSynchronize(st) != cudaSuccess) { cleanup(); set_error("mgb_generate_queue: H2D failed"); return MGB_ECUDA; }
            }
        }
    }
    cudaEventRecord(s->ev1, st);
    if (cudaStreamSynchronize(st) != cudaSuccess) { cleanup(); set_error(std::string("mgb_generate_queue: ") + cudaGetErrorString(cudaGetLastError())); return MGB_ECUDA; }
    cudaEventElapsedTime(&s->last_ms, s->ev0, s->ev1);
    s->last_launches = g_launch_counter - launches0;
    cleanup();
    if (steps_run_out) *steps_run_out = steps;
    s->prefilled = false;                              // the session's slots hold finished utterances: encode + prefill before reuse
    return MGB_OK;
}

static int codec_upload(Codec & c, const int32_t * codes, int B, int T) {
    const size_t n = (size_t)B * 8 * T;
    if (!grow((void **)&c.d_codes, &c.codes_cap, n * 4)) return MGB_ECUDA;
    if (cudaMemcpyAsync(c.d_codes, codes, n * 4, cudaMemcpyHostToDevice, (cudaStream_t)c.stream) != cudaSuccess) { set_error("codec: H2D failed"); return MGB_ECUDA; }
    return MGB_OK;
}

// ---- batched streaming synthesis ----------------------------------------------------------------------------------
// magpie_synthesize_sentence_streaming (magpie.cpp:4502-4829) for all B utterances of a session at once: the device loop runs
// `frames_per_chunk` frames, the new codes of every utterance come back in one copy, the codec decodes all utterances' chunks in
// ONE batched launch sequence and each utterance's new samples are handed to the callback.  The reference decodes every chunk
// with zero causal history (magpie.cpp:4483-4500: audible seams); codec_context_frames = N decodes each chunk together with the
// utterance's previous N frames and emits only the new samples (N >= 25 covers the codec's 24.8-frame receptive field).  As in
// the reference's streaming path the EOS frame's codes ARE decoded (magpie.cpp:4733-4742).
int mgb_stream_generate(mgb_session * ss, mgb_codec * cc, int max_steps, float temperature, int top_k, uint64_t seed, int ignore_eos,
                        int frames_per_chunk, int codec_context_frames, mgb_stream_callback cb, void * user, int32_t * n_frames_out) {
    Session * s = reinterpret_cast<Session *>(ss);
    Codec * c = reinterpret_cast<Codec *>(cc);
    if (!check_ready(s, true)) return MGB_EINVAL;
    if (!c || !cb) { set_error("mgb_stream_generate: null codec / callback"); return MGB_EINVAL; }
    if (c->device != s->m->device) { set_error("mgb_stream_generate: model and codec live on different devices"); return MGB_EINVAL; }
    const mgb_hparams & hp = s->m->hp;
    if (max_steps <= 0) max_steps = hp.max_dec_steps;
    if (temperature >= 0.01f && top_k < 1) { set_error("mgb_stream_generate: top_k must be >= 1 when sampling"); return MGB_EINVAL; }
    const int B = s->B, chunk = frames_per_chunk > 0 ? frames_per_chunk : 4, ctx = std::max(0, codec_context_frames), hop = c->hp.hop_length;
    if (s->pos + max_steps > s->max_seq) { set_error("mgb_stream_generate: KV cache too small for the requested steps"); return MGB_ERANGE; }
    LoopCfg lc;
    lc.T = max_steps; lc.temperature = temperature; lc.top_k = top_k; lc.seed = seed; lc.ignore_eos = ignore_eos != 0;
    if (!prepare_loop_buffers(*s, lc)) return MGB_ECUDA;
    cudaStream_t st = s->stream, cst = (cudaStream_t)c->stream;
    std::vector<int32_t> bos((size_t)B * 8, hp.audio_bos_id), neg(B, -1);
    if (cudaMemcpyAsync(s->d_codes, bos.data(), bos.size() * 4, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemsetAsync(s->d_step, 0, 4, st) != cudaSuccess ||
        cudaMemcpyAsync(s->d_done, neg.data(), B * 4, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) { set_error("mgb_stream_generate: init failed"); return MGB_ECUDA; }
    const bool persistent = s->loop_grid > 0 && s->loop_tables;
    if (persistent || s->mega_grid > 0) {
        if (!ensure_pages_all(*s, std::min(s->pos + max_steps + 1, s->max_seq)) || !pages_contiguous(*s, s->pos + max_steps)) {
            set_error("mgb_stream_generate: the batch-1 kernel needs contiguous cache pages"); return MGB_ERANGE;
        }
    }
    cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
    auto cleanup = [&]() { if (exec) cudaGraphExecDestroy(exec); if (graph) cudaGraphDestroy(graph); };
    if (!persistent) {
        s->attn_split = attention_plan_kv_split(hp.dec_sa_heads * B, s->pos + max_steps);
        if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { set_error("graph capture begin failed"); return MGB_ECUDA; }
        const bool okq = enqueue_iteration(*s, lc);
        const cudaError_t e = cudaStreamEndCapture(st, &graph);
        if (!okq || e != cudaSuccess || !graph) { cleanup(); if (okq) set_error("graph capture failed"); return MGB_ECUDA; }
        if (cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) { cleanup(); set_error("graph instantiate failed"); return MGB_ECUDA; }
    }
    std::vector<std::vector<int32_t>> hist(B);          // per utterance: the last <= ctx frames already emitted, frame-major
    std::vector<int> frames(B, 0), finished(B, 0);
    std::vector<int32_t> h_codes((size_t)B * chunk * 8), h_done(B, -1), cbm;
    std::vector<float> pcm;
    int t = 0, rc = MGB_OK, n_done = 0;
    float loop_ms = 0.0f;
    bool prelaunched = false;
    while (t < max_steps && n_done < B && rc == MGB_OK) {
        const int n = std::min(chunk, max_steps - t);
        int ran = n;
        if (persistent) {
            lc.n = n; lc.step0 = t;
            rc = run_loop_persistent(*s, lc, &ran);
            loop_ms += s->last_ms;
            if (rc != MGB_OK) break;
            if (cudaMemcpyAsync(h_done.data(), s->d_done, 4, cudaMemcpyDeviceToHost, st) != cudaSuccess) { set_error("D2H failed"); rc = MGB_ECUDA; break; }
        } else {
            if (!prelaunched) {
                if (!ensure_pages_all(*s, std::min(s->pos + n + 1, s->max_seq))) { rc = MGB_ERANGE; break; }
                cudaEventRecord(s->ev0, st);
                for (int i = 0; i < n && rc == MGB_OK; i++) if (cudaGraphLaunch(exec, st) != cudaSuccess) { set_error("graph launch failed"); rc = MGB_ECUDA; }
                cudaEventRecord(s->ev1, st);
                if (rc != MGB_OK) break;
            }
            prelaunched = false;
            if (cudaMemcpyAsync(h_done.data(), s->d_done, B * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess) { set_error("D2H failed"); rc = MGB_ECUDA; break; }
        }
        // the chunk's sampled codes of every utterance: rows [b][t .. t + n) of the loop buffer
        if (cudaMemcpy2DAsync(h_codes.data(), (size_t)chunk * 32, s->l_sampled + (size_t)t * 8, (size_t)max_steps * 32, (size_t)n * 32, B,
                              cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) {
            set_error(std::string("mgb_stream_generate: ") + cudaGetErrorString(cudaGetLastError())); rc = MGB_ECUDA; break;
        }
        if (!persistent) { float ms = 0.0f; cudaEventElapsedTime(&ms, s->ev0, s->ev1); loop_ms += ms; s->pos += n; }
        if (persistent && h_done[0] >= 0) h_done[0] += t;              // the kernel reports the EOS step relative to its launch
        // new frames per utterance (EOS frame included, then the utterance is finished); group equal (history, new) lengths into
        // one batched codec call -- in lock step that is every active utterance
        std::vector<int> nnew(B, 0);
        for (int b = 0; b < B; b++) {
            if (finished[b]) continue;
            const bool eos = !ignore_eos && h_done[b] >= 0 && h_done[b] < t + ran;
            nnew[b] = eos ? h_done[b] - t + 1 : ran;
            if (eos || t + ran >= max_steps) finished[b] = 1;
        }
        // the NEXT chunk's frames are generated (session stream) while this chunk is decoded by the codec (codec stream)
        {
            int nd = 0;
            for (int b = 0; b < B; b++) nd += finished[b];
            const int n2 = std::min(chunk, max_steps - (t + ran));
            if (!persistent && nd < B && n2 > 0 && ran == n) {
                if (!ensure_pages_all(*s, std::min(s->pos + n2 + 1, s->max_seq))) { rc = MGB_ERANGE; break; }
                cudaEventRecord(s->ev0, st);
                for (int i = 0; i < n2 && rc == MGB_OK; i++) if (cudaGraphLaunch(exec, st) != cudaSuccess) { set_error("graph launch failed"); rc = MGB_ECUDA; }
                cudaEventRecord(s->ev1, st);
                if (rc != MGB_OK) break;
                prelaunched = true;
            }
        }
        std::vector<char> handled(B, 0);
        for (int b0 = 0; b0 < B && rc == MGB_OK; b0++) {
            if (handled[b0] || nnew[b0] <= 0) { handled[b0] = 1; continue; }
            const int nh = (int)hist[b0].size() / 8, nn = nnew[b0], T = nh + nn;
            std::vector<int> grp;
            for (int b = b0; b < B; b++) if (!handled[b] && nnew[b] == nn && (int)hist[b].size() / 8 == nh) { grp.push_back(b); handled[b] = 1; }
            const int G = (int)grp.size();
            cbm.assign((size_t)G * 8 * T, 0);
            for (int g = 0; g < G; g++) {
                const int b = grp[g];
                for (int f = 0; f < T; f++) {
                    const int32_t * src = f < nh ? &hist[b][(size_t)f * 8] : &h_codes[((size_t)b * chunk + (f - nh)) * 8];
                    for (int q = 0; q < 8; q++) cbm[((size_t)g * 8 + q) * T + f] = src[q];
                }
            }
            if (codec_upload(*c, cbm.data(), G, T) != MGB_OK) { rc = MGB_ECUDA; break; }
            const size_t ns = (size_t)G * T * hop;
            if (!grow((void **)&c->d_pcm, &c->pcm_cap, ns * 4)) { rc = MGB_ECUDA; break; }
            if (!codec_decode_device(*c, c->d_codes, G, T, c->d_pcm, cst)) { rc = MGB_ECUDA; break; }
            pcm.resize((size_t)G * nn * hop);
            // only the new samples of every utterance come back (the context part was emitted with earlier chunks)
            if (cudaMemcpy2DAsync(pcm.data(), (size_t)nn * hop * 4, c->d_pcm + (size_t)nh * hop, (size_t)T * hop * 4, (size_t)nn * hop * 4, G,
                                  cudaMemcpyDeviceToHost, cst) != cudaSuccess || cudaStreamSynchronize(cst) != cudaSuccess) {
                set_error(std::string("mgb_stream_generate (codec): ") + cudaGetErrorString(cudaGetLastError())); rc = MGB_ECUDA; break;
            }
            for (int g = 0; g < G && rc == MGB_OK; g++) {
                const int b = grp[g];
                frames[b] += nn;
                if (ctx > 0) {
                    hist[b].insert(hist[b].end(), &h_codes[(size_t)b * chunk * 8], &h_codes[((size_t)b * chunk + nn) * 8]);
                    if ((int)hist[b].size() > ctx * 8) hist[b].erase(hist[b].begin(), hist[b].end() - (size_t)ctx * 8);
                }
                if (cb(b, pcm.data() + (size_t)g * nn * hop, nn * hop, frames[b], finished[b], user) != 0) { set_error("mgb_stream_generate: stopped by the callback"); rc = MGB_EINVAL; }
            }
        }
        n_done = 0;
        for (int b = 0; b < B; b++) n_done += finished[b];
        t += ran;
        if (ran < n) break;                                            // (persistent kernel: stopped at EOS)
    }
    if (prelaunched) cudaStreamSynchronize(st);        // (stopped by the callback with a chunk still in flight)
    cleanup();
    s->last_ms = loop_ms;
    if (n_frames_out) for (int b = 0; b < B; b++) n_frames_out[b] = frames[b];
    return rc;
}

// ---- codec ------------------------------------------------------------------------------------------
mgb_codec * mgb_codec_load(const char * path, int device) {
    if (!path) { set_error("null path"); return nullptr; }
    try { return reinterpret_cast<mgb_codec *>(load_codec(path, device)); }
    catch (const std::exception & e) { set_error(std::string("magpie_codec_init: ") + e.what()); return nullptr; }
    catch (...) { set_error("magpie_codec_init: unknown failure"); return nullptr; }
}
void mgb_codec_free(mgb_codec * c) { delete reinterpret_cast<Codec *>(c); }
int mgb_codec_get_hparams(const mgb_codec * c, mgb_codec_hparams * out) {
    if (!c || !out) return MGB_EINVAL;
    *out = reinterpret_cast<const Codec *>(c)->hp; return MGB_OK;
}
float mgb_codec_last_ms(const mgb_codec * c) { return c ? reinterpret_cast<const Codec *>(c)->last_ms : 0.0f; }
int64_t mgb_codec_last_launches(const mgb_codec * c) { return c ? reinterpret_cast<const Codec *>(c)->last_launches : 0; }

int mgb_codec_decode(mgb_codec * cc, const int32_t * codes, int batch, int n_frames, float * pcm_out) {
    Codec * c = reinterpret_cast<Codec *>(cc);
    if (!c || !codes || !pcm_out || batch <= 0 || n_frames <= 0) { set_error("magpie_codec_decode: invalid arguments"); return MGB_EINVAL; }
    if (cudaSetDevice(c->device) != cudaSuccess) { set_error("cudaSetDevice failed"); return MGB_ECUDA; }
    cudaStream_t st = (cudaStream_t)c->stream;
    int rc = codec_upload(*c, codes, batch, n_frames);
    if (rc != MGB_OK) return rc;
    const size_t ns = (size_t)batch * n_frames * c->hp.hop_length;
    if (!grow((void **)&c->d_pcm, &c->pcm_cap, ns * 4)) return MGB_ECUDA;
    // bound scratch memory: decode utterances in groups of <= ~8K frames.  All groups are enqueued first; the PCM of
    // group g is then copied back while the kernels of the later groups are still running (events per group).
    static const int group_frames = [] { const char * e = getenv("MGB_CODEC_GROUP_FRAMES"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 8192; }();
    const int per = std::max(1, group_frames / n_frames);
    const int n_groups = (batch + per - 1) / per;
    while ((int)c->group_events.size() < n_groups) {
        cudaEvent_t e = nullptr;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { set_error("codec: event creation failed"); return MGB_ECUDA; }
        c->group_events.push_back(e);
    }
    const int64_t l0 = g_launch_counter;
    cudaEventRecord((cudaEvent_t)c->ev0, st);
    for (int g = 0; g < n_groups; g++) {
        const int b0 = g * per, nb = std::min(per, batch - b0);
        if (!codec_decode_device(*c, c->d_codes + (size_t)b0 * 8 * n_frames, nb, n_frames,
                                 c->d_pcm + (size_t)b0 * n_frames * c->hp.hop_length, st)) return MGB_ECUDA;
        cudaEventRecord((cudaEvent_t)c->group_events[g], st);
    }
    cudaEventRecord((cudaEvent_t)c->ev1, st);
    for (int g = 0; g < n_groups; g++) {
        const int b0 = g * per, nb = std::min(per, batch - b0);
        const size_t off = (size_t)b0 * n_frames * c->hp.hop_length, cnt = (size_t)nb * n_frames * c->hp.hop_length;
        if (cudaEventSynchronize((cudaEvent_t)c->group_events[g]) != cudaSuccess ||
            cudaMemcpy(pcm_out + off, c->d_pcm + off, cnt * 4, cudaMemcpyDeviceToHost) != cudaSuccess) {
            set_error(std::string("magpie_codec_decode: ") + cudaGetErrorString(cudaGetLastError())); return MGB_ECUDA;
        }
    }
    if (cudaStreamSynchronize(st) != cudaSuccess) {
        set_error(std::string("magpie_codec_decode: ") + cudaGetErrorString(cudaGetLastError())); return MGB_ECUDA;
    }
    cudaEventElapsedTime(&c->last_ms, (cudaEvent_t)c->ev0, (cudaEvent_t)c->ev1);
    c->last_launches = g_launch_counter - l0;
    return MGB_OK;
}

int mgb_codec_fsq_dequantize(mgb_codec * cc, const int32_t * codes, int batch, int n_frames, float * latent_out) {
    Codec * c = reinterpret_cast<Codec *>(cc);
    if (!c || !codes || !latent_out || batch <= 0 || n_frames <= 0) { set_error("fsq_dequantize: invalid arguments"); return MGB_EINVAL; }
    if (cudaSetDevice(c->device) != cudaSuccess) { set_error("cudaSetDevice failed"); return MGB_ECUDA; }
    cudaStream_t st = (cudaStream_t)c->stream;
    int rc = codec_upload(*c, codes, batch, n_frames);
    if (rc != MGB_OK) return rc;
    const size_t n = (size_t)batch * 32 * n_frames;
    if (!grow((void **)&c->d_pcm, &c->pcm_cap, n * 4)) return MGB_ECUDA;
    if (!codec_fsq_device(c->d_codes, batch, n_frames, c->d_pcm, st)) return MGB_ECUDA;
    if (cudaMemcpyAsync(latent_out, c->d_pcm, n * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) { set_error("fsq_dequantize: D2H failed"); return MGB_ECUDA; }
    return MGB_OK;
}

}  // extern "C"
