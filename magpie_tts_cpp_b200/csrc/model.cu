// GGUF -> device weights.  Name mapping follows the reference loader's precedence
// (src/magpie.cpp:501-562, 607-667 and src/nano-codec.cpp:84-199).
#include "model.h"

#include <cstdlib>
#include <cstring>
#include <memory>

#include "common.cuh"
#include "kernels.cuh"

namespace mgb {

thread_local int64_t g_launch_counter = 0;
static thread_local std::string g_error;
void set_error(const std::string & msg) { g_error = msg; }
const std::string & get_error() { return g_error; }

namespace {

// f32 -> bf16 (round to nearest even) on the device: the host loop took 2.1 of the 2.8 s a bf16 load of the 858 MB f32 file needs
__global__ void f32_to_bf16_kernel(const float * __restrict__ src, __nv_bfloat16 * __restrict__ dst, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = __float2bfloat16_rn(src[i]);
}

struct Uploader {
    std::vector<void *> & allocs;
    bool ok = true;
    float * stage = nullptr; size_t stage_elems = 0;       // device staging buffer for the f32 image of a bf16 tensor
    ~Uploader() { if (stage) cudaFree(stage); }
    void * raw(const void * host, size_t bytes) {
        void * d = nullptr;
        if (cudaMalloc(&d, bytes ? bytes : 16) != cudaSuccess) { ok = false; set_error("cudaMalloc failed"); return nullptr; }
        allocs.push_back(d);
        if (bytes && cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
            ok = false; set_error("cudaMemcpy H2D failed"); return nullptr;
        }
        return d;
    }
    float * f32(const std::vector<float> & v) { return (float *)raw(v.data(), v.size() * 4); }
    // weight matrix in the model dtype
    void * weights(const std::vector<float> & v, int precision) {
        if (precision == MGB_PREC_F32) return raw(v.data(), v.size() * 4);
        if (v.size() > stage_elems) {
            if (stage) cudaFree(stage);
            stage = nullptr; stage_elems = 0;
            if (cudaMalloc((void **)&stage, v.size() * 4) != cudaSuccess) { ok = false; set_error("cudaMalloc failed (staging)"); return nullptr; }
            stage_elems = v.size();
        }
        void * d = nullptr;
        if (cudaMalloc(&d, v.size() ? v.size() * 2 : 16) != cudaSuccess) { ok = false; set_error("cudaMalloc failed"); return nullptr; }
        allocs.push_back(d);
        if (v.empty()) return d;
        // (legacy default stream: the copy waits for the previous tensor's conversion kernel, the kernel for the copy)
        if (cudaMemcpy(stage, v.data(), v.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) { ok = false; set_error("cudaMemcpy H2D failed"); return nullptr; }
        f32_to_bf16_kernel<<<592, 256>>>(stage, (__nv_bfloat16 *)d, v.size());
        if (cudaGetLastError() != cudaSuccess) { ok = false; set_error("bf16 conversion kernel failed"); return nullptr; }
        return d;
    }
};

bool to_f32(const GgufTensor & t, std::vector<float> & out) {
    out.resize((size_t)t.nelements());
    if (!gguf_to_f32(t, out.data())) { set_error("tensor '" + t.name + "': unsupported ggml type " + std::to_string(t.type)); return false; }
    return true;
}

int parse_idx(const char * name, const char * prefix) {
    const char * p = strstr(name, prefix);
    return p ? atoi(p + strlen(prefix)) : -1;
}

// Linear (out,in) or conv (out,in,k): ggml ne = [in,out] / [k,in,out].  k>1 is re-laid out
// tap-major [k][out][in] so every tap is a contiguous row-major matrix.
// The kernels index weights by the hyper-parameters, so every tensor's shape is checked against them BEFORE upload (a
// mismatched or corrupt file must fail here, where the reference's ggml would assert, not read device memory out of bounds).
bool make_mat(Uploader & up, const GgufTensor & t, int precision, DevMat & m, int64_t want_n, int64_t want_k, int64_t want_taps = 1) {
    {
        const int64_t k = t.n_dims == 3 ? t.ne[1] : t.ne[0], n = t.n_dims == 3 ? t.ne[2] : t.ne[1], taps = t.n_dims == 3 ? t.ne[0] : 1;
        if (t.n_dims < 2 || t.n_dims > 3 || n != want_n || k != want_k || taps != want_taps) {
            set_error("tensor '" + t.name + "': shape [" + std::to_string(n) + "][" + std::to_string(k) + "] x " + std::to_string(taps) +
                      " taps does not match the hyper-parameters (expected [" + std::to_string(want_n) + "][" + std::to_string(want_k) +
                      "] x " + std::to_string(want_taps) + ")");
            return false;
        }
    }
    std::vector<float> v;
    if (!to_f32(t, v)) return false;
    if (t.n_dims == 3 && t.ne[0] > 1) {
        int k = (int)t.ne[0], in = (int)t.ne[1], out = (int)t.ne[2];
        std::vector<float> r(v.size());
        for (int o = 0; o < out; o++)
            for (int i = 0; i < in; i++)
                for (int kk = 0; kk < k; kk++)
                    r[((size_t)kk * out + o) * in + i] = v[((size_t)o * in + i) * k + kk];
        m.N = out; m.K = in; m.taps = k;
        m.w = up.weights(r, precision);
    } else if (t.n_dims == 3) {
        m.N = (int)t.ne[2]; m.K = (int)t.ne[1]; m.taps = 1;
        m.w = up.weights(v, precision);
    } else {
        m.N = (int)t.ne[1]; m.K = (int)t.ne[0]; m.taps = 1;
        m.w = up.weights(v, precision);
    }
    if (m.K % 8 != 0) { set_error("tensor '" + t.name + "': inner dim must be a multiple of 8"); return false; }
    return up.ok;
}

// want_cols: required ne[0]; want_rows: required number of rows (product of the other dims), 0 = any (>= min_rows)
bool make_vec(Uploader & up, const GgufTensor & t, float *& dst, int64_t want_cols, int64_t want_rows, int * rows = nullptr, int64_t min_rows = 1) {
    {
        const int64_t cols = t.ne[0], r = t.nelements() / (t.ne[0] > 0 ? t.ne[0] : 1);
        const bool flat_ok = want_rows > 0 && t.nelements() == want_cols * want_rows;     // e.g. baked context stored as (speakers, C*d)
        if (!flat_ok && (cols != want_cols || (want_rows > 0 ? r != want_rows : r < min_rows))) {
            set_error("tensor '" + t.name + "': " + std::to_string(r) + " rows of " + std::to_string(cols) +
                      " do not match the hyper-parameters (expected " + (want_rows > 0 ? std::to_string(want_rows) : ">= " + std::to_string(min_rows)) +
                      " rows of " + std::to_string(want_cols) + ")");
            return false;
        }
    }
    std::vector<float> v;
    if (!to_f32(t, v)) return false;
    dst = up.f32(v);
    if (rows) *rows = (int)(t.nelements() / t.ne[0]);
    return up.ok;
}

}  // namespace

Model::~Model() {
    cudaSetDevice(device);
    for (void * p : allocations) cudaFree(p);
}

Model * load_model(const char * path, int device, int precision) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        set_error("no CUDA device available (this build has no CPU fallback)");
        return nullptr;
    }
    if (device < 0 || device >= ndev) { set_error("invalid device index"); return nullptr; }
    if (precision != MGB_PREC_F32 && precision != MGB_PREC_BF16) { set_error("invalid precision"); return nullptr; }
    if (cudaSetDevice(device) != cudaSuccess) { set_error("cudaSetDevice failed"); return nullptr; }

    GgufFile f;
    std::string err;
    if (!f.open(path, err)) { set_error(std::string("magpie_init: ") + err); return nullptr; }

    std::unique_ptr<Model> M(new Model());
    M->device = device; M->precision = precision; M->wsize = precision == MGB_PREC_F32 ? 4 : 2;
    mgb_hparams & hp = M->hp;
    // defaults = reference struct initialisers (magpie.h:35-80); keys as read at magpie.cpp:85-120
#define HP(field, def) hp.field = f.get_u32("magpie." #field, def)
    HP(d_model, 768); HP(d_ffn, 3072); HP(d_head, 64);
    HP(enc_layers, 6); HP(enc_heads, 12); HP(enc_kernel, 3);
    HP(dec_layers, 12); HP(dec_sa_heads, 12); HP(dec_xa_heads, 1); HP(dec_xa_d_head, 128); HP(dec_kernel, 1);
    HP(lt_dim, 256); HP(lt_ffn_dim, 1024); HP(lt_layers, 1); HP(lt_heads, 1);
    HP(text_vocab_size, 2380); HP(num_codebooks, 8); HP(codebook_size, 2016); HP(vocab_per_cb, 2024);
    HP(num_speakers, 5); HP(context_frames, 110);
    HP(text_bos_id, 2378); HP(text_eos_id, 2379); HP(audio_bos_id, 2016); HP(audio_eos_id, 2017);
    HP(max_dec_steps, 500); HP(sample_rate, 22050);
#undef HP
    hp.eps = f.get_f32("magpie.eps", 1e-5f);
    if (hp.d_model <= 0 || hp.d_ffn <= 0 || hp.d_ffn % 8 || hp.enc_layers <= 0 || hp.enc_layers > 64 || hp.dec_layers <= 0 || hp.dec_layers > 64 ||
        hp.enc_heads <= 0 || hp.dec_sa_heads <= 0 || hp.enc_kernel <= 0 || hp.lt_dim <= 0 || hp.lt_ffn_dim <= 0 || hp.lt_ffn_dim % 8 ||
        hp.text_vocab_size <= 0 || hp.vocab_per_cb <= 0 || hp.num_speakers <= 0 || hp.context_frames <= 0 || hp.max_dec_steps <= 0 ||
        hp.audio_bos_id < 0 || hp.audio_bos_id + 7 >= hp.vocab_per_cb || hp.audio_eos_id < 0 || hp.audio_eos_id >= hp.vocab_per_cb) {
        set_error("magpie_init: invalid hyper-parameters in the model file");
        return nullptr;
    }
    if (hp.num_codebooks != 8 || hp.lt_layers != 1 || hp.lt_heads != 1 || hp.dec_kernel != 1 ||
        hp.enc_kernel > 3 || hp.d_model % 8 || hp.d_model > 1024 || hp.lt_dim > 256 || hp.lt_dim % 8 ||
        hp.dec_xa_heads != 1 || hp.dec_xa_d_head != 128 || hp.d_model / hp.dec_sa_heads != 64 ||
        hp.d_model / hp.enc_heads != 64 || hp.lt_ffn_dim > 1024 || hp.vocab_per_cb > 2048) {
        set_error("magpie_init: unsupported architecture hyper-parameters for the sm_100a kernels");
        return nullptr;
    }
    for (const char * k : {"magpie.tokenizer.vocab", "magpie.tokenizer.dict"})
        if (const std::string * s = f.get_str(k)) M->meta_str[k] = *s;
    for (const char * k : {"magpie.tokenizer.pad", "magpie.tokenizer.oov", "magpie.tokenizer.space"})
        if (f.find(k)) M->meta_u32[k] = f.get_u32(k, -1);

    M->enc.resize(hp.enc_layers);
    M->dec.resize(hp.dec_layers);
    Uploader up{M->allocations};
    for (const GgufTensor & t : f.tensors()) {
        const char * name = t.name.c_str();
        bool ok = true;
        if (!strcmp(name, "text_embedding.weight")) ok = make_vec(up, t, M->text_emb, hp.d_model, hp.text_vocab_size);
        else if (strstr(name, "audio_embeddings.")) {
            int cb = parse_idx(name, "audio_embeddings.");
            if (cb >= 0 && cb < 8) ok = make_vec(up, t, M->audio_emb[cb], hp.d_model, hp.vocab_per_cb);
        } else if (!strcmp(name, "baked_context_embedding.weight")) ok = make_vec(up, t, M->baked_ctx, hp.d_model, (int64_t)hp.num_speakers * hp.context_frames);
        else if (!strcmp(name, "encoder.position_embeddings.weight")) ok = make_vec(up, t, M->enc_pos, hp.d_model, 0, &M->enc_pos_rows);
        else if (strstr(name, "encoder.layers.")) {     // NB: matched before decoder.layers, as the reference
            int l = parse_idx(name, "encoder.layers.");
            if (l >= 0 && l < hp.enc_layers) {
                EncLayer & L = M->enc[l];
                if (strstr(name, "norm_self.weight")) ok = make_vec(up, t, L.norm_self, hp.d_model, 1);
                else if (strstr(name, "self_attention.qkv_net.weight")) ok = make_mat(up, t, precision, L.qkv, 3 * hp.d_model, hp.d_model);
                else if (strstr(name, "self_attention.o_net.weight")) ok = make_mat(up, t, precision, L.o, hp.d_model, hp.d_model);
                else if (strstr(name, "norm_pos_ff.weight")) ok = make_vec(up, t, L.norm_ff, hp.d_model, 1);
                else if (strstr(name, "pos_ff.proj.conv.weight")) ok = make_mat(up, t, precision, L.ff1, hp.d_ffn, hp.d_model, hp.enc_kernel);
                else if (strstr(name, "pos_ff.o_net.conv.weight")) ok = make_mat(up, t, precision, L.ff2, hp.d_model, hp.d_ffn, hp.enc_kernel);
            }
        } else if (!strcmp(name, "encoder.norm_out.weight")) ok = make_vec(up, t, M->enc_norm_out, hp.d_model, 1);
        else if (!strcmp(name, "decoder.position_embeddings.weight")) ok = make_vec(up, t, M->dec_pos, hp.d_model, 0, &M->dec_pos_rows, hp.context_frames + 2);
        else if (strstr(name, "decoder.layers.")) {
            int l = parse_idx(name, "decoder.layers.");
            if (l >= 0 && l < hp.dec_layers) {
                DecLayer & L = M->dec[l];
                if (strstr(name, "norm_self.weight")) ok = make_vec(up, t, L.norm_self, hp.d_model, 1);
                else if (strstr(name, "self_attention.qkv_net.weight")) ok = make_mat(up, t, precision, L.qkv, 3 * hp.d_model, hp.d_model);
                else if (strstr(name, "self_attention.o_net.weight")) ok = make_mat(up, t, precision, L.o, hp.d_model, hp.d_model);
                else if (strstr(name, "norm_xattn_query.weight")) ok = make_vec(up, t, L.norm_xa_q, hp.d_model, 1);
                else if (strstr(name, "cross_attention.q_net.weight")) ok = make_mat(up, t, precision, L.xq, 128, hp.d_model);
                else if (strstr(name, "cross_attention.kv_net.weight")) ok = make_mat(up, t, precision, L.xkv, 256, hp.d_model);
                else if (strstr(name, "cross_attention.o_net.weight")) ok = make_mat(up, t, precision, L.xo, hp.d_model, 128);
                else if (strstr(name, "norm_xattn_memory.weight")) ok = make_vec(up, t, L.norm_xa_mem, hp.d_model, 1);
                else if (strstr(name, "norm_pos_ff.weight")) ok = make_vec(up, t, L.norm_ff, hp.d_model, 1);
                else if (strstr(name, "pos_ff.proj.conv.weight")) ok = make_mat(up, t, precision, L.ff1, hp.d_ffn, hp.d_model);
                else if (strstr(name, "pos_ff.o_net.conv.weight")) ok = make_mat(up, t, precision, L.ff2, hp.d_model, hp.d_ffn);
            }
        } else if (!strcmp(name, "decoder.norm_out.weight")) ok = make_vec(up, t, M->dec_norm_out, hp.d_model, 1);
        else if (!strcmp(name, "final_proj.weight")) ok = make_mat(up, t, precision, M->final_w, (int64_t)8 * hp.vocab_per_cb, hp.d_model);
        else if (!strcmp(name, "final_proj.bias")) ok = make_vec(up, t, M->final_b, (int64_t)8 * hp.vocab_per_cb, 1);
        else if (strstr(name, "local_transformer_in_projection.weight")) ok = make_mat(up, t, precision, M->lt_in_w, hp.lt_dim, hp.d_model);
        else if (strstr(name, "local_transformer_in_projection.bias")) ok = make_vec(up, t, M->lt_in_b, hp.lt_dim, 1);
        else if (!strcmp(name, "local_transformer.position_embeddings.weight")) ok = make_vec(up, t, M->lt_pos, hp.lt_dim, 0, &M->lt_pos_rows, 8);
        else if (strstr(name, "local_transformer.layers.0.norm_self.weight")) ok = make_vec(up, t, M->lt_norm_self, hp.lt_dim, 1);
        else if (strstr(name, "local_transformer.layers.0.self_attention.qkv_net.weight")) ok = make_mat(up, t, precision, M->lt_qkv, 3 * hp.lt_dim, hp.lt_dim);
        else if (strstr(name, "local_transformer.layers.0.self_attention.o_net.weight")) ok = make_mat(up, t, precision, M->lt_o, hp.lt_dim, hp.lt_dim);
        else if (strstr(name, "local_transformer.layers.0.norm_pos_ff.weight")) ok = make_vec(up, t, M->lt_norm_ff, hp.lt_dim, 1);
        else if (strstr(name, "local_transformer.layers.0.pos_ff.proj.conv.weight")) ok = make_mat(up, t, precision, M->lt_ff1, hp.lt_ffn_dim, hp.lt_dim);
        else if (strstr(name, "local_transformer.layers.0.pos_ff.o_net.conv.weight")) ok = make_mat(up, t, precision, M->lt_ff2, hp.lt_dim, hp.lt_ffn_dim);
        else if (strstr(name, "local_transformer_out_projections.")) {
            int cb = parse_idx(name, "local_transformer_out_projections.");
            if (cb >= 0 && cb < 8) {
                if (strstr(name, ".weight")) ok = make_mat(up, t, precision, M->lt_out_w[cb], hp.vocab_per_cb, hp.lt_dim);
                else if (strstr(name, ".bias")) ok = make_vec(up, t, M->lt_out_b[cb], hp.vocab_per_cb, 1);
            }
        }
        // anything else (context_encoder.* etc.) is loaded-but-unused in the reference: skipped here
        if (!ok || !up.ok) return nullptr;
    }

    // completeness check: every slot the hot path touches must be present
    bool complete = M->text_emb && M->baked_ctx && M->enc_pos && M->dec_pos && M->lt_pos && M->enc_norm_out &&
                    M->dec_norm_out && M->lt_in_w.w && M->lt_in_b && M->lt_norm_self && M->lt_norm_ff &&
                    M->lt_qkv.w && M->lt_o.w && M->lt_ff1.w && M->lt_ff2.w;
    for (int cb = 0; cb < 8; cb++) complete = complete && M->audio_emb[cb] && M->lt_out_w[cb].w && M->lt_out_b[cb];
    for (auto & L : M->enc) complete = complete && L.norm_self && L.norm_ff && L.qkv.w && L.o.w && L.ff1.w && L.ff2.w;
    for (auto & L : M->dec)
        complete = complete && L.norm_self && L.norm_xa_q && L.norm_xa_mem && L.norm_ff && L.qkv.w && L.o.w &&
                   L.xq.w && L.xkv.w && L.xo.w && L.ff1.w && L.ff2.w;
    if (!complete) { set_error("magpie_init: model file is missing tensors of the synthesis path"); return nullptr; }
    if (M->dec_pos_rows < hp.context_frames + 2) { set_error("magpie_init: decoder position table too short"); return nullptr; }

    // feedback table of the local transformer: P_cb[code] = in_proj . E_cb[code] + b (magpie.cpp:1274-1313),
    // computed once on the device so that the per-codebook feedback is a row gather
    for (int cb = 0; cb < 8; cb++) {
        void * t = nullptr;
        if (cudaMalloc(&t, (size_t)hp.vocab_per_cb * hp.lt_dim * sizeof(float)) != cudaSuccess) { set_error("cudaMalloc failed"); return nullptr; }
        M->allocations.push_back(t);
        M->lt_in_table[cb] = (float *)t;
        LinearArgs a;
        a.precision = precision; a.M = hp.vocab_per_cb; a.W = M->lt_in_w; a.X = M->audio_emb[cb]; a.ldx = hp.d_model;
        a.bias = M->lt_in_b; a.Y = M->lt_in_table[cb]; a.ldy = hp.lt_dim;
        if (!launch_linear(a, nullptr)) return nullptr;
    }
    if (cudaDeviceSynchronize() != cudaSuccess) { set_error("magpie_init: building the LT feedback table failed"); return nullptr; }
    // local transformer, bf16: [Wq; Wk; hi(Wo Wv); lo(Wo Wv)] -- with the O-projection folded into the value rows the attention
    // output needs no GEMV of its own (o . sum_j p_j v_j = sum_j p_j (Wo Wv) n_j); the product is kept as a bf16 hi + lo pair
    if (precision == MGB_PREC_BF16) {
        const int L = hp.lt_dim;
        void * t = nullptr;
        if (cudaMalloc(&t, (size_t)4 * L * L * 2) != cudaSuccess) { set_error("cudaMalloc failed"); return nullptr; }
        M->allocations.push_back(t);
        M->lt_qkvo = t;
        if (!launch_lt_fold_ov(M->lt_qkv.w, M->lt_o.w, L, t, nullptr) || cudaDeviceSynchronize() != cudaSuccess) {
            set_error("magpie_init: folding the LT output projection failed"); return nullptr;
        }
        // LT positions 1..7 see x = P_cb[code] + pos[cb+1], a function of (cb, code) only: their [q | k | vo] rows are tabulated
        // (7 x V x 3L f32 = 43.5 MB), so those positions need neither LayerNorm + QKV GEMV nor the exchange of its result
        void * tb = nullptr;
        if (cudaMalloc(&tb, (size_t)7 * hp.vocab_per_cb * 3 * L * sizeof(float)) != cudaSuccess) { set_error("cudaMalloc failed"); return nullptr; }
        M->allocations.push_back(tb);
        M->lt_qkv_tab = (float *)tb;
        for (int cb = 0; cb < 7; cb++)
            if (!launch_lt_qkv_table(M->lt_in_table[cb], M->lt_pos + (size_t)(cb + 1) * L, M->lt_norm_self, hp.eps, t, hp.vocab_per_cb, L,
                                     M->lt_qkv_tab + (size_t)cb * hp.vocab_per_cb * 3 * L, nullptr)) return nullptr;
        if (cudaDeviceSynchronize() != cudaSuccess) { set_error("magpie_init: building the LT QKV table failed"); return nullptr; }
    }

    // bf16 models: tensor-core tile images of the matrices the batched paths multiply with (gemm_tc.cu);
    // MGB_NO_TC=1 keeps the CUDA-core kernels (A/B tests)
    if (precision == MGB_PREC_BF16 && getenv("MGB_NO_TC") == nullptr) {
        std::vector<DevMat *> mats;
        for (auto & L : M->dec) for (DevMat * m : {&L.qkv, &L.o, &L.xq, &L.xkv, &L.xo, &L.ff1, &L.ff2}) mats.push_back(m);
        for (auto & L : M->enc) for (DevMat * m : {&L.qkv, &L.o, &L.ff1, &L.ff2}) mats.push_back(m);      // (ff1 / ff2: k = 3 causal convs, taps concatenated along k)
        mats.push_back(&M->final_w);
        mats.push_back(&M->lt_ff1); mats.push_back(&M->lt_ff2);                      // lt_cluster.cu: FFN and output projections on tcgen05
        for (int cb = 0; cb < 8; cb++) mats.push_back(&M->lt_out_w[cb]);
        for (DevMat * m : mats) {
            if (m->w == nullptr || m->K % 64 != 0) continue;
            void * t = nullptr;
            if (cudaMalloc(&t, tc_weight_tile_bytes(m->N, m->K * m->taps)) != cudaSuccess) { set_error("cudaMalloc failed (weight tiles)"); return nullptr; }
            M->allocations.push_back(t);
            if (!tc_pack_weights(m->w, m->N, m->K, t, nullptr, m->taps)) return nullptr;
            m->tiles = t;
        }
        // f16 activation images in the batched decoder step (gemm_tc.cuh pack_act2): f16 twins of the four decoder-step matrices (+175 MB)
        if (getenv("MGB_ACT_F16") == nullptr || atoi(getenv("MGB_ACT_F16")) != 0)      // default on; MGB_ACT_F16=0: bf16 hi | lo image pairs (A/B)
            for (auto & L : M->dec) for (DevMat * m : {&L.qkv, &L.o, &L.ff1, &L.ff2}) {
                if (!m->tiles || m->taps != 1) continue;
                void * t = nullptr;
                if (cudaMalloc(&t, tc_weight_tile_bytes(m->N, m->K)) != cudaSuccess) { set_error("cudaMalloc failed (f16 weight tiles)"); return nullptr; }
                M->allocations.push_back(t);
                if (!tc_pack_weights(m->w, m->N, m->K, t, nullptr, 1, true)) return nullptr;
                m->tiles16 = t;
            }
        for (auto & L : M->dec) {
            void * c = nullptr;
            if (cudaMalloc(&c, (size_t)L.qkv.N * sizeof(float)) != cudaSuccess) { set_error("cudaMalloc failed"); return nullptr; }
            M->allocations.push_back(c);
            if (!launch_row_dots(L.qkv.w, L.norm_self, L.qkv.N, L.qkv.K, (float *)c, nullptr)) return nullptr;
            L.qkv_csum = (float *)c;
            void * c1 = nullptr;
            if (L.ff1.taps == 1) {
                if (cudaMalloc(&c1, (size_t)L.ff1.N * sizeof(float)) != cudaSuccess) { set_error("cudaMalloc failed"); return nullptr; }
                M->allocations.push_back(c1);
                if (!launch_row_dots(L.ff1.w, L.norm_ff, L.ff1.N, L.ff1.K, (float *)c1, nullptr)) return nullptr;
                L.ff1_csum = (float *)c1;
            }
        }
        if (cudaDeviceSynchronize() != cudaSuccess) { set_error("magpie_init: packing the weight tiles failed"); return nullptr; }
    }

    // unique weight bytes one generated frame reads: decoder matrices + LT matrices
    int64_t el = 0, f32b = 0;
    for (auto & L : M->dec) {
        for (const DevMat * m : {&L.qkv, &L.o, &L.xq, &L.xo, &L.ff1, &L.ff2}) el += (int64_t)m->N * m->K * m->taps;
        f32b += 3 * hp.d_model * 4;
    }
    f32b += hp.d_model * 4;
    for (const DevMat * m : {&M->lt_in_w, &M->lt_qkv, &M->lt_o, &M->lt_ff1, &M->lt_ff2}) el += (int64_t)m->N * m->K;
    for (int cb = 0; cb < 8; cb++) { el += (int64_t)M->lt_out_w[cb].N * M->lt_out_w[cb].K; f32b += M->lt_out_w[cb].N * 4; }
    f32b += (hp.lt_dim * 3 + 8 * hp.lt_dim) * 4 + 8 * hp.d_model * 4;
    M->step_weight_bytes = el * (int64_t)M->wsize + f32b;
    return M.release();
}

// ---------------------------------------------------------------------------------------------
// nano-codec
// ---------------------------------------------------------------------------------------------
Codec::~Codec() {
    cudaSetDevice(device);
    for (void * p : allocations) cudaFree(p);
    for (auto & b : buf) if (b) cudaFree(b);
    for (auto & b : img) if (b) cudaFree(b);
    if (d_codes) cudaFree(d_codes);
    if (d_pcm) cudaFree(d_pcm);
    if (ev0) cudaEventDestroy((cudaEvent_t)ev0);
    if (ev1) cudaEventDestroy((cudaEvent_t)ev1);
    for (void * e : group_events) cudaEventDestroy((cudaEvent_t)e);
    if (stream) cudaStreamDestroy((cudaStream_t)stream);
}

Codec * load_codec(const char * path, int device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        set_error("no CUDA device available (this build has no CPU fallback)");
        return nullptr;
    }
    if (device < 0 || device >= ndev) { set_error("invalid device index"); return nullptr; }
    if (cudaSetDevice(device) != cudaSuccess) { set_error("cudaSetDevice failed"); return nullptr; }
    GgufFile f;
    std::string err;
    if (!f.open(path, err)) { set_error(std::string("magpie_codec_init: ") + err); return nullptr; }
    std::unique_ptr<Codec> C(new Codec());
    C->device = device;
    C->hp.sample_rate = f.get_u32("codec.sample_rate", 22050);        // nano-codec.cpp:77-81
    C->hp.num_codebooks = f.get_u32("codec.num_codebooks", 8);
    C->hp.codebook_size = f.get_u32("codec.codebook_size", 2016);
    C->hp.hop_length = f.get_u32("codec.hop_length", 1024);
    C->hp.latent_dim = f.get_u32("codec.latent_dim", 32);
    if (C->hp.num_codebooks != 8 || C->hp.latent_dim != 32 || C->hp.hop_length != 1024) {
        set_error("magpie_codec_init: unsupported codec hyper-parameters");
        return nullptr;
    }
    Uploader up{C->allocations};
    // the decoder architecture is fixed (magpie.h:666-677: 864 -> 432 -> 216 -> 108 -> 54 -> 27 channels, strides 8,8,4,2,2, residual
    // kernels 3/7/11): every tensor's element count is checked against it before upload, so a mismatched file fails here
    auto vec = [&](const GgufTensor & t, float *& dst, int64_t want, int * n = nullptr) {
        if (t.nelements() != want) {
            set_error("tensor '" + t.name + "': " + std::to_string(t.nelements()) + " elements, expected " + std::to_string(want) +
                      " for the nano-codec decoder");
            return false;
        }
        std::vector<float> v;
        if (!to_f32(t, v)) return false;
        dst = up.f32(v);
        if (n) *n = (int)v.size();
        return up.ok;
    };
    auto stage_ch = [&](int i) { return (int64_t)(C->base_ch >> (i + 1)); };       // output channels of up-sampling stage i
    for (const GgufTensor & t : f.tensors()) {
        const char * name = t.name.c_str();
        bool ok = true;
        if (strstr(name, "dec.pre.weight")) ok = vec(t, C->pre_w, (int64_t)C->base_ch * C->latent * C->pre_k);
        else if (strstr(name, "dec.pre.bias")) ok = vec(t, C->pre_b, C->base_ch);
        else if (strstr(name, "dec.post.weight")) ok = vec(t, C->post_w, stage_ch(4) * C->post_k);
        else if (strstr(name, "dec.post.bias")) ok = vec(t, C->post_b, 1);
        else if (strstr(name, "dec.post_act.alpha")) ok = vec(t, C->post_alpha, stage_ch(4) / 2, &C->n_alpha_post);
        else if (strstr(name, "dec.up.")) {
            int i = parse_idx(name, "dec.up.");
            if (i >= 0 && i < 5) {
                if (strstr(name, ".weight")) ok = vec(t, C->up_w[i], 2 * stage_ch(i) * 2 * C->up_rates[i]);
                else if (strstr(name, ".bias")) ok = vec(t, C->up_b[i], stage_ch(i));
            }
        } else if (strstr(name, "dec.act.") && strstr(name, "alpha")) {
            int i = parse_idx(name, "dec.act.");
            if (i >= 0 && i < 5) ok = vec(t, C->act_alpha[i], stage_ch(i), &C->n_alpha_act[i]);       // half of the 2 x stage_ch(i) input channels
        } else if (strstr(name, "dec.rl.")) {
            const char * p = strstr(name, "dec.rl.") + 7; int i = atoi(p);
            const char * p2 = strstr(p, ".rb."); if (!p2) continue; p2 += 4; int j = atoi(p2);
            const char * p3 = strstr(p2, ".rb."); if (!p3) continue; p3 += 4; int k = atoi(p3);
            if (i < 0 || i >= 5 || j < 0 || j >= 3 || k < 0 || k >= 3) continue;
            CodecResBlock & b = C->rb[i][j][k];
            const int64_t c = stage_ch(i), wn = c * c * C->res_k[j];
            if (strstr(name, ".in_act.alpha")) ok = vec(t, b.in_alpha, c / 2, &C->n_alpha_rb[i]);
            else if (strstr(name, ".in_conv.weight")) ok = vec(t, b.in_w, wn);
            else if (strstr(name, ".in_conv.bias")) ok = vec(t, b.in_b, c);
            else if (strstr(name, ".sk_act.alpha")) ok = vec(t, b.sk_alpha, c / 2);
            else if (strstr(name, ".sk_conv.weight")) ok = vec(t, b.sk_w, wn);
            else if (strstr(name, ".sk_conv.bias")) ok = vec(t, b.sk_b, c);
        }
        if (!ok || !up.ok) return nullptr;
    }
    bool complete = C->pre_w && C->pre_b && C->post_w && C->post_b && C->post_alpha;
    for (int i = 0; i < 5; i++) {
        complete = complete && C->act_alpha[i] && C->up_w[i] && C->up_b[i];
        for (int j = 0; j < 3; j++)
            for (int k = 0; k < 3; k++) {
                const CodecResBlock & b = C->rb[i][j][k];
                complete = complete && b.in_alpha && b.in_w && b.in_b && b.sk_alpha && b.sk_w && b.sk_b;
            }
    }
    if (!complete) { set_error("magpie_codec_init: codec file is missing decoder tensors"); return nullptr; }
    cudaStream_t st; cudaEvent_t e0, e1;
    if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) {
        set_error("magpie_codec_init: stream/event creation failed");
        return nullptr;
    }
    C->stream = st; C->ev0 = e0; C->ev1 = e1;
    return C.release();
}

}  // namespace mgb
