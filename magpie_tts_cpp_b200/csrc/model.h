// Device-resident model: weights uploaded once, laid out for the kernels.
// Tensor names / slots follow the reference loader (src/magpie.cpp:501-667); layout is fixed at
// load time so no per-step weight copies are needed (the reference re-`cont`s FFN weights every
// step, magpie.cpp:1793-1804).
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "../../include/magpie_b200.h"
#include "gguf_reader.h"

namespace mgb {

// Row-major [N][K] matrix in the model's weight dtype (f32 or bf16).  `taps` > 1: conv kernel
// re-laid out as [taps][N][K] (tap k multiplies x[t - (taps-1-k)], magpie.cpp:1825-1914).
struct DevMat {
    void * w = nullptr;
    int    N = 0, K = 0, taps = 1;
    void * tiles = nullptr;      // bf16 models: [ceil(N/128)][K/64] tiles of 128 x 64 in the tcgen05 shared-memory image (gemm_tc.cuh)
    void * tiles16 = nullptr;    // the same image with the bf16 values re-encoded as f16 (exact above 6.1e-5): operands of the f16-activation GEMMs
};

struct EncLayer { float * norm_self = nullptr, * norm_ff = nullptr; DevMat qkv, o, ff1, ff2; };
struct DecLayer {
    float * norm_self = nullptr, * norm_xa_q = nullptr, * norm_xa_mem = nullptr, * norm_ff = nullptr;
    DevMat qkv, o, xq, xkv, xo, ff1, ff2;
    float * qkv_csum = nullptr;             // bf16 models: csum[n] = sum_k Wqkv[n][k] norm_self[k] (LayerNorm folded through the QKV GEMM, kernels.cuh)
    float * ff1_csum = nullptr;             // same for W1 with norm_ff (LayerNorm folded through the FFN's first GEMM)
};

struct Model {
    mgb_hparams hp{};
    int device = 0;
    int precision = MGB_PREC_F32;
    int gelu_f16 = 1;
    size_t wsize = 4;                       // bytes per weight element

    // f32 tables (gathered rows only -> bandwidth-irrelevant, kept exact)
    float * text_emb = nullptr;             // [text_vocab][d]
    float * audio_emb[8] = {};              // [V][d] each
    float * baked_ctx = nullptr;            // [speakers][C*d]
    float * enc_pos = nullptr, * dec_pos = nullptr, * lt_pos = nullptr;
    int enc_pos_rows = 0, dec_pos_rows = 0, lt_pos_rows = 0;
    float * enc_norm_out = nullptr, * dec_norm_out = nullptr;
    std::vector<EncLayer> enc;
    std::vector<DecLayer> dec;
    DevMat final_w; float * final_b = nullptr;
    DevMat lt_in_w; float * lt_in_b = nullptr;
    float * lt_norm_self = nullptr, * lt_norm_ff = nullptr;
    DevMat lt_qkv, lt_o, lt_ff1, lt_ff2;
    DevMat lt_out_w[8]; float * lt_out_b[8] = {};
    float * lt_qkv_tab = nullptr;           // f32 [7][V][3*lt_dim]: [q | k | Wo Wv n] of LT position cb+1 for every fed code of codebook cb (row gather instead of LN + QKV GEMV)
    void * lt_qkvo = nullptr;               // bf16 [4*lt_dim][lt_dim]: [Wq; Wk; hi(Wo Wv); lo(Wo Wv)] (frame_loop.cu: the O-projection folded into V)
    float * lt_in_table[8] = {};            // P_cb = E_cb . Win^T + b  [V][lt_dim] f32 (built on the device at load)

    std::map<std::string, std::string> meta_str;    // tokenizer strings etc.
    std::map<std::string, int32_t>     meta_u32;
    std::vector<void *> allocations;
    int64_t step_weight_bytes = 0;

    ~Model();
};

// Loads + uploads. Returns nullptr and sets the thread's error on failure.
Model * load_model(const char * path, int device, int precision);

// ---- nano-codec weights (src/nano-codec.cpp:84-199 name mapping) ------------------------------
struct CodecResBlock {
    float * in_alpha = nullptr, * in_w = nullptr, * in_b = nullptr;
    float * sk_alpha = nullptr, * sk_w = nullptr, * sk_b = nullptr;
    void * in_w16 = nullptr, * sk_w16 = nullptr;       // tap-major padded f16 copies (CUDA-core direct conv)
    void * in_wt = nullptr, * sk_wt = nullptr;         // f16 tile images [split][chunk][tap][npad x 64] (tcgen05 path, codec_tc.h)
};
struct Codec {
    mgb_codec_hparams hp{};
    int device = 0;
    int latent = 32, base_ch = 864, pre_k = 7, post_k = 3;
    int up_rates[5] = {8, 8, 4, 2, 2};                 // magpie.h:672
    int res_k[3] = {3, 7, 11}, res_dil[3] = {1, 3, 5};
    float * pre_w = nullptr, * pre_b = nullptr, * post_alpha = nullptr, * post_w = nullptr, * post_b = nullptr;
    float * act_alpha[5] = {}, * up_w[5] = {}, * up_b[5] = {};
    int n_alpha_act[5] = {}, n_alpha_post = 0;
    CodecResBlock rb[5][3][3];
    int n_alpha_rb[5] = {};
    std::vector<void *> allocations;
    // scratch, grown on demand
    float * buf[6] = {}; size_t buf_elems = 0;
    void * pre_w16 = nullptr;
    void * img[4] = {}; size_t img_bytes = 0;   // tcgen05 path: time-major f16 activation images (codec_tc.h)
    bool tc_packed = false;
    int tc_mode = -1;            // 1 tensor-core pipeline, 0 CUDA-core pipeline (MGB_CODEC_NO_TC=1), -1 undecided
    int32_t * d_codes = nullptr; size_t codes_cap = 0;
    float * d_pcm = nullptr; size_t pcm_cap = 0;
    void * stream = nullptr;     // cudaStream_t
    void * ev0 = nullptr, * ev1 = nullptr;
    std::vector<void *> group_events;     // one per utterance group of a decode call (D2H of a group overlaps later groups)
    float last_ms = 0.0f; int64_t last_launches = 0;
    ~Codec();
};
Codec * load_codec(const char * path, int device);

}  // namespace mgb
