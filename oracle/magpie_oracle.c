/*
 * magpie_oracle.c -- CPU restatement of the magpie-tts.cpp per-frame synthesis hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (magpie_tts_cpp_b200/, include/) may link,
 * import or call this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / CPU baseline.
 *
 * PARITY UNPINNED: the reference's arithmetic lives in ggml (ggml-org/ggml, un-vendored and
 * unpinned: reference README.md:60-66, .gitignore:33-34) which is absent from this image, and the
 * reference's golden tensors (test_data/reference/, .bin files) are git-ignored and absent.  This file
 * therefore restates (a) the reference's graph builders op for op and (b) upstream ggml's CPU
 * rounding points from its published algorithm (f64 LayerNorm/softmax sums, f16 GELU table,
 * f16 im2col for conv_1d, f16 / Q8_0 activation rounding in mul_mat).  What *is* pinned:
 * orc_fsq_dequantize and orc_sample_top_k against the REFERENCE'S OWN fsq_dequantize_cpu / sample_top_k
 * (compiled from /root/reference by oracle/build_ref_pieces.py, golden vectors in tests/golden/ref_pieces.json,
 * tests/test_golden.py: bit-exact / identical picks), the FSQ formula and constants (nano-codec.cpp:729-742,
 * tests/test_codec_fsq.cpp:42-74), forbidden-token ids and the EOS rule (magpie.cpp:1131-1145, 4341-4348), the 1/8 audio
 * embedding scale (magpie.cpp:2769-2770); tests/ additionally cross-checks every function here
 * against an independent float64 torch restatement.
 *
 * Each function cites the reference file:line it follows (paths relative to the reference repo).
 * Plain C11 + OpenMP; build: see oracle/Makefile.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

enum { ORC_TYPE_F32 = 0, ORC_TYPE_F16 = 1, ORC_TYPE_Q8_0 = 8 };  /* ggml type ids */
enum { ACT_F32 = 0, ACT_F16 = 1, ACT_Q8_0 = 2 };                  /* how mul_mat rounds its activation */

#define MAX_LAYERS 64
#define NUM_CB 8

typedef struct {
    float * w;      /* dequantised to f32, row-major, PyTorch shape */
    int64_t ne[4];  /* ggml order: ne[0] fastest */
    int     n_dims;
    int     act;    /* ACT_* implied by the stored type (ggml mul_mat semantics) */
} mat_t;

typedef struct {
    int32_t d_model, d_ffn, d_head;
    int32_t enc_layers, enc_heads, enc_kernel;
    int32_t dec_layers, dec_sa_heads, dec_xa_heads, dec_xa_d_head, dec_kernel;
    int32_t lt_dim, lt_ffn_dim, lt_layers, lt_heads;
    int32_t text_vocab_size, num_codebooks, codebook_size, vocab_per_cb;
    int32_t num_speakers, context_frames;
    int32_t text_bos_id, text_eos_id, audio_bos_id, audio_eos_id;
    int32_t max_dec_steps, sample_rate;
    float   eps;
} orc_hparams;   /* field order == reference magpie_hparams (magpie.h:35-80) */

typedef struct { mat_t norm_self, qkv, o, norm_ff, ff_proj, ff_out; } enc_layer_t;
typedef struct {
    mat_t norm_self, qkv, o, norm_xa_q, xa_q, xa_kv, xa_o, norm_xa_mem, norm_ff, ff_proj, ff_out;
} dec_layer_t;

typedef struct {
    orc_hparams hp;
    mat_t text_emb, audio_emb[NUM_CB], baked_ctx;
    mat_t enc_pos, enc_norm_out;  enc_layer_t enc[MAX_LAYERS];
    mat_t dec_pos, dec_norm_out;  dec_layer_t dec[MAX_LAYERS];
    mat_t final_w, final_b;
    mat_t lt_in_w, lt_in_b, lt_pos, lt_norm_self, lt_qkv, lt_o, lt_norm_ff, lt_ff_proj, lt_ff_out;
    mat_t lt_out_w[NUM_CB], lt_out_b[NUM_CB];
    int gelu_f16_table;   /* 1 = emulate ggml's f16 GELU lookup table (default) */
} orc_model;

/* ------------------------------------------------------------------------------------------ */
/* f16 helpers and ggml-CPU elementwise semantics                                             */
/* ------------------------------------------------------------------------------------------ */

static inline float f16_round(float x) { return (float)(_Float16)x; }

static float    g_gelu_tab[65536];
static int      g_gelu_ready = 0;

static inline float gelu_f32(float x) {
    /* ggml_gelu_f32 [ggml-upstream]: tanh approximation */
    const float a = 0.044715f, s = 0.79788456080286535587989211986876f;
    return 0.5f * x * (1.0f + tanhf(s * x * (1.0f + a * x * x)));
}

static void gelu_table_init(void) {
    if (g_gelu_ready) return;
    for (uint32_t i = 0; i < 65536; i++) {
        uint16_t h = (uint16_t)i; _Float16 hf; memcpy(&hf, &h, 2);
        g_gelu_tab[i] = f16_round(gelu_f32((float)hf));
    }
    g_gelu_ready = 1;
}

/* ggml_vec_gelu_f32 with GGML_GELU_FP16 [ggml-upstream]: y = f32(f16(gelu(f32(f16(x))))) */
static inline float gelu_ggml(float x, int use_table) {
    if (!use_table) return gelu_f32(x);
    if (x <= -10.0f) return 0.0f;
    if (x >= 10.0f) return x;
    _Float16 hf = (_Float16)x; uint16_t h; memcpy(&h, &hf, 2);
    return g_gelu_tab[h];
}

/* ggml_norm + ggml_mul (magpie.cpp:2237-2259); sums in double [ggml-upstream] */
static void layer_norm(const float * x, const float * w, float * y, int n, float eps) {
    double sum = 0.0;
    for (int i = 0; i < n; i++) sum += (double)x[i];
    float mean = (float)(sum / n);
    double sum2 = 0.0;
    for (int i = 0; i < n; i++) { float v = x[i] - mean; y[i] = v; sum2 += (double)(v * v); }
    float variance = (float)(sum2 / n);
    const float scale = 1.0f / sqrtf(variance + eps);
    for (int i = 0; i < n; i++) y[i] = (y[i] * scale) * w[i];
}

/* ggml_soft_max over n scores, already scaled/masked [ggml-upstream]: max-subtract, expf, f64 sum */
static void soft_max(float * s, int n) {
    float mx = -INFINITY;
    for (int i = 0; i < n; i++) if (s[i] > mx) mx = s[i];
    double sum = 0.0;
    for (int i = 0; i < n; i++) {
        float e = (s[i] == -INFINITY) ? 0.0f : expf(s[i] - mx);
        s[i] = e; sum += (double)e;
    }
    float inv = (float)(1.0 / sum);
    for (int i = 0; i < n; i++) s[i] *= inv;
}

static inline float dot_f32(const float * restrict a, const float * restrict b, int n) {
    float acc[16] = {0};
    int i = 0;
    for (; i + 16 <= n; i += 16)
        for (int j = 0; j < 16; j++) acc[j] += a[i + j] * b[i + j];
    float s = 0.0f;
    for (int j = 0; j < 16; j++) s += acc[j];
    for (; i < n; i++) s += a[i] * b[i];
    return s;
}

/* Activation rounding that ggml_mul_mat applies to src1 for a given src0 type [ggml-upstream]. */
static void round_activation(const float * x, float * y, int n, int act) {
    if (act == ACT_F32) { memcpy(y, x, (size_t)n * sizeof(float)); return; }
    if (act == ACT_F16) { for (int i = 0; i < n; i++) y[i] = f16_round(x[i]); return; }
    /* Q8_0: per 32-block d = amax/127 (stored f16), q = roundf(x/d) */
    for (int b = 0; b < n; b += 32) {
        int m = n - b < 32 ? n - b : 32;
        float amax = 0.0f;
        for (int j = 0; j < m; j++) { float a = fabsf(x[b + j]); if (a > amax) amax = a; }
        float d = amax / 127.0f, id = d != 0.0f ? 1.0f / d : 0.0f;
        float d16 = f16_round(d);
        for (int j = 0; j < m; j++) y[b + j] = roundf(x[b + j] * id) * d16;
    }
}

/* ggml-CPU's mul_mat rounds the ACTIVATION row to the weight's storage type before the dot (f16 / Q8_0 blocks).  The switch
 * below turns that rounding off (weights still dequantised exactly), which isolates the WEIGHT format from ggml's activation
 * noise: the product keeps f32 activations, so against this mode it is held to the plain f32 / bf16 bars, while its distance
 * to the full ggml semantics is reported separately (tests/test_gpu_parity.py::test_quantised_gguf_weights_match_oracle). */
static int g_activation_rounding = 1;
ORC_API void orc_set_activation_rounding(int on) { g_activation_rounding = on; }

/* Y[t][n] = sum_k W[n][k] X[t][k] (+bias[n]);  ggml_mul_mat(W, X).  ldx/ldy are row strides. */
static void linear(const mat_t * W, const float * bias, const float * X, int ldx, float * Y, int ldy,
                   int T, int N, int K) {
    float * Xr = (float *)X; int ldr = ldx; float * tmp = NULL;
    if (W->act != ACT_F32 && g_activation_rounding) {
        tmp = (float *)malloc((size_t)T * K * sizeof(float));
        for (int t = 0; t < T; t++) round_activation(X + (size_t)t * ldx, tmp + (size_t)t * K, K, W->act);
        Xr = tmp; ldr = K;
    }
    const float * w = W->w;
#pragma omp parallel for schedule(static) if ((int64_t)N * K * T > 65536)
    for (int n = 0; n < N; n++) {
        const float * wr = w + (size_t)n * K;
        float b = bias ? bias[n] : 0.0f;
        for (int t = 0; t < T; t++) {
            float v = dot_f32(wr, Xr + (size_t)t * ldr, K);
            Y[(size_t)t * ldy + n] = bias ? v + b : v;
        }
    }
    free(tmp);
}

/* ------------------------------------------------------------------------------------------ */
/* model construction: name -> slot mapping (magpie.cpp:501-562, 607-667)                      */
/* ------------------------------------------------------------------------------------------ */

ORC_API orc_model * orc_model_new(const orc_hparams * hp) {
    orc_model * m = (orc_model *)calloc(1, sizeof(orc_model));
    m->hp = *hp;
    m->gelu_f16_table = 1;
    gelu_table_init();
    return m;
}

ORC_API void orc_model_set_gelu_table(orc_model * m, int on) { m->gelu_f16_table = on; }

static void mat_free(mat_t * t) { free(t->w); t->w = NULL; }

ORC_API void orc_model_free(orc_model * m) {
    if (!m) return;
    mat_t * p = (mat_t *)&m->text_emb;
    /* every mat_t between text_emb and lt_out_b is contiguous in the struct */
    size_t n = ((char *)&m->gelu_f16_table - (char *)&m->text_emb) / sizeof(mat_t);
    for (size_t i = 0; i < n; i++) mat_free(&p[i]);
    free(m);
}

static int64_t nelem(const int64_t * ne, int nd) { int64_t n = 1; for (int i = 0; i < nd; i++) n *= ne[i]; return n; }

/* Dequantise a GGUF tensor payload to f32 (Q8_0 block = f16 d + 32 x int8, along ne[0]). */
static void mat_fill(mat_t * t, const void * data, int type, int n_dims, const int64_t * ne) {
    free(t->w);
    int64_t n = nelem(ne, n_dims);
    t->w = (float *)malloc((size_t)n * sizeof(float));
    t->n_dims = n_dims;
    for (int i = 0; i < 4; i++) t->ne[i] = i < n_dims ? ne[i] : 1;
    if (type == ORC_TYPE_F32) { memcpy(t->w, data, (size_t)n * 4); t->act = ACT_F32; }
    else if (type == ORC_TYPE_F16) {
        const _Float16 * h = (const _Float16 *)data;
        for (int64_t i = 0; i < n; i++) t->w[i] = (float)h[i];
        t->act = ACT_F16;
    } else if (type == ORC_TYPE_Q8_0) {
        const uint8_t * p = (const uint8_t *)data;
        for (int64_t b = 0; b < n / 32; b++, p += 34) {
            _Float16 d; memcpy(&d, p, 2);
            const int8_t * q = (const int8_t *)(p + 2);
            for (int j = 0; j < 32; j++) t->w[b * 32 + j] = (float)d * (float)q[j];
        }
        t->act = ACT_Q8_0;
    } else { fprintf(stderr, "oracle: unsupported tensor type %d\n", type); abort(); }
}

static int parse_idx(const char * name, const char * prefix) {
    const char * p = strstr(name, prefix);
    return p ? atoi(p + strlen(prefix)) : -1;
}

ORC_API int orc_model_set_tensor(orc_model * m, const char * name, const void * data, int type,
                                 int n_dims, const int64_t * ne) {
    mat_t * slot = NULL;
    const orc_hparams * hp = &m->hp;
    /* same precedence as the reference loader, including "encoder.layers." matching before
       "decoder.layers." (so context_encoder.layers.* would land in encoder.layers[N]) */
    if (!strcmp(name, "text_embedding.weight")) slot = &m->text_emb;
    else if (strstr(name, "audio_embeddings.")) {
        int cb = parse_idx(name, "audio_embeddings.");
        if (cb >= 0 && cb < NUM_CB) slot = &m->audio_emb[cb];
    } else if (!strcmp(name, "baked_context_embedding.weight")) slot = &m->baked_ctx;
    else if (!strcmp(name, "encoder.position_embeddings.weight")) slot = &m->enc_pos;
    else if (strstr(name, "encoder.layers.")) {
        int l = parse_idx(name, "encoder.layers.");
        if (l >= 0 && l < hp->enc_layers) {
            enc_layer_t * L = &m->enc[l];
            if (strstr(name, "norm_self.weight")) slot = &L->norm_self;
            else if (strstr(name, "self_attention.qkv_net.weight")) slot = &L->qkv;
            else if (strstr(name, "self_attention.o_net.weight")) slot = &L->o;
            else if (strstr(name, "norm_pos_ff.weight")) slot = &L->norm_ff;
            else if (strstr(name, "pos_ff.proj.conv.weight")) slot = &L->ff_proj;
            else if (strstr(name, "pos_ff.o_net.conv.weight")) slot = &L->ff_out;
        }
    } else if (!strcmp(name, "encoder.norm_out.weight")) slot = &m->enc_norm_out;
    else if (!strcmp(name, "decoder.position_embeddings.weight")) slot = &m->dec_pos;
    else if (strstr(name, "decoder.layers.")) {
        int l = parse_idx(name, "decoder.layers.");
        if (l >= 0 && l < hp->dec_layers) {
            dec_layer_t * L = &m->dec[l];
            if (strstr(name, "norm_self.weight")) slot = &L->norm_self;
            else if (strstr(name, "self_attention.qkv_net.weight")) slot = &L->qkv;
            else if (strstr(name, "self_attention.o_net.weight")) slot = &L->o;
            else if (strstr(name, "norm_xattn_query.weight")) slot = &L->norm_xa_q;
            else if (strstr(name, "cross_attention.q_net.weight")) slot = &L->xa_q;
            else if (strstr(name, "cross_attention.kv_net.weight")) slot = &L->xa_kv;
            else if (strstr(name, "cross_attention.o_net.weight")) slot = &L->xa_o;
            else if (strstr(name, "norm_xattn_memory.weight")) slot = &L->norm_xa_mem;
            else if (strstr(name, "norm_pos_ff.weight")) slot = &L->norm_ff;
            else if (strstr(name, "pos_ff.proj.conv.weight")) slot = &L->ff_proj;
            else if (strstr(name, "pos_ff.o_net.conv.weight")) slot = &L->ff_out;
        }
    } else if (!strcmp(name, "decoder.norm_out.weight")) slot = &m->dec_norm_out;
    else if (!strcmp(name, "final_proj.weight")) slot = &m->final_w;
    else if (!strcmp(name, "final_proj.bias")) slot = &m->final_b;
    else if (strstr(name, "local_transformer_in_projection.weight")) slot = &m->lt_in_w;
    else if (strstr(name, "local_transformer_in_projection.bias")) slot = &m->lt_in_b;
    else if (!strcmp(name, "local_transformer.position_embeddings.weight")) slot = &m->lt_pos;
    else if (strstr(name, "local_transformer.layers.0.norm_self.weight")) slot = &m->lt_norm_self;
    else if (strstr(name, "local_transformer.layers.0.self_attention.qkv_net.weight")) slot = &m->lt_qkv;
    else if (strstr(name, "local_transformer.layers.0.self_attention.o_net.weight")) slot = &m->lt_o;
    else if (strstr(name, "local_transformer.layers.0.norm_pos_ff.weight")) slot = &m->lt_norm_ff;
    else if (strstr(name, "local_transformer.layers.0.pos_ff.proj.conv.weight")) slot = &m->lt_ff_proj;
    else if (strstr(name, "local_transformer.layers.0.pos_ff.o_net.conv.weight")) slot = &m->lt_ff_out;
    else if (strstr(name, "local_transformer_out_projections.")) {
        int cb = parse_idx(name, "local_transformer_out_projections.");
        if (cb >= 0 && cb < NUM_CB) {
            if (strstr(name, ".weight")) slot = &m->lt_out_w[cb];
            else if (strstr(name, ".bias")) slot = &m->lt_out_b[cb];
        }
    }
    if (!slot) return 0;
    mat_fill(slot, data, type, n_dims, ne);
    return 1;
}

/* ------------------------------------------------------------------------------------------ */
/* attention (magpie.cpp:1477-1575 full-seq causal; 3395-3480 single step; 1713-1767 cross)    */
/* ------------------------------------------------------------------------------------------ */

/* q: [Tq][H*dh], K,V: rows of H*dh with row strides ldk; query i may see keys j <= i + key_off
   (causal) or all nk keys (causal=0). out: [Tq][H*dh]. */
static void attention(const float * q, int ldq, const float * K, const float * V, int ldk,
                      float * out, int ldo, int Tq, int nk, int H, int dh, int causal, int key_off) {
    const float scale = 1.0f / sqrtf((float)dh);
#pragma omp parallel for collapse(2) schedule(static) if (Tq * H > 4)
    for (int i = 0; i < Tq; i++) {
        for (int h = 0; h < H; h++) {
            int n = causal ? (i + key_off + 1 < nk ? i + key_off + 1 : nk) : nk;
            float * s = (float *)malloc((size_t)n * sizeof(float));
            const float * qi = q + (size_t)i * ldq + h * dh;
            for (int j = 0; j < n; j++) s[j] = dot_f32(K + (size_t)j * ldk + h * dh, qi, dh) * scale;
            soft_max(s, n);
            float * o = out + (size_t)i * ldo + h * dh;
            for (int d = 0; d < dh; d++) o[d] = 0.0f;
            for (int j = 0; j < n; j++) {
                const float * vj = V + (size_t)j * ldk + h * dh;
                float p = s[j];
                for (int d = 0; d < dh; d++) o[d] += p * vj[d];
            }
            free(s);
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* conv-FFN (magpie.cpp:1769-1918): kernel 1 = two linears with GELU; kernel 3 = causal conv   */
/* ------------------------------------------------------------------------------------------ */

/* conv weight PyTorch (out,in,k) -> per-tap dense [k][out][in] so each tap is a plain linear */
static float * conv_taps(const mat_t * W, int out, int in, int k) {
    float * t = (float *)malloc((size_t)out * in * k * sizeof(float));
    for (int o = 0; o < out; o++)
        for (int i = 0; i < in; i++)
            for (int kk = 0; kk < k; kk++)
                t[((size_t)kk * out + o) * in + i] = W->w[((size_t)o * in + i) * k + kk];
    return t;
}

static void conv_ffn(const orc_model * m, const mat_t * proj, const mat_t * outw, const float * x,
                     float * y, int T, int d, int f, int ksize) {
    float * h = (float *)calloc((size_t)T * f, sizeof(float));
    if (ksize == 1) {
        linear(proj, NULL, x, d, h, f, T, f, d);
        for (size_t i = 0; i < (size_t)T * f; i++) h[i] = gelu_ggml(h[i], m->gelu_f16_table);
        linear(outw, NULL, h, f, y, d, T, d, f);
    } else {
        /* hidden[t] = sum_k W[:,:,k] x[t-(ksize-1)+k], zero left padding; terms added k=0,1,2 */
        float * taps = conv_taps(proj, f, d, ksize);
        float * tmp = (float *)malloc((size_t)T * f * sizeof(float));
        for (int k = 0; k < ksize; k++) {
            int sh = ksize - 1 - k;           /* x[t - sh] */
            mat_t Wk = { taps + (size_t)k * f * d, {d, f, 1, 1}, 2, proj->act };
            if (T - sh <= 0) continue;
            linear(&Wk, NULL, x, d, tmp + (size_t)sh * f, f, T - sh, f, d);
            for (int t = sh; t < T; t++)
                for (int j = 0; j < f; j++) h[(size_t)t * f + j] += tmp[(size_t)t * f + j];
        }
        free(taps);
        for (size_t i = 0; i < (size_t)T * f; i++) h[i] = gelu_ggml(h[i], m->gelu_f16_table);
        memset(y, 0, (size_t)T * d * sizeof(float));
        taps = conv_taps(outw, d, f, ksize);
        float * tmp2 = (float *)malloc((size_t)T * d * sizeof(float));
        for (int k = 0; k < ksize; k++) {
            int sh = ksize - 1 - k;
            mat_t Wk = { taps + (size_t)k * d * f, {f, d, 1, 1}, 2, outw->act };
            if (T - sh <= 0) continue;
            linear(&Wk, NULL, h, f, tmp2 + (size_t)sh * d, d, T - sh, d, f);
            for (int t = sh; t < T; t++)
                for (int j = 0; j < d; j++) y[(size_t)t * d + j] += tmp2[(size_t)t * d + j];
        }
        free(taps); free(tmp); free(tmp2);
    }
    free(h);
}

/* ------------------------------------------------------------------------------------------ */
/* text encoder (magpie.cpp:2284-2374, 1960-1995, 1929-1958)                                   */
/* ------------------------------------------------------------------------------------------ */

ORC_API int orc_encode_text(const orc_model * m, const int32_t * tokens, int E, float * enc_out) {
    const orc_hparams * hp = &m->hp;
    const int d = hp->d_model;
    if (E <= 0) return 0;
    float * x = (float *)malloc((size_t)E * d * sizeof(float));
    float * n = (float *)malloc((size_t)E * d * sizeof(float));
    float * qkv = (float *)malloc((size_t)E * 3 * d * sizeof(float));
    float * a = (float *)malloc((size_t)E * d * sizeof(float));
    float * y = (float *)malloc((size_t)E * d * sizeof(float));
    for (int t = 0; t < E; t++)
        for (int i = 0; i < d; i++)
            x[(size_t)t * d + i] = m->text_emb.w[(size_t)tokens[t] * d + i] + m->enc_pos.w[(size_t)t * d + i];
    for (int l = 0; l < hp->enc_layers; l++) {
        const enc_layer_t * L = &m->enc[l];
        for (int t = 0; t < E; t++) layer_norm(x + (size_t)t * d, L->norm_self.w, n + (size_t)t * d, d, hp->eps);
        linear(&L->qkv, NULL, n, d, qkv, 3 * d, E, 3 * d, d);
        /* the NeMo encoder is causal (magpie.cpp:1948) */
        attention(qkv, 3 * d, qkv + d, qkv + 2 * d, 3 * d, a, d, E, E, hp->enc_heads, d / hp->enc_heads, 1, 0);
        linear(&L->o, NULL, a, d, y, d, E, d, d);
        for (size_t i = 0; i < (size_t)E * d; i++) x[i] = y[i] + x[i];
        for (int t = 0; t < E; t++) layer_norm(x + (size_t)t * d, L->norm_ff.w, n + (size_t)t * d, d, hp->eps);
        conv_ffn(m, &L->ff_proj, &L->ff_out, n, y, E, d, hp->d_ffn, hp->enc_kernel);
        for (size_t i = 0; i < (size_t)E * d; i++) x[i] = y[i] + x[i];
    }
    for (int t = 0; t < E; t++) layer_norm(x + (size_t)t * d, m->enc_norm_out.w, enc_out + (size_t)t * d, d, hp->eps);
    free(x); free(n); free(qkv); free(a); free(y);
    return 1;
}

/* ------------------------------------------------------------------------------------------ */
/* decoder state: KV caches, cross-attention K/V, context prefill                              */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    const orc_model * m;
    int E, max_seq, pos;
    float * k, * v;        /* [layer][max_seq][d]  (magpie.cpp:3315-3376) */
    float * xk, * xv;      /* [layer][E][dxa] */
} orc_state;

/* one or more decoder positions through all layers; x: [T][d] already has pos-emb added.
   Writes K/V at [start, start+T) and attends causally (magpie.cpp:3484-3528, 3991-4060). */
static void decoder_layers(orc_state * s, float * x, int T, int start, float * hidden_out) {
    const orc_model * m = s->m; const orc_hparams * hp = &m->hp;
    const int d = hp->d_model, dxa = hp->dec_xa_heads * hp->dec_xa_d_head;
    float * n = (float *)malloc((size_t)T * d * sizeof(float));
    float * qkv = (float *)malloc((size_t)T * 3 * d * sizeof(float));
    float * a = (float *)malloc((size_t)T * d * sizeof(float));
    float * y = (float *)malloc((size_t)T * d * sizeof(float));
    float * xq = (float *)malloc((size_t)T * dxa * sizeof(float));
    float * xa = (float *)malloc((size_t)T * dxa * sizeof(float));
    for (int l = 0; l < hp->dec_layers; l++) {
        const dec_layer_t * L = &m->dec[l];
        float * kc = s->k + (size_t)l * s->max_seq * d, * vc = s->v + (size_t)l * s->max_seq * d;
        for (int t = 0; t < T; t++) layer_norm(x + (size_t)t * d, L->norm_self.w, n + (size_t)t * d, d, hp->eps);
        linear(&L->qkv, NULL, n, d, qkv, 3 * d, T, 3 * d, d);
        for (int t = 0; t < T; t++) {
            memcpy(kc + (size_t)(start + t) * d, qkv + (size_t)t * 3 * d + d, (size_t)d * 4);
            memcpy(vc + (size_t)(start + t) * d, qkv + (size_t)t * 3 * d + 2 * d, (size_t)d * 4);
        }
        attention(qkv, 3 * d, kc, vc, d, a, d, T, start + T, hp->dec_sa_heads, d / hp->dec_sa_heads, 1, start);
        linear(&L->o, NULL, a, d, y, d, T, d, d);
        for (size_t i = 0; i < (size_t)T * d; i++) x[i] = y[i] + x[i];
        /* cross-attention over cached encoder K/V, no mask (magpie.cpp:1713-1767) */
        for (int t = 0; t < T; t++) layer_norm(x + (size_t)t * d, L->norm_xa_q.w, n + (size_t)t * d, d, hp->eps);
        linear(&L->xa_q, NULL, n, d, xq, dxa, T, dxa, d);
        attention(xq, dxa, s->xk + (size_t)l * s->E * dxa, s->xv + (size_t)l * s->E * dxa, dxa, xa, dxa,
                  T, s->E, hp->dec_xa_heads, hp->dec_xa_d_head, 0, 0);
        linear(&L->xa_o, NULL, xa, dxa, y, d, T, d, dxa);
        for (size_t i = 0; i < (size_t)T * d; i++) x[i] = y[i] + x[i];
        for (int t = 0; t < T; t++) layer_norm(x + (size_t)t * d, L->norm_ff.w, n + (size_t)t * d, d, hp->eps);
        conv_ffn(m, &L->ff_proj, &L->ff_out, n, y, T, d, hp->d_ffn, hp->dec_kernel);
        for (size_t i = 0; i < (size_t)T * d; i++) x[i] = y[i] + x[i];
    }
    if (hidden_out)
        for (int t = 0; t < T; t++)
            layer_norm(x + (size_t)t * d, m->dec_norm_out.w, hidden_out + (size_t)t * d, d, hp->eps);
    free(n); free(qkv); free(a); free(y); free(xq); free(xa);
}

/* Steps 2-5 of magpie_synthesize_codes_graph_reuse (magpie.cpp:4089-4243): KV cache, per-layer
   cross-attention K/V (magpie.cpp:1663-1711), baked-context prefill.  max_seq <= 0 uses the
   reference's context_frames + max_dec_steps + 16. */
ORC_API orc_state * orc_state_new(const orc_model * m, const float * enc_out, int E, int speaker, int max_seq) {
    const orc_hparams * hp = &m->hp;
    const int d = hp->d_model, dxa = hp->dec_xa_heads * hp->dec_xa_d_head, C = hp->context_frames;
    orc_state * s = (orc_state *)calloc(1, sizeof(orc_state));
    s->m = m; s->E = E;
    s->max_seq = max_seq > 0 ? max_seq : C + hp->max_dec_steps + 16;
    s->k = (float *)calloc((size_t)hp->dec_layers * s->max_seq * d, sizeof(float));
    s->v = (float *)calloc((size_t)hp->dec_layers * s->max_seq * d, sizeof(float));
    s->xk = (float *)malloc((size_t)hp->dec_layers * E * dxa * sizeof(float));
    s->xv = (float *)malloc((size_t)hp->dec_layers * E * dxa * sizeof(float));
    float * n = (float *)malloc((size_t)E * d * sizeof(float));
    float * kv = (float *)malloc((size_t)E * 2 * dxa * sizeof(float));
    for (int l = 0; l < hp->dec_layers; l++) {
        const dec_layer_t * L = &m->dec[l];
        for (int t = 0; t < E; t++) layer_norm(enc_out + (size_t)t * d, L->norm_xa_mem.w, n + (size_t)t * d, d, hp->eps);
        linear(&L->xa_kv, NULL, n, d, kv, 2 * dxa, E, 2 * dxa, d);
        for (int t = 0; t < E; t++) {   /* rows 0..dxa-1 = K, dxa..2dxa-1 = V */
            memcpy(s->xk + ((size_t)l * E + t) * dxa, kv + (size_t)t * 2 * dxa, (size_t)dxa * 4);
            memcpy(s->xv + ((size_t)l * E + t) * dxa, kv + (size_t)t * 2 * dxa + dxa, (size_t)dxa * 4);
        }
    }
    free(n); free(kv);
    /* prefill: baked context row viewed as [C][d] + pos[0..C) */
    float * x = (float *)malloc((size_t)C * d * sizeof(float));
    const float * ctx = m->baked_ctx.w + (size_t)speaker * C * d;
    for (int t = 0; t < C; t++)
        for (int i = 0; i < d; i++) x[(size_t)t * d + i] = ctx[(size_t)t * d + i] + m->dec_pos.w[(size_t)t * d + i];
    decoder_layers(s, x, C, 0, NULL);
    free(x);
    s->pos = C;
    return s;
}

ORC_API void orc_state_free(orc_state * s) {
    if (!s) return;
    free(s->k); free(s->v); free(s->xk); free(s->xv); free(s);
}

ORC_API int orc_state_pos(const orc_state * s) { return s->pos; }

/* raw KV access for tests: returns pointers into the state */
ORC_API const float * orc_state_k(const orc_state * s) { return s->k; }
ORC_API const float * orc_state_v(const orc_state * s) { return s->v; }
ORC_API const float * orc_state_xk(const orc_state * s) { return s->xk; }
ORC_API const float * orc_state_xv(const orc_state * s) { return s->xv; }
ORC_API int orc_state_max_seq(const orc_state * s) { return s->max_seq; }

/* emb = (sum_cb E_cb[code_cb]) * (1/8)  (magpie.cpp:2746-2787) */
ORC_API void orc_audio_embedding(const orc_model * m, const int32_t * codes, float * emb) {
    const int d = m->hp.d_model;
    for (int i = 0; i < d; i++) {
        float s = m->audio_emb[0].w[(size_t)codes[0] * d + i];
        for (int cb = 1; cb < NUM_CB; cb++) s = s + m->audio_emb[cb].w[(size_t)codes[cb] * d + i];
        emb[i] = s * (1.0f / 8.0f);
    }
}

/* one autoregressive step: x = emb(codes) + pos[p]; 12 layers; final LN (magpie.cpp:4366-4405) */
ORC_API void orc_decoder_step(orc_state * s, const int32_t * codes, float * hidden) {
    const orc_model * m = s->m; const int d = m->hp.d_model;
    float * x = (float *)malloc((size_t)d * sizeof(float));
    orc_audio_embedding(m, codes, x);
    for (int i = 0; i < d; i++) x[i] = x[i] + m->dec_pos.w[(size_t)s->pos * d + i];
    decoder_layers(s, x, 1, s->pos, hidden);
    s->pos++;
    free(x);
}

/* final_proj (magpie.cpp:2261-2282): logits[8*V] = W h + b */
ORC_API void orc_final_proj(const orc_model * m, const float * hidden, float * logits) {
    int N = m->hp.num_codebooks * m->hp.vocab_per_cb;
    linear(&m->final_w, m->final_b.w, hidden, m->hp.d_model, logits, N, 1, N, m->hp.d_model);
}

/* ------------------------------------------------------------------------------------------ */
/* local transformer + sampler (magpie.cpp:946-1048, 1072-1317)                                */
/* ------------------------------------------------------------------------------------------ */

typedef struct { float v; int i; } scored_t;
static int scored_cmp(const void * a, const void * b) {
    const scored_t * x = (const scored_t *)a, * y = (const scored_t *)b;
    if (x->v > y->v) return -1;
    if (x->v < y->v) return 1;
    return x->i - y->i;          /* tie: lowest index first (reference order is unspecified) */
}

/* sample_top_k (magpie.cpp:1072-1109) with the uniform draw u supplied by the caller */
ORC_API int32_t orc_sample_top_k(const float * logits, int n, float temperature, int top_k, float u) {
    scored_t * sc = (scored_t *)malloc((size_t)n * sizeof(scored_t));
    for (int i = 0; i < n; i++) { sc[i].v = logits[i]; sc[i].i = i; }
    qsort(sc, (size_t)n, sizeof(scored_t), scored_cmp);
    int k = top_k < n ? top_k : n;
    if (k < 1) k = 1;
    float * p = (float *)malloc((size_t)k * sizeof(float));
    float mx = sc[0].v, sum = 0.0f;
    for (int i = 0; i < k; i++) { p[i] = expf((sc[i].v - mx) / temperature); sum += p[i]; }
    for (int i = 0; i < k; i++) p[i] /= sum;
    float cum = 0.0f; int pick = sc[k - 1].i;
    for (int i = 0; i < k; i++) { cum += p[i]; if (u < cum) { pick = sc[i].i; break; } }
    free(sc); free(p);
    return pick;
}

/* LT layer over a prefix [n][L] (already + pos); returns hidden of the LAST position only. */
static void lt_layer_last(const orc_model * m, const float * seq, int n, float * last_hidden) {
    const orc_hparams * hp = &m->hp; const int L = hp->lt_dim;
    float * nm = (float *)malloc((size_t)n * L * sizeof(float));
    float * qkv = (float *)malloc((size_t)n * 3 * L * sizeof(float));
    float * a = (float *)malloc((size_t)n * L * sizeof(float));
    float * y = (float *)malloc((size_t)n * L * sizeof(float));
    float * x = (float *)malloc((size_t)n * L * sizeof(float));
    for (int t = 0; t < n; t++) layer_norm(seq + (size_t)t * L, m->lt_norm_self.w, nm + (size_t)t * L, L, hp->eps);
    linear(&m->lt_qkv, NULL, nm, L, qkv, 3 * L, n, 3 * L, L);
    attention(qkv, 3 * L, qkv + L, qkv + 2 * L, 3 * L, a, L, n, n, hp->lt_heads, L / hp->lt_heads, 1, 0);
    linear(&m->lt_o, NULL, a, L, y, L, n, L, L);
    for (size_t i = 0; i < (size_t)n * L; i++) x[i] = y[i] + seq[i];
    for (int t = 0; t < n; t++) layer_norm(x + (size_t)t * L, m->lt_norm_ff.w, nm + (size_t)t * L, L, hp->eps);
    conv_ffn(m, &m->lt_ff_proj, &m->lt_ff_out, nm, y, n, L, hp->lt_ffn_dim, 1);
    for (int i = 0; i < L; i++) last_hidden[i] = y[(size_t)(n - 1) * L + i] + x[(size_t)(n - 1) * L + i];
    free(nm); free(qkv); free(a); free(y); free(x);
}

/* magpie_local_transformer_sample_all (magpie.cpp:1113-1317).
   forced_codes != NULL: teacher forcing -- the code fed back for cb is forced_codes[cb].
   uniforms != NULL: u[cb] for sample_top_k (ignored when temperature < 0.01).
   logits_out != NULL: receives the 8 masked logit vectors [8][V]. */
ORC_API void orc_lt_sample(const orc_model * m, const float * hidden, float temperature, int top_k,
                           int forbid_eos, const int32_t * forced_codes, const float * uniforms,
                           int32_t * sampled, int32_t * argmax_out, float * logits_out) {
    const orc_hparams * hp = &m->hp; const int L = hp->lt_dim, d = hp->d_model, V = hp->vocab_per_cb;
    float * seq = (float *)malloc((size_t)(NUM_CB + 1) * L * sizeof(float));   /* without pos */
    float * wp = (float *)malloc((size_t)(NUM_CB + 1) * L * sizeof(float));    /* with pos */
    float * h = (float *)malloc((size_t)L * sizeof(float));
    float * logits = (float *)malloc((size_t)V * sizeof(float));
    linear(&m->lt_in_w, m->lt_in_b.w, hidden, d, seq, L, 1, L, d);
    for (int cb = 0; cb < NUM_CB; cb++) {
        int n = cb + 1;
        for (int t = 0; t < n; t++)
            for (int i = 0; i < L; i++) wp[(size_t)t * L + i] = seq[(size_t)t * L + i] + m->lt_pos.w[(size_t)t * L + i];
        lt_layer_last(m, wp, n, h);
        linear(&m->lt_out_w[cb], m->lt_out_b[cb].w, h, L, logits, V, 1, V, L);
        /* forbidden: BOS and BOS+2..BOS+7; EOS too while forbid_eos (magpie.cpp:1131-1145) */
        int forb[8] = { hp->audio_bos_id, hp->audio_bos_id + 2, hp->audio_bos_id + 3, hp->audio_bos_id + 4,
                        hp->audio_bos_id + 5, hp->audio_bos_id + 6, hp->audio_bos_id + 7,
                        forbid_eos ? hp->audio_eos_id : -1 };
        for (int j = 0; j < 8; j++) if (forb[j] >= 0 && forb[j] < V) logits[forb[j]] = -INFINITY;
        int am = 0; float mv = logits[0];
        for (int i = 1; i < V; i++) if (logits[i] > mv) { mv = logits[i]; am = i; }
        argmax_out[cb] = am;
        int pick = (temperature < 0.01f) ? am
                   : orc_sample_top_k(logits, V, temperature, top_k, uniforms ? uniforms[cb] : 0.5f);
        sampled[cb] = pick;
        if (logits_out) memcpy(logits_out + (size_t)cb * V, logits, (size_t)V * 4);
        if (cb < NUM_CB - 1) {
            int fed = forced_codes ? forced_codes[cb] : pick;
            /* feedback embedding is NOT scaled by 1/8 (magpie.cpp:1285-1291) */
            linear(&m->lt_in_w, m->lt_in_b.w, m->audio_emb[cb].w + (size_t)fed * d, d,
                   seq + (size_t)(cb + 1) * L, L, 1, L, d);
        }
    }
    free(seq); free(wp); free(h); free(logits);
}

/* Full loop of magpie_synthesize_codes_graph_reuse (magpie.cpp:4063-4432).
   uniforms: [max_steps][8] or NULL. codes_out: [max_steps][8] frame-major. Returns n_frames.
   hidden_out (optional): [max_steps][d] decoder hidden fed to the LT at each step. */
ORC_API int orc_synthesize(const orc_model * m, const int32_t * tokens, int E, int speaker,
                           float temperature, int top_k, int max_steps, const float * uniforms,
                           int32_t * codes_out, float * hidden_out) {
    const orc_hparams * hp = &m->hp; const int d = hp->d_model;
    if (max_steps <= 0) max_steps = hp->max_dec_steps;
    float * enc = (float *)malloc((size_t)E * d * sizeof(float));
    if (!orc_encode_text(m, tokens, E, enc)) { free(enc); return -1; }
    orc_state * s = orc_state_new(m, enc, E, speaker, hp->context_frames + max_steps + 16);
    float * hidden = (float *)malloc((size_t)d * sizeof(float));
    int32_t bos[NUM_CB]; for (int i = 0; i < NUM_CB; i++) bos[i] = hp->audio_bos_id;
    orc_decoder_step(s, bos, hidden);
    int n_frames = 0;
    for (int step = 0; step < max_steps; step++) {
        int32_t smp[NUM_CB], am[NUM_CB];
        if (hidden_out) memcpy(hidden_out + (size_t)step * d, hidden, (size_t)d * 4);
        orc_lt_sample(m, hidden, temperature, top_k, step < 4, NULL,
                      uniforms ? uniforms + (size_t)step * NUM_CB : NULL, smp, am, NULL);
        int eos = 0;
        for (int cb = 0; cb < NUM_CB; cb++)
            if (smp[cb] == hp->audio_eos_id || am[cb] == hp->audio_eos_id) eos = 1;
        if (eos) break;                         /* EOS frame is not emitted */
        memcpy(codes_out + (size_t)n_frames * NUM_CB, smp, sizeof(smp));
        n_frames++;
        if (step + 1 >= max_steps) break;
        orc_decoder_step(s, smp, hidden);
    }
    free(hidden); free(enc); orc_state_free(s);
    return n_frames;
}

/* ------------------------------------------------------------------------------------------ */
/* nano-codec (nano-codec.cpp:376-845)                                                         */
/* ------------------------------------------------------------------------------------------ */

typedef struct { mat_t in_alpha, in_w, in_b, sk_alpha, sk_w, sk_b; } resblock_t;
typedef struct {
    mat_t pre_w, pre_b, post_alpha, post_w, post_b;
    mat_t act_alpha[5], up_w[5], up_b[5];
    resblock_t rb[5][3][3];
    int up_rates[5];
    int conv_f16;    /* 1 = emulate ggml_conv_1d's f16 im2col/kernel rounding (default) */
    int hop;
} orc_codec;

ORC_API orc_codec * orc_codec_new(void) {
    orc_codec * c = (orc_codec *)calloc(1, sizeof(orc_codec));
    const int r[5] = {8, 8, 4, 2, 2};      /* magpie.h:672 */
    memcpy(c->up_rates, r, sizeof(r));
    c->conv_f16 = 1; c->hop = 1024;
    return c;
}
ORC_API void orc_codec_set_conv_f16(orc_codec * c, int on) { c->conv_f16 = on; }

ORC_API void orc_codec_free(orc_codec * c) {
    if (!c) return;
    mat_t * p = (mat_t *)&c->pre_w;
    size_t n = ((char *)&c->up_rates - (char *)&c->pre_w) / sizeof(mat_t);
    for (size_t i = 0; i < n; i++) mat_free(&p[i]);
    free(c);
}

/* name mapping follows map_codec_tensor (nano-codec.cpp:84-199) */
ORC_API int orc_codec_set_tensor(orc_codec * c, const char * name, const void * data, int type,
                                 int n_dims, const int64_t * ne) {
    mat_t * slot = NULL;
    if (strstr(name, "dec.pre.weight")) slot = &c->pre_w;
    else if (strstr(name, "dec.pre.bias")) slot = &c->pre_b;
    else if (strstr(name, "dec.post.weight")) slot = &c->post_w;
    else if (strstr(name, "dec.post.bias")) slot = &c->post_b;
    else if (strstr(name, "dec.post_act.alpha")) slot = &c->post_alpha;
    else if (strstr(name, "dec.up.")) {
        int i = parse_idx(name, "dec.up.");
        if (i >= 0 && i < 5) { if (strstr(name, ".weight")) slot = &c->up_w[i]; else if (strstr(name, ".bias")) slot = &c->up_b[i]; }
    } else if (strstr(name, "dec.act.") && strstr(name, "alpha")) {
        int i = parse_idx(name, "dec.act.");
        if (i >= 0 && i < 5) slot = &c->act_alpha[i];
    } else if (strstr(name, "dec.rl.")) {
        const char * p = strstr(name, "dec.rl.") + 7; int i = atoi(p);
        p = strstr(p, ".rb."); if (!p) return 0; p += 4; int j = atoi(p);
        p = strstr(p, ".rb."); if (!p) return 0; p += 4; int k = atoi(p);
        if (i < 0 || i >= 5 || j < 0 || j >= 3 || k < 0 || k >= 3) return 0;
        resblock_t * b = &c->rb[i][j][k];
        if (strstr(name, ".in_act.alpha")) slot = &b->in_alpha;
        else if (strstr(name, ".in_conv.weight")) slot = &b->in_w;
        else if (strstr(name, ".in_conv.bias")) slot = &b->in_b;
        else if (strstr(name, ".sk_act.alpha")) slot = &b->sk_alpha;
        else if (strstr(name, ".sk_conv.weight")) slot = &b->sk_w;
        else if (strstr(name, ".sk_conv.bias")) slot = &b->sk_b;
    }
    if (!slot) return 0;
    mat_fill(slot, data, type, n_dims, ne);
    return 1;
}

/* fsq_dequantize_cpu (nano-codec.cpp:721-752): codes [8][T] -> latent [32][T] (T fastest) */
ORC_API void orc_fsq_dequantize(const int32_t * codes, int num_cb, int T, float * latent) {
    static const int base[4] = {1, 8, 56, 336};
    static const int levels[4] = {8, 7, 6, 6};
    for (int cb = 0; cb < num_cb; cb++)
        for (int t = 0; t < T; t++) {
            int index = codes[(size_t)cb * T + t];
            for (int dd = 0; dd < 4; dd++) {
                int nonneg = (index / base[dd]) % levels[dd];
                int half = levels[dd] / 2;
                latent[(size_t)(cb * 4 + dd) * T + t] = (float)(nonneg - half) / (float)half;
            }
        }
}

/* HalfSnake (nano-codec.cpp:376-426): first numel(alpha) channels x + sin^2(ax)/a, rest LeakyReLU(0.01) */
static void half_snake(const float * x, float * y, int C, int T, const mat_t * alpha) {
    int ns = (int)nelem(alpha->ne, alpha->n_dims);
#pragma omp parallel for schedule(static)
    for (int c = 0; c < C; c++) {
        const float * xi = x + (size_t)c * T; float * yo = y + (size_t)c * T;
        if (c < ns) {
            float a = alpha->w[c];
            for (int t = 0; t < T; t++) { float ax = xi[t] * a; float sn = sinf(ax); yo[t] = xi[t] + (sn * sn) / a; }
        } else {
            for (int t = 0; t < T; t++) yo[t] = xi[t] > 0.0f ? xi[t] : 0.01f * xi[t];
        }
    }
}

/* causal conv1d (nano-codec.cpp:429-466): y[o][t] = b[o] + sum_{i,k} w[o][i][k] x[i][t-(K-1-k)*dil];
   ggml_conv_1d rounds im2col(x) and the kernel to f16, accumulates f32 [ggml-upstream]. */
static void causal_conv1d(const orc_codec * c, const float * x, float * y, int Cin, int Cout, int T,
                          const mat_t * W, const mat_t * B, int dil) {
    int K = (int)W->ne[0];
    int pad = (K - 1) * dil;
    int f16 = c->conv_f16 || W->act == ACT_F16;
    float * xp = (float *)calloc((size_t)Cin * (T + pad), sizeof(float));
    for (int i = 0; i < Cin; i++)
        for (int t = 0; t < T; t++) {
            float v = x[(size_t)i * T + t];
            xp[(size_t)i * (T + pad) + pad + t] = f16 ? f16_round(v) : v;
        }
#pragma omp parallel for schedule(static)
    for (int o = 0; o < Cout; o++) {
        float * acc = y + (size_t)o * T;
        for (int t = 0; t < T; t++) acc[t] = 0.0f;
        for (int i = 0; i < Cin; i++) {
            const float * xi = xp + (size_t)i * (T + pad);
            for (int k = 0; k < K; k++) {
                float w = W->w[((size_t)o * Cin + i) * K + k];
                if (f16) w = f16_round(w);
                const float * xs = xi + k * dil;
                for (int t = 0; t < T; t++) acc[t] += w * xs[t];
            }
        }
        if (B && B->w) { float b = B->w[o]; for (int t = 0; t < T; t++) acc[t] += b; }
    }
    free(xp);
}

/* grouped ConvTranspose1d (nano-codec.cpp:481-565): groups = Cout, 2 inputs per group, K = 2*stride,
   keep first T*stride samples, + bias; f32 throughout. */
static void conv_transpose1d(const float * x, float * y, int Cin, int T, const mat_t * W, const mat_t * B, int stride) {
    int K = (int)W->ne[0], Cout = Cin / 2, To = T * stride;
#pragma omp parallel for schedule(static)
    for (int g = 0; g < Cout; g++) {
        float * yo = y + (size_t)g * To;
        const float * x0 = x + (size_t)(2 * g) * T, * x1 = x + (size_t)(2 * g + 1) * T;
        const float * w0 = W->w + (size_t)(2 * g) * K, * w1 = W->w + (size_t)(2 * g + 1) * K;
        for (int n = 0; n < To; n++) yo[n] = 0.0f;
        for (int t = 0; t < T; t++)
            for (int k = 0; k < K; k++) {
                int n = t * stride + k;
                if (n < To) yo[n] += w0[k] * x0[t] + w1[k] * x1[t];
            }
        if (B && B->w) { float b = B->w[g]; for (int n = 0; n < To; n++) yo[n] += b; }
    }
}

/* magpie_codec_build_decoder (nano-codec.cpp:676-715) on an FSQ latent [32][T] */
static void codec_decoder(const orc_codec * c, const float * latent, int T, float * pcm) {
    int C = (int)c->pre_w.ne[2], Cin0 = (int)c->pre_w.ne[1];
    float * cur = (float *)malloc((size_t)C * T * sizeof(float));
    causal_conv1d(c, latent, cur, Cin0, C, T, &c->pre_w, &c->pre_b, 1);
    int curT = T;
    static const int dils[3] = {1, 3, 5};
    for (int i = 0; i < 5; i++) {
        int s = c->up_rates[i], Co = C / 2, To = curT * s;
        float * act = (float *)malloc((size_t)C * curT * sizeof(float));
        half_snake(cur, act, C, curT, &c->act_alpha[i]);
        float * up = (float *)malloc((size_t)Co * To * sizeof(float));
        conv_transpose1d(act, up, C, curT, &c->up_w[i], &c->up_b[i], s);
        free(act); free(cur);
        size_t nn = (size_t)Co * To;
        float * sum = (float *)malloc(nn * sizeof(float));
        float * o = (float *)malloc(nn * sizeof(float));
        float * h = (float *)malloc(nn * sizeof(float));
        float * h2 = (float *)malloc(nn * sizeof(float));
        for (int j = 0; j < 3; j++) {
            memcpy(o, up, nn * sizeof(float));
            for (int k = 0; k < 3; k++) {      /* residual block (nano-codec.cpp:568-599) */
                const resblock_t * b = &c->rb[i][j][k];
                half_snake(o, h, Co, To, &b->in_alpha);
                causal_conv1d(c, h, h2, Co, Co, To, &b->in_w, &b->in_b, dils[k]);
                half_snake(h2, h, Co, To, &b->sk_alpha);
                causal_conv1d(c, h, h2, Co, Co, To, &b->sk_w, &b->sk_b, 1);
                for (size_t q = 0; q < nn; q++) o[q] = o[q] + h2[q];
            }
            if (j == 0) memcpy(sum, o, nn * sizeof(float));
            else for (size_t q = 0; q < nn; q++) sum[q] = sum[q] + o[q];
        }
        for (size_t q = 0; q < nn; q++) sum[q] = sum[q] * (1.0f / 3.0f);
        free(o); free(h); free(h2); free(up);
        cur = sum; C = Co; curT = To;
    }
    float * act = (float *)malloc((size_t)C * curT * sizeof(float));
    half_snake(cur, act, C, curT, &c->post_alpha);
    causal_conv1d(c, act, pcm, C, 1, curT, &c->post_w, &c->post_b, 1);
    for (int t = 0; t < curT; t++) pcm[t] = tanhf(pcm[t]);
    free(act); free(cur);
}

/* magpie_codec_decode (nano-codec.cpp:758-845): codes [8][T] codebook-major -> pcm [T*hop] */
ORC_API int orc_codec_decode(const orc_codec * c, const int32_t * codes, int T, float * pcm) {
    if (T <= 0) return 0;
    float * latent = (float *)malloc((size_t)32 * T * sizeof(float));
    orc_fsq_dequantize(codes, 8, T, latent);
    codec_decoder(c, latent, T, pcm);
    free(latent);
    return T * c->hop;
}

/* building blocks exported for layer-level cross-checks in tests/ */
ORC_API void orc_codec_half_snake(const float * x, float * y, int C, int T, const float * alpha, int n_alpha) {
    mat_t a = { (float *)alpha, {n_alpha, 1, 1, 1}, 1, ACT_F32 };
    half_snake(x, y, C, T, &a);
}
ORC_API void orc_codec_causal_conv1d(const float * x, float * y, int Cin, int Cout, int T, int K,
                                     const float * w, const float * b, int dil, int f16) {
    orc_codec c; memset(&c, 0, sizeof(c)); c.conv_f16 = f16;
    mat_t W = { (float *)w, {K, Cin, Cout, 1}, 3, ACT_F32 };
    mat_t B = { (float *)b, {Cout, 1, 1, 1}, 1, ACT_F32 };
    causal_conv1d(&c, x, y, Cin, Cout, T, &W, b ? &B : NULL, dil);
}
ORC_API void orc_codec_conv_transpose1d(const float * x, float * y, int Cin, int T, int K,
                                        const float * w, const float * b, int stride) {
    mat_t W = { (float *)w, {K, 1, Cin, 1}, 3, ACT_F32 };
    mat_t B = { (float *)b, {Cin / 2, 1, 1, 1}, 1, ACT_F32 };
    conv_transpose1d(x, y, Cin, T, &W, b ? &B : NULL, stride);
}

ORC_API void orc_layer_norm(const float * x, const float * w, float * y, int n, float eps) { layer_norm(x, w, y, n, eps); }
ORC_API float orc_gelu(float x, int table) { gelu_table_init(); return gelu_ggml(x, table); }
ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
ORC_API void orc_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
