// Parameters of the batch-1 persistent frame-loop kernel (frame_loop.cu): decoder step + local transformer +
// sampling + EOS bookkeeping for ALL frames of a generation / teacher-forced run in one cooperative launch.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace mgb {

constexpr int kLoopMaxLayers = 16;
constexpr int kLoopMaxCtx = 512;         // text tokens the folded cross-attention path handles (the first 30 from registers, the rest streamed)
constexpr int kLoopDbgPerCta = 256;      // globaltimer stamps per CTA (last frame of a launch)
constexpr int kLoopDbgStamps = 160 * kLoopDbgPerCta;

// exchange buffers (16-byte flag packets, see frame_loop.cu)
enum LoopXchg { X_QKV = 0, X_ATT, X_XA, X_XB, X_H, X_XC, T_SEQ0, T_QKV, T_X1, T_H, T_HOUT, T_AMAX, T_LOGITS, X_COUNT };

struct LoopLayer {
    const void * qkv, * o, * ff1, * ff2;     // bf16 row-major [N][K]
    const float * n_self, * n_xq, * n_ff;
    const float * xm, * xn;                  // folded cross-attention, [E][d] f32 each (xattn_fold)
};

struct FrameLoopParams {
    LoopLayer layer[kLoopMaxLayers];
    int L, E;
    float eps; int gelu_f16;
    const float * audio_emb[8]; const float * dec_pos; const float * norm_out;
    void * kcache; void * vcache; size_t kv_layer_stride;      // bf16 [L][max_seq][d]
    // local transformer
    int V;
    const void * lt_in_w; const float * lt_in_b; const float * lt_pos; const float * lt_norm_self; const float * lt_norm_ff;
    const void * lt_qkvo;            // [4*LD][LD] bf16: [Wq; Wk; hi(Wo Wv); lo(Wo Wv)] (model.cu)
    const void * lt_ff1, * lt_ff2; const void * lt_out_w[8]; const float * lt_out_b[8];
    const float * lt_in_table[8];
    const float * lt_qkv_tab;        // [7][V][3*LD] f32: [q | k | vo] of LT position cb+1 for every fed code of codebook cb
    // loop control
    int n_steps, pos0, step0, row0, min_frames, teacher, ignore_eos;
    float temperature; int top_k; unsigned long long seed;
    const float * uniforms;              // [rows][8] or null
    const int32_t * forced;              // [rows][8] or null (teacher forcing)
    const int32_t * codes_io;            // [8] codes the first step consumes
    int bos_id, eos_id;
    int32_t * sampled, * argmax;         // [rows][8]
    float * logits;                      // [rows][8][V] or null
    float * hidden_hist;                 // [rows][d] or null
    float * hidden_last;                 // [d]
    int32_t * result;                    // [0] frames run, [1] first step (relative to this launch) at which EOS was hit, or -1,
                                         // [2..9] codes the next step consumes
    uint4 * xbuf; int xoff[X_COUNT + 1]; // packet offsets inside one replica; xoff[X_COUNT] = replica stride
    unsigned * seq;                      // persistent exchange sequence number
    unsigned long long * dbg;            // optional: globaltimer stamps of CTA 0 for the last frame
    int dbg_flags;                       // profiling aids, see frame_loop.cu (0 in production)
    int max_split;                       // key splits per head, 1..6 (6 in production; smaller values exercise the multi-round scan in tests)
    int no_defer_amax;                   // 1 (MGB_LOOP_NO_DEFER): teacher forcing collects every codebook's argmax at once, as free-running generation must (A/B)
};

bool   frame_loop_shape_ok(int d, int f, int H, int ld, int lf, int V, int L);
int    frame_loop_max_grid();            // #SMs if the kernel fits one CTA per SM (and #SMs >= 147), else 0
void   frame_loop_xchg_layout(int V, int * xoff /*[X_COUNT+1]*/);
size_t frame_loop_xchg_bytes(const int * xoff);
bool   launch_frame_loop(const FrameLoopParams & p, int grid, cudaStream_t stream);

// M_l[j][i] = scale * sum_c K_l[j][c] Wq_l[c][i],  N_l[j][n] = sum_c V_l[j][c] Wo_l[n][c]   (bf16 inputs, f32 out)
bool   launch_xattn_fold(const void * xk, const void * xv, const void * wq, const void * wo, int E, int d, int dxa,
                         float scale, float * xm, float * xn, cudaStream_t stream, const int32_t * n_ctx = nullptr, int rows_per_utt = 1);

}  // namespace mgb
