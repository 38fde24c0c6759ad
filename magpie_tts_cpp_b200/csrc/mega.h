// Parameters of the batch-1 decoder-step megakernel (decoder_mega.cu).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace mgb {

constexpr int kMegaMaxLayers = 16;

struct MegaLayer {
    const void * qkv, * o, * xq, * xo, * ff1, * ff2;     // row-major [N][K], model weight dtype
    const float * n_self, * n_xq, * n_ff;
};

struct MegaParams {
    MegaLayer layer[kMegaMaxLayers];
    int L, d, f, dxa, H, n_split;
    float eps; int gelu_f16;
    const float * audio_emb[8]; const float * dec_pos; const float * norm_out;
    const int32_t * codes;                 // [8]
    const int32_t * pos; int32_t * pos_rw; int32_t * slot_rw;
    void * kcache; void * vcache; size_t kv_layer_stride;       // elements per layer
    const void * xk; const void * xv; size_t xkv_layer_stride;
    const int32_t * n_ctx;
    float * x, * q, * attn_part, * xq, * ffh, * hidden;
    unsigned * barrier;
    unsigned long long * dbg;              // optional: clock64() stamps of CTA 0 (profiling aid), 16 per layer
};

size_t mega_smem_bytes();
int    mega_max_grid(int precision);      // #SMs if the kernel fits one CTA per SM, else 0
bool   launch_decoder_mega(const MegaParams & p, int precision, int grid, cudaStream_t stream);

}  // namespace mgb
