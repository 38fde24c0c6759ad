// Tensor-core (tcgen05 / TMEM) implicit-GEMM causal conv of the nano-codec residual blocks (reference
// src/nano-codec.cpp:429-466, 568-641).  See codec_conv_tc.cu for the design.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace mgb {
namespace ctc {

constexpr int kHP = 56;            // zero rows in front of every time-major activation image: causal history of the largest conv ((11-1)*5 = 50 -> 56)
constexpr int kTile = 128;         // time steps per accumulator tile (MMA M)

// How a C -> C conv is mapped: output channels in `nsplit` groups of `nper` (padded to npad MMA columns), input
// channels in `nchunk` chunks of 64 (one 128-byte swizzle row each).
struct Geom {
    int C = 0, nsplit = 1, nper = 0, npad = 0, nchunk = 0;
    int rb = 128;           // bytes per image row: 128 (64 channels per chunk) or 64 (32 channels, stages with <= 32 channels)
    bool ok = false;
};
Geom geom_for(int C);

// Time-major f16 activation image: [B][nchunk][kHP + Tpad][64 ch] with the 16-byte channel groups of row r XOR-swizzled
// by (r & 7) (SWIZZLE_128B), so that any 8-row-aligned window is a ready-made shared-memory operand image.
// f32 tensors of the tensor-core pipeline are TIME-MAJOR rows [B][T][row_stride(C)] (padding channels stay zero)
inline int row_stride(int C) { return (C + 15) / 16 * 16; }
inline size_t act_rows(int T) { return (size_t)kHP + (size_t)(T + 4 * kTile - 1) / (4 * kTile) * (4 * kTile); }   // a work item spans up to 4 tiles
inline size_t act_bytes(int B, const Geom & g, int T) { return (size_t)B * g.nchunk * act_rows(T) * g.rb; }

size_t weight_image_bytes(const Geom & g, int K);
// f32 (Cout = C, Cin = C, K) PyTorch layout on the device -> f16 tile images [nsplit][nchunk][K][npad x 64]
bool pack_weights(const float * w, const Geom & g, int K, void * img, cudaStream_t stream);

struct ConvArgs {
    const __half * xa = nullptr;     // activated input image (time-major)
    const __half * w = nullptr;      // weight tile images
    const float * bias = nullptr;    // [C]
    const float * res = nullptr;     // optional residual, f32 rows [B][T][row_stride(C)]
    float * y = nullptr;             // optional raw output conv + bias (+ res), f32 rows
    __half * ya = nullptr;           // optional activated output image f16(half_snake(y; alpha2))
    const float * alpha2 = nullptr; int n_alpha2 = 0;
    int B = 0, T = 0, K = 0, dil = 1;
};
bool launch_conv(const Geom & g, const ConvArgs & a, cudaStream_t stream);

// HalfSnake -> grouped transposed conv (stride s) of the previous stage's output, writing the stage input `up` (f32 rows)
// and the three residual branches' first activated images.
struct UpArgs {
    const float * x[3] = {}; int n_x = 1;   // [B][T][row_stride(Cin)]; n_x == 3: the input is the mean of three branch outputs
    const float * alpha = nullptr; int n_alpha = 0;     // HalfSnake in front of the transposed conv
    const float * w = nullptr;          // [Cin][1][2s]
    const float * bias = nullptr;       // [Cin/2]
    float * up = nullptr;               // [B][T*s][row_stride(Cin/2)]
    __half * img[3] = {};
    const float * br_alpha[3] = {}; int n_br_alpha = 0;
    int B = 0, Cin = 0, T = 0, s = 1;
};
bool launch_up(const Geom & g_out, const UpArgs & a, cudaStream_t stream);

// HalfSnake -> conv (C -> 1) -> tanh on f32 rows -> pcm [B][T]
struct PostArgs {
    const float * x[3] = {};            // the three branch outputs of the last stage (their mean is the input)
    const float * alpha = nullptr; int n_alpha = 0;
    const float * w = nullptr; const float * bias = nullptr; float * pcm = nullptr;
    int B = 0, C = 0, K = 0, T = 0;
};
bool launch_post(const PostArgs & a, cudaStream_t stream);

}  // namespace ctc
}  // namespace mgb
