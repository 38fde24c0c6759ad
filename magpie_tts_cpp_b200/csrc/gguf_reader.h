// Minimal GGUF v2/v3 reader (host only).  Replaces the reference's use of ggml's gguf_* API
// (reference src/magpie.cpp:794-810, 353-398, 674-718) -- same file format, no ggml.
#pragma once
#include <cstdint>
#include <cstdio>
#include <map>
#include <string>
#include <vector>

namespace mgb {

enum GgmlType : int32_t { GGML_F32 = 0, GGML_F16 = 1, GGML_Q4_0 = 2, GGML_Q8_0 = 8 };

struct GgufTensor {
    std::string name;
    int         n_dims = 0;
    int64_t     ne[4]  = {1, 1, 1, 1};   // ggml order: ne[0] fastest (reversed PyTorch shape)
    int32_t     type   = 0;
    uint64_t    offset = 0;               // relative to data section
    const uint8_t * data = nullptr;       // into the mapped file
    int64_t nelements() const { return ne[0] * ne[1] * ne[2] * ne[3]; }
    size_t  nbytes() const;
};

struct GgufValue {
    int32_t     type = -1;    // gguf value type id
    uint64_t    u = 0;        // integer payloads
    double      f = 0.0;      // float payloads
    std::string s;            // string payload
};

class GgufFile {
public:
    ~GgufFile();
    // Returns false and fills err on failure.
    bool open(const char * path, std::string & err);
    const GgufValue * find(const std::string & key) const;
    // reference semantics: gguf_get_val_u32 on a present key, else default (magpie.cpp:74-77)
    int32_t get_u32(const std::string & key, int32_t def) const;
    float   get_f32(const std::string & key, float def) const;
    const std::string * get_str(const std::string & key) const;
    const std::vector<GgufTensor> & tensors() const { return tensors_; }
    const GgufTensor * tensor(const std::string & name) const;

private:
    std::map<std::string, GgufValue> kv_;
    std::vector<GgufTensor> tensors_;
    std::map<std::string, size_t> index_;
    void * map_ = nullptr;
    size_t map_size_ = 0;
};

// Dequantise a tensor payload (F32 / F16 / Q8_0) to float32. Returns false for unsupported types.
bool gguf_to_f32(const GgufTensor & t, float * dst);

}  // namespace mgb
