#include "gguf_reader.h"

#include <cstring>
#include <exception>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace mgb {

namespace {

enum : int32_t {
    T_U8 = 0, T_I8, T_U16, T_I16, T_U32, T_I32, T_F32, T_BOOL, T_STR, T_ARR, T_U64, T_I64, T_F64
};

struct Cursor {
    const uint8_t * p; const uint8_t * end; bool ok = true;
    template <typename T> T rd() {
        T v{};
        if (!ok || (size_t)(end - p) < sizeof(T)) { ok = false; return v; }
        memcpy(&v, p, sizeof(T)); p += sizeof(T); return v;
    }
    std::string str() {
        uint64_t n = rd<uint64_t>();
        if (!ok || n > (uint64_t)(end - p)) { ok = false; return {}; }      // remaining-size comparison: p + n cannot wrap
        std::string s((const char *)p, (size_t)n); p += n; return s;
    }
};

size_t scalar_size(int32_t t) {
    switch (t) {
        case T_U8: case T_I8: case T_BOOL: return 1;
        case T_U16: case T_I16: return 2;
        case T_U32: case T_I32: case T_F32: return 4;
        case T_U64: case T_I64: case T_F64: return 8;
        default: return 0;
    }
}

bool read_value(Cursor & c, int32_t type, GgufValue & v) {
    v.type = type;
    switch (type) {
        case T_U8:   v.u = c.rd<uint8_t>(); break;
        case T_I8:   v.u = (uint64_t)(int64_t)c.rd<int8_t>(); break;
        case T_U16:  v.u = c.rd<uint16_t>(); break;
        case T_I16:  v.u = (uint64_t)(int64_t)c.rd<int16_t>(); break;
        case T_U32:  v.u = c.rd<uint32_t>(); break;
        case T_I32:  v.u = (uint64_t)(int64_t)c.rd<int32_t>(); break;
        case T_U64:  v.u = c.rd<uint64_t>(); break;
        case T_I64:  v.u = (uint64_t)c.rd<int64_t>(); break;
        case T_BOOL: v.u = c.rd<uint8_t>(); break;
        case T_F32:  v.f = c.rd<float>(); break;
        case T_F64:  v.f = c.rd<double>(); break;
        case T_STR:  v.s = c.str(); break;
        case T_ARR: {   // arrays are skipped (the reference reads none)
            int32_t et = c.rd<int32_t>();
            uint64_t n = c.rd<uint64_t>();
            if (et == T_STR) { for (uint64_t i = 0; i < n && c.ok; i++) c.str(); }
            else {
                size_t es = scalar_size(et);
                if (es == 0 || !c.ok || n > (uint64_t)(c.end - c.p) / es) { c.ok = false; break; }
                c.p += es * n;
            }
            break;
        }
        default: c.ok = false;
    }
    return c.ok;
}

float f16_to_f32(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000) << 16, exp = (h >> 10) & 0x1f, man = h & 0x3ff, bits;
    if (exp == 0) {
        if (man == 0) bits = sign;
        else {
            int e = -1;
            do { e++; man <<= 1; } while (!(man & 0x400));
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ff) << 13);
        }
    } else if (exp == 31) bits = sign | 0x7f800000u | (man << 13);
    else bits = sign | ((exp + 112) << 23) | (man << 13);
    float f; memcpy(&f, &bits, 4); return f;
}

}  // namespace

size_t GgufTensor::nbytes() const {
    int64_t n = nelements();
    if (n < 0) return 0;
    switch (type) {
        case GGML_F32: return (size_t)n * 4;
        case GGML_F16: return (size_t)n * 2;
        case GGML_Q8_0: return (size_t)(n / 32) * 34;
        case GGML_Q4_0: return (size_t)(n / 32) * 18;
        default: return 0;
    }
}

GgufFile::~GgufFile() { if (map_) munmap(map_, map_size_); }

bool GgufFile::open(const char * path, std::string & err) {
    int fd = ::open(path, O_RDONLY);
    if (fd < 0) { err = std::string("failed to open '") + path + "'"; return false; }
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size < 24) { ::close(fd); err = "file too small"; return false; }
    map_size_ = (size_t)st.st_size;
    map_ = mmap(nullptr, map_size_, PROT_READ, MAP_PRIVATE, fd, 0);
    ::close(fd);
    if (map_ == MAP_FAILED) { map_ = nullptr; err = "mmap failed"; return false; }
    Cursor c{(const uint8_t *)map_, (const uint8_t *)map_ + map_size_};
    if (memcmp(c.p, "GGUF", 4) != 0) { err = "bad magic (not a GGUF file)"; return false; }
    c.p += 4;
    uint32_t version = c.rd<uint32_t>();
    if (version < 2 || version > 3) { err = "unsupported GGUF version"; return false; }
    uint64_t n_tensors = c.rd<uint64_t>(), n_kv = c.rd<uint64_t>();
    // every KV entry takes >= 12 bytes (key length + type), every tensor entry >= 24: counts beyond that are corrupt
    if (!c.ok || n_kv > map_size_ / 12 || n_tensors > map_size_ / 24) { err = "corrupt GGUF header (entry counts exceed the file size)"; return false; }
    uint64_t alignment = 32;
    for (uint64_t i = 0; i < n_kv && c.ok; i++) {
        std::string key = c.str();
        int32_t type = c.rd<int32_t>();
        GgufValue v;
        if (!read_value(c, type, v)) break;
        if (key == "general.alignment" && v.type == T_U32) alignment = v.u;
        kv_[key] = std::move(v);
    }
    if (!c.ok) { err = "truncated GGUF metadata"; return false; }
    if (alignment == 0 || (alignment & (alignment - 1)) != 0 || alignment > (1u << 20)) { err = "general.alignment must be a power of two"; return false; }
    try { tensors_.resize((size_t)n_tensors); }
    catch (const std::exception &) { err = "out of memory reading the tensor table"; return false; }
    for (auto & t : tensors_) {
        t.name = c.str();
        t.n_dims = (int)c.rd<uint32_t>();
        if (!c.ok || t.n_dims < 0 || t.n_dims > 4) { err = "bad tensor info"; return false; }
        int64_t total = 1;
        for (int d = 0; d < t.n_dims; d++) {
            const uint64_t ne = c.rd<uint64_t>();
            // dimensions must be positive and the element count must stay far below 2^63 (no overflow in nelements / nbytes)
            if (!c.ok || ne == 0 || ne > (uint64_t)1 << 40 || (uint64_t)total > ((uint64_t)1 << 44) / ne) { err = "tensor '" + t.name + "': bad dimensions"; return false; }
            t.ne[d] = (int64_t)ne;
            total *= (int64_t)ne;
        }
        t.type = c.rd<int32_t>();
        t.offset = c.rd<uint64_t>();
        if ((t.type == GGML_Q8_0 || t.type == GGML_Q4_0) && t.ne[0] % 32 != 0) { err = "tensor '" + t.name + "': quantised row length is not a multiple of 32"; return false; }
    }
    if (!c.ok) { err = "truncated GGUF tensor table"; return false; }
    const size_t pos = (size_t)(c.p - (const uint8_t *)map_);
    const size_t data_off = (pos + alignment - 1) / alignment * alignment;
    if (data_off > map_size_) { err = "GGUF data section starts past the end of the file"; return false; }
    const size_t data_size = map_size_ - data_off;
    for (size_t i = 0; i < tensors_.size(); i++) {
        auto & t = tensors_[i];
        const size_t nb = t.nbytes();
        if (nb == 0 && t.nelements() != 0) { err = "tensor '" + t.name + "': unsupported type"; return false; }
        if (t.offset > data_size || nb > data_size - t.offset) { err = "tensor '" + t.name + "' runs past end of file"; return false; }
        t.data = (const uint8_t *)map_ + data_off + t.offset;
        index_[t.name] = i;
    }
    return true;
}

const GgufValue * GgufFile::find(const std::string & key) const {
    auto it = kv_.find(key);
    return it == kv_.end() ? nullptr : &it->second;
}

int32_t GgufFile::get_u32(const std::string & key, int32_t def) const {
    const GgufValue * v = find(key);
    if (!v || v->type == T_STR || v->type == T_ARR || v->type == T_F32 || v->type == T_F64) return def;
    return (int32_t)v->u;
}

float GgufFile::get_f32(const std::string & key, float def) const {
    const GgufValue * v = find(key);
    if (!v || (v->type != T_F32 && v->type != T_F64)) return def;
    return (float)v->f;
}

const std::string * GgufFile::get_str(const std::string & key) const {
    const GgufValue * v = find(key);
    return (v && v->type == T_STR) ? &v->s : nullptr;
}

const GgufTensor * GgufFile::tensor(const std::string & name) const {
    auto it = index_.find(name);
    return it == index_.end() ? nullptr : &tensors_[it->second];
}

bool gguf_to_f32(const GgufTensor & t, float * dst) {
    int64_t n = t.nelements();
    if (t.type == GGML_F32) { memcpy(dst, t.data, (size_t)n * 4); return true; }
    if (t.type == GGML_F16) {
        const uint16_t * h = (const uint16_t *)t.data;
        for (int64_t i = 0; i < n; i++) dst[i] = f16_to_f32(h[i]);
        return true;
    }
    if (t.type == GGML_Q8_0) {   // block: f16 scale + 32 x int8 along ne[0]
        const uint8_t * p = t.data;
        for (int64_t b = 0; b < n / 32; b++, p += 34) {
            uint16_t hs; memcpy(&hs, p, 2);
            float d = f16_to_f32(hs);
            const int8_t * q = (const int8_t *)(p + 2);
            for (int j = 0; j < 32; j++) dst[b * 32 + j] = d * (float)q[j];
        }
        return true;
    }
    return false;
}

}  // namespace mgb
