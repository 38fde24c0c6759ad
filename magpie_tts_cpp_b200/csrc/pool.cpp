// In-process multi-GPU synthesis (include/magpie_b200.h, mgb_pool_*): one model replica, one session and one submission
// thread (hence one CUDA stream) per device; utterance i runs on device i mod G (mgb_shard_device).  Utterances are
// independent from tokens to codes (the reference is single-utterance, src/magpie.cpp:4063-4432), so there is
// NO collective and no inter-device traffic: every device runs its share as one batched session (encode -> prefill -> loop)
// and the results are scattered back to the caller's arrays in utterance order.  NCCL is not linked.
#include <algorithm>
#include <cstring>
#include <exception>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/magpie_b200.h"

namespace mgb {
void set_error(const std::string & msg);
}

struct mgb_pool {
    std::vector<mgb_model *> models;
    std::vector<int> devices;
    std::vector<float> last_ms;
    // one cached session per device, reused while the share's shape (utterances, text capacity, cache length) stays the same:
    // creating a session allocates the K/V pool and the loop buffers (tens of ms), which a serving loop must not pay per call
    struct Cached { mgb_session * s = nullptr; int nb = 0, max_text = 0, max_seq = 0; };
    std::vector<Cached> sessions;
    ~mgb_pool() {
        for (Cached & c : sessions) if (c.s) mgb_session_free(c.s);
        for (mgb_model * m : models) if (m) mgb_model_free(m);
    }
};

namespace {

struct Job {
    int n_utt = 0, max_text = 0, T = 0;
    const int32_t * tokens = nullptr, * n_tokens = nullptr, * speakers = nullptr;
    // generation
    float temperature = 0.0f; int top_k = 80; uint64_t seed = 0; int ignore_eos = 0;
    int32_t * codes_out = nullptr, * n_frames_out = nullptr;
    // teacher forcing
    const int32_t * codes_in = nullptr; int32_t * greedy_out = nullptr;
    bool teacher = false;
};

// one device's share: utterances di, di + G, di + 2G, ...
std::string run_share(mgb_pool * p, int di, const Job & j, float * ms_out) {
    const int G = (int)p->models.size();
    std::vector<int> idx;
    for (int i = 0; i < j.n_utt; i++) if (mgb_shard_device(i, G) == di) idx.push_back(i);
    *ms_out = 0.0f;
    if (idx.empty()) return "";
    const int nb = (int)idx.size();
    mgb_hparams hp;
    mgb_model_get_hparams(p->models[di], &hp);
    int local_text = 1;
    for (int i : idx) local_text = std::max(local_text, j.n_tokens[i]);
    mgb_pool::Cached & cs = p->sessions[di];
    const int max_seq = hp.context_frames + j.T + 16;
    if (!cs.s || cs.nb != nb || cs.max_text != local_text || cs.max_seq != max_seq) {
        if (cs.s) mgb_session_free(cs.s);
        cs.s = mgb_session_new(p->models[di], nb, local_text, max_seq);
        cs.nb = nb; cs.max_text = local_text; cs.max_seq = max_seq;
    }
    mgb_session * s = cs.s;
    if (!s) return mgb_last_error();
    std::string err;
    std::vector<int32_t> tok((size_t)nb * local_text, 0), nt(nb), spk(nb);
    for (int b = 0; b < nb; b++) {
        nt[b] = j.n_tokens[idx[b]]; spk[b] = j.speakers ? j.speakers[idx[b]] : 0;
        if (nt[b] > 0 && nt[b] <= j.max_text) memcpy(&tok[(size_t)b * local_text], j.tokens + (size_t)idx[b] * j.max_text, (size_t)nt[b] * 4);
    }
    if (mgb_encode_text(s, tok.data(), nt.data(), nullptr) != MGB_OK || mgb_prefill(s, spk.data()) != MGB_OK) err = mgb_last_error();
    if (err.empty()) {
        const size_t row = (size_t)j.T * 8;
        if (j.teacher) {
            std::vector<int32_t> in((size_t)nb * row), gr((size_t)nb * row);
            for (int b = 0; b < nb; b++) memcpy(&in[b * row], j.codes_in + idx[b] * row, row * 4);
            if (mgb_teacher_forced(s, in.data(), j.T, nullptr, nullptr, j.greedy_out ? gr.data() : nullptr) != MGB_OK) err = mgb_last_error();
            else if (j.greedy_out) for (int b = 0; b < nb; b++) memcpy(j.greedy_out + idx[b] * row, &gr[b * row], row * 4);
        } else {
            std::vector<int32_t> out((size_t)nb * row), nf(nb);
            if (mgb_generate(s, j.T, j.temperature, j.top_k, nullptr, j.seed + (uint64_t)di, j.ignore_eos, out.data(), nf.data(), nullptr) != MGB_OK) err = mgb_last_error();
            else for (int b = 0; b < nb; b++) {
                memcpy(j.codes_out + idx[b] * row, &out[b * row], row * 4);
                j.n_frames_out[idx[b]] = nf[b];
            }
        }
        *ms_out = mgb_session_last_loop_ms(s);
    }
    if (!err.empty()) { mgb_session_free(s); cs.s = nullptr; }      // do not reuse a session that failed
    return err;
}

int run_job(mgb_pool * p, const Job & j, float * device_ms_out) {
    if (!p || j.n_utt <= 0 || !j.tokens || !j.n_tokens || j.max_text <= 0 || j.T <= 0) { mgb::set_error("mgb_pool: invalid arguments"); return MGB_EINVAL; }
    for (int i = 0; i < j.n_utt; i++)
        if (j.n_tokens[i] <= 0 || j.n_tokens[i] > j.max_text) { mgb::set_error("mgb_pool: token count out of range"); return MGB_EINVAL; }
    const int G = (int)p->models.size();
    std::vector<std::string> errs(G);
    std::vector<std::thread> th;
    p->last_ms.assign(G, 0.0f);
    try {
        for (int di = 0; di < G; di++)
            th.emplace_back([&, di] {
                try { errs[di] = run_share(p, di, j, &p->last_ms[di]); }
                catch (const std::exception & e) { errs[di] = e.what(); }
                catch (...) { errs[di] = "unknown failure"; }
            });
    } catch (const std::exception & e) { errs[0] = std::string("thread creation failed: ") + e.what(); }
    for (std::thread & t : th) t.join();
    if (device_ms_out) for (int di = 0; di < G; di++) device_ms_out[di] = p->last_ms[di];
    for (int di = 0; di < G; di++)
        if (!errs[di].empty()) { mgb::set_error("mgb_pool (device " + std::to_string(p->devices[di]) + "): " + errs[di]); return MGB_ECUDA; }
    return MGB_OK;
}

}  // namespace

extern "C" {

mgb_pool * mgb_pool_new(const char * gguf_path, const int * devices, int n_devices, int precision) {
    if (!gguf_path) { mgb::set_error("mgb_pool_new: null path"); return nullptr; }
    try {
        std::vector<int> devs;
        if (devices && n_devices > 0) devs.assign(devices, devices + n_devices);
        else {
            const int n = mgb_device_count();
            if (n <= 0) { mgb::set_error("no CUDA device available (this build has no CPU fallback)"); return nullptr; }
            for (int d = 0; d < n; d++) devs.push_back(d);
        }
        mgb_pool * p = new mgb_pool();
        p->devices = devs;
        p->models.assign(devs.size(), nullptr);
        p->sessions.assign(devs.size(), mgb_pool::Cached());
        // replicas are loaded concurrently (each load parses the file and uploads ~0.2-0.9 GB to its own device)
        std::vector<std::string> errs(devs.size());
        std::vector<std::thread> th;
        for (size_t i = 0; i < devs.size(); i++)
            th.emplace_back([&, i] {
                p->models[i] = mgb_model_load(gguf_path, devs[i], precision);
                if (!p->models[i]) errs[i] = mgb_last_error();
            });
        for (std::thread & t : th) t.join();
        for (size_t i = 0; i < devs.size(); i++)
            if (!p->models[i]) { mgb::set_error("mgb_pool_new (device " + std::to_string(devs[i]) + "): " + errs[i]); delete p; return nullptr; }
        return p;
    } catch (const std::exception & e) { mgb::set_error(std::string("mgb_pool_new: ") + e.what()); return nullptr; }
}

void mgb_pool_free(mgb_pool * p) { delete p; }
int mgb_pool_n_devices(const mgb_pool * p) { return p ? (int)p->models.size() : MGB_EINVAL; }
mgb_model * mgb_pool_model(mgb_pool * p, int i) { return (p && i >= 0 && i < (int)p->models.size()) ? p->models[i] : nullptr; }

int mgb_pool_generate(mgb_pool * p, int n_utt, const int32_t * tokens, const int32_t * n_tokens, int max_text, const int32_t * speakers,
                      int max_steps, float temperature, int top_k, uint64_t seed, int ignore_eos,
                      int32_t * codes_out, int32_t * n_frames_out, float * device_ms_out) {
    if (!codes_out || !n_frames_out) { mgb::set_error("mgb_pool_generate: null outputs"); return MGB_EINVAL; }
    Job j;
    j.n_utt = n_utt; j.tokens = tokens; j.n_tokens = n_tokens; j.max_text = max_text; j.speakers = speakers; j.T = max_steps;
    j.temperature = temperature; j.top_k = top_k; j.seed = seed; j.ignore_eos = ignore_eos; j.codes_out = codes_out; j.n_frames_out = n_frames_out;
    return run_job(p, j, device_ms_out);
}

int mgb_pool_teacher_forced(mgb_pool * p, int n_utt, const int32_t * tokens, const int32_t * n_tokens, int max_text, const int32_t * speakers,
                            const int32_t * codes_in, int T, int32_t * greedy_out, float * device_ms_out) {
    if (!codes_in) { mgb::set_error("mgb_pool_teacher_forced: null codes"); return MGB_EINVAL; }
    Job j;
    j.n_utt = n_utt; j.tokens = tokens; j.n_tokens = n_tokens; j.max_text = max_text; j.speakers = speakers; j.T = T;
    j.teacher = true; j.codes_in = codes_in; j.greedy_out = greedy_out;
    return run_job(p, j, device_ms_out);
}

}  // extern "C"
