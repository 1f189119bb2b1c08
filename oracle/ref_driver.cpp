// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A C-ABI batch driver around the UNMODIFIED reference engine.  It is compiled
// together with the reference's own sources, read in place from
// /root/reference/engine/{include,src} (see oracle/Makefile), into
// oracle/_ref/libnnue_ref.so.  No reference source is copied into this repo.
//
// It replaces the reference's per-sample process boundary
// (engine/nnue_inference.cpp:11-65, driven by evaluate.py:153-173) with an
// in-process loop so that the oracle and the CPU baseline do not pay a
// fork/exec + model reload per image.  NNUEEvaluator keeps mutable scratch and
// is not thread-safe (engine/include/nnue_engine.h:559-570), so the driver
// holds one evaluator per worker thread.
#include <cstdint>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "nnue_engine.h"

namespace {
struct RefHandle {
    std::string path;
    std::vector<std::unique_ptr<nnue::NNUEEvaluator>> evals;
    bool ensure(size_t n) {
        while (evals.size() < n) {
            auto e = std::make_unique<nnue::NNUEEvaluator>();
            if (!e->load_model(path)) return false;
            evals.push_back(std::move(e));
        }
        return true;
    }
};
}  // namespace

extern "C" {

void *ref_load(const char *path) {
    auto *h = new RefHandle{path, {}};
    if (!h->ensure(1)) { delete h; return nullptr; }
    return h;
}

void ref_free(void *hp) { delete static_cast<RefHandle *>(hp); }

// dims[] = F, L1, L2, L3, (unused), OC, G, n_buckets
void ref_dims(void *hp, int *dims) {
    auto &e = *static_cast<RefHandle *>(hp)->evals[0];
    dims[0] = e.get_num_features(); dims[1] = e.get_l1_size(); dims[2] = e.get_l2_size();
    dims[3] = e.get_l3_size(); dims[4] = -1; dims[5] = e.get_num_channels_per_square();
    dims[6] = e.get_grid_size(); dims[7] = e.get_num_layer_stacks();
}

// images [B][H*W*3] raw floats; logits [B][NC]; density [B] computed exactly as
// nnue_inference.cpp:50-54 does.  Returns NC (>0) or a negative error.
int ref_eval_batch(void *hp, const float *imgs, int B, int H, int W, float *logits,
                   float *density, int nthreads) {
    auto *h = static_cast<RefHandle *>(hp);
    if (nthreads < 1) nthreads = 1;
    if (nthreads > B) nthreads = B > 0 ? B : 1;
    if (!h->ensure(static_cast<size_t>(nthreads))) return -1;
    const size_t img_elems = static_cast<size_t>(H) * W * 3;
    std::vector<int> nc(static_cast<size_t>(nthreads), 0);
    auto work = [&](int t) {
        auto &e = *h->evals[static_cast<size_t>(t)];
        std::vector<int> active;
        for (int b = t; b < B; b += nthreads) {
            std::vector<float> out = e.evaluate_logits(imgs + static_cast<size_t>(b) * img_elems, H, W, 0);
            if (out.empty()) { nc[static_cast<size_t>(t)] = -2; return; }
            nc[static_cast<size_t>(t)] = static_cast<int>(out.size());
            for (size_t c = 0; c < out.size(); ++c) logits[static_cast<size_t>(b) * out.size() + c] = out[c];
            e.get_active_features(active);
            const int total = e.get_total_features();
            density[b] = total > 0 ? static_cast<float>(active.size()) / total : 0.0f;
        }
    };
    if (nthreads == 1) {
        work(0);
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < nthreads; ++t) pool.emplace_back(work, t);
        for (auto &th : pool) th.join();
    }
    for (int v : nc) if (v < 0) return v;
    return B > 0 ? nc[0] : 0;
}

// Number of classes of the loaded model (probe with one zero image).
int ref_num_classes(void *hp, int H, int W) {
    auto *h = static_cast<RefHandle *>(hp);
    std::vector<float> img(static_cast<size_t>(H) * W * 3, 0.0f);
    return static_cast<int>(h->evals[0]->evaluate_logits(img.data(), H, W, 0).size());
}

// Incremental-accumulator surface (nnue_engine.cpp:739-821): refresh with `cur`
// features, then apply added/removed, and return the accumulator-driven score.
float ref_eval_incremental(void *hp, const int *feats, int n) {
    auto &e = *static_cast<RefHandle *>(hp)->evals[0];
    std::vector<int> v(feats, feats + n);
    return e.evaluate_incremental(v, 0);
}
void ref_mark_dirty(void *hp) { static_cast<RefHandle *>(hp)->evals[0]->mark_dirty(); }

}  // extern "C"
