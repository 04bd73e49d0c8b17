// dawn_front.cu -- host-side pieces either side of the hot path (SURVEY.md section 8f "next" rows):
//
//   (f1) micro-batching front: the reference submits ONE query at a time from one blocking thread
//        (/root/reference/src/search/search_service.rs:55-104, channel depth 2 at
//        src/bin/dawnsearch.rs:60).  dawn_batcher coalesces concurrent single-query callers into
//        dawn_index_search_batch calls so the tensor-core path gets real batches.
//   (f2) bulk load of the legacy `.emb` flat files: arrays of repr(C) PageEntry
//        (src/index/warc.rs:35-43; reader examples_old/document_embeddings.rs:60-71).
//   (f3) distance_limit of UdpPacket::Search (src/net/udp_packets.rs:29-39): hits with
//        distance >= limit are not returned (src/net/udp_service.rs:196-199).
//   (f4) the i24 wire codec of query embeddings (src/search/vector.rs:48-87) and the
//        normalisation gate (vector.rs:181-197), so the UDP layer can hand over raw 1152-byte
//        queries and fetch stored vectors in wire format.
//
// Pure host code (no kernels); compiled into libdawn_b200.so next to the CUDA sources.
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/dawn_index.h"

namespace {
constexpr int kDimF = DAWN_DIMENSIONS;
thread_local char g_front_err[256];
}  // namespace

extern "C" {

// ------------------------------------------------------------------ (f4) vector.rs mirrors

// src/search/vector.rs:181-183: sqrt of the sequential f32 sum of squares.
float dawn_vector_length(const float *v) {
    float acc = 0.0f;
    for (int i = 0; i < kDimF; i++) {
        float d = v[i] - 0.0f;
        acc += d * d;
    }
    return sqrtf(acc);
}

// src/search/vector.rs:185-192: finite and 0.99 < |v| < 1.01.
int dawn_is_normalized(const float *v) {
    float l = dawn_vector_length(v);
    if (!std::isfinite(l)) return 0;
    return l > 1.0f - 0.01f && l < 1.0f + 0.01f;
}

// src/search/vector.rs:194-197
void dawn_normalize(float *v) {
    float acc = 0.0f;
    for (int i = 0; i < kDimF; i++) acc += v[i] * v[i];
    float length = sqrtf(acc);
    for (int i = 0; i < kDimF; i++) v[i] /= length;
}

// src/search/vector.rs:74-86: trunc(((x+1)/2) * 0x7FFFFF), 3 bytes little endian per element.
void dawn_encode_i24(const float *v, uint8_t *out1152) {
    for (int i = 0; i < kDimF; i++) {
        double t = (((double)v[i] + 1.0) / 2.0) * (double)0x7FFFFF;
        int32_t x;
        if (t != t) x = 0;
        else if (t >= 2147483647.0) x = INT32_MAX;
        else if (t <= -2147483648.0) x = INT32_MIN;
        else x = (int32_t)t;
        out1152[i * 3 + 0] = (uint8_t)(x & 0xFF);
        out1152[i * 3 + 1] = (uint8_t)((x >> 8) & 0xFF);
        out1152[i * 3 + 2] = (uint8_t)((x >> 16) & 0xFF);
    }
}

// src/search/vector.rs:52-72 (including its `v |= 0xFF` when the top bit of the high byte is set).
// Returns DAWN_OK, or DAWN_ERR_INVALID when the decoded vector fails the normalisation gate
// (the reference's `ensure!(is_normalized(&result))`).
int dawn_decode_i24(const uint8_t *in1152, float *out384) {
    for (int i = 0; i < kDimF; i++) {
        int32_t v = 0;
        v |= (int32_t)in1152[i * 3];
        v |= (int32_t)in1152[i * 3 + 1] << 8;
        v |= (int32_t)in1152[i * 3 + 2] << 16;
        if ((in1152[i * 3 + 2] & 0x80) > 0) v |= 0xFF;
        out384[i] = (float)((double)v / (double)0x7FFFFF * 2.0 - 1.0);
    }
    return dawn_is_normalized(out384) ? DAWN_OK : DAWN_ERR_INVALID;
}

// ------------------------------------------------------------------ (f3) distance limit
// dawn_index_search_limit lives in dawn_index.cu: the limit is pushed down into the scan kernels.

// The peer side of UdpPacket::Search: raw 1152-byte i24 query in, hits below the limit out
// (udp_service.rs:174-213 without the SQLite hydration).  has_limit = 0 mirrors `None`.
int dawn_index_search_i24(dawn_index *idx, const uint8_t *query1152, size_t k, int has_limit, float distance_limit,
                          uint64_t *labels_out, float *distances_out, size_t *count_out) {
    float q[kDimF];
    if (dawn_decode_i24(query1152, q) != DAWN_OK) return DAWN_ERR_INVALID;  // "Embedding is not normalized"
    if (!has_limit) return dawn_index_search(idx, q, k, labels_out, distances_out, count_out);
    return dawn_index_search_limit(idx, q, k, distance_limit, labels_out, distances_out, count_out);
}

// Stored vector of `label` in wire format (GetEmbedding over UDP, udp_service.rs:254-276).
int dawn_index_get_i24(dawn_index *idx, uint64_t label, uint8_t *out1152) {
    float v[kDimF];
    int rc = dawn_index_get(idx, label, v);
    if (rc != DAWN_OK) return rc;
    dawn_encode_i24(v, out1152);
    return DAWN_OK;
}

// ------------------------------------------------------------------ (f2) legacy .emb bulk load

// repr(C) PageEntry, src/index/warc.rs:35-43: url_pos u64, title_pos u64, vector [f32;384],
// url_len u64, title_len u64  ->  1568 bytes, vector at offset 16.
#define DAWN_PAGE_ENTRY_BYTES 1568
#define DAWN_PAGE_ENTRY_VECTOR_OFFSET 16

// Append the vectors of n PageEntry records (an mmapped `.emb` file) with labels
// first_label, first_label+1, ...  Entries failing the normalisation gate are skipped and counted
// in *skipped (the reference's loaders bail on them, vector.rs:199-204).
int dawn_index_add_page_entries(dawn_index *idx, const void *entries, size_t n, uint64_t first_label,
                                size_t *skipped) {
    if (!entries && n) return DAWN_ERR_INVALID;
    const size_t kChunk = 4096;
    std::vector<float> vecs(kChunk * kDimF);
    std::vector<uint64_t> labels(kChunk);
    size_t bad = 0, fill = 0;
    const uint8_t *p = static_cast<const uint8_t *>(entries);
    for (size_t i = 0; i < n; i++) {
        const float *v = reinterpret_cast<const float *>(p + i * DAWN_PAGE_ENTRY_BYTES + DAWN_PAGE_ENTRY_VECTOR_OFFSET);
        float tmp[kDimF];
        memcpy(tmp, v, sizeof tmp);  // records are only 8-byte aligned
        if (!dawn_is_normalized(tmp)) {
            bad++;
            continue;
        }
        memcpy(&vecs[fill * kDimF], tmp, sizeof tmp);
        labels[fill] = first_label + i;
        if (++fill == kChunk) {
            int rc = dawn_index_add_batch(idx, labels.data(), vecs.data(), fill);
            if (rc != DAWN_OK) return rc;
            fill = 0;
        }
    }
    if (fill) {
        int rc = dawn_index_add_batch(idx, labels.data(), vecs.data(), fill);
        if (rc != DAWN_OK) return rc;
    }
    if (skipped) *skipped = bad;
    return DAWN_OK;
}

}  // extern "C"

// ------------------------------------------------------------------ (f1) micro-batching front

struct dawn_batcher {
    struct Request {
        const float *q;
        size_t k;
        uint64_t *labels;
        float *dist;
        size_t *count;
        int rc = 1;  // 1 = pending
        std::string err;
    };
    dawn_index *idx = nullptr;
    size_t max_batch = 256;
    uint32_t max_wait_us = 200;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::deque<Request *> queue;
    bool stop = false;
    std::thread worker;
    uint64_t n_batches = 0, n_queries = 0, max_seen = 0;

    void run() {
        std::vector<Request *> batch;
        std::vector<float> qbuf;
        std::vector<uint64_t> lbuf;
        std::vector<float> dbuf;
        std::vector<size_t> cbuf;
        std::unique_lock<std::mutex> lk(mu);
        while (true) {
            cv_work.wait(lk, [&] { return stop || !queue.empty(); });
            if (stop && queue.empty()) return;
            // the first request opens a window: wait for more, up to max_wait_us or max_batch
            auto deadline = std::chrono::steady_clock::now() + std::chrono::microseconds(max_wait_us);
            while (queue.size() < max_batch && !stop) {
                if (cv_work.wait_until(lk, deadline) == std::cv_status::timeout) break;
            }
            // one batch = the queued requests that share the head request's k, in arrival order
            batch.clear();
            const size_t k = queue.front()->k;
            for (auto it = queue.begin(); it != queue.end() && batch.size() < max_batch;) {
                if ((*it)->k == k) {
                    batch.push_back(*it);
                    it = queue.erase(it);
                } else {
                    ++it;
                }
            }
            lk.unlock();
            const size_t b = batch.size();
            qbuf.resize(b * kDimF);
            lbuf.resize(b * (k ? k : 1));
            dbuf.resize(b * (k ? k : 1));
            cbuf.resize(b);
            for (size_t i = 0; i < b; i++) memcpy(&qbuf[i * kDimF], batch[i]->q, kDimF * sizeof(float));
            int rc = dawn_index_search_batch(idx, qbuf.data(), b, k, lbuf.data(), dbuf.data(), cbuf.data());
            std::string err = rc == DAWN_OK ? "" : dawn_last_error();
            lk.lock();
            n_batches++;
            n_queries += b;
            if (b > max_seen) max_seen = b;
            for (size_t i = 0; i < b; i++) {
                Request *r = batch[i];
                if (rc == DAWN_OK) {
                    memcpy(r->labels, &lbuf[i * k], cbuf[i] * sizeof(uint64_t));
                    memcpy(r->dist, &dbuf[i * k], cbuf[i] * sizeof(float));
                    *r->count = cbuf[i];
                } else {
                    r->err = err;
                }
                r->rc = rc;
            }
            cv_done.notify_all();
        }
    }
};

extern "C" {

int dawn_batcher_create(dawn_index *idx, size_t max_batch, uint32_t max_wait_us, dawn_batcher **out) {
    if (!idx || !out || max_batch == 0) return DAWN_ERR_INVALID;
    dawn_batcher *b = new (std::nothrow) dawn_batcher();
    if (!b) return DAWN_ERR_INTERNAL;
    b->idx = idx;
    b->max_batch = max_batch;
    b->max_wait_us = max_wait_us;
    b->worker = std::thread([b] { b->run(); });
    *out = b;
    return DAWN_OK;
}

// Blocking and thread-safe: many threads call this with one query each; the worker answers them
// in batches.  Results are exactly those of dawn_index_search.
int dawn_batcher_search(dawn_batcher *b, const float *query384, size_t k, uint64_t *labels_out,
                        float *distances_out, size_t *count_out) {
    if (!b || !query384 || !count_out || (k && (!labels_out || !distances_out))) return DAWN_ERR_INVALID;
    dawn_batcher::Request r;
    r.q = query384;
    r.k = k;
    r.labels = labels_out;
    r.dist = distances_out;
    r.count = count_out;
    std::unique_lock<std::mutex> lk(b->mu);
    if (b->stop) return DAWN_ERR_INVALID;
    b->queue.push_back(&r);
    b->cv_work.notify_one();
    b->cv_done.wait(lk, [&] { return r.rc != 1; });
    if (r.rc != DAWN_OK) {
        strncpy(g_front_err, r.err.c_str(), sizeof g_front_err - 1);
        g_front_err[sizeof g_front_err - 1] = 0;
    }
    return r.rc;
}

const char *dawn_batcher_last_error(void) { return g_front_err; }

int dawn_batcher_stats(dawn_batcher *b, uint64_t *batches, uint64_t *queries, uint64_t *largest_batch) {
    if (!b) return DAWN_ERR_INVALID;
    std::lock_guard<std::mutex> lk(b->mu);
    if (batches) *batches = b->n_batches;
    if (queries) *queries = b->n_queries;
    if (largest_batch) *largest_batch = b->max_seen;
    return DAWN_OK;
}

void dawn_batcher_free(dawn_batcher *b) {
    if (!b) return;
    {
        std::lock_guard<std::mutex> lk(b->mu);
        b->stop = true;
    }
    b->cv_work.notify_all();
    if (b->worker.joinable()) b->worker.join();
    delete b;
}

}  // extern "C"
