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
#include <linux/futex.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <climits>
#include <cmath>
#include <condition_variable>
#include <cstring>
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
//
// Any number of threads call dawn_batcher_search with ONE query each; they are answered through
// dawn_index_search_batch in batches.  r02 rewrite, after measuring the first version under closed-loop load
// (tools/batcher_bench.cpp: 21k queries/s with 1024 callers over 10M rows, p99 1 s -- every completion woke every
// waiting caller through one mutex, and the worker gathered the queries and scattered the results itself):
//
//   * a caller reserves a slot of the OPEN batch buffer under the lock and copies its query into it there (1.5 KB), then
//     sleeps on that buffer's generation word (a futex); after the wake-up it copies its own k results out.  Gather and
//     scatter are thus done by the callers, in parallel, and a completion wakes only that batch's callers, who then touch
//     no shared lock;
//   * a batch is whatever arrived while the previous batch was on the GPU (+ at most max_wait_us after the first arrival
//     when the GPU was idle, and only once concurrent callers have been seen: a lone caller -- the reference's pattern --
//     never waits for a window).  Batches do NOT overlap on the GPU: below ~256 queries a batch costs one pass over the
//     corpus whatever its size, so two half batches side by side would cost twice one whole batch;
//   * two executor threads take turns: while one wakes its batch's callers (about a microsecond per sleeping thread), the
//     other already has the next batch on the GPU.

struct dawn_batcher {
    struct Batch {
        std::vector<float> q;         // max_batch x 384, filled by the callers
        std::vector<uint64_t> labels; // max_batch x k results, read by the callers
        std::vector<float> dist;
        std::vector<size_t> counts;
        size_t n = 0, k = 0;
        std::chrono::steady_clock::time_point t_first;
        uint32_t gen = 0;                  // generation being filled / run (under mu)
        std::atomic<uint32_t> done{0};     // last generation whose results are ready (futex word)
        std::atomic<uint32_t> readers{0};  // callers of the finished generation still copying out
        int rc = DAWN_OK;
        std::string err;
    };
    static constexpr int kRing = 4;
    dawn_index *idx = nullptr;           // exactly one of idx / multi is set
    dawn_multi *multi = nullptr;
    size_t max_batch = 256;
    uint32_t max_wait_us = 200;
    std::mutex mu;                       // open batch + counters
    std::condition_variable cv_work;     // executors: "the open batch has its first request / is full"
    std::condition_variable cv_space;    // callers: "a new batch is open"
    std::mutex gpu_mu;                   // one batch on the GPU at a time
    Batch ring[kRing];
    int open = 0;
    bool stop = false;
    size_t last_batch = 0;               // size of the previous batch: its callers are the ones about to come back
    size_t window_target = 0;            // an executor is waiting for the open batch to reach this size (0: none; under mu)
    std::thread workers[2];
    uint64_t n_batches = 0, n_queries = 0, max_seen = 0;

    static void futex_wait(std::atomic<uint32_t> *w, uint32_t seen) {
        syscall(SYS_futex, reinterpret_cast<uint32_t *>(w), FUTEX_WAIT_PRIVATE, seen, nullptr, nullptr, 0);
    }
    static void futex_wake_all(std::atomic<uint32_t> *w) {
        syscall(SYS_futex, reinterpret_cast<uint32_t *>(w), FUTEX_WAKE_PRIVATE, INT_MAX, nullptr, nullptr, 0);
    }

    void run() {
        while (true) {
            // the GPU turn: waiting here is what lets the open batch fill while the other executor's batch runs
            std::unique_lock<std::mutex> gpu(gpu_mu);
            std::unique_lock<std::mutex> lk(mu);
            const size_t n_at_turn = ring[open].n;  // joined while the previous batch was on the GPU
            const auto t_turn = std::chrono::steady_clock::now();
            cv_work.wait(lk, [&] { return stop || ring[open].n > 0; });
            Batch *b = &ring[open];
            if (b->n == 0) return;  // stop, nothing pending
            // The window.  The GPU has just become free (this thread holds the turn), and the callers of the batch that just
            // finished -- last_batch of them -- are being woken and, in a closed loop, are about to come back: give them up
            // to max_wait_us to join, counted from the start of the turn or from the batch's first arrival, whichever is
            // later, and go as soon as they are all back.  Below ~256 queries a batch costs one pass over the corpus whatever
            // its size, so two alternating half-cohorts would halve the throughput.  A lone caller is its own cohort: it
            // never waits.
            {
                const size_t expected = std::min(max_batch, n_at_turn + last_batch);
                const auto deadline = std::max(t_turn, b->t_first) + std::chrono::microseconds(max_wait_us);
                window_target = expected;
                while (b->n < expected && !stop) {
                    if (cv_work.wait_until(lk, deadline) == std::cv_status::timeout) break;
                }
                window_target = 0;
            }
            // close the batch: the next buffer of the ring opens once the callers of its last generation have left
            int next = (open + 1) % kRing;
            while (ring[next].readers.load(std::memory_order_acquire) != 0) {
                lk.unlock();
                std::this_thread::yield();
                lk.lock();
            }
            ring[next].n = 0;
            ring[next].gen++;
            open = next;
            const size_t n = b->n, k = b->k;
            last_batch = n;
            n_batches++;
            n_queries += n;
            if (n > max_seen) max_seen = n;
            // from here on the buffer cannot reopen until its n callers have copied their results out (set now, not at
            // publish time: this thread may lose the CPU between handing the GPU on and publishing)
            b->readers.store((uint32_t)n, std::memory_order_relaxed);
            lk.unlock();
            cv_space.notify_all();
            b->labels.resize(n * (k ? k : 1));
            b->dist.resize(n * (k ? k : 1));
            b->counts.resize(n);
            b->rc = multi ? dawn_multi_search_batch(multi, b->q.data(), n, k, b->labels.data(), b->dist.data(), b->counts.data())
                          : dawn_index_search_batch(idx, b->q.data(), n, k, b->labels.data(), b->dist.data(), b->counts.data());
            b->err = b->rc == DAWN_OK ? "" : dawn_last_error();
            gpu.unlock();
            // publish: one wake-up for the whole batch, issued while the other executor's batch is already running
            b->done.store(b->gen, std::memory_order_release);
            futex_wake_all(&b->done);
        }
    }
};

extern "C" {

static int batcher_create(dawn_index *idx, dawn_multi *multi, size_t max_batch, uint32_t max_wait_us, dawn_batcher **out) {
    if ((!idx && !multi) || !out || max_batch == 0) return DAWN_ERR_INVALID;
    dawn_batcher *b = new (std::nothrow) dawn_batcher();
    if (!b) return DAWN_ERR_INTERNAL;
    b->idx = idx;
    b->multi = multi;
    b->max_batch = max_batch;
    b->max_wait_us = max_wait_us;
    for (auto &r : b->ring) {
        r.q.resize(max_batch * kDimF);
        r.gen = 1;
    }
    for (auto &w : b->workers) w = std::thread([b] { b->run(); });
    *out = b;
    return DAWN_OK;
}

int dawn_batcher_create(dawn_index *idx, size_t max_batch, uint32_t max_wait_us, dawn_batcher **out) {
    return batcher_create(idx, nullptr, max_batch, max_wait_us, out);
}

// The same front over the one-process multi-GPU handle: single-query callers -> batches -> every shard -> merge.
int dawn_batcher_create_multi(dawn_multi *m, size_t max_batch, uint32_t max_wait_us, dawn_batcher **out) {
    return batcher_create(nullptr, m, max_batch, max_wait_us, out);
}

// Blocking and thread-safe: many threads call this with one query each; they are answered in batches.
// Results are exactly those of dawn_index_search.
int dawn_batcher_search(dawn_batcher *b, const float *query384, size_t k, uint64_t *labels_out,
                        float *distances_out, size_t *count_out) {
    if (!b || !query384 || !count_out || (k && (!labels_out || !distances_out))) return DAWN_ERR_INVALID;
    dawn_batcher::Batch *B;
    size_t slot;
    uint32_t gen;
    bool wake;
    {
        std::unique_lock<std::mutex> lk(b->mu);
        // one k per batch (the reference always asks for 20); a full batch, or one with another k, makes the caller wait
        // for the next one to open
        b->cv_space.wait(lk, [&] {
            const dawn_batcher::Batch &o = b->ring[b->open];
            return b->stop || o.n == 0 || (o.n < b->max_batch && o.k == k);
        });
        if (b->stop) return DAWN_ERR_INVALID;
        B = &b->ring[b->open];
        slot = B->n++;
        gen = B->gen;
        memcpy(&B->q[slot * kDimF], query384, kDimF * sizeof(float));
        if (slot == 0) {
            B->k = k;
            B->t_first = std::chrono::steady_clock::now();
        }
        wake = slot == 0 || (b->window_target && B->n >= b->window_target);
    }
    if (wake) b->cv_work.notify_one();
    for (uint32_t seen; (seen = B->done.load(std::memory_order_acquire)) != gen;) dawn_batcher::futex_wait(&B->done, seen);
    const int rc = B->rc;
    if (rc == DAWN_OK) {
        const size_t c = B->counts[slot];
        if (c) {
            memcpy(labels_out, &B->labels[slot * k], c * sizeof(uint64_t));
            memcpy(distances_out, &B->dist[slot * k], c * sizeof(float));
        }
        *count_out = c;
    } else {
        strncpy(g_front_err, B->err.c_str(), sizeof g_front_err - 1);
        g_front_err[sizeof g_front_err - 1] = 0;
    }
    B->readers.fetch_sub(1, std::memory_order_release);
    return rc;
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

// Pending requests are answered first; callers arriving afterwards get DAWN_ERR_INVALID.  The caller of free must make sure
// no thread is still inside dawn_batcher_search when it returns from its last call (as with any handle).
void dawn_batcher_free(dawn_batcher *b) {
    if (!b) return;
    {
        std::lock_guard<std::mutex> lk(b->mu);
        b->stop = true;
    }
    b->cv_work.notify_all();
    b->cv_space.notify_all();
    for (auto &w : b->workers)
        if (w.joinable()) w.join();
    for (auto &r : b->ring)  // callers of the last batches still copying their results out
        while (r.readers.load(std::memory_order_acquire) != 0) std::this_thread::yield();
    delete b;
}

}  // extern "C"
