// Closed-loop load on the micro-batching front: T caller threads, each submitting ONE query at a time through
// dawn_batcher_search -- the reference's call pattern (src/search/search_service.rs:55-104), many callers instead of one.
// Prints one JSON line per T: queries/s, per-query latency p50/p99, mean batch size; first the single-caller baseline through
// dawn_index_search (what the reference's one blocking thread would get).  After each timed window every caller thread
// compares two more batched answers, bit for bit, with dawn_index_search on the same query (not inside the window: an
// unbatched search costs a whole pass over the corpus and would eat the GPU time being measured).
//
// build: g++ -O2 -std=c++17 -I include tools/batcher_bench.cpp -o tools/bin/batcher_bench -L dawnsearch_b200/lib -ldawn_b200
//        -Wl,-rpath,$PWD/dawnsearch_b200/lib -lpthread
// usage: batcher_bench <rows> <k> <seconds per point> <max_batch> <max_wait_us> <T,T,...> [gpus]
//        gpus > 1: the corpus is sharded over devices 0..gpus-1 behind one dawn_multi handle (NCCL inside the library)
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "dawn_index.h"

using Clock = std::chrono::steady_clock;

static void make_query(std::mt19937_64 &rng, float *q) {
    std::normal_distribution<float> nd(0.f, 1.f);
    double n2 = 0;
    for (int i = 0; i < 384; i++) {
        q[i] = nd(rng);
        n2 += (double)q[i] * q[i];
    }
    const float inv = (float)(1.0 / std::sqrt(n2));
    for (int i = 0; i < 384; i++) q[i] *= inv;
}

static double pct(std::vector<double> &v, double p) {
    if (v.empty()) return 0;
    std::sort(v.begin(), v.end());
    return v[std::min(v.size() - 1, (size_t)(p * v.size()))];
}

int main(int argc, char **argv) {
    const size_t rows = argc > 1 ? strtoull(argv[1], nullptr, 10) : 10000000;
    const size_t k = argc > 2 ? strtoull(argv[2], nullptr, 10) : 10;
    const double secs = argc > 3 ? atof(argv[3]) : 3.0;
    const size_t max_batch = argc > 4 ? strtoull(argv[4], nullptr, 10) : 1024;
    const uint32_t max_wait_us = argc > 5 ? (uint32_t)atoi(argv[5]) : 100;
    std::vector<int> threads;
    {
        std::string s = argc > 6 ? argv[6] : "1,16,64,256,1024";
        size_t p = 0;
        while (p < s.size()) {
            threads.push_back(atoi(s.c_str() + p));
            p = s.find(',', p);
            if (p == std::string::npos) break;
            p++;
        }
    }
    const int gpus = argc > 7 ? atoi(argv[7]) : 1;
    dawn_index *idx = nullptr;
    dawn_multi *multi = nullptr;
    if (gpus > 1) {
        std::vector<int> devs(gpus);
        for (int g = 0; g < gpus; g++) devs[g] = g;
        if (dawn_multi_create(devs.data(), devs.size(), DAWN_SCALAR_F16, &multi) != DAWN_OK ||
            dawn_multi_reserve(multi, rows) != DAWN_OK || dawn_multi_add_synthetic(multi, 0xDA5EA2C4ull, 0, rows) != DAWN_OK) {
            fprintf(stderr, "multi: %s\n", dawn_multi_last_error());
            return 1;
        }
    } else {
        dawn_options o;
        memset(&o, 0, sizeof o);
        o.dimensions = 384;
        o.metric = DAWN_METRIC_IP;
        o.scalar = DAWN_SCALAR_F16;
        o.capacity = rows;
        if (dawn_index_create(&o, &idx) != DAWN_OK) {
            fprintf(stderr, "create: %s\n", dawn_last_error());
            return 1;
        }
        if (dawn_index_add_synthetic(idx, 0xDA5EA2C4ull, 0, rows) != DAWN_OK) {
            fprintf(stderr, "add_synthetic: %s\n", dawn_last_error());
            return 1;
        }
    }
    auto direct = [&](const float *q, uint64_t *l, float *d, size_t *c) {
        return multi ? dawn_multi_search(multi, q, k, l, d, c) : dawn_index_search(idx, q, k, l, d, c);
    };
    // --- the reference's pattern: one caller, one query at a time, straight through the index ---
    {
        std::mt19937_64 rng(7);
        std::vector<float> q(384);
        std::vector<uint64_t> l(k);
        std::vector<float> d(k);
        size_t c = 0;
        std::vector<double> lat;
        for (int i = 0; i < 20; i++) {
            make_query(rng, q.data());
            direct(q.data(), l.data(), d.data(), &c);
        }
        const auto t0 = Clock::now();
        while (std::chrono::duration<double>(Clock::now() - t0).count() < secs) {
            make_query(rng, q.data());
            const auto a = Clock::now();
            if (direct(q.data(), l.data(), d.data(), &c) != DAWN_OK) return 2;
            lat.push_back(std::chrono::duration<double, std::milli>(Clock::now() - a).count());
        }
        const double el = std::chrono::duration<double>(Clock::now() - t0).count();
        printf("{\"front\": \"%s, one caller\", \"gpus\": %d, \"rows\": %zu, \"k\": %zu, \"callers\": 1, \"qps\": %.1f, "
               "\"latency_ms_p50\": %.3f, \"latency_ms_p99\": %.3f}\n",
               multi ? "dawn_multi_search" : "dawn_index_search", gpus, rows, k, lat.size() / el, pct(lat, 0.5), pct(lat, 0.99));
        fflush(stdout);
    }
    for (int T : threads) {
        dawn_batcher *b = nullptr;
        if ((multi ? dawn_batcher_create_multi(multi, max_batch, max_wait_us, &b) : dawn_batcher_create(idx, max_batch, max_wait_us, &b)) != DAWN_OK) return 3;
        std::atomic<bool> go{false}, stop{false};
        std::atomic<uint64_t> mismatches{0}, errors{0}, checks{0};
        std::vector<std::vector<double>> lats(T);
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++) {
            th.emplace_back([&, t] {
                std::mt19937_64 rng(1000 + t);
                std::vector<float> q(384);
                std::vector<uint64_t> l(k), l2(k);
                std::vector<float> d(k), d2(k);
                size_t c = 0, c2 = 0;
                uint64_t n = 0;
                while (!go.load()) std::this_thread::yield();
                while (!stop.load()) {
                    make_query(rng, q.data());
                    const auto a = Clock::now();
                    if (dawn_batcher_search(b, q.data(), k, l.data(), d.data(), &c) != DAWN_OK) {
                        errors++;
                        break;
                    }
                    lats[t].push_back(std::chrono::duration<double, std::milli>(Clock::now() - a).count());
                    n++;
                }
                for (int rep = 0; rep < 2 && t < 64; rep++) {  // parity with the unbatched call, bit for bit
                    make_query(rng, q.data());
                    if (dawn_batcher_search(b, q.data(), k, l.data(), d.data(), &c) != DAWN_OK) errors++;
                    direct(q.data(), l2.data(), d2.data(), &c2);
                    if (c != c2 || memcmp(l.data(), l2.data(), c * 8) || memcmp(d.data(), d2.data(), c * 4)) mismatches++;
                    checks++;
                }
            });
        }
        go = true;
        std::this_thread::sleep_for(std::chrono::milliseconds(300));  // warm-up
        uint64_t b0, q0, m0;
        dawn_batcher_stats(b, &b0, &q0, &m0);
        const auto t0 = Clock::now();
        std::this_thread::sleep_for(std::chrono::duration<double>(secs));
        uint64_t b1, q1, m1;
        dawn_batcher_stats(b, &b1, &q1, &m1);
        const double el = std::chrono::duration<double>(Clock::now() - t0).count();
        stop = true;
        for (auto &x : th) x.join();
        std::vector<double> all;
        for (auto &v : lats) all.insert(all.end(), v.begin(), v.end());
        printf("{\"front\": \"dawn_batcher_search\", \"gpus\": %d, \"rows\": %zu, \"k\": %zu, \"callers\": %d, \"max_batch\": %zu, "
               "\"max_wait_us\": %u, \"qps\": %.1f, \"latency_ms_p50\": %.3f, \"latency_ms_p99\": %.3f, \"mean_batch\": %.1f, "
               "\"largest_batch\": %llu, \"parity_checks\": %llu, \"parity_mismatches\": %llu, \"errors\": %llu}\n",
               gpus, rows, k, T, max_batch, max_wait_us, (q1 - q0) / el, pct(all, 0.5), pct(all, 0.99),
               (double)(q1 - q0) / std::max<uint64_t>(1, b1 - b0), (unsigned long long)m1,
               (unsigned long long)checks.load(), (unsigned long long)mismatches.load(), (unsigned long long)errors.load());
        fflush(stdout);
        dawn_batcher_free(b);
    }
    if (idx) dawn_index_free(idx);
    if (multi) dawn_multi_free(multi);
    return 0;
}
