// provider_demo.cpp -- drives dawn::SearchProvider (include/dawn_index.hpp) the way the reference's
// SearchService drives its SearchProvider: ExtractedPage inserts one at a time, then
// search_embedding per query (src/search/search_service.rs:60-181).  Used by tests/test_cpp_mirror.py.
//   provider_demo <rows.f32> <n_rows> <queries.f32> <n_queries>   -> prints "q page_id distance_bits" lines
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>

#include "dawn_index.hpp"

static std::vector<float> read_f32(const char *path, size_t n) {
    std::vector<float> v(n);
    std::ifstream f(path, std::ios::binary);
    f.read(reinterpret_cast<char *>(v.data()), (std::streamsize)(n * sizeof(float)));
    if (!f) throw std::runtime_error(std::string("short read: ") + path);
    return v;
}

int main(int argc, char **argv) {
    try {
        if (argc == 2 && !strcmp(argv[1], "--no-gpu-check")) {
            // hosts without a GPU: construction must fail loudly, never fall back to a CPU path
            try {
                dawn::SearchProvider p(".");
                std::puts("UNEXPECTED: provider constructed");
                return 2;
            } catch (const dawn::Error &e) {
                std::printf("refused code=%d msg=%s\n", e.code, e.what());
                return 0;
            }
        }
        if (argc == 3 && !strcmp(argv[1], "--host-only")) {
            // the host mirrors of vector.rs / best_results.rs need no GPU:
            //   provider_demo --host-only <vectors.f32 (3 x 384)>  -> prints the i24 bytes (hex), the decoded bits,
            //   the normalised bits and a BestResults run
            auto v = read_f32(argv[2], 3 * dawn::EM_LEN);
            for (int i = 0; i < 3; i++) {
                std::vector<float> x(v.begin() + i * dawn::EM_LEN, v.begin() + (i + 1) * dawn::EM_LEN);
                auto wire = dawn::to24(x);
                std::printf("wire %d ", i);
                for (auto b : wire) std::printf("%02x", b);
                std::printf("\n");
                auto back = dawn::from24(wire);
                std::printf("back %d", i);
                for (float f : back) { uint32_t u; std::memcpy(&u, &f, 4); std::printf(" %u", u); }
                std::printf("\n");
                std::printf("norm %d %d\n", i, dawn::is_normalized(x.data()) ? 1 : 0);
            }
            dawn::BestResults<float> best(3);  // best_results.rs:44-79
            const float d[] = {0.5f, 0.25f, 0.75f, 0.125f, 0.25f, 0.9f};
            for (std::size_t i = 0; i < 6; i++) best.insert(dawn::NodeReference<float>{i == 4 ? 1 : i, d[i]});
            best.sort();
            std::printf("best");
            for (auto &r : best.results()) std::printf(" %zu:%g", r.id, r.distance);
            std::printf(" worst=%g\n", best.worst_distance());
            return 0;
        }
        if (argc != 5) return 64;
        const size_t n = std::stoul(argv[2]), nq = std::stoul(argv[4]);
        auto rows = read_f32(argv[1], n * dawn::EM_LEN);
        auto qs = read_f32(argv[3], nq * dawn::EM_LEN);
        dawn::SearchProvider provider(".");
        for (size_t i = 0; i < n; i++) {
            dawn::ExtractedPage page{"https://example.org/" + std::to_string(i), "title " + std::to_string(i), "text"};
            provider.insert(page, std::vector<float>(rows.begin() + i * dawn::EM_LEN, rows.begin() + (i + 1) * dawn::EM_LEN));
        }
        // inserting the same URL again is a no-op (search_provider.rs:254-263)
        provider.insert(dawn::ExtractedPage{"https://example.org/0", "dup", "dup"},
                        std::vector<float>(rows.begin(), rows.begin() + dawn::EM_LEN));
        if (provider.stats().pages_indexed != n) return 3;
        for (size_t q = 0; q < nq; q++) {
            dawn::SearchResult r = provider.search_embedding(std::vector<float>(qs.begin() + q * dawn::EM_LEN, qs.begin() + (q + 1) * dawn::EM_LEN));
            if (r.pages_searched != n) return 4;
            for (auto &p : r.pages) {
                uint32_t bits;
                memcpy(&bits, &p.distance, 4);
                std::printf("%zu %zu %u %s\n", q, p.page_id, bits, p.url.c_str());
            }
        }
        // more-like-this: a page queried with its own stored embedding returns itself first (web.rs:339)
        dawn::SearchResult like = provider.search_like(5);
        std::printf("like %zu %d\n", like.pages[0].page_id, like.pages[0].distance < 0.001f);
        // un-normalised input is rejected at the provider, as in the reference (:206)
        try {
            provider.search_embedding(std::vector<float>(dawn::EM_LEN, 1.0f));
            return 5;
        } catch (const dawn::Error &) {
        }
        // BestResults keeps the reference's semantics (best_results.rs:44-65)
        dawn::BestResults<float> best(2);
        if (best.worst_distance() != 0.0f) return 6;
        best.insert({1, 0.5f});
        best.insert({2, 0.7f});
        if (best.insert({3, 0.7f}) || !best.insert({3, 0.6f}) || best.insert({1, 0.1f})) return 7;
        best.sort();
        if (best.results()[0].id != 1 || best.results()[1].id != 3) return 8;
        return 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
}
