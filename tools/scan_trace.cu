// Per-CTA timeline of the fp16 scan kernel on a small corpus (where fixed costs dominate).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DDAWN_SCAN_TRACE -I include \
//        tools/scan_trace.cu -o tools/bin/scan_trace
// Includes the kernel source directly so the library build stays free of trace code.
#include "../dawnsearch_b200/csrc/scan_topk.cu"
#include "../dawnsearch_b200/csrc/finalize.cu"

#include <algorithm>
#include <cstdio>
#include <vector>

int main(int argc, char **argv) {
    const uint32_t n = argc > 1 ? (uint32_t)atol(argv[1]) : 100000u;
    const int kp = argc > 2 ? atoi(argv[2]) : 32;
    using namespace dawn;
    __half *corpus;
    float *q;
    Cand *partials;
    uint32_t *ctr;
    cudaMalloc(&corpus, (size_t)n * kDim * 2);
    cudaMalloc(&q, kDim * 4);
    cudaMalloc(&partials, 148 * kMaxCand * sizeof(Cand));
    cudaMalloc(&ctr, 64);
    {   // unit-ish pseudo-random rows: values in +-0.09
        std::vector<__half> h((size_t)n * kDim);
        uint64_t s = 88172645463325252ull;
        for (auto &x : h) {
            s ^= s << 13; s ^= s >> 7; s ^= s << 17;
            x = __float2half(((int)(s % 2001) - 1000) * 9e-5f);
        }
        cudaMemcpy(corpus, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
        std::vector<float> hq(kDim);
        for (int i = 0; i < kDim; i++) hq[i] = __half2float(h[5 * kDim + i]);
        cudaMemcpy(q, hq.data(), kDim * 4, cudaMemcpyHostToDevice);
    }
    ScanLaunch p{};
    p.corpus = corpus; p.labels = nullptr; p.n_rows = n; p.queries = q; p.nq = 1; p.kprime = kp;
    p.partials = partials; p.chunk_counter = ctr; p.status = ctr + 1; p.grid = 148; p.score_floor = -INFINITY;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    static unsigned long long tr[148][12];
    for (int it = 0; it < 30; it++) {
        cudaMemset(ctr, 0, 64);
        unsigned long long zero[148][12] = {};
        cudaMemcpyToSymbol(g_scan_trace, zero, sizeof(zero));
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        cudaError_t e = launch_scan_topk_f16(p, 0);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        if (e != cudaSuccess || cudaGetLastError() != cudaSuccess) { printf("launch failed\n"); return 1; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = std::min(best, ms);
        if (it == 29) printf("rows %u kp %d: last %.2f us, best %.2f us\n", n, kp, ms * 1e3f, best * 1e3f);
    }
    cudaMemcpyFromSymbol(tr, g_scan_trace, sizeof(tr));
    unsigned long long t0 = ~0ull, tend = 0;
    for (int c = 0; c < 148; c++) { t0 = std::min(t0, tr[c][0]); tend = std::max(tend, tr[c][8]); }
    printf("kernel span (first CTA entry -> last CTA exit): %.2f us\n", (tend - t0) * 1e-3);
    const char *names[9] = {"entry", "setup done", "first chunk claimed", "last bulk issued", "consumer q loaded",
                            "first stage landed", "warp0 stream end", "final prune done", "exit"};
    for (int s = 0; s < 9; s++) {
        std::vector<double> v;
        for (int c = 0; c < 148; c++) v.push_back((tr[c][s] - t0) * 1e-3);
        std::sort(v.begin(), v.end());
        printf("  %-22s min %7.2f  med %7.2f  max %7.2f us\n", names[s], v[0], v[74], v[147]);
    }
    double pr = 0, np = 0, nf = 0;
    for (int c = 0; c < 148; c++) { pr += tr[c][9] * 1e-3; np += tr[c][10]; nf += tr[c][11]; }
    printf("  in-scan prunes per CTA %.2f, %.2f us each; end-of-stream prune rounds per CTA %.2f\n", np / 148,
           np ? pr / np : 0.0, nf / 148);

    // ---- finalize (K5 + K6) over the lists the scan just wrote
    uint64_t *d_labels; float *d_dist; uint32_t *d_cnt;
    cudaMalloc(&d_labels, 128 * 8); cudaMalloc(&d_dist, 128 * 4); cudaMalloc(&d_cnt, 64);
    FinalizeLaunch f{};
    f.corpus = corpus; f.queries = q; f.nq = 1; f.partials = partials; f.n_lists = 148; f.kprime = kp;
    f.k = kp >= 32 ? 20 : 10; f.eps = 3e-5f; f.labels_out = d_labels; f.distances_out = d_dist; f.counts_out = d_cnt;
    f.flags_out = d_cnt + 1; f.scalar = 0; f.eps_q = nullptr; f.overflow = nullptr; f.counters = ctr; f.n_counters = 2;
    f.status_out = d_cnt + 2;
    best = 1e9f;
    for (int it = 0; it < 30; it++) {
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        launch_finalize(f, 0);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = std::min(best, ms);
    }
    unsigned long long ft[16];
    cudaMemcpyFromSymbol(ft, g_fin_trace, sizeof(ft));
    printf("finalize: best %.2f us by events; in-kernel timeline (us from entry):\n", best * 1e3f);
    const char *fn[9] = {"entry", "probe loaded", "bound found", "filtered", "merged list ready", "rows staged",
                         "re-scored", "results written", "exit"};
    for (int s = 0; s < 9; s++) printf("  %-20s %7.2f\n", fn[s], (ft[s] - ft[0]) * 1e-3);
    return 0;
}
