// Per-CTA timeline of the fp16 scan kernel on a small corpus (where fixed costs dominate).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DDAWN_SCAN_TRACE -I include \
//        tools/scan_trace.cu -o tools/bin/scan_trace
// Includes the kernel source directly so the library build stays free of trace code.
#include "../dawnsearch_b200/csrc/scan_topk.cu"

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
    return 0;
}
