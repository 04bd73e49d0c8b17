// dawn_index.cu -- the C ABI (include/dawn_index.h) and the host-side index object:
// corpus arena in HBM, label table, pinned staging for adds, search workspace, streams.
//
// This object takes the place of usearch's `Index` behind the reference's SearchProvider
// (/root/reference/src/search/search_provider.rs:67,102).  No CPU fallback exists: every
// compute path launches the kernels in scan_topk.cu / finalize.cu / ingest.cu.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/dawn_index.h"
#include "dawn_common.cuh"

using namespace dawn;

namespace {

thread_local std::string g_last_error;

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

constexpr size_t kStageRowsHost = 8192;  // pinned staging: 8192 vectors = 12 MB of f32
constexpr float kScanEps = 3.0e-5f;      // bound on |scan score - exact score|

// merge kernel for sharded searches, defined at the bottom of this file
cudaError_t launch_merge_results(const uint64_t *labels, const float *dist, const uint32_t *counts,
                                 int n_lists, size_t list_stride_bytes, int batch, int k, uint64_t *labels_out,
                                 float *dist_out, uint32_t *counts_out, cudaStream_t s);
cudaError_t launch_find_label(const uint64_t *labels, size_t n, uint64_t label, uint32_t *row_out,
                              cudaStream_t s);

struct EventPair {
    cudaEvent_t a, b;
    int kind;  // 0 scan, 1 finalize, 2 gemm (all rounds of one batch)
};
constexpr float kGemmAccumSlack = 6.0e-5f;  // tensor-core f32 accumulation over K=384

}  // namespace

struct dawn_index {
    std::mutex mu;
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    bool dead = false;  // sticky CUDA failure

    int scalar = DAWN_SCALAR_F16;  // storage of the corpus: fp16 rows or the blocked int8 arena
    __half *corpus = nullptr;      // fp16: [phys][384]; int8: the same pointer holds the blocked arena
    uint64_t *labels = nullptr;
    size_t size = 0;      // rows committed to the device
    size_t capacity = 0;  // logical capacity promised to the caller
    size_t phys = 0;      // rows actually allocated

    // staged adds (host, pinned) not yet on the device
    // Two buffers: while the GPU copies / converts one, the host fills the other (pipelined
    // bulk load).  h_stage / h_stage_labels / d_stage alias the buffer being filled.
    float *h_stage_buf[2] = {nullptr, nullptr};
    uint64_t *h_labels_buf[2] = {nullptr, nullptr};
    float *d_stage_buf[2] = {nullptr, nullptr};
    cudaEvent_t stage_done[2] = {nullptr, nullptr};
    bool stage_busy[2] = {false, false};
    int stage_cur = 0;
    float *h_stage = nullptr;
    uint64_t *h_stage_labels = nullptr;
    size_t staged = 0;
    float *d_stage = nullptr;

    // search workspace
    size_t q_cap = 0;  // queries
    float *d_queries = nullptr, *h_queries = nullptr;
    // results of the host API: ONE packed block (labels | distances | counts | flags | status) so that a
    // search costs a single D2H copy
    uint8_t *d_result = nullptr, *h_result = nullptr;
    size_t result_cap = 0;
    float limit_score = -INFINITY;  // 1 - distance_limit of the search being enqueued (scan paths push it down)
    bool counters_clean = false;  // finalize leaves the chunk counters / status word zeroed for the next search
    Cand *d_partials = nullptr;
    size_t partials_cap = 0;
    uint32_t *d_counters = nullptr;  // one chunk counter per scan pass, + status word at [0]
    size_t counters_cap = 0;
    uint32_t *h_word = nullptr;  // pinned scratch word
    void *d_gemm_ws = nullptr;   // K3 workspace (fp16 queries, eps, thresholds, candidate logs)
    size_t gemm_ws_cap = 0;
    // path selection: batches >= gemm_min_batch over >= gemm_min_rows rows take the tensor-core path
    int64_t gemm_min_batch = 16;
    int64_t gemm_min_rows = 65536;
    int64_t gemm_small_batch = 3;          // from this batch size on, big corpora also take the tensor path
    int64_t gemm_small_batch_rows = 2000000;
    int64_t force_path = 0;  // 0 auto, 1 scan only, 2 gemm whenever possible
    int64_t gemm_cta_group = 0;  // 0 auto, 1 = one CTA per tile, 2 = CTA pairs
    int64_t gemm_chunk_tiles = 0;  // 0 auto
    int64_t gemm_sequential_tiles = 0;
    int64_t gemm_growth = 0;  // 0 = automatic
    // int8 corpora: batches of at least this many queries go through the fp16 tensor-core tiles chunk by chunk
    // (i8_tensor.cu) instead of ceil(B/2) scan passes (16 queries = 8 passes of ~3.9 ms over a 62.5M-row shard, against
    // ~58 ms for the whole shard on the tensor path whatever the batch).  0 = never.
    int64_t i8_tensor_min_batch = 16;
    int64_t i8_tensor_chunk_rows = 4 << 20;
    __half *d_i8_scratch = nullptr;   // one dequantised chunk
    size_t i8_scratch_rows = 0;
    Cand *d_i8_lists = nullptr;       // [batch][n_chunks][k'] gathered candidate lists
    size_t i8_lists_cap = 0;
    uint32_t *d_i8_overflow = nullptr;  // [batch] log overflow in any chunk
    size_t i8_overflow_cap = 0;

    // The search workspace (partials, counters, K3 logs) is shared by all searches on this handle:
    // a search enqueued on another stream than the previous one first waits for it.
    cudaEvent_t ws_done = nullptr;
    cudaStream_t ws_stream = nullptr;
    bool ws_used = false;

    bool profiling = false;
    std::vector<EventPair> pending;
    std::vector<EventPair> free_events;
    dawn_profile prof{};
};

namespace {

#define CK(idx, expr)                                                                         \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            (idx)->dead = true;                                                               \
            return fail(DAWN_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                        __FILE__, __LINE__);                                                  \
        }                                                                                     \
    } while (0)

int check_alive(dawn_index *idx) {
    if (!idx) return fail(DAWN_ERR_INVALID, "null index handle");
    if (idx->dead) return fail(DAWN_ERR_CUDA, "index is in a failed state after an earlier CUDA error");
    cudaError_t e = cudaSetDevice(idx->device);
    if (e != cudaSuccess) {
        idx->dead = true;
        return fail(DAWN_ERR_CUDA, "cudaSetDevice(%d): %s", idx->device, cudaGetErrorString(e));
    }
    return DAWN_OK;
}

void drain_events(dawn_index *idx) {
    if (idx->pending.empty()) return;
    cudaStreamSynchronize(idx->stream);
    for (auto &p : idx->pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
            if (p.kind == 0) idx->prof.scan_ms += ms;
            else if (p.kind == 1) idx->prof.finalize_ms += ms;
            else idx->prof.gemm_ms += ms;
        }
        idx->free_events.push_back(p);
    }
    idx->pending.clear();
}

bool begin_event(dawn_index *idx, int kind, cudaStream_t s, EventPair *out) {
    if (!idx->profiling) return false;
    EventPair p;
    if (!idx->free_events.empty()) {
        p = idx->free_events.back();
        idx->free_events.pop_back();
    } else {
        if (cudaEventCreate(&p.a) != cudaSuccess || cudaEventCreate(&p.b) != cudaSuccess) return false;
    }
    p.kind = kind;
    cudaEventRecord(p.a, s);
    *out = p;
    return true;
}

void end_event(dawn_index *idx, EventPair &p, cudaStream_t s) {
    cudaEventRecord(p.b, s);
    idx->pending.push_back(p);
}

inline size_t arena_bytes(const dawn_index *idx, size_t rows) {
    return idx->scalar == DAWN_SCALAR_I8 ? i8_arena_bytes(rows) : rows * (size_t)kRowBytesF16;
}
inline uint8_t *arena_i8(const dawn_index *idx) { return reinterpret_cast<uint8_t *>(idx->corpus); }

int grow_physical(dawn_index *idx, size_t rows) {
    if (rows <= idx->phys) return DAWN_OK;
    if (rows > 0xFFFFFFF0ull) return fail(DAWN_ERR_INVALID, "capacity %zu exceeds 2^32 rows per GPU", rows);
    size_t want = rows;
    if (idx->phys > 0) {  // amortise the reference's reserve(size + 1024) pattern
        size_t geo = idx->phys + idx->phys / 2;
        if (geo > want) want = geo;
    }
    __half *nc = nullptr;
    uint64_t *nl = nullptr;
    cudaError_t e = cudaMalloc(&nc, arena_bytes(idx, want));
    if (e != cudaSuccess && want > rows) {
        cudaGetLastError();
        want = rows;
        e = cudaMalloc(&nc, arena_bytes(idx, want));
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(DAWN_ERR_CAPACITY, "cannot allocate %zu bytes of HBM for %zu vectors: %s",
                    arena_bytes(idx, want), want, cudaGetErrorString(e));
    }
    e = cudaMalloc(&nl, want * sizeof(uint64_t));
    if (e != cudaSuccess) {
        cudaGetLastError();
        cudaFree(nc);
        return fail(DAWN_ERR_CAPACITY, "cannot allocate label table for %zu vectors", want);
    }
    if (idx->size > 0) {
        CK(idx, cudaMemcpyAsync(nc, idx->corpus, arena_bytes(idx, idx->size), cudaMemcpyDeviceToDevice, idx->stream));
        CK(idx, cudaMemcpyAsync(nl, idx->labels, idx->size * sizeof(uint64_t), cudaMemcpyDeviceToDevice, idx->stream));
        CK(idx, cudaStreamSynchronize(idx->stream));
    }
    if (idx->corpus) cudaFree(idx->corpus);
    if (idx->labels) cudaFree(idx->labels);
    idx->corpus = nc;
    idx->labels = nl;
    idx->phys = want;
    return DAWN_OK;
}

// Enqueue the current staging buffer (H2D of the f32 rows, K1 convert, labels) WITHOUT waiting, and
// switch to the other buffer (waiting only if that one is still in flight).
int flush_staged_async(dawn_index *idx) {
    if (idx->staged == 0) return DAWN_OK;
    const size_t n = idx->staged;
    const int cur = idx->stage_cur;
    CK(idx, cudaMemcpyAsync(idx->d_stage, idx->h_stage, n * kDim * sizeof(float), cudaMemcpyHostToDevice, idx->stream));
    if (idx->scalar == DAWN_SCALAR_I8) CK(idx, launch_ingest_i8(idx->d_stage, arena_i8(idx), idx->size, n, idx->stream));
    else CK(idx, launch_ingest_f16(idx->d_stage, idx->corpus + idx->size * kDim, n, idx->stream));
    idx->prof.kernel_launches++;
    CK(idx, cudaMemcpyAsync(idx->labels + idx->size, idx->h_stage_labels, n * sizeof(uint64_t), cudaMemcpyHostToDevice, idx->stream));
    CK(idx, cudaEventRecord(idx->stage_done[cur], idx->stream));
    idx->stage_busy[cur] = true;
    idx->size += n;
    idx->staged = 0;
    const int nxt = cur ^ 1;
    if (idx->stage_busy[nxt]) {
        CK(idx, cudaEventSynchronize(idx->stage_done[nxt]));
        idx->stage_busy[nxt] = false;
    }
    idx->stage_cur = nxt;
    idx->h_stage = idx->h_stage_buf[nxt];
    idx->h_stage_labels = idx->h_labels_buf[nxt];
    idx->d_stage = idx->d_stage_buf[nxt];
    return DAWN_OK;
}

// Make every staged / in-flight add visible: flush what is staged and wait for the stream.
int flush_staged(dawn_index *idx) {
    if (idx->staged == 0 && !idx->stage_busy[0] && !idx->stage_busy[1]) return DAWN_OK;
    int rc = flush_staged_async(idx);
    if (rc) return rc;
    CK(idx, cudaStreamSynchronize(idx->stream));
    idx->stage_busy[0] = idx->stage_busy[1] = false;
    return DAWN_OK;
}

// DAWN_DEBUG_STAGES=1: wait for each kernel of a search separately (polling, 5 s) and say on stderr which one
// did not finish.  Debug aid only; never set in tests or benches.
bool debug_stages() {
    static const bool on = getenv("DAWN_DEBUG_STAGES") != nullptr;
    return on;
}
void debug_wait(dawn_index *idx, cudaStream_t s, const char *stage) {
    if (!debug_stages()) return;
    const auto t0 = std::chrono::steady_clock::now();
    while (true) {
        cudaError_t e = cudaStreamQuery(s);
        if (e == cudaSuccess) {
            fprintf(stderr, "[dawn debug] %s done\n", stage);
            return;
        }
        if (e != cudaErrorNotReady) {
            fprintf(stderr, "[dawn debug] %s: %s\n", stage, cudaGetErrorString(e));
            return;
        }
        if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 5.0) {
            uint32_t c[4] = {0, 0, 0, 0};
            cudaStream_t side;
            cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking);
            if (idx->d_counters) {
                cudaMemcpyAsync(c, idx->d_counters, sizeof(c), cudaMemcpyDeviceToHost, side);
                cudaStreamSynchronize(side);
            }
            fprintf(stderr, "[dawn debug] %s STILL RUNNING after 5 s (size %zu, counters %u %u %u %u)\n", stage, idx->size, c[0],
                    c[1], c[2], c[3]);
            fflush(stderr);
            return;
        }
    }
}

int choose_kprime(size_t k) {
    size_t want = k + (k / 8 > 4 ? k / 8 : 4);
    int kp = 16;
    while ((size_t)kp < want) kp <<= 1;
    return kp > kMaxCand ? kMaxCand : kp;
}

struct ResultView {
    uint64_t *labels;
    float *dist;
    uint32_t *counts, *flags, *status;
    size_t bytes;
};
constexpr size_t kDirectResultBytes = 4096;  // result blocks up to this size are written to host memory by the kernel

inline ResultView result_view(uint8_t *base, size_t batch, size_t k) {
    ResultView v;
    v.labels = reinterpret_cast<uint64_t *>(base);
    v.dist = reinterpret_cast<float *>(base + batch * k * 8);
    v.counts = reinterpret_cast<uint32_t *>(base + batch * k * 12);
    v.flags = v.counts + batch;
    v.status = v.flags + batch;
    v.bytes = batch * k * 12 + batch * 8 + 16;
    return v;
}

int ensure_query_ws(dawn_index *idx, size_t batch, size_t k) {
    if (batch > idx->q_cap) {
        size_t cap = batch < 64 ? 64 : batch;
        if (idx->d_queries) cudaFree(idx->d_queries);
        if (idx->h_queries) cudaFreeHost(idx->h_queries);
        idx->q_cap = 0;
        CK(idx, cudaMalloc(&idx->d_queries, cap * kDim * sizeof(float)));
        CK(idx, cudaMallocHost(&idx->h_queries, cap * kDim * sizeof(float)));
        idx->q_cap = cap;
    }
    const size_t need = batch * (k ? k : 1) * 12 + batch * 8 + 16;
    if (need > idx->result_cap) {
        size_t cap = need < 65536 ? 65536 : need;
        if (idx->d_result) cudaFree(idx->d_result);
        if (idx->h_result) cudaFreeHost(idx->h_result);
        idx->result_cap = 0;
        CK(idx, cudaMalloc(&idx->d_result, cap));
        CK(idx, cudaMallocHost(&idx->h_result, cap));
        idx->result_cap = cap;
    }
    return DAWN_OK;
}

// Zero the chunk counters / status word unless the previous search's finalize already did.
int prepare_counters(dawn_index *idx, size_t need, cudaStream_t s) {
    if (need > idx->counters_cap) {
        size_t cap = need < 1024 ? 1024 : need;
        if (idx->d_counters) cudaFree(idx->d_counters);
        idx->counters_cap = 0;
        CK(idx, cudaMalloc(&idx->d_counters, cap * sizeof(uint32_t)));
        idx->counters_cap = cap;
        idx->counters_clean = false;
    }
    if (!idx->counters_clean) CK(idx, cudaMemsetAsync(idx->d_counters, 0, idx->counters_cap * sizeof(uint32_t), s));
    idx->counters_clean = false;  // dirty until this search's finalize has been enqueued
    return DAWN_OK;
}

// int8 corpus, large batch: chunk by chunk through the fp16 tensor-core tiles (i8_tensor.cu), then one finalize over
// the gathered per-chunk lists with the exact int8 re-score.
constexpr float kI8DequantSlack = 5.6e-4f;  // |q.(fp16(s x8) - s x8)| <= 2^-11 ||q|| ||s x8||, plus a larger ||x~|| in the query-rounding term
int search_i8_tensor(dawn_index *idx, const float *d_queries, size_t batch, size_t k, int kprime, uint64_t *d_labels_out,
                     float *d_dist_out, uint32_t *d_counts, uint32_t *d_flags, cudaStream_t s, uint32_t *d_status_out) {
    const int grid = idx->sm_count;
    const int kp = kprime < 64 ? 64 : kprime;  // more slack: the certificate has to absorb the dequantisation rounding
    const size_t n = idx->size;
    const size_t want = (size_t)idx->i8_tensor_chunk_rows;
    const size_t n_chunks = (n + want - 1) / want;
    const size_t rpc = ((n + n_chunks - 1) / n_chunks + 255) / 256 * 256;  // rows per chunk, whole tiles
    const size_t qp = (batch + 255) / 256 * 256;
    if (rpc > idx->i8_scratch_rows) {
        if (idx->d_i8_scratch) cudaFree(idx->d_i8_scratch);
        idx->i8_scratch_rows = 0;
        CK(idx, cudaMalloc(&idx->d_i8_scratch, rpc * (size_t)kRowBytesF16));
        idx->i8_scratch_rows = rpc;
    }
    if (batch * n_chunks * kp > idx->i8_lists_cap) {
        if (idx->d_i8_lists) cudaFree(idx->d_i8_lists);
        idx->i8_lists_cap = 0;
        CK(idx, cudaMalloc(&idx->d_i8_lists, batch * n_chunks * kp * sizeof(Cand)));
        idx->i8_lists_cap = batch * n_chunks * kp;
    }
    if (batch > idx->i8_overflow_cap) {
        if (idx->d_i8_overflow) cudaFree(idx->d_i8_overflow);
        idx->i8_overflow_cap = 0;
        CK(idx, cudaMalloc(&idx->d_i8_overflow, batch * sizeof(uint32_t)));
        idx->i8_overflow_cap = batch;
    }
    const size_t need_ws = gemm_workspace_bytes((int)batch);
    if (need_ws > idx->gemm_ws_cap) {
        if (idx->d_gemm_ws) cudaFree(idx->d_gemm_ws);
        idx->gemm_ws_cap = 0;
        CK(idx, cudaMalloc(&idx->d_gemm_ws, need_ws));
        idx->gemm_ws_cap = need_ws;
    }
    if (qp * kp > idx->partials_cap) {
        if (idx->d_partials) cudaFree(idx->d_partials);
        idx->partials_cap = 0;
        CK(idx, cudaMalloc(&idx->d_partials, qp * kp * sizeof(Cand)));
        idx->partials_cap = qp * kp;
    }
    {
        int prc = prepare_counters(idx, 1, s);
        if (prc) return prc;
    }
    CK(idx, cudaMemsetAsync(idx->d_i8_overflow, 0, batch * sizeof(uint32_t), s));
    const float *eps_q = nullptr;
    EventPair evg;
    bool timedg = begin_event(idx, 2, s, &evg);
    for (size_t c = 0; c < n_chunks; c++) {
        const size_t base = c * rpc;
        const size_t rows = n - base < rpc ? n - base : rpc;
        CK(idx, launch_dequant_i8_f16(arena_i8(idx), base, rows, idx->d_i8_scratch, s));
        GemmSearch gs;
        gs.corpus = idx->d_i8_scratch;
        gs.labels = idx->labels + base;
        gs.n_rows = rows;
        gs.queries = d_queries;
        gs.n_queries = (int)batch;
        gs.kprime = kp;
        gs.grid = grid;
        gs.cta_group = (int)idx->gemm_cta_group;
        gs.chunk_tiles = (int)idx->gemm_chunk_tiles;
        gs.sequential_tiles = (int)idx->gemm_sequential_tiles;
        gs.growth = (int)idx->gemm_growth;
        gs.workspace = idx->d_gemm_ws;
        gs.final_lists = idx->d_partials;
        gs.accum_slack = kGemmAccumSlack + kI8DequantSlack;
        const uint32_t *overflow = nullptr;
        int launches = 0;
        gs.eps_out = &eps_q;
        gs.overflow_out = &overflow;
        gs.launches_out = &launches;
        CK(idx, launch_gemm_search(gs, s));
        CK(idx, launch_gather_chunk_lists(idx->d_partials, (int)batch, kp, (uint32_t)base, (int)c, (int)n_chunks, idx->d_i8_lists,
                                          overflow, idx->d_i8_overflow, s));
        idx->prof.gemm_batches++;
        idx->prof.kernel_launches += launches + 2;
    }
    if (timedg) end_event(idx, evg, s);
    FinalizeLaunch fl;
    fl.corpus = idx->corpus;
    fl.queries = d_queries;
    fl.nq = (int)batch;
    fl.partials = idx->d_i8_lists;
    fl.n_lists = (int)n_chunks;
    fl.kprime = kp;
    fl.k = (int)k;
    fl.eps = 0.f;
    fl.labels_out = d_labels_out;
    fl.distances_out = d_dist_out;
    fl.counts_out = d_counts;
    fl.flags_out = d_flags;
    fl.scalar = 1;
    fl.eps_q = eps_q;  // same queries and slack for every chunk: the last chunk's values are everybody's
    fl.overflow = idx->d_i8_overflow;
    fl.counters = idx->d_counters;
    fl.n_counters = 1;
    fl.status_out = d_status_out;
    EventPair evf;
    bool timedf = begin_event(idx, 1, s, &evf);
    CK(idx, launch_finalize(fl, s));
    if (timedf) end_event(idx, evf, s);
    idx->prof.finalize_launches++;
    idx->prof.kernel_launches++;
    idx->prof.queries += batch;
    if (idx->pending.size() > 4096) drain_events(idx);
    return DAWN_OK;
}

// Enqueue the whole search for `batch` device-resident queries on stream `s`.
int search_enqueue_impl(dawn_index *idx, const float *d_queries, size_t batch, size_t k, int kprime,
                        uint64_t *d_labels_out, float *d_dist_out, uint32_t *d_counts, uint32_t *d_flags,
                        cudaStream_t s, bool scan_only, uint32_t *d_status_out);

int search_enqueue(dawn_index *idx, const float *d_queries, size_t batch, size_t k, int kprime,
                   uint64_t *d_labels_out, float *d_dist_out, uint32_t *d_counts, uint32_t *d_flags,
                   cudaStream_t s, bool scan_only = false, uint32_t *d_status_out = nullptr) {
    if (idx->ws_used && idx->ws_stream != s) CK(idx, cudaStreamWaitEvent(s, idx->ws_done, 0));
    int rc = search_enqueue_impl(idx, d_queries, batch, k, kprime, d_labels_out, d_dist_out, d_counts, d_flags, s, scan_only,
                                 d_status_out);
    if (rc == DAWN_OK) {
        CK(idx, cudaEventRecord(idx->ws_done, s));
        idx->ws_stream = s;
        idx->ws_used = true;
        idx->counters_clean = true;  // every path ends with a finalize that zeroes the counters it used
    }
    return rc;
}

int search_enqueue_impl(dawn_index *idx, const float *d_queries, size_t batch, size_t k, int kprime,
                        uint64_t *d_labels_out, float *d_dist_out, uint32_t *d_counts, uint32_t *d_flags,
                        cudaStream_t s, bool scan_only, uint32_t *d_status_out) {
    const int grid = idx->sm_count;
    if (idx->scalar == DAWN_SCALAR_I8 && !scan_only && idx->i8_tensor_min_batch > 0 &&
        (int64_t)batch >= idx->i8_tensor_min_batch && idx->size >= 65536 && k <= 100)
        return search_i8_tensor(idx, d_queries, batch, k, kprime, d_labels_out, d_dist_out, d_counts, d_flags, s, d_status_out);
    if (idx->scalar == DAWN_SCALAR_I8) {
        // K4: int8 storage -> streaming dp4a scan, 1 or 2 queries per pass, exact f32 re-score
        const size_t need_ws = batch * (sizeof(I8Query) + sizeof(float)) + 256;
        if (need_ws > idx->gemm_ws_cap) {
            if (idx->d_gemm_ws) cudaFree(idx->d_gemm_ws);
            idx->gemm_ws_cap = 0;
            CK(idx, cudaMalloc(&idx->d_gemm_ws, need_ws));
            idx->gemm_ws_cap = need_ws;
        }
        I8Query *d_iq = static_cast<I8Query *>(idx->d_gemm_ws);
        float *d_eps = reinterpret_cast<float *>(static_cast<uint8_t *>(idx->d_gemm_ws) + batch * sizeof(I8Query));
        const size_t need_partials = batch * (size_t)grid * kprime;
        if (need_partials > idx->partials_cap) {
            if (idx->d_partials) cudaFree(idx->d_partials);
            idx->partials_cap = 0;
            CK(idx, cudaMalloc(&idx->d_partials, need_partials * sizeof(Cand)));
            idx->partials_cap = need_partials;
        }
        const size_t need_counters = batch + 1;
        {
            int prc = prepare_counters(idx, need_counters, s);
            if (prc) return prc;
        }
        CK(idx, launch_prep_queries_i8(d_queries, (int)batch, d_iq, d_eps, s));
        idx->prof.kernel_launches++;
        size_t done = 0, pass = 0;
        while (done < batch) {
            const int qt = batch - done >= 2 ? 2 : 1;
            ScanLaunchI8 sl;
            sl.corpus = arena_i8(idx);
            sl.labels = idx->labels;
            sl.n_rows = (uint32_t)idx->size;
            sl.queries = d_iq + done;
            sl.nq = qt;
            sl.kprime = kprime;
            sl.partials = idx->d_partials + done * (size_t)grid * kprime;
            sl.chunk_counter = idx->d_counters + 1 + pass;
            sl.status = idx->d_counters;
            sl.grid = grid;
            sl.eps_q = d_eps + done;
            sl.limit_score = idx->limit_score;
            EventPair ev;
            bool timed = begin_event(idx, 0, s, &ev);
            CK(idx, launch_scan_topk_i8(sl, s));
            if (timed) end_event(idx, ev, s);
            idx->prof.scan_launches++;
            idx->prof.kernel_launches++;
            done += qt;
            pass++;
        }
        FinalizeLaunch fl;
        fl.corpus = idx->corpus;
        fl.queries = d_queries;
        fl.nq = (int)batch;
        fl.partials = idx->d_partials;
        fl.n_lists = grid;
        fl.kprime = kprime;
        fl.k = (int)k;
        fl.eps = 0.f;
        fl.labels_out = d_labels_out;
        fl.distances_out = d_dist_out;
        fl.counts_out = d_counts;
        fl.flags_out = d_flags;
        fl.scalar = 1;
        fl.eps_q = d_eps;
        fl.overflow = nullptr;
        fl.counters = idx->d_counters;
        fl.n_counters = (int)need_counters;
        fl.status_out = d_status_out;
        EventPair ev;
        bool timed = begin_event(idx, 1, s, &ev);
        CK(idx, launch_finalize(fl, s));
        if (timed) end_event(idx, ev, s);
        idx->prof.finalize_launches++;
        idx->prof.kernel_launches++;
        idx->prof.queries += batch;
        if (idx->pending.size() > 4096) drain_events(idx);
        return DAWN_OK;
    }
    const bool gemm_ok = idx->size >= 1024 && batch >= 1;
    const bool use_gemm = gemm_ok && !scan_only && idx->force_path != 1 &&
                          (idx->force_path == 2 ||
                           ((int64_t)batch >= idx->gemm_min_batch && (int64_t)idx->size >= idx->gemm_min_rows) ||
                           ((int64_t)batch >= idx->gemm_small_batch && (int64_t)idx->size >= idx->gemm_small_batch_rows));
    if (use_gemm) {
        const size_t qp = (batch + 255) / 256 * 256;
        const size_t need_ws = gemm_workspace_bytes((int)batch);
        if (need_ws > idx->gemm_ws_cap) {
            if (idx->d_gemm_ws) cudaFree(idx->d_gemm_ws);
            idx->gemm_ws_cap = 0;
            CK(idx, cudaMalloc(&idx->d_gemm_ws, need_ws));
            idx->gemm_ws_cap = need_ws;
        }
        if (qp * kprime > idx->partials_cap) {
            if (idx->d_partials) cudaFree(idx->d_partials);
            idx->partials_cap = 0;
            CK(idx, cudaMalloc(&idx->d_partials, qp * kprime * sizeof(Cand)));
            idx->partials_cap = qp * kprime;
        }
        {
            int prc = prepare_counters(idx, 1, s);
            if (prc) return prc;
        }
        GemmSearch gs;
        gs.corpus = idx->corpus;
        gs.labels = idx->labels;
        gs.n_rows = idx->size;
        gs.queries = d_queries;
        gs.n_queries = (int)batch;
        gs.kprime = kprime;
        gs.grid = grid;
        gs.cta_group = (int)idx->gemm_cta_group;
        gs.chunk_tiles = (int)idx->gemm_chunk_tiles;
        gs.sequential_tiles = (int)idx->gemm_sequential_tiles;
        gs.growth = (int)idx->gemm_growth;
        gs.workspace = idx->d_gemm_ws;
        gs.final_lists = idx->d_partials;
        gs.accum_slack = kGemmAccumSlack;
        const float *eps_q = nullptr;
        const uint32_t *overflow = nullptr;
        int launches = 0;
        gs.eps_out = &eps_q;
        gs.overflow_out = &overflow;
        gs.launches_out = &launches;
        EventPair evg;
        bool timedg = begin_event(idx, 2, s, &evg);
        CK(idx, launch_gemm_search(gs, s));
        if (timedg) end_event(idx, evg, s);
        idx->prof.gemm_batches++;
        idx->prof.kernel_launches += launches;
        FinalizeLaunch fl;
        fl.corpus = idx->corpus;
        fl.queries = d_queries;
        fl.nq = (int)batch;
        fl.partials = idx->d_partials;
        fl.n_lists = 1;
        fl.kprime = kprime;
        fl.k = (int)k;
        fl.eps = 0.f;
        fl.labels_out = d_labels_out;
        fl.distances_out = d_dist_out;
        fl.counts_out = d_counts;
        fl.flags_out = d_flags;
        fl.scalar = 0;
        fl.eps_q = eps_q;
        fl.overflow = overflow;
        fl.counters = idx->d_counters;
        fl.n_counters = 1;
        fl.status_out = d_status_out;
        EventPair evf;
        bool timedf = begin_event(idx, 1, s, &evf);
        CK(idx, launch_finalize(fl, s));
        if (timedf) end_event(idx, evf, s);
        idx->prof.finalize_launches++;
        idx->prof.kernel_launches++;
        idx->prof.queries += batch;
        if (idx->pending.size() > 4096) drain_events(idx);
        return DAWN_OK;
    }
    const size_t need_partials = batch * (size_t)grid * kprime;
    if (need_partials > idx->partials_cap) {
        if (idx->d_partials) cudaFree(idx->d_partials);
        idx->partials_cap = 0;
        CK(idx, cudaMalloc(&idx->d_partials, need_partials * sizeof(Cand)));
        idx->partials_cap = need_partials;
    }
    const size_t need_counters = batch + 1;
    {
        int prc = prepare_counters(idx, need_counters, s);
        if (prc) return prc;
    }

    size_t done = 0;
    size_t pass = 0;
    const int max_qt = scan_max_queries_per_pass(kprime);
    while (done < batch) {
        size_t left = batch - done;
        int qt = left >= 4 && max_qt >= 4 ? 4 : (left >= 2 && max_qt >= 2 ? 2 : 1);
        ScanLaunch sl;
        sl.corpus = idx->corpus;
        sl.labels = idx->labels;
        sl.n_rows = (uint32_t)idx->size;
        sl.queries = d_queries + done * kDim;
        sl.nq = qt;
        sl.kprime = kprime;
        sl.partials = idx->d_partials + done * (size_t)grid * kprime;
        sl.chunk_counter = idx->d_counters + 1 + pass;
        sl.status = idx->d_counters;
        sl.grid = grid;
        sl.score_floor = idx->limit_score > -INFINITY ? idx->limit_score - 2.0f * kScanEps - 1e-6f : -INFINITY;
        EventPair ev;
        bool timed = begin_event(idx, 0, s, &ev);
        CK(idx, launch_scan_topk_f16(sl, s));
        if (timed) end_event(idx, ev, s);
        debug_wait(idx, s, "scan_topk_f16");
        idx->prof.scan_launches++;
        idx->prof.kernel_launches++;
        done += qt;
        pass++;
    }
    FinalizeLaunch fl;
    fl.corpus = idx->corpus;
    fl.queries = d_queries;
    fl.nq = (int)batch;
    fl.partials = idx->d_partials;
    fl.n_lists = grid;
    fl.kprime = kprime;
    fl.k = (int)k;
    fl.eps = kScanEps;
    fl.labels_out = d_labels_out;
    fl.distances_out = d_dist_out;
    fl.counts_out = d_counts;
    fl.flags_out = d_flags;
    fl.scalar = 0;
    fl.eps_q = nullptr;
    fl.overflow = nullptr;
    fl.counters = idx->d_counters;
    fl.n_counters = (int)need_counters;
    fl.status_out = d_status_out;
    if (debug_stages())
        fprintf(stderr, "[dawn debug] finalize: nq %d lists %d k' %d k %d out %p\n", fl.nq, fl.n_lists, fl.kprime, fl.k,
                (void *)fl.labels_out);
    EventPair ev;
    bool timed = begin_event(idx, 1, s, &ev);
    CK(idx, launch_finalize(fl, s));
    debug_wait(idx, s, "finalize (scan path)");
    if (timed) end_event(idx, ev, s);
    idx->prof.finalize_launches++;
    idx->prof.kernel_launches++;
    idx->prof.queries += batch;
    if (idx->pending.size() > 4096) drain_events(idx);
    return DAWN_OK;
}

}  // namespace

// ------------------------------------------------------------------------------ C ABI

extern "C" {

const char *dawn_last_error(void) { return g_last_error.c_str(); }
const char *dawn_version(void) { return "libdawn_b200 0.1.0 sm_100a"; }

int dawn_index_create(const dawn_options *opts, dawn_index **out) {
    if (!out) return fail(DAWN_ERR_INVALID, "out is null");
    *out = nullptr;
    dawn_options o{};
    if (opts) o = *opts;
    if (o.dimensions != 0 && o.dimensions != DAWN_DIMENSIONS)
        return fail(DAWN_ERR_INVALID, "dimensions must be %d (src/search/vector.rs:26), got %u",
                    DAWN_DIMENSIONS, o.dimensions);
    if (o.metric != DAWN_METRIC_IP) return fail(DAWN_ERR_INVALID, "only MetricKind::IP is supported");
    if (o.scalar != DAWN_SCALAR_F16 && o.scalar != DAWN_SCALAR_I8)
        return fail(DAWN_ERR_INVALID, "scalar must be DAWN_SCALAR_F16 or DAWN_SCALAR_I8, got %u", o.scalar);
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        return fail(DAWN_ERR_CUDA, "no CUDA device available (%s); libdawn_b200 has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (o.device < 0 || o.device >= n_dev)
        return fail(DAWN_ERR_INVALID, "device %d out of range (found %d devices)", o.device, n_dev);
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, o.device);
    if (e != cudaSuccess) return fail(DAWN_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return fail(DAWN_ERR_CUDA, "device %d is sm_%d%d; libdawn_b200 is built for sm_100a only", o.device,
                    prop.major, prop.minor);
    dawn_index *idx = new (std::nothrow) dawn_index();
    if (!idx) return fail(DAWN_ERR_INTERNAL, "out of host memory");
    idx->device = o.device;
    idx->scalar = (int)o.scalar;
    idx->sm_count = prop.multiProcessorCount;
    int rc = DAWN_OK;
    do {
        if ((e = cudaSetDevice(o.device)) != cudaSuccess) break;
        if ((e = cudaStreamCreateWithFlags(&idx->stream, cudaStreamNonBlocking)) != cudaSuccess) break;
        for (int b = 0; b < 2 && e == cudaSuccess; b++) {
            if ((e = cudaMallocHost(&idx->h_stage_buf[b], kStageRowsHost * kDim * sizeof(float))) != cudaSuccess) break;
            if ((e = cudaMallocHost(&idx->h_labels_buf[b], kStageRowsHost * sizeof(uint64_t))) != cudaSuccess) break;
            if ((e = cudaMalloc(&idx->d_stage_buf[b], kStageRowsHost * kDim * sizeof(float))) != cudaSuccess) break;
            if ((e = cudaEventCreateWithFlags(&idx->stage_done[b], cudaEventDisableTiming)) != cudaSuccess) break;
        }
        if (e != cudaSuccess) break;
        idx->h_stage = idx->h_stage_buf[0];
        idx->h_stage_labels = idx->h_labels_buf[0];
        idx->d_stage = idx->d_stage_buf[0];
        if ((e = cudaMallocHost(&idx->h_word, 64)) != cudaSuccess) break;
        if ((e = cudaEventCreateWithFlags(&idx->ws_done, cudaEventDisableTiming)) != cudaSuccess) break;
    } while (0);
    if (e != cudaSuccess) {
        rc = fail(DAWN_ERR_CUDA, "index setup failed: %s", cudaGetErrorString(e));
        dawn_index_free(idx);
        return rc;
    }
    if (o.capacity > 0) {
        rc = grow_physical(idx, o.capacity);
        if (rc != DAWN_OK) {
            dawn_index_free(idx);
            return rc;
        }
        idx->capacity = o.capacity;
    }
    *out = idx;
    return DAWN_OK;
}

void dawn_index_free(dawn_index *idx) {
    if (!idx) return;
    cudaSetDevice(idx->device);
    if (idx->stream) cudaStreamSynchronize(idx->stream);
    for (auto &p : idx->pending) idx->free_events.push_back(p);
    for (auto &p : idx->free_events) {
        cudaEventDestroy(p.a);
        cudaEventDestroy(p.b);
    }
    cudaFree(idx->corpus);
    cudaFree(idx->labels);
    for (int b = 0; b < 2; b++) {
        cudaFree(idx->d_stage_buf[b]);
        cudaFreeHost(idx->h_stage_buf[b]);
        cudaFreeHost(idx->h_labels_buf[b]);
        if (idx->stage_done[b]) cudaEventDestroy(idx->stage_done[b]);
    }
    cudaFree(idx->d_queries);
    cudaFreeHost(idx->h_queries);
    cudaFree(idx->d_result);
    cudaFreeHost(idx->h_result);
    cudaFree(idx->d_i8_scratch);
    cudaFree(idx->d_i8_lists);
    cudaFree(idx->d_i8_overflow);
    cudaFree(idx->d_partials);
    cudaFree(idx->d_counters);
    cudaFree(idx->d_gemm_ws);
    cudaFreeHost(idx->h_word);
    if (idx->ws_done) cudaEventDestroy(idx->ws_done);
    if (idx->stream) cudaStreamDestroy(idx->stream);
    cudaGetLastError();
    delete idx;
}

int dawn_index_reserve(dawn_index *idx, size_t n) {
    int rc = check_alive(idx);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(idx->mu);
    if (n <= idx->capacity) return DAWN_OK;
    rc = grow_physical(idx, n);
    if (rc) return rc;
    idx->capacity = n;
    return DAWN_OK;
}

int dawn_index_add_batch(dawn_index *idx, const uint64_t *labels, const float *vectors, size_t n) {
    int rc = check_alive(idx);
    if (rc) return rc;
    if (n == 0) return DAWN_OK;
    if (!labels || !vectors) return fail(DAWN_ERR_INVALID, "labels / vectors is null");
    std::lock_guard<std::mutex> lk(idx->mu);
    if (idx->size + idx->staged + n > idx->capacity)
        return fail(DAWN_ERR_CAPACITY, "add of %zu vectors exceeds capacity %zu (size %zu): reserve first", n,
                    idx->capacity, idx->size + idx->staged);
    size_t done = 0;
    while (done < n) {
        size_t room = kStageRowsHost - idx->staged;
        size_t take = n - done < room ? n - done : room;
        memcpy(idx->h_stage + idx->staged * kDim, vectors + done * kDim, take * kDim * sizeof(float));
        memcpy(idx->h_stage_labels + idx->staged, labels + done, take * sizeof(uint64_t));
        idx->staged += take;
        done += take;
        if (idx->staged == kStageRowsHost) {
            rc = flush_staged_async(idx);  // the next chunk is copied on the host while this one moves
            if (rc) return rc;
        }
    }
    return DAWN_OK;
}

int dawn_index_add(dawn_index *idx, uint64_t label, const float *vector384) {
    return dawn_index_add_batch(idx, &label, vector384, 1);
}

int dawn_index_add_synthetic(dawn_index *idx, uint64_t seed, uint64_t first_row, size_t n) {
    int rc = check_alive(idx);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(idx->mu);
    rc = flush_staged(idx);
    if (rc) return rc;
    if (idx->size + n > idx->capacity)
        return fail(DAWN_ERR_CAPACITY, "add of %zu vectors exceeds capacity %zu (size %zu): reserve first", n,
                    idx->capacity, idx->size);
    if (idx->scalar == DAWN_SCALAR_I8) CK(idx, launch_synth_i8(arena_i8(idx), idx->size, seed, first_row, n, idx->stream));
    else CK(idx, launch_synth_f16(idx->corpus + idx->size * kDim, seed, first_row, n, idx->stream));
    idx->prof.kernel_launches++;
    // labels = first_row + i + 1, written by a tiny host-free path: reuse the staging buffer
    size_t done = 0;
    while (done < n) {
        size_t take = n - done < kStageRowsHost ? n - done : kStageRowsHost;
        for (size_t i = 0; i < take; i++) idx->h_stage_labels[i] = first_row + done + i + 1;
        CK(idx, cudaMemcpyAsync(idx->labels + idx->size + done, idx->h_stage_labels, take * sizeof(uint64_t),
                                cudaMemcpyHostToDevice, idx->stream));
        CK(idx, cudaStreamSynchronize(idx->stream));
        done += take;
    }
    CK(idx, cudaStreamSynchronize(idx->stream));
    idx->size += n;
    return DAWN_OK;
}

namespace {
struct LimitScope {  // the pushed-down limit applies to one host call only
    dawn_index *idx;
    LimitScope(dawn_index *i, float distance_limit) : idx(i) {
        if (distance_limit == distance_limit && distance_limit < INFINITY)
            idx->limit_score = (float)(1.0 - (double)distance_limit);
    }
    ~LimitScope() { idx->limit_score = -INFINITY; }
};
int search_batch_host(dawn_index *idx, const float *queries, size_t batch, size_t k, float distance_limit,
                      uint64_t *labels_out, float *distances_out, size_t *counts_out);
}  // namespace

int dawn_index_search_batch(dawn_index *idx, const float *queries, size_t batch, size_t k,
                            uint64_t *labels_out, float *distances_out, size_t *counts_out) {
    return search_batch_host(idx, queries, batch, k, NAN, labels_out, distances_out, counts_out);
}

// (f3) distance_limit of UdpPacket::Search (/root/reference/src/net/udp_packets.rs:29-39): hits with
// distance >= limit are not returned (src/net/udp_service.rs:196-199).  Results are ascending, so the
// filter truncates; the limit is also pushed down into the scan kernels as a score floor, so rows that
// cannot pass it are never appended, merged or re-scored.  limit = +inf or NaN keeps everything.
int dawn_index_search_limit(dawn_index *idx, const float *query384, size_t k, float distance_limit,
                            uint64_t *labels_out, float *distances_out, size_t *count_out) {
    int rc = search_batch_host(idx, query384, 1, k, distance_limit, labels_out, distances_out, count_out);
    if (rc != DAWN_OK) return rc;
    if (distance_limit == distance_limit) {
        size_t keep = 0;
        while (keep < *count_out && distances_out[keep] < distance_limit) keep++;
        *count_out = keep;
    }
    return DAWN_OK;
}

namespace {
int search_batch_host(dawn_index *idx, const float *queries, size_t batch, size_t k, float distance_limit,
                      uint64_t *labels_out, float *distances_out, size_t *counts_out) {
    int rc = check_alive(idx);
    if (rc) return rc;
    if (batch == 0) return DAWN_OK;
    if (!queries || !counts_out) return fail(DAWN_ERR_INVALID, "queries / counts_out is null");
    if (k > DAWN_MAX_K) return fail(DAWN_ERR_INVALID, "k = %zu exceeds DAWN_MAX_K = %d", k, DAWN_MAX_K);
    if (k > 0 && (!labels_out || !distances_out)) return fail(DAWN_ERR_INVALID, "output buffer is null");
    std::lock_guard<std::mutex> lk(idx->mu);
    LimitScope limit_scope(idx, distance_limit);
    rc = flush_staged(idx);
    if (rc) return rc;
    if (k == 0 || idx->size == 0) {
        for (size_t b = 0; b < batch; b++) counts_out[b] = 0;
        return DAWN_OK;
    }
    rc = ensure_query_ws(idx, batch, k);
    if (rc) return rc;
    cudaStream_t s = idx->stream;
    memcpy(idx->h_queries, queries, batch * kDim * sizeof(float));
    CK(idx, cudaMemcpyAsync(idx->d_queries, idx->h_queries, batch * kDim * sizeof(float), cudaMemcpyHostToDevice, s));
    int kprime = choose_kprime(k);
    const uint64_t gemm_before = idx->prof.gemm_batches;
    ResultView dv = result_view(idx->d_result, batch, k), hv = result_view(idx->h_result, batch, k);
    // A small result block is written by the finalize kernel straight into the pinned host buffer (it is
    // device-accessible under unified addressing): a few hundred bytes of posted PCIe writes instead of a
    // copy-engine round trip after the kernel.  Large blocks go through one D2H copy.
    // (Passing the single query by value as a launch parameter instead of the H2D copy was tried as well:
    // no gain, the copy already overlaps the launches.)
    const bool direct = dv.bytes <= kDirectResultBytes && !getenv("DAWN_NO_DIRECT_RESULT");
    const ResultView &ov = direct ? hv : dv;
    rc = search_enqueue(idx, idx->d_queries, batch, k, kprime, ov.labels, ov.dist, ov.counts, ov.flags, s, false, ov.status);
    if (rc) return rc;
    if (!direct) CK(idx, cudaMemcpyAsync(idx->h_result, idx->d_result, dv.bytes, cudaMemcpyDeviceToHost, s));  // the one D2H
    CK(idx, cudaStreamSynchronize(s));
    if (hv.status[0] != 0) return fail(DAWN_ERR_INTERNAL, "scan kernel reported status 0x%x", hv.status[0]);
    memcpy(labels_out, hv.labels, batch * k * sizeof(uint64_t));
    memcpy(distances_out, hv.dist, batch * k * sizeof(float));
    for (size_t b = 0; b < batch; b++) counts_out[b] = hv.counts[b];

    // Exactness certificate not met (near-ties deeper than the slack, or a tensor-core-path log
    // overflow): re-run those queries through the f32 scan with the longest candidate list.
    const bool can_escalate = kprime < kMaxCand || idx->prof.gemm_batches > gemm_before;
    std::vector<size_t> redo;
    for (size_t b = 0; b < batch; b++)
        if (!(hv.flags[b] & 1u)) redo.push_back(b);
    if (!can_escalate) {
        idx->prof.uncertified += redo.size();
        return DAWN_OK;
    }
    for (size_t b : redo) {
        idx->prof.escalations++;
        ResultView d1 = result_view(idx->d_result, 1, k), h1 = result_view(idx->h_result, 1, k);
        const bool direct1 = d1.bytes <= kDirectResultBytes;
        const ResultView &o1 = direct1 ? h1 : d1;
        rc = search_enqueue(idx, idx->d_queries + b * kDim, 1, k, kMaxCand, o1.labels, o1.dist, o1.counts, o1.flags, s,
                            /*scan_only=*/true, o1.status);
        if (rc) return rc;
        if (!direct1) CK(idx, cudaMemcpyAsync(idx->h_result, idx->d_result, d1.bytes, cudaMemcpyDeviceToHost, s));
        CK(idx, cudaStreamSynchronize(s));
        memcpy(labels_out + b * k, h1.labels, k * sizeof(uint64_t));
        memcpy(distances_out + b * k, h1.dist, k * sizeof(float));
        counts_out[b] = h1.counts[0];
        if (!(h1.flags[0] & 1u)) idx->prof.uncertified++;
    }
    return DAWN_OK;
}
}  // namespace

int dawn_index_search(dawn_index *idx, const float *query384, size_t k, uint64_t *labels_out,
                      float *distances_out, size_t *count_out) {
    return dawn_index_search_batch(idx, query384, 1, k, labels_out, distances_out, count_out);
}

int dawn_index_search_device(dawn_index *idx, const float *d_queries, size_t batch, size_t k,
                             uint64_t *d_labels_out, float *d_distances_out, uint32_t *d_counts_out,
                             uint32_t *d_flags_out, void *stream) {
    int rc = check_alive(idx);
    if (rc) return rc;
    if (batch == 0) return DAWN_OK;
    if (!d_queries || !d_labels_out || !d_distances_out || !d_counts_out || !d_flags_out)
        return fail(DAWN_ERR_INVALID, "null device pointer");
    if (k == 0 || k > DAWN_MAX_K) return fail(DAWN_ERR_INVALID, "k = %zu out of range 1..%d", k, DAWN_MAX_K);
    std::lock_guard<std::mutex> lk(idx->mu);
    rc = flush_staged(idx);
    if (rc) return rc;
    cudaStream_t s = stream ? (cudaStream_t)stream : idx->stream;
    if (idx->size == 0) {
        CK(idx, cudaMemsetAsync(d_counts_out, 0, batch * sizeof(uint32_t), s));
        CK(idx, cudaMemsetAsync(d_flags_out, 0, batch * sizeof(uint32_t), s));
        return DAWN_OK;
    }
    return search_enqueue(idx, d_queries, batch, k, choose_kprime(k), d_labels_out, d_distances_out, d_counts_out,
                          d_flags_out, s);
}

size_t dawn_index_size(const dawn_index *idx) { return idx ? idx->size + idx->staged : 0; }
size_t dawn_index_capacity(const dawn_index *idx) { return idx ? idx->capacity : 0; }
size_t dawn_index_dimensions(const dawn_index *idx) { return idx ? DAWN_DIMENSIONS : 0; }

int dawn_index_get(dawn_index *idx, uint64_t label, float *vector384_out) {
    int rc = check_alive(idx);
    if (rc) return rc;
    if (!vector384_out) return fail(DAWN_ERR_INVALID, "output is null");
    std::lock_guard<std::mutex> lk(idx->mu);
    rc = flush_staged(idx);
    if (rc) return rc;
    if (idx->size == 0) return fail(DAWN_ERR_INVALID, "label %llu not found", (unsigned long long)label);
    rc = ensure_query_ws(idx, 1, 1);
    if (rc) return rc;
    cudaStream_t s = idx->stream;
    uint32_t *d_row = reinterpret_cast<uint32_t *>(idx->d_result);
    uint32_t *h_row = reinterpret_cast<uint32_t *>(idx->h_result);
    CK(idx, launch_find_label(idx->labels, idx->size, label, d_row, s));
    idx->prof.kernel_launches++;
    CK(idx, cudaMemcpyAsync(h_row, d_row, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CK(idx, cudaStreamSynchronize(s));
    if (h_row[0] == kNoRow) return fail(DAWN_ERR_INVALID, "label %llu not found", (unsigned long long)label);
    if (idx->scalar == DAWN_SCALAR_I8) CK(idx, launch_gather_f32_i8(arena_i8(idx), d_row, 1, idx->d_queries, s));
    else CK(idx, launch_gather_f32(idx->corpus, d_row, 1, idx->d_queries, s));
    idx->prof.kernel_launches++;
    CK(idx, cudaMemcpyAsync(idx->h_queries, idx->d_queries, kDim * sizeof(float), cudaMemcpyDeviceToHost, s));
    CK(idx, cudaStreamSynchronize(s));
    memcpy(vector384_out, idx->h_queries, kDim * sizeof(float));
    return DAWN_OK;
}

// ---- save / load: header, labels, fp16 rows (raw device layout) -------------------------
struct SaveHeader {
    char magic[8];  // "DAWNB200"
    uint32_t version, scalar, dims, reserved;
    uint64_t size;
};

int dawn_index_save(dawn_index *idx, const char *path) {
    int rc = check_alive(idx);
    if (rc) return rc;
    if (!path) return fail(DAWN_ERR_INVALID, "path is null");
    std::lock_guard<std::mutex> lk(idx->mu);
    rc = flush_staged(idx);
    if (rc) return rc;
    std::string tmp = std::string(path) + ".tmp";
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f) return fail(DAWN_ERR_IO, "cannot open %s for writing", tmp.c_str());
    SaveHeader h{};
    memcpy(h.magic, "DAWNB200", 8);
    h.version = 1;
    h.scalar = (uint32_t)idx->scalar;
    h.dims = kDim;
    h.size = idx->size;
    bool ok = fwrite(&h, sizeof h, 1, f) == 1;
    // stream device memory out through the pinned staging buffer
    const size_t buf_bytes = kStageRowsHost * kDim * sizeof(float);
    auto dump = [&](const void *dptr, size_t bytes) -> int {
        size_t off = 0;
        while (ok && off < bytes) {
            size_t take = bytes - off < buf_bytes ? bytes - off : buf_bytes;
            CK(idx, cudaMemcpyAsync(idx->h_stage, (const char *)dptr + off, take, cudaMemcpyDeviceToHost, idx->stream));
            CK(idx, cudaStreamSynchronize(idx->stream));
            ok = fwrite(idx->h_stage, 1, take, f) == take;
            off += take;
        }
        return DAWN_OK;
    };
    rc = dump(idx->labels, idx->size * sizeof(uint64_t));
    if (rc == DAWN_OK) rc = dump(idx->corpus, arena_bytes(idx, idx->size));
    ok = (fclose(f) == 0) && ok;
    if (rc != DAWN_OK || !ok) {
        remove(tmp.c_str());
        return rc != DAWN_OK ? rc : fail(DAWN_ERR_IO, "short write to %s", tmp.c_str());
    }
    if (rename(tmp.c_str(), path) != 0) {
        remove(tmp.c_str());
        return fail(DAWN_ERR_IO, "cannot rename %s to %s", tmp.c_str(), path);
    }
    return DAWN_OK;
}

int dawn_index_load(dawn_index *idx, const char *path) {
    int rc = check_alive(idx);
    if (rc) return rc;
    if (!path) return fail(DAWN_ERR_INVALID, "path is null");
    std::lock_guard<std::mutex> lk(idx->mu);
    FILE *f = fopen(path, "rb");
    if (!f) return fail(DAWN_ERR_IO, "cannot open %s", path);
    SaveHeader h{};
    if (fread(&h, sizeof h, 1, f) != 1 || memcmp(h.magic, "DAWNB200", 8) != 0 || h.version != 1 ||
        h.scalar != (uint32_t)idx->scalar || h.dims != kDim) {
        fclose(f);
        return fail(DAWN_ERR_IO, "%s is not a libdawn_b200 index file with this index's storage type", path);
    }
    // validate the length before touching the index, so a failed load leaves it unchanged
    // (the reference falls back to a rebuild when load fails, search_provider.rs:115-116)
    fseek(f, 0, SEEK_END);
    long long flen = ftell(f);
    long long want = (long long)sizeof h + (long long)h.size * 8 + (long long)arena_bytes(idx, h.size);
    if (flen != want) {
        fclose(f);
        return fail(DAWN_ERR_IO, "%s is truncated (%lld bytes, expected %lld)", path, flen, want);
    }
    fseek(f, sizeof h, SEEK_SET);
    rc = grow_physical(idx, h.size);
    if (rc) {
        fclose(f);
        return rc;
    }
    const size_t buf_bytes = kStageRowsHost * kDim * sizeof(float);
    bool ok = true;
    auto slurp = [&](void *dptr, size_t bytes) -> int {
        size_t off = 0;
        while (ok && off < bytes) {
            size_t take = bytes - off < buf_bytes ? bytes - off : buf_bytes;
            ok = fread(idx->h_stage, 1, take, f) == take;
            if (!ok) break;
            CK(idx, cudaMemcpyAsync((char *)dptr + off, idx->h_stage, take, cudaMemcpyHostToDevice, idx->stream));
            CK(idx, cudaStreamSynchronize(idx->stream));
            off += take;
        }
        return DAWN_OK;
    };
    idx->staged = 0;
    idx->size = 0;  // from here on the old contents are gone
    rc = slurp(idx->labels, h.size * sizeof(uint64_t));
    if (rc == DAWN_OK) rc = slurp(idx->corpus, arena_bytes(idx, h.size));
    fclose(f);
    if (rc != DAWN_OK) return rc;
    if (!ok) return fail(DAWN_ERR_IO, "read error on %s", path);
    idx->size = h.size;
    if (idx->capacity < idx->size) idx->capacity = idx->size;
    return DAWN_OK;
}

int dawn_index_set_option(dawn_index *idx, const char *key, int64_t value) {
    int rc = check_alive(idx);
    if (rc) return rc;
    if (!key) return fail(DAWN_ERR_INVALID, "key is null");
    std::lock_guard<std::mutex> lk(idx->mu);
    if (!strcmp(key, "gemm_min_batch")) idx->gemm_min_batch = value;
    else if (!strcmp(key, "gemm_min_rows")) idx->gemm_min_rows = value;
    else if (!strcmp(key, "gemm_small_batch")) idx->gemm_small_batch = value;
    else if (!strcmp(key, "gemm_small_batch_rows")) idx->gemm_small_batch_rows = value;
    else if (!strcmp(key, "force_path")) idx->force_path = value;
    else if (!strcmp(key, "gemm_cta_group")) idx->gemm_cta_group = value;
    else if (!strcmp(key, "gemm_chunk_tiles")) idx->gemm_chunk_tiles = value;
    else if (!strcmp(key, "gemm_sequential_tiles")) idx->gemm_sequential_tiles = value;
    else if (!strcmp(key, "gemm_growth")) idx->gemm_growth = value;
    else if (!strcmp(key, "i8_tensor_min_batch")) idx->i8_tensor_min_batch = value;
    else if (!strcmp(key, "i8_tensor_chunk_rows")) idx->i8_tensor_chunk_rows = value < 65536 ? 65536 : value;
    else return fail(DAWN_ERR_INVALID, "unknown option '%s'", key);
    return DAWN_OK;
}

int dawn_index_set_profiling(dawn_index *idx, int enable) {
    int rc = check_alive(idx);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(idx->mu);
    drain_events(idx);
    idx->profiling = enable != 0;
    return DAWN_OK;
}

int dawn_index_get_profile(dawn_index *idx, dawn_profile *out, int reset) {
    int rc = check_alive(idx);
    if (rc) return rc;
    if (!out) return fail(DAWN_ERR_INVALID, "out is null");
    std::lock_guard<std::mutex> lk(idx->mu);
    drain_events(idx);
    *out = idx->prof;
    if (reset) idx->prof = dawn_profile{};
    return DAWN_OK;
}

int dawn_merge_results_device(int device, const uint64_t *d_labels, const float *d_distances,
                              const uint32_t *d_counts, size_t n_lists, size_t list_stride_bytes,
                              size_t batch, size_t k, uint64_t *d_labels_out, float *d_distances_out,
                              uint32_t *d_counts_out, void *stream) {
    if (!d_labels || !d_distances || !d_counts || !d_labels_out || !d_distances_out || !d_counts_out)
        return fail(DAWN_ERR_INVALID, "null device pointer");
    if (k == 0 || k > DAWN_MAX_K || n_lists == 0 || n_lists * k > 1024)
        return fail(DAWN_ERR_INVALID, "merge shape out of range (n_lists=%zu k=%zu; n_lists*k must be <= 1024)",
                    n_lists, k);
    if (batch == 0) return DAWN_OK;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess)
        e = launch_merge_results(d_labels, d_distances, d_counts, (int)n_lists, list_stride_bytes, (int)batch, (int)k,
                                 d_labels_out, d_distances_out, d_counts_out, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(DAWN_ERR_CUDA, "merge launch failed: %s", cudaGetErrorString(e));
    return DAWN_OK;
}

}  // extern "C"

// ------------------------------------------------------------------ small kernels

namespace {

// K7 (device side of the sharded search): merge n_lists sorted result lists per query.
// Order: distance ascending, label ascending, then list index (strict).  One CTA per query,
// one thread per input entry; rank = position in own list + binary-searched counts in the others.
__global__ void __launch_bounds__(1024) merge_results_kernel(
    const uint64_t *__restrict__ labels, const float *__restrict__ dist, const uint32_t *__restrict__ counts,
    int n_lists, size_t stride_l, size_t stride_d, size_t stride_c, int batch, int k,
    uint64_t *__restrict__ labels_out, float *__restrict__ dist_out, uint32_t *__restrict__ counts_out) {
    __shared__ uint64_t s_lab[1024];
    __shared__ float s_dist[1024];
    __shared__ int s_cnt[64];
    const int qi = blockIdx.x;
    const int tid = threadIdx.x;
    const int total = n_lists * k;
    // list l's arrays start l * stride bytes after the base pointers (dense arrays or one packed
    // block per shard as it arrives from the all-gather)
    if (tid < n_lists)
        s_cnt[tid] = min((int)reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(counts) + tid * stride_c)[qi], k);
    if (tid < total) {
        const int l = tid / k, p = tid % k;
        s_lab[tid] = reinterpret_cast<const uint64_t *>(reinterpret_cast<const char *>(labels) + l * stride_l)[(size_t)qi * k + p];
        s_dist[tid] = reinterpret_cast<const float *>(reinterpret_cast<const char *>(dist) + l * stride_d)[(size_t)qi * k + p];
    }
    __syncthreads();
    int all = 0;
    for (int l = 0; l < n_lists; l++) all += s_cnt[l];
    if (tid < total) {
        const int l = tid / k, p = tid % k;
        if (p < s_cnt[l]) {
            const float d = s_dist[tid];
            const uint64_t lab = s_lab[tid];
            int rank = p;
            for (int o = 0; o < n_lists; o++) {
                if (o == l) continue;
                int lo = 0, hi = s_cnt[o];
                while (lo < hi) {  // count entries of list o that come before (d, lab, l)
                    const int mid = (lo + hi) >> 1;
                    const float od = s_dist[o * k + mid];
                    const uint64_t ol = s_lab[o * k + mid];
                    const bool before = od < d || (od == d && (ol < lab || (ol == lab && o < l)));
                    if (before) lo = mid + 1;
                    else hi = mid;
                }
                rank += lo;
            }
            if (rank < k) {
                labels_out[(size_t)qi * k + rank] = lab;
                dist_out[(size_t)qi * k + rank] = d;
            }
        }
    }
    if (tid == 0) counts_out[qi] = (uint32_t)min(all, k);
}

__global__ void __launch_bounds__(256) find_label_kernel(const uint64_t *__restrict__ labels, size_t n,
                                                         uint64_t label, uint32_t *__restrict__ row_out) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        if (labels[i] == label) atomicMin(row_out, (uint32_t)i);
}

cudaError_t launch_merge_results(const uint64_t *labels, const float *dist, const uint32_t *counts, int n_lists,
                                 size_t list_stride_bytes, int batch, int k, uint64_t *labels_out, float *dist_out,
                                 uint32_t *counts_out, cudaStream_t s) {
    if (n_lists > 64) return cudaErrorInvalidValue;
    size_t sl = list_stride_bytes, sd = list_stride_bytes, sc = list_stride_bytes;
    if (list_stride_bytes == 0) {  // dense [n_lists][batch][k] arrays
        sl = (size_t)batch * k * sizeof(uint64_t);
        sd = (size_t)batch * k * sizeof(float);
        sc = (size_t)batch * sizeof(uint32_t);
    }
    merge_results_kernel<<<batch, 1024, 0, s>>>(labels, dist, counts, n_lists, sl, sd, sc, batch, k, labels_out,
                                                dist_out, counts_out);
    return cudaGetLastError();
}

cudaError_t launch_find_label(const uint64_t *labels, size_t n, uint64_t label, uint32_t *row_out,
                              cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(row_out, 0xFF, sizeof(uint32_t), s);
    if (e != cudaSuccess) return e;
    size_t blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    find_label_kernel<<<(unsigned)blocks, 256, 0, s>>>(labels, n, label, row_out);
    return cudaGetLastError();
}

}  // namespace
