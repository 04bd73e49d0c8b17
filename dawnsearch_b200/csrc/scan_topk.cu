// scan_topk.cu -- K2: small-batch streaming scan of the fp16 corpus with a fused per-CTA top-k'.
//
// Replaces usearch's Index::search (/root/reference/src/search/search_provider.rs:214):
// instead of an approximate HNSW walk, every stored page vector is scored against the
// query (dot product, src/search/vector.rs:99-101) and the best k' rows are kept.
// Raw scores never touch HBM: the only global writes are k' candidates per CTA.
//
// Roofline: HBM.  Algorithmic bytes = n_rows * 768 per pass (1, 2 or 4 queries share a pass).
//
// Structure (one persistent CTA per SM, 17 warps):
//   warp 16, lane 0  producer: claims 256-row chunks from a global counter and streams them
//                    as 8-row (6 KB) stages into a 32-deep shared-memory ring with
//                    cp.async.bulk (TMA bulk copy), completion on one mbarrier per stage.
//   warps 0..15      consumers: warp w owns stages w, w+16, ...  Each lane reads its 24 bytes
//                    of a row (one 128-bit + one 64-bit shared load), converts to f32 and
//                    FMAs against the query columns it keeps in registers (f32 query: no
//                    query rounding), then a transposed butterfly (shuffle) reduction turns
//                    8 rows x QT queries of per-lane partials into one finished score per
//                    lane.  A score passes if it is >= the CTA's current k'-th best; passing
//                    rows (rare after warm-up) are appended to a shared buffer.
//   prune            when the buffer reaches its high-water mark all consumer warps meet at a
//                    named barrier and merge buffer + sorted list by rank counting; the k'-th
//                    score becomes the new threshold.
// A row's score is a pure function of (row bytes, query): the FMA order per lane and the
// butterfly order are fixed, so identical rows tie exactly and ties break on the label.
#include "dawn_common.cuh"

namespace dawn {

#ifdef DAWN_SCAN_TRACE  // tools/scan_trace.cu: per-CTA timeline of the fp16 scan (never defined in the library build)
__device__ unsigned long long g_scan_trace[148][12];
#define SCAN_TRACE(slot)                                                          \
    do {                                                                          \
        unsigned long long t_;                                                    \
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                    \
        if (blockIdx.x < 148) g_scan_trace[blockIdx.x][slot] = t_;                \
    } while (0)
#define SCAN_TRACE_ADD(slot, v)                                                   \
    do {                                                                          \
        if (blockIdx.x < 148) g_scan_trace[blockIdx.x][slot] += (v);              \
    } while (0)
#else
#define SCAN_TRACE(slot) do { } while (0)
#define SCAN_TRACE_ADD(slot, v) do { } while (0)
#endif

namespace {

constexpr int kConsumerWarps = 16;
constexpr int kConsumerThreads = kConsumerWarps * 32;
constexpr int kScanThreads = kConsumerThreads + 32;
constexpr int kStageRows = 8;
constexpr int kStageBytes = kStageRows * kRowBytesF16;  // 6144
constexpr int kStages = 32;
constexpr int kChunkStages = 32;
constexpr int kChunkRows = kStageRows * kChunkStages;  // 256 rows = 192 KB per claim
constexpr int kBufCap = 256;
constexpr int kHighWater = 128;  // >= kMaxCand so the first prune fills the list
static_assert(kHighWater + kConsumerWarps * kStageRows <= kBufCap, "append buffer can overflow");
static_assert(kHighWater >= kMaxCand, "first prune must be able to fill the list");

struct alignas(16) StageMeta {  // written by the producer lane with ONE 16-byte store (it is the critical path)
    uint32_t row_base;
    uint32_t n_rows;  // 0 = end-of-stream sentinel
    uint32_t stage;   // which stage the slot holds (debugging aid)
    uint32_t pad;
};

// ---- PTX wrappers ---------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                         uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void consumer_bar_sync() {
    asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory");
}

__device__ __forceinline__ void publish_stage(StageMeta *m, uint32_t row_base, uint32_t n_rows, uint32_t g) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(m)), "r"(row_base), "r"(n_rows), "r"(g), "r"(0u)
                 : "memory");
}

// Wait for this warp's next stage -- unless an append buffer reaches its high-water mark first.
// Warp w owns stages w, w+16, ...  A warp parked at the prune barrier leaves its next two landed stages
// unconsumed; the in-order producer blocks on their slots; a warp that then sat here in a plain wait for a LATER
// stage would never reach the barrier: deadlock (seen as a hang of single-query searches over 2M rows, once a slow
// append -- its label load queued behind the bulk copies -- had put one warp 32 stages behind the others).  So the
// wait gives up when a prune is due; the stage is simply waited for again afterwards.
// Returns true if the stage has landed, false if the warp has to join a prune first.
template <int QT>
__device__ __forceinline__ bool wait_stage_or_prune(uint64_t *full_bar, uint32_t parity, const uint32_t *cnt, int lane) {
    while (true) {
        const bool ok = mbar_try_wait(full_bar, parity);
        if (__all_sync(0xffffffffu, ok)) return true;
        const uint32_t c = lane < QT ? ((volatile const uint32_t *)cnt)[lane] : 0u;
        if (__any_sync(0xffffffffu, c >= (uint32_t)kHighWater)) return false;
    }
}

// ---- shared-memory layout -------------------------------------------------------------
template <int QT, int STAGES = kStages, int STAGE_BYTES = kStageBytes, int BUFCAP = kBufCap>
struct ScanSmem {
    static constexpr int kBuf = BUFCAP;
    alignas(128) uint8_t ring[STAGES][STAGE_BYTES];
    alignas(16) Cand list[QT][kMaxCand];
    alignas(16) Cand buf[QT][BUFCAP];
    alignas(8) uint64_t full_bar[STAGES];
    alignas(8) uint64_t empty_bar[STAGES];
    StageMeta meta[STAGES];
    uint32_t cnt[QT];
    uint32_t list_len[QT];
    float thr[QT];
    float floor[QT];  // scores below this are never candidates (distance_limit pushed down; -inf = none)
    uint32_t done_warps;
    uint32_t overflow;
};

// Merge the append buffer into the sorted list, all 512 consumer threads, one query at a time.
// Rank of an entry = number of entries that beat it; the list part is already sorted so only
// the buffer needs counting.  Entries are unique under cand_better, so ranks are unique.
// The streaming loop is stalled while this runs (the ring fills up and HBM goes idle), so the count
// is spread over all threads -- 4 or 2 adjacent lanes share an entry when the entries are few -- and
// its inner loop compares scores only (one LDS + two predicated adds per pair); the label / row
// tie-break is a second pass run only by entries that actually met an equal score.
template <int QT, class Smem>
__device__ __forceinline__ void prune(Smem &sm, int kprime, int tid) {
#pragma unroll 1
    for (int q = 0; q < QT; q++) {
        const int nb = min((int)((volatile uint32_t *)sm.cnt)[q], Smem::kBuf);
        const int len = (int)((volatile uint32_t *)sm.list_len)[q];
        const int total = len + nb;
        const int shift = total <= kConsumerThreads / 4 ? 2 : total <= kConsumerThreads / 2 ? 1 : 0;
        const int ent = tid >> shift;          // entry handled by this thread
        const int sub = tid & ((1 << shift) - 1);
        const bool have = ent < total && nb > 0;
        const bool from_buf = ent >= len;
        Cand e = empty_cand();
        int rank = 0, gt = 0, eq = 0;
        if (have) {
            if (!from_buf) {
                e = sm.list[q][ent];
                rank = ent;
            } else {
                e = sm.buf[q][ent - len];
                int lo = 0, hi = len;
                while (lo < hi) {
                    int mid = (lo + hi) >> 1;
                    if (cand_better(sm.list[q][mid], e)) lo = mid + 1;
                    else hi = mid;
                }
                rank = lo;
            }
            const float es = e.score;
#pragma unroll 8
            for (int j = sub; j < nb; j += 1 << shift) {
                const float s = sm.buf[q][j].score;
                gt += s > es ? 1 : 0;
                eq += s == es ? 1 : 0;
            }
        }
        for (int o = 1; o < (1 << shift); o <<= 1) {
            gt += __shfl_xor_sync(0xffffffffu, gt, o);
            eq += __shfl_xor_sync(0xffffffffu, eq, o);
        }
        rank += gt;
        if (have && sub == 0 && eq > (from_buf ? 1 : 0)) {  // equal scores: order them by (label, row)
            for (int j = 0; j < nb; j++) {
                const Cand o = sm.buf[q][j];
                if (o.score == e.score && (o.label < e.label || (o.label == e.label && o.row < e.row))) rank++;
            }
        }
        consumer_bar_sync();
        if (have && sub == 0 && rank < kprime) sm.list[q][rank] = e;
        consumer_bar_sync();
        if (tid == 0) {
            int new_len = min(total, kprime);
            sm.list_len[q] = new_len;
            sm.cnt[q] = 0;
            sm.thr[q] = new_len == kprime ? sm.list[q][kprime - 1].score : sm.floor[q];
        }
    }
    consumer_bar_sync();
}

template <int QT>
__global__ void __launch_bounds__(kScanThreads, 1)
scan_topk_f16_kernel(const __half *__restrict__ corpus, const uint64_t *__restrict__ labels,
                     uint32_t n_rows, const float *__restrict__ queries, int kprime,
                     Cand *__restrict__ partials, uint32_t *__restrict__ chunk_counter,
                     uint32_t *__restrict__ status, float score_floor) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    ScanSmem<QT> &sm = *reinterpret_cast<ScanSmem<QT> *>(smem_raw);
    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;

    if (tid == 0) {
        SCAN_TRACE(0);
        for (int s = 0; s < kStages; s++) {
            mbar_init(&sm.full_bar[s], 1);
            mbar_init(&sm.empty_bar[s], 1);
        }
        sm.done_warps = 0;
        sm.overflow = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < QT) {
        sm.cnt[tid] = 0;
        sm.list_len[tid] = 0;
        sm.thr[tid] = score_floor;
        sm.floor[tid] = score_floor;
    }
    __syncthreads();
    if (tid == 0) SCAN_TRACE(1);

    if (warp == kConsumerWarps) {
        // ===================== producer =====================
        if (lane != 0) return;
        uint32_t g = 0;
        uint32_t next = atomicAdd(chunk_counter, 1u);
        SCAN_TRACE(2);
        while (true) {
            const uint32_t chunk = next;
            const uint64_t row0 = (uint64_t)chunk * kChunkRows;
            if (row0 >= n_rows) break;
            next = atomicAdd(chunk_counter, 1u);  // claimed early: latency hides behind this chunk
            const uint32_t rows_left = n_rows - (uint32_t)row0;
            const uint32_t n_stage = min((uint32_t)kChunkStages, (rows_left + kStageRows - 1) / kStageRows);
            for (uint32_t s = 0; s < n_stage; s++, g++) {
                const uint32_t slot = g % kStages;
                mbar_wait(&sm.empty_bar[slot], ((g / kStages) & 1u) ^ 1u);
                const uint32_t rb = (uint32_t)row0 + s * kStageRows;
                const uint32_t nr = min((uint32_t)kStageRows, n_rows - rb);
                publish_stage(&sm.meta[slot], rb, nr, g);
                mbar_arrive_expect_tx(&sm.full_bar[slot], nr * kRowBytesF16);
                bulk_g2s(sm.ring[slot], corpus + (size_t)rb * kDim, nr * kRowBytesF16,
                         &sm.full_bar[slot]);
            }
        }
        SCAN_TRACE(3);
        for (int w = 0; w < kConsumerWarps; w++, g++) {  // one end-of-stream stage per warp
            const uint32_t slot = g % kStages;
            mbar_wait(&sm.empty_bar[slot], ((g / kStages) & 1u) ^ 1u);
            publish_stage(&sm.meta[slot], 0u, 0u, g);
            mbar_arrive(&sm.full_bar[slot]);
        }
        return;
    }

    // ===================== consumers =====================
    constexpr int V = kStageRows * QT;                 // values reduced per warp-iteration
    constexpr int LOGV = (V == 8) ? 3 : (V == 16) ? 4 : 5;
    constexpr int REP = 32 / V;                        // lanes holding the same finished value
    const int my_idx = lane >> (5 - LOGV);
    const int my_r = my_idx / QT;
    const int my_q = my_idx % QT;
    const bool is_rep = (lane & (REP - 1)) == 0;

    // Query columns owned by this lane: [8*lane, 8*lane+8) and [256+4*lane, 256+4*lane+4).
    float qa[QT][8], qb[QT][4];
#pragma unroll
    for (int q = 0; q < QT; q++) {
        const float4 *qp = reinterpret_cast<const float4 *>(queries + (size_t)q * kDim);
        float4 t0 = qp[lane * 2], t1 = qp[lane * 2 + 1], t2 = qp[64 + lane];
        qa[q][0] = t0.x; qa[q][1] = t0.y; qa[q][2] = t0.z; qa[q][3] = t0.w;
        qa[q][4] = t1.x; qa[q][5] = t1.y; qa[q][6] = t1.z; qa[q][7] = t1.w;
        qb[q][0] = t2.x; qb[q][1] = t2.y; qb[q][2] = t2.z; qb[q][3] = t2.w;
    }
    uint32_t qmask[QT];
#pragma unroll
    for (int q = 0; q < QT; q++) qmask[q] = __ballot_sync(0xffffffffu, my_q == q && is_rep);

    float thr = score_floor;
    uint32_t g = warp;
    if (tid == 0) SCAN_TRACE(4);
    while (true) {
        // Join a prune if any query's buffer reached the high-water mark -- seen here or while waiting for data.
        const uint32_t slot = g % kStages;
        bool landed;
        {
            uint32_t c = lane < QT ? ((volatile uint32_t *)sm.cnt)[lane] : 0u;
            landed = !__any_sync(0xffffffffu, c >= (uint32_t)kHighWater) &&
                     wait_stage_or_prune<QT>(&sm.full_bar[slot], (g / kStages) & 1u, sm.cnt, lane);
        }
        if (!landed) {
#ifdef DAWN_SCAN_TRACE
            unsigned long long tp0, tp1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tp0));
#endif
            consumer_bar_sync();
            prune<QT>(sm, kprime, tid);
            thr = ((volatile float *)sm.thr)[my_q];
#ifdef DAWN_SCAN_TRACE
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tp1));
            if (tid == 0) { SCAN_TRACE_ADD(9, tp1 - tp0); SCAN_TRACE_ADD(10, 1ull); }
#endif
            continue;  // stage g has not been consumed: wait for it again
        }
        if (g == 0 && lane == 0) SCAN_TRACE(5);
        const uint32_t row_base = sm.meta[slot].row_base;
        const uint32_t n_stage_rows = sm.meta[slot].n_rows;
        if (n_stage_rows == 0) break;

        const uint8_t *sp = sm.ring[slot];
        float v[V];
#pragma unroll
        for (int r = 0; r < kStageRows; r++) {
            const uint4 a = *reinterpret_cast<const uint4 *>(sp + r * kRowBytesF16 + lane * 16);
            const uint2 b = *reinterpret_cast<const uint2 *>(sp + r * kRowBytesF16 + 512 + lane * 8);
            const float2 x0 = __half22float2(*reinterpret_cast<const __half2 *>(&a.x));
            const float2 x1 = __half22float2(*reinterpret_cast<const __half2 *>(&a.y));
            const float2 x2 = __half22float2(*reinterpret_cast<const __half2 *>(&a.z));
            const float2 x3 = __half22float2(*reinterpret_cast<const __half2 *>(&a.w));
            const float2 x4 = __half22float2(*reinterpret_cast<const __half2 *>(&b.x));
            const float2 x5 = __half22float2(*reinterpret_cast<const __half2 *>(&b.y));
#pragma unroll
            for (int q = 0; q < QT; q++) {
                float acc = x0.x * qa[q][0];
                acc = fmaf(x0.y, qa[q][1], acc);
                acc = fmaf(x1.x, qa[q][2], acc);
                acc = fmaf(x1.y, qa[q][3], acc);
                acc = fmaf(x2.x, qa[q][4], acc);
                acc = fmaf(x2.y, qa[q][5], acc);
                acc = fmaf(x3.x, qa[q][6], acc);
                acc = fmaf(x3.y, qa[q][7], acc);
                acc = fmaf(x4.x, qb[q][0], acc);
                acc = fmaf(x4.y, qb[q][1], acc);
                acc = fmaf(x5.x, qb[q][2], acc);
                acc = fmaf(x5.y, qb[q][3], acc);
                v[r * QT + q] = acc;
            }
        }
        // The stage's bytes are in registers now: hand the slot back to the producer.
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty_bar[slot]);

        // Transposed butterfly: V per-lane partials -> one finished sum per lane.
#pragma unroll
        for (int s = 0; s < LOGV; s++) {
            const int off = 16 >> s;
            const int half = V >> (s + 1);
            const bool up = (lane & off) != 0;
#pragma unroll
            for (int i = 0; i < half; i++) {
                const float keep = up ? v[i + half] : v[i];
                const float send = up ? v[i] : v[i + half];
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
        }
#pragma unroll
        for (int off = 16 >> LOGV; off >= 1; off >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
        const float score = v[0];

        const bool pass = is_rep && (uint32_t)my_r < n_stage_rows && score >= thr;
        const uint32_t m = __ballot_sync(0xffffffffu, pass);
        if (m != 0) {
            const uint32_t row = row_base + my_r;
            Cand c;
            c.score = score;
            c.row = row;
            c.label = 0;
            if (pass) c.label = labels ? labels[row] : (uint64_t)row + 1;
#pragma unroll
            for (int q = 0; q < QT; q++) {
                const uint32_t mq = m & qmask[q];
                if (mq == 0) continue;
                const int leader = __ffs(mq) - 1;
                uint32_t base = 0;
                if (lane == leader) base = atomicAdd(&sm.cnt[q], (uint32_t)__popc(mq));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (pass && my_q == q) {
                    const uint32_t pos = base + __popc(mq & ((1u << lane) - 1u));
                    if (pos < (uint32_t)kBufCap) sm.buf[q][pos] = c;
                    else sm.overflow = 1;
                }
            }
        }
        g += kConsumerWarps;
    }

    // End of this warp's stream: keep joining prunes until every consumer warp is done.
    __syncwarp();
    if (tid == 0) SCAN_TRACE(6);
    if (lane == 0) atomicAdd(&sm.done_warps, 1u);
    while (true) {
        consumer_bar_sync();
        const uint32_t done = *((volatile uint32_t *)&sm.done_warps);
        prune<QT>(sm, kprime, tid);
        if (tid == 0) SCAN_TRACE_ADD(11, 1ull);
        if (done == (uint32_t)kConsumerWarps) break;
    }
    if (tid == 0) SCAN_TRACE(7);

    // One sorted, sentinel-padded list of k' candidates per query per CTA.
    for (int i = tid; i < QT * kprime; i += kConsumerThreads) {
        const int q = i / kprime, j = i % kprime;
        Cand c = j < (int)sm.list_len[q] ? sm.list[q][j] : empty_cand();
        partials[((size_t)q * gridDim.x + blockIdx.x) * kprime + j] = c;
    }
    if (tid == 0 && sm.overflow) atomicOr(status, 1u);  // cannot happen (static_assert above)
    if (tid == 0) SCAN_TRACE(8);
}


// ======================================================================================
// K4: int8-stored corpus (config C5: 500M x 384 i8, ~24 GB per GPU).
//
// Storage is blocked: 8 rows x 384 int8 followed by their 8 f32 scales (kI8BlockBytes = 3104), so
// one TMA bulk copy brings rows AND scales, and a stage is released before any prune rendezvous
// exactly as in the fp16 kernel (a warp never holds a ring slot while it waits for other warps).  Precedents in the reference: distance_i8
// (src/search/vector.rs:157-163) and ScalarKind::F8 in examples_old/search_usearch.rs:38; the
// scheme (per-row absmax/127 scale, round half away like vector.rs:30-32) is restated in
// oracle/dawn_oracle.c:dawn_oracle_store_i8.
//
// Roofline: HBM, 388 algorithmic bytes per row per pass.  With half the bytes per row the
// issue budget per row halves too, so the dot product is integer: the f32 query is split into
// two int8 vectors, q ~= s1*hi + s2*lo (residual ~3e-6 per element), and each row costs six
// dp4a per lane per query.  Integer sums are exact, the score s_row*(s1*HI + s2*LO) is a pure
// function of (row, query), and eps_q = ||q - s1*hi - s2*lo|| (Cauchy-Schwarz) bounds its
// distance from the oracle's f32 score, which finalize.cu recomputes exactly.
constexpr int kI8StageBlocks = 2;               // two 8-row blocks per stage: one 6,208-byte bulk copy
constexpr int kI8StageRows = kI8BlockRows * kI8StageBlocks;
constexpr int kI8StageBytes = kI8BlockBytes * kI8StageBlocks;
constexpr int kI8Stages = 32;                   // 32 x 6208 B = 194 KB ring
constexpr int kI8ChunkStages = 16;              // 256 rows per claim
constexpr int kI8SubRows = kI8BlockRows;        // rows reduced per butterfly
constexpr int kI8BufCap = 384;                  // a warp appends <= 16 rows per query between prune checks
static_assert(kHighWater + kConsumerWarps * kI8StageRows <= kI8BufCap, "append buffer can overflow");
static_assert(kMaxCand + kI8BufCap <= kConsumerThreads, "prune handles one entry per thread");

template <int QT>
using ScanSmemI8 = ScanSmem<QT, kI8Stages, kI8StageBytes, kI8BufCap>;

template <int QT>
__global__ void __launch_bounds__(kScanThreads, 1)
scan_topk_i8_kernel(const uint8_t *__restrict__ corpus, const uint64_t *__restrict__ labels, uint32_t n_rows,
                    const I8Query *__restrict__ queries, int kprime, Cand *__restrict__ partials,
                    uint32_t *__restrict__ chunk_counter, uint32_t *__restrict__ status,
                    const float *__restrict__ eps_q, float limit_score) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    ScanSmemI8<QT> &sm = *reinterpret_cast<ScanSmemI8<QT> *>(smem_raw);
    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < kI8Stages; s++) {
            mbar_init(&sm.full_bar[s], 1);
            mbar_init(&sm.empty_bar[s], 1);
        }
        sm.done_warps = 0;
        sm.overflow = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < QT) {
        sm.cnt[tid] = 0;
        sm.list_len[tid] = 0;
        // distance_limit pushed down: a row whose scan score is below limit_score - 2 eps_q has an exact
        // distance above the limit
        const float f = limit_score > __int_as_float(0xff800000) ? limit_score - 2.0f * eps_q[tid] - 1e-6f
                                                                  : __int_as_float(0xff800000);
        sm.thr[tid] = f;
        sm.floor[tid] = f;
    }
    __syncthreads();

    const uint32_t n_blocks = (n_rows + kI8StageRows - 1) / kI8StageRows;  // stage-sized groups of blocks
    if (warp == kConsumerWarps) {
        // ===================== producer: two 3,104-byte blocks (16 rows) per stage =====================
        if (lane != 0) return;
        uint32_t g = 0;
        uint32_t next = atomicAdd(chunk_counter, 1u);
        while (true) {
            const uint32_t chunk = next;
            const uint64_t blk0 = (uint64_t)chunk * kI8ChunkStages;
            if (blk0 >= n_blocks) break;
            next = atomicAdd(chunk_counter, 1u);
            const uint32_t n_stage = min((uint32_t)kI8ChunkStages, n_blocks - (uint32_t)blk0);
            for (uint32_t s = 0; s < n_stage; s++, g++) {
                const uint32_t slot = g % kI8Stages;
                mbar_wait(&sm.empty_bar[slot], ((g / kI8Stages) & 1u) ^ 1u);
                const uint32_t blk = (uint32_t)blk0 + s;
                const uint32_t rb = blk * kI8StageRows;
                const uint32_t nr = min((uint32_t)kI8StageRows, n_rows - rb);
                // copy only the 8-row blocks that exist (the arena is allocated in whole blocks)
                const uint32_t bytes = ((nr + kI8BlockRows - 1) / kI8BlockRows) * kI8BlockBytes;
                publish_stage(&sm.meta[slot], rb, nr, g);
                mbar_arrive_expect_tx(&sm.full_bar[slot], bytes);
                bulk_g2s(sm.ring[slot], corpus + (size_t)blk * kI8StageBytes, bytes, &sm.full_bar[slot]);
            }
        }
        for (int w = 0; w < kConsumerWarps; w++, g++) {
            const uint32_t slot = g % kI8Stages;
            mbar_wait(&sm.empty_bar[slot], ((g / kI8Stages) & 1u) ^ 1u);
            publish_stage(&sm.meta[slot], 0u, 0u, g);
            mbar_arrive(&sm.full_bar[slot]);
        }
        return;
    }

    // ===================== consumers =====================
    constexpr int V = kI8SubRows * QT * 2;              // (row, query, hi|lo) sums per butterfly
    constexpr int LOGV = (V == 16) ? 4 : 5;
    constexpr int REP = 32 / V;
    const int my_idx = lane >> (5 - LOGV);
    const int my_part = my_idx & 1;                     // 0 = hi sum, 1 = lo sum
    const int my_r = (my_idx >> 1) / QT;
    const int my_q = (my_idx >> 1) % QT;
    const bool is_rep = my_part == 0 && (lane & (REP - 1)) == 0;

    // This lane's query bytes: columns [8l, 8l+8) and [256+4l, 256+4l+4), hi and lo parts.
    int qh[QT][3], ql[QT][3];
    float s1[QT], s2[QT];
#pragma unroll
    for (int q = 0; q < QT; q++) {
        const I8Query &iq = queries[q];
        const uint2 h = *reinterpret_cast<const uint2 *>(iq.hi + lane * 8);
        const uint2 l = *reinterpret_cast<const uint2 *>(iq.lo + lane * 8);
        qh[q][0] = (int)h.x; qh[q][1] = (int)h.y;
        qh[q][2] = *reinterpret_cast<const int *>(iq.hi + 256 + lane * 4);
        ql[q][0] = (int)l.x; ql[q][1] = (int)l.y;
        ql[q][2] = *reinterpret_cast<const int *>(iq.lo + 256 + lane * 4);
        s1[q] = iq.s1;
        s2[q] = iq.s2;
    }
    float my_s1 = s1[0], my_s2 = s2[0];
#pragma unroll
    for (int q = 1; q < QT; q++)
        if (my_q == q) { my_s1 = s1[q]; my_s2 = s2[q]; }
    uint32_t qmask[QT];
#pragma unroll
    for (int q = 0; q < QT; q++) qmask[q] = __ballot_sync(0xffffffffu, my_q == q && is_rep);

    float thr = ((volatile float *)sm.thr)[my_q];
    uint32_t g = warp;
    while (true) {
        // join a prune if any query's buffer reached the high-water mark, seen here or while waiting for data
        // (no slot is held here; see wait_stage_or_prune for why the wait must be interruptible)
        const uint32_t slot = g % kI8Stages;
        bool landed;
        {
            uint32_t c = lane < QT ? ((volatile uint32_t *)sm.cnt)[lane] : 0u;
            landed = !__any_sync(0xffffffffu, c >= (uint32_t)kHighWater) &&
                     wait_stage_or_prune<QT>(&sm.full_bar[slot], (g / kI8Stages) & 1u, sm.cnt, lane);
        }
        if (!landed) {
            consumer_bar_sync();
            prune<QT>(sm, kprime, tid);
            thr = ((volatile float *)sm.thr)[my_q];
            continue;
        }
        const uint32_t row_base = sm.meta[slot].row_base;
        const uint32_t n_stage_rows = sm.meta[slot].n_rows;
        if (n_stage_rows == 0) break;
        const uint8_t *ring_slot = sm.ring[slot];
#pragma unroll 1
        for (int sub = 0; sub < kI8StageBlocks; sub++) {
            if ((uint32_t)(sub * kI8SubRows) >= n_stage_rows) {  // second block absent in a ragged last stage
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.empty_bar[slot]);
                break;
            }
            const uint8_t *sp = ring_slot + sub * kI8BlockBytes;
            int v[V];
#pragma unroll
            for (int r = 0; r < kI8SubRows; r++) {
                const uint8_t *rp = sp + r * kDim;
                const uint2 a = *reinterpret_cast<const uint2 *>(rp + lane * 8);
                const int b = *reinterpret_cast<const int *>(rp + 256 + lane * 4);
#pragma unroll
                for (int q = 0; q < QT; q++) {
                    v[(r * QT + q) * 2 + 0] = __dp4a(b, qh[q][2], __dp4a((int)a.y, qh[q][1], __dp4a((int)a.x, qh[q][0], 0)));
                    v[(r * QT + q) * 2 + 1] = __dp4a(b, ql[q][2], __dp4a((int)a.y, ql[q][1], __dp4a((int)a.x, ql[q][0], 0)));
                }
            }
            const float row_scale = *reinterpret_cast<const float *>(sp + kI8BlockRows * kDim + my_r * 4);
            if (sub == kI8StageBlocks - 1) {  // the stage's last bytes are in registers: hand the slot back
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.empty_bar[slot]);
            }
#pragma unroll
            for (int s = 0; s < LOGV; s++) {
                const int off = 16 >> s;
                const int half = V >> (s + 1);
                const bool up = (lane & off) != 0;
#pragma unroll
                for (int i = 0; i < half; i++) {
                    const int keep = up ? v[i + half] : v[i];
                    const int send = up ? v[i] : v[i + half];
                    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
            }
#pragma unroll
            for (int off = 16 >> LOGV; off >= 1; off >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
            const int other = __shfl_xor_sync(0xffffffffu, v[0], REP);  // the lane holding the other part
            const int hi = my_part == 0 ? v[0] : other;
            const int lo = my_part == 0 ? other : v[0];
            float acc = __fmul_rn(my_s1, (float)hi);
            acc = __fmaf_rn(my_s2, (float)lo, acc);
            const float score = __fmul_rn(row_scale, acc);

            const uint32_t r_in_stage = sub * kI8SubRows + my_r;
            const bool pass = is_rep && r_in_stage < n_stage_rows && score >= thr;
            const uint32_t m = __ballot_sync(0xffffffffu, pass);
            if (m != 0) {
                const uint32_t row = row_base + r_in_stage;
                Cand c;
                c.score = score;
                c.row = row;
                c.label = 0;
                if (pass) c.label = labels ? labels[row] : (uint64_t)row + 1;
#pragma unroll
                for (int q = 0; q < QT; q++) {
                    const uint32_t mq = m & qmask[q];
                    if (mq == 0) continue;
                    const int leader = __ffs(mq) - 1;
                    uint32_t base = 0;
                    if (lane == leader) base = atomicAdd(&sm.cnt[q], (uint32_t)__popc(mq));
                    base = __shfl_sync(0xffffffffu, base, leader);
                    if (pass && my_q == q) {
                        const uint32_t pos = base + __popc(mq & ((1u << lane) - 1u));
                        if (pos < (uint32_t)kI8BufCap) sm.buf[q][pos] = c;
                        else sm.overflow = 1;
                    }
                }
            }
        }
        g += kConsumerWarps;
    }

    __syncwarp();
    if (lane == 0) atomicAdd(&sm.done_warps, 1u);
    while (true) {
        consumer_bar_sync();
        const uint32_t done = *((volatile uint32_t *)&sm.done_warps);
        prune<QT>(sm, kprime, tid);
        if (done == (uint32_t)kConsumerWarps) break;
    }
    for (int i = tid; i < QT * kprime; i += kConsumerThreads) {
        const int q = i / kprime, j = i % kprime;
        Cand c = j < (int)sm.list_len[q] ? sm.list[q][j] : empty_cand();
        partials[((size_t)q * gridDim.x + blockIdx.x) * kprime + j] = c;
    }
    if (tid == 0 && sm.overflow) atomicOr(status, 1u);
}

// f32 query -> two int8 vectors + scales + eps (one 128-thread block per query).
__global__ void __launch_bounds__(128) prep_queries_i8_kernel(const float *__restrict__ q32, int n_queries,
                                                              I8Query *__restrict__ out, float *__restrict__ eps_q) {
    const int qi = blockIdx.x;
    if (qi >= n_queries) return;
    __shared__ float red[4];
    const float *q = q32 + (size_t)qi * kDim;
    auto block_max = [&](float x) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, off));
        __syncthreads();
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = x;
        __syncthreads();
        return fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    };
    float x[3], amax = 0.f;
#pragma unroll
    for (int j = 0; j < 3; j++) {
        x[j] = q[threadIdx.x + 128 * j];
        amax = fmaxf(amax, fabsf(x[j]));
    }
    amax = block_max(amax);
    const float s1 = amax > 0.f ? amax / 127.0f : 1.0f;
    float r[3], rmax = 0.f;
    int hi[3];
#pragma unroll
    for (int j = 0; j < 3; j++) {
        hi[j] = max(-127, min(127, (int)rintf(x[j] / s1)));
        r[j] = fmaf(-s1, (float)hi[j], x[j]);
        rmax = fmaxf(rmax, fabsf(r[j]));
    }
    rmax = block_max(rmax);
    const float s2 = rmax > 0.f ? rmax / 127.0f : 1.0f;
    float err2 = 0.f;
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const int lo = max(-127, min(127, (int)rintf(r[j] / s2)));
        const float e = fmaf(-s2, (float)lo, r[j]);
        err2 += e * e;
        out[qi].hi[threadIdx.x + 128 * j] = (int8_t)hi[j];
        out[qi].lo[threadIdx.x + 128 * j] = (int8_t)lo;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) err2 += __shfl_xor_sync(0xffffffffu, err2, off);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = err2;
    __syncthreads();
    if (threadIdx.x == 0) {
        out[qi].s1 = s1;
        out[qi].s2 = s2;
        // |sum e_i x_i| <= ||e|| * ||x||; dequantised rows have norm < 1.02; + f32 rounding of the
        // three-operation score formula
        eps_q[qi] = sqrtf(red[0] + red[1] + red[2] + red[3]) * 1.03f + 3.0e-5f;  // 384 * 2^-24 * 1.03 = 2.4e-5 for the sequential re-score alone
    }
}

template <int QT>
cudaError_t launch_i8_qt(const ScanLaunchI8 &p, cudaStream_t s) {
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    const size_t smem = sizeof(ScanSmemI8<QT>);
    if (dev < 64 && !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(scan_topk_i8_kernel<QT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    scan_topk_i8_kernel<QT><<<p.grid, kScanThreads, smem, s>>>(p.corpus, p.labels, p.n_rows, p.queries, p.kprime,
                                                              p.partials, p.chunk_counter, p.status, p.eps_q,
                                                              p.limit_score);
    return cudaGetLastError();
}

template <int QT>
cudaError_t launch_qt(const ScanLaunch &p, cudaStream_t s) {
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    const size_t smem = sizeof(ScanSmem<QT>);
    if (dev < 64 && !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(scan_topk_f16_kernel<QT>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    scan_topk_f16_kernel<QT><<<p.grid, kScanThreads, smem, s>>>(
        p.corpus, p.labels, p.n_rows, p.queries, p.kprime, p.partials, p.chunk_counter, p.status, p.score_floor);
    return cudaGetLastError();
}

}  // namespace

int scan_max_queries_per_pass(int /*kprime*/) { return 4; }

cudaError_t launch_scan_topk_f16(const ScanLaunch &p, cudaStream_t s) {
    if (p.kprime < 1 || p.kprime > kMaxCand) return cudaErrorInvalidValue;
    switch (p.nq) {
        case 1: return launch_qt<1>(p, s);
        case 2: return launch_qt<2>(p, s);
        case 4: return launch_qt<4>(p, s);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_prep_queries_i8(const float *q32, int n_queries, I8Query *out, float *eps_q, cudaStream_t s) {
    if (n_queries <= 0) return cudaSuccess;
    prep_queries_i8_kernel<<<n_queries, 128, 0, s>>>(q32, n_queries, out, eps_q);
    return cudaGetLastError();
}

cudaError_t launch_scan_topk_i8(const ScanLaunchI8 &p, cudaStream_t s) {
    if (p.kprime < 1 || p.kprime > kMaxCand) return cudaErrorInvalidValue;
    switch (p.nq) {
        case 1: return launch_i8_qt<1>(p, s);
        case 2: return launch_i8_qt<2>(p, s);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace dawn
