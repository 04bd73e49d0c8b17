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

struct StageMeta {
    uint32_t row_base;
    uint32_t n_rows;  // 0 = end-of-stream sentinel
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

// ---- shared-memory layout -------------------------------------------------------------
template <int QT>
struct ScanSmem {
    alignas(128) uint8_t ring[kStages][kStageBytes];
    alignas(16) Cand list[QT][kMaxCand];
    alignas(16) Cand buf[QT][kBufCap];
    alignas(8) uint64_t full_bar[kStages];
    alignas(8) uint64_t empty_bar[kStages];
    StageMeta meta[kStages];
    uint32_t cnt[QT];
    uint32_t list_len[QT];
    float thr[QT];
    uint32_t done_warps;
    uint32_t overflow;
};

// Merge the append buffer into the sorted list, all 512 consumer threads, one query at a time.
// Rank of an entry = number of entries that beat it; the list part is already sorted so only
// the buffer needs counting.  Entries are unique under cand_better, so ranks are unique.
template <int QT>
__device__ __forceinline__ void prune(ScanSmem<QT> &sm, int kprime, int tid) {
#pragma unroll 1
    for (int q = 0; q < QT; q++) {
        const int nb = min((int)((volatile uint32_t *)sm.cnt)[q], kBufCap);
        const int len = (int)((volatile uint32_t *)sm.list_len)[q];
        const int total = len + nb;
        const bool have = tid < total;
        Cand e = empty_cand();
        int rank = 0;
        if (have && nb > 0) {
            if (tid < len) {
                e = sm.list[q][tid];
                rank = tid;
            } else {
                e = sm.buf[q][tid - len];
                int lo = 0, hi = len;
                while (lo < hi) {
                    int mid = (lo + hi) >> 1;
                    if (cand_better(sm.list[q][mid], e)) lo = mid + 1;
                    else hi = mid;
                }
                rank = lo;
            }
            for (int j = 0; j < nb; j++) rank += cand_better(sm.buf[q][j], e) ? 1 : 0;
        }
        consumer_bar_sync();
        if (have && nb > 0 && rank < kprime) sm.list[q][rank] = e;
        consumer_bar_sync();
        if (tid == 0) {
            int new_len = min(total, kprime);
            sm.list_len[q] = new_len;
            sm.cnt[q] = 0;
            sm.thr[q] = new_len == kprime ? sm.list[q][kprime - 1].score : __int_as_float(0xff800000);
        }
    }
    consumer_bar_sync();
}

template <int QT>
__global__ void __launch_bounds__(kScanThreads, 1)
scan_topk_f16_kernel(const __half *__restrict__ corpus, const uint64_t *__restrict__ labels,
                     uint32_t n_rows, const float *__restrict__ queries, int kprime,
                     Cand *__restrict__ partials, uint32_t *__restrict__ chunk_counter,
                     uint32_t *__restrict__ status) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    ScanSmem<QT> &sm = *reinterpret_cast<ScanSmem<QT> *>(smem_raw);
    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;

    if (tid == 0) {
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
        sm.thr[tid] = __int_as_float(0xff800000);
    }
    __syncthreads();

    if (warp == kConsumerWarps) {
        // ===================== producer =====================
        if (lane != 0) return;
        uint32_t g = 0;
        uint32_t next = atomicAdd(chunk_counter, 1u);
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
                sm.meta[slot].row_base = rb;
                sm.meta[slot].n_rows = nr;
                mbar_arrive_expect_tx(&sm.full_bar[slot], nr * kRowBytesF16);
                bulk_g2s(sm.ring[slot], corpus + (size_t)rb * kDim, nr * kRowBytesF16,
                         &sm.full_bar[slot]);
            }
        }
        for (int w = 0; w < kConsumerWarps; w++, g++) {  // one end-of-stream stage per warp
            const uint32_t slot = g % kStages;
            mbar_wait(&sm.empty_bar[slot], ((g / kStages) & 1u) ^ 1u);
            sm.meta[slot].row_base = 0;
            sm.meta[slot].n_rows = 0;
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

    float thr = __int_as_float(0xff800000);
    uint32_t g = warp;
    while (true) {
        // Join a prune if any query's buffer reached the high-water mark.
        {
            uint32_t c = lane < QT ? ((volatile uint32_t *)sm.cnt)[lane] : 0u;
            if (__any_sync(0xffffffffu, c >= (uint32_t)kHighWater)) {
                consumer_bar_sync();
                prune<QT>(sm, kprime, tid);
                thr = ((volatile float *)sm.thr)[my_q];
                continue;
            }
        }
        const uint32_t slot = g % kStages;
        mbar_wait(&sm.full_bar[slot], (g / kStages) & 1u);
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
    if (lane == 0) atomicAdd(&sm.done_warps, 1u);
    while (true) {
        consumer_bar_sync();
        const uint32_t done = *((volatile uint32_t *)&sm.done_warps);
        prune<QT>(sm, kprime, tid);
        if (done == (uint32_t)kConsumerWarps) break;
    }

    // One sorted, sentinel-padded list of k' candidates per query per CTA.
    for (int i = tid; i < QT * kprime; i += kConsumerThreads) {
        const int q = i / kprime, j = i % kprime;
        Cand c = j < (int)sm.list_len[q] ? sm.list[q][j] : empty_cand();
        partials[((size_t)q * gridDim.x + blockIdx.x) * kprime + j] = c;
    }
    if (tid == 0 && sm.overflow) atomicOr(status, 1u);  // cannot happen (static_assert above)
}

template <int QT>
cudaError_t launch_qt(const ScanLaunch &p, cudaStream_t s) {
    static bool configured = false;
    const size_t smem = sizeof(ScanSmem<QT>);
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(scan_topk_f16_kernel<QT>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    scan_topk_f16_kernel<QT><<<p.grid, kScanThreads, smem, s>>>(
        p.corpus, p.labels, p.n_rows, p.queries, p.kprime, p.partials, p.chunk_counter, p.status);
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

}  // namespace dawn
