// gemm_topk.cu -- K3: large-batch path.  queries x corpus is a dense contraction, run on the
// 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM) with the top-k selection
// fused into the epilogue, so raw scores never reach HBM.
//
// Replaces usearch's Index::search (/root/reference/src/search/search_provider.rs:214) for a
// batch of queries (the batching front-end of SURVEY.md section 8f-1 is what produces batches).
//
// Roofline: tensor pipe for batch >~ 256 (2*B*N*384 flop), HBM below that (the corpus is read
// once per 128-query tile: N*768 bytes).
//
// Shape of one CTA (256 threads, 1 CTA/SM, persistent over work units):
//   A operand  = one 128-query tile, fp16, K-major, resident in shared memory (6 k-blocks of
//                128 rows x 128 B, TMA SWIZZLE_128B)                                   96 KB
//   B operand  = corpus tiles of 256 rows, streamed k-block by k-block (256 rows x 128 B =
//                32 KB per stage) through a 3-stage TMA/mbarrier ring                  96 KB
//   D          = 128 x 256 f32 in TMEM, two buffers (512 columns) so the epilogue of tile i
//                overlaps the MMAs of tile i+1
//   warp 0     TMA producer (one lane)         warp 1   MMA issuer (one lane) + TMEM alloc
//   warps 4-7  epilogue: thread = one query (TMEM lane); tcgen05.ld 32 columns at a time,
//              max-reduce, compare with the query's threshold (a register); survivors (rare)
//              are appended to the query's candidate log in global memory.
//
// Selection across the corpus runs in geometrically growing ROUNDS of tiles (4 tiles = 1024 rows, x8, ...;
// tiles are visited in a strided permutation so each round is a uniform sample of the corpus):
// round 0 logs everything, select_topk_kernel then keeps the best k' per query and publishes
// the k'-th score as the threshold for the next round, so a round appends ~7k' candidates per
// query regardless of its size.  The log is a superset of the top-k' under the fp16-query
// scores; finalize.cu re-scores the k' survivors exactly and certifies the result with
// eps_q = ||q - fp16(q)|| + accumulation slack (Cauchy-Schwarz, rows have norm <= 1.01).
#include <cuda.h>

#include <cmath>

#include "dawn_common.cuh"
#include "gemm_pipe.cuh"

namespace dawn {

namespace {

constexpr int kGemmThreads = 256;
constexpr int BM = 128;      // queries per tile (UMMA M)
constexpr int BN = 256;      // corpus rows per tile (UMMA N)
constexpr int BK = 64;       // fp16 elements per k-block = 128 B = one swizzle atom row
constexpr int kKBlocks = kDim / BK;  // 6
constexpr int kUmmaK = 16;
constexpr int kMaxStagesB = 6;              // 3 x 32 KB (one CTA) or 6 x 16 KB (CTA pair)
constexpr int kBRingBytes = 3 * BN * BK * 2;  // 98304 either way
constexpr int kABlockBytes = BM * BK * 2;   // 16384
constexpr int kABytes = kABlockBytes * kKBlocks;  // 98304
constexpr int kTmemCols = 512;
constexpr int kStageCap = 16;  // survivors a query thread parks in shared memory before one atomic flush
constexpr int kStagingBytes = kStageCap * 128 * 8;  // 16 KB
constexpr int kSmemBytes = 1024 + kABytes + kBRingBytes + 256 + kStagingBytes;

using namespace pipe;

// Instruction descriptor: D=f32, A=B=f16, both K-major, N>>3 at bit 17, M>>4 at bit 24
// (M = 128 for one CTA, 256 for a CTA pair).
__host__ __device__ constexpr uint32_t make_idesc(int m) { return (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }

struct GemmSmem {  // offsets from the 1024-aligned base
    static constexpr int a_off = 0;
    static constexpr int b_off = kABytes;
    static constexpr int bar_off = kABytes + kBRingBytes;
    static constexpr int staging_off = bar_off + 256;
    // barriers (8 B each): full[6], empty[6], tmem_full[2], tmem_empty[2], a_full, a_free ; then tmem ptr
};

// Move a query thread's parked survivors to its global candidate log: one atomic reserves the
// slots, the copies are plain stores.  stage is [kStageCap][128] uint2, column = epilogue thread.
__device__ __noinline__ void flush_staged(uint32_t stage_smem, int col, uint32_t n, uint2 *__restrict__ log_q,
                                          uint32_t *__restrict__ cnt_q, uint32_t *__restrict__ overflow_q, int cap) {
    uint32_t slot = atomicAdd(cnt_q, n);
    for (uint32_t e = 0; e < n; e++, slot++) {
        uint2 val;
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(val.x), "=r"(val.y) : "r"(stage_smem + (e * 128 + col) * 8));
        if (slot < (uint32_t)cap) log_q[slot] = val;
        else *overflow_q = 1u;
    }
}

// CG = 1: one CTA per tile (M = 128 queries).  CG = 2: a CTA pair shares every MMA (cta_group::2,
// M = 256 queries, each CTA holds 128 of them and streams only HALF of each 256-row corpus tile),
// which halves the L2->shared traffic per SM -- the bound of the one-CTA kernel at large batch.
template <int CG>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_topk_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_x,
                 uint32_t tile_begin, uint32_t tile_end, uint32_t n_rows, uint32_t n_tiles_total, uint32_t perm_mult,
                 int n_qtiles, int n_queries,
                 int chunk_tiles, const float *__restrict__ thr_g, uint32_t *__restrict__ cnt_g,
                 uint2 *__restrict__ log_g, uint32_t *__restrict__ overflow_g, int log_cap, uint32_t *__restrict__ chunk_arrive) {
    constexpr int kStagesB = 3 * CG;
    constexpr int kBRows = BN / CG;                 // corpus rows this CTA streams per tile
    constexpr int kBStageBytes = kBRows * BK * 2;   // 32 KB or 16 KB
    constexpr uint32_t kIdesc = make_idesc(BM * CG);
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_smem = base + GemmSmem::a_off;
    const uint32_t b_smem = base + GemmSmem::b_off;
    const uint32_t bars = base + GemmSmem::bar_off;
    auto full_bar = [&](int s) { return bars + 8 * s; };
    auto empty_bar = [&](int s) { return bars + 8 * (kMaxStagesB + s); };
    auto tfull_bar = [&](int a) { return bars + 8 * (2 * kMaxStagesB + a); };
    auto tempty_bar = [&](int a) { return bars + 8 * (2 * kMaxStagesB + 2 + a); };
    const uint32_t a_full_bar = bars + 8 * (2 * kMaxStagesB + 4);
    const uint32_t a_free_bar = bars + 8 * (2 * kMaxStagesB + 5);
    const uint32_t tmem_ptr_smem = bars + 8 * (2 * kMaxStagesB + 6);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0u;  // 0 = leader (issues the MMAs)
    const uint32_t unit_first = blockIdx.x / CG;                 // CTA pairs walk the units together
    const uint32_t unit_stride = gridDim.x / CG;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStagesB; s++) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; a++) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 4 * CG);  // one arrival per epilogue warp (of both CTAs)
        }
        mbar_init(a_full_bar, 1);
        mbar_init(a_free_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM allocation: one full warp (the same warp in both CTAs of a pair), all 512 columns
        if constexpr (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_smem),
                         "r"((uint32_t)kTmemCols));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_smem),
                         "r"((uint32_t)kTmemCols));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
    }
    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all();
    else __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

    // Work units: (query tile t, chunk of corpus tiles); unit u -> t = u % n_qtiles (fastest, so
    // CTAs that stream the same rows run side by side and share them through L2).
    // Tiles are visited in a strided permutation of the corpus (phys = (i * perm_mult) % n_tiles_total,
    // perm_mult coprime to n_tiles_total): every round is a uniform sample of the whole corpus, so the
    // thresholds learnt in early rounds are representative whatever order the pages were stored in.
    const uint32_t n_tiles = tile_end - tile_begin;
    // (the 64-bit modulo is taken once per work unit; inside a unit the physical tile advances by perm_mult mod total)
    auto phys_tile = [&](uint32_t tile) { return (uint32_t)(((uint64_t)(tile_begin + tile) * perm_mult) % n_tiles_total); };
    auto phys_next = [&](uint32_t pt) {
        const uint32_t nx = pt + perm_mult;  // perm_mult < n_tiles_total <= 2^24: no overflow
        return nx >= n_tiles_total ? nx - n_tiles_total : nx;
    };
    const uint32_t n_chunks = (n_tiles + chunk_tiles - 1) / chunk_tiles;
    const uint32_t n_units = n_chunks * (uint32_t)n_qtiles;

    if (warp == 0 && lane == 0) {
        // ===================== TMA producer (both CTAs of a pair) =====================
        // Pair mode: every load signals the LEADER's barrier; only the leader posts the expected
        // byte count (for both CTAs); each CTA waits for its own copy of the empty barriers, which
        // the MMA commit multicasts.
        uint32_t g = 0;         // B stage counter
        uint32_t n_reload = 0;  // A (query tile) loads issued by this CTA
        int cur_t = -1;
        for (uint32_t u = unit_first; u < n_units; u += unit_stride) {
            const int t = (int)(u % (uint32_t)n_qtiles);
            const uint32_t chunk = u / (uint32_t)n_qtiles;
            if (t != cur_t) {
                // the MMAs that read the previous query tile must have retired (a_free is committed
                // by the MMA warp exactly when the next unit needs a different tile)
                if (n_reload > 0) mbar_wait(a_free_bar, (n_reload - 1) & 1u);
                const int qrow = t * BM * CG + (int)cta_rank * BM;
                if constexpr (CG == 2) {
                    if (cta_rank == 0) mbar_expect_tx(a_full_bar, 2 * kABytes);
                    const uint32_t lead = mapa_rank(a_full_bar, 0);
                    for (int kb = 0; kb < kKBlocks; kb++)
                        tma_load_2d_pair(a_smem + kb * kABlockBytes, &tmap_q, kb * BK, qrow, lead);
                } else {
                    mbar_expect_tx(a_full_bar, kABytes);
                    for (int kb = 0; kb < kKBlocks; kb++)
                        tma_load_2d(a_smem + kb * kABlockBytes, &tmap_q, kb * BK, qrow, a_full_bar);
                }
                cur_t = t;
                n_reload++;
            }
            if (chunk_arrive != nullptr && chunk < (uint32_t)kArriveSlots)  // start the chunk together with its other query tiles
                chunk_rendezvous(chunk_arrive + chunk, (uint32_t)n_qtiles, cta_rank == 0);
            const uint32_t tile0 = chunk * chunk_tiles;
            const uint32_t tile1 = min(n_tiles, tile0 + chunk_tiles);
            uint32_t pt = phys_tile(tile0);
            for (uint32_t tile = tile0; tile < tile1; tile++, pt = phys_next(pt)) {
                const int row0 = (int)(pt * (uint32_t)BN) + (int)cta_rank * kBRows;
                for (int kb = 0; kb < kKBlocks; kb++, g++) {
                    const uint32_t s = g % kStagesB;
                    mbar_wait(empty_bar(s), ((g / kStagesB) & 1u) ^ 1u);
                    if constexpr (CG == 2) {
                        if (cta_rank == 0) mbar_expect_tx(full_bar(s), 2 * kBStageBytes);
                        tma_load_2d_pair(b_smem + s * kBStageBytes, &tmap_x, kb * BK, row0, mapa_rank(full_bar(s), 0));
                    } else {
                        mbar_expect_tx(full_bar(s), kBStageBytes);
                        tma_load_2d(b_smem + s * kBStageBytes, &tmap_x, kb * BK, row0, full_bar(s));
                    }
                }
            }
        }
    } else if (warp == 1 && lane == 0 && cta_rank == 0) {
        // ===================== MMA issuer (leader CTA only) =====================
        uint32_t g = 0, tile_ctr = 0, a_loads = 0;
        int cur_t = -1;
        for (uint32_t u = unit_first; u < n_units; u += unit_stride) {
            const int t = (int)(u % (uint32_t)n_qtiles);
            const uint32_t chunk = u / (uint32_t)n_qtiles;
            if (t != cur_t) {
                mbar_wait(a_full_bar, a_loads & 1u);
                a_loads++;
                cur_t = t;
            }
            const uint32_t tile0 = chunk * chunk_tiles;
            const uint32_t tile1 = min(n_tiles, tile0 + chunk_tiles);
            for (uint32_t tile = tile0; tile < tile1; tile++, tile_ctr++) {
                const uint32_t acc = tile_ctr & 1u;
                mbar_wait(tempty_bar(acc), ((tile_ctr >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < kKBlocks; kb++, g++) {
                    const uint32_t s = g % kStagesB;
                    mbar_wait(full_bar(s), (g / kStagesB) & 1u);
                    tc_fence_after();
                    const uint64_t adesc = make_kmajor_sw128_desc(a_smem + kb * kABlockBytes);
                    const uint64_t bdesc = make_kmajor_sw128_desc(b_smem + s * kBStageBytes);
#pragma unroll
                    for (int k = 0; k < BK / kUmmaK; k++) {
                        // advance 16 elements = 32 B inside the swizzle atom: +2 in 16-byte units
                        if constexpr (CG == 2)
                            tc_mma_f16_pair(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), kIdesc,
                                            (uint32_t)((kb | k) != 0));
                        else
                            tc_mma_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), kIdesc,
                                       (uint32_t)((kb | k) != 0));
                    }
                    if constexpr (CG == 2) tc_commit_pair(empty_bar(s));  // stage free in both CTAs
                    else tc_commit(empty_bar(s));
                }
                if constexpr (CG == 2) tc_commit_pair(tfull_bar(acc));  // accumulators ready in both CTAs
                else tc_commit(tfull_bar(acc));
            }
            const uint32_t u_next = u + unit_stride;
            if (u_next < n_units && (int)(u_next % (uint32_t)n_qtiles) != t) {
                if constexpr (CG == 2) tc_commit_pair(a_free_bar);
                else tc_commit(a_free_bar);
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: threshold filter =====================
        const int quarter = warp & 3;  // TMEM lane quarter this warp may read
        const int col = quarter * 32 + lane;
        const uint32_t stage_smem = base + GemmSmem::staging_off;
        uint32_t n_st = 0;
        uint32_t tile_ctr = 0;
        const uint32_t tempty_lead0 = CG == 2 ? mapa_rank(tempty_bar(0), 0) : 0u;  // the MMA issuer waits on the leader's
        const uint32_t tempty_lead1 = CG == 2 ? mapa_rank(tempty_bar(1), 0) : 0u;
        for (uint32_t u = unit_first; u < n_units; u += unit_stride) {
            const int t = (int)(u % (uint32_t)n_qtiles);
            const uint32_t chunk = u / (uint32_t)n_qtiles;
            const int q = t * BM * CG + (int)cta_rank * BM + quarter * 32 + lane;
            const bool q_valid = q < n_queries;
            const float thr = q_valid ? thr_g[q] : __int_as_float(0x7f800000);
            uint2 *log_q = log_g + (size_t)q * log_cap;
            const uint32_t tile0 = chunk * chunk_tiles;
            const uint32_t tile1 = min(n_tiles, tile0 + chunk_tiles);
            uint32_t pt = phys_tile(tile0);
            for (uint32_t tile = tile0; tile < tile1; tile++, tile_ctr++, pt = phys_next(pt)) {
                const uint32_t acc = tile_ctr & 1u;
                const uint32_t row0 = pt * (uint32_t)BN;
                mbar_wait(tfull_bar(acc), (tile_ctr >> 1) & 1u);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN;
                const int lim_tile = (int)n_rows - (int)row0;  // valid columns of this tile (the last physical tile is ragged)
                // One 32-column chunk: 4 group maxima -> overall max; only groups that reach the
                // threshold are examined element by element.  Survivors are parked in shared memory.
                auto process = [&](const uint32_t (&v)[32], int c) {
                    float g[4];
#pragma unroll
                    for (int gi = 0; gi < 4; gi++) {
                        float m0 = fmaxf(__uint_as_float(v[8 * gi]), __uint_as_float(v[8 * gi + 1]));
                        float m1 = fmaxf(__uint_as_float(v[8 * gi + 2]), __uint_as_float(v[8 * gi + 3]));
                        float m2 = fmaxf(__uint_as_float(v[8 * gi + 4]), __uint_as_float(v[8 * gi + 5]));
                        float m3 = fmaxf(__uint_as_float(v[8 * gi + 6]), __uint_as_float(v[8 * gi + 7]));
                        g[gi] = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
                    }
                    const float m = fmaxf(fmaxf(g[0], g[1]), fmaxf(g[2], g[3]));
                    if (q_valid && m >= thr) {
#pragma unroll
                        for (int gi = 0; gi < 4; gi++) {
                            if (g[gi] >= thr) {
#pragma unroll
                                for (int j = 0; j < 8; j++) {
                                    const int i = 8 * gi + j;
                                    if (__uint_as_float(v[i]) >= thr && c * 32 + i < lim_tile) {
                                        asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(stage_smem + (n_st * 128 + col) * 8),
                                                     "r"(v[i]), "r"(row0 + c * 32 + i)
                                                     : "memory");
                                        if (++n_st == (uint32_t)kStageCap) {
                                            flush_staged(stage_smem, col, n_st, log_q, cnt_g + q, overflow_g + q, log_cap);
                                            n_st = 0;
                                        }
                                    }
                                }
                            }
                        }
                    }
                    __syncwarp();
                };
                // software pipeline: the next chunk's TMEM load is in flight while this one is filtered
                uint32_t v0[32], v1[32];
                tc_ld_32x32b_x32(taddr, v0);
#pragma unroll 1
                for (int c = 0; c < BN / 32; c += 2) {
                    tc_wait_ld();
                    tc_ld_32x32b_x32(taddr + (c + 1) * 32, v1);
                    process(v0, c);
                    tc_wait_ld();
                    if (c + 2 < BN / 32) tc_ld_32x32b_x32(taddr + (c + 2) * 32, v0);
                    process(v1, c + 1);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if constexpr (CG == 2) mbar_arrive_cluster(acc ? tempty_lead1 : tempty_lead0);
                    else mbar_arrive(tempty_bar(acc));
                }
                // Flush together: when any lane's park is half full, every lane with survivors
                // flushes now, so the warp pays ONE atomic round trip instead of one per lane.
                if (__any_sync(0xffffffffu, n_st >= (uint32_t)(kStageCap / 2))) {
                    if (n_st) flush_staged(stage_smem, col, n_st, log_q, cnt_g + q, overflow_g + q, log_cap);
                    n_st = 0;
                    __syncwarp();
                }
            }
            if (n_st) {  // the next unit belongs to another query tile
                flush_staged(stage_smem, col, n_st, log_q, cnt_g + q, overflow_g + q, log_cap);
                n_st = 0;
            }
            __syncwarp();
        }
    }

    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all();  // the peer may still be reading this CTA's operands / barriers
    else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        __syncwarp();
        if constexpr (CG == 2)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kTmemCols));
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kTmemCols));
    }
}

// ---- query preparation: f32 -> fp16 (padded to a multiple of 128 rows) and eps_q -------------
__global__ void __launch_bounds__(128) prep_queries_kernel(const float *__restrict__ q32, int n_queries,
                                                           int n_padded, __half *__restrict__ q16,
                                                           float *__restrict__ eps_q, float accum_slack) {
    const int q = blockIdx.x;
    __shared__ float red[4];
    float err2 = 0.f;
    for (int c = threadIdx.x; c < kDim; c += blockDim.x) {
        float x = q < n_queries ? q32[(size_t)q * kDim + c] : 0.f;
        __half h = __float2half_rn(x);
        q16[(size_t)q * kDim + c] = h;
        float d = x - __half2float(h);
        err2 += d * d;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) err2 += __shfl_xor_sync(0xffffffffu, err2, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = err2;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = red[0] + red[1] + red[2] + red[3];
        // |sum (q_i - q16_i) x_i| <= ||q - q16|| * ||x||, stored rows have norm < 1.011 (the reference's
        // gate is 1.01, vector.rs:185-192, plus fp16 rounding); 1.02 also covers the f32 rounding here.
        eps_q[q] = q < n_queries ? sqrtf(s) * 1.02f + accum_slack : 0.f;
    }
    (void)n_padded;
}

// ---- select: keep the best k' log entries of a query, publish the k'-th score as threshold ----
// 64-bit radix select (8 passes over a 256-bin shared histogram) on key = (ordered score, ~row):
// descending key order == score descending, row ascending.  Keys are unique (rows are), so
// exactly k' entries satisfy key >= K.  The kept entries are written back UNSORTED.

__global__ void __launch_bounds__(kSelThreads) select_topk_kernel(uint2 *__restrict__ log_g, uint32_t *__restrict__ cnt_g,
                                                                  float *__restrict__ thr_g,
                                                                  uint32_t *__restrict__ overflow_g, int log_cap, int kp,
                                                                  const uint64_t *__restrict__ labels,
                                                                  Cand *__restrict__ final_lists,
                                                                  const float *__restrict__ eps_q, float limit_score,
                                                                  uint32_t *__restrict__ arrive) {
    __shared__ unsigned long long keys[kSelCap];
    clear_arrive_slots(arrive, blockIdx.x, gridDim.x, threadIdx.x, kSelThreads);
    __shared__ unsigned long long s_prefix;
    __shared__ uint32_t hist[256];
    __shared__ uint32_t s_need, s_out;
    const int q = blockIdx.x;
    const int tid = threadIdx.x;
    uint2 *log_q = log_g + (size_t)q * log_cap;
    const uint32_t cnt = cnt_g[q];
    const int n = (int)min(cnt, (uint32_t)log_cap);
    for (int i = tid; i < n; i += kSelThreads) {
        const uint2 e = log_q[i];
        keys[i] = ((unsigned long long)float_to_ordered(__uint_as_float(e.x)) << 32) | (unsigned long long)(~e.y);
    }
    if (tid == 0) {
        s_prefix = 0ull;
        s_need = (uint32_t)kp;
        s_out = 0u;
    }
    __syncthreads();
    // keep everything when there are fewer than k' entries
    const unsigned long long K = n >= kp ? radix_select_kth(keys, n, kp, hist, &s_prefix, &s_need, tid) : 0ull;
    // compaction (order does not matter downstream)
    for (int i = tid; i < n; i += kSelThreads) {
        const unsigned long long key = keys[i];
        if (key >= K) {
            const uint32_t pos = atomicAdd(&s_out, 1u);
            if (pos < (uint32_t)kp) {
                const float sc = ordered_to_float((uint32_t)(key >> 32));
                const uint32_t row = ~(uint32_t)key;
                log_q[pos] = make_uint2(__float_as_uint(sc), row);
                if (final_lists) {
                    Cand c;
                    c.score = sc;
                    c.row = row;
                    c.label = labels[row];
                    final_lists[(size_t)q * kp + pos] = c;
                }
            }
        }
    }
    __syncthreads();
    const int keep = min(n, kp);
    if (final_lists)
        for (int i = keep + tid; i < kp; i += kSelThreads) final_lists[(size_t)q * kp + i] = empty_cand();
    if (tid == 0) {
        cnt_g[q] = (uint32_t)keep;
        float thr = n >= kp ? ordered_to_float((uint32_t)(K >> 32)) : __int_as_float(0xff800000);
        // distance_limit pushed down (udp_packets.rs:29-39): a row whose fp16-query score is below
        // limit_score - eps_q has an exact distance above the limit and would be dropped by the caller anyway
        if (limit_score > __int_as_float(0xff800000)) thr = fmaxf(thr, limit_score - eps_q[q] - 1e-6f);
        thr_g[q] = thr;
        if (cnt > (uint32_t)log_cap) overflow_g[q] = 1u;
    }
}

// ---- host side ----------------------------------------------------------------------------
// [rows][384] fp16 row-major viewed as a 2D tensor, box = 64 columns x box_rows, 128 B swizzle.
bool make_tmap(CUtensorMap *map, const void *base, uint64_t rows, uint32_t box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {(cuuint64_t)kDim, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)kRowBytesF16};
    cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

}  // namespace

size_t gemm_workspace_bytes(int n_queries) {
    const size_t qp = ((size_t)n_queries + 2 * BM - 1) / (2 * BM) * (2 * BM);  // pair mode pads to 256
    return qp * kDim * sizeof(__half) + qp * (4 * sizeof(float)) + qp * (size_t)kSelCap * sizeof(uint2) + 1024 +
           kArriveSlots * sizeof(uint32_t);
}

template <int CG>
static cudaError_t launch_gemm_round(int grid, cudaStream_t s, const CUtensorMap &tmap_q, const CUtensorMap &tmap_x,
                                     uint32_t tile_begin, uint32_t tile_end, uint32_t n_rows, uint32_t n_tiles_total,
                                     uint32_t perm_mult, int n_qtiles, int n_queries, int chunk, const float *thr,
                                     uint32_t *cnt, uint2 *log, uint32_t *overflow, uint32_t *arrive) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = kSmemBytes;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int log_cap = kSelCap;
    return cudaLaunchKernelEx(&cfg, gemm_topk_kernel<CG>, tmap_q, tmap_x, tile_begin, tile_end, n_rows, n_tiles_total,
                              perm_mult, n_qtiles, n_queries, chunk, thr, cnt, log, overflow, log_cap, arrive);
}

cudaError_t launch_gemm_search(const GemmSearch &p, cudaStream_t s) {
    if (p.n_queries <= 0 || p.n_rows == 0) return cudaErrorInvalidValue;
    // CTA pairs (cta_group::2, 256-query tiles) once there is more than one 128-query tile
    int cg = p.cta_group;
    if (cg != 1 && cg != 2) cg = p.n_queries > BM ? 2 : 1;
    if (p.grid % 2) cg = 1;
    const int qtile = BM * cg;
    const int qp = (p.n_queries + qtile - 1) / qtile * qtile;
    const int n_qtiles = qp / qtile;
    const int workers = p.grid / cg;
    // carve the workspace
    uint8_t *w = static_cast<uint8_t *>(p.workspace);
    __half *q16 = reinterpret_cast<__half *>(w);
    w += (size_t)qp * kDim * sizeof(__half);
    float *eps_q = reinterpret_cast<float *>(w);
    w += (size_t)qp * sizeof(float);
    float *thr = reinterpret_cast<float *>(w);
    w += (size_t)qp * sizeof(float);
    uint32_t *cnt = reinterpret_cast<uint32_t *>(w);
    w += (size_t)qp * sizeof(uint32_t);
    uint32_t *overflow = reinterpret_cast<uint32_t *>(w);
    w += (size_t)qp * sizeof(uint32_t);
    w = reinterpret_cast<uint8_t *>(((uintptr_t)w + 255) & ~(uintptr_t)255);
    uint2 *log = reinterpret_cast<uint2 *>(w);
    // chunk rendezvous counters (gemm_pipe.cuh): only when several query tiles share the corpus; "gemm_unit_sync" = 0 turns it off
    uint32_t *arrive = (n_qtiles > 1 && !p.no_unit_sync) ? reinterpret_cast<uint32_t *>(w + (size_t)qp * kSelCap * sizeof(uint2)) : nullptr;

    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(gemm_topk_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(gemm_topk_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e != cudaSuccess) return e;
        configured[dev] = true;
    }
    CUtensorMap tmap_q, tmap_x;
    if (!make_tmap(&tmap_q, q16, (uint64_t)qp, BM) || !make_tmap(&tmap_x, p.corpus, p.n_rows, BN / cg))
        return cudaErrorInvalidValue;

    cudaError_t e;
    // thresholds start at -inf (0xff800000), counters and overflow flags at 0
    if ((e = cudaMemsetAsync(cnt, 0, (size_t)qp * 2 * sizeof(uint32_t), s)) != cudaSuccess) return e;
    prep_queries_kernel<<<qp, 128, 0, s>>>(p.queries, p.n_queries, qp, q16, eps_q, p.accum_slack);
    {
        // -inf thresholds: written by a select pass over empty logs (cnt == 0 -> thr = -inf)
        select_topk_kernel<<<qp, kSelThreads, 0, s>>>(log, cnt, thr, overflow, kSelCap, p.kprime, p.labels, nullptr, eps_q,
                                                      p.limit_score, arrive);
    }
    int launches = 2;
    // rounds of rows: [0,1024), then x`growth` each time ((growth-1)*k' survivors per query per
    // round, which must fit the 2048-entry log with margin); every boundary is a multiple of the
    // tile height.  Compute-bound batches use x8 (fewest survivors to handle); HBM-bound batches
    // (one query tile: the epilogue has slack) use up to x32 to save rounds.
    // With k' = 128 a x8 round leaves ~900 survivors per query and the epilogue's slow path becomes the
    // bottleneck of the early rounds; x4 measured ~5 % faster at batch 1024, k = 100 (tools/ab_gemm.py).
    // k' = 16: x16 (every round that is not launched saves its fixed ~30-40 us: launch, prologue, pipeline fill, select --
    // profiles/r02_launches_12m5_*.csv).  The pattern: a round should leave no more than ~250-400 survivors per query
    // ((growth - 1) * k'); k' = 32 (the reference's k = 20) at x16 leaves 480 and measured 1-6 % slower than x8 in three
    // same-box A/Bs on a 12.5M-row shard, no difference at 100M rows (profiles/r02_ab_growth.txt).
    uint64_t growth = p.kprime > 64 ? 4 : (p.kprime > 16 ? 8 : 16);
    if (n_qtiles == 1 && cg == 1) {
        growth = 32;
        while (growth > 8 && (growth - 1) * (uint64_t)p.kprime * 5 / 4 + (uint64_t)p.kprime > (uint64_t)kSelCap) growth /= 2;
    }
    if (p.growth >= 2) {  // tuning override, clamped so that a round's survivors fit the log
        growth = (uint64_t)p.growth;
        while (growth > 2 && (growth - 1) * (uint64_t)p.kprime * 5 / 4 + (uint64_t)p.kprime > (uint64_t)kSelCap) growth /= 2;
    }
    const uint64_t total_tiles = (p.n_rows + BN - 1) / BN;
    // stride of the visiting permutation: about 0.618 * total, made coprime to total
    uint64_t mult = (uint64_t)((double)total_tiles * 0.6180339887498949) | 1ull;
    auto gcd = [](uint64_t a, uint64_t b) { while (b) { uint64_t t = a % b; a = b; b = t; } return a; };
    while (mult > 1 && gcd(mult, total_tiles) != 1) mult += 2;
    if (total_tiles <= 2 || p.sequential_tiles) mult = 1;  // mult = 1: physical order (tuning / A-B only)
    mult %= total_tiles > 0 ? total_tiles : 1;
    if (mult == 0) mult = 1;
    uint64_t begin = 0, end = 1024 / BN;  // in (permuted) tiles: round 0 = 4 tiles = 1024 rows
    if (p.debug_raw_scores) {
        if (p.n_rows > (uint64_t)kSelCap) return cudaErrorInvalidValue;
        end = total_tiles;  // one round, thresholds still at -inf: every valid (query,row) score lands in the log
    }
    while (begin < total_tiles) {
        if (end > total_tiles || end + end / 4 > total_tiles) end = total_tiles;  // fold a short last round into this one
        const uint64_t n_tiles = end - begin;
        // Tiles per work unit: ~4 units per worker for small rounds, at most 64 tiles (A/B on a B200:
        // 16..100 tiles are equivalent within noise, >= 200 costs ~3 %: tools/ab_gemm.py).
        uint64_t chunk = (n_tiles * (uint64_t)n_qtiles + (uint64_t)workers * 4 - 1) / ((uint64_t)workers * 4);
        if (chunk < 1) chunk = 1;
        if (chunk > 64) chunk = 64;
        if (p.chunk_tiles > 0) chunk = (uint64_t)p.chunk_tiles;  // tuning override
        if (cg == 2)
            e = launch_gemm_round<2>(p.grid, s, tmap_q, tmap_x, (uint32_t)begin, (uint32_t)end, (uint32_t)p.n_rows,
                                     (uint32_t)total_tiles, (uint32_t)mult, n_qtiles, p.n_queries, (int)chunk, thr, cnt, log, overflow, arrive);
        else
            e = launch_gemm_round<1>(p.grid, s, tmap_q, tmap_x, (uint32_t)begin, (uint32_t)end, (uint32_t)p.n_rows,
                                     (uint32_t)total_tiles, (uint32_t)mult, n_qtiles, p.n_queries, (int)chunk, thr, cnt, log, overflow, arrive);
        if (e != cudaSuccess) return e;
        if (p.debug_raw_scores) {
            if (p.debug_log_out) *p.debug_log_out = log;
            if (p.debug_cnt_out) *p.debug_cnt_out = cnt;
            if (p.debug_q16_out) *p.debug_q16_out = q16;
            launches += 1;
            break;
        }
        const bool last = end >= total_tiles;
        select_topk_kernel<<<qp, kSelThreads, 0, s>>>(log, cnt, thr, overflow, kSelCap, p.kprime, p.labels,
                                                      last ? p.final_lists : nullptr, eps_q, p.limit_score, arrive);
        launches += 2;
        begin = end;
        end = end * growth;
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (p.eps_out) *p.eps_out = eps_q;
    if (p.overflow_out) *p.overflow_out = overflow;
    if (p.launches_out) *p.launches_out = launches;
    return cudaSuccess;
}

}  // namespace dawn
