// gemm_pipe.cuh -- inline-PTX building blocks shared by the tensor-core kernels (gemm_topk.cu: kind::f16,
// gemm_i8.cu: kind::i8): mbarriers, TMA tensor loads, tcgen05 MMA / commit / TMEM loads, CTA-pair helpers,
// shared-memory matrix descriptors.  sm_100a only.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dawn {
namespace pipe {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
            "r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// tcgen05.commit: the mbarrier gets one arrival when every MMA issued so far by this thread retires.
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- CTA-pair (cta_group::2) variants ----------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) inside CTA `rank`
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
// Arrival on a barrier of another CTA of the cluster (the epilogue warps of a CTA pair tell the leader's MMA issuer that a
// TMEM accumulator may be overwritten).  Default semantics (.release at CTA scope), the form CUTLASS's ClusterBarrier::arrive
// uses: what has to be ordered before the arrival are this warp's tcgen05.ld's, and tcgen05.wait::ld +
// tcgen05.fence::before_thread_sync do that.  The former .release.cluster cost a MEMBAR.ALL.GPU + ERRBAR per tile per warp --
// 25 % of the epilogue warps' stall samples in the first ncu capture of gemm_i8_topk_kernel (profiles/r02_*).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
#ifdef DAWN_ARRIVE_RELEASE_CLUSTER  // A/B builds only (make OUT=../lib_ab EXTRA=-DDAWN_ARRIVE_RELEASE_CLUSTER)
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
#else
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
#endif
}
// TMA load issued by either CTA of a pair; completion bytes are counted on the LEADER's barrier.
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap *map, int c0, int c1,
                                                 uint32_t leader_bar_cluster_addr) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
            "r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(leader_bar_cluster_addr)
        : "memory");
}
// commit of the pair's MMAs: one arrival on the same barrier in BOTH CTAs
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void tc_mma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// Shared-memory matrix descriptor for a K-major operand stored as rows of 128 B with the TMA
// 128-byte swizzle: 8-row groups are 1024 B apart (SBO), version 1 (Blackwell), layout 2.
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);        // start address, 16-byte units
    d |= (uint64_t)1 << 16;                            // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset: 8 rows x 128 B
    d |= (uint64_t)1 << 46;                            // descriptor version
    d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
    return d;
}

// kind::i8: A, B int8 (signed), accumulators s32 in TMEM.  Same operand descriptors as kind::f16; one MMA covers
// K = 32 elements = 32 bytes of every operand row.
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_mma_i8_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::
            "r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2,
                                                 uint32_t leader_bar_cluster_addr) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::
            "r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(leader_bar_cluster_addr)
        : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = (EncodeTiledFn)p;
    }
    return fn;
}

// ---- rendezvous of the workers that share a corpus chunk ----------------------------------------------------------------
// The T workers (CTAs / CTA pairs) that hold the T query tiles of one corpus chunk read the same rows; they share them through
// L2 only if they stream the chunk at the same time.  Left alone they drift apart over a launch -- ncu then shows the corpus
// read from DRAM up to 1.9x -- so each worker's producer registers at the chunk's counter and waits until all T have (or 30 us
// have passed: the wait is a performance hint, never a correctness condition, and two launches that share the SMs must not
// be able to block each other).  Measured (final round of a 20M-row index, batch 1024, same box): DRAM 1.90x -> 1.003x of the
// algorithmic bytes, 4.5 % less time at a 7 % higher clock (profiles/r02_chunk_rendezvous_ab.txt).
constexpr int kArriveSlots = 8192;  // counters of one round; zeroed by the select kernel that precedes it
__device__ __forceinline__ void chunk_rendezvous(uint32_t *ctr, uint32_t sharers, bool registers) {
    if (registers) atomicAdd(ctr, 1u);
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (true) {
        uint32_t seen;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory");
        if (seen >= sharers) break;
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 30000ull) break;
        __nanosleep(64);
    }
}
// every block of a select launch clears its share of the next round's counters
__device__ __forceinline__ void clear_arrive_slots(uint32_t *arrive, int block, int n_blocks, int tid, int n_threads) {
    if (arrive == nullptr) return;
    const int per = (kArriveSlots + n_blocks - 1) / n_blocks;
    for (int i = tid; i < per; i += n_threads) {
        const int j = block * per + i;
        if (j < kArriveSlots) arrive[j] = 0u;
    }
}

// ---- selection between rounds (shared by select_topk_kernel and select_i8_kernel) ----------------
constexpr int kSelThreads = 256;

// The kp-th largest of n distinct 64-bit keys in shared memory, by an MSB-first radix select (8 passes over a
// 256-bin histogram).  All kSelThreads threads call; s_prefix / s_need / hist are shared scratch (s_prefix = 0 and
// s_need = kp on entry).  Shared with gemm_i8.cu.
__device__ __forceinline__ unsigned long long radix_select_kth(const unsigned long long *keys, int n, int kp, uint32_t *hist,
                                                               unsigned long long *s_prefix, uint32_t *s_need, int tid) {
    (void)kp;
    for (int byte = 7; byte >= 0; byte--) {
        hist[tid] = 0u;  // kSelThreads == 256
        __syncthreads();
        const unsigned long long prefix = *s_prefix;
        for (int i = tid; i < n; i += kSelThreads) {
            const unsigned long long key = keys[i];
            if (byte == 7 || (key >> (8 * (byte + 1))) == prefix) atomicAdd(&hist[(uint32_t)(key >> (8 * byte)) & 0xFFu], 1u);
        }
        __syncthreads();
        if (tid < 32) {  // digits from 255 down; lane L owns digits 255-8L .. 248-8L
            uint32_t mine = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) mine += hist[255 - 8 * tid - j];
            uint32_t incl = mine;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, incl, off);
                if (tid >= off) incl += o;
            }
            const uint32_t need = *s_need;
            const uint32_t hit = __ballot_sync(0xffffffffu, incl >= need);
            __syncwarp();  // every lane has read *s_need before the owner of the digit rewrites it (the ballot already orders this; explicit for racecheck)
            if (tid == __ffs(hit) - 1) {
                uint32_t cum = incl - mine;
                for (int j = 0; j < 8; j++) {
                    const int d = 255 - 8 * tid - j;
                    if (cum + hist[d] >= need) {
                        *s_prefix = (prefix << 8) | (unsigned long long)d;
                        *s_need = need - cum;
                        break;
                    }
                    cum += hist[d];
                }
            }
        }
        __syncthreads();
    }
    return *s_prefix;
}
constexpr int kSelCap = 2048;  // candidate-log capacity per query == keys held in shared memory by a select CTA

}  // namespace pipe
}  // namespace dawn
