// i8_tensor.cu -- large batches over an int8-stored corpus on the tensor cores, by way of the fp16 tiles.
//
// K4 (scan_topk.cu) answers 1-2 queries per pass over the int8 arena; a batch of B queries costs ceil(B/2) passes.
// For big batches the corpus is instead taken in chunks: a chunk is dequantised into an fp16 scratch
// (x~ = fp16(s_row * x8), HBM-bound), the tcgen05 rounds of gemm_topk.cu run over the scratch and leave k'
// candidates per query, the lists of all chunks are gathered (row ids made global, each list sorted) and
// finalize.cu selects from them and re-scores the survivors exactly from the int8 arena -- so labels and distances
// are still bit-identical to oracle/dawn_oracle.c:dawn_oracle_search_i8.
// The rounding of x~ moves a score by at most 2^-11 * ||q|| * ||s x8|| <= 5.5e-4; the caller adds that to the
// accumulation slack, i.e. to every eps_q the certificate uses.
//
// Used for batches of at least "i8_tensor_min_batch" queries (dawn_index_set_option, default 16) over at least 65,536 rows.
// Measured on a B200: 62.5M rows (one shard of config C5), batch 1024: 57.4 ms (k = 10), 60.9 ms (k = 100), no
// escalations -- the scan needs 512 passes of 3.9 ms for the same batch.  Precedent for int8 storage in the reference:
// ScalarKind::F8 in /root/reference/examples_old/search_usearch.rs:38, distance_i8 in src/search/vector.rs:157-163.
#include "dawn_common.cuh"

namespace dawn {

namespace {

// One thread per 16 stored bytes: 16 int8 -> 16 fp16 (two 16-byte stores).
__global__ void __launch_bounds__(256) dequant_i8_f16_kernel(const uint8_t *__restrict__ arena, size_t first_row,
                                                             size_t n_rows, __half *__restrict__ out) {
    constexpr int kPieces = kDim / 16;  // 24
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_rows * kPieces) return;
    const size_t r = t / kPieces;
    const int piece = (int)(t % kPieces);
    const size_t row = first_row + r;
    const uint4 u = *reinterpret_cast<const uint4 *>(arena + i8_row_offset(row) + (size_t)piece * 16);
    const float scale = *reinterpret_cast<const float *>(arena + i8_scale_offset(row));
    const int8_t *b = reinterpret_cast<const int8_t *>(&u);
    uint32_t w[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const __half2 h2 = __floats2half2_rn(__fmul_rn(scale, (float)b[2 * i]), __fmul_rn(scale, (float)b[2 * i + 1]));
        w[i] = *reinterpret_cast<const uint32_t *>(&h2);
    }
    uint4 *dst = reinterpret_cast<uint4 *>(out + r * kDim + (size_t)piece * 16);
    dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
    dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
}

// One CTA per query: the chunk's candidate list -> slot `chunk` of the query's gathered lists, rows made global,
// sorted by cand_better (finalize's tree-merge fallback needs sorted lists), empty slots last.
__global__ void __launch_bounds__(kMaxCand) gather_chunk_lists_kernel(const Cand *__restrict__ chunk_lists, int kp,
                                                                      uint32_t row_offset, int chunk, int n_chunks,
                                                                      Cand *__restrict__ gathered,
                                                                      const uint32_t *__restrict__ overflow,
                                                                      uint32_t *__restrict__ overflow_any) {
    __shared__ Cand s[kMaxCand];
    const int q = blockIdx.x, tid = threadIdx.x;
    Cand c = empty_cand();
    if (tid < kp) {
        c = chunk_lists[(size_t)q * kp + tid];
        if (c.row != kNoRow) c.row += row_offset;
        else c = empty_cand();
        s[tid] = c;
    }
    __syncthreads();
    if (tid < kp) {
        int rank = 0;
        const bool valid = c.row != kNoRow;
        for (int j = 0; j < kp; j++) {
            const Cand o = s[j];
            const bool o_valid = o.row != kNoRow;
            if (valid) rank += (o_valid && cand_better(o, c)) ? 1 : 0;
            else rank += (o_valid || j < tid) ? 1 : 0;  // empty slots keep their relative order behind the valid ones
        }
        gathered[((size_t)q * n_chunks + chunk) * kp + rank] = c;
    }
    if (tid == 0 && overflow[q]) overflow_any[q] = 1u;
}

}  // namespace

cudaError_t launch_dequant_i8_f16(const uint8_t *arena, size_t first_row, size_t n_rows, __half *out, cudaStream_t s) {
    if (n_rows == 0) return cudaSuccess;
    const size_t threads = n_rows * (kDim / 16);
    dequant_i8_f16_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(arena, first_row, n_rows, out);
    return cudaGetLastError();
}

cudaError_t launch_gather_chunk_lists(const Cand *chunk_lists, int n_queries, int kp, uint32_t row_offset, int chunk,
                                      int n_chunks, Cand *gathered, const uint32_t *overflow, uint32_t *overflow_any,
                                      cudaStream_t s) {
    if (n_queries <= 0) return cudaSuccess;
    if (kp < 1 || kp > kMaxCand) return cudaErrorInvalidValue;
    gather_chunk_lists_kernel<<<n_queries, kMaxCand, 0, s>>>(chunk_lists, kp, row_offset, chunk, n_chunks, gathered, overflow,
                                                           overflow_any);
    return cudaGetLastError();
}

}  // namespace dawn
