/*
 * dawn_oracle.h -- CPU oracle for the DawnSearch vector top-k hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.  The product library (libdawn_b200.so) never links, loads or
 * calls anything in this directory and has no CPU fallback.
 *
 * PARITY UNPINNED: the reference (dawn-search/dawnsearch v0.2.0) holds no tests,
 * fixtures or golden vectors for this path (.github/workflows/build.yml:32 "No tests
 * yet!") and its arithmetic lives in the un-vendored third-party crate
 * usearch = "0.22.3" (Cargo.toml:36, Cargo.lock:3755-3762), an *approximate* HNSW
 * index that cannot be built here (no cargo/rustc, no network).  What this oracle
 * restates is therefore (i) the reference's own scalar semantics in
 * src/search/vector.rs and src/search/best_results.rs, and (ii) north_star's
 * definition of correctness: an exact f32 brute force over the stored vectors.
 * The only pins available are the reference's runtime invariants,
 * which tests/test_oracle.py checks.
 *
 * All citations are relative to /root/reference/.
 */
#ifndef DAWN_ORACLE_H
#define DAWN_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DAWN_ORACLE_EM_LEN 384 /* src/search/vector.rs:26 EM_LEN */

/* ---- src/search/vector.rs ------------------------------------------------ */

/* vector.rs:95-97  Distance<f32,f32>::distance: sequential f32 sum of (a-b)^2. */
float dawn_oracle_distance_l2sq(const float *a, const float *b);
/* vector.rs:99-101 distance_ip: sequential f32 sum of a*b (mul and add rounded separately). */
float dawn_oracle_distance_ip(const float *a, const float *b);
/* vector.rs:128-134 distance_cosine: 1.0f - sequential f32 dot. */
float dawn_oracle_distance_cosine(const float *a, const float *b);
/* vector.rs:181-183 vector_length: sqrt of distance to the zero vector. */
float dawn_oracle_vector_length(const float *v);
/* vector.rs:185-192 is_normalized: finite and 0.99 < |v| < 1.01. */
int dawn_oracle_is_normalized(const float *v);
/* vector.rs:194-197 normalize: divide by sqrt(sequential f32 sum of squares). */
void dawn_oracle_normalize(float *v);
/* vector.rs:30-32 f32_to_i16: round-half-away(x * 32767) saturating cast. */
int16_t dawn_oracle_f32_to_i16(float x);
/* vector.rs:104-116 Distance<i16,u64>: i64 sums. */
uint64_t dawn_oracle_distance_i16(const int16_t *a, const int16_t *b);
uint64_t dawn_oracle_distance_ip_i16(const int16_t *a, const int16_t *b);
/* vector.rs:149-155 distance_reduced: u32 sum of squared i16 differences, as f32. */
float dawn_oracle_distance_reduced(const float *a, const float *b);
/* vector.rs:157-163 distance_i8: u32 sum of squared differences. */
uint32_t dawn_oracle_distance_i8(const int8_t *a, const int8_t *b);
/* vector.rs:74-86 to24: trunc(((x+1)/2) * 0x7FFFFF) little endian, 3 bytes per element. */
void dawn_oracle_to24(const float *v, uint8_t *out1152);
/* vector.rs:52-72 from24: decode (incl. the reference's `v |= 0xFF` quirk when the top
 * bit of the high byte is set); returns 0 when the decoded vector is not normalised. */
int dawn_oracle_from24(const uint8_t *in1152, float *out);

/* ---- stored-vector formats (ours; north_star: fp16 corpus, optional i8) ----- */

/* IEEE binary32 -> binary16, round to nearest even, software bit manipulation. */
uint16_t dawn_oracle_f32_to_f16(float x);
float dawn_oracle_f16_to_f32(uint16_t h);
void dawn_oracle_store_f16(const float *rows, size_t n, uint16_t *out);
/* i8 row: scale = absmax/127 (1.0 if absmax==0), q = round-half-away(x/scale)
 * (rounding mode follows the reference's precedent vector.rs:30-32). */
void dawn_oracle_store_i8(const float *rows, size_t n, int8_t *out, float *scales);

/* ---- synthetic corpus: bit-exact on CPU and GPU (integer hash + IEEE f64) ---- */

/* Row `row` of the synthetic corpus with seed `seed`: f32 unit vector. */
void dawn_oracle_synth_row_f32(uint64_t seed, uint64_t row, float *out384);
void dawn_oracle_synth_rows_f16(uint64_t seed, uint64_t first_row, size_t n, uint16_t *out);

/* ---- the hot path: exact top-k ------------------------------------------------ */

/* Exact search over fp16-stored rows.  score = sequential f32 sum of q[i]*f32(x[i])
 * (vector.rs:128-134 order), total order (distance asc, label asc, row asc) on the emitted
 * distance = 1.0f - score.  labels==NULL means label = row + 1
 * (SQLite rowids, search_provider.rs:275).  Returns count = min(k, n). */
size_t dawn_oracle_search_f16(const uint16_t *corpus, const uint64_t *labels, size_t n,
                              const float *query, size_t k, uint64_t *labels_out,
                              float *distances_out);
/* Same over f32-stored rows (config C1: the reference's ScalarKind::F32 corpus). */
size_t dawn_oracle_search_f32(const float *corpus, const uint64_t *labels, size_t n,
                              const float *query, size_t k, uint64_t *labels_out,
                              float *distances_out);
/* Same over i8-stored rows: score = scale[row] * sequential f32 sum of q[i]*f32(x[i]). */
size_t dawn_oracle_search_i8(const int8_t *corpus, const float *scales, const uint64_t *labels,
                             size_t n, const float *query, size_t k, uint64_t *labels_out,
                             float *distances_out);
/* Exact score of one stored fp16 row (used to re-check GPU results at full size). */
float dawn_oracle_score_f16(const uint16_t *row, const float *query);

/* ---- src/search/best_results.rs ------------------------------------------- */

typedef struct dawn_oracle_best_results dawn_oracle_best_results;
/* best_results.rs:35-43 */
dawn_oracle_best_results *dawn_oracle_best_new(size_t size);
void dawn_oracle_best_free(dawn_oracle_best_results *b);
/* best_results.rs:44-65 insert: dedupe by id; once full replace worst only on strict <. */
int dawn_oracle_best_insert(dawn_oracle_best_results *b, uint64_t id, float distance);
/* best_results.rs:71-79 sort ascending by distance (stable). */
void dawn_oracle_best_sort(dawn_oracle_best_results *b);
size_t dawn_oracle_best_len(const dawn_oracle_best_results *b);
/* best_results.rs:93-95 worst_distance (0 until full: the quirk in SURVEY.md section 5). */
float dawn_oracle_best_worst_distance(const dawn_oracle_best_results *b);
void dawn_oracle_best_get(const dawn_oracle_best_results *b, size_t i, uint64_t *id,
                          float *distance);

/* ---- cpu_scan.c: threaded SIMD exact scan (the timed CPU baseline) ---------- */

/* Same contract and same results as dawn_oracle_search_f16, for `nq` queries:
 * a SIMD candidate pass keeps k+slack rows per thread, candidates are then re-scored
 * with dawn_oracle_score_f16 and ordered by the oracle's total order.
 * Returns 0 on success; *certified is 0 if the slack bound could not prove exactness. */
int dawn_cpu_scan_f16(const uint16_t *corpus, const uint64_t *labels, size_t n,
                      const float *queries, size_t nq, size_t k, int threads,
                      uint64_t *labels_out, float *distances_out, size_t *counts_out,
                      int *certified);
int dawn_cpu_scan_threads_default(void);

#ifdef __cplusplus
}
#endif
#endif
