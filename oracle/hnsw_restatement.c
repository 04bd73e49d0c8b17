/*
 * hnsw_restatement.c -- "HNSW restatement (NOT USearch 0.22.3)".
 *
 * TEST / BASELINE INFRASTRUCTURE ONLY (see dawn_oracle.h).
 *
 * The reference answers searches with the third-party crate usearch = "0.22.3"
 * (Cargo.toml:36, Cargo.lock:3755-3762): a C++ HNSW index configured at
 * src/search/search_provider.rs:35-42 with dimensions 384, MetricKind::IP, ScalarKind::F32 and
 * connectivity / expansion_add / expansion_search = 0, i.e. the library defaults.  That crate is
 * not vendored under /root/reference and cannot be built here (no cargo, no network), so this file
 * restates the published HNSW algorithm (Malkov & Yashunin, 2016) with USearch's documented
 * defaults -- M = 16 links per node on upper levels, 2M = 32 on level 0, efConstruction = 128,
 * efSearch = 64, level multiplier 1/ln(M), inner-product distance 1 - dot on f32 vectors,
 * single-threaded like the reference's search thread (src/bin/dawnsearch.rs:76-78).
 * It exists to report recall@k of "the reference's kind of index" against the exact answer and to
 * time a CPU approximate search next to the GPU numbers.  It is NOT bit-compatible with USearch
 * (different neighbour heuristic details, level RNG, tie handling) and is labelled as such
 * wherever its numbers appear.
 */
#include <immintrin.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define EM 384

typedef struct {
    float d;
    uint32_t id;
} hn_t;

typedef struct {  /* binary heap; max-heap if sign=+1 on d, min-heap via negated compare */
    hn_t *v;
    size_t n, cap;
} heap_t;

static void heap_reserve(heap_t *h, size_t cap) {
    if (cap > h->cap) {
        h->v = (hn_t *)realloc(h->v, cap * sizeof(hn_t));
        h->cap = cap;
    }
}
static void heap_push(heap_t *h, hn_t x, int maxheap) {
    if (h->n == h->cap) heap_reserve(h, h->cap ? h->cap * 2 : 64);
    size_t i = h->n++;
    while (i > 0) {
        size_t p = (i - 1) / 2;
        int up = maxheap ? (x.d > h->v[p].d) : (x.d < h->v[p].d);
        if (!up) break;
        h->v[i] = h->v[p];
        i = p;
    }
    h->v[i] = x;
}
static hn_t heap_pop(heap_t *h, int maxheap) {
    hn_t top = h->v[0];
    hn_t x = h->v[--h->n];
    size_t i = 0;
    for (;;) {
        size_t l = 2 * i + 1, r = l + 1, c;
        if (l >= h->n) break;
        if (r < h->n) c = maxheap ? (h->v[r].d > h->v[l].d ? r : l) : (h->v[r].d < h->v[l].d ? r : l);
        else c = l;
        int down = maxheap ? (h->v[c].d > x.d) : (h->v[c].d < x.d);
        if (!down) break;
        h->v[i] = h->v[c];
        i = c;
    }
    if (h->n) h->v[i] = x;
    return top;
}

typedef struct dawn_hnsw {
    size_t M, M0, efc, efs;
    double mult;
    uint64_t rng;
    size_t n, cap;
    float *vec;        /* [cap][384] */
    uint64_t *label;   /* [cap] */
    int *level;        /* [cap] */
    uint32_t **links;  /* [cap] -> per node: for each level l: count + slots (M0 at l=0, M above) */
    int max_level;
    uint32_t entry;
    uint32_t *visited;
    uint32_t epoch;
    heap_t cand, top, tmp;
} dawn_hnsw;

/* Scratch of one searching thread (the build uses the copy embedded in dawn_hnsw). */
typedef struct {
    uint32_t *visited;
    uint32_t epoch;
    heap_t cand, top, tmp;
} hn_ctx;

static float (*dot_fn)(const float *, const float *);

static float dot_scalar(const float *a, const float *b) {
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < EM; i += 8)
        for (int j = 0; j < 8; j++) s[j] += a[i + j] * b[i + j];
    return ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
}
__attribute__((target("avx2,fma"))) static float dot_avx2(const float *a, const float *b) {
    __m256 s0 = _mm256_setzero_ps(), s1 = s0, s2 = s0, s3 = s0;
    for (int i = 0; i < EM; i += 32) {
        s0 = _mm256_fmadd_ps(_mm256_loadu_ps(a + i), _mm256_loadu_ps(b + i), s0);
        s1 = _mm256_fmadd_ps(_mm256_loadu_ps(a + i + 8), _mm256_loadu_ps(b + i + 8), s1);
        s2 = _mm256_fmadd_ps(_mm256_loadu_ps(a + i + 16), _mm256_loadu_ps(b + i + 16), s2);
        s3 = _mm256_fmadd_ps(_mm256_loadu_ps(a + i + 24), _mm256_loadu_ps(b + i + 24), s3);
    }
    __m256 s = _mm256_add_ps(_mm256_add_ps(s0, s1), _mm256_add_ps(s2, s3));
    __m128 lo = _mm_add_ps(_mm256_castps256_ps128(s), _mm256_extractf128_ps(s, 1));
    lo = _mm_add_ps(lo, _mm_movehl_ps(lo, lo));
    lo = _mm_add_ss(lo, _mm_shuffle_ps(lo, lo, 1));
    return _mm_cvtss_f32(lo);
}
static void pick_dot(void) {
    if (dot_fn) return;
    __builtin_cpu_init();
    dot_fn = (__builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma")) ? dot_avx2 : dot_scalar;
}

static inline float dist(const dawn_hnsw *h, const float *q, uint32_t id) {
    return 1.0f - dot_fn(q, h->vec + (size_t)id * EM); /* MetricKind::IP, search_provider.rs:37 */
}
static inline size_t cap_at(const dawn_hnsw *h, int l) { return l == 0 ? h->M0 : h->M; }
static inline uint32_t *links_at(const dawn_hnsw *h, uint32_t id, int l) {
    uint32_t *p = h->links[id];
    if (l == 0) return p;
    return p + (1 + h->M0) + (size_t)(l - 1) * (1 + h->M);
}

dawn_hnsw *dawn_hnsw_new(size_t M, size_t efc, size_t efs, uint64_t seed) {
    pick_dot();
    dawn_hnsw *h = (dawn_hnsw *)calloc(1, sizeof *h);
    h->M = M ? M : 16;
    h->M0 = 2 * h->M;
    h->efc = efc ? efc : 128;
    h->efs = efs ? efs : 64;
    h->mult = 1.0 / log((double)h->M);
    h->rng = seed ? seed : 0x9E3779B97F4A7C15ull;
    h->max_level = -1;
    return h;
}

void dawn_hnsw_free(dawn_hnsw *h) {
    if (!h) return;
    for (size_t i = 0; i < h->n; i++) free(h->links[i]);
    free(h->links);
    free(h->vec);
    free(h->label);
    free(h->level);
    free(h->visited);
    free(h->cand.v);
    free(h->top.v);
    free(h->tmp.v);
    free(h);
}

size_t dawn_hnsw_size(const dawn_hnsw *h) { return h->n; }
void dawn_hnsw_set_ef_search(dawn_hnsw *h, size_t efs) { h->efs = efs; }

static double rnd01(dawn_hnsw *h) { /* splitmix64 */
    uint64_t z = (h->rng += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return ((double)(z >> 11) + 0.5) / 9007199254740992.0;
}

/* ef-bounded best-first search on one level; result left in h->top (max-heap on distance). */
static void search_layer(dawn_hnsw *h, const float *q, uint32_t ep, float ep_d, size_t ef, int l) {
    h->epoch++;
    h->cand.n = h->top.n = 0;
    hn_t e = {ep_d, ep};
    heap_push(&h->cand, e, 0);
    heap_push(&h->top, e, 1);
    h->visited[ep] = h->epoch;
    while (h->cand.n) {
        hn_t c = heap_pop(&h->cand, 0);
        if (c.d > h->top.v[0].d && h->top.n >= ef) break;
        uint32_t *lk = links_at(h, c.id, l);
        for (uint32_t i = 0; i < lk[0]; i++) {
            uint32_t nb = lk[1 + i];
            if (h->visited[nb] == h->epoch) continue;
            h->visited[nb] = h->epoch;
            float d = dist(h, q, nb);
            if (h->top.n < ef || d < h->top.v[0].d) {
                hn_t x = {d, nb};
                heap_push(&h->cand, x, 0);
                heap_push(&h->top, x, 1);
                if (h->top.n > ef) heap_pop(&h->top, 1);
            }
        }
    }
}

/* the same walk with a caller-owned scratch (read-only on the graph: safe from many threads) */
static void search_layer_ctx(const dawn_hnsw *h, hn_ctx *c, const float *q, uint32_t ep, float ep_d, size_t ef, int l) {
    c->epoch++;
    c->cand.n = c->top.n = 0;
    hn_t e = {ep_d, ep};
    heap_push(&c->cand, e, 0);
    heap_push(&c->top, e, 1);
    c->visited[ep] = c->epoch;
    while (c->cand.n) {
        hn_t x0 = heap_pop(&c->cand, 0);
        if (x0.d > c->top.v[0].d && c->top.n >= ef) break;
        uint32_t *lk = links_at(h, x0.id, l);
        for (uint32_t i = 0; i < lk[0]; i++) {
            uint32_t nb = lk[1 + i];
            if (c->visited[nb] == c->epoch) continue;
            c->visited[nb] = c->epoch;
            float d = dist(h, q, nb);
            if (c->top.n < ef || d < c->top.v[0].d) {
                hn_t x = {d, nb};
                heap_push(&c->cand, x, 0);
                heap_push(&c->top, x, 1);
                if (c->top.n > ef) heap_pop(&c->top, 1);
            }
        }
    }
}

static int cmp_hn(const void *a, const void *b) {
    const hn_t *x = (const hn_t *)a, *y = (const hn_t *)b;
    if (x->d != y->d) return x->d < y->d ? -1 : 1;
    return x->id < y->id ? -1 : (x->id > y->id);
}

/* Neighbour selection heuristic (Malkov & Yashunin alg. 4, no extension, no pruned keep):
 * walk candidates by increasing distance, keep one only if it is closer to the base than to
 * every neighbour kept so far. */
static size_t select_neighbors(dawn_hnsw *h, hn_t *cands, size_t n, size_t m, uint32_t *out) {
    qsort(cands, n, sizeof(hn_t), cmp_hn);
    size_t kept = 0;
    for (size_t i = 0; i < n && kept < m; i++) {
        int good = 1;
        const float *cv = h->vec + (size_t)cands[i].id * EM;
        for (size_t j = 0; j < kept; j++) {
            float d = 1.0f - dot_fn(cv, h->vec + (size_t)out[j] * EM);
            if (d < cands[i].d) {
                good = 0;
                break;
            }
        }
        if (good) out[kept++] = cands[i].id;
    }
    return kept;
}

int dawn_hnsw_add(dawn_hnsw *h, uint64_t label, const float *v) {
    if (h->n == h->cap) {
        size_t nc = h->cap ? h->cap * 2 : 1024;
        h->vec = (float *)realloc(h->vec, nc * EM * sizeof(float));
        h->label = (uint64_t *)realloc(h->label, nc * sizeof(uint64_t));
        h->level = (int *)realloc(h->level, nc * sizeof(int));
        h->links = (uint32_t **)realloc(h->links, nc * sizeof(uint32_t *));
        h->visited = (uint32_t *)realloc(h->visited, nc * sizeof(uint32_t));
        memset(h->visited + h->cap, 0, (nc - h->cap) * sizeof(uint32_t));
        h->cap = nc;
    }
    uint32_t id = (uint32_t)h->n;
    memcpy(h->vec + (size_t)id * EM, v, EM * sizeof(float));
    h->label[id] = label;
    int lvl = (int)(-log(rnd01(h)) * h->mult);
    h->level[id] = lvl;
    size_t words = (1 + h->M0) + (size_t)lvl * (1 + h->M);
    h->links[id] = (uint32_t *)calloc(words, sizeof(uint32_t));
    h->n++;
    if (h->max_level < 0) {
        h->max_level = lvl;
        h->entry = id;
        return 0;
    }
    uint32_t ep = h->entry;
    float ep_d = dist(h, v, ep);
    for (int l = h->max_level; l > lvl; l--) { /* greedy descent */
        int changed = 1;
        while (changed) {
            changed = 0;
            uint32_t *lk = links_at(h, ep, l);
            for (uint32_t i = 0; i < lk[0]; i++) {
                float d = dist(h, v, lk[1 + i]);
                if (d < ep_d) {
                    ep_d = d;
                    ep = lk[1 + i];
                    changed = 1;
                }
            }
        }
    }
    uint32_t sel[64];
    for (int l = lvl < h->max_level ? lvl : h->max_level; l >= 0; l--) {
        search_layer(h, v, ep, ep_d, h->efc, l);
        size_t nc = h->top.n;
        heap_reserve(&h->tmp, nc + 64);
        memcpy(h->tmp.v, h->top.v, nc * sizeof(hn_t));
        size_t m = select_neighbors(h, h->tmp.v, nc, h->M, sel); /* new node links to <= M on every level */
        uint32_t *mine = links_at(h, id, l);
        mine[0] = (uint32_t)m;
        memcpy(mine + 1, sel, m * sizeof(uint32_t));
        ep = h->tmp.v[0].id; /* closest found: entry for the next level down */
        ep_d = h->tmp.v[0].d;
        for (size_t j = 0; j < m; j++) { /* back links, shrink with the same heuristic when full */
            uint32_t nb = sel[j];
            uint32_t *lk = links_at(h, nb, l);
            size_t capl = cap_at(h, l);
            if (lk[0] < capl) {
                lk[1 + lk[0]++] = id;
            } else {
                hn_t pool[65];
                const float *nv = h->vec + (size_t)nb * EM;
                for (uint32_t t = 0; t < lk[0]; t++) {
                    pool[t].id = lk[1 + t];
                    pool[t].d = 1.0f - dot_fn(nv, h->vec + (size_t)lk[1 + t] * EM);
                }
                pool[lk[0]].id = id;
                pool[lk[0]].d = 1.0f - dot_fn(nv, v);
                uint32_t keep[64];
                size_t kn = select_neighbors(h, pool, lk[0] + 1, capl, keep);
                lk[0] = (uint32_t)kn;
                memcpy(lk + 1, keep, kn * sizeof(uint32_t));
            }
        }
    }
    if (lvl > h->max_level) {
        h->max_level = lvl;
        h->entry = id;
    }
    return 0;
}

/* usearch Index::search(query, k): approximate top-k, ascending distance = 1 - dot. */
size_t dawn_hnsw_search(dawn_hnsw *h, const float *q, size_t k, uint64_t *labels_out, float *dist_out) {
    if (h->n == 0 || k == 0) return 0;
    uint32_t ep = h->entry;
    float ep_d = dist(h, q, ep);
    for (int l = h->max_level; l > 0; l--) {
        int changed = 1;
        while (changed) {
            changed = 0;
            uint32_t *lk = links_at(h, ep, l);
            for (uint32_t i = 0; i < lk[0]; i++) {
                float d = dist(h, q, lk[1 + i]);
                if (d < ep_d) {
                    ep_d = d;
                    ep = lk[1 + i];
                    changed = 1;
                }
            }
        }
    }
    size_t ef = h->efs > k ? h->efs : k;
    search_layer(h, q, ep, ep_d, ef, 0);
    size_t n = h->top.n;
    heap_reserve(&h->tmp, n + 1);
    memcpy(h->tmp.v, h->top.v, n * sizeof(hn_t));
    qsort(h->tmp.v, n, sizeof(hn_t), cmp_hn);
    if (n > k) n = k;
    for (size_t i = 0; i < n; i++) {
        labels_out[i] = h->label[h->tmp.v[i].id];
        dist_out[i] = h->tmp.v[i].d;
    }
    return n;
}

/* ---- batch entry points (no per-vector FFI overhead; several searching threads) ---------------- */

int dawn_hnsw_add_batch(dawn_hnsw *h, const uint64_t *labels, const float *vecs, size_t n) {
    for (size_t i = 0; i < n; i++)
        if (dawn_hnsw_add(h, labels[i], vecs + i * EM)) return -1;
    return 0;
}

static size_t search_ctx(const dawn_hnsw *h, hn_ctx *c, const float *q, size_t k, uint64_t *labels_out, float *dist_out) {
    if (h->n == 0 || k == 0) return 0;
    uint32_t ep = h->entry;
    float ep_d = dist(h, q, ep);
    for (int l = h->max_level; l > 0; l--) {
        int changed = 1;
        while (changed) {
            changed = 0;
            uint32_t *lk = links_at(h, ep, l);
            for (uint32_t i = 0; i < lk[0]; i++) {
                float d = dist(h, q, lk[1 + i]);
                if (d < ep_d) {
                    ep_d = d;
                    ep = lk[1 + i];
                    changed = 1;
                }
            }
        }
    }
    size_t ef = h->efs > k ? h->efs : k;
    search_layer_ctx(h, c, q, ep, ep_d, ef, 0);
    size_t n = c->top.n;
    heap_reserve(&c->tmp, n + 1);
    memcpy(c->tmp.v, c->top.v, n * sizeof(hn_t));
    qsort(c->tmp.v, n, sizeof(hn_t), cmp_hn);
    if (n > k) n = k;
    for (size_t i = 0; i < n; i++) {
        labels_out[i] = h->label[c->tmp.v[i].id];
        dist_out[i] = c->tmp.v[i].d;
    }
    return n;
}

#include <pthread.h>
typedef struct {
    const dawn_hnsw *h;
    const float *q;
    size_t nq, k, t, nt;
    uint64_t *labels;
    float *dist;
    uint64_t *counts;
} mt_job;

static void *mt_worker(void *arg) {
    mt_job *j = (mt_job *)arg;
    hn_ctx c;
    memset(&c, 0, sizeof c);
    c.visited = (uint32_t *)calloc(j->h->cap ? j->h->cap : 1, sizeof(uint32_t));
    for (size_t i = j->t; i < j->nq; i += j->nt)
        j->counts[i] = search_ctx(j->h, &c, j->q + i * EM, j->k, j->labels + i * j->k, j->dist + i * j->k);
    free(c.visited);
    free(c.cand.v);
    free(c.top.v);
    free(c.tmp.v);
    return NULL;
}

/* nq independent searches spread over `threads` threads (the graph is read-only while searching).
 * threads = 1 is the reference's own model: one search thread (src/bin/dawnsearch.rs:76-78). */
int dawn_hnsw_search_batch(const dawn_hnsw *h, const float *queries, size_t nq, size_t k, int threads, uint64_t *labels_out,
                           float *dist_out, uint64_t *counts_out) {
    if (threads < 1) threads = 1;
    if ((size_t)threads > nq) threads = (int)(nq ? nq : 1);
    pthread_t th[256];
    mt_job jobs[256];
    if (threads > 256) threads = 256;
    for (int t = 0; t < threads; t++) {
        jobs[t] = (mt_job){h, queries, nq, k, (size_t)t, (size_t)threads, labels_out, dist_out, counts_out};
        if (t > 0) pthread_create(&th[t], NULL, mt_worker, &jobs[t]);
    }
    mt_worker(&jobs[0]);
    for (int t = 1; t < threads; t++) pthread_join(th[t], NULL);
    return 0;
}
