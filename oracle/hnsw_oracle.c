/*
 * hnsw_oracle.c — CPU restatement of hnsw-clj's distance core.  TEST INFRASTRUCTURE ONLY.
 *
 * This file restates, in plain C, the arithmetic of the reference's hot path (flat exact search,
 * IVF-FLAT build + list scan, k-means assign/update, HNSW neighbour-candidate scoring) exactly as
 * the Clojure source computes it: sequential fp64 accumulation in index order, separate multiply
 * and add (compile with -ffp-contract=off, never -ffast-math), stable sorts, java.util.Random.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  Nothing under hnsw_clj_b200/ links, imports or calls it.
 *
 * PARITY PINNING: the reference (pure Clojure, no JVM in the build container) cannot be executed
 * here, and its own tests hold only three known-answer checks on this path
 * (test/hnsw/core_test.clj:9-31).  Those KATs, plus the published java.util.Random(42) values,
 * are what pin this file (tests/test_oracle.py).  Top-k id lists, IVF partitions and recall are
 * "parity unpinned" by the reference's tests: for them this restatement IS the pin.
 *
 * Every function cites the reference file:line (relative to the reference repo root) it follows.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_EXPORT __attribute__((visibility("default")))

enum { ORC_COSINE = 0, ORC_L2 = 1, ORC_IP = 2 };

/* ------------------------------------------------------------------------------------------
 * A.1 pairwise arithmetic
 * ---------------------------------------------------------------------------------------- */

/* src/hnsw/ultra_fast.clj:53-95 (cosine-distance-ultra): three sequential accumulators; the 4x
 * unrolled n-ary `+` is a left fold, i.e. strict index order.  Guard at :92-95 returns 1.0 unless
 * both squared norms are > 0. */
ORC_EXPORT double orc_cosine_distance_ultra(const double *a, const double *b, int64_t d) {
    double dot = 0.0, n1 = 0.0, n2 = 0.0;
    for (int64_t i = 0; i < d; ++i) {
        double x = a[i], y = b[i];
        dot = dot + x * y;
        n1 = n1 + x * x;
        n2 = n2 + y * y;
    }
    if (n1 > 0.0 && n2 > 0.0) return 1.0 - dot / (sqrt(n1) * sqrt(n2));
    return 1.0;
}

/* src/hnsw/simd.clj:129-147 (cosine-distance-direct): same sums, guard is `zero? magnitude`. */
ORC_EXPORT double orc_cosine_distance_direct(const double *a, const double *b, int64_t d) {
    double dot = 0.0, n1 = 0.0, n2 = 0.0;
    for (int64_t i = 0; i < d; ++i) {
        double x = a[i], y = b[i];
        dot = dot + x * y;
        n1 = n1 + x * x;
        n2 = n2 + y * y;
    }
    double mag = sqrt(n1) * sqrt(n2);
    if (mag == 0.0) return 1.0;
    return 1.0 - dot / mag;
}

/* src/hnsw/ultra_fast.clj:43-51 (euclidean-distance-ultra) == src/hnsw/simd.clj:149-160. */
ORC_EXPORT double orc_euclidean_distance(const double *a, const double *b, int64_t d) {
    double s = 0.0;
    for (int64_t i = 0; i < d; ++i) {
        double t = a[i] - b[i];
        s = s + t * t;
    }
    return sqrt(s);
}

/* src/hnsw/simd_optimized.clj:283-293 (dot-product, fallback branch). */
ORC_EXPORT double orc_dot(const double *a, const double *b, int64_t d) {
    double s = 0.0;
    for (int64_t i = 0; i < d; ++i) s = s + a[i] * b[i];
    return s;
}

/* src/hnsw/ann/partition/ivf_flat.clj:173-176 and src/hnsw/simd_optimized.clj:206-216. */
ORC_EXPORT double orc_norm(const double *a, int64_t d) {
    double s = 0.0;
    for (int64_t i = 0; i < d; ++i) s = s + a[i] * a[i];
    return sqrt(s);
}

/* fp32-stored rows: the reference holds fp32-representable values in double[];
 * widening float->double is exact, so these are the same arithmetic on the same values. */
static inline double dot_fq(const float *v, const double *q, int64_t d) {
    double s = 0.0;
    for (int64_t i = 0; i < d; ++i) s = s + (double)v[i] * q[i];
    return s;
}
static inline double sumsq_f(const float *v, int64_t d) {
    double s = 0.0;
    for (int64_t i = 0; i < d; ++i) s = s + (double)v[i] * (double)v[i];
    return s;
}
static inline double sumsq_d(const double *v, int64_t d) {
    double s = 0.0;
    for (int64_t i = 0; i < d; ++i) s = s + v[i] * v[i];
    return s;
}
static inline double l2_fq(const float *v, const double *q, int64_t d) {
    /* argument order (query, vector) as at every reference call site; (a-b)^2 == (b-a)^2 bitwise */
    double s = 0.0;
    for (int64_t i = 0; i < d; ++i) {
        double t = q[i] - (double)v[i];
        s = s + t * t;
    }
    return sqrt(s);
}
/* cosine-distance-ultra(x, c) with x an fp32 row and c an fp64 vector (centroid or widened query) */
static inline double cos_ultra_fq(const float *x, const double *c, int64_t d) {
    double dot = 0.0, n1 = 0.0, n2 = 0.0;
    for (int64_t i = 0; i < d; ++i) {
        double a = (double)x[i], b = c[i];
        dot = dot + a * b;
        n1 = n1 + a * a;
        n2 = n2 + b * b;
    }
    if (n1 > 0.0 && n2 > 0.0) return 1.0 - dot / (sqrt(n1) * sqrt(n2));
    return 1.0;
}
static inline double dist_fn_fq(int metric, const float *x, const double *c, int64_t d) {
    return metric == ORC_L2 ? l2_fq(x, c, d) : cos_ultra_fq(x, c, d);
}

ORC_EXPORT void orc_row_norms_f32(const float *rows, int64_t n, int64_t d, double *out) {
    for (int64_t i = 0; i < n; ++i) out[i] = sqrt(sumsq_f(rows + i * d, d));
}

/* ------------------------------------------------------------------------------------------
 * A.8 java.util.Random
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    uint64_t seed;
    int have_next_gaussian;
    double next_gaussian;
} orc_rng;

#define JR_MULT 0x5DEECE66DULL
#define JR_MASK ((1ULL << 48) - 1)

ORC_EXPORT void orc_rng_init(orc_rng *r, int64_t seed) {
    r->seed = ((uint64_t)seed ^ JR_MULT) & JR_MASK;
    r->have_next_gaussian = 0;
    r->next_gaussian = 0.0;
}
static inline int32_t jr_next(orc_rng *r, int bits) {
    r->seed = (r->seed * JR_MULT + 0xBULL) & JR_MASK;
    return (int32_t)(uint32_t)(r->seed >> (48 - bits)); /* (int)(seed >>> (48 - bits)) */
}
ORC_EXPORT int32_t orc_rng_next_int(orc_rng *r) { return jr_next(r, 32); }
ORC_EXPORT int32_t orc_rng_next_int_bound(orc_rng *r, int32_t bound) {
    int32_t x = jr_next(r, 31);
    int32_t m = bound - 1;
    if ((bound & m) == 0) return (int32_t)(((int64_t)bound * (int64_t)x) >> 31);
    for (int32_t u = x;; u = jr_next(r, 31)) {
        x = u % bound;
        /* Java: u - r + m < 0 in wrapping int32 arithmetic */
        int32_t t = (int32_t)((uint32_t)u - (uint32_t)x + (uint32_t)m);
        if (t >= 0) break;
    }
    return x;
}
ORC_EXPORT double orc_rng_next_double(orc_rng *r) {
    int64_t hi = (int64_t)jr_next(r, 26);
    int64_t lo = (int64_t)jr_next(r, 27);
    return (double)((hi << 27) + lo) * 0x1.0p-53;
}
/* StrictMath.log = fdlibm's __ieee754_log (e_log.c, "FreeBSD msun / Sun fdlibm 5.3"; the JDK ships a port of it,
 * java.lang.FdLibm since JDK 21, the C original before): argument reduction x = 2^k (1+f), s = f/(2+f),
 * log(1+f) = f - (hfsq - s (hfsq + R(s^2))) with the degree-14 Remez polynomial Lg1..Lg7.  Restated here because
 * java.util.Random.nextGaussian (the LSH projection matrices, src/hnsw/ann/hash/hybrid_lsh.clj:24-31) goes through
 * it and libm's log is only faithfully rounded, not bit-identical.  Pinned by new Random(42).nextGaussian() ==
 * 1.1419053154730547 (tests/test_oracle.py). */
ORC_EXPORT double orc_strict_log(double x) {
    static const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
                        two54 = 1.80143985094819840000e+16, Lg1 = 6.666666666666735130e-01,
                        Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01, Lg4 = 2.222219843214978396e-01,
                        Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01, Lg7 = 1.479819860511658591e-01;
    union { double f; uint64_t u; } c;
    c.f = x;
    int32_t hx = (int32_t)(c.u >> 32);
    uint32_t lx = (uint32_t)c.u;
    int32_t k = 0, i, j;
    if (hx < 0x00100000) { /* x < 2^-1022 */
        if (((hx & 0x7fffffff) | lx) == 0) return -INFINITY;
        if (hx < 0) return NAN;
        k -= 54;
        x *= two54;
        c.f = x;
        hx = (int32_t)(c.u >> 32);
    }
    if (hx >= 0x7ff00000) return x + x;
    k += (hx >> 20) - 1023;
    hx &= 0x000fffff;
    i = (hx + 0x95f64) & 0x100000;
    c.f = x;
    c.u = (c.u & 0xffffffffull) | ((uint64_t)(uint32_t)(hx | (i ^ 0x3ff00000)) << 32); /* normalize x or x/2 */
    x = c.f;
    k += (i >> 20);
    double f = x - 1.0, dk, R, s, z, w, t1, t2, hfsq;
    if ((0x000fffff & (2 + hx)) < 3) { /* |f| < 2^-20 */
        if (f == 0.0) {
            if (k == 0) return 0.0;
            dk = (double)k;
            return dk * ln2_hi + dk * ln2_lo;
        }
        R = f * f * (0.5 - 0.33333333333333333 * f);
        if (k == 0) return f - R;
        dk = (double)k;
        return dk * ln2_hi - ((R - dk * ln2_lo) - f);
    }
    s = f / (2.0 + f);
    dk = (double)k;
    z = s * s;
    i = hx - 0x6147a;
    w = z * z;
    j = 0x6b851 - hx;
    t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
    t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
    i |= j;
    R = t2 + t1;
    if (i > 0) {
        hfsq = 0.5 * f * f;
        if (k == 0) return f - (hfsq - s * (hfsq + R));
        return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
    }
    if (k == 0) return f - s * (f - R);
    return dk * ln2_hi - ((s * (f - R) - dk * ln2_lo) - f);
}

/* Marsaglia polar method as in java.util.Random.nextGaussian: StrictMath.log (above) and StrictMath.sqrt
 * (correctly rounded, = C sqrt). */
ORC_EXPORT double orc_rng_next_gaussian(orc_rng *r) {
    if (r->have_next_gaussian) {
        r->have_next_gaussian = 0;
        return r->next_gaussian;
    }
    double v1, v2, s;
    do {
        v1 = 2.0 * orc_rng_next_double(r) - 1.0;
        v2 = 2.0 * orc_rng_next_double(r) - 1.0;
        s = v1 * v1 + v2 * v2;
    } while (s >= 1.0 || s == 0.0);
    double mul = sqrt(-2.0 * orc_strict_log(s) / s);
    r->next_gaussian = v2 * mul;
    r->have_next_gaussian = 1;
    return v1 * mul;
}
ORC_EXPORT int64_t orc_rng_sizeof(void) { return (int64_t)sizeof(orc_rng); }

/* test/data_generator.clj:28-87 (generate-dataset).  distribution: 0 gaussian, 1 uniform,
 * 2 unit, 3 clustered.  Output is fp64, row-major. */
ORC_EXPORT void orc_gen_dataset(int64_t n, int64_t d, int distribution, int32_t num_clusters,
                                double noise, int64_t seed, double *out) {
    orc_rng r;
    orc_rng_init(&r, seed);
    double *centers = NULL;
    if (distribution == 3) {
        centers = (double *)malloc(sizeof(double) * (size_t)num_clusters * (size_t)d);
        for (int64_t i = 0; i < (int64_t)num_clusters * d; ++i) centers[i] = orc_rng_next_gaussian(&r);
    }
    for (int64_t i = 0; i < n; ++i) {
        double *v = out + i * d;
        if (distribution == 0) {
            for (int64_t j = 0; j < d; ++j) v[j] = orc_rng_next_gaussian(&r);
        } else if (distribution == 1) {
            for (int64_t j = 0; j < d; ++j) v[j] = 2.0 * orc_rng_next_double(&r) - 1.0;
        } else if (distribution == 2) {
            for (int64_t j = 0; j < d; ++j) v[j] = orc_rng_next_gaussian(&r);
            double s = 0.0; /* (reduce + (map #(* % %) v)) */
            for (int64_t j = 0; j < d; ++j) s = s + v[j] * v[j];
            double nrm = sqrt(s);
            if (nrm != 0.0)
                for (int64_t j = 0; j < d; ++j) v[j] = v[j] / nrm;
        } else {
            /* (nth centers (.nextInt rng num-clusters)) is evaluated before the per-dim noise */
            const double *c = centers + (int64_t)orc_rng_next_int_bound(&r, num_clusters) * d;
            for (int64_t j = 0; j < d; ++j) v[j] = c[j] + noise * orc_rng_next_gaussian(&r);
        }
    }
    free(centers);
}

/* ------------------------------------------------------------------------------------------
 * stable merge sort of (distance, payload) records: the reference sorts with Collections/sort
 * (TimSort) and sort-by (Arrays.sort on objects): both stable, so any stable sort reproduces
 * the order.  Comparison is Double/compare-like on distance; NaN sorts last.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    double dist;
    int64_t id;
} orc_hit;

static inline int hit_less(const orc_hit *a, const orc_hit *b) {
    /* strict "a before b"; NaN greater than everything */
    if (isnan(a->dist)) return 0;
    if (isnan(b->dist)) return 1;
    return a->dist < b->dist;
}
static void stable_sort_hits(orc_hit *a, int64_t n, orc_hit *tmp) {
    if (n < 2) return;
    if (n <= 16) { /* insertion sort (stable) */
        for (int64_t i = 1; i < n; ++i) {
            orc_hit x = a[i];
            int64_t j = i;
            while (j > 0 && hit_less(&x, &a[j - 1])) {
                a[j] = a[j - 1];
                --j;
            }
            a[j] = x;
        }
        return;
    }
    int64_t h = n / 2;
    stable_sort_hits(a, h, tmp);
    stable_sort_hits(a + h, n - h, tmp);
    int64_t i = 0, j = h, k = 0;
    while (i < h && j < n) {
        if (hit_less(&a[j], &a[i])) tmp[k++] = a[j++];
        else tmp[k++] = a[i++];
    }
    while (i < h) tmp[k++] = a[i++];
    while (j < n) tmp[k++] = a[j++];
    memcpy(a, tmp, sizeof(orc_hit) * (size_t)n);
}

/* simple static thread fan-out, mirroring src/hnsw/helper/parallel_search.clj:15-49 (one task per
 * query on a fixed pool, results in query order) */
typedef void (*orc_task_fn)(void *ctx, int64_t begin, int64_t end);
typedef struct {
    orc_task_fn fn;
    void *ctx;
    int64_t n;
    int64_t next;
    int64_t chunk;
    pthread_mutex_t mu;
} orc_pool;
static void *pool_worker(void *p) {
    orc_pool *pool = (orc_pool *)p;
    for (;;) {
        pthread_mutex_lock(&pool->mu);
        int64_t b = pool->next;
        pool->next += pool->chunk;
        pthread_mutex_unlock(&pool->mu);
        if (b >= pool->n) break;
        int64_t e = b + pool->chunk < pool->n ? b + pool->chunk : pool->n;
        pool->fn(pool->ctx, b, e);
    }
    return NULL;
}
static void parallel_for(int64_t n, int nthreads, int64_t chunk, orc_task_fn fn, void *ctx) {
    if (nthreads <= 1 || n <= 1) {
        fn(ctx, 0, n);
        return;
    }
    orc_pool pool = {fn, ctx, n, 0, chunk < 1 ? 1 : chunk, PTHREAD_MUTEX_INITIALIZER};
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], NULL, pool_worker, &pool);
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    free(th);
}

/* ------------------------------------------------------------------------------------------
 * A.3 exact flat search
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    const float *rows;
    int64_t n, d;
    const float *queries;
    int64_t k;
    int metric;
    int64_t *out_ids;
    double *out_dist;
} flat_ctx;

/* src/hnsw/bench.clj:72-84 (compute-exact-knn): per row dot, sqrt(sum v^2), sqrt(sum q^2);
 * dist = 1 - dot/(nv*nq) (no zero guard); stable sort-by :distance; take k.  Also the shape of
 * src/hnsw/simd_optimized.clj:271-280 (top-k-distances: all distances, full sort, take k).
 * metric L2 / IP (extension): distance = euclid resp. -dot, same sort. */
static void flat_task(void *p, int64_t qb, int64_t qe) {
    flat_ctx *c = (flat_ctx *)p;
    int64_t n = c->n, d = c->d, k = c->k;
    orc_hit *hits = (orc_hit *)malloc(sizeof(orc_hit) * (size_t)(n > 0 ? n : 1));
    orc_hit *tmp = (orc_hit *)malloc(sizeof(orc_hit) * (size_t)(n > 0 ? n : 1));
    double *q = (double *)malloc(sizeof(double) * (size_t)d);
    for (int64_t qi = qb; qi < qe; ++qi) {
        for (int64_t j = 0; j < d; ++j) q[j] = (double)c->queries[qi * d + j];
        double nq = sqrt(sumsq_d(q, d));
        for (int64_t i = 0; i < n; ++i) {
            const float *v = c->rows + i * d;
            double dist;
            if (c->metric == ORC_COSINE) {
                double dot = dot_fq(v, q, d);
                double nv = sqrt(sumsq_f(v, d));
                dist = 1.0 - dot / (nv * nq);
            } else if (c->metric == ORC_L2) {
                dist = l2_fq(v, q, d);
            } else {
                dist = -dot_fq(v, q, d);
            }
            hits[i].dist = dist;
            hits[i].id = i;
        }
        stable_sort_hits(hits, n, tmp);
        for (int64_t j = 0; j < k; ++j) {
            if (j < n) {
                c->out_ids[qi * k + j] = hits[j].id;
                c->out_dist[qi * k + j] = hits[j].dist;
            } else {
                c->out_ids[qi * k + j] = -1;
                c->out_dist[qi * k + j] = INFINITY;
            }
        }
    }
    free(hits);
    free(tmp);
    free(q);
}
ORC_EXPORT void orc_exact_knn_f32(const float *rows, int64_t n, int64_t d, const float *queries,
                                  int64_t nq, int64_t k, int metric, int64_t *out_ids,
                                  double *out_dist, int nthreads) {
    flat_ctx c = {rows, n, d, queries, k, metric, out_ids, out_dist};
    parallel_for(nq, nthreads, 1, flat_task, &c);
}

/* src/hnsw/bench.clj:86-92 (calc-recall): |approx ∩ exact| / |exact| per query, then the mean
 * (measure-recall :124-132).  ids < 0 are padding and ignored. */
ORC_EXPORT double orc_recall(const int64_t *approx, const int64_t *exact, int64_t nq, int64_t k) {
    double total = 0.0;
    for (int64_t q = 0; q < nq; ++q) {
        int64_t ne = 0, hit = 0;
        for (int64_t j = 0; j < k; ++j) {
            int64_t e = exact[q * k + j];
            if (e < 0) continue;
            ++ne;
            for (int64_t t = 0; t < k; ++t)
                if (approx[q * k + t] == e) {
                    ++hit;
                    break;
                }
        }
        total += ne ? (double)hit / (double)ne : 1.0;
    }
    return nq ? total / (double)nq : 1.0;
}

/* ------------------------------------------------------------------------------------------
 * A.5 k-means++  (src/hnsw/ann/partition/ivf_flat.clj:32-60)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    const float *rows;
    int64_t d;
    const double *c;
    int metric;
    double *mind;
} kpp_ctx;
static void kpp_task(void *p, int64_t b, int64_t e) {
    kpp_ctx *c = (kpp_ctx *)p;
    for (int64_t i = b; i < e; ++i) {
        double dist = dist_fn_fq(c->metric, c->rows + i * c->d, c->c, c->d);
        if (dist < c->mind[i]) c->mind[i] = dist; /* (reduce min Double/MAX_VALUE ...) */
    }
}
/* Incremental form: d_i = min(prev d_i, dist(x_i, newest centroid)).  `min` is exact, so this is
 * bit-identical to the reference's min over ALL chosen centroids (:43-49) at O(N*k*D) instead of
 * O(N*k^2*D).  orc_kmeanspp_init_literal below is the literal form, kept for cross-checking. */
static void kmeanspp_weighted(const float *rows, int64_t n, int64_t d, int32_t nlist, int metric, int64_t seed,
                              int squared, int64_t *out_seed_rows, int nthreads) {
    orc_rng r;
    orc_rng_init(&r, seed); /* (Random. 42) :37 */
    double *mind = (double *)malloc(sizeof(double) * (size_t)n);
    double *cvec = (double *)malloc(sizeof(double) * (size_t)d);
    for (int64_t i = 0; i < n; ++i) mind[i] = 1.7976931348623157e308; /* Double/MAX_VALUE */
    int64_t pick = orc_rng_next_int_bound(&r, (int32_t)n); /* :40 */
    out_seed_rows[0] = pick;
    for (int32_t t = 1; t < nlist; ++t) {
        for (int64_t j = 0; j < d; ++j) cvec[j] = (double)rows[pick * d + j];
        kpp_ctx c = {rows, d, cvec, metric, mind};
        parallel_for(n, nthreads, 4096, kpp_task, &c);
        double sum = 0.0; /* :51-52, row order */
        for (int64_t i = 0; i < n; ++i) sum = sum + (squared ? mind[i] * mind[i] : mind[i]);
        double rr = orc_rng_next_double(&r) * sum; /* :53 */
        double cum = 0.0;
        int64_t i = 0;
        for (;; ++i) { /* :54-58 */
            double w = squared ? mind[i] * mind[i] : mind[i];
            if (cum + w >= rr) break;
            cum = cum + w;
            if (i == n - 1) break; /* reference would throw ArrayIndexOutOfBounds; clamp */
        }
        pick = i;
        out_seed_rows[t] = pick;
    }
    free(mind);
    free(cvec);
}
ORC_EXPORT void orc_kmeanspp_init(const float *rows, int64_t n, int64_t d, int32_t nlist, int metric,
                                  int64_t seed, int64_t *out_seed_rows, int nthreads) {
    kmeanspp_weighted(rows, n, d, nlist, metric, seed, 1, out_seed_rows, nthreads);
}
/* Lightning's seeding (src/hnsw/ann/partition/lightning.clj:86-109): the same walk with the weights d_i instead of
 * d_i^2 (:100-106) and the minimum taken over all chosen centroids (:93-97; the incremental min is bit-identical). */
ORC_EXPORT void orc_lightning_seeds(const float *rows, int64_t n, int64_t d, int32_t nlist, int metric,
                                    int64_t seed, int64_t *out_seed_rows, int nthreads) {
    kmeanspp_weighted(rows, n, d, nlist, metric, seed, 0, out_seed_rows, nthreads);
}

ORC_EXPORT void orc_kmeanspp_init_literal(const float *rows, int64_t n, int64_t d, int32_t nlist,
                                          int metric, int64_t seed, int64_t *out_seed_rows) {
    orc_rng r;
    orc_rng_init(&r, seed);
    double *dist = (double *)malloc(sizeof(double) * (size_t)n);
    double *cents = (double *)malloc(sizeof(double) * (size_t)nlist * (size_t)d);
    int64_t pick = orc_rng_next_int_bound(&r, (int32_t)n);
    out_seed_rows[0] = pick;
    for (int64_t j = 0; j < d; ++j) cents[j] = (double)rows[pick * d + j];
    for (int32_t t = 1; t < nlist; ++t) {
        for (int64_t i = 0; i < n; ++i) {
            double m = 1.7976931348623157e308;
            for (int32_t c = 0; c < t; ++c) {
                double x = dist_fn_fq(metric, rows + i * d, cents + (int64_t)c * d, d);
                if (x < m) m = x;
            }
            dist[i] = m;
        }
        double sum = 0.0;
        for (int64_t i = 0; i < n; ++i) sum = sum + dist[i] * dist[i];
        double rr = orc_rng_next_double(&r) * sum;
        double cum = 0.0;
        int64_t i = 0;
        for (;; ++i) {
            double dsq = dist[i] * dist[i];
            if (cum + dsq >= rr) break;
            cum = cum + dsq;
            if (i == n - 1) break;
        }
        pick = i;
        out_seed_rows[t] = pick;
        for (int64_t j = 0; j < d; ++j) cents[(int64_t)t * d + j] = (double)rows[pick * d + j];
    }
    free(dist);
    free(cents);
}

/* ------------------------------------------------------------------------------------------
 * A.6 Lloyd  (src/hnsw/ann/partition/ivf_flat.clj:79-131)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    const float *rows;
    int64_t d;
    const double *cents;
    int32_t nlist;
    int metric;
    int32_t *assign;
} assign_ctx;
/* :79-90 assign-to-nearest-centroid: strict `<` from Double/MAX_VALUE, first minimum wins. */
static void assign_task(void *p, int64_t b, int64_t e) {
    assign_ctx *c = (assign_ctx *)p;
    for (int64_t i = b; i < e; ++i) {
        double best = 1.7976931348623157e308;
        int32_t bi = 0;
        for (int32_t j = 0; j < c->nlist; ++j) {
            double x = dist_fn_fq(c->metric, c->rows + i * c->d, c->cents + (int64_t)j * c->d, c->d);
            if (x < best) {
                best = x;
                bi = j;
            }
        }
        c->assign[i] = bi;
    }
}
ORC_EXPORT void orc_assign(const float *rows, int64_t n, int64_t d, const double *cents, int32_t nlist,
                           int metric, int32_t *out_assign, int nthreads) {
    assign_ctx c = {rows, d, cents, nlist, metric, out_assign};
    parallel_for(n, nthreads, 256, assign_task, &c);
}

/* :66-77 compute-centroid over members in data (row) order; :112-116 empty keeps previous. */
ORC_EXPORT void orc_update_centroids(const float *rows, int64_t n, int64_t d, const int32_t *assign,
                                     int32_t nlist, double *cents /* in: previous, out: new */) {
    double *sum = (double *)calloc((size_t)nlist * (size_t)d, sizeof(double));
    int64_t *cnt = (int64_t *)calloc((size_t)nlist, sizeof(int64_t));
    for (int64_t i = 0; i < n; ++i) {
        double *s = sum + (int64_t)assign[i] * d;
        const float *v = rows + i * d;
        for (int64_t j = 0; j < d; ++j) s[j] = s[j] + (double)v[j];
        cnt[assign[i]]++;
    }
    for (int32_t c = 0; c < nlist; ++c) {
        if (cnt[c] == 0) continue;
        double nv = (double)cnt[c];
        for (int64_t j = 0; j < d; ++j) cents[(int64_t)c * d + j] = sum[(int64_t)c * d + j] / nv;
    }
    free(sum);
    free(cnt);
}

/* :92-131 partition-vectors-kmeans.  seed_rows == NULL -> run k-means++ (Random(seed)); otherwise
 * use the given seed rows (lets callers share one init between oracle and device). */
ORC_EXPORT void orc_kmeans(const float *rows, int64_t n, int64_t d, int32_t nlist, int32_t iters,
                           int metric, int64_t seed, const int64_t *seed_rows, double *out_cents,
                           int32_t *out_assign, int nthreads) {
    int64_t *sr = (int64_t *)malloc(sizeof(int64_t) * (size_t)nlist);
    if (seed_rows) memcpy(sr, seed_rows, sizeof(int64_t) * (size_t)nlist);
    else orc_kmeanspp_init(rows, n, d, nlist, metric, seed, sr, nthreads);
    for (int32_t c = 0; c < nlist; ++c)
        for (int64_t j = 0; j < d; ++j) out_cents[(int64_t)c * d + j] = (double)rows[sr[c] * d + j];
    free(sr);
    for (int32_t it = 0; it < iters; ++it) { /* :100-117, no convergence test */
        orc_assign(rows, n, d, out_cents, nlist, metric, out_assign, nthreads);
        orc_update_centroids(rows, n, d, out_assign, nlist, out_cents);
    }
    orc_assign(rows, n, d, out_cents, nlist, metric, out_assign, nthreads); /* :119-131 */
}

/* build-lightning-index with :smart-partition? true (src/hnsw/ann/partition/lightning.clj:84-130): seeds as above,
 * every row goes to its nearest SEED (assign-to-partition :31-44: strict <, first minimum), partitions keep data order
 * (:117-120), the routing centroids are the partition means (compute-centroid :17-29), a zero vector for an empty
 * partition (:123-126).  No Lloyd rounds.  Search is search-lightning (:184-298) = the IVF-FLAT search with
 * nprobe = max(1, (int)(partitions * percent)) (:262) and centroid routing (:265-274). */
ORC_EXPORT void orc_lightning_build(const float *rows, int64_t n, int64_t d, int32_t nlist, int metric, int64_t seed,
                                    double *out_cents, int32_t *out_assign, int nthreads) {
    int64_t *sr = (int64_t *)malloc(sizeof(int64_t) * (size_t)nlist);
    orc_lightning_seeds(rows, n, d, nlist, metric, seed, sr, nthreads);
    double *seeds = (double *)malloc(sizeof(double) * (size_t)nlist * (size_t)d);
    for (int32_t c = 0; c < nlist; ++c)
        for (int64_t j = 0; j < d; ++j) seeds[(int64_t)c * d + j] = (double)rows[sr[c] * d + j];
    orc_assign(rows, n, d, seeds, nlist, metric, out_assign, nthreads);
    for (int64_t x = 0; x < (int64_t)nlist * d; ++x) out_cents[x] = 0.0;
    orc_update_centroids(rows, n, d, out_assign, nlist, out_cents); /* empty keeps what is there: the zero vector */
    free(seeds);
    free(sr);
}

/* ------------------------------------------------------------------------------------------
 * A.2 / A.4 / A.7 IVF-FLAT search (src/hnsw/ann/partition/ivf_flat.clj:217-294)
 * The index is given as centroids + per-row assignment; lists hold their rows in row order
 * (:126-129).  list_offsets/list_rows is the CSR form of that.
 * ---------------------------------------------------------------------------------------- */
ORC_EXPORT void orc_build_lists(const int32_t *assign, int64_t n, int32_t nlist, int64_t *list_offsets,
                                int64_t *list_rows) {
    for (int32_t c = 0; c <= nlist; ++c) list_offsets[c] = 0;
    for (int64_t i = 0; i < n; ++i) list_offsets[assign[i] + 1]++;
    for (int32_t c = 0; c < nlist; ++c) list_offsets[c + 1] += list_offsets[c];
    int64_t *cur = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nlist > 0 ? nlist : 1));
    if (nlist > 0) memcpy(cur, list_offsets, sizeof(int64_t) * (size_t)nlist);
    for (int64_t i = 0; i < n; ++i) list_rows[cur[assign[i]]++] = i;
    free(cur);
}

typedef struct {
    const float *rows;
    int64_t n, d;
    const double *cents;
    int32_t nlist;
    const int64_t *list_offsets;
    const int64_t *list_rows;
    const double *norms;
    const float *queries;
    int64_t k;
    int32_t nprobe;
    int coarse_metric;
    int64_t *out_ids;
    double *out_dist;
    int32_t *out_probes; /* optional [nq x nprobe], -1 padded */
} ivf_ctx;

static void ivf_task(void *p, int64_t qb, int64_t qe) {
    ivf_ctx *c = (ivf_ctx *)p;
    int64_t d = c->d, k = c->k;
    int32_t nlist = c->nlist;
    int32_t np = c->nprobe < nlist ? c->nprobe : nlist;
    double *q = (double *)malloc(sizeof(double) * (size_t)d);
    orc_hit *ch = (orc_hit *)malloc(sizeof(orc_hit) * (size_t)(nlist > 0 ? nlist : 1));
    orc_hit *ctmp = (orc_hit *)malloc(sizeof(orc_hit) * (size_t)(nlist > 0 ? nlist : 1));
    int64_t maxlist = 1;
    for (int32_t l = 0; l < nlist; ++l) {
        int64_t s = c->list_offsets[l + 1] - c->list_offsets[l];
        if (s > maxlist) maxlist = s;
    }
    orc_hit *lh = (orc_hit *)malloc(sizeof(orc_hit) * (size_t)maxlist);
    orc_hit *ltmp = (orc_hit *)malloc(sizeof(orc_hit) * (size_t)maxlist);
    int64_t mcap = (int64_t)np * 2 * k + 1;
    orc_hit *merged = (orc_hit *)malloc(sizeof(orc_hit) * (size_t)mcap);
    orc_hit *mtmp = (orc_hit *)malloc(sizeof(orc_hit) * (size_t)mcap);
    for (int64_t qi = qb; qi < qe; ++qi) {
        for (int64_t j = 0; j < d; ++j) q[j] = (double)c->queries[qi * d + j];
        /* :261-269 coarse: distance-fn(query, centroid) for all centroids, stable sort, take */
        for (int32_t l = 0; l < nlist; ++l) {
            const double *cv = c->cents + (int64_t)l * d;
            ch[l].dist = c->coarse_metric == ORC_L2 ? orc_euclidean_distance(q, cv, d)
                                                    : orc_cosine_distance_ultra(q, cv, d);
            ch[l].id = l;
        }
        stable_sort_hits(ch, nlist, ctmp);
        /* :275-278 query norm */
        double qn = sqrt(sumsq_d(q, d));
        int64_t m = 0;
        for (int32_t pr = 0; pr < np; ++pr) { /* :281-288 lists in probe order */
            int32_t l = (int32_t)ch[pr].id;
            if (c->out_probes) c->out_probes[qi * c->nprobe + pr] = l;
            int64_t b = c->list_offsets[l], e = c->list_offsets[l + 1];
            /* :217-234 search-partition: always cosine via precomputed norms, no zero guard */
            for (int64_t t = b; t < e; ++t) {
                int64_t row = c->list_rows[t];
                double dot = dot_fq(c->rows + row * d, q, d);
                lh[t - b].dist = 1.0 - dot / (qn * c->norms[row]);
                lh[t - b].id = row;
            }
            stable_sort_hits(lh, e - b, ltmp);
            int64_t take = (e - b) < 2 * k ? (e - b) : 2 * k; /* (* k 2) :284 */
            memcpy(merged + m, lh, sizeof(orc_hit) * (size_t)take);
            m += take;
        }
        if (c->out_probes)
            for (int32_t pr = np; pr < c->nprobe; ++pr) c->out_probes[qi * c->nprobe + pr] = -1;
        stable_sort_hits(merged, m, mtmp); /* :291-294 */
        for (int64_t j = 0; j < k; ++j) {
            if (j < m) {
                c->out_ids[qi * k + j] = merged[j].id;
                c->out_dist[qi * k + j] = merged[j].dist;
            } else {
                c->out_ids[qi * k + j] = -1;
                c->out_dist[qi * k + j] = INFINITY;
            }
        }
    }
    free(q);
    free(ch);
    free(ctmp);
    free(lh);
    free(ltmp);
    free(merged);
    free(mtmp);
}

ORC_EXPORT void orc_ivf_search(const float *rows, int64_t n, int64_t d, const double *cents,
                               int32_t nlist, const int64_t *list_offsets, const int64_t *list_rows,
                               const double *norms, const float *queries, int64_t nq, int64_t k,
                               int32_t nprobe, int coarse_metric, int64_t *out_ids, double *out_dist,
                               int32_t *out_probes, int nthreads) {
    ivf_ctx c = {rows, n,      d, cents,  nlist,         list_offsets, list_rows, norms,
                 queries, k, nprobe, coarse_metric, out_ids, out_dist, out_probes};
    parallel_for(nq, nthreads, 1, ivf_task, &c);
}

/* ------------------------------------------------------------------------------------------
 * HNSW (src/hnsw/ultra_fast.clj).  Restated with two declared deviations, both forced:
 *  (1) level assignment uses a SEEDED java.util.Random (the reference's is an unseeded
 *      thread-local, :139-147, so its graphs are not reproducible at all);
 *  (2) neighbour sets iterate in insertion order (the reference iterates java.util.HashSet<String>
 *      in String-hash bucket order, which depends on the id strings).
 * java.util.PriorityQueue's sift-up/sift-down are restated exactly, so heap-array order (which
 * the reference leaks through `.forEach nearest`, :207-212, and `(take m candidates)`, :255)
 * matches a JVM run with the same comparison results.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    double dist;
    int32_t id;
} hcand;
typedef struct {
    hcand *a;
    int32_t size, cap;
    int reversed;
} jpq;
static inline int jpq_cmp(const jpq *q, const hcand *x, const hcand *y) {
    /* Candidate.compareTo = Double/compare (:115-118); reverseOrder swaps */
    const hcand *l = q->reversed ? y : x, *r = q->reversed ? x : y;
    if (l->dist < r->dist) return -1;
    if (l->dist > r->dist) return 1;
    return 0; /* NaN / signed-zero refinements of Double.compare are not exercised */
}
static void jpq_init(jpq *q, int32_t cap, int reversed) {
    q->cap = cap < 4 ? 4 : cap;
    q->a = (hcand *)malloc(sizeof(hcand) * (size_t)q->cap);
    q->size = 0;
    q->reversed = reversed;
}
static void jpq_free(jpq *q) { free(q->a); }
static void jpq_offer(jpq *q, hcand x) { /* PriorityQueue.offer -> siftUp */
    if (q->size == q->cap) {
        q->cap *= 2;
        q->a = (hcand *)realloc(q->a, sizeof(hcand) * (size_t)q->cap);
    }
    int32_t k = q->size++;
    while (k > 0) {
        int32_t parent = (k - 1) >> 1;
        if (jpq_cmp(q, &x, &q->a[parent]) >= 0) break;
        q->a[k] = q->a[parent];
        k = parent;
    }
    q->a[k] = x;
}
static hcand jpq_poll(jpq *q) { /* PriorityQueue.poll -> siftDown of the last element */
    hcand result = q->a[0];
    int32_t n = --q->size;
    if (n > 0) {
        hcand x = q->a[n];
        int32_t k = 0, half = n >> 1;
        while (k < half) {
            int32_t child = 2 * k + 1, right = child + 1;
            if (right < n && jpq_cmp(q, &q->a[child], &q->a[right]) > 0) child = right;
            if (jpq_cmp(q, &x, &q->a[child]) <= 0) break;
            q->a[k] = q->a[child];
            k = child;
        }
        q->a[k] = x;
    }
    return result;
}

typedef struct {
    int32_t *ids;
    int32_t size, cap;
} nbset; /* insertion-ordered set */
static int nbset_contains(const nbset *s, int32_t id) {
    for (int32_t i = 0; i < s->size; ++i)
        if (s->ids[i] == id) return 1;
    return 0;
}
static void nbset_add(nbset *s, int32_t id) {
    if (nbset_contains(s, id)) return;
    if (s->size == s->cap) {
        s->cap = s->cap ? s->cap * 2 : 8;
        s->ids = (int32_t *)realloc(s->ids, sizeof(int32_t) * (size_t)s->cap);
    }
    s->ids[s->size++] = id;
}

typedef struct {
    const float *rows;
    int64_t n, d;
    int metric;
    int32_t M, maxM, efc;
    double ml;
    int32_t *level;  /* per node */
    nbset **nbrs;    /* per node: array[level+1] */
    int32_t entry;   /* -1 if empty */
    int64_t count;
    orc_rng rng;
    uint32_t *visit_stamp; /* visited set as stamps */
    uint32_t stamp;
    int64_t n_dist; /* distance evaluations (statistics) */
} orc_hnsw;

static double hnsw_dist_q(orc_hnsw *g, const double *q, int32_t id) {
    g->n_dist++;
    return dist_fn_fq(g->metric, g->rows + (int64_t)id * g->d, q, g->d);
    /* cosine-distance-ultra(query, vector): dot/norm sums are symmetric in the arguments */
}

/* :151-212 search-layer-ultra.  Returns ids in `nearest` heap-array order. */
static int32_t hnsw_search_layer(orc_hnsw *g, const double *q, const int32_t *eps, int32_t neps,
                                 int32_t num_closest, int32_t level, int32_t *out) {
    jpq cand, nearest;
    jpq_init(&cand, num_closest, 0);
    jpq_init(&nearest, num_closest, 1);
    if (++g->stamp == 0) {
        memset(g->visit_stamp, 0, sizeof(uint32_t) * (size_t)g->n);
        g->stamp = 1;
    }
    for (int32_t i = 0; i < neps; ++i) { /* :162-167 */
        hcand c = {hnsw_dist_q(g, q, eps[i]), eps[i]};
        g->visit_stamp[eps[i]] = g->stamp;
        jpq_offer(&cand, c);
        jpq_offer(&nearest, c);
    }
    while (cand.size > 0) { /* :170 */
        hcand cur = jpq_poll(&cand);
        /* :175-178: no break on failure, keeps polling */
        if (!(nearest.size < num_closest ||
              cur.dist <= (nearest.size == 0 ? 1.7976931348623157e308 : nearest.a[0].dist)))
            continue;
        if (level > g->level[cur.id]) continue; /* :181 */
        nbset *s = &g->nbrs[cur.id][level];
        for (int32_t t = 0; t < s->size; ++t) { /* :185-204 */
            int32_t nb = s->ids[t];
            if (g->visit_stamp[nb] == g->stamp) continue;
            g->visit_stamp[nb] = g->stamp;
            double dist = hnsw_dist_q(g, q, nb);
            if (nearest.size < num_closest ||
                dist < (nearest.size == 0 ? 1.7976931348623157e308 : nearest.a[0].dist)) {
                hcand c = {dist, nb};
                jpq_offer(&cand, c);
                jpq_offer(&nearest, c);
                if (nearest.size > num_closest) jpq_poll(&nearest);
            }
        }
    }
    int32_t m = nearest.size;
    for (int32_t i = 0; i < m; ++i) out[i] = nearest.a[i].id; /* :207-212 heap-array order */
    jpq_free(&cand);
    jpq_free(&nearest);
    return m;
}

/* :279-299 prune-connections-ultra: stable sort-by distance over the set's iteration order,
 * clear, re-add the closest max-conns. */
static void hnsw_prune(orc_hnsw *g, int32_t node, int32_t level, int32_t max_conns) {
    nbset *s = &g->nbrs[node][level];
    if (s->size <= max_conns) return;
    int64_t d = g->d;
    double *nv = (double *)malloc(sizeof(double) * (size_t)d);
    for (int64_t j = 0; j < d; ++j) nv[j] = (double)g->rows[(int64_t)node * d + j];
    orc_hit *h = (orc_hit *)malloc(sizeof(orc_hit) * (size_t)s->size);
    orc_hit *tmp = (orc_hit *)malloc(sizeof(orc_hit) * (size_t)s->size);
    for (int32_t i = 0; i < s->size; ++i) {
        h[i].dist = hnsw_dist_q(g, nv, s->ids[i]);
        h[i].id = s->ids[i];
    }
    stable_sort_hits(h, s->size, tmp);
    for (int32_t i = 0; i < max_conns; ++i) s->ids[i] = (int32_t)h[i].id;
    s->size = max_conns;
    free(nv);
    free(h);
    free(tmp);
}

ORC_EXPORT orc_hnsw *orc_hnsw_create(const float *rows, int64_t n, int64_t d, int metric, int32_t M,
                                     int32_t ef_construction, int64_t level_seed) {
    orc_hnsw *g = (orc_hnsw *)calloc(1, sizeof(orc_hnsw));
    g->rows = rows;
    g->n = n;
    g->d = d;
    g->metric = metric;
    g->M = M;
    g->maxM = 2 * M; /* :131 */
    g->efc = ef_construction;
    g->ml = 1.0 / log(2.0); /* :133 */
    g->level = (int32_t *)calloc((size_t)(n > 0 ? n : 1), sizeof(int32_t));
    g->nbrs = (nbset **)calloc((size_t)(n > 0 ? n : 1), sizeof(nbset *));
    g->entry = -1;
    g->visit_stamp = (uint32_t *)calloc((size_t)(n > 0 ? n : 1), sizeof(uint32_t));
    orc_rng_init(&g->rng, level_seed);
    return g;
}
ORC_EXPORT void orc_hnsw_free(orc_hnsw *g) {
    if (!g) return;
    for (int64_t i = 0; i < g->n; ++i)
        if (g->nbrs[i]) {
            for (int32_t l = 0; l <= g->level[i]; ++l) free(g->nbrs[i][l].ids);
            free(g->nbrs[i]);
        }
    free(g->level);
    free(g->nbrs);
    free(g->visit_stamp);
    free(g);
}

/* :216-275 insert-single, rows inserted in row order by orc_hnsw_build (:303-330). */
static void hnsw_insert(orc_hnsw *g, int32_t id) {
    /* :143-147 (long)(ml * -log(nextDouble)) */
    int32_t level = (int32_t)(int64_t)(g->ml * (-log(orc_rng_next_double(&g->rng))));
    g->level[id] = level;
    g->nbrs[id] = (nbset *)calloc((size_t)level + 1, sizeof(nbset));
    g->count++;
    if (g->entry < 0) g->entry = id;
    if (g->count > 1) {
        int32_t entry = g->entry;
        int32_t entry_level = g->level[entry];
        int64_t d = g->d;
        double *v = (double *)malloc(sizeof(double) * (size_t)d);
        for (int64_t j = 0; j < d; ++j) v[j] = (double)g->rows[(int64_t)id * d + j];
        int32_t cap = g->efc > 1 ? g->efc : 1;
        int32_t *nearest = (int32_t *)malloc(sizeof(int32_t) * (size_t)(cap + 1));
        int32_t *cands = (int32_t *)malloc(sizeof(int32_t) * (size_t)(cap + 1));
        int32_t nn = 1;
        nearest[0] = entry;
        for (int32_t lc = level < entry_level ? level : entry_level; lc >= 0; --lc) { /* :242 */
            int32_t nc = hnsw_search_layer(g, v, nearest, nn, lc > 0 ? 1 : g->efc, lc, cands);
            int32_t m = lc == 0 ? g->maxM : g->M;
            int32_t take = nc < m ? nc : m; /* (take m candidates) of heap-array order :255 */
            for (int32_t t = 0; t < take; ++t) {
                int32_t nb = cands[t];
                if (lc <= g->level[nb]) {
                    nbset_add(&g->nbrs[id][lc], nb);
                    nbset_add(&g->nbrs[nb][lc], id);
                    if (g->nbrs[nb][lc].size > m) hnsw_prune(g, nb, lc, m);
                }
            }
            memcpy(nearest, cands, sizeof(int32_t) * (size_t)nc);
            nn = nc;
        }
        free(v);
        free(nearest);
        free(cands);
    }
    if (level > g->level[g->entry]) g->entry = id; /* :271-273 */
}
ORC_EXPORT void orc_hnsw_build(orc_hnsw *g) {
    for (int64_t i = 0; i < g->n; ++i) hnsw_insert(g, (int32_t)i);
}
/* Replace the graph by a given one (test infrastructure for graphs built elsewhere, e.g. the bulk k-NN graph of the
 * benchmark): per-node level, entry point and one CSR adjacency per level in neighbour iteration order — the same
 * arrays orc_hnsw_export_level produces.  The traversal (:151-212, :346-374) then runs on it unchanged. */
ORC_EXPORT void orc_hnsw_import(orc_hnsw *g, const int32_t *levels, int32_t entry, int32_t max_level,
                                const int64_t *const *level_offsets, const int32_t *const *level_ids) {
    for (int64_t i = 0; i < g->n; ++i) {
        if (g->nbrs[i]) {
            for (int32_t l = 0; l <= g->level[i]; ++l) free(g->nbrs[i][l].ids);
            free(g->nbrs[i]);
        }
        int32_t lv = levels[i];
        g->level[i] = lv;
        g->nbrs[i] = (nbset *)calloc((size_t)lv + 1, sizeof(nbset));
        for (int32_t l = 0; l <= lv && l <= max_level; ++l) {
            int64_t a = level_offsets[l][i], b = level_offsets[l][i + 1];
            nbset *s = &g->nbrs[i][l];
            s->size = s->cap = (int32_t)(b - a);
            if (b > a) {
                s->ids = (int32_t *)malloc(sizeof(int32_t) * (size_t)(b - a));
                memcpy(s->ids, level_ids[l] + a, sizeof(int32_t) * (size_t)(b - a));
            }
        }
    }
    g->entry = entry;
    g->count = g->n;
}
ORC_EXPORT int32_t orc_hnsw_entry(const orc_hnsw *g) { return g->entry; }
ORC_EXPORT int32_t orc_hnsw_max_level(const orc_hnsw *g) { return g->entry < 0 ? -1 : g->level[g->entry]; }
ORC_EXPORT void orc_hnsw_levels(const orc_hnsw *g, int32_t *out) {
    memcpy(out, g->level, sizeof(int32_t) * (size_t)g->n);
}
/* adjacency of one level as CSR (offsets[n+1], ids); nodes absent from the level get no edges.
 * Call with ids == NULL to size. */
ORC_EXPORT int64_t orc_hnsw_export_level(const orc_hnsw *g, int32_t level, int64_t *offsets, int32_t *ids) {
    int64_t tot = 0;
    for (int64_t i = 0; i < g->n; ++i) {
        if (offsets) offsets[i] = tot;
        if (g->nbrs[i] && level <= g->level[i]) {
            const nbset *s = &g->nbrs[i][level];
            if (ids) memcpy(ids + tot, s->ids, sizeof(int32_t) * (size_t)s->size);
            tot += s->size;
        }
    }
    if (offsets) offsets[g->n] = tot;
    return tot;
}

/* :346-374 search-knn with ef = (max k 50) (:355); `ef_override` > 0 replaces it (extension used
 * by config C5's efSearch=128; the reference has no such knob, src/hnsw/wip/search_config.clj:14). */
ORC_EXPORT int32_t orc_hnsw_search(orc_hnsw *g, const float *query, int32_t k, int32_t ef_override,
                                   int64_t *out_ids, double *out_dist) {
    for (int32_t j = 0; j < k; ++j) {
        out_ids[j] = -1;
        out_dist[j] = INFINITY;
    }
    if (g->entry < 0 || g->count == 0) return 0;
    int64_t d = g->d;
    double *q = (double *)malloc(sizeof(double) * (size_t)d);
    for (int64_t j = 0; j < d; ++j) q[j] = (double)query[j];
    int32_t ef = ef_override > 0 ? ef_override : (k > 50 ? k : 50);
    int32_t cap = ef > 1 ? ef : 1;
    int32_t *nearest = (int32_t *)malloc(sizeof(int32_t) * (size_t)(cap + 1));
    int32_t *cands = (int32_t *)malloc(sizeof(int32_t) * (size_t)(cap + 1));
    int32_t nn = 1;
    nearest[0] = g->entry;
    for (int32_t level = g->level[g->entry]; level >= 0; --level) {
        int32_t nc = hnsw_search_layer(g, q, nearest, nn, level > 0 ? 1 : ef, level, cands);
        memcpy(nearest, cands, sizeof(int32_t) * (size_t)nc);
        nn = nc;
    }
    /* :364-370 re-score, stable sort-by :distance, take k */
    orc_hit *h = (orc_hit *)malloc(sizeof(orc_hit) * (size_t)(nn > 0 ? nn : 1));
    orc_hit *tmp = (orc_hit *)malloc(sizeof(orc_hit) * (size_t)(nn > 0 ? nn : 1));
    for (int32_t i = 0; i < nn; ++i) {
        h[i].dist = hnsw_dist_q(g, q, nearest[i]);
        h[i].id = nearest[i];
    }
    stable_sort_hits(h, nn, tmp);
    int32_t m = nn < k ? nn : k;
    for (int32_t j = 0; j < m; ++j) {
        out_ids[j] = h[j].id;
        out_dist[j] = h[j].dist;
    }
    free(q);
    free(nearest);
    free(cands);
    free(h);
    free(tmp);
    return m;
}
ORC_EXPORT int64_t orc_hnsw_dist_evals(const orc_hnsw *g) { return g->n_dist; }

/* Batched neighbour-candidate scoring: scores[p] = distance-fn(query[pair_query[p]], row[pair_row[p]])
 * — the call at src/hnsw/ultra_fast.clj:192, batched as sketched in
 * src/hnsw/wip/parallel_build.clj:106-118. */
ORC_EXPORT void orc_gather_score(const float *rows, int64_t d, const float *queries, const int32_t *pair_query,
                                 const int32_t *pair_row, int64_t npairs, int metric, double *out) {
    double *q = (double *)malloc(sizeof(double) * (size_t)d);
    for (int64_t p = 0; p < npairs; ++p) {
        const float *qs = queries + (int64_t)pair_query[p] * d;
        for (int64_t j = 0; j < d; ++j) q[j] = (double)qs[j];
        const float *v = rows + (int64_t)pair_row[p] * d;
        if (metric == ORC_IP) out[p] = dot_fq(v, q, d);
        else out[p] = dist_fn_fq(metric, v, q, d);
    }
    free(q);
}

/* ------------------------------------------------------------------------------------------
 * Hybrid LSH (src/hnsw/ann/hash/hybrid_lsh.clj) — §8 f3: its bucket scan is search-partition's arithmetic (A.2) over a
 * candidate set chosen by hashing.  8 tables x 12 bits (:12-14); only the first 12 of the 64 projection rows of a table
 * reach the bucket id (hash-to-bucket-id, :47-55).
 * ---------------------------------------------------------------------------------------- */
enum { ORC_LSH_TABLES = 8, ORC_LSH_BITS = 12, ORC_LSH_PROJ = 64 };

/* generate-random-matrix x NUM-HASH-TABLES from one Random(42) (:24-31, :77-81): table by table, row by row.
 * out: [8][64][d] fp64. */
ORC_EXPORT void orc_lsh_matrices(int64_t d, int64_t seed, double *out) {
    orc_rng r;
    orc_rng_init(&r, seed);
    for (int64_t i = 0; i < (int64_t)ORC_LSH_TABLES * ORC_LSH_PROJ * d; ++i) out[i] = orc_rng_next_gaussian(&r);
}

/* compute-hash-vector + hash-to-bucket-id (:33-55): bit i = (sum_j v[j] * row_i[j] >= 0.0), i < 12. */
static int32_t lsh_bucket(const double *v, const double *table, int64_t d) {
    int32_t id = 0;
    for (int i = 0; i < ORC_LSH_BITS; ++i) {
        const double *row = table + (int64_t)i * d;
        double sum = 0.0;
        for (int64_t j = 0; j < d; ++j) sum = sum + v[j] * row[j];
        if (sum >= 0.0) id |= 1 << i;
    }
    return id;
}
/* bucket ids of every row in every table: out [n][8] int32 */
ORC_EXPORT void orc_lsh_hash(const float *rows, int64_t n, int64_t d, const double *matrices, int32_t *out) {
    double *v = (double *)malloc(sizeof(double) * (size_t)d);
    for (int64_t r = 0; r < n; ++r) {
        for (int64_t j = 0; j < d; ++j) v[j] = (double)rows[r * d + j];
        for (int t = 0; t < ORC_LSH_TABLES; ++t)
            out[r * ORC_LSH_TABLES + t] = lsh_bucket(v, matrices + (int64_t)t * ORC_LSH_PROJ * d, d);
    }
    free(v);
}

/* search-bucket-brute-force (:147-193) appended to `res`: every member if the bucket holds <= limit rows (bucket order),
 * else the stable sort by distance cut to limit.  members = row ids of the bucket in insertion (data) order. */
static int64_t lsh_scan_bucket(const float *rows, int64_t d, const double *norms, const int64_t *members, int64_t size,
                               const double *q, double qnorm, int64_t limit, orc_hit *res, orc_hit *tmp) {
    for (int64_t i = 0; i < size; ++i) {
        const float *v = rows + members[i] * d;
        double dot = 0.0;
        for (int64_t j = 0; j < d; ++j) dot = dot + (double)v[j] * q[j];
        res[i].dist = 1.0 - dot / (qnorm * norms[members[i]]);
        res[i].id = members[i];
    }
    if (size <= limit) return size;
    stable_sort_hits(res, size, tmp);
    return limit;
}

/* search-hybrid (:195-259; multiprobe = 0) and search-hybrid-multiprobe (:261-342; multiprobe = 1) for one query.
 * Candidates are concatenated table by table (the sequential branch; the parallel branch of search-hybrid adds them in
 * the same order, :217-220, with k*3 instead of k*2 per bucket — pass main_mult = 3; the parallel branch of the
 * multi-probe search lets its tasks interleave, which only reorders equal distances of different rows), then
 * deduplicated by id keeping the first, stable-sorted, cut to k.
 * bucket CSR per table: off [8][4097], members [8][n]. */
ORC_EXPORT void orc_lsh_search(const float *rows, int64_t n, int64_t d, const double *norms, const double *matrices,
                               const int64_t *bucket_off, const int64_t *bucket_members, const float *queries, int64_t nq,
                               int64_t k, int32_t num_probes, int32_t probe_radius, int multiprobe, int32_t main_mult,
                               int64_t *out_ids, double *out_dist) {
    const int64_t nb = (int64_t)1 << ORC_LSH_BITS;
    const int probes = num_probes < ORC_LSH_TABLES ? num_probes : ORC_LSH_TABLES;
    const int radius = probe_radius < ORC_LSH_BITS ? probe_radius : ORC_LSH_BITS;
    double *q = (double *)malloc(sizeof(double) * (size_t)d);
    orc_hit *cand = (orc_hit *)malloc(sizeof(orc_hit) * (size_t)(n * (1 + ORC_LSH_BITS) * ORC_LSH_TABLES + 1));
    orc_hit *buf = (orc_hit *)malloc(sizeof(orc_hit) * (size_t)(n + 1));
    orc_hit *tmp = (orc_hit *)malloc(sizeof(orc_hit) * (size_t)(n * (1 + ORC_LSH_BITS) * ORC_LSH_TABLES + 1));
    unsigned char *seen = (unsigned char *)malloc((size_t)n + 1);
    for (int64_t qi = 0; qi < nq; ++qi) {
        for (int64_t j = 0; j < d; ++j) q[j] = (double)queries[qi * d + j];
        const double qnorm = sqrt(sumsq_d(q, d)); /* compute-vector-norm, :57-64 */
        int64_t nc = 0;
        for (int t = 0; t < probes; ++t) {
            const int32_t base = lsh_bucket(q, matrices + (int64_t)t * ORC_LSH_PROJ * d, d);
            const int64_t *off = bucket_off + (int64_t)t * (nb + 1);
            const int64_t *mem = bucket_members + (int64_t)t * n;
            int64_t got = lsh_scan_bucket(rows, d, norms, mem + off[base], off[base + 1] - off[base], q, qnorm,
                                          k * (multiprobe ? 2 : main_mult), buf, tmp);
            if (off[base + 1] > off[base]) { /* (when bucket ...): an absent bucket adds nothing */
                memcpy(cand + nc, buf, sizeof(orc_hit) * (size_t)got);
                nc += got;
            }
            if (multiprobe)
                for (int bit = 0; bit < radius; ++bit) { /* :301-308 */
                    const int32_t nbk = (base ^ (1 << bit)) & (int32_t)(nb - 1);
                    if (off[nbk + 1] == off[nbk]) continue;
                    got = lsh_scan_bucket(rows, d, norms, mem + off[nbk], off[nbk + 1] - off[nbk], q, qnorm, k, buf, tmp);
                    memcpy(cand + nc, buf, sizeof(orc_hit) * (size_t)got);
                    nc += got;
                }
        }
        memset(seen, 0, (size_t)n);
        int64_t nu = 0;
        for (int64_t i = 0; i < nc; ++i)
            if (!seen[cand[i].id]) {
                seen[cand[i].id] = 1;
                cand[nu++] = cand[i];
            }
        stable_sort_hits(cand, nu, tmp);
        for (int64_t j = 0; j < k; ++j) {
            out_ids[qi * k + j] = j < nu ? cand[j].id : -1;
            out_dist[qi * k + j] = j < nu ? cand[j].dist : INFINITY;
        }
    }
    free(q);
    free(cand);
    free(buf);
    free(tmp);
    free(seen);
}

/* ------------------------------------------------------------------------------------------
 * a4: the float[] Vector-API variants (src/hnsw/simd.clj:18-115) and their only consumer, PCAF
 * (src/hnsw/ann/dimreduct/pcaf.clj).
 *
 * Per SPECIES-LENGTH chunk the reference multiplies fp32 lanes (FloatVector.mul), reduces them with
 * reduceLanes(VectorOperators/ADD) and adds the (float) chunk sum into a double accumulator
 * ((+ sum (double (.reduceLanes ...))), simd.clj:33, :56, :96-98); the tail (len mod lanes) is scalar in
 * double (:37-43, :100-110).  The JDK leaves the lane order of a floating-point ADD reduction unspecified;
 * restated here is the order of its scalar fallback and of HotSpot's ordered AddReductionVF: fp32 adds
 * left to right starting from 0.0f.  `lanes` = SPECIES_PREFERRED.length() of the machine the reference ran on
 * (4 / 8 / 16; 8 = AVX2 is the default of the tests).  Another lane order or width moves the result by
 * a few fp32 ulps of a chunk sum: the north star's 1e-5 relative bound for fp32 covers it
 * (tests/test_oracle.py::test_simd_lane_order_is_within_the_fp32_bound).
 * ---------------------------------------------------------------------------------------- */
static inline float lane_sum_mul(const float *a, const float *b, int lanes) {
    float s = 0.0f;
    for (int i = 0; i < lanes; ++i) {
        float p = a[i] * b[i];
        s = s + p;
    }
    return s;
}
/* dot-product-simd-optimized, simd.clj:18-43 */
ORC_EXPORT double orc_simd_dot(const float *a, const float *b, int64_t d, int32_t lanes) {
    const int64_t ub = d - d % lanes;
    double sum = 0.0;
    for (int64_t i = 0; i < ub; i += lanes) sum = sum + (double)lane_sum_mul(a + i, b + i, lanes);
    for (int64_t j = ub; j < d; ++j) sum = sum + (double)a[j] * (double)b[j];
    return sum;
}
/* euclidean-distance-simd-optimized, simd.clj:45-71: diff and square in fp32 lanes */
ORC_EXPORT double orc_simd_euclidean(const float *a, const float *b, int64_t d, int32_t lanes) {
    const int64_t ub = d - d % lanes;
    double sum = 0.0;
    for (int64_t i = 0; i < ub; i += lanes) {
        float s = 0.0f;
        for (int l = 0; l < lanes; ++l) {
            float df = a[i + l] - b[i + l];
            float sq = df * df;
            s = s + sq;
        }
        sum = sum + (double)s;
    }
    for (int64_t j = ub; j < d; ++j) {
        double df = (double)a[j] - (double)b[j];
        sum = sum + df * df;
    }
    return sqrt(sum);
}
/* cosine-distance-simd-optimized, simd.clj:73-115: three accumulators, (zero? magnitude) -> 1.0 */
ORC_EXPORT double orc_simd_cosine(const float *a, const float *b, int64_t d, int32_t lanes) {
    const int64_t ub = d - d % lanes;
    double dot = 0.0, na = 0.0, nb = 0.0;
    for (int64_t i = 0; i < ub; i += lanes) {
        dot = dot + (double)lane_sum_mul(a + i, b + i, lanes);
        na = na + (double)lane_sum_mul(a + i, a + i, lanes);
        nb = nb + (double)lane_sum_mul(b + i, b + i, lanes);
    }
    for (int64_t j = ub; j < d; ++j) {
        double x = (double)a[j], y = (double)b[j];
        dot = dot + x * y;
        na = na + x * x;
        nb = nb + y * y;
    }
    double mag = sqrt(na) * sqrt(nb);
    if (mag == 0.0) return 1.0;
    return 1.0 - dot / mag;
}
/* the same three with another lane order (pairwise tree): used only to show the sensitivity to the unspecified order */
ORC_EXPORT double orc_simd_cosine_tree(const float *a, const float *b, int64_t d, int32_t lanes) {
    const int64_t ub = d - d % lanes;
    double acc[3] = {0.0, 0.0, 0.0};
    float t[3][16];
    for (int64_t i = 0; i < ub; i += lanes) {
        for (int l = 0; l < lanes; ++l) {
            t[0][l] = a[i + l] * b[i + l];
            t[1][l] = a[i + l] * a[i + l];
            t[2][l] = b[i + l] * b[i + l];
        }
        for (int w = lanes / 2; w >= 1; w /= 2)
            for (int v = 0; v < 3; ++v)
                for (int l = 0; l < w; ++l) t[v][l] = t[v][l] + t[v][l + w];
        for (int v = 0; v < 3; ++v) acc[v] = acc[v] + (double)t[v][0];
    }
    for (int64_t j = ub; j < d; ++j) {
        double x = (double)a[j], y = (double)b[j];
        acc[0] = acc[0] + x * y;
        acc[1] = acc[1] + x * x;
        acc[2] = acc[2] + y * y;
    }
    double mag = sqrt(acc[1]) * sqrt(acc[2]);
    if (mag == 0.0) return 1.0;
    return 1.0 - acc[0] / mag;
}

/* create-random-projection, pcaf.clj:33-46: Random(42), scale = (float)(1 / sqrt(target)), matrix[i] =
 * scale * (float) nextGaussian — Clojure multiplies two floats in double and aset narrows the product to float.
 * out: [target][original] fp32. */
ORC_EXPORT void orc_pcaf_matrix(int64_t original_dim, int64_t target_dim, int64_t seed, float *out) {
    orc_rng r;
    orc_rng_init(&r, seed);
    const float scale = (float)(1.0 / sqrt((double)target_dim));
    for (int64_t i = 0; i < target_dim * original_dim; ++i) {
        const float g = (float)orc_rng_next_gaussian(&r);
        out[i] = (float)((double)scale * (double)g);
    }
}
/* project-vector-simd, pcaf.clj:48-81: result[i] = (float) of the chunked dot of matrix row i and the vector (the loop
 * local `sum` is initialised with (float 0.0), which Clojure widens to a double loop local; the tail multiplies two
 * floats in double). */
ORC_EXPORT void orc_pcaf_project(const float *matrix, int64_t original_dim, int64_t target_dim, const float *rows, int64_t n,
                                 int32_t lanes, float *out) {
    for (int64_t r = 0; r < n; ++r)
        for (int64_t t = 0; t < target_dim; ++t)
            out[r * target_dim + t] = (float)orc_simd_dot(matrix + t * original_dim, rows + r * original_dim, original_dim, lanes);
}
/* search-pcaf-parallel, pcaf.clj:195-253, for a batch: phase 1 = simd cosine of the projected query against every
 * projected row, stable sort, take min(k-filter, 3k) (:229-230); phase 2 = simd cosine in the full dimension for those
 * (:236-243), stable sort of the candidates in phase-1 order (:246-250), take k.  The reference iterates a hash map
 * in phase 1, so exact low-dim distance ties fall in hash order there; rows in data order here. */
ORC_EXPORT void orc_pcaf_search(const float *rows, const float *low_rows, int64_t n, int64_t d, int64_t t, const float *matrix,
                                const float *queries, int64_t nq, int64_t k, int64_t k_filter, int32_t lanes,
                                int64_t *out_ids, double *out_dist) {
    orc_hit *h = (orc_hit *)malloc(sizeof(orc_hit) * (size_t)(n + 1));
    orc_hit *tmp = (orc_hit *)malloc(sizeof(orc_hit) * (size_t)(n + 1));
    float *lq = (float *)malloc(sizeof(float) * (size_t)t);
    for (int64_t qi = 0; qi < nq; ++qi) {
        const float *q = queries + qi * d;
        orc_pcaf_project(matrix, d, t, q, 1, lanes, lq);
        for (int64_t r = 0; r < n; ++r) {
            h[r].dist = orc_simd_cosine(lq, low_rows + r * t, t, lanes);
            h[r].id = r;
        }
        stable_sort_hits(h, n, tmp);
        int64_t c = k_filter < 3 * k ? k_filter : 3 * k;
        if (c > n) c = n;
        for (int64_t i = 0; i < c; ++i) h[i].dist = orc_simd_cosine(q, rows + h[i].id * d, d, lanes);
        stable_sort_hits(h, c, tmp);
        for (int64_t j = 0; j < k; ++j) {
            out_ids[qi * k + j] = j < c ? h[j].id : -1;
            out_dist[qi * k + j] = j < c ? h[j].dist : INFINITY;
        }
    }
    free(h);
    free(tmp);
    free(lq);
}
