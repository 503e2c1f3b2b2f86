/* TEST INFRASTRUCTURE — CPU oracle, NOT product code.
 *
 * Plain-C restatement of the algorithms on the three hot paths of zyx7k/HAI-25-RAG-on-Edge, used only by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as the CHECKER.
 * The product (libvsb200.so) never links, loads or calls anything in this file.
 *
 * Paths restated (file:line relative to /root/reference):
 *   HP1  exact L2          cpu/cpu_baseline.cpp:95-153, 222-248
 *   HP2  IVF search        qidk_ivf/android/app/main/jni/IVFIndex.cpp:269-358, 449-496, 640-859
 *   INT8 brute force       qidk_bruteforce/android/app/main/jni/QnnRunner.cpp:13-55, main.cpp:36-71
 *
 * Pinning status:
 *   HP1  pinned against the UNMODIFIED reference compiled here (oracle/_ref/ref_driver, built by
 *        oracle/Makefile) — tests/golden/hp1_*.npz were produced by that binary.
 *   HP2  parity unpinned: IVFIndex.cpp needs arm_neon.h + the QNN SDK and has a syntax defect at :498;
 *        the reference holds no golden vectors for it.
 *   INT8 parity unpinned: the MatMul runs inside the proprietary QNN HTP runtime; only the input quantiser
 *        and the u8 top-k are visible in the reference, the requantisation rule is defined by this project.
 *
 * Build: see oracle/Makefile (gcc -O3 -fopenmp -mavx2 -mfma -ffp-contract=off; every fused multiply-add the
 * restatement wants is written as an explicit fmaf()).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------------ */
/* HP1: exact L2                                                                                    */
/* ------------------------------------------------------------------------------------------------ */

/* cpu_baseline.cpp:95-114 compute_norm_avx2: 8 lane accumulators, lane l sums v[l], v[l+8], ... with FMA,
 * lanes then added in order 0..7, scalar tail afterwards (tail contracted to FMA by g++ -O3 -mfma). */
static float orc_norm_row(const float* v, int dim) {
    float lane[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int i = 0;
    for (; i + 7 < dim; i += 8)
        for (int l = 0; l < 8; ++l) lane[l] = fmaf(v[i + l], v[i + l], lane[l]);
    float s = lane[0] + lane[1] + lane[2] + lane[3] + lane[4] + lane[5] + lane[6] + lane[7];
    for (; i < dim; ++i) s = fmaf(v[i], v[i], s);
    return s;
}

/* cpu_baseline.cpp:116-125 compute_norms */
ORC_EXPORT void orc_norms(const float* data, int64_t rows, int dim, float* norms) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < rows; ++i) norms[i] = orc_norm_row(data + i * (int64_t)dim, dim);
}

/* Dot product standing in for cblas_sgemm M=1 (cpu_baseline.cpp:229-237).  OpenBLAS is an un-vendored,
 * unpinned dependency (`-lopenblas`, cpu/README.md:101); its summation order is not part of the
 * reference, so this is *a* fp32 order: 16 lane accumulators by (d mod 16) with FMA, pairwise lane tree.
 * Contract vs the real reference: bit-exact on integer-valued data, <=1e-5 relative otherwise. */
static inline float orc_dot16(const float* a, const float* b, int dim) {
    float lane[16];
    for (int l = 0; l < 16; ++l) lane[l] = 0.f;
    int i = 0;
    for (; i + 15 < dim; i += 16)
        for (int l = 0; l < 16; ++l) lane[l] = fmaf(a[i + l], b[i + l], lane[l]);
    for (int w = 8; w >= 1; w >>= 1)
        for (int l = 0; l < w; ++l) lane[l] = lane[l] + lane[l + w];
    float s = lane[0];
    for (; i < dim; ++i) s = fmaf(a[i], b[i], s);
    return s;
}

typedef struct {
    float v;
    int32_t id;
} orc_pair;

/* cpu_baseline.cpp:127-153 select_topk, literal: seed with first k, track arg-max slot, replace on strict <,
 * rescan, final ascending sort by dist only (std::sort is unstable: order among equal dists unspecified —
 * here: insertion sort on dist, stable w.r.t. slot order). Requires k <= N (UB in the reference otherwise). */
static void orc_select_topk_literal(const float* d, int64_t N, int k, orc_pair* top) {
    for (int i = 0; i < k; ++i) {
        top[i].v = d[i];
        top[i].id = (int32_t)i;
    }
    int mx = 0;
    for (int i = 1; i < k; ++i)
        if (top[i].v > top[mx].v) mx = i;
    for (int64_t j = k; j < N; ++j) {
        if (d[j] < top[mx].v) {
            top[mx].v = d[j];
            top[mx].id = (int32_t)j;
            mx = 0;
            for (int i = 1; i < k; ++i)
                if (top[i].v > top[mx].v) mx = i;
        }
    }
    for (int i = 1; i < k; ++i) {
        orc_pair t = top[i];
        int p = i - 1;
        while (p >= 0 && top[p].v > t.v) {
            top[p + 1] = top[p];
            --p;
        }
        top[p + 1] = t;
    }
}

/* Canonical total order used by the product: (value asc, id asc) for smallest-k. Bounded insertion. */
static void orc_select_topk_canonical(const float* d, int64_t N, int k, orc_pair* top) {
    int n = 0;
    for (int64_t j = 0; j < N; ++j) {
        float v = d[j];
        if (n == k && !(v < top[k - 1].v)) continue; /* ids ascend, so ties lose against kept entries */
        int p = (n < k) ? n : k - 1;
        while (p > 0 && top[p - 1].v > v) {
            top[p] = top[p - 1];
            --p;
        }
        top[p].v = v;
        top[p].id = (int32_t)j;
        if (n < k) ++n;
    }
    for (int i = n; i < k; ++i) {
        top[i].v = INFINITY;
        top[i].id = -1;
    }
}

/* The hot loop of run_benchmark (cpu_baseline.cpp:222-248) for all queries.
 * mode 0 = literal select_topk tie behaviour, 1 = canonical (dist asc, id asc).
 * bnorms/qnorms may be NULL (computed here).  OpenMP over queries (the reference is serial over queries and
 * threads inside OpenBLAS; total work is identical). Returns 0, or -1 on bad arguments. */
ORC_EXPORT int orc_exact_search(const float* base, int64_t nb, int dim, const float* queries, int64_t nq, int k,
                                int mode, const float* bnorms_in, int32_t* out_ids, float* out_dists) {
    if (k <= 0 || nb <= 0 || nq < 0 || (mode == 0 && k > nb)) return -1;
    float* bn = NULL;
    if (!bnorms_in) {
        bn = (float*)malloc(sizeof(float) * (size_t)nb);
        orc_norms(base, nb, dim, bn);
    }
    const float* bnorms = bnorms_in ? bnorms_in : bn;
#pragma omp parallel
    {
        float* dist = (float*)malloc(sizeof(float) * (size_t)nb);
        orc_pair* top = (orc_pair*)malloc(sizeof(orc_pair) * (size_t)k);
#pragma omp for schedule(dynamic, 1)
        for (int64_t i = 0; i < nq; ++i) {
            const float* q = queries + i * (int64_t)dim;
            const float qn = orc_norm_row(q, dim);
            for (int64_t j = 0; j < nb; ++j) {
                float dot = orc_dot16(q, base + j * (int64_t)dim, dim);
                /* cpu_baseline.cpp:241  q_norm + B_norms[j] - 2.0f * distances[j]; 2*x is exact so the
                 * contracted (fnmadd) and uncontracted forms round identically. */
                dist[j] = fmaf(-2.0f, dot, qn + bnorms[j]);
            }
            if (mode == 0)
                orc_select_topk_literal(dist, nb, k, top);
            else
                orc_select_topk_canonical(dist, nb, k, top);
            for (int t = 0; t < k; ++t) {
                out_ids[i * k + t] = top[t].id;
                out_dists[i * k + t] = top[t].v;
            }
        }
        free(dist);
        free(top);
    }
    free(bn);
    return 0;
}

/* Distances of given (query, id) pairs, for the tie-aware comparator: recompute what the oracle would
 * report for an id the product returned. */
ORC_EXPORT void orc_exact_distances_at(const float* base, int dim, const float* queries, int64_t nq, int k,
                                       const int32_t* ids, float* out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nq; ++i) {
        const float* q = queries + i * (int64_t)dim;
        const float qn = orc_norm_row(q, dim);
        for (int t = 0; t < k; ++t) {
            int32_t id = ids[i * k + t];
            if (id < 0) {
                out[i * k + t] = INFINITY;
                continue;
            }
            const float* x = base + (int64_t)id * dim;
            out[i * k + t] = fmaf(-2.0f, orc_dot16(q, x, dim), qn + orc_norm_row(x, dim));
        }
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* HP2: IVF                                                                                         */
/* ------------------------------------------------------------------------------------------------ */

/* IVFIndex.cpp:278-357 computeDotProductsContiguous, one row: four accumulators (NEON float32x4 lanes),
 * lane c sums q[4j+c]*v[4j+c] for j = 0..dim/4-1 with vmlaq_f32 (FMLA on AArch64 => fused), then
 * vaddvq_f32 = (l0+l1)+(l2+l3).  All three unrolled variants (8/4/1 rows) use this order. */
static inline float orc_dot_neon4(const float* q, const float* v, int dim) {
    float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
    for (int j = 0; j < dim; j += 4) {
        l0 = fmaf(q[j + 0], v[j + 0], l0);
        l1 = fmaf(q[j + 1], v[j + 1], l1);
        l2 = fmaf(q[j + 2], v[j + 2], l2);
        l3 = fmaf(q[j + 3], v[j + 3], l3);
    }
    return (l0 + l1) + (l2 + l3);
}

/* Coarse scores S[b][c] = q_b . centroid_c  (create_ivf_model.py:45-64 MatMul; IVFIndex.cpp:653-708).
 * The reference runs this on the HTP (fp16 inside, unpinned); the restatement and the product both use
 * the fine scan's fp32 order so that the probe sets are bit-identical. */
ORC_EXPORT void orc_ivf_coarse(const float* queries, int64_t nq, const float* centroids, int nlist, int dim,
                               float* scores) {
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < nq; ++b)
        for (int c = 0; c < nlist; ++c)
            scores[b * nlist + c] = orc_dot_neon4(queries + b * (int64_t)dim, centroids + (int64_t)c * dim, dim);
}

static int orc_cmp_desc(const void* a, const void* b) {
    const orc_pair* x = (const orc_pair*)a;
    const orc_pair* y = (const orc_pair*)b;
    if (x->v > y->v) return -1;
    if (x->v < y->v) return 1;
    return (x->id > y->id) - (x->id < y->id);
}

/* Top-nprobe of one query's coarse scores, largest first (IVFIndex.cpp:598-599 partial_sort,
 * :711-712 nth_element — set membership at ties and order inside the set are unspecified there; here
 * canonical (score desc, cluster id asc)). */
ORC_EXPORT void orc_ivf_select_probes(const float* scores, int nlist, int nprobe, int32_t* probes) {
    orc_pair* p = (orc_pair*)malloc(sizeof(orc_pair) * (size_t)nlist);
    for (int c = 0; c < nlist; ++c) {
        p[c].v = scores[c];
        p[c].id = c;
    }
    qsort(p, (size_t)nlist, sizeof(orc_pair), orc_cmp_desc);
    for (int i = 0; i < nprobe; ++i) probes[i] = p[i].id;
    free(p);
}

/* binary min-heap on .v only (front = smallest), mirrors std::make/push/pop_heap with comparator a.first > b.first */
static void orc_heap_sift_down(orc_pair* h, int n, int i) {
    for (;;) {
        int l = 2 * i + 1, r = l + 1, m = i;
        if (l < n && h[l].v < h[m].v) m = l;
        if (r < n && h[r].v < h[m].v) m = r;
        if (m == i) return;
        orc_pair t = h[i];
        h[i] = h[m];
        h[m] = t;
        i = m;
    }
}

/* IVFIndex::searchBatch, reordered (list-contiguous) mode, IVFIndex.cpp:674-782, plus scattered mode
 * :783-846 (same candidates and arithmetic; rows addressed through cluster_indices).
 *   vectors      [N x dim]   list-contiguous rows when reorder_map != NULL, original order otherwise
 *   offsets      [nlist+1]
 *   id_map       reordered mode: reorder_to_original [N]; scattered mode: cluster_indices [N]
 *   reordered    1/0
 *   mode 0: literal heap (insert on strict >, arrival order = probe order then position in list)
 *   mode 1: canonical (score desc, original id asc)
 * Outputs padded with id -1 / score -inf beyond counts[b] = min(k, candidates).  Returns total candidates
 * (IVFIndex.cpp:728,858). nprobe is clamped to nlist (:647). */
ORC_EXPORT int64_t orc_ivf_search(const float* vectors, const int32_t* offsets, const int32_t* id_map, int reordered,
                                  int nlist, int dim, const float* coarse_scores, const float* queries, int64_t nq,
                                  int k, int nprobe, int mode, int32_t* out_ids, float* out_scores,
                                  int32_t* out_counts) {
    if (nprobe > nlist) nprobe = nlist;
    int64_t total = 0;
#pragma omp parallel reduction(+ : total)
    {
        int32_t* probes = (int32_t*)malloc(sizeof(int32_t) * (size_t)nprobe);
        orc_pair* heap = (orc_pair*)malloc(sizeof(orc_pair) * (size_t)(k > 0 ? k : 1));
#pragma omp for schedule(dynamic, 4)
        for (int64_t b = 0; b < nq; ++b) {
            const float* q = queries + b * (int64_t)dim;
            orc_ivf_select_probes(coarse_scores + b * nlist, nlist, nprobe, probes);
            int64_t cand = 0;
            for (int i = 0; i < nprobe; ++i) cand += offsets[probes[i] + 1] - offsets[probes[i]];
            total += cand;
            int kact = (int)(cand < k ? cand : k);
            int n = 0;
            for (int i = 0; i < nprobe && kact > 0; ++i) {
                int32_t s = offsets[probes[i]], e = offsets[probes[i] + 1];
                for (int32_t r = s; r < e; ++r) {
                    int32_t row = reordered ? r : id_map[r];
                    int32_t oid = reordered ? id_map[r] : id_map[r];
                    float sc = orc_dot_neon4(q, vectors + (int64_t)row * dim, dim);
                    if (mode == 0) {
                        if (n < kact) {
                            heap[n].v = sc;
                            heap[n].id = oid;
                            if (++n == kact)
                                for (int t = kact / 2 - 1; t >= 0; --t) orc_heap_sift_down(heap, kact, t);
                        } else if (sc > heap[0].v) {
                            heap[0].v = sc;
                            heap[0].id = oid;
                            orc_heap_sift_down(heap, kact, 0);
                        }
                    } else {
                        /* canonical bounded insertion on (score desc, id asc) */
                        if (n == kact) {
                            orc_pair* w = &heap[kact - 1];
                            if (!(sc > w->v || (sc == w->v && oid < w->id))) continue;
                        }
                        int p = (n < kact) ? n : kact - 1;
                        while (p > 0 && (heap[p - 1].v < sc || (heap[p - 1].v == sc && heap[p - 1].id > oid))) {
                            heap[p] = heap[p - 1];
                            --p;
                        }
                        heap[p].v = sc;
                        heap[p].id = oid;
                        if (n < kact) ++n;
                    }
                }
            }
            if (mode == 0) qsort(heap, (size_t)n, sizeof(orc_pair), orc_cmp_desc);
            for (int t = 0; t < k; ++t) {
                out_ids[b * k + t] = t < n ? heap[t].id : -1;
                out_scores[b * k + t] = t < n ? heap[t].v : -INFINITY;
            }
            out_counts[b] = n;
        }
        free(probes);
        free(heap);
    }
    return total;
}

/* Inner products of given (query, original id) pairs in the scan's order (tie-aware comparator). */
ORC_EXPORT void orc_ivf_scores_at(const float* vectors_original_order, int dim, const float* queries, int64_t nq,
                                  int k, const int32_t* ids, float* out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nq; ++i)
        for (int t = 0; t < k; ++t) {
            int32_t id = ids[i * k + t];
            out[i * k + t] = id < 0 ? -INFINITY
                                    : orc_dot_neon4(queries + i * (int64_t)dim,
                                                    vectors_original_order + (int64_t)id * dim, dim);
        }
}

/* k-means assignment step used by the builder (create_ivf_model.py:102-110 -> sklearn KMeans, L2):
 * label = argmin_c ||x - c||^2 evaluated as cn[c] - 2 x.c (x norm is constant per row), ties -> lowest c. */
ORC_EXPORT void orc_kmeans_assign(const float* x, int64_t n, const float* centroids, int nlist, int dim,
                                  int32_t* labels, float* best_dist) {
    float* cn = (float*)malloc(sizeof(float) * (size_t)nlist);
    orc_norms(centroids, nlist, dim, cn);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const float* v = x + i * (int64_t)dim;
        const float xn = orc_norm_row(v, dim);
        float bd = INFINITY;
        int32_t bl = -1;
        for (int c = 0; c < nlist; ++c) {
            float d = fmaf(-2.0f, orc_dot16(v, centroids + (int64_t)c * dim, dim), xn + cn[c]);
            if (d < bd) {
                bd = d;
                bl = c;
            }
        }
        labels[i] = bl;
        if (best_dist) best_dist[i] = bd;
    }
    free(cn);
}

/* ------------------------------------------------------------------------------------------------ */
/* INT8 brute force                                                                                 */
/* ------------------------------------------------------------------------------------------------ */

/* QnnRunner.cpp:13-55 quantize_buffer_neon: x*inv_scale (vmulq_n_f32), +0.5f (vaddq_f32, NOT fused),
 * convert toward zero (vcvtq_s32_f32), saturate to [0,255] (vqmovn/vqmovun).  inv_scale = 1.0f/scale is
 * computed by the caller in fp32 (QnnRunner.cpp:544,585,619). */
ORC_EXPORT void orc_quantize_u8(const float* src, int64_t count, float inv_scale, uint8_t* dst) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < count; ++i) {
        volatile float m = src[i] * inv_scale; /* volatile: keep the product rounded to fp32 */
        float t = m + 0.5f;
        int32_t qv;
        if (t != t)
            qv = 0; /* NaN converts to 0 on AArch64 */
        else if (t >= 2147483648.0f)
            qv = 2147483647;
        else if (t <= -2147483648.0f)
            qv = (int32_t)(-2147483647 - 1);
        else
            qv = (int32_t)t; /* C cast truncates toward zero */
        dst[i] = (uint8_t)(qv < 0 ? 0 : (qv > 255 ? 255 : qv));
    }
}

/* Requantisation rule DEFINED by this project for the opaque HTP MatMul:
 *   acc = sum_d q_u8[d]*w_u8[d]           (int32, exact)
 *   out = sat_u8( floor( fl(fl(acc)*m) + 0.5f ) ),  m = fl(fl(s_in*s_w)/s_out)  — all fp32, not fused. */
static inline uint8_t orc_requant(int32_t acc, float m) {
    volatile float p = (float)acc * m;
    float t = floorf(p + 0.5f);
    return (uint8_t)(t < 0.f ? 0 : (t > 255.f ? 255 : (int)t));
}

ORC_EXPORT float orc_int8_multiplier(float s_in, float s_w, float s_out) {
    volatile float a = s_in * s_w;
    return a / s_out;
}

/* Raw u8 score matrix (the equivalent of QnnRunner::getRawOutputBuffer after executeBatchRaw). */
ORC_EXPORT void orc_int8_scores(const uint8_t* base_u8, int64_t nb, int dim, const uint8_t* q_u8, int64_t nq,
                                float m, uint8_t* out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nq; ++i)
        for (int64_t j = 0; j < nb; ++j) {
            int32_t acc = 0;
            for (int d = 0; d < dim; ++d) acc += (int32_t)q_u8[i * dim + d] * (int32_t)base_u8[j * (int64_t)dim + d];
            out[i * nb + j] = orc_requant(acc, m);
        }
}

/* main.cpp:36-57 find_top_k_int8: size-k min-heap on similarity (front = current min), insert on strict >,
 * final sort descending by similarity.  mode 0 literal (heap order decides ties), 1 canonical
 * (score desc, id asc). Scores are computed on the fly (no [nq x nb] matrix). */
ORC_EXPORT int orc_int8_search(const uint8_t* base_u8, int64_t nb, int dim, const uint8_t* q_u8, int64_t nq, int k,
                               float m, int mode, int32_t* out_ids, uint8_t* out_scores) {
    if (k <= 0 || nb <= 0) return -1;
#pragma omp parallel
    {
        orc_pair* heap = (orc_pair*)malloc(sizeof(orc_pair) * (size_t)k);
#pragma omp for schedule(dynamic, 1)
        for (int64_t i = 0; i < nq; ++i) {
            int n = 0;
            const uint8_t* q = q_u8 + i * dim;
            for (int64_t j = 0; j < nb; ++j) {
                const uint8_t* w = base_u8 + j * (int64_t)dim;
                int32_t acc = 0;
                for (int d = 0; d < dim; ++d) acc += (int32_t)q[d] * (int32_t)w[d];
                float sc = (float)orc_requant(acc, m);
                if (mode == 0) {
                    if (n < k) {
                        /* push_heap */
                        int c = n++;
                        heap[c].v = sc;
                        heap[c].id = (int32_t)j;
                        while (c > 0) {
                            int p = (c - 1) / 2;
                            if (heap[p].v > heap[c].v) {
                                orc_pair t = heap[p];
                                heap[p] = heap[c];
                                heap[c] = t;
                                c = p;
                            } else
                                break;
                        }
                    } else if (sc > heap[0].v) {
                        heap[0].v = sc;
                        heap[0].id = (int32_t)j;
                        orc_heap_sift_down(heap, k, 0);
                    }
                } else {
                    if (n == k && !(sc > heap[k - 1].v)) continue; /* ids ascend: ties lose */
                    int p = (n < k) ? n : k - 1;
                    while (p > 0 && heap[p - 1].v < sc) {
                        heap[p] = heap[p - 1];
                        --p;
                    }
                    heap[p].v = sc;
                    heap[p].id = (int32_t)j;
                    if (n < k) ++n;
                }
            }
            if (mode == 0) qsort(heap, (size_t)n, sizeof(orc_pair), orc_cmp_desc);
            for (int t = 0; t < k; ++t) {
                out_ids[i * k + t] = t < n ? heap[t].id : -1;
                out_scores[i * k + t] = t < n ? (uint8_t)heap[t].v : 0;
            }
        }
        free(heap);
    }
    return 0;
}

ORC_EXPORT int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
