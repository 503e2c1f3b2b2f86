// C ABI of the IVF path (include/vsb200.h, "IVF" section): index handle (from arrays or from the reference's on-disk
// directory format), two-stage search, and the k-means builder whose assignment step reuses the fused
// tensor-core distance + arg-min kernel of the exact path.
#include <cuda.h>
#include <cuda_runtime.h>
#include <sys/stat.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <random>
#include <string>
#include <vector>

#include "../host/ivf_io.hpp"
#include "kernels.cuh"
#include "vsb_common.cuh"

using namespace vsb;

struct vs_ivf {
    int device = 0;
    int64_t n = 0;
    int nlist = 0;
    int dim = 0;
    float avg_cluster_size = 0.f;
    float* d_vectors = nullptr;    // [n x 128] list-contiguous
    int32_t* d_offsets = nullptr;  // [nlist+1]
    int32_t* d_idmap = nullptr;    // [n] list position -> original id
    int32_t* d_order = nullptr;    // [nlist] lists by descending length (work order of the list-major scan)
    float* d_centroids = nullptr;  // [nlist x 128]
    CUtensorMap tmV;
    // tensor-core list-major scan: TF32 hi / lo split of the vectors (hi aliases d_vectors and lo is absent when every
    // component is TF32-exact, e.g. integer SIFT data; lo is then a zero array created on first need)
    float* d_vhi = nullptr;
    float* d_vlo = nullptr;
    bool v_exact = false;
    CUtensorMap tmVhi, tmVlo;
    int* h_flag = nullptr;  // pinned: "queries are not TF32-exact"
    cudaStream_t stream = nullptr;
    DevBuf q, scores, probes, out_ids, out_scores, out_counts, total, lm_ws, part_key, part_id, qhi, qlo, gthr;
    int num_sms = 148;
    unsigned long long* h_total = nullptr;  // pinned
    bool profile = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool ev_valid = false;
    int last_launches = 0;
};

static int ivf_free(vs_ivf* h) {
    if (!h) return VS_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->d_vhi && h->d_vhi != h->d_vectors) cudaFree(h->d_vhi);
    if (h->d_vlo) cudaFree(h->d_vlo);
    for (void* p : {(void*)h->d_vectors, (void*)h->d_offsets, (void*)h->d_idmap, (void*)h->d_centroids, (void*)h->d_order})
        if (p) cudaFree(p);
    for (DevBuf* b : {&h->q, &h->scores, &h->probes, &h->out_ids, &h->out_scores, &h->out_counts, &h->total, &h->lm_ws,
                      &h->part_key, &h->part_id, &h->qhi, &h->qlo, &h->gthr})
        b->release();
    if (h->h_total) cudaFreeHost(h->h_total);
    if (h->h_flag) cudaFreeHost(h->h_flag);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return VS_OK;
}

// vectors: [n x dim] list-contiguous rows; offsets [nlist+1]; id_map [n]; centroids [nlist x dim] — host pointers
static int ivf_create_impl(vs_ivf_t** out, const float* vectors, int64_t n, int dim, const int32_t* offsets, int nlist,
                           const int32_t* id_map, const float* centroids, int device) {
    if (!out) return fail(VS_ERR_INVALID, "out handle is NULL");
    *out = nullptr;
    if (!vectors || !offsets || !id_map || !centroids || n <= 0 || nlist <= 0) return fail(VS_ERR_INVALID, "NULL array or empty index");
    if (dim != 128) return fail(VS_ERR_UNSUPPORTED, "only dim == 128 (SIFT shape) is implemented");
    if (n > 0x7fffffffLL) return fail(VS_ERR_INVALID, "ids must fit int32");
    if (offsets[0] != 0 || offsets[nlist] != n) return fail(VS_ERR_INVALID, "cluster offsets do not cover [0, n)");
    for (int c = 0; c < nlist; ++c)
        if (offsets[c + 1] < offsets[c]) return fail(VS_ERR_INVALID, "cluster offsets are not monotone");
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) {
        cudaGetLastError();
        return fail(VS_ERR_CUDA, "no CUDA device available (libvsb200 has no CPU fallback)");
    }
    if (device < 0 || device >= cnt) return fail(VS_ERR_INVALID, "bad device ordinal");
    VSB_CUDA(cudaSetDevice(device));
    vs_ivf* h = new (std::nothrow) vs_ivf();
    if (!h) return fail(VS_ERR_NOMEM, "host allocation failed");
    h->device = device;
    h->n = n;
    h->nlist = nlist;
    h->dim = dim;
    h->avg_cluster_size = (float)((double)n / nlist);
    auto body = [&]() -> int {
        VSB_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        // one extra chunk of zero rows behind the last list: TMA boxes never run past the allocation
        const size_t pad_rows = (size_t)ivf_scan_rows_per_chunk();
        VSB_CUDA(cudaMalloc((void**)&h->d_vectors, sizeof(float) * ((size_t)n + pad_rows) * dim));
        VSB_CUDA(cudaMemsetAsync(h->d_vectors + (size_t)n * dim, 0, sizeof(float) * pad_rows * dim, h->stream));
        VSB_CUDA(cudaMalloc((void**)&h->d_offsets, sizeof(int32_t) * ((size_t)nlist + 1)));
        VSB_CUDA(cudaMalloc((void**)&h->d_idmap, sizeof(int32_t) * (size_t)n));
        VSB_CUDA(cudaMalloc((void**)&h->d_centroids, sizeof(float) * (size_t)nlist * dim));
        VSB_CUDA(cudaMalloc((void**)&h->d_order, sizeof(int32_t) * (size_t)nlist));
        {
            std::vector<int32_t> order((size_t)nlist);
            for (int c = 0; c < nlist; ++c) order[(size_t)c] = c;
            std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) {
                return offsets[a + 1] - offsets[a] > offsets[b + 1] - offsets[b];
            });
            VSB_CUDA(cudaMemcpy(h->d_order, order.data(), sizeof(int32_t) * (size_t)nlist, cudaMemcpyHostToDevice));
        }
        VSB_CUDA(cudaMemcpyAsync(h->d_vectors, vectors, sizeof(float) * (size_t)n * dim, cudaMemcpyHostToDevice, h->stream));
        VSB_CUDA(cudaMemcpyAsync(h->d_offsets, offsets, sizeof(int32_t) * ((size_t)nlist + 1), cudaMemcpyHostToDevice, h->stream));
        VSB_CUDA(cudaMemcpyAsync(h->d_idmap, id_map, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
        VSB_CUDA(cudaMemcpyAsync(h->d_centroids, centroids, sizeof(float) * (size_t)nlist * dim, cudaMemcpyHostToDevice, h->stream));
        VSB_TRY(make_tmap_2d(&h->tmV, h->d_vectors, (uint64_t)n + pad_rows, 128, 4, (uint32_t)ivf_scan_rows_per_chunk()));
        VSB_TRY(ivf_set_attributes());
        VSB_TRY(ivf_lm_set_attributes());
        cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device);
        VSB_TRY(h->total.reserve(sizeof(unsigned long long)));
        VSB_CUDA(cudaMallocHost((void**)&h->h_total, sizeof(unsigned long long)));
        VSB_CUDA(cudaMallocHost((void**)&h->h_flag, sizeof(int)));
        // TF32 split of the vectors for the tensor-core scan (rows incl. the zero padding behind the last list)
        {
            const int64_t rows = n + (int64_t)pad_rows;
            VSB_TRY(h->gthr.reserve(sizeof(int)));
            int* flag = h->gthr.as<int>();
            VSB_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), h->stream));
            VSB_TRY(launch_prep_rows(h->d_vectors, n, 128, nullptr, nullptr, nullptr, flag, h->stream));
            VSB_CUDA(cudaMemcpyAsync(h->h_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
            VSB_CUDA(cudaStreamSynchronize(h->stream));
            h->v_exact = *h->h_flag == 0;
            if (h->v_exact) {
                h->d_vhi = h->d_vectors;
            } else {
                VSB_CUDA(cudaMalloc((void**)&h->d_vhi, sizeof(float) * (size_t)rows * 128));
                VSB_CUDA(cudaMalloc((void**)&h->d_vlo, sizeof(float) * (size_t)rows * 128));
                VSB_TRY(launch_prep_rows(h->d_vectors, rows, 128, nullptr, h->d_vhi, h->d_vlo, nullptr, h->stream));
                VSB_TRY(make_tmap_2d(&h->tmVlo, h->d_vlo, (uint64_t)rows, 128, 4, 128));
            }
            VSB_TRY(make_tmap_2d(&h->tmVhi, h->d_vhi, (uint64_t)rows, 128, 4, 128));
            if (!h->d_vlo) h->tmVlo = h->tmVhi;
        }
        VSB_CUDA(cudaStreamSynchronize(h->stream));
        return VS_OK;
    };
    const int rc = body();
    if (rc != VS_OK) {
        const std::string keep = vs_last_error();
        ivf_free(h);
        return fail(rc, keep);
    }
    *out = h;
    return VS_OK;
}

static int ivf_search_core(vs_ivf* h, const float* q_dev, int64_t nq, int k, int nprobe, int32_t* out_ids, float* out_scores,
                           int32_t* out_counts, cudaStream_t st) {
    h->last_launches = 0;
    if (nq == 0) return VS_OK;
    nprobe = std::min(nprobe, h->nlist);  // IVFIndex.cpp:584,647
    VSB_TRY(h->scores.reserve(sizeof(float) * (size_t)nq * h->nlist));
    VSB_TRY(h->probes.reserve(sizeof(int32_t) * (size_t)nq * nprobe));
    VSB_CUDA(cudaMemsetAsync(h->total.p, 0, sizeof(unsigned long long), st));
    VSB_TRY(launch_ivf_coarse(q_dev, nq, h->d_centroids, h->nlist, h->scores.as<float>(), st));
    VSB_TRY(launch_ivf_probes(h->scores.as<float>(), nq, h->nlist, nprobe, h->probes.as<int32_t>(), st));
    // large batches: group the (query, list) pairs by list and scan list-major (K8, FFMA-bound) instead of
    // query-major (K6, bound by streaming each probed list once per query).  VSB_IVF_LM=0/1 forces either path.
    const char* lm_env = getenv("VSB_IVF_LM");
    // measured (tools/ivf_crossover.py, 1M x 128, nlist 1024): K8 wins once a list is probed by ~6 queries on average
    // (nprobe 8: from ~700 queries), and for nprobe >= 16 at any batch (K6 walks a query's lists one after the other)
    const bool lm_auto = (int64_t)nq * nprobe >= 6 * (int64_t)h->nlist || (nprobe >= 16 && nq >= 8);
    const bool list_major = lm_env ? atoi(lm_env) != 0 : (lm_auto && round_up_ktop(k) != 0);
    // list-major on the tensor cores (default) or with FFMA (K8, VSB_IVF_TC=0: scores bit-identical by construction)
    const char* tc_env = getenv("VSB_IVF_TC");
    const bool use_tc = list_major && !(tc_env && atoi(tc_env) == 0) && k <= 30;
    if (h->profile) VSB_CUDA(cudaEventRecord(h->ev0, st));
    if (use_tc) {
        // candidates: fused TF32 tensor-core GEMM + top-k per (query, probe slot) over the list-major work items; the merge
        // re-scores the best k + 2 per query in the reference's NEON order, so the returned scores are bit-identical to
        // IVFIndex.cpp:278-357 and to the FFMA kernels
        const int ktop = round_up_ktop(std::min(k + 2, kMaxRegK));
        const int64_t n_pairs = nq * nprobe;
        const int n_lists = nprobe * 3;
        if (n_lists > 96) return fail(VS_ERR_UNSUPPORTED, "IVF tensor-core scan: nprobe > 32 (use VSB_IVF_TC=0)");
        VSB_TRY(h->lm_ws.reserve(sizeof(int32_t) * ivf_tc_workspace_ints(nq, nprobe, h->nlist)));
        VSB_TRY(h->qhi.reserve(sizeof(float) * (size_t)(n_pairs + 128) * 128));
        VSB_TRY(h->qlo.reserve(sizeof(float) * (size_t)(n_pairs + 128) * 128));
        // candidates: what the (query, probe slot, epilogue group) lists kept, appended per query (<= cap by construction)
        const int cap = n_lists * ktop;
        VSB_TRY(h->part_key.reserve(sizeof(uint2) * (size_t)nq * cap));
        VSB_TRY(h->part_id.reserve(sizeof(int32_t) * (size_t)nq));
        VSB_TRY(h->gthr.reserve(sizeof(int32_t) * (size_t)nq));
        VSB_CUDA(cudaMemsetAsync(h->part_id.p, 0, sizeof(int32_t) * (size_t)nq, st));
        const int4* items;
        const int32_t *n_items, *pairs;
        int32_t* q_cand;
        int* q_flag;
        VSB_TRY(launch_ivf_tc_prep(q_dev, h->d_offsets, h->nlist, h->d_order, h->probes.as<int32_t>(), nq, nprobe, h->lm_ws.as<int32_t>(),
                                   h->qhi.as<float>(), h->qlo.as<float>(), &items, &n_items, &pairs, &q_cand, &q_flag, st));
        bool split3 = true;
        if (h->v_exact) {  // 1xTF32 is bit-identical to 3xTF32 iff the query lo parts are all zero too (one 4-byte round trip)
            VSB_CUDA(cudaMemcpyAsync(h->h_flag, q_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
            VSB_CUDA(cudaStreamSynchronize(st));
            split3 = *h->h_flag != 0;
            if (split3 && !h->d_vlo) {  // a real zero lo operand keeps the arithmetic honest
                const int64_t rows = h->n + (int64_t)ivf_scan_rows_per_chunk();
                VSB_CUDA(cudaMalloc((void**)&h->d_vlo, sizeof(float) * (size_t)rows * 128));
                VSB_CUDA(cudaMemsetAsync(h->d_vlo, 0, sizeof(float) * (size_t)rows * 128, st));
                VSB_TRY(make_tmap_2d(&h->tmVlo, h->d_vlo, (uint64_t)rows, 128, 4, 128));
            }
        }
        VSB_CUDA(cudaMemsetAsync(h->gthr.p, 0x7f, sizeof(int32_t) * (size_t)nq, st));
        CUtensorMap tmQhi, tmQlo;
        VSB_TRY(make_tmap_2d(&tmQhi, h->qhi.p, (uint64_t)(n_pairs + 128), 128, 4, 128));
        VSB_TRY(make_tmap_2d(&tmQlo, h->qlo.p, (uint64_t)(n_pairs + 128), 128, 4, 128));
        VSB_TRY(launch_exact_tc_ivf(tmQhi, tmQlo, h->tmVhi, h->tmVlo, items, n_items, pairs, nprobe, h->gthr.as<int32_t>(), (int)nq, ktop,
                                    split3, h->part_id.as<int32_t>(), h->part_key.p, cap, h->num_sms, st));
        VSB_TRY(launch_ivf_counts(q_cand, nq, k, out_counts, h->total.as<unsigned long long>(), st));
        if (h->profile) {
            VSB_CUDA(cudaEventRecord(h->ev1, st));
            h->ev_valid = true;
        }
        VSB_TRY(launch_filter_merge_ivf(h->part_key.p, h->part_id.as<int32_t>(), cap, nq, k, out_scores, out_ids, h->d_vectors, q_dev,
                                        h->d_idmap, st));
        h->last_launches = 10;
        return VS_OK;
    }
    if (list_major) {
        const int ktop = round_up_ktop(k);
        if (ktop == 0) return fail(VS_ERR_UNSUPPORTED, "IVF search: k > 32 is not implemented");
        VSB_TRY(h->lm_ws.reserve(sizeof(int32_t) * ivf_lm_workspace_ints(nq, nprobe, h->nlist)));
        VSB_TRY(h->part_key.reserve(sizeof(float) * (size_t)nprobe * nq * ktop));
        VSB_TRY(h->part_id.reserve(sizeof(int32_t) * (size_t)nprobe * nq * ktop));
        VSB_TRY(launch_ivf_listmajor(q_dev, h->d_vectors, h->d_offsets, h->d_idmap, h->nlist, h->d_order, h->probes.as<int32_t>(), nq, nprobe,
                                     k, h->lm_ws.as<int32_t>(), h->part_key.as<float>(), h->part_id.as<int32_t>(), out_counts,
                                     h->total.as<unsigned long long>(), h->num_sms, st));
        if (h->profile) {
            VSB_CUDA(cudaEventRecord(h->ev1, st));
            h->ev_valid = true;
        }
        // keys are -score: merge ascending, store +score (padding: -inf / -1)
        VSB_TRY(launch_merge_lists(h->part_key.as<float>(), h->part_id.as<int32_t>(), nprobe, nq, ktop, ktop, k, 0, 0, 1, out_scores,
                                   out_ids, k, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, st));
        h->last_launches = 8;
        return VS_OK;
    }
    VSB_TRY(launch_ivf_scan(h->tmV, q_dev, h->probes.as<int32_t>(), h->d_offsets, h->d_idmap, nq, nprobe, k, out_scores, out_ids,
                            out_counts, h->total.as<unsigned long long>(), st));
    if (h->profile) {
        VSB_CUDA(cudaEventRecord(h->ev1, st));
        h->ev_valid = true;
    }
    h->last_launches = 3;
    return VS_OK;
}

// ---------------------------------------------------------------------------------------------------------
// k-means builder helpers
// ---------------------------------------------------------------------------------------------------------
namespace {

// one CTA per list: mean of the member rows in list order, accumulated in double (deterministic)
__global__ void __launch_bounds__(128) list_mean_kernel(const float* __restrict__ x, const int32_t* __restrict__ offsets,
                                                        const int32_t* __restrict__ members, float* __restrict__ cent,
                                                        double* __restrict__ shift2) {
    const int c = blockIdx.x;
    const int d = threadIdx.x;
    const int s = offsets[c], e = offsets[c + 1];
    if (e <= s) return;  // empty cluster keeps its previous centre
    double acc = 0.0;
    for (int r = s; r < e; ++r) acc += (double)__ldg(x + (size_t)__ldg(members + r) * 128 + d);
    const float nc = (float)(acc / (double)(e - s));
    const float oc = cent[(size_t)c * 128 + d];
    cent[(size_t)c * 128 + d] = nc;
    const double df = (double)nc - (double)oc;
    double v = df * df;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((d & 31) == 0) atomicAdd(shift2, v);
}

void labels_to_lists(const std::vector<int32_t>& labels, int nlist, std::vector<int32_t>& offsets, std::vector<int32_t>& members) {
    const size_t n = labels.size();
    offsets.assign((size_t)nlist + 1, 0);
    for (size_t i = 0; i < n; ++i) offsets[(size_t)labels[i] + 1]++;
    for (int c = 0; c < nlist; ++c) offsets[c + 1] += offsets[c];
    members.resize(n);
    std::vector<int32_t> cur(offsets.begin(), offsets.end() - 1);
    for (size_t i = 0; i < n; ++i) members[(size_t)cur[labels[i]]++] = (int32_t)i;  // ascending ids inside a list
}

}  // namespace

extern "C" {

int vs_ivf_create(vs_ivf_t** out, const float* vectors_list_order, int64_t n, int dim, const int32_t* cluster_offsets,
                  int nlist, const int32_t* position_to_id, const float* centroids, int device) {
    return ivf_create_impl(out, vectors_list_order, n, dim, cluster_offsets, nlist, position_to_id, centroids, device);
}

int vs_ivf_open(vs_ivf_t** out, const char* index_dir, int device) {
    if (!out) return fail(VS_ERR_INVALID, "out handle is NULL");
    *out = nullptr;
    if (!index_dir) return fail(VS_ERR_INVALID, "index_dir is NULL");
    const std::string dir(index_dir);
    std::string json, e;
    if (!(e = vsb_io::read_text(dir + "/ivf_config.json", json)).empty()) return fail(VS_ERR_IO, e);
    vsb_io::IvfConfig cfg;
    if (!(e = vsb_io::parse_ivf_config(json, cfg)).empty()) return fail(VS_ERR_IO, e);
    // the values become int / int32 ids below: refuse what does not fit before any array is sized from them
    if (cfg.n_vectors == 0 || cfg.n_vectors > 0x7fffffffull || cfg.n_clusters == 0 || cfg.n_clusters > 0x7fffffffull ||
        cfg.dim == 0 || cfg.dim > 65536)
        return fail(VS_ERR_IO, "ivf_config.json: n_vectors / n_clusters / dim out of range");
    std::vector<size_t> shape;
    std::vector<int32_t> offsets, idx;
    std::vector<float> vectors, centroids;
    if (!(e = vsb_io::load_npy_i32(dir + "/cluster_offsets.npy", offsets, shape)).empty())
        return fail(VS_ERR_IO, "Failed to load cluster offsets: " + e);
    if (offsets.size() != cfg.n_clusters + 1) return fail(VS_ERR_IO, "cluster_offsets.npy does not have n_clusters+1 entries");
    if (cfg.reordered) {
        if (!(e = vsb_io::load_npy_i32(dir + "/reorder_to_original.npy", idx, shape)).empty())
            return fail(VS_ERR_IO, "Failed to load reorder map: " + e);
        if (!(e = vsb_io::load_npy_f32(dir + "/vectors_reordered.npy", vectors, shape)).empty())
            return fail(VS_ERR_IO, "Cannot open reordered vectors file: " + e);
    } else {
        if (!(e = vsb_io::load_npy_i32(dir + "/cluster_indices.npy", idx, shape)).empty())
            return fail(VS_ERR_IO, "Failed to load cluster indices: " + e);
        std::vector<float> orig;
        if (vsb_io::file_exists(dir + "/vectors.npy")) {
            if (!(e = vsb_io::load_npy_f32(dir + "/vectors.npy", orig, shape)).empty()) return fail(VS_ERR_IO, e);
        } else {
            std::string raw;
            if (!(e = vsb_io::read_text(dir + "/vectors.bin", raw)).empty())
                return fail(VS_ERR_IO, "Cannot open vectors file: " + dir + "/vectors.npy or vectors.bin");
            orig.resize(raw.size() / 4);
            std::memcpy(orig.data(), raw.data(), orig.size() * 4);
        }
        if (orig.size() != cfg.n_vectors * cfg.dim) return fail(VS_ERR_IO, "vectors file does not match n_vectors x dim");
        if (idx.size() != cfg.n_vectors) return fail(VS_ERR_IO, "cluster_indices.npy does not have n_vectors entries");
        // scattered layout -> list-contiguous rows (what the device scan streams)
        vectors.resize(orig.size());
        for (size_t r = 0; r < idx.size(); ++r) {
            if (idx[r] < 0 || (size_t)idx[r] >= cfg.n_vectors) return fail(VS_ERR_IO, "cluster index out of range");
            std::memcpy(&vectors[r * cfg.dim], &orig[(size_t)idx[r] * cfg.dim], cfg.dim * 4);
        }
    }
    if (vectors.size() != cfg.n_vectors * cfg.dim || idx.size() != cfg.n_vectors)
        return fail(VS_ERR_IO, "index arrays do not match n_vectors x dim");
    // The reference feeds the coarse stage from centroids.bin, a QNN context binary of the MatMul graph; the
    // same numbers live in centroids.npy, written by both builders.
    if (!(e = vsb_io::load_npy_f32(dir + "/centroids.npy", centroids, shape)).empty())
        return fail(VS_ERR_IO, "Failed to load centroids: " + e);
    if (centroids.size() != cfg.n_clusters * cfg.dim) return fail(VS_ERR_IO, "centroids.npy does not match n_clusters x dim");
    const int rc = ivf_create_impl(out, vectors.data(), (int64_t)cfg.n_vectors, (int)cfg.dim, offsets.data(), (int)cfg.n_clusters,
                                   idx.data(), centroids.data(), device);
    if (rc == VS_OK && cfg.avg_cluster_size > 0.f) (*out)->avg_cluster_size = cfg.avg_cluster_size;
    return rc;
}

int vs_ivf_destroy(vs_ivf_t* h) { return ivf_free(h); }
int64_t vs_ivf_num_vectors(const vs_ivf_t* h) { return h ? h->n : 0; }
int vs_ivf_num_clusters(const vs_ivf_t* h) { return h ? h->nlist : 0; }
int vs_ivf_dim(const vs_ivf_t* h) { return h ? h->dim : 0; }
float vs_ivf_avg_cluster_size(const vs_ivf_t* h) { return h ? h->avg_cluster_size : 0.f; }

int vs_ivf_search_dev(vs_ivf_t* h, const float* queries_dev, int64_t nq, int k, int nprobe, int32_t* out_ids_dev,
                      float* out_scores_dev, int32_t* out_counts_dev, void* stream) {
    if (!h) return fail(VS_ERR_INVALID, "handle is NULL");
    if (nq < 0 || k <= 0 || nprobe <= 0) return fail(VS_ERR_INVALID, "nq < 0, k <= 0 or nprobe <= 0");
    if (k > kMaxRegK) return fail(VS_ERR_UNSUPPORTED, "IVF search: k > 32 is not implemented");
    if (nq > 0 && (!queries_dev || !out_ids_dev || !out_scores_dev || !out_counts_dev)) return fail(VS_ERR_INVALID, "NULL buffer");
    VSB_CUDA(cudaSetDevice(h->device));
    return ivf_search_core(h, queries_dev, nq, k, nprobe, out_ids_dev, out_scores_dev, out_counts_dev,
                           stream ? (cudaStream_t)stream : h->stream);
}

int vs_ivf_search(vs_ivf_t* h, const float* queries, int64_t nq, int k, int nprobe, int32_t* out_ids, float* out_scores,
                  int32_t* out_counts, uint64_t* total_candidates) {
    if (!h) return fail(VS_ERR_INVALID, "handle is NULL");
    if (nq < 0 || k <= 0 || nprobe <= 0) return fail(VS_ERR_INVALID, "nq < 0, k <= 0 or nprobe <= 0");
    if (k > kMaxRegK) return fail(VS_ERR_UNSUPPORTED, "IVF search: k > 32 is not implemented");
    if (total_candidates) *total_candidates = 0;
    if (nq == 0) return VS_OK;
    if (!queries || !out_ids || !out_scores) return fail(VS_ERR_INVALID, "NULL buffer");
    VSB_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    VSB_TRY(h->q.reserve(sizeof(float) * (size_t)nq * 128));
    VSB_TRY(h->out_ids.reserve(sizeof(int32_t) * (size_t)nq * k));
    VSB_TRY(h->out_scores.reserve(sizeof(float) * (size_t)nq * k));
    VSB_TRY(h->out_counts.reserve(sizeof(int32_t) * (size_t)nq));
    VSB_CUDA(cudaMemcpyAsync(h->q.p, queries, sizeof(float) * (size_t)nq * 128, cudaMemcpyHostToDevice, st));
    VSB_TRY(ivf_search_core(h, h->q.as<float>(), nq, k, nprobe, h->out_ids.as<int32_t>(), h->out_scores.as<float>(),
                            h->out_counts.as<int32_t>(), st));
    VSB_CUDA(cudaMemcpyAsync(out_ids, h->out_ids.p, sizeof(int32_t) * (size_t)nq * k, cudaMemcpyDeviceToHost, st));
    VSB_CUDA(cudaMemcpyAsync(out_scores, h->out_scores.p, sizeof(float) * (size_t)nq * k, cudaMemcpyDeviceToHost, st));
    if (out_counts) VSB_CUDA(cudaMemcpyAsync(out_counts, h->out_counts.p, sizeof(int32_t) * (size_t)nq, cudaMemcpyDeviceToHost, st));
    VSB_CUDA(cudaMemcpyAsync(h->h_total, h->total.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    VSB_CUDA(cudaStreamSynchronize(st));
    if (total_candidates) *total_candidates = (uint64_t)*h->h_total;
    return VS_OK;
}

int vs_ivf_coarse_scores(vs_ivf_t* h, const float* queries, int64_t nq, float* out_scores) {
    if (!h) return fail(VS_ERR_INVALID, "handle is NULL");
    if (nq < 0) return fail(VS_ERR_INVALID, "nq < 0");
    if (nq == 0) return VS_OK;
    if (!queries || !out_scores) return fail(VS_ERR_INVALID, "NULL buffer");
    VSB_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    VSB_TRY(h->q.reserve(sizeof(float) * (size_t)nq * 128));
    VSB_TRY(h->scores.reserve(sizeof(float) * (size_t)nq * h->nlist));
    VSB_CUDA(cudaMemcpyAsync(h->q.p, queries, sizeof(float) * (size_t)nq * 128, cudaMemcpyHostToDevice, st));
    VSB_TRY(launch_ivf_coarse(h->q.as<float>(), nq, h->d_centroids, h->nlist, h->scores.as<float>(), st));
    VSB_CUDA(cudaMemcpyAsync(out_scores, h->scores.p, sizeof(float) * (size_t)nq * h->nlist, cudaMemcpyDeviceToHost, st));
    VSB_CUDA(cudaStreamSynchronize(st));
    return VS_OK;
}

int vs_ivf_set_profile(vs_ivf_t* h, int enable) {
    if (!h) return fail(VS_ERR_INVALID, "handle is NULL");
    VSB_CUDA(cudaSetDevice(h->device));
    if (enable && !h->ev0) {
        VSB_CUDA(cudaEventCreate(&h->ev0));
        VSB_CUDA(cudaEventCreate(&h->ev1));
    }
    h->profile = enable != 0;
    h->ev_valid = false;
    return VS_OK;
}

int vs_ivf_last_kernel_ms(vs_ivf_t* h, float* ms) {
    if (!h || !ms) return fail(VS_ERR_INVALID, "NULL argument");
    if (!h->ev_valid) return fail(VS_ERR_INVALID, "no profiled search yet (vs_ivf_set_profile)");
    VSB_CUDA(cudaEventSynchronize(h->ev1));
    VSB_CUDA(cudaEventElapsedTime(ms, h->ev0, h->ev1));
    return VS_OK;
}

// k-means (Lloyd, L2) on the GPU + inverted lists + the reference's index directory.
int vs_ivf_build(const float* base, int64_t n, int dim, int nlist, int max_iter, uint64_t seed, const char* out_dir,
                 int reordered, int device, const float* init_centroids, int* out_nlist, int* out_iters,
                 double* out_inertia) {
    if (!base || n <= 0 || nlist <= 0 || max_iter < 0 || !out_dir) return fail(VS_ERR_INVALID, "bad arguments");
    if (dim != 128) return fail(VS_ERR_UNSUPPORTED, "only dim == 128 (SIFT shape) is implemented");
    if (n > 0x7fffffffLL) return fail(VS_ERR_INVALID, "ids must fit int32");
    if (nlist > n / 10) nlist = (int)std::max<int64_t>(16, n / 100);  // create_ivf_model.py:97-99
    if (nlist > n) nlist = (int)n;
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) {
        cudaGetLastError();
        return fail(VS_ERR_CUDA, "no CUDA device available (libvsb200 has no CPU fallback)");
    }
    if (device < 0 || device >= cnt) return fail(VS_ERR_INVALID, "bad device ordinal");
    VSB_CUDA(cudaSetDevice(device));

    std::vector<float> cent((size_t)nlist * dim);
    float* d_base = nullptr;
    float* d_cent = nullptr;
    int32_t *d_lab = nullptr, *d_prev = nullptr, *d_off = nullptr, *d_mem = nullptr, *d_ws = nullptr, *d_changed = nullptr;
    float* d_dist = nullptr;
    double *d_shift = nullptr, *d_part = nullptr;
    vs_exact_t* cidx = nullptr;
    std::vector<int32_t> labels((size_t)n), offsets((size_t)nlist + 1), members((size_t)n);
    int iters = 0;
    double inertia = 0.0;
    auto cleanup = [&]() {
        if (cidx) vs_exact_destroy(cidx);
        for (void* p : {(void*)d_base, (void*)d_cent, (void*)d_lab, (void*)d_prev, (void*)d_off, (void*)d_mem, (void*)d_ws,
                        (void*)d_changed, (void*)d_dist, (void*)d_shift, (void*)d_part})
            if (p) cudaFree(p);
    };
    auto body = [&]() -> int {
        VSB_CUDA(cudaMalloc((void**)&d_base, sizeof(float) * (size_t)n * dim));
        VSB_CUDA(cudaMalloc((void**)&d_cent, sizeof(float) * (size_t)nlist * dim));
        VSB_CUDA(cudaMalloc((void**)&d_lab, sizeof(int32_t) * (size_t)n));
        VSB_CUDA(cudaMalloc((void**)&d_prev, sizeof(int32_t) * (size_t)n));
        VSB_CUDA(cudaMalloc((void**)&d_dist, sizeof(float) * (size_t)n));
        VSB_CUDA(cudaMalloc((void**)&d_off, sizeof(int32_t) * ((size_t)nlist + 1)));
        VSB_CUDA(cudaMalloc((void**)&d_mem, sizeof(int32_t) * (size_t)n));
        VSB_CUDA(cudaMalloc((void**)&d_ws, sizeof(int32_t) * multisplit_workspace_ints(n, nlist)));
        VSB_CUDA(cudaMalloc((void**)&d_changed, sizeof(int32_t)));
        VSB_CUDA(cudaMalloc((void**)&d_shift, sizeof(double) * 2));
        VSB_CUDA(cudaMalloc((void**)&d_part, sizeof(double) * (size_t)ceil_div64(n, 1024)));
        VSB_CUDA(cudaMemcpy(d_base, base, sizeof(float) * (size_t)n * dim, cudaMemcpyHostToDevice));
        if (init_centroids) {
            VSB_CUDA(cudaMemcpy(d_cent, init_centroids, sizeof(float) * (size_t)nlist * dim, cudaMemcpyHostToDevice));
        } else {
            // sklearn's default init: greedy k-means++ (2 + ln k candidates per centre drawn ~ D^2), on the device
            VSB_TRY(launch_kmeanspp(d_base, n, nlist, seed, d_cent, nullptr));
        }
        VSB_CUDA(cudaMemset(d_prev, 0xff, sizeof(int32_t) * (size_t)n));  // -1: "no label yet"
        // data variance for sklearn's relative tolerance (tol = 1e-4 * mean per-feature variance)
        double var = 0.0;
        {
            std::vector<double> s1((size_t)dim, 0.0), s2((size_t)dim, 0.0);
            for (int64_t i = 0; i < n; ++i)
                for (int d = 0; d < dim; ++d) {
                    const double v = base[(size_t)i * dim + d];
                    s1[(size_t)d] += v;
                    s2[(size_t)d] += v * v;
                }
            for (int d = 0; d < dim; ++d) var += s2[(size_t)d] / n - (s1[(size_t)d] / n) * (s1[(size_t)d] / n);
            var /= dim;
        }
        const double tol = 1e-4 * var;
        for (int it = 0;; ++it) {
            // ---- assignment: fused tensor-core distance + arg-min over the centroids (exact path, k = 1)
            if (!cidx)
                VSB_TRY(vs_exact_create_dev(&cidx, d_cent, nlist, dim, device, 0));
            else
                VSB_TRY(vs_exact_refresh(cidx));  // centroids moved: norms / fp16 copy / TF32 split recomputed in place
            VSB_TRY(vs_exact_search_dev(cidx, d_base, n, 1, VS_PREC_FP32_3XTF32, d_lab, d_dist, nullptr));
            // labels changed? (counted on the device: 4 bytes come back, not the labels)
            VSB_CUDA(cudaMemsetAsync(d_changed, 0, sizeof(int32_t), nullptr));
            VSB_CUDA(cudaDeviceSynchronize());  // the search ran on the handle's stream
            VSB_TRY(launch_labels_changed(d_lab, d_prev, n, d_changed, nullptr));
            int32_t changed = 0;
            VSB_CUDA(cudaMemcpy(&changed, d_changed, sizeof(int32_t), cudaMemcpyDeviceToHost));
            if (it >= max_iter || changed == 0) {
                iters = it;
                break;
            }
            // ---- update: inverted lists (stable counting sort on the device), per-list means
            VSB_TRY(launch_multisplit(d_lab, n, nlist, d_ws, d_off, d_mem, nullptr));
            VSB_CUDA(cudaMemsetAsync(d_shift, 0, sizeof(double), nullptr));
            list_mean_kernel<<<nlist, 128>>>(d_base, d_off, d_mem, d_cent, d_shift);
            VSB_CUDA(cudaGetLastError());
            double shift2 = 0.0;
            VSB_CUDA(cudaMemcpy(&shift2, d_shift, sizeof(double), cudaMemcpyDeviceToHost));
            if (shift2 <= tol) max_iter = std::min(max_iter, it + 1);  // converged: one final assignment, then stop
        }
        // final state: labels, lists, inertia (sum of the exact fp32 distances to the assigned centroid, fixed order)
        VSB_TRY(launch_multisplit(d_lab, n, nlist, d_ws, d_off, d_mem, nullptr));
        VSB_TRY(launch_inertia(d_dist, n, d_part, d_shift + 1, nullptr));
        VSB_CUDA(cudaMemcpy(labels.data(), d_lab, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost));
        VSB_CUDA(cudaMemcpy(offsets.data(), d_off, sizeof(int32_t) * ((size_t)nlist + 1), cudaMemcpyDeviceToHost));
        VSB_CUDA(cudaMemcpy(members.data(), d_mem, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost));
        VSB_CUDA(cudaMemcpy(&inertia, d_shift + 1, sizeof(double), cudaMemcpyDeviceToHost));
        VSB_CUDA(cudaMemcpy(cent.data(), d_cent, sizeof(float) * (size_t)nlist * dim, cudaMemcpyDeviceToHost));
        return VS_OK;
    };
    int rc = body();
    const std::string keep = rc == VS_OK ? "" : vs_last_error();
    cleanup();
    if (rc != VS_OK) return fail(rc, keep);

    // ---- files (create_ivf_model.py:114-166, create_ivf_model_reordered.py:108-169)
    int32_t mn = 0x7fffffff, mx = 0;
    std::vector<int32_t> sizes((size_t)nlist);
    for (int c = 0; c < nlist; ++c) {
        sizes[(size_t)c] = offsets[(size_t)c + 1] - offsets[(size_t)c];
        mn = std::min(mn, sizes[(size_t)c]);
        mx = std::max(mx, sizes[(size_t)c]);
    }
    const std::string dir(out_dir);
    ::mkdir(dir.c_str(), 0777);
    std::string e;
    char js[512];
    std::snprintf(js, sizeof js,
                  "{\n  \"n_vectors\": %lld,\n  \"n_clusters\": %d,\n  \"dim\": %d,\n  \"batch_size\": 1,\n"
                  "  \"avg_cluster_size\": %.6f,\n  \"min_cluster_size\": %d,\n  \"max_cluster_size\": %d%s\n}",
                  (long long)n, nlist, dim, (double)n / nlist, mn, mx, reordered ? ",\n  \"reordered\": true" : "");
    {
        FILE* f = std::fopen((dir + "/ivf_config.json").c_str(), "w");
        if (!f) return fail(VS_ERR_IO, "cannot create " + dir + "/ivf_config.json");
        std::fputs(js, f);
        std::fclose(f);
    }
    const std::vector<size_t> sh_off{(size_t)nlist + 1}, sh_n{(size_t)n}, sh_c{(size_t)nlist, (size_t)dim}, sh_v{(size_t)n, (size_t)dim},
        sh_l{(size_t)nlist};
    if (!(e = vsb_io::save_npy(dir + "/cluster_offsets.npy", offsets.data(), "<i4", sh_off, 4)).empty()) return fail(VS_ERR_IO, e);
    if (!(e = vsb_io::save_npy(dir + "/centroids.npy", cent.data(), "<f4", sh_c, 4)).empty()) return fail(VS_ERR_IO, e);
    if (!(e = vsb_io::save_npy(dir + "/cluster_ids.npy", labels.data(), "<i4", sh_n, 4)).empty()) return fail(VS_ERR_IO, e);
    if (reordered) {
        std::vector<float> rv((size_t)n * dim);
        for (int64_t r = 0; r < n; ++r) std::memcpy(&rv[(size_t)r * dim], base + (size_t)members[(size_t)r] * dim, (size_t)dim * 4);
        if (!(e = vsb_io::save_npy(dir + "/vectors_reordered.npy", rv.data(), "<f4", sh_v, 4)).empty()) return fail(VS_ERR_IO, e);
        if (!(e = vsb_io::save_npy(dir + "/reorder_to_original.npy", members.data(), "<i4", sh_n, 4)).empty()) return fail(VS_ERR_IO, e);
        if (!(e = vsb_io::save_npy(dir + "/cluster_sizes.npy", sizes.data(), "<i4", sh_l, 4)).empty()) return fail(VS_ERR_IO, e);
    } else {
        if (!(e = vsb_io::save_npy(dir + "/cluster_indices.npy", members.data(), "<i4", sh_n, 4)).empty()) return fail(VS_ERR_IO, e);
        if (!(e = vsb_io::save_npy(dir + "/vectors.npy", base, "<f4", sh_v, 4)).empty()) return fail(VS_ERR_IO, e);
    }
    if (out_nlist) *out_nlist = nlist;
    if (out_iters) *out_iters = iters;
    if (out_inertia) *out_inertia = inertia;
    return VS_OK;
}

}  // extern "C"
