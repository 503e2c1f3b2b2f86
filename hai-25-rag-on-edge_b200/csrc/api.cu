// C ABI of libvsb200 (include/vsb200.h): handle management, workspace, TMA descriptors, search orchestration.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "exact_handle.cuh"
#include "kernels.cuh"
#include "vsb_common.cuh"

namespace vsb {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point query: libvsb200.so then has no link-time
// dependency on libcuda and still loads (for symbol checks) on machines without a driver.
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static encode_tiled_fn get_encode_fn() {
    static encode_tiled_fn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (encode_tiled_fn)p;
    });
    return fn;
}

// 2D row-major [rows x cols] matrix of `elem_bytes` elements, box = 128 B x box_rows, 128-B swizzle, OOB -> 0
int make_tmap_fold(CUtensorMap* out, const void* gptr, uint64_t rows, uint32_t box_rows) {
    encode_tiled_fn fn = get_encode_fn();
    if (!fn) return fail(VS_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t gdim[2] = {(cuuint64_t)TC_FOLD_COLS, rows};
    cuuint64_t gstr[1] = {(cuuint64_t)TC_FOLD_COLS * 2};
    cuuint32_t box[2] = {(cuuint32_t)TC_FOLD_COLS, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(gptr), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(VS_ERR_CUDA, "cuTensorMapEncodeTiled (norm block) failed with code " + std::to_string((int)r));
    return VS_OK;
}

int make_tmap_2d(CUtensorMap* out, const void* gptr, uint64_t rows, uint64_t cols, int elem_bytes, uint32_t box_rows) {
    encode_tiled_fn fn = get_encode_fn();
    if (!fn) return fail(VS_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    const CUtensorMapDataType dt = elem_bytes == 4   ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                   : elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                                     : CU_TENSOR_MAP_DATA_TYPE_UINT8;
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstr[1] = {cols * (uint64_t)elem_bytes};
    cuuint32_t box[2] = {(cuuint32_t)(128 / elem_bytes), box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, dt, 2, const_cast<void*>(gptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(VS_ERR_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
    return VS_OK;
}

}  // namespace vsb

using namespace vsb;

int exact_free(vs_exact* h) {
    if (!h) return VS_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->owns_base && h->d_base) cudaFree((void*)h->d_base);
    if (h->d_hi && h->d_hi != h->d_base) cudaFree(h->d_hi);
    if (h->d_lo) cudaFree(h->d_lo);
    if (h->d_f16) cudaFree(h->d_f16);
    if (h->d_fold) cudaFree(h->d_fold);
    if (h->d_norm) cudaFree(h->d_norm);
    for (DevBuf* b : {&h->q, &h->qhi, &h->qlo, &h->qf16, &h->qnorm, &h->part_key, &h->part_id, &h->lbk, &h->lbi, &h->out_ids,
                      &h->out_keys, &h->flag, &h->gthr, &h->qparams, &h->qfold, &h->unc_list, &h->fb_q, &h->fb_ids, &h->fb_keys,
                      &h->f_smin, &h->f_thr, &h->f_cnt, &h->f_cand})
        b->release();
    for (auto& g : h->graphs) cudaGraphExecDestroy(g.exec);
    for (DevBuf* b : {&h->g_q, &h->g_qnorm, &h->g_part_key, &h->g_part_id, &h->g_ids, &h->g_keys}) b->release();
    if (h->hp_q) cudaFreeHost(h->hp_q);
    if (h->hp_ids) cudaFreeHost(h->hp_ids);
    if (h->hp_keys) cudaFreeHost(h->hp_keys);
    if (h->h_flag) cudaFreeHost(h->h_flag);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->ev_pre) cudaEventDestroy(h->ev_pre);
    if (h->ev_cert) cudaEventDestroy(h->ev_cert);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return VS_OK;
}

// flag buffer layout (device, 4 words): [0] "not TF32-exact" flag, [1] uncertified-query count, [2] abs-max bits
// (queries), [3] scratch for max reductions at build time
static int exact_build(vs_exact* h) {
    const int64_t n = h->n;
    const int dim = h->dim;
    const int64_t n_pad = ceil_div64(n, 128) * 128;
    // (vs_exact_refresh rebuilds into the buffers of the first build: same shape)
    if (!h->d_norm) VSB_CUDA(cudaMalloc((void**)&h->d_norm, sizeof(float) * (size_t)n_pad));
    VSB_TRY(launch_fill_f32(h->d_norm, n_pad, __builtin_inff(), h->stream));
    VSB_TRY(h->flag.reserve(4 * sizeof(int)));
    if (!h->h_flag) VSB_CUDA(cudaMallocHost((void**)&h->h_flag, 4 * sizeof(int)));
    VSB_CUDA(cudaMemsetAsync(h->flag.p, 0, 4 * sizeof(int), h->stream));
    // norms in the reference's order + "is every component TF32-representable"
    VSB_TRY(launch_prep_rows(h->d_base, n, dim, h->d_norm, nullptr, nullptr, h->flag.as<int>(), h->stream));
    if (dim == 128) {
        // scaled fp16 copy for the certified candidate pass: abs-max -> power-of-two scale; max norm for the bound
        float* scratch = reinterpret_cast<float*>(h->flag.as<int>() + 3);
        VSB_TRY(launch_absmax_f32(h->d_base, n * dim, scratch, h->stream));
        VSB_CUDA(cudaMemcpyAsync(h->h_flag, h->flag.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        VSB_CUDA(cudaStreamSynchronize(h->stream));
        h->base_exact = (h->h_flag[0] == 0);
        float absmax;
        memcpy(&absmax, &h->h_flag[3], sizeof(float));
        h->s_b = f16_scale_host(absmax);
        VSB_CUDA(cudaMemsetAsync(scratch, 0, sizeof(float), h->stream));
        VSB_TRY(launch_absmax_f32(h->d_norm, n, scratch, h->stream));
        VSB_CUDA(cudaMemcpyAsync(&h->h_flag[3], scratch, sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        if (!h->d_f16) VSB_CUDA(cudaMalloc(&h->d_f16, 2 * (size_t)n * dim));
        VSB_TRY(launch_to_half_scaled(h->d_base, n * dim, h->s_b, nullptr, h->d_f16, h->stream));
        // the norm term as an extra K = 16 operand block: three fp16 pieces of s_b^2 ||x||^2 / 2 per row
        if (!h->d_fold) VSB_CUDA(cudaMalloc(&h->d_fold, (size_t)n_pad * TC_FOLD_COLS * 2));
        VSB_TRY(launch_norm_pieces(h->d_norm, n, n_pad, h->s_b, h->d_fold, h->stream));
        VSB_CUDA(cudaStreamSynchronize(h->stream));
        memcpy(&h->bn_max, &h->h_flag[3], sizeof(float));
        VSB_TRY(make_tmap_2d(&h->tmB16.hi, h->d_f16, (uint64_t)n, 128, 2, 128));
        VSB_TRY(make_tmap_fold(&h->tmB16.lo, h->d_fold, (uint64_t)n_pad, 128));
        h->tmB16.hi_half = h->tmB16.hi;  // CTA pairs are not used by the fp16 pass
        h->tmB16.lo_half = h->tmB16.lo;
        VSB_TRY(tc_set_attributes());
    } else {
        VSB_CUDA(cudaStreamSynchronize(h->stream));
    }
    return VS_OK;
}

// TF32 hi/lo split of the base (2x the base bytes), built the first time a TF32 search needs it
static int exact_ensure_split(vs_exact* h, bool need_lo, cudaStream_t st) {
    if (!h->split_ready) {
        if (h->base_exact && !h->d_lo) {  // hi == x and lo == 0: no copy
            h->d_hi = const_cast<float*>(h->d_base);
        } else {
            if (!h->d_hi || h->d_hi == h->d_base) VSB_CUDA(cudaMalloc((void**)&h->d_hi, sizeof(float) * (size_t)h->n * 128));
            if (!h->d_lo) VSB_CUDA(cudaMalloc((void**)&h->d_lo, sizeof(float) * (size_t)h->n * 128));
            VSB_TRY(launch_prep_rows(h->d_base, h->n, 128, nullptr, h->d_hi, h->d_lo, nullptr, st));
            VSB_TRY(make_tmap_2d(&h->tmB32.lo, h->d_lo, (uint64_t)h->n, 128, 4, 128));
            VSB_TRY(make_tmap_2d(&h->tmB32.lo_half, h->d_lo, (uint64_t)h->n, 128, 4, 64));
        }
        VSB_TRY(make_tmap_2d(&h->tmB32.hi, h->d_hi, (uint64_t)h->n, 128, 4, 128));
        VSB_TRY(make_tmap_2d(&h->tmB32.hi_half, h->d_hi, (uint64_t)h->n, 128, 4, 64));
        if (!h->d_lo) {
            h->tmB32.lo = h->tmB32.hi;
            h->tmB32.lo_half = h->tmB32.hi_half;
        }
        h->split_ready = true;
    }
    if (need_lo && !h->d_lo) {  // 3x search over a TF32-exact base: a real zero lo operand keeps the arithmetic honest
        VSB_CUDA(cudaMalloc((void**)&h->d_lo, sizeof(float) * (size_t)h->n * 128));
        VSB_CUDA(cudaMemsetAsync(h->d_lo, 0, sizeof(float) * (size_t)h->n * 128, st));
        VSB_TRY(make_tmap_2d(&h->tmB32.lo, h->d_lo, (uint64_t)h->n, 128, 4, 128));
        VSB_TRY(make_tmap_2d(&h->tmB32.lo_half, h->d_lo, (uint64_t)h->n, 128, 4, 64));
    }
    return VS_OK;
}

int exact_create_common(vs_exact_t** out, const float* base, bool on_device, int64_t n, int dim, int device,
                               int64_t id_base) {
    if (!out) return fail(VS_ERR_INVALID, "out handle is NULL");
    *out = nullptr;
    if (!base || n <= 0) return fail(VS_ERR_INVALID, "base is NULL or n <= 0");
    if (dim != 128) return fail(VS_ERR_UNSUPPORTED, "only dim == 128 (SIFT shape) is implemented");
    if (id_base < 0 || id_base + n > 0x7fffffffLL) return fail(VS_ERR_INVALID, "ids must fit int32");
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) {
        cudaGetLastError();
        return fail(VS_ERR_CUDA, "no CUDA device available (libvsb200 has no CPU fallback)");
    }
    if (device < 0 || device >= cnt) return fail(VS_ERR_INVALID, "bad device ordinal");
    VSB_CUDA(cudaSetDevice(device));
    vs_exact* h = new (std::nothrow) vs_exact();
    if (!h) return fail(VS_ERR_NOMEM, "host allocation failed");
    h->device = device;
    h->n = n;
    h->dim = dim;
    h->id_base = id_base;
    cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device);
    int rc = VS_OK;
    do {
        if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
            rc = fail(VS_ERR_CUDA, "cudaStreamCreate failed");
            break;
        }
        if (on_device) {
            h->d_base = base;
            h->owns_base = false;
        } else {
            float* d = nullptr;
            cudaError_t e = cudaMalloc((void**)&d, sizeof(float) * (size_t)n * dim);
            if (e != cudaSuccess) {
                rc = fail(VS_ERR_NOMEM, std::string("cudaMalloc(base): ") + cudaGetErrorString(e));
                break;
            }
            h->d_base = d;
            h->owns_base = true;
            e = cudaMemcpyAsync(d, base, sizeof(float) * (size_t)n * dim, cudaMemcpyHostToDevice, h->stream);
            if (e != cudaSuccess) {
                rc = fail(VS_ERR_CUDA, std::string("H2D(base): ") + cudaGetErrorString(e));
                break;
            }
        }
        rc = exact_build(h);
    } while (0);
    if (rc != VS_OK) {
        std::string keep = g_err;
        exact_free(h);
        g_err = keep;
        return rc;
    }
    *out = h;
    return VS_OK;
}

// The fp16 candidate pass is a THRESHOLD FILTER: every row whose fp16 key lies below a per-query threshold thr[q] becomes a
// candidate (unordered (key, id) pairs appended to the query's array); nothing is kept sorted inside the tensor-core kernel.
// Any threshold is legal, because the merge certifies each query against the bound it actually used ("no row outside the
// kept candidates has a key below B"); a threshold that turns out too tight only sends the query to the fp32 fallback.  thr[q]
// comes from a SAMPLE pass over every stride-th base tile: the m-th smallest of the query's group minima has at least m
// sampled rows at or below it, i.e. about m * stride (~ 128) rows of the whole base — enough margin for k <= 16, few enough
// to keep the candidate traffic negligible.
constexpr int kF16CandCap = 512;      // candidates kept per query (more: the query is redone on the fp32 path)
constexpr int kF16SampleOff = 3;      // first sampled tile
constexpr int kF16MinSampleTiles = 24;

// Sample geometry: one base tile in 16 and the 8th smallest group minimum (~ 8 x 16 rows below the threshold) for bases of
// >= 400 tiles (~ 51 K rows); one tile in 4 and the 16th smallest group minimum below that (few groups: several of the
// smallest sampled keys share a group, which only loosens the threshold).
struct F16Sample {
    int stride, m, tiles;
};
static F16Sample f16_sample_plan(int64_t n) {
    const int64_t n_tiles = ceil_div64(n, 128);
    F16Sample s;
    s.stride = n_tiles >= 400 ? 16 : 4;
    s.m = n_tiles >= 400 ? 8 : 16;
    if (const char* e = getenv("VSB_F16_STRIDE")) s.stride = std::max(1, atoi(e));
    if (const char* e = getenv("VSB_F16_M")) s.m = std::max(1, atoi(e));
    s.tiles = n_tiles > kF16SampleOff ? (int)((n_tiles - kF16SampleOff + s.stride - 1) / s.stride) : 0;
    return s;
}
// bases of <= kF16CandCap rows need no threshold at all; otherwise the sample must be large enough to mean something
static bool f16_pass_supported(int64_t n) { return n <= kF16CandCap || f16_sample_plan(n).tiles >= kF16MinSampleTiles; }

// Query prep (norms, abs-max -> power-of-two scale, fp16 copy, bound constants), sample pass, thresholds, filter pass.
// Leaves the candidates in h->f_cand / h->f_cnt and the thresholds in h->f_thr.
static int exact_f16_candidate_pass(vs_exact* h, const float* q_dev, int64_t nq, cudaStream_t st) {
    int* flag = h->flag.as<int>();
    VSB_TRY(h->qnorm.reserve(sizeof(float) * (size_t)nq));
    VSB_TRY(h->qf16.reserve(2 * (size_t)nq * 128));
    VSB_TRY(h->qparams.reserve(sizeof(TcQueryParams)));
    VSB_TRY(h->qfold.reserve((size_t)128 * TC_FOLD_COLS * 2));
    VSB_TRY(h->f_thr.reserve(sizeof(float) * (size_t)nq));
    VSB_TRY(h->f_cnt.reserve(sizeof(int32_t) * (size_t)nq));
    VSB_TRY(h->f_cand.reserve(sizeof(uint2) * (size_t)nq * kF16CandCap));
    VSB_CUDA(cudaMemsetAsync(flag + 1, 0, 2 * sizeof(int), st));
    VSB_TRY(launch_query_prep(q_dev, nq, h->qnorm.as<float>(), reinterpret_cast<float*>(flag + 2), st));
    TcQueryParams* qp = h->qparams.as<TcQueryParams>();
    VSB_TRY(launch_tc_query_params(reinterpret_cast<const float*>(flag + 2), h->s_b, h->bn_max, qp, h->qfold.p, st));
    VSB_TRY(launch_to_half_scaled(q_dev, nq * 128, 1.f, qp, h->qf16.p, st));  // the NEGATED scaled copy
    CUtensorMap tmA, tmAe;
    VSB_TRY(make_tmap_2d(&tmA, h->qf16.p, (uint64_t)nq, 128, 2, 128));
    VSB_TRY(make_tmap_fold(&tmAe, h->qfold.p, 128, 128));
    h->ev_pre_valid = false;
    if (h->profile && h->n > kF16CandCap) {
        VSB_CUDA(cudaEventRecord(h->ev_pre, st));
        h->ev_pre_valid = true;
    }
    if (h->n <= kF16CandCap) {
        VSB_TRY(launch_tc_fill_thr((int)nq, h->f_thr.as<float>(), h->f_cnt.as<int32_t>(), st));
    } else {
        const F16Sample sp = f16_sample_plan(h->n);
        const int stride = sp.stride, m = sp.m, g_tiles = sp.tiles;
        if (g_tiles < 1) return fail(VS_ERR_UNSUPPORTED, "fp16 candidate pass: base too small for the sample pass");
        const TcPlan splan = tc_make_plan((int64_t)g_tiles * 128, nq, h->num_sms, 2);
        const int n_groups = splan.n_splits * tc_sample_groups_per_split();
        VSB_TRY(h->f_smin.reserve(sizeof(float) * (size_t)n_groups * nq));
        VSB_TRY(launch_exact_tc_f16(tmA, tmAe, h->tmB16, (int)nq, h->n, splan, true, stride, kF16SampleOff, h->f_smin.as<float>(), nullptr,
                                    nullptr, nullptr, 0, st));
        VSB_TRY(launch_tc_select_thr(h->f_smin.as<float>(), n_groups, (int)nq, m, h->f_thr.as<float>(), h->f_cnt.as<int32_t>(), st));
    }
    // the filter pass holds two query tiles per unit (every base tile feeds both): its plan counts tile PAIRS as columns
    const TcPlan plan = tc_make_plan(h->n, ceil_div64(ceil_div64(nq, 128), 2) * 128, h->num_sms, 2);
    if (h->profile) VSB_CUDA(cudaEventRecord(h->ev0, st));  // the dominant kernel alone: the filter pass
    VSB_TRY(launch_exact_tc_f16(tmA, tmAe, h->tmB16, (int)nq, h->n, plan, false, 1, 0, nullptr, h->f_thr.as<float>(),
                                h->f_cnt.as<int32_t>(), h->f_cand.p, kF16CandCap, st));
    if (h->profile) {
        VSB_CUDA(cudaEventRecord(h->ev1, st));
        h->ev_valid = true;
    }
    return VS_OK;
}

// Certified candidate pass: the scaled fp16 tensor-core kernel collects every row below the query's threshold (key error
// bounded by cert_a*sqrt(qn)+cert_b), the merge kernel keeps the <= 32 best of them, recomputes their distances in exact
// fp32, ranks them and certifies the top k; the (rare) uncertified queries are redone on the 3xTF32 / FFMA path.  The
// 4-byte count of uncertified queries is waited for in exact_certified_finish.
static int exact_search_certified(vs_exact* h, const float* q_dev, int64_t nq, int k, int32_t* out_ids, float* out_dists,
                                  cudaStream_t st, int32_t* unc_dev) {
    int* flag = h->flag.as<int>();
    VSB_TRY(h->unc_list.reserve(sizeof(int32_t) * (size_t)nq));
    VSB_TRY(exact_f16_candidate_pass(h, q_dev, nq, st));
    TcQueryParams* qp = h->qparams.as<TcQueryParams>();
    // count of uncertified queries: the handle's own word (zeroed by the candidate pass), or the caller's (trailer of an
    // exchange block, already zeroed)
    int32_t* unc = unc_dev ? unc_dev : flag + 1;
    VSB_TRY(launch_filter_merge(h->f_cand.p, h->f_cnt.as<int32_t>(), kF16CandCap, h->f_thr.as<float>(), nq, k, h->id_base, out_dists,
                                out_ids, k, h->d_base, h->d_norm, q_dev, h->qnorm.as<float>(), qp, unc, h->unc_list.as<int32_t>(), st));
    h->last_launches = h->n <= kF16CandCap ? 6 : 7;
    h->last_precision = VS_PREC_F16_CERTIFIED;
    h->last_fallback = 0;
    VSB_CUDA(cudaMemcpyAsync(h->h_flag + 1, unc, sizeof(int), cudaMemcpyDeviceToHost, st));
    if (!h->ev_cert) VSB_CUDA(cudaEventCreateWithFlags(&h->ev_cert, cudaEventDisableTiming));
    VSB_CUDA(cudaEventRecord(h->ev_cert, st));
    h->cert_pending = true;
    h->pend_q = q_dev;
    h->pend_nq = nq;
    h->pend_k = k;
    h->pend_ids = out_ids;
    h->pend_dists = out_dists;
    h->pend_st = st;
    return VS_OK;
}

// Second half of the certified search: waits for the 4-byte count of uncertified queries (an event recorded right
// after its copy: work enqueued later on the stream is not waited for) and redoes those queries on the fp32 path.
int exact_certified_finish(vs_exact* h, int* n_redone) {
    if (n_redone) *n_redone = 0;
    if (!h->cert_pending) return VS_OK;
    h->cert_pending = false;
    const float* q_dev = h->pend_q;
    const int k = h->pend_k;
    int32_t* out_ids = h->pend_ids;
    float* out_dists = h->pend_dists;
    cudaStream_t st = h->pend_st;
    VSB_CUDA(cudaEventSynchronize(h->ev_cert));
    const int n_unc = h->h_flag[1];
    if (n_redone) *n_redone = n_unc;
    h->last_fallback = n_unc;
    if (n_unc > 0) {
        VSB_TRY(h->fb_q.reserve(sizeof(float) * (size_t)n_unc * 128));
        VSB_TRY(h->fb_ids.reserve(sizeof(int32_t) * (size_t)n_unc * k));
        VSB_TRY(h->fb_keys.reserve(sizeof(float) * (size_t)n_unc * k));
        VSB_TRY(launch_gather_rows(q_dev, h->unc_list.as<int32_t>(), n_unc, h->fb_q.as<float>(), st));
        const int launches = h->last_launches;
        const bool prof = h->profile;
        h->profile = false;  // keep the timing of the dominant (candidate) kernel
        const int rc = exact_search_core(h, h->fb_q.as<float>(), n_unc, k, n_unc <= 8 ? VS_PREC_FP32_FFMA : VS_PREC_FP32_3XTF32,
                                         h->fb_ids.as<int32_t>(), h->fb_keys.as<float>(), st);
        h->profile = prof;
        VSB_TRY(rc);
        VSB_TRY(launch_scatter_results(h->fb_keys.as<float>(), h->fb_ids.as<int32_t>(), h->unc_list.as<int32_t>(), n_unc, k,
                                       out_dists, out_ids, st));
        h->last_launches += launches + 2;
        h->last_precision = VS_PREC_F16_CERTIFIED;
        h->last_fallback = n_unc;
    }
    return VS_OK;
}

// One group of <= 32 results per query: [pass]
int exact_search_core(vs_exact* h, const float* q_dev, int64_t nq, int k, int precision, int32_t* out_ids,
                      float* out_dists, cudaStream_t st, bool defer_certification, int32_t* unc_dev) {
    if (h->cert_pending) return fail(VS_ERR_INVALID, "vs_exact_search_dev_finish() of the previous search was not called");
    if (h->broken) return fail(VS_ERR_INVALID, "the handle is unusable: vs_exact_refresh() failed");
    h->last_launches = 0;
    h->last_fallback = 0;
    if (nq == 0) return VS_OK;
    const int dim = h->dim;
    int prec = precision;
    if (prec != VS_PREC_AUTO && prec != VS_PREC_FP32_3XTF32 && prec != VS_PREC_FP32_FFMA && prec != VS_PREC_TF32_1X &&
        prec != VS_PREC_F16_CERTIFIED)
        return fail(VS_ERR_INVALID, "unknown precision");
    if (nq > 0x7fffffff / 128) return fail(VS_ERR_INVALID, "nq too large");
    if (prec == VS_PREC_F16_CERTIFIED && k > 16) return fail(VS_ERR_UNSUPPORTED, "certified fp16 candidate pass needs k <= 16");
    // AUTO (measured on B200, top-10, whole call, tools/small_batch.py -> profiles/r2_small_batch_sweep.txt): <= 8 queries: one
    // FFMA pass over the fp32 base (0.17-0.28 ms at 1M rows; host calls replay a CUDA graph); from 9 queries on the certified
    // fp16 path wins at every batch size (device-pointer calls) — it streams the half-size fp16 base: 0.16 ms up to 256 queries, 0.21 ms at 512,
    // 0.95 ms at 4096, against 0.27 / 0.54 / 3.5 ms for 3xTF32 (125K rows: 0.10-0.26 ms against 0.13-0.52)
    // Host calls: FFMA graph 0.167 / 0.175 / 0.212 / 0.295 ms at batch 1 / 2 / 4 / 8, fp16 path 0.173-0.175 ms at all of them:
    // the fp16 path takes over from 3 queries (where it applies); otherwise FFMA up to 8 queries, TF32 beyond
    constexpr int64_t kAutoFfmaMax = 8, kAutoF16Min = 3;
    if (prec == VS_PREC_AUTO && nq >= kAutoF16Min && k <= 16 && dim == 128 && f16_pass_supported(h->n)) prec = VS_PREC_F16_CERTIFIED;
    // bases too small for a meaningful sample pass (513 .. ~13 K rows) take the 3xTF32 path: same answer, and just as fast there
    if (prec == VS_PREC_F16_CERTIFIED && dim == 128 && !f16_pass_supported(h->n)) prec = VS_PREC_FP32_3XTF32;
    const bool want_tc = (prec == VS_PREC_FP32_3XTF32 || prec == VS_PREC_TF32_1X || prec == VS_PREC_F16_CERTIFIED ||
                          (prec == VS_PREC_AUTO && nq > kAutoFfmaMax));
    if (want_tc && dim != 128) return fail(VS_ERR_UNSUPPORTED, "tensor-core path needs dim == 128");

    VSB_TRY(h->qnorm.reserve(sizeof(float) * (size_t)nq));
    if (prec == VS_PREC_F16_CERTIFIED) {
        VSB_TRY(exact_search_certified(h, q_dev, nq, k, out_ids, out_dists, st, unc_dev));
        return defer_certification ? VS_OK : exact_certified_finish(h, nullptr);
    }
    const int passes = (k + kMaxRegK - 1) / kMaxRegK;
    // single pass: keep a couple of spare candidates beyond k so that the exact refine can repair a k-th/k+1-th
    // swap caused by the tensor-core rounding bias
    const int ktop = passes == 1 ? round_up_ktop(std::min(k + 2, kMaxRegK)) : kMaxRegK;
    if (passes > 1) {
        VSB_TRY(h->lbk.reserve(sizeof(float) * (size_t)nq));
        VSB_TRY(h->lbi.reserve(sizeof(int32_t) * (size_t)nq));
    }

    if (want_tc) {
        VSB_TRY(h->qhi.reserve(sizeof(float) * (size_t)nq * 128));
        VSB_TRY(h->qlo.reserve(sizeof(float) * (size_t)nq * 128));
        VSB_CUDA(cudaMemsetAsync(h->flag.p, 0, sizeof(int), st));
        VSB_TRY(launch_prep_rows(q_dev, nq, 128, h->qnorm.as<float>(), h->qhi.as<float>(), h->qlo.as<float>(),
                                 h->flag.as<int>(), st));
        h->last_launches++;
        bool split3 = true;
        if (prec == VS_PREC_TF32_1X) {
            split3 = false;
        } else if (prec == VS_PREC_AUTO) {
            if (h->base_exact) {  // 1xTF32 is bit-identical to 3xTF32 iff the query lo parts are all zero too
                VSB_CUDA(cudaMemcpyAsync(h->h_flag, h->flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
                VSB_CUDA(cudaStreamSynchronize(st));
                split3 = (h->h_flag[0] != 0);
            }
        }
        VSB_TRY(exact_ensure_split(h, split3, st));
        h->last_precision = split3 ? VS_PREC_FP32_3XTF32 : VS_PREC_TF32_1X;
        const TcPlan plan = tc_make_plan(h->n, nq, h->num_sms, split3 ? 1 : 0);
        const int n_lists = plan.n_splits * tc_lists_per_split(split3 ? 1 : 0);
        VSB_TRY(h->part_key.reserve(sizeof(float) * (size_t)n_lists * nq * ktop));
        VSB_TRY(h->part_id.reserve(sizeof(int32_t) * (size_t)n_lists * nq * ktop));
        VSB_TRY(h->gthr.reserve(sizeof(int32_t) * (size_t)nq));
        CUtensorMap tmA_hi, tmA_lo;
        VSB_TRY(make_tmap_2d(&tmA_hi, h->qhi.p, (uint64_t)nq, 128, 4, 128));
        VSB_TRY(make_tmap_2d(&tmA_lo, h->qlo.p, (uint64_t)nq, 128, 4, 128));
        for (int pass = 0; pass < passes; ++pass) {
            const int kk = std::min(kMaxRegK, k - pass * kMaxRegK);
            const bool lb = pass > 0;
            // shared per-query thresholds start at a huge finite value (0x7f7f7f7f in the ordered-int encoding)
            VSB_CUDA(cudaMemsetAsync(h->gthr.p, 0x7f, sizeof(int32_t) * (size_t)nq, st));
            if (h->profile && pass == 0) VSB_CUDA(cudaEventRecord(h->ev0, st));
            VSB_TRY(launch_exact_tc(tmA_hi, tmA_lo, h->tmB32, h->d_norm, h->gthr.as<int32_t>(), (int)nq, h->n, plan,
                                    ktop, split3 ? 1 : 0, nullptr, lb ? h->lbk.as<float>() : nullptr,
                                    lb ? h->lbi.as<int32_t>() : nullptr, h->part_key.as<float>(), h->part_id.as<int32_t>(), st));
            if (h->profile && pass == 0) {
                VSB_CUDA(cudaEventRecord(h->ev1, st));
                h->ev_valid = true;
                h->ev_pre_valid = false;
            }
            VSB_TRY(launch_merge_lists(h->part_key.as<float>(), h->part_id.as<int32_t>(), n_lists, nq, ktop,
                                       passes == 1 ? ktop : kk, passes == 1 ? k : kk, h->id_base, 0, 0, out_dists,
                                       out_ids, k, pass * kMaxRegK, passes > 1 ? h->lbk.as<float>() : nullptr,
                                       passes > 1 ? h->lbi.as<int32_t>() : nullptr, h->d_base, h->d_norm, q_dev,
                                       h->qnorm.as<float>(), st));
            h->last_launches += 2;
        }
        if (passes > 1) {
            VSB_TRY(launch_sort_rows(out_dists, out_ids, nq, k, st));
            h->last_launches++;
        }
        return VS_OK;
    }

    // ---- FFMA streaming path: groups of <= 8 queries, one pass over the base per group (and per 32 results)
    h->last_precision = VS_PREC_FP32_FFMA;
    VSB_TRY(launch_prep_rows(q_dev, nq, dim, h->qnorm.as<float>(), nullptr, nullptr, nullptr, st));
    h->last_launches++;
    const int max_ctas = stream_num_ctas(h->device, 1);
    VSB_TRY(h->part_key.reserve(sizeof(float) * (size_t)max_ctas * 8 * ktop));
    VSB_TRY(h->part_id.reserve(sizeof(int32_t) * (size_t)max_ctas * 8 * ktop));
    for (int64_t q0 = 0; q0 < nq; q0 += 8) {
        const int g = (int)std::min<int64_t>(8, nq - q0);
        const int n_ctas = stream_num_ctas(h->device, g);
        for (int pass = 0; pass < passes; ++pass) {
            const int kk = std::min(kMaxRegK, k - pass * kMaxRegK);
            const bool lb = pass > 0;
            if (h->profile && pass == 0 && q0 == 0) VSB_CUDA(cudaEventRecord(h->ev0, st));
            VSB_TRY(launch_exact_stream(h->d_base, h->d_norm, h->n, q_dev + q0 * dim, h->qnorm.as<float>() + q0, g, ktop,
                                        lb ? h->lbk.as<float>() + q0 : nullptr, lb ? h->lbi.as<int32_t>() + q0 : nullptr,
                                        h->part_key.as<float>(), h->part_id.as<int32_t>(), n_ctas, st));
            if (h->profile && pass == 0 && q0 == 0) {
                VSB_CUDA(cudaEventRecord(h->ev1, st));
                h->ev_valid = true;
                h->ev_pre_valid = false;
            }
            VSB_TRY(launch_merge_lists(h->part_key.as<float>(), h->part_id.as<int32_t>(), n_ctas, g, ktop,
                                       passes == 1 ? ktop : kk, passes == 1 ? k : kk, h->id_base, 0, 0,
                                       out_dists + q0 * k, out_ids + q0 * k, k, pass * kMaxRegK,
                                       passes > 1 ? h->lbk.as<float>() + q0 : nullptr,
                                       passes > 1 ? h->lbi.as<int32_t>() + q0 : nullptr, h->d_base, h->d_norm,
                                       q_dev + q0 * dim, h->qnorm.as<float>() + q0, st));
            h->last_launches += 2;
        }
    }
    if (passes > 1) {
        VSB_TRY(launch_sort_rows(out_dists, out_ids, nq, k, st));
        h->last_launches++;
    }
    return VS_OK;
}

// Batch <= 8, k <= 32 through the host-buffer entry point: latency is launch-bound (three kernels, two copies and a
// synchronisation around a ~0.1 ms scan), so the whole sequence is captured once per (nq, k) and replayed as one graph.
// The graph works on its own (fixed-size) workspaces and pinned staging buffers, so larger searches in between cannot move
// the buffers it has baked in.
static int exact_search_small_graph(vs_exact* h, const float* queries, int64_t nq, int k, int32_t* out_ids, float* out_dists) {
    cudaStream_t st = h->stream;
    if (!h->hp_q) {
        VSB_CUDA(cudaMallocHost((void**)&h->hp_q, sizeof(float) * 8 * 128));
        VSB_CUDA(cudaMallocHost((void**)&h->hp_ids, sizeof(int32_t) * 8 * kMaxRegK));
        VSB_CUDA(cudaMallocHost((void**)&h->hp_keys, sizeof(float) * 8 * kMaxRegK));
        const int max_ctas = stream_num_ctas(h->device, 1);
        VSB_TRY(h->g_q.reserve(sizeof(float) * 8 * 128));
        VSB_TRY(h->g_qnorm.reserve(sizeof(float) * 8));
        VSB_TRY(h->g_part_key.reserve(sizeof(float) * (size_t)max_ctas * 8 * kMaxRegK));
        VSB_TRY(h->g_part_id.reserve(sizeof(int32_t) * (size_t)max_ctas * 8 * kMaxRegK));
        VSB_TRY(h->g_ids.reserve(sizeof(int32_t) * 8 * kMaxRegK));
        VSB_TRY(h->g_keys.reserve(sizeof(float) * 8 * kMaxRegK));
    }
    cudaGraphExec_t exec = nullptr;
    for (auto& g : h->graphs)
        if (g.nq == nq && g.k == k) exec = g.exec;
    if (!exec) {
        // capture: the same calls as the ordinary path, on the graph's own workspaces
        std::swap(h->qnorm, h->g_qnorm);
        std::swap(h->part_key, h->g_part_key);
        std::swap(h->part_id, h->g_part_id);
        cudaGraph_t graph = nullptr;
        int rc = VS_OK;
        cudaError_t e = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
        if (e != cudaSuccess) rc = fail(VS_ERR_CUDA, std::string("cudaStreamBeginCapture: ") + cudaGetErrorString(e));
        if (rc == VS_OK) {
            const size_t qb = sizeof(float) * (size_t)nq * 128, rb = (size_t)nq * k * 4;
            e = cudaMemcpyAsync(h->g_q.p, h->hp_q, qb, cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess)
                rc = exact_search_core(h, h->g_q.as<float>(), nq, k, VS_PREC_FP32_FFMA, h->g_ids.as<int32_t>(), h->g_keys.as<float>(), st);
            if (e == cudaSuccess && rc == VS_OK) e = cudaMemcpyAsync(h->hp_ids, h->g_ids.p, rb, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess && rc == VS_OK) e = cudaMemcpyAsync(h->hp_keys, h->g_keys.p, rb, cudaMemcpyDeviceToHost, st);
            const cudaError_t e2 = cudaStreamEndCapture(st, &graph);  // always ends the capture
            if (rc == VS_OK && e != cudaSuccess) rc = fail(VS_ERR_CUDA, std::string("graph capture: ") + cudaGetErrorString(e));
            if (rc == VS_OK && e2 != cudaSuccess) rc = fail(VS_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e2));
        }
        std::swap(h->qnorm, h->g_qnorm);
        std::swap(h->part_key, h->g_part_key);
        std::swap(h->part_id, h->g_part_id);
        if (rc == VS_OK) {
            e = cudaGraphInstantiate(&exec, graph, 0);
            if (e != cudaSuccess) rc = fail(VS_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
        }
        if (graph) cudaGraphDestroy(graph);
        if (rc != VS_OK) {
            cudaGetLastError();
            return rc;
        }
        h->graphs.push_back({nq, k, exec});
    }
    memcpy(h->hp_q, queries, sizeof(float) * (size_t)nq * 128);
    VSB_CUDA(cudaGraphLaunch(exec, st));
    VSB_CUDA(cudaStreamSynchronize(st));
    memcpy(out_ids, h->hp_ids, (size_t)nq * k * 4);
    memcpy(out_dists, h->hp_keys, (size_t)nq * k * 4);
    h->last_launches = 3;
    h->last_precision = VS_PREC_FP32_FFMA;
    h->last_fallback = 0;
    return VS_OK;
}

extern "C" {

const char* vs_last_error(void) { return g_err.c_str(); }
int vs_abi_version(void) { return VSB200_ABI_VERSION; }

int vs_device_count(int* count) {
    if (!count) return fail(VS_ERR_INVALID, "count is NULL");
    *count = 0;
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) {
        cudaGetLastError();
        return fail(VS_ERR_CUDA, "no CUDA device available");
    }
    *count = c;
    return c > 0 ? VS_OK : fail(VS_ERR_CUDA, "no CUDA device available");
}

int vs_exact_create(vs_exact_t** out, const float* base, int64_t n, int dim, int device, int64_t id_base) {
    return exact_create_common(out, base, false, n, dim, device, id_base);
}
int vs_exact_create_dev(vs_exact_t** out, const float* base_dev, int64_t n, int dim, int device, int64_t id_base) {
    return exact_create_common(out, base_dev, true, n, dim, device, id_base);
}
int vs_exact_destroy(vs_exact_t* h) { return exact_free(h); }

int vs_exact_refresh(vs_exact_t* h) {
    if (!h) return fail(VS_ERR_INVALID, "handle is NULL");
    VSB_CUDA(cudaSetDevice(h->device));
    VSB_CUDA(cudaStreamSynchronize(h->stream));
    // same pointer, same shape: norms, fp16 copy, norm block and (on the next TF32 search) the hi/lo split are recomputed
    // IN PLACE, nothing is freed or reallocated (the k-means builder calls this once per iteration)
    if (h->d_hi == h->d_base) h->d_hi = nullptr;
    h->split_ready = false;
    const int rc = exact_build(h);
    h->broken = rc != VS_OK;  // a failed rebuild leaves no usable buffers: every later search is refused
    return rc;
}
int64_t vs_exact_size(const vs_exact_t* h) { return h ? h->n : 0; }
int vs_exact_dim(const vs_exact_t* h) { return h ? h->dim : 0; }
int vs_exact_base_is_tf32_exact(const vs_exact_t* h) { return h && h->base_exact ? 1 : 0; }

int vs_exact_search_dev(vs_exact_t* h, const float* queries_dev, int64_t nq, int k, int precision, int32_t* out_ids_dev,
                        float* out_dists_dev, void* stream) {
    if (!h) return fail(VS_ERR_INVALID, "handle is NULL");
    if (nq < 0 || k <= 0) return fail(VS_ERR_INVALID, "nq < 0 or k <= 0");
    if ((int64_t)k > h->n) return fail(VS_ERR_INVALID, "k > n (undefined in the reference, cpu_baseline.cpp:129-131)");
    if (nq > 0 && (!queries_dev || !out_ids_dev || !out_dists_dev)) return fail(VS_ERR_INVALID, "NULL buffer");
    VSB_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    return exact_search_core(h, queries_dev, nq, k, precision, out_ids_dev, out_dists_dev, st);
}

int vs_exact_search_dev_begin(vs_exact_t* h, const float* queries_dev, int64_t nq, int k, int precision,
                              int32_t* out_ids_dev, float* out_dists_dev, void* stream) {
    if (!h) return fail(VS_ERR_INVALID, "handle is NULL");
    if (nq < 0 || k <= 0) return fail(VS_ERR_INVALID, "nq < 0 or k <= 0");
    if ((int64_t)k > h->n) return fail(VS_ERR_INVALID, "k > n (undefined in the reference, cpu_baseline.cpp:129-131)");
    if (nq > 0 && (!queries_dev || !out_ids_dev || !out_dists_dev)) return fail(VS_ERR_INVALID, "NULL buffer");
    VSB_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    return exact_search_core(h, queries_dev, nq, k, precision, out_ids_dev, out_dists_dev, st, true);
}

int vs_exact_search_dev_finish(vs_exact_t* h, int* n_redone) {
    if (!h) return fail(VS_ERR_INVALID, "handle is NULL");
    VSB_CUDA(cudaSetDevice(h->device));
    return exact_certified_finish(h, n_redone);
}

int vs_exact_search_f32(vs_exact_t* h, const float* queries, int64_t nq, int k, int precision, int32_t* out_ids,
                        float* out_dists) {
    if (!h) return fail(VS_ERR_INVALID, "handle is NULL");
    if (nq < 0 || k <= 0) return fail(VS_ERR_INVALID, "nq < 0 or k <= 0");
    if ((int64_t)k > h->n) return fail(VS_ERR_INVALID, "k > n (undefined in the reference, cpu_baseline.cpp:129-131)");
    if (nq == 0) return VS_OK;
    if (!queries || !out_ids || !out_dists) return fail(VS_ERR_INVALID, "NULL buffer");
    VSB_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    const bool auto_f16 = precision == VS_PREC_AUTO && nq >= 3 && k <= 16 && f16_pass_supported(h->n);  // exact_search_core's rule
    if (nq <= 8 && k <= kMaxRegK && h->dim == 128 && ((precision == VS_PREC_AUTO && !auto_f16) || precision == VS_PREC_FP32_FFMA) &&
        !h->profile && !h->cert_pending && !h->broken && !getenv("VSB_NO_GRAPH"))
        return exact_search_small_graph(h, queries, nq, k, out_ids, out_dists);
    VSB_TRY(h->q.reserve(sizeof(float) * (size_t)nq * h->dim));
    VSB_TRY(h->out_ids.reserve(sizeof(int32_t) * (size_t)nq * k));
    VSB_TRY(h->out_keys.reserve(sizeof(float) * (size_t)nq * k));
    VSB_CUDA(cudaMemcpyAsync(h->q.p, queries, sizeof(float) * (size_t)nq * h->dim, cudaMemcpyHostToDevice, st));
    // the download is enqueued BEFORE the host waits for the certification count (no idle gap behind the merge kernel); the rare
    // redo of uncertified queries rewrites rows, so the copies are then repeated
    VSB_TRY(exact_search_core(h, h->q.as<float>(), nq, k, precision, h->out_ids.as<int32_t>(), h->out_keys.as<float>(), st, true));
    int redone = 0;
    for (int pass = 0; pass < 2; ++pass) {
        VSB_CUDA(cudaMemcpyAsync(out_ids, h->out_ids.p, sizeof(int32_t) * (size_t)nq * k, cudaMemcpyDeviceToHost, st));
        VSB_CUDA(cudaMemcpyAsync(out_dists, h->out_keys.p, sizeof(float) * (size_t)nq * k, cudaMemcpyDeviceToHost, st));
        if (pass == 0) VSB_TRY(exact_certified_finish(h, &redone));
        if (redone == 0) break;
    }
    VSB_CUDA(cudaStreamSynchronize(st));
    return VS_OK;
}

int vs_exact_debug_f16_candidates(vs_exact_t* h, const float* queries, int64_t nq, int32_t* out_ids, float* out_keys,
                                  float* out_bound) {
    if (!h || !queries || !out_ids || !out_keys || !out_bound) return fail(VS_ERR_INVALID, "NULL argument");
    if (nq <= 0 || nq > 0x7fffffff / 128) return fail(VS_ERR_INVALID, "bad nq");
    if (h->dim != 128) return fail(VS_ERR_UNSUPPORTED, "tensor-core path needs dim == 128");
    if (h->cert_pending) return fail(VS_ERR_INVALID, "a search is in flight");
    VSB_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    const int ktop = kMaxRegK;
    VSB_TRY(h->q.reserve(sizeof(float) * (size_t)nq * 128));
    VSB_TRY(h->out_ids.reserve(sizeof(int32_t) * (size_t)nq * ktop));
    VSB_TRY(h->out_keys.reserve(sizeof(float) * (size_t)nq * ktop));
    VSB_CUDA(cudaMemcpyAsync(h->q.p, queries, sizeof(float) * (size_t)nq * 128, cudaMemcpyHostToDevice, st));
    if (!f16_pass_supported(h->n)) return fail(VS_ERR_UNSUPPORTED, "the fp16 candidate pass needs <= 512 or >= ~13 K base rows");
    VSB_TRY(exact_f16_candidate_pass(h, h->q.as<float>(), nq, st));
    // the (up to) 32 best candidate keys per query exactly as the tensor-core pass ranked them: no refine, no certification
    VSB_TRY(launch_filter_merge(h->f_cand.p, h->f_cnt.as<int32_t>(), kF16CandCap, h->f_thr.as<float>(), nq, ktop, 0,
                                h->out_keys.as<float>(), h->out_ids.as<int32_t>(), ktop, nullptr, nullptr, nullptr, nullptr, nullptr,
                                nullptr, nullptr, st));
    std::vector<float> qn((size_t)nq);
    TcQueryParams qp;
    VSB_CUDA(cudaMemcpyAsync(out_ids, h->out_ids.p, sizeof(int32_t) * (size_t)nq * ktop, cudaMemcpyDeviceToHost, st));
    VSB_CUDA(cudaMemcpyAsync(out_keys, h->out_keys.p, sizeof(float) * (size_t)nq * ktop, cudaMemcpyDeviceToHost, st));
    VSB_CUDA(cudaMemcpyAsync(qn.data(), h->qnorm.p, sizeof(float) * (size_t)nq, cudaMemcpyDeviceToHost, st));
    VSB_CUDA(cudaMemcpyAsync(&qp, h->qparams.p, sizeof(qp), cudaMemcpyDeviceToHost, st));
    VSB_CUDA(cudaStreamSynchronize(st));
    if (!qp.fold_ok) return fail(VS_ERR_UNSUPPORTED, "query / base magnitudes too far apart for the fp16 candidate pass");
    for (int64_t i = 0; i < nq * ktop; ++i) out_keys[i] *= qp.key_unscale;  // accumulator units -> distance units (exact)
    for (int64_t i = 0; i < nq; ++i) out_bound[i] = qp.cert_a * sqrtf(qn[(size_t)i]) + qp.cert_b;
    return VS_OK;
}

int vs_exact_last_launches(const vs_exact_t* h, int* n_kernels, int* precision_used) {
    if (!h) return fail(VS_ERR_INVALID, "handle is NULL");
    if (n_kernels) *n_kernels = h->last_launches;
    if (precision_used) *precision_used = h->last_precision;
    return VS_OK;
}

int vs_exact_last_fallbacks(const vs_exact_t* h, int* n_queries) {
    if (!h || !n_queries) return fail(VS_ERR_INVALID, "NULL argument");
    *n_queries = h->last_fallback;
    return VS_OK;
}

int vs_exact_set_profile(vs_exact_t* h, int enable) {
    if (!h) return fail(VS_ERR_INVALID, "handle is NULL");
    VSB_CUDA(cudaSetDevice(h->device));
    if (enable && !h->ev0) {
        VSB_CUDA(cudaEventCreate(&h->ev0));
        VSB_CUDA(cudaEventCreate(&h->ev1));
        VSB_CUDA(cudaEventCreate(&h->ev_pre));
    }
    h->profile = enable != 0;
    h->ev_valid = false;
    return VS_OK;
}

int vs_exact_last_kernel_ms(vs_exact_t* h, float* ms) {
    if (!h || !ms) return fail(VS_ERR_INVALID, "NULL argument");
    if (!h->ev_valid) return fail(VS_ERR_INVALID, "no profiled search yet (vs_exact_set_profile)");
    VSB_CUDA(cudaEventSynchronize(h->ev1));
    VSB_CUDA(cudaEventElapsedTime(ms, h->ev0, h->ev1));
    return VS_OK;
}

int vs_exact_last_prepass_ms(vs_exact_t* h, float* ms) {
    if (!h || !ms) return fail(VS_ERR_INVALID, "NULL argument");
    *ms = 0.f;
    if (!h->ev_valid) return fail(VS_ERR_INVALID, "no profiled search yet (vs_exact_set_profile)");
    if (!h->ev_pre_valid) return VS_OK;
    VSB_CUDA(cudaEventSynchronize(h->ev0));
    VSB_CUDA(cudaEventElapsedTime(ms, h->ev_pre, h->ev0));
    return VS_OK;
}

int vs_merge_topk_dev(const int32_t* ids_dev, const float* keys_dev, int n_shards, int64_t nq, int k, int smallest,
                      int32_t* out_ids_dev, float* out_keys_dev, void* stream) {
    if (!ids_dev || !keys_dev || !out_ids_dev || !out_keys_dev) return fail(VS_ERR_INVALID, "NULL buffer");
    if (n_shards <= 0 || nq < 0 || k <= 0) return fail(VS_ERR_INVALID, "bad sizes");
    if (k > kMaxRegK)
        return launch_merge_shards(keys_dev, ids_dev, n_shards, nq, k, smallest ? 0 : 1, out_keys_dev, out_ids_dev,
                                   (cudaStream_t)stream);
    return launch_merge_lists(keys_dev, ids_dev, n_shards, nq, k, k, k, 0, smallest ? 0 : 1, smallest ? 0 : 1,
                              out_keys_dev, out_ids_dev, k, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                              (cudaStream_t)stream);
}

int vs_synth_fill_dev(float* out_dev, int64_t row0, int64_t nrows, int dim, int law, uint64_t seed, uint64_t centre_seed,
                      void* stream) {
    if (!out_dev || nrows < 0 || dim <= 0) return fail(VS_ERR_INVALID, "bad arguments");
    return launch_synth(out_dev, row0, nrows, dim, law, seed, centre_seed, (cudaStream_t)stream);
}

int vs_host_alloc(void** out, size_t bytes) {
    if (!out) return fail(VS_ERR_INVALID, "out is NULL");
    *out = nullptr;
    cudaError_t e = cudaMallocHost(out, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(VS_ERR_CUDA, std::string("cudaMallocHost: ") + cudaGetErrorString(e));
    }
    return VS_OK;
}
int vs_host_free(void* p) {
    if (p) cudaFreeHost(p);
    return VS_OK;
}

}  // extern "C"
