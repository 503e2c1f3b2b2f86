// C ABI of the INT8 brute-force path (include/vsb200.h, "INT8" section).  Replaces the QnnRunner + find_top_k_int8
// pair of the reference's qidk_rag_demo (qidk_bruteforce/android/app/main/jni/QnnRunner.h:20-52, QnnRunner.cpp:529-638,
// main.cpp:36-71): the base is quantised once to u8 at create time (the ONNX initializer of create_model.py:57-87,
// quantised by the QNN converter), queries are quantised per call with quantize_buffer_neon's rule, and one fused
// kernel does MatMul + requantisation + largest-k.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <new>
#include <string>

#include <cstdlib>

#include "kernels.cuh"
#include "vsb_common.cuh"

using namespace vsb;

struct vs_int8 {
    int device = 0;
    int num_sms = 148;
    int64_t n = 0;
    int dim = 0;
    int64_t id_base = 0;
    float s_in = 0.f, s_w = 0.f, s_out = 0.f;
    float inv_in = 0.f;  // 1.0f / s_in, computed in fp32 like QnnRunner.cpp:544
    float m = 0.f;       // requantisation multiplier fl(fl(s_in*s_w)/s_out)
    uint8_t* d_base = nullptr;  // [n_pad x 128] u8, zero rows beyond n
    CUtensorMap tmB;
    cudaStream_t stream = nullptr;
    DevBuf q, q_u8, part_key, part_id, gthr, out_ids, out_keys, out_u8, raw;
    bool profile = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool ev_valid = false;
    int last_launches = 0;
};

static int int8_free(vs_int8* h) {
    if (!h) return VS_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->d_base) cudaFree(h->d_base);
    for (DevBuf* b : {&h->q, &h->q_u8, &h->part_key, &h->part_id, &h->gthr, &h->out_ids, &h->out_keys, &h->out_u8, &h->raw})
        b->release();
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return VS_OK;
}

static float mul_rn(float a, float b) {
    volatile float r = a * b;  // keep the product rounded to fp32 (no contraction)
    return r;
}

static int int8_create_common(vs_int8_t** out, const float* base, bool on_device, int64_t n, int dim, float in_scale,
                              float w_scale, float out_scale, int device, int64_t id_base) {
    if (!out) return fail(VS_ERR_INVALID, "out handle is NULL");
    *out = nullptr;
    if (!base || n <= 0) return fail(VS_ERR_INVALID, "base is NULL or n <= 0");
    if (dim != 128) return fail(VS_ERR_UNSUPPORTED, "only dim == 128 (SIFT shape) is implemented");
    if (!(in_scale > 0.f) || !(out_scale > 0.f) || !std::isfinite(in_scale) || !std::isfinite(out_scale))
        return fail(VS_ERR_INVALID, "in_scale and out_scale must be positive and finite");
    if (std::isnan(w_scale) || std::isinf(w_scale)) return fail(VS_ERR_INVALID, "w_scale must be finite (<= 0 selects max(base)/255)");
    if (id_base < 0 || id_base + n > 0x7fffffffLL) return fail(VS_ERR_INVALID, "ids must fit int32");
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) {
        cudaGetLastError();
        return fail(VS_ERR_CUDA, "no CUDA device available (libvsb200 has no CPU fallback)");
    }
    if (device < 0 || device >= cnt) return fail(VS_ERR_INVALID, "bad device ordinal");
    VSB_CUDA(cudaSetDevice(device));
    vs_int8* h = new (std::nothrow) vs_int8();
    if (!h) return fail(VS_ERR_NOMEM, "host allocation failed");
    h->device = device;
    h->n = n;
    h->dim = dim;
    h->id_base = id_base;
    cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device);
    float* d_stage = nullptr;
    float* d_max = nullptr;
    auto body = [&]() -> int {
        VSB_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        const int64_t n_pad = ceil_div64(n, 128) * 128;
        VSB_CUDA(cudaMalloc((void**)&h->d_base, (size_t)n_pad * 128));
        VSB_CUDA(cudaMemsetAsync(h->d_base + (size_t)n * 128, 0, (size_t)(n_pad - n) * 128, h->stream));
        VSB_CUDA(cudaMalloc((void**)&d_max, sizeof(float)));
        VSB_CUDA(cudaMemsetAsync(d_max, 0, sizeof(float), h->stream));
        // host base: staged through the device in chunks of <= 4M rows (2 GB fp32), twice when the weight scale has
        // to be derived from max(base) first
        const int64_t chunk = std::min<int64_t>(n, 1 << 22);
        if (!on_device) VSB_CUDA(cudaMalloc((void**)&d_stage, sizeof(float) * (size_t)chunk * 128));
        if (!(w_scale > 0.f)) {
            for (int64_t r0 = 0; r0 < n; r0 += chunk) {
                const int64_t rows = std::min(chunk, n - r0);
                const float* src = base + (size_t)r0 * 128;
                if (!on_device) {
                    VSB_CUDA(cudaMemcpyAsync(d_stage, src, sizeof(float) * (size_t)rows * 128, cudaMemcpyHostToDevice, h->stream));
                    src = d_stage;
                }
                VSB_TRY(launch_max_f32(src, rows * 128, d_max, h->stream));
            }
            float mx = 0.f;
            VSB_CUDA(cudaMemcpyAsync(&mx, d_max, sizeof(float), cudaMemcpyDeviceToHost, h->stream));
            VSB_CUDA(cudaStreamSynchronize(h->stream));
            w_scale = mx > 0.f ? mx / 255.0f : 1.0f;  // min/max encoding with offset 0 (SIFT components are >= 0)
        }
        h->s_in = in_scale;
        h->s_w = w_scale;
        h->s_out = out_scale;
        h->inv_in = 1.0f / in_scale;
        h->m = mul_rn(in_scale, w_scale) / out_scale;
        if (!(h->m > 0.f) || !std::isfinite(h->m)) return fail(VS_ERR_INVALID, "scales give a non-positive requantisation multiplier");
        const float inv_w = 1.0f / w_scale;
        for (int64_t r0 = 0; r0 < n; r0 += chunk) {
            const int64_t rows = std::min(chunk, n - r0);
            const float* src = base + (size_t)r0 * 128;
            if (!on_device) {
                VSB_CUDA(cudaMemcpyAsync(d_stage, src, sizeof(float) * (size_t)rows * 128, cudaMemcpyHostToDevice, h->stream));
                src = d_stage;
            }
            VSB_TRY(launch_quantize_u8(src, rows * 128, inv_w, h->d_base + (size_t)r0 * 128, h->stream));
        }
        VSB_TRY(make_tmap_2d(&h->tmB, h->d_base, (uint64_t)n_pad, 128, 1, 128));
        VSB_TRY(int8_set_attributes());
        VSB_CUDA(cudaStreamSynchronize(h->stream));
        return VS_OK;
    };
    const int rc = body();
    if (d_stage) cudaFree(d_stage);
    if (d_max) cudaFree(d_max);
    if (rc != VS_OK) {
        const std::string keep = vs_last_error();
        int8_free(h);
        return fail(rc, keep);
    }
    *out = h;
    return VS_OK;
}

// q_dev fp32 [nq x 128] -> out_ids [nq x k], out_scores u8 [nq x k] (device)
static int int8_search_core(vs_int8* h, const float* q_dev, int64_t nq, int k, int32_t* out_ids, uint8_t* out_scores,
                            cudaStream_t st) {
    h->last_launches = 0;
    if (nq == 0) return VS_OK;
    if (nq > 0x7fffffff / 128) return fail(VS_ERR_INVALID, "nq too large");
    const int ktop = round_up_ktop(k);
    if (ktop == 0) return fail(VS_ERR_UNSUPPORTED, "INT8 search: k > 32 is not implemented");
    const int64_t nq_pad = ceil_div64(nq, 128) * 128;
    VSB_TRY(h->q_u8.reserve((size_t)nq_pad * 128));
    // small batches: 4 (2) copies of the quantised queries fill the 128-row tile, one per TMEM lane quadrant (pair)
    const int rep = nq <= 32 ? 4 : (nq <= 64 ? 2 : 1);
    if (rep > 1) VSB_CUDA(cudaMemsetAsync(h->q_u8.p, 0, 128 * 128, st));
    VSB_TRY(launch_quantize_u8(q_dev, nq * 128, h->inv_in, h->q_u8.as<uint8_t>(), st));
    for (int c = 1; c < rep; ++c)
        VSB_CUDA(cudaMemcpyAsync(h->q_u8.as<uint8_t>() + (size_t)c * (128 / rep) * 128, h->q_u8.p, (size_t)nq * 128,
                                 cudaMemcpyDeviceToDevice, st));
    // batches of more than one query tile: two query tiles per unit share every base tile (int8_tc_pair_kernel; the plan then
    // counts tile PAIRS as its columns); VSB_INT8_PAIR=0 keeps one tile per unit
    const char* pe = getenv("VSB_INT8_PAIR");
    const bool pair = nq > 128 && !(pe && atoi(pe) == 0);
    const int64_t n_mt = ceil_div64(nq, 128);
    const TcPlan plan = tc_make_plan(h->n, pair ? ceil_div64(n_mt, 2) * 128 : nq, h->num_sms, 2);
    const int n_lists = pair ? plan.n_splits : plan.n_splits * int8_lists_per_split() * rep;
    VSB_TRY(h->part_key.reserve(sizeof(float) * (size_t)n_lists * nq * ktop));
    VSB_TRY(h->part_id.reserve(sizeof(int32_t) * (size_t)n_lists * nq * ktop));
    VSB_TRY(h->gthr.reserve(sizeof(int32_t) * (size_t)nq));
    VSB_TRY(h->out_keys.reserve(sizeof(float) * (size_t)nq * k));
    VSB_CUDA(cudaMemsetAsync(h->gthr.p, 0x7f, sizeof(int32_t) * (size_t)nq, st));
    CUtensorMap tmA;
    VSB_TRY(make_tmap_2d(&tmA, h->q_u8.p, (uint64_t)(rep > 1 ? 128 : nq), 128, 1, 128));
    if (h->profile) VSB_CUDA(cudaEventRecord(h->ev0, st));
    VSB_TRY(launch_int8_tc(tmA, h->tmB, h->gthr.as<int32_t>(), h->m, (int)nq, h->n, plan, ktop, pair ? 0 : rep, h->part_key.as<float>(),
                           h->part_id.as<int32_t>(), st));
    if (h->profile) {
        VSB_CUDA(cudaEventRecord(h->ev1, st));
        h->ev_valid = true;
    }
    // keys are -score: merge ascending, store +score
    VSB_TRY(launch_merge_lists(h->part_key.as<float>(), h->part_id.as<int32_t>(), n_lists, nq, ktop, ktop, k, h->id_base, 0, 1,
                               h->out_keys.as<float>(), out_ids, k, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                               st));
    VSB_TRY(launch_scores_to_u8(h->out_keys.as<float>(), nq * k, out_scores, st));
    h->last_launches = 4;
    return VS_OK;
}

extern "C" {

int vs_int8_create(vs_int8_t** out, const float* base, int64_t n, int dim, float in_scale, float w_scale, float out_scale,
                   int device, int64_t id_base) {
    return int8_create_common(out, base, false, n, dim, in_scale, w_scale, out_scale, device, id_base);
}
int vs_int8_create_dev(vs_int8_t** out, const float* base_dev, int64_t n, int dim, float in_scale, float w_scale,
                       float out_scale, int device, int64_t id_base) {
    return int8_create_common(out, base_dev, true, n, dim, in_scale, w_scale, out_scale, device, id_base);
}
int vs_int8_destroy(vs_int8_t* h) { return int8_free(h); }
int64_t vs_int8_num_docs(const vs_int8_t* h) { return h ? h->n : 0; }
int vs_int8_dim(const vs_int8_t* h) { return h ? h->dim : 0; }
float vs_int8_output_scale(const vs_int8_t* h) { return h ? h->s_out : 0.f; }
int vs_int8_scales(const vs_int8_t* h, float* in_scale, float* w_scale, float* out_scale, float* multiplier) {
    if (!h) return fail(VS_ERR_INVALID, "handle is NULL");
    if (in_scale) *in_scale = h->s_in;
    if (w_scale) *w_scale = h->s_w;
    if (out_scale) *out_scale = h->s_out;
    if (multiplier) *multiplier = h->m;
    return VS_OK;
}

int vs_int8_search_dev(vs_int8_t* h, const float* queries_dev, int64_t nq, int k, int32_t* out_ids_dev,
                       uint8_t* out_scores_dev, void* stream) {
    if (!h) return fail(VS_ERR_INVALID, "handle is NULL");
    if (nq < 0 || k <= 0) return fail(VS_ERR_INVALID, "nq < 0 or k <= 0");
    if ((int64_t)k > h->n) return fail(VS_ERR_INVALID, "k > number of documents");
    if (nq > 0 && (!queries_dev || !out_ids_dev || !out_scores_dev)) return fail(VS_ERR_INVALID, "NULL buffer");
    VSB_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    return int8_search_core(h, queries_dev, nq, k, out_ids_dev, out_scores_dev, st);
}

int vs_int8_search(vs_int8_t* h, const float* queries, int64_t nq, int k, int32_t* out_ids, uint8_t* out_scores) {
    if (!h) return fail(VS_ERR_INVALID, "handle is NULL");
    if (nq < 0 || k <= 0) return fail(VS_ERR_INVALID, "nq < 0 or k <= 0");
    if ((int64_t)k > h->n) return fail(VS_ERR_INVALID, "k > number of documents");
    if (nq == 0) return VS_OK;
    if (!queries || !out_ids || !out_scores) return fail(VS_ERR_INVALID, "NULL buffer");
    VSB_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    VSB_TRY(h->q.reserve(sizeof(float) * (size_t)nq * 128));
    VSB_TRY(h->out_ids.reserve(sizeof(int32_t) * (size_t)nq * k));
    VSB_TRY(h->out_u8.reserve((size_t)nq * k));
    VSB_CUDA(cudaMemcpyAsync(h->q.p, queries, sizeof(float) * (size_t)nq * 128, cudaMemcpyHostToDevice, st));
    VSB_TRY(int8_search_core(h, h->q.as<float>(), nq, k, h->out_ids.as<int32_t>(), h->out_u8.as<uint8_t>(), st));
    VSB_CUDA(cudaMemcpyAsync(out_ids, h->out_ids.p, sizeof(int32_t) * (size_t)nq * k, cudaMemcpyDeviceToHost, st));
    VSB_CUDA(cudaMemcpyAsync(out_scores, h->out_u8.p, (size_t)nq * k, cudaMemcpyDeviceToHost, st));
    VSB_CUDA(cudaStreamSynchronize(st));
    return VS_OK;
}

int vs_int8_scores_raw(vs_int8_t* h, const float* queries, int64_t nq, uint8_t* out_scores) {
    if (!h) return fail(VS_ERR_INVALID, "handle is NULL");
    if (nq < 0) return fail(VS_ERR_INVALID, "nq < 0");
    if (nq == 0) return VS_OK;
    if (!queries || !out_scores) return fail(VS_ERR_INVALID, "NULL buffer");
    if ((double)nq * (double)h->n > 4e9) return fail(VS_ERR_UNSUPPORTED, "raw score matrix larger than 4 GB: use vs_int8_search");
    VSB_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    VSB_TRY(h->q.reserve(sizeof(float) * (size_t)nq * 128));
    VSB_TRY(h->q_u8.reserve((size_t)ceil_div64(nq, 128) * 128 * 128));
    VSB_TRY(h->raw.reserve((size_t)nq * h->n));
    VSB_CUDA(cudaMemcpyAsync(h->q.p, queries, sizeof(float) * (size_t)nq * 128, cudaMemcpyHostToDevice, st));
    VSB_TRY(launch_quantize_u8(h->q.as<float>(), nq * 128, h->inv_in, h->q_u8.as<uint8_t>(), st));
    VSB_TRY(launch_int8_scores(h->d_base, h->n, h->q_u8.as<uint8_t>(), nq, h->m, h->raw.as<uint8_t>(), st));
    VSB_CUDA(cudaMemcpyAsync(out_scores, h->raw.p, (size_t)nq * h->n, cudaMemcpyDeviceToHost, st));
    VSB_CUDA(cudaStreamSynchronize(st));
    return VS_OK;
}

int vs_int8_quantize(const float* src, int64_t count, float scale, uint8_t* dst, int device) {
    if (count < 0 || !(scale > 0.f)) return fail(VS_ERR_INVALID, "count < 0 or scale <= 0");
    if (count == 0) return VS_OK;
    if (!src || !dst) return fail(VS_ERR_INVALID, "NULL buffer");
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) {
        cudaGetLastError();
        return fail(VS_ERR_CUDA, "no CUDA device available (libvsb200 has no CPU fallback)");
    }
    if (device < 0 || device >= cnt) return fail(VS_ERR_INVALID, "bad device ordinal");
    VSB_CUDA(cudaSetDevice(device));
    float* d_src = nullptr;
    uint8_t* d_dst = nullptr;
    int rc = VS_OK;
    do {
        if (cudaMalloc((void**)&d_src, sizeof(float) * (size_t)count) != cudaSuccess ||
            cudaMalloc((void**)&d_dst, (size_t)count + 4) != cudaSuccess) {
            rc = fail(VS_ERR_NOMEM, "cudaMalloc failed");
            break;
        }
        if (cudaMemcpy(d_src, src, sizeof(float) * (size_t)count, cudaMemcpyHostToDevice) != cudaSuccess) {
            rc = fail(VS_ERR_CUDA, "H2D failed");
            break;
        }
        rc = launch_quantize_u8(d_src, count, 1.0f / scale, d_dst, nullptr);
        if (rc != VS_OK) break;
        if (cudaMemcpy(dst, d_dst, (size_t)count, cudaMemcpyDeviceToHost) != cudaSuccess) rc = fail(VS_ERR_CUDA, "D2H failed");
    } while (0);
    if (d_src) cudaFree(d_src);
    if (d_dst) cudaFree(d_dst);
    return rc;
}

int vs_int8_set_profile(vs_int8_t* h, int enable) {
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

int vs_int8_last_kernel_ms(vs_int8_t* h, float* ms) {
    if (!h || !ms) return fail(VS_ERR_INVALID, "NULL argument");
    if (!h->ev_valid) return fail(VS_ERR_INVALID, "no profiled search yet (vs_int8_set_profile)");
    VSB_CUDA(cudaEventSynchronize(h->ev1));
    VSB_CUDA(cudaEventElapsedTime(ms, h->ev0, h->ev1));
    return VS_OK;
}

}  // extern "C"
