// The exact-search handle and the internal entry points shared by api.cu (single-shard C ABI) and api_group.cu
// (shard groups: several shards per device, several devices per process).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <vector>

#include "kernels.cuh"
#include "vsb_common.cuh"

using vsb::DevBuf;

struct vs_exact {
    int device = 0;
    int num_sms = 148;
    int64_t n = 0;
    int dim = 0;
    int64_t id_base = 0;
    bool owns_base = false;
    const float* d_base = nullptr;  // [n x dim] fp32
    float* d_hi = nullptr;          // TF32 split, built on first use (dim == 128 only); d_hi aliases d_base when the
    float* d_lo = nullptr;          // base is TF32-exact
    bool split_ready = false;
    void* d_fold = nullptr;         // [n_pad x 16] fp16: the norm term s_b^2 ||x||^2 / 2 as an extra K = 16 operand block
    void* d_f16 = nullptr;          // [n x 128] fp16 copy scaled by s_b (candidate pass)
    float s_b = 1.f;                // power-of-two scale of the fp16 copy
    float bn_max = 0.f;             // max ||x||^2
    float* d_norm = nullptr;        // [n_tiles*128] +inf padded
    bool base_exact = false;
    bool broken = false;            // vs_exact_refresh() failed half-way
    vsb::TcBaseMaps tmB32;          // TF32 hi / lo operand maps (fp32 containers)
    vsb::TcBaseMaps tmB16;          // scaled fp16 copy (hi == lo)
    cudaStream_t stream = nullptr;
    // workspace (grow-only)
    DevBuf q, qhi, qlo, qf16, qnorm, part_key, part_id, lbk, lbi, out_ids, out_keys, flag, gthr, qparams, qfold, unc_list, fb_q,
        fb_ids, fb_keys;
    // fp16 threshold-filter candidate pass: group minima of the sample pass, per-query thresholds, candidate counters / arrays
    DevBuf f_smin, f_thr, f_cnt, f_cand;
    // batch <= 8 host calls (vs_exact_search_f32): the whole call — H2D, query norms, streaming kernel, merge, D2H — as ONE
    // CUDA graph per (nq, k), over workspaces and pinned staging buffers that only the graphs touch
    struct SmallGraph {
        int64_t nq;
        int k;
        cudaGraphExec_t exec;
    };
    std::vector<SmallGraph> graphs;
    DevBuf g_q, g_qnorm, g_part_key, g_part_id, g_ids, g_keys;
    float* hp_q = nullptr;       // pinned [8 x 128]
    int32_t* hp_ids = nullptr;   // pinned [8 x 32]
    float* hp_keys = nullptr;    // pinned [8 x 32]
    int* h_flag = nullptr;  // pinned: [0] exactness flag, [1] uncertified count
    int last_launches = 0;
    int last_precision = 0;
    int last_fallback = 0;
    // optional CUDA-event timing of the dominant kernel (bench.py's roofline line)
    bool profile = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_pre = nullptr;  // ev_pre .. ev0: sample pass + thresholds (fp16 path)
    bool ev_valid = false, ev_pre_valid = false;
    // certified search split in two (vs_exact_search_dev_begin / _finish): what finish needs to redo uncertified queries
    cudaEvent_t ev_cert = nullptr;
    bool cert_pending = false;
    const float* pend_q = nullptr;
    int64_t pend_nq = 0;
    int pend_k = 0;
    int32_t* pend_ids = nullptr;
    float* pend_dists = nullptr;
    cudaStream_t pend_st = nullptr;
};


int exact_create_common(vs_exact_t** out, const float* base, bool on_device, int64_t n, int dim, int device, int64_t id_base);
int exact_free(vs_exact* h);
// unc_dev: device word that receives the number of uncertified queries of a certified search (zeroed by the caller;
// nullptr = the handle's own word)
int exact_search_core(vs_exact* h, const float* q_dev, int64_t nq, int k, int precision, int32_t* out_ids, float* out_dists,
                      cudaStream_t st, bool defer_certification = false, int32_t* unc_dev = nullptr);
int exact_certified_finish(vs_exact* h, int* n_redone);
