// Launch wrappers implemented in the .cu files of libvsb200 (internal interface between api.cu and the kernels).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vsb {

// per-call constants of the scaled-fp16 candidate pass, derived on the device from the queries (no host round trip)
struct TcQueryParams {
    float s_q;          // power-of-two scale of the fp16 query copy (the copy holds -s_q * q)
    float key_unscale;  // key = acc * key_unscale,  2 / (s_q * s_b): the accumulator already contains the norm term
    float cert_a;       // |key_f16 - key_exact| <= cert_a * sqrt(qn) + cert_b
    float cert_b;
    float bn_max;
    int fold_ok;        // 0: s_q / s_b is outside the range in which the norm block's query-side constants are fp16 numbers
                        // -> the pass cannot run, every query is reported uncertified (and redone on the fp32 path)
};
// The norm term of the candidate pass travels through the tensor core as an extra K = 16 block (exact_tc.cuh):
// base side  e[row][16] fp16 = { p1 x8, p2 x4, p3, 0, 0, 0 },  U = s_b^2 * ||x||^2 / 2 = 2^13 * p1 + 2^2 * p2 + 2^-8 * p3
// query side a[16]      fp16 = { 2^(rho+10) x8, 2^rho x4, 2^(rho-8), 0, 0, 0 },  rho = log2(s_q / s_b)
constexpr int TC_FOLD_COLS = 16;
int launch_norm_pieces(const float* bnorm, int64_t n, int64_t n_pad, float s_b, void* e_half, cudaStream_t st);
int launch_absmax_f32(const float* x, int64_t count, float* out_zeroed, cudaStream_t st);
float f16_scale_host(float absmax);
// dim == 128 only: norms (reference summation order) + abs-max of the whole batch in one pass
int launch_query_prep(const float* x, int64_t rows, float* norms, float* absmax_zeroed, cudaStream_t st);
// also writes the query-side norm block qe_half[128][16] (every row the same)
int launch_tc_query_params(const float* q_absmax, float s_b, float bn_max, TcQueryParams* out, void* qe_half, cudaStream_t st);
// out = fp16(x * scale), or fp16(x * -scale_dev->s_q) when scale_dev != nullptr (the NEGATED scaled query copy)
int launch_to_half_scaled(const float* x, int64_t count, float scale, const TcQueryParams* scale_dev, void* out_half,
                          cudaStream_t st);
int launch_gather_rows(const float* src, const int32_t* idx, int n_idx, float* dst, cudaStream_t st);
int launch_scatter_results(const float* key, const int32_t* id, const int32_t* idx, int n_idx, int k, float* out_key,
                           int32_t* out_id, cudaStream_t st);

// prep.cu ---------------------------------------------------------------------------------------
// ||x||^2 per row in the reference's summation order (cpu_baseline.cpp:95-114) and, when hi/lo are non-null,
// the TF32 split x = hi + lo (hi = rna_tf32(x), lo = rna_tf32(x - hi)).  *not_tf32_exact is OR-ed with 1 when
// some lo != 0.  norms may be null.
int launch_prep_rows(const float* x, int64_t rows, int dim, float* norms, float* hi, float* lo, int* not_tf32_exact,
                     cudaStream_t st);
int launch_fill_f32(float* p, int64_t n, float v, cudaStream_t st);

// exact_stream.cu -------------------------------------------------------------------------------
// K2: HBM-streaming FFMA scan for up to 8 queries at a time. Writes sorted partial lists
// part[(cta*nq + q)*ktop + i]; returns the number of partial lists per query in *n_parts.
int stream_num_ctas(int device, int nq);
int launch_exact_stream(const float* base, const float* bnorm, int64_t n, const float* q, const float* qnorm, int nq,
                        int ktop, const float* lb_key, const int32_t* lb_id, float* part_key, int32_t* part_id,
                        int n_ctas, cudaStream_t st);

// exact_tc.cu -----------------------------------------------------------------------------------
struct TcPlan {
    int n_tiles, n_mtiles, n_splits, tiles_per_split, grid;
    int cl;  // CTAs per cluster: 2 = CTA pairs sharing the streamed base tiles (TMA multicast), 1 = independent CTAs
};
// tensor maps of a streamed base operand: hi (or the only) part and lo part, as full-tile boxes (128 rows) and as
// half-tile boxes (64 rows: each CTA of a pair fetches one half and multicasts it)
struct TcBaseMaps {
    CUtensorMap hi, lo, hi_half, lo_half;
};
TcPlan tc_make_plan(int64_t n, int64_t nq, int num_sms, int mode = 0);  // mode: exact_tc.cuh TcMode
// mode: 0 = 1xTF32, 1 = 3xTF32, 2 = scaled fp16 candidate pass (exact_tc.cuh TcMode; keys stay in accumulator units,
// tmA_lo = the query-side norm block [128 x 16] fp16, tmB.lo = the base-side norm block [n_pad x 16] fp16, both SWIZZLE_32B)
int launch_exact_tc(const CUtensorMap& tmA_hi, const CUtensorMap& tmA_lo, const TcBaseMaps& tmB, const float* bnorm,
                    int32_t* gthr, int nq, int64_t n_rows, const TcPlan& plan,
                    int ktop, int mode, const float* key_scale_dev, const float* lb_key, const int32_t* lb_id, float* part_key,
                    int32_t* part_id, cudaStream_t st);
int launch_exact_tc_ivf(const CUtensorMap& tmA_hi, const CUtensorMap& tmA_lo, const CUtensorMap& tmB_hi, const CUtensorMap& tmB_lo,
                        const int4* items, const int32_t* n_items, const int32_t* pairs, int nprobe, int32_t* gthr, int nq, int ktop,
                        bool split3, int32_t* cand_cnt, void* cand, int cand_cap, int num_sms, cudaStream_t st);
// the fp16 threshold-filter candidate pass: sample pass -> thresholds -> filter pass -> filter merge
int launch_exact_tc_f16(const CUtensorMap& tmA, const CUtensorMap& tmA_fold, const TcBaseMaps& tmB, int nq, int64_t n_rows,
                        const TcPlan& plan, bool sample, int tile_stride, int tile_off, float* smin, const float* thr,
                        int32_t* cand_cnt, void* cand, int cand_cap, cudaStream_t st);
int tc_sample_groups_per_split();
int launch_tc_select_thr(const float* smin, int n_groups, int nq, int m, float* thr, int32_t* cand_cnt, cudaStream_t st);
int launch_tc_fill_thr(int nq, float* thr, int32_t* cand_cnt, cudaStream_t st);
int launch_filter_merge(const void* cand, const int32_t* cand_cnt, int cap, const float* thr, int64_t nq, int k, int64_t id_base,
                        float* out_key, int32_t* out_id, int out_stride, const float* rf_base, const float* rf_bnorm, const float* rf_q,
                        const float* rf_qnorm, const TcQueryParams* cert_qp, int32_t* uncert_count, int32_t* uncert_list,
                        cudaStream_t st);
int launch_filter_merge_ivf(const void* cand, const int32_t* cand_cnt, int cap, int64_t nq, int k, float* out_scores, int32_t* out_ids,
                            const float* vectors, const float* q, const int32_t* ivf_idmap, cudaStream_t st);
int tc_lists_per_split(int mode);  // partial lists written per (split, query): 1 (TC_F16, mode 2) or 3
int tc_set_attributes();  // opt-in to > 48 KB dynamic shared memory for every instantiation

// kernels.cu (K3) ----------------------------------------------------------------------------
// Merges n_lists sorted lists of `list_len` (<= 32) entries per query (layout [list][nq][list_len]): selects the
// `nsel` best by the candidate keys, optionally recomputes their distances in exact fp32 (rf_* non-null; local
// ids), orders them by (key asc, id asc) and writes the first k to out[q*out_stride + out_off ...]; adds id_base
// to valid ids; neg_in / neg_out flip the key sign on load / store (descending scores are handled as ascending
// negated keys).  lb_*_out receive the last selected candidate (exclusive lower bound of the next pass, k > 32).
// cert_qp != nullptr (needs the refine): queries whose answer cannot be certified complete from the bounded-error
// candidate pass are appended to uncert_list / counted in *uncert_count (zeroed by the caller).
int launch_merge_lists(const float* part_key, const int32_t* part_id, int n_lists, int64_t nq, int list_len, int nsel,
                       int k, int64_t id_base, int neg_in, int neg_out, float* out_key, int32_t* out_id, int out_stride,
                       int out_off, float* lb_key_out, int32_t* lb_id_out, const float* rf_base, const float* rf_bnorm,
                       const float* rf_q, const float* rf_qnorm, cudaStream_t st, const TcQueryParams* cert_qp = nullptr,
                       int32_t* uncert_count = nullptr, int32_t* uncert_list = nullptr,
                       // lists inside per-shard exchange blocks: list l starts list_stride 4-byte words after list l-1
                       // (0 = dense [list][nq][len]); trailer: word 0 of each block's trailer, summed into *trailer_total_out
                       size_t list_stride = 0, const int32_t* trailer = nullptr, int32_t* trailer_total_out = nullptr,
                       // IVF re-score (rf_base = list-contiguous vectors, rf_q = queries, ids = row positions): the selected
                       // candidates get their inner product in the reference's NEON order (IVFIndex.cpp:278-357), key = -score,
                       // and are ranked / returned with their ORIGINAL ids ivf_idmap[position]
                       const int32_t* ivf_idmap = nullptr);
int launch_sort_rows(float* key, int32_t* id, int64_t nq, int k, cudaStream_t st);
// G <= 32 sorted per-shard lists [G][nq][k] of any k -> the k best per query, canonical order (neg: largest key first)
int launch_merge_shards(const float* in_key, const int32_t* in_id, int n_shards, int64_t nq, int k, int neg, float* out_key,
                        int32_t* out_id, cudaStream_t st, size_t list_stride = 0, const int32_t* trailer = nullptr,
                        int32_t* trailer_total_out = nullptr);

// synth.cu --------------------------------------------------------------------------------------
int launch_synth(float* out, int64_t row0, int64_t nrows, int dim, int law, uint64_t seed, uint64_t centre_seed,
                 cudaStream_t st);

// kmeans.cu -------------------------------------------------------------------------------------
// greedy k-means++ seeding (sklearn's rule: 2 + floor(ln k) candidates per centre drawn ~ D^2, keep the one that lowers the
// potential most), deterministic for a given seed; base_dev [n x 128] -> cent_dev [k x 128]; synchronises `st`
int launch_kmeanspp(const float* base_dev, int64_t n, int k, uint64_t seed, float* cent_dev, cudaStream_t st);
// *changed_zeroed += number of rows with cur != prev; prev = cur
int launch_labels_changed(const int32_t* cur, int32_t* prev, int64_t n, int32_t* changed_zeroed, cudaStream_t st);
// stable counting sort by label: offsets [nlist+1] (CSR), members [n] (ascending row ids inside a list, np.where order)
size_t multisplit_workspace_ints(int64_t n, int nlist);
int launch_multisplit(const int32_t* lab, int64_t n, int nlist, int32_t* ws, int32_t* offsets, int32_t* members, cudaStream_t st);
// *out = sum of max(dist, 0) in double, fixed order; part_ws: ceil(n/1024) doubles
int launch_inertia(const float* dist, int64_t n, double* part_ws, double* out, cudaStream_t st);

// ivf.cu ----------------------------------------------------------------------------------------
int launch_ivf_coarse(const float* q, int64_t nq, const float* cent, int nlist, float* scores, cudaStream_t st);
int launch_ivf_probes(const float* scores, int64_t nq, int nlist, int nprobe, int32_t* probes, cudaStream_t st);
int launch_ivf_scan(const CUtensorMap& tmV, const float* q, const int32_t* probes, const int32_t* offsets,
                    const int32_t* id_map, int64_t nq, int nprobe, int k, float* out_scores, int32_t* out_ids,
                    int32_t* out_counts, unsigned long long* total, cudaStream_t st);
// K8 list-major scan for large batches (ivf_lm.cu): one sorted list per (query, probe slot) -> part_key/part_id
// [nprobe][nq][round_up_ktop(k)] (key = -score), counts and total candidates; ws = ivf_lm_workspace_ints() ints
int ivf_lm_set_attributes();
size_t ivf_lm_workspace_ints(int64_t nq, int nprobe, int nlist);
int launch_ivf_listmajor(const float* q, const float* vectors, const int32_t* offsets, const int32_t* id_map, int nlist,
                         const int32_t* list_order /* lists by descending length */, const int32_t* probes, int64_t nq, int nprobe, int k, int32_t* ws, float* part_key, int32_t* part_id,
                         int32_t* out_counts, unsigned long long* total, int num_sms, cudaStream_t st);
// tensor-core list-major scan: pair grouping into work items of <= 128 pairs (long lists first) + gathered TF32-split queries
size_t ivf_tc_workspace_ints(int64_t nq, int nprobe, int nlist);
int launch_ivf_tc_prep(const float* q, const int32_t* offsets, int nlist, const int32_t* list_order, const int32_t* probes, int64_t nq,
                       int nprobe, int32_t* ws, float* qhi, float* qlo, const int4** items_out, const int32_t** n_items_out,
                       const int32_t** pairs_out, int32_t** q_cand_out, int** q_not_exact_out, cudaStream_t st);
int launch_ivf_counts(const int32_t* q_cand, int64_t nq, int k, int32_t* out_counts, unsigned long long* total, cudaStream_t st);
int ivf_set_attributes();
int ivf_scan_rows_per_chunk();

// int8_tc.cu -----------------------------------------------------------------------------------
// K4: u8 = sat(trunc(x * inv_scale + 0.5)) (QnnRunner.cpp:13-55); K5: fused u8 x u8 -> s32 tcgen05 GEMM + requant +
// largest-k, partial lists keyed by -score; helpers for the raw score matrix and the weight scale.
int launch_quantize_u8(const float* src, int64_t count, float inv_scale, uint8_t* dst, cudaStream_t st);
int launch_scores_to_u8(const float* src, int64_t count, uint8_t* dst, cudaStream_t st);
int launch_int8_scores(const uint8_t* base, int64_t n, const uint8_t* q, int64_t nq, float m, uint8_t* out, cudaStream_t st);
int launch_max_f32(const float* x, int64_t count, float* out_zeroed, cudaStream_t st);
int int8_set_attributes();
int int8_lists_per_split();
int launch_int8_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, int32_t* gthr, float m, int nq, int64_t n, const TcPlan& plan,
                   int ktop, int rep, float* part_key, int32_t* part_id, cudaStream_t st);

// api.cu helpers shared with api_ivf.cu / api_int8.cu
int make_tmap_2d(CUtensorMap* out, const void* gptr, uint64_t rows, uint64_t cols, int elem_bytes, uint32_t box_rows);
// [rows x 16] fp16 (32-byte rows), box = 32 B x box_rows, SWIZZLE_32B: the K = 16 norm block of the fp16 candidate pass
int make_tmap_fold(CUtensorMap* out, const void* gptr, uint64_t rows, uint32_t box_rows);

int round_up_ktop(int k);  // smallest supported register-list size >= k (1,5,10,16,32), 0 if k > 32

}  // namespace vsb
