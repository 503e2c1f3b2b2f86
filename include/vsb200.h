/* vsb200 — C ABI of the B200-native vector-search hot paths (libvsb200.so).
 *
 * Drop-in boundary for the two data-parallel hot paths of zyx7k/HAI-25-RAG-on-Edge (file:line relative
 * to the reference tree):
 *
 *   exact L2 kNN   replaces the triple  compute_norms -> cblas_sgemm(M=1) -> select_topk  inside
 *                  run_benchmark()                     cpu/cpu_baseline.cpp:116-153, 211-248
 *   IVF search     replaces IVFIndex::IVFIndex / search / searchBatch and the QNN coarse MatMul
 *                  qidk_ivf/android/app/main/jni/IVFIndex.h:19-54, IVFIndex.cpp:154-267, 572-859
 *                  and the builder build_ivf_index()   qidk_ivf/prepare/create_ivf_model.py:86-175
 *   INT8 brute     replaces QnnRunner::executeRaw/executeBatchRaw + find_top_k_int8
 *                  qidk_bruteforce/android/app/main/jni/QnnRunner.h:20-52, QnnRunner.cpp:13-55,529-638,
 *                  main.cpp:36-71
 *
 * Conventions: plain pointers and sizes, no C++ or torch types; every function returns a vs_status
 * (0 = ok) and never throws; vs_last_error() gives the message of the last failure on the calling thread.
 * Host buffers are owned by the caller; *_dev entry points take device pointers valid on the handle's
 * device and enqueue on the given CUDA stream (cudaStream_t passed as void*, NULL = the handle's own
 * stream) without synchronising.  A handle is not thread-safe (one in-flight search per handle);
 * different handles are independent.  There is NO CPU fallback: without a CUDA device every compute
 * entry point fails with VS_ERR_CUDA.
 *
 * Result order is canonical and deterministic: exact L2 -> (distance ascending, id ascending);
 * IVF / INT8 -> (score descending, id ascending).  Ids are global row ids (shard id_base added).
 */
#ifndef VSB200_H
#define VSB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VSB200_ABI_VERSION 1

#if defined(__GNUC__)
#define VSB_API __attribute__((visibility("default")))
#else
#define VSB_API
#endif

typedef enum vs_status {
    VS_OK = 0,
    VS_ERR_INVALID = 1,     /* bad argument (NULL, k<=0, k>n, dim mismatch, ...) */
    VS_ERR_CUDA = 2,        /* CUDA runtime/driver failure or no device */
    VS_ERR_IO = 3,          /* index directory / file problem */
    VS_ERR_NOMEM = 4,
    VS_ERR_UNSUPPORTED = 5  /* shape outside what the kernels implement */
} vs_status;

/* Arithmetic of the exact path's dot products (distances are always combined in fp32 as
 * (qn + bn) - 2*dot, cpu_baseline.cpp:241). */
typedef enum vs_precision {
    VS_PREC_AUTO = 0,        /* <= 2 queries (<= 8 where the fp16 pass does not apply) -> FFMA stream; >= 3 queries, k <= 16, base >= ~13 K rows -> certified fp16 pass
                                (below); otherwise 1xTF32 when every operand is exactly representable in TF32 (integer
                                SIFT data: bit-identical to fp32), else 3xTF32.  Thresholds measured on B200. */
    VS_PREC_FP32_3XTF32 = 1, /* tcgen05 kind::tf32, hi/lo split, 3 products, fp32 accumulate in TMEM.  The tensor-core keys
                                rank the candidates (k + 2 of them when k <= 30, exactly k beyond), every returned
                                distance is recomputed in plain fp32: ids equal the fp32 reference's except inside
                                distance ties within 1e-5 relative (the north star's tolerance; not certified) */
    VS_PREC_FP32_FFMA = 2,   /* CUDA-core FFMA streaming kernel (HBM-bound; any batch, slow for large ones) */
    VS_PREC_TF32_1X = 3,     /* single TF32 product; exact only for TF32-representable data */
    VS_PREC_F16_CERTIFIED = 4 /* fp32-faithful results from a cheaper tensor-core pass: tcgen05 kind::f16 on power-of-two
                                scaled fp16 copies collects every row whose key lies below a per-query threshold (taken
                                from a sample pass over 1/16 of the base; key error rigorously bounded), the <= 32 best
                                are recomputed in exact fp32 and the top k is certified complete per query (every
                                excluded row is provably farther than the k-th result); queries that cannot be certified
                                are redone with 3xTF32 / FFMA.  k <= 16; bases of 513 .. ~13 K rows take 3xTF32 instead
                                (no meaningful sample).  Synchronises the stream once. */
} vs_precision;

VSB_API const char* vs_last_error(void);
VSB_API int vs_abi_version(void);
/* number of visible CUDA devices (0 and VS_ERR_CUDA when there is none) */
VSB_API int vs_device_count(int* count);

/* ---------------------------------------------------------------------------------------------- */
/* Exact L2 kNN                                                                                    */
/* ---------------------------------------------------------------------------------------------- */
typedef struct vs_exact vs_exact_t;

/* Builds the device-resident index for base[n x dim] (row-major fp32, host memory): copies the rows,
 * precomputes ||x||^2 in the reference's summation order (cpu_baseline.cpp:95-114) and the TF32 hi/lo split.
 * device = CUDA ordinal; id_base is added to every returned id (row-sharded multi-GPU use).  dim must be 128 (the
 * SIFT shape of every BASELINE config): any other dim fails with VS_ERR_UNSUPPORTED. */
VSB_API int vs_exact_create(vs_exact_t** out, const float* base, int64_t n, int dim, int device, int64_t id_base);
/* Same, base already resident on `device` (not copied; must outlive the handle). */
VSB_API int vs_exact_create_dev(vs_exact_t** out, const float* base_dev, int64_t n, int dim, int device, int64_t id_base);
VSB_API int vs_exact_destroy(vs_exact_t* h);
/* The rows behind a vs_exact_create_dev handle were modified in place (same pointer, same shape): recompute the
 * norms and the TF32 split.  Used by the k-means builder, whose centroids move every iteration. */
VSB_API int vs_exact_refresh(vs_exact_t* h);
VSB_API int64_t vs_exact_size(const vs_exact_t* h);
VSB_API int vs_exact_dim(const vs_exact_t* h);
/* 1 if every base component is exactly representable in TF32 (then 1xTF32 == 3xTF32 == fp32 bit for bit) */
VSB_API int vs_exact_base_is_tf32_exact(const vs_exact_t* h);

/* queries[nq x dim] host -> out_ids[nq x k] (int32), out_dists[nq x k] (squared L2, ascending).
 * Timed span equivalent to the reference's (cpu_baseline.cpp:220-257): query upload + search + download. */
VSB_API int vs_exact_search_f32(vs_exact_t* h, const float* queries, int64_t nq, int k, int precision,
                        int32_t* out_ids, float* out_dists);
/* Device-pointer variant; asynchronous on `stream`. */
VSB_API int vs_exact_search_dev(vs_exact_t* h, const float* queries_dev, int64_t nq, int k, int precision,
                        int32_t* out_ids_dev, float* out_dists_dev, void* stream);
/* The same in two halves, for callers that want to enqueue more work (e.g. the all-gather of a sharded search) before
 * the host waits for the certification count of VS_PREC_F16_CERTIFIED / AUTO: _begin enqueues and returns without any
 * host synchronisation; _finish waits for the count only (not for work enqueued after _begin), redoes the uncertified
 * queries on the same stream and reports how many result rows it rewrote (0 in the common case: whatever consumed the
 * results in between is still valid).  Exactly one _finish per _begin. */
VSB_API int vs_exact_search_dev_begin(vs_exact_t* h, const float* queries_dev, int64_t nq, int k, int precision,
                                      int32_t* out_ids_dev, float* out_dists_dev, void* stream);
VSB_API int vs_exact_search_dev_finish(vs_exact_t* h, int* n_redone);
/* Introspection for benchmarks: kernels launched / tensor-core or streaming kernel device time of the last
 * search is measured by the caller with CUDA events on the stream; this returns how many kernels the last
 * search launched and which precision path AUTO resolved to. */
VSB_API int vs_exact_last_launches(const vs_exact_t* h, int* n_kernels, int* precision_used);
/* number of queries of the last search that the certified path could not certify and redid on the fp32 path */
VSB_API int vs_exact_last_fallbacks(const vs_exact_t* h, int* n_queries);
/* When enabled, every search brackets its dominant kernel (the fused distance+top-k kernel: tcgen05 or FFMA
 * stream) with CUDA events on the launching stream; vs_exact_last_kernel_ms waits for and returns that duration. */
VSB_API int vs_exact_set_profile(vs_exact_t* h, int enable);
VSB_API int vs_exact_last_kernel_ms(vs_exact_t* h, float* ms);
/* VS_PREC_F16_CERTIFIED only: device time of what runs before the dominant (filter) kernel to produce the per-query
 * thresholds — the sample pass over one base tile in 16 and the threshold selection.  0 for the other paths. */
VSB_API int vs_exact_last_prepass_ms(vs_exact_t* h, float* ms);

/* Test hook for the certification bound of VS_PREC_F16_CERTIFIED: runs ONLY the fp16 tensor-core candidate generation
 * (sample pass, thresholds, filter pass) and returns, per query, the up to 32 candidates the filter merge keeps (24 .. 32
 * when more rows lie below the threshold) as the kernel ranked them — out_ids[nq x 32] (local row ids, -1 padded),
 * out_keys[nq x 32] (the kernel's keys  ||x||^2 - 2 q.x  in distance units, ascending) — and out_bound[nq], the bound
 * E_q = cert_a*sqrt(||q||^2) + cert_b the certificate assumes for |key - exact key| (tests/test_exact_gpu.py measures
 * the actual error against float64). Host buffers. */
VSB_API int vs_exact_debug_f16_candidates(vs_exact_t* h, const float* queries, int64_t nq, int32_t* out_ids,
                                          float* out_keys, float* out_bound);

/* Merge per-shard results gathered from G shards (layout [G][nq][k], as produced by an all-gather of every
 * rank's out_ids/out_dists) into the global top-k per query, canonical order. smallest!=0: keys ascending
 * (L2); smallest==0: keys descending (inner product). Device pointers, asynchronous on `stream`. */
VSB_API int vs_merge_topk_dev(const int32_t* ids_dev, const float* keys_dev, int n_shards, int64_t nq, int k,
                      int smallest, int32_t* out_ids_dev, float* out_keys_dev, void* stream);

/* ---------------------------------------------------------------------------------------------- */
/* Row-sharded exact search (SURVEY.md §8b "n_gpus", §8e): base rows partitioned into shards, queries    */
/* replicated, every shard returns its local top-k, ONE exchange step, merge.  The canonical (distance, id) */
/* order makes the merged answer independent of the number of shards.                                  */
/* ---------------------------------------------------------------------------------------------- */
/* Exchange block of one shard for nq queries x k results: ids [nq x k] int32 | keys [nq x k] fp32 | 16-byte trailer
 * (word 0 = number of queries the shard could not certify).  ids and keys travel in ONE collective and the trailer
 * carries the "was anything redone" information, so the exchange needs no second collective and no host round trip. */
VSB_API size_t vs_topk_block_bytes(int64_t nq, int k);
/* Merge n_shards blocks (block s at blocks_dev + s*block_stride) into the global top-k per query; total_dev (may be
 * NULL) receives the sum of the trailers' counts.  Any k; device pointers, asynchronous on `stream`. */
VSB_API int vs_merge_blocks_dev(const void* blocks_dev, int n_shards, size_t block_stride, int64_t nq, int k, int smallest,
                                int32_t* out_ids_dev, float* out_keys_dev, int32_t* total_dev, void* stream);

/* Push form of the exchange step over NVLink peer memory (one process per GPU whose gathered buffers are mapped into each
 * other's address space: CUDA IPC / symmetric memory).  vs_push_block_dev: ONE kernel stores `bytes` (a multiple of 16; this
 * participant's slots; 0 = signal only) from src_dev to the same offset of every peer's gathered buffer — dst_dev is a HOST
 * array of n_dst (<= 16) device pointers — and, once every store is fenced at system scope, writes `epoch` into each peer's
 * flag word flag_dst[i] (peer-mapped, one word per sender).  counter_dev: a zeroed 4-byte scratch word on this device.
 * vs_wait_flags_dev: a one-block kernel that returns when the n flag words at flags_dev (skipping index `self`) have all
 * reached `epoch` — the receiving half of the barrier.  Both are asynchronous on `stream`; together they replace the NCCL
 * all-gather (hai-25-rag-on-edge_b200/sharded.py).  Epochs must increase by one per exchange on every participant. */
VSB_API int vs_push_block_dev(const void* src_dev, void* const* dst_dev, int n_dst, size_t bytes, void* const* flag_dst,
                              uint32_t epoch, uint32_t* counter_dev, void* stream);
VSB_API int vs_wait_flags_dev(const uint32_t* flags_dev, int n, int self, uint32_t epoch, void* stream);

/* The shards that live on ONE device, as one participant of the exchange (one process per GPU: bench.py / torchrun;
 * or several shards on one GPU).  The group does not own the shard handles.  Slots first_slot .. first_slot+n_local-1
 * of the gathered buffer [n_slots][vs_topk_block_bytes(nq, k)] belong to this group.
 *   begin   enqueues the local searches; every shard writes its block IN PLACE into its slot (no host synchronisation)
 *   <the caller's exchange: an in-place all-gather of the slots; nothing when every shard is local>
 *   merge   enqueues the merge of all slots and the 4-byte total of the uncertified counts on its way to the host
 *   finish  waits for that total only.  0 in the common case.  Otherwise the local shards redo their uncertified queries
 *           on the fp32 path (rewriting rows of their blocks) and *need_reexchange = 1: exchange, merge and finish again.
 *           Every participant sees the same total, so all of them take the same decision without a collective. */
typedef struct vs_exact_group vs_exact_group_t;
VSB_API int vs_exact_group_create_from(vs_exact_group_t** out, int n_local, vs_exact_t* const* shards, int n_slots,
                                       int first_slot);
VSB_API int vs_exact_group_destroy(vs_exact_group_t* g);
VSB_API int vs_exact_group_begin(vs_exact_group_t* g, const float* queries_dev, int64_t nq, int k, int precision,
                                 void* gathered_dev, void* stream);
VSB_API int vs_exact_group_merge(vs_exact_group_t* g, int32_t* out_ids_dev, float* out_dists_dev);
VSB_API int vs_exact_group_finish(vs_exact_group_t* g, int* need_reexchange);

/* Single process, n_gpus devices (0 = all visible), shards_per_gpu shards on each (>= 1; > 1 is for tests and for
 * bases whose per-GPU share should be searched in pieces): the drop-in for run_benchmark()'s hot triple
 * (cpu/cpu_baseline.cpp:177-257) on a multi-GPU box.  Contiguous row ranges, one worker thread and one stream per GPU,
 * ncclCommInitAll + one grouped ncclAllGather of the exchange blocks over NVLink (NCCL is loaded with dlopen at the
 * first multi-GPU create; n_gpus = 1 needs no NCCL), device 0 merges and returns.  Host buffers in and out; the
 * queries cross PCIe once (every GPU uploads its slice, the slices are replicated over NVLink). */
typedef struct vs_exact_mgpu vs_exact_mgpu_t;
VSB_API int vs_exact_mgpu_create(vs_exact_mgpu_t** out, const float* base, int64_t n, int dim, int n_gpus, int shards_per_gpu);
VSB_API int vs_exact_mgpu_destroy(vs_exact_mgpu_t* m);
VSB_API int vs_exact_mgpu_num_gpus(const vs_exact_mgpu_t* m);
VSB_API int vs_exact_mgpu_num_shards(const vs_exact_mgpu_t* m);
VSB_API int vs_exact_mgpu_search_f32(vs_exact_mgpu_t* m, const float* queries, int64_t nq, int k, int precision,
                                     int32_t* out_ids, float* out_dists);
/* exchanges of the last search (1 unless a shard had to redo uncertified queries) / whether anything was redone */
VSB_API int vs_exact_mgpu_last_stats(const vs_exact_mgpu_t* m, int* n_exchanges, int* redone);
VSB_API int vs_exact_mgpu_set_profile(vs_exact_mgpu_t* m, int enable);
VSB_API int vs_exact_mgpu_last_kernel_ms(vs_exact_mgpu_t* m, float* ms);   /* slowest shard's fused kernel */

/* ---------------------------------------------------------------------------------------------- */
/* IVF two-stage search (inner-product metric, like the reference)                                 */
/* ---------------------------------------------------------------------------------------------- */
typedef struct vs_ivf vs_ivf_t;

/* Index from host arrays: vectors in list-contiguous order [n x dim] (the reference's "reordered" layout,
 * vectors_reordered.npy), CSR cluster_offsets [nlist+1], position_to_id [n] (reorder_to_original.npy) and
 * centroids [nlist x dim].  Replaces IVFIndex::IVFIndex's loaders (IVFIndex.cpp:154-267). */
VSB_API int vs_ivf_create(vs_ivf_t** out, const float* vectors_list_order, int64_t n, int dim,
                          const int32_t* cluster_offsets, int nlist, const int32_t* position_to_id,
                          const float* centroids, int device);
/* Index from the reference's on-disk directory (ivf_config.json + .npy files); both the
 * "reordered" and the scattered (cluster_indices.npy + vectors.npy|vectors.bin) layouts; centroids come from
 * centroids.npy (the reference loads the same numbers as a QNN context binary, centroids.bin). */
VSB_API int vs_ivf_open(vs_ivf_t** out, const char* index_dir, int device);
VSB_API int vs_ivf_destroy(vs_ivf_t* h);
VSB_API int64_t vs_ivf_num_vectors(const vs_ivf_t* h);   /* IVFIndex::getNumVectors  (IVFIndex.h:51) */
VSB_API int vs_ivf_num_clusters(const vs_ivf_t* h);      /* IVFIndex::getNumClusters (IVFIndex.h:52) */
VSB_API int vs_ivf_dim(const vs_ivf_t* h);               /* IVFIndex::getDim         (IVFIndex.h:53) */
VSB_API float vs_ivf_avg_cluster_size(const vs_ivf_t* h);/* IVFIndex::getAvgClusterSize (IVFIndex.h:54) */

/* IVFIndex::searchBatch (IVFIndex.h:45-48, IVFIndex.cpp:640-859): coarse query x centroid scores, top-nprobe
 * lists (nprobe clamped to nlist), fine scan, top-k by inner product.  out_ids[nq x k] original ids (-1 padded),
 * out_scores[nq x k] descending (-inf padded), out_counts[nq] = min(k, candidates) (may be NULL),
 * *total_candidates = rows scanned over the whole batch (the reference's return value). k <= 32. */
VSB_API int vs_ivf_search(vs_ivf_t* h, const float* queries, int64_t nq, int k, int nprobe, int32_t* out_ids,
                          float* out_scores, int32_t* out_counts, uint64_t* total_candidates);
VSB_API int vs_ivf_search_dev(vs_ivf_t* h, const float* queries_dev, int64_t nq, int k, int nprobe,
                              int32_t* out_ids_dev, float* out_scores_dev, int32_t* out_counts_dev, void* stream);
/* Raw coarse scores [nq x nlist] (what QnnRunner::getRawOutputBuffer holds after executeBatchRaw on the float
 * centroid model, IVFIndex.cpp:657-663). */
VSB_API int vs_ivf_coarse_scores(vs_ivf_t* h, const float* queries, int64_t nq, float* out_scores);
VSB_API int vs_ivf_set_profile(vs_ivf_t* h, int enable);       /* CUDA-event timing of the list-scan kernel */
VSB_API int vs_ivf_last_kernel_ms(vs_ivf_t* h, float* ms);

/* build_ivf_index (qidk_ivf/prepare/create_ivf_model.py:86-175, create_ivf_model_reordered.py:82-177): Lloyd
 * k-means (L2) — assignment = the fused tensor-core distance + arg-min kernel, update = per-list means — then the
 * inverted lists and the index directory (same file names, dtypes and ivf_config.json keys; no ONNX/QNN
 * artefacts).  nlist is adjusted like the reference (nlist > n/10 -> max(16, n/100)).  init_centroids
 * [nlist x dim] may be NULL (seeded sample of rows); with init_centroids and max_iter = 0 the given centroids are
 * used unchanged (parity mode). reordered != 0 writes the list-contiguous layout. */
VSB_API int vs_ivf_build(const float* base, int64_t n, int dim, int nlist, int max_iter, uint64_t seed,
                         const char* out_dir, int reordered, int device, const float* init_centroids,
                         int* out_nlist, int* out_iters, double* out_inertia);

/* ---------------------------------------------------------------------------------------------- */
/* INT8 brute force (the QNN u8 MatMul path)                                                       */
/* ---------------------------------------------------------------------------------------------- */
typedef struct vs_int8 vs_int8_t;

/* Replaces QnnRunner::QnnRunner (qidk_bruteforce/.../QnnRunner.h:20, QnnRunner.cpp:57-83): where the reference
 * loads a QNN context binary whose MatMul weights are the u8-quantised documents (create_model.py:57-87 +
 * convert_to_qnn.sh), this quantises base[n x dim] (fp32, raw SIFT values) once with
 *   w = sat_u8(trunc(x * (1/w_scale) + 0.5))      (w_scale <= 0: max(base)/255, the converter's min/max rule)
 * and keeps it resident.  in_scale / out_scale are the u8 encodings of the graph input and output
 * (QnnRunner.cpp:70-71 defaults 0.6627451 and 1013.4312; offsets 0).  Scores are
 *   sat_u8(floor(fl(fl(acc) * m) + 0.5)),  acc = sum q_u8*w_u8 (int32),  m = fl(fl(in_scale*w_scale)/out_scale)
 * — the requantisation rule is defined by this project (the HTP's is proprietary), see DESIGN.md. */
VSB_API int vs_int8_create(vs_int8_t** out, const float* base, int64_t n, int dim, float in_scale, float w_scale,
                           float out_scale, int device, int64_t id_base);
VSB_API int vs_int8_create_dev(vs_int8_t** out, const float* base_dev, int64_t n, int dim, float in_scale, float w_scale,
                               float out_scale, int device, int64_t id_base);
VSB_API int vs_int8_destroy(vs_int8_t* h);
VSB_API int64_t vs_int8_num_docs(const vs_int8_t* h);      /* QnnRunner::getNumDocs     (QnnRunner.h:49) */
VSB_API int vs_int8_dim(const vs_int8_t* h);               /* QnnRunner::getDim         (QnnRunner.h:48) */
VSB_API float vs_int8_output_scale(const vs_int8_t* h);    /* QnnRunner::getOutputScale (QnnRunner.h:51) */
VSB_API int vs_int8_scales(const vs_int8_t* h, float* in_scale, float* w_scale, float* out_scale, float* multiplier);
/* executeBatchRaw + find_top_k_int8 fused (QnnRunner.cpp:597-638, main.cpp:36-71): fp32 queries are quantised with
 * quantize_buffer_neon's rule (QnnRunner.cpp:13-55); out_ids[nq x k], out_scores[nq x k] raw u8, ordered
 * (score descending, id ascending); the printed score of the reference is out_scores * out_scale (main.cpp:185).
 * Any nq (the reference's fixed model batch + zero padding, main.cpp:206-211, is not needed). k <= 32. */
VSB_API int vs_int8_search(vs_int8_t* h, const float* queries, int64_t nq, int k, int32_t* out_ids, uint8_t* out_scores);
VSB_API int vs_int8_search_dev(vs_int8_t* h, const float* queries_dev, int64_t nq, int k, int32_t* out_ids_dev,
                               uint8_t* out_scores_dev, void* stream);
/* The whole raw u8 score matrix [nq x n] (QnnRunner::getRawOutputBuffer after executeBatchRaw, QnnRunner.h:41);
 * for tests and small n only (nq*n <= 4e9). */
VSB_API int vs_int8_scores_raw(vs_int8_t* h, const float* queries, int64_t nq, uint8_t* out_scores);
/* quantize_buffer_neon on the device for host buffers (QnnRunner.cpp:13-55), scale as in the reference's callers
 * (inv_scale = 1.0f / scale in fp32). */
VSB_API int vs_int8_quantize(const float* src, int64_t count, float scale, uint8_t* dst, int device);
VSB_API int vs_int8_set_profile(vs_int8_t* h, int enable);     /* CUDA-event timing of the fused INT8 kernel */
VSB_API int vs_int8_last_kernel_ms(vs_int8_t* h, float* ms);

/* Seeded synthetic SIFT-shaped rows generated on the device (bit-identical to the numpy generator in
 * hai-25-rag-on-edge_b200/synth.py). law: 0 "sift", 1 "cont", 2 "mix". */
VSB_API int vs_synth_fill_dev(float* out_dev, int64_t row0, int64_t nrows, int dim, int law, uint64_t seed,
                      uint64_t centre_seed, void* stream);

/* Pinned host memory helpers (so that host<->device copies inside the timed span run at full PCIe rate). */
VSB_API int vs_host_alloc(void** out, size_t bytes);
VSB_API int vs_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* VSB200_H */
