// IVF two-stage search kernels (replaces IVFIndex::search / searchBatch, qidk_ivf/android/app/main/jni/IVFIndex.cpp:572-859).
//
//   ivf_coarse_kernel    S[b][c] = q_b . centroid_c  — the query x centroid MatMul the reference runs on the QNN HTP
//                        (IVFIndex.cpp:653-708; graph = create_ivf_model.py:45-64)
//   ivf_probe_kernel     top-nprobe clusters per query, largest score first (IVFIndex.cpp:598-599, 711-712)
//   ivf_scan_kernel (K6) fine scan of the probed inverted lists with a per-query top-k
//                        (computeDotProductsContiguous :269-358 + the heap of :733-779)
//
// Metric is inner product, largest = best (the reference ranks by raw dot products).  Every dot product uses the
// reference's fp32 order — four accumulators by (d mod 4) advanced with FMA over d = 0,4,8,..., combined as
// (l0+l1)+(l2+l3) (NEON vmlaq_f32 / vaddvq_f32) — so scores and probe sets are bit-identical to the CPU
// restatement and result order is the canonical (score desc, original id asc).
//
// K6 layout: vectors are list-contiguous in HBM ([N][128] fp32).  One CTA per query walks its nprobe lists in
// chunks of 64 rows; each chunk is four TMA boxes (64 rows x 128 B, SWIZZLE_128B) into a 3-stage shared-memory
// ring, so global reads are full 128-B lines issued by the TMA engine while 64 threads (one row each) read their
// row back with conflict-free LDS.128 (the swizzle XORs the 16-B unit index with row%8) and the query as a
// shared-memory broadcast.  HBM-bound: the per-row math (128 FMA) is ~20x below what the streaming rate needs.
#include <cuda.h>

#include <cstdlib>

#include "kernels.cuh"
#include "vsb_common.cuh"

namespace vsb {

// ------------------------------------------------------------------------------------------------
// coarse scores
// ------------------------------------------------------------------------------------------------
constexpr int CO_QT = 16;    // queries per block (each block re-reads its 128 centroids: fewer, fatter blocks)
constexpr int CO_CT = 128;   // centroids per block (one per thread)

__global__ void __launch_bounds__(CO_CT) ivf_coarse_kernel(const float* __restrict__ q, int64_t nq,
                                                           const float* __restrict__ cent, int nlist,
                                                           float* __restrict__ scores) {
    __shared__ float4 sq[CO_QT][32];
    const int64_t q0 = (int64_t)blockIdx.y * CO_QT;
    const int c = blockIdx.x * CO_CT + threadIdx.x;
    for (int i = threadIdx.x; i < CO_QT * 32; i += CO_CT) {
        const int qi = i >> 5;
        sq[qi][i & 31] = (q0 + qi < nq) ? __ldg(reinterpret_cast<const float4*>(q + (q0 + qi) * 128) + (i & 31))
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    if (c >= nlist) return;
    float l[CO_QT][4];
#pragma unroll
    for (int qi = 0; qi < CO_QT; ++qi) l[qi][0] = l[qi][1] = l[qi][2] = l[qi][3] = 0.f;
    const float4* cr = reinterpret_cast<const float4*>(cent + (size_t)c * 128);
#pragma unroll 4
    for (int j = 0; j < 32; ++j) {
        const float4 x = __ldg(cr + j);
#pragma unroll
        for (int qi = 0; qi < CO_QT; ++qi) {
            const float4 v = sq[qi][j];
            l[qi][0] = fmaf(v.x, x.x, l[qi][0]);
            l[qi][1] = fmaf(v.y, x.y, l[qi][1]);
            l[qi][2] = fmaf(v.z, x.z, l[qi][2]);
            l[qi][3] = fmaf(v.w, x.w, l[qi][3]);
        }
    }
#pragma unroll
    for (int qi = 0; qi < CO_QT; ++qi)
        if (q0 + qi < nq)
            scores[(q0 + qi) * nlist + c] = __fadd_rn(__fadd_rn(l[qi][0], l[qi][1]), __fadd_rn(l[qi][2], l[qi][3]));
}

// The same scores for batches: a block = 128 centroids x 32 queries.  The centroid tile is staged in shared memory TRANSPOSED
// (float4 j of centroid c at [j][c], pitch 129 so that both the staging stores and the compute loads are conflict-free) with
// coalesced global loads, the queries as float4 broadcasts; thread (c, half) keeps 16 queries x 4 accumulators in registers:
// 17 shared-memory loads per 64 FMAs instead of one global load per 64.  Every score is the same chain of operations as above.
constexpr int CB_CT = 128, CB_QT = 32, CB_PITCH = 129;
constexpr int CB_SMEM = (32 * CB_PITCH + CB_QT * 32) * 16;  // 82 KB: two blocks per SM

__global__ void __launch_bounds__(256, 2) ivf_coarse_tiled_kernel(const float* __restrict__ q, int64_t nq,
                                                                  const float* __restrict__ cent, int nlist,
                                                                  float* __restrict__ scores) {
    extern __shared__ float4 cb_smem[];
    float4* sC = cb_smem;                   // [32][CB_PITCH]
    float4* sQ = cb_smem + 32 * CB_PITCH;   // [CB_QT][32]
    const int c0 = blockIdx.x * CB_CT;
    const int64_t q0 = (int64_t)blockIdx.y * CB_QT;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = warp; r < CB_CT; r += 8) {  // one centroid row per warp and step: 512 coalesced bytes
        const int c = c0 + r;
        sC[lane * CB_PITCH + r] = c < nlist ? __ldg(reinterpret_cast<const float4*>(cent + (size_t)c * 128) + lane)
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int r = warp; r < CB_QT; r += 8)
        sQ[r * 32 + lane] = (q0 + r < nq) ? __ldg(reinterpret_cast<const float4*>(q + (q0 + r) * 128) + lane)
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    const int cl = threadIdx.x & (CB_CT - 1);
    const int half = threadIdx.x >> 7;  // queries half * 16 .. + 15
    float l[16][4];
#pragma unroll
    for (int qi = 0; qi < 16; ++qi) l[qi][0] = l[qi][1] = l[qi][2] = l[qi][3] = 0.f;
#pragma unroll 2
    for (int j = 0; j < 32; ++j) {
        const float4 x = sC[j * CB_PITCH + cl];
#pragma unroll
        for (int qi = 0; qi < 16; ++qi) {
            const float4 v = sQ[(half * 16 + qi) * 32 + j];
            l[qi][0] = fmaf(v.x, x.x, l[qi][0]);
            l[qi][1] = fmaf(v.y, x.y, l[qi][1]);
            l[qi][2] = fmaf(v.z, x.z, l[qi][2]);
            l[qi][3] = fmaf(v.w, x.w, l[qi][3]);
        }
    }
    const int c = c0 + cl;
    if (c >= nlist) return;
#pragma unroll
    for (int qi = 0; qi < 16; ++qi) {
        const int64_t qq = q0 + half * 16 + qi;
        if (qq < nq) scores[qq * nlist + c] = __fadd_rn(__fadd_rn(l[qi][0], l[qi][1]), __fadd_rn(l[qi][2], l[qi][3]));
    }
}

// ------------------------------------------------------------------------------------------------
// probe selection: one warp per query, nprobe rounds of arg-max over the score row held in shared memory
// (canonical order: score desc, cluster id asc)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) ivf_probe_kernel(const float* __restrict__ scores, int64_t nq, int nlist, int nprobe,
                                                        int32_t* __restrict__ probes) {
    extern __shared__ float s_sc[];  // [4][nlist]
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int64_t q = (int64_t)blockIdx.x * 4 + wib;
    if (q >= nq) return;
    float* sc = s_sc + (size_t)wib * nlist;
    const float NINF = __int_as_float(0xff800000);
    // NaN scores (a non-finite query or centroid; undefined in the reference's nth_element) rank last, like -inf: the
    // consumed-entry marker below is NaN, so a NaN input must not reach the selection
    for (int c = lane; c < nlist; c += 32) {
        const float v = scores[q * nlist + c];
        sc[c] = v == v ? v : NINF;
    }
    __syncwarp();
    for (int r = 0; r < nprobe; ++r) {
        float bv = NINF;
        int bc = 0x7fffffff;  // none yet
        for (int c = lane; c < nlist; c += 32) {
            const float v = sc[c];  // consumed entries are NaN and fail every comparison
            if (v > bv || (bc == 0x7fffffff && v == v)) {
                bv = v;
                bc = c;  // c ascends: the first of an equal-score run is kept
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
            if (ov > bv || (ov == bv && oc < bc)) {
                bv = ov;
                bc = oc;
            }
        }
        if (lane == 0) {
            probes[q * nprobe + r] = bc;  // always a valid list: nprobe <= nlist unconsumed non-NaN entries exist
            sc[bc] = __int_as_float(0x7fc00000);  // NaN: never compares greater/equal again
        }
        __syncwarp();
    }
}

// The same selection for nprobe <= 32 and nlist <= 1024 with ~5x fewer instructions: the score row sits in registers (NPER per
// lane, cluster = lane + 32 j) as order-reversed integer keys, a radix descent from the highest differing bit finds the
// bucket that holds the nprobe-th best score (stopping as soon as the whole bucket is needed), everything before the bucket
// and the first clusters of the bucket are compacted through shared memory, and a rank-by-counting pass puts the nprobe
// clusters in the canonical order (score desc, cluster id asc).  Same output as ivf_probe_kernel, bit for bit.
template <int NPER>
__global__ void __launch_bounds__(128) ivf_probe_radix_kernel(const float* __restrict__ scores, int64_t nq, int nlist, int nprobe,
                                                              int32_t* __restrict__ probes) {
    __shared__ uint32_t s_u[4][32];
    __shared__ int32_t s_c[4][32];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int64_t q = (int64_t)blockIdx.x * 4 + wib;
    if (q >= nq) return;
    uint32_t u[NPER];  // smaller = better; 0xffffffff = no such cluster
    uint32_t lo = 0xffffffffu, hi = 0u;
#pragma unroll
    for (int j = 0; j < NPER; ++j) {
        const int c = lane + 32 * j;
        u[j] = 0xffffffffu;
        if (c < nlist) {
            float v = scores[q * nlist + c];
            v = v == v ? v : __int_as_float(0xff800000);  // NaN ranks last, like -inf
            const uint32_t b = __float_as_uint(v);
            u[j] = (b & 0x80000000u) ? b : ~(b | 0x80000000u);  // = ~(monotone key): descending scores ascend
            lo = min(lo, u[j]);
            hi = max(hi, u[j]);
        }
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    int need = nprobe, size = nlist;
    uint32_t prefix = lo;
    int bit = -1;
    if (lo != hi) {
        bit = 31 - __clz(lo ^ hi);
        prefix = bit == 31 ? 0u : (lo & ~((2u << bit) - 1u));
        while (bit >= 0 && need < size) {
            const uint32_t m = (bit == 31 ? 0u : ~((2u << bit) - 1u)) | (1u << bit);  // the bucket's bits and this one
            int c0 = 0;
#pragma unroll
            for (int j = 0; j < NPER; ++j) c0 += (((u[j] ^ prefix) & m) == 0u) ? 1 : 0;
            c0 = __reduce_add_sync(0xffffffffu, c0);
            if (need <= c0) {
                size = c0;
            } else {
                need -= c0;
                size -= c0;
                prefix |= 1u << bit;
            }
            --bit;
        }
    }
    // clusters with u < prefix are all selected (nprobe - need of them); of the bucket [prefix, prefix | low] the first `need`
    const uint32_t low = bit >= 0 ? ((2u << bit) - 1u) : 0u;
    const unsigned lt_mask = (1u << lane) - 1u;
    int base_l = 0, base_b = nprobe - need;
#pragma unroll
    for (int j = 0; j < NPER; ++j) {
        const bool lt = u[j] < prefix;
        const bool inb = !lt && (u[j] - prefix) <= low && u[j] != 0xffffffffu;
        const unsigned bl = __ballot_sync(0xffffffffu, lt);
        const unsigned bb = __ballot_sync(0xffffffffu, inb);
        if (lt) {
            const int pos = base_l + __popc(bl & lt_mask);
            s_u[wib][pos] = u[j];
            s_c[wib][pos] = lane + 32 * j;
        }
        if (inb) {
            const int pos = base_b + __popc(bb & lt_mask);
            if (pos < nprobe) {
                s_u[wib][pos] = u[j];
                s_c[wib][pos] = lane + 32 * j;
            }
        }
        base_l += __popc(bl);
        base_b += __popc(bb);
    }
    __syncwarp();
    if (lane < nprobe) {
        const uint32_t ur = s_u[wib][lane];
        const int32_t cr = s_c[wib][lane];
        int rank = 0;
        for (int t = 0; t < nprobe; ++t) {
            const uint32_t ut = s_u[wib][t];
            const int32_t ct = s_c[wib][t];
            rank += (ut < ur || (ut == ur && ct < cr)) ? 1 : 0;
        }
        probes[q * nprobe + rank] = cr;
    }
}

// ------------------------------------------------------------------------------------------------
// K6: list scan
// ------------------------------------------------------------------------------------------------
constexpr int IV_ROWS = 64;                       // rows per chunk = threads per CTA
constexpr int IV_STAGES = 3;
constexpr int IV_KB_BYTES = IV_ROWS * 128;        // one k-block box: 64 rows x 128 B
constexpr int IV_STAGE_BYTES = 4 * IV_KB_BYTES;   // 32 KB
constexpr int IV_SMEM = IV_STAGES * IV_STAGE_BYTES + 3072;  // ring + query/barriers/merge scratch + 1024-B alignment slack

struct IvfScanParams {
    const float* q;          // [nq][128]
    const int32_t* probes;   // [nq][nprobe]
    const int32_t* offsets;  // [nlist+1]
    const int32_t* id_map;   // [N] list position -> original id
    int64_t nq;
    int nprobe;
    int k;
    float* out_scores;       // [nq][k] descending, -inf padded
    int32_t* out_ids;        // [nq][k] original ids, -1 padded
    int32_t* out_counts;     // [nq] min(k, candidates)
    unsigned long long* total_candidates;
};

template <int KTOP>
__global__ void __launch_bounds__(IV_ROWS, 2) ivf_scan_kernel(const __grid_constant__ CUtensorMap tmV, const IvfScanParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* ring = smem;                                             // [IV_STAGES][4 k-blocks][64 rows][128 B]
    float4* sq = (float4*)(smem + IV_STAGES * IV_STAGE_BYTES);        // query, 512 B
    uint64_t* full = (uint64_t*)((uint8_t*)sq + 512);                 // [IV_STAGES]
    float* s_key = (float*)(full + IV_STAGES);                        // [2][KTOP]
    int32_t* s_id = (int32_t*)(s_key + 2 * KTOP);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int64_t qi = blockIdx.x;

    if (tid < 32) sq[tid] = __ldg(reinterpret_cast<const float4*>(p.q + qi * 128) + tid);
    if (tid == 0) {
        for (int s = 0; s < IV_STAGES; ++s) mbar_init(&full[s], 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmV);
    }
    __syncthreads();

    // chunk enumeration state, replicated in every thread: (probe index, row offset inside that list)
    const int32_t* pr = p.probes + qi * p.nprobe;
    // loader cursor (thread 0 only) and consumer cursor (all threads) walk the same sequence
    int ld_probe = 0, ld_row = 0, ld_start = 0, ld_len = 0;
    int cs_probe = 0, cs_row = 0, cs_start = 0, cs_len = 0;
    auto next_list = [&](int& probe, int& start, int& len) {
        // advance to the next non-empty probed list; len = 0 when exhausted
        len = 0;
        while (probe < p.nprobe) {
            const int c = __ldg(pr + probe);
            start = __ldg(p.offsets + c);
            len = __ldg(p.offsets + c + 1) - start;
            if (len > 0) return;
            ++probe;
        }
    };
    next_list(cs_probe, cs_start, cs_len);
    ld_probe = cs_probe;
    ld_start = cs_start;
    ld_len = cs_len;

    auto issue = [&](int stage) {  // thread 0: load the chunk at the loader cursor, advance the cursor
        if (ld_len == 0) return;
        uint8_t* dst = ring + stage * IV_STAGE_BYTES;
        mbar_expect_tx(&full[stage], (uint32_t)IV_STAGE_BYTES);
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) tma_load_2d(dst + kb * IV_KB_BYTES, &tmV, &full[stage], kb * 32, ld_start + ld_row);
        ld_row += IV_ROWS;
        if (ld_row >= ld_len) {
            ++ld_probe;
            ld_row = 0;
            next_list(ld_probe, ld_start, ld_len);
        }
    };
    if (tid == 0) {
        for (int s = 0; s < IV_STAGES - 1; ++s) issue(s);
    }

    RegTopK<KTOP> top;  // keys = -score (smallest first), ids = original ids
    top.init();
    unsigned long long cand = 0;
    int stage = 0;
    uint32_t phase = 0;
    int ld_stage = IV_STAGES - 1;
    while (cs_len > 0) {
        mbar_wait(&full[stage], phase);
        const int rows_here = min(IV_ROWS, cs_len - cs_row);
        if (tid == 0) cand += (unsigned long long)rows_here;
        if (tid < rows_here) {
            const uint8_t* base = ring + stage * IV_STAGE_BYTES + tid * 128;
            const int sw = tid & 7;
            float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float4 qv = sq[j];
                const float4 xv = *reinterpret_cast<const float4*>(base + (j >> 3) * IV_KB_BYTES + (((j & 7) ^ sw) << 4));
                l0 = fmaf(qv.x, xv.x, l0);
                l1 = fmaf(qv.y, xv.y, l1);
                l2 = fmaf(qv.z, xv.z, l2);
                l3 = fmaf(qv.w, xv.w, l3);
            }
            const float key = -__fadd_rn(__fadd_rn(l0, l1), __fadd_rn(l2, l3));
            if (key <= top.threshold()) top.insert_any(key, __ldg(p.id_map + cs_start + cs_row + tid));
        }
        // advance the consumer cursor
        cs_row += IV_ROWS;
        if (cs_row >= cs_len) {
            ++cs_probe;
            cs_row = 0;
            next_list(cs_probe, cs_start, cs_len);
        }
        __syncthreads();  // everyone is done with `stage` ... and with the stage consumed one iteration ago
        if (tid == 0) {
            issue(ld_stage);
        }
        if (++ld_stage == IV_STAGES) ld_stage = 0;
        if (++stage == IV_STAGES) { stage = 0; phase ^= 1; }
    }

    // merge the 64 per-thread lists
    warp_merge_lists<KTOP>(top, KTOP, s_key + warp * KTOP, s_id + warp * KTOP);
    __syncthreads();
    if (warp == 0) {
        RegTopK<KTOP> mine;
        mine.init();
        if (lane < 2) {
#pragma unroll
            for (int i = 0; i < KTOP; ++i) {
                mine.key[i] = s_key[lane * KTOP + i];
                mine.id[i] = s_id[lane * KTOP + i];
            }
        }
        __syncwarp();
        warp_merge_lists<KTOP>(mine, KTOP, s_key, s_id);  // lane 0 writes; safe: all lanes loaded above
        __syncwarp();
        const float NINF = __int_as_float(0xff800000);
        int nvalid = 0;
        for (int i = lane; i < p.k; i += 32) {
            const int32_t id = i < KTOP ? s_id[i] : -1;
            p.out_ids[qi * p.k + i] = id;
            p.out_scores[qi * p.k + i] = id >= 0 ? -s_key[i] : NINF;
            nvalid += id >= 0;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nvalid += __shfl_xor_sync(0xffffffffu, nvalid, o);
        if (lane == 0) {
            p.out_counts[qi] = nvalid;
            atomicAdd(p.total_candidates, cand);
        }
    }
}

int launch_ivf_coarse(const float* q, int64_t nq, const float* cent, int nlist, float* scores, cudaStream_t st) {
    if (nq <= 0) return VS_OK;
    if (nq >= 256) {  // batches: shared-memory tiles
        dim3 gridb((unsigned)((nlist + CB_CT - 1) / CB_CT), (unsigned)ceil_div64(nq, CB_QT));
        if (gridb.y > 65535) return fail(VS_ERR_UNSUPPORTED, "coarse: too many queries in one call");
        VSB_CUDA(cudaFuncSetAttribute(ivf_coarse_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CB_SMEM));  // per device
        ivf_coarse_tiled_kernel<<<gridb, 256, CB_SMEM, st>>>(q, nq, cent, nlist, scores);
        VSB_CUDA(cudaGetLastError());
        return VS_OK;
    }
    dim3 grid((unsigned)((nlist + CO_CT - 1) / CO_CT), (unsigned)ceil_div64(nq, CO_QT));
    if (grid.y > 65535) return fail(VS_ERR_UNSUPPORTED, "coarse: too many queries in one call");
    ivf_coarse_kernel<<<grid, CO_CT, 0, st>>>(q, nq, cent, nlist, scores);
    VSB_CUDA(cudaGetLastError());
    return VS_OK;
}

int launch_ivf_probes(const float* scores, int64_t nq, int nlist, int nprobe, int32_t* probes, cudaStream_t st) {
    if (nq <= 0) return VS_OK;
    if (nprobe <= 32 && nlist <= 1024 && !getenv("VSB_IVF_PROBE_ROUNDS")) {
        const unsigned blocks = (unsigned)ceil_div64(nq, 4);
        const int nper = (nlist + 31) / 32;
        if (nper <= 2)
            ivf_probe_radix_kernel<2><<<blocks, 128, 0, st>>>(scores, nq, nlist, nprobe, probes);
        else if (nper <= 8)
            ivf_probe_radix_kernel<8><<<blocks, 128, 0, st>>>(scores, nq, nlist, nprobe, probes);
        else if (nper <= 16)
            ivf_probe_radix_kernel<16><<<blocks, 128, 0, st>>>(scores, nq, nlist, nprobe, probes);
        else
            ivf_probe_radix_kernel<32><<<blocks, 128, 0, st>>>(scores, nq, nlist, nprobe, probes);
        VSB_CUDA(cudaGetLastError());
        return VS_OK;
    }
    const size_t smem = (size_t)4 * nlist * sizeof(float);
    if (smem > 48 * 1024) return fail(VS_ERR_UNSUPPORTED, "probe selection: nlist > 3072 not implemented");
    ivf_probe_kernel<<<(unsigned)ceil_div64(nq, 4), 128, smem, st>>>(scores, nq, nlist, nprobe, probes);
    VSB_CUDA(cudaGetLastError());
    return VS_OK;
}

int ivf_scan_rows_per_chunk() { return IV_ROWS; }

int ivf_set_attributes() {
    VSB_CUDA(cudaFuncSetAttribute(ivf_scan_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, IV_SMEM));
    VSB_CUDA(cudaFuncSetAttribute(ivf_scan_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, IV_SMEM));
    VSB_CUDA(cudaFuncSetAttribute(ivf_scan_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, IV_SMEM));
    VSB_CUDA(cudaFuncSetAttribute(ivf_scan_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, IV_SMEM));
    VSB_CUDA(cudaFuncSetAttribute(ivf_scan_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, IV_SMEM));
    return VS_OK;
}

int launch_ivf_scan(const CUtensorMap& tmV, const float* q, const int32_t* probes, const int32_t* offsets,
                    const int32_t* id_map, int64_t nq, int nprobe, int k, float* out_scores, int32_t* out_ids,
                    int32_t* out_counts, unsigned long long* total, cudaStream_t st) {
    if (nq <= 0) return VS_OK;
    IvfScanParams p{q, probes, offsets, id_map, nq, nprobe, k, out_scores, out_ids, out_counts, total};
    const int ktop = round_up_ktop(k);
    const unsigned grid = (unsigned)nq;
    switch (ktop) {
        case 1: ivf_scan_kernel<1><<<grid, IV_ROWS, IV_SMEM, st>>>(tmV, p); break;
        case 5: ivf_scan_kernel<5><<<grid, IV_ROWS, IV_SMEM, st>>>(tmV, p); break;
        case 10: ivf_scan_kernel<10><<<grid, IV_ROWS, IV_SMEM, st>>>(tmV, p); break;
        case 16: ivf_scan_kernel<16><<<grid, IV_ROWS, IV_SMEM, st>>>(tmV, p); break;
        case 32: ivf_scan_kernel<32><<<grid, IV_ROWS, IV_SMEM, st>>>(tmV, p); break;
        default: return fail(VS_ERR_UNSUPPORTED, "IVF search: k > 32 is not implemented");
    }
    VSB_CUDA(cudaGetLastError());
    return VS_OK;
}

}  // namespace vsb
