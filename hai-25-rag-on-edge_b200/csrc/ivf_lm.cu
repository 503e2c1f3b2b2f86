// K8 — LIST-MAJOR IVF fine scan for large query batches.
//
// The reference scans, per query, the nprobe inverted lists it selected (IVFIndex.cpp:715-779): at 10 000 queries
// x nprobe 32 over 1024 lists every list is scanned ~312 times.  K6 (ivf.cu) keeps that per-query formulation and is
// bound by streaming the lists (out of L2 for the most part).  Here the (query, list) pairs are grouped BY LIST:
// a CTA takes one list and a tile of 32 of the queries that probe it, keeps the queries and a 128-row chunk of the
// list in shared memory and computes the 32 x 128 scores from registers (4 queries x 4 rows per thread), so a list
// row is read from shared memory once per FOUR queries and from L2/HBM once per 32 — the scan becomes FFMA-bound.
//
// Arithmetic is unchanged: every score is its own chain — four accumulators by (d mod 4) advanced with FMA over
// d = 0,4,8,..., combined as (l0+l1)+(l2+l3) (computeDotProductsContiguous, IVFIndex.cpp:278-357) — so scores are
// bit-identical to K6 and to the CPU restatement, and the result order is the canonical (score desc, id asc).
//
// Selection: the warp that computed the 128 scores of a query (warp = group of 4 queries) folds them 32 at a time
// into that query's top-k, held across the lanes (lane i = entry i): ballot of "beats the k-th entry", then per
// qualifying row a ballot/popc position and a shuffle-up shift.  One sorted list per (query, probe) goes to global
// memory; the ordinary merge kernel (K3) combines the nprobe lists of a query.
#include <cuda.h>

#include "kernels.cuh"
#include "vsb_common.cuh"

namespace vsb {

constexpr int LM_QT = 32;        // queries per work item (8 warps x 4)
constexpr int LM_RT = 128;       // list rows per chunk (32 lanes x 4)
constexpr int LM_THREADS = 256;
constexpr int LM_V_BYTES = LM_RT * 512;   // one chunk of rows: 64 KB
// one chunk buffer per CTA: two CTAs share an SM (16 warps), one CTA's load overlaps the other's arithmetic
constexpr int LM_SMEM = LM_V_BYTES + LM_QT * 512 + LM_RT * 4 + LM_QT * 4 + 64;

// ---- pair grouping -------------------------------------------------------------------------------------------------
// per list: number of (query, probe) pairs; per query: candidates = sum of the probed lists' lengths
__global__ void lm_count_kernel(const int32_t* __restrict__ probes, int64_t n_pairs, int nprobe, const int32_t* __restrict__ offsets,
                                int32_t* __restrict__ list_cnt, int32_t* __restrict__ q_cand) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = probes[i];
        atomicAdd(list_cnt + c, 1);
        atomicAdd(q_cand + i / nprobe, offsets[c + 1] - offsets[c]);
    }
}

// one block: exclusive scan of the pair counts (pair_start), of the item counts (item_start) and the item table
// `order`: the lists by descending length — the items of long lists come first in the work queue (no long tail)
__global__ void __launch_bounds__(1024) lm_scan_kernel(const int32_t* __restrict__ list_cnt, int nlist,
                                                       const int32_t* __restrict__ order,
                                                       const int32_t* __restrict__ offsets, int32_t* __restrict__ pair_start,
                                                       int32_t* __restrict__ cursor, int4* __restrict__ items,
                                                       int32_t* __restrict__ n_items, int qt /* queries per work item */) {
    __shared__ int s_pairs[1024], s_items[1024];
    const int t = threadIdx.x;
    const int per = (nlist + 1023) / 1024;
    const int c0 = t * per, c1 = min(c0 + per, nlist);
    int np = 0, ni = 0;
    for (int i = c0; i < c1; ++i) {
        const int c = order[i];
        np += list_cnt[c];
        ni += (list_cnt[c] + qt - 1) / qt;
    }
    s_pairs[t] = np;
    s_items[t] = ni;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {  // Hillis-Steele inclusive scan
        const int a = t >= o ? s_pairs[t - o] : 0, b = t >= o ? s_items[t - o] : 0;
        __syncthreads();
        s_pairs[t] += a;
        s_items[t] += b;
        __syncthreads();
    }
    int pp = s_pairs[t] - np, ii = s_items[t] - ni;
    for (int i = c0; i < c1; ++i) {
        const int c = order[i];
        pair_start[c] = pp;
        cursor[c] = pp;
        const int cnt = list_cnt[c];
        const int r0 = offsets[c], rl = offsets[c + 1] - r0;
        for (int q0 = 0; q0 < cnt; q0 += qt)  // {first pair, queries, first row, rows}: one 16-byte load per item
            items[ii++] = make_int4(pp + q0, min(qt, cnt - q0), r0, rl);
        pp += cnt;
    }
    if (t == 1023) *n_items = s_items[1023];
}

__global__ void lm_fill_kernel(const int32_t* __restrict__ probes, int64_t n_pairs, int32_t* __restrict__ cursor,
                               int32_t* __restrict__ pairs) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs; i += (int64_t)gridDim.x * blockDim.x)
        pairs[atomicAdd(cursor + probes[i], 1)] = (int32_t)i;  // i = query * nprobe + probe slot
}

__global__ void lm_counts_kernel(const int32_t* __restrict__ q_cand, int64_t nq, int k, int32_t* __restrict__ out_counts,
                                 unsigned long long* __restrict__ total) {
    unsigned long long s = 0;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (int64_t)gridDim.x * blockDim.x) {
        const int c = q_cand[q];
        out_counts[q] = c < k ? c : k;
        s += (unsigned long long)c;
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(total, s);
}

// ---- the scan ------------------------------------------------------------------------------------------------------
struct LmParams {
    const float* q;            // [nq][128]
    const float* vectors;      // [n][128] list-contiguous
    const int32_t* offsets;    // [nlist+1]
    const int32_t* id_map;     // [n]
    const int32_t* pairs;      // [nq*nprobe] grouped by list: query * nprobe + slot
    const int4* items;         // {first pair, queries (<= 32), first row, rows} per work item
    const int32_t* n_items;
    int32_t* next_item;        // work counter (zeroed by the host)
    int32_t* gthr;             // [nq] k-th best score found so far in ANY list of the query (ordered-int encoding, zeroed
                               // = "none"): a row that scores below it cannot be in the query's final top-k
    float* part_key;           // [nprobe][nq][KTOP]  key = -score, ascending
    int32_t* part_id;
    int64_t nq;
    int nprobe;
};

// order-preserving float -> int with 0 below every finite value ("none yet"): positive encodings for all floats
__device__ __forceinline__ int32_t lm_encode(float f) {
    const int32_t i = __float_as_int(f);
    const int32_t o = i >= 0 ? i : i ^ 0x7fffffff;      // monotone signed int
    return (o >> 1) + 0x40000001;                        // halve the range, shift above zero (ties may merge: a bound only)
}
__device__ __forceinline__ float lm_decode(int32_t e) {
    if (e == 0) return __int_as_float(0xff800000);
    const int32_t o = (e - 0x40000001) << 1;              // the low bit was dropped: decode to the smaller value
    return __int_as_float(o >= 0 ? o : o ^ 0x7fffffff);
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

// scores of 4 queries (shared-memory rows at qbase, +512 B each) x NJ*32 list rows (this lane's rows lane + 32*j at vbase,
// 16-byte units XOR-swizzled with the lane) -> out[u][j] (row lane + 32*j); the reference's summation order per score
template <int NJ>
__device__ __forceinline__ void lm_scores(uint32_t qbase, uint32_t vbase, int lane, float (&out)[4][4]) {
    float acc[4][NJ][4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[u][j][0] = acc[u][j][1] = acc[u][j][2] = acc[u][j][3] = 0.f;
#pragma unroll 2
    for (int c4 = 0; c4 < 32; ++c4) {
        float4 qv[4], xv[NJ];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(qv[u].x), "=f"(qv[u].y), "=f"(qv[u].z), "=f"(qv[u].w)
                         : "r"(qbase + (uint32_t)u * 512u + (uint32_t)c4 * 16u));
#pragma unroll
        for (int j = 0; j < NJ; ++j)
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(xv[j].x), "=f"(xv[j].y), "=f"(xv[j].z), "=f"(xv[j].w)
                         : "r"(vbase + (uint32_t)j * (32u * 512u) + (uint32_t)((c4 ^ lane) << 4)));
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                acc[u][j][0] = fmaf(qv[u].x, xv[j].x, acc[u][j][0]);
                acc[u][j][1] = fmaf(qv[u].y, xv[j].y, acc[u][j][1]);
                acc[u][j][2] = fmaf(qv[u].z, xv[j].z, acc[u][j][2]);
                acc[u][j][3] = fmaf(qv[u].w, xv[j].w, acc[u][j][3]);
            }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < NJ; ++j)
            out[u][j] = __fadd_rn(__fadd_rn(acc[u][j][0], acc[u][j][1]), __fadd_rn(acc[u][j][2], acc[u][j][3]));
}

template <int KTOP>
__global__ void __launch_bounds__(LM_THREADS, 2) ivf_lm_kernel(const LmParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    uint8_t* sV = smem;                                   // [128 rows][32 float4, slot c4 ^ (row & 31)]
    float* sQ = (float*)(smem + LM_V_BYTES);              // [32 queries][128]
    int32_t* sId = (int32_t*)(sQ + LM_QT * 128);          // [128] original ids of the chunk's rows
    int32_t* sPair = sId + LM_RT;                         // [32] pair index of each query slot, -1 = unused
    int32_t* sItem = sPair + LM_QT;                       // [1] index of the next work item, [4..7] its record

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;      // = query group: queries 4*warp .. 4*warp+3 of the tile
    const uint32_t sV_u = smem_u32(sV), sQ_u = smem_u32(sQ);
    const float NINF = __int_as_float(0xff800000);

    const int n_items = __ldg(p.n_items);
    if (tid == 0) {
        const int first = atomicAdd(p.next_item, 1);
        sItem[0] = first;
        if (first < n_items) *reinterpret_cast<int4*>(sItem + 4) = __ldg(p.items + first);
    }
    for (;;) {
        __syncthreads();  // the previous item's shared memory is no longer read; the next item's record is in place
        const int item = sItem[0];
        if (item >= n_items) break;
        const int4 rec = *reinterpret_cast<const int4*>(sItem + 4);
        const int pair0 = rec.x, nqt = rec.y, r_start = rec.z, r_len = rec.w;
        __syncthreads();  // everyone holds the record: thread 0 may overwrite it at the end of this item
        // claim the item after this one now: the atomic's round trip overlaps this item's loads and arithmetic
        int nxt = 0;
        if (tid == 0) nxt = atomicAdd(p.next_item, 1);
        const int n_chunks = (r_len + LM_RT - 1) / LM_RT;

        auto load_chunk = [&](int c) {  // rows [c*128, +128) of the list -> sV (swizzled), ids -> sId
            const int rows = min(LM_RT, r_len - c * LM_RT);
            const float4* src = reinterpret_cast<const float4*>(p.vectors + (size_t)(r_start + c * LM_RT) * 128);
            const uint32_t dst = sV_u;
#pragma unroll 4
            for (int e = tid; e < LM_RT * 32; e += LM_THREADS) {
                const int row = e >> 5, c4 = e & 31;
                if (row < rows) cp_async16(dst + (uint32_t)row * 512u + (uint32_t)((c4 ^ (row & 31)) << 4), src + e);
            }
            if (tid < LM_RT) sId[tid] = tid < rows ? __ldg(p.id_map + r_start + c * LM_RT + tid) : -1;
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        if (n_chunks > 0) load_chunk(0);
        // query tile
        if (tid < LM_QT) sPair[tid] = tid < nqt ? __ldg(p.pairs + pair0 + tid) : -1;
        __syncthreads();
        for (int e = tid; e < LM_QT * 32; e += LM_THREADS) {
            const int qi = e >> 5, c4 = e & 31;
            const int pr = sPair[qi];
            reinterpret_cast<float4*>(sQ)[e] = pr >= 0 ? __ldg(reinterpret_cast<const float4*>(p.q + (size_t)(pr / p.nprobe) * 128) + c4)
                                                       : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // per-warp state: the top-k lists of its 4 queries, entry i in lane i (score desc, id asc; id -1 = empty)
        float ls[4];
        int32_t li[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            ls[u] = NINF;
            li[u] = -1;
        }
        // bound from the other lists of the same queries (other CTAs, earlier items): most lists add nothing to a
        // query's final top-k, and with this bound their rows fail one compare instead of filling a fresh list
        float gs[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int pr = sPair[4 * warp + u];
            gs[u] = pr >= 0 ? lm_decode(__ldcg(p.gthr + pr / p.nprobe)) : NINF;
        }
        const bool warp_has_queries = 4 * warp < nqt;
        for (int c = 0; c < n_chunks; ++c) {
            if (c > 0) load_chunk(c);
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();  // chunk c (and, for c == 0, the query tile) is in shared memory
            if (warp_has_queries) {
            // ---- scores: queries 4*warp+u, rows lane + 32*j for the 32-row groups that exist in this chunk
            const uint32_t vbase = sV_u + (uint32_t)lane * 512u;
            const uint32_t qbase = sQ_u + (uint32_t)(4 * warp) * 512u;
            float scv[4][4];  // scores stay in registers: lane = row within the 32-row group j, as the selection needs them
            switch ((min(LM_RT, r_len - c * LM_RT) + 31) >> 5) {
                case 1: lm_scores<1>(qbase, vbase, lane, scv); break;
                case 2: lm_scores<2>(qbase, vbase, lane, scv); break;
                case 3: lm_scores<3>(qbase, vbase, lane, scv); break;
                default: lm_scores<4>(qbase, vbase, lane, scv); break;
            }
            // ---- selection
            const int rows = min(LM_RT, r_len - c * LM_RT);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (4 * warp + u >= nqt) break;  // warp-uniform
                float thr_s = __shfl_sync(0xffffffffu, ls[u], KTOP - 1);
                int32_t thr_i = __shfl_sync(0xffffffffu, li[u], KTOP - 1);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (32 * j >= rows) break;  // warp-uniform
                    const int r = 32 * j + lane;
                    const float sc = scv[u][j];
                    const int32_t id = sId[r];
                    // canonical order: larger score first, equal scores by smaller id; an empty k-th entry admits all
                    unsigned m = __ballot_sync(0xffffffffu, r < rows && sc >= gs[u] &&
                                                                (thr_i < 0 || sc > thr_s || (sc == thr_s && id < thr_i)));
                    while (m) {
                        const int jj = __ffs(m) - 1;
                        m &= m - 1;
                        const float x = __shfl_sync(0xffffffffu, sc, jj);
                        const int32_t xi = __shfl_sync(0xffffffffu, id, jj);
                        if (!(thr_i < 0 || x > thr_s || (x == thr_s && xi < thr_i))) continue;  // the bound moved
                        const bool before = li[u] >= 0 && (ls[u] > x || (ls[u] == x && li[u] < xi));
                        const int pos = __popc(__ballot_sync(0xffffffffu, before));
                        const float ns = __shfl_up_sync(0xffffffffu, ls[u], 1);
                        const int32_t ni = __shfl_up_sync(0xffffffffu, li[u], 1);
                        if (lane > pos) {
                            ls[u] = ns;
                            li[u] = ni;
                        } else if (lane == pos) {
                            ls[u] = x;
                            li[u] = xi;
                        }
                        thr_s = __shfl_sync(0xffffffffu, ls[u], KTOP - 1);
                        thr_i = __shfl_sync(0xffffffffu, li[u], KTOP - 1);
                    }
                }
                // a full list bounds the query's k-th best score: share it, and pick up what the others found
                if (thr_i >= 0 && thr_s > gs[u]) {
                    gs[u] = thr_s;
                    if (lane == 0) atomicMax(p.gthr + sPair[4 * warp + u] / p.nprobe, lm_encode(thr_s));
                }
            }
            }
            __syncthreads();  // everyone is done with sV / sId before the next load overwrites them
        }
        if (tid == 0) {
            sItem[0] = nxt;
            if (nxt < n_items) *reinterpret_cast<int4*>(sItem + 4) = __ldg(p.items + nxt);
        }
        // ---- one sorted list per (query, probe slot)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int pr = sPair[4 * warp + u];
            if (pr < 0) continue;
            const int64_t qi = pr / p.nprobe;
            const int slot = pr - (int)qi * p.nprobe;
            if (lane < KTOP) {
                const size_t o = ((size_t)slot * p.nq + qi) * KTOP + lane;
                p.part_key[o] = li[u] >= 0 ? -ls[u] : __int_as_float(0x7f800000);
                p.part_id[o] = li[u];
            }
        }
    }
}

int ivf_lm_set_attributes() {
    VSB_CUDA(cudaFuncSetAttribute(ivf_lm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, LM_SMEM + 128));
    VSB_CUDA(cudaFuncSetAttribute(ivf_lm_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, LM_SMEM + 128));
    VSB_CUDA(cudaFuncSetAttribute(ivf_lm_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, LM_SMEM + 128));
    VSB_CUDA(cudaFuncSetAttribute(ivf_lm_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, LM_SMEM + 128));
    VSB_CUDA(cudaFuncSetAttribute(ivf_lm_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, LM_SMEM + 128));
    return VS_OK;
}

size_t ivf_lm_workspace_ints(int64_t nq, int nprobe, int nlist) {
    const size_t n_pairs = (size_t)nq * nprobe;
    const size_t max_items = n_pairs / LM_QT + (size_t)nlist + 1;
    // list_cnt[nlist] | q_cand[nq] | gthr[nq] | next_item, n_items | pair_start[nlist+1] | cursor[nlist] | pairs | (pad) | items (int4)
    return (size_t)nlist + 2 * (size_t)nq + 2 + (size_t)nlist + 1 + (size_t)nlist + n_pairs + 4 + 4 * max_items;
}

// ws: ivf_lm_workspace_ints() ints; part_key / part_id: [nprobe][nq][round_up_ktop(k)]
int launch_ivf_listmajor(const float* q, const float* vectors, const int32_t* offsets, const int32_t* id_map, int nlist,
                         const int32_t* list_order, const int32_t* probes, int64_t nq, int nprobe, int k, int32_t* ws, float* part_key, int32_t* part_id,
                         int32_t* out_counts, unsigned long long* total, int num_sms, cudaStream_t st) {
    if (nq <= 0) return VS_OK;
    const int ktop = round_up_ktop(k);
    if (ktop == 0) return fail(VS_ERR_UNSUPPORTED, "IVF search: k > 32 is not implemented");
    const int64_t n_pairs = nq * nprobe;
    if (n_pairs > 0x7fffffff) return fail(VS_ERR_UNSUPPORTED, "IVF list-major scan: too many (query, probe) pairs");
    const size_t max_items = (size_t)n_pairs / LM_QT + (size_t)nlist + 1;
    int32_t* list_cnt = ws;
    int32_t* q_cand = list_cnt + nlist;
    int32_t* gthr = q_cand + nq;
    int32_t* next_item = gthr + nq;
    int32_t* n_items = next_item + 1;
    int32_t* pair_start = n_items + 1;
    int32_t* cursor = pair_start + nlist + 1;
    int32_t* pairs = cursor + nlist;
    int4* items = reinterpret_cast<int4*>(ws + ((pairs + n_pairs - ws + 3) & ~(ptrdiff_t)3));  // 16-byte aligned (ws is)
    (void)max_items;
    VSB_CUDA(cudaMemsetAsync(ws, 0, sizeof(int32_t) * ((size_t)nlist + 2 * (size_t)nq + 2), st));
    const unsigned gb = (unsigned)std::min<int64_t>(ceil_div64(n_pairs, 256), 148 * 8);
    lm_count_kernel<<<gb, 256, 0, st>>>(probes, n_pairs, nprobe, offsets, list_cnt, q_cand);
    lm_scan_kernel<<<1, 1024, 0, st>>>(list_cnt, nlist, list_order, offsets, pair_start, cursor, items, n_items, LM_QT);
    lm_fill_kernel<<<gb, 256, 0, st>>>(probes, n_pairs, cursor, pairs);
    LmParams p{q, vectors, offsets, id_map, pairs, items, n_items, next_item, gthr, part_key, part_id, nq, nprobe};
    const int grid = 2 * num_sms;  // two resident CTAs per SM
    switch (ktop) {
        case 1: ivf_lm_kernel<1><<<grid, LM_THREADS, LM_SMEM + 128, st>>>(p); break;
        case 5: ivf_lm_kernel<5><<<grid, LM_THREADS, LM_SMEM + 128, st>>>(p); break;
        case 10: ivf_lm_kernel<10><<<grid, LM_THREADS, LM_SMEM + 128, st>>>(p); break;
        case 16: ivf_lm_kernel<16><<<grid, LM_THREADS, LM_SMEM + 128, st>>>(p); break;
        default: ivf_lm_kernel<32><<<grid, LM_THREADS, LM_SMEM + 128, st>>>(p); break;
    }
    lm_counts_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(nq, 256), 148), 256, 0, st>>>(q_cand, nq, k, out_counts, total);
    VSB_CUDA(cudaGetLastError());
    return VS_OK;
}

// ---- tensor-core list-major scan (exact_tc.cuh, IVF = true) ------------------------------------------------------
// queries of the pairs, in pair order, split for the TF32 tensor core (hi = rna_tf32(q), lo = rna_tf32(q - hi));
// *not_exact |= 1 when some lo != 0.  One warp per pair.
__global__ void __launch_bounds__(256) lm_gather_split_kernel(const float* __restrict__ q, const int32_t* __restrict__ pairs,
                                                              int64_t n_pairs, int nprobe, float* __restrict__ qhi,
                                                              float* __restrict__ qlo, int* __restrict__ not_exact) {
    const int lane = threadIdx.x & 31;
    int inexact = 0;
    for (int64_t pi = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); pi < n_pairs; pi += (int64_t)gridDim.x * 8) {
        const int query = pairs[pi] / nprobe;
        const float4 v = __ldg(reinterpret_cast<const float4*>(q + (size_t)query * 128) + lane);
        float4 h, l;
        uint32_t t;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.x)); h.x = __uint_as_float(t);
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.y)); h.y = __uint_as_float(t);
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.z)); h.z = __uint_as_float(t);
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.w)); h.w = __uint_as_float(t);
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.x - h.x)); l.x = __uint_as_float(t);
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.y - h.y)); l.y = __uint_as_float(t);
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.z - h.z)); l.z = __uint_as_float(t);
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.w - h.w)); l.w = __uint_as_float(t);
        reinterpret_cast<float4*>(qhi + (size_t)pi * 128)[lane] = h;
        reinterpret_cast<float4*>(qlo + (size_t)pi * 128)[lane] = l;
        inexact |= (l.x != 0.f) | (l.y != 0.f) | (l.z != 0.f) | (l.w != 0.f);
    }
    if (__any_sync(0xffffffffu, inexact) && lane == 0) atomicOr(not_exact, 1);
}

size_t ivf_tc_workspace_ints(int64_t nq, int nprobe, int nlist) {
    const size_t n_pairs = (size_t)nq * nprobe;
    // list_cnt[nlist] q_cand[nq] n_items[1] flag[1] pair_start[nlist+1] cursor[nlist] pairs[n_pairs] (+ alignment) items[4 x (n_pairs/128 + nlist + 1)]
    return (size_t)nlist * 3 + (size_t)nq + 8 + n_pairs + 4 + 4 * (n_pairs / 128 + (size_t)nlist + 1);
}

// Pair grouping for the tensor-core scan: work items of <= 128 pairs, long lists first; gathers and splits the queries.
// Outputs (inside ws): *items, *n_items (device), *pairs; qhi/qlo [(n_pairs + 128) x 128]; q_not_exact (device flag).
int launch_ivf_tc_prep(const float* q, const int32_t* offsets, int nlist, const int32_t* list_order, const int32_t* probes, int64_t nq,
                       int nprobe, int32_t* ws, float* qhi, float* qlo, const int4** items_out, const int32_t** n_items_out,
                       const int32_t** pairs_out, int32_t** q_cand_out, int** q_not_exact_out, cudaStream_t st) {
    const int64_t n_pairs = nq * nprobe;
    if (n_pairs > 0x7fffffff - 256) return fail(VS_ERR_UNSUPPORTED, "IVF list-major scan: too many (query, probe) pairs");
    int32_t* list_cnt = ws;
    int32_t* q_cand = list_cnt + nlist;
    int32_t* n_items = q_cand + nq;
    int32_t* flag = n_items + 1;
    int32_t* pair_start = flag + 1;
    int32_t* cursor = pair_start + nlist + 1;
    int32_t* pairs = cursor + nlist;
    int4* items = reinterpret_cast<int4*>(ws + ((pairs + n_pairs - ws + 3) & ~(ptrdiff_t)3));
    VSB_CUDA(cudaMemsetAsync(ws, 0, sizeof(int32_t) * ((size_t)nlist + (size_t)nq + 2), st));
    const unsigned gb = (unsigned)std::min<int64_t>(ceil_div64(n_pairs, 256), 148 * 8);
    lm_count_kernel<<<gb, 256, 0, st>>>(probes, n_pairs, nprobe, offsets, list_cnt, q_cand);
    lm_scan_kernel<<<1, 1024, 0, st>>>(list_cnt, nlist, list_order, offsets, pair_start, cursor, items, n_items, 128);
    lm_fill_kernel<<<gb, 256, 0, st>>>(probes, n_pairs, cursor, pairs);
    lm_gather_split_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(n_pairs, 8), 148 * 16), 256, 0, st>>>(q, pairs, n_pairs, nprobe, qhi,
                                                                                                           qlo, flag);
    VSB_CUDA(cudaGetLastError());
    *items_out = items;
    *n_items_out = n_items;
    *pairs_out = pairs;
    *q_cand_out = q_cand;
    *q_not_exact_out = flag;
    return VS_OK;
}

int launch_ivf_counts(const int32_t* q_cand, int64_t nq, int k, int32_t* out_counts, unsigned long long* total, cudaStream_t st) {
    lm_counts_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(nq, 256), 148), 256, 0, st>>>(q_cand, nq, k, out_counts, total);
    VSB_CUDA(cudaGetLastError());
    return VS_OK;
}

}  // namespace vsb
