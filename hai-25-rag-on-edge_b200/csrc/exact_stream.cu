// K2 — HBM-streaming exact L2 scan for small query batches (<= 8 queries per pass), CUDA-core FFMA in fp32.
// Replaces, for batch-1 style calls, the reference's per-query sgemm(M=1) + distance loop + select_topk
// (cpu/cpu_baseline.cpp:222-248): one pass over the base (N*512 B + N*4 B norms) per group of queries.
//
// Layout: a warp reads 4 consecutive rows (2 KB) per step; the 8 lanes of a row group read one contiguous
// 128-B line per LDG.128 (lane `sub` owns float4 columns sub, sub+8, sub+16, sub+24).  Two steps are kept in
// flight (8 x 16 B per lane).  Partial dots are reduced over the 8 lanes with xor-shuffles; lane `sub` then
// owns query `sub` and keeps its top-k list in registers.  Lists are merged inside the CTA (warp merge, then
// one warp per query over the per-warp results) and one sorted list per (CTA, query) goes to global memory;
// merge_lists_kernel (K3) finishes.
#include "kernels.cuh"
#include "vsb_common.cuh"

namespace vsb {

constexpr int ST_THREADS = 256;
constexpr int ST_WARPS = ST_THREADS / 32;
constexpr int ST_QB_MAX = 8;  // queries per pass
// QB <= 2: light register footprint -> 2 CTAs/SM, 2 steps in flight; QB 4/8: 1 CTA/SM, 4 steps in flight.
// Either way ~64 KB of loads are outstanding per SM (Little: 6.5 TB/s x ~800 ns / 148 SMs ~ 35 KB).
__host__ __device__ constexpr int st_ctas_per_sm(int qb) { return qb <= 2 ? 2 : 1; }
__host__ __device__ constexpr int st_unroll(int qb) { return qb <= 2 ? 2 : 4; }

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}

template <int KTOP, int ST_QB, bool HAS_LB>
__global__ void __launch_bounds__(ST_THREADS, st_ctas_per_sm(ST_QB))
exact_stream_kernel(const float* __restrict__ base, const float* __restrict__ bnorm, int64_t n,
                    const float* __restrict__ q, const float* __restrict__ qnorm, int nq,
                    const float* __restrict__ lb_key, const int32_t* __restrict__ lb_id, float* __restrict__ part_key,
                    int32_t* __restrict__ part_id) {
    __shared__ float s_key[ST_WARPS][ST_QB][KTOP];
    __shared__ int32_t s_id[ST_WARPS][ST_QB][KTOP];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int sub = lane & 7;   // column slot inside the row / query owned by this lane
    const int g = lane >> 3;    // row inside the 4-row step
    const float INF = __int_as_float(0x7f800000);

    // query fragments: qv[qi][i] = float4 column (i*8 + sub) of query qi (zeros beyond nq)
    float4 qv[ST_QB][4];
#pragma unroll
    for (int qi = 0; qi < ST_QB; ++qi)
#pragma unroll
        for (int i = 0; i < 4; ++i)
            qv[qi][i] = qi < nq ? __ldg(reinterpret_cast<const float4*>(q + (size_t)qi * 128) + i * 8 + sub)
                                : make_float4(0.f, 0.f, 0.f, 0.f);
    const bool owner = sub < nq;
    const float my_qn = owner ? __ldg(qnorm + sub) : 0.f;
    float lbk = -INF;
    int32_t lbi = -1;
    if (HAS_LB && owner) {
        lbk = __ldg(lb_key + sub);
        lbi = __ldg(lb_id + sub);
    }
    RegTopK<KTOP> top;
    top.init();

    const int64_t n_steps = (n + 3) >> 2;
    const int64_t wglobal = (int64_t)blockIdx.x * ST_WARPS + warp;
    const int64_t wtotal = (int64_t)gridDim.x * ST_WARPS;

    constexpr int U = st_unroll(ST_QB);
    for (int64_t s0 = wglobal; s0 < n_steps; s0 += U * wtotal) {
        float4 x[U][4];
        int64_t rows[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t step = s0 + u * wtotal;
            const int64_t row = step * 4 + g;
            rows[u] = (step < n_steps && row < n) ? row : -1;
            const float4* src = reinterpret_cast<const float4*>(base + (rows[u] >= 0 ? rows[u] : 0) * 128);
#pragma unroll
            for (int i = 0; i < 4; ++i) x[u][i] = ldg_stream(src + i * 8 + sub);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const float bn = rows[u] >= 0 ? __ldg(bnorm + rows[u]) : INF;
            float mydot = 0.f;
#pragma unroll
            for (int qi = 0; qi < ST_QB; ++qi) {
                float p = 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    p = fmaf(qv[qi][i].x, x[u][i].x, p);
                    p = fmaf(qv[qi][i].y, x[u][i].y, p);
                    p = fmaf(qv[qi][i].z, x[u][i].z, p);
                    p = fmaf(qv[qi][i].w, x[u][i].w, p);
                }
                p += __shfl_xor_sync(0xffffffffu, p, 1);
                p += __shfl_xor_sync(0xffffffffu, p, 2);
                p += __shfl_xor_sync(0xffffffffu, p, 4);
                if (sub == qi) mydot = p;
                if (qi + 1 >= nq) break;  // uniform: nq is a kernel argument
            }
            // (qn + bn) - 2*dot   cpu_baseline.cpp:241
            float d = fmaf(-2.0f, mydot, my_qn + bn);
            if (HAS_LB) {
                const bool after = d > lbk || (d == lbk && (int32_t)rows[u] > lbi);
                d = after ? d : INF;
            }
            if (owner && rows[u] >= 0 && d < top.threshold()) top.insert(d, (int32_t)rows[u]);
        }
    }

    // ---- in-CTA merge: per warp, per query (lanes with sub == qi hold that query's lists; rows were visited in
    // ascending order per lane, so each list is canonical) -------------------------------------------------------
    for (int qi = 0; qi < nq; ++qi) {
        RegTopK<KTOP> mine;
        if (sub == qi) {
            mine = top;
        } else {
            mine.init();
        }
        warp_merge_lists<KTOP>(mine, KTOP, &s_key[warp][qi][0], &s_id[warp][qi][0]);
    }
    __syncthreads();
    if (warp < nq) {
        RegTopK<KTOP> mine;
        mine.init();
        if (lane < ST_WARPS) {
#pragma unroll
            for (int i = 0; i < KTOP; ++i) {
                mine.key[i] = s_key[lane][warp][i];
                mine.id[i] = s_id[lane][warp][i];
            }
        }
        float* pk = part_key + ((size_t)blockIdx.x * nq + warp) * KTOP;
        int32_t* pi = part_id + ((size_t)blockIdx.x * nq + warp) * KTOP;
        warp_merge_lists<KTOP>(mine, KTOP, pk, pi);
    }
}

static int stream_qb(int nq) { return nq <= 1 ? 1 : nq <= 2 ? 2 : nq <= 4 ? 4 : 8; }

int stream_num_ctas(int device, int nq) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    return sms * st_ctas_per_sm(stream_qb(nq));
}

int launch_exact_stream(const float* base, const float* bnorm, int64_t n, const float* q, const float* qnorm, int nq,
                        int ktop, const float* lb_key, const int32_t* lb_id, float* part_key, int32_t* part_id,
                        int n_ctas, cudaStream_t st) {
    if (nq < 1 || nq > ST_QB_MAX) return fail(VS_ERR_INVALID, "stream kernel handles 1..8 queries per pass");
    if (lb_key && ktop != 32) return fail(VS_ERR_INVALID, "stream: lower bound needs the 32-entry list");
    const int qb = stream_qb(nq);
#define VSB_STREAM_LAUNCH(KT, QB, LB)                                                                              \
    exact_stream_kernel<KT, QB, LB><<<n_ctas, ST_THREADS, 0, st>>>(base, bnorm, n, q, qnorm, nq, lb_key, lb_id, \
                                                                   part_key, part_id)
#define VSB_STREAM_QB(KT, LB)                      \
    switch (qb) {                                  \
        case 1: VSB_STREAM_LAUNCH(KT, 1, LB); break; \
        case 2: VSB_STREAM_LAUNCH(KT, 2, LB); break; \
        case 4: VSB_STREAM_LAUNCH(KT, 4, LB); break; \
        default: VSB_STREAM_LAUNCH(KT, 8, LB); break; \
    }
    switch (ktop) {
        case 1: VSB_STREAM_QB(1, false) break;
        case 5: VSB_STREAM_QB(5, false) break;
        case 10: VSB_STREAM_QB(10, false) break;
        case 16: VSB_STREAM_QB(16, false) break;
        case 32:
            if (lb_key) {
                VSB_STREAM_QB(32, true)
            } else {
                VSB_STREAM_QB(32, false)
            }
            break;
        default:
            return fail(VS_ERR_INVALID, "stream: unsupported list size");
    }
#undef VSB_STREAM_QB
#undef VSB_STREAM_LAUNCH
    VSB_CUDA(cudaGetLastError());
    return VS_OK;
}

}  // namespace vsb
