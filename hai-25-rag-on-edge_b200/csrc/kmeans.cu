// k-means++ seeding and the device-side bookkeeping of Lloyd's iterations for the IVF index builder
// (qidk_ivf/prepare/create_ivf_model.py:102-119: sklearn KMeans(n_clusters, random_state=42, n_init=1, max_iter=100),
// i.e. greedy k-means++ initialisation + Lloyd; lists = np.where(cluster_ids == i), ascending ids).
//
// Seeding = sklearn's _kmeans_plusplus: the first centre uniformly at random, then per new centre 2 + floor(ln k) candidates
// drawn with probability proportional to D^2 (squared distance to the nearest centre chosen so far); the candidate that
// lowers the potential sum(D^2) most is kept.  One step = four launches, no host round trip:
//   locate   (1 block)   turns the step's uniform numbers into row indices by a two-level prefix search over D^2
//   pass     (grid)      distances of every row to the L candidates (one coalesced read of the base), candidate potentials
//   pick     (1 block)   fixed-order sum of the per-block potentials, arg-min, records the chosen row
//   apply    (grid)      D^2 = min(D^2, distance to the chosen candidate), per-1024-row sums of D^2 for the next locate
// Every reduction runs in a fixed order: the seeding is a pure function of (data, seed).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <random>
#include <vector>

#include "kernels.cuh"
#include "vsb_common.cuh"

namespace vsb {

constexpr int KPP_MAXL = 16;      // candidates per step (2 + ln k <= 16 up to k = 1.2 M)
constexpr int KPP_SEG = 1024;     // rows per D^2 segment sum
constexpr int KPP_PASS_BLOCKS = 148 * 8;

// dcand[j][i] = ||x_i - c_j||^2 (fp32, components summed by a fixed lane tree); part[blk][j] = sum over the block's rows of
// min(D_i, dcand[j][i]) in double
__global__ void __launch_bounds__(256) kpp_pass_kernel(const float* __restrict__ base, int64_t n, const int32_t* __restrict__ cand,
                                                       int L, const float* __restrict__ D, float* __restrict__ dcand,
                                                       double* __restrict__ part) {
    __shared__ float4 s_c[KPP_MAXL][32];
    __shared__ double s_pot[8][KPP_MAXL];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < L * 32; i += blockDim.x)
        s_c[i >> 5][i & 31] = __ldg(reinterpret_cast<const float4*>(base + (size_t)cand[i >> 5] * 128) + (i & 31));
    __syncthreads();
    double pot[KPP_MAXL];
#pragma unroll
    for (int j = 0; j < KPP_MAXL; ++j) pot[j] = 0.0;
    const int64_t warps = (int64_t)gridDim.x * 8;
    for (int64_t r = (int64_t)blockIdx.x * 8 + wib; r < n; r += warps) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(base + (size_t)r * 128) + lane);
        const float dr = D[r];
#pragma unroll
        for (int j = 0; j < KPP_MAXL; ++j) {
            if (j >= L) break;
            const float4 c = s_c[j][lane];
            const float a = x.x - c.x, b = x.y - c.y, cc = x.z - c.z, d = x.w - c.w;
            float p = __fmul_rn(a, a);
            p = fmaf(b, b, p);
            p = fmaf(cc, cc, p);
            p = fmaf(d, d, p);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) p = __fadd_rn(p, __shfl_down_sync(0xffffffffu, p, o));
            if (lane == 0) {
                dcand[(size_t)j * n + r] = p;
                pot[j] += (double)fminf(dr, p);
            }
        }
    }
    if (lane == 0)
        for (int j = 0; j < L; ++j) s_pot[wib][j] = pot[j];
    __syncthreads();
    if (threadIdx.x < L) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += s_pot[w][threadIdx.x];
        part[(size_t)blockIdx.x * KPP_MAXL + threadIdx.x] = s;
    }
}

// one block: total potential per candidate (fixed order over the blocks), arg-min (first minimum), chosen[step] = its row
__global__ void kpp_pick_kernel(const double* __restrict__ part, int n_blocks, int L, const int32_t* __restrict__ cand,
                                int32_t* __restrict__ chosen, int step, int32_t* __restrict__ best_j) {
    __shared__ double s_tot[KPP_MAXL];
    if (threadIdx.x < L) {
        double s = 0.0;
        for (int b = 0; b < n_blocks; ++b) s += part[(size_t)b * KPP_MAXL + threadIdx.x];
        s_tot[threadIdx.x] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int bj = 0;
        for (int j = 1; j < L; ++j)
            if (s_tot[j] < s_tot[bj]) bj = j;
        *best_j = bj;
        chosen[step] = cand[bj];
    }
}

// D_i = min(D_i, dcand[*best_j][i]); seg[s] = sum of D over rows [1024 s, 1024 s + 1024) (double, fixed tree)
__global__ void __launch_bounds__(KPP_SEG) kpp_apply_kernel(float* __restrict__ D, const float* __restrict__ dcand, int64_t n,
                                                            const int32_t* __restrict__ best_j, double* __restrict__ seg) {
    __shared__ double s[KPP_SEG];
    const int64_t i = (int64_t)blockIdx.x * KPP_SEG + threadIdx.x;
    double v = 0.0;
    if (i < n) {
        const float d = fminf(D[i], dcand[(size_t)(*best_j) * n + i]);
        D[i] = d;
        v = (double)d;
    }
    s[threadIdx.x] = v;
    __syncthreads();
    for (int o = KPP_SEG / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) seg[blockIdx.x] = s[0];
}

// one block of 1024 threads: cand[j] = the row whose cumulative D^2 interval contains u[j] * sum(D^2)
__global__ void __launch_bounds__(1024) kpp_locate_kernel(const float* __restrict__ D, int64_t n, const double* __restrict__ seg,
                                                          int n_seg, const float* __restrict__ u, int L, int32_t* __restrict__ cand) {
    __shared__ double s_part[1024], s_excl[1024], s_scan[1024];
    __shared__ int s_seg[KPP_MAXL];
    __shared__ double s_res[KPP_MAXL];
    __shared__ int s_first;
    const int t = threadIdx.x;
    const int per = (n_seg + 1023) / 1024;
    const int b0 = min(t * per, n_seg), b1 = min(b0 + per, n_seg);
    double mine = 0.0;
    for (int b = b0; b < b1; ++b) mine += seg[b];
    s_part[t] = mine;
    s_scan[t] = mine;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {  // inclusive scan
        const double a = t >= o ? s_scan[t - o] : 0.0;
        __syncthreads();
        s_scan[t] += a;
        __syncthreads();
    }
    s_excl[t] = t ? s_scan[t - 1] : 0.0;  // thread t owns [s_scan[t-1], s_scan[t]): contiguous intervals, one owner per target
    const double total = s_scan[1023];
    if (t < L) {
        s_seg[t] = -1;
        s_res[t] = 0.0;
    }
    __syncthreads();
    for (int j = 0; j < L; ++j) {
        const double target = (double)u[j] * total;
        if (target >= s_excl[t] && target < s_scan[t]) {  // exactly one thread owns the target
            double acc = s_excl[t];
            int b = b0;
            for (; b < b1 - 1; ++b) {
                if (target < acc + seg[b]) break;
                acc += seg[b];
            }
            s_seg[j] = b;
            s_res[j] = target - acc;
        }
    }
    __syncthreads();
    for (int j = 0; j < L; ++j) {
        int sgm = s_seg[j];
        if (sgm < 0) {  // target == total by rounding (or an all-zero potential): the last segment
            sgm = n_seg - 1;
            if (t == 0) s_res[j] = 1e300;
        }
        __syncthreads();
        const int64_t i = (int64_t)sgm * KPP_SEG + t;
        const double v = i < n ? (double)D[i] : 0.0;
        s_scan[t] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const double a = t >= o ? s_scan[t - o] : 0.0;
            __syncthreads();
            s_scan[t] += a;
            __syncthreads();
        }
        // first row whose inclusive sum exceeds the residual; none (rounding): the last row of the segment with D > 0
        const bool hit = s_scan[t] > s_res[j] && v > 0.0;
        if (t == 0) s_first = 1 << 30;
        __syncthreads();
        if (hit) atomicMin(&s_first, t);
        __syncthreads();
        if (t == 0) {
            int pick = s_first;
            if (pick == (1 << 30)) {  // fall back to the last valid row of the segment
                const int64_t last = min((int64_t)sgm * KPP_SEG + KPP_SEG, n) - 1;
                pick = (int)(last - (int64_t)sgm * KPP_SEG);
            }
            cand[j] = (int32_t)((int64_t)sgm * KPP_SEG + pick);
        }
        __syncthreads();
    }
}

__global__ void kpp_gather_kernel(const float* __restrict__ base, const int32_t* __restrict__ chosen, int k, float* __restrict__ cent) {
    const int c = blockIdx.x;
    if (c < k) reinterpret_cast<float4*>(cent + (size_t)c * 128)[threadIdx.x] =
        __ldg(reinterpret_cast<const float4*>(base + (size_t)chosen[c] * 128) + threadIdx.x);
}

// greedy k-means++ (sklearn's rule) over base_dev[n x 128] -> cent_dev[k x 128]
int launch_kmeanspp(const float* base_dev, int64_t n, int k, uint64_t seed, float* cent_dev, cudaStream_t st) {
    if (n <= 0 || k <= 0 || k > n) return fail(VS_ERR_INVALID, "k-means++: bad sizes");
    const int L = std::min(KPP_MAXL, 2 + (int)std::log((double)k));
    const int n_seg = (int)ceil_div64(n, KPP_SEG);
    const int pass_blocks = (int)std::min<int64_t>(KPP_PASS_BLOCKS, ceil_div64(n, 8));
    DevBuf D, dcand, part, seg, u, cand, chosen, bestj;
    auto cleanup = [&]() {
        for (DevBuf* b : {&D, &dcand, &part, &seg, &u, &cand, &chosen, &bestj}) b->release();
    };
    auto body = [&]() -> int {
        VSB_TRY(D.reserve(sizeof(float) * (size_t)n));
        VSB_TRY(dcand.reserve(sizeof(float) * (size_t)n * L));
        VSB_TRY(part.reserve(sizeof(double) * (size_t)pass_blocks * KPP_MAXL));
        VSB_TRY(seg.reserve(sizeof(double) * (size_t)n_seg));
        VSB_TRY(u.reserve(sizeof(float) * (size_t)k * L));
        VSB_TRY(cand.reserve(sizeof(int32_t) * KPP_MAXL));
        VSB_TRY(chosen.reserve(sizeof(int32_t) * (size_t)k));
        VSB_TRY(bestj.reserve(sizeof(int32_t)));
        std::mt19937_64 rng(seed);
        std::vector<float> hu((size_t)k * L);
        for (float& v : hu) v = (float)((rng() >> 40) * (1.0 / 16777216.0));  // 24 random bits in [0, 1)
        const int32_t first = (int32_t)(rng() % (uint64_t)n);
        VSB_CUDA(cudaMemcpyAsync(u.p, hu.data(), sizeof(float) * hu.size(), cudaMemcpyHostToDevice, st));
        VSB_CUDA(cudaMemcpyAsync(cand.p, &first, sizeof(int32_t), cudaMemcpyHostToDevice, st));
        VSB_TRY(launch_fill_f32(D.as<float>(), n, __builtin_inff(), st));
        VSB_CUDA(cudaStreamSynchronize(st));  // hu / first are host temporaries
        for (int step = 0; step < k; ++step) {
            const int l = step == 0 ? 1 : L;
            if (step > 0) {
                kpp_locate_kernel<<<1, 1024, 0, st>>>(D.as<float>(), n, seg.as<double>(), n_seg, u.as<float>() + (size_t)step * L, l,
                                                      cand.as<int32_t>());
            }
            kpp_pass_kernel<<<pass_blocks, 256, 0, st>>>(base_dev, n, cand.as<int32_t>(), l, D.as<float>(), dcand.as<float>(),
                                                         part.as<double>());
            kpp_pick_kernel<<<1, 32, 0, st>>>(part.as<double>(), pass_blocks, l, cand.as<int32_t>(), chosen.as<int32_t>(), step,
                                              bestj.as<int32_t>());
            kpp_apply_kernel<<<n_seg, KPP_SEG, 0, st>>>(D.as<float>(), dcand.as<float>(), n, bestj.as<int32_t>(), seg.as<double>());
        }
        VSB_CUDA(cudaGetLastError());
        kpp_gather_kernel<<<k, 32, 0, st>>>(base_dev, chosen.as<int32_t>(), k, cent_dev);
        VSB_CUDA(cudaGetLastError());
        VSB_CUDA(cudaStreamSynchronize(st));
        return VS_OK;
    };
    const int rc = body();
    cleanup();
    return rc;
}

// ------------------------------------------------------------------------------------------------
// Lloyd bookkeeping on the device
// ------------------------------------------------------------------------------------------------
// number of rows whose label changed (and the copy prev = cur)
__global__ void labels_changed_kernel(const int32_t* __restrict__ cur, int32_t* __restrict__ prev, int64_t n, int32_t* __restrict__ changed) {
    int c = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t a = cur[i];
        c += a != prev[i];
        prev[i] = a;
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(changed, c);
}
int launch_labels_changed(const int32_t* cur, int32_t* prev, int64_t n, int32_t* changed_zeroed, cudaStream_t st) {
    labels_changed_kernel<<<148 * 4, 256, 0, st>>>(cur, prev, n, changed_zeroed);
    VSB_CUDA(cudaGetLastError());
    return VS_OK;
}

// Stable counting sort of the rows by label = the inverted lists with ascending ids inside a list (np.where order):
//   count   per 1024-row block, rows per label                           -> cnt[label][block]
//   scan    exclusive prefix over (label major, block minor) + the CSR offsets of the lists
//   scatter row i of block b goes to  start[label][b] + (number of earlier rows of b with the same label)
__global__ void __launch_bounds__(1024) ms_count_kernel(const int32_t* __restrict__ lab, int64_t n, int nlist, int n_blocks,
                                                        int32_t* __restrict__ cnt) {
    extern __shared__ int32_t s_hist[];
    for (int c = threadIdx.x; c < nlist; c += 1024) s_hist[c] = 0;
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    if (i < n) atomicAdd(&s_hist[lab[i]], 1);
    __syncthreads();
    for (int c = threadIdx.x; c < nlist; c += 1024) cnt[(size_t)c * n_blocks + blockIdx.x] = s_hist[c];
}
__global__ void __launch_bounds__(1024) ms_scan_kernel(int32_t* __restrict__ cnt, int64_t total, int nlist, int n_blocks,
                                                       int32_t* __restrict__ offsets) {
    __shared__ int32_t s[1024];
    const int t = threadIdx.x;
    const int64_t per = (total + 1023) / 1024;
    const int64_t a = t * per < total ? t * per : total, b = a + per < total ? a + per : total;
    int32_t mine = 0;
    for (int64_t i = a; i < b; ++i) mine += cnt[i];
    s[t] = mine;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int32_t v = t >= o ? s[t - o] : 0;
        __syncthreads();
        s[t] += v;
        __syncthreads();
    }
    int32_t run = s[t] - mine;
    for (int64_t i = a; i < b; ++i) {
        const int32_t c = cnt[i];
        cnt[i] = run;
        if (i % n_blocks == 0) offsets[i / n_blocks] = run;  // first block of a label = start of its list
        run += c;
    }
    if (t == 1023) offsets[nlist] = s[1023];
}
__global__ void __launch_bounds__(1024) ms_scatter_kernel(const int32_t* __restrict__ lab, int64_t n, int nlist, int n_blocks,
                                                          const int32_t* __restrict__ start, int32_t* __restrict__ members) {
    extern __shared__ int32_t s_run[];  // rows of this block already placed, per label
    for (int c = threadIdx.x; c < nlist; c += 1024) s_run[c] = 0;
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int32_t l = i < n ? lab[i] : -1;
    for (int w = 0; w < 32; ++w) {  // warps in order: stable
        if (wib == w) {
            const unsigned peers = __match_any_sync(0xffffffffu, l);
            if (l >= 0) {
                const int rank = __popc(peers & ((1u << lane) - 1u));
                const int before = s_run[l];
                members[(size_t)start[(size_t)l * n_blocks + blockIdx.x] + before + rank] = (int32_t)i;
                __syncwarp(peers);
                if (rank == 0) s_run[l] = before + __popc(peers);
            }
        }
        __syncthreads();
    }
}
// lab[n] (values in [0, nlist)) -> offsets[nlist+1], members[n] (ascending ids inside every list); ws: nlist * ceil(n/1024) ints
size_t multisplit_workspace_ints(int64_t n, int nlist) { return (size_t)nlist * (size_t)ceil_div64(n, 1024); }
int launch_multisplit(const int32_t* lab, int64_t n, int nlist, int32_t* ws, int32_t* offsets, int32_t* members, cudaStream_t st) {
    const int n_blocks = (int)ceil_div64(n, 1024);
    const size_t smem = sizeof(int32_t) * (size_t)nlist;
    if (smem > 200 * 1024) return fail(VS_ERR_UNSUPPORTED, "multisplit: too many lists");
    if (smem > 48 * 1024) {  // per device, hence not cached in a static
        VSB_CUDA(cudaFuncSetAttribute(ms_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        VSB_CUDA(cudaFuncSetAttribute(ms_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
    ms_count_kernel<<<n_blocks, 1024, smem, st>>>(lab, n, nlist, n_blocks, ws);
    ms_scan_kernel<<<1, 1024, 0, st>>>(ws, (int64_t)nlist * n_blocks, nlist, n_blocks, offsets);
    ms_scatter_kernel<<<n_blocks, 1024, smem, st>>>(lab, n, nlist, n_blocks, ws, members);
    VSB_CUDA(cudaGetLastError());
    return VS_OK;
}

// sum over the rows of max(dist, 0) in double, fixed order (per-block partials, then one thread)
__global__ void __launch_bounds__(1024) inertia_part_kernel(const float* __restrict__ dist, int64_t n, double* __restrict__ part) {
    __shared__ double s[1024];
    const int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    s[threadIdx.x] = i < n ? (double)fmaxf(dist[i], 0.f) : 0.0;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.x] = s[0];
}
__global__ void inertia_sum_kernel(const double* __restrict__ part, int n_parts, double* __restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < n_parts; ++i) s += part[i];
        *out = s;
    }
}
int launch_inertia(const float* dist, int64_t n, double* part_ws, double* out, cudaStream_t st) {
    const int nb = (int)ceil_div64(n, 1024);
    inertia_part_kernel<<<nb, 1024, 0, st>>>(dist, n, part_ws);
    inertia_sum_kernel<<<1, 32, 0, st>>>(part_ws, nb, out);
    VSB_CUDA(cudaGetLastError());
    return VS_OK;
}

}  // namespace vsb
