// Host-side planning and launch of K1 (exact_tc.cuh).
#include "exact_tc.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "kernels.cuh"

namespace vsb {

// Split the base into n_splits contiguous tile ranges so that n_mtiles * n_splits units fill the grid in an
// (almost) integral number of rounds.  More splits = better balance but more partial lists and more list
// warm-up; each split keeps at least 16 tiles unless the problem is tiny.
TcPlan tc_make_plan(int64_t n, int64_t nq, int num_sms, int mode) {
    TcPlan pl{};
    pl.n_tiles = (int)ceil_div64(n, TC_BN);
    pl.n_mtiles = (int)ceil_div64(nq, TC_BM);
    // CTA pairs (two query tiles share every base tile through TMA multicast, VSB_TC_CL=2) are implemented and give
    // identical results, but measured no gain on B200 (tools/tc_cluster_ab.py, profiles/r2_tc_cluster_ab.txt: the sweep is
    // bound by shared-memory bandwidth and the epilogue, not by L2 reads) and couple the two CTAs' stalls: independent CTAs
    // stay the default
    pl.cl = 1;
    if (const char* e = getenv("VSB_TC_CL")) pl.cl = atoi(e) == 2 && num_sms >= 2 && mode != TC_F16 ? 2 : 1;
    const int cols = pl.cl == 2 ? (pl.n_mtiles + 1) / 2 : pl.n_mtiles;   // unit columns
    const int workers = pl.cl == 2 ? num_sms / 2 : num_sms;              // CTAs or CTA pairs
    // few query tiles (small batches): allow enough splits for two units per worker; many query tiles: cap the number of
    // partial lists per query
    const int want = std::max(64, (int)ceil_div64(2 * (int64_t)workers, cols));
    const int max_splits = std::max(1, std::min(pl.n_tiles / 16, want));
    int best_s = 1;
    double best_cost = 1e30;
    for (int s = 1; s <= std::max(1, max_splits); ++s) {
        const int tps = (pl.n_tiles + s - 1) / s;
        const int s_eff = (pl.n_tiles + tps - 1) / tps;
        const int64_t units = (int64_t)cols * s_eff;
        const int64_t rounds = ceil_div64(units, workers);
        // time ~ rounds * tiles per unit (+ a small per-unit overhead measured in tiles)
        const double cost = (double)rounds * (tps + 2.0);
        if (cost < best_cost - 1e-9) {
            best_cost = cost;
            best_s = s_eff;
        }
    }
    pl.tiles_per_split = (pl.n_tiles + best_s - 1) / best_s;
    pl.n_splits = (pl.n_tiles + pl.tiles_per_split - 1) / pl.tiles_per_split;
    pl.grid = (int)std::min<int64_t>((int64_t)cols * pl.n_splits, workers) * pl.cl;
    return pl;
}

// shared-memory layout of an instantiation (the fp16 filter pass, KTOP == 32, holds two query tiles per unit)
template <int KTOP, int MODE>
using TcS = TcSmem<MODE, MODE == TC_F16 && KTOP == 32>;

template <int KTOP, int MODE, bool HAS_LB>
static int set_attr_one() {
    VSB_CUDA(cudaFuncSetAttribute(exact_tc_kernel<KTOP, MODE, HAS_LB, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  TcS<KTOP, MODE>::TOTAL));
    if constexpr (MODE != TC_F16)
        VSB_CUDA(cudaFuncSetAttribute(exact_tc_kernel<KTOP, MODE, HAS_LB, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      TcS<KTOP, MODE>::TOTAL));
    return VS_OK;
}

// one launch: CL == 2 as clusters of two CTAs (the pair shares the streamed base tiles)
template <int KTOP, int MODE, bool HAS_LB>
static int launch_one(const TcPlan& plan, const CUtensorMap& tmA_hi, const CUtensorMap& tmA_lo, const CUtensorMap& tmB_hi,
                      const CUtensorMap& tmB_lo, const TcParams& p, cudaStream_t st) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)plan.grid);
    cfg.blockDim = dim3((unsigned)TcS<KTOP, MODE>::THREADS);
    cfg.dynamicSmemBytes = TcS<KTOP, MODE>::TOTAL;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)plan.cl;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if constexpr (MODE != TC_F16) {
        if (plan.cl == 2) {
            VSB_CUDA(cudaLaunchKernelEx(&cfg, exact_tc_kernel<KTOP, MODE, HAS_LB, 2>, tmA_hi, tmA_lo, tmB_hi, tmB_lo, p));
            return VS_OK;
        }
    }
    if (plan.cl != 1) return fail(VS_ERR_INVALID, "tc: CTA pairs are not available for this mode");
    VSB_CUDA(cudaLaunchKernelEx(&cfg, exact_tc_kernel<KTOP, MODE, HAS_LB, 1>, tmA_hi, tmA_lo, tmB_hi, tmB_lo, p));
    return VS_OK;
}

// IVF list-major scan on the tensor cores (exact_tc.cuh, IVF = true): tmA_* = gathered / split queries in pair order,
// tmB_* = list-contiguous vectors (hi / lo), work items and pair table produced by launch_ivf_tc_prep.  split3 = 3xTF32.
template <int KTOP, int MODE>
static int launch_ivf_one(int grid, const CUtensorMap& tmA_hi, const CUtensorMap& tmA_lo, const CUtensorMap& tmB_hi,
                          const CUtensorMap& tmB_lo, const TcParams& p, cudaStream_t st) {
    VSB_CUDA(cudaFuncSetAttribute(exact_tc_kernel<KTOP, MODE, false, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  TcSmem<MODE>::TOTAL));  // per device, hence not cached in a static
    exact_tc_kernel<KTOP, MODE, false, 1, true><<<grid, TcSmem<MODE>::THREADS, TcSmem<MODE>::TOTAL, st>>>(tmA_hi, tmA_lo, tmB_hi, tmB_lo, p);
    VSB_CUDA(cudaGetLastError());
    return VS_OK;
}

int launch_exact_tc_ivf(const CUtensorMap& tmA_hi, const CUtensorMap& tmA_lo, const CUtensorMap& tmB_hi, const CUtensorMap& tmB_lo,
                        const int4* items, const int32_t* n_items, const int32_t* pairs, int nprobe, int32_t* gthr, int nq, int ktop,
                        bool split3, int32_t* cand_cnt, void* cand, int cand_cap, int num_sms, cudaStream_t st) {
    TcParams p{};
    p.gthr = gthr;
    p.cand_cnt = cand_cnt;
    p.cand = reinterpret_cast<uint2*>(cand);
    p.cand_cap = cand_cap;
    p.nq = nq;
    p.items = items;
    p.n_items = n_items;
    p.pairs = pairs;
    p.nprobe = nprobe;
    p.n_mtiles = 1;
    p.n_splits = 1;
    p.tiles_per_split = 1;
#define VSB_IVF_CASE(KT)                                                                              \
    case KT:                                                                                          \
        return split3 ? launch_ivf_one<KT, TC_TF32X3>(num_sms, tmA_hi, tmA_lo, tmB_hi, tmB_lo, p, st) \
                      : launch_ivf_one<KT, TC_TF32X1>(num_sms, tmA_hi, tmA_lo, tmB_hi, tmB_lo, p, st);
    switch (ktop) {
        VSB_IVF_CASE(5)
        VSB_IVF_CASE(10)
        VSB_IVF_CASE(16)
        VSB_IVF_CASE(32)
        default:
            return fail(VS_ERR_INVALID, "ivf tc: unsupported list size");
    }
#undef VSB_IVF_CASE
}

int tc_sample_groups_per_split() { return TC_SAMPLE_GROUPS; }
int tc_lists_per_split(int mode) { return mode == TC_F16 ? 1 : TC_EPI_GROUPS; }

int tc_set_attributes() {
    VSB_TRY((set_attr_one<1, TC_TF32X1, false>()));
    VSB_TRY((set_attr_one<1, TC_TF32X3, false>()));
    VSB_TRY((set_attr_one<5, TC_TF32X1, false>()));
    VSB_TRY((set_attr_one<5, TC_TF32X3, false>()));
    VSB_TRY((set_attr_one<10, TC_TF32X1, false>()));
    VSB_TRY((set_attr_one<10, TC_TF32X3, false>()));
    VSB_TRY((set_attr_one<16, TC_TF32X1, false>()));
    VSB_TRY((set_attr_one<16, TC_TF32X3, false>()));
    VSB_TRY((set_attr_one<32, TC_TF32X1, false>()));
    VSB_TRY((set_attr_one<32, TC_TF32X3, false>()));
    VSB_TRY((set_attr_one<32, TC_TF32X1, true>()));
    VSB_TRY((set_attr_one<32, TC_TF32X3, true>()));
    VSB_TRY((set_attr_one<32, TC_F16, false>()));
    VSB_TRY((set_attr_one<1, TC_F16, false>()));
    return VS_OK;
}

// mode: TC_TF32X1 / TC_TF32X3 (tmA_lo / tmB_lo used by X3 only) / TC_F16 (tmA_hi / tmB_hi are the fp16 maps,
// *key_scale_dev = -2 / (s_q * s_b); list sizes 16 and 32 only).  tmB: full-tile boxes (128 rows) and half-tile boxes
// (64 rows, used when plan.cl == 2)
int launch_exact_tc(const CUtensorMap& tmA_hi, const CUtensorMap& tmA_lo, const TcBaseMaps& tmB, const float* bnorm, int32_t* gthr, int nq,
                    int64_t n_rows, const TcPlan& plan,
                    int ktop, int mode, const float* key_scale_dev, const float* lb_key, const int32_t* lb_id, float* part_key,
                    int32_t* part_id, cudaStream_t st) {
    TcParams p{};
    p.bnorm = bnorm;
    p.gthr = gthr;
    p.lb_key = lb_key;
    p.lb_id = lb_id;
    p.part_key = part_key;
    p.part_id = part_id;
    p.nq = nq;
    p.n_tiles = plan.n_tiles;
    p.n_mtiles = plan.n_mtiles;
    p.n_splits = plan.n_splits;
    p.tiles_per_split = plan.tiles_per_split;
    p.n_rem = (int)(n_rows % TC_BN);
    p.key_scale_ptr = key_scale_dev;
    if ((int64_t)plan.n_tiles != ceil_div64(n_rows, TC_BN)) return fail(VS_ERR_INVALID, "tc: plan does not match the row count");
    {
        const char* e = getenv("VSB_TC_DBG");
        p.dbg = e ? atoi(e) : 0;
        const char* qb = getenv("VSB_TC_QBATCH");
        p.qbatch = qb ? atoi(qb) : 0;
    }
    static unsigned long long* d_stats = nullptr;
    const bool want_stats = getenv("VSB_TC_STATS") != nullptr;
    if (want_stats) {
        if (!d_stats) VSB_CUDA(cudaMalloc((void**)&d_stats, 16 * sizeof(unsigned long long)));
        VSB_CUDA(cudaMemsetAsync(d_stats, 0, 16 * sizeof(unsigned long long), st));
        p.stats = d_stats;
    }
    if (lb_key && ktop != 32) return fail(VS_ERR_INVALID, "tc: lower bound needs the 32-entry list");
    if (mode == TC_F16) return fail(VS_ERR_INVALID, "tc: the fp16 candidate pass is launched by launch_exact_tc_f16");
    const CUtensorMap& tmB_hi = plan.cl == 2 ? tmB.hi_half : tmB.hi;
    const CUtensorMap& tmB_lo = plan.cl == 2 ? tmB.lo_half : tmB.lo;
#define VSB_TC_LAUNCH(KT, MD, LB) VSB_TRY((launch_one<KT, MD, LB>(plan, tmA_hi, tmA_lo, tmB_hi, tmB_lo, p, st)))
#define VSB_TC_CASE(KT)                              \
    case KT:                                         \
        if (mode == TC_TF32X3)                       \
            VSB_TC_LAUNCH(KT, TC_TF32X3, false);     \
        else                                         \
            VSB_TC_LAUNCH(KT, TC_TF32X1, false);     \
        break;
    {
        switch (ktop) {
            VSB_TC_CASE(1)
            VSB_TC_CASE(5)
            VSB_TC_CASE(10)
            VSB_TC_CASE(16)
            case 32:
                if (lb_key) {
                    if (mode == TC_TF32X3)
                        VSB_TC_LAUNCH(32, TC_TF32X3, true);
                    else
                        VSB_TC_LAUNCH(32, TC_TF32X1, true);
                } else {
                    if (mode == TC_TF32X3)
                        VSB_TC_LAUNCH(32, TC_TF32X3, false);
                    else
                        VSB_TC_LAUNCH(32, TC_TF32X1, false);
                }
                break;
            default:
                return fail(VS_ERR_INVALID, "tc: unsupported list size");
        }
    }
#undef VSB_TC_CASE
#undef VSB_TC_LAUNCH
    VSB_CUDA(cudaGetLastError());
    if (want_stats) {
        unsigned long long hs[16];
        VSB_CUDA(cudaMemcpyAsync(hs, d_stats, sizeof(hs), cudaMemcpyDeviceToHost, st));
        VSB_CUDA(cudaStreamSynchronize(st));
        const double chunks = (double)plan.n_mtiles * 4.0 * plan.n_tiles * 4.0;  // warp-chunks of the whole sweep
        fprintf(stderr, "[tc stats] mode=%d ktop=%d splits=%d: warp slow-path entries %llu (%.2f%% of warp-chunks), lane entries %llu, "
                "qualifying %llu, inserts %llu (%.1f per query)\n", mode, ktop, plan.n_splits, hs[0], 100.0 * hs[0] / chunks, hs[1],
                hs[2], hs[3], (double)hs[3] / nq);
    }
    return VS_OK;
}

// The fp16 candidate pass (exact_tc.cuh, TC_F16).  sample = true: walks the tiles tile_off + t * tile_stride, t < plan.n_tiles,
// and writes the group minima smin[plan.n_splits * TC_SAMPLE_GROUPS][nq]; sample = false: walks every tile (plan made for
// n_rows) and appends every (query, row) pair with key < thr[query] to the query's candidate array.
int launch_exact_tc_f16(const CUtensorMap& tmA, const CUtensorMap& tmA_fold, const TcBaseMaps& tmB, int nq, int64_t n_rows,
                        const TcPlan& plan, bool sample, int tile_stride, int tile_off, float* smin, const float* thr,
                        int32_t* cand_cnt, void* cand, int cand_cap, cudaStream_t st) {
    TcParams p{};
    p.nq = nq;
    p.n_tiles = plan.n_tiles;
    p.n_mtiles = plan.n_mtiles;
    p.n_splits = plan.n_splits;
    p.tiles_per_split = plan.tiles_per_split;
    p.n_rem = (int)(n_rows % TC_BN);
    p.n_tiles_real = (int)ceil_div64(n_rows, TC_BN);
    p.tile_stride = sample ? tile_stride : 1;
    p.tile_off = sample ? tile_off : 0;
    p.smin = smin;
    p.thr = thr;
    p.cand_cnt = cand_cnt;
    p.cand = reinterpret_cast<uint2*>(cand);
    p.cand_cap = cand_cap;
    if (plan.cl != 1) return fail(VS_ERR_INVALID, "tc: the fp16 pass runs with independent CTAs");
    if (!sample && p.n_tiles != p.n_tiles_real) return fail(VS_ERR_INVALID, "tc: plan does not match the row count");
    if (sample && (int64_t)(p.n_tiles - 1) * tile_stride + tile_off >= p.n_tiles_real) return fail(VS_ERR_INVALID, "tc: sample tiles out of range");
    {
        const char* e = getenv("VSB_TC_DBG");
        p.dbg = e ? atoi(e) : 0;
        const char* qb = getenv("VSB_TC_QBATCH");
        p.qbatch = qb ? atoi(qb) : 0;
    }
    static unsigned long long* d_stats = nullptr;
    const bool want_stats = !sample && getenv("VSB_TC_STATS") != nullptr;
    if (want_stats) {
        if (!d_stats) VSB_CUDA(cudaMalloc((void**)&d_stats, 16 * sizeof(unsigned long long)));
        VSB_CUDA(cudaMemsetAsync(d_stats, 0, 16 * sizeof(unsigned long long), st));
        p.stats = d_stats;
    }
    if (sample)
        VSB_TRY((launch_one<1, TC_F16, false>(plan, tmA, tmA_fold, tmB.hi, tmB.lo, p, st)));
    else
        VSB_TRY((launch_one<32, TC_F16, false>(plan, tmA, tmA_fold, tmB.hi, tmB.lo, p, st)));
    VSB_CUDA(cudaGetLastError());
    if (want_stats) {
        unsigned long long hs[16];
        VSB_CUDA(cudaMemcpyAsync(hs, d_stats, sizeof(hs), cudaMemcpyDeviceToHost, st));
        VSB_CUDA(cudaStreamSynchronize(st));
        const double chunks = (double)plan.n_mtiles * 4.0 * plan.n_tiles * 4.0;  // warp-chunks of the whole sweep
        fprintf(stderr, "[tc stats] f16 filter pass, splits=%d: warp hand-offs %llu (%.2f%% of warp-chunks), queue entries %llu (%.1f per "
                "query), keeper batches %llu (%.1f entries each)\n", plan.n_splits, hs[0], 100.0 * hs[0] / chunks, hs[1], (double)hs[1] / nq,
                hs[5], hs[5] ? (double)hs[4] / hs[5] : 0.0);
    }
    return VS_OK;
}

}  // namespace vsb
