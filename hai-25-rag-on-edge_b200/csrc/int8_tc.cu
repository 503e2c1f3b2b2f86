// INT8 brute force (the QNN path of the reference re-expressed for B200):
//   K4  quantize_u8_kernel   QnnRunner.cpp:13-55 quantize_buffer_neon: u8 = sat(trunc(x * (1/scale) + 0.5))
//   K5  int8_tc_kernel       the HTP MatMul (QnnRunner.cpp:628 graphExecute; graph = create_model.py:57-87, u8 in/out
//                            per quant_overrides.json) as tcgen05 kind::i8 u8 x u8 -> s32 tiles, fused with the
//                            requantisation to u8 and find_top_k_int8 (main.cpp:36-57): the [B x N] u8 score matrix
//                            never exists.
// Requantisation rule (defined by this project; the HTP's is proprietary):
//   score = sat_u8(floor(fl(fl(acc) * m) + 0.5)),  m = fl(fl(s_in * s_w) / s_out), ranked (score desc, id asc).
// The rule is monotone in acc, so the epilogue filters raw accumulators against the smallest accumulator value
// that still reaches the score currently needed (one integer compare per element) and requantises only the rare
// candidates.
//
// Tile: 128 queries x 128 base rows, K = 128 u8 = one 128-byte swizzle row per vector: one TMA box (16 KB) and four
// K=32 MMAs per tile.  Same warp roles / unit decomposition / threshold sharing as exact_tc.cuh.
#include <cuda.h>

#include <algorithm>

#include "kernels.cuh"
#include "vsb_common.cuh"

namespace vsb {

constexpr int I8_BM = 128, I8_BN = 128;
constexpr int I8_TILE_BYTES = 128 * 128;  // 128 rows x 128 B
constexpr int I8_NSTAGE = 8;
constexpr int I8_NACC = 4;
constexpr int I8_EPI_GROUPS = 2;
constexpr int I8_GCOLS = I8_BN / I8_EPI_GROUPS;
constexpr int I8_THREADS = 64 + 128 * I8_EPI_GROUPS;
constexpr int I8_SMEM = I8_TILE_BYTES * (1 + I8_NSTAGE) + 1024 + 1024;
constexpr int I8_THR_REFRESH = 8;

__device__ __forceinline__ int requant_u8(int32_t acc, float m) {
    const float t = floorf(__fadd_rn(__fmul_rn((float)acc, m), 0.5f));
    return t < 0.f ? 0 : (t > 255.f ? 255 : (int)t);
}
// smallest accumulator whose requantised score is >= s (the rule is monotone non-decreasing in acc)
__device__ __forceinline__ int32_t min_acc_for_score(int s, float m) {
    if (s <= 0) return 0;
    if (s > 255) return 0x7fffffff;
    int32_t a = (int32_t)ceilf(((float)s - 0.5f) / m);
    if (a < 0) a = 0;
    while (a > 0 && requant_u8(a - 1, m) >= s) --a;
    while (requant_u8(a, m) < s) ++a;
    return a;
}
__device__ __forceinline__ int32_t acc_bound_for_thr(float thr, float m) {
    // a candidate needs key = -score < thr  <=>  score > -thr
    if (!(thr < __int_as_float(0x7f800000))) return 0;
    const float need = floorf(-thr) + 1.0f;
    return min_acc_for_score(need > 256.f ? 256 : (need < 0.f ? 0 : (int)need), m);
}

struct I8Params {
    int32_t* gthr;
    float* part_key;   // [n_splits*2][nq][KTOP], key = -score
    int32_t* part_id;
    float m;           // requantisation multiplier
    int nq;
    int64_t n;
    int n_tiles, n_mtiles, n_splits, tiles_per_split;
    int rep;           // small batches (nq <= 128 / rep): the query tile holds `rep` copies of the queries, copy r folds the
                       // 32-column chunks c with c % rep == r of every base tile, so all four TMEM lane quadrants (= all
                       // four SM sub-partitions) share the epilogue instead of one; 1 = no replication
};

template <int KTOP, int REP>
__global__ void __launch_bounds__(I8_THREADS, 1)
int8_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const I8Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + I8_TILE_BYTES;
    uint64_t* bars = (uint64_t*)(sB + I8_NSTAGE * I8_TILE_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = full + I8_NSTAGE;
    uint64_t* acc_full = empty + I8_NSTAGE;
    uint64_t* acc_empty = acc_full + I8_NACC;
    uint64_t* a_full = acc_empty + I8_NACC;
    uint64_t* a_empty = a_full + 1;
    uint32_t* tmem_slot = (uint32_t*)(a_empty + 1);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < I8_NSTAGE; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < I8_NACC; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 4 * I8_EPI_GROUPS);
        }
        mbar_init(a_full, 1);
        mbar_init(a_empty, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, I8_NACC * I8_BN);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int n_units = p.n_mtiles * p.n_splits;

    if (warp == 0) {
        // whole warp walks the loop (uniform addresses), one elected lane issues the copies
        const bool leader = elect_one();
        {
            if (leader) {
                tma_prefetch_desc(&tmA);
                tma_prefetch_desc(&tmB);
            }
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++it) {
                const int m_tile = unit % p.n_mtiles;
                const int split = unit / p.n_mtiles;
                mbar_wait(a_empty, (uint32_t)((it & 1) ^ 1));
                if (leader) {
                    mbar_expect_tx(a_full, (uint32_t)I8_TILE_BYTES);
                    tma_load_2d(sA, &tmA, a_full, 0, m_tile * I8_BM);
                }
                const int t0 = split * p.tiles_per_split;
                const int t1 = min(t0 + p.tiles_per_split, p.n_tiles);
                for (int t = t0; t < t1; ++t) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    if (leader) {
                        mbar_expect_tx(&full[stage], (uint32_t)I8_TILE_BYTES);
                        tma_load_2d(sB + stage * I8_TILE_BYTES, &tmB, &full[stage], 0, t * I8_BN);
                    }
                    if (++stage == I8_NSTAGE) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        const bool leader = elect_one();  // the same lane issues every MMA and commit
        {
            constexpr uint32_t idesc = umma_idesc(kIdescCS32, kIdescU8, I8_BM, I8_BN);
            const uint64_t a_desc = umma_desc_sw128(smem_u32(sA));
            const uint32_t sB_u = smem_u32(sB);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            int it = 0;
            for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++it) {
                const int split = unit / p.n_mtiles;
                const int t0 = split * p.tiles_per_split;
                const int t1 = min(t0 + p.tiles_per_split, p.n_tiles);
                mbar_wait(a_full, (uint32_t)(it & 1));
                tc_fence_after();
                for (int t = t0; t < t1; ++t) {
                    mbar_wait(&acc_empty[acc], acc_phase ^ 1);
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * I8_BN);
                    const uint64_t b_desc = umma_desc_sw128(sB_u + stage * I8_TILE_BYTES);
                    if (leader) {
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)  // K = 32 u8 = 32 B per step
                            tc_mma_i8(d_tmem, a_desc + 2 * ks, b_desc + 2 * ks, idesc, ks > 0 ? 1u : 0u);
                        tc_commit(&empty[stage]);
                        tc_commit(&acc_full[acc]);
                    }
                    if (++stage == I8_NSTAGE) { stage = 0; phase ^= 1; }
                    if (++acc == I8_NACC) { acc = 0; acc_phase ^= 1; }
                }
                if (leader) tc_commit(a_empty);
            }
        }
    } else {
        const int quad = warp & 3;
        const int grp = (warp - 2) >> 2;
        const int row = quad * 32 + lane;
        const float INF = __int_as_float(0x7f800000);
        constexpr int CH = I8_GCOLS / 32;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
            const int m_tile = unit % p.n_mtiles;
            const int split = unit / p.n_mtiles;
            const int t0 = split * p.tiles_per_split;
            const int t1 = min(t0 + p.tiles_per_split, p.n_tiles);
            constexpr int rows_per_copy = I8_BM / REP;
            const int rix = row / rows_per_copy;                       // which copy of the queries this row belongs to
            const int q = m_tile * I8_BM + (row - rix * rows_per_copy);
            const bool valid = q < p.nq;
            const bool quad_live = m_tile * I8_BM + (quad * 32) % rows_per_copy < p.nq;  // idle quadrants only handshake
            RegTopK<KTOP> top;
            top.init();
            float cap = INF, thr = INF;
            int32_t bound = 0;
            for (int t = t0; t < t1; ++t) {
                const int rel = (t - t0) & (I8_THR_REFRESH - 1);
                if (rel == 0 && valid) {
                    cap = fminf(cap, ordered_to_float(__ldcg(p.gthr + q)));
                    thr = fminf(top.threshold(), cap + 1.0f);  // keys are integers: next key above cap
                    if (!(cap < 1e30f)) thr = top.threshold();
                    bound = acc_bound_for_thr(thr, p.m);
                }
                mbar_wait(&acc_full[acc], acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * I8_BN + grp * I8_GCOLS);
                uint32_t r[2][32];
                auto fold = [&](const uint32_t (&rr)[32], const int col0) {
                    int32_t a[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) a[j] = (int32_t)rr[j];
                    int32_t mx[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) mx[j] = max(a[j], a[j + 16]);
#pragma unroll
                    for (int w = 8; w >= 1; w >>= 1)
#pragma unroll
                        for (int j = 0; j < w; ++j) mx[j] = max(mx[j], mx[j + w]);
                    if (mx[0] >= bound) {
                        uint32_t mask = 0;
#pragma unroll
                        for (int j = 0; j < 32; ++j) mask |= (a[j] >= bound) ? (1u << j) : 0u;
                        while (mask) {
                            const int j = __ffs(mask) - 1;
                            mask &= mask - 1;
                            const int32_t av = select32(a, j);
                            if (av >= bound && (int64_t)(col0 + j) < p.n) {
                                const float key = -(float)requant_u8(av, p.m);
                                if (key < thr) {
                                    top.insert(key, col0 + j);
                                    thr = fminf(thr, top.threshold());
                                    bound = acc_bound_for_thr(thr, p.m);
                                }
                            }
                        }
                    }
                };
                if (REP == 1) {
                    if (quad_live) tmem_ld32(taddr, r[0]);
#pragma unroll
                    for (int c = 0; c < CH; ++c) {
                        if (!quad_live) break;
                        tc_wait_ld();
                        if (c + 1 < CH) tmem_ld32(taddr + (c + 1) * 32, r[(c + 1) & 1]);
                        fold(r[c & 1], t * I8_BN + grp * I8_GCOLS + c * 32);
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < CH; ++c) {  // only the chunks of this copy
                        if (!quad_live || ((grp * CH + c) & (REP - 1)) != rix) continue;
                        tmem_ld32(taddr + c * 32, r[0]);
                        tc_wait_ld();
                        fold(r[0], t * I8_BN + grp * I8_GCOLS + c * 32);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[acc]);
                if (++acc == I8_NACC) { acc = 0; acc_phase ^= 1; }
                if (valid && rel == I8_THR_REFRESH - 1 && top.threshold() < cap) {
                    atomicMin(p.gthr + q, float_to_ordered(top.threshold()));
                    cap = top.threshold();
                }
            }
            if (valid) {
                if (top.threshold() < cap) atomicMin(p.gthr + q, float_to_ordered(top.threshold()));
                const size_t list = ((size_t)split * I8_EPI_GROUPS + grp) * REP + rix;
                float* pk = p.part_key + (list * p.nq + q) * KTOP;
                int32_t* pi = p.part_id + (list * p.nq + q) * KTOP;
#pragma unroll
                for (int i = 0; i < KTOP; ++i) {
                    pk[i] = top.key[i];
                    pi[i] = top.id[i];
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, I8_NACC * I8_BN);
}

// ------------------------------------------------------------------------------------------------
// K5, large batches: TWO query tiles per unit share every base tile.  With one query tile per CTA a 16 KB base tile feeds only
// four MMAs (~256 tensor cycles): the TMA fill rate of an SM (~64 B/clk, profiles/r2_sm_limits_tmem_tma.txt) equals the MMA
// rate and the two contend for shared-memory bandwidth.  With a pair, a tile feeds eight MMAs; epilogue group g owns query
// tile g of the pair (all 128 columns of its accumulator), TMEM holds two pair-buffers of 2 x 128 columns, and a unit writes
// ONE list per query instead of two.
// ------------------------------------------------------------------------------------------------
constexpr int I8P_NSTAGE = 8;
constexpr int I8P_SMEM = I8_TILE_BYTES * (2 + I8P_NSTAGE) + 1024 + 1024;

template <int KTOP>
__global__ void __launch_bounds__(I8_THREADS, 1)
int8_tc_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const I8Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;  // two query tiles
    uint8_t* sB = smem + 2 * I8_TILE_BYTES;
    uint64_t* bars = (uint64_t*)(sB + I8P_NSTAGE * I8_TILE_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = full + I8P_NSTAGE;
    uint64_t* acc_full = empty + I8P_NSTAGE;   // [2] pair-buffers
    uint64_t* acc_empty = acc_full + 2;
    uint64_t* a_full = acc_empty + 2;
    uint64_t* a_empty = a_full + 1;
    uint32_t* tmem_slot = (uint32_t*)(a_empty + 1);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < I8P_NSTAGE; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 4 * I8_EPI_GROUPS);
        }
        mbar_init(a_full, 1);
        mbar_init(a_empty, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int n_pairs = p.n_mtiles;  // the plan counts tile PAIRS as its query-tile columns
    const int n_units = n_pairs * p.n_splits;

    if (warp == 0) {
        const bool leader = elect_one();
        if (leader) {
            tma_prefetch_desc(&tmA);
            tma_prefetch_desc(&tmB);
        }
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++it) {
            const int pair = unit % n_pairs;
            const int split = unit / n_pairs;
            mbar_wait(a_empty, (uint32_t)((it & 1) ^ 1));
            if (leader) {
                mbar_expect_tx(a_full, (uint32_t)(2 * I8_TILE_BYTES));
                tma_load_2d(sA, &tmA, a_full, 0, (2 * pair) * I8_BM);
                tma_load_2d(sA + I8_TILE_BYTES, &tmA, a_full, 0, (2 * pair + 1) * I8_BM);  // beyond nq: zero fill
            }
            const int t0 = split * p.tiles_per_split;
            const int t1 = min(t0 + p.tiles_per_split, p.n_tiles);
            for (int t = t0; t < t1; ++t) {
                mbar_wait(&empty[stage], phase ^ 1);
                if (leader) {
                    mbar_expect_tx(&full[stage], (uint32_t)I8_TILE_BYTES);
                    tma_load_2d(sB + stage * I8_TILE_BYTES, &tmB, &full[stage], 0, t * I8_BN);
                }
                if (++stage == I8P_NSTAGE) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        const bool leader = elect_one();
        constexpr uint32_t idesc = umma_idesc(kIdescCS32, kIdescU8, I8_BM, I8_BN);
        const uint64_t a0_desc = umma_desc_sw128(smem_u32(sA));
        const uint64_t a1_desc = umma_desc_sw128(smem_u32(sA + I8_TILE_BYTES));
        const uint32_t sB_u = smem_u32(sB);
        int stage = 0, buf = 0;
        uint32_t phase = 0, buf_phase = 0;
        int it = 0;
        for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++it) {
            const int split = unit / n_pairs;
            const int t0 = split * p.tiles_per_split;
            const int t1 = min(t0 + p.tiles_per_split, p.n_tiles);
            mbar_wait(a_full, (uint32_t)(it & 1));
            tc_fence_after();
            for (int t = t0; t < t1; ++t) {
                mbar_wait(&acc_empty[buf], buf_phase ^ 1);
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * 2 * I8_BN);
                const uint64_t b_desc = umma_desc_sw128(sB_u + stage * I8_TILE_BYTES);
                if (leader) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) tc_mma_i8(d_tmem, a0_desc + 2 * ks, b_desc + 2 * ks, idesc, ks > 0 ? 1u : 0u);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) tc_mma_i8(d_tmem + I8_BN, a1_desc + 2 * ks, b_desc + 2 * ks, idesc, ks > 0 ? 1u : 0u);
                    tc_commit(&empty[stage]);
                    tc_commit(&acc_full[buf]);
                }
                if (++stage == I8P_NSTAGE) { stage = 0; phase ^= 1; }
                if (++buf == 2) { buf = 0; buf_phase ^= 1; }
            }
            if (leader) tc_commit(a_empty);
        }
    } else {
        const int quad = warp & 3;
        const int grp = (warp - 2) >> 2;  // = which query tile of the pair
        const int row = quad * 32 + lane;
        const float INF = __int_as_float(0x7f800000);
        constexpr int CH = I8_BN / 32;
        int buf = 0;
        uint32_t buf_phase = 0;
        for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
            const int pair = unit % n_pairs;
            const int split = unit / n_pairs;
            const int t0 = split * p.tiles_per_split;
            const int t1 = min(t0 + p.tiles_per_split, p.n_tiles);
            const int m_tile = 2 * pair + grp;
            const int q = m_tile * I8_BM + row;
            const bool valid = q < p.nq;
            const bool quad_live = m_tile * I8_BM + quad * 32 < p.nq;  // idle quadrants only handshake
            RegTopK<KTOP> top;
            top.init();
            float cap = INF, thr = INF;
            int32_t bound = 0;
            for (int t = t0; t < t1; ++t) {
                const int rel = (t - t0) & (I8_THR_REFRESH - 1);
                if (rel == 0 && valid) {
                    cap = fminf(cap, ordered_to_float(__ldcg(p.gthr + q)));
                    thr = fminf(top.threshold(), cap + 1.0f);  // keys are integers: next key above cap
                    if (!(cap < 1e30f)) thr = top.threshold();
                    bound = acc_bound_for_thr(thr, p.m);
                }
                mbar_wait(&acc_full[buf], buf_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * 2 * I8_BN + grp * I8_BN);
                uint32_t r[2][32];
                auto fold = [&](const uint32_t (&rr)[32], const int col0) {
                    int32_t a[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) a[j] = (int32_t)rr[j];
                    int32_t mx[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) mx[j] = max(a[j], a[j + 16]);
#pragma unroll
                    for (int w = 8; w >= 1; w >>= 1)
#pragma unroll
                        for (int j = 0; j < w; ++j) mx[j] = max(mx[j], mx[j + w]);
                    if (mx[0] >= bound) {
                        uint32_t mask = 0;
#pragma unroll
                        for (int j = 0; j < 32; ++j) mask |= (a[j] >= bound) ? (1u << j) : 0u;
                        while (mask) {
                            const int j = __ffs(mask) - 1;
                            mask &= mask - 1;
                            const int32_t av = select32(a, j);
                            if (av >= bound && (int64_t)(col0 + j) < p.n) {
                                const float key = -(float)requant_u8(av, p.m);
                                if (key < thr) {
                                    top.insert(key, col0 + j);
                                    thr = fminf(thr, top.threshold());
                                    bound = acc_bound_for_thr(thr, p.m);
                                }
                            }
                        }
                    }
                };
                if (quad_live) tmem_ld32(taddr, r[0]);
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    if (!quad_live) break;
                    tc_wait_ld();
                    if (c + 1 < CH) tmem_ld32(taddr + (c + 1) * 32, r[(c + 1) & 1]);
                    fold(r[c & 1], t * I8_BN + c * 32);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[buf]);
                if (++buf == 2) { buf = 0; buf_phase ^= 1; }
                if (valid && rel == I8_THR_REFRESH - 1 && top.threshold() < cap) {
                    atomicMin(p.gthr + q, float_to_ordered(top.threshold()));
                    cap = top.threshold();
                }
            }
            if (valid) {
                if (top.threshold() < cap) atomicMin(p.gthr + q, float_to_ordered(top.threshold()));
                float* pk = p.part_key + ((size_t)split * p.nq + q) * KTOP;  // one list per (split, query)
                int32_t* pi = p.part_id + ((size_t)split * p.nq + q) * KTOP;
#pragma unroll
                for (int i = 0; i < KTOP; ++i) {
                    pk[i] = top.key[i];
                    pi[i] = top.id[i];
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// K4 quantiser, float->u8 score conversion, raw score matrix (tests), max reduction (weight scale)
// ------------------------------------------------------------------------------------------------
__global__ void quantize_u8_kernel(const float* __restrict__ src, int64_t count, float inv_scale, uint8_t* __restrict__ dst) {
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < count; i += (int64_t)gridDim.x * blockDim.x * 4) {
        uint32_t packed = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (i + e < count) {
                const float t = __fadd_rn(__fmul_rn(__ldg(src + i + e), inv_scale), 0.5f);  // multiply, then add (not fused)
                int v;
                if (t != t) v = 0;
                else if (t >= 2147483648.0f) v = 0x7fffffff;
                else if (t <= -2147483648.0f) v = (int)0x80000000;
                else v = __float2int_rz(t);  // vcvtq_s32_f32 truncates toward zero
                v = v < 0 ? 0 : (v > 255 ? 255 : v);
                packed |= (uint32_t)v << (8 * e);
            }
        }
        if (i + 3 < count) {
            *reinterpret_cast<uint32_t*>(dst + i) = packed;
        } else {
            for (int e = 0; e < 4 && i + e < count; ++e) dst[i + e] = (uint8_t)(packed >> (8 * e));
        }
    }
}

__global__ void scores_f32_to_u8_kernel(const float* __restrict__ src, int64_t count, uint8_t* __restrict__ dst) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = src[i];
        dst[i] = (v >= 0.f && v <= 255.f) ? (uint8_t)v : 0;  // padding (-inf) -> 0
    }
}

__global__ void int8_scores_kernel(const uint8_t* __restrict__ base, int64_t n, const uint8_t* __restrict__ q, int64_t nq, float m,
                                   uint8_t* __restrict__ out) {
    const int64_t total = n * nq;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t qi = e / n, j = e % n;
        const uint32_t* a = reinterpret_cast<const uint32_t*>(q + qi * 128);
        const uint32_t* b = reinterpret_cast<const uint32_t*>(base + j * 128);
        uint32_t acc = 0;
#pragma unroll 8
        for (int w = 0; w < 32; ++w) acc = __dp4a(__ldg(a + w), __ldg(b + w), acc);  // unsigned x unsigned
        out[e] = (uint8_t)requant_u8((int32_t)acc, m);
    }
}

__global__ void max_f32_kernel(const float* __restrict__ x, int64_t count, float* __restrict__ out) {
    float mx = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
        mx = fmaxf(mx, __ldg(x + i));
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(mx));  // non-negative floats order as ints
}

int launch_quantize_u8(const float* src, int64_t count, float inv_scale, uint8_t* dst, cudaStream_t st) {
    if (count <= 0) return VS_OK;
    const int64_t blocks = std::min<int64_t>(ceil_div64(count, 4 * 256), 148 * 16);
    quantize_u8_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, count, inv_scale, dst);
    VSB_CUDA(cudaGetLastError());
    return VS_OK;
}
int launch_scores_to_u8(const float* src, int64_t count, uint8_t* dst, cudaStream_t st) {
    if (count <= 0) return VS_OK;
    scores_f32_to_u8_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(count, 256), 148 * 8), 256, 0, st>>>(src, count, dst);
    VSB_CUDA(cudaGetLastError());
    return VS_OK;
}
int launch_int8_scores(const uint8_t* base, int64_t n, const uint8_t* q, int64_t nq, float m, uint8_t* out, cudaStream_t st) {
    if (n * nq <= 0) return VS_OK;
    int8_scores_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(n * nq, 256), 148 * 16), 256, 0, st>>>(base, n, q, nq, m, out);
    VSB_CUDA(cudaGetLastError());
    return VS_OK;
}
int launch_max_f32(const float* x, int64_t count, float* out_zeroed, cudaStream_t st) {
    if (count <= 0) return VS_OK;
    max_f32_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(count, 256), 148 * 8), 256, 0, st>>>(x, count, out_zeroed);
    VSB_CUDA(cudaGetLastError());
    return VS_OK;
}

template <int KTOP>
static int int8_set_attr_k() {
    VSB_CUDA(cudaFuncSetAttribute(int8_tc_kernel<KTOP, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, I8_SMEM));
    VSB_CUDA(cudaFuncSetAttribute(int8_tc_kernel<KTOP, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, I8_SMEM));
    VSB_CUDA(cudaFuncSetAttribute(int8_tc_kernel<KTOP, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, I8_SMEM));
    VSB_CUDA(cudaFuncSetAttribute(int8_tc_pair_kernel<KTOP>, cudaFuncAttributeMaxDynamicSharedMemorySize, I8P_SMEM));
    return VS_OK;
}
int int8_set_attributes() {
    VSB_TRY(int8_set_attr_k<1>());
    VSB_TRY(int8_set_attr_k<5>());
    VSB_TRY(int8_set_attr_k<10>());
    VSB_TRY(int8_set_attr_k<16>());
    VSB_TRY(int8_set_attr_k<32>());
    return VS_OK;
}
int int8_lists_per_split() { return I8_EPI_GROUPS; }

// pair mode (rep == 0): plan.n_mtiles counts PAIRS of query tiles; one list per (split, query)
int launch_int8_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, int32_t* gthr, float m, int nq, int64_t n, const TcPlan& plan,
                   int ktop, int rep, float* part_key, int32_t* part_id, cudaStream_t st) {
    if (rep == 0) {
        I8Params pp{gthr, part_key, part_id, m, nq, n, plan.n_tiles, plan.n_mtiles, plan.n_splits, plan.tiles_per_split, 1};
#define VSB_I8P_LAUNCH(KT) \
    case KT: int8_tc_pair_kernel<KT><<<plan.grid, I8_THREADS, I8P_SMEM, st>>>(tmA, tmB, pp); break;
        switch (ktop) {
            VSB_I8P_LAUNCH(1)
            VSB_I8P_LAUNCH(5)
            VSB_I8P_LAUNCH(10)
            VSB_I8P_LAUNCH(16)
            VSB_I8P_LAUNCH(32)
            default: return fail(VS_ERR_UNSUPPORTED, "INT8 search: k > 32 is not implemented");
        }
#undef VSB_I8P_LAUNCH
        VSB_CUDA(cudaGetLastError());
        return VS_OK;
    }
    if (rep != 1 && rep != 2 && rep != 4) return fail(VS_ERR_INVALID, "int8: replication must be 1, 2 or 4");
    if (rep > 1 && nq > I8_BM / rep) return fail(VS_ERR_INVALID, "int8: too many queries for the replication factor");
    I8Params p{gthr, part_key, part_id, m, nq, n, plan.n_tiles, plan.n_mtiles, plan.n_splits, plan.tiles_per_split, rep};
#define VSB_I8_LAUNCH(KT)                                                                     \
    case KT:                                                                                  \
        if (rep == 1) int8_tc_kernel<KT, 1><<<plan.grid, I8_THREADS, I8_SMEM, st>>>(tmA, tmB, p);      \
        else if (rep == 2) int8_tc_kernel<KT, 2><<<plan.grid, I8_THREADS, I8_SMEM, st>>>(tmA, tmB, p); \
        else int8_tc_kernel<KT, 4><<<plan.grid, I8_THREADS, I8_SMEM, st>>>(tmA, tmB, p);               \
        break;
    switch (ktop) {
        VSB_I8_LAUNCH(1)
        VSB_I8_LAUNCH(5)
        VSB_I8_LAUNCH(10)
        VSB_I8_LAUNCH(16)
        VSB_I8_LAUNCH(32)
        default: return fail(VS_ERR_UNSUPPORTED, "INT8 search: k > 32 is not implemented");
    }
#undef VSB_I8_LAUNCH
    VSB_CUDA(cudaGetLastError());
    return VS_OK;
}

}  // namespace vsb
