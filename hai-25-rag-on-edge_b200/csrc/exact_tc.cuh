// K1 — fused distance GEMM + selection on the 5th-generation tensor cores (tcgen05, accumulators in TMEM, operands
// fed by TMA), replacing the reference's per-query  cblas_sgemm(M=1) -> distance loop -> select_topk
// (cpu/cpu_baseline.cpp:222-248) for a whole batch of queries.  The Q x N distance matrix never exists in
// memory: each epilogue thread owns one query row of the accumulator tile (a TMEM lane), 32 columns at a time.
//
// Mapping: queries -> M (TMEM lanes, 128 per CTA), base rows -> N (TMEM columns, 128 per tile), K = dim = 128.
//   1xTF32 : 16 MMAs per tile (exact when operands are TF32-representable, e.g. integer SIFT data)
//   3xTF32 : q.x ~= q_lo.x_hi + q_hi.x_hi + q_hi.x_lo per k-block, 48 MMAs per tile, fp32 accumulate
//   F16    : 8 kind::f16 MMAs on power-of-two scaled fp16 copies + 1 for the norm block: the accumulator is the key
// TF32 modes: ranking key bn - 2*dot, a sorted register list per (unit, epilogue group, query), thresholds shared between
// groups and CTAs; the merge kernel recomputes the reported distances of the survivors in exact fp32 (kernels.cu).
// F16 mode: a candidate generator built as a threshold filter — sample pass (KTOP == 1: group minima over one base tile in
// 16), per-query threshold (tc_select_thr_kernel), filter pass (KTOP == 32: every row below the threshold is appended to the
// query's candidate array, two query tiles per unit), filter merge + certificate (kernels.cu); see DESIGN.md §3 / §4.
//
// Work decomposition: unit = (query tile [pair], base split); units are ordered split-major so that the CTAs resident
// at any moment sweep the same base panel (it streams from HBM once and is re-read from L2 by the other query tiles).
//
// Warp roles: warp 0 = TMA producer (one lane), warp 1 = TMEM allocator + MMA issuer (one lane), warps 2-3 idle,
// then three epilogue warpgroups that take accumulator tiles round-robin (TMEM lane quadrant = warp % 4), and in F16 mode
// a fifth warpgroup of four keeper warps (one per quadrant) that drain the epilogue warps' candidate queues.
#pragma once
#include <cuda.h>

#include "vsb_common.cuh"

namespace vsb {

constexpr int TC_BM = 128;        // queries per CTA tile
constexpr int TC_BN = 128;        // base rows per accumulator tile
constexpr int TC_KB_BYTES = 128 * 128;  // one k-block of a 128-row operand tile: 128 rows x 128 B
constexpr int TC_NACC = 4;        // accumulator buffers in TMEM (4 x 128 columns = all 512)
constexpr int TC_EPI_GROUPS = 3;  // epilogue warpgroups; tiles rotate over them
constexpr int TC_THREADS = 128 + 128 * TC_EPI_GROUPS;  // warpgroup 0: TMA producer, MMA issuer, two idle warps
constexpr int TC_REGS_CTRL = 80, TC_REGS_EPI = 144;   // setmaxnreg budgets: 128*80 + 384*144 == 64K registers
// TC_F16 adds a fifth warpgroup of four keeper warps (one per TMEM lane quadrant): 640 threads start with 96 registers
// (61440); control drops to 40, the keepers to 64, the three epilogue warpgroups rise to 120 (128*40 + 128*64 + 384*120 <=
// 61440: setmaxnreg.inc can only take what the CTA itself released)
constexpr int TC_THREADS_Q = TC_THREADS + 128;
constexpr int TC_REGS_CTRL_Q = 40, TC_REGS_KEEP_Q = 64, TC_REGS_EPI_Q = 120;
constexpr int TC_SAMPLE_SUB = 8;                                  // sample pass: running minima per thread and unit
constexpr int TC_SAMPLE_GROUPS = TC_EPI_GROUPS * TC_SAMPLE_SUB;  // groups of sampled rows per unit
#ifndef VSB_TC_QN
#define VSB_TC_QN 32
#endif
#ifndef VSB_TC_F16_STAGES
#define VSB_TC_F16_STAGES 3
#endif
#ifndef VSB_TC_F16_PAIR_STAGES
#define VSB_TC_F16_PAIR_STAGES 2
#endif
constexpr int TC_QN = VSB_TC_QN;      // candidate-queue entries per EPILOGUE WARP (one queue per (quadrant, group): no slot atomics)
constexpr int TC_QBATCH = 24;         // queued entries that make a keeper batch worthwhile
constexpr int TC_QENTRY = 144;        // bytes per entry: the 32 keys of a chunk + {query, first column, threshold, -}
constexpr int TC_THR_REFRESH = 4; // tiles of one group between reads of the shared threshold (power of two)

// Operand arithmetic of the tensor-core pass
//   TC_TF32X1  kind::tf32, one product                    (exact when operands are TF32-representable)
//   TC_TF32X3  kind::tf32, hi/lo split, three products    (fp32-faithful)
//   TC_F16     kind::f16 on power-of-two-scaled fp16 copies: a CANDIDATE generator whose key error is bounded
//              (api.cu derives the bound); the merge kernel refines the candidates in exact fp32 and certifies,
//              per query, that no true neighbour can have been missed.
enum TcMode : int { TC_TF32X1 = 0, TC_TF32X3 = 1, TC_F16 = 2 };

constexpr int TC_FOLD_BYTES = 128 * 32;  // the K = 16 norm block of a 128-row operand tile: 128 rows x 32 B (SWIZZLE_32B)

// PAIR (TC_F16 filter pass): two query tiles per unit share every base tile.  With one, a 36 KB base tile feeds 9 MMAs (576
// tensor cycles): the SM's TMA fill rate (~64 B/clk, profiles/r2_sm_limits_tmem_tma.txt) equals the MMA rate and the two
// contend for shared-memory bandwidth; with a pair it feeds 18.
template <int MODE, bool PAIR = false>
struct TcSmem {
    static constexpr bool SPLIT3 = MODE == TC_TF32X3;
    static constexpr int NKB = MODE == TC_F16 ? 2 : 4;  // 128-byte k-blocks per row (128 fp16 = 256 B, 128 fp32 = 512 B)
    // TC_F16 folds the norm term into the MMA: one extra K = 16 block on both operands (kernels.cuh, TC_FOLD_COLS)
    static constexpr bool FOLD = MODE == TC_F16;
    static constexpr int A_BYTES = (SPLIT3 || PAIR ? 2 : 1) * NKB * TC_KB_BYTES + (FOLD ? TC_FOLD_BYTES : 0);
    // TC_F16 keeps ONE candidate list per query row in the registers of dedicated list-keeper warps that are fed
    // through per-quadrant shared-memory queues; the other modes keep three per-thread register lists per row in the
    // epilogue warps themselves (their query tile leaves no room for the queues: 64 / 128 KB)
    static constexpr bool SMEM_LIST = MODE == TC_F16;
    static constexpr int THREADS = SMEM_LIST ? TC_THREADS_Q : TC_THREADS;
    // ring of base operand stages: TC_F16 one whole tile per stage (two k-blocks + the norm block, 36 KB), the TF32 modes one
    // k-block per stage
    static constexpr int NSTAGE = MODE == TC_F16 ? (PAIR ? VSB_TC_F16_PAIR_STAGES : VSB_TC_F16_STAGES) : (SPLIT3 ? 5 : 8);
    static constexpr int BSTAGE = MODE == TC_F16 ? 2 * TC_KB_BYTES + TC_FOLD_BYTES : TC_KB_BYTES;
    static constexpr int B_BYTES = NSTAGE * BSTAGE;
    static constexpr int NORM_BYTES = FOLD ? 0 : TC_NACC * TC_BN * 4;
    static constexpr int PUB_BYTES = SMEM_LIST ? 0 : TC_EPI_GROUPS * TC_BM * 8;  // per (group, query row): {key, unit tag}
    static constexpr int STAGE_BYTES = SMEM_LIST ? 4 * TC_EPI_GROUPS * TC_QN * TC_QENTRY : 0;  // candidate queues, one per epilogue warp
    static constexpr int AUX_BYTES = SMEM_LIST ? 4 * TC_EPI_GROUPS * TC_QN * 4 + 128 : 0;  // ready words, final tails, heads
    static constexpr int BAR_BYTES = 1024;
    static constexpr int TOTAL = A_BYTES + B_BYTES + NORM_BYTES + PUB_BYTES + STAGE_BYTES + AUX_BYTES + BAR_BYTES +
                                 1024;  // + slack for 1024-B alignment
    static_assert(TOTAL <= 227 * 1024, "shared memory budget");
};

struct TcParams {
    const float* bnorm;  // [n_tiles*128], +inf beyond n
    const float* lb_key; // optional per-query exclusive lower bound (multi-pass k > 32), or nullptr
    const int32_t* lb_id;
    int32_t* gthr;       // [nq] shared thresholds (order-preserving int encoding), preset to a huge value (TF32 modes)
    float* part_key;     // [n_splits*TC_EPI_GROUPS][nq][KTOP]
    int32_t* part_id;
    int nq;
    int n_tiles;         // ceil(n / 128)
    int n_mtiles;        // ceil(nq / 128)
    int n_splits;
    int tiles_per_split;
    int n_rem;           // rows of the last base tile (n % 128, 0 = full): TC_F16 masks the columns beyond (the norm block cannot
                         // carry +inf)
    const float* key_scale_ptr;  // unused (TC_F16 keys stay in accumulator units: key = acc * 2 / (s_q * s_b), kernels.cuh)
    unsigned long long* stats;  // debug counters (VSB_TC_STATS) or nullptr: [0] warp slow-path entries, [1] lane entries,
                         // [2] qualifying elements, [3] insertions
    int qbatch;          // TC_F16: queued rows that make a batch worth folding (0 = default)
    // TC_F16 (threshold-filter candidate pass, see the kernel comment).  The base tiles walked are tile_off + t * tile_stride
    // for t < n_tiles (the sample pass walks every tile_stride-th tile; n_tiles_real = ceil(n / 128) locates the ragged tile)
    int tile_stride, tile_off, n_tiles_real;
    float* smin;         // sample pass (KTOP == 1): [n_splits * TC_SAMPLE_GROUPS][nq] smallest key of each group of sampled rows
    const float* thr;    // filter pass (KTOP == 32): [nq] per-query key thresholds; every row with key < thr is a candidate
    int32_t* cand_cnt;   // [nq] candidates found (may exceed cand_cap: overflow, the query is then not certified)
    uint2* cand;         // [nq][cand_cap] {key bits, local row id}, unordered
    int cand_cap;
    // IVF = true: work items {first pair, pairs, first row, rows}, their number, pair -> query * nprobe + probe slot
    const int4* items;
    const int32_t* n_items;
    const int32_t* pairs;
    int nprobe;
    int dbg;             // timing experiments only (VSB_TC_DBG): 1 no epilogue work, 2 no hand-off, 4 no MMA, 8 no B loads,
                         // 64 keeper consumes without scanning, 128 hand-off writes the header only,
                         // 16 epilogue = TMEM loads only, 32 epilogue = math only (no TMEM loads)
};

// smallest float strictly greater than x (x finite or +inf; +inf maps to itself)
__device__ __forceinline__ float next_up(float x) {
    if (x == 0.f) return __int_as_float(1);
    if (!(x < __int_as_float(0x7f800000))) return x;
    const int32_t i = __float_as_int(x);
    return __int_as_float(i > 0 ? i + 1 : i - 1);
}

// 16-byte shared-memory load with an explicit shared-space address (a generic-pointer LD costs a longer scoreboard wait)
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}
// volatile accesses with an explicit shared-space address: a generic `volatile` access compiles to LD/ST.E.STRONG.SYS,
// which costs hundreds of cycles
__device__ __forceinline__ uint32_t ldsv_u32(const void* p) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ float ldsv_f32(const void* p) { return __uint_as_float(ldsv_u32(p)); }
__device__ __forceinline__ void stsv_u32(void* p, uint32_t x) {
    asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(x) : "memory");
}
__device__ __forceinline__ uint2 lds64_volatile(uint32_t addr) {
    uint2 v;
    asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts64_volatile(uint32_t addr, uint2 v) {
    asm volatile("st.volatile.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// CL = CTAs per cluster (1 or 2).  CL == 2: the two CTAs of a pair work on two DIFFERENT query tiles and the SAME base
// tiles; each CTA fetches half of every base k-block (64 rows) and TMA multicasts it into both CTAs' rings, so the
// L2 -> SM traffic of the streamed operand halves (with one query tile per CTA the fp16 / 1xTF32 sweeps are bound by the
// chip-wide L2 read rate, not by the tensor pipe).  A ring stage is refilled only when BOTH CTAs' MMAs have read it:
// the `empty` barriers take two arrivals, the second one being the peer's multicast tcgen05.commit.  tmB_* are then the
// 64-row-box tensor maps.
//
// IVF = true (TF32 modes only): the same machine as the list-major fine scan of the IVF search (IVFIndex::searchBatch's
// per-query list scan, qidk_ivf/android/app/main/jni/IVFIndex.cpp:715-779, regrouped by list).  A unit is a work item
// {first pair, pairs (<= 128), first row, rows}: the A tile holds the (hi/lo split) queries of up to 128 (query, probe slot)
// pairs that probe ONE inverted list, gathered contiguously in pair order; the base tiles are that list's rows (a TMA box
// may start at any row).  Key = -2 q.x (inner product, largest first; no norm term); thresholds are shared per QUERY
// across all its lists; one sorted list per (probe slot, epilogue group, query) goes to the merge kernel, ids = row
// positions in the list-contiguous array.
template <int KTOP, int MODE, bool HAS_LB, int CL, bool IVF = false>
__global__ void __launch_bounds__(TcSmem<MODE>::THREADS, 1)
exact_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                const TcParams p) {
    constexpr bool PAIR = MODE == TC_F16 && KTOP == 32;  // the filter pass; p.n_mtiles then counts PAIRS of query tiles
    using S = TcSmem<MODE, PAIR>;
    constexpr int NSTAGE = S::NSTAGE;
    constexpr bool SPLIT3 = S::SPLIT3;
    constexpr int TC_NKB = S::NKB;
    constexpr int KB_ELEMS = MODE == TC_F16 ? 64 : 32;  // elements per 128-byte k-block
    constexpr bool SAMPLE = MODE == TC_F16 && KTOP == 1;  // TC_F16: KTOP == 1 is the sample pass, KTOP == 32 the filter pass
    static_assert(MODE != TC_F16 || KTOP == 1 || KTOP == 32, "TC_F16: sample (1) or filter (32) pass");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + S::A_BYTES;
    float* sN = (float*)(sB + S::B_BYTES);      // [TC_NACC][128] base norms of the tile in accumulator slot i
    uint2* sPub = (uint2*)((uint8_t*)sN + S::NORM_BYTES);  // [TC_EPI_GROUPS][TC_BM]
    uint8_t* sQueue = (uint8_t*)sPub + S::PUB_BYTES;              // [4][TC_EPI_GROUPS][TC_QN] entries of TC_QENTRY bytes   (SMEM_LIST)
    int* sReady = (int*)(sQueue + S::STAGE_BYTES);                // [4][TC_EPI_GROUPS][TC_QN] sequence number + 1 of the entry in the slot
    int* sTail = sReady + (S::SMEM_LIST ? 4 * TC_EPI_GROUPS * TC_QN : 0);  // [4][TC_EPI_GROUPS] entries produced (written once, at the end)
    int* sHead = sTail + 16;                                      // [4][TC_EPI_GROUPS] entries consumed so far
    uint64_t* bars = (uint64_t*)(sQueue + S::STAGE_BYTES + S::AUX_BYTES);
    uint64_t* full = bars;                    // [NSTAGE]  TMA -> MMA
    uint64_t* empty = full + NSTAGE;          // [NSTAGE]  MMA -> TMA
    uint64_t* acc_full = empty + NSTAGE;      // [TC_NACC] MMA -> epilogue
    uint64_t* acc_empty = acc_full + TC_NACC; // [TC_NACC] epilogue -> MMA, and -> producer (norm slot free)
    uint64_t* n_full = acc_empty + TC_NACC;   // [TC_NACC] norms landed
    uint64_t* a_full = n_full + TC_NACC;      // query tile landed
    uint64_t* a_empty = a_full + 1;           // query tile no longer read by the tensor core
    uint64_t* u_done = a_empty + 1;           // (SMEM_LIST) the epilogue warps have queued every candidate of the unit
    uint64_t* u_flushed = u_done + 1;         // (SMEM_LIST) the keepers have written out and reset the lists
    uint32_t* tmem_slot = (uint32_t*)(u_flushed + 1);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int cta_rank = CL == 2 ? (int)cluster_ctarank() : 0;
    const int worker = CL == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;      // a CTA, or a CTA pair
    const int n_workers = CL == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int n_mt = CL == 2 ? (p.n_mtiles + 1) >> 1 : p.n_mtiles;               // unit columns: query tiles or tile pairs

    if (threadIdx.x == 0) {
        for (int i = 0; i < NSTAGE; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], CL);
        }
        for (int i = 0; i < TC_NACC; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 4);  // the four warps of the group that consumed the tile
            mbar_init(&n_full[i], 1);
        }
        mbar_init(a_full, 1);
        mbar_init(a_empty, 1);
        mbar_init(u_done, 4 * TC_EPI_GROUPS);
        mbar_init(u_flushed, 4);
        if (S::SMEM_LIST) {
            for (int i = 0; i < 4 * TC_EPI_GROUPS * TC_QN; ++i) sReady[i] = 0;
            for (int i = 0; i < 32; ++i) sTail[i] = 0;  // tails and heads
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, TC_NACC * TC_BN);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    if (CL == 2) cluster_sync_all();  // the peer's barriers are initialised before anything of ours can signal them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // unit = (query tile [pair], base split), split-major; the odd tile out of an odd tile count is paired with a ghost
    // (all rows beyond nq: the TMA fills zeros, nothing is written)
    const int n_units = IVF ? __ldg(p.n_items) : n_mt * p.n_splits;
    static_assert(!IVF || (MODE != TC_F16 && !HAS_LB && CL == 1), "IVF units run on the TF32 path with independent CTAs");

    if (warp < 4) {
      if constexpr (S::SMEM_LIST)
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TC_REGS_CTRL_Q));
      else
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TC_REGS_CTRL));
      if (warp == 0) {
        // ===================================== TMA producer =====================================
        // the whole warp walks the loop (warp-uniform addresses stay in uniform registers); one elected lane
        // issues the asynchronous copies and barrier operations
        const bool leader = elect_one();
        {
            if (leader) {
                tma_prefetch_desc(&tmA_hi);
                tma_prefetch_desc(&tmB_hi);
                if (SPLIT3) {
                    tma_prefetch_desc(&tmA_lo);
                    tma_prefetch_desc(&tmB_lo);
                }
            }
            if constexpr (MODE == TC_F16) {
              // one ring stage = one whole base tile: k-blocks 0 and 1 (64 fp16 each) + the norm block (16 fp16, SWIZZLE_32B);
              // tmA_lo / tmB_lo are the norm-block maps
              static_assert(CL == 1, "the fp16 pass runs with independent CTAs");
              if (leader) {
                  tma_prefetch_desc(&tmA_lo);
                  tma_prefetch_desc(&tmB_lo);
              }
              int stage = 0;
              uint32_t phase = 0;
              int it = 0;
              for (int unit = worker; unit < n_units; unit += n_workers, ++it) {
                  const int m_tile = unit % n_mt;
                  const int split = unit / n_mt;
                  mbar_wait(a_empty, (uint32_t)((it & 1) ^ 1));
                  if (leader) {
                      mbar_expect_tx(a_full, (uint32_t)S::A_BYTES);
                      if constexpr (PAIR) {  // query tiles 2 m and 2 m + 1 (the latter may lie beyond nq: zero fill)
                          tma_load_2d(sA, &tmA_hi, a_full, 0, 2 * m_tile * TC_BM);
                          tma_load_2d(sA + TC_KB_BYTES, &tmA_hi, a_full, 64, 2 * m_tile * TC_BM);
                          tma_load_2d(sA + 2 * TC_KB_BYTES, &tmA_hi, a_full, 0, (2 * m_tile + 1) * TC_BM);
                          tma_load_2d(sA + 3 * TC_KB_BYTES, &tmA_hi, a_full, 64, (2 * m_tile + 1) * TC_BM);
                          tma_load_2d(sA + 4 * TC_KB_BYTES, &tmA_lo, a_full, 0, 0);
                      } else {
                          tma_load_2d(sA, &tmA_hi, a_full, 0, m_tile * TC_BM);
                          tma_load_2d(sA + TC_KB_BYTES, &tmA_hi, a_full, 64, m_tile * TC_BM);
                          tma_load_2d(sA + 2 * TC_KB_BYTES, &tmA_lo, a_full, 0, 0);  // the same 128 x 16 constants for every tile
                      }
                  }
                  const int t0 = split * p.tiles_per_split;
                  const int t1 = min(t0 + p.tiles_per_split, p.n_tiles);
                  for (int t = t0; t < t1; ++t) {
                      mbar_wait(&empty[stage], phase ^ 1);
                      if (leader) {
                          uint8_t* dst = sB + stage * S::BSTAGE;
                          if (p.dbg & 8) {
                              mbar_arrive(&full[stage]);
                          } else {
                              mbar_expect_tx(&full[stage], (uint32_t)S::BSTAGE);
                              const int row0 = (t * p.tile_stride + p.tile_off) * TC_BN;
                              tma_load_2d(dst, &tmB_hi, &full[stage], 0, row0);
                              tma_load_2d(dst + TC_KB_BYTES, &tmB_hi, &full[stage], 64, row0);
                              tma_load_2d(dst + 2 * TC_KB_BYTES, &tmB_lo, &full[stage], 0, row0);
                          }
                      }
                      if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                  }
              }
            } else {
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            int it = 0;
            for (int unit = worker; unit < n_units; unit += n_workers, ++it) {
                int a_row0, b_row0, n_t;  // first query row of the A tile, first base row, base tiles of the unit
                if constexpr (IVF) {
                    const int4 rec = __ldg(p.items + unit);
                    a_row0 = rec.x;
                    b_row0 = rec.z;
                    n_t = (rec.w + TC_BN - 1) / TC_BN;
                } else {
                    const int m_tile = (unit % n_mt) * CL + cta_rank;
                    const int split = unit / n_mt;
                    const int t0 = split * p.tiles_per_split;
                    a_row0 = m_tile * TC_BM;
                    b_row0 = t0 * TC_BN;
                    n_t = min(t0 + p.tiles_per_split, p.n_tiles) - t0;
                }
                mbar_wait(a_empty, (uint32_t)((it & 1) ^ 1));
                if (leader) {
                    mbar_expect_tx(a_full, (uint32_t)S::A_BYTES);
#pragma unroll
                    for (int kb = 0; kb < TC_NKB; ++kb) {
                        tma_load_2d(sA + kb * TC_KB_BYTES, &tmA_hi, a_full, kb * KB_ELEMS, a_row0);
                        if (SPLIT3) tma_load_2d(sA + (TC_NKB + kb) * TC_KB_BYTES, &tmA_lo, a_full, kb * 32, a_row0);
                    }
                }
                for (int ti = 0; ti < n_t; ++ti) {
                    const int t_row = b_row0 + ti * TC_BN;  // first base row of the tile
                    if constexpr (!IVF) {
                        // norms of this tile go to the slot of the accumulator the tile will use
                        mbar_wait(&acc_empty[acc], acc_phase ^ 1);
                        if (leader) {
                            mbar_expect_tx(&n_full[acc], (uint32_t)(TC_BN * 4));
                            bulk_load_1d(sN + acc * TC_BN, p.bnorm + (size_t)t_row, TC_BN * 4, &n_full[acc]);
                        }
                        if (++acc == TC_NACC) { acc = 0; acc_phase ^= 1; }
                    }
#pragma unroll
                    for (int kb = 0; kb < TC_NKB; ++kb) {
                        mbar_wait(&empty[stage], phase ^ 1);
                        if (leader) {
                            if (p.dbg & 8) {
                                mbar_arrive(&full[stage]);
                            } else {
                                mbar_expect_tx(&full[stage], (uint32_t)TC_KB_BYTES);
                                if (CL == 2)
                                    tma_load_2d_mcast(sB + stage * TC_KB_BYTES + cta_rank * (TC_KB_BYTES / 2), &tmB_hi, &full[stage],
                                                      kb * KB_ELEMS, t_row + cta_rank * (TC_BN / 2), (uint16_t)3);
                                else
                                    tma_load_2d(sB + stage * TC_KB_BYTES, &tmB_hi, &full[stage], kb * KB_ELEMS, t_row);
                            }
                        }
                        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                        if (SPLIT3) {
                            mbar_wait(&empty[stage], phase ^ 1);
                            if (leader) {
                                if (p.dbg & 8) {
                                    mbar_arrive(&full[stage]);
                                } else {
                                    mbar_expect_tx(&full[stage], (uint32_t)TC_KB_BYTES);
                                    if (CL == 2)
                                        tma_load_2d_mcast(sB + stage * TC_KB_BYTES + cta_rank * (TC_KB_BYTES / 2), &tmB_lo, &full[stage],
                                                          kb * 32, t_row + cta_rank * (TC_BN / 2), (uint16_t)3);
                                    else
                                        tma_load_2d(sB + stage * TC_KB_BYTES, &tmB_lo, &full[stage], kb * 32, t_row);
                                }
                            }
                            if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                        }
                    }
                }
            }
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer =======================================
        const bool leader = elect_one();  // the same lane issues every MMA and every commit
        {
            constexpr uint32_t idesc = umma_idesc(kIdescCF32, MODE == TC_F16 ? kIdescF16 : kIdescTF32, TC_BM, TC_BN);
            const uint32_t sA_u = smem_u32(sA);
            const uint32_t sB_u = smem_u32(sB);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            int it = 0;
            if constexpr (MODE == TC_F16) {
              // 9 instructions per tile: 2 k-blocks x 4 (K = 16 each) of  (-s_q q) . (s_b x)  and the norm block
              // (2^rho constants) . (pieces of s_b^2 ||x||^2 / 2):  acc = (s_q s_b / 2) (||x||^2 - 2 q.x)
              constexpr int NH = PAIR ? 2 : 1;  // query tiles per unit
              const uint64_t a_e = umma_desc_sw32(sA_u + NH * 2 * TC_KB_BYTES);
              for (int unit = worker; unit < n_units; unit += n_workers, ++it) {
                  const int split = unit / n_mt;
                  const int t0 = split * p.tiles_per_split;
                  const int t1 = min(t0 + p.tiles_per_split, p.n_tiles);
                  mbar_wait(a_full, (uint32_t)(it & 1));
                  tc_fence_after();
                  for (int t = t0; t < t1; ++t) {
                      mbar_wait(&full[stage], phase);
                      const uint32_t b_u = sB_u + stage * S::BSTAGE;
#pragma unroll
                      for (int h = 0; h < NH; ++h) {  // one accumulator per query tile of the unit
                          mbar_wait(&acc_empty[acc], acc_phase ^ 1);
                          tc_fence_after();
                          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * TC_BN);
                          if (!(p.dbg & 4) && leader) {
#pragma unroll
                              for (int kb = 0; kb < 2; ++kb) {
                                  const uint64_t a = umma_desc_sw128(sA_u + (2 * h + kb) * TC_KB_BYTES);
                                  const uint64_t b = umma_desc_sw128(b_u + kb * TC_KB_BYTES);
#pragma unroll
                                  for (int ks = 0; ks < 4; ++ks) tc_mma_f16(d_tmem, a + 2 * ks, b + 2 * ks, idesc, (kb | ks) ? 1u : 0u);
                              }
                              tc_mma_f16(d_tmem, a_e, umma_desc_sw32(b_u + 2 * TC_KB_BYTES), idesc, 1u);
                          }
                          if (leader) tc_commit(&acc_full[acc]);
                          if (++acc == TC_NACC) { acc = 0; acc_phase ^= 1; }
                      }
                      if (leader) tc_commit(&empty[stage]);
                      if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                  }
                  if (leader) tc_commit(a_empty);
              }
            } else
            for (int unit = worker; unit < n_units; unit += n_workers, ++it) {
                int n_t;
                if constexpr (IVF) {
                    n_t = (__ldg(p.items + unit).w + TC_BN - 1) / TC_BN;
                } else {
                    const int t0 = (unit / n_mt) * p.tiles_per_split;
                    n_t = min(t0 + p.tiles_per_split, p.n_tiles) - t0;
                }
                mbar_wait(a_full, (uint32_t)(it & 1));
                tc_fence_after();
                for (int ti = 0; ti < n_t; ++ti) {
                    mbar_wait(&acc_empty[acc], acc_phase ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * TC_BN);
#pragma unroll
                    for (int kb = 0; kb < TC_NKB; ++kb) {
                        const uint64_t a_hi = umma_desc_sw128(sA_u + kb * TC_KB_BYTES);
                        const uint64_t a_lo = umma_desc_sw128(sA_u + (TC_NKB + kb) * TC_KB_BYTES);
                        // ---- stage holding x_hi[kb]
                        mbar_wait(&full[stage], phase);
                        tc_fence_after();
                        {
                            const uint64_t b = umma_desc_sw128(sB_u + stage * TC_KB_BYTES);
                            if (SPLIT3 && !(p.dbg & 4) && leader) {
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks)    // q_lo . x_hi   (small term first)
                                    tc_mma_tf32(d_tmem, a_lo + 2 * ks, b + 2 * ks, idesc, (kb | ks) ? 1u : 0u);
                            }
                            if (!(p.dbg & 4) && leader) {
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks) {  // q_hi . x_hi   (32 bytes of K per instruction)
                                    if (MODE == TC_F16)
                                        tc_mma_f16(d_tmem, a_hi + 2 * ks, b + 2 * ks, idesc, (kb | ks) ? 1u : 0u);
                                    else
                                        tc_mma_tf32(d_tmem, a_hi + 2 * ks, b + 2 * ks, idesc, (SPLIT3 || (kb | ks)) ? 1u : 0u);
                                }
                            }
                        }
                        if (leader) { if (CL == 2) tc_commit_mcast(&empty[stage], (uint16_t)3); else tc_commit(&empty[stage]); }
                        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                        if (SPLIT3) {
                            // ---- stage holding x_lo[kb]
                            mbar_wait(&full[stage], phase);
                            tc_fence_after();
                            const uint64_t b = umma_desc_sw128(sB_u + stage * TC_KB_BYTES);
                            if (!(p.dbg & 4) && leader) {
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks)    // q_hi . x_lo
                                    tc_mma_tf32(d_tmem, a_hi + 2 * ks, b + 2 * ks, idesc, 1u);
                            }
                            if (leader) { if (CL == 2) tc_commit_mcast(&empty[stage], (uint16_t)3); else tc_commit(&empty[stage]); }
                            if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                        }
                    }
                    if (leader) tc_commit(&acc_full[acc]);
                    if (++acc == TC_NACC) { acc = 0; acc_phase ^= 1; }
                }
                if (leader) tc_commit(a_empty);
            }
        }
      }
    } else {
      if constexpr (S::SMEM_LIST) {
       if (warp >= 4 + 4 * TC_EPI_GROUPS) {
        // ===================================== candidate writers (TC_F16 filter pass) =============
        // Keeper warp `quad` drains the queue of TMEM lane quadrant `quad`.  An entry = the 32 keys of one 32-column chunk
        // of one query row whose minimum lies below the query's threshold.  Lane i of a batch takes entry i: it scans the
        // 32 keys (lane-rotated order, conflict-free banks), reserves room in the query's candidate array with ONE
        // atomicAdd and appends the (key, row id) pairs that pass.  No lists, no ordering: the merge kernel selects.
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TC_REGS_KEEP_Q));
        if constexpr (!SAMPLE) {
        const int quad = warp & 3;
        const int qbatch = p.qbatch > 0 ? p.qbatch : TC_QBATCH;
        // the quadrant's three queues (one per epilogue group); a batch takes entries from all of them
        int next[TC_EPI_GROUPS], final_tail[TC_EPI_GROUPS];
#pragma unroll
        for (int g = 0; g < TC_EPI_GROUPS; ++g) {
            next[g] = 0;         // entries consumed so far
            final_tail[g] = -1;  // entries produced in total, known once every epilogue warp has finished its last unit
        }
        bool done = false;
        while (true) {
            int n[TC_EPI_GROUPS], total = 0;
            bool half_full = false;
#pragma unroll
            for (int g = 0; g < TC_EPI_GROUPS; ++g) {
                const int my = next[g] + lane;
                const int* ready = sReady + (quad * TC_EPI_GROUPS + g) * TC_QN;
                const unsigned rm = __ballot_sync(0xffffffffu, lane < TC_QN && (int)ldsv_u32(ready + (my & (TC_QN - 1))) == my + 1);
                n[g] = rm == 0xffffffffu ? 32 : __ffs(~rm) - 1;  // consecutive ready entries
                total += n[g];
                half_full |= n[g] >= TC_QN / 2;
            }
            if (total == 0 || (total < qbatch && !done && !half_full)) {
                // nothing, or not yet a worthwhile batch (a batch has a mostly fixed cost) and no queue is filling up
                if (!done) {
                    if (mbar_try_wait(u_done, 0u)) {
                        done = true;
#pragma unroll
                        for (int g = 0; g < TC_EPI_GROUPS; ++g) final_tail[g] = (int)ldsv_u32(sTail + quad * TC_EPI_GROUPS + g);
                    } else {
                        __nanosleep(total == 0 ? 64 : 100);
                    }
                    continue;
                }
                if (total == 0) {
                    bool all = true;
#pragma unroll
                    for (int g = 0; g < TC_EPI_GROUPS; ++g) all &= next[g] == final_tail[g];
                    if (all) break;
                    continue;
                }
            }
            asm volatile("fence.acq_rel.cta;" ::: "memory");
            // lanes [0, t0) take queue 0, [t0, t0 + t1) queue 1, ...
            int take[TC_EPI_GROUPS], left = 32, first = 0, my_g = -1, my_idx = 0;
#pragma unroll
            for (int g = 0; g < TC_EPI_GROUPS; ++g) {
                take[g] = min(n[g], left);
                left -= take[g];
                if (my_g < 0 && lane < first + take[g]) {
                    my_g = g;
                    my_idx = next[g] + lane - first;
                }
                first += take[g];
            }
            const bool act = my_g >= 0;
            const uint32_t e = smem_u32(sQueue) + (uint32_t)(((quad * TC_EPI_GROUPS + (act ? my_g : 0)) * TC_QN + (my_idx & (TC_QN - 1))) * TC_QENTRY);
            int qg = 0, col0 = 0, pad_;
            float thr_e = -__int_as_float(0x7f800000);
            if (act) asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(qg), "=r"(col0), "=f"(thr_e), "=r"(pad_) : "r"(e + 128u) : "memory");
            unsigned qm = 0;
            if (!(p.dbg & 64)) {  // (timing experiment 64: consume without looking)
#pragma unroll
            for (int j0 = 0; j0 < 32; j0 += 8) {  // eight loads in flight, then their tests
                float v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = lds_f32(e + (uint32_t)((j0 + u + lane) & 31) * 4u);
#pragma unroll
                for (int u = 0; u < 8; ++u) qm |= (v[u] < thr_e) ? (1u << ((j0 + u + lane) & 31)) : 0u;
            }
            }
            int pos = 0;
            if (qm != 0) pos = atomicAdd(p.cand_cnt + qg, __popc(qm));
            uint2* dst = p.cand + (size_t)qg * p.cand_cap;
            while (qm != 0) {
                const int j = __ffs(qm) - 1;
                qm &= qm - 1;
                const float x = lds_f32(e + (uint32_t)j * 4u);
                if (pos < p.cand_cap) __stcg(dst + pos, make_uint2(__float_as_uint(x), (uint32_t)(col0 + j)));
                ++pos;
            }
            __syncwarp();
#pragma unroll
            for (int g = 0; g < TC_EPI_GROUPS; ++g) {
                next[g] += take[g];
                if (lane == 0 && take[g] > 0) stsv_u32(sHead + quad * TC_EPI_GROUPS + g, (uint32_t)next[g]);  // the slots may be reused
            }
            if (p.stats && lane == 0) {
                atomicAdd(p.stats + 4, (unsigned long long)(32 - left));
                atomicAdd(p.stats + 5, 1ull);
            }
        }
        }
       } else {
        // ===================================== epilogue (TC_F16) ==================================
        // Filter pass: per 32-column chunk the keys (the accumulator IS the key: norm and dot product were combined by the
        // tensor core), a min tree and ONE test of the row minimum against the query's threshold thr[q], which is fixed for
        // the whole launch — no lists, no bounds to refresh.  A chunk that passes is handed to the quadrant's keeper as a whole
        // (eight 16-byte stores by all passing lanes together).
        // Sample pass (KTOP == 1): no threshold yet; every thread keeps TC_SAMPLE_SUB running minima over the tiles it sees
        // (tile sequence number mod TC_SAMPLE_SUB) and writes them out per unit: the group minima from which
        // tc_select_thr_kernel derives thr[q].
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(TC_REGS_EPI_Q));
        const int quad = warp & 3;
        const int grp = (warp - 4) >> 2;
        const int row = quad * 32 + lane;
        const float INF = __int_as_float(0x7f800000);
        constexpr int CH = TC_BN / 32;
        // this warp's own queue: slots are handed out from a register counter, the consumer's head is re-read only when the
        // queue looks full
        const uint32_t queue_u = smem_u32(sQueue + (quad * TC_EPI_GROUPS + grp) * TC_QN * TC_QENTRY);
        const uint32_t ready_u = smem_u32(sReady + (quad * TC_EPI_GROUPS + grp) * TC_QN);
        const int* my_head = sHead + quad * TC_EPI_GROUPS + grp;
        int qtail = 0, head_seen = 0;
        int tcount = 0;
        for (int unit = worker; unit < n_units; unit += n_workers) {
            constexpr int NH = PAIR ? 2 : 1;  // query tiles per unit: every base tile yields NH accumulator tiles
            const int m_tile = (unit % n_mt) * NH;
            const int split = unit / n_mt;
            const int t0 = split * p.tiles_per_split;
            const int t1 = min(t0 + p.tiles_per_split, p.n_tiles);
            int qh[NH];
            bool validh[NH], liveh[NH];
            float thrh[NH];
#pragma unroll
            for (int h = 0; h < NH; ++h) {
                qh[h] = (m_tile + h) * TC_BM + row;
                validh[h] = qh[h] < p.nq;
                liveh[h] = (m_tile + h) * TC_BM + quad * 32 < p.nq;
                thrh[h] = -INF;  // rows beyond the last query never pass
                if (!SAMPLE && validh[h]) thrh[h] = __ldg(p.thr + qh[h]);
            }
            float gmin[TC_SAMPLE_SUB];
#pragma unroll
            for (int u = 0; u < TC_SAMPLE_SUB; ++u) gmin[u] = INF;
            const int nv = (t1 - t0) * NH;  // accumulator tiles of the unit, in MMA order: base tile v / NH, query tile v % NH
            int first = (grp - tcount % TC_EPI_GROUPS + TC_EPI_GROUPS) % TC_EPI_GROUPS;
            int j = 0;
            for (int i = first; i < nv; i += TC_EPI_GROUPS, ++j) {
                const int hsel = PAIR ? (i & 1) : 0;
                const int q = PAIR ? (hsel ? qh[NH - 1] : qh[0]) : qh[0];
                const bool valid = PAIR ? (hsel ? validh[NH - 1] : validh[0]) : validh[0];
                const bool quad_live = PAIR ? (hsel ? liveh[NH - 1] : liveh[0]) : liveh[0];
                const float thr = PAIR ? (hsel ? thrh[NH - 1] : thrh[0]) : thrh[0];
                const int t = (t0 + i / NH) * p.tile_stride + p.tile_off;  // the base tile
                const int tc = tcount + i;
                const int acc = tc & (TC_NACC - 1);
                const uint32_t acc_phase = (uint32_t)(tc / TC_NACC) & 1u;
                mbar_wait(&acc_full[acc], acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * TC_BN);
                uint32_t r[2][32];
                const bool skip = (p.dbg & 1) || !quad_live;
                // the last base tile may be ragged: its missing rows arrive as zeros (TMA fill) and would rank as key 0
                const int live_cols = (t == p.n_tiles_real - 1 && p.n_rem) ? p.n_rem : TC_BN;
                float tmin = INF;
                if (!skip) tmem_ld32(taddr, r[0]);
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    if (skip) break;
                    tc_wait_ld();
                    if (c + 1 < CH) tmem_ld32(taddr + (c + 1) * 32, r[(c + 1) & 1]);
                    if (p.dbg & 16) continue;                // timing experiment: TMEM loads without the arithmetic
                    float d[32];
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj) d[jj] = __uint_as_float(r[c & 1][jj]);
                    if (live_cols < TC_BN) {
#pragma unroll
                        for (int jj = 0; jj < 32; ++jj) d[jj] = (c * 32 + jj < live_cols) ? d[jj] : INF;
                    }
                    float m[16];  // min tree (the compiler folds it into 3-input FMNMX3)
#pragma unroll
                    for (int jj = 0; jj < 16; ++jj) m[jj] = fminf(d[jj], d[jj + 16]);
#pragma unroll
                    for (int w = 8; w >= 1; w >>= 1)
#pragma unroll
                        for (int jj = 0; jj < w; ++jj) m[jj] = fminf(m[jj], m[jj + w]);
                    if constexpr (SAMPLE) {
                        tmin = fminf(tmin, m[0]);
                        continue;
                    }
                    const bool hit = m[0] < thr;
                    const unsigned todo = __ballot_sync(0xffffffffu, hit);
                    if (todo != 0 && !(p.dbg & 2)) {
                        if (p.stats && lane == 0) {
                            atomicAdd(p.stats + 0, 1ull);
                            atomicAdd(p.stats + 1, (unsigned long long)__popc(todo));
                        }
                        if (hit) {
                            const int idx = qtail + __popc(todo & ((1u << lane) - 1u));
                            while (idx - head_seen >= TC_QN) {  // looks full: refresh the head; really full: leave the issue slots to the keeper
                                head_seen = (int)ldsv_u32(my_head);
                                if (idx - head_seen >= TC_QN) __nanosleep(100);
                            }
                            const uint32_t e = queue_u + (uint32_t)(idx & (TC_QN - 1)) * TC_QENTRY;
                            if (!(p.dbg & 128)) {  // (timing experiment 128: header only)
#pragma unroll
                            for (int j4 = 0; j4 < 8; ++j4)
                                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(e + (uint32_t)(j4 << 4)),
                                             "f"(d[4 * j4 + 0]), "f"(d[4 * j4 + 1]), "f"(d[4 * j4 + 2]), "f"(d[4 * j4 + 3]) : "memory");
                            }
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(e + 128u), "r"(q), "r"(t * TC_BN + c * 32),
                                         "f"(thr), "r"(0) : "memory");
                            // release: the entry is complete before its sequence number shows
                            asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(ready_u + (uint32_t)(idx & (TC_QN - 1)) * 4u), "r"(idx + 1) : "memory");
                        }
                        qtail += __popc(todo);
                        __syncwarp();
                    }
                }
                if constexpr (SAMPLE) {
#pragma unroll
                    for (int u = 0; u < TC_SAMPLE_SUB; ++u) gmin[u] = (j & (TC_SAMPLE_SUB - 1)) == u ? fminf(gmin[u], tmin) : gmin[u];
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[acc]);
            }
            tcount += nv;
            if constexpr (SAMPLE) {
                if (validh[0]) {
#pragma unroll
                    for (int u = 0; u < TC_SAMPLE_SUB; ++u)
                        p.smin[((size_t)(split * TC_EPI_GROUPS + grp) * TC_SAMPLE_SUB + u) * p.nq + qh[0]] = gmin[u];
                }
            }
        }
        __syncwarp();
        if (!SAMPLE && lane == 0) {
            stsv_u32(sTail + quad * TC_EPI_GROUPS + grp, (uint32_t)qtail);
            mbar_arrive(u_done);  // every candidate of this warp is in its queue, the final count is published (release)
        }
       }
      } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(TC_REGS_EPI));
        // ===================================== epilogue ==========================================
        // Tiles rotate over the TC_EPI_GROUPS warpgroups (the CTA's running tile count modulo the group count), so
        // several tiles are in flight in the epilogue and each warp's TMEM-load / barrier latencies are covered by
        // the other groups' arithmetic.  A thread owns one query row of its group's tiles (all 128 columns).
        const int quad = warp & 3;
        const int grp = (warp - 4) >> 2;
        const int row = quad * 32 + lane;
        const float INF = __int_as_float(0x7f800000);
        const float key_scale = MODE == TC_F16 ? __ldg(p.key_scale_ptr) : -2.0f;
        constexpr int CH = TC_BN / 32;        // 32-column chunks per thread per tile
        constexpr bool DB = KTOP <= 16;       // double-buffered TMEM loads while the register budget allows
        int tcount = 0;                       // tiles this CTA has gone through before the current unit
        for (int unit = worker; unit < n_units; unit += n_workers) {
            int q, n_t, b_row0, b_rows = 0x7fffffff, split = 0, slot = 0;
            bool valid, quad_live;
            if constexpr (IVF) {
                const int4 rec = __ldg(p.items + unit);  // {first pair, pairs, first row, rows}
                valid = row < rec.y;
                quad_live = quad * 32 < rec.y;
                const int pair = valid ? __ldg(p.pairs + rec.x + row) : 0;
                q = pair / p.nprobe;
                slot = pair - q * p.nprobe;
                b_row0 = rec.z;
                b_rows = rec.w;
                n_t = (rec.w + TC_BN - 1) / TC_BN;
            } else {
                const int m_tile = (unit % n_mt) * CL + cta_rank;
                split = unit / n_mt;
                const int t0 = split * p.tiles_per_split;
                q = m_tile * TC_BM + row;
                valid = q < p.nq;
                // whole 32-lane quadrant beyond the last query (small batches): barriers only, no TMEM traffic
                quad_live = m_tile * TC_BM + quad * 32 < p.nq;
                b_row0 = t0 * TC_BN;
                n_t = min(t0 + p.tiles_per_split, p.n_tiles) - t0;
            }
            float lbk = -INF;
            int32_t lbi = -1;
            if (HAS_LB && valid) {
                lbk = __ldg(p.lb_key + q);
                lbi = __ldg(p.lb_id + q);
            }
            RegTopK<KTOP> top;
            top.init();
            // Threshold sharing.  The groups of a CTA hold disjoint lists of the same query: once each of them keeps
            // at least SUB = ceil(KTOP / groups) entries, the largest of their SUB-th best keys bounds the query's
            // KTOP-th best key (groups x SUB >= KTOP rows lie at or below it).  Each thread posts its SUB-th best key,
            // tagged with the unit it belongs to (groups run a few tiles apart and may be in different units), in
            // shared memory; the combined bound also goes to the global per-query array read by the other CTAs.
            // cap = best such bound seen so far.  Keys equal to it may still belong to the canonical answer (smaller
            // id), hence the strict test against next_up(cap).
            constexpr int SUB = (KTOP + TC_EPI_GROUPS - 1) / TC_EPI_GROUPS;
            const uint32_t my_pub = smem_u32(sPub + grp * TC_BM + row);
            sts64_volatile(my_pub, make_uint2(__float_as_uint(INF), (uint32_t)unit));
            float cap = INF;
            float thr = INF;  // invariant: thr == min(top.threshold(), next_up(cap))
            // shared threshold: the value consumed at a refresh point was requested one refresh earlier
            int32_t pending = valid ? __ldcg(p.gthr + q) : 0x7f7f7f7f;
            int first = (grp - tcount % TC_EPI_GROUPS + TC_EPI_GROUPS) % TC_EPI_GROUPS;
            int j = 0;  // tiles of this unit seen by this group
            for (int i = first; i < n_t; i += TC_EPI_GROUPS, ++j) {
                const int tc = tcount + i;
                const int acc = tc & (TC_NACC - 1);
                const uint32_t acc_phase = (uint32_t)(tc / TC_NACC) & 1u;
                const int rel = j & (TC_THR_REFRESH - 1);
                if (valid) {
                    // bound from the sibling groups' posts (every tile: three 8-byte shared loads)
                    float comb = top.key[SUB - 1];
#pragma unroll
                    for (int g = 1; g < TC_EPI_GROUPS; ++g) {
                        const uint2 o = lds64_volatile(smem_u32(sPub + ((grp + g) % TC_EPI_GROUPS) * TC_BM + row));
                        comb = fmaxf(comb, o.y == (uint32_t)unit ? __uint_as_float(o.x) : INF);
                    }
                    if (comb < cap) {
                        cap = comb;
                        if (rel == TC_THR_REFRESH - 1) atomicMin(p.gthr + q, float_to_ordered(comb));
                    }
                    if (rel == 0) {
                        cap = fminf(cap, ordered_to_float(pending));
                        pending = __ldcg(p.gthr + q);
                    }
                    thr = fminf(top.threshold(), next_up(cap));
                }
                if constexpr (!IVF) mbar_wait(&n_full[acc], acc_phase);
                mbar_wait(&acc_full[acc], acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * TC_BN);
                const uint32_t bn_s = smem_u32(sN + acc * TC_BN);
                const int live_cols = b_rows - i * TC_BN;  // IVF: rows of the list left in this tile (the next list follows)
                uint32_t r[DB ? 2 : 1][32];
                const bool skip = (p.dbg & 1) || !quad_live;
                if (!skip) tmem_ld32(taddr, r[0]);
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    if (skip) break;
                    tc_wait_ld();
                    if (DB && c + 1 < CH) tmem_ld32(taddr + (c + 1) * 32, r[(c + 1) & 1]);
                    const int col0 = b_row0 + i * TC_BN + c * 32;
                    float d[32];
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        const uint32_t* rr = r[DB ? (c & 1) : 0] + 4 * j4;
                        if constexpr (IVF) {  // inner product, largest first: key = -2 q.x; columns past the list's end never rank
                            d[4 * j4 + 0] = c * 32 + 4 * j4 + 0 < live_cols ? -2.0f * __uint_as_float(rr[0]) : INF;
                            d[4 * j4 + 1] = c * 32 + 4 * j4 + 1 < live_cols ? -2.0f * __uint_as_float(rr[1]) : INF;
                            d[4 * j4 + 2] = c * 32 + 4 * j4 + 2 < live_cols ? -2.0f * __uint_as_float(rr[2]) : INF;
                            d[4 * j4 + 3] = c * 32 + 4 * j4 + 3 < live_cols ? -2.0f * __uint_as_float(rr[3]) : INF;
                            continue;
                        }
                        const float4 bn = lds128(bn_s + (uint32_t)(c * 32 + 4 * j4) * 4u);  // smem broadcast
                        if (MODE == TC_F16) {
                            d[4 * j4 + 0] = fmaf(key_scale, __uint_as_float(rr[0]), bn.x);
                            d[4 * j4 + 1] = fmaf(key_scale, __uint_as_float(rr[1]), bn.y);
                            d[4 * j4 + 2] = fmaf(key_scale, __uint_as_float(rr[2]), bn.z);
                            d[4 * j4 + 3] = fmaf(key_scale, __uint_as_float(rr[3]), bn.w);
                        } else {
                            d[4 * j4 + 0] = fmaf(-2.0f, __uint_as_float(rr[0]), bn.x);
                            d[4 * j4 + 1] = fmaf(-2.0f, __uint_as_float(rr[1]), bn.y);
                            d[4 * j4 + 2] = fmaf(-2.0f, __uint_as_float(rr[2]), bn.z);
                            d[4 * j4 + 3] = fmaf(-2.0f, __uint_as_float(rr[3]), bn.w);
                        }
                    }
                    if (!DB && c + 1 < CH) tmem_ld32(taddr + (c + 1) * 32, r[0]);  // r[0] is dead from here on
                    if (p.dbg & 16) continue;
                    if (HAS_LB) {
#pragma unroll
                        for (int jj = 0; jj < 32; ++jj) {
                            const bool after = d[jj] > lbk || (d[jj] == lbk && (col0 + jj) > lbi);
                            d[jj] = after ? d[jj] : INF;
                        }
                    }
                    float m[16];  // pairwise min tree (depth 5)
#pragma unroll
                    for (int jj = 0; jj < 16; ++jj) m[jj] = fminf(d[jj], d[jj + 16]);
#pragma unroll
                    for (int w = 8; w >= 1; w >>= 1)
#pragma unroll
                        for (int jj = 0; jj < w; ++jj) m[jj] = fminf(m[jj], m[jj + w]);
                    if (m[0] < thr && !(p.dbg & 2)) {
                        // rare path: bit mask of the qualifying columns, then one insertion per set bit
                        uint32_t mask = 0;
#pragma unroll
                        for (int jj = 0; jj < 32; ++jj) mask |= (d[jj] < thr) ? (1u << jj) : 0u;
                        if (p.stats) {
                            if ((__activemask() & ((1u << lane) - 1)) == 0) atomicAdd(p.stats + 0, 1ull);
                            atomicAdd(p.stats + 1, 1ull);
                            atomicAdd(p.stats + 2, (unsigned long long)__popc(mask));
                        }
                        while (mask) {
                            const int jj = __ffs(mask) - 1;
                            mask &= mask - 1;
                            const float v = select32(d, jj);
                            if (v < thr) {
                                if (p.stats) atomicAdd(p.stats + 3, 1ull);
                                top.insert(v, col0 + jj);
                                thr = fminf(thr, top.threshold());
                            }
                        }
                        // +inf until SUB entries are kept
                        sts64_volatile(my_pub, make_uint2(__float_as_uint(top.key[SUB - 1]), (uint32_t)unit));
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[acc]);
            }
            tcount += n_t;
            if (IVF && valid) {
                // what this (query, list, group) kept goes to the query's candidate array (at most KTOP per thread, so
                // nprobe * TC_EPI_GROUPS * KTOP slots per query can never overflow)
                if (cap < INF) atomicMin(p.gthr + q, float_to_ordered(cap));
                int n_kept = 0;
#pragma unroll
                for (int i = 0; i < KTOP; ++i) n_kept += top.id[i] >= 0 ? 1 : 0;
                if (n_kept > 0) {
                    const int at = atomicAdd(p.cand_cnt + q, n_kept);
                    uint2* dst = p.cand + (size_t)q * p.cand_cap + at;
#pragma unroll
                    for (int i = 0; i < KTOP; ++i)
                        if (i < n_kept) __stcg(dst + i, make_uint2(__float_as_uint(top.key[i]), (uint32_t)top.id[i]));
                }
            } else if (valid) {
                if (cap < INF) atomicMin(p.gthr + q, float_to_ordered(cap));
                const size_t list = (size_t)split * TC_EPI_GROUPS + grp;
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
    }
    tc_fence_before();
    __syncthreads();
    if (CL == 2) cluster_sync_all();  // no CTA leaves while its peer's multicasts / arrivals may still target it
    if (warp == 1) tmem_dealloc(tmem_base, TC_NACC * TC_BN);
}

}  // namespace vsb
