// Shared device/host helpers for libvsb200: error plumbing, sm_100a PTX wrappers (mbarrier, TMA, tcgen05),
// register-resident top-k lists and a warp-level k-selection over candidate arrays.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/vsb200.h"

namespace vsb {

// ------------------------------------------------------------------------------------------------
// host-side error plumbing
// ------------------------------------------------------------------------------------------------
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);

#define VSB_CUDA(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return ::vsb::fail(VS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));        \
    } while (0)

#define VSB_TRY(expr)                  \
    do {                               \
        int _rc = (expr);              \
        if (_rc != VS_OK) return _rc;  \
    } while (0)

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

constexpr int kMaxRegK = 32;  // largest k kept in a per-thread register list

// restores the caller's current CUDA device on scope exit (entry points that touch several devices)
struct DeviceGuard {
    int prev = -1;
    DeviceGuard() {
        if (cudaGetDevice(&prev) != cudaSuccess) {
            prev = -1;
            cudaGetLastError();
        }
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// grow-only device buffer
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return VS_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        const size_t want = bytes + bytes / 4;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) return fail(VS_ERR_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
        cap = want;
        return VS_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T* as() const { return (T*)p; }
};


#if defined(__CUDACC__)
// ------------------------------------------------------------------------------------------------
// PTX wrappers (sm_100a)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// one lane of the (fully active) warp
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// 2D TMA tile load global -> shared, completion on an mbarrier (transaction bytes)
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar, int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// the same tile delivered to the same shared-memory offset of every CTA in `cta_mask` of the cluster; each destination
// CTA's mbarrier (same offset) receives the complete_tx for the bytes it got
__device__ __forceinline__ void tma_load_2d_mcast(void* smem_dst, const void* tmap, uint64_t* bar, int32_t c0, int32_t c1,
                                                  uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// all threads of all CTAs of the cluster (release / acquire)
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
// 1D bulk copy global -> shared (contiguous bytes, 16-B aligned and a multiple of 16)
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- tcgen05 / TMEM ----
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {  // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // same warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// all previously issued tcgen05.mma of this thread arrive on `bar` when complete (implies fence::before_thread_sync)
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// the same, arriving on the barrier at this offset in every CTA of `cta_mask` (operand stages shared by a CTA pair)
__device__ __forceinline__ void tc_commit_mcast(uint64_t* bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, kind::tf32 (K = 8 per instruction), issued by ONE thread
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// u8 x u8 -> s32 (K = 32 per instruction)
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns: thread t of the warp gets row (lane base + t), columns c..c+31
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}

// Shared-memory matrix descriptor, K-major operand, 128-byte swizzle: rows of 128 B (32 fp32 / 128 u8 along K),
// 8-row groups 1024 B apart (SBO), version 1 (sm_100), layout type 2 (SWIZZLE_128B).  The tile base must be
// 1024-B aligned; stepping along K inside the 128-B row adds the byte offset to the start address.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);   // start address        bits [0,14)
    d |= (uint64_t)0 << 16;                       // leading byte offset  bits [16,30) (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset   bits [32,46)
    d |= (uint64_t)1 << 46;                       // descriptor version   bits [46,48)
    d |= (uint64_t)2 << 61;                       // SWIZZLE_128B         bits [61,64)
    return d;
}
// The same for 32-byte swizzle: rows of 32 B (16 fp16 along K = one K = 16 instruction), 8-row groups 256 B apart.
__device__ __forceinline__ uint64_t umma_desc_sw32(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)(256 >> 4) << 32;              // stride byte offset
    d |= (uint64_t)1 << 46;                       // descriptor version
    d |= (uint64_t)6 << 61;                       // SWIZZLE_32B
    return d;
}
// Instruction descriptor: D fp32 (or s32), A/B format, both K-major, M x N tile.
__host__ __device__ constexpr uint32_t umma_idesc(uint32_t c_fmt, uint32_t ab_fmt, uint32_t M, uint32_t N) {
    return (c_fmt << 4) | (ab_fmt << 7) | (ab_fmt << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
constexpr uint32_t kIdescCF32 = 1, kIdescCS32 = 2, kIdescTF32 = 2, kIdescF16 = 0, kIdescU8 = 0;

// ------------------------------------------------------------------------------------------------
// Per-thread sorted top-k list in registers: smallest keys first, ties keep the earlier insertion first
// (callers feed ids in ascending order, which yields the canonical (key asc, id asc) order).
// ------------------------------------------------------------------------------------------------
template <int K>
struct RegTopK {
    float key[K];
    int32_t id[K];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int i = 0; i < K; ++i) {
            key[i] = __int_as_float(0x7f800000);
            id[i] = -1;
        }
    }
    __device__ __forceinline__ float threshold() const { return key[K - 1]; }
    // precondition: v < threshold().  Every slot is computed independently (no serial bubble chain):
    // entries with key <= v stay, the first slot whose key > v receives v, later slots shift down by one.
    __device__ __forceinline__ void insert(float v, int32_t i) {
        bool keep[K];
#pragma unroll
        for (int s = 0; s < K; ++s) keep[s] = key[s] <= v;
#pragma unroll
        for (int s = K - 1; s > 0; --s) {
            const float nk = keep[s - 1] ? v : key[s - 1];
            const int32_t ni = keep[s - 1] ? i : id[s - 1];
            key[s] = keep[s] ? key[s] : nk;
            id[s] = keep[s] ? id[s] : ni;
        }
        if (!keep[0]) {
            key[0] = v;
            id[0] = i;
        }
    }
    // general insert honouring (key, id) order for arbitrary arrival order
    __device__ __forceinline__ void insert_any(float v, int32_t i) {
        if (!(v < key[K - 1] || (v == key[K - 1] && (uint32_t)i < (uint32_t)id[K - 1]))) return;
        key[K - 1] = v;
        id[K - 1] = i;
#pragma unroll
        for (int s = K - 1; s > 0; --s) {
            const bool sw = key[s] < key[s - 1] || (key[s] == key[s - 1] && (uint32_t)id[s] < (uint32_t)id[s - 1]);
            const float a = key[s], b = key[s - 1];
            const int32_t ia = id[s], ib = id[s - 1];
            key[s] = sw ? b : a;
            key[s - 1] = sw ? a : b;
            id[s] = sw ? ib : ia;
            id[s - 1] = sw ? ia : ib;
        }
    }
};

// value d[j] for a run-time j without dynamic register indexing: 5-level multiplexer (31 selects)
template <class T>
__device__ __forceinline__ T select32(const T (&d)[32], int j) {
    T a[16], b[8], c[4];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = (j & 16) ? d[i + 16] : d[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) b[i] = (j & 8) ? a[i + 8] : a[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i] = (j & 4) ? b[i + 4] : b[i];
    const T e0 = (j & 2) ? c[2] : c[0];
    const T e1 = (j & 2) ? c[3] : c[1];
    return (j & 1) ? e1 : e0;
}

// order-preserving float <-> int mapping (so that atomicMin on int orders floats, negatives included)
__device__ __forceinline__ int32_t float_to_ordered(float f) {
    const int32_t i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int32_t i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// (key, id) lexicographic "a before b"; id compared unsigned so that the -1 padding sorts last
__device__ __forceinline__ bool pair_less(float ka, int32_t ia, float kb, int32_t ib) {
    return ka < kb || (ka == kb && (uint32_t)ia < (uint32_t)ib);
}

// pop-min merge of the lists held by the lanes of a warp; `mine` = this lane takes part. Lane 0 writes k entries.
template <int KTOP>
__device__ __forceinline__ void warp_merge_lists(RegTopK<KTOP>& L, int k, float* out_key, int32_t* out_id) {
    const int lane = threadIdx.x & 31;
    const float INF = __int_as_float(0x7f800000);
    for (int r = 0; r < k; ++r) {
        float hk = L.key[0];
        int32_t hid = L.id[0];
        int src = lane;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ok = __shfl_xor_sync(0xffffffffu, hk, o);
            const int32_t oi = __shfl_xor_sync(0xffffffffu, hid, o);
            const int os = __shfl_xor_sync(0xffffffffu, src, o);
            if (pair_less(ok, oi, hk, hid)) {
                hk = ok;
                hid = oi;
                src = os;
            }
        }
        if (lane == 0) {
            out_key[r] = hid >= 0 ? hk : INF;
            out_id[r] = hid;
        }
        if (src == lane && hid >= 0) {
#pragma unroll
            for (int i = 0; i + 1 < KTOP; ++i) {
                L.key[i] = L.key[i + 1];
                L.id[i] = L.id[i + 1];
            }
            L.key[KTOP - 1] = INF;
            L.id[KTOP - 1] = -1;
        }
    }
}

// One warp selects the k smallest (key, id) pairs out of C candidates living in shared or global memory
// (keys are overwritten with +inf as they are consumed).  Results written by lane 0 in canonical order.
__device__ __forceinline__ void warp_select_k(float* ck, int32_t* cid, int C, int k, float* out_key, int32_t* out_id,
                                              int out_stride) {
    const int lane = threadIdx.x & 31;
    const float INF = __int_as_float(0x7f800000);
    for (int r = 0; r < k; ++r) {
        float bk = INF;
        int32_t bi = -1;
        int bpos = -1;
        for (int c = lane; c < C; c += 32) {
            const float kk = ck[c];
            const int32_t ii = cid[c];
            if (ii >= 0 && (bpos < 0 || pair_less(kk, ii, bk, bi))) {
                bk = kk;
                bi = ii;
                bpos = c;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ok = __shfl_xor_sync(0xffffffffu, bk, o);
            const int32_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
            const int op = __shfl_xor_sync(0xffffffffu, bpos, o);
            if (op >= 0 && (bpos < 0 || pair_less(ok, oi, bk, bi))) {
                bk = ok;
                bi = oi;
                bpos = op;
            }
        }
        if (lane == 0) {
            out_key[r * out_stride] = bpos >= 0 ? bk : INF;
            out_id[r * out_stride] = bpos >= 0 ? bi : -1;
            if (bpos >= 0) cid[bpos] = -1;  // consumed
        }
        __syncwarp();
    }
}
#endif  // __CUDACC__

}  // namespace vsb
