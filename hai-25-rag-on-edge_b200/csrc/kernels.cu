// Index-build / query-prep kernels, list merge (K3) and the device-side synthetic data generator.
#include <cuda_fp16.h>

#include "kernels.cuh"

#include <type_traits>
#include "vsb_common.cuh"

namespace vsb {

int round_up_ktop(int k) {
    if (k <= 1) return 1;
    if (k <= 5) return 5;
    if (k <= 10) return 10;
    if (k <= 16) return 16;
    if (k <= 32) return 32;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// prep: norms (reference order) + TF32 hi/lo split
// ------------------------------------------------------------------------------------------------
// 8 threads per row; thread l accumulates v[l], v[l+8], ... with FMA exactly like lane l of the AVX2 register in
// compute_norm_avx2 (cpu_baseline.cpp:99-102); lanes are then added in order 0..7 (:106-107), the scalar tail
// (:109-111, contracted to FMA by g++ -O3 -mfma) follows.  Each step the 8 threads read 32 contiguous bytes.
__device__ __forceinline__ float round_tf32(float a) {  // round to nearest (ties away), low 13 mantissa bits zero
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(a));
    return __uint_as_float(r);
}

__global__ void __launch_bounds__(256) prep_rows_kernel(const float* __restrict__ x, int64_t rows, int dim,
                                                        float* __restrict__ norms, float* __restrict__ hi,
                                                        float* __restrict__ lo, int* __restrict__ not_exact) {
    const int l = threadIdx.x & 7;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const bool active = row < rows;
    const float* v = x + (active ? row : 0) * (int64_t)dim;
    float acc = 0.f;
    int inexact = 0;
    const int dim8 = dim & ~7;
    if (active) {
        for (int i = l; i < dim8; i += 8) {
            const float a = __ldg(v + i);
            acc = fmaf(a, a, acc);
            if (hi) {
                const float h = round_tf32(a);
                const float r = round_tf32(a - h);  // a - h is exact
                hi[row * (int64_t)dim + i] = h;
                lo[row * (int64_t)dim + i] = r;
                inexact |= (r != 0.f);
            } else if (not_exact) {
                inexact |= (round_tf32(a) != a);
            }
        }
    }
    // lanes summed 0..7 in order by the first thread of the group
    float s = acc;
#pragma unroll
    for (int j = 1; j < 8; ++j) {
        const float o = __shfl_sync(0xffffffffu, acc, (threadIdx.x & 31 & ~7) + j);
        if (l == 0) s = s + o;
    }
    if (active && l == 0) {
        for (int i = dim8; i < dim; ++i) {
            const float a = __ldg(v + i);
            s = fmaf(a, a, s);
            if (hi) {
                const float h = round_tf32(a);
                const float r = round_tf32(a - h);
                hi[row * (int64_t)dim + i] = h;
                lo[row * (int64_t)dim + i] = r;
                inexact |= (r != 0.f);
            } else if (not_exact) {
                inexact |= (round_tf32(a) != a);
            }
        }
        if (norms) norms[row] = s;
    }
    if (not_exact && __any_sync(0xffffffffu, inexact) && (threadIdx.x & 31) == 0) atomicOr(not_exact, 1);
}

int launch_prep_rows(const float* x, int64_t rows, int dim, float* norms, float* hi, float* lo, int* not_tf32_exact,
                     cudaStream_t st) {
    if (rows <= 0) return VS_OK;
    const int64_t threads = rows * 8;
    const int64_t blocks = ceil_div64(threads, 256);
    prep_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, rows, dim, norms, hi, lo, not_tf32_exact);
    VSB_CUDA(cudaGetLastError());
    return VS_OK;
}

// Query prep of the certified path in one pass: ||q||^2 in the reference's order (the eight partial sums of
// compute_norm_avx2, cpu_baseline.cpp:99-107, kept by ONE thread per row, then added 0..7) and the batch's abs-max.
// Queries are few (<= a few MB): a thread per row with 16-byte loads beats the 8-threads-per-row base kernel here.
__global__ void __launch_bounds__(128) query_prep_kernel(const float* __restrict__ x, int64_t rows, float* __restrict__ norms,
                                                         float* __restrict__ absmax_out) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float mx = 0.f;
    if (row < rows) {
        const float4* v = reinterpret_cast<const float4*>(x + row * 128);
        float acc[8];
#pragma unroll
        for (int l = 0; l < 8; ++l) acc[l] = 0.f;
#pragma unroll 4
        for (int i = 0; i < 16; ++i) {
            const float4 a = __ldg(v + 2 * i), b = __ldg(v + 2 * i + 1);
            acc[0] = fmaf(a.x, a.x, acc[0]); acc[1] = fmaf(a.y, a.y, acc[1]);
            acc[2] = fmaf(a.z, a.z, acc[2]); acc[3] = fmaf(a.w, a.w, acc[3]);
            acc[4] = fmaf(b.x, b.x, acc[4]); acc[5] = fmaf(b.y, b.y, acc[5]);
            acc[6] = fmaf(b.z, b.z, acc[6]); acc[7] = fmaf(b.w, b.w, acc[7]);
            mx = fmaxf(mx, fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))));
            mx = fmaxf(mx, fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w))));
        }
        float s = acc[0];
#pragma unroll
        for (int l = 1; l < 8; ++l) s = __fadd_rn(s, acc[l]);
        norms[row] = s;
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0 && mx > 0.f) atomicMax(reinterpret_cast<int*>(absmax_out), __float_as_int(mx));
}
int launch_query_prep(const float* x, int64_t rows, float* norms, float* absmax_zeroed, cudaStream_t st) {
    if (rows <= 0) return VS_OK;
    query_prep_kernel<<<(unsigned)ceil_div64(rows, 128), 128, 0, st>>>(x, rows, norms, absmax_zeroed);
    VSB_CUDA(cudaGetLastError());
    return VS_OK;
}

__global__ void fill_f32_kernel(float* p, int64_t n, float v) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}
int launch_fill_f32(float* p, int64_t n, float v, cudaStream_t st) {
    if (n <= 0) return VS_OK;
    const int64_t blocks = ceil_div64(n, 256);
    fill_f32_kernel<<<(unsigned)(blocks > 1184 ? 1184 : blocks), 256, 0, st>>>(p, n, v);
    VSB_CUDA(cudaGetLastError());
    return VS_OK;
}

// ------------------------------------------------------------------------------------------------
// scaled fp16 operand copies for the candidate pass (TC_F16) and the per-call certification constants
// ------------------------------------------------------------------------------------------------
__global__ void absmax_f32_kernel(const float* __restrict__ x, int64_t count, float* __restrict__ out) {
    float mx = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
        mx = fmaxf(mx, fabsf(__ldg(x + i)));  // fmaxf drops NaN: garbage in, garbage out
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(mx));  // non-negative floats order as ints
}
int launch_absmax_f32(const float* x, int64_t count, float* out_zeroed, cudaStream_t st) {
    if (count <= 0) return VS_OK;
    const int64_t blocks = ceil_div64(count, 256);
    absmax_f32_kernel<<<(unsigned)(blocks > 148 * 8 ? 148 * 8 : blocks), 256, 0, st>>>(x, count, out_zeroed);
    VSB_CUDA(cudaGetLastError());
    return VS_OK;
}

// power-of-two scale that brings absmax into [2^10, 2^11): fp16 keeps 11 significant bits for everything above 2^-14 in
// scaled units (2^-25 of the largest magnitude), integers up to 2048 stay exact, and the norm term s_b^2 * bn / 2 <= 2^28
// fits three fp16 pieces (below)
__host__ __device__ inline float f16_scale_for(float absmax) {
    if (!(absmax > 0.f) || !(absmax < 3.0e38f)) return 1.0f;
    int e;
    frexpf(absmax, &e);  // absmax = m * 2^e, m in [0.5, 1)
    int se = 11 - e;
    se = se < -100 ? -100 : (se > 100 ? 100 : se);
    return ldexpf(1.0f, se);
}
float f16_scale_host(float absmax) { return f16_scale_for(absmax); }

// Base side of the norm block: U = s_b^2 * bn / 2 (<= 2^28) as three fp16 pieces of 11 bits each,
// U = 2^13 * p1 + 2^2 * p2 + 2^-8 * p3 (residual <= 2^-5: 33 bits cover the 24-bit fp32 norm exactly unless a piece lands in the
// fp16 subnormals, whose spacing is 2^-32 in U units).  Row layout {p1 x8, p2 x4, p3, 0, 0, 0}; rows >= n are zero (their
// columns are masked in the kernel's last tile).
__global__ void norm_pieces_kernel(const float* __restrict__ bnorm, int64_t n, int64_t n_pad, float s_b, __half* __restrict__ e) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_pad) return;
    __half p1 = __float2half_rn(0.f), p2 = p1, p3 = p1;
    if (r < n) {
        const double U = 0.5 * (double)bnorm[r] * (double)s_b * (double)s_b;
        p1 = __double2half(U * (1.0 / 8192.0));
        const double r1 = U - (double)__half2float(p1) * 8192.0;
        p2 = __double2half(r1 * 0.25);
        const double r2 = r1 - (double)__half2float(p2) * 4.0;
        p3 = __double2half(r2 * 256.0);
    }
    const __half z = __float2half_rn(0.f);
    __half row[TC_FOLD_COLS] = {p1, p1, p1, p1, p1, p1, p1, p1, p2, p2, p2, p2, p3, z, z, z};
    uint4* dst = reinterpret_cast<uint4*>(e + r * TC_FOLD_COLS);
    dst[0] = *reinterpret_cast<uint4*>(row);
    dst[1] = *reinterpret_cast<uint4*>(row + 8);
}
int launch_norm_pieces(const float* bnorm, int64_t n, int64_t n_pad, float s_b, void* e_half, cudaStream_t st) {
    if (n_pad <= 0) return VS_OK;
    norm_pieces_kernel<<<(unsigned)ceil_div64(n_pad, 256), 256, 0, st>>>(bnorm, n, n_pad, s_b, (__half*)e_half);
    VSB_CUDA(cudaGetLastError());
    return VS_OK;
}

// TcQueryParams (kernels.cuh): scale of the query copy, key factor, the two certification constants and the query side
// of the norm block.
//   |key_f16 - key_exact| <= cert_a * sqrt(qn) + cert_b   with, per element, |x^ - x| <= eps|x| + u (eps = 2^-11,
//   u = 2^-25 / scale: half the fp16 subnormal spacing), Cauchy-Schwarz for sum|q||x| <= sqrt(qn * bn_max) and
//   sum|v| <= sqrt(128 * ||v||^2), plus 2^-15 * (sum|q||x| + bn/2) for the tensor core's fp32 accumulation of the dot
//   product AND the norm block (in key units: 2^-15 * (2 sqrt(qn bn_max) + bn_max)), plus 2^-23 * bn_max for the residual of
//   the three-piece norm, times 1.02, plus the fp32 rounding of the refined distances themselves (4e-6 * (qn + bn_max), folded
//   into cert_b per query by the merge kernel).  tests/test_exact_gpu.py measures the actual error against this bound.
__global__ void tc_query_params_kernel(const float* __restrict__ q_absmax, float s_b, float bn_max, TcQueryParams* out,
                                       __half* __restrict__ qe) {
    __shared__ __half row[TC_FOLD_COLS];
    const float s_q = f16_scale_for(*q_absmax);
    if (threadIdx.x == 0) {
        const float eps = 1.0f / 2048.0f;
        const float A = 2.0f * eps + eps * eps + 1.0f / 32768.0f;
        const float u_q = ldexpf(1.0f, -25) / s_q, u_b = ldexpf(1.0f, -25) / s_b;
        int eq, eb;
        frexpf(s_q, &eq);
        frexpf(s_b, &eb);
        const int rho = eq - eb;  // s_q / s_b = 2^rho (both are powers of two)
        TcQueryParams r;
        r.s_q = s_q;
        r.key_unscale = 2.0f / (s_q * s_b);
        r.cert_a = 2.04f * (A * sqrtf(bn_max) + u_b * sqrtf(128.0f));
        r.cert_b = 2.04f * (u_q * sqrtf(128.0f * bn_max) + 128.0f * u_q * u_b) + 1.02f * bn_max * (1.0f / 32768.0f + 1.0f / 8388608.0f);
        r.bn_max = bn_max;
        r.fold_ok = (rho <= 5 && rho >= -16) ? 1 : 0;
        *out = r;
        const __half a1 = __float2half_rn(r.fold_ok ? ldexpf(1.0f, rho + 10) : 0.f);
        const __half a2 = __float2half_rn(r.fold_ok ? ldexpf(1.0f, rho) : 0.f);
        const __half a3 = __float2half_rn(r.fold_ok ? ldexpf(1.0f, rho - 8) : 0.f);  // down to 2^-24: exact as a subnormal
        const __half z = __float2half_rn(0.f);
        for (int i = 0; i < TC_FOLD_COLS; ++i) row[i] = i < 8 ? a1 : (i < 12 ? a2 : (i == 12 ? a3 : z));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 128 * TC_FOLD_COLS; i += blockDim.x) qe[i] = row[i % TC_FOLD_COLS];
}
int launch_tc_query_params(const float* q_absmax, float s_b, float bn_max, TcQueryParams* out, void* qe_half, cudaStream_t st) {
    tc_query_params_kernel<<<1, 128, 0, st>>>(q_absmax, s_b, bn_max, out, (__half*)qe_half);
    VSB_CUDA(cudaGetLastError());
    return VS_OK;
}

// out[i] = fp16_rn(x[i] * scale); scale either given or -s_q read from *scale_dev (the query scale is derived on the device;
// the query copy is NEGATED so that the accumulator  -s_q s_b q.x + s_q s_b bn / 2  is the key times s_q s_b / 2)
__global__ void to_half_scaled_kernel(const float* __restrict__ x, int64_t count, float scale, const TcQueryParams* scale_dev,
                                      __half* __restrict__ out) {
    const float s = scale_dev ? -scale_dev->s_q : scale;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < count; i += (int64_t)gridDim.x * blockDim.x * 4) {
        if (i + 3 < count) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(x + i));
            __half2 a = __floats2half2_rn(v.x * s, v.y * s);
            __half2 b = __floats2half2_rn(v.z * s, v.w * s);
            uint2 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&a);
            pk.y = *reinterpret_cast<uint32_t*>(&b);
            *reinterpret_cast<uint2*>(out + i) = pk;
        } else {
            for (int e = 0; e < 4 && i + e < count; ++e) out[i + e] = __float2half_rn(__ldg(x + i + e) * s);
        }
    }
}
int launch_to_half_scaled(const float* x, int64_t count, float scale, const TcQueryParams* scale_dev, void* out, cudaStream_t st) {
    if (count <= 0) return VS_OK;
    const int64_t blocks = ceil_div64(count, 4 * 256);
    to_half_scaled_kernel<<<(unsigned)(blocks > 148 * 16 ? 148 * 16 : blocks), 256, 0, st>>>(x, count, scale, scale_dev,
                                                                                         (__half*)out);
    VSB_CUDA(cudaGetLastError());
    return VS_OK;
}

// rows gather / result scatter for the fallback of uncertified queries
__global__ void gather_rows_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx, int n_idx, float* __restrict__ dst) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= n_idx) return;
    const int lane = threadIdx.x & 31;
    reinterpret_cast<float4*>(dst + (size_t)r * 128)[lane] = __ldg(reinterpret_cast<const float4*>(src + (size_t)idx[r] * 128) + lane);
}
__global__ void scatter_results_kernel(const float* __restrict__ key, const int32_t* __restrict__ id, const int32_t* __restrict__ idx,
                                       int n_idx, int k, float* __restrict__ out_key, int32_t* __restrict__ out_id) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_idx * k) return;
    const int r = i / k, c = i % k;
    out_key[(size_t)idx[r] * k + c] = key[i];
    out_id[(size_t)idx[r] * k + c] = id[i];
}
int launch_gather_rows(const float* src, const int32_t* idx, int n_idx, float* dst, cudaStream_t st) {
    if (n_idx <= 0) return VS_OK;
    gather_rows_kernel<<<(unsigned)((n_idx + 3) / 4), 128, 0, st>>>(src, idx, n_idx, dst);
    VSB_CUDA(cudaGetLastError());
    return VS_OK;
}
int launch_scatter_results(const float* key, const int32_t* id, const int32_t* idx, int n_idx, int k, float* out_key,
                           int32_t* out_id, cudaStream_t st) {
    if (n_idx <= 0) return VS_OK;
    scatter_results_kernel<<<(unsigned)((n_idx * k + 255) / 256), 256, 0, st>>>(key, id, idx, n_idx, k, out_key, out_id);
    VSB_CUDA(cudaGetLastError());
    return VS_OK;
}

// ------------------------------------------------------------------------------------------------
// K3: merge sorted partial lists (+ optional exact fp32 refine).  One warp per query; lane j folds lists
// j, j+32, ... into a register list, then the warp pops the global minimum `nsel` times.
//
// Refine (rf.base != nullptr): the tensor-core accumulator does not round to nearest (measured on B200: a
// systematic ~2e-6 relative truncation bias of the dot product on continuous data), so the fused kernel's keys
// are treated as a candidate ranking; the nsel selected candidates get their distance recomputed in plain fp32
// — rows read cooperatively, four FMAs per lane, a fixed pairwise lane tree (warp_refine_dots), then (qn + bn) - 2*dot
// (cpu_baseline.cpp:241) — and are re-sorted by (distance, id).  Reported distances are then fp32-faithful regardless of which
// kernel generated the candidates.
// ------------------------------------------------------------------------------------------------
// Lists that live in per-shard exchange blocks (block = ids [nq x k] | keys [nq x k] | 16-byte trailer, see
// vs_topk_block_bytes): list l starts list_stride 4-byte words after list l-1 (0 = densely packed [list][nq][len]);
// trailer != nullptr: word 0 of every block's trailer (the shard's uncertified-query count) is summed into *total_out.
struct BlockArgs {
    size_t list_stride;
    const int32_t* trailer;
    int32_t* total_out;
};

struct RefineArgs {
    const float* base;   // [n x 128] fp32, or nullptr = no refine
    const float* bnorm;
    const float* q;      // [nq x 128]
    const float* qnorm;
    // certification of a bounded-error candidate pass (TC_F16), or qp == nullptr: a query is certified when no row
    // outside its nsel candidates can beat the k-th refined distance:  qn + a_last - E_q > D_k  (a_last = largest
    // candidate key, D_k = k-th smallest refined distance), or when fewer than nsel candidates exist at all.
    const TcQueryParams* qp;
    int32_t* uncert_count;  // number of uncertified queries
    int32_t* uncert_list;   // their indices
    const int32_t* ivf_idmap;  // IVF re-score mode (see launch_merge_lists)
    // threshold-filter candidates (filter_merge_kernel): `lastk` passed to merge_tail is a key bound B such that every row
    // that is NOT among the selected candidates has key >= B; the query is always checked (there is no "fewer than nsel rows
    // exist" shortcut) and fails outright when its candidate array overflowed
    bool filter = false;
    bool overflow = false;
};

// q . x in the reference's NEON order (IVFIndex.cpp:278-357 computeDotProductsContiguous): four accumulators by d mod 4,
// FMA over d = 0, 4, 8, ..., then (l0 + l1) + (l2 + l3) — bit-identical to the scan kernels and the CPU restatement
__device__ __forceinline__ float dot_neon4_128(const float* __restrict__ q, const float* __restrict__ x) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 8
    for (int c4 = 0; c4 < 32; ++c4) {
        const float4 qv = __ldg(reinterpret_cast<const float4*>(q) + c4);
        const float4 xv = __ldg(reinterpret_cast<const float4*>(x) + c4);
        a0 = fmaf(qv.x, xv.x, a0);
        a1 = fmaf(qv.y, xv.y, a1);
        a2 = fmaf(qv.z, xv.z, a2);
        a3 = fmaf(qv.w, xv.w, a3);
    }
    return __fadd_rn(__fadd_rn(a0, a1), __fadd_rn(a2, a3));
}

// The same dot product with the row staged in shared memory (16-byte aligned)
__device__ __forceinline__ float dot_neon4_128_smem(const float* __restrict__ q, const float* x_smem) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 8
    for (int c4 = 0; c4 < 32; ++c4) {
        const float4 qv = __ldg(reinterpret_cast<const float4*>(q) + c4);
        const float4 xv = *reinterpret_cast<const float4*>(x_smem + 4 * c4);
        a0 = fmaf(qv.x, xv.x, a0);
        a1 = fmaf(qv.y, xv.y, a1);
        a2 = fmaf(qv.z, xv.z, a2);
        a3 = fmaf(qv.w, xv.w, a3);
    }
    return __fadd_rn(__fadd_rn(a0, a1), __fadd_rn(a2, a3));
}
constexpr int kRowPitch = 132;  // floats per staged row: 528 B, so that eight lanes' 16-byte reads hit distinct bank groups

// Exact fp32 dot products q . x_id of a warp's (up to 32) candidates, lane r holding candidate r's local row id (or -1).
// Every row is read COOPERATIVELY — lane i takes components 4i .. 4i+3 (one coalesced 512-byte request per row) — and the
// 32 partial sums are added in a FIXED tree (lane i + lane i+16, then +8, +4, +2, +1), so the value of a (query, row) pair
// does not depend on the lane the candidate sits in: every kernel path, shard count and list position gives the same bits.
// Returns, in lane r, the dot product of candidate r.
__device__ __forceinline__ float warp_refine_dots(const float* __restrict__ q_row, const float* __restrict__ base, int nsel,
                                                  int32_t myid, int lane) {
    const float4 qv = __ldg(reinterpret_cast<const float4*>(q_row) + lane);
    float mydot = 0.f;
    for (int r0 = 0; r0 < nsel; r0 += 4) {
        float4 x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {  // four independent row loads in flight
            const int32_t id = __shfl_sync(0xffffffffu, myid, (r0 + u) & 31);
            x[u] = (r0 + u < nsel && id >= 0) ? __ldg(reinterpret_cast<const float4*>(base + (size_t)id * 128) + lane)
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float p = __fmul_rn(x[u].x, qv.x);
            p = fmaf(x[u].y, qv.y, p);
            p = fmaf(x[u].z, qv.z, p);
            p = fmaf(x[u].w, qv.w, p);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) p = __fadd_rn(p, __shfl_down_sync(0xffffffffu, p, o));
            const float tot = __shfl_sync(0xffffffffu, p, 0);
            if (lane == r0 + u) mydot = tot;
        }
    }
    return mydot;
}

// Common tail of the merge kernels.  In: lane r holds candidate r (myk, myid; -1 = none) of the nsel selected by the candidate
// keys, lastk / lasti = the last (largest) selected candidate.  Does: lower bound for the next pass, exact fp32 refine,
// (distance, id) ranking, certification, output rows.
__device__ __forceinline__ void merge_tail(float myk, int32_t myid, float lastk, int32_t lasti, int64_t q, int lane, int nsel, int k,
                                           int64_t id_base, int neg_out, float* __restrict__ out_key, int32_t* __restrict__ out_id,
                                           int out_stride, int out_off, float* __restrict__ lb_key_out,
                                           int32_t* __restrict__ lb_id_out, const RefineArgs& rf,
                                           float* row_scratch = nullptr /* per-warp shared memory, nsel x kRowPitch floats */) {
    const float INF = __int_as_float(0x7f800000);
    if (lb_key_out && lane == 0) {  // exclusive lower bound for the next pass (local ids, candidate-ranking keys)
        lb_key_out[q] = lasti >= 0 ? lastk : INF;
        lb_id_out[q] = lasti >= 0 ? lasti : 0x7fffffff;
    }
    if (rf.ivf_idmap) {
        if (row_scratch) {
            // the candidates' rows come in with coalesced 512-byte loads (lane i takes components 4i .. 4i+3 of every row); each
            // lane then walks ITS row in the reference's order out of shared memory
            __syncwarp();
            for (int r = 0; r < nsel; ++r) {
                const int32_t id = __shfl_sync(0xffffffffu, myid, r);
                if (id >= 0)
                    *reinterpret_cast<float4*>(row_scratch + r * kRowPitch + 4 * lane) =
                        __ldg(reinterpret_cast<const float4*>(rf.base + (size_t)id * 128) + lane);
            }
            __syncwarp();
            if (myid >= 0) {
                myk = -dot_neon4_128_smem(rf.q + (size_t)q * 128, row_scratch + lane * kRowPitch);
                myid = __ldg(rf.ivf_idmap + myid);
            }
        } else if (myid >= 0) {
            myk = -dot_neon4_128(rf.q + (size_t)q * 128, rf.base + (size_t)myid * 128);
            myid = __ldg(rf.ivf_idmap + myid);
        }
    } else if (rf.base) {
        const float dot = warp_refine_dots(rf.q + (size_t)q * 128, rf.base, nsel, myid, lane);
        if (myid >= 0) myk = fmaf(-2.0f, dot, __fadd_rn(__ldg(rf.qnorm + q), __ldg(rf.bnorm + myid)));
    }
    // rank by counting (ids are unique; padding sorts last and is never written)
    int rank = 0;
#pragma unroll 1
    for (int j = 0; j < nsel; ++j) {
        const float ok = __shfl_sync(0xffffffffu, myk, j);
        const int32_t oi = __shfl_sync(0xffffffffu, myid, j);
        rank += (oi >= 0 && pair_less(ok, oi, myk, myid)) ? 1 : 0;
    }
    const int n_valid = __popc(__ballot_sync(0xffffffffu, myid >= 0));
    if (rf.qp) {
        // lastk = key of the last popped candidate = largest candidate key (lists pop in ascending key order)
        const unsigned holder = __ballot_sync(0xffffffffu, myid >= 0 && rank == k - 1);
        const float dk = __shfl_sync(0xffffffffu, myk, holder ? __ffs(holder) - 1 : 0);
        if (lane == 0 && (rf.filter || n_valid == nsel || !rf.qp->fold_ok)) {
            const float qn = __ldg(rf.qnorm + q);
            const float e = rf.qp->cert_a * sqrtf(qn) + rf.qp->cert_b + 4e-6f * (qn + rf.qp->bn_max);
            // candidate keys are in accumulator units (x s_q s_b / 2, an exact power-of-two factor)
            const bool ok = rf.qp->fold_ok && !rf.overflow && holder != 0 && (qn + lastk * rf.qp->key_unscale) - e > dk;
            if (!ok) rf.uncert_list[atomicAdd(rf.uncert_count, 1)] = (int32_t)q;
        }
    }
    float* ok_row = out_key + q * out_stride + out_off;
    int32_t* oi_row = out_id + q * out_stride + out_off;
    if (myid >= 0 && rank < k) {
        ok_row[rank] = neg_out ? -myk : myk;
        oi_row[rank] = (int32_t)(myid + id_base);
    }
    for (int r = n_valid + lane; r < k; r += 32) {
        ok_row[r] = neg_out ? -INF : INF;
        oi_row[r] = -1;
    }
}

template <int KTOP>
__global__ void __launch_bounds__(128) merge_lists_kernel(const float* __restrict__ part_key,
                                                          const int32_t* __restrict__ part_id, int n_lists, int64_t nq,
                                                          int list_len, int nsel, int k, int64_t id_base, int neg_in,
                                                          int neg_out, float* __restrict__ out_key,
                                                          int32_t* __restrict__ out_id, int out_stride, int out_off,
                                                          float* __restrict__ lb_key_out, int32_t* __restrict__ lb_id_out,
                                                          RefineArgs rf, BlockArgs ba) {
    __shared__ float s_key[4][32];
    __shared__ int32_t s_id[4][32];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
    if (ba.trailer && blockIdx.x == 0 && threadIdx.x == 0) {  // exchange blocks: total of the shards' uncertified counts
        int32_t tot = 0;
        for (int l = 0; l < n_lists; ++l) tot += ba.trailer[(size_t)l * ba.list_stride];
        *ba.total_out = tot;
    }
    if (q >= nq) return;
    const size_t lstride = ba.list_stride ? ba.list_stride : (size_t)nq * list_len;
    RegTopK<KTOP> L;
    L.init();
    bool first = true;
    for (int l = lane; l < n_lists; l += 32) {
        const float* pk = part_key + (size_t)l * lstride + (size_t)q * list_len;
        const int32_t* pi = part_id + (size_t)l * lstride + (size_t)q * list_len;
        if (first) {  // lists are already sorted: adopt the first one as is
#pragma unroll
            for (int i = 0; i < KTOP; ++i) {
                if (i < list_len) {
                    const float kk = pk[i];
                    L.key[i] = neg_in ? -kk : kk;
                    L.id[i] = pi[i];
                }
            }
            first = false;
        } else {
#pragma unroll 1
            for (int i = 0; i < list_len; ++i) {
                const int32_t id = pi[i];
                if (id < 0) break;
                const float kk = neg_in ? -pk[i] : pk[i];
                // the partial list is sorted: once an element does not make it, none of the following ones can
                if (!pair_less(kk, id, L.key[KTOP - 1], L.id[KTOP - 1])) break;
                L.insert_any(kk, id);
            }
        }
    }
    const float INF = __int_as_float(0x7f800000);
    float lastk = INF;
    int32_t lasti = -1;
    for (int r = 0; r < nsel; ++r) {
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
            s_key[wib][r] = hid >= 0 ? hk : INF;
            s_id[wib][r] = hid;
        }
        lastk = hk;
        lasti = hid;
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
    __syncwarp();
    // ---- this lane's candidate (nsel <= 32)
    float myk = INF;
    int32_t myid = -1;
    if (lane < nsel) {
        myk = s_key[wib][lane];
        myid = s_id[wib][lane];
    }
    merge_tail(myk, myid, lastk, lasti, q, lane, nsel, k, id_base, neg_out, out_key, out_id, out_stride, out_off, lb_key_out,
               lb_id_out, rf);
}

// The same for n_lists <= 96 (every tensor-core / IVF / exchange merge): the lists are staged in shared memory with coalesced
// loads (a list = 128 contiguous bytes), lane j walks list j with a cursor, and every round the warp takes the smallest head
// (shuffle arg-min) — no register-list shifting, ~1/4 of the instructions of the general kernel.
__global__ void __launch_bounds__(128) merge_small_kernel(const float* __restrict__ part_key, const int32_t* __restrict__ part_id,
                                                          int n_lists, int64_t nq, int list_len, int nsel, int k, int64_t id_base,
                                                          int neg_in, int neg_out, float* __restrict__ out_key,
                                                          int32_t* __restrict__ out_id, int out_stride, int out_off,
                                                          float* __restrict__ lb_key_out, int32_t* __restrict__ lb_id_out,
                                                          RefineArgs rf, BlockArgs ba, int warp_floats) {
    extern __shared__ float sm_lists[];  // per warp (warp_floats each): keys [n_lists][32] then ids [n_lists][32]
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
    if (ba.trailer && blockIdx.x == 0 && threadIdx.x == 0) {
        int32_t tot = 0;
        for (int l = 0; l < n_lists; ++l) tot += ba.trailer[(size_t)l * ba.list_stride];
        *ba.total_out = tot;
    }
    if (q >= nq) return;
    const float INF = __int_as_float(0x7f800000);
    const size_t lstride = ba.list_stride ? ba.list_stride : (size_t)nq * list_len;
    float* sk = sm_lists + (size_t)wib * warp_floats;
    int32_t* si = reinterpret_cast<int32_t*>(sk + n_lists * 32);
    // staging: element e = l * list_len + i of the query's n_lists * list_len entries; eight independent loads per lane in
    // flight (a serial load -> store loop is latency-bound: 96 lists took ~0.3 ms per 10 K queries)
    {
        const int total = n_lists * list_len;
        for (int e0 = 0; e0 < total; e0 += 8 * 32) {
            float kk[8];
            int32_t ii[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int e = e0 + u * 32 + lane;
                kk[u] = INF;
                ii[u] = -1;
                if (e < total) {
                    const int l = e / list_len, i = e - l * list_len;
                    kk[u] = part_key[(size_t)l * lstride + (size_t)q * list_len + i];
                    ii[u] = part_id[(size_t)l * lstride + (size_t)q * list_len + i];
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int e = e0 + u * 32 + lane;
                if (e < total) {
                    const int l = e / list_len, i = e - l * list_len;
                    sk[l * 32 + i] = ii[u] >= 0 ? (neg_in ? -kk[u] : kk[u]) : INF;
                    si[l * 32 + i] = ii[u];
                }
            }
        }
        // (slots [list_len, 32) of a list stay unwritten: the cursors stop at list_len)
    }
    __syncwarp();
    // lane j walks lists j, j + 32, j + 64 (n_lists <= 96) with one cursor each
    int cur[3] = {0, 0, 0};
    float hk[3];
    int32_t hid[3];
#pragma unroll
    for (int u = 0; u < 3; ++u) {
        const int l = lane + 32 * u;
        hid[u] = l < n_lists ? si[l * 32] : -1;
        hk[u] = hid[u] >= 0 ? sk[l * 32] : INF;
    }
    float myk = INF, lastk = INF;
    int32_t myid = -1, lasti = -1;
    for (int r = 0; r < nsel; ++r) {
        // this lane's best head, then the warp's
        int bu = 0;
        if (pair_less(hk[1], hid[1], hk[bu], hid[bu])) bu = 1;
        if (pair_less(hk[2], hid[2], hk[bu], hid[bu])) bu = 2;
        float bk = bu == 0 ? hk[0] : (bu == 1 ? hk[1] : hk[2]);
        int32_t bi = bu == 0 ? hid[0] : (bu == 1 ? hid[1] : hid[2]);
        int src = lane;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ok = __shfl_xor_sync(0xffffffffu, bk, o);
            const int32_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
            const int os = __shfl_xor_sync(0xffffffffu, src, o);
            if (pair_less(ok, oi, bk, bi)) {
                bk = ok;
                bi = oi;
                src = os;
            }
        }
        if (lane == r) {
            myk = bi >= 0 ? bk : INF;
            myid = bi;
        }
        lastk = bk;
        lasti = bi;
        if (src == lane && bi >= 0) {
#pragma unroll
            for (int u = 0; u < 3; ++u)
                if (u == bu) {
                    const int l = lane + 32 * u;
                    ++cur[u];
                    hid[u] = cur[u] < list_len ? si[l * 32 + cur[u]] : -1;
                    hk[u] = hid[u] >= 0 ? sk[l * 32 + cur[u]] : INF;
                }
        }
    }
    // IVF re-score: the lists are consumed; their shared memory (>= nsel x kRowPitch floats per warp, see the launch) takes the rows
    merge_tail(myk, myid, lastk, lasti, q, lane, nsel, k, id_base, neg_out, out_key, out_id, out_stride, out_off, lb_key_out,
               lb_id_out, rf, rf.ivf_idmap ? sm_lists + (size_t)wib * warp_floats : nullptr);
}

// ---- threshold-filter candidate pass (exact_tc.cuh, TC_F16) ----------------------------------------------------------
// thr[q] = the m-th smallest of the query's group minima (sample pass): at least m sampled rows have a key <= thr[q], so about
// m * (rows / sampled rows) rows of the whole base lie below it.  Fewer than m finite minima: +inf (everything is a
// candidate; the candidate array then overflows unless the base is tiny).  Also zeroes the candidate counters.
// One warp per query, eight consecutive queries per block (they share the 32-byte sectors of smin[g][.]).
__global__ void __launch_bounds__(256) tc_select_thr_kernel(const float* __restrict__ smin, int n_groups, int nq, int m,
                                                            float* __restrict__ thr, int32_t* __restrict__ cand_cnt) {
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (q >= nq) return;
    const float INF = __int_as_float(0x7f800000);
    constexpr int NV = 12;  // group minima held per lane (n_groups <= 384: read once); more are re-read every round
    float v[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int g = lane + 32 * j;
        v[j] = g < n_groups ? smin[(size_t)g * nq + q] : INF;
    }
    // m rounds of "smallest (value, group) after the previous pick"; (INF, .) entries stand for missing groups
    float pk = -INF;
    int pg = -1;
    for (int r = 0; r < m; ++r) {
        float bk = INF;
        int bg = 0x7fffffff;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int g = lane + 32 * j;
            const bool after = v[j] > pk || (v[j] == pk && g > pg);
            if (after && v[j] < INF && (v[j] < bk || (v[j] == bk && g < bg))) {
                bk = v[j];
                bg = g;
            }
        }
        for (int g = lane + 32 * NV; g < n_groups; g += 32) {
            const float x = smin[(size_t)g * nq + q];
            const bool after = x > pk || (x == pk && g > pg);
            if (after && x < INF && (x < bk || (x == bk && g < bg))) {
                bk = x;
                bg = g;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ok = __shfl_xor_sync(0xffffffffu, bk, o);
            const int og = __shfl_xor_sync(0xffffffffu, bg, o);
            if (ok < bk || (ok == bk && og < bg)) {
                bk = ok;
                bg = og;
            }
        }
        pk = bk;
        pg = bg;
        if (bg == 0x7fffffff) {  // fewer than m groups with a finite minimum
            pk = INF;
            break;
        }
    }
    if (lane == 0) {
        thr[q] = pk;
        cand_cnt[q] = 0;
    }
}
int launch_tc_select_thr(const float* smin, int n_groups, int nq, int m, float* thr, int32_t* cand_cnt, cudaStream_t st) {
    if (nq <= 0) return VS_OK;
    tc_select_thr_kernel<<<(unsigned)ceil_div64(nq, 8), 256, 0, st>>>(smin, n_groups, nq, m, thr, cand_cnt);
    VSB_CUDA(cudaGetLastError());
    return VS_OK;
}
// thr[q] = +inf, counters zeroed: bases of at most cand_cap rows need no sample pass
__global__ void tc_fill_thr_kernel(int nq, float* __restrict__ thr, int32_t* __restrict__ cand_cnt) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < nq) {
        thr[q] = __int_as_float(0x7f800000);
        cand_cnt[q] = 0;
    }
}
int launch_tc_fill_thr(int nq, float* thr, int32_t* cand_cnt, cudaStream_t st) {
    if (nq <= 0) return VS_OK;
    tc_fill_thr_kernel<<<(unsigned)ceil_div64(nq, 256), 256, 0, st>>>(nq, thr, cand_cnt);
    VSB_CUDA(cudaGetLastError());
    return VS_OK;
}

__device__ __forceinline__ uint32_t key_to_u32(float f) {  // monotone: a < b  <=>  key_to_u32(a) < key_to_u32(b)
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float u32_to_key(uint32_t u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

// A pivot P over the warp's NPER x 32 keys u[] (0xffffffff = empty slot) with at most 32 keys below it, found by a radix
// descent from the highest bit in which the keys differ.  *n_below = keys < P.  The descent stops early once >= 24 keys lie
// below the current bucket, min_keep = 24 in the certified search (any P with <= 32 keys below it is a valid bound; a few
// candidates fewer cost nothing); a descent
// that runs through bit 0 returns the exact 32nd smallest key and *n_ties = how many keys == P complete the 32.
template <int NPER>
__device__ __forceinline__ uint32_t warp_pivot32(const uint32_t (&u)[NPER], int* n_below, int* n_ties, int min_keep) {
    uint32_t lo = 0xffffffffu, hi = 0u;
#pragma unroll
    for (int j = 0; j < NPER; ++j) {
        if (u[j] != 0xffffffffu) {
            lo = min(lo, u[j]);
            hi = max(hi, u[j]);
        }
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    const uint32_t diff = lo ^ hi;
    int need = 32;
    uint32_t prefix = lo;
    if (diff == 0) {  // all keys equal
        *n_below = 0;
        *n_ties = 32;
        return lo;
    }
    const int top = 31 - __clz(diff);
    prefix = top == 31 ? 0u : (lo & ~((2u << top) - 1u));  // the common high bits
    for (int bit = top; bit >= 0; --bit) {
        const uint32_t hmask = bit == 31 ? 0u : ~((2u << bit) - 1u);
        int c0 = 0;
#pragma unroll
        for (int j = 0; j < NPER; ++j) c0 += (((u[j] ^ prefix) & hmask) == 0u && !((u[j] >> bit) & 1u)) ? 1 : 0;
        c0 = __reduce_add_sync(0xffffffffu, c0);
        if (need > c0) {  // the whole 0-branch lies below the pivot
            need -= c0;
            prefix |= 1u << bit;
            if (32 - need >= min_keep) {
                *n_below = 32 - need;
                *n_ties = 0;
                return prefix;
            }
        }
    }
    *n_below = 32 - need;
    *n_ties = need;
    return prefix;
}

// Merge step of the threshold-filter candidate pass: one warp per query.  The query's candidates (every row with
// key < thr[q], unordered, cnt of them) are cut down to <= 32 by key, then merge_tail refines them in exact fp32, ranks,
// certifies against the bound B (= thr[q] when all candidates were kept, else the pivot key: no row outside the kept
// ones has a key below it) and writes the k results.
// Also the merge of the tensor-core IVF scan (rf.ivf_idmap: candidates = what the (query, probe slot) lists kept, thr ==
// nullptr, at least min_keep = k + 2 of them survive the cut and are re-scored in the reference's order out of rows staged
// in dynamic shared memory).  More than 1536 candidates (IVF with k > 14 only, rare): the same descent with the keys re-read from memory.
template <bool BIG>  // BIG: register-resident selection up to 1536 candidates (IVF), else up to 512
__global__ void __launch_bounds__(128) filter_merge_kernel(const uint2* __restrict__ cand, const int32_t* __restrict__ cand_cnt,
                                                           int cap, const float* __restrict__ thr, int64_t nq, int k, int64_t id_base,
                                                           int neg_out, int min_keep, float* __restrict__ out_key,
                                                           int32_t* __restrict__ out_id, int out_stride, RefineArgs rf) {
    extern __shared__ __align__(16) float fm_rows[];  // IVF re-score only: [4][32 x kRowPitch]
    __shared__ uint32_t s_key[4][32];
    __shared__ int32_t s_id[4][32];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int64_t q = (int64_t)blockIdx.x * 4 + wib;
    if (q >= nq) return;
    const float INF = __int_as_float(0x7f800000);
    const int c_raw = __ldg(cand_cnt + q);
    const int c = min(c_raw, cap);
    const uint2* src = cand + (size_t)q * cap;
    float myk = INF, bound = thr ? __ldg(thr + q) : INF;
    int32_t myid = -1;
    const unsigned lt_mask = (1u << lane) - 1u;
    if (c <= 32) {
        if (lane < c) {
            const uint2 e = __ldcg(src + lane);
            myk = __uint_as_float(e.x);
            myid = (int32_t)e.y;
        }
    } else if (c <= (BIG ? 1536 : 512)) {
        uint32_t pivot;
        int n_below, n_ties;
        auto run = [&](auto tag) {
            constexpr int NPER = decltype(tag)::value;
            uint32_t u[NPER];
            int32_t id[NPER];
#pragma unroll
            for (int j = 0; j < NPER; ++j) {
                const int i = lane + 32 * j;
                u[j] = 0xffffffffu;
                id[j] = -1;
                if (i < c) {
                    const uint2 e = __ldcg(src + i);
                    u[j] = key_to_u32(__uint_as_float(e.x));
                    id[j] = (int32_t)e.y;
                }
            }
            pivot = warp_pivot32<NPER>(u, &n_below, &n_ties, min_keep);
            int base = 0;
#pragma unroll
            for (int j = 0; j < NPER; ++j) {
                const bool lt = u[j] < pivot;
                const unsigned b = __ballot_sync(0xffffffffu, lt);
                if (lt) {
                    const int pos = base + __popc(b & lt_mask);
                    s_key[wib][pos] = u[j];
                    s_id[wib][pos] = id[j];
                }
                base += __popc(b);
            }
            if (n_ties > 0) {
#pragma unroll
                for (int j = 0; j < NPER; ++j) {
                    const bool eq = u[j] == pivot;
                    const unsigned b = __ballot_sync(0xffffffffu, eq);
                    if (eq) {
                        const int pos = base + __popc(b & lt_mask);
                        if (pos < 32) {
                            s_key[wib][pos] = u[j];
                            s_id[wib][pos] = id[j];
                        }
                    }
                    base = min(32, base + __popc(b));
                }
            }
            __syncwarp();
            if (lane < base) {
                myk = u32_to_key(s_key[wib][lane]);
                myid = s_id[wib][lane];
            }
        };
        if (c <= 128)
            run(std::integral_constant<int, 4>{});
        else if (c <= 256)
            run(std::integral_constant<int, 8>{});
        else if (c <= 512)
            run(std::integral_constant<int, 16>{});
        else if constexpr (BIG) {
            if (c <= 1024)
                run(std::integral_constant<int, 32>{});
            else
                run(std::integral_constant<int, 48>{});
        }
        bound = fminf(bound, u32_to_key(pivot));
    } else {
        // exact 32nd smallest key by a full radix descent over the keys in memory, then the same compaction
        uint32_t prefix = 0u;
        int need = 32;
        for (int bit = 31; bit >= 0; --bit) {
            const uint32_t m = (bit == 31 ? 0u : ~((2u << bit) - 1u)) | (1u << bit);
            int c0 = 0;
            for (int i = lane; i < c; i += 32) c0 += (((key_to_u32(__uint_as_float(__ldcg(&src[i].x))) ^ prefix) & m) == 0u) ? 1 : 0;
            c0 = __reduce_add_sync(0xffffffffu, c0);
            if (need > c0) {
                need -= c0;
                prefix |= 1u << bit;
            }
        }
        int base = 0, ties = 32 - need;  // keys < prefix land in [0, 32 - need), ties behind them
        for (int i0 = 0; i0 < c; i0 += 32) {
            const int i = i0 + lane;
            uint32_t u = 0xffffffffu;
            int32_t id = -1;
            if (i < c) {
                const uint2 e = __ldcg(src + i);
                u = key_to_u32(__uint_as_float(e.x));
                id = (int32_t)e.y;
            }
            const bool lt = u < prefix, eq = u == prefix && i < c;
            const unsigned bl = __ballot_sync(0xffffffffu, lt), be = __ballot_sync(0xffffffffu, eq);
            if (lt) {
                const int pos = base + __popc(bl & lt_mask);
                s_key[wib][pos] = u;
                s_id[wib][pos] = id;
            }
            if (eq) {
                const int pos = ties + __popc(be & lt_mask);
                if (pos < 32) {
                    s_key[wib][pos] = u;
                    s_id[wib][pos] = id;
                }
            }
            base += __popc(bl);
            ties = min(32, ties + __popc(be));
        }
        __syncwarp();
        myk = u32_to_key(s_key[wib][lane]);  // c > 1536 >= 32: all 32 slots are filled
        myid = s_id[wib][lane];
        bound = fminf(bound, u32_to_key(prefix));
    }
    rf.filter = true;
    rf.overflow = c_raw > cap;
    merge_tail(myk, myid, bound, 0, q, lane, 32, k, id_base, neg_out, out_key, out_id, out_stride, 0, nullptr, nullptr, rf,
               rf.ivf_idmap ? fm_rows + (size_t)wib * 32 * kRowPitch : nullptr);
}

// cand [nq][cap] {key bits, local id}, cand_cnt [nq], thr [nq] -> out [nq][out_stride] (k written per query).  rf_base == nullptr:
// no refine and no certification (debug: the candidates as the tensor-core pass ranked them)
int launch_filter_merge(const void* cand, const int32_t* cand_cnt, int cap, const float* thr, int64_t nq, int k, int64_t id_base,
                        float* out_key, int32_t* out_id, int out_stride, const float* rf_base, const float* rf_bnorm, const float* rf_q,
                        const float* rf_qnorm, const TcQueryParams* cert_qp, int32_t* uncert_count, int32_t* uncert_list,
                        cudaStream_t st) {
    if (nq <= 0) return VS_OK;
    if (k > 32 || cap > 512 || cap < 32) return fail(VS_ERR_INVALID, "filter merge: need k <= 32 and 32 <= cap <= 512");
    if (cert_qp && (!rf_base || !uncert_count || !uncert_list)) return fail(VS_ERR_INVALID, "merge: certification needs the refine");
    RefineArgs rf{rf_base, rf_bnorm, rf_q, rf_qnorm, cert_qp, uncert_count, uncert_list, nullptr};
    filter_merge_kernel<false><<<(unsigned)ceil_div64(nq, 4), 128, 0, st>>>(reinterpret_cast<const uint2*>(cand), cand_cnt, cap, thr, nq, k, id_base,
                                                                    0, 24, out_key, out_id, out_stride, rf);
    VSB_CUDA(cudaGetLastError());
    return VS_OK;
}

// The merge of the tensor-core IVF scan: candidates {key = -2 q.x as the tensor core saw it, row position}, at most cap per
// query by construction (no overflow); the best >= k + 2 are re-scored in the reference's order (vectors = list-contiguous
// rows, ivf_idmap = position -> original id) and the k best written, scores descending
int launch_filter_merge_ivf(const void* cand, const int32_t* cand_cnt, int cap, int64_t nq, int k, float* out_scores, int32_t* out_ids,
                            const float* vectors, const float* q, const int32_t* ivf_idmap, cudaStream_t st) {
    if (nq <= 0) return VS_OK;
    if (k > 30) return fail(VS_ERR_INVALID, "ivf filter merge: need k <= 30");
    RefineArgs rf{vectors, nullptr, q, nullptr, nullptr, nullptr, nullptr, ivf_idmap};
    const int smem = 4 * 32 * kRowPitch * (int)sizeof(float);
    VSB_CUDA(cudaFuncSetAttribute(filter_merge_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));  // per device
    filter_merge_kernel<true><<<(unsigned)ceil_div64(nq, 4), 128, smem, st>>>(reinterpret_cast<const uint2*>(cand), cand_cnt, cap, nullptr, nq, k, 0,
                                                                       1, k + 2, out_scores, out_ids, k, rf);
    VSB_CUDA(cudaGetLastError());
    return VS_OK;
}

int launch_merge_lists(const float* part_key, const int32_t* part_id, int n_lists, int64_t nq, int list_len, int nsel,
                       int k, int64_t id_base, int neg_in, int neg_out, float* out_key, int32_t* out_id, int out_stride,
                       int out_off, float* lb_key_out, int32_t* lb_id_out, const float* rf_base, const float* rf_bnorm,
                       const float* rf_q, const float* rf_qnorm, cudaStream_t st, const TcQueryParams* cert_qp,
                       int32_t* uncert_count, int32_t* uncert_list, size_t list_stride, const int32_t* trailer,
                       int32_t* trailer_total_out, const int32_t* ivf_idmap) {
    if (nq <= 0) return VS_OK;
    if (nsel > list_len || k > nsel || nsel > 32) return fail(VS_ERR_INVALID, "merge: need k <= nsel <= list length <= 32");
    if (cert_qp && (!rf_base || !uncert_count || !uncert_list)) return fail(VS_ERR_INVALID, "merge: certification needs the refine");
    const unsigned blocks = (unsigned)ceil_div64(nq, 4);
    RefineArgs rf{rf_base, rf_bnorm, rf_q, rf_qnorm, cert_qp, uncert_count, uncert_list, ivf_idmap};
    BlockArgs ba{list_stride, trailer, trailer_total_out};
    if (ivf_idmap && (!rf_base || !rf_q || n_lists > 96)) return fail(VS_ERR_INVALID, "merge: IVF re-score needs vectors, queries and <= 96 lists");
    if (n_lists <= 96 && round_up_ktop(list_len) != 0) {
        // per warp: the staged lists, reused for the candidates' rows by the IVF re-score
        const int warp_floats = std::max(n_lists * 64, ivf_idmap ? ((nsel * kRowPitch + 3) & ~3) : 0);
        const size_t smem = (size_t)4 * warp_floats * sizeof(float);  // <= 96 KB
        if (smem > 48 * 1024)  // per device, hence not cached in a static
            VSB_CUDA(cudaFuncSetAttribute(merge_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 96 * 64 * (int)sizeof(float)));
        merge_small_kernel<<<blocks, 128, smem, st>>>(part_key, part_id, n_lists, nq, list_len, nsel, k, id_base, neg_in, neg_out,
                                                     out_key, out_id, out_stride, out_off, lb_key_out, lb_id_out, rf, ba, warp_floats);
        VSB_CUDA(cudaGetLastError());
        return VS_OK;
    }
#define VSB_MERGE_CASE(KT)                                                                                            \
    case KT:                                                                                                          \
        merge_lists_kernel<KT><<<blocks, 128, 0, st>>>(part_key, part_id, n_lists, nq, list_len, nsel, k, id_base,    \
                                                       neg_in, neg_out, out_key, out_id, out_stride, out_off,         \
                                                       lb_key_out, lb_id_out, rf, ba);                                \
        break;
    switch (round_up_ktop(list_len)) {
        VSB_MERGE_CASE(1)
        VSB_MERGE_CASE(5)
        VSB_MERGE_CASE(10)
        VSB_MERGE_CASE(16)
        VSB_MERGE_CASE(32)
        default:
            return fail(VS_ERR_INVALID, "merge: unsupported list size");
    }
#undef VSB_MERGE_CASE
    VSB_CUDA(cudaGetLastError());
    return VS_OK;
}

// Merge of G <= 32 per-shard result lists of any length k (the exchange step of the row-sharded search when k > 32):
// in[g][q][0..k) sorted in canonical order ((key asc | desc), id asc), -1 ids = padding at the tail.  One warp per
// query, lane g walks list g; every round the warp picks the best head with a shuffle arg-min and that lane advances.
__global__ void __launch_bounds__(128) merge_shards_kernel(const float* __restrict__ in_key, const int32_t* __restrict__ in_id,
                                                           int n_shards, int64_t nq, int k, int neg,
                                                           float* __restrict__ out_key, int32_t* __restrict__ out_id,
                                                           BlockArgs ba) {
    const int lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (ba.trailer && blockIdx.x == 0 && threadIdx.x == 0) {
        int32_t tot = 0;
        for (int l = 0; l < n_shards; ++l) tot += ba.trailer[(size_t)l * ba.list_stride];
        *ba.total_out = tot;
    }
    if (q >= nq) return;
    const float INF = __int_as_float(0x7f800000);
    const size_t lstride = ba.list_stride ? ba.list_stride : (size_t)nq * k;
    const float* pk = in_key + (size_t)min(lane, n_shards - 1) * lstride + (size_t)q * k;
    const int32_t* pi = in_id + (size_t)min(lane, n_shards - 1) * lstride + (size_t)q * k;
    int pos = lane < n_shards ? 0 : k;
    float hk = INF;
    int32_t hid = -1;
    if (pos < k) {
        hid = pi[0];
        hk = hid >= 0 ? (neg ? -pk[0] : pk[0]) : INF;
    }
    for (int r = 0; r < k; ++r) {
        float bk = hk;
        int32_t bi = hid;
        int src = lane;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ok = __shfl_xor_sync(0xffffffffu, bk, o);
            const int32_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
            const int os = __shfl_xor_sync(0xffffffffu, src, o);
            if (pair_less(ok, oi, bk, bi)) {
                bk = ok;
                bi = oi;
                src = os;
            }
        }
        if (lane == 0) {
            out_key[q * k + r] = bi >= 0 ? (neg ? -bk : bk) : (neg ? -INF : INF);
            out_id[q * k + r] = bi;
        }
        if (src == lane && bi >= 0) {
            ++pos;
            hid = pos < k ? pi[pos] : -1;
            hk = hid >= 0 ? (neg ? -pk[pos] : pk[pos]) : INF;
        }
    }
}

int launch_merge_shards(const float* in_key, const int32_t* in_id, int n_shards, int64_t nq, int k, int neg, float* out_key,
                        int32_t* out_id, cudaStream_t st, size_t list_stride, const int32_t* trailer, int32_t* trailer_total_out) {
    if (nq <= 0) return VS_OK;
    if (n_shards > 32) return fail(VS_ERR_UNSUPPORTED, "merge: more than 32 shards");
    BlockArgs ba{list_stride, trailer, trailer_total_out};
    merge_shards_kernel<<<(unsigned)ceil_div64(nq, 4), 128, 0, st>>>(in_key, in_id, n_shards, nq, k, neg, out_key, out_id, ba);
    VSB_CUDA(cudaGetLastError());
    return VS_OK;
}

// Final (key, id) sort of each row of out[nq][k] for multi-pass results (k > 32): passes are ordered by the
// candidate ranking, refined keys may reorder neighbours across a pass boundary.  One warp per row.
__global__ void __launch_bounds__(128) sort_rows_kernel(float* __restrict__ key, int32_t* __restrict__ id, int64_t nq, int k) {
    extern __shared__ uint8_t sm[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int64_t q = (int64_t)blockIdx.x * 4 + wib;
    if (q >= nq) return;
    float* sk = reinterpret_cast<float*>(sm) + (size_t)wib * k;
    int32_t* si = reinterpret_cast<int32_t*>(sm + (size_t)4 * k * sizeof(float)) + (size_t)wib * k;
    for (int c = lane; c < k; c += 32) {
        sk[c] = key[q * k + c];
        si[c] = id[q * k + c];
    }
    __syncwarp();
    for (int c = lane; c < k; c += 32) {
        const float mk = sk[c];
        const int32_t mi = si[c];
        if (mi < 0) continue;  // padding already sits at the end
        int rank = 0;
        for (int j = 0; j < k; ++j) rank += (si[j] >= 0 && pair_less(sk[j], si[j], mk, mi)) ? 1 : 0;
        key[q * k + rank] = mk;
        id[q * k + rank] = mi;
    }
}

int launch_sort_rows(float* key, int32_t* id, int64_t nq, int k, cudaStream_t st) {
    if (nq <= 0) return VS_OK;
    const size_t smem = (size_t)4 * k * 8;
    if (smem > 48 * 1024) return fail(VS_ERR_UNSUPPORTED, "k too large for the final sort");
    sort_rows_kernel<<<(unsigned)ceil_div64(nq, 4), 128, smem, st>>>(key, id, nq, k);
    VSB_CUDA(cudaGetLastError());
    return VS_OK;
}

// ------------------------------------------------------------------------------------------------
// synthetic data (must stay bit-identical to hai-25-rag-on-edge_b200/synth.py)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ uint64_t hash_idx(uint64_t seed, uint64_t idx) { return mix64(idx + seed * 0x9E3779B97F4A7C15ull); }
__device__ __forceinline__ int sift_from_hash(uint64_t h) { return (int)(((h & 0xFF) * ((h >> 8) & 0xFF)) >> 9); }

__global__ void synth_kernel(float* __restrict__ out, int64_t row0, int64_t nrows, int dim, int law, uint64_t seed,
                             uint64_t centre_seed) {
    const int64_t total = nrows * dim;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t r = (uint64_t)(row0 + e / dim);
        const uint64_t c = (uint64_t)(e % dim);
        const uint64_t h = hash_idx(seed, r * (uint64_t)dim + c);
        float v;
        if (law == 0) {
            v = (float)sift_from_hash(h);
        } else if (law == 1) {
            const float frac = (float)((h >> 16) & 0xFFFF) / 65536.0f;
            v = (float)sift_from_hash(h) + (frac - 0.5f);
        } else {
            const uint64_t cid = hash_idx(seed ^ 0x5BD1E995ull, r) % 4096ull;
            const int centre = sift_from_hash(hash_idx(centre_seed, cid * (uint64_t)dim + c));
            const int s = (int)((h >> 16) & 0xFF) + (int)((h >> 24) & 0xFF) + (int)((h >> 32) & 0xFF) + (int)((h >> 40) & 0xFF);
            int val = centre + s / 12 - 42;
            val = val < 0 ? 0 : (val > 218 ? 218 : val);
            v = (float)val;
        }
        out[e] = v;
    }
}

int launch_synth(float* out, int64_t row0, int64_t nrows, int dim, int law, uint64_t seed, uint64_t centre_seed,
                 cudaStream_t st) {
    if (nrows <= 0) return VS_OK;
    if (law < 0 || law > 2) return fail(VS_ERR_INVALID, "synth: unknown law");
    synth_kernel<<<148 * 8, 256, 0, st>>>(out, row0, nrows, dim, law, seed, centre_seed);
    VSB_CUDA(cudaGetLastError());
    return VS_OK;
}

}  // namespace vsb
