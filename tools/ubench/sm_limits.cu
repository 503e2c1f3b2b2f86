// Micro-benchmarks of the two per-SM rates that bound the fused distance + top-k kernel's epilogue and operand feed
// on B200 (sm_100a), measured with clock64 inside persistent 1-CTA-per-SM kernels:
//   tmem   TMEM -> register read rate (tcgen05.ld 32x32b.x32 / .x64, 4..16 warps, 1 or 2 loads in flight)
//   tma    shared-memory fill rate of TMA tile loads out of L2 (16 KB boxes, 128-byte swizzle, 8-stage ring, no consumer
//          work): every CTA streams the SAME panel (the kernel's access pattern), DISTINCT panels, or CTA pairs that
//          fetch half a box each and multicast it
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o sm_limits tools/ubench/sm_limits.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                      \
    do {                                                                           \
        cudaError_t e_ = (x);                                                      \
        if (e_ != cudaSuccess) {                                                   \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                               \
        }                                                                          \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {  // arrive on CTA `cta`'s barrier at this offset
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(cta));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// ---------------------------------------------------------------------------------------------------- TMEM reads
template <int SHAPE, int INFLIGHT>
__global__ void __launch_bounds__(512, 1) tmem_read_kernel(int iters, unsigned long long* cycles, uint32_t* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int f = 0; f < INFLIGHT; ++f) {
            const uint32_t col = (uint32_t)(((i * INFLIGHT + f) * (SHAPE == 64 ? 64 : 32)) & 511 & ~(SHAPE == 64 ? 63 : 31));
            if (SHAPE == 32) {
                uint32_t r[32];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                      "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                      "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                      "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(base + col) : "memory");
                if (f == INFLIGHT - 1) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                acc ^= r[0] ^ r[31];
            } else if (SHAPE == 64) {
                uint32_t r[64];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
                    "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
                    "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                      "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                      "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                      "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]),
                      "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]),
                      "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]),
                      "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]),
                      "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
                    : "r"(base + col) : "memory");
                if (f == INFLIGHT - 1) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                acc ^= r[0] ^ r[63];
            } else {  // 16x256b.x8: 16 lanes x 256 bits x 8 = 4 KB per instruction, two instructions cover the quadrant's 32 lanes
                uint32_t r[32];
                const uint32_t a = slot + ((uint32_t)((warp & 3) * 32 + (f & 1) * 16) << 16) + col;
                asm volatile(
                    "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                      "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                      "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                      "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(a) : "memory");
                if (f == INFLIGHT - 1) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                acc ^= r[0] ^ r[31];
            }
        }
    }
    const long long t1 = clock64();
    if (acc == 0x12345678u) sink[0] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}

// ---------------------------------------------------------------------------------------------------- TMA fill rate
constexpr int STAGES = 8, BOX = 128 * 128;  // 16 KB: 128 rows x 128 B
// mode 0: every CTA streams the same panel; 1: CTA b streams its own panel; 2: CTA pairs, each CTA fetches half of every
// box (64 rows) and multicasts it to both
template <int MODE>
__global__ void __launch_bounds__(64, 1) tma_fill_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tm_half,
                                                         int boxes, int rows_total, unsigned long long* cycles) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full = (uint64_t*)(smem + STAGES * BOX);
    uint64_t* empty = full + STAGES;
    uint32_t rank = 0;
    if (MODE == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], MODE == 2 ? 2 : 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (MODE == 2) {
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    const int tiles = rows_total / 128;
    const int first = MODE == 1 ? (int)((blockIdx.x * 977u) % (unsigned)tiles) : 0;
    const long long t0 = clock64();
    if (threadIdx.x == 0) {  // producer
        int stage = 0;
        uint32_t phase = 0;
        for (int i = 0; i < boxes; ++i) {
            mbar_wait(&empty[stage], phase ^ 1);
            const int tile = (first + (i >> 1)) % tiles;  // two k-blocks (of 128 B) per 128-row tile, like the fp16 sweep
            const int c0 = (i & 1) * 64;
            mbar_expect_tx(&full[stage], BOX);
            if (MODE == 2)
                asm volatile(
                    "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
                    ::"r"(smem_u32(smem + stage * BOX + rank * (BOX / 2))), "l"(&tm_half), "r"(smem_u32(&full[stage])), "r"(c0),
                      "r"(tile * 128 + (int)rank * 64), "h"((uint16_t)3) : "memory");
            else
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                             ::"r"(smem_u32(smem + stage * BOX)), "l"(&tm), "r"(smem_u32(&full[stage])), "r"(c0), "r"(tile * 128) : "memory");
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (threadIdx.x == 32) {  // consumer: frees the stage as soon as it has landed
        int stage = 0;
        uint32_t phase = 0;
        for (int i = 0; i < boxes; ++i) {
            mbar_wait(&full[stage], phase);
            if (MODE == 2) {
                mbar_arrive_cluster(&empty[stage], 0);
                mbar_arrive_cluster(&empty[stage], 1);
            } else {
                mbar_arrive(&empty[stage]);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (MODE == 2) {
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
}

typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static double mean_cycles(unsigned long long* d, int n) {
    std::vector<unsigned long long> h(n);
    CK(cudaMemcpy(h.data(), d, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    double s = 0;
    for (auto v : h) s += (double)v;
    return s / n;
}

template <int SHAPE, int INFLIGHT>
static void run_tmem(int warps, unsigned long long* d_cyc, uint32_t* d_sink, int sms) {
    const int iters = 20000 / INFLIGHT;
    tmem_read_kernel<SHAPE, INFLIGHT><<<sms, warps * 32>>>(100, d_cyc, d_sink);
    tmem_read_kernel<SHAPE, INFLIGHT><<<sms, warps * 32>>>(iters, d_cyc, d_sink);
    CK(cudaDeviceSynchronize());
    const double cyc = mean_cycles(d_cyc, sms);
    const double bytes = (double)warps * iters * INFLIGHT * (SHAPE == 64 ? 8192.0 : 4096.0);
    printf("tmem  shape=%-10s warps=%2d inflight=%d : %7.1f B/clk/SM  (%.0f cycles per 4 KB warp-load)\n",
           SHAPE == 32 ? "32x32b.x32" : SHAPE == 64 ? "32x32b.x64" : "16x256b.x8", warps, INFLIGHT, bytes / cyc,
           cyc / (iters * INFLIGHT) / (SHAPE == 64 ? 2 : 1));
}

int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    unsigned long long* d_cyc;
    uint32_t* d_sink;
    CK(cudaMalloc(&d_cyc, 1024 * sizeof(unsigned long long)));
    CK(cudaMalloc(&d_sink, 64));
    for (int warps : {4, 8, 12, 16}) {
        run_tmem<32, 1>(warps, d_cyc, d_sink, sms);
        run_tmem<32, 2>(warps, d_cyc, d_sink, sms);
    }
    run_tmem<64, 1>(4, d_cyc, d_sink, sms);
    run_tmem<64, 1>(8, d_cyc, d_sink, sms);

    // ---- TMA fill
    void* fnp = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
    encode_fn enc = (encode_fn)fnp;
    const int rows = 1 << 20;  // 1M rows x 128 fp16 = 256 MB (does not fit L2: first touch streams from HBM, sharers hit L2)
    void* d_base;
    CK(cudaMalloc(&d_base, (size_t)rows * 256));
    CK(cudaMemset(d_base, 1, (size_t)rows * 256));
    CUtensorMap tm, tmh;
    cuuint64_t gdim[2] = {128, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {256};
    cuuint32_t box[2] = {64, 128}, boxh[2] = {64, 64}, estr[2] = {1, 1};
    if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d_base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
        enc(&tmh, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d_base, gdim, gstr, boxh, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
        printf("tensor map encode failed\n");
        return 1;
    }
    const int smem = STAGES * BOX + 1024 + 256;
    CK(cudaFuncSetAttribute(tma_fill_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(tma_fill_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(tma_fill_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int boxes = 8192;  // 128 MB per CTA
    for (int rep = 0; rep < 2; ++rep) {
        for (int grid : {1, 2, 37, 74, 148}) {
            if (grid > sms) continue;
            tma_fill_kernel<0><<<grid, 64, smem>>>(tm, tmh, boxes, rows, d_cyc);
            CK(cudaDeviceSynchronize());
            const double c0 = mean_cycles(d_cyc, grid);
            tma_fill_kernel<1><<<grid, 64, smem>>>(tm, tmh, boxes, rows, d_cyc);
            CK(cudaDeviceSynchronize());
            const double c1 = mean_cycles(d_cyc, grid);
            double c2 = 0;
            if (grid % 2 == 0) {
                cudaLaunchConfig_t cfg{};
                cfg.gridDim = dim3(grid);
                cfg.blockDim = dim3(64);
                cfg.dynamicSmemBytes = smem;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = 2;
                at[0].val.clusterDim.y = at[0].val.clusterDim.z = 1;
                cfg.attrs = at;
                cfg.numAttrs = 1;
                CK(cudaLaunchKernelEx(&cfg, tma_fill_kernel<2>, tm, tmh, boxes, rows, d_cyc));
                CK(cudaDeviceSynchronize());
                c2 = mean_cycles(d_cyc, grid);
            }
            if (rep == 1)
                printf("tma   grid=%3d : same panel %6.1f B/clk/SM | distinct panels %6.1f B/clk/SM | pair multicast %6.1f B/clk/SM "
                       "(bytes landed per SM; chip-wide x grid)\n",
                       grid, (double)boxes * BOX / c0, (double)boxes * BOX / c1, c2 > 0 ? (double)boxes * BOX / c2 : 0.0);
        }
    }
    return 0;
}
