// Row-sharded exact search behind the C ABI (SURVEY.md §8b "n_gpus", §8e): the reference's run_benchmark()
// (cpu/cpu_baseline.cpp:177-181) has one base and one query loop; here the base rows are partitioned into shards,
// every shard runs the fused search on its rows, and ONE exchange step makes every participant see every shard's
// local top-k.
//
//   exchange block (per shard)   ids [nq x k] int32 | keys [nq x k] fp32 | 16-byte trailer (word 0 = number of queries
//                                the shard could not certify).  A shard writes its block straight into slot `s` of the
//                                gathered buffer [n_slots][block]: the all-gather is in place, ids and keys travel in one
//                                collective, and the "was anything redone?" question needs no collective of its own.
//   vs_exact_group               the shards that live on ONE device + the begin / merge / finish sequence around the
//                                exchange.  The exchange itself is the caller's: nothing (all shards local), a
//                                torch.distributed all-gather (one process per GPU: bench.py / sharded.py), or NCCL
//                                inside vs_exact_mgpu.
//   vs_exact_mgpu                single process, one worker thread + one stream per GPU, ncclCommInitAll, grouped
//                                ncclAllGather over NVLink (NCCL is dlopen'ed: libvsb200.so keeps loading without it).
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "exact_handle.cuh"
#include "kernels.cuh"
#include "vsb_common.cuh"

using namespace vsb;

namespace {

inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }
inline size_t trailer_off(int64_t nq, int k) { return align16((size_t)nq * k * 8); }
inline size_t block_bytes(int64_t nq, int k) { return trailer_off(nq, k) + 16; }

// ------------------------------------------------------------------------------------------------
// NCCL through dlopen (no link-time dependency)
// ------------------------------------------------------------------------------------------------
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string why;
};

NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            api.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (api.lib) break;
        }
        if (!api.lib) {
            const char* e = dlerror();
            api.why = std::string("libnccl.so.2 could not be loaded: ") + (e ? e : "?");
            return;
        }
        auto sym = [&](const char* n) {
            void* p = dlsym(api.lib, n);
            if (!p && api.why.empty()) api.why = std::string("NCCL symbol missing: ") + n;
            return p;
        };
        api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    });
    return &api;
}

#define VSB_NCCL(expr)                                                                                           \
    do {                                                                                                         \
        ncclResult_t _r = (expr);                                                                                \
        if (_r != ncclSuccess)                                                                                   \
            return ::vsb::fail(VS_ERR_CUDA, std::string(#expr) + ": " + nccl_api()->GetErrorString(_r));          \
    } while (0)

}  // namespace

// ------------------------------------------------------------------------------------------------
// vs_exact_group: the shards of one device
// ------------------------------------------------------------------------------------------------
struct vs_exact_group {
    int device = 0;
    int n_slots = 0, first_slot = 0;
    std::vector<vs_exact*> shards;
    bool owns_shards = false;
    DevBuf total;           // [1] sum of the slots' uncertified counts (written by the merge kernel)
    int* h_total = nullptr; // pinned copy
    cudaEvent_t ev_total = nullptr;
    // search in flight
    bool begun = false, merged = false;
    int64_t nq = 0;
    int k = 0;
    uint8_t* gathered = nullptr;
    cudaStream_t st = nullptr;
};

static int group_free(vs_exact_group* g) {
    if (!g) return VS_OK;
    cudaSetDevice(g->device);
    if (g->owns_shards)
        for (vs_exact* s : g->shards) exact_free(s);
    g->total.release();
    if (g->h_total) cudaFreeHost(g->h_total);
    if (g->ev_total) cudaEventDestroy(g->ev_total);
    delete g;
    return VS_OK;
}

static int group_from(vs_exact_group** out, const std::vector<vs_exact*>& shards, bool owns, int n_slots, int first_slot) {
    if (!out) return fail(VS_ERR_INVALID, "out handle is NULL");
    *out = nullptr;
    if (shards.empty()) return fail(VS_ERR_INVALID, "a group needs at least one shard");
    if (first_slot < 0 || first_slot + (int)shards.size() > n_slots) return fail(VS_ERR_INVALID, "slots out of range");
    if (n_slots > 32 * 1024) return fail(VS_ERR_INVALID, "too many shards");
    for (vs_exact* s : shards)
        if (!s || s->device != shards[0]->device || s->dim != shards[0]->dim)
            return fail(VS_ERR_INVALID, "the shards of a group must live on one device and share dim");
    vs_exact_group* g = new (std::nothrow) vs_exact_group();
    if (!g) return fail(VS_ERR_NOMEM, "host allocation failed");
    g->device = shards[0]->device;
    g->shards = shards;
    g->owns_shards = false;  // until everything below succeeded: a failed create leaves the caller's shards alone
    g->n_slots = n_slots;
    g->first_slot = first_slot;
    int rc = VS_OK;
    do {
        if (cudaSetDevice(g->device) != cudaSuccess) { rc = fail(VS_ERR_CUDA, "cudaSetDevice failed"); break; }
        if ((rc = g->total.reserve(sizeof(int))) != VS_OK) break;
        if (cudaMallocHost((void**)&g->h_total, sizeof(int)) != cudaSuccess) { rc = fail(VS_ERR_CUDA, "cudaMallocHost failed"); break; }
        if (cudaEventCreateWithFlags(&g->ev_total, cudaEventDisableTiming) != cudaSuccess) { rc = fail(VS_ERR_CUDA, "cudaEventCreate failed"); break; }
    } while (0);
    if (rc != VS_OK) {
        group_free(g);
        return rc;
    }
    g->owns_shards = owns;
    *out = g;
    return VS_OK;
}

// local searches -> the shards' blocks inside `gathered` (no host synchronisation)
static int group_begin(vs_exact_group* g, const float* q_dev, int64_t nq, int k, int precision, void* gathered, cudaStream_t st) {
    if (g->begun) return fail(VS_ERR_INVALID, "vs_exact_group_finish() of the previous search was not called");
    if (nq <= 0 || k <= 0) return fail(VS_ERR_INVALID, "nq <= 0 or k <= 0");
    if (!q_dev || !gathered) return fail(VS_ERR_INVALID, "NULL buffer");
    VSB_CUDA(cudaSetDevice(g->device));
    const size_t B = block_bytes(nq, k), toff = trailer_off(nq, k);
    for (size_t i = 0; i < g->shards.size(); ++i) {
        vs_exact* s = g->shards[i];
        if ((int64_t)k > s->n) return fail(VS_ERR_INVALID, "k exceeds the rows of a shard");
        uint8_t* blk = (uint8_t*)gathered + (size_t)(g->first_slot + (int)i) * B;
        VSB_CUDA(cudaMemsetAsync(blk + toff, 0, 16, st));
        VSB_TRY(exact_search_core(s, q_dev, nq, k, precision, (int32_t*)blk, (float*)(blk + (size_t)nq * k * 4), st, true,
                                  (int32_t*)(blk + toff)));
    }
    g->begun = true;
    g->merged = false;
    g->nq = nq;
    g->k = k;
    g->gathered = (uint8_t*)gathered;
    g->st = st;
    return VS_OK;
}

static int merge_blocks(const void* blocks, int n_shards, size_t stride, int64_t nq, int k, int smallest, int32_t* out_ids,
                        float* out_keys, int32_t* total_out, cudaStream_t st) {
    if (stride % 4 != 0 || stride < block_bytes(nq, k)) return fail(VS_ERR_INVALID, "bad block stride");
    const int32_t* ids = (const int32_t*)blocks;
    const float* keys = (const float*)((const uint8_t*)blocks + (size_t)nq * k * 4);
    const int32_t* trailer = total_out ? (const int32_t*)((const uint8_t*)blocks + trailer_off(nq, k)) : nullptr;
    const int neg = smallest ? 0 : 1;
    if (k > kMaxRegK) return launch_merge_shards(keys, ids, n_shards, nq, k, neg, out_keys, out_ids, st, stride / 4, trailer, total_out);
    return launch_merge_lists(keys, ids, n_shards, nq, k, k, k, 0, neg, neg, out_keys, out_ids, k, 0, nullptr, nullptr, nullptr,
                              nullptr, nullptr, nullptr, st, nullptr, nullptr, nullptr, stride / 4, trailer, total_out);
}

// merge of all slots (after the exchange) + the total of the uncertified counts on its way to the host
static int group_merge(vs_exact_group* g, int32_t* out_ids, float* out_dists) {
    if (!g->begun) return fail(VS_ERR_INVALID, "vs_exact_group_begin() was not called");
    if (!out_ids || !out_dists) return fail(VS_ERR_INVALID, "NULL buffer");
    VSB_CUDA(cudaSetDevice(g->device));
    VSB_TRY(merge_blocks(g->gathered, g->n_slots, block_bytes(g->nq, g->k), g->nq, g->k, 1, out_ids, out_dists,
                         g->total.as<int32_t>(), g->st));
    VSB_CUDA(cudaMemcpyAsync(g->h_total, g->total.p, sizeof(int), cudaMemcpyDeviceToHost, g->st));
    VSB_CUDA(cudaEventRecord(g->ev_total, g->st));
    g->merged = true;
    return VS_OK;
}

// redo of the local shards' uncertified queries (rewrites rows of their blocks, zeroes the trailers)
static int group_redo_local(vs_exact_group* g, int* n_redone) {
    VSB_CUDA(cudaSetDevice(g->device));
    int tot = 0;
    const size_t B = block_bytes(g->nq, g->k), toff = trailer_off(g->nq, g->k);
    for (size_t i = 0; i < g->shards.size(); ++i) {
        int n = 0;
        VSB_TRY(exact_certified_finish(g->shards[i], &n));
        tot += n;
        if (g->begun && n > 0)
            VSB_CUDA(cudaMemsetAsync(g->gathered + (size_t)(g->first_slot + (int)i) * B + toff, 0, 16, g->st));
    }
    if (n_redone) *n_redone = tot;
    return VS_OK;
}

// waits for the total only; > 0: some shard (here or elsewhere) has queries to redo -> exchange + merge again
static int group_finish(vs_exact_group* g, int* need_reexchange) {
    if (need_reexchange) *need_reexchange = 0;
    if (!g->begun) return VS_OK;
    if (!g->merged) return fail(VS_ERR_INVALID, "vs_exact_group_merge() was not called");
    VSB_CUDA(cudaSetDevice(g->device));
    VSB_CUDA(cudaEventSynchronize(g->ev_total));
    const int tot = *g->h_total;
    VSB_TRY(group_redo_local(g, nullptr));
    if (tot > 0) {
        g->merged = false;  // stays begun: the caller exchanges and merges again, then calls finish again
        if (need_reexchange) *need_reexchange = 1;
        return VS_OK;
    }
    g->begun = false;
    return VS_OK;
}

// ------------------------------------------------------------------------------------------------
// vs_exact_mgpu: one process, several GPUs
// ------------------------------------------------------------------------------------------------
namespace {

// one persistent worker thread per device; run(fn) executes fn(d) on every worker and waits
class DevicePool {
  public:
    explicit DevicePool(const std::vector<int>& devices) : devices_(devices), rc_(devices.size(), VS_OK), err_(devices.size()) {
        for (size_t i = 0; i < devices.size(); ++i) threads_.emplace_back([this, i] { loop(i); });
    }
    ~DevicePool() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
            ++gen_;
        }
        cv_.notify_all();
        for (auto& t : threads_) t.join();
    }
    int run(const std::function<int(int)>& fn) {
        if (devices_.size() == 1) {  // no hand-off latency on a single GPU
            cudaSetDevice(devices_[0]);
            return fn(0);
        }
        {
            std::lock_guard<std::mutex> lk(m_);
            fn_ = &fn;
            left_ = (int)devices_.size();
            ++gen_;
        }
        cv_.notify_all();
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [this] { return left_ == 0; });
        for (size_t i = 0; i < rc_.size(); ++i)
            if (rc_[i] != VS_OK) return fail(rc_[i], err_[i]);
        return VS_OK;
    }

  private:
    void loop(size_t i) {
        cudaSetDevice(devices_[i]);
        uint64_t seen = 0;
        for (;;) {
            const std::function<int(int)>* fn;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                fn = fn_;
            }
            const int rc = (*fn)((int)i);
            const std::string e = rc != VS_OK ? vs_last_error() : "";
            {
                std::lock_guard<std::mutex> lk(m_);
                rc_[i] = rc;
                err_[i] = e;
                if (--left_ == 0) done_.notify_all();
            }
        }
    }
    std::vector<int> devices_;
    std::vector<std::thread> threads_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    const std::function<int(int)>* fn_ = nullptr;
    uint64_t gen_ = 0;
    int left_ = 0;
    bool stop_ = false;
    std::vector<int> rc_;
    std::vector<std::string> err_;
};

struct DevCtx {
    int device = 0;
    vs_exact_group* grp = nullptr;
    cudaStream_t st = nullptr;
    DevBuf q, gathered, out_ids, out_dists;
    ncclComm_t comm = nullptr;
};

}  // namespace

struct vs_exact_mgpu {
    int n_gpus = 0, spg = 1, n_slots = 0, dim = 0;
    int64_t n = 0;
    std::vector<DevCtx> dev;
    DevicePool* pool = nullptr;
    bool profile = false;
    int last_redone = 0;
    int last_exchanges = 0;
};

static int mgpu_free(vs_exact_mgpu* m) {
    if (!m) return VS_OK;
    for (DevCtx& c : m->dev) {
        cudaSetDevice(c.device);
        if (c.st) cudaStreamSynchronize(c.st);
        if (c.comm) nccl_api()->CommDestroy(c.comm);
        group_free(c.grp);
        c.q.release();
        c.gathered.release();
        c.out_ids.release();
        c.out_dists.release();
        if (c.st) cudaStreamDestroy(c.st);
    }
    delete m->pool;
    delete m;
    return VS_OK;
}

// all-gather in place over the devices' buffers: device d contributes bytes [d*count, (d+1)*count) of its buffer
static int mgpu_allgather(vs_exact_mgpu* m, const std::function<uint8_t*(DevCtx&)>& buf, size_t count) {
    if (m->n_gpus == 1) return VS_OK;
    NcclApi* nc = nccl_api();
    VSB_NCCL(nc->GroupStart());
    for (int d = 0; d < m->n_gpus; ++d) {
        DevCtx& c = m->dev[d];
        uint8_t* b = buf(c);
        ncclResult_t r = nc->AllGather(b + (size_t)d * count, b, count, ncclChar, c.comm, c.st);
        if (r != ncclSuccess) {
            nc->GroupEnd();
            return fail(VS_ERR_CUDA, std::string("ncclAllGather: ") + nc->GetErrorString(r));
        }
    }
    VSB_NCCL(nc->GroupEnd());
    return VS_OK;
}

namespace {
struct PushDst {
    void* p[16];
};
// every 16-byte unit of the source goes to the same offset of every destination (peer GPUs over NVLink); the block that
// finishes last (all stores of all blocks fenced at system scope) raises this sender's flag word on every peer
__global__ void __launch_bounds__(256) push_block_kernel(const uint4* __restrict__ src, PushDst dst, int n_dst, size_t n16,
                                                         PushDst flag, uint32_t epoch, uint32_t* __restrict__ counter) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = src[i];
        for (int d = 0; d < n_dst; ++d) reinterpret_cast<uint4*>(dst.p[d])[i] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(counter, 1u);
        if (prev == gridDim.x - 1) {
            atomicExch(counter, 0u);
            __threadfence_system();
            for (int d = 0; d < n_dst; ++d)
                asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag.p[d]), "r"(epoch) : "memory");
        }
    }
}
__global__ void wait_flags_kernel(const uint32_t* flags, int n, int self, uint32_t epoch) {
    const int t = threadIdx.x;
    if (t < n && t != self) {
        uint32_t v;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + t) : "memory");
            if ((int32_t)(v - epoch) >= 0) break;
            __nanosleep(40);
        } while (true);
    }
}
}  // namespace

extern "C" {

int vs_push_block_dev(const void* src_dev, void* const* dst_dev, int n_dst, size_t bytes, void* const* flag_dst, uint32_t epoch,
                      uint32_t* counter_dev, void* stream) {
    if (n_dst < 0 || n_dst > 16) return fail(VS_ERR_INVALID, "push: at most 16 destinations");
    if (n_dst == 0) return VS_OK;
    if (!flag_dst || !counter_dev || (bytes > 0 && (!src_dev || !dst_dev))) return fail(VS_ERR_INVALID, "NULL buffer");
    if (bytes % 16 != 0 || ((uintptr_t)src_dev & 15)) return fail(VS_ERR_INVALID, "push: 16-byte granularity");
    PushDst d{}, f{};
    for (int i = 0; i < n_dst; ++i) {
        if (bytes > 0 && (!dst_dev[i] || ((uintptr_t)dst_dev[i] & 15))) return fail(VS_ERR_INVALID, "push: bad destination pointer");
        if (!flag_dst[i] || ((uintptr_t)flag_dst[i] & 3)) return fail(VS_ERR_INVALID, "push: bad flag pointer");
        d.p[i] = bytes > 0 ? dst_dev[i] : nullptr;
        f.p[i] = flag_dst[i];
    }
    const size_t n16 = bytes / 16;
    const unsigned blocks = (unsigned)std::max<size_t>(1, std::min<size_t>((n16 + 255) / 256, 148 * 4));
    push_block_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(src_dev), d, n_dst, n16, f, epoch,
                                                               counter_dev);
    VSB_CUDA(cudaGetLastError());
    return VS_OK;
}

int vs_wait_flags_dev(const uint32_t* flags_dev, int n, int self, uint32_t epoch, void* stream) {
    if (!flags_dev || n < 1 || n > 32) return fail(VS_ERR_INVALID, "wait: 1 .. 32 flag words");
    wait_flags_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(flags_dev, n, self, epoch);
    VSB_CUDA(cudaGetLastError());
    return VS_OK;
}

size_t vs_topk_block_bytes(int64_t nq, int k) { return (nq > 0 && k > 0) ? block_bytes(nq, k) : 0; }

int vs_merge_blocks_dev(const void* blocks_dev, int n_shards, size_t block_stride, int64_t nq, int k, int smallest,
                        int32_t* out_ids_dev, float* out_keys_dev, int32_t* total_dev, void* stream) {
    if (!blocks_dev || !out_ids_dev || !out_keys_dev) return fail(VS_ERR_INVALID, "NULL buffer");
    if (n_shards <= 0 || nq < 0 || k <= 0) return fail(VS_ERR_INVALID, "bad sizes");
    if (nq == 0) return VS_OK;
    return merge_blocks(blocks_dev, n_shards, block_stride, nq, k, smallest, out_ids_dev, out_keys_dev, total_dev, (cudaStream_t)stream);
}

int vs_exact_group_create_from(vs_exact_group_t** out, int n_local, vs_exact_t* const* shards, int n_slots, int first_slot) {
    if (!shards || n_local <= 0) return fail(VS_ERR_INVALID, "no shards");
    return group_from(out, std::vector<vs_exact*>(shards, shards + n_local), false, n_slots, first_slot);
}
int vs_exact_group_destroy(vs_exact_group_t* g) { return group_free(g); }

int vs_exact_group_begin(vs_exact_group_t* g, const float* queries_dev, int64_t nq, int k, int precision, void* gathered_dev,
                         void* stream) {
    if (!g) return fail(VS_ERR_INVALID, "handle is NULL");
    return group_begin(g, queries_dev, nq, k, precision, gathered_dev, stream ? (cudaStream_t)stream : g->shards[0]->stream);
}
int vs_exact_group_merge(vs_exact_group_t* g, int32_t* out_ids_dev, float* out_dists_dev) {
    if (!g) return fail(VS_ERR_INVALID, "handle is NULL");
    return group_merge(g, out_ids_dev, out_dists_dev);
}
int vs_exact_group_finish(vs_exact_group_t* g, int* need_reexchange) {
    if (!g) return fail(VS_ERR_INVALID, "handle is NULL");
    return group_finish(g, need_reexchange);
}

// ---- single process, n_gpus devices ------------------------------------------------------------
int vs_exact_mgpu_create(vs_exact_mgpu_t** out, const float* base, int64_t n, int dim, int n_gpus, int shards_per_gpu) {
    if (!out) return fail(VS_ERR_INVALID, "out handle is NULL");
    *out = nullptr;
    if (!base || n <= 0) return fail(VS_ERR_INVALID, "base is NULL or n <= 0");
    DeviceGuard guard;  // the caller's current device is left as it was
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) {
        cudaGetLastError();
        return fail(VS_ERR_CUDA, "no CUDA device available (libvsb200 has no CPU fallback)");
    }
    if (n_gpus == 0) n_gpus = cnt;
    if (n_gpus < 0 || n_gpus > cnt) return fail(VS_ERR_INVALID, "n_gpus exceeds the visible devices");
    if (shards_per_gpu <= 0) shards_per_gpu = 1;
    const int n_slots = n_gpus * shards_per_gpu;
    if ((int64_t)n_slots > n) return fail(VS_ERR_INVALID, "more shards than rows");
    if (n_gpus > 1 && !nccl_api()->lib) return fail(VS_ERR_CUDA, nccl_api()->why);
    if (n_gpus > 1 && !nccl_api()->why.empty()) return fail(VS_ERR_CUDA, nccl_api()->why);
    vs_exact_mgpu* m = new (std::nothrow) vs_exact_mgpu();
    if (!m) return fail(VS_ERR_NOMEM, "host allocation failed");
    m->n_gpus = n_gpus;
    m->spg = shards_per_gpu;
    m->n_slots = n_slots;
    m->dim = dim;
    m->n = n;
    m->dev.resize(n_gpus);
    std::vector<int> devs(n_gpus);
    for (int d = 0; d < n_gpus; ++d) devs[d] = m->dev[d].device = d;
    m->pool = new DevicePool(devs);
    // every device uploads and indexes its own shards (in parallel: one worker per device)
    int rc = m->pool->run([&](int d) -> int {
        DevCtx& c = m->dev[d];
        VSB_CUDA(cudaSetDevice(c.device));
        VSB_CUDA(cudaStreamCreateWithFlags(&c.st, cudaStreamNonBlocking));
        std::vector<vs_exact*> shards;
        for (int i = 0; i < shards_per_gpu; ++i) {
            const int slot = d * shards_per_gpu + i;
            const int64_t r0 = n * slot / n_slots, r1 = n * (slot + 1) / n_slots;
            vs_exact_t* s = nullptr;
            const int r = exact_create_common(&s, base + (size_t)r0 * dim, false, r1 - r0, dim, c.device, r0);
            if (r != VS_OK) {
                for (vs_exact* x : shards) exact_free(x);
                return r;
            }
            shards.push_back(s);
        }
        const int r = group_from(&c.grp, shards, true, n_slots, d * shards_per_gpu);
        if (r != VS_OK)
            for (vs_exact* x : shards) exact_free(x);
        return r;
    });
    if (rc == VS_OK && n_gpus > 1) {
        std::vector<ncclComm_t> comms(n_gpus);
        ncclResult_t r = nccl_api()->CommInitAll(comms.data(), n_gpus, devs.data());
        if (r != ncclSuccess) rc = fail(VS_ERR_CUDA, std::string("ncclCommInitAll: ") + nccl_api()->GetErrorString(r));
        else
            for (int d = 0; d < n_gpus; ++d) m->dev[d].comm = comms[d];
    }
    if (rc == VS_OK && n_gpus > 1) {
        // NCCL sets its channels and peer connections up lazily, inside the first collective (hundreds of milliseconds):
        // do that here, as part of the untimed index build, with a tiny all-gather over the buffers the searches will use
        rc = m->pool->run([&](int d) -> int { return m->dev[d].gathered.reserve((size_t)n_gpus * 256); });
        if (rc == VS_OK) rc = mgpu_allgather(m, [](DevCtx& c) { return c.gathered.as<uint8_t>(); }, 256);
        if (rc == VS_OK)
            rc = m->pool->run([&](int d) -> int {
                VSB_CUDA(cudaStreamSynchronize(m->dev[d].st));
                return VS_OK;
            });
    }
    if (rc != VS_OK) {
        const std::string keep = vs_last_error();
        mgpu_free(m);
        set_error(keep);
        return rc;
    }
    *out = m;
    return VS_OK;
}

int vs_exact_mgpu_destroy(vs_exact_mgpu_t* m) {
    DeviceGuard guard;
    return mgpu_free(m);
}
int vs_exact_mgpu_num_gpus(const vs_exact_mgpu_t* m) { return m ? m->n_gpus : 0; }
int vs_exact_mgpu_num_shards(const vs_exact_mgpu_t* m) { return m ? m->n_slots : 0; }

int vs_exact_mgpu_search_f32(vs_exact_mgpu_t* m, const float* queries, int64_t nq, int k, int precision, int32_t* out_ids,
                             float* out_dists) {
    if (!m) return fail(VS_ERR_INVALID, "handle is NULL");
    if (nq < 0 || k <= 0) return fail(VS_ERR_INVALID, "nq < 0 or k <= 0");
    if (nq == 0) return VS_OK;
    if (!queries || !out_ids || !out_dists) return fail(VS_ERR_INVALID, "NULL buffer");
    DeviceGuard guard;
    const int G = m->n_gpus;
    const int dim = m->dim;
    const size_t B = block_bytes(nq, k);
    const int64_t q_slice = (nq + G - 1) / G;  // rows of the query slice a device uploads
    m->last_redone = 0;
    m->last_exchanges = 0;
    // 1. every device uploads ITS slice of the queries over its own PCIe link (the batch crosses PCIe once in total)
    VSB_TRY(m->pool->run([&](int d) -> int {
        DevCtx& c = m->dev[d];
        VSB_TRY(c.q.reserve(sizeof(float) * (size_t)q_slice * G * dim));
        VSB_TRY(c.gathered.reserve(B * (size_t)m->n_slots));
        if (d == 0) {
            VSB_TRY(c.out_ids.reserve(sizeof(int32_t) * (size_t)nq * k));
            VSB_TRY(c.out_dists.reserve(sizeof(float) * (size_t)nq * k));
        }
        const int64_t r0 = std::min<int64_t>(nq, q_slice * d), r1 = std::min<int64_t>(nq, q_slice * (d + 1));
        if (r1 > r0)
            VSB_CUDA(cudaMemcpyAsync(c.q.as<float>() + (size_t)r0 * dim, queries + (size_t)r0 * dim,
                                     sizeof(float) * (size_t)(r1 - r0) * dim, cudaMemcpyHostToDevice, c.st));
        return VS_OK;
    }));
    // 2. ... and the slices are replicated over NVLink
    VSB_TRY(mgpu_allgather(m, [](DevCtx& c) { return c.q.as<uint8_t>(); }, sizeof(float) * (size_t)q_slice * dim));
    // 3. local fused searches, each shard's block written in place into the device's gathered buffer
    VSB_TRY(m->pool->run([&](int d) -> int {
        DevCtx& c = m->dev[d];
        return group_begin(c.grp, c.q.as<float>(), nq, k, precision, c.gathered.p, c.st);
    }));
    DevCtx& c0 = m->dev[0];
    for (int round = 0; round < 3; ++round) {
        // 4. the exchange: one all-gather of the blocks (ids, keys and uncertified counts together)
        VSB_TRY(mgpu_allgather(m, [](DevCtx& c) { return c.gathered.as<uint8_t>(); }, B * (size_t)m->spg));
        ++m->last_exchanges;
        // 5. device 0 merges and returns
        VSB_CUDA(cudaSetDevice(c0.device));
        VSB_TRY(group_merge(c0.grp, c0.out_ids.as<int32_t>(), c0.out_dists.as<float>()));
        VSB_CUDA(cudaMemcpyAsync(out_ids, c0.out_ids.p, sizeof(int32_t) * (size_t)nq * k, cudaMemcpyDeviceToHost, c0.st));
        VSB_CUDA(cudaMemcpyAsync(out_dists, c0.out_dists.p, sizeof(float) * (size_t)nq * k, cudaMemcpyDeviceToHost, c0.st));
        int need = 0;
        VSB_TRY(group_finish(c0.grp, &need));  // waits for the 4-byte total only
        if (!need) {
            // the other devices' shards had nothing to redo either (their counts are part of the total)
            VSB_TRY(m->pool->run([&](int d) -> int {
                if (d == 0) return VS_OK;
                DevCtx& c = m->dev[d];
                VSB_TRY(group_redo_local(c.grp, nullptr));
                c.grp->begun = false;
                return VS_OK;
            }));
            break;
        }
        if (round == 2) return fail(VS_ERR_CUDA, "certification did not settle");
        // 6. rare: some shard could not certify some queries -> every device redoes ITS uncertified queries on the fp32
        //    path (device 0 already did inside finish), then the blocks are exchanged and merged again
        std::vector<int> redone(G, 0);
        VSB_TRY(m->pool->run([&](int d) -> int {
            if (d == 0) return VS_OK;
            return group_redo_local(m->dev[d].grp, &redone[d]);
        }));
        m->last_redone = 1;
    }
    VSB_CUDA(cudaSetDevice(c0.device));
    VSB_CUDA(cudaStreamSynchronize(c0.st));
    return VS_OK;
}

int vs_exact_mgpu_last_stats(const vs_exact_mgpu_t* m, int* n_exchanges, int* redone) {
    if (!m) return fail(VS_ERR_INVALID, "handle is NULL");
    if (n_exchanges) *n_exchanges = m->last_exchanges;
    if (redone) *redone = m->last_redone;
    return VS_OK;
}

int vs_exact_mgpu_set_profile(vs_exact_mgpu_t* m, int enable) {
    if (!m) return fail(VS_ERR_INVALID, "handle is NULL");
    for (DevCtx& c : m->dev)
        for (vs_exact* s : c.grp->shards) VSB_TRY(vs_exact_set_profile(s, enable));
    m->profile = enable != 0;
    return VS_OK;
}

// slowest shard's dominant-kernel time of the last search
int vs_exact_mgpu_last_kernel_ms(vs_exact_mgpu_t* m, float* ms) {
    if (!m || !ms) return fail(VS_ERR_INVALID, "NULL argument");
    float mx = 0.f;
    for (DevCtx& c : m->dev)
        for (vs_exact* s : c.grp->shards) {
            float t = 0.f;
            VSB_TRY(vs_exact_last_kernel_ms(s, &t));
            mx = std::max(mx, t);
        }
    *ms = mx;
    return VS_OK;
}

}  // extern "C"
