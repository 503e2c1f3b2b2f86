// `cpu_baseline` — drop-in for the reference's exact-kNN benchmark program (cpu/cpu_baseline.cpp) with the hot
// triple compute_norms -> cblas_sgemm -> select_topk replaced by libvsb200 (include/vsb200.h).
//
//   cpu_baseline <base.fvecs> <query.fvecs> <k> <results.txt> [--batch B] [--precision auto|3xtf32|ffma|1xtf32|f16cert]
//                [--device D] [--gpus N] [--shards-per-gpu S]
//       the CLI documented in cpu/README.md:83-103 (the reference's main() ignores argv, cpu_baseline.cpp:323).
//       --gpus N (0 = all visible): base rows sharded over N GPUs inside this one process (vs_exact_mgpu_*: a worker
//       thread and a stream per GPU, one NCCL all-gather of the per-shard top-k, merge on GPU 0).  Exit status 1 when
//       the run fails.
//   cpu_baseline
//       no arguments: the reference's hard-coded runs — siftsmall/ and sift/ in the working directory, k = 5,
//       siftsmall_results.txt / sift_results.txt (cpu_baseline.cpp:329-345); a missing dataset is reported and skipped.
//
// Output: results file `Query i: (idx, dist) ...` with default ostream float formatting (cpu_baseline.cpp:167-172),
// neighbours ascending by squared L2 distance; the metrics block of cpu_baseline.cpp:270-312 on stdout, with the
// per-query spans replaced by per-batch spans (one fused kernel does distance and top-k, so there is no separate
// "distance" and "top-k" time; the timed span is query upload + search + download, like the reference's query loop).
#include <chrono>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "../../include/vsb200.h"
#include "vsb_io.hpp"

namespace {

struct Options {
    int64_t batch = 0;  // 0 = all queries in one call
    int precision = VS_PREC_AUTO;
    int device = 0;
    int gpus = -1;            // -1: single-GPU handle on `device`; >= 0: vs_exact_mgpu over that many GPUs (0 = all)
    int shards_per_gpu = 1;
};

bool write_results(const std::string& path, const std::vector<int32_t>& ids, const std::vector<float>& dists, size_t nq,
                   int k) {
    std::ofstream out(path);
    if (!out) {
        std::cerr << "Error: Cannot open output file " << path << std::endl;
        return false;
    }
    std::vector<char> buffer(1 << 20);
    out.rdbuf()->pubsetbuf(buffer.data(), (std::streamsize)buffer.size());
    for (size_t i = 0; i < nq; ++i) {
        out << "Query " << i << ":";
        for (int t = 0; t < k; ++t) out << " (" << ids[i * k + t] << ", " << dists[i * k + t] << ")";
        out << "\n";
    }
    return out.good();
}

void print_stats(const char* title, const vsbio::TimingStats& s, const char* unit) {
    std::cout << "\n" << title << ":" << std::endl;
    std::cout << "  Average latency: " << s.mean * 1000.0 << " ms/" << unit << std::endl;
    std::cout << "  Std deviation: " << s.std_dev * 1000.0 << " ms" << std::endl;
    std::cout << "  Min latency: " << s.min_val * 1000.0 << " ms" << std::endl;
    std::cout << "  Max latency: " << s.max_val * 1000.0 << " ms" << std::endl;
    std::cout << "  P50 latency: " << s.p50 * 1000.0 << " ms" << std::endl;
    std::cout << "  P95 latency: " << s.p95 * 1000.0 << " ms" << std::endl;
    std::cout << "  P99 latency: " << s.p99 * 1000.0 << " ms" << std::endl;
}

// same role and argument meaning as run_benchmark() (cpu_baseline.cpp:177-181); errors are reported and the run is
// abandoned, like the reference (the return value only feeds the exit status of the 4-argument CLI)
bool run_benchmark(const std::string& dataset_name, const std::string& base_file, const std::string& query_file, int k,
                   const std::string& output_file, const Options& opt) {
    using clock = std::chrono::high_resolution_clock;
    std::cout << "\n========================================" << std::endl;
    std::cout << "Processing dataset: " << dataset_name << std::endl;
    std::cout << "========================================" << std::endl;

    std::vector<float> B, Q;
    size_t B_rows = 0, Q_rows = 0;
    int B_dim = 0, Q_dim = 0;
    try {
        std::cout << "Reading base vectors from " << base_file << "..." << std::endl;
        vsbio::read_fvecs(base_file, B, B_rows, B_dim);
        std::cout << "Reading query vectors from " << query_file << "..." << std::endl;
        vsbio::read_fvecs(query_file, Q, Q_rows, Q_dim);
    } catch (const std::exception& e) {
        std::cerr << "Error: " << e.what() << std::endl;
        return false;
    }
    if (B_dim != Q_dim) {
        std::cerr << "Error: Dimension mismatch between base and query vectors." << std::endl;
        return false;
    }
    if (k <= 0 || (size_t)k > B_rows) {
        std::cerr << "Error: k must be in [1, number of base vectors]." << std::endl;
        return false;
    }
    std::cout << "Loaded " << B_rows << " base vectors and " << Q_rows << " queries (dim " << B_dim << ")" << std::endl;

    // index build = base upload + norm precompute: outside the timed region, like compute_norms (cpu_baseline.cpp:211-212)
    vs_exact_t* index = nullptr;
    vs_exact_mgpu_t* multi = nullptr;
    const auto build_start = clock::now();
    const int rc_create = opt.gpus >= 0
                              ? vs_exact_mgpu_create(&multi, B.data(), (int64_t)B_rows, B_dim, opt.gpus, opt.shards_per_gpu)
                              : vs_exact_create(&index, B.data(), (int64_t)B_rows, B_dim, opt.device, 0);
    if (rc_create != VS_OK) {
        std::cerr << "Error: " << vs_last_error() << std::endl;
        return false;
    }
    std::vector<int32_t> ids(Q_rows * (size_t)k);
    std::vector<float> dists(Q_rows * (size_t)k);
    const int64_t batch = opt.batch > 0 ? opt.batch : (int64_t)std::max<size_t>(Q_rows, 1);
    // still index build: one untimed search of the first batch, so that the device workspaces, the kernel modules and (multi-GPU)
    // the NCCL channels exist before the timed region — the GPU counterpart of the reference's precomputed norms
    if (Q_rows > 0) {
        const int64_t nb = (int64_t)std::min<size_t>((size_t)batch, Q_rows);
        const int rc = multi ? vs_exact_mgpu_search_f32(multi, Q.data(), nb, k, opt.precision, ids.data(), dists.data())
                             : vs_exact_search_f32(index, Q.data(), nb, k, opt.precision, ids.data(), dists.data());
        if (rc != VS_OK) {
            std::cerr << "Error: " << vs_last_error() << std::endl;
            if (multi) vs_exact_mgpu_destroy(multi);
            else vs_exact_destroy(index);
            return false;
        }
    }
    const double build_s = std::chrono::duration<double>(clock::now() - build_start).count();

    std::vector<double> batch_times;
    int precision_used = 0, launches = 0;

    const auto total_start = clock::now();
    for (size_t i = 0; i < Q_rows; i += (size_t)batch) {
        const int64_t nb = (int64_t)std::min<size_t>((size_t)batch, Q_rows - i);
        const auto t0 = clock::now();
        const int rc = multi ? vs_exact_mgpu_search_f32(multi, &Q[i * (size_t)Q_dim], nb, k, opt.precision, &ids[i * (size_t)k],
                                                        &dists[i * (size_t)k])
                             : vs_exact_search_f32(index, &Q[i * (size_t)Q_dim], nb, k, opt.precision, &ids[i * (size_t)k],
                                                   &dists[i * (size_t)k]);
        if (rc != VS_OK) {
            std::cerr << "Error: " << vs_last_error() << std::endl;
            if (multi) vs_exact_mgpu_destroy(multi);
            else vs_exact_destroy(index);
            return false;
        }
        batch_times.push_back(std::chrono::duration<double>(clock::now() - t0).count());
    }
    const double total_time = std::chrono::duration<double>(clock::now() - total_start).count();
    int n_gpus = 1, n_shards = 1;
    if (multi) {
        n_gpus = vs_exact_mgpu_num_gpus(multi);
        n_shards = vs_exact_mgpu_num_shards(multi);
        precision_used = opt.precision;
        vs_exact_mgpu_destroy(multi);
    } else {
        vs_exact_last_launches(index, &launches, &precision_used);
        vs_exact_destroy(index);
    }

    static const char* prec_names[] = {"auto", "fp32 (3xTF32 tensor-core split)", "fp32 (FFMA stream)", "1xTF32",
                                       "fp16 candidate pass + exact fp32 refine, certified"};
    std::cout << "\n=== B200 RAG Performance Metrics ===" << std::endl;
    std::cout << "\nDataset Information:" << std::endl;
    std::cout << "  Number of queries: " << Q_rows << std::endl;
    std::cout << "  Number of documents: " << B_rows << std::endl;
    std::cout << "  Dimension: " << Q_dim << std::endl;
    std::cout << "  Top-K: " << k << std::endl;
    std::cout << "  Batch size: " << batch << std::endl;
    std::cout << "  Arithmetic: " << prec_names[precision_used >= 0 && precision_used <= 4 ? precision_used : 0] << std::endl;
    std::cout << "  GPUs: " << n_gpus << " (" << n_shards << " row shard" << (n_shards > 1 ? "s" : "") << ")" << std::endl;
    std::cout << "  Index build (untimed, upload + norms): " << build_s << " s" << std::endl;
    std::cout << "\nOverall Performance:" << std::endl;
    std::cout << "  Total execution time: " << total_time << " s" << std::endl;
    std::cout << "  Throughput: " << (total_time > 0 ? (double)Q_rows / total_time : 0.0) << " queries/sec" << std::endl;
    print_stats("End-to-End Per Batch (upload + fused distance/top-k + download)", vsbio::compute_statistics(batch_times), "batch");
    std::cout << "\nWriting results to " << output_file << "..." << std::endl;
    if (!write_results(output_file, ids, dists, Q_rows, k)) {
        std::cerr << "Failed to write results!" << std::endl;
        return false;
    }
    std::cout << "\nDone processing " << dataset_name << "!\n" << std::endl;
    return true;
}

}  // namespace

int main(int argc, char* argv[]) {
    std::cout << "=== B200 drop-in for the CPU baseline k-NN search ===" << std::endl;
    int ndev = 0;
    if (vs_device_count(&ndev) != VS_OK) {
        std::cerr << "Error: " << vs_last_error() << " (libvsb200 has no CPU fallback)" << std::endl;
        return 1;
    }
    std::cout << "Using libvsb200 ABI " << vs_abi_version() << ", " << ndev << " CUDA device(s)" << std::endl;
    std::cout << "====================================\n" << std::endl;

    Options opt;
    std::vector<std::string> pos;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (a == "--batch" && i + 1 < argc) {
            opt.batch = std::atoll(argv[++i]);
        } else if (a == "--device" && i + 1 < argc) {
            opt.device = std::atoi(argv[++i]);
        } else if (a == "--gpus" && i + 1 < argc) {
            opt.gpus = std::atoi(argv[++i]);
        } else if (a == "--shards-per-gpu" && i + 1 < argc) {
            opt.shards_per_gpu = std::atoi(argv[++i]);
        } else if (a == "--precision" && i + 1 < argc) {
            const std::string p = argv[++i];
            if (p == "auto") opt.precision = VS_PREC_AUTO;
            else if (p == "3xtf32") opt.precision = VS_PREC_FP32_3XTF32;
            else if (p == "ffma") opt.precision = VS_PREC_FP32_FFMA;
            else if (p == "1xtf32") opt.precision = VS_PREC_TF32_1X;
            else if (p == "f16cert") opt.precision = VS_PREC_F16_CERTIFIED;
            else {
                std::cerr << "unknown --precision " << p << std::endl;
                return 1;
            }
        } else {
            pos.push_back(a);
        }
    }
    if (pos.size() == 4) {
        if (!run_benchmark(pos[0], pos[0], pos[1], std::atoi(pos[2].c_str()), pos[3], opt)) return 1;
    } else if (pos.empty()) {  // the reference's hard-coded runs: a failing dataset is reported and skipped (rc 0)
        const int k = 5;
        run_benchmark("SIFT-small", "siftsmall/siftsmall_base.fvecs", "siftsmall/siftsmall_query.fvecs", k, "siftsmall_results.txt", opt);
        run_benchmark("SIFT", "sift/sift_base.fvecs", "sift/sift_query.fvecs", k, "sift_results.txt", opt);
    } else {
        std::cerr << "Usage: " << argv[0] << " <base.fvecs> <query.fvecs> <k> <results.txt> [--batch B] [--precision P] [--device D] [--gpus N]"
                  << std::endl;
        return 1;
    }
    std::cout << "\n========================================" << std::endl;
    std::cout << "All benchmarks completed!" << std::endl;
    std::cout << "========================================" << std::endl;
    return 0;
}
