// `qidk_ivf_search` — drop-in for the reference's IVF benchmark driver (qidk_ivf/android/app/main/jni/main_ivf.cpp):
//   qidk_ivf_search <index_dir> <queries.fvecs> <results_dir> <backend.so> <top_k> [nprobe=16] [groundtruth.ivecs] [batch=1]
// same positional arguments (backend.so is accepted and ignored), same results.txt line format (4-decimal scores,
// main_ivf.cpp:179-183) and the same metrics.txt sections and labels (main_ivf.cpp:212-273), so existing sweep
// scripts keep parsing the output.  Recall@k = hits / k against the first k ground-truth ids (main_ivf.cpp:52-59).
#include <sys/stat.h>

#include <algorithm>
#include <chrono>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <set>
#include <string>
#include <vector>

#include "IVFIndex.h"
#include "vsb_io.hpp"

static double compute_recall(const std::vector<int>& predicted, const int32_t* gt, int gt_k, int k) {
    std::set<int> truth(gt, gt + std::min(k, gt_k));
    int hits = 0;
    for (int i = 0; i < std::min(k, (int)predicted.size()); ++i) hits += truth.count(predicted[(size_t)i]) ? 1 : 0;
    return (double)hits / k;
}

int main(int argc, char* argv[]) {
    if (argc < 6) {
        std::cerr << "Usage: " << argv[0]
                  << " <index_dir> <queries.fvecs> <results_dir> <backend.so> <top_k> [nprobe] [groundtruth.ivecs] [batch_size]"
                  << std::endl;
        return 1;
    }
    const std::string index_dir = argv[1], query_file = argv[2], results_dir = argv[3], backend_path = argv[4];
    const int TOP_K = std::stoi(argv[5]);
    const int NPROBE = argc > 6 ? std::stoi(argv[6]) : 16;
    const std::string gt_file = argc > 7 ? argv[7] : "";
    const int BATCH_SIZE = argc > 8 ? std::max(1, std::stoi(argv[8])) : 1;
    try {
        mkdir(results_dir.c_str(), 0755);
        const std::string results_txt = results_dir + "/results.txt", metrics_txt = results_dir + "/metrics.txt";

        std::cout << "Loading queries..." << std::endl;
        std::vector<float> queries;
        size_t num_queries = 0;
        int query_dim = 0;
        vsbio::read_fvecs(query_file, queries, num_queries, query_dim);
        std::cout << "Loaded " << num_queries << " queries." << std::endl;

        std::vector<int32_t> gt;
        size_t gt_rows = 0;
        int gt_k = 0;
        if (!gt_file.empty()) {
            std::cout << "Loading ground truth..." << std::endl;
            vsbio::read_ivecs(gt_file, gt, gt_rows, gt_k);
            std::cout << "Loaded " << gt_rows << " ground truth entries." << std::endl;
        }

        std::cout << "Loading IVF index..." << std::endl;
        IVFIndex ivf(index_dir, backend_path);
        if (query_dim != (int)ivf.getDim())
            throw std::runtime_error("Query dim (" + std::to_string(query_dim) + ") != index dim (" + std::to_string(ivf.getDim()) + ")");

        std::ofstream results_file(results_txt);
        if (!results_file) throw std::runtime_error("Cannot open results file: " + results_txt);

        double total_centroid_ms = 0, total_gather_ms = 0, total_fine_ms = 0, total_search_ms = 0, total_recall = 0;
        size_t total_candidates = 0;
        std::vector<double> latencies;
        latencies.reserve(num_queries);
        std::cout << "\nProcessing " << num_queries << " queries with nprobe=" << NPROBE << ", batch_size=" << BATCH_SIZE << "..."
                  << std::endl;
        const auto total_start = std::chrono::high_resolution_clock::now();
        std::vector<float> batch_queries((size_t)BATCH_SIZE * query_dim);
        for (size_t i = 0; i < num_queries; i += (size_t)BATCH_SIZE) {
            const size_t cur = std::min((size_t)BATCH_SIZE, num_queries - i);
            std::copy(queries.begin() + i * query_dim, queries.begin() + (i + cur) * query_dim, batch_queries.begin());
            std::fill(batch_queries.begin() + cur * query_dim, batch_queries.end(), 0.0f);  // zero padding, main_ivf.cpp:149-153
            std::vector<std::vector<int>> batch_indices;
            std::vector<std::vector<float>> batch_scores;
            IVFIndex::SearchTiming timing;
            total_candidates += ivf.searchBatch(batch_queries, BATCH_SIZE, TOP_K, NPROBE, batch_indices, batch_scores, timing);
            total_centroid_ms += timing.centroid_search_ms;
            total_gather_ms += timing.gather_ms;
            total_fine_ms += timing.fine_search_ms;
            total_search_ms += timing.total_ms;
            for (size_t j = 0; j < cur; ++j) {
                latencies.push_back(timing.total_ms);
                if (!gt.empty() && i + j < gt_rows)
                    total_recall += compute_recall(batch_indices[j], &gt[(i + j) * (size_t)gt_k], gt_k, TOP_K);
                results_file << "Query " << (i + j) << ":";
                for (size_t t = 0; t < batch_indices[j].size(); ++t)
                    results_file << " (" << batch_indices[j][t] << ", " << std::fixed << std::setprecision(4) << batch_scores[j][t] << ")";
                results_file << "\n";
            }
            if ((i + cur) % 100 == 0 || i + cur >= num_queries)
                std::cout << "Processed " << (i + cur) << "/" << num_queries << " queries." << std::endl;
        }
        const double total_s = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - total_start).count();
        results_file.close();

        const double avg_latency = total_search_ms / num_queries;
        const double avg_candidates = (double)total_candidates / num_queries;
        const double avg_recall = gt.empty() ? 0.0 : total_recall / num_queries;
        const double throughput = num_queries / total_s;
        std::sort(latencies.begin(), latencies.end());
        const double p50 = latencies[latencies.size() / 2];
        const double p95 = latencies[(size_t)(latencies.size() * 0.95)];
        const double p99 = latencies[(size_t)(latencies.size() * 0.99)];
        const double speedup_candidates = (double)ivf.getNumVectors() / avg_candidates;

        std::ofstream m(metrics_txt);
        if (!m) throw std::runtime_error("Cannot open metrics file: " + metrics_txt);
        m << std::fixed << std::setprecision(6);
        m << "=== IVF Search Performance Metrics ===\n\n";
        m << "Index Configuration:\n";
        m << "  Total vectors: " << ivf.getNumVectors() << "\n  Number of clusters: " << ivf.getNumClusters() << "\n  Dimension: "
          << ivf.getDim() << "\n  nprobe: " << NPROBE << "\n  top_k: " << TOP_K << "\n  batch_size: " << BATCH_SIZE << "\n\n";
        m << "Query Statistics:\n";
        m << "  Number of queries: " << num_queries << "\n  Avg candidates searched: " << avg_candidates
          << "\n  Candidate reduction: " << speedup_candidates << "x\n\n";
        if (!gt.empty()) m << "Accuracy:\n  Recall@" << TOP_K << ": " << (avg_recall * 100.0) << "%\n\n";
        m << "Latency:\n";
        m << "  Avg per query (amortized): " << avg_latency << " ms\n";
        m << "  Avg centroid search (NPU): " << (total_centroid_ms / num_queries) << " ms\n";
        m << "  Avg gather: " << (total_gather_ms / num_queries) << " ms\n";
        m << "  Avg fine search (NEON): " << (total_fine_ms / num_queries) << " ms\n";
        m << "  Batch P50: " << p50 << " ms\n  Batch P95: " << p95 << " ms\n  Batch P99: " << p99 << " ms\n\n";
        m << "Throughput:\n  Total time: " << total_s << " s\n  QPS: " << throughput << "\n\n";
        const double centroid_flops = 2.0 * ivf.getDim() * ivf.getNumClusters();
        const double fine_flops = 2.0 * ivf.getDim() * avg_candidates;
        const double flops_per_query = centroid_flops + fine_flops;
        m << "Compute:\n";
        m << "  FLOPs per query (centroid): " << std::scientific << centroid_flops << "\n";
        m << "  FLOPs per query (fine): " << fine_flops << "\n";
        m << "  FLOPs per query (total): " << flops_per_query << "\n";
        m << "  Avg GFLOPS: " << std::fixed << (flops_per_query / 1e9) / (avg_latency / 1000.0) << "\n";
        m << "  Total GFLOPS: " << flops_per_query * num_queries / (total_s * 1e9) << "\n\n";
        m << "Time Breakdown:\n";
        const double total_ms = total_s * 1000.0;
        m << "  Centroid search (NPU): " << (total_centroid_ms / total_ms * 100.0) << "%\n";
        m << "  Gather candidates: " << (total_gather_ms / total_ms * 100.0) << "%\n";
        m << "  Fine search (NEON): " << (total_fine_ms / total_ms * 100.0) << "%\n";
        m.close();

        std::cout << "\n=== IVF Search Complete ===" << std::endl;
        std::cout << "Throughput: " << throughput << " QPS" << std::endl;
        std::cout << "Avg latency: " << avg_latency << " ms" << std::endl;
        std::cout << "Avg candidates: " << avg_candidates << " (" << speedup_candidates << "x reduction)" << std::endl;
        if (!gt.empty()) std::cout << "Recall@" << TOP_K << ": " << (avg_recall * 100.0) << "%" << std::endl;
        std::cout << "\nResults saved to: " << results_txt << std::endl;
        std::cout << "Metrics saved to: " << metrics_txt << std::endl;
    } catch (const std::exception& e) {
        std::cerr << "FATAL ERROR: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
