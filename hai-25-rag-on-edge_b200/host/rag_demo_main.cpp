// `qidk_rag_demo` — drop-in for the reference's INT8 brute-force driver (qidk_bruteforce/android/app/main/jni/main.cpp):
//   qidk_rag_demo <context_binary> <queries.fvecs> <results_dir> <backend.so> <documents.fvecs> <top_k> [batch_size]
// Same positional arguments.  The context binary and backend are accepted and ignored: the "model" is built from
// documents.fvecs (the reference bakes the same file into the ONNX MatMul, create_model.py:57-87).  batch_size takes
// the place of the batch dimension the reference reads from the model (main.cpp:112-118; default 1).
// results.txt: `Query i: (index, u8*output_scale)` with 4 decimals (main.cpp:183-187); metrics.txt keeps the
// reference's section names (main.cpp:320-390) with the NPU / CPU split collapsed: top-k is fused into the kernel.
#include <sys/stat.h>

#include <algorithm>
#include <chrono>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <string>
#include <vector>

#include "QnnRunner.h"
#include "vsb_io.hpp"

int main(int argc, char* argv[]) {
    if (argc != 7 && argc != 8) {
        std::cerr << "Usage: " << argv[0]
                  << " <context_binary> <queries.fvecs> <results_dir> <backend.so> <documents.fvecs> <top_k> [batch_size]" << std::endl;
        return 1;
    }
    const std::string query_file = argv[2], results_dir = argv[3], backend_path = argv[4], doc_file = argv[5];
    const int TOP_K = std::stoi(argv[6]);
    const size_t batch = argc == 8 ? (size_t)std::max(1, std::stoi(argv[7])) : 1;
    try {
        mkdir(results_dir.c_str(), 0755);
        const std::string results_txt = results_dir + "/results.txt", metrics_txt = results_dir + "/metrics.txt";
        std::cout << "Loading queries..." << std::endl;
        std::vector<float> queries;
        size_t num_queries = 0;
        int query_dim = 0;
        vsbio::read_fvecs(query_file, queries, num_queries, query_dim);
        std::cout << "Loaded " << num_queries << " queries." << std::endl;
        std::cout << "Loading documents..." << std::endl;
        QnnRunner runner(doc_file, backend_path, batch);
        std::cout << "Loaded " << runner.getNumDocs() << " documents." << std::endl;
        std::cout << "Model configuration:" << std::endl;
        std::cout << "  Batch size: " << runner.getBatchSize() << std::endl;
        std::cout << "  Dimension: " << runner.getDim() << std::endl;
        std::cout << "  Number of documents: " << runner.getNumDocs() << std::endl;
        if (runner.getDim() != (size_t)query_dim) throw std::runtime_error("Model input dim and query dim mismatch!");
        std::ofstream results_file(results_txt);
        if (!results_file) throw std::runtime_error("Cannot open output file: " + results_txt);
        const float output_scale = runner.getOutputScale();

        std::vector<double> batch_ms, kernel_ms;
        std::vector<int32_t> ids;
        std::vector<uint8_t> scores;
        const auto total_start = std::chrono::high_resolution_clock::now();
        for (size_t b0 = 0; b0 < num_queries; b0 += batch) {
            const size_t nb = std::min(batch, num_queries - b0);  // no zero padding needed: any batch size is accepted
            ExecutionTiming t;
            runner.searchTopK(&queries[b0 * (size_t)query_dim], nb, TOP_K, ids, scores, t);
            batch_ms.push_back(t.total_ms);
            kernel_ms.push_back(t.graph_execute_ms);
            for (size_t i = 0; i < nb; ++i) {
                results_file << "Query " << (b0 + i) << ":";
                for (int r = 0; r < TOP_K; ++r)
                    results_file << " (" << ids[i * TOP_K + r] << ", " << std::fixed << std::setprecision(4)
                                 << (float)scores[i * TOP_K + r] * output_scale << ")";
                results_file << "\n";
            }
        }
        const double total_s = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - total_start).count();
        results_file.close();

        const vsbio::TimingStats bs = vsbio::compute_statistics(batch_ms);
        double kernel_total = 0;
        for (double v : kernel_ms) kernel_total += v;
        const double flops = 2.0 * (double)num_queries * (double)runner.getNumDocs() * (double)runner.getDim();
        const double bytes_per_batch = (double)runner.getNumDocs() * (double)runner.getDim();  // u8 documents streamed once per batch
        std::ofstream m(metrics_txt);
        if (!m) throw std::runtime_error("Cannot open metrics file: " + metrics_txt);
        m << std::fixed << std::setprecision(6);
        m << "=== RAG Performance Metrics (INT8, B200) ===\n\n";
        m << "Configuration:\n  Number of queries: " << num_queries << "\n  Number of documents: " << runner.getNumDocs()
          << "\n  Dimension: " << runner.getDim() << "\n  Top-K: " << TOP_K << "\n  Batch size: " << batch << "\n\n";
        m << "Latency:\n  Avg per batch: " << bs.mean << " ms\n  P50: " << bs.p50 << " ms\n  P95: " << bs.p95 << " ms\n  P99: "
          << bs.p99 << " ms\n  Avg fused kernel (quantize + MatMul + top-k): " << kernel_total / std::max<size_t>(1, kernel_ms.size())
          << " ms\n\n";
        m << "Throughput:\n  Total time: " << total_s << " s\n  QPS: " << num_queries / total_s << "\n\n";
        m << "Compute:\n  Total operations: " << std::scientific << flops << "\n  GOPS (wall): " << std::fixed << flops / total_s / 1e9
          << "\n  GOPS (kernel): " << (kernel_total > 0 ? flops / (kernel_total * 1e-3) / 1e9 : 0.0)
          << "\n  Operational intensity: " << (2.0 * (double)batch * bytes_per_batch) / bytes_per_batch << " ops/byte\n";
        m.close();
        std::cout << "\n=== Search Complete ===" << std::endl;
        std::cout << "Throughput: " << num_queries / total_s << " QPS" << std::endl;
        std::cout << "Results saved to: " << results_txt << std::endl;
        std::cout << "Metrics saved to: " << metrics_txt << std::endl;
    } catch (const std::exception& e) {
        std::cerr << "FATAL ERROR: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
