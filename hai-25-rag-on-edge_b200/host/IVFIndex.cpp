#include "IVFIndex.h"

#include <chrono>
#include <stdexcept>

#include "../../include/vsb200.h"

namespace {
[[noreturn]] void raise(const char* what) { throw std::runtime_error(std::string(what) + ": " + vs_last_error()); }
}  // namespace

IVFIndex::IVFIndex(const std::string& indexDir, const std::string& /*backendPath*/, int device) {
    if (vs_ivf_open(&m_handle, indexDir.c_str(), device) != VS_OK) raise("IVFIndex: cannot load index");
    m_numVectors = (size_t)vs_ivf_num_vectors(m_handle);
    m_numClusters = (size_t)vs_ivf_num_clusters(m_handle);
    m_dim = (size_t)vs_ivf_dim(m_handle);
    m_avgClusterSize = vs_ivf_avg_cluster_size(m_handle);
    vs_ivf_set_profile(m_handle, 1);
}

IVFIndex::~IVFIndex() {
    if (m_handle) vs_ivf_destroy(m_handle);
}

size_t IVFIndex::search(const std::vector<float>& query, int k, int nprobe, std::vector<int>& indices, std::vector<float>& scores) {
    SearchTiming timing;
    return search(query, k, nprobe, indices, scores, timing);
}

size_t IVFIndex::search(const std::vector<float>& query, int k, int nprobe, std::vector<int>& indices, std::vector<float>& scores,
                        SearchTiming& timing) {
    std::vector<std::vector<int>> all_i;
    std::vector<std::vector<float>> all_s;
    const size_t cand = searchBatch(query, 1, k, nprobe, all_i, all_s, timing);
    indices = std::move(all_i[0]);
    scores = std::move(all_s[0]);
    return cand;
}

size_t IVFIndex::searchBatch(const std::vector<float>& queries, int batchSize, int k, int nprobe,
                             std::vector<std::vector<int>>& allIndices, std::vector<std::vector<float>>& allScores,
                             SearchTiming& timing) {
    const auto t0 = std::chrono::high_resolution_clock::now();
    if (batchSize < 0 || queries.size() != (size_t)batchSize * m_dim)
        throw std::runtime_error("IVFIndex::searchBatch: query buffer size " + std::to_string(queries.size()) +
                                 " != batchSize * dim (" + std::to_string((size_t)batchSize * m_dim) + ")");
    m_ids.resize((size_t)batchSize * (size_t)k);
    m_scores.resize((size_t)batchSize * (size_t)k);
    m_counts.resize((size_t)batchSize);
    uint64_t total = 0;
    if (vs_ivf_search(m_handle, queries.data(), batchSize, k, nprobe, m_ids.data(), m_scores.data(), m_counts.data(), &total) != VS_OK)
        raise("IVFIndex::searchBatch");
    allIndices.assign((size_t)batchSize, {});
    allScores.assign((size_t)batchSize, {});
    for (int b = 0; b < batchSize; ++b) {
        const int n = m_counts[(size_t)b];  // min(k, candidates), IVFIndex.cpp:457,735
        allIndices[(size_t)b].assign(m_ids.begin() + (size_t)b * k, m_ids.begin() + (size_t)b * k + n);
        allScores[(size_t)b].assign(m_scores.begin() + (size_t)b * k, m_scores.begin() + (size_t)b * k + n);
    }
    float fine_ms = 0.f;
    vs_ivf_last_kernel_ms(m_handle, &fine_ms);
    timing.total_ms = std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t0).count();
    timing.fine_search_ms = fine_ms;
    timing.gather_ms = 0.0;
    timing.centroid_search_ms = std::max(0.0, timing.total_ms - fine_ms);  // coarse kernels + copies + host glue
    return (size_t)total;
}
