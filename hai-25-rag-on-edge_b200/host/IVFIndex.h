// IVFIndex — C++ class with the reference's public surface (qidk_ivf/android/app/main/jni/IVFIndex.h:14-54) on top of
// libvsb200's C ABI: same constructor arguments, same search / searchBatch signatures and return value (total
// candidates scanned), same getters, exceptions (std::runtime_error) on failure like the reference's loaders
// (IVFIndex.cpp:184-258).  Scores are inner products, descending; ids are original base ids (IVFIndex.cpp:771-779).
#ifndef VSB_IVFINDEX_H
#define VSB_IVFINDEX_H

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

struct vs_ivf;

class IVFIndex {
public:
    // indexDir: the reference's index directory (ivf_config.json + .npy files).  backendPath is accepted for
    // source compatibility (the reference dlopens libQnnHtp.so for the coarse MatMul, IVFIndex.cpp:167-169); the
    // coarse search here is a CUDA kernel and the argument is ignored.  device = CUDA ordinal.
    explicit IVFIndex(const std::string& indexDir, const std::string& backendPath = "./libQnnHtp.so", int device = 0);
    ~IVFIndex();
    IVFIndex(const IVFIndex&) = delete;
    IVFIndex& operator=(const IVFIndex&) = delete;

    size_t search(const std::vector<float>& query, int k, int nprobe, std::vector<int>& indices, std::vector<float>& scores);

    struct SearchTiming {
        double centroid_search_ms = 0.0;  // coarse scores + probe selection (device)
        double gather_ms = 0.0;           // always 0: lists are contiguous on the device, nothing is gathered
        double fine_search_ms = 0.0;      // list-scan kernel (device)
        double total_ms = 0.0;            // wall clock of the call, copies included
    };

    size_t search(const std::vector<float>& query, int k, int nprobe, std::vector<int>& indices, std::vector<float>& scores,
                  SearchTiming& timing);

    // queries: batchSize x dim, flat.  Every one of the batchSize rows is searched (callers zero-pad short batches,
    // main_ivf.cpp:149-161) and counted in the return value, like the reference (IVFIndex.cpp:728,858).
    size_t searchBatch(const std::vector<float>& queries, int batchSize, int k, int nprobe,
                       std::vector<std::vector<int>>& allIndices, std::vector<std::vector<float>>& allScores,
                       SearchTiming& timing);

    size_t getNumVectors() const { return m_numVectors; }
    size_t getNumClusters() const { return m_numClusters; }
    size_t getDim() const { return m_dim; }
    float getAvgClusterSize() const { return m_avgClusterSize; }

private:
    vs_ivf* m_handle = nullptr;
    size_t m_numVectors = 0, m_numClusters = 0, m_dim = 0;
    float m_avgClusterSize = 0.f;
    std::vector<int32_t> m_ids;
    std::vector<float> m_scores;
    std::vector<int32_t> m_counts;
};

#endif  // VSB_IVFINDEX_H
