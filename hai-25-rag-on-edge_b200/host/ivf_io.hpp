// Host-side I/O for the on-disk IVF index directory of the reference:
//   ivf_config.json, cluster_offsets.npy, cluster_indices.npy | reorder_to_original.npy,
//   vectors.npy | vectors.bin | vectors_reordered.npy, centroids.npy, cluster_ids.npy, cluster_sizes.npy
// written by qidk_ivf/prepare/create_ivf_model.py:135-166 / create_ivf_model_reordered.py:142-169 and read by
// IVFIndex::loadConfig / loadClusterData / loadVectors (qidk_ivf/android/app/main/jni/IVFIndex.cpp:181-267).
// Plain C++17, no CUDA.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace vsb_io {

struct IvfConfig {
    size_t n_vectors = 0, n_clusters = 0, dim = 0;
    float avg_cluster_size = 0.f;
    bool reordered = false;
};

// Every function returns an empty string on success, else the error message.
std::string read_text(const std::string& path, std::string& out);
std::string parse_ivf_config(const std::string& json, IvfConfig& cfg);
std::string load_npy_f32(const std::string& path, std::vector<float>& data, std::vector<size_t>& shape);
std::string load_npy_i32(const std::string& path, std::vector<int32_t>& data, std::vector<size_t>& shape);
std::string save_npy(const std::string& path, const void* data, const char* descr, const std::vector<size_t>& shape,
                     size_t elem_bytes);
std::string read_fvecs(const std::string& path, std::vector<float>& data, size_t& rows, size_t& dim);
std::string read_ivecs(const std::string& path, std::vector<int32_t>& data, size_t& rows, size_t& dim);
bool file_exists(const std::string& path);

}  // namespace vsb_io
