#include "QnnRunner.h"

#include <chrono>
#include <stdexcept>

#include "../../include/vsb200.h"
#include "vsb_io.hpp"

namespace {
constexpr float kInputScale = 0.6627451f;    // QnnRunner.cpp:70
constexpr float kOutputScale = 1013.4312f;   // QnnRunner.cpp:71
[[noreturn]] void raise(const char* what) { throw std::runtime_error(std::string(what) + ": " + vs_last_error()); }
double ms_since(std::chrono::high_resolution_clock::time_point t0) {
    return std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t0).count();
}
}  // namespace

QnnRunner::QnnRunner(const std::string& documentsPath, const std::string& /*backendPath*/, size_t batchSize, int device)
    : m_batch(batchSize ? batchSize : 1) {
    std::vector<float> docs;
    size_t rows = 0;
    int dim = 0;
    vsbio::read_fvecs(documentsPath, docs, rows, dim);
    if (rows == 0) throw std::runtime_error("QnnRunner: no documents in " + documentsPath);
    if (vs_int8_create(&m_handle, docs.data(), (int64_t)rows, dim, kInputScale, 0.0f, kOutputScale, device, 0) != VS_OK)
        raise("QnnRunner: cannot build the INT8 index");
    m_dim = (size_t)dim;
    m_numDocs = rows;
    m_outputScale = vs_int8_output_scale(m_handle);
    vs_int8_set_profile(m_handle, 1);
}

QnnRunner::~QnnRunner() {
    if (m_handle) vs_int8_destroy(m_handle);
}

void QnnRunner::executeRaw(const std::vector<float>& query, ExecutionTiming& timing) {
    if (query.size() != m_dim) throw std::runtime_error("Query size mismatch");  // QnnRunner.cpp:537
    const auto t0 = std::chrono::high_resolution_clock::now();
    m_output.resize(m_numDocs);
    if (vs_int8_scores_raw(m_handle, query.data(), 1, m_output.data()) != VS_OK) raise("QnnRunner::executeRaw");
    timing = ExecutionTiming{};
    timing.total_ms = timing.graph_execute_ms = ms_since(t0);
}

void QnnRunner::executeBatchRaw(const std::vector<float>& batch_queries, ExecutionTiming& timing) {
    if (batch_queries.size() != m_batch * m_dim) throw std::runtime_error("Batch query size mismatch");  // QnnRunner.cpp:578
    const auto t0 = std::chrono::high_resolution_clock::now();
    m_output.resize(m_batch * m_numDocs);
    if (vs_int8_scores_raw(m_handle, batch_queries.data(), (int64_t)m_batch, m_output.data()) != VS_OK)
        raise("QnnRunner::executeBatchRaw");
    timing = ExecutionTiming{};
    timing.total_ms = timing.graph_execute_ms = ms_since(t0);
}

void QnnRunner::searchTopK(const float* queries, size_t nq, int k, std::vector<int32_t>& ids, std::vector<uint8_t>& scores,
                           ExecutionTiming& timing) {
    const auto t0 = std::chrono::high_resolution_clock::now();
    ids.resize(nq * (size_t)k);
    scores.resize(nq * (size_t)k);
    if (vs_int8_search(m_handle, queries, (int64_t)nq, k, ids.data(), scores.data()) != VS_OK) raise("QnnRunner::searchTopK");
    timing = ExecutionTiming{};
    float ms = 0.f;
    if (nq > 0 && vs_int8_last_kernel_ms(m_handle, &ms) == VS_OK) timing.graph_execute_ms = ms;
    timing.total_ms = ms_since(t0);
}
