// QnnRunner — the brute-force INT8 runner of the reference's qidk_rag_demo (qidk_bruteforce/android/app/main/jni/
// QnnRunner.h:18-52) re-hosted on libvsb200.  Where the reference loads a QNN context binary whose MatMul weights are
// the u8-quantised documents, this class quantises a documents .fvecs file at construction (same u8 encodings:
// input scale 0.6627451, output scale 1013.4312, QnnRunner.cpp:70-71; weights by min/max).  executeRaw /
// executeBatchRaw keep their contract — after the call getRawOutputBuffer() holds the raw u8 scores
// [batch x num_docs] — and searchTopK() is the fused form (quantise + MatMul + requant + find_top_k_int8 in one
// kernel, no score matrix), which is what the drop-in driver uses.  Errors throw std::runtime_error like the reference.
#ifndef VSB_QNNRUNNER_H
#define VSB_QNNRUNNER_H

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

struct vs_int8;

struct ExecutionTiming {
    double quantize_ms = 0.0;       // folded into the fused kernel: reported as 0
    double graph_execute_ms = 0.0;  // device time of the fused INT8 kernel (or wall time of the raw-score call)
    double dequantize_ms = 0.0;     // raw u8 is returned: 0
    double total_ms = 0.0;
};

class QnnRunner {
public:
    // documentsPath: .fvecs of the documents (replaces the context binary); backendPath accepted and ignored;
    // batchSize: the fixed batch of the reference's model (getBatchSize(); callers zero-pad, main.cpp:206-211).
    QnnRunner(const std::string& documentsPath, const std::string& backendPath = "./libQnnHtp.so", size_t batchSize = 1,
              int device = 0);
    ~QnnRunner();
    QnnRunner(const QnnRunner&) = delete;
    QnnRunner& operator=(const QnnRunner&) = delete;

    void executeRaw(const std::vector<float>& query, ExecutionTiming& timing);
    void executeBatchRaw(const std::vector<float>& batch_queries, ExecutionTiming& timing);
    const uint8_t* getRawOutputBuffer() const { return m_output.data(); }
    size_t getOutputSize() const { return m_output.size(); }
    float getOutputScale() const { return m_outputScale; }
    size_t getBatchSize() const { return m_batch; }
    size_t getDim() const { return m_dim; }
    size_t getNumDocs() const { return m_numDocs; }

    // fused path: nq queries (flat, nq x dim) -> ids / raw u8 scores [nq x k], (score desc, id asc)
    void searchTopK(const float* queries, size_t nq, int k, std::vector<int32_t>& ids, std::vector<uint8_t>& scores,
                    ExecutionTiming& timing);

private:
    vs_int8* m_handle = nullptr;
    size_t m_batch = 1, m_dim = 0, m_numDocs = 0;
    float m_outputScale = 0.f;
    std::vector<uint8_t> m_output;
};

#endif  // VSB_QNNRUNNER_H
