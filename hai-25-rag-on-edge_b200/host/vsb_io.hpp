// Host-side file formats and statistics shared by the drop-in executables (C++17, no CUDA).
//   fvecs / ivecs   little-endian records [int32 d][d x 4 bytes]; all records share d, a trailing partial record is an
//                   error ("File seems truncated.") — same acceptance rules as the reference readers
//                   (cpu/cpu_baseline.cpp:31-58, qidk_ivf/.../main_ivf.cpp:18-50), read in one bulk pass with size_t indexing
//   timing stats    mean / population std-dev / min / max / p50 / p95 / p99 by index n*p/100 (cpu_baseline.cpp:60-93)
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace vsbio {

template <class T>
inline void read_vecs(const std::string& path, std::vector<T>& data, size_t& rows, int& dim) {
    static_assert(sizeof(T) == 4, "fvecs / ivecs hold 4-byte elements");
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("Cannot open file " + path);
    std::fseek(f, 0, SEEK_END);
    const long long bytes = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    rows = 0;
    dim = 0;
    data.clear();
    if (bytes == 0) {
        std::fclose(f);
        return;
    }
    int32_t d = 0;
    if (bytes < 4 || std::fread(&d, 4, 1, f) != 1 || d <= 0) {
        std::fclose(f);
        throw std::runtime_error("File seems truncated.");
    }
    const size_t rec = 4 + (size_t)d * 4;
    if ((size_t)bytes % rec != 0) {
        std::fclose(f);
        throw std::runtime_error("File seems truncated.");
    }
    rows = (size_t)bytes / rec;
    dim = d;
    data.resize(rows * (size_t)d);
    std::fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> buf;
    const size_t chunk_rows = std::max<size_t>(1, (64u << 20) / rec);
    buf.resize(chunk_rows * rec);
    for (size_t r0 = 0; r0 < rows; r0 += chunk_rows) {
        const size_t n = std::min(chunk_rows, rows - r0);
        if (std::fread(buf.data(), rec, n, f) != n) {
            std::fclose(f);
            throw std::runtime_error("File seems truncated.");
        }
        for (size_t i = 0; i < n; ++i) {
            int32_t di;
            std::memcpy(&di, buf.data() + i * rec, 4);
            if (di != d) {
                std::fclose(f);
                throw std::runtime_error("Inconsistent dimension.");
            }
            std::memcpy(&data[(r0 + i) * (size_t)d], buf.data() + i * rec + 4, (size_t)d * 4);
        }
    }
    std::fclose(f);
}

inline void read_fvecs(const std::string& path, std::vector<float>& data, size_t& rows, int& dim) {
    read_vecs<float>(path, data, rows, dim);
}
inline void read_ivecs(const std::string& path, std::vector<int32_t>& data, size_t& rows, int& dim) {
    read_vecs<int32_t>(path, data, rows, dim);
}

struct TimingStats {
    double mean = 0, std_dev = 0, min_val = 0, max_val = 0, p50 = 0, p95 = 0, p99 = 0;
};

inline TimingStats compute_statistics(std::vector<double> t) {
    TimingStats s;
    if (t.empty()) return s;
    std::sort(t.begin(), t.end());
    const size_t n = t.size();
    s.min_val = t.front();
    s.max_val = t.back();
    double sum = 0;
    for (double v : t) sum += v;
    s.mean = sum / (double)n;
    double sq = 0;
    for (double v : t) sq += (v - s.mean) * (v - s.mean);
    s.std_dev = std::sqrt(sq / (double)n);
    s.p50 = t[n * 50 / 100];
    s.p95 = t[n * 95 / 100];
    s.p99 = t[n * 99 / 100];
    return s;
}

}  // namespace vsbio
