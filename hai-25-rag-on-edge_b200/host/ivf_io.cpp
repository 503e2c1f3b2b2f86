#include "ivf_io.hpp"

#include <sys/stat.h>

#include <cctype>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>

namespace vsb_io {

bool file_exists(const std::string& path) {
    struct stat st;
    return ::stat(path.c_str(), &st) == 0;
}

std::string read_text(const std::string& path, std::string& out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return "Cannot open config file: " + path;
    std::ostringstream ss;
    ss << f.rdbuf();
    out = ss.str();
    return "";
}

// `"key" : value` located by substring search, like the reference's mini parser (IVFIndex.cpp:13-50); unlike it a
// malformed boolean is an error instead of "true".
static bool find_value(const std::string& json, const std::string& key, size_t& pos) {
    pos = json.find("\"" + key + "\"");
    if (pos == std::string::npos) return false;
    pos = json.find(':', pos);
    if (pos == std::string::npos) return false;
    ++pos;
    while (pos < json.size() && std::isspace((unsigned char)json[pos])) ++pos;
    return pos < json.size();
}

std::string parse_ivf_config(const std::string& json, IvfConfig& cfg) {
    size_t p;
    auto get_size = [&](const char* key, size_t& v) -> bool {
        if (!find_value(json, key, p) || !std::isdigit((unsigned char)json[p])) return false;
        v = std::strtoull(json.c_str() + p, nullptr, 10);
        return true;
    };
    if (!get_size("n_vectors", cfg.n_vectors)) return "Missing n_vectors in config";
    if (!get_size("n_clusters", cfg.n_clusters)) return "Missing n_clusters in config";
    if (!get_size("dim", cfg.dim)) return "Missing dim in config";
    if (find_value(json, "avg_cluster_size", p)) cfg.avg_cluster_size = std::strtof(json.c_str() + p, nullptr);
    cfg.reordered = false;
    if (find_value(json, "reordered", p)) {
        if (json.compare(p, 4, "true") == 0)
            cfg.reordered = true;
        else if (json.compare(p, 5, "false") == 0)
            cfg.reordered = false;
        else
            return "Malformed boolean for \"reordered\" in config";
    }
    return "";
}

// NPY v1/v2/v3, little-endian C-order only; dtype checked against `want` ("<f4" / "<i4").
static std::string load_npy(const std::string& path, const char* want, size_t elem, std::vector<uint8_t>& raw,
                            std::vector<size_t>& shape) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return "cannot open " + path;
    char magic[6];
    f.read(magic, 6);
    if (!f || std::memcmp(magic, "\x93NUMPY", 6) != 0) return "not an NPY file: " + path;
    uint8_t ver[2];
    f.read((char*)ver, 2);
    uint32_t hlen = 0;
    if (ver[0] == 1) {
        uint16_t h16;
        f.read((char*)&h16, 2);
        hlen = h16;
    } else {
        f.read((char*)&hlen, 4);  // the reference truncates this to 16 bits (IVFIndex.cpp:72); we do not
    }
    if (!f || hlen > (1u << 24)) return "bad NPY header in " + path;
    std::string hdr(hlen, '\0');
    f.read(&hdr[0], hlen);
    if (!f) return "truncated NPY header in " + path;
    if (hdr.find(want) == std::string::npos) return std::string("NPY dtype is not ") + want + ": " + path;
    if (hdr.find("'fortran_order': True") != std::string::npos) return "fortran-ordered NPY not supported: " + path;
    size_t s0 = hdr.find("'shape': (");
    if (s0 == std::string::npos) return "NPY header has no shape: " + path;
    s0 += 10;
    const size_t s1 = hdr.find(')', s0);
    if (s1 == std::string::npos) return "NPY shape tuple is not closed: " + path;
    // the payload cannot be larger than what is left of the file: a corrupt header must not drive the allocation
    const std::streampos data_pos = f.tellg();
    f.seekg(0, std::ios::end);
    const std::streampos end_pos = f.tellg();
    f.seekg(data_pos);
    if (!f || end_pos < data_pos) return "cannot size " + path;
    const size_t avail = (size_t)(end_pos - data_pos);
    shape.clear();
    size_t total = 1;
    for (size_t i = s0; i < s1;) {
        while (i < s1 && !std::isdigit((unsigned char)hdr[i])) ++i;
        if (i >= s1) break;
        char* end = nullptr;
        const size_t d = std::strtoull(hdr.c_str() + i, &end, 10);
        shape.push_back(d);
        if (d != 0 && total > avail / d) return "NPY shape exceeds the file size: " + path;  // also catches overflow
        total *= d;
        i = (size_t)(end - hdr.c_str());
    }
    if (total > avail / elem) return "NPY shape exceeds the file size: " + path;
    raw.resize(total * elem);
    f.read((char*)raw.data(), (std::streamsize)raw.size());
    if ((size_t)f.gcount() != raw.size()) return "truncated NPY data in " + path;
    return "";
}

std::string load_npy_f32(const std::string& path, std::vector<float>& data, std::vector<size_t>& shape) {
    std::vector<uint8_t> raw;
    std::string e = load_npy(path, "<f4", 4, raw, shape);
    if (!e.empty()) return e;
    data.resize(raw.size() / 4);
    std::memcpy(data.data(), raw.data(), raw.size());
    return "";
}

std::string load_npy_i32(const std::string& path, std::vector<int32_t>& data, std::vector<size_t>& shape) {
    std::vector<uint8_t> raw;
    std::string e = load_npy(path, "<i4", 4, raw, shape);
    if (!e.empty()) return e;
    data.resize(raw.size() / 4);
    std::memcpy(data.data(), raw.data(), raw.size());
    return "";
}

// NPY v1.0 writer; header padded so that the data starts at a multiple of 64 bytes (what numpy writes).
std::string save_npy(const std::string& path, const void* data, const char* descr, const std::vector<size_t>& shape,
                     size_t elem_bytes) {
    std::string hdr = std::string("{'descr': '") + descr + "', 'fortran_order': False, 'shape': (";
    size_t total = 1;
    for (size_t i = 0; i < shape.size(); ++i) {
        hdr += std::to_string(shape[i]);
        if (shape.size() == 1 || i + 1 < shape.size()) hdr += ",";
        if (i + 1 < shape.size()) hdr += " ";
        total *= shape[i];
    }
    hdr += "), }";
    const size_t unpadded = 10 + hdr.size() + 1;
    hdr.append((64 - unpadded % 64) % 64, ' ');
    hdr += "\n";
    std::ofstream f(path, std::ios::binary);
    if (!f) return "cannot create " + path;
    const uint16_t h16 = (uint16_t)hdr.size();
    f.write("\x93NUMPY\x01\x00", 8);
    f.write((const char*)&h16, 2);
    f.write(hdr.data(), (std::streamsize)hdr.size());
    f.write((const char*)data, (std::streamsize)(total * elem_bytes));
    return f.good() ? "" : "write failed: " + path;
}

template <class T>
static std::string read_vecs(const std::string& path, std::vector<T>& data, size_t& rows, size_t& dim) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) return "Cannot open file " + path;
    const size_t bytes = (size_t)f.tellg();
    f.seekg(0);
    rows = dim = 0;
    data.clear();
    if (bytes == 0) return "";
    int32_t d = 0;
    f.read((char*)&d, 4);
    if (!f || d <= 0) return "Bad record header in " + path;
    dim = (size_t)d;
    const size_t rec = 4 + dim * 4;
    if (bytes % rec != 0) return "File seems truncated.";  // cpu_baseline.cpp:53-56
    rows = bytes / rec;
    data.resize(rows * dim);
    std::vector<uint8_t> buf(std::min<size_t>(rows, 4096) * rec);
    f.seekg(0);
    size_t done = 0;
    while (done < rows) {
        const size_t n = std::min<size_t>(4096, rows - done);
        f.read((char*)buf.data(), (std::streamsize)(n * rec));
        if ((size_t)f.gcount() != n * rec) return "File seems truncated.";
        for (size_t r = 0; r < n; ++r) {
            int32_t dd;
            std::memcpy(&dd, buf.data() + r * rec, 4);
            if ((size_t)dd != dim) return "Inconsistent dimension.";  // cpu_baseline.cpp:43-46
            std::memcpy(&data[(done + r) * dim], buf.data() + r * rec + 4, dim * 4);
        }
        done += n;
    }
    return "";
}

std::string read_fvecs(const std::string& path, std::vector<float>& data, size_t& rows, size_t& dim) {
    return read_vecs<float>(path, data, rows, dim);
}
std::string read_ivecs(const std::string& path, std::vector<int32_t>& data, size_t& rows, size_t& dim) {
    return read_vecs<int32_t>(path, data, rows, dim);
}

}  // namespace vsb_io
