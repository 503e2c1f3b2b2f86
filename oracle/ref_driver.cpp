// TEST INFRASTRUCTURE — not product code.
//
// Driver linked against the UNMODIFIED reference translation unit
// /root/reference/cpu/cpu_baseline.cpp (compiled with -Dmain=ref_main, see oracle/Makefile).
// The reference's main() ignores argv and hard-codes k=5 (cpu_baseline.cpp:323-345), so to run it on
// arbitrary files / k we call its own entry points:
//
//   bench <name> <base.fvecs> <query.fvecs> <k> <results.txt>
//        -> run_benchmark(...)            (cpu_baseline.cpp:177-321; its own timed loop and text output)
//   dump  <base.fvecs> <query.fvecs> <k> <out.bin>
//        -> read_fvecs / compute_norms / cblas_sgemm / select_topk called in the order the
//           reference's hot loop calls them (cpu_baseline.cpp:211-248) and the exact float32
//           distances + int32 ids written in binary (results.txt only keeps 6 significant digits).
//           Layout: int32 nq, int32 k, then nq*k int32 ids, nq*k float32 dists, nq float32 qnorms,
//           then int32 nb and nb float32 base norms.
//
// Only declarations of the reference's symbols appear here; their definitions stay in /root/reference.
#include <cblas.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

struct Result {  // must match the reference's layout/name for linkage (cpu_baseline.cpp:13-19)
    float dist;
    int idx;
    bool operator<(const Result& o) const { return dist < o.dist; }
};
bool read_fvecs(const std::string& filename, std::vector<float>& data, int& rows, int& dim);
void compute_norms(const std::vector<float>& data, std::vector<float>& norms, int rows, int dim);
void select_topk(const float* distances, int N, int k, std::vector<Result>& topk);
void run_benchmark(const std::string& dataset_name, const std::string& base_file,
                   const std::string& query_file, int k, const std::string& output_file);

static int usage() {
    std::fprintf(stderr,
                 "usage: ref_driver bench <name> <base.fvecs> <query.fvecs> <k> <results.txt>\n"
                 "       ref_driver dump <base.fvecs> <query.fvecs> <k> <out.bin>\n"
                 "       ref_driver config\n");
    return 2;
}

int main(int argc, char** argv) {
    if (argc < 2) return usage();
    if (!std::strcmp(argv[1], "config")) {
        std::printf("%s\n", scipy_openblas_get_config());
        return 0;
    }
    if (!std::strcmp(argv[1], "bench")) {
        if (argc != 7) return usage();
        run_benchmark(argv[2], argv[3], argv[4], std::atoi(argv[5]), argv[6]);
        return 0;
    }
    if (!std::strcmp(argv[1], "dump")) {
        if (argc != 6) return usage();
        std::vector<float> Q, B, qn, bn;
        int nq = 0, nb = 0, dq = 0, db = 0;
        if (!read_fvecs(argv[2], B, nb, db) || !read_fvecs(argv[3], Q, nq, dq) || dq != db) return 1;
        const int k = std::atoi(argv[4]);
        compute_norms(Q, qn, nq, dq);
        compute_norms(B, bn, nb, db);
        std::vector<int> ids((size_t)nq * k);
        std::vector<float> dists((size_t)nq * k);
        std::vector<float> row(nb);
        std::vector<Result> top;
        for (int i = 0; i < nq; ++i) {
            cblas_sgemm(CblasRowMajor, CblasNoTrans, CblasTrans, 1, nb, dq, 1.0f, &Q[(size_t)i * dq], dq,
                        B.data(), dq, 0.0f, row.data(), nb);
            for (int j = 0; j < nb; ++j) row[j] = qn[i] + bn[j] - 2.0f * row[j];
            select_topk(row.data(), nb, k, top);
            for (int t = 0; t < k; ++t) {
                ids[(size_t)i * k + t] = top[t].idx;
                dists[(size_t)i * k + t] = top[t].dist;
            }
        }
        FILE* f = std::fopen(argv[5], "wb");
        if (!f) return 1;
        std::fwrite(&nq, 4, 1, f);
        std::fwrite(&k, 4, 1, f);
        std::fwrite(ids.data(), 4, ids.size(), f);
        std::fwrite(dists.data(), 4, dists.size(), f);
        std::fwrite(qn.data(), 4, qn.size(), f);
        std::fwrite(&nb, 4, 1, f);
        std::fwrite(bn.data(), 4, bn.size(), f);
        std::fclose(f);
        return 0;
    }
    return usage();
}
