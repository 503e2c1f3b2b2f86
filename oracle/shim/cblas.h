/* TEST INFRASTRUCTURE — not product code.
 *
 * Minimal <cblas.h> stand-in so that the UNMODIFIED reference cpu/cpu_baseline.cpp
 * (which does `#include <cblas.h>` and calls cblas_sgemm at cpu_baseline.cpp:229-237)
 * can be compiled in an image that has no libopenblas-dev.  The symbol is mapped onto
 * the OpenBLAS that ships inside scipy's wheel (LP64, symbols prefixed `scipy_`).
 * Nothing here restates reference code; it only declares the standard CBLAS prototype.
 */
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
typedef enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_ORDER;
typedef enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 } CBLAS_TRANSPOSE;
void scipy_cblas_sgemm(CBLAS_ORDER order, CBLAS_TRANSPOSE ta, CBLAS_TRANSPOSE tb, int M, int N, int K,
                       float alpha, const float* A, int lda, const float* B, int ldb, float beta,
                       float* C, int ldc);
void scipy_openblas_set_num_threads(int n);
char* scipy_openblas_get_config(void);
#define cblas_sgemm scipy_cblas_sgemm
#ifdef __cplusplus
}
#endif
