/* Minimal cblas.h for building the unmodified reference (src/matrix.cpp:5,112)
 * against the OpenBLAS shared object bundled in the image's opencv wheel, which
 * ships no headers.  Test infrastructure only (oracle/_ref). */
#ifndef GVC_SHIM_CBLAS_H
#define GVC_SHIM_CBLAS_H
#ifdef __cplusplus
extern "C" {
#endif
enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 };
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 };
void cblas_sgemm(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE ta, enum CBLAS_TRANSPOSE tb,
                 int M, int N, int K, float alpha, const float *A, int lda,
                 const float *B, int ldb, float beta, float *C, int ldc);
void openblas_set_num_threads(int n);
int openblas_get_num_threads(void);
char *openblas_get_config(void);
char *openblas_get_corename(void);
#ifdef __cplusplus
}
#endif
#endif
