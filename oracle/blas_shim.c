/* TEST INFRASTRUCTURE (oracle build) -- not part of the product path.
 *
 * The reference CPU engine calls BLAS/LAPACK through seven Fortran forwarding
 * shims spral_c_d* (reference: src/ssids/cpu/cpu_iface.f90:100-173, prototypes
 * src/ssids/cpu/kernels/wrappers.cxx:10-18).  No Fortran compiler exists in
 * this image, so the same seven entry points are provided here in C and
 * forwarded to the LP64 OpenBLAS that ships inside scipy.libs (symbols carry a
 * scipy_ prefix; hidden Fortran character lengths are passed last).
 */
#include <stddef.h>

void scipy_dgemm_(char*, char*, int*, int*, int*, double*, const double*, int*,
                  const double*, int*, double*, double*, int*, size_t, size_t);
void scipy_dpotrf_(char*, int*, double*, int*, int*, size_t);
void scipy_dsytrf_(char*, int*, double*, int*, int*, double*, int*, int*, size_t);
void scipy_dtrsm_(char*, char*, char*, char*, int*, int*, const double*,
                  const double*, int*, double*, int*, size_t, size_t, size_t, size_t);
void scipy_dsyrk_(char*, char*, int*, int*, double*, const double*, int*,
                  double*, double*, int*, size_t, size_t);
void scipy_dtrsv_(char*, char*, char*, int*, const double*, int*, double*, int*,
                  size_t, size_t, size_t);
void scipy_dgemv_(char*, int*, int*, const double*, const double*, int*,
                  const double*, int*, const double*, double*, int*, size_t);
void scipy_openblas_set_num_threads(int);

void spral_c_dgemm(char* ta, char* tb, int* m, int* n, int* k, double* alpha,
                   const double* a, int* lda, const double* b, int* ldb,
                   double* beta, double* c, int* ldc) {
   scipy_dgemm_(ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, 1, 1);
}
void spral_c_dpotrf(char* uplo, int* n, double* a, int* lda, int* info) {
   scipy_dpotrf_(uplo, n, a, lda, info, 1);
}
void spral_c_dsytrf(char* uplo, int* n, double* a, int* lda, int* ipiv,
                    double* work, int* lwork, int* info) {
   scipy_dsytrf_(uplo, n, a, lda, ipiv, work, lwork, info, 1);
}
void spral_c_dtrsm(char* side, char* uplo, char* transa, char* diag, int* m,
                   int* n, const double* alpha, const double* a, int* lda,
                   double* b, int* ldb) {
   scipy_dtrsm_(side, uplo, transa, diag, m, n, alpha, a, lda, b, ldb, 1, 1, 1, 1);
}
void spral_c_dsyrk(char* uplo, char* trans, int* n, int* k, double* alpha,
                   const double* a, int* lda, double* beta, double* c, int* ldc) {
   scipy_dsyrk_(uplo, trans, n, k, alpha, a, lda, beta, c, ldc, 1, 1);
}
void spral_c_dtrsv(char* uplo, char* trans, char* diag, int* n, const double* a,
                   int* lda, double* x, int* incx) {
   scipy_dtrsv_(uplo, trans, diag, n, a, lda, x, incx, 1, 1, 1);
}
void spral_c_dgemv(char* trans, int* m, int* n, const double* alpha,
                   const double* a, int* lda, const double* x, int* incx,
                   const double* beta, double* y, int* incy) {
   scipy_dgemv_(trans, m, n, alpha, a, lda, x, incx, beta, y, incy, 1);
}

/* Task parallelism comes from SSIDS' OpenMP tasks; keep BLAS serial. */
void oracle_blas_single_thread(void) { scipy_openblas_set_num_threads(1); }
/* the solves of the reference are a serial loop over the nodes: their parallelism is the BLAS library's */
void oracle_blas_set_threads(int n) { scipy_openblas_set_num_threads(n < 1 ? 1 : n); }
