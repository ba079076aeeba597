/* Stand-in for Intel MKL's mkl_spblas.h -- ORACLE BUILD ONLY; see mkl_types.h in this directory. */
#ifndef RCHOL_B200_MKLSHIM_SPBLAS_H
#define RCHOL_B200_MKLSHIM_SPBLAS_H
#include "mkl_types.h"
#ifdef __cplusplus
extern "C" {
#endif
/* numeric values are those of oneMKL's public enums (verified against the 3x3 KAT) */
typedef enum { SPARSE_STATUS_SUCCESS = 0 } sparse_status_t;
typedef enum { SPARSE_INDEX_BASE_ZERO = 0, SPARSE_INDEX_BASE_ONE = 1 } sparse_index_base_t;
typedef enum { SPARSE_OPERATION_NON_TRANSPOSE = 10, SPARSE_OPERATION_TRANSPOSE = 11 } sparse_operation_t;
typedef enum { SPARSE_MATRIX_TYPE_GENERAL = 20, SPARSE_MATRIX_TYPE_SYMMETRIC = 21,
               SPARSE_MATRIX_TYPE_TRIANGULAR = 23 } sparse_matrix_type_t;
typedef enum { SPARSE_FILL_MODE_LOWER = 40, SPARSE_FILL_MODE_UPPER = 41 } sparse_fill_mode_t;
typedef enum { SPARSE_DIAG_NON_UNIT = 50, SPARSE_DIAG_UNIT = 51 } sparse_diag_type_t;
struct matrix_descr {
  sparse_matrix_type_t type;
  sparse_fill_mode_t mode;
  sparse_diag_type_t diag;
};
struct rchol_b200_ilp64_handle;
typedef struct rchol_b200_ilp64_handle *sparse_matrix_t;

sparse_status_t rchol_b200_mkl_create_csr(sparse_matrix_t *A, sparse_index_base_t indexing, MKL_INT rows,
                                          MKL_INT cols, MKL_INT *rows_start, MKL_INT *rows_end,
                                          MKL_INT *col_indx, double *values);
sparse_status_t rchol_b200_mkl_mv(sparse_operation_t op, double alpha, const sparse_matrix_t A,
                                  struct matrix_descr descr, const double *x, double beta, double *y);
sparse_status_t rchol_b200_mkl_trsv(sparse_operation_t op, double alpha, const sparse_matrix_t A,
                                    struct matrix_descr descr, const double *x, double *y);
sparse_status_t rchol_b200_mkl_destroy(sparse_matrix_t A);
#define mkl_sparse_d_create_csr rchol_b200_mkl_create_csr
#define mkl_sparse_d_mv rchol_b200_mkl_mv
#define mkl_sparse_d_trsv rchol_b200_mkl_trsv
#define mkl_sparse_destroy rchol_b200_mkl_destroy
#ifdef __cplusplus
}
#endif
#endif
