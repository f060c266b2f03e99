// oracle/ref_build/cusparse_csrmv_shim.h  --  TEST INFRASTRUCTURE, not product code.
//
// Force-included (nvcc -include) when compiling the two reference files that call
// the legacy cuSPARSE csrmv entry points removed in CUDA 11
// (/root/reference/src/linop/block_sparse.cu:156,190,224,258 and
//  /root/reference/src/prox/prox_ind_range.cu:210,236,266,292).
// It re-creates the two removed names on top of the generic cusparseSpMV API so the
// reference sources compile untouched.  Only result parity matters here, not speed.
#pragma once
#include <cuda_runtime.h>
#include <cusparse.h>

namespace pb_ref_shim {

template <typename T> struct cuda_type;
template <> struct cuda_type<float>  { static constexpr cudaDataType v = CUDA_R_32F; };
template <> struct cuda_type<double> { static constexpr cudaDataType v = CUDA_R_64F; };

template <typename T>
inline cusparseStatus_t legacy_csrmv(cusparseHandle_t h, cusparseOperation_t op,
                                     int m, int n, int nnz, const T* alpha,
                                     const T* val, const int* rowptr, const int* colind,
                                     const T* x, const T* beta, T* y)
{
  cusparseSpMatDescr_t A = nullptr;
  cusparseDnVecDescr_t vx = nullptr, vy = nullptr;
  const int xlen = (op == CUSPARSE_OPERATION_NON_TRANSPOSE) ? n : m;
  const int ylen = (op == CUSPARSE_OPERATION_NON_TRANSPOSE) ? m : n;
  cusparseStatus_t st;
  st = cusparseCreateCsr(&A, m, n, nnz, const_cast<int*>(rowptr), const_cast<int*>(colind),
                         const_cast<T*>(val), CUSPARSE_INDEX_32I, CUSPARSE_INDEX_32I,
                         CUSPARSE_INDEX_BASE_ZERO, cuda_type<T>::v);
  if (st != CUSPARSE_STATUS_SUCCESS) return st;
  st = cusparseCreateDnVec(&vx, xlen, const_cast<T*>(x), cuda_type<T>::v);
  if (st != CUSPARSE_STATUS_SUCCESS) { cusparseDestroySpMat(A); return st; }
  st = cusparseCreateDnVec(&vy, ylen, y, cuda_type<T>::v);
  if (st != CUSPARSE_STATUS_SUCCESS) { cusparseDestroyDnVec(vx); cusparseDestroySpMat(A); return st; }
  size_t ws = 0;
  void* buf = nullptr;
  st = cusparseSpMV_bufferSize(h, op, alpha, A, vx, beta, vy, cuda_type<T>::v,
                               CUSPARSE_SPMV_ALG_DEFAULT, &ws);
  if (st == CUSPARSE_STATUS_SUCCESS) {
    if (ws > 0) cudaMalloc(&buf, ws);
    st = cusparseSpMV(h, op, alpha, A, vx, beta, vy, cuda_type<T>::v,
                      CUSPARSE_SPMV_ALG_DEFAULT, buf);
    if (buf) cudaFree(buf);
  }
  cusparseDestroyDnVec(vy);
  cusparseDestroyDnVec(vx);
  cusparseDestroySpMat(A);
  return st;
}

}  // namespace pb_ref_shim

inline cusparseStatus_t cusparseScsrmv(cusparseHandle_t h, cusparseOperation_t op, int m, int n,
                                       int nnz, const float* alpha, cusparseMatDescr_t,
                                       const float* val, const int* rowptr, const int* colind,
                                       const float* x, const float* beta, float* y)
{ return pb_ref_shim::legacy_csrmv<float>(h, op, m, n, nnz, alpha, val, rowptr, colind, x, beta, y); }

inline cusparseStatus_t cusparseDcsrmv(cusparseHandle_t h, cusparseOperation_t op, int m, int n,
                                       int nnz, const double* alpha, cusparseMatDescr_t,
                                       const double* val, const int* rowptr, const int* colind,
                                       const double* x, const double* beta, double* y)
{ return pb_ref_shim::legacy_csrmv<double>(h, op, m, n, nnz, alpha, val, rowptr, colind, x, beta, y); }
