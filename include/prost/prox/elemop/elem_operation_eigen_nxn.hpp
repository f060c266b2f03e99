// prost/prox/elemop/elem_operation_eigen_nxn.hpp -- ElemOperationEigenNxN<T, FUN_1D>: prox of sum_i h(lambda_i) of a symmetrised n x n matrix, n <= 32 (the reference's N_MAX)
// (reference: elem_operation_eigen_nxn.hpp).
#ifndef PROST_ELEM_OPERATION_EIGEN_NXN_HPP_
#define PROST_ELEM_OPERATION_EIGEN_NXN_HPP_

#include "prost/prox/elemop/elem_operation.hpp"
#include "prost/prox/elemop/function_2d.hpp"

namespace prost {

template <typename T, class FUN_1D>
struct ElemOperationEigenNxN : public ElemOperation<0, 7> {
  static const int kKind = detail::kElemOpSpectral;
  static const int kSpectralKind = PB_SPECTRAL_EIGEN_NXN;
  static const int kFunctionId = FUN_1D::kFunctionId;
  static const int kFunction2D = 0;
};

}  // namespace prost

#endif
