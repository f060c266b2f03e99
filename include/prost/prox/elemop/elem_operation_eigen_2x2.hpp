// prost/prox/elemop/elem_operation_eigen_2x2.hpp -- ElemOperationEigen2x2<T, FUN_1D>: prox of sum_i h(lambda_i) of a symmetrised 2 x 2 matrix
// (reference: elem_operation_eigen_2x2.hpp:94-146).
#ifndef PROST_ELEM_OPERATION_EIGEN_2X2_HPP_
#define PROST_ELEM_OPERATION_EIGEN_2X2_HPP_

#include "prost/prox/elemop/elem_operation.hpp"
#include "prost/prox/elemop/function_2d.hpp"

namespace prost {

template <typename T, class FUN_1D>
struct ElemOperationEigen2x2 : public ElemOperation<0, 7> {
  static const int kKind = detail::kElemOpSpectral;
  static const int kSpectralKind = PB_SPECTRAL_EIGEN_2X2;
  static const int kFunctionId = FUN_1D::kFunctionId;
  static const int kFunction2D = 0;
};

}  // namespace prost

#endif
