// prost/prox/elemop/elem_operation_eigen_3x3.hpp -- ElemOperationEigen3x3<T, FUN_1D>: prox of sum_i h(lambda_i) of a symmetrised 3 x 3 matrix
// (reference: elem_operation_eigen_3x3.hpp:302-377).
#ifndef PROST_ELEM_OPERATION_EIGEN_3X3_HPP_
#define PROST_ELEM_OPERATION_EIGEN_3X3_HPP_

#include "prost/prox/elemop/elem_operation.hpp"
#include "prost/prox/elemop/function_2d.hpp"

namespace prost {

template <typename T, class FUN_1D>
struct ElemOperationEigen3x3 : public ElemOperation<0, 7> {
  static const int kKind = detail::kElemOpSpectral;
  static const int kSpectralKind = PB_SPECTRAL_EIGEN_3X3;
  static const int kFunctionId = FUN_1D::kFunctionId;
  static const int kFunction2D = 0;
};

}  // namespace prost

#endif
