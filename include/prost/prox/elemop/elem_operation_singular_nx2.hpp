// prost/prox/elemop/elem_operation_singular_nx2.hpp -- ElemOperationSingularNx2<T, FUN_2D>: prox of a Function2D of the singular values of an N x 2 matrix
// (reference: elem_operation_singular_nx2.hpp:32-150).
#ifndef PROST_ELEM_OPERATION_SINGULAR_NX2_HPP_
#define PROST_ELEM_OPERATION_SINGULAR_NX2_HPP_

#include "prost/prox/elemop/elem_operation.hpp"
#include "prost/prox/elemop/function_2d.hpp"

namespace prost {

template <typename T, class FUN_2D>
struct ElemOperationSingularNx2 : public ElemOperation<0, 7> {
  static const int kKind = detail::kElemOpSpectral;
  static const int kSpectralKind = PB_SPECTRAL_SINGULAR_NX2;
  static const int kFunctionId = FUN_2D::kFunctionId;
  static const int kFunction2D = FUN_2D::kFunction2D;
};

}  // namespace prost

#endif
