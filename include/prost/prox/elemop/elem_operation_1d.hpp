// prost/prox/elemop/elem_operation_1d.hpp -- ElemOperation1D<T, FUN_1D>:
// prox of c*f(ax - b) + dx + (e/2)x^2 per element (reference: elem_operation_1d.hpp:36-59).
#ifndef PROST_ELEM_OPERATION_1D_HPP_
#define PROST_ELEM_OPERATION_1D_HPP_

#include "prost/prox/elemop/elem_operation.hpp"

namespace prost {

template <typename T, class FUN_1D>
struct ElemOperation1D : public ElemOperation<1, 7> {
  static const int kKind = detail::kElemOp1D;
  static const int kFunctionId = FUN_1D::kFunctionId;
};

}  // namespace prost

#endif
