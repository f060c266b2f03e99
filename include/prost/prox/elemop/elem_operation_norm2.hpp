// prost/prox/elemop/elem_operation_norm2.hpp -- ElemOperationNorm2<T, FUN_1D>:
// prox of h(|x|_2) over dim-vectors (reference: elem_operation_norm2.hpp:39-88).
#ifndef PROST_ELEM_OPERATION_NORM2_HPP_
#define PROST_ELEM_OPERATION_NORM2_HPP_

#include "prost/prox/elemop/elem_operation.hpp"

namespace prost {

template <typename T, class FUN_1D>
struct ElemOperationNorm2 : public ElemOperation<0, 7> {
  static const int kKind = detail::kElemOpNorm2;
  static const int kFunctionId = FUN_1D::kFunctionId;
};

}  // namespace prost

#endif
