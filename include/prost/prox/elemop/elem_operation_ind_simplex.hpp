// prost/prox/elemop/elem_operation_ind_simplex.hpp -- ElemOperationIndSimplex<T>:
// projection onto the unit simplex per group (reference: elem_operation_ind_simplex.hpp:47-115).
#ifndef PROST_ELEM_OPERATION_SIMPLEX_HPP_
#define PROST_ELEM_OPERATION_SIMPLEX_HPP_

#include "prost/prox/elemop/elem_operation.hpp"

namespace prost {

template <typename T>
struct ElemOperationIndSimplex : public ElemOperation<0, 0> {
  static const int kKind = detail::kElemOpIndSimplex;
  static const int kFunctionId = 0;
};

}  // namespace prost

#endif
