// prost/prox/elemop/elem_operation_ind_sum.hpp -- ElemOperationIndSum<T>: projection of every group onto the
// sum-to-one constraint (reference: elem_operation_ind_sum.hpp:38-58).
#ifndef PROST_ELEM_OPERATION_IND_SUM_HPP_
#define PROST_ELEM_OPERATION_IND_SUM_HPP_

#include "prost/prox/elemop/elem_operation.hpp"

namespace prost {

template <typename T>
struct ElemOperationIndSum : public ElemOperation<0, 0> {
  static const int kKind = detail::kElemOpIndSum;
  static const int kFunctionId = 0;
};

}  // namespace prost

#endif
