// prost/prox/elemop/elem_operation.hpp -- tags describing element operations
// (reference: include/prost/prox/elemop/elem_operation.hpp:28-40).
#ifndef PROST_ELEM_OPERATION_HPP_
#define PROST_ELEM_OPERATION_HPP_

#include "prost/common.hpp"

namespace prost {

namespace detail {
enum ElemOpKind { kElemOp1D, kElemOpNorm2, kElemOpIndSimplex, kElemOpIndSum, kElemOpSpectral };
}

/// DIM = 0 means "taken from the constructor"; COEFFS_COUNT = number of per-element coefficient arrays.
template <size_t DIM = 0, size_t COEFFS_COUNT = 0>
struct ElemOperation {
  static const size_t kCoeffsCount = COEFFS_COUNT;
  static const size_t kDim = DIM;
};

}  // namespace prost

#endif
