// prost/prox/elemop/elem_operation_mass_norm.hpp -- ElemOperationMass4 / Mass5<T, conjugate>: prox of the mass norm
// of a 2-vector in R^4 (dim 6, one coefficient: the cost) / R^5 (dim 10, no coefficients), or projection onto the
// unit ball of the comass norm (conjugate = true) (reference: elem_operation_mass_norm.hpp:17-186).
#ifndef PROST_ELEM_OPERATION_MASS_NORM_HPP_
#define PROST_ELEM_OPERATION_MASS_NORM_HPP_

#include "prost/prox/elemop/elem_operation.hpp"

namespace prost {

template <typename T, bool conjugate>
struct ElemOperationMass4 : public ElemOperation<6, 1> {
  static const int kKind = detail::kElemOpSpectral;
  static const int kSpectralKind = conjugate ? PB_SPECTRAL_COMASS4_BALL : PB_SPECTRAL_MASS4;
  static const int kFunctionId = PB_FUN_ZERO;
  static const int kFunction2D = 0;
};

template <typename T, bool conjugate>
struct ElemOperationMass5 : public ElemOperation<10, 0> {
  static const int kKind = detail::kElemOpSpectral;
  static const int kSpectralKind = conjugate ? PB_SPECTRAL_COMASS5_BALL : PB_SPECTRAL_MASS5;
  static const int kFunctionId = PB_FUN_ZERO;
  static const int kFunction2D = 0;
};

}  // namespace prost

#endif
