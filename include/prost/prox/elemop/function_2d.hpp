// prost/prox/elemop/function_2d.hpp -- the Function2D family as compile-time tags
// (reference: include/prost/prox/elemop/function_2d.hpp:28-101); arithmetic in prost_b200/csrc/pb_spectral.cu.
#ifndef PROST_FUNCTION_2D_HPP_
#define PROST_FUNCTION_2D_HPP_

#include "prost/prox/elemop/function_1d.hpp"

namespace prost {

template <typename T, class FUN_1D>
struct Function2DSum1D {
  static const int kFunction2D = 0;
  static const int kFunctionId = FUN_1D::kFunctionId;
};

template <typename T>
struct Function2DIndL1Ball {
  static const int kFunction2D = 1;
  static const int kFunctionId = PB_FUN_ZERO;
};

template <typename T, class OTHER_FUN_2D>
struct Function2DMoreau {          // only Function2DMoreau<T, Function2DIndL1Ball<T>> is registered (factory.cpp:60)
  static const int kFunction2D = 2;
  static const int kFunctionId = PB_FUN_ZERO;
};

}  // namespace prost

#endif
