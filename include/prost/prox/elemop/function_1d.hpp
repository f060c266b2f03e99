// prost/prox/elemop/function_1d.hpp -- the Function1D family as compile-time tags
// (reference: include/prost/prox/elemop/function_1d.hpp:34-326).  The arithmetic lives in
// prost_b200/csrc/pb_math.cuh; the tag only carries the pb_function1d id.
#ifndef PROST_FUNCTION_1D_HPP_
#define PROST_FUNCTION_1D_HPP_

#include "prost/common.hpp"

namespace prost {

#define PROST_FUNCTION_1D_TAG(NAME, ID) \
  template <typename T> struct NAME { static const int kFunctionId = ID; }

PROST_FUNCTION_1D_TAG(Function1DZero, PB_FUN_ZERO);
PROST_FUNCTION_1D_TAG(Function1DAbs, PB_FUN_ABS);
PROST_FUNCTION_1D_TAG(Function1DSquare, PB_FUN_SQUARE);
PROST_FUNCTION_1D_TAG(Function1DIndLeq0, PB_FUN_IND_LEQ0);
PROST_FUNCTION_1D_TAG(Function1DIndGeq0, PB_FUN_IND_GEQ0);
PROST_FUNCTION_1D_TAG(Function1DIndEq0, PB_FUN_IND_EQ0);
PROST_FUNCTION_1D_TAG(Function1DIndBox01, PB_FUN_IND_BOX01);
PROST_FUNCTION_1D_TAG(Function1DMaxPos0, PB_FUN_MAX_POS0);
PROST_FUNCTION_1D_TAG(Function1DL0, PB_FUN_L0);
PROST_FUNCTION_1D_TAG(Function1DHuber, PB_FUN_HUBER);
PROST_FUNCTION_1D_TAG(Function1DLq, PB_FUN_LQ);
PROST_FUNCTION_1D_TAG(Function1DLqPlusEps, PB_FUN_LQ_PLUS_EPS);
PROST_FUNCTION_1D_TAG(Function1DTruncQuad, PB_FUN_TRUNC_QUAD);
PROST_FUNCTION_1D_TAG(Function1DTruncLinear, PB_FUN_TRUNC_LINEAR);

#undef PROST_FUNCTION_1D_TAG

}  // namespace prost

#endif
