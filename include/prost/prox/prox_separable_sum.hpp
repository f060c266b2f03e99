// prost/prox/prox_separable_sum.hpp -- ProxSeparableSum<T>
// (reference: include/prost/prox/prox_separable_sum.hpp:50-77).
#ifndef PROST_PROX_SEPARABLE_SUM_HPP_
#define PROST_PROX_SEPARABLE_SUM_HPP_

#include "prost/prox/prox.hpp"

namespace prost {

/// Sum of `count` functions of `dim` variables each; interleaved: x0 y0 x1 y1 ..., planar: x0 x1 .. y0 y1 ..
template <typename T>
class ProxSeparableSum : public Prox<T> {
 public:
  ProxSeparableSum(size_t index, size_t count, size_t dim, bool interleaved, bool diagsteps)
      : Prox<T>(index, count * dim, diagsteps), count_(count), dim_(dim), interleaved_(interleaved) {}
  size_t dim() const { return dim_; }
  size_t count() const { return count_; }
  bool interleaved() const { return interleaved_; }

 protected:
  size_t count_, dim_;
  bool interleaved_;
};

}  // namespace prost

#endif
