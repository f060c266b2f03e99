// prost/prox/prox_ind_halfspace.hpp -- ProxIndHalfspace<T>: projection onto a^T x <= b per group
// (reference: include/prost/prox/prox_ind_halfspace.hpp:41-52, src/prox/prox_ind_halfspace.cu).
#ifndef PROST_PROX_IND_HALFSPACE_HPP_
#define PROST_PROX_IND_HALFSPACE_HPP_

#include "prost/prox/prox_separable_sum.hpp"

namespace prost {

template <typename T>
class ProxIndHalfspace : public ProxSeparableSum<T> {
 public:
  ProxIndHalfspace(size_t index, size_t count, size_t dim, bool interleaved, bool diagsteps, std::vector<T> a,
                   std::vector<T> b)
      : ProxSeparableSum<T>(index, count, dim, interleaved, diagsteps), a_(a.begin(), a.end()),
        b_(b.begin(), b.end()) {}

 protected:
  virtual pb_prox* create() {
    pb_prox* h = nullptr;
    detail::check(pb_prox_create_ind_halfspace(detail::context(), this->index_, this->count_, this->dim_,
                                               this->interleaved_, this->diagsteps_, a_.data(), a_.size(), b_.data(),
                                               b_.size(), &h));
    return h;
  }
  std::vector<float> a_, b_;
};

}  // namespace prost

#endif
