// prost/prox/prox_ind_soc.hpp -- ProxIndSOC<T>: projection onto the second-order cone alpha |x|_2 <= y
// (reference: include/prost/prox/prox_ind_soc.hpp:39-48, src/prox/prox_ind_soc.cu; alpha = 1 only).
#ifndef PROST_PROX_IND_SOC_HPP_
#define PROST_PROX_IND_SOC_HPP_

#include "prost/prox/prox_separable_sum.hpp"

namespace prost {

template <typename T>
class ProxIndSOC : public ProxSeparableSum<T> {
 public:
  ProxIndSOC(size_t index, size_t count, size_t dim, bool interleaved, bool diagsteps, T alpha)
      : ProxSeparableSum<T>(index, count, dim, interleaved, diagsteps), alpha_(alpha) {}

 protected:
  virtual pb_prox* create() {
    pb_prox* h = nullptr;
    detail::check(pb_prox_create_ind_soc(detail::context(), this->index_, this->count_, this->dim_,
                                         this->interleaved_, this->diagsteps_, static_cast<float>(alpha_), &h));
    return h;
  }
  T alpha_;
};

}  // namespace prost

#endif
