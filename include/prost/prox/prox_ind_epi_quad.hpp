// prost/prox/prox_ind_epi_quad.hpp -- ProxIndEpiQuad<T>: projection onto a x^T x + b^T x + c <= y
// (reference: include/prost/prox/prox_ind_epi_quad.hpp:42-51, src/prox/prox_ind_epi_quad.cu).
#ifndef PROST_PROX_IND_EPI_QUAD_HPP_
#define PROST_PROX_IND_EPI_QUAD_HPP_

#include "prost/prox/prox_separable_sum.hpp"

namespace prost {

template <typename T>
class ProxIndEpiQuad : public ProxSeparableSum<T> {
 public:
  ProxIndEpiQuad(size_t index, size_t count, size_t dim, bool interleaved, bool diagsteps, std::vector<T> a,
                 std::vector<T> b, std::vector<T> c)
      : ProxSeparableSum<T>(index, count, dim, interleaved, diagsteps), a_(a), b_(b), c_(c) {}

 protected:
  virtual pb_prox* create() {
    pb_prox* h = nullptr;
    detail::check(pb_prox_create_ind_epi_quad(detail::context(), this->index_, this->count_, this->dim_,
                                              this->interleaved_, this->diagsteps_, a_.data(), a_.size(),
                                              b_.data(), b_.size(), c_.data(), c_.size(), &h));
    return h;
  }
  std::vector<T> a_, b_, c_;
};

}  // namespace prost

#endif
