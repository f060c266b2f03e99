// prost/prox/prox_zero.hpp -- ProxZero<T>: identity prox (reference: include/prost/prox/prox_zero.hpp:34).
#ifndef PROST_PROX_ZERO_HPP_
#define PROST_PROX_ZERO_HPP_

#include "prost/prox/prox.hpp"

namespace prost {

template <typename T>
class ProxZero : public Prox<T> {
 public:
  ProxZero(size_t index, size_t size) : Prox<T>(index, size, true) {}

 protected:
  virtual pb_prox* create() {
    pb_prox* h = nullptr;
    detail::check(pb_prox_create_zero(detail::context(), this->index_, this->size_, &h));
    return h;
  }
};

}  // namespace prost

#endif
