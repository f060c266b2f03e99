// prost/prox/prox_moreau.hpp -- ProxMoreau<T>: prox of the conjugate through Moreau's identity
// (reference: include/prost/prox/prox_moreau.hpp:37, src/prox/prox_moreau.cu:98-134).
#ifndef PROST_PROX_MOREAU_HPP_
#define PROST_PROX_MOREAU_HPP_

#include "prost/prox/prox.hpp"

namespace prost {

template <typename T>
class ProxMoreau : public Prox<T> {
 public:
  explicit ProxMoreau(std::shared_ptr<Prox<T> > conjugate) : Prox<T>(*conjugate), conjugate_(conjugate) {}

 protected:
  virtual pb_prox* create() {
    pb_prox* h = nullptr;
    detail::check(pb_prox_create_moreau(detail::context(), conjugate_->handle(), &h));
    return h;
  }
  std::shared_ptr<Prox<T> > conjugate_;
};

}  // namespace prost

#endif
