// prost/prox/prox_permute.hpp -- ProxPermute<T>: gather, inner prox, scatter
// (reference: include/prost/prox/prox_permute.hpp:37, src/prox/prox_permute.cu:101-145).
#ifndef PROST_PROX_PERMUTE_HPP_
#define PROST_PROX_PERMUTE_HPP_

#include "prost/prox/prox.hpp"

namespace prost {

template <typename T>
class ProxPermute : public Prox<T> {
 public:
  ProxPermute(std::shared_ptr<Prox<T> > base_prox, const std::vector<int>& perm)
      : Prox<T>(*base_prox), base_prox_(base_prox), perm_(perm) {}

 protected:
  virtual pb_prox* create() {
    pb_prox* h = nullptr;
    detail::check(pb_prox_create_permute(detail::context(), base_prox_->handle(), perm_.data(), perm_.size(), &h));
    return h;
  }
  std::shared_ptr<Prox<T> > base_prox_;
  std::vector<int> perm_;
};

}  // namespace prost

#endif
