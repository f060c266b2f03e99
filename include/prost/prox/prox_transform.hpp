// prost/prox/prox_transform.hpp -- ProxTransform<T>: prox of  c f(a x - b) + <d, x> + (e/2)|x|^2  through the
// prox of f (reference: include/prost/prox/prox_transform.hpp:38-44, src/prox/prox_transform.cu:27-226).
#ifndef PROST_PROX_TRANSFORM_HPP_
#define PROST_PROX_TRANSFORM_HPP_

#include <vector>

#include "prost/prox/prox.hpp"

namespace prost {

template <typename T>
class ProxTransform : public Prox<T> {
 public:
  ProxTransform(std::shared_ptr<Prox<T> > inner_fn, const std::vector<T>& a, const std::vector<T>& b,
                const std::vector<T>& c, const std::vector<T>& d, const std::vector<T>& e)
      : Prox<T>(*inner_fn), inner_fn_(inner_fn) {
    const std::vector<T>* src[5] = {&a, &b, &c, &d, &e};
    for (int k = 0; k < 5; ++k) coeffs_[k].assign(src[k]->begin(), src[k]->end());
  }

 protected:
  virtual pb_prox* create() {
    const float* ptrs[5];
    size_t lens[5];
    for (int k = 0; k < 5; ++k) {
      ptrs[k] = coeffs_[k].data();
      lens[k] = coeffs_[k].size();
    }
    pb_prox* h = nullptr;
    detail::check(pb_prox_create_transform(detail::context(), inner_fn_->handle(), ptrs, lens, &h));
    return h;
  }
  std::shared_ptr<Prox<T> > inner_fn_;
  std::vector<float> coeffs_[5];
};

}  // namespace prost

#endif
