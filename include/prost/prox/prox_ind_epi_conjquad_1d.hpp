// prost/prox/prox_ind_epi_conjquad_1d.hpp -- ProxIndEpiConjQuad1D<T>: projection of (x, y) pairs onto the epigraph
// of the conjugate of a u^2 + b u + c restricted to [alpha, beta] (sublabel-accurate lifting, CVPR 2016).
// The reference keeps this class outside its tree (cmake/CustomSources.cmake.example:8-14:
// ../../preciserelaxation/src/cvpr2016/prost/prox_ind_epi_conjquad_1d.hpp); the constructor below follows the
// coefficient-struct convention of that family (count pairs, one coefficient vector per name).  Parity unpinned.
#ifndef PROST_PROX_IND_EPI_CONJQUAD_1D_HPP_
#define PROST_PROX_IND_EPI_CONJQUAD_1D_HPP_

#include "prost/prox/prox_separable_sum.hpp"

namespace prost {

template <typename T>
struct EpiConjQuadCoeffs {
  std::vector<T> a, b, c, alpha, beta;
};

template <typename T>
class ProxIndEpiConjQuad1D : public ProxSeparableSum<T> {
 public:
  ProxIndEpiConjQuad1D(size_t index, size_t count, bool interleaved, const EpiConjQuadCoeffs<T>& coeffs)
      : ProxSeparableSum<T>(index, count, 2, interleaved, false) {
    const std::vector<T>* src[5] = {&coeffs.a, &coeffs.b, &coeffs.c, &coeffs.alpha, &coeffs.beta};
    for (int k = 0; k < 5; ++k) co_[k].assign(src[k]->begin(), src[k]->end());
  }

 protected:
  virtual pb_prox* create() {
    const float* ptrs[5];
    size_t lens[5];
    for (int k = 0; k < 5; ++k) { ptrs[k] = co_[k].data(); lens[k] = co_[k].size(); }
    pb_prox* h = nullptr;
    detail::check(pb_prox_create_ind_epi_conjquad_1d(detail::context(), this->index_, this->count_, this->interleaved_,
                                                     this->diagsteps_, ptrs, lens, &h));
    return h;
  }
  std::vector<float> co_[5];
};

}  // namespace prost

#endif
