// prost/prox/prox_elem_operation.hpp -- ProxElemOperation<T, ELEM_OPERATION>
// (reference: include/prost/prox/prox_elem_operation.hpp:33-102).
#ifndef PROST_PROX_ELEM_OPERATION_HPP_
#define PROST_PROX_ELEM_OPERATION_HPP_

#include <array>

#include "prost/prox/elemop/elem_operation.hpp"
#include "prost/prox/prox_separable_sum.hpp"

namespace prost {

namespace detail {
// kSpectralKind / kFunction2D of the spectral tags, 0 for every other element operation (SFINAE on the member)
template <class OP> int spectral_kind(decltype(OP::kSpectralKind)*) { return OP::kSpectralKind; }
template <class OP> int spectral_kind(...) { return 0; }
template <class OP> int function_2d(decltype(OP::kFunction2D)*) { return OP::kFunction2D; }
template <class OP> int function_2d(...) { return 0; }
}  // namespace detail

template <typename T, class ELEM_OPERATION>
class ProxElemOperation : public ProxSeparableSum<T> {
 public:
  /// Operations without coefficients (ElemOperationIndSimplex, ElemOperationIndSum).
  ProxElemOperation(size_t index, size_t count, size_t dim, bool interleaved, bool diagsteps)
      : ProxSeparableSum<T>(index, count, ELEM_OPERATION::kDim <= 0 ? dim : ELEM_OPERATION::kDim, interleaved,
                            diagsteps) {
    static_assert(ELEM_OPERATION::kCoeffsCount == 0, "this element operation needs its coefficients");
  }

  /// Operations with coefficient arrays a,b,c,d,e,alpha,beta; each of length 1 or count.
  ProxElemOperation(size_t index, size_t count, size_t dim, bool interleaved, bool diagsteps,
                    std::array<std::vector<T>, ELEM_OPERATION::kCoeffsCount> coeffs)
      : ProxSeparableSum<T>(index, count, ELEM_OPERATION::kDim <= 0 ? dim : ELEM_OPERATION::kDim, interleaved,
                            diagsteps),
        coeffs_(coeffs.begin(), coeffs.end()) {}

 protected:
  virtual pb_prox* create() {
    pb_prox* h = nullptr;
    pb_context* ctx = detail::context();
    if (ELEM_OPERATION::kKind == detail::kElemOpIndSimplex) {
      detail::check(pb_prox_create_ind_simplex(ctx, this->index_, this->count_, this->dim_, this->interleaved_,
                                               this->diagsteps_, &h));
      return h;
    }
    if (ELEM_OPERATION::kKind == detail::kElemOpIndSum) {
      detail::check(pb_prox_create_ind_sum(ctx, this->index_, this->count_, this->dim_, this->interleaved_,
                                           this->diagsteps_, &h));
      return h;
    }
    if (ELEM_OPERATION::kKind == detail::kElemOpSpectral && coeffs_.size() < 7) {
      // mass / comass norms: zero or one coefficient array (the cost); the ABI takes the usual seven
      static const float defaults[7] = {1.f, 0.f, 1.f, 0.f, 0.f, 0.f, 0.f};
      std::vector<std::vector<T> > padded(7);
      for (int k = 0; k < 7; ++k) padded[k].assign(1, defaults[k]);
      for (size_t k = 0; k < coeffs_.size(); ++k) padded[k] = coeffs_[k];
      coeffs_.swap(padded);
    }
    if (coeffs_.size() != 7) throw Exception("ProxElemOperation: expected 7 coefficient arrays.");
    const float* ptrs[7];
    size_t lens[7];
    for (int k = 0; k < 7; ++k) { ptrs[k] = coeffs_[k].data(); lens[k] = coeffs_[k].size(); }
    if (ELEM_OPERATION::kKind == detail::kElemOpSpectral)
      detail::check(pb_prox_create_spectral(ctx, detail::spectral_kind<ELEM_OPERATION>(0), this->index_, this->count_,
                                            this->dim_, this->interleaved_, this->diagsteps_,
                                            ELEM_OPERATION::kFunctionId, detail::function_2d<ELEM_OPERATION>(0), ptrs,
                                            lens, &h));
    else if (ELEM_OPERATION::kKind == detail::kElemOp1D)
      detail::check(pb_prox_create_elem_1d(ctx, this->index_, this->count_, this->dim_, this->interleaved_,
                                           this->diagsteps_, ELEM_OPERATION::kFunctionId, ptrs, lens, &h));
    else
      detail::check(pb_prox_create_elem_norm2(ctx, this->index_, this->count_, this->dim_, this->interleaved_,
                                              this->diagsteps_, ELEM_OPERATION::kFunctionId, ptrs, lens, &h));
    return h;
  }
  std::vector<std::vector<T> > coeffs_;
};

}  // namespace prost

#endif
