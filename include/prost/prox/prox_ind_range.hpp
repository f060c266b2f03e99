// prost/prox/prox_ind_range.hpp -- ProxIndRange<T>: projection onto the range of a sparse matrix A,
// x = A (A^T A)^{-1} A^T x0 (reference: include/prost/prox/prox_ind_range.hpp:37-50, src/prox/prox_ind_range.cu).
#ifndef PROST_PROX_IND_RANGE_HPP_
#define PROST_PROX_IND_RANGE_HPP_

#include <stdint.h>

#include "prost/prox/prox.hpp"

namespace prost {

template <typename T>
class ProxIndRange : public Prox<T> {
 public:
  ProxIndRange(size_t index, size_t size, bool diagsteps) : Prox<T>(index, size, diagsteps), m_(0), n_(0), nnz_(0) {}

  // call before Initialize (CSC of A, like BlockSparse::CreateFromCSC)
  void setA(int m, int n, int nnz, const std::vector<T>& val, const std::vector<int32_t>& ptr,
            const std::vector<int32_t>& ind) {
    m_ = m; n_ = n; nnz_ = nnz;
    val_.assign(val.begin(), val.end());
    ptr_ = ptr;
    ind_ = ind;
  }
  void setAA(int m, int n, const std::vector<T>& val) {
    if (m != n) throw Exception("ProxIndRange: Matrix 'AA' must be square!");
    if (m != n_) throw Exception("ProxIndRange: Matrix 'AA' must fit dimension of 'A'!");
    aa_.assign(val.begin(), val.end());
  }

 protected:
  virtual pb_prox* create() {
    pb_prox* h = nullptr;
    detail::check(pb_prox_create_ind_range(detail::context(), this->index_, this->size_, this->diagsteps_, m_, n_, nnz_,
                                           val_.data(), ptr_.data(), ind_.data(), aa_.data(), &h));
    return h;
  }
  int m_, n_, nnz_;
  std::vector<float> val_, aa_;
  std::vector<int32_t> ptr_, ind_;
};

}  // namespace prost

#endif
