// prost/prox/prox.hpp -- Prox<T>: base class of all proximal operators
// (reference: include/prost/prox/prox.hpp:39-135).
#ifndef PROST_PROX_HPP_
#define PROST_PROX_HPP_

#include "prost/common.hpp"

namespace prost {

template <typename T>
class Prox : detail::require_float<T> {
 public:
  Prox(size_t index, size_t size, bool diagsteps)
      : index_(index), size_(size), diagsteps_(diagsteps), handle_(nullptr) {}
  Prox(const Prox<T>& other)
      : index_(other.index_), size_(other.size_), diagsteps_(other.diagsteps_), handle_(nullptr) {}
  virtual ~Prox() { if (handle_) pb_prox_destroy(handle_); }

  virtual void Initialize() { handle(); }
  virtual void Release() {}

  /// Host-vector evaluation (prox.cu:45-71): result = prox_{tau * diag(tau_diag) f}(arg) on the index
  /// range of this prox; returns device milliseconds.
  double Eval(std::vector<T>& result, const std::vector<T>& arg, const std::vector<T>& tau_diag, T tau) {
    if (arg.size() != tau_diag.size()) throw Exception("Prox::Eval: arg and tau_diag differ in size.");
    result.resize(arg.size());
    double ms = 0;
    detail::check(pb_prox_eval_host(handle(), result.data(), arg.data(), tau_diag.data(), arg.size(), tau, 0, &ms));
    return ms;
  }

  virtual size_t gpu_mem_amount() const { return pb_prox_gpu_mem_amount(const_cast<Prox*>(this)->handle()); }
  size_t index() const { return index_; }
  size_t size() const { return size_; }
  size_t end() const { return index_ + size_ - 1; }
  bool diagsteps() const { return diagsteps_; }

  pb_prox* handle() {
    if (!handle_) handle_ = create();
    return handle_;
  }

 protected:
  virtual pb_prox* create() = 0;
  size_t index_, size_;
  bool diagsteps_;
  pb_prox* handle_;
};

}  // namespace prost

#endif
