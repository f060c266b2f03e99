// prost/backend/backend.hpp -- Backend<T>: abstract primal-dual algorithm
// (reference: include/prost/backend/backend.hpp:37-95).
#ifndef PROST_BACKEND_HPP_
#define PROST_BACKEND_HPP_

#include "prost/common.hpp"
#include "prost/solver.hpp"

namespace prost {

template <typename T> class Problem;

template <typename T>
class Backend : detail::require_float<T> {
 public:
  Backend() : handle_(nullptr) {}
  virtual ~Backend() { if (handle_) pb_backend_destroy(handle_); }

  void SetProblem(shared_ptr<Problem<T> > problem) { problem_ = problem; }
  void SetOptions(const typename Solver<T>::Options& opts) { solver_opts_ = opts; }

  /// Allocates the iterates on the GPU and applies Solver::Options::x0 / y0.
  virtual void Initialize() {
    if (handle_) { pb_backend_destroy(handle_); handle_ = nullptr; }
    handle_ = create();
    detail::check(pb_backend_initialize(handle_, solver_opts_.x0.data(), solver_opts_.x0.size(),
                                        solver_opts_.y0.data(), solver_opts_.y0.size()));
  }
  virtual void PerformIteration() { need(); detail::check(pb_backend_iterate(handle_, 1)); }
  virtual void Release() {}

  /// Copies the current primal-dual pair (x, y) to pre-sized host vectors.
  virtual void current_solution(vector<T>& primal_sol, vector<T>& dual_sol) {
    need();
    detail::check(pb_backend_current_solution(handle_, primal_sol.data(), nullptr, dual_sol.data(), nullptr));
  }
  /// Copies (x, z, y, w) with the constraint variables z = Kx and w = -K^T y estimates.
  virtual void current_solution(vector<T>& primal_x, vector<T>& primal_z, vector<T>& dual_y, vector<T>& dual_w) {
    need();
    detail::check(pb_backend_current_solution(handle_, primal_x.data(), primal_z.data(), dual_y.data(), dual_w.data()));
  }

  virtual T primal_residual() const { return res(0); }
  virtual T dual_residual() const { return res(1); }
  virtual T primal_var_norm() const { return res(2); }
  virtual T dual_var_norm() const { return res(3); }
  virtual T eps_primal() const { return res(4); }
  virtual T eps_dual() const { return res(5); }
  virtual size_t gpu_mem_amount() const { need(); return pb_backend_gpu_mem_amount(handle_); }

  pb_backend* handle() { need(); return handle_; }

 protected:
  virtual pb_backend* create() = 0;
  void need() const { if (!handle_) throw Exception("Backend has not been initialized."); }
  T res(int k) const {
    need();
    float out[6];
    detail::check(pb_backend_residuals(handle_, out));
    return out[k];
  }
  static pb_solver_options solver_options_c(const typename Solver<T>::Options& o) { return Solver<T>::to_c(o); }

  shared_ptr<Problem<T> > problem_;
  typename Solver<T>::Options solver_opts_;
  pb_backend* handle_;
};

}  // namespace prost

#endif
