// prost/backend/backend_admm.hpp -- BackendADMM<T>: graph-projection ADMM with inexact CG projections
// (reference: include/prost/backend/backend_admm.hpp:38-63, src/backend/backend_admm.cu).
#ifndef PROST_BACKEND_ADMM_HPP_
#define PROST_BACKEND_ADMM_HPP_

#include "prost/backend/backend.hpp"
#include "prost/problem.hpp"

namespace prost {

template <typename T>
class BackendADMM : public Backend<T> {
 public:
  struct Options {
    double rho0;
    double alpha;                              ///< over-relaxation
    double cg_tol_pow, cg_tol_min, cg_tol_max;
    int cg_max_iter;
    int residual_iter;
    T arb_delta, arb_tau, arb_gamma;
  };

  explicit BackendADMM(const typename BackendADMM<T>::Options& opts) : opts_(opts) {}
  virtual ~BackendADMM() {}

 protected:
  virtual pb_backend* create() {
    pb_admm_options o;
    o.rho0 = opts_.rho0; o.alpha = opts_.alpha; o.cg_tol_pow = opts_.cg_tol_pow; o.cg_tol_min = opts_.cg_tol_min;
    o.cg_tol_max = opts_.cg_tol_max; o.cg_max_iter = opts_.cg_max_iter; o.residual_iter = opts_.residual_iter;
    o.arb_delta = opts_.arb_delta; o.arb_tau = opts_.arb_tau; o.arb_gamma = opts_.arb_gamma;
    const pb_solver_options so = Backend<T>::solver_options_c(this->solver_opts_);
    pb_backend* h = nullptr;
    detail::check(pb_admm_create(detail::context(), this->problem_->handle(), &o, &so, &h));
    return h;
  }
  typename BackendADMM<T>::Options opts_;
};

}  // namespace prost

#endif
