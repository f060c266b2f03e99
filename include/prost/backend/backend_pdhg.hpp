// prost/backend/backend_pdhg.hpp -- BackendPDHG<T>: primal-dual hybrid gradient
// (reference: include/prost/backend/backend_pdhg.hpp:40-100, src/backend/backend_pdhg.cu).
#ifndef PROST_BACKEND_PDHG_HPP_
#define PROST_BACKEND_PDHG_HPP_

#include "prost/backend/backend.hpp"
#include "prost/problem.hpp"

namespace prost {

template <typename T>
class BackendPDHG : public Backend<T> {
 public:
  enum StepsizeVariant {
    kPDHGStepsAlg1 = 1,           ///< constant steps
    kPDHGStepsAlg2,               ///< accelerated steps for strongly convex g
    kPDHGStepsResidualGoldstein,  ///< residual balancing (Goldstein, Esser)
    kPDHGStepsResidualBoyd,       ///< residual converging (Fougner, Boyd)
  };

  struct Options {
    double tau0, sigma0;
    int residual_iter;             ///< residuals are refreshed every residual_iter iterations
    bool scale_steps_operator;     ///< rescale so that tau*sigma*|K|^2 = 1
    T alg2_gamma;
    T arg_alpha0, arg_nu, arg_delta;
    T arb_delta, arb_tau;
    typename BackendPDHG<T>::StepsizeVariant stepsize_variant;
  };

  explicit BackendPDHG(const typename BackendPDHG<T>::Options& opts) : opts_(opts) {}
  virtual ~BackendPDHG() {}

 protected:
  virtual pb_backend* create() {
    pb_pdhg_options o;
    pb_pdhg_default_options(&o);
    o.tau0 = opts_.tau0; o.sigma0 = opts_.sigma0; o.residual_iter = opts_.residual_iter;
    o.scale_steps_operator = opts_.scale_steps_operator ? 1 : 0;
    o.alg2_gamma = opts_.alg2_gamma; o.arg_alpha0 = opts_.arg_alpha0; o.arg_nu = opts_.arg_nu;
    o.arg_delta = opts_.arg_delta; o.arb_delta = opts_.arb_delta; o.arb_tau = opts_.arb_tau;
    o.stepsize_variant = static_cast<int>(opts_.stepsize_variant);
    const pb_solver_options so = Backend<T>::solver_options_c(this->solver_opts_);
    pb_backend* h = nullptr;
    detail::check(pb_pdhg_create(detail::context(), this->problem_->handle(), &o, &so, &h));
    return h;
  }
  typename BackendPDHG<T>::Options opts_;
};

}  // namespace prost

#endif
