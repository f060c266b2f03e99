// pb_problem.cuh -- Problem (graph-form problem container): include/prost/problem.hpp:64-114,
// src/problem.cu.
#pragma once

#include <memory>
#include <string>
#include <vector>

#include "pb_linop.cuh"
#include "pb_prox.cuh"

namespace pb {

typedef std::vector<std::shared_ptr<Prox>> ProxList;

class Problem {
 public:
  enum Scaling { kScalingIdentity, kScalingAlpha, kScalingCustom };

  explicit Problem(Context* ctx) : ctx_(ctx), linop_(new LinearOperator(ctx)) {}

  void add_block(std::shared_ptr<Block> b) { linop_->add_block(std::move(b)); }
  void add_prox_g(std::shared_ptr<Prox> p) { prox_g_.push_back(std::move(p)); }
  void add_prox_f(std::shared_ptr<Prox> p) { prox_f_.push_back(std::move(p)); }
  void add_prox_gstar(std::shared_ptr<Prox> p) { prox_gstar_.push_back(std::move(p)); }
  void add_prox_fstar(std::shared_ptr<Prox> p) { prox_fstar_.push_back(std::move(p)); }
  void set_dimensions(size_t nrows, size_t ncols) { nrows_ = nrows; ncols_ = ncols; dims_set_ = true; }
  void set_scaling_alpha(float alpha) { scaling_type_ = kScalingAlpha; scaling_alpha_ = alpha; }
  void set_scaling_identity() { scaling_type_ = kScalingIdentity; }
  void set_scaling_custom(const float* left, size_t nl, const float* right, size_t nr);

  void initialize();
  void dualize();
  float normest(float tol, int max_iters, const float* h_x0);

  size_t nrows() const { return nrows_; }
  size_t ncols() const { return ncols_; }
  size_t gpu_mem_amount() const;
  bool dualized() const { return dualized_; }
  bool initialized() const { return initialized_; }

  LinearOperator* linop() const { return linop_.get(); }
  // K apply honouring Dualize(): primal problem K / K^T, dual problem -K^T / -K
  void apply_K(float* d_res, const float* d_rhs, bool adjoint);

  const ProxList& prox_g() const { return prox_g_; }
  const ProxList& prox_f() const { return prox_f_; }
  const ProxList& prox_gstar() const { return prox_gstar_; }
  const ProxList& prox_fstar() const { return prox_fstar_; }

  // Sigma (left, nrows) and T (right, ncols)
  const float* scaling_left() const { return d_left_.data(); }
  const float* scaling_right() const { return d_right_.data(); }
  const std::vector<float>& scaling_left_host() const { return left_host_; }
  const std::vector<float>& scaling_right_host() const { return right_host_; }
  ScaleRef left_ref() const { return left_uniform_ ? ScaleRef{nullptr, left_host_.empty() ? 1.f : left_host_[0]} : ScaleRef{d_left_.data(), 1.f}; }
  ScaleRef right_ref() const { return right_uniform_ ? ScaleRef{nullptr, right_host_.empty() ? 1.f : right_host_[0]} : ScaleRef{d_right_.data(), 1.f}; }
  Context* ctx() const { return ctx_; }

 private:
  void average_preconditioners(std::vector<float>& precond, const ProxList& prox);

  Context* ctx_;
  std::shared_ptr<LinearOperator> linop_;
  ProxList prox_g_, prox_f_, prox_gstar_, prox_fstar_;
  size_t nrows_ = 0, ncols_ = 0;
  bool dims_set_ = false;
  // the reference leaves scaling_type_ uninitialised in C++ (problem.hpp:125); MATLAB always
  // sets alpha = 1 (matlab/+prost/problem.m:10), which is the default here (Appendix B #19)
  Scaling scaling_type_ = kScalingAlpha;
  float scaling_alpha_ = 1.f;
  std::vector<float> left_host_, right_host_;
  DeviceBuffer<float> d_left_, d_right_;
  bool left_uniform_ = false, right_uniform_ = false;
  bool dualized_ = false;
  bool initialized_ = false;
};

}  // namespace pb
