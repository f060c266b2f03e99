// prost/problem.hpp -- Problem<T>: min_x max_y g(x) + <Kx, y> - f*(y)
// (reference: include/prost/problem.hpp:64-114, src/problem.cu).
#ifndef PROST_PROBLEM_HPP_
#define PROST_PROBLEM_HPP_

#include "prost/linop/block.hpp"
#include "prost/linop/linearoperator.hpp"
#include "prost/prox/prox.hpp"

namespace prost {

template <typename T>
class Problem : detail::require_float<T> {
 public:
  typedef vector<shared_ptr<Prox<T> > > ProxList;

  Problem() : linop_(new LinearOperator<T>()), nrows_(0), ncols_(0), dims_set_(false), scaling_(kAlpha),
              alpha_(1), handle_(nullptr) {}
  virtual ~Problem() { if (handle_) pb_problem_destroy(handle_); }

  void AddBlock(shared_ptr<Block<T> > block) { linop_->AddBlock(block); }
  void AddProx_g(shared_ptr<Prox<T> > prox) { prox_g_.push_back(prox); }
  void AddProx_f(shared_ptr<Prox<T> > prox) { prox_f_.push_back(prox); }
  void AddProx_gstar(shared_ptr<Prox<T> > prox) { prox_gstar_.push_back(prox); }
  void AddProx_fstar(shared_ptr<Prox<T> > prox) { prox_fstar_.push_back(prox); }

  /// Builds the operator, fills uncovered ranges with identity proxes, checks the prox domains and
  /// computes the diagonal preconditioners (problem.cu:195-323).
  void Initialize() {
    if (handle_) { pb_problem_destroy(handle_); handle_ = nullptr; }
    detail::check(pb_problem_create(detail::context(), &handle_));
    const std::vector<std::shared_ptr<Block<T> > >& blocks = linop_->blocks();
    for (size_t i = 0; i < blocks.size(); ++i) detail::check(pb_problem_add_block(handle_, blocks[i]->handle()));
    for (size_t i = 0; i < prox_g_.size(); ++i) detail::check(pb_problem_add_prox_g(handle_, prox_g_[i]->handle()));
    for (size_t i = 0; i < prox_f_.size(); ++i) detail::check(pb_problem_add_prox_f(handle_, prox_f_[i]->handle()));
    for (size_t i = 0; i < prox_gstar_.size(); ++i)
      detail::check(pb_problem_add_prox_gstar(handle_, prox_gstar_[i]->handle()));
    for (size_t i = 0; i < prox_fstar_.size(); ++i)
      detail::check(pb_problem_add_prox_fstar(handle_, prox_fstar_[i]->handle()));
    if (dims_set_) detail::check(pb_problem_set_dimensions(handle_, nrows_, ncols_));
    if (scaling_ == kAlpha) detail::check(pb_problem_set_scaling_alpha(handle_, alpha_));
    else if (scaling_ == kIdentity) detail::check(pb_problem_set_scaling_identity(handle_));
    else detail::check(pb_problem_set_scaling_custom(handle_, left_.data(), left_.size(), right_.data(), right_.size()));
    detail::check(pb_problem_initialize(handle_));
    linop_->Initialize();
    nrows_ = pb_problem_nrows(handle_);
    ncols_ = pb_problem_ncols(handle_);
  }
  void Release() {}

  /// left / right are Sigma^{1/2} and Tau^{1/2} (the library stores their squares, problem.cu:344-364).
  void SetScalingCustom(const vector<T>& left, const vector<T>& right) { scaling_ = kCustom; left_ = left; right_ = right; }
  /// Pock-Chambolle diagonal preconditioners (ICCV '11) with exponent alpha.
  void SetScalingAlpha(T alpha) { scaling_ = kAlpha; alpha_ = alpha; }
  void SetScalingIdentity() { scaling_ = kIdentity; }
  void SetDimensions(size_t nrows, size_t ncols) { nrows_ = nrows; ncols_ = ncols; dims_set_ = true; }

  shared_ptr<LinearOperator<T> > linop() const { return linop_; }
  const ProxList& prox_f() const { return prox_f_; }
  const ProxList& prox_g() const { return prox_g_; }
  const ProxList& prox_fstar() const { return prox_fstar_; }
  const ProxList& prox_gstar() const { return prox_gstar_; }
  size_t nrows() const { return nrows_; }
  size_t ncols() const { return ncols_; }
  size_t gpu_mem_amount() const { need(); return pb_problem_gpu_mem_amount(handle_); }

  /// Power iteration for |Sigma^{1/2} K Tau^{1/2}| (problem.cu:428-500).
  T normest(T tol = 1e-6, int max_iters = 100) {
    need();
    float out = 0;
    detail::check(pb_problem_normest(handle_, tol, max_iters, nullptr, &out));
    return out;
  }

  /// Swaps g <-> f*, f <-> g*, K <-> -K^T; call after Initialize() (problem.cu:538-547).
  void Dualize() {
    need();
    detail::check(pb_problem_dualize(handle_));
    prox_g_.swap(prox_fstar_);
    prox_gstar_.swap(prox_f_);
    std::swap(nrows_, ncols_);
  }

  /// Diagonal preconditioners Sigma (nrows) and Tau (ncols) as computed by Initialize().
  void scaling(vector<T>& left, vector<T>& right) const {
    need();
    left.resize(nrows_);
    right.resize(ncols_);
    detail::check(pb_problem_get_scaling(handle_, left.data(), right.data()));
  }

  pb_problem* handle() { need(); return handle_; }

 protected:
  void need() const { if (!handle_) throw Exception("Problem has not been initialized."); }
  enum Scaling { kIdentity, kAlpha, kCustom };
  shared_ptr<LinearOperator<T> > linop_;
  ProxList prox_g_, prox_f_, prox_gstar_, prox_fstar_;
  size_t nrows_, ncols_;
  bool dims_set_;
  Scaling scaling_;
  T alpha_;
  vector<T> left_, right_;
  pb_problem* handle_;
};

}  // namespace prost

#endif
