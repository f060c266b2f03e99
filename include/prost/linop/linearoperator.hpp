// prost/linop/linearoperator.hpp -- LinearOperator<T>
// (reference: include/prost/linop/linearoperator.hpp:36-90).
#ifndef PROST_LINEAROPERATOR_HPP_
#define PROST_LINEAROPERATOR_HPP_

#include "prost/linop/block.hpp"

namespace prost {

template <typename T>
class LinearOperator : detail::require_float<T> {
 public:
  LinearOperator() : handle_(nullptr), initialized_(false) {}
  virtual ~LinearOperator() { if (handle_) pb_linop_destroy(handle_); }

  void AddBlock(std::shared_ptr<Block<T> > block) { blocks_.push_back(block); }

  /// Computes the operator size and rejects overlapping blocks (linearoperator.cu:83-125).
  virtual void Initialize() {
    if (handle_) { pb_linop_destroy(handle_); handle_ = nullptr; }
    detail::check(pb_linop_create(detail::context(), &handle_));
    for (size_t i = 0; i < blocks_.size(); ++i) detail::check(pb_linop_add_block(handle_, blocks_[i]->handle()));
    detail::check(pb_linop_initialize(handle_));
    initialized_ = true;
  }
  virtual void Release() {}

  /// Host-vector overloads (linearoperator.cu:172-220): result is resized; returns device ms.
  double Eval(std::vector<T>& result, const std::vector<T>& rhs) { return eval(result, rhs, false); }
  double EvalAdjoint(std::vector<T>& result, const std::vector<T>& rhs) { return eval(result, rhs, true); }

  T row_sum(size_t row, T alpha) const { need(); return pb_linop_row_sum(handle_, row, alpha); }
  T col_sum(size_t col, T alpha) const { need(); return pb_linop_col_sum(handle_, col, alpha); }
  size_t nrows() const { need(); return pb_linop_nrows(handle_); }
  size_t ncols() const { need(); return pb_linop_ncols(handle_); }
  size_t gpu_mem_amount() const {
    size_t mem = 0;
    for (size_t i = 0; i < blocks_.size(); ++i) mem += blocks_[i]->gpu_mem_amount();
    return mem;
  }
  const std::vector<std::shared_ptr<Block<T> > >& blocks() const { return blocks_; }

 protected:
  void need() const { if (!initialized_) throw Exception("LinearOperator has not been initialized."); }
  double eval(std::vector<T>& result, const std::vector<T>& rhs, bool transpose) {
    need();
    const size_t nin = transpose ? nrows() : ncols();
    if (rhs.size() != nin) throw Exception("LinearOperator::Eval: right-hand side has the wrong size.");
    result.resize(transpose ? ncols() : nrows());
    double ms = 0;
    detail::check(pb_linop_eval_host(handle_, result.data(), rhs.data(), transpose ? 1 : 0, &ms));
    return ms;
  }
  std::vector<std::shared_ptr<Block<T> > > blocks_;
  pb_linop* handle_;
  bool initialized_;
};

}  // namespace prost

#endif
