// prost/linop/block.hpp -- Block<T>: one block of the linear operator
// (reference: include/prost/linop/block.hpp:37-83).
#ifndef PROST_BLOCK_HPP_
#define PROST_BLOCK_HPP_

#include "prost/common.hpp"

namespace prost {

template <typename T>
class Block : detail::require_float<T> {
 public:
  Block(size_t row, size_t col, size_t nrows, size_t ncols)
      : row_(row), col_(col), nrows_(nrows), ncols_(ncols), handle_(nullptr) {}
  virtual ~Block() { if (handle_) pb_block_destroy(handle_); }

  /// Uploads the block's data (the reference's H2D copies happen here as well).
  virtual void Initialize() { handle(); }
  virtual void Release() {}

  /// Sum over a row / column of |K_ij|^alpha, block-local indices (block.hpp:63-70).
  virtual T row_sum(size_t row, T alpha) const { return pb_block_row_sum(const_cast<Block*>(this)->handle(), row, alpha); }
  virtual T col_sum(size_t col, T alpha) const { return pb_block_col_sum(const_cast<Block*>(this)->handle(), col, alpha); }
  virtual size_t gpu_mem_amount() const { return pb_block_gpu_mem_amount(const_cast<Block*>(this)->handle()); }

  size_t row() const { return row_; }
  size_t col() const { return col_; }
  size_t nrows() const { return nrows_; }
  size_t ncols() const { return ncols_; }

  /// C-ABI handle (created on first use).
  pb_block* handle() {
    if (!handle_) handle_ = create();
    return handle_;
  }

 protected:
  virtual pb_block* create() = 0;
  size_t row_, col_, nrows_, ncols_;
  pb_block* handle_;
};

}  // namespace prost

#endif
