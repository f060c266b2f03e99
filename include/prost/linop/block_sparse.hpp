// prost/linop/block_sparse.hpp -- BlockSparse<T> (reference: include/prost/linop/block_sparse.hpp:43-51).
#ifndef PROST_BLOCK_SPARSE_HPP_
#define PROST_BLOCK_SPARSE_HPP_

#include "prost/linop/block.hpp"

namespace prost {

template <typename T>
class BlockSparse : public Block<T> {
  BlockSparse(size_t row, size_t col, size_t nrows, size_t ncols) : Block<T>(row, col, nrows, ncols), nnz_(0) {}

 public:
  /// Compressed-sparse-column input, as MATLAB stores sparse matrices.
  static BlockSparse<T>* CreateFromCSC(size_t row, size_t col, int m, int n, int nnz, const vector<T>& val,
                                       const vector<int32_t>& ptr, const vector<int32_t>& ind) {
    BlockSparse<T>* b = new BlockSparse<T>(row, col, m, n);
    b->nnz_ = nnz;
    b->val_ = val;
    b->ptr_ = ptr;
    b->ind_ = ind;
    return b;
  }

 protected:
  virtual pb_block* create() {
    pb_block* h = nullptr;
    detail::check(pb_block_create_sparse_csc(detail::context(), this->row_, this->col_,
                                             static_cast<int>(this->nrows_), static_cast<int>(this->ncols_), nnz_,
                                             val_.data(), ptr_.data(), ind_.data(), &h));
    return h;
  }
  int nnz_;
  vector<T> val_;
  vector<int32_t> ptr_, ind_;
};

}  // namespace prost

#endif
