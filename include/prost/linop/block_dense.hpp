// prost/linop/block_dense.hpp -- BlockDense<T> (reference: include/prost/linop/block_dense.hpp:42-43).
#ifndef PROST_BLOCK_DENSE_HPP_
#define PROST_BLOCK_DENSE_HPP_

#include "prost/linop/block.hpp"

namespace prost {

template <typename T>
class BlockDense : public Block<T> {
  BlockDense(size_t row, size_t col, size_t nrows, size_t ncols) : Block<T>(row, col, nrows, ncols) {}

 public:
  /// `data` is column-major with leading dimension nrows.
  static BlockDense<T>* CreateFromColFirstData(size_t row, size_t col, size_t nrows, size_t ncols,
                                               const std::vector<T>& data) {
    BlockDense<T>* b = new BlockDense<T>(row, col, nrows, ncols);
    b->data_ = data;
    return b;
  }

 protected:
  virtual pb_block* create() {
    pb_block* h = nullptr;
    detail::check(pb_block_create_dense(detail::context(), this->row_, this->col_, this->nrows_, this->ncols_,
                                        data_.data(), &h));
    return h;
  }
  std::vector<T> data_;
};

}  // namespace prost

#endif
