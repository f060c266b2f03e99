// prost/linop/block_zero.hpp -- BlockZero<T> (reference: include/prost/linop/block_zero.hpp).
#ifndef PROST_BLOCK_ZERO_HPP_
#define PROST_BLOCK_ZERO_HPP_

#include "prost/linop/block.hpp"

namespace prost {

template <typename T>
class BlockZero : public Block<T> {
 public:
  BlockZero(size_t row, size_t col, size_t nrows, size_t ncols) : Block<T>(row, col, nrows, ncols) {}

 protected:
  virtual pb_block* create() {
    pb_block* h = nullptr;
    detail::check(pb_block_create_zero(detail::context(), this->row_, this->col_, this->nrows_, this->ncols_, &h));
    return h;
  }
};

}  // namespace prost

#endif
