// prost/linop/block_dense_kron_id.hpp -- BlockDenseKronId<T>: kron(K, I_diaglength) for a small dense K
// (reference: include/prost/linop/block_dense_kron_id.hpp, src/linop/block_dense_kron_id.cu).
#ifndef PROST_BLOCK_DENSE_KRON_ID_HPP_
#define PROST_BLOCK_DENSE_KRON_ID_HPP_

#include <vector>

#include "prost/linop/block.hpp"

namespace prost {

template <typename T>
class BlockDenseKronId : public Block<T> {
  BlockDenseKronId(size_t row, size_t col, size_t nrows, size_t ncols) : Block<T>(row, col, nrows, ncols) {}

 public:
  /// `data` is the mat_nrows x mat_ncols factor, column-major; the block is that times `diaglength` in both directions.
  static BlockDenseKronId<T>* CreateFromColFirstData(size_t diaglength, size_t row, size_t col, size_t nrows, size_t ncols,
                                          const std::vector<T>& data) {
    BlockDenseKronId<T>* b = new BlockDenseKronId<T>(row, col, nrows * diaglength, ncols * diaglength);
    b->diaglength_ = diaglength;
    b->mat_nrows_ = nrows;
    b->mat_ncols_ = ncols;
    b->data_.assign(data.begin(), data.end());
    return b;
  }

 protected:
  virtual pb_block* create() {
    pb_block* h = nullptr;
    detail::check(pb_block_create_dense_kron_id(detail::context(), diaglength_, this->row_, this->col_, mat_nrows_, mat_ncols_,
                  data_.data(), &h));
    return h;
  }
  size_t diaglength_, mat_nrows_, mat_ncols_;
  std::vector<float> data_;
};

}  // namespace prost

#endif
