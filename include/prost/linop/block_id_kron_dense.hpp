// prost/linop/block_id_kron_dense.hpp -- BlockIdKronDense<T>: kron(I_diaglength, K) for a small dense K
// (reference: include/prost/linop/block_id_kron_dense.hpp, src/linop/block_id_kron_dense.cu).
#ifndef PROST_BLOCK_ID_KRON_DENSE_HPP_
#define PROST_BLOCK_ID_KRON_DENSE_HPP_

#include <vector>

#include "prost/linop/block.hpp"

namespace prost {

template <typename T>
class BlockIdKronDense : public Block<T> {
  BlockIdKronDense(size_t row, size_t col, size_t nrows, size_t ncols) : Block<T>(row, col, nrows, ncols) {}

 public:
  /// `data` is the mat_nrows x mat_ncols factor, column-major; the block is that times `diaglength` in both directions.
  static BlockIdKronDense<T>* CreateFromColFirstData(size_t diaglength, size_t row, size_t col, size_t nrows, size_t ncols,
                                          const std::vector<T>& data) {
    BlockIdKronDense<T>* b = new BlockIdKronDense<T>(row, col, nrows * diaglength, ncols * diaglength);
    b->diaglength_ = diaglength;
    b->mat_nrows_ = nrows;
    b->mat_ncols_ = ncols;
    b->data_.assign(data.begin(), data.end());
    return b;
  }

 protected:
  virtual pb_block* create() {
    pb_block* h = nullptr;
    detail::check(pb_block_create_id_kron_dense(detail::context(), diaglength_, this->row_, this->col_, mat_nrows_, mat_ncols_,
                  data_.data(), &h));
    return h;
  }
  size_t diaglength_, mat_nrows_, mat_ncols_;
  std::vector<float> data_;
};

}  // namespace prost

#endif
