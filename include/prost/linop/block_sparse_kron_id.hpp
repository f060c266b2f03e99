// prost/linop/block_sparse_kron_id.hpp -- BlockSparseKronId<T>: kron(K, I_diaglength) for a sparse factor K
// (reference: include/prost/linop/block_sparse_kron_id.hpp:39-48, src/linop/block_sparse_kron_id.cu).
#ifndef PROST_BLOCK_SPARSE_KRON_ID_HPP_
#define PROST_BLOCK_SPARSE_KRON_ID_HPP_

#include <cstdint>
#include <vector>

#include "prost/linop/block.hpp"

namespace prost {

template <typename T>
class BlockSparseKronId : public Block<T> {
  BlockSparseKronId(size_t row, size_t col, size_t nrows, size_t ncols) : Block<T>(row, col, nrows, ncols) {}

 public:
  /// Compressed-sparse-column factor (m x n, int32 indices), as MATLAB stores sparse matrices.
  static BlockSparseKronId<T>* CreateFromCSC(size_t row, size_t col, size_t diaglength, int m, int n, int nnz,
                                 const std::vector<T>& val, const std::vector<int32_t>& ptr,
                                 const std::vector<int32_t>& ind) {
    BlockSparseKronId<T>* b = new BlockSparseKronId<T>(row, col, static_cast<size_t>(m) * diaglength, static_cast<size_t>(n) * diaglength);
    b->diaglength_ = diaglength;
    b->m_ = m;
    b->n_ = n;
    b->nnz_ = nnz;
    b->val_.assign(val.begin(), val.end());
    b->ptr_ = ptr;
    b->ind_ = ind;
    return b;
  }

 protected:
  virtual pb_block* create() {
    pb_block* h = nullptr;
    detail::check(pb_block_create_sparse_kron_id(detail::context(), this->row_, this->col_, diaglength_, m_, n_, nnz_, val_.data(),
                  ptr_.data(), ind_.data(), &h));
    return h;
  }
  size_t diaglength_;
  int m_, n_, nnz_;
  std::vector<float> val_;
  std::vector<int32_t> ptr_, ind_;
};

}  // namespace prost

#endif
