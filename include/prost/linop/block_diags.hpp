// prost/linop/block_diags.hpp -- BlockDiags<T> (reference: include/prost/linop/block_diags.hpp:40-57).
#ifndef PROST_BLOCK_DIAGS_HPP_
#define PROST_BLOCK_DIAGS_HPP_

#include "prost/linop/block.hpp"

namespace prost {

/// Sum of constant diagonals: K[r, r + offsets[d]] = factors[d].
template <typename T>
class BlockDiags : public Block<T> {
 public:
  BlockDiags(size_t row, size_t col, size_t nrows, size_t ncols, size_t ndiags,
             const std::vector<ssize_t>& offsets, const std::vector<T>& factors)
      : Block<T>(row, col, nrows, ncols), ndiags_(ndiags), offsets_(offsets.begin(), offsets.end()),
        factors_(factors) {}

  /// The reference keeps all diagonals in 1024 __constant__ slots and needs this reset per problem
  /// (prost.cpp:74); here diagonals live with their block, so this is a no-op kept for source
  /// compatibility.
  static void ResetConstMem() {}

 protected:
  virtual pb_block* create() {
    pb_block* h = nullptr;
    detail::check(pb_block_create_diags(detail::context(), this->row_, this->col_, this->nrows_, this->ncols_,
                                        ndiags_, offsets_.data(), factors_.data(), &h));
    return h;
  }
  size_t ndiags_;
  std::vector<int64_t> offsets_;
  std::vector<T> factors_;
};

}  // namespace prost

#endif
