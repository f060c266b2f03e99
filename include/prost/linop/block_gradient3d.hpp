// prost/linop/block_gradient3d.hpp -- BlockGradient3D<T>
// (reference: include/prost/linop/block_gradient3d.hpp, src/linop/block_gradient3d.cu).
#ifndef PROST_BLOCK_GRADIENT3D_HPP_
#define PROST_BLOCK_GRADIENT3D_HPP_

#include "prost/linop/block.hpp"

namespace prost {

/// Forward-difference gradient of an nx x ny image with L labels/channels
/// (planar y + x*ny + l*nx*ny or label-first l + y*L + x*ny*L); third component along l with a Dirichlet boundary at l = L-1.
template <typename T>
class BlockGradient3D : public Block<T> {
 public:
  BlockGradient3D(size_t row, size_t col, size_t nx, size_t ny, size_t L, bool label_first)
      : Block<T>(row, col, nx * ny * L * 3, nx * ny * L), nx_(nx), ny_(ny), L_(L), label_first_(label_first) {}

 protected:
  virtual pb_block* create() {
    pb_block* h = nullptr;
    detail::check(pb_block_create_gradient3d(detail::context(), this->row_, this->col_, nx_, ny_, L_,
                                              label_first_ ? 1 : 0, &h));
    return h;
  }
  size_t nx_, ny_, L_;
  bool label_first_;
};

}  // namespace prost

#endif
