// pb_linop.cuh -- linear-operator blocks: POD device descriptors with pointwise
// ("gather form") row/column products, and the host-side block objects.
//
// Every block kind the fused passes understand is evaluated pointwise: (K u)[r] and
// (K^T p)[c] are computed from the operand vectors directly, so an operator apply never
// needs a zero-fill plus read-modify-write accumulation pass (reference:
// linearoperator.cu:134-170 + `+=` in every block kernel) and can be fused with the
// prox-argument arithmetic and the prox itself.
#pragma once

#include <cstdint>
#include <memory>
#include <vector>

#include "pb_common.cuh"

namespace pb {

enum BlockKind : int {
  kBlockZero = 0,
  kBlockGradient2D = 1,
  kBlockGradient3D = 2,
  kBlockDiags = 3,
  kBlockSparse = 4,
  kBlockDense = 5,
  kBlockDenseKronId = 6,
  kBlockIdKronDense = 7,
  kBlockSparseKronId = 8,
  kBlockIdKronSparse = 9,
};

// POD view of one block, passed by value to kernels.
struct BlockDesc {
  int kind = kBlockZero;
  int label_first = 0;
  uint32_t row = 0, col = 0, nrows = 0, ncols = 0;   // fused paths require sizes < 2^31
  // gradient geometry
  uint32_t nx = 0, ny = 0, L = 0;
  uint32_t plane = 0;            // nx*ny*L (= ncols of a gradient block)
  FastDiv div_ny, div_L, div_nxny, div_nyL, div_plane;
  // diags
  int ndiags = 0;
  const long long* offsets = nullptr;   // device, sorted ascending
  const float* factors = nullptr;       // device
  // sparse: CSR of K and CSR of K^T (reference stores both: block_sparse.cu:81-109)
  const int* ptr = nullptr; const int* ind = nullptr; const float* val = nullptr;
  const int* ptr_t = nullptr; const int* ind_t = nullptr; const float* val_t = nullptr;
  // dense, column-major, lda = nrows
  const float* dense = nullptr;
};

// true for the kinds whose products are cheap pointwise gathers (fusable)
__host__ __device__ inline bool block_is_stencil(int kind) {
  return kind == kBlockZero || kind == kBlockGradient2D || kind == kBlockGradient3D ||
         kind == kBlockDiags;
}

#ifdef __CUDACC__

// decode linear index of a gradient block's domain into (x, y, l)
__device__ __forceinline__ void grad_decode(const BlockDesc& b, uint32_t idx, uint32_t& x,
                                            uint32_t& y, uint32_t& l) {
  if (b.label_first) {           // idx = l + y*L + x*ny*L
    uint32_t rem;
    b.div_nyL.divmod(idx, x, rem);
    b.div_L.divmod(rem, y, l);
  } else {                       // idx = y + x*ny + l*nx*ny
    uint32_t rem;
    b.div_nxny.divmod(idx, l, rem);
    b.div_ny.divmod(rem, x, y);
  }
}

// (K u)[r] for r local to the block; u points at the block's first column.
// Gradient: block_gradient2d.cu:25-78, block_gradient3d.cu:25-81.  Diags: block_diags.cu:36-65.
__device__ __forceinline__ float block_row_dot(const BlockDesc& b, uint32_t r,
                                               const float* __restrict__ u) {
  switch (b.kind) {
    case kBlockGradient2D:
    case kBlockGradient3D: {
      uint32_t comp, idx;
      b.div_plane.divmod(r, comp, idx);
      uint32_t x, y, l;
      grad_decode(b, idx, x, y, l);
      const float v = u[idx];
      const uint32_t sy = b.label_first ? b.L : 1u;
      const uint32_t sx = b.label_first ? b.ny * b.L : b.ny;
      if (comp == 0) return (x < b.nx - 1) ? u[idx + sx] - v : 0.f;
      if (comp == 1) return (y < b.ny - 1) ? u[idx + sy] - v : 0.f;
      const uint32_t sl = b.label_first ? 1u : b.nx * b.ny;      // 3-D only: Dirichlet at l = L-1
      return (l < b.L - 1) ? u[idx + sl] - v : -v;
    }
    case kBlockDiags: {
      float acc = 0.f;
      for (int i = 0; i < b.ndiags; ++i) {
        const long long c = static_cast<long long>(r) + b.offsets[i];
        if (c < 0) continue;
        if (c >= static_cast<long long>(b.ncols)) break;   // offsets are sorted
        acc += u[c] * b.factors[i];
      }
      return acc;
    }
    case kBlockSparse: {
      float acc = 0.f;
      for (int k = b.ptr[r]; k < b.ptr[r + 1]; ++k) acc += b.val[k] * u[b.ind[k]];
      return acc;
    }
    default: return 0.f;
  }
}

// (K^T p)[c] for c local to the block; p points at the block's first row.
// Gradient adjoint = minus divergence: block_gradient2d.cu:80-139, block_gradient3d.cu:83-150.
// Diags adjoint: block_diags.cu:67-96.
__device__ __forceinline__ float block_col_dot(const BlockDesc& b, uint32_t c,
                                               const float* __restrict__ p) {
  switch (b.kind) {
    case kBlockGradient2D:
    case kBlockGradient3D: {
      uint32_t x, y, l;
      grad_decode(b, c, x, y, l);
      const uint32_t sy = b.label_first ? b.L : 1u;
      const uint32_t sx = b.label_first ? b.ny * b.L : b.ny;
      const float* p1 = p;                 // x-component plane
      const float* p2 = p + b.plane;       // y-component plane
      float divy = (y < b.ny - 1) ? p2[c] : 0.f;
      if (y > 0) divy -= p2[c - sy];
      float divx = (x < b.nx - 1) ? p1[c] : 0.f;
      if (x > 0) divx -= p1[c - sx];
      if (b.kind == kBlockGradient2D) return -(divx + divy);
      const uint32_t sl = b.label_first ? 1u : b.nx * b.ny;
      const float* p3 = p + 2u * static_cast<size_t>(b.plane);
      float divl = p3[c];
      if (l > 0) divl -= p3[c - sl];
      return -(divx + divy + divl);
    }
    case kBlockDiags: {
      float acc = 0.f;
      const long long cc = static_cast<long long>(c);
      for (int i = 0; i < b.ndiags; ++i) {
        const long long ofs = b.offsets[i];
        if (ofs > cc) break;
        const long long r = cc - ofs;
        if (r < static_cast<long long>(b.nrows)) acc += p[r] * b.factors[i];
      }
      return acc;
    }
    case kBlockSparse: {
      float acc = 0.f;
      for (int k = b.ptr_t[c]; k < b.ptr_t[c + 1]; ++k) acc += b.val_t[k] * p[b.ind_t[k]];
      return acc;
    }
    default: return 0.f;
  }
}

#endif  // __CUDACC__

// ---- host-side block objects ----------------------------------------------------------------

class Block {
 public:
  Block(Context* ctx, size_t row, size_t col, size_t nrows, size_t ncols)
      : ctx_(ctx), row_(row), col_(col), nrows_(nrows), ncols_(ncols) {}
  virtual ~Block() {}

  size_t row() const { return row_; }
  size_t col() const { return col_; }
  size_t nrows() const { return nrows_; }
  size_t ncols() const { return ncols_; }

  virtual int kind() const = 0;
  virtual float row_sum(size_t row, float alpha) const = 0;
  virtual float col_sum(size_t col, float alpha) const = 0;
  virtual size_t gpu_mem_amount() const { return 0; }
  // true when row_sum / col_sum do not depend on the index (gradient, zero)
  virtual bool uniform_sums() const { return false; }

  // d_res[0:nrows] += K d_rhs[0:ncols] (pointers already offset to the block: block.cu:46-56)
  virtual void eval_local_add(float* d_res, const float* d_rhs) = 0;
  // d_res[0:ncols] += K^T d_rhs[0:nrows] (block.cu:58-68)
  virtual void eval_adjoint_local_add(float* d_res, const float* d_rhs) = 0;

  // Optional: d_res[0:nrows] = K d_rhs (resp. d_res[0:ncols] = K^T d_rhs), i.e. the block OVERWRITES its output
  // range instead of accumulating into it.  LinearOperator::eval uses it for beta = 0 when no other block writes the
  // same outputs, which saves the zero fill and the read of the read-modify-write (two thirds of the output traffic
  // of a block with few inputs per output).  Return false when not implemented.
  virtual bool eval_local_set(float*, const float*) { return false; }
  virtual bool eval_adjoint_local_set(float*, const float*) { return false; }

  // descriptor for pointwise evaluation inside fused kernels
  virtual BlockDesc desc() const;

 protected:
  Context* ctx_;
  size_t row_, col_, nrows_, ncols_;
};

std::shared_ptr<Block> make_block_gradient(Context* ctx, bool three_d, size_t row, size_t col,
                                           size_t nx, size_t ny, size_t L, bool label_first);
std::shared_ptr<Block> make_block_diags(Context* ctx, size_t row, size_t col, size_t nrows,
                                        size_t ncols, size_t ndiags, const int64_t* offsets,
                                        const float* factors);
std::shared_ptr<Block> make_block_sparse_csc(Context* ctx, size_t row, size_t col, int m, int n,
                                             int nnz, const float* val, const int32_t* ptr,
                                             const int32_t* ind);
std::shared_ptr<Block> make_block_dense(Context* ctx, size_t row, size_t col, size_t nrows,
                                        size_t ncols, const float* data);
// kron(K, I_d) (BlockDenseKronId) and kron(I_d, K) (BlockIdKronDense); K is mat_nrows x mat_ncols, column-major
std::shared_ptr<Block> make_block_dense_kron(Context* ctx, bool id_first, size_t diaglength, size_t row, size_t col,
                                             size_t mat_nrows, size_t mat_ncols, const float* data);
// the same two products for a sparse factor given in CSC (BlockSparseKronId / BlockIdKronSparse::CreateFromCSC)
std::shared_ptr<Block> make_block_sparse_kron(Context* ctx, bool id_first, size_t diaglength, size_t row, size_t col,
                                              int m, int n, int nnz, const float* val, const int32_t* ptr,
                                              const int32_t* ind);
std::shared_ptr<Block> make_block_zero(Context* ctx, size_t row, size_t col, size_t nrows,
                                       size_t ncols);

// Dense Kronecker products on the tensor cores (pb_kron_tc.cu): tcgen05.mma kind::tf32 with the 3 x TF32 split,
// accumulators in TMEM.  `supported` decides per call (factor shape, alignment); BlockDenseKron falls back to its
// fp32 kernels otherwise.
struct KronTensorCore {
  struct Packed {               // hi / lo parts of the factor in the shared-memory operand layout, zero padded
    DeviceBuffer<float> hi, lo;
    uint32_t n_pad = 0, k_pad = 0;
    bool ready = false;
  };
  static bool supported(bool id_first, uint32_t n_out, uint32_t n_in, size_t d, const float* res, const float* rhs);
  static size_t smem_bytes(bool id_first, uint32_t n_out, uint32_t n_pad, uint32_t k_pad);
  // factor entries K(o, i) = k[o * so + i * si] (host memory)
  static void pack(Context* ctx, const float* k, uint32_t n_out, uint32_t n_in, uint32_t so, uint32_t si, Packed& out);
  static void launch(Context* ctx, bool id_first, const Packed& f, float* res, const float* rhs, uint32_t n_out,
                     uint32_t n_in, size_t d, bool set);
};

// LinearOperator: include/prost/linop/linearoperator.hpp:36-90
class LinearOperator {
 public:
  explicit LinearOperator(Context* ctx) : ctx_(ctx) {}
  void add_block(std::shared_ptr<Block> b) { blocks_.push_back(std::move(b)); }
  void initialize();                                         // sizes + overlap check
  void eval(float* d_result, const float* d_rhs, float beta, bool transpose, bool negate = false);
  float row_sum(size_t row, float alpha) const;
  float col_sum(size_t col, float alpha) const;
  void row_sums(float alpha, std::vector<float>& out) const;  // all rows, one sweep per block
  void col_sums(float alpha, std::vector<float>& out) const;
  size_t nrows() const { return nrows_; }
  size_t ncols() const { return ncols_; }
  size_t gpu_mem_amount() const;
  const std::vector<std::shared_ptr<Block>>& blocks() const { return blocks_; }
  bool all_stencil() const;
  Context* ctx() const { return ctx_; }

 private:
  Context* ctx_;
  std::vector<std::shared_ptr<Block>> blocks_;
  size_t nrows_ = 0, ncols_ = 0;
  // the blocks' output ranges are pairwise disjoint (rows: forward, columns: adjoint): a block may then overwrite
  // its outputs when beta = 0 (Block::eval_local_set)
  bool disjoint_rows_ = false, disjoint_cols_ = false;
};

}  // namespace pb
