// pb_linop.cu -- block objects and the unfused operator applies (LinearOperator::Eval /
// EvalAdjoint of the reference, linearoperator.cu:134-170).  The fused PDHG passes in
// pb_fused.cu evaluate the same pointwise products from pb_linop.cuh directly.
#include "pb_linop.cuh"

#include <algorithm>
#include <cmath>
#include <numeric>

namespace pb {

// ---- kernels ----------------------------------------------------------------------------------

// res[i] += (K rhs)[i] or (K^T rhs)[i], one thread per output element.
template <bool kTranspose>
__global__ void __launch_bounds__(kBlock) block_apply_add_kernel(BlockDesc b, float* __restrict__ res,
                                                                  const float* __restrict__ rhs,
                                                                  float sign, const int* __restrict__ skip) {
  if (skip && *skip) return;
  const uint32_t n = kTranspose ? b.ncols : b.nrows;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float v = kTranspose ? block_col_dot(b, i, rhs) : block_row_dot(b, i, rhs);
    res[i] += sign * v;
  }
}

__global__ void __launch_bounds__(kBlock) scale_kernel(float* __restrict__ v, size_t n, float beta,
                                                       const int* __restrict__ skip) {
  if (skip && *skip) return;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    v[i] = beta * v[i];
}

// CSR SpMV, one warp per row: lanes stride the row's nonzeros (coalesced val/ind reads),
// shuffle-reduce, lane 0 accumulates.  res[r] += sign * sum_k val[k] * x[ind[k]].
__global__ void __launch_bounds__(kBlock) csr_spmv_add_kernel(const int* __restrict__ ptr,
                                                               const int* __restrict__ ind,
                                                               const float* __restrict__ val,
                                                               uint32_t nrows, float* __restrict__ res,
                                                               const float* __restrict__ x, float sign,
                                                               const int* __restrict__ skip) {
  if (skip && *skip) return;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warps_per_grid = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < nrows; r += warps_per_grid) {
    const int beg = ptr[r], end = ptr[r + 1];
    float acc = 0.f;
    for (int k = beg + lane; k < end; k += 32) acc += val[k] * __ldg(x + ind[k]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if (lane == 0) res[r] += sign * acc;
  }
}

// CSR SpMV for short rows: `kGroup` lanes per row, two rows per group and step.  With uniformly random columns the
// kernel is bound by the L2 gather of x (one 32-byte sector per nonzero, profiles/r02_operators.md), i.e. by how many
// gathers are in flight: a lane loads the (column, value) pairs of up to kUnroll of its entries of BOTH rows before
// the first gather is issued, so 2 * kUnroll independent gathers leave each thread back to back (round 1's form had
// one dependent ptr -> ind -> x chain per thread).
template <int kGroup, int kUnroll, int kRows>
__global__ void __launch_bounds__(kBlock) csr_spmv_add_group_kernel(
    const int* __restrict__ ptr, const int* __restrict__ ind, const float* __restrict__ val,
    uint32_t nrows, float* __restrict__ res, const float* __restrict__ x, float sign,
    const int* __restrict__ skip) {
  if (skip && *skip) return;
  const uint32_t sub = threadIdx.x % kGroup;
  const uint32_t groups_per_grid = (gridDim.x * blockDim.x) / kGroup;
  // all lanes of a warp iterate the same number of times so the shuffles stay converged
  const uint32_t first = (blockIdx.x * blockDim.x + threadIdx.x) / kGroup;
  const uint32_t iters = (nrows + kRows * groups_per_grid - 1) / (kRows * groups_per_grid);
  for (uint32_t it = 0; it < iters; ++it) {
    uint32_t r[kRows];
    int k[kRows], end[kRows];
    float acc[kRows], old[kRows];
    bool more = false;
#pragma unroll
    for (int j = 0; j < kRows; ++j) {
      r[j] = first + (kRows * it + j) * groups_per_grid;
      k[j] = end[j] = 0;
      acc[j] = old[j] = 0.f;
      if (r[j] < nrows) {
        k[j] = __ldg(ptr + r[j]) + (int)sub;
        end[j] = __ldg(ptr + r[j] + 1);
        if (sub == 0) old[j] = res[r[j]];
      }
      more |= k[j] < end[j];
    }
    while (more) {
      int c[kRows][kUnroll];
      float v[kRows][kUnroll];
#pragma unroll
      for (int j = 0; j < kRows; ++j)
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          const int kk = k[j] + u * kGroup;
          const bool ok = kk < end[j];
          c[j][u] = ok ? __ldcs(ind + kk) : 0;
          v[j][u] = ok ? __ldcs(val + kk) : 0.f;
        }
      more = false;
#pragma unroll
      for (int j = 0; j < kRows; ++j) {
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) acc[j] += v[j][u] * __ldg(x + c[j][u]);
        k[j] += kUnroll * kGroup;
        more |= k[j] < end[j];
      }
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < kRows; ++j) {
#pragma unroll
      for (int o = kGroup / 2; o > 0; o >>= 1) acc[j] += __shfl_down_sync(0xffffffffu, acc[j], o, kGroup);
      if (sub == 0 && r[j] < nrows) res[r[j]] = old[j] + sign * acc[j];
    }
  }
}

// Dense y += A x, A column-major (lda = M).  Thread per row (coalesced down a column),
// columns split over blockIdx.y; partials land in `part[split][M]` and are folded in a fixed
// order by dense_fold_kernel, so the result is deterministic.
__global__ void __launch_bounds__(kBlock) dense_gemv_n_kernel(const float* __restrict__ A, uint32_t M,
                                                               uint32_t N, uint32_t cols_per_split,
                                                               const float* __restrict__ x,
                                                               float* __restrict__ part,
                                                               const int* __restrict__ skip) {
  if (skip && *skip) return;
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= M) return;
  const uint32_t c0 = blockIdx.y * cols_per_split;
  const uint32_t c1 = min(N, c0 + cols_per_split);
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
  uint32_t c = c0;
  for (; c + 3 < c1; c += 4) {
    acc0 += A[(size_t)c * M + r] * x[c];
    acc1 += A[(size_t)(c + 1) * M + r] * x[c + 1];
    acc2 += A[(size_t)(c + 2) * M + r] * x[c + 2];
    acc3 += A[(size_t)(c + 3) * M + r] * x[c + 3];
  }
  for (; c < c1; ++c) acc0 += A[(size_t)c * M + r] * x[c];
  part[(size_t)blockIdx.y * M + r] = (acc0 + acc1) + (acc2 + acc3);
}

__global__ void __launch_bounds__(kBlock) dense_fold_kernel(const float* __restrict__ part, uint32_t M,
                                                             uint32_t splits, float* __restrict__ res,
                                                             float sign, const int* __restrict__ skip) {
  if (skip && *skip) return;
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= M) return;
  float acc = 0.f;
  for (uint32_t s = 0; s < splits; ++s) acc += part[(size_t)s * M + r];
  res[r] += sign * acc;
}

// Dense y += A^T x: one warp per column (contiguous in memory), shuffle-reduce.
__global__ void __launch_bounds__(kBlock) dense_gemv_t_kernel(const float* __restrict__ A, uint32_t M,
                                                               uint32_t N, const float* __restrict__ x,
                                                               float* __restrict__ res, float sign,
                                                               const int* __restrict__ skip) {
  if (skip && *skip) return;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warps_per_grid = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < N; c += warps_per_grid) {
    const float* col = A + (size_t)c * M;
    float acc0 = 0.f, acc1 = 0.f;
    uint32_t r = lane;
    for (; r + 32 < M; r += 64) {
      acc0 += col[r] * x[r];
      acc1 += col[r + 32] * x[r + 32];
    }
    if (r < M) acc0 += col[r] * x[r];
    float acc = acc0 + acc1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if (lane == 0) res[c] += sign * acc;
  }
}

// ---- Block base -------------------------------------------------------------------------------

static uint32_t checked_u32(size_t v, const char* what) {
  if (v >= (1ull << 31)) fail(PB_ERR_UNSUPPORTED, std::string(what) + " exceeds 2^31-1 elements");
  return static_cast<uint32_t>(v);
}

BlockDesc Block::desc() const {
  BlockDesc d;
  d.kind = kind();
  d.row = checked_u32(row_, "block row offset");
  d.col = checked_u32(col_, "block column offset");
  d.nrows = checked_u32(nrows_, "block rows");
  d.ncols = checked_u32(ncols_, "block columns");
  return d;
}

static void launch_generic_add(Context* ctx, const BlockDesc& d, bool transpose, float* res,
                               const float* rhs, float sign) {
  const size_t n = transpose ? d.ncols : d.nrows;
  if (n == 0) return;
  const unsigned grid = std::min<size_t>(grid_for(n), (size_t)ctx->num_sms * 32);
  if (transpose)
    block_apply_add_kernel<true><<<grid, kBlock, 0, ctx->stream>>>(d, res, rhs, sign, ctx->skip_flag);
  else
    block_apply_add_kernel<false><<<grid, kBlock, 0, ctx->stream>>>(d, res, rhs, sign, ctx->skip_flag);
  PB_CHECK_LAUNCH();
  ctx->launches++;
}

// ---- gradient (2-D and 3-D) -----------------------------------------------------------------------

class BlockGradient : public Block {
 public:
  BlockGradient(Context* ctx, bool three_d, size_t row, size_t col, size_t nx, size_t ny, size_t L,
                bool label_first)
      : Block(ctx, row, col, nx * ny * L * (three_d ? 3 : 2), nx * ny * L),
        three_d_(three_d), nx_(nx), ny_(ny), L_(L), label_first_(label_first) {
    if (nx == 0 || ny == 0 || L == 0) fail(PB_ERR_INVALID, "BlockGradient: empty grid");
    checked_u32(nrows_, "gradient block rows");
  }
  int kind() const override { return three_d_ ? kBlockGradient3D : kBlockGradient2D; }
  // constants, independent of the boundary rows (block_gradient2d.cu:153-163, 3d:164-174)
  float row_sum(size_t, float) const override { return 2.f; }
  float col_sum(size_t, float) const override { return three_d_ ? 6.f : 4.f; }
  bool uniform_sums() const override { return true; }

  BlockDesc desc() const override {
    BlockDesc d = Block::desc();
    d.label_first = label_first_ ? 1 : 0;
    d.nx = (uint32_t)nx_; d.ny = (uint32_t)ny_; d.L = (uint32_t)L_;
    d.plane = (uint32_t)(nx_ * ny_ * L_);
    d.div_ny = FastDiv(ny_);
    d.div_L = FastDiv(L_);
    d.div_nxny = FastDiv(nx_ * ny_);
    d.div_nyL = FastDiv(ny_ * L_);
    d.div_plane = FastDiv(nx_ * ny_ * L_);
    return d;
  }
  void eval_local_add(float* res, const float* rhs) override {
    launch_generic_add(ctx_, desc(), false, res, rhs, 1.f);
  }
  void eval_adjoint_local_add(float* res, const float* rhs) override {
    launch_generic_add(ctx_, desc(), true, res, rhs, 1.f);
  }

 private:
  bool three_d_;
  size_t nx_, ny_, L_;
  bool label_first_;
};

// ---- diagonals ------------------------------------------------------------------------------------

class BlockDiags : public Block {
 public:
  BlockDiags(Context* ctx, size_t row, size_t col, size_t nrows, size_t ncols, size_t ndiags,
             const int64_t* offsets, const float* factors)
      : Block(ctx, row, col, nrows, ncols) {
    // sort by offset (the forward loop stops at the first out-of-range column and relies on
    // ascending offsets: block_diags.cu:58-59,110-118); factors are always float (hpp:82)
    std::vector<size_t> order(ndiags);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(),
                     [&](size_t a, size_t b) { return offsets[a] < offsets[b]; });
    offsets_.resize(ndiags);
    factors_.resize(ndiags);
    for (size_t i = 0; i < ndiags; ++i) {
      offsets_[i] = offsets[order[i]];
      factors_[i] = factors[order[i]];
    }
    d_offsets_.assign(offsets_, ctx->stream);
    d_factors_.assign(factors_, ctx->stream);
  }
  int kind() const override { return kBlockDiags; }
  // square block whose diagonals all sit at offset 0 (identity-type blocks, +block/identity.m:12-13):
  // every row and every column sees every diagonal, so the sums do not depend on the index
  bool uniform_sums() const override {
    if (nrows_ != ncols_ || nrows_ == 0) return false;
    for (long long o : offsets_)
      if (o != 0) return false;
    return true;
  }
  float row_sum(size_t row, float alpha) const override {
    float sum = 0;
    for (size_t i = 0; i < offsets_.size(); ++i) {
      const long long c = (long long)row + offsets_[i];
      if (c < 0) continue;
      if (c >= (long long)ncols_) break;
      sum += std::pow(std::abs(factors_[i]), alpha);
    }
    return sum;
  }
  float col_sum(size_t col, float alpha) const override {
    float sum = 0;
    for (size_t i = 0; i < offsets_.size(); ++i) {
      const long long ofs = offsets_[i];
      if (ofs > (long long)col) break;
      if ((long long)col - ofs < (long long)nrows_) sum += std::pow(std::abs(factors_[i]), alpha);
    }
    return sum;
  }
  size_t gpu_mem_amount() const override {
    return offsets_.size() * (sizeof(long long) + sizeof(float));
  }
  BlockDesc desc() const override {
    BlockDesc d = Block::desc();
    d.ndiags = (int)offsets_.size();
    d.offsets = d_offsets_.data();
    d.factors = d_factors_.data();
    return d;
  }
  void eval_local_add(float* res, const float* rhs) override {
    launch_generic_add(ctx_, desc(), false, res, rhs, 1.f);
  }
  void eval_adjoint_local_add(float* res, const float* rhs) override {
    // NOTE: the reference sizes this launch by nrows (block_diags.cu:211) and silently skips
    // columns >= nrows; all ncols columns are computed here.
    launch_generic_add(ctx_, desc(), true, res, rhs, 1.f);
  }

 private:
  std::vector<long long> offsets_;
  std::vector<float> factors_;
  DeviceBuffer<long long> d_offsets_;
  DeviceBuffer<float> d_factors_;
};

// ---- sparse ---------------------------------------------------------------------------------------

class BlockSparse : public Block {
 public:
  BlockSparse(Context* ctx, size_t row, size_t col, int m, int n, int nnz, const float* val,
              const int32_t* ptr, const int32_t* ind)
      : Block(ctx, row, col, (size_t)m, (size_t)n), nnz_(nnz) {
    if (m < 0 || n < 0 || nnz < 0) fail(PB_ERR_INVALID, "BlockSparse: negative size");
    if (ptr[0] != 0 || ptr[n] != nnz) fail(PB_ERR_INVALID, "BlockSparse: inconsistent CSC column pointer");
    // CSC of K == CSR of K^T
    ptr_t_.assign(ptr, ptr + n + 1);
    ind_t_.assign(ind, ind + nnz);
    val_t_.assign(val, val + nnz);
    for (int k = 0; k < nnz; ++k)
      if (ind[k] < 0 || ind[k] >= m) fail(PB_ERR_INVALID, "BlockSparse: row index out of range");
    // CSR of K by a counting-sort transpose; within a row the entries come out in ascending
    // column order, like the reference's host csr2csc (common.cu:54-82)
    ptr_.assign(m + 1, 0);
    for (int k = 0; k < nnz; ++k) ptr_[ind[k] + 1]++;
    for (int r = 0; r < m; ++r) ptr_[r + 1] += ptr_[r];
    ind_.resize(nnz);
    val_.resize(nnz);
    std::vector<int> fill(ptr_.begin(), ptr_.end() - 1);
    for (int c = 0; c < n; ++c)
      for (int k = ptr[c]; k < ptr[c + 1]; ++k) {
        const int dst = fill[ind[k]]++;
        ind_[dst] = c;
        val_[dst] = val[k];
      }
    d_ptr_.assign(ptr_, ctx->stream);
    d_ind_.assign(ind_, ctx->stream);
    d_val_.assign(val_, ctx->stream);
    d_ptr_t_.assign(ptr_t_, ctx->stream);
    d_ind_t_.assign(ind_t_, ctx->stream);
    d_val_t_.assign(val_t_, ctx->stream);
  }
  int kind() const override { return kBlockSparse; }
  float row_sum(size_t row, float alpha) const override {
    float sum = 0;
    for (int k = ptr_[row]; k < ptr_[row + 1]; ++k) sum += std::pow(std::abs(val_[k]), alpha);
    return sum;
  }
  float col_sum(size_t col, float alpha) const override {
    float sum = 0;
    for (int k = ptr_t_[col]; k < ptr_t_[col + 1]; ++k) sum += std::pow(std::abs(val_t_[k]), alpha);
    return sum;
  }
  size_t gpu_mem_amount() const override {
    return 2 * (size_t)nnz_ * (sizeof(int32_t) + sizeof(float)) +
           (nrows_ + ncols_ + 2) * sizeof(int32_t);
  }
  BlockDesc desc() const override {
    BlockDesc d = Block::desc();
    d.ptr = d_ptr_.data(); d.ind = d_ind_.data(); d.val = d_val_.data();
    d.ptr_t = d_ptr_t_.data(); d.ind_t = d_ind_t_.data(); d.val_t = d_val_t_.data();
    return d;
  }
  void spmv(const DeviceBuffer<int>& ptr, const DeviceBuffer<int>& ind, const DeviceBuffer<float>& val,
            size_t rows, float* res, const float* x) {
    if (rows == 0) return;
    const double avg = rows ? (double)nnz_ / (double)rows : 0.0;
    const unsigned cap = ctx_->num_sms * 16;
    // lanes per row / entries per lane / rows per step: 3 or 4 rows per step, 2 or 8 lanes per row at 12 nnz and 8
    // lanes at 48 nnz all measure within 4 % of these (profiles/r02_operators.md)
#define PB_SPMV(G, U, R)                                                                                           \
  do {                                                                                                             \
    const unsigned grid = std::min<size_t>(grid_for((rows * G + R - 1) / R), cap);                                 \
    csr_spmv_add_group_kernel<G, U, R><<<grid, kBlock, 0, ctx_->stream>>>(ptr.data(), ind.data(), val.data(),      \
                                                                         (uint32_t)rows, res, x, 1.f, ctx_->skip_flag); \
  } while (0)
    if (avg <= 6.0) {
      PB_SPMV(4, 2, 2);
    } else if (avg <= 16.0) {
      PB_SPMV(4, 4, 2);
    } else if (avg <= 32.0) {
      PB_SPMV(8, 4, 2);
    } else if (avg <= 128.0) {
      PB_SPMV(16, 4, 2);
    } else {
      const unsigned grid = std::min<size_t>(grid_for(rows * 32), cap);
      csr_spmv_add_kernel<<<grid, kBlock, 0, ctx_->stream>>>(ptr.data(), ind.data(), val.data(),
                                                             (uint32_t)rows, res, x, 1.f, ctx_->skip_flag);
    }
#undef PB_SPMV
    PB_CHECK_LAUNCH();
    ctx_->launches++;
  }
  void eval_local_add(float* res, const float* rhs) override {
    spmv(d_ptr_, d_ind_, d_val_, nrows_, res, rhs);
  }
  void eval_adjoint_local_add(float* res, const float* rhs) override {
    spmv(d_ptr_t_, d_ind_t_, d_val_t_, ncols_, res, rhs);
  }

 private:
  int nnz_;
  std::vector<int> ptr_, ind_, ptr_t_, ind_t_;
  std::vector<float> val_, val_t_;
  DeviceBuffer<int> d_ptr_, d_ind_, d_ptr_t_, d_ind_t_;
  DeviceBuffer<float> d_val_, d_val_t_;
};

// ---- dense ----------------------------------------------------------------------------------------

class BlockDense : public Block {
 public:
  BlockDense(Context* ctx, size_t row, size_t col, size_t nrows, size_t ncols, const float* data)
      : Block(ctx, row, col, nrows, ncols), host_(data, data + nrows * ncols) {
    checked_u32(nrows, "dense rows");
    checked_u32(ncols, "dense columns");
    d_data_.assign(host_, ctx->stream);
    // column splits so that a 4096x4096 block still fills the machine
    const size_t row_ctas = (nrows + kBlock - 1) / kBlock;
    size_t want = ((size_t)ctx->num_sms * 4 + row_ctas - 1) / std::max<size_t>(row_ctas, 1);
    splits_ = (uint32_t)std::max<size_t>(1, std::min<size_t>(want, std::max<size_t>(1, ncols / 64)));
    cols_per_split_ = (uint32_t)((ncols + splits_ - 1) / splits_);
    splits_ = (uint32_t)((ncols + cols_per_split_ - 1) / std::max<uint32_t>(cols_per_split_, 1));
    if (splits_ == 0) splits_ = 1;
    d_part_.resize((size_t)splits_ * nrows);
  }
  int kind() const override { return kBlockDense; }
  float row_sum(size_t row, float alpha) const override {
    float sum = 0;
    for (size_t c = 0; c < ncols_; ++c) sum += std::pow(std::abs(host_[c * nrows_ + row]), alpha);
    return sum;
  }
  float col_sum(size_t col, float alpha) const override {
    float sum = 0;
    for (size_t r = 0; r < nrows_; ++r) sum += std::pow(std::abs(host_[col * nrows_ + r]), alpha);
    return sum;
  }
  size_t gpu_mem_amount() const override { return nrows_ * ncols_ * sizeof(float); }
  BlockDesc desc() const override {
    BlockDesc d = Block::desc();
    d.dense = d_data_.data();
    return d;
  }
  void eval_local_add(float* res, const float* rhs) override {
    if (nrows_ == 0 || ncols_ == 0) return;
    dim3 grid(grid_for(nrows_), splits_);
    dense_gemv_n_kernel<<<grid, kBlock, 0, ctx_->stream>>>(d_data_.data(), (uint32_t)nrows_,
                                                           (uint32_t)ncols_, cols_per_split_, rhs,
                                                           d_part_.data(), ctx_->skip_flag);
    PB_CHECK_LAUNCH();
    dense_fold_kernel<<<grid_for(nrows_), kBlock, 0, ctx_->stream>>>(d_part_.data(), (uint32_t)nrows_,
                                                                     splits_, res, 1.f, ctx_->skip_flag);
    PB_CHECK_LAUNCH();
    ctx_->launches += 2;
  }
  void eval_adjoint_local_add(float* res, const float* rhs) override {
    if (nrows_ == 0 || ncols_ == 0) return;
    const unsigned grid = std::min<size_t>(grid_for(ncols_ * 32), (size_t)ctx_->num_sms * 16);
    dense_gemv_t_kernel<<<grid, kBlock, 0, ctx_->stream>>>(d_data_.data(), (uint32_t)nrows_,
                                                           (uint32_t)ncols_, rhs, res, 1.f, ctx_->skip_flag);
    PB_CHECK_LAUNCH();
    ctx_->launches++;
  }

 private:
  std::vector<float> host_;
  DeviceBuffer<float> d_data_, d_part_;
  uint32_t splits_ = 1, cols_per_split_ = 0;
};

// ---- Kronecker products with an identity (block_dense_kron_id.cu, block_id_kron_dense.cu) --------------
// K is a small n_out x n_in matrix read through strides (so, si) so that one kernel serves K and K^T.  All four
// products stream vectors of d * n floats with d in the millions: they are HBM bound as long as every vector element
// crosses HBM once, so the kernels are organised around that (round 1's one-thread-per-output forms re-read the
// input once per output row or per 8 rows and ran at 3 - 16 % of the HBM peak, profiles/r02_operators.md):
//
// kron(K, I_d):  res[o*d + k] (+)= sum_i K(o, i) rhs[i*d + k]  -- the GEMM  Res (n_out x d) = K X (n_in x d).
//   A thread owns VEC consecutive columns k (128-bit loads / stores) and kKronRowTile output rows in registers; the
//   CTA stages its rows of K in shared memory once (read back as 128-bit broadcasts), X streams through once per
//   row tile (once in total for n_out <= 16).  Same summation order as the reference (i ascending, then += / =).
constexpr int kKronRowTile = 16;
constexpr int kKronSmemFloats = 8192;       // staged factor entries per CTA (32 KB)

template <int VEC, bool SET>
__global__ void __launch_bounds__(kBlock) kron_k_id_kernel(float* __restrict__ res, const float* __restrict__ rhs,
                                                           const float* __restrict__ K, uint32_t n_out, uint32_t n_in,
                                                           size_t d, uint32_t so, uint32_t si, bool k_in_smem,
                                                           const int* __restrict__ skip) {
  if (skip && *skip) return;
  extern __shared__ __align__(16) float kron_smem[];          // Ks[i][r], r = row of the tile (fastest)
  const uint32_t o0 = blockIdx.y * kKronRowTile;
  if (k_in_smem) {
    for (uint32_t e = threadIdx.x; e < n_in * kKronRowTile; e += blockDim.x) {
      const uint32_t i = e / kKronRowTile, r = e % kKronRowTile;
      kron_smem[e] = (o0 + r < n_out) ? K[(size_t)(o0 + r) * so + (size_t)i * si] : 0.f;
    }
    __syncthreads();
  }
  const size_t k = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * VEC;
  if (k >= d) return;
  float acc[kKronRowTile][VEC];
#pragma unroll
  for (int r = 0; r < kKronRowTile; ++r)
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[r][v] = 0.f;
  for (uint32_t i = 0; i < n_in; ++i) {
    float x[VEC];
    if (VEC == 4) {
      const float4 t = *reinterpret_cast<const float4*>(rhs + (size_t)i * d + k);
      x[0] = t.x; x[1 % VEC] = t.y; x[2 % VEC] = t.z; x[3 % VEC] = t.w;
    } else {
      x[0] = rhs[(size_t)i * d + k];
    }
    float kr[kKronRowTile];
    if (k_in_smem) {
#pragma unroll
      for (int r = 0; r < kKronRowTile; r += 4) {
        const float4 t = *reinterpret_cast<const float4*>(kron_smem + i * kKronRowTile + r);
        kr[r] = t.x; kr[r + 1] = t.y; kr[r + 2] = t.z; kr[r + 3] = t.w;
      }
    } else {
#pragma unroll
      for (int r = 0; r < kKronRowTile; ++r)
        kr[r] = (o0 + r < n_out) ? __ldg(K + (size_t)(o0 + r) * so + (size_t)i * si) : 0.f;
    }
#pragma unroll
    for (int r = 0; r < kKronRowTile; ++r)
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[r][v] += kr[r] * x[v];
  }
#pragma unroll
  for (int r = 0; r < kKronRowTile; ++r) {
    if (o0 + r >= n_out) break;
    float* out = res + (size_t)(o0 + r) * d + k;
    if (VEC == 4) {
      float4 t = SET ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<const float4*>(out);
      t.x += acc[r][0]; t.y += acc[r][1 % VEC]; t.z += acc[r][2 % VEC]; t.w += acc[r][3 % VEC];
      *reinterpret_cast<float4*>(out) = t;
    } else {
      out[0] = SET ? acc[r][0] : out[0] + acc[r][0];
    }
  }
}

// kron(I_d, K):  res[b*n_out + o] (+)= sum_i K(o, i) rhs[b*n_in + i]  -- d independent small products on contiguous
// segments.  A CTA takes kIdTileB consecutive segments: their inputs are ONE contiguous run of memory, staged into
// shared memory with coalesced loads next to the factor; thread (segment, o) then reads its K row (padded rows: no
// bank conflicts between the o of a warp) and the segment's inputs as broadcasts, and the outputs of the tile are
// again one contiguous run.
// (warp tiles: every warp stages, computes and stores its own run of segments -- no CTA-wide barrier in the loop)
template <bool SET>
__global__ void __launch_bounds__(kBlock) kron_id_k_kernel(float* __restrict__ res, const float* __restrict__ rhs,
                                                           const float* __restrict__ K, uint32_t n_out, uint32_t n_in,
                                                           size_t d, uint32_t so, uint32_t si, uint32_t tile_b,
                                                           const int* __restrict__ skip) {
  if (skip && *skip) return;
  extern __shared__ __align__(16) float kron_smem[];
  const uint32_t pitch = (n_out + 3u) & ~3u;                   // Ks[i][o], o fastest, rows padded to 4 (zeros)
  const uint32_t groups = pitch / 4;                           // a lane owns 4 consecutive outputs of one segment
  const uint32_t xp = n_in + 1;                                // padded segment pitch: conflict-free across segments
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
  float* Ks = kron_smem;
  float* xs = kron_smem + (size_t)n_in * pitch + (size_t)warp * tile_b * xp;      // this warp's xs[segment][i]
  for (uint32_t e = threadIdx.x; e < n_in * pitch; e += blockDim.x) {
    const uint32_t i = e / pitch, o = e % pitch;
    Ks[e] = o < n_out ? K[(size_t)o * so + (size_t)i * si] : 0.f;
  }
  __syncthreads();
  const size_t stride = (size_t)gridDim.x * warps * tile_b;
  for (size_t b0 = ((size_t)blockIdx.x * warps + warp) * tile_b; b0 < d; b0 += stride) {
    const uint32_t nb = (uint32_t)min((size_t)tile_b, d - b0);
    __syncwarp();                                              // previous tile consumed
    for (uint32_t e = lane; e < nb * n_in; e += 32) xs[(e / n_in) * xp + e % n_in] = rhs[b0 * n_in + e];
    __syncwarp();
    for (uint32_t e = lane; e < nb * groups; e += 32) {
      const uint32_t bl = e / groups, o0 = (e % groups) * 4;
      const float* x = xs + bl * xp;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      for (uint32_t i = 0; i < n_in; ++i) {                    // one 128-bit K read + one x read per 4 FMAs
        const float4 k4 = *reinterpret_cast<const float4*>(Ks + i * pitch + o0);
        const float xv = x[i];
        s0 += k4.x * xv; s1 += k4.y * xv; s2 += k4.z * xv; s3 += k4.w * xv;
      }
      float* out = res + (b0 + bl) * n_out + o0;
      const float sv[4] = {s0, s1, s2, s3};
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (o0 + j < n_out) out[j] = SET ? sv[j] : out[j] + sv[j];
    }
  }
}

// fall-back for factors that do not fit in shared memory: one thread per output element
template <bool SET>
__global__ void __launch_bounds__(kBlock) kron_id_k_big_kernel(float* __restrict__ res, const float* __restrict__ rhs,
                                                               const float* __restrict__ K, uint32_t n_out,
                                                               uint32_t n_in, size_t d, uint32_t so, uint32_t si,
                                                               const int* __restrict__ skip) {
  if (skip && *skip) return;
  const size_t total = d * n_out;
  for (size_t tx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; tx < total; tx += (size_t)gridDim.x * blockDim.x) {
    const size_t b = tx / n_out;
    const uint32_t o = (uint32_t)(tx - b * n_out);
    const float* x = rhs + b * n_in;
    float sum = 0.f;
    for (uint32_t i = 0; i < n_in; ++i) sum += __ldg(K + (size_t)o * so + (size_t)i * si) * x[i];
    res[tx] = SET ? sum : res[tx] + sum;
  }
}

class BlockDenseKron : public Block {
 public:
  BlockDenseKron(Context* ctx, bool id_first, size_t diaglength, size_t row, size_t col, size_t mr, size_t mc,
                 const float* data)
      : Block(ctx, row, col, mr * diaglength, mc * diaglength), id_first_(id_first), d_(diaglength), mr_(mr), mc_(mc),
        host_(data, data + mr * mc) {
    checked_u32(mr, "Kronecker factor rows");
    checked_u32(mc, "Kronecker factor columns");
    if (diaglength == 0) fail(PB_ERR_INVALID, "Kronecker block: diaglength must be positive");
    d_data_.assign(host_, ctx->stream);
  }
  int kind() const override { return id_first_ ? kBlockIdKronDense : kBlockDenseKronId; }
  // block_dense_kron_id.cu:100-121 (row / diaglength), block_id_kron_dense.cu (row % mat_nrows)
  float row_sum(size_t row, float alpha) const override {
    const size_t r = id_first_ ? row % mr_ : row / d_;
    float sum = 0;
    for (size_t i = 0; i < mc_; ++i) sum += std::pow(std::abs(host_[i * mr_ + r]), alpha);
    return sum;
  }
  float col_sum(size_t col, float alpha) const override {
    const size_t c = id_first_ ? col % mc_ : col / d_;
    float sum = 0;
    for (size_t i = 0; i < mr_; ++i) sum += std::pow(std::abs(host_[i + c * mr_]), alpha);
    return sum;
  }
  size_t gpu_mem_amount() const override { return host_.size() * sizeof(float); }
  void eval_local_add(float* res, const float* rhs) override { apply(res, rhs, false, false); }
  void eval_adjoint_local_add(float* res, const float* rhs) override { apply(res, rhs, true, false); }
  bool eval_local_set(float* res, const float* rhs) override { apply(res, rhs, false, true); return true; }
  bool eval_adjoint_local_set(float* res, const float* rhs) override { apply(res, rhs, true, true); return true; }

 private:
  template <class Kernel>
  static void allow_smem(Kernel kernel, size_t bytes) {
    if (bytes > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  }
  void apply(float* res, const float* rhs, bool transpose, bool set) {
    if (mr_ == 0 || mc_ == 0) return;
    // K(o, i): column-major K[i*mr + o]; transposed K^T(o, i) = K[o*mr + i]
    const uint32_t n_out = (uint32_t)(transpose ? mc_ : mr_), n_in = (uint32_t)(transpose ? mr_ : mc_);
    const uint32_t so = transpose ? (uint32_t)mr_ : 1u, si = transpose ? 1u : (uint32_t)mr_;
    cudaStream_t st = ctx_->stream;
    const float* K = d_data_.data();
    const int* skip = ctx_->skip_flag;
    if (KronTensorCore::supported(id_first_, n_out, n_in, d_, res, rhs)) {
      KronTensorCore::Packed& f = tc_[transpose ? 1 : 0];
      if (!f.ready) KronTensorCore::pack(ctx_, host_.data(), n_out, n_in, so, si, f);
      KronTensorCore::launch(ctx_, id_first_, f, res, rhs, n_out, n_in, d_, set);
      PB_CHECK_LAUNCH();
      ctx_->launches++;
      return;
    }
    if (!id_first_) {
      const bool vec4 = d_ % 4 == 0 && ((reinterpret_cast<uintptr_t>(res) | reinterpret_cast<uintptr_t>(rhs)) & 15u) == 0;
      const size_t threads = vec4 ? d_ / 4 : d_;
      const dim3 grid(grid_for(threads), (n_out + kKronRowTile - 1) / kKronRowTile);
      if (grid.y > 65535u) fail(PB_ERR_UNSUPPORTED, "Kronecker block: factor too large for this kernel");
      const bool in_smem = (size_t)n_in * kKronRowTile <= (size_t)kKronSmemFloats;
      const size_t smem = in_smem ? (size_t)n_in * kKronRowTile * sizeof(float) : 0;
#define PB_KRON(V, S) kron_k_id_kernel<V, S><<<grid, kBlock, smem, st>>>(res, rhs, K, n_out, n_in, d_, so, si, in_smem, skip)
      if (vec4) { if (set) PB_KRON(4, true); else PB_KRON(4, false); }
      else { if (set) PB_KRON(1, true); else PB_KRON(1, false); }
#undef PB_KRON
    } else {
      const size_t k_floats = (size_t)n_in * ((n_out + 3u) & ~3u);
      // segments per tile: about one output per thread and pass, bounded by the shared memory left for the inputs
      // segments per warp tile: up to 32, bounded by ~4 KB of staged inputs per warp
      const size_t warps = kBlock / 32;
      const size_t tile_b = std::max<size_t>(1, std::min<size_t>(32, 1024 / std::max<uint32_t>(n_in, 1u)));
      const size_t smem = (k_floats + warps * tile_b * (n_in + 1)) * sizeof(float);
      if (smem <= 96 * 1024) {
        const size_t tiles = (d_ + tile_b - 1) / tile_b;
        const unsigned grid = (unsigned)std::min<size_t>((tiles + warps - 1) / warps, (size_t)ctx_->num_sms * 8);
        if (set) {
          allow_smem(kron_id_k_kernel<true>, smem);
          kron_id_k_kernel<true><<<grid, kBlock, smem, st>>>(res, rhs, K, n_out, n_in, d_, so, si, (uint32_t)tile_b, skip);
        } else {
          allow_smem(kron_id_k_kernel<false>, smem);
          kron_id_k_kernel<false><<<grid, kBlock, smem, st>>>(res, rhs, K, n_out, n_in, d_, so, si, (uint32_t)tile_b, skip);
        }
      } else {
        const unsigned grid = (unsigned)std::min<size_t>(grid_for(d_ * n_out), (size_t)ctx_->num_sms * 32);
        if (set) kron_id_k_big_kernel<true><<<grid, kBlock, 0, st>>>(res, rhs, K, n_out, n_in, d_, so, si, skip);
        else kron_id_k_big_kernel<false><<<grid, kBlock, 0, st>>>(res, rhs, K, n_out, n_in, d_, so, si, skip);
      }
    }
    PB_CHECK_LAUNCH();
    ctx_->launches++;
  }
  bool id_first_;
  size_t d_, mr_, mc_;
  std::vector<float> host_;
  DeviceBuffer<float> d_data_;
  KronTensorCore::Packed tc_[2];      // forward / adjoint factor for the tensor-core path, packed on first use
};

// ---- the same products for a sparse factor (block_sparse_kron_id.cu:28-52, block_id_kron_sparse.cu) -------
// CSR of the factor (or of its transpose for the adjoint); entries of a row in ascending column order, so the sums
// run in the reference's order.
// kron(K, I_d):  res[o*d + k] (+)= sum_{j in row o} val[j] rhs[ind[j]*d + k].  A thread owns VEC consecutive k of one
// output row (128-bit loads of the referenced input rows, 128-bit store): contiguous d-segments are streamed by a
// warp instead of one element per thread (block_sparse_kron_id.cu:27-52 is the design this replaces).
template <int VEC, bool SET>
__global__ void __launch_bounds__(kBlock) kron_sparse_id_kernel(float* __restrict__ res, const float* __restrict__ rhs,
                                                                const int* __restrict__ ptr, const int* __restrict__ ind,
                                                                const float* __restrict__ val, uint32_t n_out, size_t d,
                                                                const int* __restrict__ skip) {
  if (skip && *skip) return;
  const size_t per_row = d / VEC, total = per_row * n_out;
  for (size_t tx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; tx < total; tx += (size_t)gridDim.x * blockDim.x) {
    const size_t o = tx / per_row, k = (tx - o * per_row) * VEC;
    float sum[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) sum[v] = 0.f;
    const int stop = ptr[o + 1];
    for (int j = ptr[o]; j < stop; ++j) {
      const float a = val[j];
      const float* x = rhs + (size_t)ind[j] * d + k;
      if (VEC == 4) {
        const float4 t = *reinterpret_cast<const float4*>(x);
        sum[0] += a * t.x; sum[1 % VEC] += a * t.y; sum[2 % VEC] += a * t.z; sum[3 % VEC] += a * t.w;
      } else {
        sum[0] += a * x[0];
      }
    }
    float* out = res + o * d + k;
    if (VEC == 4) {
      float4 t = SET ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<const float4*>(out);
      t.x += sum[0]; t.y += sum[1 % VEC]; t.z += sum[2 % VEC]; t.w += sum[3 % VEC];
      *reinterpret_cast<float4*>(out) = t;
    } else {
      out[0] = SET ? sum[0] : out[0] + sum[0];
    }
  }
}

// kron(I_d, K):  res[b*n_out + o] (+)= sum_{j in row o} val[j] rhs[b*n_in + ind[j]].  Like the dense form: a CTA stages
// the inputs of tile_b consecutive segments (one contiguous run) in shared memory, the gathers then hit shared
// memory and the tile's outputs are one contiguous run.
template <bool SET>
__global__ void __launch_bounds__(kBlock) kron_id_sparse_kernel(float* __restrict__ res, const float* __restrict__ rhs,
                                                                const int* __restrict__ ptr, const int* __restrict__ ind,
                                                                const float* __restrict__ val, uint32_t n_out,
                                                                uint32_t n_in, size_t d, uint32_t tile_b,
                                                                const int* __restrict__ skip) {
  if (skip && *skip) return;
  extern __shared__ __align__(16) float kron_smem[];
  const uint32_t xp = n_in + 1;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
  float* xs = kron_smem + (size_t)warp * tile_b * xp;          // this warp's xs[segment][i]
  const size_t stride = (size_t)gridDim.x * warps * tile_b;
  for (size_t b0 = ((size_t)blockIdx.x * warps + warp) * tile_b; b0 < d; b0 += stride) {
    const uint32_t nb = (uint32_t)min((size_t)tile_b, d - b0);
    __syncwarp();
    for (uint32_t e = lane; e < nb * n_in; e += 32) xs[(e / n_in) * xp + e % n_in] = rhs[b0 * n_in + e];
    __syncwarp();
    for (uint32_t e = lane; e < nb * n_out; e += 32) {
      const uint32_t bl = e / n_out, o = e % n_out;
      const float* x = xs + bl * xp;
      float sum = 0.f;
      const int stop = ptr[o + 1];
      for (int j = ptr[o]; j < stop; ++j) sum += val[j] * x[ind[j]];
      float* out = res + b0 * n_out + e;
      *out = SET ? sum : *out + sum;
    }
  }
}

// fall-back when a segment tile does not fit in shared memory
template <bool SET>
__global__ void __launch_bounds__(kBlock) kron_id_sparse_big_kernel(float* __restrict__ res,
                                                                    const float* __restrict__ rhs,
                                                                    const int* __restrict__ ptr,
                                                                    const int* __restrict__ ind,
                                                                    const float* __restrict__ val, uint32_t n_out,
                                                                    uint32_t n_in, size_t d,
                                                                    const int* __restrict__ skip) {
  if (skip && *skip) return;
  const size_t total = d * n_out;
  for (size_t tx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; tx < total; tx += (size_t)gridDim.x * blockDim.x) {
    const size_t b = tx / n_out, o = tx - b * n_out;
    const float* x = rhs + b * n_in;
    float sum = 0.f;
    const int stop = ptr[o + 1];
    for (int j = ptr[o]; j < stop; ++j) sum += val[j] * x[ind[j]];
    res[tx] = SET ? sum : res[tx] + sum;
  }
}

class BlockSparseKron : public Block {
 public:
  BlockSparseKron(Context* ctx, bool id_first, size_t diaglength, size_t row, size_t col, int m, int n, int nnz,
                  const float* val, const int32_t* ptr, const int32_t* ind)
      : Block(ctx, row, col, (size_t)std::max(m, 0) * diaglength, (size_t)std::max(n, 0) * diaglength),
        id_first_(id_first), d_(diaglength), m_(m), n_(n) {
    if (m < 0 || n < 0 || nnz < 0) fail(PB_ERR_INVALID, "Kronecker block: negative size");
    if (diaglength == 0) fail(PB_ERR_INVALID, "Kronecker block: diaglength must be positive");
    if (ptr[0] != 0 || ptr[n] != nnz) fail(PB_ERR_INVALID, "Kronecker block: inconsistent CSC column pointer");
    for (int k = 0; k < nnz; ++k)
      if (ind[k] < 0 || ind[k] >= m) fail(PB_ERR_INVALID, "Kronecker block: row index out of range");
    // CSC of K == CSR of K^T; CSR of K by a counting-sort transpose (ascending columns within a row, like the
    // reference's host csr2csc, common.cu:54-82)
    ptr_t_.assign(ptr, ptr + n + 1);
    ind_t_.assign(ind, ind + nnz);
    val_t_.assign(val, val + nnz);
    ptr_.assign(m + 1, 0);
    for (int k = 0; k < nnz; ++k) ptr_[ind[k] + 1]++;
    for (int r = 0; r < m; ++r) ptr_[r + 1] += ptr_[r];
    ind_.resize(nnz);
    val_.resize(nnz);
    std::vector<int> fill(ptr_.begin(), ptr_.end() - 1);
    for (int c = 0; c < n; ++c)
      for (int k = ptr[c]; k < ptr[c + 1]; ++k) {
        const int dst = fill[ind[k]]++;
        ind_[dst] = c;
        val_[dst] = val[k];
      }
    d_ptr_.assign(ptr_, ctx->stream);
    d_ind_.assign(ind_, ctx->stream);
    d_val_.assign(val_, ctx->stream);
    d_ptr_t_.assign(ptr_t_, ctx->stream);
    d_ind_t_.assign(ind_t_, ctx->stream);
    d_val_t_.assign(val_t_, ctx->stream);
  }
  int kind() const override { return id_first_ ? kBlockIdKronSparse : kBlockSparseKronId; }
  // block_sparse_kron_id.cu (row / diaglength), block_id_kron_sparse.cu:127-150 (row % mat_nrows)
  float row_sum(size_t row, float alpha) const override {
    const size_t r = id_first_ ? row % (size_t)m_ : row / d_;
    float sum = 0;
    for (int k = ptr_[r]; k < ptr_[r + 1]; ++k) sum += std::pow(std::abs(val_[k]), alpha);
    return sum;
  }
  float col_sum(size_t col, float alpha) const override {
    const size_t c = id_first_ ? col % (size_t)n_ : col / d_;
    float sum = 0;
    for (int k = ptr_t_[c]; k < ptr_t_[c + 1]; ++k) sum += std::pow(std::abs(val_t_[k]), alpha);
    return sum;
  }
  size_t gpu_mem_amount() const override {
    return 2 * val_.size() * (sizeof(int32_t) + sizeof(float)) + (size_t)(m_ + n_ + 2) * sizeof(int32_t);
  }
  void eval_local_add(float* res, const float* rhs) override { apply(res, rhs, d_ptr_, d_ind_, d_val_, m_, n_, false); }
  void eval_adjoint_local_add(float* res, const float* rhs) override {
    apply(res, rhs, d_ptr_t_, d_ind_t_, d_val_t_, n_, m_, false);
  }
  bool eval_local_set(float* res, const float* rhs) override {
    apply(res, rhs, d_ptr_, d_ind_, d_val_, m_, n_, true);
    return true;
  }
  bool eval_adjoint_local_set(float* res, const float* rhs) override {
    apply(res, rhs, d_ptr_t_, d_ind_t_, d_val_t_, n_, m_, true);
    return true;
  }

 private:
  void apply(float* res, const float* rhs, const DeviceBuffer<int>& ptr, const DeviceBuffer<int>& ind,
             const DeviceBuffer<float>& val, int n_out, int n_in, bool set) {
    if (n_out == 0 || n_in == 0) return;
    cudaStream_t st = ctx_->stream;
    const int* skip = ctx_->skip_flag;
    if (val.size() == 0) {                       // empty factor: K = 0
      if (set) PB_CUDA(cudaMemsetAsync(res, 0, d_ * (size_t)n_out * sizeof(float), st));
      return;
    }
    if (!id_first_) {
      const bool vec4 = d_ % 4 == 0 && ((reinterpret_cast<uintptr_t>(res) | reinterpret_cast<uintptr_t>(rhs)) & 15u) == 0;
      const size_t threads = (vec4 ? d_ / 4 : d_) * (size_t)n_out;
      const unsigned grid = (unsigned)std::min<size_t>(grid_for(threads), (size_t)ctx_->num_sms * 32);
#define PB_KRON(V, S) kron_sparse_id_kernel<V, S><<<grid, kBlock, 0, st>>>(res, rhs, ptr.data(), ind.data(), val.data(), (uint32_t)n_out, d_, skip)
      if (vec4) { if (set) PB_KRON(4, true); else PB_KRON(4, false); }
      else { if (set) PB_KRON(1, true); else PB_KRON(1, false); }
#undef PB_KRON
    } else {
      const size_t warps = kBlock / 32;
      const size_t tile_b = std::max<size_t>(1, std::min<size_t>(32, 1024 / (size_t)std::max(n_in, 1)));
      const size_t smem = warps * tile_b * ((size_t)n_in + 1) * sizeof(float);
      if (smem <= 48 * 1024) {
        const size_t tiles = (d_ + tile_b - 1) / tile_b;
        const unsigned grid = (unsigned)std::min<size_t>((tiles + warps - 1) / warps, (size_t)ctx_->num_sms * 8);
        if (set)
          kron_id_sparse_kernel<true><<<grid, kBlock, smem, st>>>(res, rhs, ptr.data(), ind.data(), val.data(),
                                                                  (uint32_t)n_out, (uint32_t)n_in, d_, (uint32_t)tile_b, skip);
        else
          kron_id_sparse_kernel<false><<<grid, kBlock, smem, st>>>(res, rhs, ptr.data(), ind.data(), val.data(),
                                                                   (uint32_t)n_out, (uint32_t)n_in, d_, (uint32_t)tile_b, skip);
      } else {
        const unsigned grid = (unsigned)std::min<size_t>(grid_for(d_ * (size_t)n_out), (size_t)ctx_->num_sms * 32);
        if (set)
          kron_id_sparse_big_kernel<true><<<grid, kBlock, 0, st>>>(res, rhs, ptr.data(), ind.data(), val.data(),
                                                                   (uint32_t)n_out, (uint32_t)n_in, d_, skip);
        else
          kron_id_sparse_big_kernel<false><<<grid, kBlock, 0, st>>>(res, rhs, ptr.data(), ind.data(), val.data(),
                                                                    (uint32_t)n_out, (uint32_t)n_in, d_, skip);
      }
    }
    PB_CHECK_LAUNCH();
    ctx_->launches++;
  }
  bool id_first_;
  size_t d_;
  int m_, n_;
  std::vector<int> ptr_, ind_, ptr_t_, ind_t_;
  std::vector<float> val_, val_t_;
  DeviceBuffer<int> d_ptr_, d_ind_, d_ptr_t_, d_ind_t_;
  DeviceBuffer<float> d_val_, d_val_t_;
};

// ---- zero -----------------------------------------------------------------------------------------

class BlockZero : public Block {
 public:
  using Block::Block;
  int kind() const override { return kBlockZero; }
  float row_sum(size_t, float) const override { return 0.f; }
  float col_sum(size_t, float) const override { return 0.f; }
  bool uniform_sums() const override { return true; }
  void eval_local_add(float*, const float*) override {}
  void eval_adjoint_local_add(float*, const float*) override {}
};

// ---- factories ------------------------------------------------------------------------------------

std::shared_ptr<Block> make_block_gradient(Context* ctx, bool three_d, size_t row, size_t col,
                                           size_t nx, size_t ny, size_t L, bool label_first) {
  return std::make_shared<BlockGradient>(ctx, three_d, row, col, nx, ny, L, label_first);
}
std::shared_ptr<Block> make_block_diags(Context* ctx, size_t row, size_t col, size_t nrows,
                                        size_t ncols, size_t ndiags, const int64_t* offsets,
                                        const float* factors) {
  return std::make_shared<BlockDiags>(ctx, row, col, nrows, ncols, ndiags, offsets, factors);
}
std::shared_ptr<Block> make_block_sparse_csc(Context* ctx, size_t row, size_t col, int m, int n,
                                             int nnz, const float* val, const int32_t* ptr,
                                             const int32_t* ind) {
  return std::make_shared<BlockSparse>(ctx, row, col, m, n, nnz, val, ptr, ind);
}
std::shared_ptr<Block> make_block_dense(Context* ctx, size_t row, size_t col, size_t nrows,
                                        size_t ncols, const float* data) {
  return std::make_shared<BlockDense>(ctx, row, col, nrows, ncols, data);
}
std::shared_ptr<Block> make_block_dense_kron(Context* ctx, bool id_first, size_t diaglength, size_t row, size_t col,
                                             size_t mat_nrows, size_t mat_ncols, const float* data) {
  return std::make_shared<BlockDenseKron>(ctx, id_first, diaglength, row, col, mat_nrows, mat_ncols, data);
}
std::shared_ptr<Block> make_block_sparse_kron(Context* ctx, bool id_first, size_t diaglength, size_t row, size_t col,
                                              int m, int n, int nnz, const float* val, const int32_t* ptr,
                                              const int32_t* ind) {
  return std::make_shared<BlockSparseKron>(ctx, id_first, diaglength, row, col, m, n, nnz, val, ptr, ind);
}
std::shared_ptr<Block> make_block_zero(Context* ctx, size_t row, size_t col, size_t nrows,
                                       size_t ncols) {
  return std::make_shared<BlockZero>(ctx, row, col, nrows, ncols);
}

// ---- LinearOperator -------------------------------------------------------------------------------

static bool rect_overlap(size_t x1, size_t y1, size_t x2, size_t y2, size_t a1, size_t b1, size_t a2,
                         size_t b2) {
  return (x1 <= a2) && (x2 >= a1) && (y1 <= b2) && (y2 >= b1);
}

void LinearOperator::initialize() {
  nrows_ = ncols_ = 0;
  bool overlap = false;
  for (size_t i = 0; i < blocks_.size(); ++i) {
    const Block& a = *blocks_[i];
    nrows_ = std::max(nrows_, a.row() + a.nrows());
    ncols_ = std::max(ncols_, a.col() + a.ncols());
    for (size_t j = i + 1; j < blocks_.size(); ++j) {
      const Block& b = *blocks_[j];
      if (a.nrows() == 0 || a.ncols() == 0 || b.nrows() == 0 || b.ncols() == 0) continue;
      overlap |= rect_overlap(a.col(), a.row(), a.col() + a.ncols() - 1, a.row() + a.nrows() - 1,
                              b.col(), b.row(), b.col() + b.ncols() - 1, b.row() + b.nrows() - 1);
    }
  }
  if (overlap)
    fail(PB_ERR_INVALID, "Blocks are overlapping inside the linear operator. Recheck the indices.");
  auto disjoint = [&](bool rows) {
    std::vector<std::pair<size_t, size_t>> r;
    for (auto& b : blocks_) {
      const size_t lo = rows ? b->row() : b->col(), n = rows ? b->nrows() : b->ncols();
      if (b->nrows() && b->ncols()) r.emplace_back(lo, lo + n);
    }
    std::sort(r.begin(), r.end());
    for (size_t i = 1; i < r.size(); ++i)
      if (r[i].first < r[i - 1].second) return false;
    return true;
  };
  disjoint_rows_ = disjoint(true);
  disjoint_cols_ = disjoint(false);
}

void LinearOperator::eval(float* d_result, const float* d_rhs, float beta, bool transpose, bool negate) {
  ctx_->bind();
  const size_t nout = transpose ? ncols_ : nrows_;
  if (nout == 0) return;
  if (beta == 0.f && !negate && !ctx_->skip_flag && (transpose ? disjoint_cols_ : disjoint_rows_)) {
    // nobody else writes a block's outputs: blocks that can overwrite do so, the others (and the gaps between the
    // blocks) see zeros first
    size_t covered = 0;      // outputs [0, covered) are dealt with; blocks are visited in output order
    std::vector<Block*> order;
    for (auto& b : blocks_) order.push_back(b.get());
    std::sort(order.begin(), order.end(), [&](Block* a, Block* b) {
      return (transpose ? a->col() : a->row()) < (transpose ? b->col() : b->row());
    });
    for (Block* b : order) {
      if (b->nrows() == 0 || b->ncols() == 0) continue;
      const size_t lo = transpose ? b->col() : b->row(), n = transpose ? b->ncols() : b->nrows();
      if (lo > covered) PB_CUDA(cudaMemsetAsync(d_result + covered, 0, (lo - covered) * sizeof(float), ctx_->stream));
      const bool done = transpose ? b->eval_adjoint_local_set(d_result + b->col(), d_rhs + b->row())
                                  : b->eval_local_set(d_result + b->row(), d_rhs + b->col());
      if (!done) {
        PB_CUDA(cudaMemsetAsync(d_result + lo, 0, n * sizeof(float), ctx_->stream));
        if (transpose) b->eval_adjoint_local_add(d_result + b->col(), d_rhs + b->row());
        else b->eval_local_add(d_result + b->row(), d_rhs + b->col());
      }
      covered = std::max(covered, lo + n);
    }
    if (nout > covered) PB_CUDA(cudaMemsetAsync(d_result + covered, 0, (nout - covered) * sizeof(float), ctx_->stream));
    return;
  }
  if (beta == 0.f) {
    PB_CUDA(cudaMemsetAsync(d_result, 0, nout * sizeof(float), ctx_->stream));
  } else if (beta != 1.f) {
    const unsigned grid = std::min<size_t>(grid_for(nout), (size_t)ctx_->num_sms * 32);
    scale_kernel<<<grid, kBlock, 0, ctx_->stream>>>(d_result, nout, beta, ctx_->skip_flag);
    PB_CHECK_LAUNCH();
    ctx_->launches++;
  }
  if (!negate) {
    for (auto& b : blocks_) {
      if (transpose)
        b->eval_adjoint_local_add(d_result + b->col(), d_rhs + b->row());
      else
        b->eval_local_add(d_result + b->row(), d_rhs + b->col());
    }
  } else {
    // DualLinearOperator (dual_linearoperator.cu:38-80): result = beta*result - K^T rhs, realised
    // as  result <- -( -beta*result + K^T rhs ).  Only used when solve_dual_problem is set.
    const unsigned grid = std::min<size_t>(grid_for(nout), (size_t)ctx_->num_sms * 32);
    if (beta != 0.f) {
      scale_kernel<<<grid, kBlock, 0, ctx_->stream>>>(d_result, nout, -1.f, ctx_->skip_flag);
      PB_CHECK_LAUNCH();
      ctx_->launches++;
    }
    for (auto& b : blocks_) {
      if (transpose)
        b->eval_adjoint_local_add(d_result + b->col(), d_rhs + b->row());
      else
        b->eval_local_add(d_result + b->row(), d_rhs + b->col());
    }
    scale_kernel<<<grid, kBlock, 0, ctx_->stream>>>(d_result, nout, -1.f, ctx_->skip_flag);
    PB_CHECK_LAUNCH();
    ctx_->launches++;
  }
}

float LinearOperator::row_sum(size_t row, float alpha) const {
  float sum = 0;
  for (auto& b : blocks_) {
    if (row < b->row() || row >= b->row() + b->nrows()) continue;
    sum += b->row_sum(row - b->row(), alpha);
  }
  return sum;
}

float LinearOperator::col_sum(size_t col, float alpha) const {
  float sum = 0;
  for (auto& b : blocks_) {
    if (col < b->col() || col >= b->col() + b->ncols()) continue;
    sum += b->col_sum(col - b->col(), alpha);
  }
  return sum;
}

// All sums in one sweep per block (the reference makes nrows+ncols virtual calls through
// LinearOperator::row_sum, problem.cu:262-287; the accumulation order per row -- block list
// order -- is the same).
void LinearOperator::row_sums(float alpha, std::vector<float>& out) const {
  out.assign(nrows_, 0.f);
  for (auto& b : blocks_) {
    if (b->uniform_sums()) {
      const float v = b->nrows() ? b->row_sum(0, alpha) : 0.f;
      for (size_t r = 0; r < b->nrows(); ++r) out[b->row() + r] += v;
    } else {
      for (size_t r = 0; r < b->nrows(); ++r) out[b->row() + r] += b->row_sum(r, alpha);
    }
  }
}

void LinearOperator::col_sums(float alpha, std::vector<float>& out) const {
  out.assign(ncols_, 0.f);
  for (auto& b : blocks_) {
    if (b->uniform_sums()) {
      const float v = b->ncols() ? b->col_sum(0, alpha) : 0.f;
      for (size_t c = 0; c < b->ncols(); ++c) out[b->col() + c] += v;
    } else {
      for (size_t c = 0; c < b->ncols(); ++c) out[b->col() + c] += b->col_sum(c, alpha);
    }
  }
}

size_t LinearOperator::gpu_mem_amount() const {
  size_t mem = 0;
  for (auto& b : blocks_) mem += b->gpu_mem_amount();
  return mem;
}

bool LinearOperator::all_stencil() const {
  for (auto& b : blocks_)
    if (!block_is_stencil(b->kind())) return false;
  return true;
}

}  // namespace pb
