// oracle/prost_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement (C++/OpenMP) of the reference's PDHG / ADMM hot path, used ONLY as the checker in
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  The
// product (prost_b200/) never links, imports or calls anything in this directory.
//
// It follows the reference's UNFUSED structure and operation order (9 state vectors, one loop
// per reference kernel) so that it can be read side by side with /root/reference:
//   operators   src/linop/block_gradient2d.cu:25-139, block_gradient3d.cu:25-150,
//               block_diags.cu:36-119, block_sparse.cu:33-68 (+ common.cu:54-82), block_dense.cu:82-188,
//               linearoperator.cu:134-170, block.cu:46-68
//   proxes      include/prost/prox/elemop/function_1d.hpp:34-326, elem_operation_1d.hpp:36-59,
//               elem_operation_norm2.hpp:39-88, elem_operation_ind_simplex.hpp:47-115,
//               src/prox/prox_ind_epi_quad.cu:42-79 + include/prost/prox/helper.hpp:44-105,
//               src/prox/prox_moreau.cu:29-134, prox_permute.cu:30-145, prox_zero.cu:36-48,
//               include/prost/prox/vector.hpp:42-48, prox_elem_operation.inl:32-94
//   problem     src/problem.cu:92-158 (zero fill), :262-306 (scaling), :502-536 (averaging)
//   PDHG        src/backend/backend_pdhg.cu:38-186, 199-309, 311-489, 513-563; backend.hpp:71-74
//   ADMM        src/backend/backend_admm.cu:52-272, 285-665, 697-741; include/prost/cgls.hpp:222-371
// Parity pinning: against outputs of the reference itself (oracle/_ref, run on the GPU box) stored
// under tests/golden/, and against the closed forms of the reference's MATLAB tests
// (matlab/+prost/+test/*.m) restated in numpy inside tests/.
//
// Build: g++ -O2 -ffp-contract=off -fopenmp -shared -fPIC (see oracle/Makefile).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace orc {

typedef std::vector<float> vec;
static thread_local std::string g_err;

// ------------------------------------------------------------------------------------------------
// Function1D family (function_1d.hpp)
// ------------------------------------------------------------------------------------------------
enum Fn { ZERO, ABS, SQUARE, IND_LEQ0, IND_GEQ0, IND_EQ0, IND_BOX01, MAX_POS0, L0, HUBER, LQ,
          LQ_PLUS_EPS, TRUNC_QUAD, TRUNC_LINEAR };

static float fn_abs(float x0, float tau) {                 // :50-62
  if (x0 >= tau) return x0 - tau;
  else if (x0 <= -tau) return x0 + tau;
  return 0;
}
static float fn_square(float x0, float tau) { return x0 / (1. + tau); }   // :64-74 (double literal)
static float fn_l0(float x0, float tau) { return (x0 * x0 > 2 * tau) ? x0 : 0; }   // :143-155

static float lq_newton(float t0, float alpha, float q, float eps) {       // :171-191
  float t = t0, delta = 0;
  do {
    const float power = std::pow(t, q);
    const float dF1 = t - 1 + alpha * q * power / t;
    const float dF2 = 1 + alpha * q * (q - 1) * power / (t * t);
    delta = dF1 / dF2;
    t = t - delta;
  } while (delta > eps);
  return t;
}
static float lq_half(float alpha) {                                        // :193-202
  const float sqrt3 = std::sqrt(3.f);
  const float PI_half = 1.5707963267948966192313216916397514420985846996875529f;
  const float s = 2 * (std::sin((std::acos(alpha * 3 * sqrt3 / 4) + PI_half) / 3)) / sqrt3;
  return s * s;
}
static float fn_lq(float x0, float tau, float alpha) {                     // :204-257
  if (alpha == 1) return fn_abs(x0, tau);
  if (alpha == 0) return fn_l0(x0, tau);
  const float eps = 1e-5f;
  float t = 0;
  if (std::fabs(x0) > 0) {
    float factor = tau * std::pow(std::fabs(x0), alpha - 2);
    if (alpha < 1) {
      const float t2 = 2 * (alpha - 1) / (alpha - 2);
      if (factor < 0.5 * (1 - (t2 - 1) * (t2 - 1)) / std::pow(t2, alpha)) {
        if (alpha == 0.5) t = lq_half(factor);
        else t = lq_newton(1, factor, alpha, eps);
      }
    } else {
      t = lq_newton(1, factor, alpha, eps);
    }
  }
  return t * std::fabs(x0);
}

static float fn_eval(int fn, float x0, float tau, float alpha, float beta) {
  switch (fn) {
    case ZERO: return x0;
    case ABS: return fn_abs(x0, tau);
    case SQUARE: return fn_square(x0, tau);
    case IND_LEQ0: return x0 > 0. ? 0.f : x0;
    case IND_GEQ0: return x0 < 0. ? 0.f : x0;
    case IND_EQ0: return 0;
    case IND_BOX01: return x0 > 1. ? 1.f : (x0 < 0. ? 0.f : x0);
    case MAX_POS0: return x0 > tau ? x0 - tau : (x0 < 0. ? x0 : 0.f);
    case L0: return fn_l0(x0, tau);
    case HUBER: {                                                          // :157-169
      float result = (x0 / tau) / (1. + alpha / tau);
      result /= std::max(1.f, std::fabs(result));
      return x0 - tau * result;
    }
    case LQ: return fn_lq(x0, tau, alpha);
    case LQ_PLUS_EPS: return 0;                                            // :293-306 (stub)
    case TRUNC_QUAD: {                                                     // :273-291
      const float x_sq = fn_square(x0, 2 * tau * alpha);
      const float en_sq = alpha * x_sq * x_sq + (x_sq - x0) * (x_sq - x0) / (2 * tau);
      return en_sq < beta ? x_sq : x0;
    }
    case TRUNC_LINEAR: {                                                   // :308-326
      const float x_shrink = fn_abs(x0, tau * alpha);
      const float en = (x_shrink - x0) * (x_shrink - x0) / (2 * tau) + alpha * std::fabs(x_shrink);
      return en < beta ? x_shrink : x0;
    }
  }
  return x0;
}

// ------------------------------------------------------------------------------------------------
// Blocks
// ------------------------------------------------------------------------------------------------
struct Block {
  size_t row, col, nrows, ncols;
  Block(size_t r, size_t c, size_t nr, size_t nc) : row(r), col(c), nrows(nr), ncols(nc) {}
  virtual ~Block() {}
  virtual void eval_add(float* res, const float* rhs) const = 0;          // res += K rhs (local ptrs)
  virtual void eval_adj_add(float* res, const float* rhs) const = 0;      // res += K^T rhs
  virtual float row_sum(size_t r, float alpha) const = 0;
  virtual float col_sum(size_t c, float alpha) const = 0;
};

struct BlockGradient : Block {      // block_gradient2d.cu / block_gradient3d.cu
  size_t nx, ny, L;
  bool label_first, three_d;
  BlockGradient(size_t r, size_t c, size_t nx_, size_t ny_, size_t L_, bool lf, bool td)
      : Block(r, c, nx_ * ny_ * L_ * (td ? 3 : 2), nx_ * ny_ * L_), nx(nx_), ny(ny_), L(L_),
        label_first(lf), three_d(td) {}
  void eval_add(float* res, const float* rhs) const override {
    const size_t N = nx * ny * L;
#pragma omp parallel for collapse(2) schedule(static)
    for (size_t x = 0; x < nx; ++x)
      for (size_t yt = 0; yt < ny * L; ++yt) {
        size_t y, l, idx, idx_gx, idx_gy, idx_gl;
        if (label_first) { l = yt % L; y = yt / L; idx = l + y * L + x * ny * L; idx_gy = idx + L; idx_gx = idx + ny * L; idx_gl = idx + 1; }
        else { y = yt % ny; l = yt / ny; idx = y + x * ny + l * nx * ny; idx_gy = idx + 1; idx_gx = idx + ny; idx_gl = idx + ny * nx; }
        const float val_pt = rhs[idx];
        float gx = 0, gy = 0;
        if (y < ny - 1) gy = rhs[idx_gy] - val_pt;
        if (x < nx - 1) gx = rhs[idx_gx] - val_pt;
        res[idx] += gx;
        res[idx + N] += gy;
        if (three_d) {
          float gl;
          if (l < L - 1) gl = rhs[idx_gl] - val_pt;
          else gl = -val_pt;                                                // Dirichlet (3d:73-76)
          res[idx + 2 * N] += gl;
        }
      }
  }
  void eval_adj_add(float* res, const float* rhs) const override {
    const size_t N = nx * ny * L;
#pragma omp parallel for collapse(2) schedule(static)
    for (size_t x = 0; x < nx; ++x)
      for (size_t yt = 0; yt < ny * L; ++yt) {
        size_t y, l, idx, sy, sx, sl;
        if (label_first) { y = yt / L; l = yt % L; idx = l + y * L + x * ny * L; sy = L; sx = ny * L; sl = 1; }
        else { y = yt % ny; l = yt / ny; idx = y + x * ny + l * nx * ny; sy = 1; sx = ny; sl = nx * ny; }
        float divx, divy;
        if (y < ny - 1) divy = rhs[idx + N]; else divy = 0;
        if (y > 0) divy -= rhs[idx + N - sy];
        if (x < nx - 1) divx = rhs[idx]; else divx = 0;
        if (x > 0) divx -= rhs[idx - sx];
        if (!three_d) {
          res[idx] -= (divx + divy);
        } else {
          float divl = rhs[idx + 2 * N];
          if (l > 0) divl -= rhs[idx + 2 * N - sl];
          res[idx] -= (divx + divy + divl);
        }
      }
  }
  float row_sum(size_t, float) const override { return 2; }
  float col_sum(size_t, float) const override { return three_d ? 6 : 4; }
};

struct BlockDiags : Block {         // block_diags.cu
  std::vector<long long> ofs;
  std::vector<float> fac;
  BlockDiags(size_t r, size_t c, size_t nr, size_t nc, size_t nd, const int64_t* o, const float* f)
      : Block(r, c, nr, nc), ofs(o, o + nd), fac(f, f + nd) {
    for (size_t i = 0; i < nd; i++)                                        // :110-118
      for (size_t j = i; j < nd; j++)
        if (ofs[i] > ofs[j]) { std::swap(ofs[i], ofs[j]); std::swap(fac[i], fac[j]); }
  }
  void eval_add(float* res, const float* rhs) const override {
#pragma omp parallel for schedule(static)
    for (size_t r = 0; r < nrows; ++r) {
      float result = 0;
      for (size_t i = 0; i < ofs.size(); i++) {
        const long long c = (long long)r + ofs[i];
        if (c < 0) continue;
        if (c >= (long long)ncols) break;
        result += rhs[c] * fac[i];
      }
      res[r] += result;
    }
  }
  void eval_adj_add(float* res, const float* rhs) const override {
    // all ncols columns (the reference launch is sized by nrows, block_diags.cu:211)
#pragma omp parallel for schedule(static)
    for (size_t c = 0; c < ncols; ++c) {
      float result = 0;
      for (size_t i = 0; i < ofs.size(); i++) {
        const long long o = ofs[i], cc = (long long)c;
        if (o <= cc && (cc - o) < (long long)nrows && (cc - o) >= 0) result += rhs[cc - o] * fac[i];
        if (o > cc) break;
      }
      res[c] += result;
    }
  }
  float row_sum(size_t r, float alpha) const override {
    float sum = 0;
    for (size_t i = 0; i < ofs.size(); i++) {
      const long long c = (long long)r + ofs[i];
      if (c < 0) continue;
      if ((size_t)c >= ncols) break;
      sum += std::pow(std::abs(fac[i]), alpha);
    }
    return sum;
  }
  float col_sum(size_t c, float alpha) const override {
    float sum = 0;
    const long long sc = (long long)c;
    for (size_t i = 0; i < ofs.size(); i++) {
      const long long o = ofs[i];
      if (o <= sc && (sc - o) < (long long)nrows && (sc - o) >= 0) sum += std::pow(std::abs(fac[i]), alpha);
      if (o > sc) break;
    }
    return sum;
  }
};

struct BlockSparse : Block {        // block_sparse.cu (CSC in, CSR of K and of K^T kept)
  std::vector<int> ptr, ind, ptr_t, ind_t;
  vec val, val_t;
  BlockSparse(size_t r, size_t c, int m, int n, int nnz, const float* v, const int32_t* p, const int32_t* i)
      : Block(r, c, m, n), ptr_t(p, p + n + 1), ind_t(i, i + nnz), val_t(v, v + nnz) {
    ptr.assign(m + 1, 0);
    ind.resize(nnz);
    val.resize(nnz);
    for (int k = 0; k < nnz; ++k) ptr[ind_t[k] + 1]++;
    for (int q = 0; q < m; ++q) ptr[q + 1] += ptr[q];
    std::vector<int> pos(ptr.begin(), ptr.end() - 1);
    for (int col_ = 0; col_ < n; ++col_)
      for (int k = ptr_t[col_]; k < ptr_t[col_ + 1]; ++k) {
        const int d = pos[ind_t[k]]++;
        ind[d] = col_;
        val[d] = val_t[k];
      }
  }
  static void spmv(const std::vector<int>& p, const std::vector<int>& i, const vec& v, size_t rows,
                   float* res, const float* x) {
#pragma omp parallel for schedule(static)
    for (size_t r = 0; r < rows; ++r) {
      float acc = 0;
      for (int k = p[r]; k < p[r + 1]; ++k) acc += v[k] * x[i[k]];
      res[r] += acc;
    }
  }
  void eval_add(float* res, const float* rhs) const override { spmv(ptr, ind, val, nrows, res, rhs); }
  void eval_adj_add(float* res, const float* rhs) const override { spmv(ptr_t, ind_t, val_t, ncols, res, rhs); }
  float row_sum(size_t r, float alpha) const override {
    float s = 0;
    for (int k = ptr[r]; k < ptr[r + 1]; ++k) s += std::pow(std::abs(val[k]), alpha);
    return s;
  }
  float col_sum(size_t c, float alpha) const override {
    float s = 0;
    for (int k = ptr_t[c]; k < ptr_t[c + 1]; ++k) s += std::pow(std::abs(val_t[k]), alpha);
    return s;
  }
};

// SURVEY.md 8(f) row 3, oracle side only so far: Kronecker products of a small matrix K (mr x mc, column-major) with
// an identity of size d.  id_first = false: kron(K, I_d) (block_dense_kron_id.cu:28-64, block_sparse_kron_id.cu),
// id_first = true: kron(I_d, K) (block_id_kron_dense.cu:28-64, block_id_kron_sparse.cu).  The sparse variants are
// the same operators with K given in CSC; the oracle stores K densely (zeros add exactly 0 to the float sums).
struct BlockKron : Block {
  vec k;
  size_t mr, mc, d;
  bool id_first;
  BlockKron(size_t r, size_t c, size_t mr_, size_t mc_, size_t d_, bool idf, const float* data)
      : Block(r, c, mr_ * d_, mc_ * d_), k(data, data + mr_ * mc_), mr(mr_), mc(mc_), d(d_), id_first(idf) {}
  void eval_add(float* res, const float* rhs) const override {
#pragma omp parallel for schedule(static)
    for (size_t tx = 0; tx < d * mr; ++tx) {
      float sum = 0;
      if (!id_first) {
        const size_t row = tx / d, ofs = tx % d;
        for (size_t i = 0; i < mc; ++i) sum += k[i * mr + row] * rhs[i * d + ofs];
      } else {
        const size_t row = tx % mr, ofs = (tx / mr) * mc;
        for (size_t i = 0; i < mc; ++i) sum += k[i * mr + row] * rhs[ofs + i];
      }
      res[tx] += sum;
    }
  }
  void eval_adj_add(float* res, const float* rhs) const override {
#pragma omp parallel for schedule(static)
    for (size_t tx = 0; tx < d * mc; ++tx) {
      float sum = 0;
      if (!id_first) {
        const size_t colm = tx / d, ofs = tx % d;
        for (size_t i = 0; i < mr; ++i) sum += k[i + colm * mr] * rhs[i * d + ofs];
      } else {
        const size_t colm = tx % mc, ofs = (tx / mc) * mr;
        for (size_t i = 0; i < mr; ++i) sum += k[i + colm * mr] * rhs[ofs + i];
      }
      res[tx] += sum;
    }
  }
  float row_sum(size_t r, float alpha) const override {                 // block_dense_kron_id.cu:100-109, id_kron_*: row % mr
    const size_t row = id_first ? r % mr : r / d;
    float s = 0;
    for (size_t i = 0; i < mc; ++i) s += std::pow(std::abs(k[i * mr + row]), alpha);
    return s;
  }
  float col_sum(size_t c, float alpha) const override {
    const size_t colm = id_first ? c % mc : c / d;
    float s = 0;
    for (size_t i = 0; i < mr; ++i) s += std::pow(std::abs(k[i + colm * mr]), alpha);
    return s;
  }
};

struct BlockDense : Block {         // block_dense.cu (column-major, gemv N / T)
  vec a;
  BlockDense(size_t r, size_t c, size_t nr, size_t nc, const float* d) : Block(r, c, nr, nc), a(d, d + nr * nc) {}
  void eval_add(float* res, const float* rhs) const override {
#pragma omp parallel for schedule(static)
    for (size_t r = 0; r < nrows; ++r) {
      double acc = 0;
      for (size_t c = 0; c < ncols; ++c) acc += (double)a[c * nrows + r] * rhs[c];
      res[r] += (float)acc;
    }
  }
  void eval_adj_add(float* res, const float* rhs) const override {
#pragma omp parallel for schedule(static)
    for (size_t c = 0; c < ncols; ++c) {
      double acc = 0;
      for (size_t r = 0; r < nrows; ++r) acc += (double)a[c * nrows + r] * rhs[r];
      res[c] += (float)acc;
    }
  }
  float row_sum(size_t r, float alpha) const override {
    float s = 0;
    for (size_t c = 0; c < ncols; ++c) s += std::pow(std::abs(a[c * nrows + r]), alpha);
    return s;
  }
  float col_sum(size_t c, float alpha) const override {
    float s = 0;
    for (size_t r = 0; r < nrows; ++r) s += std::pow(std::abs(a[c * nrows + r]), alpha);
    return s;
  }
};

struct BlockZero : Block {
  using Block::Block;
  void eval_add(float*, const float*) const override {}
  void eval_adj_add(float*, const float*) const override {}
  float row_sum(size_t, float) const override { return 0; }
  float col_sum(size_t, float) const override { return 0; }
};

// ------------------------------------------------------------------------------------------------
// Proxes
// ------------------------------------------------------------------------------------------------
struct Prox {
  size_t index, size;
  bool diagsteps;
  Prox(size_t i, size_t s, bool d) : index(i), size(s), diagsteps(d) {}
  virtual ~Prox() {}
  size_t end() const { return index + size - 1; }
  // local pointers (already offset by index), like Prox::EvalLocal
  virtual void eval_local(float* res, const float* arg, const float* tau_diag, float tau, bool invert) = 0;
  virtual void sep(std::vector<std::tuple<size_t, size_t, size_t>>& s) const { s.emplace_back(index, size, 1); }
  void eval(float* res, const float* arg, const float* td, float tau, bool invert) {   // prox.cu:26-43
    eval_local(res + index, arg + index, td + index, tau, invert);
  }
};

struct ProxZero : Prox {            // prox_zero.cu
  ProxZero(size_t i, size_t s) : Prox(i, s, true) {}
  void eval_local(float* res, const float* arg, const float*, float, bool) override {
    std::copy(arg, arg + size, res);
  }
};

struct ProxSeparable : Prox {       // prox_separable_sum.hpp
  size_t count, dim;
  bool interleaved;
  ProxSeparable(size_t i, size_t c, size_t d, bool il, bool ds) : Prox(i, c * d, ds), count(c), dim(d), interleaved(il) {}
  size_t at(size_t tx, size_t i) const { return interleaved ? (tx * dim + i) : (tx + count * i); }   // vector.hpp:42-48
  void sep(std::vector<std::tuple<size_t, size_t, size_t>>& s) const override {
    for (size_t i = 0; i < count; i++) {
      if (interleaved) s.emplace_back(index + i * dim, dim, 1);
      else s.emplace_back(index + i, dim, count);
    }
  }
};

struct ProxElem7 : ProxSeparable {  // 1D and Norm2 families with 7 coefficients
  int fn;
  bool norm2;
  vec coeffs[7];
  ProxElem7(bool n2, size_t i, size_t c, size_t d, bool il, bool ds, int f, const float* const* co, const size_t* len)
      : ProxSeparable(i, c, n2 ? d : 1, il, ds), fn(f), norm2(n2) {
    for (int k = 0; k < 7; ++k) coeffs[k].assign(co[k], co[k] + len[k]);
  }
  void eval_local(float* res, const float* arg, const float* td, float tau_scal, bool invert) override {
#pragma omp parallel for schedule(static)
    for (size_t tx = 0; tx < count; ++tx) {
      float c[7];
      for (int k = 0; k < 7; ++k) c[k] = coeffs[k].size() > 1 ? coeffs[k][tx] : coeffs[k][0];   // inl:82-89
      if (!norm2) {                                                        // elem_operation_1d.hpp:36-59
        const size_t e = at(tx, 0);
        float tau = invert ? (1. / (tau_scal * td[e])) : (tau_scal * td[e]);
        if (c[0] == 0 || c[2] == 0) {
          res[e] = (arg[e] - tau * c[3]) / (1 + tau * c[4]);
        } else {
          const float prox_arg = ((c[0] * (arg[e] - c[3] * tau)) / (1. + tau * c[4])) - c[1];
          const float step = (c[2] * c[0] * c[0] * tau) / (1. + tau * c[4]);
          res[e] = (fn_eval(fn, prox_arg, step, c[5], c[6]) + c[1]) / c[0];
        }
      } else {                                                             // elem_operation_norm2.hpp:39-88
        float norm = 0;
        for (size_t i = 0; i < dim; i++) { const float v = arg[at(tx, i)]; norm += v * v; }
        if (norm > 0) {
          norm = std::sqrt(norm);
          const float t0 = td[at(tx, 0)];
          float tau = invert ? (1. / (tau_scal * t0)) : (tau_scal * t0);
          const float prox_arg = ((c[0] * (norm - c[3] * tau)) / (1. + tau * c[4])) - c[1];
          const float step = (c[2] * c[0] * c[0] * tau) / (1. + tau * c[4]);
          const float prox_result = (fn_eval(fn, prox_arg, step, c[5], c[6]) + c[1]) / c[0];
          for (size_t i = 0; i < dim; i++) res[at(tx, i)] = prox_result * arg[at(tx, i)] / norm;
        } else {
          for (size_t i = 0; i < dim; i++) res[at(tx, i)] = 0;
        }
      }
    }
  }
};

// ---- spectral element operations ---------------------------------------------------------
// singular_nx2 and eigen_2x2 follow the reference statement by statement (elem_operation_singular_nx2.hpp:41-150,
// function_2d.hpp:28-101; elem_operation_eigen_2x2.hpp:30-146 incl. its dlaev2-style rotation).  eigen_3x3 / eigen_nxn:
// the reference diagonalises with Kopp's Cardano + cross-product routine (elem_operation_eigen_3x3.hpp:31-300) resp.
// EISPACK tred2 / tql2 (elem_operation_eigen_nxn.hpp); a spectral function V f(Lambda) V^T does not depend on the
// eigen-solver, so this restatement uses cyclic Jacobi rotations in double and is pinned on numpy's eigh (the
// closed form of the reference's own test_prox_sum_eigen_3x3.m / _nxn.m) and on the live reference.
static void oracle_ind_l1_ball(float y1, float y2, float& x1, float& x2, float alpha) {      // function_2d.hpp:43-82
  float v1 = std::abs(y1), v2 = std::abs(y2);
  if (v1 + v2 <= alpha) { x1 = y1; x2 = y2; return; }
  float mu1, mu2;
  if (v1 < v2) { mu1 = v2; mu2 = v1; } else { mu1 = v1; mu2 = v2; }
  float l = 0.5 * (mu2 - mu1 + alpha);
  char rho = 2;
  if (l <= 0.) rho = 1;
  float theta = (1. / rho) * (mu1 + (rho == 2 ? mu2 : 0.) - alpha);
  mu1 = std::max(v1 - theta, 0.f);
  mu2 = std::max(v2 - theta, 0.f);
  x1 = ((0.f < y1) - (y1 < 0.f)) * mu1;
  x2 = ((0.f < y2) - (y2 < 0.f)) * mu2;
}

static void jacobi_eig(std::vector<double>& A, std::vector<double>& V, int n) {     // A symmetric n x n, row-major
  V.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) V[i * n + i] = 1.0;
  for (int sweep = 0; sweep < 32; ++sweep) {
    double off = 0, diag = 0;
    for (int i = 0; i < n; ++i) { diag += A[i * n + i] * A[i * n + i]; for (int j = i + 1; j < n; ++j) off += A[i * n + j] * A[i * n + j]; }
    if (off <= 1e-32 * diag || off == 0) break;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = A[p * n + q];
        if (apq == 0) continue;
        const double theta = (A[q * n + q] - A[p * n + p]) / (2 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::abs(theta) + std::sqrt(theta * theta + 1));
        const double c = 1 / std::sqrt(t * t + 1), sn = t * c;
        for (int r = 0; r < n; ++r) {            // A <- A J
          const double arp = A[r * n + p], arq = A[r * n + q];
          A[r * n + p] = c * arp - sn * arq;
          A[r * n + q] = sn * arp + c * arq;
        }
        for (int r = 0; r < n; ++r) {            // A <- J^T A
          const double apr = A[p * n + r], aqr = A[q * n + r];
          A[p * n + r] = c * apr - sn * aqr;
          A[q * n + r] = sn * apr + c * aqr;
        }
        for (int r = 0; r < n; ++r) {
          const double vp = V[r * n + p], vq = V[r * n + q];
          V[r * n + p] = c * vp - sn * vq;
          V[r * n + q] = sn * vp + c * vq;
        }
      }
  }
}

struct ProxSpectral : ProxSeparable {
  int kind, fn, fn2d;      // kind: 0 singular_nx2, 1 eigen_2x2, 2 eigen_3x3, 3 eigen_nxn, 4-7 mass4 / comass4 / mass5 / comass5
  vec coeffs[7];
  ProxSpectral(int k, size_t i, size_t c, size_t d, bool il, bool ds, int f, int f2, const float* const* co, const size_t* len)
      : ProxSeparable(i, c, d, il, ds), kind(k), fn(f), fn2d(f2) {
    for (int j = 0; j < 7; ++j) coeffs[j].assign(co[j], co[j] + len[j]);
  }
  // one eigenvalue through the scaled Function1D (eigen_2x2.hpp:112-126)
  double eig_prox(double lam, double tau, const float* c) const {
    if (c[0] == 0 || c[2] == 0) return (lam - tau * c[3]) / (1 + tau * c[4]);
    const double p = ((c[0] * (lam - c[3] * tau)) / (1. + tau * c[4])) - c[1];
    const double step = (c[2] * c[0] * c[0] * tau) / (1. + tau * c[4]);
    return (fn_eval(fn, static_cast<float>(p), static_cast<float>(step), c[5], c[6]) + c[1]) / c[0];
  }
  void eval_local(float* res, const float* arg, const float* td, float tau_scal, bool invert) override {
#pragma omp parallel for schedule(static)
    for (size_t tx = 0; tx < count; ++tx) {
      float c[7];
      for (int k = 0; k < 7; ++k) c[k] = coeffs[k].size() > 1 ? coeffs[k][tx] : coeffs[k][0];
      const float td0 = td[at(tx, 0)];
      const double tau = invert ? (1. / (tau_scal * td0)) : (tau_scal * td0);
      if (kind == 0) {                                           // elem_operation_singular_nx2.hpp:41-150
        const size_t n = dim / 2;
        double d11 = 0., d12 = 0., d22 = 0.;
        for (size_t i = 0; i < n; i++) {
          const float a1 = arg[at(tx, i)], a2 = arg[at(tx, n + i)];
          d11 += a1 * a1; d12 += a1 * a2; d22 += a2 * a2;
        }
        const double trace = d11 + d22, det = d11 * d22 - d12 * d12;
        const double d = std::sqrt(std::max(0., 0.25 * trace * trace - det));
        const double lmax = std::max(0., 0.5 * trace + d), lmin = std::max(0., 0.5 * trace - d);
        const double smax = std::sqrt(lmax), smin = std::sqrt(lmin);
        double s1, s2;
        if (c[0] == 0 || c[2] == 0) {
          s1 = (smax - tau * c[3]) / (1. + tau * c[4]);
          s2 = (smin - tau * c[3]) / (1. + tau * c[4]);
        } else {
          const float y1 = ((c[0] * (smax - c[3] * tau)) / (1. + tau * c[4])) - c[1];
          const float y2 = ((c[0] * (smin - c[3] * tau)) / (1. + tau * c[4])) - c[1];
          const float step = (c[2] * c[0] * c[0] * tau) / (1. + tau * c[4]);
          float x1, x2;
          if (fn2d == 0) { x1 = fn_eval(fn, y1, step, c[5], c[6]); x2 = fn_eval(fn, y2, step, c[5], c[6]); }
          else if (fn2d == 1) oracle_ind_l1_ball(y1, y2, x1, x2, c[5]);
          else {                                                 // Function2DMoreau :85-101
            float r1, r2;
            oracle_ind_l1_ball(y1 / step, y2 / step, r1, r2, c[5]);
            x1 = y1 - step * r1;
            x2 = y2 - step * r2;
          }
          s1 = (x1 + c[1]) / c[0];
          s2 = (x2 + c[1]) / c[0];
        }
        if (smax > 0) {
          double v11, v12, v21, v22;                             // :97-123
          if (d12 == 0.0) {
            if (d11 >= d22) { v11 = 1; v21 = 0; v12 = 0; v22 = 1; } else { v11 = 0; v21 = 1; v12 = 1; v22 = 0; }
          } else {
            v11 = lmax - d22; v21 = d12;
            const double l1 = std::hypot(v11, v21);
            v11 /= l1; v21 /= l1;
            v12 = lmin - d22; v22 = d12;
            const double l2 = std::hypot(v12, v22);
            v12 /= l2; v22 /= l2;
          }
          s1 /= smax;
          s2 = (smin > 0.0) ? (s2 / smin) : 0.0;
          const double t11 = s1 * v11 * v11 + s2 * v12 * v12, t12 = s1 * v11 * v21 + s2 * v12 * v22;
          const double t21 = s1 * v21 * v11 + s2 * v22 * v12, t22 = s1 * v21 * v21 + s2 * v22 * v22;
          for (size_t i = 0; i < n; i++) {
            const float a1 = arg[at(tx, i)], a2 = arg[at(tx, n + i)];
            res[at(tx, i)] = a1 * t11 + a2 * t21;
            res[at(tx, n + i)] = a1 * t12 + a2 * t22;
          }
        } else {
          for (size_t i = 0; i < 2 * n; i++) res[at(tx, i)] = 0;
          res[at(tx, 0)] = s1;
          res[at(tx, n + 1)] = s2;
        }
      } else if (kind == 1) {                                    // elem_operation_eigen_2x2.hpp:30-146
        const double A = arg[at(tx, 0)], B = (arg[at(tx, 1)] + arg[at(tx, 2)]) / 2, C = arg[at(tx, 3)];
        const double sm = A + C, df = A - C, rt = std::sqrt(df * df + 4.0 * B * B);
        double rt1, rt2, cs, sn, t;
        if (sm > 0.0) { rt1 = 0.5 * (sm + rt); t = 1.0 / rt1; rt2 = (A * t) * C - (B * t) * B; }
        else if (sm < 0.0) { rt2 = 0.5 * (sm - rt); t = 1.0 / rt2; rt1 = (A * t) * C - (B * t) * B; }
        else { rt1 = 0.5 * rt; rt2 = -0.5 * rt; }
        cs = df > 0.0 ? df + rt : df - rt;
        if (std::abs(cs) > 2.0 * std::abs(B)) { t = -2.0 * B / cs; sn = 1.0 / std::sqrt(1.0 + t * t); cs = t * sn; }
        else if (std::abs(B) == 0.0) { cs = 1.0; sn = 0.0; }
        else { t = -0.5 * cs / B; cs = 1.0 / std::sqrt(1.0 + t * t); sn = t * cs; }
        if (df > 0.0) { t = cs; cs = -sn; sn = t; }
        rt1 = eig_prox(rt1, tau, c);
        rt2 = eig_prox(rt2, tau, c);
        const double t11 = rt1 * cs * cs + rt2 * sn * sn, t12 = rt1 * cs * sn - sn * rt2 * cs, t22 = rt1 * sn * sn + rt2 * cs * cs;
        res[at(tx, 0)] = t11; res[at(tx, 1)] = t12; res[at(tx, 2)] = t12; res[at(tx, 3)] = t22;
      } else if (kind >= 4) {                                    // mass4 / comass4 / mass5 / comass5
        // elem_operation_mass_norm.hpp:17-186.  The reference tridiagonalises the skew-symmetric matrix M of the
        // 2-vector and takes a 2 x 2 SVD; the prox is U f(Sigma) V^T = M h(M^T M), evaluated here through the
        // symmetric eigenproblem of M^T M (pinned on numpy's svd and on the live reference).
        const int nm = kind <= 5 ? 4 : 5;
        const bool conj = kind == 5 || kind == 7;
        const double tau_m = kind == 4 ? (invert ? (1. / (tau_scal * c[0] * td0)) : (tau_scal * c[0] * td0)) : tau;
        std::vector<double> M((size_t)nm * nm, 0.0), S((size_t)nm * nm, 0.0), V;
        int e = 0;
        for (int i = 0; i < nm; ++i)
          for (int j = i + 1; j < nm; ++j) { const double v = arg[at(tx, e++)]; M[i * nm + j] = v; M[j * nm + i] = -v; }
        for (int i = 0; i < nm; ++i)
          for (int j = 0; j < nm; ++j) { double t = 0; for (int k = 0; k < nm; ++k) t += M[k * nm + i] * M[k * nm + j]; S[i * nm + j] = t; }
        jacobi_eig(S, V, nm);
        std::vector<double> w(nm);
        for (int k = 0; k < nm; ++k) {
          const double sg = std::sqrt(std::max(S[k * nm + k], 0.0));
          const double f = conj ? std::min(sg, 1.0) : std::max(sg - tau_m, 0.0);
          w[k] = sg > 1e-150 ? f / sg : 0.0;
        }
        e = 0;
        for (int i = 0; i < nm; ++i)
          for (int j = i + 1; j < nm; ++j) {
            double t = 0;
            for (int l = 0; l < nm; ++l) {
              double h = 0;
              for (int k = 0; k < nm; ++k) h += V[l * nm + k] * V[j * nm + k] * w[k];
              t += M[i * nm + l] * h;
            }
            res[at(tx, e++)] = t;
          }
      } else {                                                   // eigen_3x3 / eigen_nxn (see the note above)
        int n = 1;
        while ((size_t)n * n < dim) ++n;
        std::vector<double> A((size_t)n * n), V;
        for (int i = 0; i < n; ++i)
          for (int j = i; j < n; ++j) {
            const float u = arg[at(tx, i * n + j)], l = arg[at(tx, j * n + i)];
            A[i * n + j] = A[j * n + i] = i == j ? static_cast<double>(u) : static_cast<double>(u + l) / 2.;
          }
        jacobi_eig(A, V, n);
        std::vector<double> f(n);
        for (int k = 0; k < n; ++k) f[k] = eig_prox(A[k * n + k], tau, c);
        for (int i = 0; i < n; ++i)
          for (int j = i; j < n; ++j) {
            double t = 0;
            for (int k = 0; k < n; ++k) t += V[i * n + k] * V[j * n + k] * f[k];
            res[at(tx, i * n + j)] = t;
            res[at(tx, j * n + i)] = t;
          }
      }
    }
  }
};

struct ProxSimplex : ProxSeparable {   // elem_operation_ind_simplex.hpp:47-115
  using ProxSeparable::ProxSeparable;
  void eval_local(float* res, const float* arg, const float*, float, bool) override {
#pragma omp parallel
    {
      vec local(dim);
#pragma omp for schedule(static)
      for (size_t tx = 0; tx < count; ++tx) {
        for (size_t i = 0; i < dim; i++) local[i] = arg[at(tx, i)];
        // descending shell sort with the reference's gap sequence (:94-115)
        const int gaps[6] = {132, 57, 23, 10, 4, 1};
        for (int k = 0; k < 6; k++) {
          const int gap = gaps[k];
          for (int i = gap; i < (int)dim; i++) {
            const float temp = local[i];
            int j = i;
            for (; (j >= gap) && (local[j - gap] <= temp); j -= gap) local[j] = local[j - gap];
            local[j] = temp;
          }
        }
        bool bget = false;
        float tmpsum = 0, tmax = 0;
        for (int ii = 1; ii <= (int)dim - 1; ii++) {
          tmpsum += local[ii - 1];
          tmax = (tmpsum - 1.) / (float)ii;
          if (tmax >= local[ii]) { bget = true; break; }
        }
        if (!bget) tmax = (tmpsum + local[dim - 1] - 1.0) / (float)dim;
        for (size_t i = 0; i < dim; i++) res[at(tx, i)] = std::max(arg[at(tx, i)] - tmax, 0.f);
      }
    }
  }
};

// helper.hpp:44-105 (x0 and x alias the same planar storage in the caller, like the reference)
static void project_epi_quad_nd(float* xbase, size_t stride, size_t dim, float y0, float alpha, float& y) {
  float sq_norm_x0 = 0;
  for (size_t i = 0; i < dim; i++) sq_norm_x0 += xbase[i * stride] * xbase[i * stride];
  const float norm_x0 = std::sqrt(sq_norm_x0);
  if (y0 >= alpha * sq_norm_x0) { y = y0; return; }
  const float a = 2. * alpha * norm_x0;
  const float b = 2. * (1. - 2. * alpha * y0) / 3.;
  float d, v;
  if (b < 0) {
    const float sq = std::pow(-b, static_cast<float>(3. / 2.));
    d = (a - sq) * (a + sq);
  } else {
    d = a * a + b * b * b;
  }
  if (d >= 0) {
    const float c = std::pow(a + std::sqrt(d), static_cast<float>(1. / 3.));
    if (std::fabs(c) > 1e-6) v = c - b / c;
    else v = 0;
  } else {
    v = 2 * std::sqrt(-b) * std::cos(std::acos(a / std::pow(-b, static_cast<float>(3. / 2.))) / static_cast<float>(3.));
  }
  if (norm_x0 > 0) {
    for (size_t i = 0; i < dim; i++) xbase[i * stride] = (v / (2. * alpha)) * (xbase[i * stride] / norm_x0);
  } else {
    for (size_t i = 0; i < dim; i++) xbase[i * stride] = 0;
  }
  float sq_norm_x = 0;
  for (size_t i = 0; i < dim; i++) sq_norm_x += xbase[i * stride] * xbase[i * stride];
  y = alpha * sq_norm_x;
}

struct ProxIndSum : ProxSeparable {    // elem_operation_ind_sum.hpp:38-58: projection onto { sum_i x_i = 1 } per group
  ProxIndSum(size_t i, size_t c, size_t d, bool il, bool ds) : ProxSeparable(i, c, d, il, ds) {}
  void eval_local(float* res, const float* arg, const float*, float, bool) override {
#pragma omp parallel for schedule(static)
    for (size_t tx = 0; tx < count; ++tx) {
      float tl = 0;
      for (size_t k = 0; k < dim; ++k) tl += arg[at(tx, k)];
      tl = static_cast<float>((tl - 1.) / static_cast<float>(dim));    // `1.` promotes to double (:50)
      for (size_t k = 0; k < dim; ++k) res[at(tx, k)] = arg[at(tx, k)] - tl;
    }
  }
};

// SURVEY.md 8(f) row 2, oracle side only so far (the CUDA kernels follow once these are pinned on a GPU run of the
// reference): projections onto a halfspace and onto the second-order cone, always planar like the reference.
struct ProxIndHalfspace : ProxSeparable {   // prox_ind_halfspace.cu:34-92: { x | <a, x> <= b } per group
  vec a, b;
  ProxIndHalfspace(size_t i, size_t c, size_t d, bool il, bool ds, const float* pa, size_t na, const float* pb, size_t nb)
      : ProxSeparable(i, c, d, il, ds), a(pa, pa + na), b(pb, pb + nb) {}
  void eval_local(float* res, const float* arg, const float*, float, bool) override {
    const bool a_per_group = a.size() == count * dim;     // else one normal of `dim` entries for all groups (:79-88)
#pragma omp parallel for schedule(static)
    for (size_t tx = 0; tx < count; ++tx) {
      const float t = b.size() == count ? b[tx] : b[0];
      auto n = [&](size_t k) { return a_per_group ? a[tx + count * k] : a[k]; };
      float sq_norm = 0, iprod = 0;
      for (size_t k = 0; k < dim; ++k) { sq_norm += n(k) * n(k); iprod += n(k) * arg[tx + count * k]; }   // :42-47
      for (size_t k = 0; k < dim; ++k)
        res[tx + count * k] = arg[tx + count * k] - (std::max(0.f, iprod - t) / sq_norm) * n(k);           // :49-51
    }
  }
};

// ProxIndSum (prox_ind_sum.cu:33-145, prox_ind_sum.hpp:37-62): index-list groups; the zero prox elsewhere; the
// second list's launch is sized by the FIRST list's group count (:135), so only its first
// ceil(count / 256) * 256 groups are ever projected.
struct ProxIndSumIndexed : Prox {
  size_t count[2], dim[2];
  std::vector<size_t> inds[2];
  float total[2];
  bool two;
  ProxIndSumIndexed(size_t i, size_t s, size_t c, size_t d, const unsigned long long* in, float t, size_t c2, size_t d2,
                    const unsigned long long* in2, float t2)
      : Prox(i, s, true), two(in2 != nullptr) {
    count[0] = c; dim[0] = d; total[0] = t; inds[0].assign(in, in + c * d);
    count[1] = c2; dim[1] = d2; total[1] = t2;
    if (two) inds[1].assign(in2, in2 + c2 * d2);
  }
  void eval_local(float* res, const float* arg, const float* td, float tau, bool invert) override {
    std::copy(arg, arg + size, res);                                             // :113-115
    for (int l = 0; l < (two ? 2 : 1); ++l) {
      const size_t n_run = l == 0 ? count[0] : std::min(count[1], (count[0] + 255) / 256 * 256);
      const std::vector<size_t>& ix = inds[l];
      const size_t d = dim[l];
      for (size_t tx = 0; tx < n_run; ++tx) {                                    // ProxIndSumKernel :47-67
        float sum_arg = 0, sum_tau = 0;
        for (size_t k = 0; k < d; ++k) {
          float mytau = td[ix[tx * d + k]] * tau;
          if (invert) mytau = static_cast<float>(1. / mytau);
          sum_arg += arg[ix[tx * d + k]];
          sum_tau += mytau;
        }
        for (size_t k = 0; k < d; ++k) {
          float mytau = td[ix[tx * d + k]] * tau;
          if (invert) mytau = static_cast<float>(1. / mytau);
          res[ix[tx * d + k]] = arg[ix[tx * d + k]] - mytau * (sum_arg - total[l]) / sum_tau;
        }
      }
    }
  }
};

// ProxIndEpiConjQuad1D ("ProxEpiConjQuadr" in the north star).  PARITY UNPINNED: the reference tree only names the
// class (cmake/CustomSources.cmake.example:8-14: ../../preciserelaxation/src/cvpr2016/prost/prox_ind_epi_conjquad_1d.cu,
// no version pinned, not vendored).  Restated from the published definition (Moellenhoff, Laude, Moeller, Lellmann,
// Cremers: Sublabel-accurate relaxation of nonconvex energies, CVPR 2016, eq. (17)-(21): the dual constraint set of
// a piecewise quadratic data term is the epigraph of the conjugate of every piece) on top of the in-tree helpers it
// uses (helper.hpp:112-183 ProjectEpiQuad1d / ProjectEpiQuadGeneral1d, :185-215 halfspace projection).  Pinned only
// by a double-precision brute-force projection (tests/test_oracle_closed_forms.py).
static void project_epi_quad_1d(float x0, float y0, float alpha, float& x, float& y) {      // helper.hpp:112-157
  if (y0 >= alpha * (x0 * x0)) { x = x0; y = y0; return; }
  const float a = static_cast<float>(2. * alpha * std::abs(x0));
  const float b = static_cast<float>(2. * (1. - 2. * alpha * y0) / 3.);
  float d, v;
  if (b < 0) {
    const float sq = std::pow(-b, static_cast<float>(3. / 2.));
    d = (a - sq) * (a + sq);
  } else {
    d = a * a + b * b * b;
  }
  if (d >= 0) {
    const float c = std::pow(a + std::sqrt(d), static_cast<float>(1. / 3.));
    v = c - b / c;
  } else {
    v = 2 * std::sqrt(-b) * std::cos(std::acos(a / std::pow(-b, static_cast<float>(3. / 2.))) / static_cast<float>(3.));
  }
  if (x0 > 0) x = static_cast<float>(v / (2. * alpha));
  else if (x0 < 0) x = static_cast<float>(-v / (2. * alpha));
  else x = 0;
  y = alpha * x * x;
}

struct ProxIndEpiConjQuad1D : ProxSeparable {
  vec co[5];                               // a, b, c, alpha, beta: 1 or count entries
  ProxIndEpiConjQuad1D(size_t i, size_t cnt, bool il, bool ds, const float* const* c, const size_t* len)
      : ProxSeparable(i, cnt, 2, il, ds) {
    for (int k = 0; k < 5; ++k) co[k].assign(c[k], c[k] + len[k]);
  }
  void eval_local(float* res, const float* arg, const float*, float, bool) override {
#pragma omp parallel for schedule(static)
    for (size_t tx = 0; tx < count; ++tx) {
      auto at = [&](int k) { return co[k].size() == 1 ? co[k][0] : co[k][tx]; };
      const float a = at(0), b = at(1), c = at(2), alpha = at(3), beta = at(4);
      const float x0 = arg[this->at(tx, 0)], y0 = arg[this->at(tx, 1)];
      const float x1 = 2 * a * alpha + b, x2 = 2 * a * beta + b;
      const float r1 = (a * alpha + b) * alpha + c, r2 = (a * beta + b) * beta + c;
      const float y1 = alpha * x1 - r1, y2 = beta * x2 - r2;
      const float s1 = (x0 - x1) + alpha * (y0 - y1), s2 = (x0 - x2) + beta * (y0 - y2);
      float x, y;
      if (s1 <= 0) {
        const float viol = std::max(0.f, alpha * x0 - y0 - r1) / (alpha * alpha + 1.f);
        x = x0 - viol * alpha;
        y = y0 + viol;
      } else if (s2 >= 0) {
        const float viol = std::max(0.f, beta * x0 - y0 - r2) / (beta * beta + 1.f);
        x = x0 - viol * beta;
        y = y0 + viol;
      } else if (a > 0) {
        const float p = 1.f / (4 * a), q = -b / (2 * a), r = b * b / (4 * a) - c;
        float tx_, ty_;                                                  // ProjectEpiQuadGeneral1d, helper.hpp:160-183
        project_epi_quad_1d(static_cast<float>(x0 + q / (2. * p)), static_cast<float>(y0 + q * q / (4. * p) - r), p, tx_, ty_);
        x = static_cast<float>(tx_ - q / (2. * p));
        y = static_cast<float>(ty_ - q * q / (4. * p) + r);
      } else {
        const bool inside = y0 >= std::max(alpha * (x0 - b), beta * (x0 - b)) - c;
        x = inside ? x0 : b;
        y = inside ? y0 : -c;
      }
      res[this->at(tx, 0)] = x;
      res[this->at(tx, 1)] = y;
    }
  }
};

// ProxIndRange (prox_ind_range.cu:28-300): x = A (A^T A)^{-1} A^T x0 with A sparse (CSC) and AA = A^T A dense.
// The reference: csrmv with A^T (float), potrf / potrs of AA (float, cusolverDn), csrmv with A.  Restated with float
// products accumulated per row like a CSR SpMV and a Cholesky solve in double (pinned on the closed form of
// test_prox_ind_range.m at its 1e-4 and on the live reference).
struct ProxIndRange : Prox {
  int m, n;
  std::vector<int> ptr, ind;       // CSC of A
  vec val;
  std::vector<double> L;           // Cholesky factor of AA, row-major lower
  ProxIndRange(size_t i, size_t sz, bool ds, int m_, int n_, int nnz, const float* v, const int* p, const int* ix,
               const float* aa)
      : Prox(i, sz, ds), m(m_), n(n_), ptr(p, p + n_ + 1), ind(ix, ix + nnz), val(v, v + nnz), L((size_t)n_ * n_, 0.0) {
    const size_t N = n;
    for (size_t j = 0; j < N; ++j) {
      double d = aa[j * N + j];
      for (size_t k = 0; k < j; ++k) d -= L[j * N + k] * L[j * N + k];
      L[j * N + j] = std::sqrt(d);
      for (size_t r = j + 1; r < N; ++r) {
        double t = aa[j * N + r];
        for (size_t k = 0; k < j; ++k) t -= L[r * N + k] * L[j * N + k];
        L[r * N + j] = t / L[j * N + j];
      }
    }
  }
  void eval_local(float* res, const float* arg, const float*, float, bool) override {
    const size_t N = n;
    std::vector<double> t(N);
    for (size_t c = 0; c < N; ++c) {                               // A^T x0: column c of A dotted with x0
      float acc = 0;
      for (int k = ptr[c]; k < ptr[c + 1]; ++k) acc += val[k] * arg[ind[k]];
      t[c] = acc;
    }
    for (size_t r = 0; r < N; ++r) { double v = t[r]; for (size_t k = 0; k < r; ++k) v -= L[r * N + k] * t[k]; t[r] = v / L[r * N + r]; }
    for (size_t r = N; r-- > 0;) { double v = t[r]; for (size_t k = r + 1; k < N; ++k) v -= L[k * N + r] * t[k]; t[r] = v / L[r * N + r]; }
    std::vector<double> out(m, 0.0);
    for (size_t c = 0; c < N; ++c) {
      const float tc = static_cast<float>(t[c]);
      for (int k = ptr[c]; k < ptr[c + 1]; ++k) out[ind[k]] += static_cast<double>(val[k] * tc);
    }
    for (int r = 0; r < m; ++r) res[r] = static_cast<float>(out[r]);
  }
};

struct ProxIndSOC : ProxSeparable {         // prox_ind_soc.cu:33-77: { (x, y) | |x|_2 <= y }, alpha = 1 only
  ProxIndSOC(size_t i, size_t c, size_t d, bool il, bool ds) : ProxSeparable(i, c, d, il, ds) {}
  void eval_local(float* res, const float* arg, const float*, float, bool) override {
#pragma omp parallel for schedule(static)
    for (size_t tx = 0; tx < count; ++tx) {
      const float y0 = arg[count * (dim - 1) + tx];
      float norm_x0 = 0;
      for (size_t k = 0; k + 1 < dim; ++k) norm_x0 += arg[tx + count * k] * arg[tx + count * k];
      norm_x0 = std::sqrt(norm_x0);
      float fac, y;
      if (norm_x0 <= y0) { fac = 1; y = y0; }                                   // inside the cone
      else if (norm_x0 <= -y0) { fac = 0; y = 0; }                              // inside the polar cone
      else { fac = (y0 + norm_x0) / (2 * norm_x0); y = fac * norm_x0; }
      for (size_t k = 0; k + 1 < dim; ++k)
        res[tx + count * k] = (norm_x0 <= y0) ? arg[tx + count * k] : fac * arg[tx + count * k];
      res[count * (dim - 1) + tx] = y;
    }
  }
};

struct ProxEpiQuad : ProxSeparable {   // prox_ind_epi_quad.cu:42-79 (always planar)
  vec a, b, c;
  ProxEpiQuad(size_t i, size_t cnt, size_t d, bool il, bool ds, const float* a_, size_t na, const float* b_,
              size_t nb, const float* c_, size_t nc)
      : ProxSeparable(i, cnt, d, il, ds), a(a_, a_ + na), b(b_, b_ + nb), c(c_, c_ + nc) {}
  void eval_local(float* res, const float* arg, const float*, float, bool) override {
#pragma omp parallel for schedule(static)
    for (size_t tx = 0; tx < count; ++tx) {
      const size_t dx = dim - 1;
      float* x = res + tx;                       // stride count
      const float y0 = arg[count * dx + tx];
      const float av = a.size() == 1 ? a[0] : a[tx];
      const float cv = c.size() == 1 ? c[0] : c[tx];
      float sq_norm_b = 0;
      for (size_t i = 0; i < dx; i++) {
        const float val = b[tx + count * i];
        x[i * count] = arg[tx + count * i] + (val / (2 * av));
        sq_norm_b += val * val;
      }
      float y;
      project_epi_quad_nd(x, count, dx, y0 - cv + (sq_norm_b / (4 * av)), av, y);
      for (size_t i = 0; i < dx; i++) x[i * count] -= b[tx + count * i] / (2 * av);
      res[count * dx + tx] = y + cv - (sq_norm_b / (4 * av));
    }
  }
};

struct ProxMoreau : Prox {          // prox_moreau.cu:98-134
  std::shared_ptr<Prox> inner;
  vec scaled;
  explicit ProxMoreau(std::shared_ptr<Prox> p) : Prox(p->index, p->size, p->diagsteps), inner(p), scaled(p->size) {}
  void sep(std::vector<std::tuple<size_t, size_t, size_t>>& s) const override { inner->sep(s); }
  void eval_local(float* res, const float* arg, const float* td, float tau, bool invert) override {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < size; ++i) scaled[i] = invert ? arg[i] * (tau * td[i]) : arg[i] / (tau * td[i]);
    inner->eval_local(res, scaled.data(), td, tau, !invert);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < size; ++i) {
      if (invert) res[i] = arg[i] - res[i] / (tau * td[i]);
      else res[i] = arg[i] - tau * td[i] * res[i];
    }
  }
};

struct ProxTransform : Prox {       // prox_transform.cu:27-226: prox of c f(a x - b) + <d, x> + (e/2)|x|^2
  std::shared_ptr<Prox> inner;
  vec co[5], scaled_arg, scaled_tau;   // a, b, c, d, e: one value or one per element
  ProxTransform(std::shared_ptr<Prox> p, const float* const* coeffs, const size_t* len)
      : Prox(p->index, p->size, p->diagsteps), inner(p), scaled_arg(p->size), scaled_tau(p->size) {
    for (int k = 0; k < 5; ++k) co[k].assign(coeffs[k], coeffs[k] + len[k]);
  }
  void sep(std::vector<std::tuple<size_t, size_t, size_t>>& s) const override { inner->sep(s); }
  float at(int k, size_t i) const { return co[k].size() > 1 ? co[k][i] : co[k][0]; }
  void eval_local(float* res, const float* arg, const float* td, float tau, bool invert) override {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < size; ++i) {
      float tau2 = tau * td[i];                                      // :40-42
      if (invert) tau2 = 1 / tau2;
      const float a = at(0, i), b = at(1, i), c = at(2, i), d = at(3, i), e = at(4, i);
      scaled_arg[i] = (a * (arg[i] - tau2 * d)) / (1 + tau2 * e) - b;  // :49
      scaled_tau[i] = (a * a * c * tau2) / (1 + tau2 * e);             // :75
    }
    inner->eval_local(res, scaled_arg.data(), scaled_tau.data(), 1.f, false);   // :201-210
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < size; ++i) res[i] = (res[i] + at(1, i)) / at(0, i);  // :94
  }
};

struct ProxPermute : Prox {         // prox_permute.cu:101-145
  std::shared_ptr<Prox> inner;
  std::vector<int> perm;
  vec permuted;
  ProxPermute(std::shared_ptr<Prox> p, const int* pm, size_t n)
      : Prox(p->index, p->size, p->diagsteps), inner(p), perm(pm, pm + n), permuted(n) {}
  void sep(std::vector<std::tuple<size_t, size_t, size_t>>& s) const override { inner->sep(s); }
  void eval_local(float* res, const float* arg, const float* td, float tau, bool invert) override {
    for (size_t i = 0; i < perm.size(); ++i) res[i] = arg[perm[i]];
    inner->eval_local(permuted.data(), res, td, tau, invert);
    for (size_t i = 0; i < perm.size(); ++i) res[perm[i]] = permuted[i];
  }
};

// ------------------------------------------------------------------------------------------------
// Problem + PDHG
// ------------------------------------------------------------------------------------------------
struct Problem {
  std::vector<std::shared_ptr<Block>> blocks;
  std::vector<std::shared_ptr<Prox>> pool;                  // every prox ever created (ids)
  std::vector<std::shared_ptr<Prox>> prox[4];               // g, f, gstar, fstar
  size_t nrows = 0, ncols = 0, lin_rows = 0, lin_cols = 0;
  bool dims_set = false;
  int scaling = 1;                                          // 0 identity, 1 alpha, 2 custom
  float alpha = 1;
  vec left, right;

  void linop(float* res, const float* rhs, bool transpose) const {     // linearoperator.cu:134-170, beta = 0
    std::fill(res, res + (transpose ? ncols : nrows), 0.f);
    for (auto& b : blocks) {
      if (transpose) b->eval_adj_add(res + b->col, rhs + b->row);
      else b->eval_add(res + b->row, rhs + b->col);
    }
  }
  float row_sum(size_t r, float al) const {
    float s = 0;
    for (auto& b : blocks) { if (r < b->row || r >= b->row + b->nrows) continue; s += b->row_sum(r - b->row, al); }
    return s;
  }
  float col_sum(size_t c, float al) const {
    float s = 0;
    for (auto& b : blocks) { if (c < b->col || c >= b->col + b->ncols) continue; s += b->col_sum(c - b->col, al); }
    return s;
  }
  static void add_zero(std::vector<std::shared_ptr<Prox>>& ps, size_t n) {   // problem.cu:92-158
    if (ps.empty()) return;
    auto s = ps;
    std::sort(s.begin(), s.end(), [](auto& a, auto& b) { return a->index < b->index; });
    if (s[0]->index > 0) ps.push_back(std::make_shared<ProxZero>(0, s[0]->index));
    for (size_t i = 0; i + 1 < s.size(); i++)
      if (s[i]->end() < s[i + 1]->index - 1)
        ps.push_back(std::make_shared<ProxZero>(s[i]->end() + 1, s[i + 1]->index - s[i]->end() - 1));
    if (s.back()->end() < n - 1) ps.push_back(std::make_shared<ProxZero>(s.back()->end() + 1, (n - 1) - s.back()->end()));
  }
  static void average(vec& pre, const std::vector<std::shared_ptr<Prox>>& ps) {   // problem.cu:502-536
    std::vector<std::tuple<size_t, size_t, size_t>> ics;
    for (auto& p : ps) if (!p->diagsteps) p->sep(ics);
    for (auto& t : ics) {
      const size_t idx = std::get<0>(t), cnt = std::get<1>(t), st = std::get<2>(t);
      float avg = 0;
      for (size_t c = 0; c < cnt; c++) avg += pre[idx + c * st];
      avg /= static_cast<float>(cnt);
      for (size_t c = 0; c < cnt; c++) pre[idx + c * st] = avg;
    }
  }
  int initialize() {
    lin_rows = lin_cols = 0;
    for (auto& b : blocks) { lin_rows = std::max(lin_rows, b->row + b->nrows); lin_cols = std::max(lin_cols, b->col + b->ncols); }
    if (!dims_set) { nrows = lin_rows; ncols = lin_cols; }
    if (lin_rows > nrows || lin_cols > ncols) { g_err = "linear operator larger than variables"; return -1; }
    add_zero(prox[1], nrows); add_zero(prox[0], ncols); add_zero(prox[3], nrows); add_zero(prox[2], ncols);
    if (scaling == 1) {                                     // problem.cu:262-287
      left.assign(nrows, 0); right.assign(ncols, 0);
      float value = 1;
      for (size_t r = 0; r < nrows; r++) { const float rs = row_sum(r, alpha); if (rs > 0) value = 1. / rs; left[r] = value; }
      for (size_t c = 0; c < ncols; c++) { const float cs = col_sum(c, 2. - alpha); if (cs > 0) value = 1. / cs; right[c] = value; }
    } else if (scaling == 0) {
      left.assign(nrows, 1); right.assign(ncols, 1);
    } else if (left.size() != nrows || right.size() != ncols) {
      g_err = "custom scaling has wrong size"; return -1;
    }
    average(right, prox[0].empty() ? prox[2] : prox[0]);
    average(left, prox[1].empty() ? prox[3] : prox[1]);
    return 0;
  }
};

struct Pdhg {
  Problem* P;
  // options
  double tau0, sigma0; int residual_iter; float alg2_gamma, arg_alpha0, arg_nu, arg_delta, arb_delta, arb_tau; int variant;
  float tol_rel_p, tol_rel_d, tol_abs_p, tol_abs_d;
  // state (backend_pdhg.hpp:105-156)
  vec x, y, x_prev, y_prev, temp, kx, kty, kx_prev, kty_prev;
  float tau, sigma, theta, arg_alpha; int arb_l, arb_u; size_t iteration;
  float primal_residual = 0, dual_residual = 0, primal_var_norm = 0, dual_var_norm = 0;
  std::vector<std::shared_ptr<Prox>> prox_g, prox_fstar;

  float eps_primal() const { return std::sqrt(P->nrows) * tol_abs_p + tol_rel_p * primal_var_norm; }   // backend.hpp:71
  float eps_dual() const { return std::sqrt(P->ncols) * tol_abs_d + tol_rel_d * dual_var_norm; }       // backend.hpp:74

  int init(const float* x0, size_t nx0, const float* y0, size_t ny0) {     // :199-309 (no normest)
    const size_t m = P->nrows, n = P->ncols;
    x.assign(n, 0); x_prev.assign(n, 0); kty.assign(n, 0); kty_prev.assign(n, 0);
    y.assign(m, 0); y_prev.assign(m, 0); kx.assign(m, 0); kx_prev.assign(m, 0); temp.assign(std::max(m, n), 0);
    iteration = 0; tau = tau0; sigma = sigma0; theta = 1; arb_l = arb_u = 0; arg_alpha = arg_alpha0;
    prox_g.clear(); prox_fstar.clear();
    if (P->prox[0].empty()) { if (P->prox[2].empty()) { g_err = "Neither prox_g nor prox_gstar specified."; return -1; }
      for (auto& p : P->prox[2]) prox_g.push_back(std::make_shared<ProxMoreau>(p)); } else prox_g = P->prox[0];
    if (P->prox[3].empty()) { if (P->prox[1].empty()) { g_err = "Neither prox_f nor prox_fstar specified."; return -1; }
      for (auto& p : P->prox[1]) prox_fstar.push_back(std::make_shared<ProxMoreau>(p)); } else prox_fstar = P->prox[3];
    if (nx0) { if (nx0 != n) { g_err = "Initial primal solution has wrong size."; return -1; } x.assign(x0, x0 + n); x_prev = x; }
    if (ny0) { if (ny0 != m) { g_err = "Initial dual solution has wrong size."; return -1; } y.assign(y0, y0 + m); y_prev = y; }
    return 0;
  }

  void iterate() {                                          // PerformIteration :311-381
    const size_t m = P->nrows, n = P->ncols;
    const float* T = P->right.data();
    const float* S = P->left.data();
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) temp[i] = x[i] - tau * T[i] * kty[i];                       // :38-51
    x.swap(x_prev);
    for (auto& p : prox_g) p->eval(x.data(), temp.data(), T, tau, false);
    kx.swap(kx_prev);
    P->linop(kx.data(), x.data(), false);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < m; ++i) temp[i] = y[i] + sigma * S[i] * ((1 + theta) * kx[i] - theta * kx_prev[i]);   // :54-70
    y.swap(y_prev);
    for (auto& p : prox_fstar) p->eval(y.data(), temp.data(), S, sigma, false);
    update_residuals_and_stepsizes();
    iteration++;
    kty.swap(kty_prev);
    P->linop(kty.data(), y.data(), true);
  }

  void update_residuals_and_stepsizes() {                   // :383-489
    const size_t m = P->nrows, n = P->ncols;
    const float* T = P->right.data();
    const float* S = P->left.data();
    if (iteration == 0 || (iteration % (size_t)(long long)residual_iter) == 0) {
      double p0 = 0, p1 = 0, d0 = 0, d1 = 0;
#pragma omp parallel for schedule(static) reduction(+ : p0, p1)
      for (size_t i = 0; i < m; ++i) {                      // primal_residual_transform :97-120
        const float sd = S[i];
        const float z_hat = (y_prev[i] - y[i]) / (sigma * std::sqrt(sd)) + std::sqrt(sd) * ((1 + theta) * kx[i] - theta * kx_prev[i]);
        const float diff = z_hat - std::sqrt(sd) * kx[i];
        p0 += diff * diff; p1 += z_hat * z_hat;
      }
#pragma omp parallel for schedule(static) reduction(+ : d0, d1)
      for (size_t i = 0; i < n; ++i) {                      // dual_residual_transform :73-94
        const float td = T[i];
        const float w_hat = (x_prev[i] - x[i]) / (tau * std::sqrt(td)) - std::sqrt(td) * kty_prev[i];
        const float diff = w_hat + std::sqrt(td) * kty[i];
        d0 += diff * diff; d1 += w_hat * w_hat;
      }
      primal_residual = std::sqrt((float)p0); primal_var_norm = std::sqrt((float)p1);
      dual_residual = std::sqrt((float)d0); dual_var_norm = std::sqrt((float)d1);
      const float eps_p = eps_primal(), eps_d = eps_dual();
      if (variant == 3) {                                   // Goldstein :443-460
        const float scale = eps_d / eps_p;
        if (dual_residual > (scale * primal_residual * arg_delta)) { tau = tau / (1 - arg_alpha); sigma = sigma * (1 - arg_alpha); arg_alpha = arg_alpha * arg_nu; }
        if (dual_residual < (scale * primal_residual / arg_delta)) { tau = tau * (1 - arg_alpha); sigma = sigma / (1 - arg_alpha); arg_alpha = arg_alpha * arg_nu; }
      } else if (variant == 4) {                            // Boyd :462-476
        if ((dual_residual < eps_d) && (arb_tau * iteration > arb_l)) { tau /= arb_delta; sigma *= arb_delta; arb_u = iteration; }
        else if ((primal_residual < eps_p) && (arb_tau * iteration > arb_u)) { tau *= arb_delta; sigma /= arb_delta; arb_l = iteration; }
      }
    }
    if (variant == 2) {                                     // Alg2 :483-488
      theta = 1. / std::sqrt(1. + 2. * alg2_gamma * tau);
      tau = theta * tau;
      sigma = sigma / theta;
    }
  }

  void solution(float* hx, float* hz, float* hy, float* hw) {   // current_solution :513-563
    const size_t m = P->nrows, n = P->ncols;
    if (hx) std::copy(x.begin(), x.end(), hx);
    if (hy) std::copy(y.begin(), y.end(), hy);
    if (hw) for (size_t i = 0; i < n; ++i) hw[i] = (x_prev[i] - x[i]) / (P->right[i] * tau) - kty_prev[i];
    if (hz) for (size_t i = 0; i < m; ++i) hz[i] = (y_prev[i] - y[i]) / (sigma * P->left[i]) + (1 + theta) * kx[i] - theta * kx_prev[i];
  }
};

// ------------------------------------------------------------------------------------------------
// ADMM (graph projection splitting) + CGLS: src/backend/backend_admm.cu:52-272, 285-665,
// include/prost/cgls.hpp:222-371.  Statement by statement, same temporaries and aliases.
// ------------------------------------------------------------------------------------------------
struct Admm {
  Problem* P;
  // options (backend_admm.hpp:38-63)
  double rho0, alpha, cg_tol_pow, cg_tol_min, cg_tol_max; int cg_max_iter, residual_iter;
  float arb_delta, arb_tau, arb_gamma;
  float tol_rel_p, tol_rel_d, tol_abs_p, tol_abs_d;
  // state (backend_admm.hpp:83-99)
  vec x_half, z_half, x_proj, z_proj, x_dual, z_dual, temp1, temp2, temp3;
  float rho, delta; int arb_u, arb_l; size_t iteration;
  float primal_residual = 0, dual_residual = 0, primal_var_norm = 0, dual_var_norm = 0;
  std::vector<std::shared_ptr<Prox>> prox_g, prox_f;
  int last_cg_iters = 0;
  long total_cg_iters = 0;

  float eps_primal() const { return std::sqrt(P->nrows) * tol_abs_p + tol_rel_p * primal_var_norm; }   // backend.hpp:71
  float eps_dual() const { return std::sqrt(P->ncols) * tol_abs_d + tol_rel_d * dual_var_norm; }       // backend.hpp:74

  int init() {                                              // Initialize :285-352
    const size_t m = P->nrows, n = P->ncols;
    x_half.assign(n, 0); x_proj.assign(n, 0); x_dual.assign(n, 0);
    z_half.assign(m, 0); z_proj.assign(m, 0); z_dual.assign(m, 0);
    temp1.assign(n, 0); temp2.assign(m, 0); temp3.assign(std::max(m, n), 0);
    prox_g.clear(); prox_f.clear();
    if (P->prox[0].empty()) { if (P->prox[2].empty()) { g_err = "Neither prox_g nor prox_gstar specified."; return -1; }
      for (auto& p : P->prox[2]) prox_g.push_back(std::make_shared<ProxMoreau>(p)); } else prox_g = P->prox[0];
    if (P->prox[1].empty()) { if (P->prox[3].empty()) { g_err = "Neither prox_f nor prox_fstar specified."; return -1; }
      for (auto& p : P->prox[3]) prox_f.push_back(std::make_shared<ProxMoreau>(p)); } else prox_f = P->prox[1];
    delta = arb_delta; rho = rho0; iteration = 0; arb_u = arb_l = 0;
    return 0;
  }

  // linearoperator.cu:134-170 with a general beta
  void linop(float* res, const float* rhs, float beta, bool transpose) const {
    const size_t nout = transpose ? P->ncols : P->nrows;
    if (beta == 0) std::fill(res, res + nout, 0.f);
    else if (beta != 1) for (size_t i = 0; i < nout; ++i) res[i] = beta * res[i];
    for (auto& b : P->blocks) {
      if (transpose) b->eval_adj_add(res + b->col, rhs + b->row);
      else b->eval_add(res + b->row, rhs + b->col);
    }
  }

  // GemvPrecondK::operator() :198-272: y := alpha op(Sigma^{1/2} K Tau^{1/2}) x + beta y
  void gemv(char op, float al, const vec& x, float be, vec& y) {
    const size_t m = P->nrows, n = P->ncols;
    const float* T = P->right.data();
    const float* S = P->left.data();
    if (op == 'n') {
      for (size_t i = 0; i < n; ++i) temp3[i] = std::sqrt(T[i]) * x[i];                       // gemv_functor1
      for (size_t i = 0; i < m; ++i) y[i] = (be / (al * std::sqrt(S[i]))) * y[i];             // gemv_functor2
      linop(y.data(), temp3.data(), 1, false);
      for (size_t i = 0; i < m; ++i) y[i] = al * std::sqrt(S[i]) * y[i];                      // gemv_functor3
    } else {
      for (size_t i = 0; i < m; ++i) temp3[i] = std::sqrt(S[i]) * x[i];
      for (size_t i = 0; i < n; ++i) y[i] = (be / (al * std::sqrt(T[i]))) * y[i];
      linop(y.data(), temp3.data(), 1, true);
      for (size_t i = 0; i < n; ++i) y[i] = al * std::sqrt(T[i]) * y[i];
    }
  }

  static double nrm2d(const vec& v, size_t n) {             // cgls.hpp:166-189 (double accumulation)
    double s = 0;
    for (size_t i = 0; i < n; ++i) s += static_cast<double>(v[i]) * static_cast<double>(v[i]);
    return std::sqrt(s);
  }
  static float nrm2f(const vec& v, size_t n) {              // backend_admm.cu:46-50 (cublasSnrm2 -> float)
    double s = 0;
    for (size_t i = 0; i < n; ++i) s += static_cast<double>(v[i]) * static_cast<double>(v[i]);
    return static_cast<float>(std::sqrt(s));
  }

  // cgls::Solve :222-371 with shift = 1 on the preconditioned operator
  int cgls(const vec& b, vec& x, double shift, double tol, int maxit, vec& p, vec& q, vec& r, vec& s, int& iterations) {
    const size_t m = P->nrows, n = P->ncols;
    double gamma, normp, normq, norms, norms0, normx, xmax;
    int k = 0, flag = 0, indefinite = 0;
    const float kNegOne = -1.f, kZero = 0.f, kOne = 1.f, kNegShift = static_cast<float>(-shift);
    const double kEps = std::numeric_limits<float>::epsilon();
    std::copy(b.begin(), b.begin() + m, r.begin());
    std::copy(x.begin(), x.begin() + n, s.begin());
    normx = nrm2d(x, n);
    if (normx > 0.) gemv('n', kNegOne, x, kOne, r);
    gemv('t', kOne, r, kNegShift, s);
    std::copy(s.begin(), s.begin() + n, p.begin());
    norms = nrm2d(s, n);
    norms0 = norms;
    gamma = norms0 * norms0;
    normx = nrm2d(x, n);
    xmax = normx;
    if (norms < kEps) flag = 1;
    for (k = 0; k < maxit && !flag; ++k) {
      gemv('n', kOne, p, kZero, q);
      normp = nrm2d(p, n);
      normq = nrm2d(q, m);
      double delta_ = normq * normq + shift * normp * normp;
      if (delta_ <= 0.) indefinite = 1;
      if (delta_ == 0.) delta_ = kEps;
      const float al = static_cast<float>(gamma / delta_);
      const float neg_al = static_cast<float>(-gamma / delta_);
      for (size_t i = 0; i < n; ++i) x[i] = al * p[i] + x[i];          // cublasSaxpy
      for (size_t i = 0; i < m; ++i) r[i] = neg_al * q[i] + r[i];
      std::copy(x.begin(), x.begin() + n, s.begin());
      gemv('t', kOne, r, kNegShift, s);
      norms = nrm2d(s, n);
      const double gamma1 = gamma;
      gamma = norms * norms;
      const float be = static_cast<float>(gamma / gamma1);
      for (size_t i = 0; i < n; ++i) s[i] = be * p[i] + s[i];
      std::copy(s.begin(), s.begin() + n, p.begin());
      normx = nrm2d(x, n);
      xmax = std::max(xmax, normx);
      const bool converged = (norms <= norms0 * tol) || (normx * tol >= 1.);
      if (converged) break;
    }
    const double shrink = normx / xmax;
    if (k == maxit) flag = 2;
    else if (indefinite) flag = 3;
    else if (shrink * shrink <= tol) flag = 4;
    iterations = k;
    return flag;
  }

  void iterate() {                                          // PerformIteration :354-665
    const size_t m = P->nrows, n = P->ncols;
    const float* T = P->right.data();
    const float* S = P->left.data();
    const float al = static_cast<float>(alpha);
    for (size_t i = 0; i < n; ++i)                          // temp1_functor :52-66
      temp1[i] = (al * x_half[i] + (1 - al) * x_proj[i] + x_dual[i]) / std::sqrt(T[i]);
    for (size_t i = 0; i < m; ++i)                          // temp2_functor :68-79
      temp2[i] = std::sqrt(S[i]) * (z_half[i] + z_dual[i]);
    std::copy(temp2.begin(), temp2.end(), z_dual.begin());  // :394-395, tmp_proj_arg aliases z_dual
    vec& tmp_proj_arg = z_dual;
    std::copy(temp3.begin(), temp3.begin() + n, x_proj.begin());   // warm start :398
    gemv('n', -1, temp1, 1, tmp_proj_arg);                  // :401
    double cg_tol = cg_tol_min / std::pow(static_cast<float>(iteration + 1), cg_tol_pow);   // :403-405
    cg_tol = std::max(cg_tol, cg_tol_max);
    vec& tmp_p = x_half; vec& tmp_q = z_half; vec& tmp_r = z_proj; vec& tmp_s = x_dual;     // :411-414
    cgls(tmp_proj_arg, x_proj, 1, cg_tol, cg_max_iter, tmp_p, tmp_q, tmp_r, tmp_s, last_cg_iters);
    total_cg_iters += last_cg_iters;
    std::copy(x_proj.begin(), x_proj.end(), temp3.begin()); // :439
    for (size_t i = 0; i < n; ++i) x_proj[i] = std::sqrt(T[i]) * (x_proj[i] + temp1[i]);    // x_proj_functor
    linop(z_proj.data(), x_proj.data(), 0, false);          // :456
    for (size_t i = 0; i < n; ++i) x_dual[i] = temp1[i] * std::sqrt(T[i]) - x_proj[i];      // x_dual_functor
    for (size_t i = 0; i < m; ++i) z_dual[i] = temp2[i] / std::sqrt(S[i]) - z_proj[i];      // z_dual_functor
    for (size_t i = 0; i < n; ++i) temp1[i] = x_proj[i] - x_dual[i];
    for (auto& p : prox_g) p->eval(x_half.data(), temp1.data(), T, 1 / rho, false);         // :505-506
    for (size_t i = 0; i < m; ++i) temp2[i] = z_proj[i] - z_dual[i];
    for (auto& p : prox_f) p->eval(z_half.data(), temp2.data(), S, rho, true);              // :522-523
    iteration++;

    if (iteration == 0 || (iteration % (size_t)(long long)residual_iter) == 0) {            // :529
      std::copy(z_half.begin(), z_half.end(), temp2.begin());
      linop(temp2.data(), x_half.data(), -1, false);        // temp2 = K x_half - z_half
      for (size_t i = 0; i < m; ++i) temp2[i] = std::sqrt(S[i]) * temp2[i];
      primal_residual = nrm2f(temp2, m);
      for (size_t i = 0; i < m; ++i) temp2[i] = std::sqrt(S[i]) * z_half[i];
      primal_var_norm = nrm2f(temp2, m);
      for (size_t i = 0; i < n; ++i)                        // get_dual_functor(rho, -1) :184-196
        temp1[i] = -rho * std::pow(T[i], -1.f) * (x_half[i] - x_proj[i] + x_dual[i]);
      {
        vec tw(n);                                          // the reference writes these n values into temp2_ (needs m >= n)
        for (size_t i = 0; i < n; ++i) tw[i] = std::sqrt(T[i]) * temp1[i];
        dual_var_norm = nrm2f(tw, n);
      }
      for (size_t i = 0; i < m; ++i)                        // get_dual_functor(rho, 1)
        temp2[i] = -rho * std::pow(S[i], 1.f) * (z_half[i] - z_proj[i] + z_dual[i]);
      linop(temp1.data(), temp2.data(), 1, true);           // w + K^T y
      for (size_t i = 0; i < n; ++i) temp1[i] = std::sqrt(T[i]) * temp1[i];
      dual_residual = nrm2f(temp1, n);
      const float eps_p = eps_primal(), eps_d = eps_dual();
      const float rho_prev = rho;
      if ((dual_residual < eps_d) && (arb_tau * iteration > arb_l)) { rho *= delta; delta *= arb_gamma; arb_u = iteration; }
      else if ((primal_residual < eps_p) && (arb_tau * iteration > arb_u)) { rho /= delta; delta *= arb_gamma; arb_l = iteration; }
      if (std::abs(rho - rho_prev) > 1e-7) {                // :646-659
        const float f = rho_prev / rho;
        for (size_t i = 0; i < n; ++i) x_dual[i] = f * x_dual[i];
        for (size_t i = 0; i < m; ++i) z_dual[i] = f * z_dual[i];
      }
    }
  }

  void solution(float* hx, float* hz, float* hy, float* hw) {   // current_solution :697-741
    const size_t m = P->nrows, n = P->ncols;
    const float* T = P->right.data();
    const float* S = P->left.data();
    if (hw) for (size_t i = 0; i < n; ++i) hw[i] = -rho * std::pow(T[i], -1.f) * (x_half[i] - x_proj[i] + x_dual[i]);
    if (hy) for (size_t i = 0; i < m; ++i) hy[i] = -rho * std::pow(S[i], 1.f) * (z_half[i] - z_proj[i] + z_dual[i]);
    if (hx) std::copy(x_half.begin(), x_half.end(), hx);
    if (hz) std::copy(z_half.begin(), z_half.end(), hz);
  }
};

}  // namespace orc

// ------------------------------------------------------------------------------------------------
// C interface for ctypes
// ------------------------------------------------------------------------------------------------
using namespace orc;
extern "C" {

const char* orc_last_error() { return g_err.c_str(); }
int orc_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}

void* orc_problem_new() { return new Problem(); }
void orc_problem_free(void* p) { delete static_cast<Problem*>(p); }
#define PP static_cast<Problem*>(p)

void orc_add_gradient(void* p, int three_d, size_t row, size_t col, size_t nx, size_t ny, size_t L, int lf) {
  PP->blocks.push_back(std::make_shared<BlockGradient>(row, col, nx, ny, L, lf != 0, three_d != 0));
}
void orc_add_diags(void* p, size_t row, size_t col, size_t nr, size_t nc, size_t nd, const int64_t* o, const float* f) {
  PP->blocks.push_back(std::make_shared<BlockDiags>(row, col, nr, nc, nd, o, f));
}
void orc_add_sparse_csc(void* p, size_t row, size_t col, int m, int n, int nnz, const float* v, const int32_t* ptr, const int32_t* ind) {
  PP->blocks.push_back(std::make_shared<BlockSparse>(row, col, m, n, nnz, v, ptr, ind));
}
void orc_add_dense(void* p, size_t row, size_t col, size_t nr, size_t nc, const float* d) {
  PP->blocks.push_back(std::make_shared<BlockDense>(row, col, nr, nc, d));
}
void orc_add_kron(void* p, size_t row, size_t col, size_t mr, size_t mc, size_t d, int id_first, const float* data) {
  PP->blocks.push_back(std::make_shared<BlockKron>(row, col, mr, mc, d, id_first != 0, data));
}
void orc_add_zero(void* p, size_t row, size_t col, size_t nr, size_t nc) {
  PP->blocks.push_back(std::make_shared<BlockZero>(row, col, nr, nc));
}

static int push(Problem* P, std::shared_ptr<Prox> q) { P->pool.push_back(q); return (int)P->pool.size() - 1; }
int orc_prox_elem(void* p, int norm2, size_t idx, size_t count, size_t dim, int il, int ds, int fn,
                  const float* const* coeffs, const size_t* len) {
  return push(PP, std::make_shared<ProxElem7>(norm2 != 0, idx, count, dim, il != 0, ds != 0, fn, coeffs, len));
}
int orc_prox_simplex(void* p, size_t idx, size_t count, size_t dim, int il, int ds) {
  return push(PP, std::make_shared<ProxSimplex>(idx, count, dim, il != 0, ds != 0));
}
int orc_prox_ind_sum(void* p, size_t idx, size_t count, size_t dim, int il, int ds) {
  return push(PP, std::make_shared<ProxIndSum>(idx, count, dim, il != 0, ds != 0));
}
int orc_prox_ind_halfspace(void* p, size_t idx, size_t count, size_t dim, int il, int ds, const float* a, size_t na,
                           const float* b, size_t nb) {
  return push(PP, std::make_shared<ProxIndHalfspace>(idx, count, dim, il != 0, ds != 0, a, na, b, nb));
}
int orc_prox_ind_sum_indexed(void* p, size_t idx, size_t size, size_t count, size_t dim, const unsigned long long* inds,
                             float total, size_t count2, size_t dim2, const unsigned long long* inds2, float total2) {
  return push(PP, std::make_shared<ProxIndSumIndexed>(idx, size, count, dim, inds, total, count2, dim2, inds2, total2));
}
int orc_prox_ind_epi_conjquad_1d(void* p, size_t idx, size_t count, int il, int ds, const float* const* coeffs,
                                 const size_t* len) {
  return push(PP, std::make_shared<ProxIndEpiConjQuad1D>(idx, count, il != 0, ds != 0, coeffs, len));
}
int orc_prox_spectral(void* p, int kind, size_t idx, size_t count, size_t dim, int il, int ds, int fn, int fn2d,
                      const float* const* coeffs, const size_t* len) {
  return push(PP, std::make_shared<ProxSpectral>(kind, idx, count, dim, il != 0, ds != 0, fn, fn2d, coeffs, len));
}
int orc_prox_ind_range(void* p, size_t idx, size_t size, int ds, int m, int n, int nnz, const float* val, const int* ptr,
                       const int* ind, const float* aa) {
  return push(PP, std::make_shared<ProxIndRange>(idx, size, ds != 0, m, n, nnz, val, ptr, ind, aa));
}
int orc_prox_ind_soc(void* p, size_t idx, size_t count, size_t dim, int il, int ds) {
  return push(PP, std::make_shared<ProxIndSOC>(idx, count, dim, il != 0, ds != 0));
}
int orc_prox_epi_quad(void* p, size_t idx, size_t count, size_t dim, int il, int ds, const float* a, size_t na,
                      const float* b, size_t nb, const float* c, size_t nc) {
  return push(PP, std::make_shared<ProxEpiQuad>(idx, count, dim, il != 0, ds != 0, a, na, b, nb, c, nc));
}
int orc_prox_moreau(void* p, int inner) { return push(PP, std::make_shared<ProxMoreau>(PP->pool[inner])); }
int orc_prox_permute(void* p, int inner, const int* perm, size_t n) {
  return push(PP, std::make_shared<ProxPermute>(PP->pool[inner], perm, n));
}
int orc_prox_transform(void* p, int inner, const float* const* coeffs, const size_t* len) {
  return push(PP, std::make_shared<ProxTransform>(PP->pool[inner], coeffs, len));
}
int orc_prox_zero(void* p, size_t idx, size_t size) { return push(PP, std::make_shared<ProxZero>(idx, size)); }
void orc_set_prox(void* p, int which, int id) { PP->prox[which].push_back(PP->pool[id]); }
void orc_set_dims(void* p, size_t nrows, size_t ncols) { PP->nrows = nrows; PP->ncols = ncols; PP->dims_set = true; }
void orc_set_scaling_alpha(void* p, float a) { PP->scaling = 1; PP->alpha = a; }
void orc_set_scaling_identity(void* p) { PP->scaling = 0; }
void orc_set_scaling_custom(void* p, const float* l, size_t nl, const float* r, size_t nr) {   // problem.cu:344-364
  PP->scaling = 2;
  PP->left.resize(nl); PP->right.resize(nr);
  for (size_t i = 0; i < nl; ++i) PP->left[i] = l[i] * l[i];
  for (size_t i = 0; i < nr; ++i) PP->right[i] = r[i] * r[i];
}
int orc_initialize(void* p) { return PP->initialize(); }
size_t orc_nrows(void* p) { return PP->nrows; }
size_t orc_ncols(void* p) { return PP->ncols; }
// sizes of the operator alone (usable before orc_initialize)
void orc_linop_size(void* p, size_t* nrows, size_t* ncols) {
  size_t r = 0, c = 0;
  for (auto& b : PP->blocks) { r = std::max(r, b->row + b->nrows); c = std::max(c, b->col + b->ncols); }
  *nrows = r; *ncols = c;
}
void orc_linop_eval(void* p, float* res, const float* rhs, int transpose) {
  if (PP->nrows == 0 && PP->ncols == 0) orc_linop_size(p, &PP->nrows, &PP->ncols);
  PP->linop(res, rhs, transpose != 0);
}
void orc_row_sums(void* p, float alpha, float* out, size_t n) { for (size_t r = 0; r < n; ++r) out[r] = PP->row_sum(r, alpha); }
void orc_col_sums(void* p, float alpha, float* out, size_t n) { for (size_t c = 0; c < n; ++c) out[c] = PP->col_sum(c, alpha); }
void orc_get_scaling(void* p, float* l, float* r) {
  std::copy(PP->left.begin(), PP->left.end(), l);
  std::copy(PP->right.begin(), PP->right.end(), r);
}
void orc_prox_eval(void* p, int id, float* res, const float* arg, const float* td, float tau, int invert) {
  PP->pool[id]->eval(res, arg, td, tau, invert != 0);
}

void* orc_pdhg_new(void* p, double tau0, double sigma0, int residual_iter, float alg2_gamma, float arg_alpha0,
                   float arg_nu, float arg_delta, float arb_delta, float arb_tau, int variant, float tol_rel_p,
                   float tol_rel_d, float tol_abs_p, float tol_abs_d) {
  Pdhg* s = new Pdhg();
  s->P = PP;
  s->tau0 = tau0; s->sigma0 = sigma0; s->residual_iter = residual_iter; s->alg2_gamma = alg2_gamma;
  s->arg_alpha0 = arg_alpha0; s->arg_nu = arg_nu; s->arg_delta = arg_delta; s->arb_delta = arb_delta;
  s->arb_tau = arb_tau; s->variant = variant;
  s->tol_rel_p = tol_rel_p; s->tol_rel_d = tol_rel_d; s->tol_abs_p = tol_abs_p; s->tol_abs_d = tol_abs_d;
  return s;
}
#define SS static_cast<Pdhg*>(s)
void orc_pdhg_free(void* s) { delete SS; }
int orc_pdhg_init(void* s, const float* x0, size_t nx0, const float* y0, size_t ny0) { return SS->init(x0, nx0, y0, ny0); }
void orc_pdhg_iterate(void* s, int n) { for (int i = 0; i < n; ++i) SS->iterate(); }
void orc_pdhg_residuals(void* s, float* out) {
  out[0] = SS->primal_residual; out[1] = SS->dual_residual; out[2] = SS->primal_var_norm; out[3] = SS->dual_var_norm;
  out[4] = SS->eps_primal(); out[5] = SS->eps_dual();
}
void orc_pdhg_stepsizes(void* s, double* out) { out[0] = SS->tau; out[1] = SS->sigma; out[2] = SS->theta; }
void orc_pdhg_solution(void* s, float* x, float* z, float* y, float* w) { SS->solution(x, z, y, w); }

void* orc_admm_new(void* p, double rho0, double alpha, double cg_tol_pow, double cg_tol_min, double cg_tol_max,
                   int cg_max_iter, int residual_iter, float arb_delta, float arb_tau, float arb_gamma,
                   float tol_rel_p, float tol_rel_d, float tol_abs_p, float tol_abs_d) {
  Admm* s = new Admm();
  s->P = PP;
  s->rho0 = rho0; s->alpha = alpha; s->cg_tol_pow = cg_tol_pow; s->cg_tol_min = cg_tol_min; s->cg_tol_max = cg_tol_max;
  s->cg_max_iter = cg_max_iter; s->residual_iter = residual_iter; s->arb_delta = arb_delta; s->arb_tau = arb_tau;
  s->arb_gamma = arb_gamma;
  s->tol_rel_p = tol_rel_p; s->tol_rel_d = tol_rel_d; s->tol_abs_p = tol_abs_p; s->tol_abs_d = tol_abs_d;
  return s;
}
#define AA static_cast<Admm*>(s)
void orc_admm_free(void* s) { delete AA; }
int orc_admm_init(void* s) { return AA->init(); }
void orc_admm_iterate(void* s, int n) { for (int i = 0; i < n; ++i) AA->iterate(); }
void orc_admm_residuals(void* s, float* out) {
  out[0] = AA->primal_residual; out[1] = AA->dual_residual; out[2] = AA->primal_var_norm; out[3] = AA->dual_var_norm;
  out[4] = AA->eps_primal(); out[5] = AA->eps_dual();
}
void orc_admm_stepsizes(void* s, double* out) { out[0] = AA->rho; out[1] = AA->delta; out[2] = (double)AA->total_cg_iters; }
void orc_admm_solution(void* s, float* x, float* z, float* y, float* w) { AA->solution(x, z, y, w); }

}  // extern "C"
