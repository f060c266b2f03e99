// pb_prox.cuh -- separable proximal operators: POD descriptors, the in-register group
// operations, the generic pass kernel (shared by the unfused Prox::Eval path and the fused
// PDHG passes) and the host-side prox objects.
//
// Reference: ProxElemOperationKernel + Vector<T> addressing (prox_elem_operation.inl:32-94,
// vector.hpp:42-48), one thread per group of `dim` elements; ProxMoreau (prox_moreau.cu:98-134)
// is folded into the same kernel as a pre/post scaling in registers instead of two extra
// passes over a scratch vector.
#pragma once

#include <memory>
#include <tuple>
#include <vector>

#include "pb_common.cuh"
#include "pb_math.cuh"

namespace pb {

enum ProxKind : int {
  kProxZero = 0,
  kProxElem1D = 1,
  kProxNorm2 = 2,
  kProxSimplex = 3,
  kProxEpiQuad = 4,
  kProxMoreau = 5,
  kProxPermute = 6,
  kProxTransform = 7,
  kProxIndSum = 8,
  kProxIndHalfspace = 9,
  kProxIndSOC = 10,
  kProxIndSumIndexed = 11,
  kProxIndEpiConjQuad1D = 12,
  kProxSpectral = 13,
  kProxIndRange = 14,
};

// per-element vector or scalar (ElemOpCoefficients: prox_elem_operation.hpp:104-109)
struct CoeffRef {
  const float* ptr[7];
  float val[7];
};

// diagonal preconditioner entries: per-element vector, or one scalar when uniform
struct ScaleRef {
  const float* ptr;
  float val;
  __device__ __forceinline__ float at(uint32_t e) const { return ptr ? __ldg(ptr + e) : val; }
};

// POD view of a leaf prox (optionally evaluated through Moreau's identity)
struct ProxDesc {
  int kind = kProxZero;
  int fn = 0;                 // pb_function1d for Elem1D / Norm2
  uint32_t index = 0, count = 0, dim = 1;
  int interleaved = 0;
  int moreau = 0;
  CoeffRef coeffs = {};
  // epigraph-quadratic coefficients: a, c scalar-or-vector (count), b vector count*(dim-1), planar
  const float* epi_a = nullptr; const float* epi_b = nullptr; const float* epi_c = nullptr;
  float epi_a_val = 0.f, epi_c_val = 0.f;
};

constexpr int kMaxRegDim = 64;   // largest group held in registers; larger dims use the slow path
constexpr int kMaxEpiRegDim = 8; // epigraph projections are instantiated for dim <= 8 only

// smallest instantiated register capacity >= dim for a prox kind, 0 if none
inline int dim_cap(size_t dim, int kind = kProxZero) {
  if (kind == kProxEpiQuad && dim > (size_t)kMaxEpiRegDim) return 0;
  const int caps[] = {1, 2, 4, 8, 16, 32, 64};
  for (int c : caps)
    if ((size_t)c >= dim) return c;
  return 0;
}

#ifdef __CUDACC__

// global element index of component i of group tx (Vector<T>::operator[], vector.hpp:42-48;
// ProxIndEpiQuad is always planar, prox_ind_epi_quad.cu:54-57)
__device__ __forceinline__ uint32_t elem_index(const ProxDesc& p, uint32_t tx, uint32_t i) {
  const bool il = p.interleaved && p.kind != kProxEpiQuad;
  return p.index + (il ? tx * p.dim + i : tx + p.count * i);
}

__device__ __forceinline__ void load_coeffs(const CoeffRef& c, uint32_t tx, Coeffs7& out) {
#pragma unroll
  for (int k = 0; k < 7; ++k) out.v[k] = c.ptr[k] ? __ldg(c.ptr[k] + tx) : c.val[k];
}

// Simplex projection of v[0:dim] (elem_operation_ind_simplex.hpp:47-115): sort descending,
// first i with (sum_{k<i} s_k - 1)/i >= s_i, else (sum - 1)/dim; result max(v - t, 0).
// The sort is a compare-exchange network on a register copy (the reference shell-sorts a
// 4 KB per-thread local array); the scan is the reference's, including the double `1.`.
template <int CAP>
__device__ __forceinline__ float simplex_threshold(const float (&v)[CAP], uint32_t dim) {
  float s[CAP];
#pragma unroll
  for (int i = 0; i < CAP; ++i) s[i] = (i < (int)dim) ? v[i] : -INFINITY;
  // odd-even merge sort would be fewer exchanges; bitonic keeps the index math trivial and
  // fully unrollable: CAP is a power of two.
#pragma unroll
  for (int k = 2; k <= CAP; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
      for (int i = 0; i < CAP; ++i) {
        const int l = i ^ j;
        if (l > i) {
          const bool desc = ((i & k) == 0);     // descending overall
          const float a = s[i], b = s[l];
          const float hi = fmaxf(a, b), lo = fminf(a, b);
          s[i] = desc ? hi : lo;
          s[l] = desc ? lo : hi;
        }
      }
    }
  }
  bool found = false;
  float tmpsum = 0.f, tmax = 0.f;
#pragma unroll
  for (int ii = 1; ii < CAP; ++ii) {
    if (!found && ii <= (int)dim - 1) {
      tmpsum += s[ii - 1];
      tmax = static_cast<float>((static_cast<double>(tmpsum) - 1.0) / static_cast<double>((float)ii));
      if (tmax >= s[ii]) found = true;
    }
  }
  if (!found) {
    float last = s[0];
#pragma unroll
    for (int i = 1; i < CAP; ++i)
      if (i == (int)dim - 1) last = s[i];
    tmax = static_cast<float>((static_cast<double>(tmpsum + last) - 1.0) /
                              static_cast<double>((float)dim));
  }
  return tmax;
}

// Leaf operation on one group held in registers: v (argument) is replaced by the prox.
// td0 = tau_diag of the group's first component (all in-tree ops read only tau_diag[0],
// Appendix B #10).
// KIND >= 0 fixes the prox kind at compile time (specialised kernels); KIND < 0 dispatches on
// p.kind at run time (generic kernels).
// FN >= 0 likewise fixes the Function1D member of the 1D / Norm2 families.
// `pre` (optional) supplies coefficients already held in registers (uniform scalars in the
// specialised kernels); otherwise they are fetched per group with load_coeffs.
template <int CAP, int KIND = -1, int FN = -1>
__device__ __forceinline__ void leaf_apply(const ProxDesc& p, uint32_t tx, float (&v)[CAP],
                                           float tau_scal, float td0, bool invert,
                                           const Coeffs7* pre = nullptr) {
  const uint32_t dim = p.dim;
  const int kind = KIND >= 0 ? KIND : p.kind;
  const int fn = FN >= 0 ? FN : p.fn;
  switch (kind) {
    case kProxElem1D: {
      if (CAP == 1) {                       // dim is always 1 (ElemOperation1D::kDim)
        Coeffs7 c;
        if (pre) c = *pre; else load_coeffs(p.coeffs, tx, c);
        v[0] = elem1d_apply(fn, v[0], tau_scal, td0, invert, c);
      }
      break;
    }
    case kProxNorm2: {
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < CAP; ++i)
        if (i < (int)dim) sq += v[i] * v[i];
      if (sq > 0.f) {
        const float norm = sqrtf(sq);
        Coeffs7 c;
        if (pre) c = *pre; else load_coeffs(p.coeffs, tx, c);
        const float tau = effective_tau(tau_scal, td0, invert);
        const float r = scaled_fun_prox(fn, norm, tau, c);
#pragma unroll
        for (int i = 0; i < CAP; ++i)
          if (i < (int)dim) v[i] = r * v[i] / norm;
      } else {
#pragma unroll
        for (int i = 0; i < CAP; ++i) v[i] = 0.f;
      }
      break;
    }
    case kProxSimplex: {
      const float t = simplex_threshold<CAP>(v, dim);
#pragma unroll
      for (int i = 0; i < CAP; ++i)
        if (i < (int)dim) v[i] = fmaxf(v[i] - t, 0.f);
      break;
    }
    case kProxEpiQuad: if (CAP >= 2 && CAP <= kMaxEpiRegDim) {
      // prox_ind_epi_quad.cu:42-79: shift by b/(2a), project onto y >= a|x|^2, shift back.
      // components 0..dim-2 are x, component dim-1 is y.
      const float a = p.epi_a ? __ldg(p.epi_a + tx) : p.epi_a_val;
      const float c = p.epi_c ? __ldg(p.epi_c + tx) : p.epi_c_val;
      const SharedDivisor by_2a(2 * a);       // b / (2a), v / (2a), sqb / (4a): exact reciprocal when a is 2^k
      float bb[CAP];
      float sqb = 0.f, sqx = 0.f, y0 = 0.f;
#pragma unroll
      for (int i = 0; i < CAP; ++i) {
        bb[i] = 0.f;
        if (i < (int)dim - 1) {
          const float b = __ldg(p.epi_b + tx + (size_t)p.count * i);
          bb[i] = by_2a(b);
          v[i] = v[i] + bb[i];
          sqb += b * b;
          sqx += v[i] * v[i];
        } else if (i == (int)dim - 1) {
          y0 = v[i];
        }
      }
      const float shift = by_2a.pow2 ? sqb * (0.5f * by_2a.r) : sqb / (4 * a);
      const float ys = y0 - c + shift;
      float vv;
      bool inside;
      project_epi_quad(sqx, ys, a, vv, inside);
      float y;
      if (inside) {
        y = ys;
      } else {
        const float norm = sqrtf(sqx);
        float sq_new = 0.f;
        const double scale = by_2a.quotient(static_cast<double>(vv));
#pragma unroll
        for (int i = 0; i < CAP; ++i)
          if (i < (int)dim - 1) {
            v[i] = (norm > 0.f) ? static_cast<float>(scale * static_cast<double>(v[i] / norm)) : 0.f;
            sq_new += v[i] * v[i];
          }
        y = a * sq_new;
      }
#pragma unroll
      for (int i = 0; i < CAP; ++i) {
        if (i < (int)dim - 1) v[i] -= bb[i];
        else if (i == (int)dim - 1) v[i] = y + c - shift;
      }
      break;
    } else break;
    default: break;   // kProxZero: identity
  }
}

// Leaf, or Moreau's identity around the leaf (prox_moreau.cu:29-61, 110-133):
//   s = arg / (tau T)  (arg * tau T when inverted);  r = prox_leaf(s; !invert);
//   res = arg - tau T r  (arg - r / (tau T) when inverted),  T per component.
template <int CAP, int KIND = -1, int FN = -1>
__device__ __forceinline__ void group_apply(const ProxDesc& p, uint32_t tx, float (&v)[CAP],
                                            const float (&td)[CAP], float tau_scal, bool invert,
                                            const Coeffs7* pre = nullptr) {
  if (!p.moreau) {
    leaf_apply<CAP, KIND, FN>(p, tx, v, tau_scal, td[0], invert, pre);
    return;
  }
  float a[CAP];
#pragma unroll
  for (int i = 0; i < CAP; ++i) {
    a[i] = v[i];
    const float t = tau_scal * td[i];
    v[i] = invert ? v[i] * t : v[i] / t;
  }
  leaf_apply<CAP, KIND, FN>(p, tx, v, tau_scal, td[0], !invert, pre);
#pragma unroll
  for (int i = 0; i < CAP; ++i) {
    if (invert) v[i] = a[i] - v[i] / (tau_scal * td[i]);
    else v[i] = a[i] - tau_scal * td[i] * v[i];
  }
}

// Generic pass kernel.  `Src` supplies the prox argument of every element and may observe the
// result (residual accumulation in the fused passes).  The source itself stays in the kernel
// parameter (constant) space; its mutable per-thread part is `Src::Regs`:
//   float  begin(r)                 scalar step for this pass, resets accumulators
//   float  load(r, e, i)            argument of global element e (component slot i)
//   void   post(r, e, i, result)    called after the prox
//   void   finish(r)                block-level epilogue (residual partials)
// KIND >= 0 fixes the prox kind at compile time (no run-time dispatch, no dead cases in the kernel).
template <int CAP, class Src, int KIND = -1>
__global__ void __launch_bounds__(kBlock) prox_pass_kernel(const ProxDesc p, const Src src,
                                                           float* __restrict__ out,
                                                           const ScaleRef tdiag, const bool invert) {
  typename Src::Regs r;
  const float tau = src.begin(r);
  for (uint32_t tx = blockIdx.x * blockDim.x + threadIdx.x; tx < p.count;
       tx += gridDim.x * blockDim.x) {
    float v[CAP], td[CAP];
#pragma unroll
    for (int i = 0; i < CAP; ++i) {
      v[i] = 0.f;
      td[i] = 1.f;
      if (i < (int)p.dim) {
        const uint32_t e = elem_index(p, tx, i);
        td[i] = tdiag.at(e);
        v[i] = src.load(r, e, i);
      }
    }
    group_apply<CAP, KIND>(p, tx, v, td, tau, invert);
#pragma unroll
    for (int i = 0; i < CAP; ++i) {
      if (i < (int)p.dim) {
        const uint32_t e = elem_index(p, tx, i);
        out[e] = v[i];
        src.post(r, e, i, v[i]);
      }
    }
  }
  src.finish(r);
}

// Same pass for planar groups of dimension 2 (the (x, y) pairs of ProxIndEpiQuad on the identity rows of the lifting
// config), four consecutive groups per thread: the two components of four neighbouring groups are 16 contiguous
// bytes each, so arguments and results move as 128-bit vectors (a quarter of the load / store instructions and of
// their address arithmetic; the pass is instruction bound, profiles/r02_lifting.md).  Same arithmetic per group.
// Requires uniform step-size diagonals, count % 4 == 0 and 16-byte aligned rows (checked by the launcher).
template <class Src, int KIND>
__global__ void __launch_bounds__(kBlock) prox_pass_pairs4_kernel(const ProxDesc p, const Src src,
                                                                  float* __restrict__ out, const float tdiag_val,
                                                                  const bool invert) {
  typename Src::Regs r;
  const float tau = src.begin(r);
  const uint32_t quads = p.count >> 2;
  for (uint32_t q4 = blockIdx.x * blockDim.x + threadIdx.x; q4 < quads; q4 += gridDim.x * blockDim.x) {
    const uint32_t tx = 4 * q4;
    const uint32_t e0 = p.index + tx, e1 = e0 + p.count;
    float a0[4], a1[4];
    src.load4(r, e0, a0);
    src.load4(r, e1, a1);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float v[2] = {a0[q], a1[q]}, td[2] = {tdiag_val, tdiag_val};
      group_apply<2, KIND>(p, tx + q, v, td, tau, invert);
      a0[q] = v[0];
      a1[q] = v[1];
    }
    *reinterpret_cast<float4*>(out + e0) = make_float4(a0[0], a0[1], a0[2], a0[3]);
    *reinterpret_cast<float4*>(out + e1) = make_float4(a1[0], a1[1], a1[2], a1[3]);
  }
}

// argument read from memory: the unfused Prox::Eval path
struct MemSource {
  const float* __restrict__ arg;
  float tau_scal;
  struct Regs {};
  __device__ __forceinline__ float begin(Regs&) const { return tau_scal; }
  __device__ __forceinline__ float load(Regs&, uint32_t e, int) const { return arg[e]; }
  __device__ __forceinline__ void post(Regs&, uint32_t, int, float) const {}
  __device__ __forceinline__ void finish(Regs&) const {}
};

#endif  // __CUDACC__

// ---- host-side prox objects -------------------------------------------------------------------

class Prox {
 public:
  Prox(Context* ctx, size_t index, size_t size, bool diagsteps)
      : ctx_(ctx), index_(index), size_(size), diagsteps_(diagsteps) {}
  virtual ~Prox() {}

  size_t index() const { return index_; }
  size_t size() const { return size_; }
  size_t end() const { return index_ + size_ - 1; }
  bool diagsteps() const { return diagsteps_; }
  virtual int kind() const = 0;
  virtual size_t gpu_mem_amount() const { return 0; }
  // (index, count, stride) groups over which preconditioners are averaged (prox.cu:73-78,
  // prox_separable_sum.hpp:65-77)
  virtual void get_separable_structure(std::vector<std::tuple<size_t, size_t, size_t>>& sep) const {
    sep.emplace_back(index_, size_, 1);
  }

  // Common size of ALL groups of get_separable_structure(), or 0 when they differ / are unknown.
  // Lets Problem::initialize average a constant preconditioner without enumerating the groups.
  virtual size_t uniform_group_size() const { return size_; }

  // Prox::Eval (prox.cu:26-43): full-length device vectors; slices by index.
  void eval(float* d_result, const float* d_arg, const float* d_tau_diag, float tau, bool invert) {
    eval_local(d_result + index_, d_arg + index_, d_tau_diag + index_, tau, invert);
  }
  // pointers already offset to the prox' first element
  virtual void eval_local(float* d_res, const float* d_arg, const float* d_tau, float tau,
                          bool invert) = 0;

  // Descriptor with index relative to a vector that starts at element `base` of the global
  // one (0 for full vectors, index() for local pointers).  Returns false when the prox is not
  // a register-resident leaf (then only eval_local can evaluate it).
  virtual bool leaf_desc(ProxDesc& out, size_t base) const { (void)out; (void)base; return false; }

 protected:
  Context* ctx_;
  size_t index_, size_;
  bool diagsteps_;
};

std::shared_ptr<Prox> make_prox_elem(Context* ctx, int kind, size_t index, size_t count, size_t dim,
                                     bool interleaved, bool diagsteps, int function,
                                     const float* const coeffs[7], const size_t coeff_len[7]);
std::shared_ptr<Prox> make_prox_simplex(Context* ctx, size_t index, size_t count, size_t dim,
                                        bool interleaved, bool diagsteps);
std::shared_ptr<Prox> make_prox_epi_quad(Context* ctx, size_t index, size_t count, size_t dim,
                                         bool interleaved, bool diagsteps, const float* a, size_t na,
                                         const float* b, size_t nb, const float* c, size_t nc);
std::shared_ptr<Prox> make_prox_moreau(Context* ctx, std::shared_ptr<Prox> inner);
// ProxElemOperation<T, ElemOperationIndSum<T>> (elem_operation_ind_sum.hpp:38-58): sum-to-one projection per group
std::shared_ptr<Prox> make_prox_ind_sum(Context* ctx, size_t index, size_t count, size_t dim, bool interleaved,
                                        bool diagsteps);
// ProxIndSum(index,size,count,dim,inds,sum[,count2,dim2,inds2,sum2]) (prox_ind_sum.hpp:37-62): index-list groups;
// inds2 == nullptr: one list
std::shared_ptr<Prox> make_prox_ind_sum_indexed(Context* ctx, size_t index, size_t size, size_t count, size_t dim,
                                                const unsigned long long* inds, float total, size_t count2,
                                                size_t dim2, const unsigned long long* inds2, float total2);
// spectral element operations (pb_spectral.cu): singular values of N x 2 matrices, eigenvalues of symmetric
// 2 x 2 / 3 x 3 / n x n matrices (elem_operation_singular_nx2.hpp, elem_operation_eigen_{2x2,3x3,nxn}.hpp)
std::shared_ptr<Prox> make_prox_spectral(Context* ctx, int kind, size_t index, size_t count, size_t dim,
                                         bool interleaved, bool diagsteps, int function_1d, int function_2d,
                                         const float* const coeffs[7], const size_t coeff_len[7]);
// ProxIndRange (prox_ind_range.hpp:37-50): projection onto the range of a sparse m x n matrix A given as CSC, with the
// dense column-major AA = A^T A
std::shared_ptr<Prox> make_prox_ind_range(Context* ctx, size_t index, size_t size, bool diagsteps, int m, int n, int nnz,
                                          const float* val, const int32_t* ptr, const int32_t* ind, const float* aa);
// ProxIndEpiConjQuad1D (external to the reference tree, cmake/CustomSources.cmake.example:8-14; parity unpinned):
// projection of (x, y) pairs onto the epigraph of the conjugate of a u^2 + b u + c restricted to [alpha, beta]
std::shared_ptr<Prox> make_prox_ind_epi_conjquad_1d(Context* ctx, size_t index, size_t count, bool interleaved,
                                                    bool diagsteps, const float* const coeffs[5], const size_t len[5]);
// ProxIndHalfspace(index,count,dim,interleaved,diagsteps,a,b) (prox_ind_halfspace.hpp:41-52), ProxIndSOC(..., alpha)
// (prox_ind_soc.hpp:39-48); both address planar groups regardless of `interleaved`, like the reference kernels
std::shared_ptr<Prox> make_prox_ind_halfspace(Context* ctx, size_t index, size_t count, size_t dim, bool interleaved,
                                              bool diagsteps, const float* a, size_t na, const float* b, size_t nb);
std::shared_ptr<Prox> make_prox_ind_soc(Context* ctx, size_t index, size_t count, size_t dim, bool interleaved,
                                        bool diagsteps, float alpha);
// ProxTransform (prox_transform.hpp:38-44): a, b, c, d, e with one value or one value per element
std::shared_ptr<Prox> make_prox_transform(Context* ctx, std::shared_ptr<Prox> inner, const float* const coeffs[5],
                                          const size_t coeff_len[5]);
std::shared_ptr<Prox> make_prox_permute(Context* ctx, std::shared_ptr<Prox> inner, const int* perm,
                                        size_t n);
std::shared_ptr<Prox> make_prox_zero(Context* ctx, size_t index, size_t size);

// launches prox_pass_kernel<CAP, MemSource> for a leaf descriptor (used by pb_prox.cu)
void launch_leaf_unfused(Context* ctx, const ProxDesc& d, float* d_res, const float* d_arg,
                         const float* d_tau, float tau, bool invert);

}  // namespace pb
