// pb_prox.cu -- host-side prox objects and the unfused Prox::Eval path.
#include "pb_prox.cuh"
#include "pb_linop.cuh"

#include <algorithm>
#include <cstring>

namespace pb {

// ---- slow path for groups that do not fit in registers (dim > kMaxRegDim) ---------------------
// One thread per group, arguments re-read from memory (L1/L2 resident) instead of staged in a
// per-thread local array like the reference's T[1024] (elem_operation_ind_simplex.hpp:28,63).

__global__ void __launch_bounds__(kBlock) prox_large_kernel(ProxDesc p, float* __restrict__ res,
                                                            const float* __restrict__ arg,
                                                            const float* __restrict__ tdiag,
                                                            float tau_scal, bool invert) {
  const uint32_t tx = blockIdx.x * blockDim.x + threadIdx.x;
  if (tx >= p.count) return;
  const uint32_t dim = p.dim;
  const bool moreau = p.moreau != 0;
  const bool inv = moreau ? !invert : invert;
  // argument as seen by the leaf (Moreau pre-scaling applied on the fly)
  auto leaf_arg = [&](uint32_t i) -> float {
    const uint32_t e = elem_index(p, tx, i);
    const float a = arg[e];
    if (!moreau) return a;
    const float t = tau_scal * tdiag[e];
    return invert ? a * t : a / t;
  };
  auto store = [&](uint32_t i, float r) {
    const uint32_t e = elem_index(p, tx, i);
    if (moreau) {
      const float t = tau_scal * tdiag[e];
      r = invert ? arg[e] - r / t : arg[e] - t * r;
    }
    res[e] = r;
  };
  const float td0 = tdiag[elem_index(p, tx, 0)];

  if (p.kind == kProxNorm2) {
    float sq = 0.f;
    for (uint32_t i = 0; i < dim; ++i) { const float v = leaf_arg(i); sq += v * v; }
    if (sq > 0.f) {
      const float norm = sqrtf(sq);
      Coeffs7 c;
      load_coeffs(p.coeffs, tx, c);
      const float tau = effective_tau(tau_scal, td0, inv);
      const float r = scaled_fun_prox(p.fn, norm, tau, c);
      for (uint32_t i = 0; i < dim; ++i) store(i, r * leaf_arg(i) / norm);
    } else {
      for (uint32_t i = 0; i < dim; ++i) store(i, 0.f);
    }
  } else if (p.kind == kProxSimplex) {
    // Michelot's fixed point on the active set; same unique threshold as the sort-based scan.
    float sum = 0.f;
    for (uint32_t i = 0; i < dim; ++i) sum += leaf_arg(i);
    uint32_t cnt = dim;
    float t = static_cast<float>((static_cast<double>(sum) - 1.0) / static_cast<double>((float)cnt));
    for (uint32_t it = 0; it < dim; ++it) {
      float s2 = 0.f;
      uint32_t c2 = 0;
      for (uint32_t i = 0; i < dim; ++i) {
        const float v = leaf_arg(i);
        if (v > t) { s2 += v; ++c2; }
      }
      if (c2 == cnt || c2 == 0) break;
      cnt = c2;
      t = static_cast<float>((static_cast<double>(s2) - 1.0) / static_cast<double>((float)cnt));
    }
    for (uint32_t i = 0; i < dim; ++i) store(i, fmaxf(leaf_arg(i) - t, 0.f));
  } else if (p.kind == kProxEpiQuad) {
    const float a = p.epi_a ? p.epi_a[tx] : p.epi_a_val;
    const float c = p.epi_c ? p.epi_c[tx] : p.epi_c_val;
    float sqb = 0.f, sqx = 0.f;
    for (uint32_t i = 0; i + 1 < dim; ++i) {
      const float b = p.epi_b[tx + (size_t)p.count * i];
      const float xs = leaf_arg(i) + b / (2 * a);
      sqb += b * b;
      sqx += xs * xs;
    }
    const float shift = sqb / (4 * a);
    const float ys = leaf_arg(dim - 1) - c + shift;
    float vv;
    bool inside;
    project_epi_quad(sqx, ys, a, vv, inside);
    float y = ys;
    const float norm = sqrtf(sqx);
    const double scale = static_cast<double>(vv) / (2.0 * static_cast<double>(a));
    float sq_new = 0.f;
    for (uint32_t i = 0; i + 1 < dim; ++i) {
      const float b = p.epi_b[tx + (size_t)p.count * i];
      float xs = leaf_arg(i) + b / (2 * a);
      if (!inside) {
        xs = (norm > 0.f) ? static_cast<float>(scale * static_cast<double>(xs / norm)) : 0.f;
        sq_new += xs * xs;
      }
      store(i, xs - b / (2 * a));
    }
    if (!inside) y = a * sq_new;
    store(dim - 1, y + c - shift);
  } else {   // 1D ops and zero have dim == 1 and never come here
    for (uint32_t i = 0; i < dim; ++i) store(i, leaf_arg(i));
  }
}

// ---- unfused leaf launch ------------------------------------------------------------------------

template <int CAP>
static void launch_cap(Context* ctx, const ProxDesc& d, float* res, const float* arg, const float* tdiag,
                       float tau, bool invert) {
  MemSource src{arg, tau};
  ScaleRef sr{tdiag, 1.f};
  const unsigned grid = std::min<size_t>(grid_for(d.count), (size_t)ctx->num_sms * 16);
  prox_pass_kernel<CAP, MemSource><<<grid, kBlock, 0, ctx->stream>>>(d, src, res, sr, invert);
  PB_CHECK_LAUNCH();
  ctx->launches++;
}

void launch_leaf_unfused(Context* ctx, const ProxDesc& d, float* res, const float* arg,
                         const float* tdiag, float tau, bool invert) {
  if (d.count == 0) return;
  switch (dim_cap(d.dim, d.kind)) {
    case 1: launch_cap<1>(ctx, d, res, arg, tdiag, tau, invert); break;
    case 2: launch_cap<2>(ctx, d, res, arg, tdiag, tau, invert); break;
    case 4: launch_cap<4>(ctx, d, res, arg, tdiag, tau, invert); break;
    case 8: launch_cap<8>(ctx, d, res, arg, tdiag, tau, invert); break;
    case 16: launch_cap<16>(ctx, d, res, arg, tdiag, tau, invert); break;
    case 32: launch_cap<32>(ctx, d, res, arg, tdiag, tau, invert); break;
    case 64: launch_cap<64>(ctx, d, res, arg, tdiag, tau, invert); break;
    default: {
      prox_large_kernel<<<grid_for(d.count), kBlock, 0, ctx->stream>>>(d, res, arg, tdiag, tau, invert);
      PB_CHECK_LAUNCH();
      ctx->launches++;
    }
  }
}

// ---- small helper kernels ---------------------------------------------------------------------

__global__ void __launch_bounds__(kBlock) moreau_prescale_kernel(float* __restrict__ out,
                                                                 const float* __restrict__ arg,
                                                                 const float* __restrict__ td, size_t n,
                                                                 float tau, bool invert) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    const float t = tau * td[i];
    out[i] = invert ? arg[i] * t : arg[i] / t;
  }
}

__global__ void __launch_bounds__(kBlock) moreau_postscale_kernel(float* __restrict__ res,
                                                                  const float* __restrict__ arg,
                                                                  const float* __restrict__ td, size_t n,
                                                                  float tau, bool invert) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    if (invert) res[i] = arg[i] - res[i] / (tau * td[i]);
    else res[i] = arg[i] - tau * td[i] * res[i];
  }
}

// res[i] = arg[perm[i]]  /  res[perm[i]] = arg[i]   (prox_permute.cu:30-48)
__global__ void __launch_bounds__(kBlock) permute_kernel(float* __restrict__ res,
                                                         const float* __restrict__ arg,
                                                         const int* __restrict__ perm, size_t n,
                                                         bool inverse) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    if (inverse) res[perm[i]] = arg[i];
    else res[i] = arg[perm[i]];
  }
}

static unsigned stream_grid(Context* ctx, size_t n) {
  return std::min<size_t>(grid_for(n), (size_t)ctx->num_sms * 32);
}

// ---- leaf proxes --------------------------------------------------------------------------------

class ProxLeaf : public Prox {
 public:
  ProxLeaf(Context* ctx, int kind, size_t index, size_t count, size_t dim, bool interleaved,
           bool diagsteps)
      : Prox(ctx, index, count * dim, diagsteps), kind_(kind), count_(count), dim_(dim),
        interleaved_(interleaved) {
    if (index + count * dim >= (1ull << 31)) fail(PB_ERR_UNSUPPORTED, "prox range exceeds 2^31-1");
  }
  int kind() const override { return kind_; }
  size_t uniform_group_size() const override { return kind_ == kProxZero ? size_ : dim_; }
  void get_separable_structure(std::vector<std::tuple<size_t, size_t, size_t>>& sep) const override {
    if (kind_ == kProxZero) { sep.emplace_back(index_, size_, 1); return; }
    // ProxSeparableSum (prox_separable_sum.hpp:65-77)
    for (size_t i = 0; i < count_; ++i) {
      if (interleaved_) sep.emplace_back(index_ + i * dim_, dim_, 1);
      else sep.emplace_back(index_ + i, dim_, count_);
    }
  }
  bool leaf_desc(ProxDesc& d, size_t base) const override {
    d = ProxDesc();
    d.kind = kind_;
    d.index = (uint32_t)(index_ - base);
    d.count = (uint32_t)count_;
    d.dim = (uint32_t)dim_;
    d.interleaved = interleaved_ ? 1 : 0;
    fill_desc(d);
    return true;
  }
  void eval_local(float* res, const float* arg, const float* td, float tau, bool invert) override {
    ctx_->bind();
    ProxDesc d;
    leaf_desc(d, index_);
    launch_leaf_unfused(ctx_, d, res, arg, td, tau, invert);
  }

 protected:
  virtual void fill_desc(ProxDesc&) const {}
  int kind_;
  size_t count_, dim_;
  bool interleaved_;
};

// ProxElemOperation with 7 coefficients (1D / Norm2 families) or none (simplex)
class ProxElem : public ProxLeaf {
 public:
  ProxElem(Context* ctx, int kind, size_t index, size_t count, size_t dim, bool interleaved,
           bool diagsteps, int function, const float* const coeffs[7], const size_t coeff_len[7])
      : ProxLeaf(ctx, kind, index, count, kind == kProxElem1D ? 1 : dim, interleaved, diagsteps),
        fn_(function) {
    if (function < 0 || function >= PB_FUN_COUNT_) fail(PB_ERR_INVALID, "unknown Function1D id");
    for (int k = 0; k < 7; ++k) {
      const size_t len = coeffs ? coeff_len[k] : 0;
      if (len == 0) fail(PB_ERR_INVALID, "ProxElemOperation: empty coefficient array");
      if (len > 1) {
        // reference: coeffs_[i].size() > 1 => device vector indexed by the group id
        // (prox_elem_operation.inl:82-89,156-167)
        if (len < count) fail(PB_ERR_INVALID, "ProxElemOperation: coefficient array shorter than count");
        // straight from the caller's buffer (one DMA when it is pinned), no intermediate host copy
        d_coeffs_[k].resize(len);
        upload_from_host(ctx, d_coeffs_[k].data(), coeffs[k], len);
        val_[k] = 0.f;
      } else {
        val_[k] = coeffs[k][0];
      }
    }
  }
  size_t gpu_mem_amount() const override {
    size_t mem = 0;
    for (int k = 0; k < 7; ++k)
      if (d_coeffs_[k].size()) mem += count_ * sizeof(float);
    return mem;
  }

 protected:
  void fill_desc(ProxDesc& d) const override {
    d.fn = fn_;
    for (int k = 0; k < 7; ++k) {
      d.coeffs.ptr[k] = d_coeffs_[k].size() ? d_coeffs_[k].data() : nullptr;
      d.coeffs.val[k] = val_[k];
    }
  }

 private:
  int fn_;
  float val_[7];
  DeviceBuffer<float> d_coeffs_[7];
};

class ProxEpiQuad : public ProxLeaf {
 public:
  ProxEpiQuad(Context* ctx, size_t index, size_t count, size_t dim, bool interleaved, bool diagsteps,
              const float* a, size_t na, const float* b, size_t nb, const float* c, size_t nc)
      : ProxLeaf(ctx, kProxEpiQuad, index, count, dim, interleaved, diagsteps) {
    // checks and messages of ProxIndEpiQuad::Initialize (prox_ind_epi_quad.cu:135-151)
    if (dim < 2) fail(PB_ERR_INVALID, "Wrong input: ind_epi_quad needs dim >= 2!");
    if (na != count && na != 1)
      fail(PB_ERR_INVALID, "Wrong input: Coefficient a has to have dimension count or 1!");
    for (size_t i = 0; i < na; ++i)
      if (a[i] <= 0) fail(PB_ERR_INVALID, "Wrong input: Coefficient a must be greater 0!");
    if (nb != count * (dim - 1))
      fail(PB_ERR_INVALID, "Wrong input: Coefficient b has to have dimension count*(dim-1)!");
    if (nc != count && nc != 1)
      fail(PB_ERR_INVALID, "Wrong input: Coefficient c has to have dimension count or 1!");
    if (na != 1) d_a_.assign(std::vector<float>(a, a + na), ctx->stream); else a_val_ = a[0];
    if (nc != 1) d_c_.assign(std::vector<float>(c, c + nc), ctx->stream); else c_val_ = c[0];
    d_b_.assign(std::vector<float>(b, b + nb), ctx->stream);
  }
  size_t gpu_mem_amount() const override {
    return (d_a_.size() + d_b_.size() + d_c_.size()) * sizeof(float);
  }

 protected:
  void fill_desc(ProxDesc& d) const override {
    d.interleaved = 0;   // the kernel addresses planar regardless (prox_ind_epi_quad.cu:54-57)
    d.epi_a = d_a_.size() ? d_a_.data() : nullptr;
    d.epi_b = d_b_.data();
    d.epi_c = d_c_.size() ? d_c_.data() : nullptr;
    d.epi_a_val = a_val_;
    d.epi_c_val = c_val_;
  }

 private:
  DeviceBuffer<float> d_a_, d_b_, d_c_;
  float a_val_ = 1.f, c_val_ = 0.f;
};

// ---- wrappers -----------------------------------------------------------------------------------

class ProxMoreau : public Prox {
 public:
  ProxMoreau(Context* ctx, std::shared_ptr<Prox> inner)
      : Prox(ctx, inner->index(), inner->size(), inner->diagsteps()), inner_(std::move(inner)) {}
  int kind() const override { return kProxMoreau; }
  size_t gpu_mem_amount() const override {
    ProxDesc d;
    const bool folded = inner_->leaf_desc(d, 0) && !d.moreau;
    return (folded ? 0 : size_ * sizeof(float)) + inner_->gpu_mem_amount();
  }
  size_t uniform_group_size() const override { return inner_->uniform_group_size(); }
  void get_separable_structure(std::vector<std::tuple<size_t, size_t, size_t>>& sep) const override {
    inner_->get_separable_structure(sep);
  }
  bool leaf_desc(ProxDesc& d, size_t base) const override {
    if (!inner_->leaf_desc(d, base)) return false;
    if (d.moreau) return false;          // Moreau of Moreau: generic path
    d.moreau = 1;
    return true;
  }
  void eval_local(float* res, const float* arg, const float* td, float tau, bool invert) override {
    ctx_->bind();
    ProxDesc d;
    if (leaf_desc(d, index_)) {          // pre/post scaling folded into the leaf kernel
      launch_leaf_unfused(ctx_, d, res, arg, td, tau, invert);
      return;
    }
    // generic: the reference's three steps (prox_moreau.cu:98-134)
    if (scaled_.size() != size_) scaled_.resize(size_);
    if (size_ == 0) return;
    const unsigned grid = stream_grid(ctx_, size_);
    moreau_prescale_kernel<<<grid, kBlock, 0, ctx_->stream>>>(scaled_.data(), arg, td, size_, tau, invert);
    PB_CHECK_LAUNCH();
    inner_->eval_local(res, scaled_.data(), td, tau, !invert);
    moreau_postscale_kernel<<<grid, kBlock, 0, ctx_->stream>>>(res, arg, td, size_, tau, invert);
    PB_CHECK_LAUNCH();
    ctx_->launches += 2;
  }

 private:
  std::shared_ptr<Prox> inner_;
  DeviceBuffer<float> scaled_;
};

// ---- ElemOperationIndSum (elem_operation_ind_sum.hpp:38-58): projection of every group onto { sum_i x_i = 1 }.
// One thread per group, the group is read twice (second time from L1 / L2); tau is ignored like in the reference.
__global__ void __launch_bounds__(kBlock) ind_sum_kernel(float* __restrict__ res, const float* __restrict__ arg,
                                                         size_t count, size_t dim, bool interleaved) {
  for (size_t tx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; tx < count; tx += (size_t)gridDim.x * blockDim.x) {
    const size_t first = interleaved ? tx * dim : tx, stride = interleaved ? 1 : count;   // vector.hpp:42-48
    float tl = 0;
    for (size_t i = 0; i < dim; ++i) tl += arg[first + i * stride];
    tl = static_cast<float>((tl - 1.) / static_cast<float>(dim));      // `1.` promotes to double (:50)
    for (size_t i = 0; i < dim; ++i) res[first + i * stride] = arg[first + i * stride] - tl;
  }
}

class ProxIndSum : public Prox {
 public:
  ProxIndSum(Context* ctx, size_t index, size_t count, size_t dim, bool interleaved, bool diagsteps)
      : Prox(ctx, index, count * dim, diagsteps), count_(count), dim_(dim), interleaved_(interleaved) {
    if (dim == 0) fail(PB_ERR_INVALID, "elem_operation:ind_sum needs dim >= 1");
    if (index + count * dim >= (1ull << 31)) fail(PB_ERR_UNSUPPORTED, "prox range exceeds 2^31-1");
  }
  int kind() const override { return kProxIndSum; }
  size_t uniform_group_size() const override { return dim_; }
  void get_separable_structure(std::vector<std::tuple<size_t, size_t, size_t>>& sep) const override {
    for (size_t i = 0; i < count_; ++i) {                 // ProxSeparableSum (prox_separable_sum.hpp:65-77)
      if (interleaved_) sep.emplace_back(index_ + i * dim_, dim_, 1);
      else sep.emplace_back(index_ + i, dim_, count_);
    }
  }
  void eval_local(float* res, const float* arg, const float*, float, bool) override {
    ctx_->bind();
    if (count_ == 0) return;
    ind_sum_kernel<<<stream_grid(ctx_, count_), kBlock, 0, ctx_->stream>>>(res, arg, count_, dim_, interleaved_);
    PB_CHECK_LAUNCH();
    ctx_->launches++;
  }

 private:
  size_t count_, dim_;
  bool interleaved_;
};

// ---- ProxIndHalfspace (prox_ind_halfspace.cu:34-92): projection of every group onto { x | <a, x> <= b }; planar
// groups (element k of group tx at tx + count*k); `a` holds one normal per group (count*dim, planar) or one normal
// for all groups (dim); `b` one offset per group or one for all.  tau is ignored like in the reference.
__global__ void __launch_bounds__(kBlock) ind_halfspace_kernel(float* __restrict__ res, const float* __restrict__ arg,
                                                               size_t count, size_t dim, const float* __restrict__ a,
                                                               const float* __restrict__ b, bool a_per_group,
                                                               bool b_per_group) {
  for (size_t tx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; tx < count; tx += (size_t)gridDim.x * blockDim.x) {
    const float t = b_per_group ? b[tx] : b[0];
    const float* n = a_per_group ? a + tx : a;
    const size_t n_stride = a_per_group ? count : 1;
    float sq_norm = 0, iprod = 0;
    for (size_t k = 0; k < dim; ++k) {                                   // ProjectHalfspace :42-47
      const float nk = n[k * n_stride];
      sq_norm += nk * nk;
      iprod += nk * arg[tx + count * k];
    }
    for (size_t k = 0; k < dim; ++k)                                     // :49-51
      res[tx + count * k] = arg[tx + count * k] - (fmaxf(0.f, iprod - t) / sq_norm) * n[k * n_stride];
  }
}

// ---- ProxIndSOC (prox_ind_soc.cu:33-77): projection onto { (x, y) | |x|_2 <= y }; x = components 0..dim-2 (planar),
// y = component dim-1; alpha = 1 only, like the reference (ProxIndSOC::Initialize).
__global__ void __launch_bounds__(kBlock) ind_soc_kernel(float* __restrict__ res, const float* __restrict__ arg,
                                                         size_t count, size_t dim) {
  for (size_t tx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; tx < count; tx += (size_t)gridDim.x * blockDim.x) {
    const float y0 = arg[count * (dim - 1) + tx];
    float norm_x0 = 0;
    for (size_t k = 0; k + 1 < dim; ++k) norm_x0 += arg[tx + count * k] * arg[tx + count * k];
    norm_x0 = sqrtf(norm_x0);
    if (norm_x0 <= y0) {
      for (size_t k = 0; k + 1 < dim; ++k) res[tx + count * k] = arg[tx + count * k];
      res[count * (dim - 1) + tx] = y0;
    } else if (norm_x0 <= -y0) {
      for (size_t k = 0; k + 1 < dim; ++k) res[tx + count * k] = 0;
      res[count * (dim - 1) + tx] = 0;
    } else {
      const float fac = (y0 + norm_x0) / (2 * norm_x0);
      for (size_t k = 0; k + 1 < dim; ++k) res[tx + count * k] = fac * arg[tx + count * k];
      res[count * (dim - 1) + tx] = fac * norm_x0;
    }
  }
}

// common part of the planar-group projections that live outside the fused machinery
class ProxGroupProjection : public Prox {
 public:
  ProxGroupProjection(Context* ctx, size_t index, size_t count, size_t dim, bool interleaved, bool diagsteps)
      : Prox(ctx, index, count * dim, diagsteps), count_(count), dim_(dim), interleaved_(interleaved) {
    if (index + count * dim >= (1ull << 31)) fail(PB_ERR_UNSUPPORTED, "prox range exceeds 2^31-1");
  }
  size_t uniform_group_size() const override { return dim_; }
  void get_separable_structure(std::vector<std::tuple<size_t, size_t, size_t>>& sep) const override {
    for (size_t i = 0; i < count_; ++i) {                 // ProxSeparableSum (prox_separable_sum.hpp:65-77)
      if (interleaved_) sep.emplace_back(index_ + i * dim_, dim_, 1);
      else sep.emplace_back(index_ + i, dim_, count_);
    }
  }

 protected:
  size_t count_, dim_;
  bool interleaved_;
};

class ProxIndHalfspace : public ProxGroupProjection {
 public:
  ProxIndHalfspace(Context* ctx, size_t index, size_t count, size_t dim, bool interleaved, bool diagsteps,
                   const float* a, size_t na, const float* b, size_t nb)
      : ProxGroupProjection(ctx, index, count, dim, interleaved, diagsteps) {
    // checks and messages of ProxIndHalfspace::Initialize (prox_ind_halfspace.cu:132-137)
    if (!a || !b || (na != count * dim && na != dim))
      fail(PB_ERR_INVALID, "Wrong input: Coefficient a has to have dimension count*dim or dim!");
    if (nb != count && nb != 1) fail(PB_ERR_INVALID, "Wrong input: Coefficient b has to have dimension count or 1!");
    a_per_group_ = na == count * dim;          // (count == 1: both readings address the same elements)
    b_per_group_ = nb == count;
    d_a_.resize(na);
    d_b_.resize(nb);
    upload_from_host(ctx, d_a_.data(), a, na);
    upload_from_host(ctx, d_b_.data(), b, nb);
  }
  int kind() const override { return kProxIndHalfspace; }
  size_t gpu_mem_amount() const override { return (d_a_.size() + d_b_.size()) * sizeof(float); }
  void eval_local(float* res, const float* arg, const float*, float, bool) override {
    ctx_->bind();
    if (count_ == 0) return;
    ind_halfspace_kernel<<<stream_grid(ctx_, count_), kBlock, 0, ctx_->stream>>>(
        res, arg, count_, dim_, d_a_.data(), d_b_.data(), a_per_group_, b_per_group_);
    PB_CHECK_LAUNCH();
    ctx_->launches++;
  }

 private:
  DeviceBuffer<float> d_a_, d_b_;
  bool a_per_group_ = false, b_per_group_ = false;
};

class ProxIndSOC : public ProxGroupProjection {
 public:
  ProxIndSOC(Context* ctx, size_t index, size_t count, size_t dim, bool interleaved, bool diagsteps, float alpha)
      : ProxGroupProjection(ctx, index, count, dim, interleaved, diagsteps) {
    if (alpha != 1) fail(PB_ERR_INVALID, "ProxIndSOC: Only alpha = 1 implemented right now.");   // prox_ind_soc.cu:117-119
    if (dim < 1) fail(PB_ERR_INVALID, "ProxIndSOC needs dim >= 1");
  }
  int kind() const override { return kProxIndSOC; }
  void eval_local(float* res, const float* arg, const float*, float, bool) override {
    ctx_->bind();
    if (count_ == 0) return;
    ind_soc_kernel<<<stream_grid(ctx_, count_), kBlock, 0, ctx_->stream>>>(res, arg, count_, dim_);
    PB_CHECK_LAUNCH();
    ctx_->launches++;
  }
};

// ---- ProxIndSum (prox_ind_sum.cu:33-70, prox_ind_sum.hpp:37-62): projection, in the metric of the step sizes,
// of index-list groups onto { sum_i x[inds[g*dim + i]] = total }: x_j - tau_j (sum_arg - total) / sum_tau.  Elements
// that no list mentions are copied (the zero prox).  One thread per group like the reference (the members of a group
// are scattered, so there is nothing to coalesce but the index list itself, which is stored member-major here:
// entry i of group g at i*count + g).  Same float expressions in the same order as the reference kernel.
__global__ void __launch_bounds__(kBlock) ind_sum_indexed_kernel(float* __restrict__ res, const float* __restrict__ arg,
                                                                 const float* __restrict__ td,
                                                                 const uint32_t* __restrict__ inds, size_t count,
                                                                 size_t n_run, size_t dim, float total, float tau,
                                                                 bool invert) {
  for (size_t tx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; tx < n_run; tx += (size_t)gridDim.x * blockDim.x) {
    float sum_arg = 0, sum_tau = 0;
    for (size_t i = 0; i < dim; ++i) {
      const uint32_t j = inds[i * count + tx];
      float mytau = td[j] * tau;
      if (invert) mytau = static_cast<float>(1. / mytau);
      sum_arg += arg[j];
      sum_tau += mytau;
    }
    for (size_t i = 0; i < dim; ++i) {
      const uint32_t j = inds[i * count + tx];
      float mytau = td[j] * tau;
      if (invert) mytau = static_cast<float>(1. / mytau);
      res[j] = arg[j] - mytau * (sum_arg - total) / sum_tau;
    }
  }
}

class ProxIndSumIndexed : public Prox {
 public:
  struct List {
    size_t count = 0, dim = 0, n_run = 0;
    float total = 0;
    DeviceBuffer<uint32_t> inds;
  };
  // diagsteps is always true (prox_ind_sum.hpp:44: Prox<T>(index, size, true))
  ProxIndSumIndexed(Context* ctx, size_t index, size_t size, size_t count, size_t dim, const unsigned long long* inds,
                    float total, size_t count2, size_t dim2, const unsigned long long* inds2, float total2)
      : Prox(ctx, index, size, true) {
    if (index + size >= (1ull << 31)) fail(PB_ERR_UNSUPPORTED, "prox range exceeds 2^31-1");
    fill(lists_[0], count, dim, inds, total);
    two_ = inds2 != nullptr;
    if (two_) {
      fill(lists_[1], count2, dim2, inds2, total2);
      // the reference sizes the second launch with the FIRST list's group count (prox_ind_sum.cu:135):
      // groups of the second list beyond that many 256-thread blocks are never projected
      const size_t covered = (count + 255) / 256 * 256;
      lists_[1].n_run = std::min(count2, covered);
    }
  }
  int kind() const override { return kProxIndSumIndexed; }
  size_t uniform_group_size() const override { return size_; }
  size_t gpu_mem_amount() const override {
    return (lists_[0].inds.size() + lists_[1].inds.size()) * sizeof(size_t);       // prox_ind_sum.cu:88-90
  }
  void eval_local(float* res, const float* arg, const float* td, float tau, bool invert) override {
    ctx_->bind();
    if (size_ == 0) return;
    // zero prox on the other indices (prox_ind_sum.cu:113-115)
    if (res != arg) PB_CUDA(cudaMemcpyAsync(res, arg, size_ * sizeof(float), cudaMemcpyDeviceToDevice, ctx_->stream));
    for (int k = 0; k < (two_ ? 2 : 1); ++k) {
      const List& l = lists_[k];
      if (l.n_run == 0) continue;
      ind_sum_indexed_kernel<<<stream_grid(ctx_, l.n_run), kBlock, 0, ctx_->stream>>>(
          res, arg, td, l.inds.data(), l.count, l.n_run, l.dim, l.total, tau, invert);
      PB_CHECK_LAUNCH();
      ctx_->launches++;
    }
  }

 private:
  void fill(List& l, size_t count, size_t dim, const unsigned long long* inds, float total) {
    if (!inds && count * dim != 0) fail(PB_ERR_INVALID, "ProxIndSum: dimensions dont fit");   // prox_ind_sum.cu:74-75
    l.count = l.n_run = count;
    l.dim = dim;
    l.total = total;
    std::vector<uint32_t> t(count * dim);
    for (size_t g = 0; g < count; ++g)
      for (size_t i = 0; i < dim; ++i) {
        const unsigned long long j = inds[g * dim + i];
        if (j >= size_) fail(PB_ERR_INVALID, "ProxIndSum: index outside the prox range");
        t[i * count + g] = static_cast<uint32_t>(j);
      }
    l.inds.resize(t.size());
    if (!t.empty()) l.inds.upload(t.data(), t.size(), ctx_->stream);
    PB_CUDA(cudaStreamSynchronize(ctx_->stream));
  }
  List lists_[2];
  bool two_ = false;
};

// ---- ProxIndEpiConjQuad1D ("ProxEpiConjQuadr" of the sublabel-accurate lifting, Moellenhoff et al. CVPR 2016) ----------
// PARITY UNPINNED: the reference names this prox only in cmake/CustomSources.cmake.example:8-14 (un-vendored repository
// preciserelaxation/src/cvpr2016/prost/prox_ind_epi_conjquad_1d.cu); its source is not under /root/reference.  Built
// from the published definition and the in-tree pieces it rests on (helper.hpp:112-183 ProjectEpiQuad1d /
// ProjectEpiQuadGeneral1d, :185-215 ProjectHalfspace), validated against a double-precision brute-force projection
// (tests/test_oracle_closed_forms.py, tests/test_zz_gpu_next_rows.py).
//
// Per group (x, y): Euclidean projection onto epi(rho*) with rho(u) = a u^2 + b u + c on [alpha, beta], a >= 0:
//   rho*(x) = alpha x - rho(alpha)            x <= x1 = 2 a alpha + b      (ray of slope alpha)
//           = (x - b)^2 / (4a) - c            x1 <= x <= x2 = 2 a beta + b (parabola arc; a = 0: the vertex (b, -c))
//           = beta x - rho(beta)              x >= x2                      (ray of slope beta)
// rho* is C^1, so the three boundary pieces own the strips between the normals at the two junctions: the sign of
// the tangential coordinate (x0 - xj) + slope_j (y0 - yj) selects the piece.
__device__ __forceinline__ void project_epi_quad_1d(float x0, float y0, float alpha, float& x, float& y) {
  if (y0 >= alpha * (x0 * x0)) { x = x0; y = y0; return; }                  // helper.hpp:121-124
  const float a = static_cast<float>(2. * static_cast<double>(alpha) * static_cast<double>(fabsf(x0)));
  const float b = static_cast<float>(2. * (1. - 2. * static_cast<double>(alpha) * static_cast<double>(y0)) / 3.);
  float d, v;
  if (b < 0) {
    const float sq = powf(-b, 1.5f);
    d = (a - sq) * (a + sq);
  } else {
    d = a * a + b * b * b;
  }
  if (d >= 0) {
    const float c = powf(a + sqrtf(d), static_cast<float>(1. / 3.));
    v = c - b / c;
  } else {
    v = 2 * sqrtf(-b) * cosf(acosf(a / powf(-b, 1.5f)) / 3.f);
  }
  if (x0 > 0) x = static_cast<float>(static_cast<double>(v) / (2. * static_cast<double>(alpha)));
  else if (x0 < 0) x = static_cast<float>(-static_cast<double>(v) / (2. * static_cast<double>(alpha)));
  else x = 0;
  y = alpha * x * x;
}

__device__ __forceinline__ void project_epi_conjquad_1d(float x0, float y0, float a, float b, float c, float alpha,
                                                        float beta, float& x, float& y) {
  const float x1 = 2 * a * alpha + b, x2 = 2 * a * beta + b;               // junction abscissae
  const float r1 = (a * alpha + b) * alpha + c, r2 = (a * beta + b) * beta + c;   // rho(alpha), rho(beta)
  const float y1 = alpha * x1 - r1, y2 = beta * x2 - r2;                   // rho*(x1), rho*(x2)
  const float s1 = (x0 - x1) + alpha * (y0 - y1);                          // tangential coordinates at the junctions
  const float s2 = (x0 - x2) + beta * (y0 - y2);
  if (s1 <= 0) {                                                           // ray 1: halfspace alpha x - y <= rho(alpha)
    const float viol = fmaxf(0.f, alpha * x0 - y0 - r1) / (alpha * alpha + 1.f);
    x = x0 - viol * alpha;
    y = y0 + viol;
  } else if (s2 >= 0) {                                                    // ray 2
    const float viol = fmaxf(0.f, beta * x0 - y0 - r2) / (beta * beta + 1.f);
    x = x0 - viol * beta;
    y = y0 + viol;
  } else if (a > 0) {                                                      // parabola y >= p x^2 + q x + r
    const float p = 1.f / (4 * a), q = -b / (2 * a), r = b * b / (4 * a) - c;
    float tx, ty;                                                          // ProjectEpiQuadGeneral1d, helper.hpp:160-183
    project_epi_quad_1d(static_cast<float>(x0 + q / (2. * p)), static_cast<float>(y0 + q * q / (4. * p) - r), p, tx, ty);
    x = static_cast<float>(tx - q / (2. * p));
    y = static_cast<float>(ty - q * q / (4. * p) + r);
  } else {                                                                 // a = 0: normal cone of the vertex (b, -c)
    const bool inside = y0 >= fmaxf(alpha * (x0 - b), beta * (x0 - b)) - c;
    x = inside ? x0 : b;
    y = inside ? y0 : -c;
  }
}

struct ConjQuadCoeffs {
  const float* ptr[5];      // a, b, c, alpha, beta per group, or null
  float val[5];
  __device__ __forceinline__ float at(int k, size_t i) const { return ptr[k] ? ptr[k][i] : val[k]; }
};

__global__ void __launch_bounds__(kBlock) ind_epi_conjquad_1d_kernel(float* __restrict__ res,
                                                                     const float* __restrict__ arg, size_t count,
                                                                     bool interleaved, const ConjQuadCoeffs co) {
  for (size_t tx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; tx < count; tx += (size_t)gridDim.x * blockDim.x) {
    const size_t ix = interleaved ? 2 * tx : tx, iy = interleaved ? 2 * tx + 1 : tx + count;
    float x, y;
    project_epi_conjquad_1d(arg[ix], arg[iy], co.at(0, tx), co.at(1, tx), co.at(2, tx), co.at(3, tx), co.at(4, tx), x, y);
    res[ix] = x;
    res[iy] = y;
  }
}

class ProxIndEpiConjQuad1D : public ProxGroupProjection {
 public:
  ProxIndEpiConjQuad1D(Context* ctx, size_t index, size_t count, bool interleaved, bool diagsteps,
                       const float* const coeffs[5], const size_t len[5])
      : ProxGroupProjection(ctx, index, count, 2, interleaved, diagsteps) {
    static const char* names[5] = {"a", "b", "c", "alpha", "beta"};
    for (int k = 0; k < 5; ++k) {
      if (!coeffs[k] || (len[k] != 1 && len[k] != count))
        fail(PB_ERR_INVALID, std::string("ProxIndEpiConjQuad1D: coefficient ") + names[k] + " needs 1 or count entries");
      co_.ptr[k] = nullptr;
      co_.val[k] = coeffs[k][0];
      if (len[k] > 1) {
        d_[k].resize(len[k]);
        upload_from_host(ctx, d_[k].data(), coeffs[k], len[k]);
        co_.ptr[k] = d_[k].data();
      }
    }
    for (size_t i = 0; i < std::max(len[0], std::max(len[3], len[4])); ++i) {
      const float a = coeffs[0][len[0] > 1 ? i : 0], al = coeffs[3][len[3] > 1 ? i : 0], be = coeffs[4][len[4] > 1 ? i : 0];
      if (!(a >= 0.f)) fail(PB_ERR_INVALID, "ProxIndEpiConjQuad1D: a must be >= 0 (convex pieces)");
      if (!(al <= be)) fail(PB_ERR_INVALID, "ProxIndEpiConjQuad1D: needs alpha <= beta");
    }
  }
  int kind() const override { return kProxIndEpiConjQuad1D; }
  size_t gpu_mem_amount() const override {
    size_t n = 0;
    for (auto& d : d_) n += d.size();
    return n * sizeof(float);
  }
  void eval_local(float* res, const float* arg, const float*, float, bool) override {
    ctx_->bind();
    if (count_ == 0) return;
    ind_epi_conjquad_1d_kernel<<<stream_grid(ctx_, count_), kBlock, 0, ctx_->stream>>>(res, arg, count_, interleaved_, co_);
    PB_CHECK_LAUNCH();
    ctx_->launches++;
  }

 private:
  ConjQuadCoeffs co_;
  DeviceBuffer<float> d_[5];
};

// ---- ProxIndRange (prox_ind_range.cu:28-300, prox_ind_range.hpp:37-50): projection onto the range of a sparse
// m x n matrix A,  x = A (A^T A)^{-1} A^T x0,  with the dense AA = A^T A supplied by the caller.  tau is ignored.
// The reference factors AA with cusolverDn potrf at Initialize and runs csrmv, potrs, csrmv per evaluation.  Here
// (A^T A)^{-1} is formed ONCE on the host (Cholesky + triangular solves in double; AA is n x n with n in the
// hundreds) and an evaluation is three fully parallel kernels: CSR SpMV with A^T, dense GEMV, CSR SpMV with A --
// no sequential triangular solve on the device and no cuSOLVER / cuSPARSE dependency.
class ProxIndRange : public Prox {
 public:
  ProxIndRange(Context* ctx, size_t index, size_t size, bool diagsteps, int m, int n, int nnz, const float* val,
               const int32_t* ptr, const int32_t* ind, const float* aa)
      : Prox(ctx, index, size, diagsteps), m_(m), n_(n) {
    if (m < 0 || n < 0 || nnz < 0 || !ptr || (nnz > 0 && (!val || !ind)) || !aa)
      fail(PB_ERR_INVALID, "ProxIndRange: bad matrix arguments");
    if ((size_t)m != size) fail(PB_ERR_INVALID, "ProxIndRange: the number of rows of 'A' must equal the prox size");
    if (n > 4096) fail(PB_ERR_UNSUPPORTED, "ProxIndRange: more than 4096 columns (host-side factorisation)");
    // Cholesky AA = L L^T in double; same failure as an indefinite potrf
    const size_t N = (size_t)n;
    std::vector<double> L(N * N, 0.0);
    for (size_t j = 0; j < N; ++j) {
      double dsum = aa[j * N + j];
      for (size_t k = 0; k < j; ++k) dsum -= L[j * N + k] * L[j * N + k];
      if (!(dsum > 0.0)) fail(PB_ERR_INVALID, "ProxIndRange: matrix 'AA' is not positive definite");
      const double ljj = std::sqrt(dsum);
      L[j * N + j] = ljj;
      for (size_t i = j + 1; i < N; ++i) {
        double v = aa[j * N + i];                         // column-major, lower triangle: AA(i, j)
        for (size_t k = 0; k < j; ++k) v -= L[i * N + k] * L[j * N + k];
        L[i * N + j] = v / ljj;
      }
    }
    // inverse by solving L L^T X = I column by column
    std::vector<float> inv(N * N);
    std::vector<double> col(N);
    for (size_t c = 0; c < N; ++c) {
      for (size_t i = 0; i < N; ++i) {                    // forward: L y = e_c
        double v = i == c ? 1.0 : 0.0;
        for (size_t k = 0; k < i; ++k) v -= L[i * N + k] * col[k];
        col[i] = v / L[i * N + i];
      }
      for (size_t ii = N; ii-- > 0;) {                    // backward: L^T x = y
        double v = col[ii];
        for (size_t k = ii + 1; k < N; ++k) v -= L[k * N + ii] * col[k];
        col[ii] = v / L[ii * N + ii];
      }
      for (size_t i = 0; i < N; ++i) inv[c * N + i] = static_cast<float>(col[i]);      // column-major
    }
    a_ = make_block_sparse_csc(ctx, 0, 0, m, n, nnz, val, ptr, ind);
    inv_ = make_block_dense(ctx, 0, 0, N, N, inv.data());
    t1_.resize(std::max<size_t>(N, 1));
    t2_.resize(std::max<size_t>(N, 1));
  }
  int kind() const override { return kProxIndRange; }
  size_t gpu_mem_amount() const override { return a_->gpu_mem_amount() + inv_->gpu_mem_amount() + 2 * (size_t)n_ * sizeof(float); }
  void eval_local(float* res, const float* arg, const float*, float, bool) override {
    ctx_->bind();
    if (m_ == 0) return;
    cudaStream_t s = ctx_->stream;
    PB_CUDA(cudaMemsetAsync(t1_.data(), 0, (size_t)n_ * sizeof(float), s));
    PB_CUDA(cudaMemsetAsync(t2_.data(), 0, (size_t)n_ * sizeof(float), s));
    PB_CUDA(cudaMemsetAsync(res, 0, (size_t)m_ * sizeof(float), s));
    if (n_ == 0) return;
    a_->eval_adjoint_local_add(t1_.data(), arg);          // A^T x0
    inv_->eval_local_add(t2_.data(), t1_.data());         // (A^T A)^{-1} .
    a_->eval_local_add(res, t2_.data());                  // A .
  }

 private:
  int m_, n_;
  std::shared_ptr<Block> a_, inv_;
  DeviceBuffer<float> t1_, t2_;
};

// ---- ProxTransform (prox_transform.cu:27-226): prox of  c f(a x - b) + <d, x> + (e/2)|x|^2  through the prox of f.
// Same three element-wise steps around the inner prox as the reference, same float expressions.
struct TransformCoeffs {
  const float* ptr[5];      // a, b, c, d, e per element, or null
  float val[5];
  __device__ __forceinline__ float at(int k, size_t i) const { return ptr[k] ? ptr[k][i] : val[k]; }
};

__global__ void __launch_bounds__(kBlock) transform_prescale_kernel(float* __restrict__ scaled_arg,
                                                                    float* __restrict__ scaled_tau,
                                                                    const float* __restrict__ arg,
                                                                    const float* __restrict__ td,
                                                                    const TransformCoeffs co, size_t n, float tau,
                                                                    bool invert) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float tau2 = tau * td[i];
    if (invert) tau2 = 1 / tau2;
    const float a = co.at(0, i), b = co.at(1, i), c = co.at(2, i), d = co.at(3, i), e = co.at(4, i);
    scaled_arg[i] = (a * (arg[i] - tau2 * d)) / (1 + tau2 * e) - b;      // ProxTransformPrescaleArgument :49
    scaled_tau[i] = (a * a * c * tau2) / (1 + tau2 * e);                 // ProxTransformPrescaleStepSize :75
  }
}

__global__ void __launch_bounds__(kBlock) transform_postscale_kernel(float* __restrict__ res,
                                                                     const TransformCoeffs co, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    res[i] = (res[i] + co.at(1, i)) / co.at(0, i);                       // ProxTransformPostscale :94
}

class ProxTransform : public Prox {
 public:
  ProxTransform(Context* ctx, std::shared_ptr<Prox> inner, const float* const coeffs[5], const size_t len[5])
      : Prox(ctx, inner->index(), inner->size(), inner->diagsteps()), inner_(std::move(inner)) {
    static const char* names = "abcde";
    for (int k = 0; k < 5; ++k) {
      const size_t n = coeffs ? len[k] : 0;
      if (n == 0 || !coeffs[k]) fail(PB_ERR_INVALID, "ProxTransform: empty coefficient array");
      if (n > 1 && n < size_) {
        std::ostringstream ss;
        ss << "ProxTransform: coefficient '" << names[k] << "' has " << n << " elements, expected 1 or " << size_ << ".";
        fail(PB_ERR_INVALID, ss.str());
      }
      if (k == 0)
        for (size_t i = 0; i < n; ++i)
          if (coeffs[0][i] == 0.f)
            fail(PB_ERR_INVALID,
                 "ProxTransform: Vector 'a' isn't allowed to contain zero element. (Division by zero)");
      co_.ptr[k] = nullptr;
      co_.val[k] = coeffs[k][0];
      if (n > 1) {
        d_co_[k].resize(size_);
        upload_from_host(ctx, d_co_[k].data(), coeffs[k], size_);
        co_.ptr[k] = d_co_[k].data();
      }
    }
    scaled_arg_.resize(size_);
    scaled_tau_.resize(size_);
  }
  int kind() const override { return kProxTransform; }
  size_t gpu_mem_amount() const override {
    size_t mem = 2 * size_ * sizeof(float) + inner_->gpu_mem_amount();
    for (int k = 0; k < 5; ++k) mem += d_co_[k].size() * sizeof(float);
    return mem;
  }
  size_t uniform_group_size() const override { return inner_->uniform_group_size(); }
  void get_separable_structure(std::vector<std::tuple<size_t, size_t, size_t>>& sep) const override {
    inner_->get_separable_structure(sep);
  }
  void eval_local(float* res, const float* arg, const float* td, float tau, bool invert) override {
    ctx_->bind();
    if (size_ == 0) return;
    const unsigned grid = stream_grid(ctx_, size_);
    transform_prescale_kernel<<<grid, kBlock, 0, ctx_->stream>>>(scaled_arg_.data(), scaled_tau_.data(), arg, td, co_,
                                                                 size_, tau, invert);
    PB_CHECK_LAUNCH();
    // the inner prox sees the transformed step as its diagonal and tau = 1, never inverted (:201-210)
    inner_->eval_local(res, scaled_arg_.data(), scaled_tau_.data(), 1.f, false);
    transform_postscale_kernel<<<grid, kBlock, 0, ctx_->stream>>>(res, co_, size_);
    PB_CHECK_LAUNCH();
    ctx_->launches += 2;
  }

 private:
  std::shared_ptr<Prox> inner_;
  TransformCoeffs co_;
  DeviceBuffer<float> d_co_[5], scaled_arg_, scaled_tau_;
};

class ProxPermute : public Prox {
 public:
  ProxPermute(Context* ctx, std::shared_ptr<Prox> inner, const int* perm, size_t n)
      : Prox(ctx, inner->index(), inner->size(), inner->diagsteps()), inner_(std::move(inner)) {
    if (n != inner_->size()) {
      std::ostringstream ss;
      ss << "Permutation vector has wrong size (" << n << ") instead of " << inner_->size() << ".";
      fail(PB_ERR_INVALID, ss.str());
    }
    for (size_t i = 0; i < n; ++i)
      if (perm[i] < 0 || (size_t)perm[i] >= n) fail(PB_ERR_INVALID, "Permutation index out of range.");
    d_perm_.assign(std::vector<int>(perm, perm + n), ctx->stream);
    permuted_.resize(n);
  }
  int kind() const override { return kProxPermute; }
  size_t gpu_mem_amount() const override {
    return size_ * (sizeof(float) + sizeof(int)) + inner_->gpu_mem_amount();
  }
  size_t uniform_group_size() const override { return inner_->uniform_group_size(); }
  void get_separable_structure(std::vector<std::tuple<size_t, size_t, size_t>>& sep) const override {
    inner_->get_separable_structure(sep);
  }
  void eval_local(float* res, const float* arg, const float* td, float tau, bool invert) override {
    ctx_->bind();
    if (size_ == 0) return;
    const unsigned grid = stream_grid(ctx_, size_);
    // gather into res, inner prox res -> permuted_ with the UN-permuted tau (Appendix B #12),
    // scatter back (prox_permute.cu:101-145)
    permute_kernel<<<grid, kBlock, 0, ctx_->stream>>>(res, arg, d_perm_.data(), size_, false);
    PB_CHECK_LAUNCH();
    inner_->eval_local(permuted_.data(), res, td, tau, invert);
    permute_kernel<<<grid, kBlock, 0, ctx_->stream>>>(res, permuted_.data(), d_perm_.data(), size_, true);
    PB_CHECK_LAUNCH();
    ctx_->launches += 2;
  }

 private:
  std::shared_ptr<Prox> inner_;
  DeviceBuffer<int> d_perm_;
  DeviceBuffer<float> permuted_;
};

// ---- factories --------------------------------------------------------------------------------

std::shared_ptr<Prox> make_prox_elem(Context* ctx, int kind, size_t index, size_t count, size_t dim,
                                     bool interleaved, bool diagsteps, int function,
                                     const float* const coeffs[7], const size_t coeff_len[7]) {
  return std::make_shared<ProxElem>(ctx, kind, index, count, dim, interleaved, diagsteps, function,
                                    coeffs, coeff_len);
}
std::shared_ptr<Prox> make_prox_simplex(Context* ctx, size_t index, size_t count, size_t dim,
                                        bool interleaved, bool diagsteps) {
  return std::make_shared<ProxLeaf>(ctx, kProxSimplex, index, count, dim, interleaved, diagsteps);
}
std::shared_ptr<Prox> make_prox_epi_quad(Context* ctx, size_t index, size_t count, size_t dim,
                                         bool interleaved, bool diagsteps, const float* a, size_t na,
                                         const float* b, size_t nb, const float* c, size_t nc) {
  return std::make_shared<ProxEpiQuad>(ctx, index, count, dim, interleaved, diagsteps, a, na, b, nb, c, nc);
}
std::shared_ptr<Prox> make_prox_moreau(Context* ctx, std::shared_ptr<Prox> inner) {
  return std::make_shared<ProxMoreau>(ctx, std::move(inner));
}
std::shared_ptr<Prox> make_prox_ind_sum(Context* ctx, size_t index, size_t count, size_t dim, bool interleaved,
                                        bool diagsteps) {
  return std::make_shared<ProxIndSum>(ctx, index, count, dim, interleaved, diagsteps);
}
std::shared_ptr<Prox> make_prox_ind_sum_indexed(Context* ctx, size_t index, size_t size, size_t count, size_t dim,
                                                const unsigned long long* inds, float total, size_t count2,
                                                size_t dim2, const unsigned long long* inds2, float total2) {
  return std::make_shared<ProxIndSumIndexed>(ctx, index, size, count, dim, inds, total, count2, dim2, inds2, total2);
}

std::shared_ptr<Prox> make_prox_ind_epi_conjquad_1d(Context* ctx, size_t index, size_t count, bool interleaved,
                                                    bool diagsteps, const float* const coeffs[5], const size_t len[5]) {
  return std::make_shared<ProxIndEpiConjQuad1D>(ctx, index, count, interleaved, diagsteps, coeffs, len);
}

std::shared_ptr<Prox> make_prox_ind_range(Context* ctx, size_t index, size_t size, bool diagsteps, int m, int n, int nnz,
                                          const float* val, const int32_t* ptr, const int32_t* ind, const float* aa) {
  return std::make_shared<ProxIndRange>(ctx, index, size, diagsteps, m, n, nnz, val, ptr, ind, aa);
}

std::shared_ptr<Prox> make_prox_ind_halfspace(Context* ctx, size_t index, size_t count, size_t dim, bool interleaved,
                                              bool diagsteps, const float* a, size_t na, const float* b, size_t nb) {
  return std::make_shared<ProxIndHalfspace>(ctx, index, count, dim, interleaved, diagsteps, a, na, b, nb);
}
std::shared_ptr<Prox> make_prox_ind_soc(Context* ctx, size_t index, size_t count, size_t dim, bool interleaved,
                                        bool diagsteps, float alpha) {
  return std::make_shared<ProxIndSOC>(ctx, index, count, dim, interleaved, diagsteps, alpha);
}
std::shared_ptr<Prox> make_prox_transform(Context* ctx, std::shared_ptr<Prox> inner, const float* const coeffs[5],
                                          const size_t coeff_len[5]) {
  return std::make_shared<ProxTransform>(ctx, std::move(inner), coeffs, coeff_len);
}
std::shared_ptr<Prox> make_prox_permute(Context* ctx, std::shared_ptr<Prox> inner, const int* perm,
                                        size_t n) {
  return std::make_shared<ProxPermute>(ctx, std::move(inner), perm, n);
}
std::shared_ptr<Prox> make_prox_zero(Context* ctx, size_t index, size_t size) {
  // ProxZero(index,size): diagsteps = true (prox_zero.cu:27-30); modelled as a dim-1 identity leaf
  return std::make_shared<ProxLeaf>(ctx, kProxZero, index, size, 1, false, true);
}

}  // namespace pb
