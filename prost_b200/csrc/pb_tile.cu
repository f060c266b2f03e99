// pb_tile.cu -- one whole PDHG iteration in ONE pass over HBM for planar 2-D gradient operators.
//
// The two-pass schedule (pb_stencil.cuh) moves 44 B per pixel and iteration for ROF: the dual pass
// re-reads x+, x and y that the primal pass had in registers a moment earlier.  Here a CTA owns a
// TX x TY tile of the image:
//   phase A  x+ = prox_g(x - tau T K^T y) on the tile PLUS one halo column (x = cx+TX) and one
//            halo row (y = cy+TY): the forward differences of the dual step need x+ there.  x+ and
//            x of the extended tile go to shared memory; the tile's own x+ is written to HBM.
//   phase B  y+ = prox_f*(y + sigma S ((1+theta) K x+ - theta K x)) on the tile, with K x+ and K x
//            taken from shared memory.
// HBM traffic per pixel: read y (2), x, f; write x+, y+ (2) = 7 floats = 28 B instead of 44 B; the
// halo re-computation costs (TX+1)(TY+1)/(TX TY) - 1 = 3.9 % extra arithmetic and loads that hit L2
// (they are a neighbouring tile's primary loads).  The halo values of x+ are recomputed with exactly
// the same instructions as in the tile that owns them, so the result is bit-identical to the
// two-pass kernels, which in turn follow the reference operation by operation
// (backend_pdhg.cu:38-70, 311-381; block_gradient2d.cu:25-139).
//
// Used on iterations that neither refresh the residuals nor need the iteration-0 special cases
// (K^T y := 0, K x_prev := 0); those run the two-pass kernels on the same ping-pong buffers.
#include "pb_stencil.cuh"

namespace pb {

namespace {

constexpr int kTX = 32;             // tile columns
constexpr int kTY = 128;            // tile rows (y is the contiguous direction): one warp = one column
constexpr int kTileThreads = 256;   // 8 warps -> 4 column sweeps per tile
constexpr int kSRow = kTY + 4;      // shared row pitch in floats: tile + halo row, 16-byte aligned

// prox_g on VEC lanes: ProxElemOperation<ElemOperation1D<FN>> with scalar weights and an optional
// per-pixel b (same code path as grad_primal_body's scalar-weight branch)
template <int VEC, int FN>
__device__ __forceinline__ void elem1d_lanes(const ProxDesc& p, const Coeffs7& c0, const bool simple,
                                             const float tau, const float Tval, const uint32_t e,
                                             float (&arg)[VEC]) {
  float bv[VEC];
  if (p.coeffs.ptr[1]) {
    VecIO<VEC>::ld(p.coeffs.ptr[1] + e, bv);
  } else {
#pragma unroll
    for (int j = 0; j < VEC; ++j) bv[j] = c0.v[1];
  }
  const int fn = FN >= 0 ? FN : p.fn;
  if (simple) {
    const float tau_eff = effective_tau(tau, Tval, false);
#pragma unroll
    for (int j = 0; j < VEC; ++j)
      arg[j] = scaled_fun_prox_simple(fn, arg[j], tau_eff, bv[j], c0.v[2], c0.v[5], c0.v[6]);
  } else {
    Coeffs7 c = c0;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      c.v[1] = bv[j];
      arg[j] = elem1d_apply(fn, arg[j], tau, Tval, false, c);
    }
  }
}

// x+ at VEC consecutive y of column gx (label plane l); returns x (old) in xo and x+ in xn
template <int VEC, int FN>
__device__ __forceinline__ void primal_point(const GradGeom& g, const ProxDesc& pg, const Coeffs7& cg,
                                             const bool simple, const float* __restrict__ x,
                                             const float* __restrict__ y, const float tau, const float Tval,
                                             const uint32_t gx, const uint32_t gy, const uint32_t l,
                                             float (&xo)[VEC], float (&xn)[VEC]) {
  const uint32_t idx = gy + gx * g.ny + l * g.nxny;
  float k[VEC];
  VecIO<VEC>::ld(x + idx, xo);
  grad_adj<VEC, false, false, false>(g, y, nullptr, idx, gx, gy, l, k);
#pragma unroll
  for (int j = 0; j < VEC; ++j) xn[j] = primal_prox_arg(xo[j], tau, Tval, k[j]);
  elem1d_lanes<VEC, FN>(pg, cg, simple, tau, Tval, idx, xn);
}

template <int FN_G, int FN_F>
__global__ void __launch_bounds__(kTileThreads, 3) grad2d_iteration_tile_kernel(
    const GradGeom g, const ProxDesc pg, const ProxDesc pf, const float* __restrict__ x,
    const float* __restrict__ y, const float Tval, const float Sval, const PdhgState* __restrict__ st,
    const uint32_t tiles_y, float* __restrict__ x_out, float* __restrict__ y_out) {
  __shared__ __align__(16) float sxn[kTX + 1][kSRow];
  __shared__ __align__(16) float sxo[kTX + 1][kSRow];

  const float tau = st->tau, sigma = st->sigma, theta = st->theta;
  const uint32_t tile = blockIdx.x;
  const uint32_t tx = tile / tiles_y, ty = tile - tx * tiles_y;
  const uint32_t cx = tx * kTX, cy = ty * kTY;
  const uint32_t l = blockIdx.y;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t ly = lane * 4, gy = cy + ly;

  Coeffs7 cg, cf;
#pragma unroll
  for (int k = 0; k < 7; ++k) { cg.v[k] = pg.coeffs.val[k]; cf.v[k] = pf.coeffs.val[k]; }
  const bool g_simple = coeffs_simple(cg) && cg.v[2] != 0.f;

  // ---- phase A: x+ on the extended tile ---------------------------------------------------------
#pragma unroll
  for (int s = 0; s < kTX / 8; ++s) {
    const uint32_t col = warp + 8 * s, gx = cx + col;
    if (gx < g.nx && gy < g.ny) {
      float xo[4], xn[4];
      primal_point<4, FN_G>(g, pg, cg, g_simple, x, y, tau, Tval, gx, gy, l, xo, xn);
      VecIO<4>::st(x_out + gy + gx * g.ny + l * g.nxny, xn);
      VecIO<4>::st(&sxn[col][ly], xn);
      VecIO<4>::st(&sxo[col][ly], xo);
    }
  }
  if (warp == 0) {                       // halo column x = cx + TX (owned by the tile to the right)
    const uint32_t gx = cx + kTX;
    if (gx < g.nx && gy < g.ny) {
      float xo[4], xn[4];
      primal_point<4, FN_G>(g, pg, cg, g_simple, x, y, tau, Tval, gx, gy, l, xo, xn);
      VecIO<4>::st(&sxn[kTX][ly], xn);
      VecIO<4>::st(&sxo[kTX][ly], xo);
    }
  } else if (warp == 1) {                // halo row y = cy + TY (owned by the tile below)
    const uint32_t gx = cx + lane, hy = cy + kTY;
    if (gx < g.nx && hy < g.ny) {
      float xo[1], xn[1];
      primal_point<1, FN_G>(g, pg, cg, g_simple, x, y, tau, Tval, gx, hy, l, xo, xn);
      sxn[lane][kTY] = xn[0];
      sxo[lane][kTY] = xo[0];
    }
  }
  __syncthreads();

  // ---- phase B: y+ on the tile --------------------------------------------------------------------
  const int fn_f = FN_F >= 0 ? FN_F : pf.fn;
  const float tau_f = effective_tau(sigma, Sval, false);
  const bool f_simple = coeffs_simple(cf);
#pragma unroll
  for (int s = 0; s < kTX / 8; ++s) {
    const uint32_t col = warp + 8 * s, gx = cx + col;
    if (gx < g.nx && gy < g.ny) {
      const uint32_t idx = gy + gx * g.ny + l * g.nxny;
      float arg[2][4];
      float cn[4], co[4], rn[4], ro[4];
      VecIO<4>::ld(&sxn[col][ly], cn);
      VecIO<4>::ld(&sxo[col][ly], co);
      float k1x[4], k0x[4], k1y[4], k0y[4];
      if (gx < g.nx - 1) {               // grad_fwd: gx = u[idx+ny] - u[idx], 0 on the last column
        VecIO<4>::ld(&sxn[col + 1][ly], rn);
        VecIO<4>::ld(&sxo[col + 1][ly], ro);
#pragma unroll
        for (int j = 0; j < 4; ++j) { k1x[j] = rn[j] - cn[j]; k0x[j] = ro[j] - co[j]; }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) { k1x[j] = 0.f; k0x[j] = 0.f; }
      }
#pragma unroll
      for (int j = 0; j < 3; ++j) { k1y[j] = cn[j + 1] - cn[j]; k0y[j] = co[j + 1] - co[j]; }
      if (gy + 4 < g.ny) {               // gy = u[idx+1] - u[idx], 0 on the last row
        k1y[3] = sxn[col][ly + 4] - cn[3];
        k0y[3] = sxo[col][ly + 4] - co[3];
      } else {
        k1y[3] = 0.f;
        k0y[3] = 0.f;
      }
      float y1[4], y2[4];
      VecIO<4>::ld(y + idx, y1);
      VecIO<4>::ld(y + g.plane + idx, y2);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        arg[0][j] = dual_prox_arg(y1[j], sigma, Sval, dual_extrapolate(theta, k1x[j], k0x[j]));
        arg[1][j] = dual_prox_arg(y2[j], sigma, Sval, dual_extrapolate(theta, k1y[j], k0y[j]));
      }
      if (f_simple) norm2_lanes<4, 2, true>(fn_f, arg, cf, tau_f);
      else norm2_lanes<4, 2, false>(fn_f, arg, cf, tau_f);
      VecIO<4>::st(y_out + idx, arg[0]);
      VecIO<4>::st(y_out + g.plane + idx, arg[1]);
    }
  }
}

template <int FN_G>
void tile_launch_f(Context* ctx, dim3 grid, const GradGeom& g, const ProxDesc& pg, const ProxDesc& pf,
                   const float* x, const float* y, float Tval, float Sval, const PdhgState* st,
                   uint32_t tiles_y, float* x_out, float* y_out) {
  if (pf.fn == PB_FUN_IND_LEQ0)
    grad2d_iteration_tile_kernel<FN_G, PB_FUN_IND_LEQ0><<<grid, kTileThreads, 0, ctx->stream>>>(
        g, pg, pf, x, y, Tval, Sval, st, tiles_y, x_out, y_out);
  else
    grad2d_iteration_tile_kernel<FN_G, -1><<<grid, kTileThreads, 0, ctx->stream>>>(
        g, pg, pf, x, y, Tval, Sval, st, tiles_y, x_out, y_out);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

// Can the whole iteration run as one tiled pass?  K = one planar BlockGradient2D (no identity rows),
// prox_g = one Elem1D over all columns with scalar weights (b may be per pixel), prox_f* = one
// Norm2 over the (gx, gy) pair of every voxel with scalar weights, uniform T and Sigma.
bool tile_iteration_supported(const StencilPlan& plan, const std::vector<ProxDesc>& gd,
                              const std::vector<ProxDesc>& fd, ScaleRef T, ScaleRef S) {
  if (!plan.ok || plan.three_d) return false;
  const GradGeom& g = plan.geom;
  if (g.has_id || g.halo.has_left || g.halo.has_right) return false;
  if (g.ny % 4 != 0 || g.plane % 4 != 0 || g.L > 65535u) return false;
  if (T.ptr || S.ptr) return false;
  if (gd.size() != 1 || fd.size() != 1) return false;
  const ProxDesc& pg = gd[0];
  const ProxDesc& pf = fd[0];
  if (pg.kind != kProxElem1D || pg.moreau || pg.index != 0 || pg.dim != 1 || pg.count != g.plane) return false;
  for (int k = 0; k < 7; ++k)
    if (k != 1 && pg.coeffs.ptr[k]) return false;
  if (pg.coeffs.ptr[1] && !aligned16(pg.coeffs.ptr[1])) return false;
  if (pf.kind != kProxNorm2 || pf.moreau || pf.interleaved || pf.index != 0 || pf.dim != 2 ||
      pf.count != g.plane)
    return false;
  for (int k = 0; k < 7; ++k)
    if (pf.coeffs.ptr[k]) return false;
  return true;
}

void tile_iteration_launch(Context* ctx, const StencilPlan& plan, const ProxDesc& pg, const ProxDesc& pf,
                           const float* x, const float* y, ScaleRef T, ScaleRef S, const PdhgState* st,
                           float* x_out, float* y_out) {
  const GradGeom& g = plan.geom;
  const uint32_t tiles_x = (g.nx + kTX - 1) / kTX, tiles_y = (g.ny + kTY - 1) / kTY;
  const dim3 grid(tiles_x * tiles_y, g.L, 1);
  if (pg.fn == PB_FUN_SQUARE)
    tile_launch_f<PB_FUN_SQUARE>(ctx, grid, g, pg, pf, x, y, T.val, S.val, st, tiles_y, x_out, y_out);
  else if (pg.fn == PB_FUN_ABS)
    tile_launch_f<PB_FUN_ABS>(ctx, grid, g, pg, pf, x, y, T.val, S.val, st, tiles_y, x_out, y_out);
  else
    tile_launch_f<-1>(ctx, grid, g, pg, pf, x, y, T.val, S.val, st, tiles_y, x_out, y_out);
  PB_CHECK_LAUNCH();
  ctx->launches++;
}

}  // namespace pb
