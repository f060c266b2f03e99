// pb_stencil_staged.cuh -- per-pixel label groups (lifted multilabel energies): the two stencil passes with
// their operands staged through shared memory by cp.async.
//
// With L labels per pixel (planar layout idx = pix + l*nx*ny) one thread owns one pixel: the simplex projection
// needs all L labels of the pixel, the Norm2 ball all 2L gradient components.  The plain kernels
// (pb_stencil.cuh, CAPL > 1) issue the 5-8 global loads of every label into registers; at L = 32 that is 150-250
// registers per thread, 8-12 resident warps per SM and only a few loads in flight per thread -- ncu: 0.9-1.0 TB/s,
// long-scoreboard stalls 8-17 per issue (profiles/r01_lifting.md).  Here every thread first issues ALL its loads
// as 4-byte cp.async copies into its own column of a shared-memory array (operand a, label l, thread t at
// [(a*CAPL + l)*kStagedBlock + t]: conflict-free, and private to the thread, so no CTA barrier is needed, only
// cp.async.wait_all), i.e. 50-90 KB per CTA are in flight at once without occupying registers; then it computes
// from shared memory with exactly the arithmetic of grad_adj / grad_fwd / grad_primal_body / grad_dual_body
// (results are bit-identical; boundary rules become zero-filled copies: x - 0 == x exactly).
//
// Slab decomposition: same halo protocol as the plain kernels (halo_wait before the edge column's copies,
// edge column stored to the neighbour, halo_signal at the end).
#pragma once

#include "pb_stencil.cuh"

namespace pb {

constexpr int kStagedBlock = 64;

#ifdef __CUDACC__

__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc, bool pred) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  const int bytes = pred ? 4 : 0;                     // 0: nothing is read, the destination is zero-filled
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// (K^T p)(pix, l) from the five staged values (grad_adj<1, false, HAS_ID, .>):
//   a1 = p1[idx] (0 on the last column), a2 = p1[idx-ny] (0 / halo on column 0), a3 = p2[idx] (0 on the last
//   row), a4 = p2[idx-1] (0 on row 0), a5 = identity rows
template <bool HAS_ID>
__device__ __forceinline__ float staged_adj(float a1, float a2, float a3, float a4, float a5, float id_factor) {
  const float divx = a1 - a2;
  const float divy = a3 - a4;
  float out = -(divx + divy);
  if (HAS_ID) out = __fadd_rn(out, __fmul_rn(a5, id_factor));
  return out;
}

// ---- primal pass: x+ = proj_simplex( x - tau T K^T y ) over the L <= CAPL labels of a pixel -----------------
// operands per label: x, p1, p1-left, p2, p2-up (+ identity rows) (+ the same five of y_prev when CHECK)
template <int CAPL, bool HAS_ID, bool CHECK, bool SLAB>
__global__ void __launch_bounds__(kStagedBlock) grad_primal_simplex_staged_kernel(
    const GradGeom g, const ProxDesc p, const float* __restrict__ x, const float* __restrict__ y,
    const float* __restrict__ y_prev, const float Tval, const PdhgState* __restrict__ st, const int ktyprev_zero,
    double* __restrict__ partials, float* __restrict__ x_out) {
  extern __shared__ __align__(16) float staged_smem[];
  constexpr int NA = HAS_ID ? 6 : 5;                  // arrays of the current dual iterate (incl. x)
  constexpr int B = kStagedBlock;
  const float tau = st->tau;
  double acc0 = 0.0, acc1 = 0.0;
  const uint32_t total = g.q * g.nx;                  // VEC = 1: q = ny
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  bool edge = false;
  if (t < total) {
    uint32_t xx, y0;
    g.div_q.divmod(t, xx, y0);
    const uint32_t pix = y0 + xx * g.ny;
    const int nl = static_cast<int>(g.L);
    edge = SLAB && g.halo.has_left && xx == 0;
    if (SLAB && edge) halo_wait(g.halo);
    const bool has1 = xx < g.nx - 1 || (SLAB && g.halo.has_right);
    const bool left_in = xx > 0, left_halo = SLAB && !left_in && g.halo.has_left;
    const bool has3 = y0 + 1 != g.ny, has4 = y0 > 0;
    const bool prev = CHECK && !ktyprev_zero;
    float* s = staged_smem + threadIdx.x;
    auto stage_dual = [&](const float* __restrict__ q, const float* __restrict__ q_halo, int a0, int li,
                          uint32_t idx) {
      const float* q1 = q;
      const float* q2 = q + g.plane;
      cp_async4(s + ((a0 + 0) * CAPL + li) * B, q1 + idx, has1);
      cp_async4(s + ((a0 + 1) * CAPL + li) * B,
                left_in ? q1 + idx - g.ny : (left_halo ? q_halo + y0 + li * g.ny : q1), left_in || left_halo);
      cp_async4(s + ((a0 + 2) * CAPL + li) * B, q2 + idx, has3);
      cp_async4(s + ((a0 + 3) * CAPL + li) * B, has4 ? q2 + idx - 1 : q2, has4);
      if (HAS_ID) cp_async4(s + ((a0 + 4) * CAPL + li) * B, q + g.id_row + idx, true);
    };
#pragma unroll
    for (int li = 0; li < CAPL; ++li) {
      if (li < nl) {
        const uint32_t idx = pix + li * g.nxny;
        cp_async4(s + (0 * CAPL + li) * B, x + idx, true);
        stage_dual(y, g.halo.in_a, 1, li, idx);
        if (prev) stage_dual(y_prev, g.halo.in_b, NA, li, idx);
      }
    }
    cp_async_wait_all();

    float v[CAPL], tdl[CAPL];
#pragma unroll
    for (int li = 0; li < CAPL; ++li) {
      v[li] = 0.f;
      tdl[li] = Tval;
      if (li < nl) {
        const float k = staged_adj<HAS_ID>(s[(1 * CAPL + li) * B], s[(2 * CAPL + li) * B], s[(3 * CAPL + li) * B],
                                           s[(4 * CAPL + li) * B], HAS_ID ? s[(5 * CAPL + li) * B] : 0.f,
                                           g.id_factor);
        v[li] = primal_prox_arg(s[(0 * CAPL + li) * B], tau, Tval, k);
      }
    }
    group_apply<CAPL, kProxSimplex, -1>(p, pix, v, tdl, tau, false);
#pragma unroll
    for (int li = 0; li < CAPL; ++li) {
      if (li < nl) {
        const uint32_t idx = pix + li * g.nxny;
        x_out[idx] = v[li];
        if (SLAB && edge) g.halo.out[y0 + li * g.ny] = v[li];       // new column 0 -> left neighbour
        if (CHECK) {
          // dual residual (backend_pdhg.cu:73-94), same expressions as grad_primal_body
          const float xo = s[(0 * CAPL + li) * B];
          const float k = staged_adj<HAS_ID>(s[(1 * CAPL + li) * B], s[(2 * CAPL + li) * B],
                                             s[(3 * CAPL + li) * B], s[(4 * CAPL + li) * B],
                                             HAS_ID ? s[(5 * CAPL + li) * B] : 0.f, g.id_factor);
          float kp = 0.f;
          if (prev)
            kp = staged_adj<HAS_ID>(s[((NA + 0) * CAPL + li) * B], s[((NA + 1) * CAPL + li) * B],
                                    s[((NA + 2) * CAPL + li) * B], s[((NA + 3) * CAPL + li) * B],
                                    HAS_ID ? s[((NA + 4) * CAPL + li) * B] : 0.f, g.id_factor);
          const float sq = sqrtf(Tval);
          const float w_hat = (xo - v[li]) / (tau * sq) - sq * kp;
          const float diff = w_hat + sq * k;
          acc0 += static_cast<double>(diff * diff);
          acc1 += static_cast<double>(w_hat * w_hat);
        }
      }
    }
  }
  if (SLAB) halo_signal(g.halo, edge);
  if (CHECK) {
    block_sum2(acc0, acc1);
    if (threadIdx.x == 0) { partials[2 * blockIdx.x] = acc0; partials[2 * blockIdx.x + 1] = acc1; }
  }
}

inline size_t primal_staged_smem(int capl, bool has_id, bool check_prev) {
  const int na = has_id ? 6 : 5;
  return static_cast<size_t>(na + (check_prev ? na - 1 : 0)) * capl * kStagedBlock * sizeof(float);
}

// ---- dual pass on the gradient rows: y+ = prox_Norm2( y + sigma S ((1+theta) K x+ - theta K x) ) -----------
// group = the 2L gradient components of a pixel (component c*L + l), scalar weights, uniform Sigma.
// operands per label: x+ (centre, right, down), x (centre, right, down), y.gx, y.gy
template <int CAPL, int FN, bool CHECK, bool SLAB>
__global__ void __launch_bounds__(kStagedBlock) grad_dual_norm2_staged_kernel(
    const GradGeom g, const ProxDesc p, const float* __restrict__ y, const float* __restrict__ xn,
    const float* __restrict__ xo, const float Sval, const PdhgState* __restrict__ st, const int kxprev_zero,
    double* __restrict__ partials, float* __restrict__ y_out) {
  extern __shared__ __align__(16) float staged_smem[];
  constexpr int B = kStagedBlock;
  constexpr int CAP = 2 * CAPL;
  const float sigma = st->sigma, theta = st->theta;
  double acc0 = 0.0, acc1 = 0.0;
  const uint32_t total = g.q * g.nx;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  bool edge = false;
  if (t < total) {
    uint32_t xx, y0;
    g.div_q.divmod(t, xx, y0);
    if (SLAB) xx = g.nx - 1 - xx;                      // edge column first (see grad_dual_body)
    const uint32_t pix = y0 + xx * g.ny;
    const int nl = static_cast<int>(g.L);
    edge = SLAB && g.halo.has_right && xx == g.nx - 1;
    if (SLAB && edge) halo_wait(g.halo);
    const bool right_in = xx < g.nx - 1, right_halo = SLAB && !right_in && g.halo.has_right;
    const bool has_r = right_in || right_halo, has_d = y0 + 1 < g.ny;
    float* s = staged_smem + threadIdx.x;
    auto stage_primal = [&](const float* __restrict__ u, const float* __restrict__ u_halo, int a0, int li,
                            uint32_t idx) {
      cp_async4(s + ((a0 + 0) * CAPL + li) * B, u + idx, true);
      cp_async4(s + ((a0 + 1) * CAPL + li) * B,
                right_in ? u + idx + g.ny : (right_halo ? u_halo + y0 + li * g.ny : u), has_r);
      cp_async4(s + ((a0 + 2) * CAPL + li) * B, has_d ? u + idx + 1 : u, has_d);
    };
#pragma unroll
    for (int li = 0; li < CAPL; ++li) {
      if (li < nl) {
        const uint32_t idx = pix + li * g.nxny;
        stage_primal(xn, g.halo.in_a, 0, li, idx);
        if (!kxprev_zero) stage_primal(xo, g.halo.in_b, 3, li, idx);
        cp_async4(s + (6 * CAPL + li) * B, y + idx, true);
        cp_async4(s + (7 * CAPL + li) * B, y + g.plane + idx, true);
      }
    }
    cp_async_wait_all();

    // grad_fwd<1, false, .>: gx = right - centre (0 on the last column), gy = down - centre (0 on the last row)
    auto k_of = [&](int a0, int li, float& kx, float& ky) {
      const float c = s[((a0 + 0) * CAPL + li) * B];
      kx = has_r ? s[((a0 + 1) * CAPL + li) * B] - c : 0.f;
      ky = has_d ? s[((a0 + 2) * CAPL + li) * B] - c : 0.f;
    };
    float arg[CAP][1];
#pragma unroll
    for (int i = 0; i < CAP; ++i) arg[i][0] = 0.f;       // unused label slots do not change a 2-norm
#pragma unroll
    for (int li = 0; li < CAPL; ++li) {
      if (li < nl) {
        float k1x, k1y, k0x = 0.f, k0y = 0.f;
        k_of(0, li, k1x, k1y);
        if (!kxprev_zero) k_of(3, li, k0x, k0y);
        arg[li][0] = dual_prox_arg(s[(6 * CAPL + li) * B], sigma, Sval, dual_extrapolate(theta, k1x, k0x));
        arg[CAPL + li][0] = dual_prox_arg(s[(7 * CAPL + li) * B], sigma, Sval, dual_extrapolate(theta, k1y, k0y));
      }
    }
    Coeffs7 c;
#pragma unroll
    for (int k = 0; k < 7; ++k) c.v[k] = p.coeffs.val[k];
    const int fn = FN >= 0 ? FN : p.fn;
    const float tau_eff = effective_tau(sigma, Sval, false);
    if (coeffs_simple(c)) norm2_lanes<1, CAP, true>(fn, arg, c, tau_eff);
    else norm2_lanes<1, CAP, false>(fn, arg, c, tau_eff);
#pragma unroll
    for (int li = 0; li < CAPL; ++li) {
      if (li < nl) {
        const uint32_t idx = pix + li * g.nxny;
        y_out[idx] = arg[li][0];
        y_out[g.plane + idx] = arg[CAPL + li][0];
        if (SLAB && edge) g.halo.out[y0 + li * g.ny] = arg[li][0];   // last gx column -> right neighbour
        if (CHECK) {
          // primal residual (backend_pdhg.cu:97-120), same expressions as grad_dual_body
          float k1[2], k0[2] = {0.f, 0.f};
          k_of(0, li, k1[0], k1[1]);
          if (!kxprev_zero) k_of(3, li, k0[0], k0[1]);
          const float sq = sqrtf(Sval);
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const float yo = s[((6 + cc) * CAPL + li) * B];
            const float ext = dual_extrapolate(theta, k1[cc], k0[cc]);
            const float z_hat = (yo - arg[cc * CAPL + li][0]) / (sigma * sq) + sq * ext;
            const float diff = z_hat - sq * k1[cc];
            acc0 += static_cast<double>(diff * diff);
            acc1 += static_cast<double>(z_hat * z_hat);
          }
        }
      }
    }
  }
  if (SLAB) halo_signal(g.halo, edge);
  if (CHECK) {
    block_sum2(acc0, acc1);
    if (threadIdx.x == 0) { partials[2 * blockIdx.x] = acc0; partials[2 * blockIdx.x + 1] = acc1; }
  }
}

inline size_t dual_staged_smem(int capl) { return static_cast<size_t>(8) * capl * kStagedBlock * sizeof(float); }

#endif  // __CUDACC__

}  // namespace pb
