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
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc, bool pred) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  const int bytes = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Cooperative staging (COOP, columns of ny % kStagedBlock == 0 pixels, slabs and refresh iterations included): a CTA owns kStagedBlock
// consecutive pixels of ONE image column, so every operand of a label is a contiguous, 16-byte aligned row of
// kStagedBlock floats.  Instead of one 4-byte cp.async per thread, operand and label (with its own 64-bit address
// arithmetic: ~40 of the ~100 instructions per pixel and label, profiles/r02_lifting.md), the CTA copies each row with
// kStagedBlock / 4 16-byte cp.async: thread (label slot tid / 16, 16-byte piece tid % 16) issues ONE copy per operand
// and chunk.  The shifted operands (down / up neighbour) are the same row read at +-1, plus one extra element.  Rows
// are [4 pad][kStagedBlock][4 pad] floats; three chunk buffers and one CTA barrier per chunk (buffer (c + 1) % 3 was
// last read for chunk c - 2, which every thread finished before the barrier of chunk c - 1).
constexpr int kCoopRow = 64 + 8;

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

// Labels are staged in chunks of LC: chunk c+1 is in flight (its own cp.async group, second buffer) while chunk c
// is turned into prox arguments that stay in registers, so a thread holds 2 * LC labels of operands in shared
// memory instead of all CAPL.  Staging everything at once (round 1: 768 B per thread at 32 labels) left 6-8
// resident warps per SM and 0.33 issued instructions per cycle and scheduler (profiles/r01_lifting.md); with
// LC = 4 the footprint is 192-256 B per thread and the register file, not shared memory, bounds the occupancy.
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int kStagedChunk = 4;

// ---- primal pass: x+ = proj_simplex( x - tau T K^T y ) over the L <= CAPL labels of a pixel -----------------
// operands per label: x, p1, p1-left, p2, p2-up (+ identity rows) (+ the same dual operands of y_prev when CHECK)
template <int CAPL, int LC, bool HAS_ID, bool CHECK, bool SLAB, bool COOP = false>
__global__ void __launch_bounds__(kStagedBlock, CHECK ? 2 : 8) grad_primal_simplex_staged_kernel(
    const GradGeom g, const ProxDesc p, const float* __restrict__ x, const float* __restrict__ y,
    const float* __restrict__ y_prev, const float Tval, const PdhgState* __restrict__ st, const int ktyprev_zero,
    double* __restrict__ partials, float* __restrict__ x_out) {
  extern __shared__ __align__(16) float staged_smem[];
  constexpr int NA = HAS_ID ? 6 : 5;                  // arrays of the current iterate (x + dual operands)
  constexpr int NT = NA + (CHECK ? NA - 1 : 0);       // + the dual operands of y_prev
  constexpr int NCH = (CAPL + LC - 1) / LC;
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
    static_assert(!COOP || (LC == 4 && kStagedBlock == 64), "cooperative staging: 4 label slots x 16 pieces = 64 threads");
    // COOP: three chunk buffers and one barrier per chunk; the refresh variant (CHECK) keeps two buffers and pays a
    // second barrier per chunk instead, because its keep[] columns already fill the shared memory
    constexpr int NBUF = (COOP && !CHECK) ? 3 : 2;
    constexpr int RS = COOP ? kCoopRow : B;             // floats per (buffer, array, label slot) row
    float* s = staged_smem + threadIdx.x + (COOP ? 4 : 0);
    // chunk buffer b, array a, label slot j of the chunk; (CHECK) keep[k][label]: xo, K^T y, K^T y_prev
    auto at = [&](int b, int a, int j) -> float* { return s + ((b * NT + a) * LC + j) * RS; };
    float* keep = staged_smem + NBUF * NT * LC * RS + threadIdx.x;
    // COOP: this thread copies 16-byte piece `ck` of the rows of label slot `jrow`
    const uint32_t jrow = threadIdx.x >> 4, ck = threadIdx.x & 15u;
    const uint32_t cta_pix = pix - threadIdx.x;        // first pixel of the CTA (same column)
    const bool cta_first = y0 == threadIdx.x, cta_last = y0 - threadIdx.x + B == g.ny;
    auto stage_dual = [&](const float* __restrict__ q, const float* __restrict__ q_halo, int b, int a0, int j,
                          int li, uint32_t idx) {
      const float* q1 = q;
      const float* q2 = q + g.plane;
      cp_async4(at(b, a0 + 0, j), q1 + idx, has1);
      cp_async4(at(b, a0 + 1, j),
                left_in ? q1 + idx - g.ny : (left_halo ? q_halo + y0 + li * g.ny : q1), left_in || left_halo);
      cp_async4(at(b, a0 + 2, j), q2 + idx, has3);
      cp_async4(at(b, a0 + 3, j), has4 ? q2 + idx - 1 : q2, has4);
      if (HAS_ID) cp_async4(at(b, a0 + 4, j), q + g.id_row + idx, true);
    };
    auto stage_chunk = [&](int c) {
      const int b = c % NBUF;
      if (COOP) {
        const int li = c * LC + (int)jrow;
        const bool ok = li < CAPL && li < nl;
        const uint32_t eo = (ok ? li : 0) * g.nxny + cta_pix + 4 * ck;
        float* const dst = staged_smem + ((b * NT) * LC + jrow) * RS + 4 + 4 * ck;   // array 0, this slot, this piece
        cp_async16(dst, x + eo, ok);
        if (has1) cp_async16(dst + 1 * LC * RS, y + eo, ok);
        if (left_in) cp_async16(dst + 2 * LC * RS, y + eo - g.ny, ok);
        else if (left_halo)         // column 0 of a slab: the left neighbour's last y.gx column, [label][ny]
          cp_async16(dst + 2 * LC * RS, g.halo.in_a + (ok ? li : 0) * g.ny + (y0 - threadIdx.x) + 4 * ck, ok);
        cp_async16(dst + 3 * LC * RS, y + g.plane + eo, ok);
        if (HAS_ID) cp_async16(dst + 5 * LC * RS, y + g.id_row + eo, ok);
        // p2 one pixel up of the CTA's first pixel (the previous CTA's last pixel, same column)
        if (ck == 0) cp_async4(dst + 3 * LC * RS - 1, y + g.plane + eo - 1, ok && !cta_first);
        if (prev) {                 // the same dual operands of y_prev (arrays NA ..)
          float* const dp = dst + NA * LC * RS;
          if (has1) cp_async16(dp, y_prev + eo, ok);
          if (left_in) cp_async16(dp + 1 * LC * RS, y_prev + eo - g.ny, ok);
          else if (left_halo)
            cp_async16(dp + 1 * LC * RS, g.halo.in_b + (ok ? li : 0) * g.ny + (y0 - threadIdx.x) + 4 * ck, ok);
          cp_async16(dp + 2 * LC * RS, y_prev + g.plane + eo, ok);
          if (HAS_ID) cp_async16(dp + 4 * LC * RS, y_prev + g.id_row + eo, ok);
          if (ck == 0) cp_async4(dp + 2 * LC * RS - 1, y_prev + g.plane + eo - 1, ok && !cta_first);
        }
        cp_async_commit();
        return;
      }
#pragma unroll
      for (int j = 0; j < LC; ++j) {
        const int li = c * LC + j;
        if (li < CAPL && li < nl) {
          const uint32_t idx = pix + li * g.nxny;
          cp_async4(at(b, 0, j), x + idx, true);
          stage_dual(y, g.halo.in_a, b, 1, j, li, idx);
          if (prev) stage_dual(y_prev, g.halo.in_b, b, NA, j, li, idx);
        }
      }
      cp_async_commit();
    };
    stage_chunk(0);
    float v[CAPL], tdl[CAPL];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      if (c + 1 < NCH) stage_chunk(c + 1);
      else cp_async_commit();                          // empty group: uniform accounting for wait_group 1
      cp_async_wait_group<1>();
      if (COOP) __syncthreads();
      const int b = c % NBUF;
#pragma unroll
      for (int j = 0; j < LC; ++j) {
        const int li = c * LC + j;
        if (li < CAPL) {
          v[li] = 0.f;
          tdl[li] = Tval;
          if (li < nl) {
            // COOP: operands that do not exist (last column / row, first column / row) read as the zeros the
            // per-thread copies stage for them
            const float k = COOP ? staged_adj<HAS_ID>(has1 ? *at(b, 1, j) : 0.f,
                                                      (left_in || left_halo) ? *at(b, 2, j) : 0.f,
                                                      has3 ? *at(b, 3, j) : 0.f, has4 ? at(b, 3, j)[-1] : 0.f,
                                                      HAS_ID ? *at(b, 5, j) : 0.f, g.id_factor)
                                 : staged_adj<HAS_ID>(*at(b, 1, j), *at(b, 2, j), *at(b, 3, j), *at(b, 4, j),
                                               HAS_ID ? *at(b, 5, j) : 0.f, g.id_factor);
            const float xo = *at(b, 0, j);
            v[li] = primal_prox_arg(xo, tau, Tval, k);
            if (CHECK) {
              float kp = 0.f;
              if (prev)
                kp = COOP ? staged_adj<HAS_ID>(has1 ? *at(b, NA + 0, j) : 0.f,
                                               (left_in || left_halo) ? *at(b, NA + 1, j) : 0.f,
                                               has3 ? *at(b, NA + 2, j) : 0.f, has4 ? at(b, NA + 2, j)[-1] : 0.f,
                                               HAS_ID ? *at(b, NA + 4, j) : 0.f, g.id_factor)
                          : staged_adj<HAS_ID>(*at(b, NA + 0, j), *at(b, NA + 1, j), *at(b, NA + 2, j), *at(b, NA + 3, j),
                                        HAS_ID ? *at(b, NA + 4, j) : 0.f, g.id_factor);
              keep[(0 * CAPL + li) * B] = xo;
              keep[(1 * CAPL + li) * B] = k;
              keep[(2 * CAPL + li) * B] = kp;
            }
          }
        }
      }
      if (COOP && NBUF == 2) __syncthreads();           // buffer b is free for chunk c + 2
    }
    group_apply<CAPL, kProxSimplex, -1>(p, pix, v, tdl, tau, false);
    const float sq = sqrtf(Tval);
#pragma unroll
    for (int li = 0; li < CAPL; ++li) {
      if (li < nl) {
        const uint32_t idx = pix + li * g.nxny;
        x_out[idx] = v[li];
        if (SLAB && edge) g.halo.out[y0 + li * g.ny] = v[li];       // new column 0 -> left neighbour
        if (CHECK) {
          // dual residual (backend_pdhg.cu:73-94), same expressions as grad_primal_body
          const float xo = keep[(0 * CAPL + li) * B], k = keep[(1 * CAPL + li) * B], kp = keep[(2 * CAPL + li) * B];
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

inline size_t primal_staged_smem(int capl, int lc, bool has_id, bool check, bool coop = false) {
  const int na = has_id ? 6 : 5;
  const int nt = na + (check ? na - 1 : 0);
  if (coop)
    return static_cast<size_t>((check ? 2 : 3) * nt * lc) * kCoopRow * sizeof(float) +
           static_cast<size_t>(check ? 3 * capl : 0) * kStagedBlock * sizeof(float);
  return static_cast<size_t>(2 * nt * lc + (check ? 3 * capl : 0)) * kStagedBlock * sizeof(float);
}

// ---- dual pass on the gradient rows: y+ = prox_Norm2( y + sigma S ((1+theta) K x+ - theta K x) ) -----------
// group = the 2L gradient components of a pixel (component c*L + l), scalar weights, uniform Sigma.
// operands per label: x+ (centre, right, down), x (centre, right, down), y.gx, y.gy
template <int CAPL, int LC, int FN, bool CHECK, bool SLAB, bool COOP = false>
__global__ void __launch_bounds__(kStagedBlock, CHECK ? 2 : 6) grad_dual_norm2_staged_kernel(
    const GradGeom g, const ProxDesc p, const float* __restrict__ y, const float* __restrict__ xn,
    const float* __restrict__ xo, const float Sval, const PdhgState* __restrict__ st, const int kxprev_zero,
    double* __restrict__ partials, float* __restrict__ y_out) {
  extern __shared__ __align__(16) float staged_smem[];
  constexpr int B = kStagedBlock;
  constexpr int CAP = 2 * CAPL;
  constexpr int NT = 8;
  constexpr int NCH = (CAPL + LC - 1) / LC;
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
    static_assert(!COOP || (LC == 4 && kStagedBlock == 64), "cooperative staging: 4 label slots x 16 pieces = 64 threads");
    constexpr int NBUF = (COOP && !CHECK) ? 3 : 2;      // see grad_primal_simplex_staged_kernel
    constexpr int RS = COOP ? kCoopRow : B;             // floats per (buffer, array, label slot) row
    float* s = staged_smem + threadIdx.x + (COOP ? 4 : 0);
    auto at = [&](int b, int a, int j) -> float* { return s + ((b * NT + a) * LC + j) * RS; };
    // (CHECK) keep[k][label]: y.gx, y.gy, extrapolated K x (x, y), K x+ (x, y)
    float* keep = staged_smem + NBUF * NT * LC * RS + threadIdx.x;
    // COOP: this thread copies 16-byte piece `ck` of the rows of label slot `jrow`
    const uint32_t jrow = threadIdx.x >> 4, ck = threadIdx.x & 15u;
    const uint32_t cta_pix = pix - threadIdx.x;        // first pixel of the CTA (same column)
    const bool cta_last = y0 - threadIdx.x + B == g.ny;
    auto stage_primal = [&](const float* __restrict__ u, const float* __restrict__ u_halo, int b, int a0, int j,
                            int li, uint32_t idx) {
      cp_async4(at(b, a0 + 0, j), u + idx, true);
      cp_async4(at(b, a0 + 1, j),
                right_in ? u + idx + g.ny : (right_halo ? u_halo + y0 + li * g.ny : u), has_r);
      cp_async4(at(b, a0 + 2, j), has_d ? u + idx + 1 : u, has_d);
    };
    auto stage_chunk = [&](int c) {
      const int b = c % NBUF;
      if (COOP) {
        const int li = c * LC + (int)jrow;
        const bool ok = li < CAPL && li < nl;
        const uint32_t eo = (ok ? li : 0) * g.nxny + cta_pix + 4 * ck;
        float* const dst = staged_smem + ((b * NT) * LC + jrow) * RS + 4 + 4 * ck;   // array 0, this slot, this piece
        // right neighbour column: inside the slab, or (last column of a slab) the right neighbour's column 0 of
        // x+ / x, [label][ny]
        const uint32_t ho = (ok ? li : 0) * g.ny + (y0 - threadIdx.x) + 4 * ck;
        cp_async16(dst, xn + eo, ok);
        if (right_in) cp_async16(dst + 1 * LC * RS, xn + eo + g.ny, ok);
        else if (right_halo) cp_async16(dst + 1 * LC * RS, g.halo.in_a + ho, ok);
        if (!kxprev_zero) {
          cp_async16(dst + 3 * LC * RS, xo + eo, ok);
          if (right_in) cp_async16(dst + 4 * LC * RS, xo + eo + g.ny, ok);
          else if (right_halo) cp_async16(dst + 4 * LC * RS, g.halo.in_b + ho, ok);
        }
        cp_async16(dst + 6 * LC * RS, y + eo, ok);
        cp_async16(dst + 7 * LC * RS, y + g.plane + eo, ok);
        // x one pixel below the CTA's last pixel (the next CTA's first pixel, same column)
        if (ck == 15) {
          cp_async4(dst + 4, xn + eo + 4, ok && !cta_last);
          if (!kxprev_zero) cp_async4(dst + 3 * LC * RS + 4, xo + eo + 4, ok && !cta_last);
        }
        cp_async_commit();
        return;
      }
#pragma unroll
      for (int j = 0; j < LC; ++j) {
        const int li = c * LC + j;
        if (li < CAPL && li < nl) {
          const uint32_t idx = pix + li * g.nxny;
          stage_primal(xn, g.halo.in_a, b, 0, j, li, idx);
          if (!kxprev_zero) stage_primal(xo, g.halo.in_b, b, 3, j, li, idx);
          cp_async4(at(b, 6, j), y + idx, true);
          cp_async4(at(b, 7, j), y + g.plane + idx, true);
        }
      }
      cp_async_commit();
    };
    // grad_fwd<1, false, .>: gx = right - centre (0 on the last column), gy = down - centre (0 on the last row)
    auto k_of = [&](int b, int a0, int j, float& kx, float& ky) {
      const float c = *at(b, a0 + 0, j);
      kx = has_r ? *at(b, a0 + 1, j) - c : 0.f;
      ky = has_d ? (COOP ? at(b, a0 + 0, j)[1] : *at(b, a0 + 2, j)) - c : 0.f;
    };
    stage_chunk(0);
    float arg[CAP][1];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      if (c + 1 < NCH) stage_chunk(c + 1);
      else cp_async_commit();
      cp_async_wait_group<1>();
      if (COOP) __syncthreads();
      const int b = c % NBUF;
#pragma unroll
      for (int j = 0; j < LC; ++j) {
        const int li = c * LC + j;
        if (li < CAPL) {
          arg[li][0] = 0.f;                            // unused label slots do not change a 2-norm
          arg[CAPL + li][0] = 0.f;
          if (li < nl) {
            float k1x, k1y, k0x = 0.f, k0y = 0.f;
            k_of(b, 0, j, k1x, k1y);
            if (!kxprev_zero) k_of(b, 3, j, k0x, k0y);
            const float ex = dual_extrapolate(theta, k1x, k0x), ey = dual_extrapolate(theta, k1y, k0y);
            const float ygx = *at(b, 6, j), ygy = *at(b, 7, j);
            arg[li][0] = dual_prox_arg(ygx, sigma, Sval, ex);
            arg[CAPL + li][0] = dual_prox_arg(ygy, sigma, Sval, ey);
            if (CHECK) {
              keep[(0 * CAPL + li) * B] = ygx; keep[(1 * CAPL + li) * B] = ygy;
              keep[(2 * CAPL + li) * B] = ex;  keep[(3 * CAPL + li) * B] = ey;
              keep[(4 * CAPL + li) * B] = k1x; keep[(5 * CAPL + li) * B] = k1y;
            }
          }
        }
      }
      if (COOP && NBUF == 2) __syncthreads();           // buffer b is free for chunk c + 2
    }
    Coeffs7 c;
#pragma unroll
    for (int k = 0; k < 7; ++k) c.v[k] = p.coeffs.val[k];
    const int fn = FN >= 0 ? FN : p.fn;
    const float tau_eff = effective_tau(sigma, Sval, false);
    if (coeffs_simple(c)) norm2_lanes<1, CAP, true>(fn, arg, c, tau_eff);
    else norm2_lanes<1, CAP, false>(fn, arg, c, tau_eff);
    const float sq = sqrtf(Sval);
#pragma unroll
    for (int li = 0; li < CAPL; ++li) {
      if (li < nl) {
        const uint32_t idx = pix + li * g.nxny;
        y_out[idx] = arg[li][0];
        y_out[g.plane + idx] = arg[CAPL + li][0];
        if (SLAB && edge) g.halo.out[y0 + li * g.ny] = arg[li][0];   // last gx column -> right neighbour
        if (CHECK) {
          // primal residual (backend_pdhg.cu:97-120), same expressions as grad_dual_body
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const float yo = keep[((0 + cc) * CAPL + li) * B];
            const float ext = keep[((2 + cc) * CAPL + li) * B];
            const float k1 = keep[((4 + cc) * CAPL + li) * B];
            const float z_hat = (yo - arg[cc * CAPL + li][0]) / (sigma * sq) + sq * ext;
            const float diff = z_hat - sq * k1;
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

inline size_t dual_staged_smem(int capl, int lc, bool check, bool coop = false) {
  if (coop)
    return static_cast<size_t>((check ? 2 : 3) * 8 * lc) * kCoopRow * sizeof(float) +
           static_cast<size_t>(check ? 6 * capl : 0) * kStagedBlock * sizeof(float);
  return static_cast<size_t>(2 * 8 * lc + (check ? 6 * capl : 0)) * kStagedBlock * sizeof(float);
}

#endif  // __CUDACC__

}  // namespace pb
