// pb_math.cuh -- scalar prox arithmetic shared by the unfused and the fused kernels.
//
// Semantics follow the reference's device functors; where the reference promotes to double
// through `1.` / `2.` literals (Appendix B #11 of SURVEY.md) the same promotion is kept so
// that float results round identically:
//   Function1D family   include/prost/prox/elemop/function_1d.hpp:34-326
//   1D elem operation   include/prost/prox/elemop/elem_operation_1d.hpp:36-59
//   Norm2 elem op       include/prost/prox/elemop/elem_operation_norm2.hpp:39-88
//   epigraph projection include/prost/prox/helper.hpp:44-105
#pragma once

#include <cuda_runtime.h>
#include <math.h>

#include "prost_b200.h"

namespace pb {

// ---- PDHG prox arguments --------------------------------------------------------------------
// The contraction pattern is pinned to what nvcc emits for the reference's thrust functors
// (cuobjdump of backend_pdhg.cu: primal  FMUL t = tau*T; FFMA x - t*kty;  dual  FMUL u = theta*kx_prev;
// FFMA ext = (1+theta)*kx - u; FMUL s = sigma*S; FFMA y + s*ext), so that every execution mode
// rounds exactly like the reference instead of depending on per-kernel compiler choices.
__device__ __forceinline__ float primal_prox_arg(float x, float tau, float T, float kty) {
  return __fmaf_rn(-__fmul_rn(tau, T), kty, x);
}
__device__ __forceinline__ float dual_extrapolate(float theta, float kx, float kx_prev) {
  return __fmaf_rn(1 + theta, kx, -__fmul_rn(theta, kx_prev));
}
__device__ __forceinline__ float dual_prox_arg(float y, float sigma, float S, float ext) {
  return __fmaf_rn(__fmul_rn(sigma, S), ext, y);
}

// ---- Function1D: prox_{tau f}(x0) -------------------------------------------------------

__device__ __forceinline__ float f1d_abs(float x0, float tau) {
  if (x0 >= tau) return x0 - tau;
  if (x0 <= -tau) return x0 + tau;
  return 0.f;
}

__device__ __forceinline__ float f1d_square(float x0, float tau) {
  // x0 / (1. + tau): the literal makes this a DOUBLE division in the reference.  It is evaluated
  // as x0 * (1 / (1 + tau)) in double: the reciprocal is loop-invariant wherever tau is (uniform
  // lambda and preconditioner), and the float result differs from the true double quotient only
  // when that quotient lies within 2^-52 of a float rounding boundary (probability ~2^-28).
  return static_cast<float>(static_cast<double>(x0) * (1.0 / (1.0 + static_cast<double>(tau))));
}

__device__ __forceinline__ float f1d_l0(float x0, float tau) {
  return (x0 * x0 > 2 * tau) ? x0 : 0.f;
}

// Newton iteration for 0.5 (t-1)^2 + alpha t^q (function_1d.hpp:171-191)
__device__ inline float lq_newton(float t0, float alpha, float q, float eps) {
  float t = t0, delta;
  do {
    const float pw = powf(t, q);
    const float d1 = t - 1 + alpha * q * pw / t;
    const float d2 = 1 + alpha * q * (q - 1) * pw / (t * t);
    delta = d1 / d2;
    t = t - delta;
  } while (delta > eps);
  return t;
}

// closed form for q = 1/2 (function_1d.hpp:193-202)
__device__ inline float lq_half(float alpha) {
  const float sqrt3 = sqrtf(3.f);
  const float pi_half = 1.5707963267948966f;
  const float s = 2 * sinf((acosf(alpha * 3 * sqrt3 / 4) + pi_half) / 3) / sqrt3;
  return s * s;
}

__device__ inline float f1d_lq(float x0, float tau, float alpha) {
  if (alpha == 1) return f1d_abs(x0, tau);
  if (alpha == 0) return f1d_l0(x0, tau);
  const float eps = 1e-5f;                              // Function1DLq<float>::eps
  float t = 0;
  const float ax = fabsf(x0);
  if (ax > 0) {
    const float factor = tau * powf(ax, alpha - 2);
    if (alpha < 1) {
      const float t2 = 2 * (alpha - 1) / (alpha - 2);
      // 0.5 literal => double comparison in the reference
      const double bound = 0.5 * static_cast<double>(1 - (t2 - 1) * (t2 - 1)) /
                           static_cast<double>(powf(t2, alpha));
      if (static_cast<double>(factor) < bound)
        t = (alpha == 0.5f) ? lq_half(factor) : lq_newton(1.f, factor, alpha, eps);
    } else {
      t = lq_newton(1.f, factor, alpha, eps);
    }
  }
  return t * ax;
}

__device__ __forceinline__ float fun1d(int fn, float x0, float tau, float alpha, float beta) {
  switch (fn) {
    case PB_FUN_ZERO: return x0;
    case PB_FUN_ABS: return f1d_abs(x0, tau);
    case PB_FUN_SQUARE: return f1d_square(x0, tau);
    case PB_FUN_IND_LEQ0: return x0 > 0.f ? 0.f : x0;
    case PB_FUN_IND_GEQ0: return x0 < 0.f ? 0.f : x0;
    case PB_FUN_IND_EQ0: return 0.f;
    case PB_FUN_IND_BOX01: return x0 > 1.f ? 1.f : (x0 < 0.f ? 0.f : x0);
    case PB_FUN_MAX_POS0: return x0 > tau ? x0 - tau : (x0 < 0.f ? x0 : 0.f);
    case PB_FUN_L0: return f1d_l0(x0, tau);
    case PB_FUN_HUBER: {
      // (x0 / tau) / (1. + alpha / tau), then clamp to the unit ball (function_1d.hpp:157-169)
      float r = static_cast<float>(static_cast<double>(x0 / tau) /
                                   (1.0 + static_cast<double>(alpha / tau)));
      r /= fmaxf(1.f, fabsf(r));
      return x0 - tau * r;
    }
    case PB_FUN_LQ: return f1d_lq(x0, tau, alpha);
    case PB_FUN_LQ_PLUS_EPS: return 0.f;                // stub in the reference (:293-306)
    case PB_FUN_TRUNC_QUAD: {
      const float xs = f1d_square(x0, 2 * tau * alpha);
      const float en = alpha * xs * xs + (xs - x0) * (xs - x0) / (2 * tau);
      return en < beta ? xs : x0;
    }
    case PB_FUN_TRUNC_LINEAR: {
      const float xs = f1d_abs(x0, tau * alpha);
      const float en = (xs - x0) * (xs - x0) / (2 * tau) + alpha * fabsf(xs);
      return en < beta ? xs : x0;
    }
    default: return x0;
  }
}

// ---- c*f(ax - b) + dx + (e/2) x^2 plumbing common to the 1D and Norm2 operations ----------

struct Coeffs7 {           // a, b, c, d, e, alpha, beta of one element
  float v[7];
};

// effective step: tau_scal * tau_diag, or its reciprocal when invert_tau (double 1. literal)
__device__ __forceinline__ float effective_tau(float tau_scal, float tau_diag, bool invert) {
  const float t = tau_scal * tau_diag;
  return invert ? static_cast<float>(1.0 / static_cast<double>(t)) : t;
}

// Scaled argument / step of the generalised prox; r is the scalar the function acts on
// (the element itself for 1D, the group norm for Norm2).
__device__ __forceinline__ float scaled_fun_prox(int fn, float r, float tau, const Coeffs7& c) {
  const float a = c.v[0], b = c.v[1], cc = c.v[2], d = c.v[3], e = c.v[4];
  const float num_arg = a * (r - d * tau);
  const float num_step = cc * a * a * tau;
  float prox_arg, step;
  if (e == 0.f) {
    // denominator 1. + tau*0 == 1 exactly: the double ops round like their float versions
    prox_arg = __fsub_rn(num_arg, b);
    step = num_step;
  } else {
    const double rden = 1.0 / (1.0 + static_cast<double>(tau * e));     // shared by both quotients
    prox_arg = static_cast<float>(static_cast<double>(num_arg) * rden - static_cast<double>(b));
    step = static_cast<float>(static_cast<double>(num_step) * rden);
  }
  const float res = fun1d(fn, prox_arg, step, c.v[5], c.v[6]) + b;
  return a == 1.f ? res : res / a;        // res / 1 == res exactly
}

// Same as scaled_fun_prox for a == 1, d == 0, e == 0 (the default weights of the reference's
// function constructors): 1*(r - 0*tau) == r, c*1*1*tau == c*tau and x/1 == x hold exactly in
// IEEE arithmetic, so dropping those operations does not change a single bit.
__device__ __forceinline__ float scaled_fun_prox_simple(int fn, float r, float tau, float b, float cc,
                                                        float alpha, float beta) {
  return fun1d(fn, __fsub_rn(r, b), cc * tau, alpha, beta) + b;
}

__device__ __forceinline__ bool coeffs_simple(const Coeffs7& c) {
  return c.v[0] == 1.f && c.v[3] == 0.f && c.v[4] == 0.f;
}

// p / n for several p sharing one divisor: rn = RN(1/n) is computed once with an IEEE division,
// each quotient then costs one multiply and two FMAs (Markstein's correction step, which
// reproduces the correctly rounded quotient of the full division).
__device__ __forceinline__ float div_shared(float p, float n, float rn) {
  const float q0 = __fmul_rn(p, rn);
  const float rem = __fmaf_rn(-q0, n, p);
  return __fmaf_rn(rem, rn, q0);
}

// ElemOperation1D::operator() on one element
__device__ __forceinline__ float elem1d_apply(int fn, float arg, float tau_scal, float tau_diag,
                                              bool invert, const Coeffs7& c) {
  const float tau = effective_tau(tau_scal, tau_diag, invert);
  if (c.v[0] == 0.f || c.v[2] == 0.f) return (arg - tau * c.v[3]) / (1 + tau * c.v[4]);
  return scaled_fun_prox(fn, arg, tau, c);
}

// ---- projection onto the epigraph of y >= alpha ||x||^2 (helper.hpp:44-105) ----------------
// Returns the scale s such that x = s * x0 and writes y; `passthrough` is set when the
// point is already inside (then x = x0, y = y0 exactly).
//
// Same cubic, same branches as the reference (Cardano for a non-negative discriminant, the trigonometric form
// otherwise).  The reference evaluates it in float with powf(., 1.5), powf(., 1/3) and double-precision divisions;
// that is ~300 issue slots per point and made the identity-row pass of the lifting config compute bound
// (profiles/r02_lifting.md).  Here: t sqrtf(t) for t^1.5 and cbrtf for the cube root (both at least as accurate as
// powf: <= 1 ulp), the exact float product for 2 alpha |x| (a double product of two floats rounded to float IS the
// float product), one double multiply by 1/3 instead of the double division.  Differences against the reference's
// expressions are of the order of one float ulp of v (tests: 1e-5 on iterates, 2e-5 on the projection itself).
__device__ inline void project_epi_quad(float sq_norm_x0, float y0, float alpha, float& v_out,
                                        bool& inside) {
  inside = (y0 >= alpha * sq_norm_x0);
  v_out = 0.f;
  if (inside) return;
  const float norm_x0 = sqrtf(sq_norm_x0);
  const float a = __fmul_rn(2.f * alpha, norm_x0);
  const float b = static_cast<float>(
      (2.0 - 4.0 * static_cast<double>(alpha) * static_cast<double>(y0)) * (1.0 / 3.0));
  float d, v, sq = 0.f;
  if (b < 0) {
    sq = -b * sqrtf(-b);
    d = (a - sq) * (a + sq);
  } else {
    d = a * a + b * b * b;
  }
  if (d >= 0) {
    const float c = cbrtf(a + sqrtf(d));
    v = (fabsf(c) > 1e-6f) ? c - b / c : 0.f;
  } else {
    v = 2 * sqrtf(-b) * cosf(acosf(a / sq) * (1.f / 3.f));
  }
  v_out = v;
}

// x / d for a divisor that is used several times: when d is a power of two the quotient is the exact product with
// 1 / d (built from the exponent bits), otherwise the IEEE division.  Bit-identical to x / d either way.
struct SharedDivisor {
  float d, r;
  bool pow2;
  __device__ __forceinline__ explicit SharedDivisor(float dv) : d(dv) {
    const uint32_t bits = __float_as_uint(dv), e = (bits >> 23) & 0xffu;
    pow2 = (bits & 0x007fffffu) == 0u && e >= 2u && e <= 252u;
    r = __uint_as_float((bits & 0x80000000u) | ((254u - e) << 23));
  }
  __device__ __forceinline__ float operator()(float x) const { return pow2 ? x * r : x / d; }
  __device__ __forceinline__ double quotient(double x) const {
    return pow2 ? x * static_cast<double>(r) : x / static_cast<double>(d);
  }
};

}  // namespace pb
