// pb_spectral.cu -- spectral element operations: proxes of functions of the singular
// values of N x 2 matrices and of the eigenvalues of symmetric 2 x 2, 3 x 3 and n x n matrices, one small matrix
// per group of a ProxSeparableSum.
//
// Reference: include/prost/prox/elemop/elem_operation_singular_nx2.hpp:32-150 (+ function_2d.hpp:28-101),
// elem_operation_eigen_2x2.hpp:28-146, elem_operation_eigen_3x3.hpp:31-377, elem_operation_eigen_nxn.hpp,
// elem_operation_mass_norm.hpp:17-186;
// mex names "elem_operation:singular_nx2:{sum_1d:<fun>, ind_l1_ball, moreau:ind_l1_ball}",
// "elem_operation:eigen_{2x2,3x3,nxn}:<fun>" (factory.cpp:49-102).
//
// What is kept from the reference: the interface (group layout, 7 coefficient arrays, tau_diag[0] of the group,
// invert_tau), double-precision internals, the scaled-prox plumbing around the Function1D / Function2D member
// (same float / double mix expression by expression), and for singular_nx2 the rank-deficient conventions
// (:116-147).  What is NOT kept: the eigen-solvers.  A prox of a spectral function is T = V f(Lambda) V^T, which
// does not depend on the choice of eigenvectors, so
//   * 2 x 2: Sylvester's formula  T = f(l2) I + (f(l1) - f(l2)) / (l1 - l2) (M - l2 I)  -- no eigenvectors, no
//     branches on the rotation (the reference carries LAPACK's dlaev2 logic, eigen_2x2.hpp:30-92);
//   * 3 x 3 and n x n (n <= 8; run-time loops up to the reference's N_MAX = 32): cyclic Jacobi rotations in double on registers / local memory, which are
//     unconditionally stable for (nearly) degenerate spectra (the reference uses Kopp's Cardano + cross-product
//     routine with its threshold cascades for 3 x 3, :31-300, and EISPACK tred2 / tql2 for n x n).
// Results agree with the reference to rounding (tests: numpy eigh / svd closed forms as in the reference's own
// test_prox_sum_eigen_*.m, and the live reference build on the GPU box).
#include <algorithm>

#include "pb_prox.cuh"

namespace pb {

namespace {

unsigned stream_grid(Context* ctx, size_t n) {
  return static_cast<unsigned>(std::min<size_t>(grid_for(n), (size_t)ctx->num_sms * 16));
}

enum { kFun2DSum1D = 0, kFun2DIndL1Ball = 1, kFun2DMoreauIndL1Ball = 2 };

struct SpectralDesc {
  int kind;            // pb_spectral_kind
  int fn;              // Function1D member
  int fn2d;            // Function2D member (singular_nx2 only)
  uint32_t count, dim, n;
  int interleaved;
  CoeffRef coeffs;
};

__device__ __forceinline__ size_t elem_at(const SpectralDesc& p, size_t tx, uint32_t i) {
  return p.interleaved ? tx * p.dim + i : tx + (size_t)p.count * i;      // vector.hpp:42-48
}

__device__ __forceinline__ void load7(const CoeffRef& c, size_t tx, float (&v)[7]) {
#pragma unroll
  for (int k = 0; k < 7; ++k) v[k] = c.ptr[k] ? c.ptr[k][tx] : c.val[k];
}

// step size of the group (elem_operation_eigen_2x2.hpp:110, singular_nx2.hpp:67)
__device__ __forceinline__ double group_tau(float tau_scal, float td0, bool invert) {
  return invert ? (1. / static_cast<double>(tau_scal * td0)) : static_cast<double>(tau_scal * td0);
}

// eigen_*: one eigenvalue through c f(a x - b) + d x + (e/2) x^2 (eigen_2x2.hpp:112-126: double argument and step,
// float Function1D, float back-substitution)
__device__ __forceinline__ double eig_prox(int fn, double lam, double tau, const float (&c)[7]) {
  if (c[0] == 0.f || c[2] == 0.f) return (lam - tau * c[3]) / (1 + tau * c[4]);
  const double p = ((c[0] * (lam - c[3] * tau)) / (1. + tau * c[4])) - c[1];
  const double step = (c[2] * c[0] * c[0] * tau) / (1. + tau * c[4]);
  return (fun1d(fn, static_cast<float>(p), static_cast<float>(step), c[5], c[6]) + c[1]) / c[0];
}

// Function2DIndL1Ball (function_2d.hpp:43-82): projection of (y1, y2) onto the l1 ball of radius alpha
__device__ __forceinline__ void ind_l1_ball(float y1, float y2, float& x1, float& x2, float alpha) {
  const float v1 = fabsf(y1), v2 = fabsf(y2);
  if (v1 + v2 <= alpha) { x1 = y1; x2 = y2; return; }
  const float mu1 = fmaxf(v1, v2), mu2 = fminf(v1, v2);
  const float l = static_cast<float>(0.5 * static_cast<double>(mu2 - mu1 + alpha));
  const int rho = (static_cast<double>(l) <= 0.) ? 1 : 2;
  const float theta = static_cast<float>((1. / rho) * static_cast<double>(mu1 + (rho == 2 ? mu2 : 0.f) - alpha));
  const float m1 = static_cast<float>(fmax(static_cast<double>(v1 - theta), 0.));
  const float m2 = static_cast<float>(fmax(static_cast<double>(v2 - theta), 0.));
  x1 = static_cast<float>((0.f < y1) - (y1 < 0.f)) * m1;
  x2 = static_cast<float>((0.f < y2) - (y2 < 0.f)) * m2;
}

__device__ __forceinline__ void fun2d(int fn2d, int fn, float y1, float y2, float& x1, float& x2, float tau, float alpha,
                                      float beta) {
  if (fn2d == kFun2DSum1D) {                         // Function2DSum1D :28-40
    x1 = fun1d(fn, y1, tau, alpha, beta);
    x2 = fun1d(fn, y2, tau, alpha, beta);
  } else if (fn2d == kFun2DIndL1Ball) {
    ind_l1_ball(y1, y2, x1, x2, alpha);
  } else {                                           // Function2DMoreau<IndL1Ball> :85-101
    float r1, r2;
    ind_l1_ball(y1 / tau, y2 / tau, r1, r2, alpha);
    x1 = y1 - tau * r1;
    x2 = y2 - tau * r2;
  }
}

// ---- singular values of an n x 2 matrix (columns arg[0..n), arg[n..2n)) ----------------------------------------
__global__ void __launch_bounds__(kBlock) spectral_singular_nx2_kernel(const SpectralDesc p, float* __restrict__ res,
                                                                       const float* __restrict__ arg,
                                                                       const float* __restrict__ td, float tau_scal,
                                                                       bool invert) {
  for (size_t tx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; tx < p.count; tx += (size_t)gridDim.x * blockDim.x) {
    const uint32_t n = p.dim / 2;
    float c[7];
    load7(p.coeffs, tx, c);
    double d11 = 0., d12 = 0., d22 = 0.;                       // D = A^T A (:47-53: float products, double sums)
    for (uint32_t i = 0; i < n; ++i) {
      const float a1 = arg[elem_at(p, tx, i)], a2 = arg[elem_at(p, tx, n + i)];
      d11 += a1 * a1;
      d12 += a1 * a2;
      d22 += a2 * a2;
    }
    const double trace = d11 + d22, det = d11 * d22 - d12 * d12;
    const double d = sqrt(fmax(0., 0.25 * trace * trace - det));
    const double lmax = fmax(0., 0.5 * trace + d), lmin = fmax(0., 0.5 * trace - d);
    const double smax = sqrt(lmax), smin = sqrt(lmin);
    const double tau = group_tau(tau_scal, td[elem_at(p, tx, 0)], invert);
    double s1, s2;
    if (c[0] == 0.f || c[2] == 0.f) {                          // :70-73
      s1 = (smax - tau * c[3]) / (1. + tau * c[4]);
      s2 = (smin - tau * c[3]) / (1. + tau * c[4]);
    } else {                                                   // :74-95
      const float y1 = static_cast<float>(((c[0] * (smax - c[3] * tau)) / (1. + tau * c[4])) - c[1]);
      const float y2 = static_cast<float>(((c[0] * (smin - c[3] * tau)) / (1. + tau * c[4])) - c[1]);
      const float step = static_cast<float>((c[2] * c[0] * c[0] * tau) / (1. + tau * c[4]));
      float x1, x2;
      fun2d(p.fn2d, p.fn, y1, y2, x1, x2, step, c[5], c[6]);
      s1 = (x1 + c[1]) / c[0];
      s2 = (x2 + c[1]) / c[0];
    }
    if (smax > 0) {
      // T = V Sigma^+ Sigma_p V^T for the symmetric D: with g1 = s1 / smax, g2 = s2 / smin (0 if smin = 0), Sylvester:
      //   T = g2 I + (g1 - g2) / (lmax - lmin) (D - lmin I)     (D = lmax: T = g1 I on its range)
      const double g1 = s1 / smax, g2 = (smin > 0.0) ? (s2 / smin) : 0.0;
      double t11, t12, t22;
      const double gap = lmax - lmin;
      if (gap > 1e-14 * lmax) {
        const double w = (g1 - g2) / gap;
        t11 = g2 + w * (d11 - lmin);
        t12 = w * d12;
        t22 = g2 + w * (d22 - lmin);
      } else {                                                 // equal singular values: any orthonormal V
        t11 = t22 = g1;
        t12 = 0.;
      }
      for (uint32_t i = 0; i < n; ++i) {
        const float a1 = arg[elem_at(p, tx, i)], a2 = arg[elem_at(p, tx, n + i)];
        res[elem_at(p, tx, i)] = static_cast<float>(a1 * t11 + a2 * t12);
        res[elem_at(p, tx, n + i)] = static_cast<float>(a1 * t12 + a2 * t22);
      }
    } else {                                                   // zero matrix (:141-147)
      for (uint32_t i = 0; i < 2 * n; ++i) res[elem_at(p, tx, i)] = 0.f;
      res[elem_at(p, tx, 0)] = static_cast<float>(s1);
      res[elem_at(p, tx, n + 1)] = static_cast<float>(s2);
    }
  }
}

// ---- eigenvalues of a symmetric 2 x 2 matrix [arg0, (arg1 + arg2)/2; ., arg3] -----------------------------------
__global__ void __launch_bounds__(kBlock) spectral_eigen_2x2_kernel(const SpectralDesc p, float* __restrict__ res,
                                                                    const float* __restrict__ arg,
                                                                    const float* __restrict__ td, float tau_scal,
                                                                    bool invert) {
  for (size_t tx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; tx < p.count; tx += (size_t)gridDim.x * blockDim.x) {
    float c[7];
    load7(p.coeffs, tx, c);
    const double A = arg[elem_at(p, tx, 0)], C = arg[elem_at(p, tx, 3)];
    const double B = (arg[elem_at(p, tx, 1)] + arg[elem_at(p, tx, 2)]) / 2;          // float sum like :107
    const double sm = A + C, df = A - C;
    const double rt = sqrt(df * df + 4.0 * B * B);
    const double l1 = 0.5 * (sm + rt), l2 = 0.5 * (sm - rt);                         // l1 >= l2
    const double tau = group_tau(tau_scal, td[elem_at(p, tx, 0)], invert);
    const double f1 = eig_prox(p.fn, l1, tau, c), f2 = eig_prox(p.fn, l2, tau, c);
    double t11, t12, t22;
    if (rt > 1e-300) {                               // Sylvester: T = f2 I + (f1 - f2) / (l1 - l2) (M - l2 I)
      const double w = (f1 - f2) / rt;
      t11 = f2 + w * (A - l2);
      t12 = w * B;
      t22 = f2 + w * (C - l2);
    } else {
      t11 = t22 = f1;
      t12 = 0.;
    }
    res[elem_at(p, tx, 0)] = static_cast<float>(t11);
    res[elem_at(p, tx, 1)] = static_cast<float>(t12);
    res[elem_at(p, tx, 2)] = static_cast<float>(t12);
    res[elem_at(p, tx, 3)] = static_cast<float>(t22);
  }
}

// ---- eigenvalues of a symmetric N x N matrix: cyclic Jacobi in double -----------------------------------------------
// A (upper triangle used) is diagonalised in place, V accumulates the rotations.  Converges quadratically; 3 x 3
// needs 4-5 sweeps, 8 x 8 about 8.
template <int N>
__device__ __forceinline__ void jacobi_eig(double (&A)[N][N], double (&V)[N][N], int n) {
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) V[i][j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 16; ++sweep) {
    double off = 0.0, diag = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      if (i < n) diag += A[i][i] * A[i][i];
#pragma unroll
      for (int j = i + 1; j < N; ++j)
        if (j < n) off += A[i][j] * A[i][j];
    }
    if (off <= 1e-32 * diag || off == 0.0) break;
#pragma unroll
    for (int pi = 0; pi < N - 1; ++pi) {
#pragma unroll
      for (int q = pi + 1; q < N; ++q) {
        if (q >= n) continue;
        const double apq = A[pi][q];
        if (apq == 0.0) continue;
        // rotation that annihilates A[p][q] (Rutishauser's stable form)
        const double theta = (A[q][q] - A[pi][pi]) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
        A[pi][pi] -= t * apq;
        A[q][q] += t * apq;
        A[pi][q] = 0.0;
#pragma unroll
        for (int r = 0; r < N; ++r) {
          if (r >= n) continue;
          if (r != pi && r != q) {
            // upper-triangle storage: element (min, max)
            double& arp = r < pi ? A[r][pi] : A[pi][r];
            double& arq = r < q ? A[r][q] : A[q][r];
            const double vp = arp, vq = arq;
            arp = cs * vp - sn * vq;
            arq = sn * vp + cs * vq;
          }
          const double vp = V[r][pi], vq = V[r][q];
          V[r][pi] = cs * vp - sn * vq;
          V[r][q] = sn * vp + cs * vq;
        }
      }
    }
  }
}

template <int N>
__global__ void __launch_bounds__(kBlock) spectral_eigen_nxn_kernel(const SpectralDesc p, float* __restrict__ res,
                                                                    const float* __restrict__ arg,
                                                                    const float* __restrict__ td, float tau_scal,
                                                                    bool invert) {
  for (size_t tx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; tx < p.count; tx += (size_t)gridDim.x * blockDim.x) {
    const int n = static_cast<int>(p.n);
    float c[7];
    load7(p.coeffs, tx, c);
    double A[N][N], V[N][N];
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j) {
        A[i][j] = 0.0;
        if (i < n && j < n && j >= i) {              // symmetrised input (eigen_3x3.hpp:311-325, eigen_nxn.hpp:283-290)
          const float u = arg[elem_at(p, tx, i * n + j)], l = arg[elem_at(p, tx, j * n + i)];
          A[i][j] = i == j ? static_cast<double>(u) : static_cast<double>(u + l) / 2.;
        }
      }
    jacobi_eig<N>(A, V, n);
    const double tau = group_tau(tau_scal, td[elem_at(p, tx, 0)], invert);
    double f[N];
#pragma unroll
    for (int k = 0; k < N; ++k) f[k] = k < n ? eig_prox(p.fn, A[k][k], tau, c) : 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = i; j < N; ++j) {
        if (j >= n) continue;
        double t = 0.0;                               // T = V f(Lambda) V^T
#pragma unroll
        for (int k = 0; k < N; ++k)
          if (k < n) t += V[i][k] * V[j][k] * f[k];
        res[elem_at(p, tx, i * n + j)] = static_cast<float>(t);
        res[elem_at(p, tx, j * n + i)] = static_cast<float>(t);
      }
  }
}

// 8 < n <= 32 (the reference's N_MAX, elem_operation_eigen_nxn.hpp:11): the same cyclic Jacobi iteration with run-time
// loops on thread-local arrays (full symmetric storage).  Like the reference's tred2 / tql2 on double V[32][32] this
// lives in local memory; such sizes are rare and far from any hot path.
constexpr int kSpectralMaxN = 32;
__global__ void __launch_bounds__(64) spectral_eigen_big_kernel(const SpectralDesc p, float* __restrict__ res,
                                                                const float* __restrict__ arg,
                                                                const float* __restrict__ td, float tau_scal,
                                                                bool invert) {
  for (size_t tx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; tx < p.count; tx += (size_t)gridDim.x * blockDim.x) {
    const int n = static_cast<int>(p.n);
    float c[7];
    load7(p.coeffs, tx, c);
    double A[kSpectralMaxN][kSpectralMaxN], V[kSpectralMaxN][kSpectralMaxN];
#pragma unroll 1
    for (int i = 0; i < n; ++i)
#pragma unroll 1
      for (int j = i; j < n; ++j) {
        const float u = arg[elem_at(p, tx, i * n + j)], l = arg[elem_at(p, tx, j * n + i)];
        A[i][j] = A[j][i] = i == j ? static_cast<double>(u) : static_cast<double>(u + l) / 2.;
        V[i][j] = V[j][i] = i == j ? 1.0 : 0.0;
      }
#pragma unroll 1
    for (int sweep = 0; sweep < 24; ++sweep) {
      double off = 0.0, diag = 0.0;
#pragma unroll 1
      for (int i = 0; i < n; ++i) {
        diag += A[i][i] * A[i][i];
#pragma unroll 1
        for (int j = i + 1; j < n; ++j) off += A[i][j] * A[i][j];
      }
      if (off <= 1e-32 * diag || off == 0.0) break;
#pragma unroll 1
      for (int pi = 0; pi < n - 1; ++pi)
#pragma unroll 1
        for (int q = pi + 1; q < n; ++q) {
          const double apq = A[pi][q];
          if (apq == 0.0) continue;
          const double theta = (A[q][q] - A[pi][pi]) / (2.0 * apq);
          const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
          const double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
          A[pi][pi] -= t * apq;
          A[q][q] += t * apq;
          A[pi][q] = A[q][pi] = 0.0;
#pragma unroll 1
          for (int r = 0; r < n; ++r) {
            if (r != pi && r != q) {
              const double vp = A[r][pi], vq = A[r][q];
              A[r][pi] = A[pi][r] = cs * vp - sn * vq;
              A[r][q] = A[q][r] = sn * vp + cs * vq;
            }
            const double vp = V[r][pi], vq = V[r][q];
            V[r][pi] = cs * vp - sn * vq;
            V[r][q] = sn * vp + cs * vq;
          }
        }
    }
    const double tau = group_tau(tau_scal, td[elem_at(p, tx, 0)], invert);
#pragma unroll 1
    for (int k = 0; k < n; ++k) A[k][k] = eig_prox(p.fn, A[k][k], tau, c);     // f(lambda_k) in place
#pragma unroll 1
    for (int i = 0; i < n; ++i)
#pragma unroll 1
      for (int j = i; j < n; ++j) {
        double t = 0.0;                               // T = V f(Lambda) V^T
#pragma unroll 1
        for (int k = 0; k < n; ++k) t += V[i][k] * V[j][k] * A[k][k];
        res[elem_at(p, tx, i * n + j)] = static_cast<float>(t);
        res[elem_at(p, tx, j * n + i)] = static_cast<float>(t);
      }
  }
}

// ---- mass / comass norms of 2-vectors in R^4 and R^5 (elem_operation_mass_norm.hpp:17-186) ---------------------
// The argument (6 resp. 10 numbers) is the upper triangle of a skew-symmetric matrix M; the mass norm is the sum of
// its singular-value pairs, the prox shrinks them (comass ball: clamps them to 1).  With M = U Sigma V^T:
// prox(M) = U f(Sigma) V^T = M h(M^T M),  h = V diag(f(sigma_k) / sigma_k) V^T  -- a spectral function of the
// SYMMETRIC matrix M^T M, evaluated with the Jacobi rotations above (the reference tridiagonalises the skew matrix
// with Householder / Givens steps and a 2 x 2 SVD, :40-83, :120-175).
template <int NM, bool CONJ>
__global__ void __launch_bounds__(kBlock) spectral_mass_kernel(const SpectralDesc p, float* __restrict__ res,
                                                               const float* __restrict__ arg,
                                                               const float* __restrict__ td, float tau_scal,
                                                               bool invert) {
  for (size_t tx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; tx < p.count; tx += (size_t)gridDim.x * blockDim.x) {
    float c[7];
    load7(p.coeffs, tx, c);
    const float ts = NM == 4 ? tau_scal * c[0] : tau_scal;                 // weighted mass norm (:27)
    const double tau = group_tau(ts, td[elem_at(p, tx, 0)], invert);
    double M[NM][NM];
    int e = 0;
#pragma unroll
    for (int i = 0; i < NM; ++i) {
      M[i][i] = 0.0;
#pragma unroll
      for (int j = i + 1; j < NM; ++j) {
        const double v = arg[elem_at(p, tx, e++)];
        M[i][j] = v;
        M[j][i] = -v;
      }
    }
    double S[NM][NM], V[NM][NM];
#pragma unroll
    for (int i = 0; i < NM; ++i)
#pragma unroll
      for (int j = 0; j < NM; ++j) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < NM; ++k) t += M[k][i] * M[k][j];              // M^T M
        S[i][j] = t;
      }
    jacobi_eig<NM>(S, V, NM);
    double w[NM];
#pragma unroll
    for (int k = 0; k < NM; ++k) {
      const double lam = fmax(S[k][k], 0.0), sg = sqrt(lam);
      const double f = CONJ ? fmin(sg, 1.0) : fmax(sg - tau, 0.0);
      w[k] = sg > 1e-150 ? f / sg : 0.0;
    }
    e = 0;
#pragma unroll
    for (int i = 0; i < NM; ++i)
#pragma unroll
      for (int j = i + 1; j < NM; ++j) {
        double t = 0.0;                                                    // (M h)_{ij} = sum_l M_il h_lj
#pragma unroll
        for (int l = 0; l < NM; ++l) {
          double h = 0.0;
#pragma unroll
          for (int k = 0; k < NM; ++k) h += V[l][k] * V[j][k] * w[k];
          t += M[i][l] * h;
        }
        res[elem_at(p, tx, e++)] = static_cast<float>(t);
      }
  }
}

class ProxSpectral : public Prox {
 public:
  ProxSpectral(Context* ctx, int kind, size_t index, size_t count, size_t dim, bool interleaved, bool diagsteps,
               int fn, int fn2d, const float* const coeffs[7], const size_t len[7])
      : Prox(ctx, index, count * dim, diagsteps), count_(count), dim_(dim), interleaved_(interleaved) {
    if (index + count * dim >= (1ull << 31)) fail(PB_ERR_UNSUPPORTED, "prox range exceeds 2^31-1");
    desc_.kind = kind;
    desc_.fn = fn;
    desc_.fn2d = fn2d;
    desc_.count = static_cast<uint32_t>(count);
    desc_.dim = static_cast<uint32_t>(dim);
    desc_.interleaved = interleaved ? 1 : 0;
    desc_.n = 0;
    switch (kind) {
      case PB_SPECTRAL_SINGULAR_NX2:
        if (dim < 4 || dim % 2) fail(PB_ERR_INVALID, "singular_nx2: dim must be even and >= 4 (an N x 2 matrix, N >= 2)");
        if (fn2d < kFun2DSum1D || fn2d > kFun2DMoreauIndL1Ball) fail(PB_ERR_INVALID, "singular_nx2: unknown Function2D");
        break;
      case PB_SPECTRAL_EIGEN_2X2:
        if (dim != 4) fail(PB_ERR_INVALID, "eigen_2x2: dim must be 4");
        desc_.n = 2;
        break;
      case PB_SPECTRAL_EIGEN_3X3:
        if (dim != 9) fail(PB_ERR_INVALID, "eigen_3x3: dim must be 9");
        desc_.n = 3;
        break;
      case PB_SPECTRAL_EIGEN_NXN: {
        uint32_t n = 1;
        while ((size_t)n * n < dim) ++n;
        if ((size_t)n * n != dim) fail(PB_ERR_INVALID, "eigen_nxn: dim must be a square number");
        if (n > 32) fail(PB_ERR_UNSUPPORTED, "eigen_nxn: matrices larger than 32 x 32 are not supported");  // N_MAX
        desc_.n = n;
        break;
      }
      case PB_SPECTRAL_MASS4:
      case PB_SPECTRAL_COMASS4_BALL:
        if (dim != 6) fail(PB_ERR_INVALID, "mass4 / ind_comass4_ball: dim must be 6");
        break;
      case PB_SPECTRAL_MASS5:
      case PB_SPECTRAL_COMASS5_BALL:
        if (dim != 10) fail(PB_ERR_INVALID, "mass5 / ind_comass5_ball: dim must be 10");
        break;
      default: fail(PB_ERR_INVALID, "unknown spectral operation");
    }
    if (fn < 0 || fn >= PB_FUN_COUNT_) fail(PB_ERR_INVALID, "unknown Function1D");
    for (int k = 0; k < 7; ++k) {
      if (!coeffs[k] || (len[k] != 1 && len[k] != count))
        fail(PB_ERR_INVALID, "spectral operation: every coefficient needs 1 or count entries");
      desc_.coeffs.ptr[k] = nullptr;
      desc_.coeffs.val[k] = coeffs[k][0];
      if (len[k] > 1) {
        d_co_[k].resize(len[k]);
        upload_from_host(ctx, d_co_[k].data(), coeffs[k], len[k]);
        desc_.coeffs.ptr[k] = d_co_[k].data();
      }
    }
  }
  int kind() const override { return kProxSpectral; }
  size_t uniform_group_size() const override { return dim_; }
  void get_separable_structure(std::vector<std::tuple<size_t, size_t, size_t>>& sep) const override {
    for (size_t i = 0; i < count_; ++i) {                 // ProxSeparableSum (prox_separable_sum.hpp:65-77)
      if (interleaved_) sep.emplace_back(index_ + i * dim_, dim_, 1);
      else sep.emplace_back(index_ + i, dim_, count_);
    }
  }
  size_t gpu_mem_amount() const override {
    size_t n = 0;
    for (auto& d : d_co_) n += d.size();
    return n * sizeof(float);
  }
  void eval_local(float* res, const float* arg, const float* td, float tau, bool invert) override {
    ctx_->bind();
    if (count_ == 0) return;
    const unsigned grid = stream_grid(ctx_, count_);
    cudaStream_t s = ctx_->stream;
    switch (desc_.kind) {
      case PB_SPECTRAL_SINGULAR_NX2:
        spectral_singular_nx2_kernel<<<grid, kBlock, 0, s>>>(desc_, res, arg, td, tau, invert);
        break;
      case PB_SPECTRAL_EIGEN_2X2:
        spectral_eigen_2x2_kernel<<<grid, kBlock, 0, s>>>(desc_, res, arg, td, tau, invert);
        break;
      case PB_SPECTRAL_MASS4: spectral_mass_kernel<4, false><<<grid, kBlock, 0, s>>>(desc_, res, arg, td, tau, invert); break;
      case PB_SPECTRAL_COMASS4_BALL: spectral_mass_kernel<4, true><<<grid, kBlock, 0, s>>>(desc_, res, arg, td, tau, invert); break;
      case PB_SPECTRAL_MASS5: spectral_mass_kernel<5, false><<<grid, kBlock, 0, s>>>(desc_, res, arg, td, tau, invert); break;
      case PB_SPECTRAL_COMASS5_BALL: spectral_mass_kernel<5, true><<<grid, kBlock, 0, s>>>(desc_, res, arg, td, tau, invert); break;
      default:
        if (desc_.n <= 3) spectral_eigen_nxn_kernel<3><<<grid, kBlock, 0, s>>>(desc_, res, arg, td, tau, invert);
        else if (desc_.n <= 5) spectral_eigen_nxn_kernel<5><<<grid, kBlock, 0, s>>>(desc_, res, arg, td, tau, invert);
        else if (desc_.n <= 8) spectral_eigen_nxn_kernel<8><<<grid, kBlock, 0, s>>>(desc_, res, arg, td, tau, invert);
        else spectral_eigen_big_kernel<<<(unsigned)std::min<size_t>((desc_.count + 63) / 64, (size_t)ctx_->num_sms * 8), 64, 0, s>>>(
            desc_, res, arg, td, tau, invert);
        break;
    }
    PB_CHECK_LAUNCH();
    ctx_->launches++;
  }

 private:
  SpectralDesc desc_;
  size_t count_, dim_;
  bool interleaved_;
  DeviceBuffer<float> d_co_[7];
};

}  // namespace

std::shared_ptr<Prox> make_prox_spectral(Context* ctx, int kind, size_t index, size_t count, size_t dim,
                                         bool interleaved, bool diagsteps, int function_1d, int function_2d,
                                         const float* const coeffs[7], const size_t coeff_len[7]) {
  return std::make_shared<ProxSpectral>(ctx, kind, index, count, dim, interleaved, diagsteps, function_1d, function_2d,
                                        coeffs, coeff_len);
}

}  // namespace pb
