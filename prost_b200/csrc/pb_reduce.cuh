// pb_reduce.cuh -- warp-shuffle block reductions with double accumulators.
//
// The reference reduces residual norms with thrust::transform_reduce in float
// (backend_pdhg.cu:392-431).  Here every thread accumulates in double, a CTA combines its
// threads with warp shuffles + one shared-memory hop, and writes ONE partial per CTA; a
// single-CTA kernel folds the partials in index order, so results are deterministic and do
// not lose digits at 10^8 elements.
#pragma once

#include <cuda_runtime.h>

namespace pb {

#ifdef __CUDACC__

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// Sums (a, b) over the CTA; the result is valid in thread 0.  All threads must call.
__device__ __forceinline__ void block_sum2(double& a, double& b) {
  __shared__ double sh[2][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  a = warp_sum(a);
  b = warp_sum(b);
  if (lane == 0) { sh[0][warp] = a; sh[1][warp] = b; }
  __syncthreads();
  if (warp == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    a = lane < nw ? sh[0][lane] : 0.0;
    b = lane < nw ? sh[1][lane] : 0.0;
    a = warp_sum(a);
    b = warp_sum(b);
  }
  __syncthreads();
}

// out[0..1] = sum over `n` pairs of partials, folded in index order by one CTA.
__device__ __forceinline__ void fold_partials2(const double* __restrict__ part, unsigned n,
                                               double& a, double& b) {
  a = 0.0;
  b = 0.0;
  for (unsigned i = threadIdx.x; i < n; i += blockDim.x) {
    a += part[2 * i];
    b += part[2 * i + 1];
  }
  block_sum2(a, b);
}

#endif

}  // namespace pb
