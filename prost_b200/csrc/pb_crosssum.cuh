// pb_crosssum.cuh -- sum of a few doubles over the ranks of one box from INSIDE a kernel.
//
// Every rank owns [2][kMaxReduceRanks][4] double slots and one sequence word per writer in its IPC block
// (pb_comm.cuh); all blocks are peer-mapped.  The calling thread (one per rank: the last CTA's thread 0 of a
// reduction kernel) stores its values into slot [count & 1][my rank] of EVERY rank over NVLink, then the
// sequence number `count`; it waits until every rank's sequence word in its own block has reached `count` and
// adds the slots in rank order, so all ranks obtain identical bits.  `count` is a device-resident counter that
// every rank advances by one per executed reduction (the ranks execute the same sequence of reductions), which
// keeps the two slot sets strictly alternating: a writer can only be one reduction ahead of a reader, because
// completing reduction c requires every rank's sequence word for c, which a rank publishes after it has read
// the slots of c - 1.
// Replaces a fold kernel + ncclAllReduce of 4 doubles + a finalize kernel per reduction.
#pragma once

#include <cuda_runtime.h>

namespace pb {

constexpr int kMaxReduceRanks = 8;

struct CrossSum {
  int world = 1, rank = 0;
  unsigned* count = nullptr;                     // local: number of reductions executed so far
  const double* red_in = nullptr;                // local slots [2][kMaxReduceRanks][4]
  const unsigned* red_flag_in = nullptr;         // local sequence words [kMaxReduceRanks]
  double* red_out[kMaxReduceRanks] = {};         // every rank's slots (own included)
  unsigned* red_flag_out[kMaxReduceRanks] = {};
  int* error = nullptr;                          // set when a wait times out (~2 s): no GPU hang
};

#ifdef __CUDACC__

// ONE thread per rank.  v[0..3] in, sums over ranks out.
__device__ __forceinline__ void cross_rank_sum4(const CrossSum& cs, double (&v)[4]) {
  if (cs.world <= 1) return;
  const unsigned seq = *cs.count + 1u;
  *cs.count = seq;
  const unsigned slot = (seq & 1u) * kMaxReduceRanks;
  for (int r = 0; r < cs.world; ++r) {
    double* o = cs.red_out[r] + (size_t)(slot + cs.rank) * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) __stcg(o + j, v[j]);
  }
  __threadfence_system();
  for (int r = 0; r < cs.world; ++r)
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(cs.red_flag_out[r] + cs.rank), "r"(seq) : "memory");
#pragma unroll
  for (int j = 0; j < 4; ++j) v[j] = 0.0;
  for (int r = 0; r < cs.world; ++r) {
    unsigned f;
    unsigned long long t0 = 0;
    for (unsigned spins = 0;; ++spins) {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(f) : "l"(cs.red_flag_in + r) : "memory");
      if ((int)(f - seq) >= 0) break;
      if ((spins & 1023u) == 1023u) {
        if (cs.error && *reinterpret_cast<volatile int*>(cs.error)) break;
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > 2000000000ull) { if (cs.error) atomicExch(cs.error, 1); break; }
      }
    }
    const double* in = cs.red_in + (size_t)(slot + r) * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] += __ldcg(in + j);
  }
}

#endif  // __CUDACC__

}  // namespace pb
