// pb_comm.cuh -- one-process-per-GPU communicator for the slab decomposition.
//
// The reference is single-GPU (no NCCL/MPI/streams anywhere); this layer is what
// lets the fused PDHG passes run on a column slab of a larger grid:
//   * stencil halos travel peer-to-peer: every rank owns a small block of device memory (two
//     ping-pong slots per direction + sequence flags) that its neighbours map through CUDA IPC,
//     and the producing KERNEL stores its edge column straight into the neighbour's slot over
//     NVLink, then publishes a sequence number; the consuming kernel's edge threads spin on that
//     number, so the transfer overlaps the interior of both passes and no copy kernel, no
//     NCCL call and no host synchronisation sits on the iteration path;
//   * the four residual sums are combined with one ncclAllReduce (4 doubles) on the iteration
//     stream at residual_iter boundaries only; every rank then runs the identical step-size
//     state machine on identical bits.
// If CUDA IPC is unavailable (e.g. PB_HALO=nccl, or the mapping fails) the same kernels write the
// edge column to a local staging buffer and the halos move with ncclSend/ncclRecv between the passes.
//
// NCCL is loaded lazily (dlopen libnccl.so.2) so that single-GPU users carry no dependency and a
// process that already loaded torch's bundled NCCL shares that copy.
#pragma once

#include "pb_common.cuh"
#include "pb_crosssum.cuh"

namespace pb {

// device-visible control words of one rank (lives at the start of the IPC block)
struct HaloFlags {
  unsigned x_seq;        // sequence number of the newest x halo written by the RIGHT neighbour
  unsigned y_seq;        // sequence number of the newest y halo written by the LEFT neighbour
  unsigned done_primal;  // local: edge CTAs of the running primal pass that finished
  unsigned done_dual;    // local: same for the dual pass
  int error;             // set by a kernel whose halo wait timed out
  unsigned red_count;    // local: reductions over ranks executed so far (pb_crosssum.cuh)
  int pad[10];
};

class Comm {
 public:
  // collective over all ranks; `id` = 128 bytes from unique_id() on rank 0
  Comm(Context* ctx, int rank, int world, const void* id);
  ~Comm();
  static void unique_id(void* out128);

  int rank() const { return rank_; }
  int world() const { return world_; }
  bool p2p() const { return p2p_; }
  bool has_left() const { return rank_ > 0; }
  bool has_right() const { return rank_ + 1 < world_; }
  Context* ctx() const { return ctx_; }

  // collective: (re)allocates the halo block for edge columns of `col_floats` floats and maps the
  // neighbours' blocks; resets the flags.  All ranks must pass the same size.
  void ensure_halo(size_t col_floats);
  size_t col_floats() const { return col_floats_; }

  // in-stream sum over ranks of n doubles (device memory, in place)
  void allreduce_sum(double* d_buf, size_t n);
  // in-stream sum over ranks of n floats: the partial K^T r vectors of the row-sharded ADMM
  void allreduce_sum_f32(float* d_buf, size_t n);
  // host-side conveniences (synchronise the stream)
  void allreduce_sum_host(double* h_buf, size_t n);
  void barrier();

  // --- halo plumbing used by BackendPDHG (see pb_pdhg.cu: slab mode) -------------------------------
  // local slots: what THIS rank reads
  float* x_slot(unsigned seq) const { return x_in_ + (seq & 1u) * col_floats_; }   // from the right
  float* y_slot(unsigned seq) const { return y_in_ + (seq & 1u) * col_floats_; }   // from the left
  // where THIS rank writes its edge columns: the neighbour's slot (p2p) or the local staging buffer
  float* x_out(unsigned seq) const;   // my column 0 of x      -> left neighbour's x slot
  float* y_out(unsigned seq) const;   // my last column of y.gx -> right neighbour's y slot
  HaloFlags* flags() const { return flags_; }
  unsigned* left_x_seq() const;       // left neighbour's  flags->x_seq (p2p only, else null)
  unsigned* right_y_seq() const;      // right neighbour's flags->y_seq (p2p only, else null)
  // NCCL staging mode: move the staged edge column to the neighbour (no-ops in p2p mode)
  void exchange_x(unsigned seq);
  void exchange_y(unsigned seq);
  // pass sequence numbers (identical on every rank: all ranks run the same passes)
  unsigned x_seq = 0, y_seq = 0;
  // --- flag-in-data slots of the one-pass ring kernel (RingHalo, pb_stencil.cuh): two 16-byte lines per four rows
  const uint4* x_ll(unsigned seq) const { return x_ll_ + (seq & 1u) * ll_lines(); }      // local, from the right
  const uint4* y_ll(unsigned seq) const { return y_ll_ + (seq % 3u) * ll_lines(); }      // local, from the left
  uint4* x_ll_out(unsigned seq) const;         // the left neighbour's x slot (p2p only)
  uint4* y_ll_out(unsigned seq) const;         // the right neighbour's y slot
  size_t ll_lines() const { return (col_floats_ + 3) / 4 * 2; }        // 16-byte lines per slot
  // after a two-pass iteration (plain slots + sequence words): repack the received columns x(xs), y(ys) into the
  // local flag-in-data slots, so that a following one-pass iteration finds them there
  void pack_ll(unsigned xs, unsigned ys);
  // before the plain slots are read on the host side of current_solution: the reverse, for one sequence number
  void unpack_ll_x(unsigned seq);
  void unpack_ll_y(unsigned seq);
  // throws if a kernel reported a halo timeout
  void check_error();

  // --- in-kernel sum over ranks (RingFinish, pb_stencil.cuh): every rank's block carries
  // [2][kMaxReduceRanks][4] double slots + one sequence word per writer; all blocks are peer-mapped
  bool reduce_p2p() const;                    // p2p mode, world <= kMaxReduceRanks, every block mapped
  const double* red_in() const;               // local slots
  const unsigned* red_flag_in() const;        // local sequence words
  double* red_out(int r) const;               // rank r's slots (own block for r == rank)
  unsigned* red_flag_out(int r) const;
  unsigned* red_count() const;                // local device counter of executed reductions (pb_crosssum.cuh)
  // descriptor for kernels that sum over ranks themselves (world 1: no-op descriptor)
  struct CrossSum cross_sum() const;

 private:
  void release_halo();
  Context* ctx_;
  int rank_, world_;
  void* nccl_ = nullptr;               // ncclComm_t
  bool p2p_ = true;
  size_t col_floats_ = 0;
  // local IPC block: [header: HaloFlags, reduce words and slots | x_in[2][col] | y_in[2][col] | x_ll[2][2 col] |
  // y_ll[3][2 col]]
  void* block_ = nullptr;
  HaloFlags* flags_ = nullptr;
  float* x_in_ = nullptr;
  float* y_in_ = nullptr;
  uint4* x_ll_ = nullptr;              // [2][ll_lines]
  uint4* y_ll_ = nullptr;              // [3][ll_lines]
  void* left_block_ = nullptr;         // neighbours' blocks mapped into this process (p2p)
  void* right_block_ = nullptr;
  std::vector<void*> peer_blocks_;     // every rank's block (own: block_), p2p only
  DeviceBuffer<float> stage_x_, stage_y_;   // staging mode
  DeviceBuffer<double> scratch_;
};

}  // namespace pb
