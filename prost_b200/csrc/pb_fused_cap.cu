// pb_fused_cap.cu -- one translation unit per register capacity PB_CAP (1,2,4,...,64) so the
// fused-pass instantiations compile in parallel.  See pb_fused.cuh for the kernels.
#include "pb_fused.cuh"

#ifndef PB_CAP
#error "compile with -DPB_CAP=<1|2|4|8|16|32|64>"
#endif

namespace pb {

#define PB_CAT2(a, b) a##b
#define PB_CAT(a, b) PB_CAT2(a, b)

template <bool CHECK>
static void primal_launch(Context* ctx, unsigned grid, const ProxDesc& d, const BlockList& bl,
                          const float* x, const float* y, const float* y_prev, ScaleRef T,
                          const PdhgState* st, bool kty_zero, bool ktyprev_zero, double* partials,
                          float* x_out) {
  PrimalSource<PB_CAP, CHECK> src;
  src.x = x; src.y = y; src.y_prev = y_prev; src.T = T; src.st = st; src.bl = bl;
  src.kty_zero = kty_zero; src.ktyprev_zero = ktyprev_zero; src.partials = partials;
  prox_pass_kernel<PB_CAP, PrimalSource<PB_CAP, CHECK>><<<grid, kBlock, 0, ctx->stream>>>(d, src, x_out, T, false);
}

template <bool CHECK>
static void dual_launch(Context* ctx, unsigned grid, const ProxDesc& d, const BlockList& bl,
                        const float* y, const float* x_new, const float* x_old, ScaleRef S,
                        const PdhgState* st, bool kxprev_zero, double* partials, float* y_out) {
  DualSource<PB_CAP, CHECK> src;
  src.y = y; src.x_new = x_new; src.x_old = x_old; src.S = S; src.st = st; src.bl = bl;
  src.kxprev_zero = kxprev_zero; src.partials = partials;
  prox_pass_kernel<PB_CAP, DualSource<PB_CAP, CHECK>><<<grid, kBlock, 0, ctx->stream>>>(d, src, y_out, S, false);
}

void PB_CAT(fused_primal_cap_, PB_CAP)(Context* ctx, unsigned grid, const ProxDesc& d, const BlockList& bl,
                                       const float* x, const float* y, const float* y_prev, ScaleRef T,
                                       const PdhgState* st, bool kty_zero, bool ktyprev_zero, bool check,
                                       double* partials, float* x_out) {
  if (check) primal_launch<true>(ctx, grid, d, bl, x, y, y_prev, T, st, kty_zero, ktyprev_zero, partials, x_out);
  else primal_launch<false>(ctx, grid, d, bl, x, y, y_prev, T, st, kty_zero, ktyprev_zero, partials, x_out);
}

void PB_CAT(fused_dual_cap_, PB_CAP)(Context* ctx, unsigned grid, const ProxDesc& d, const BlockList& bl,
                                     const float* y, const float* x_new, const float* x_old, ScaleRef S,
                                     const PdhgState* st, bool kxprev_zero, bool check, double* partials,
                                     float* y_out) {
  if (check) dual_launch<true>(ctx, grid, d, bl, y, x_new, x_old, S, st, kxprev_zero, partials, y_out);
  else dual_launch<false>(ctx, grid, d, bl, y, x_new, x_old, S, st, kxprev_zero, partials, y_out);
}

}  // namespace pb
