// pb_admm.cu -- BackendADMM (graph-projection ADMM + CGLS).  Placeholder until the ADMM row of
// SURVEY.md section 8 (a23) is built; creation fails loudly instead of silently doing nothing.
#include "pb_backend.cuh"

namespace pb {

std::shared_ptr<Backend> make_backend_admm(Context*, std::shared_ptr<Problem>, const pb_admm_options&,
                                           const pb_solver_options&) {
  fail(PB_ERR_UNSUPPORTED, "BackendADMM is not implemented yet");
}

}  // namespace pb
