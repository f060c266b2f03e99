// prost/solver.hpp -- Solver<T>: iteration loop, stopping criteria and callbacks
// (reference: include/prost/solver.hpp:39-117, src/solver.cu).
#ifndef PROST_SOLVER_HPP_
#define PROST_SOLVER_HPP_

#include "prost/common.hpp"

namespace prost {

template <typename T> class Problem;
template <typename T> class Backend;

template <typename T>
class Solver : detail::require_float<T> {
 public:
  struct Options {
    T tol_rel_primal, tol_rel_dual, tol_abs_primal, tol_abs_dual;
    int max_iters;
    /// how often the intermediate-solution callback runs within max_iters
    int num_cback_calls;
    bool verbose;
    vector<T> x0, y0;
    bool solve_dual_problem;
  };

  enum ConvergenceResult { kConverged, kStoppedMaxIters, kStoppedUser };

  /// (iteration, primal solution, dual solution) -> true to stop as converged
  typedef function<bool(int, const vector<T>&, const vector<T>&)> IntermCallback;
  /// called once per iteration; true terminates the solver
  typedef function<bool()> StoppingCallback;

  Solver(shared_ptr<Problem<T> > problem, shared_ptr<Backend<T> > backend);
  virtual ~Solver() {}

  void Initialize();
  typename Solver<T>::ConvergenceResult Solve();
  void Release();

  void SetOptions(const typename Solver<T>::Options& opts) { opts_ = opts; }
  void SetStoppingCallback(const StoppingCallback& cb) { stopping_cb_ = cb; }
  void SetIntermCallback(const IntermCallback& cb) { interm_cb_ = cb; }

  const vector<T>& cur_primal_sol() const { return opts_.solve_dual_problem ? cur_dual_sol_ : cur_primal_sol_; }
  const vector<T>& cur_dual_sol() const { return opts_.solve_dual_problem ? cur_primal_sol_ : cur_dual_sol_; }
  const vector<T>& cur_primal_constr_sol() const {
    return opts_.solve_dual_problem ? cur_dual_constr_sol_ : cur_primal_constr_sol_;
  }
  const vector<T>& cur_dual_constr_sol() const {
    return opts_.solve_dual_problem ? cur_primal_constr_sol_ : cur_dual_constr_sol_;
  }
  int iterations() const { return iterations_; }

 protected:
  static int stop_trampoline(void* user) {
    Solver<T>* s = static_cast<Solver<T>*>(user);
    return (s->stopping_cb_ && s->stopping_cb_()) ? 1 : 0;
  }
  static int interm_trampoline(void* user, int it, const float* p, size_t np, const float* d, size_t nd) {
    Solver<T>* s = static_cast<Solver<T>*>(user);
    if (!s->interm_cb_) return 0;
    const vector<T> pv(p, p + np), dv(d, d + nd);
    return s->interm_cb_(it, pv, dv) ? 1 : 0;
  }
  static pb_solver_options to_c(const Options& o) {
    pb_solver_options c;
    c.tol_rel_primal = o.tol_rel_primal; c.tol_rel_dual = o.tol_rel_dual;
    c.tol_abs_primal = o.tol_abs_primal; c.tol_abs_dual = o.tol_abs_dual;
    c.max_iters = o.max_iters; c.num_cback_calls = o.num_cback_calls;
    c.verbose = o.verbose ? 1 : 0; c.solve_dual_problem = o.solve_dual_problem ? 1 : 0;
    return c;
  }

  Options opts_;
  shared_ptr<Problem<T> > problem_;
  shared_ptr<Backend<T> > backend_;
  vector<T> cur_primal_sol_, cur_dual_sol_, cur_primal_constr_sol_, cur_dual_constr_sol_;
  IntermCallback interm_cb_;
  StoppingCallback stopping_cb_;
  int iterations_;

  template <typename U> friend class Backend;
};

}  // namespace prost

#include "prost/backend/backend.hpp"
#include "prost/problem.hpp"

namespace prost {

template <typename T>
Solver<T>::Solver(shared_ptr<Problem<T> > problem, shared_ptr<Backend<T> > backend)
    : problem_(problem), backend_(backend), iterations_(0) {
  opts_.tol_rel_primal = opts_.tol_rel_dual = opts_.tol_abs_primal = opts_.tol_abs_dual = 1e-4f;
  opts_.max_iters = 1000;
  opts_.num_cback_calls = 10;
  opts_.verbose = true;
  opts_.solve_dual_problem = false;
}

template <typename T>
void Solver<T>::Initialize() {
  try {
    problem_->Initialize();
  } catch (Exception& e) {
    throw Exception(string("Failed to initialize the problem. Reason: ") + e.what());
  }
  if (opts_.solve_dual_problem) {
    problem_->Dualize();
    opts_.x0.swap(opts_.y0);
  }
  try {
    backend_->SetProblem(problem_);
    backend_->SetOptions(opts_);
    backend_->Initialize();
  } catch (Exception& e) {
    throw Exception(string("Failed to initialize the backend. Reason: ") + e.what());
  }
  if (opts_.verbose) {
    const size_t mem = problem_->gpu_mem_amount() + backend_->gpu_mem_amount();
    std::cout << "# primal variables: " << problem_->ncols() << std::endl;
    std::cout << "# dual variables: " << problem_->nrows() << std::endl;
    std::cout << "Memory requirements: " << mem / (1024 * 1024) << "MB." << std::endl;
  }
  cur_primal_sol_.resize(problem_->ncols());
  cur_primal_constr_sol_.resize(problem_->nrows());
  cur_dual_sol_.resize(problem_->nrows());
  cur_dual_constr_sol_.resize(problem_->ncols());
}

template <typename T>
typename Solver<T>::ConvergenceResult Solver<T>::Solve() {
  const pb_solver_options c = to_c(opts_);
  int result = PB_STOPPED_MAX_ITERS, iters = 0;
  detail::check(pb_solver_solve(backend_->handle(), &c, &Solver<T>::stop_trampoline, &Solver<T>::interm_trampoline,
                                this, cur_primal_sol_.data(), cur_primal_constr_sol_.data(),
                                cur_dual_sol_.data(), cur_dual_constr_sol_.data(), &result, &iters));
  iterations_ = iters;
  if (opts_.solve_dual_problem) {       // restore the original problem (solver.cu:198-203)
    problem_->Dualize();
    opts_.x0.swap(opts_.y0);
  }
  return result == PB_CONVERGED ? kConverged : (result == PB_STOPPED_USER ? kStoppedUser : kStoppedMaxIters);
}

template <typename T>
void Solver<T>::Release() {
  problem_->Release();
  backend_->Release();
}

}  // namespace prost

#endif
