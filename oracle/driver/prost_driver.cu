// oracle/driver/prost_driver.cu -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A MATLAB-free driver that builds a problem through prost's PUBLIC C++ API only
// (Problem / Block* / Prox* / BackendPDHG / Solver, include/prost/*.hpp) from a small text
// description and either solves it, applies the linear operator, or evaluates a prox.
//
// The SAME source is compiled twice:
//   * against the unmodified reference headers + oracle/_ref/libprost_ref.a   -> oracle/_ref/prost_ref_driver
//     (the live GPU oracle: the reference's own CUDA kernels on the same B200)
//   * against this repo's include/prost/*.hpp + libprost_b200.so               -> prost_b200/lib/prost_b200_driver
//     (shows that an existing problem definition drops in unchanged)
// tests/test_reference_parity.py runs both on identical inputs and compares the dumps.
//
// Description format (one statement per line, arrays are raw little-endian files next to it):
//   dims <nrows> <ncols>
//   block gradient2d|gradient3d <row> <col> <nx> <ny> <L> <label_first>
//   block diags <row> <col> <nrows> <ncols> <ndiags> <offsets.i64> <factors.f32>
//   block sparse <row> <col> <m> <n> <nnz> <val.f32> <ptr.i32> <ind.i32>        (CSC)
//   block dense <row> <col> <nrows> <ncols> <data.f32>                           (column-major)
//   block dense_kron_id|id_kron_dense <row> <col> <mat_nrows> <mat_ncols> <diaglength> <data.f32 column-major>
//   block sparse_kron_id|id_kron_sparse <row> <col> <diaglength> <m> <n> <nnz> <val.f32> <ptr.i32> <ind.i32>   (CSC)
//   block zero <row> <col> <nrows> <ncols>
//   prox g|f|gstar|fstar|eval <PROX>
//     PROX := elem1d|norm2 <fun> <idx> <count> <dim> <interleaved> <diagsteps> <c0> .. <c6>
//           | simplex <idx> <count> <dim> <interleaved> <diagsteps>
//           | indsum <idx> <count> <dim> <interleaved> <diagsteps>
//           | halfspace <idx> <count> <dim> <interleaved> <diagsteps> <a> <b>
//           | spectral <singular_nx2|eigen_2x2|eigen_3x3|eigen_nxn> <fun> <idx> <count> <dim> <interleaved> <diagsteps> <7 coeffs>
//           | massnorm <mass4|ind_comass4_ball|mass5|ind_comass5_ball> <idx> <count> <dim> <interleaved> <diagsteps> <cost>
//           | indrange <idx> <size> <diagsteps> <m> <n> <nnz> <val file> <ptr file> <ind file> <AA file (n*n, column-major)>
//           | indsumidx <idx> <size> <n_lists: 1|2> { <dim> <inds file (u64)> <n_inds> <sum> } x n_lists
//           | soc <idx> <count> <dim> <interleaved> <diagsteps> <alpha>
//           | epiquad <idx> <count> <dim> <interleaved> <diagsteps> <a> <b> <c>
//           | moreau <PROX> | permute <perm.i32> <n> <PROX> | zero <idx> <size>
//           | transform <a> <b> <c> <d> <e> <PROX>
//     coefficient := s:<value> | f:<file>:<length>
//   scaling alpha <a> | identity | custom <left.f32> <right.f32>
//   pdhg <tau0> <sigma0> <residual_iter> <scale_steps_operator> <alg2_gamma> <arg_alpha0> <arg_nu>
//        <arg_delta> <arb_delta> <arb_tau> <variant 1..4>
//   admm <rho0> <alpha> <cg_tol_pow> <cg_tol_min> <cg_tol_max> <cg_max_iter> <residual_iter> <arb_delta>
//        <arb_tau> <arb_gamma>                                  (selects BackendADMM instead of BackendPDHG)
//   solver <tol_rel_p> <tol_rel_d> <tol_abs_p> <tol_abs_d> <max_iters> <num_cback> <x0|-> <y0|-> <solve_dual>
//   action solve | linop <in.f32> <transpose> | prox <arg.f32> <taudiag.f32> <tau>
//   out <prefix>
#include <array>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "prost/backend/backend_admm.hpp"
#include "prost/backend/backend_pdhg.hpp"
#include "prost/exception.hpp"
#include "prost/linop/block_dense.hpp"
#include "prost/linop/block_dense_kron_id.hpp"
#include "prost/linop/block_id_kron_dense.hpp"
#include "prost/linop/block_id_kron_sparse.hpp"
#include "prost/linop/block_sparse_kron_id.hpp"
#include "prost/linop/block_diags.hpp"
#include "prost/linop/block_gradient2d.hpp"
#include "prost/linop/block_gradient3d.hpp"
#include "prost/linop/block_sparse.hpp"
#include "prost/linop/block_zero.hpp"
#include "prost/linop/linearoperator.hpp"
#include "prost/problem.hpp"
#include "prost/prox/elemop/elem_operation_1d.hpp"
#include "prost/prox/elemop/elem_operation_ind_simplex.hpp"
#include "prost/prox/elemop/elem_operation_ind_sum.hpp"
#include "prost/prox/elemop/elem_operation_norm2.hpp"
#include "prost/prox/elemop/function_1d.hpp"
#include "prost/prox/prox_elem_operation.hpp"
#include "prost/prox/prox_ind_epi_quad.hpp"
#include "prost/prox/prox_moreau.hpp"
#include "prost/prox/prox_ind_halfspace.hpp"
#include "prost/prox/prox_ind_sum.hpp"
#include "prost/prox/prox_ind_range.hpp"
#include "prost/prox/elemop/elem_operation_singular_nx2.hpp"
#include "prost/prox/elemop/elem_operation_eigen_2x2.hpp"
#include "prost/prox/elemop/elem_operation_eigen_3x3.hpp"
#include "prost/prox/elemop/elem_operation_eigen_nxn.hpp"
#include "prost/prox/elemop/function_2d.hpp"
#include "prost/prox/elemop/elem_operation_mass_norm.hpp"
#include "prost/prox/prox_ind_soc.hpp"
#include "prost/prox/prox_transform.hpp"
#include "prost/prox/prox_permute.hpp"
#include "prost/prox/prox_zero.hpp"
#include "prost/solver.hpp"

using namespace prost;
#ifdef PROST_DRIVER_DOUBLE
typedef double real;      // reference only: double-precision trajectory for conditioning studies
#else
typedef float real;
#endif

static std::string g_dir;

template <typename T>
static std::vector<T> read_file(const std::string& name, size_t n) {
  std::vector<T> v(n);
  std::ifstream f(g_dir + "/" + name, std::ios::binary);
  if (!f) { std::cerr << "cannot open " << name << std::endl; std::exit(2); }
  f.read(reinterpret_cast<char*>(v.data()), n * sizeof(T));
  if (static_cast<size_t>(f.gcount()) != n * sizeof(T)) { std::cerr << "short read " << name << std::endl; std::exit(2); }
  return v;
}

// real-valued arrays are always stored as float32 on disk
static std::vector<real> read_real(const std::string& name, size_t n) {
  std::vector<float> f = read_file<float>(name, n);
  return std::vector<real>(f.begin(), f.end());
}

static void write_file(const std::string& path, const std::vector<real>& v) {
  std::vector<float> f32(v.begin(), v.end());
  std::ofstream f(path, std::ios::binary);
  f.write(reinterpret_cast<const char*>(f32.data()), f32.size() * sizeof(float));
}

// s:<value>  or  f:<file>:<length>
static std::vector<real> coeff(const std::string& tok) {
  if (tok.compare(0, 2, "s:") == 0) return std::vector<real>(1, static_cast<real>(std::atof(tok.c_str() + 2)));
  const size_t c2 = tok.rfind(':');
  return read_real(tok.substr(2, c2 - 2), static_cast<size_t>(std::atoll(tok.c_str() + c2 + 1)));
}

template <template <typename> class FUN>
static Prox<real>* make_elem(bool norm2, size_t idx, size_t count, size_t dim, bool il, bool ds,
                             const std::array<std::vector<real>, 7>& c) {
  if (norm2) return new ProxElemOperation<real, ElemOperationNorm2<real, FUN<real>>>(idx, count, dim, il, ds, c);
  return new ProxElemOperation<real, ElemOperation1D<real, FUN<real>>>(idx, count, dim, il, ds, c);
}

static std::shared_ptr<Prox<real>> parse_prox(std::istringstream& in) {
  std::string kind;
  in >> kind;
  if (kind == "elem1d" || kind == "norm2") {
    std::string fun;
    size_t idx, count, dim;
    int il, ds;
    in >> fun >> idx >> count >> dim >> il >> ds;
    std::array<std::vector<real>, 7> c;
    for (int k = 0; k < 7; ++k) { std::string t; in >> t; c[k] = coeff(t); }
    const bool n2 = kind == "norm2";
    Prox<real>* p = nullptr;
    if (fun == "zero") p = make_elem<Function1DZero>(n2, idx, count, dim, il, ds, c);
    else if (fun == "abs") p = make_elem<Function1DAbs>(n2, idx, count, dim, il, ds, c);
    else if (fun == "square") p = make_elem<Function1DSquare>(n2, idx, count, dim, il, ds, c);
    else if (fun == "ind_leq0") p = make_elem<Function1DIndLeq0>(n2, idx, count, dim, il, ds, c);
    else if (fun == "ind_geq0") p = make_elem<Function1DIndGeq0>(n2, idx, count, dim, il, ds, c);
    else if (fun == "ind_eq0") p = make_elem<Function1DIndEq0>(n2, idx, count, dim, il, ds, c);
    else if (fun == "ind_box01") p = make_elem<Function1DIndBox01>(n2, idx, count, dim, il, ds, c);
    else if (fun == "max_pos0") p = make_elem<Function1DMaxPos0>(n2, idx, count, dim, il, ds, c);
    else if (fun == "l0") p = make_elem<Function1DL0>(n2, idx, count, dim, il, ds, c);
    else if (fun == "huber") p = make_elem<Function1DHuber>(n2, idx, count, dim, il, ds, c);
    else if (fun == "lq") p = make_elem<Function1DLq>(n2, idx, count, dim, il, ds, c);
    else if (fun == "lq_plus_eps") p = make_elem<Function1DLqPlusEps>(n2, idx, count, dim, il, ds, c);
    else if (fun == "truncquad") p = make_elem<Function1DTruncQuad>(n2, idx, count, dim, il, ds, c);
    else if (fun == "trunclin") p = make_elem<Function1DTruncLinear>(n2, idx, count, dim, il, ds, c);
    else { std::cerr << "unknown function " << fun << std::endl; std::exit(2); }
    return std::shared_ptr<Prox<real>>(p);
  }
  if (kind == "simplex") {
    size_t idx, count, dim;
    int il, ds;
    in >> idx >> count >> dim >> il >> ds;
    return std::shared_ptr<Prox<real>>(
        new ProxElemOperation<real, ElemOperationIndSimplex<real>>(idx, count, dim, il, ds));
  }
  if (kind == "spectral") {
    std::string op, fun;
    size_t idx, count, dim;
    int il, ds;
    in >> op >> fun >> idx >> count >> dim >> il >> ds;
    std::array<std::vector<real>, 7> c;
    for (int k = 0; k < 7; ++k) { std::string tok; in >> tok; c[k] = coeff(tok); }
    Prox<real>* p = nullptr;
#define PB_SPEC_EIG(OPNAME, CLASS)                                                                                       \
    if (op == OPNAME) {                                                                                                \
      if (fun == "zero") p = new ProxElemOperation<real, CLASS<real, Function1DZero<real>>>(idx, count, dim, il, ds, c);          \
      else if (fun == "abs") p = new ProxElemOperation<real, CLASS<real, Function1DAbs<real>>>(idx, count, dim, il, ds, c);       \
      else if (fun == "square") p = new ProxElemOperation<real, CLASS<real, Function1DSquare<real>>>(idx, count, dim, il, ds, c); \
      else if (fun == "ind_leq0") p = new ProxElemOperation<real, CLASS<real, Function1DIndLeq0<real>>>(idx, count, dim, il, ds, c); \
      else if (fun == "ind_geq0") p = new ProxElemOperation<real, CLASS<real, Function1DIndGeq0<real>>>(idx, count, dim, il, ds, c); \
      else if (fun == "ind_box01") p = new ProxElemOperation<real, CLASS<real, Function1DIndBox01<real>>>(idx, count, dim, il, ds, c); \
      else if (fun == "huber") p = new ProxElemOperation<real, CLASS<real, Function1DHuber<real>>>(idx, count, dim, il, ds, c);   \
    }
    PB_SPEC_EIG("eigen_2x2", ElemOperationEigen2x2)
    PB_SPEC_EIG("eigen_3x3", ElemOperationEigen3x3)
    PB_SPEC_EIG("eigen_nxn", ElemOperationEigenNxN)
#undef PB_SPEC_EIG
    if (op == "singular_nx2") {
#define PB_SPEC_SV(FUNNAME, ...) \
      if (fun == FUNNAME) p = new ProxElemOperation<real, ElemOperationSingularNx2<real, __VA_ARGS__>>(idx, count, dim, il, ds, c);
      PB_SPEC_SV("sum_1d:zero", Function2DSum1D<real, Function1DZero<real>>)
      PB_SPEC_SV("sum_1d:abs", Function2DSum1D<real, Function1DAbs<real>>)
      PB_SPEC_SV("sum_1d:square", Function2DSum1D<real, Function1DSquare<real>>)
      PB_SPEC_SV("sum_1d:ind_leq0", Function2DSum1D<real, Function1DIndLeq0<real>>)
      PB_SPEC_SV("sum_1d:ind_box01", Function2DSum1D<real, Function1DIndBox01<real>>)
      PB_SPEC_SV("sum_1d:huber", Function2DSum1D<real, Function1DHuber<real>>)
      PB_SPEC_SV("ind_l1_ball", Function2DIndL1Ball<real>)
      PB_SPEC_SV("moreau:ind_l1_ball", Function2DMoreau<real, Function2DIndL1Ball<real>>)
#undef PB_SPEC_SV
    }
    if (!p) { std::cerr << "unknown spectral operation " << op << " " << fun << std::endl; std::exit(2); }
    return std::shared_ptr<Prox<real>>(p);
  }
  if (kind == "massnorm") {
    std::string op, cost;
    size_t idx, count, dim;
    int il, ds;
    in >> op >> idx >> count >> dim >> il >> ds >> cost;
    std::array<std::vector<real>, 1> c1;
    c1[0] = coeff(cost);
    if (op == "mass4") return std::shared_ptr<Prox<real>>(new ProxElemOperation<real, ElemOperationMass4<real, false>>(idx, count, dim, il, ds, c1));
    if (op == "ind_comass4_ball") return std::shared_ptr<Prox<real>>(new ProxElemOperation<real, ElemOperationMass4<real, true>>(idx, count, dim, il, ds, c1));
    if (op == "mass5") return std::shared_ptr<Prox<real>>(new ProxElemOperation<real, ElemOperationMass5<real, false>>(idx, count, dim, il, ds));
    if (op == "ind_comass5_ball") return std::shared_ptr<Prox<real>>(new ProxElemOperation<real, ElemOperationMass5<real, true>>(idx, count, dim, il, ds));
    std::cerr << "unknown mass norm " << op << std::endl;
    std::exit(2);
  }
  if (kind == "indrange") {
    size_t idx, size;
    int ds, m, n, nnz;
    std::string fv, fp, fi, fa;
    in >> idx >> size >> ds >> m >> n >> nnz >> fv >> fp >> fi >> fa;
    ProxIndRange<real>* p = new ProxIndRange<real>(idx, size, ds != 0);
    const std::vector<float> v32 = read_file<float>(fv, nnz), a32 = read_file<float>(fa, (size_t)n * n);
    p->setA(m, n, nnz, std::vector<real>(v32.begin(), v32.end()), read_file<int32_t>(fp, n + 1), read_file<int32_t>(fi, nnz));
    p->setAA(n, n, std::vector<real>(a32.begin(), a32.end()));
    return std::shared_ptr<Prox<real>>(p);
  }
  if (kind == "indsumidx") {
    size_t idx, size;
    int lists;
    in >> idx >> size >> lists;
    size_t dim[2] = {0, 0}, n[2] = {0, 0};
    double total[2] = {0, 0};
    std::vector<size_t> inds[2];
    for (int l = 0; l < lists && l < 2; ++l) {
      std::string file;
      in >> dim[l] >> file >> n[l] >> total[l];
      const std::vector<unsigned long long> raw = read_file<unsigned long long>(file, n[l]);
      inds[l].assign(raw.begin(), raw.end());
    }
    if (lists == 1)
      return std::shared_ptr<Prox<real>>(
          new ProxIndSum<real>(idx, size, n[0] / dim[0], dim[0], inds[0], static_cast<real>(total[0])));
    return std::shared_ptr<Prox<real>>(new ProxIndSum<real>(idx, size, n[0] / dim[0], dim[0], inds[0],
                                                            static_cast<real>(total[0]), n[1] / dim[1], dim[1], inds[1],
                                                            static_cast<real>(total[1])));
  }
  if (kind == "halfspace") {
    size_t idx, count, dim;
    int il, ds;
    std::string a, b;
    in >> idx >> count >> dim >> il >> ds >> a >> b;
    return std::shared_ptr<Prox<real>>(new ProxIndHalfspace<real>(idx, count, dim, il, ds, coeff(a), coeff(b)));
  }
  if (kind == "soc") {
    size_t idx, count, dim;
    int il, ds;
    double alpha;
    in >> idx >> count >> dim >> il >> ds >> alpha;
    return std::shared_ptr<Prox<real>>(new ProxIndSOC<real>(idx, count, dim, il, ds, static_cast<real>(alpha)));
  }
  if (kind == "indsum") {
    size_t idx, count, dim;
    int il, ds;
    in >> idx >> count >> dim >> il >> ds;
    return std::shared_ptr<Prox<real>>(new ProxElemOperation<real, ElemOperationIndSum<real>>(idx, count, dim, il, ds));
  }
  if (kind == "epiquad") {
    size_t idx, count, dim;
    int il, ds;
    std::string a, b, c;
    in >> idx >> count >> dim >> il >> ds >> a >> b >> c;
    return std::shared_ptr<Prox<real>>(new ProxIndEpiQuad<real>(idx, count, dim, il, ds, coeff(a), coeff(b), coeff(c)));
  }
  if (kind == "moreau") return std::shared_ptr<Prox<real>>(new ProxMoreau<real>(parse_prox(in)));
  if (kind == "transform") {
    std::string a, b, c, d, e;
    in >> a >> b >> c >> d >> e;
    const std::vector<real> va = coeff(a), vb = coeff(b), vc = coeff(c), vd = coeff(d), ve = coeff(e);
    return std::shared_ptr<Prox<real>>(new ProxTransform<real>(parse_prox(in), va, vb, vc, vd, ve));
  }
  if (kind == "permute") {
    std::string file;
    size_t n;
    in >> file >> n;
    std::vector<int32_t> p32 = read_file<int32_t>(file, n);
    std::vector<int> perm(p32.begin(), p32.end());
    return std::shared_ptr<Prox<real>>(new ProxPermute<real>(parse_prox(in), perm));
  }
  if (kind == "zero") {
    size_t idx, size;
    in >> idx >> size;
    return std::shared_ptr<Prox<real>>(new ProxZero<real>(idx, size));
  }
  std::cerr << "unknown prox kind " << kind << std::endl;
  std::exit(2);
}

static std::shared_ptr<Block<real>> parse_block(std::istringstream& in) {
  std::string kind;
  size_t row, col;
  in >> kind >> row >> col;
  if (kind == "gradient2d" || kind == "gradient3d") {
    size_t nx, ny, L;
    int lf;
    in >> nx >> ny >> L >> lf;
    if (kind == "gradient2d") return std::shared_ptr<Block<real>>(new BlockGradient2D<real>(row, col, nx, ny, L, lf != 0));
    return std::shared_ptr<Block<real>>(new BlockGradient3D<real>(row, col, nx, ny, L, lf != 0));
  }
  if (kind == "diags") {
    size_t nrows, ncols, nd;
    std::string fo, ff;
    in >> nrows >> ncols >> nd >> fo >> ff;
    std::vector<int64_t> o64 = read_file<int64_t>(fo, nd);
    std::vector<ssize_t> ofs(o64.begin(), o64.end());
    return std::shared_ptr<Block<real>>(new BlockDiags<real>(row, col, nrows, ncols, nd, ofs, read_real(ff, nd)));
  }
  if (kind == "sparse") {
    int m, n, nnz;
    std::string fv, fp, fi;
    in >> m >> n >> nnz >> fv >> fp >> fi;
    return std::shared_ptr<Block<real>>(BlockSparse<real>::CreateFromCSC(
        row, col, m, n, nnz, read_real(fv, nnz), read_file<int32_t>(fp, n + 1), read_file<int32_t>(fi, nnz)));
  }
  if (kind == "dense") {
    size_t nrows, ncols;
    std::string fd;
    in >> nrows >> ncols >> fd;
    return std::shared_ptr<Block<real>>(
        BlockDense<real>::CreateFromColFirstData(row, col, nrows, ncols, read_real(fd, nrows * ncols)));
  }
  if (kind == "sparse_kron_id" || kind == "id_kron_sparse") {
    size_t diaglength;
    int m, n, nnz;
    std::string fv, fp, fi;
    in >> diaglength >> m >> n >> nnz >> fv >> fp >> fi;
    const std::vector<real> val = read_real(fv, nnz);
    const std::vector<int32_t> ptr = read_file<int32_t>(fp, n + 1), ind = read_file<int32_t>(fi, nnz);
    if (kind == "sparse_kron_id")
      return std::shared_ptr<Block<real>>(BlockSparseKronId<real>::CreateFromCSC(row, col, diaglength, m, n, nnz, val, ptr, ind));
    return std::shared_ptr<Block<real>>(BlockIdKronSparse<real>::CreateFromCSC(row, col, diaglength, m, n, nnz, val, ptr, ind));
  }
  if (kind == "dense_kron_id" || kind == "id_kron_dense") {
    size_t nrows, ncols, diaglength;
    std::string fd;
    in >> nrows >> ncols >> diaglength >> fd;
    const std::vector<real> data = read_real(fd, nrows * ncols);
    if (kind == "dense_kron_id")
      return std::shared_ptr<Block<real>>(
          BlockDenseKronId<real>::CreateFromColFirstData(diaglength, row, col, nrows, ncols, data));
    return std::shared_ptr<Block<real>>(
        BlockIdKronDense<real>::CreateFromColFirstData(diaglength, row, col, nrows, ncols, data));
  }
  if (kind == "zero") {
    size_t nrows, ncols;
    in >> nrows >> ncols;
    return std::shared_ptr<Block<real>>(new BlockZero<real>(row, col, nrows, ncols));
  }
  std::cerr << "unknown block kind " << kind << std::endl;
  std::exit(2);
}

int main(int argc, char** argv) {
  if (argc < 2) { std::cerr << "usage: " << argv[0] << " <description.txt>" << std::endl; return 2; }
  const std::string spec = argv[1];
  const size_t slash = spec.rfind('/');
  g_dir = slash == std::string::npos ? "." : spec.substr(0, slash);
  std::ifstream fin(spec);
  if (!fin) { std::cerr << "cannot open " << spec << std::endl; return 2; }

  try {
    BlockDiags<real>::ResetConstMem();      // the mex does this before every solve (prost.cpp:74)
    std::shared_ptr<Problem<real>> problem(new Problem<real>());
    std::shared_ptr<LinearOperator<real>> linop(new LinearOperator<real>());
    std::shared_ptr<Prox<real>> eval_prox;
    problem->SetScalingAlpha(1);
    size_t nrows = 0, ncols = 0;
    BackendPDHG<real>::Options po;
    po.tau0 = 1; po.sigma0 = 1; po.residual_iter = 1; po.scale_steps_operator = false; po.alg2_gamma = 0;
    po.arg_alpha0 = 0.5f; po.arg_nu = 0.95f; po.arg_delta = 1.5f; po.arb_delta = 1.05f; po.arb_tau = 0.8f;
    po.stepsize_variant = BackendPDHG<real>::kPDHGStepsResidualBoyd;
    BackendADMM<real>::Options ao;                        // matlab/+prost/+backend/admm.m:3-13
    ao.rho0 = 1; ao.alpha = 1.7; ao.cg_tol_pow = 1.3; ao.cg_tol_min = 1e-5; ao.cg_tol_max = 1e-8; ao.cg_max_iter = 10;
    ao.residual_iter = 1; ao.arb_delta = 1.05f; ao.arb_tau = 0.8f; ao.arb_gamma = 1.01f;
    bool use_admm = false;
    Solver<real>::Options so;
    so.tol_rel_primal = so.tol_rel_dual = so.tol_abs_primal = so.tol_abs_dual = 1e-4f;
    so.max_iters = 100; so.num_cback_calls = 0; so.verbose = false; so.solve_dual_problem = false;
    std::string action = "solve", out = "out", a1, a2, a3;

    std::string line;
    while (std::getline(fin, line)) {
      std::istringstream in(line);
      std::string key;
      if (!(in >> key) || key[0] == '#') continue;
      if (key == "dims") { in >> nrows >> ncols; problem->SetDimensions(nrows, ncols); }
      else if (key == "block") { auto b = parse_block(in); problem->AddBlock(b); linop->AddBlock(b); }
      else if (key == "prox") {
        std::string which;
        in >> which;
        auto p = parse_prox(in);
        if (which == "g") problem->AddProx_g(p);
        else if (which == "f") problem->AddProx_f(p);
        else if (which == "gstar") problem->AddProx_gstar(p);
        else if (which == "fstar") problem->AddProx_fstar(p);
        else eval_prox = p;
      }
      else if (key == "scaling") {
        std::string kind;
        in >> kind;
        if (kind == "alpha") { real a; in >> a; problem->SetScalingAlpha(a); }
        else if (kind == "identity") problem->SetScalingIdentity();
        else { std::string fl, fr; in >> fl >> fr; problem->SetScalingCustom(read_real(fl, nrows), read_real(fr, ncols)); }
      }
      else if (key == "pdhg") {
        int sso, variant;
        in >> po.tau0 >> po.sigma0 >> po.residual_iter >> sso >> po.alg2_gamma >> po.arg_alpha0 >> po.arg_nu >>
            po.arg_delta >> po.arb_delta >> po.arb_tau >> variant;
        po.scale_steps_operator = sso != 0;
        po.stepsize_variant = static_cast<BackendPDHG<real>::StepsizeVariant>(variant);
      }
      else if (key == "admm") {
        in >> ao.rho0 >> ao.alpha >> ao.cg_tol_pow >> ao.cg_tol_min >> ao.cg_tol_max >> ao.cg_max_iter >>
            ao.residual_iter >> ao.arb_delta >> ao.arb_tau >> ao.arb_gamma;
        use_admm = true;
      }
      else if (key == "solver") {
        std::string fx, fy;
        int dual;
        in >> so.tol_rel_primal >> so.tol_rel_dual >> so.tol_abs_primal >> so.tol_abs_dual >> so.max_iters >>
            so.num_cback_calls >> fx >> fy >> dual;
        so.solve_dual_problem = dual != 0;
        if (fx != "-") so.x0 = read_real(fx, ncols);
        if (fy != "-") so.y0 = read_real(fy, nrows);
      }
      else if (key == "action") { in >> action >> a1 >> a2 >> a3; }
      else if (key == "out") { in >> out; }
      else { std::cerr << "unknown statement " << key << std::endl; return 2; }
    }

    std::ofstream info(out + "_info.txt");
    if (action == "linop") {
      linop->Initialize();
      const bool transpose = std::atoi(a2.c_str()) != 0;
      std::vector<real> rhs = read_real(a1, transpose ? linop->nrows() : linop->ncols());
      std::vector<real> res;
      const double ms = transpose ? linop->EvalAdjoint(res, rhs) : linop->Eval(res, rhs);
      write_file(out + "_res.f32", res);
      std::vector<real> rs(linop->nrows()), cs(linop->ncols());
      for (size_t r = 0; r < rs.size(); ++r) rs[r] = linop->row_sum(r, 1);
      for (size_t c = 0; c < cs.size(); ++c) cs[c] = linop->col_sum(c, 1);
      write_file(out + "_rowsum.f32", rs);
      write_file(out + "_colsum.f32", cs);
      info << "nrows " << linop->nrows() << "\nncols " << linop->ncols() << "\nms " << ms << "\n";
    } else if (action == "prox") {
      eval_prox->Initialize();
      const size_t n = eval_prox->index() + eval_prox->size();
      std::ifstream probe(g_dir + "/" + a1, std::ios::binary | std::ios::ate);
      const size_t len = static_cast<size_t>(probe.tellg()) / sizeof(float);
      std::vector<real> arg = read_real(a1, len), td = read_real(a2, len), res;
      const double ms = eval_prox->Eval(res, arg, td, static_cast<real>(std::atof(a3.c_str())));
      write_file(out + "_res.f32", res);
      info << "n " << n << "\nms " << ms << "\n";
    } else {
      std::shared_ptr<Backend<real>> backend;
      if (use_admm) backend = std::shared_ptr<Backend<real>>(new BackendADMM<real>(ao));
      else backend = std::shared_ptr<Backend<real>>(new BackendPDHG<real>(po));
      Solver<real> solver(problem, backend);
      solver.SetOptions(so);
      int iterations = 0;                // the stopping callback runs once per iteration (solver.cu:147)
      solver.SetStoppingCallback([&iterations]() { ++iterations; return false; });
      int cbacks = 0;
      solver.SetIntermCallback([&cbacks](int, const std::vector<real>&, const std::vector<real>&) { ++cbacks; return false; });
      solver.Initialize();
      const auto t0 = std::chrono::steady_clock::now();
      const int result = static_cast<int>(solver.Solve());
      const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
      write_file(out + "_x.f32", solver.cur_primal_sol());
      write_file(out + "_y.f32", solver.cur_dual_sol());
      write_file(out + "_z.f32", solver.cur_primal_constr_sol());
      write_file(out + "_w.f32", solver.cur_dual_constr_sol());
      info.precision(9);
      info << "result " << result << "\nprimal_residual " << backend->primal_residual() << "\ndual_residual "
           << backend->dual_residual() << "\nprimal_var_norm " << backend->primal_var_norm() << "\ndual_var_norm "
           << backend->dual_var_norm() << "\neps_primal " << backend->eps_primal() << "\neps_dual "
           << backend->eps_dual() << "\nsolve_ms " << ms << "\ncallbacks " << cbacks << "\niterations " << iterations << "\n";
      solver.Release();
    }
  } catch (const Exception& e) {
    std::cerr << "prost::Exception: " << e.what() << std::endl;
    return 3;
  }
  return 0;
}
