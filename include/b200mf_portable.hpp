// b200mf_portable.hpp -- header-only C++ shim that gives the C ABI of libb200mf.so the shape of
// deal.II's Portable::MatrixFree / operator / SolverCG surface for the hot path, so that code
// written against
//     Portable::MatrixFree<dim,Number>        (matrix_free/portable_matrix_free.h:185)
//     HelmholtzOperator / LaplaceOperator     (examples/step-64/step-64.cc:225-370,
//                                              tests/performance/timing_matrix_free_kokkos.cc:129-152)
//     SolverCG + DiagonalMatrix / PreconditionChebyshev   (lac/solver_cg.h, lac/precondition.h)
// can switch by changing a namespace.  It contains no deal.II code and needs no deal.II headers:
// a deal.II-side adapter fills b200::ReinitData from DoFHandler / AffineConstraints / Mapping
// exactly where Portable::MatrixFree::internal_reinit builds its per-cell arrays
// (portable_matrix_free.templates.h:267-346, 1366-1425); see INTEGRATION.md.
//
// Vectors are raw device pointers of length n_owned + n_ghost in LA::d::Vector layout
// (LinearAlgebra::distributed::Vector<Number, MemorySpace::Default>::get_values()).
#ifndef B200MF_PORTABLE_HPP
#define B200MF_PORTABLE_HPP

#include <cstdint>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "b200mf.h"

namespace b200 {

// AssertThrow of the reference becomes an exception carrying b200mf_last_error()
struct Exception : std::runtime_error {
  int code;
  Exception(int c, const std::string &what) : std::runtime_error(what), code(c) {}
};
struct NoConvergence : Exception { // SolverControl::NoConvergence (lac/solver_control.h)
  int last_step;
  double last_residual;
  NoConvergence(int steps, double res)
      : Exception(B200MF_ERR_NOCONVERGENCE, "iterative method did not converge"), last_step(steps),
        last_residual(res) {}
};
inline void check(int code) {
  if (code != B200MF_OK) throw Exception(code, b200mf_last_error());
}

template <typename Number>
constexpr int number_code() {
  static_assert(std::is_same<Number, double>::value || std::is_same<Number, float>::value,
                "Number must be double or float");
  return std::is_same<Number, double>::value ? B200MF_F64 : B200MF_F32;
}

// The arrays Portable::MatrixFree::reinit extracts from (mapping, dof_handler, constraints, quad).
struct ReinitData {
  int degree = 1;
  int n_q_points_1d = 0; // 0 => degree + 1
  std::uint64_t n_cells = 0, n_owned_dofs = 0, n_ghost_dofs = 0, n_cells_interior = 0;
  const std::uint32_t *local_to_global = nullptr; // [n_cells][(p+1)^dim], lexicographic, local
  const std::uint16_t *constraint_mask = nullptr; // ConstraintKinds per cell or null
  const double *cell_vertices = nullptr;          // MappingQ1: [n_cells][2^dim][dim]
  const double *inv_jacobian = nullptr, *JxW = nullptr; // or the arrays PMF stores itself
  const std::uint32_t *constrained_dofs = nullptr;
  std::uint64_t n_constrained_dofs = 0;
};

template <int dim, typename Number>
class MatrixFree {
public:
  struct AdditionalData { // Portable::MatrixFree::AdditionalData (portable_matrix_free.h:209-268)
    bool use_coloring = false;                        // the engine always scatters with atomics
    bool overlap_communication_computation = true;    // honoured by the distributed driver
  };

  MatrixFree() = default;
  MatrixFree(const MatrixFree &) = delete;
  MatrixFree &operator=(const MatrixFree &) = delete;
  ~MatrixFree() { clear(); }

  void reinit(const ReinitData &d, const AdditionalData & = AdditionalData()) {
    clear();
    b200mf_setup_desc s{};
    s.dim = dim; s.degree = d.degree; s.n_q_points_1d = d.n_q_points_1d > 0 ? d.n_q_points_1d : d.degree + 1;
    s.number = number_code<Number>();
    s.n_cells = d.n_cells; s.n_owned_dofs = d.n_owned_dofs; s.n_ghost_dofs = d.n_ghost_dofs;
    s.local_to_global = d.local_to_global; s.constraint_mask = d.constraint_mask;
    if (d.cell_vertices) { s.geometry = B200MF_GEOMETRY_Q1_VERTICES; s.cell_vertices = d.cell_vertices; }
    else { s.geometry = B200MF_GEOMETRY_JACOBIANS; s.inv_jacobian = d.inv_jacobian; s.JxW = d.JxW; }
    s.constrained_dofs = d.constrained_dofs; s.n_constrained_dofs = d.n_constrained_dofs;
    s.n_cells_interior = d.n_cells_interior;
    check(b200mf_setup_create(&s, &setup_));
    n_local_ = d.n_owned_dofs + d.n_ghost_dofs; n_owned_ = d.n_owned_dofs;
  }
  void clear() { if (setup_) { b200mf_setup_destroy(setup_); setup_ = nullptr; } }

  // cell_loop(func, src, dst) with the recognised functor family (see b200mf_operator)
  void cell_loop(const b200mf_operator &op, const Number *src, Number *dst, void *stream = nullptr) const {
    check(b200mf_cell_loop(setup_, &op, dst, src, stream));
  }
  void copy_constrained_values(const Number *src, Number *dst, void *stream = nullptr) const {
    check(b200mf_copy_constrained_values(setup_, dst, src, stream));
  }
  void set_constrained_values(Number value, Number *dst, void *stream = nullptr) const {
    check(b200mf_set_constrained_values(setup_, dst, double(value), stream));
  }
  std::uint64_t n_local_dofs() const { return n_local_; }   // size initialize_dof_vector() gives
  std::uint64_t locally_owned_size() const { return n_owned_; }
  const b200mf_setup *get_setup() const { return setup_; }

private:
  b200mf_setup *setup_ = nullptr;
  std::uint64_t n_local_ = 0, n_owned_ = 0;
};

// (c_grad grad u, grad v) + (c_mass u, v): LaplaceOperator / HelmholtzOperator::vmult
template <int dim, typename Number>
class Operator {
public:
  Operator(const MatrixFree<dim, Number> &mf, const Number *grad_coefficient = nullptr,
           const Number *mass_coefficient = nullptr, double grad_constant = 1.0, double mass_constant = 0.0)
      : mf_(mf), op_{grad_coefficient, mass_coefficient, grad_constant, mass_constant} {}
  void vmult(Number *dst, const Number *src, void *stream = nullptr) const {
    check(b200mf_vmult(mf_.get_setup(), &op_, dst, src, stream));
  }
  void compute_diagonal(Number *diag, void *stream = nullptr) const { // MatrixFreeTools::compute_diagonal
    check(b200mf_compute_diagonal(mf_.get_setup(), &op_, diag, stream));
  }
  std::uint64_t m() const { return mf_.locally_owned_size(); }
  const b200mf_operator &functor() const { return op_; }
  const MatrixFree<dim, Number> &matrix_free() const { return mf_; }

private:
  const MatrixFree<dim, Number> &mf_;
  b200mf_operator op_;
};

struct SolverControl { // lac/solver_control.h
  int max_steps = 100;
  double tolerance = 1e-10;
  int last_step_ = 0;
  double last_value_ = 0.0;
  SolverControl(int n = 100, double tol = 1e-10) : max_steps(n), tolerance(tol) {}
  int last_step() const { return last_step_; }
  double last_value() const { return last_value_; }
};

struct PreconditionChebyshevData { // PreconditionChebyshev::AdditionalData (lac/precondition.h:2121-2175)
  int degree = 1;
  double smoothing_range = 0.0;
  int eig_cg_n_iterations = 8;
  double safety_factor = 1.2;
};

template <int dim, typename Number>
class SolverCG { // SolverCG<VectorType>::solve(A, x, b, preconditioner) (lac/solver_cg.h:1391)
public:
  explicit SolverCG(SolverControl &c) : control_(c) {}
  // inverse_diagonal == nullptr: PreconditionIdentity; chebyshev == nullptr: Jacobi
  void solve(const Operator<dim, Number> &A, Number *x, const Number *b, const Number *inverse_diagonal,
             const PreconditionChebyshevData *chebyshev = nullptr, void *stream = nullptr) {
    b200mf_solver_desc d{};
    d.preconditioner = inverse_diagonal == nullptr ? B200MF_PRECOND_NONE
                       : (chebyshev ? B200MF_PRECOND_CHEBYSHEV : B200MF_PRECOND_JACOBI);
    d.inverse_diagonal = inverse_diagonal;
    if (chebyshev) {
      d.chebyshev_degree = chebyshev->degree; d.smoothing_range = chebyshev->smoothing_range;
      d.eig_cg_n_iterations = chebyshev->eig_cg_n_iterations; d.safety_factor = chebyshev->safety_factor;
    }
    d.tolerance = control_.tolerance; d.max_iterations = control_.max_steps;
    b200mf_solver_result r{};
    const int rc = b200mf_cg_solve(A.matrix_free().get_setup(), &A.functor(), &d, x, b, &r, stream);
    control_.last_step_ = r.iterations; control_.last_value_ = r.residual;
    if (rc == B200MF_ERR_NOCONVERGENCE) throw NoConvergence(r.iterations, r.residual);
    check(rc);
  }

private:
  SolverControl &control_;
};

} // namespace b200
#endif
