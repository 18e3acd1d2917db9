// step-64 on the engine, from deal.II host code: the Helmholtz problem of examples/step-64/step-64.cc
// (variable coefficient 10 / (0.05 + 2 |x|^2), zero Dirichlet boundary, FE_Q(fe_degree), one refinement
// ball so that the mesh has hanging nodes in the later cycles), set up with deal.II's own Triangulation /
// DoFHandler / AffineConstraints, solved twice with the UNMODIFIED deal.II SolverCG + Jacobi:
//   (1) reference: CPU MatrixFree operator (FEEvaluation in MatrixFree::cell_loop) on host vectors,
//   (2) engine:    HelmholtzOperator::vmult = dst = 0; cell_loop; copy_constrained_values
//                  (step-64.cc:313-325) through b200::dealii_adapter::MatrixFree -> libb200mf.so on
//                  device vectors.
// It prints, per cycle, both iteration counts and solution norms and exits non-zero when the iteration
// counts differ by more than 1 or the solutions by more than 1e-9 (relative, l2).
//
// Built by oracle/ref_drivers/build.sh where the reference's headers exist (needs deal.II + CUDA
// runtime headers); run on the GPU box by tests/test_dealii_adapter_gpu.py.
#include <deal.II/base/function.h>
#include <deal.II/base/quadrature_lib.h>

#include <deal.II/dofs/dof_handler.h>
#include <deal.II/dofs/dof_tools.h>

#include <deal.II/fe/fe_q.h>
#include <deal.II/fe/mapping_q1.h>

#include <deal.II/grid/grid_generator.h>
#include <deal.II/grid/tria.h>

#include <deal.II/lac/affine_constraints.h>
#include <deal.II/lac/diagonal_matrix.h>
#include <deal.II/lac/la_parallel_vector.h>
#include <deal.II/lac/precondition.h>
#include <deal.II/lac/solver_cg.h>

#include <deal.II/matrix_free/fe_evaluation.h>
#include <deal.II/matrix_free/matrix_free.h>
#include <deal.II/matrix_free/tools.h>

#include <deal.II/numerics/vector_tools.h>

#include <cstdio>

#include "b200mf_dealii.hpp"

using namespace dealii;

template <int dim>
double coefficient(const Point<dim> &p) { return 10. / (0.05 + 2. * p.square()); } // step-64.cc:92-105

// ---- (1) the reference operator on the host
template <int dim, int fe_degree>
class CpuHelmholtz {
public:
  using VectorType = LinearAlgebra::distributed::Vector<double, MemorySpace::Host>;
  CpuHelmholtz(const Mapping<dim> &mapping, const DoFHandler<dim> &dof, const AffineConstraints<double> &constraints) {
    typename MatrixFree<dim, double>::AdditionalData ad;
    ad.tasks_parallel_scheme = MatrixFree<dim, double>::AdditionalData::none;
    ad.mapping_update_flags = update_values | update_gradients | update_JxW_values | update_quadrature_points;
    mf.reinit(mapping, dof, constraints, QGauss<1>(fe_degree + 1), ad);
    FEEvaluation<dim, fe_degree> phi(mf);
    coef.reinit(mf.n_cell_batches(), phi.n_q_points);
    for (unsigned int cell = 0; cell < mf.n_cell_batches(); ++cell) {
      phi.reinit(cell);
      for (unsigned int q = 0; q < phi.n_q_points; ++q) {
        const auto pq = phi.quadrature_point(q);
        VectorizedArray<double> c = 0.;
        for (unsigned int v = 0; v < VectorizedArray<double>::size(); ++v) {
          Point<dim> pt;
          for (unsigned int d = 0; d < dim; ++d) pt[d] = pq[d][v];
          c[v] = coefficient(pt);
        }
        coef(cell, q) = c;
      }
    }
  }
  void initialize_dof_vector(VectorType &v) const { mf.initialize_dof_vector(v); }
  void vmult(VectorType &dst, const VectorType &src) const {
    mf.cell_loop(&CpuHelmholtz::local_apply, this, dst, src, true);
    for (const auto i : mf.get_constrained_dofs()) dst.local_element(i) = src.local_element(i);
  }
  void inverse_diagonal(VectorType &inv) const {
    mf.initialize_dof_vector(inv);
    MatrixFreeTools::compute_diagonal<dim, fe_degree, fe_degree + 1, 1, double, VectorizedArray<double>>(
      mf, inv, [this](auto &phi) {
        const unsigned int cell = phi.get_current_cell_index();
        phi.evaluate(EvaluationFlags::values | EvaluationFlags::gradients);
        for (unsigned int q = 0; q < phi.n_q_points; ++q) {
          phi.submit_value(coef(cell, q) * phi.get_value(q), q);
          phi.submit_gradient(phi.get_gradient(q), q);
        }
        phi.integrate(EvaluationFlags::values | EvaluationFlags::gradients);
      });
    for (auto &v : inv) v = (v != 0.) ? 1. / v : 1.;
  }

private:
  void local_apply(const MatrixFree<dim, double> &data, VectorType &dst, const VectorType &src,
                   const std::pair<unsigned int, unsigned int> &range) const {
    FEEvaluation<dim, fe_degree> phi(data);
    for (unsigned int cell = range.first; cell < range.second; ++cell) {
      phi.reinit(cell);
      phi.read_dof_values(src);
      phi.evaluate(EvaluationFlags::values | EvaluationFlags::gradients);
      for (unsigned int q = 0; q < phi.n_q_points; ++q) {
        phi.submit_value(coef(cell, q) * phi.get_value(q), q);
        phi.submit_gradient(phi.get_gradient(q), q);
      }
      phi.integrate(EvaluationFlags::values | EvaluationFlags::gradients);
      phi.distribute_local_to_global(dst);
    }
  }
  MatrixFree<dim, double> mf;
  Table<2, VectorizedArray<double>> coef;
};

// ---- (2) step-64's HelmholtzOperator on the engine (step-64.cc:225-370)
template <int dim, int fe_degree>
class HelmholtzOperator {
public:
  using VectorType = b200::dealii_adapter::Vector<double>;
  HelmholtzOperator(const Mapping<dim> &mapping, const DoFHandler<dim> &dof, const AffineConstraints<double> &constraints) {
    mf_data.reinit(mapping, dof, constraints, QGauss<1>(fe_degree + 1));
    mf_data.evaluate_coefficients([](const Point<dim> &p) { return coefficient(p); }, coef);
    op = std::make_unique<b200::Operator<dim, double>>(mf_data, nullptr, coef.get_values(), 1.0, 0.0);
  }
  void initialize_dof_vector(VectorType &v) const { mf_data.initialize_dof_vector(v); }
  void vmult(VectorType &dst, const VectorType &src) const { op->vmult(dst.get_values(), src.get_values()); }
  void compute_diagonal() {
    mf_data.initialize_dof_vector(inverse_diagonal.get_vector());
    op->compute_diagonal(inverse_diagonal.get_vector().get_values());
    // 1 / diag on the host: setup cost only
    Vector<double> h;
    inverse_diagonal.get_vector().export_to_host(h);
    for (auto &v : h) v = 1. / v;
    inverse_diagonal.get_vector().import_from_host(h);
  }
  types::global_dof_index m() const { return mf_data.get_vector_partitioner()->size(); }
  b200::dealii_adapter::DiagonalPreconditioner<double> inverse_diagonal;

private:
  b200::dealii_adapter::MatrixFree<dim, double> mf_data;
  b200::dealii_adapter::Vector<double> coef;
  std::unique_ptr<b200::Operator<dim, double>> op;
};

template <int dim, int fe_degree>
int run_cycle(unsigned int cycle, bool hanging) {
  Triangulation<dim> tria;
  GridGenerator::hyper_cube(tria, 0., 1.);
  tria.refine_global(2 + cycle);
  if (hanging) {
    for (const auto &cell : tria.active_cell_iterators())
      if (cell->center().distance(Point<dim>(0.5, 0.5, 0.5)) < 0.3) cell->set_refine_flag();
    tria.execute_coarsening_and_refinement();
  }
  const FE_Q<dim> fe(fe_degree);
  const MappingQ1<dim> mapping;
  DoFHandler<dim> dof(tria);
  dof.distribute_dofs(fe);
  AffineConstraints<double> constraints;
  DoFTools::make_hanging_node_constraints(dof, constraints);
  VectorTools::interpolate_boundary_values(mapping, dof, 0, Functions::ZeroFunction<dim>(), constraints);
  constraints.close();
  const unsigned int n = dof.n_dofs();

  // rhs = 1 tested against the basis, as step-64 assembles it (step-64.cc:520-560), on the host
  Vector<double> rhs(n);
  {
    const QGauss<dim> quad(fe_degree + 1);
    FEValues<dim> fev(mapping, fe, quad, update_values | update_JxW_values);
    Vector<double> cell_rhs(fe.n_dofs_per_cell());
    std::vector<types::global_dof_index> idx(fe.n_dofs_per_cell());
    for (const auto &cell : dof.active_cell_iterators()) {
      cell_rhs = 0;
      fev.reinit(cell);
      for (unsigned int q = 0; q < quad.size(); ++q)
        for (unsigned int i = 0; i < fe.n_dofs_per_cell(); ++i) cell_rhs(i) += fev.shape_value(i, q) * fev.JxW(q);
      cell->get_dof_indices(idx);
      constraints.distribute_local_to_global(cell_rhs, idx, rhs);
    }
  }
  const double tol = 1e-12 * rhs.l2_norm(); // step-64.cc:600

  // (1) reference
  unsigned int it_cpu = 0;
  double norm_cpu = 0;
  Vector<double> sol_cpu(n);
  {
    CpuHelmholtz<dim, fe_degree> A(mapping, dof, constraints);
    typename CpuHelmholtz<dim, fe_degree>::VectorType x, b;
    A.initialize_dof_vector(x);
    A.initialize_dof_vector(b);
    for (unsigned int i = 0; i < n; ++i) b.local_element(i) = rhs(i);
    DiagonalMatrix<typename CpuHelmholtz<dim, fe_degree>::VectorType> jacobi;
    A.inverse_diagonal(jacobi.get_vector());
    SolverControl control(n, tol);
    SolverCG<typename CpuHelmholtz<dim, fe_degree>::VectorType> cg(control);
    cg.solve(A, x, b, jacobi);
    it_cpu = control.last_step();
    for (unsigned int i = 0; i < n; ++i) sol_cpu(i) = x.local_element(i);
    constraints.distribute(sol_cpu);
    norm_cpu = sol_cpu.l2_norm();
  }
  // (2) engine
  unsigned int it_gpu = 0;
  double norm_gpu = 0, diff = 0;
  {
    HelmholtzOperator<dim, fe_degree> A(mapping, dof, constraints);
    A.compute_diagonal();
    typename HelmholtzOperator<dim, fe_degree>::VectorType x, b;
    A.initialize_dof_vector(x);
    b.import_from_host(rhs);
    SolverControl control(n, tol);
    SolverCG<typename HelmholtzOperator<dim, fe_degree>::VectorType> cg(control);
    cg.solve(A, x, b, A.inverse_diagonal);
    it_gpu = control.last_step();
    Vector<double> sol;
    x.export_to_host(sol);
    constraints.distribute(sol);
    norm_gpu = sol.l2_norm();
    sol -= sol_cpu;
    diff = sol.l2_norm();
  }
  std::printf("cycle %u%s: %u cells, %u DoFs | deal.II CPU MatrixFree: %u CG iterations, |u| = %.10g | "
              "engine via adapter: %u CG iterations, |u| = %.10g | |u_engine - u_cpu| / |u_cpu| = %.3e\n",
              cycle, hanging ? " (hanging nodes)" : "", tria.n_active_cells(), n, it_cpu, norm_cpu, it_gpu, norm_gpu,
              diff / norm_cpu);
  const bool ok = (it_cpu > it_gpu ? it_cpu - it_gpu : it_gpu - it_cpu) <= 1 && diff <= 1e-9 * norm_cpu;
  return ok ? 0 : 1;
}

int main() {
  int rc = 0;
  try {
    for (unsigned int cycle = 0; cycle < 3; ++cycle) rc |= run_cycle<3, 3>(cycle, false);   // shipped degree
    rc |= run_cycle<3, 5>(0, false);                                                        // BASELINE configs[2]: Q5
    rc |= run_cycle<3, 3>(1, true);                                                         // with hanging nodes
  } catch (const std::exception &e) {
    std::fprintf(stderr, "exception: %s\n", e.what());
    return 2;
  }
  std::printf(rc == 0 ? "step-64 through libb200mf.so: OK\n" : "step-64 through libb200mf.so: MISMATCH\n");
  return rc;
}
