// step-37 on the engine, from deal.II host code: -div(a grad u) = 1 with a = 1 / (0.05 + 2 |x|^2), zero
// Dirichlet boundary, FE_Q(degree) on a globally refined hyper_cube, solved twice with the UNMODIFIED
// deal.II SolverCG (double):
//   (1) reference: MatrixFreeOperators::LaplaceOperator on the host, preconditioned by deal.II's own
//       PreconditionMG = Multigrid + MGTransferMatrixFree + PreconditionChebyshev smoothers on float levels
//       (examples/step-37/step-37.cc:950-1060);
//   (2) engine: the system operator through b200::dealii_adapter::MatrixFree (CPU-MatrixFree treatment of
//       constrained dofs), preconditioned by b200::dealii_adapter::PreconditionMG<dim, float>, which builds
//       the same hierarchy from the same DoFHandler (distribute_mg_dofs, MGConstrainedDoFs) and runs the
//       V-cycle on the device.
// It prints both iteration counts and solution norms per refinement and exits non-zero when the iteration
// counts differ by more than 1 or the solutions by more than 1e-8 (relative, l2).
//
// Built by oracle/ref_drivers/build.sh where the reference's headers exist; run on the GPU box by
// tests/test_dealii_adapter_gpu.py.
#include <deal.II/base/quadrature_lib.h>

#include <deal.II/dofs/dof_handler.h>
#include <deal.II/dofs/dof_tools.h>

#include <deal.II/fe/fe_q.h>
#include <deal.II/fe/mapping_q1.h>

#include <deal.II/grid/grid_generator.h>
#include <deal.II/grid/tria.h>

#include <deal.II/lac/affine_constraints.h>
#include <deal.II/lac/la_parallel_vector.h>
#include <deal.II/lac/precondition.h>
#include <deal.II/lac/solver_cg.h>

#include <deal.II/matrix_free/fe_evaluation.h>
#include <deal.II/matrix_free/matrix_free.h>
#include <deal.II/matrix_free/operators.h>

#include <deal.II/multigrid/mg_coarse.h>
#include <deal.II/multigrid/mg_constrained_dofs.h>
#include <deal.II/multigrid/mg_matrix.h>
#include <deal.II/multigrid/mg_smoother.h>
#include <deal.II/multigrid/mg_tools.h>
#include <deal.II/multigrid/mg_transfer_matrix_free.h>
#include <deal.II/multigrid/multigrid.h>

#include <deal.II/numerics/vector_tools.h>

#include <cstdio>

#include "b200mf_dealii.hpp"

using namespace dealii;

template <int dim>
double coefficient(const Point<dim> &p) { return 1. / (0.05 + 2. * p.square()); } // step-37.cc:100-125

template <int dim, int degree, typename Number>
std::shared_ptr<Table<2, VectorizedArray<Number>>> coefficient_table(const MatrixFree<dim, Number> &mf) {
  FEEvaluation<dim, degree, degree + 1, 1, Number> phi(mf);
  auto table = std::make_shared<Table<2, VectorizedArray<Number>>>(mf.n_cell_batches(), phi.n_q_points);
  for (unsigned int cell = 0; cell < mf.n_cell_batches(); ++cell) {
    phi.reinit(cell);
    for (unsigned int q = 0; q < phi.n_q_points; ++q) {
      const auto xq = phi.quadrature_point(q);
      VectorizedArray<Number> a;
      for (unsigned int v = 0; v < VectorizedArray<Number>::size(); ++v) {
        Point<dim> p;
        for (unsigned int d = 0; d < dim; ++d) p[d] = xq[d][v];
        a[v] = coefficient(p);
      }
      (*table)(cell, q) = a;
    }
  }
  return table;
}

// the system operator on the engine: step-37's LaplaceOperator::vmult = MatrixFreeOperators::Base::vmult
template <int dim>
class EngineLaplace {
public:
  using VectorType = b200::dealii_adapter::Vector<double>;
  EngineLaplace(const Mapping<dim> &mapping, const DoFHandler<dim> &dof, const AffineConstraints<double> &constraints,
                unsigned int degree) {
    typename b200::dealii_adapter::MatrixFree<dim, double>::AdditionalData ad;
    ad.eliminate_constrained_dofs = true;
    mf.reinit(mapping, dof, constraints, QGauss<1>(degree + 1), ad);
    mf.evaluate_coefficients([](const Point<dim> &p) { return coefficient(p); }, coef);
    op = std::make_unique<b200::Operator<dim, double>>(mf, coef.get_values(), nullptr, 1.0, 0.0);
  }
  void initialize_dof_vector(VectorType &v) const { mf.initialize_dof_vector(v); }
  void vmult(VectorType &dst, const VectorType &src) const { op->vmult(dst.get_values(), src.get_values()); }

private:
  b200::dealii_adapter::MatrixFree<dim, double> mf;
  b200::dealii_adapter::Vector<double> coef;
  std::unique_ptr<b200::Operator<dim, double>> op;
};

template <int dim, int degree>
int run(unsigned int refinements) {
  using SystemVector = LinearAlgebra::distributed::Vector<double>;
  using LevelVector = LinearAlgebra::distributed::Vector<float>;
  using SystemMatrix = MatrixFreeOperators::LaplaceOperator<dim, degree, degree + 1, 1, SystemVector>;
  using LevelMatrix = MatrixFreeOperators::LaplaceOperator<dim, degree, degree + 1, 1, LevelVector>;

  Triangulation<dim> tria(Triangulation<dim>::limit_level_difference_at_vertices);
  GridGenerator::hyper_cube(tria, 0., 1.);
  tria.refine_global(refinements);
  const FE_Q<dim> fe(degree);
  const MappingQ1<dim> mapping;
  DoFHandler<dim> dof(tria);
  dof.distribute_dofs(fe);
  dof.distribute_mg_dofs();
  const unsigned int n = dof.n_dofs(), n_levels = tria.n_global_levels();
  AffineConstraints<double> constraints;
  VectorTools::interpolate_boundary_values(mapping, dof, 0, Functions::ZeroFunction<dim>(), constraints);
  constraints.close();
  MGConstrainedDoFs mg_constrained_dofs;
  mg_constrained_dofs.initialize(dof);
  mg_constrained_dofs.make_zero_boundary_constraints(dof, {0});

  // ---- (1) the reference
  SystemMatrix system_matrix;
  {
    typename MatrixFree<dim, double>::AdditionalData data;
    data.tasks_parallel_scheme = MatrixFree<dim, double>::AdditionalData::none;
    data.mapping_update_flags = update_gradients | update_JxW_values | update_quadrature_points;
    auto mf = std::make_shared<MatrixFree<dim, double>>();
    mf->reinit(mapping, dof, constraints, QGauss<1>(degree + 1), data);
    system_matrix.initialize(mf);
    system_matrix.set_coefficient(coefficient_table<dim, degree, double>(*mf));
  }
  SystemVector rhs, x_ref;
  system_matrix.initialize_dof_vector(rhs);
  system_matrix.initialize_dof_vector(x_ref);
  {
    const auto &mf = *system_matrix.get_matrix_free();
    FEEvaluation<dim, degree, degree + 1, 1, double> phi(mf);
    for (unsigned int cell = 0; cell < mf.n_cell_batches(); ++cell) {
      phi.reinit(cell);
      for (unsigned int q = 0; q < phi.n_q_points; ++q) phi.submit_value(make_vectorized_array<double>(1.0), q);
      phi.integrate(EvaluationFlags::values);
      phi.distribute_local_to_global(rhs);
    }
    rhs.compress(VectorOperation::add);
  }
  const double tol = 1e-12 * rhs.l2_norm(); // step-37.cc:1066
  unsigned int it_ref = 0;
  {
    MGLevelObject<LevelMatrix> mg_matrices(0, n_levels - 1);
    for (unsigned int level = 0; level < n_levels; ++level) {
      AffineConstraints<double> level_constraints;
      for (const types::global_dof_index i : mg_constrained_dofs.get_boundary_indices(level)) level_constraints.constrain_dof_to_zero(i);
      level_constraints.close();
      typename MatrixFree<dim, float>::AdditionalData data;
      data.tasks_parallel_scheme = MatrixFree<dim, float>::AdditionalData::none;
      data.mapping_update_flags = update_gradients | update_JxW_values | update_quadrature_points;
      data.mg_level = level;
      auto mf = std::make_shared<MatrixFree<dim, float>>();
      mf->reinit(mapping, dof, level_constraints, QGauss<1>(degree + 1), data);
      mg_matrices[level].initialize(mf, mg_constrained_dofs, level);
      mg_matrices[level].set_coefficient(coefficient_table<dim, degree, float>(*mf));
    }
    MGTransferMatrixFree<dim, float> mg_transfer(mg_constrained_dofs);
    mg_transfer.build(dof);
    using Smoother = PreconditionChebyshev<LevelMatrix, LevelVector>;
    mg::SmootherRelaxation<Smoother, LevelVector> mg_smoother;
    MGLevelObject<typename Smoother::AdditionalData> smoother_data(0, n_levels - 1);
    for (unsigned int level = 0; level < n_levels; ++level) {
      if (level > 0) {
        smoother_data[level].smoothing_range = 15.;
        smoother_data[level].degree = 5;
        smoother_data[level].eig_cg_n_iterations = 10;
      } else {
        smoother_data[0].smoothing_range = 1e-3;
        smoother_data[0].degree = numbers::invalid_unsigned_int;
        smoother_data[0].eig_cg_n_iterations = mg_matrices[0].m();
      }
      mg_matrices[level].compute_diagonal();
      smoother_data[level].preconditioner = mg_matrices[level].get_matrix_diagonal_inverse();
    }
    mg_smoother.initialize(mg_matrices, smoother_data);
    MGCoarseGridApplySmoother<LevelVector> mg_coarse;
    mg_coarse.initialize(mg_smoother);
    mg::Matrix<LevelVector> mg_matrix(mg_matrices);
    Multigrid<LevelVector> mg(mg_matrix, mg_coarse, mg_transfer, mg_smoother, mg_smoother);
    PreconditionMG<dim, LevelVector, MGTransferMatrixFree<dim, float>> preconditioner(dof, mg, mg_transfer);
    SolverControl control(100, tol);
    SolverCG<SystemVector> cg(control);
    cg.solve(system_matrix, x_ref, rhs, preconditioner);
    it_ref = control.last_step();
  }

  // ---- (2) the engine under the same SolverCG
  unsigned int it_gpu = 0;
  double diff = 0, norm_gpu = 0;
  {
    EngineLaplace<dim> A(mapping, dof, constraints, degree);
    b200::dealii_adapter::PreconditionMG<dim, float> preconditioner;
    preconditioner.reinit(mapping, dof, mg_constrained_dofs, QGauss<1>(degree + 1),
                          [](const Point<dim> &p) { return float(coefficient(p)); });
    typename EngineLaplace<dim>::VectorType x, b;
    A.initialize_dof_vector(x);
    Vector<double> host(n);
    for (unsigned int i = 0; i < n; ++i) host(i) = rhs.local_element(i);
    b.import_from_host(host);
    SolverControl control(100, tol);
    SolverCG<typename EngineLaplace<dim>::VectorType> cg(control);
    cg.solve(A, x, b, preconditioner);
    it_gpu = control.last_step();
    x.export_to_host(host);
    norm_gpu = host.l2_norm();
    for (unsigned int i = 0; i < n; ++i) host(i) -= x_ref.local_element(i);
    diff = host.l2_norm();
  }
  std::printf("Q%d, %u refinements: %u cells, %u DoFs, %u levels | deal.II GMG on the host: %u CG iterations, |u| = %.10g | "
              "engine PreconditionMG: %u CG iterations, |u| = %.10g | |u_engine - u_ref| / |u_ref| = %.3e\n",
              degree, refinements, tria.n_active_cells(), n, n_levels, it_ref, x_ref.l2_norm(), it_gpu, norm_gpu,
              diff / x_ref.l2_norm());
  const bool ok = (it_ref > it_gpu ? it_ref - it_gpu : it_gpu - it_ref) <= 1 && diff <= 1e-8 * x_ref.l2_norm();
  return ok ? 0 : 1;
}

int main() {
  int rc = 0;
  try {
    for (unsigned int r = 2; r <= 4; ++r) rc |= run<3, 2>(r); // step-37's shipped degree
    rc |= run<3, 4>(3);                                       // BASELINE configs[1]'s degree
    rc |= run<2, 3>(5);
  } catch (const std::exception &e) {
    std::fprintf(stderr, "exception: %s\n", e.what());
    return 2;
  }
  std::printf(rc == 0 ? "step-37 through libb200mf.so: OK\n" : "step-37 through libb200mf.so: MISMATCH\n");
  return rc;
}
