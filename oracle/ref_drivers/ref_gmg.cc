// ref_gmg -- TEST INFRASTRUCTURE: links the UNMODIFIED reference library (oracle/_ref/install) and runs
// the reference's matrix-free geometric multigrid on the problem of examples/step-37:
//
//   -div(a grad u) = 1 on (0,1)^dim, u = 0 on the boundary, FE_Q(degree), refine_global(r),
//   a = 1 ("constant") or a = 1 / (0.05 + 2 |x|^2) ("step37"),
//   SolverCG (double) preconditioned by one V-cycle (PreconditionMG + Multigrid, level vectors in
//   float or double) with MGTransferMatrixFree, PreconditionChebyshev smoothers (degree 5, range 15,
//   10 Lanczos iterations) and the Chebyshev "solver" on level 0 (tolerance 1e-3, full Lanczos),
//
// i.e. the classes and parameters of step-37.cc:950-1060, with the library's own
// MatrixFreeOperators::LaplaceOperator as level and system operator.  It dumps what
// tests/test_multigrid_gpu.py compares the engine's multigrid with: per-level dof numbering and
// eigenvalue estimates, one application of the transfer operators, one V-cycle, the CG iteration count
// and the solution.
//
// usage: ref_gmg <dim> <refinements> <levels: f32|f64> <coef: constant|step37> <outdir> [timing] [deform=<a>]
// (timing: no array dumps, relative tolerance 1e-6, best of two solves, one JSON line on stdout)
// Compiled once per degree (-DREF_DEGREE=k).  Output: raw little-endian arrays + manifest.json.
#include <deal.II/base/quadrature_lib.h>

#include <deal.II/dofs/dof_handler.h>
#include <deal.II/dofs/dof_tools.h>

#include <deal.II/fe/fe_q.h>
#include <deal.II/fe/fe_tools.h>
#include <deal.II/fe/mapping_q1.h>

#include <deal.II/grid/grid_generator.h>
#include <deal.II/grid/grid_tools.h>
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

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <string>

#ifndef REF_DEGREE
#  error "compile with -DREF_DEGREE=<degree>"
#endif

using namespace dealii;

struct Dump
{
  std::string   dir;
  std::ofstream manifest;
  bool          first  = true;
  bool          arrays = true; // false: scalars only (timing runs on meshes too large to dump)
  explicit Dump(const std::string &d)
    : dir(d)
    , manifest(d + "/manifest.json")
  {
    manifest << "{";
  }
  ~Dump()
  {
    manifest << "\n}\n";
  }
  void
  key(const std::string &k)
  {
    manifest << (first ? "\n" : ",\n") << "  \"" << k << "\": ";
    first = false;
  }
  void
  scalar(const std::string &k, double v)
  {
    key(k);
    char buf[64];
    std::snprintf(buf, sizeof(buf), "%.17g", v);
    manifest << buf;
  }
  template <typename T>
  void
  array(const std::string &k, const std::vector<T> &v, const char *dtype)
  {
    if (!arrays)
      return;
    std::ofstream f(dir + "/" + k + ".bin", std::ios::binary);
    f.write(reinterpret_cast<const char *>(v.data()), sizeof(T) * v.size());
    key(k);
    manifest << "{\"file\": \"" << k << ".bin\", \"dtype\": \"" << dtype << "\", \"size\": " << v.size() << "}";
  }
  template <typename Number>
  void
  vector(const std::string &k, const LinearAlgebra::distributed::Vector<Number> &v)
  {
    if (!arrays)
      return;
    std::vector<double> out(v.size());
    for (unsigned int i = 0; i < v.size(); ++i)
      out[i] = v.local_element(i);
    array(k, out, "float64");
  }
};

template <int dim>
double
coefficient_value(const bool variable, const Point<dim> &p)
{
  return variable ? 1. / (0.05 + 2. * p.square()) : 1.;
}

template <int dim, int degree, typename Number>
std::shared_ptr<Table<2, VectorizedArray<Number>>>
evaluate_coefficient(const MatrixFree<dim, Number> &mf, const bool variable)
{
  FEEvaluation<dim, degree, degree + 1, 1, Number> phi(mf);
  auto table = std::make_shared<Table<2, VectorizedArray<Number>>>(mf.n_cell_batches(), phi.n_q_points);
  for (unsigned int cell = 0; cell < mf.n_cell_batches(); ++cell)
    {
      phi.reinit(cell);
      for (unsigned int q = 0; q < phi.n_q_points; ++q)
        {
          const auto            xq = phi.quadrature_point(q);
          VectorizedArray<Number> a;
          for (unsigned int v = 0; v < VectorizedArray<Number>::size(); ++v)
            {
              Point<dim> p;
              for (unsigned int d = 0; d < dim; ++d)
                p[d] = xq[d][v];
              a[v] = coefficient_value(variable, p);
            }
          (*table)(cell, q) = a;
        }
    }
  return table;
}

template <int dim, int degree, typename LevelNumber>
void
run(const unsigned int refinements, const bool variable, const std::string &outdir, const bool timing,
    const double deformation)
{
  using SystemVector = LinearAlgebra::distributed::Vector<double>;
  using LevelVector  = LinearAlgebra::distributed::Vector<LevelNumber>;
  using SystemMatrix = MatrixFreeOperators::LaplaceOperator<dim, degree, degree + 1, 1, SystemVector>;
  using LevelMatrix  = MatrixFreeOperators::LaplaceOperator<dim, degree, degree + 1, 1, LevelVector>;

  Dump dump(outdir);
  dump.arrays = !timing;
  Triangulation<dim> tria(Triangulation<dim>::limit_level_difference_at_vertices);
  GridGenerator::hyper_cube(tria, 0., 1.);
  tria.refine_global(refinements);
  if (deformation != 0.)
    // the displacement of the engine's mesh generator (B200MF_DEFORM_SINE): every cell of every level becomes a
    // general cell (the vertices of all levels are moved, like the level meshes the engine generates)
    GridTools::transform(
      [deformation](const Point<dim> &p) {
        double s = deformation;
        for (unsigned int d = 0; d < dim; ++d)
          s *= std::sin(numbers::PI * p[d]);
        Point<dim> q = p;
        for (unsigned int d = 0; d < dim; ++d)
          q[d] += s;
        return q;
      },
      tria);
  dump.scalar("deformation", deformation);
  const FE_Q<dim>    fe(degree);
  const MappingQ1<dim> mapping;
  DoFHandler<dim>    dof_handler(tria);
  dof_handler.distribute_dofs(fe);
  dof_handler.distribute_mg_dofs();
  const unsigned int n_levels = tria.n_global_levels();
  dump.scalar("dim", dim);
  dump.scalar("degree", degree);
  dump.scalar("refinements", refinements);
  dump.scalar("n_levels", n_levels);
  dump.scalar("n_dofs", dof_handler.n_dofs());
  dump.scalar("variable_coefficient", variable);
  dump.scalar("level_number_bytes", sizeof(LevelNumber));

  // ---- system operator
  AffineConstraints<double> constraints;
  constraints.reinit(dof_handler.locally_owned_dofs(), DoFTools::extract_locally_relevant_dofs(dof_handler));
  VectorTools::interpolate_boundary_values(mapping, dof_handler, 0, Functions::ZeroFunction<dim>(), constraints);
  constraints.close();
  SystemMatrix system_matrix;
  {
    typename MatrixFree<dim, double>::AdditionalData data;
    data.tasks_parallel_scheme = MatrixFree<dim, double>::AdditionalData::none;
    data.mapping_update_flags  = update_gradients | update_JxW_values | update_quadrature_points;
    auto mf = std::make_shared<MatrixFree<dim, double>>();
    mf->reinit(mapping, dof_handler, constraints, QGauss<1>(degree + 1), data);
    system_matrix.initialize(mf);
    system_matrix.set_coefficient(evaluate_coefficient<dim, degree, double>(*mf, variable));
  }

  // ---- level operators (step-37.cc:812-850)
  MGConstrainedDoFs mg_constrained_dofs;
  mg_constrained_dofs.initialize(dof_handler);
  mg_constrained_dofs.make_zero_boundary_constraints(dof_handler, {0});
  MGLevelObject<LevelMatrix> mg_matrices(0, n_levels - 1);
  const std::vector<unsigned int> l2h = FETools::lexicographic_to_hierarchic_numbering<dim>(degree);
  for (unsigned int level = 0; level < n_levels; ++level)
    {
      AffineConstraints<double> level_constraints(dof_handler.locally_owned_mg_dofs(level),
                                                  DoFTools::extract_locally_relevant_level_dofs(dof_handler, level));
      for (const types::global_dof_index i : mg_constrained_dofs.get_boundary_indices(level))
        level_constraints.constrain_dof_to_zero(i);
      level_constraints.close();
      typename MatrixFree<dim, LevelNumber>::AdditionalData data;
      data.tasks_parallel_scheme = MatrixFree<dim, LevelNumber>::AdditionalData::none;
      data.mapping_update_flags  = update_gradients | update_JxW_values | update_quadrature_points;
      data.mg_level              = level;
      auto mf = std::make_shared<MatrixFree<dim, LevelNumber>>();
      mf->reinit(mapping, dof_handler, level_constraints, QGauss<1>(degree + 1), data);
      mg_matrices[level].initialize(mf, mg_constrained_dofs, level);
      mg_matrices[level].set_coefficient(evaluate_coefficient<dim, degree, LevelNumber>(*mf, variable));

      // the level numbering, per level cell in lexicographic local order (compared with the engine's
      // mesh generator on the mesh of that level)
      std::vector<std::uint32_t>           l2g;
      std::vector<types::global_dof_index> idx(fe.n_dofs_per_cell());
      if (!timing)
      for (const auto &cell : dof_handler.mg_cell_iterators_on_level(level))
        {
          cell->get_mg_dof_indices(idx);
          for (unsigned int i = 0; i < idx.size(); ++i)
            l2g.push_back(idx[l2h[i]]);
        }
      dump.array("level_l2g_" + std::to_string(level), l2g, "uint32");
    }

  MGTransferMatrixFree<dim, LevelNumber> mg_transfer(mg_constrained_dofs);
  mg_transfer.build(dof_handler);

  // ---- smoothers (step-37.cc:956-988)
  using Smoother = PreconditionChebyshev<LevelMatrix, LevelVector>;
  mg::SmootherRelaxation<Smoother, LevelVector>     mg_smoother;
  MGLevelObject<typename Smoother::AdditionalData> smoother_data(0, n_levels - 1);
  for (unsigned int level = 0; level < n_levels; ++level)
    {
      if (level > 0)
        {
          smoother_data[level].smoothing_range     = 15.;
          smoother_data[level].degree              = 5;
          smoother_data[level].eig_cg_n_iterations = 10;
        }
      else
        {
          smoother_data[0].smoothing_range     = 1e-3;
          smoother_data[0].degree              = numbers::invalid_unsigned_int;
          smoother_data[0].eig_cg_n_iterations = mg_matrices[0].m();
        }
      mg_matrices[level].compute_diagonal();
      smoother_data[level].preconditioner = mg_matrices[level].get_matrix_diagonal_inverse();
    }
  mg_smoother.initialize(mg_matrices, smoother_data);
  for (unsigned int level = 0; level < n_levels; ++level)
    {
      LevelVector v;
      mg_matrices[level].initialize_dof_vector(v);
      const auto info = mg_smoother[level].estimate_eigenvalues(v);
      dump.scalar("eig_min_" + std::to_string(level), info.min_eigenvalue_estimate);
      dump.scalar("eig_max_" + std::to_string(level), info.max_eigenvalue_estimate);
      dump.scalar("cheb_degree_" + std::to_string(level), info.degree);
      dump.scalar("eig_cg_iterations_" + std::to_string(level), info.cg_iterations);
      dump.vector("level_inverse_diagonal_" + std::to_string(level),
                  mg_matrices[level].get_matrix_diagonal_inverse()->get_vector());
    }
  MGCoarseGridApplySmoother<LevelVector> mg_coarse;
  mg_coarse.initialize(mg_smoother);

  // ---- transfer: one prolongation and one restriction between the two finest levels
  if (n_levels > 1)
    {
      const unsigned int top = n_levels - 1;
      LevelVector        coarse, fine;
      mg_matrices[top - 1].initialize_dof_vector(coarse);
      mg_matrices[top].initialize_dof_vector(fine);
      for (unsigned int i = 0; i < coarse.size(); ++i)
        coarse.local_element(i) = 0.25 * ((i * 7 + 3) % 13) - 1.0;
      for (const types::global_dof_index i : mg_constrained_dofs.get_boundary_indices(top - 1))
        coarse[i] = 0;
      dump.vector("prolongate_src", coarse);
      mg_transfer.prolongate(top, fine, coarse);
      dump.vector("prolongate_dst", fine);
      for (unsigned int i = 0; i < fine.size(); ++i)
        fine.local_element(i) = 0.125 * ((i * 5 + 1) % 17) - 1.0;
      for (const types::global_dof_index i : mg_constrained_dofs.get_boundary_indices(top))
        fine[i] = 0;
      dump.vector("restrict_src", fine);
      coarse = 0;
      mg_transfer.restrict_and_add(top, coarse, fine);
      dump.vector("restrict_dst", coarse);
    }

  // ---- V-cycle preconditioner and the solve (step-37.cc:1021-1075)
  mg::Matrix<LevelVector> mg_matrix(mg_matrices);
  Multigrid<LevelVector>  mg(mg_matrix, mg_coarse, mg_transfer, mg_smoother, mg_smoother);
  PreconditionMG<dim, LevelVector, MGTransferMatrixFree<dim, LevelNumber>> preconditioner(dof_handler, mg, mg_transfer);

  SystemVector rhs, solution, z;
  system_matrix.initialize_dof_vector(rhs);
  system_matrix.initialize_dof_vector(solution);
  system_matrix.initialize_dof_vector(z);
  {
    // rhs_i = (phi_i, 1), constrained entries zero (step-37.cc:868-885)
    const auto                                   &mf = *system_matrix.get_matrix_free();
    FEEvaluation<dim, degree, degree + 1, 1, double> phi(mf);
    for (unsigned int cell = 0; cell < mf.n_cell_batches(); ++cell)
      {
        phi.reinit(cell);
        for (unsigned int q = 0; q < phi.n_q_points; ++q)
          phi.submit_value(make_vectorized_array<double>(1.0), q);
        phi.integrate(EvaluationFlags::values);
        phi.distribute_local_to_global(rhs);
      }
    rhs.compress(VectorOperation::add);
  }
  dump.vector("rhs", rhs);
  preconditioner.vmult(z, rhs);
  dump.vector("vcycle_of_rhs", z);

  const double           rel_tol = timing ? 1e-6 : 1e-12; // timing runs use bench.py's --cg-rel-tol default
  SolverControl          control(100, rel_tol * rhs.l2_norm());
  SolverCG<SystemVector> cg(control);
  double                 best_seconds = 1e300;
  for (unsigned int rep = 0; rep < (timing ? 2u : 1u); ++rep)
    {
      solution = 0;
      constraints.set_zero(solution);
      const auto t0 = std::chrono::steady_clock::now();
      cg.solve(system_matrix, solution, rhs, preconditioner);
      best_seconds = std::min(best_seconds, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    }
  constraints.distribute(solution);
  dump.scalar("solve_seconds", best_seconds);
  if (timing)
    std::printf("{\"ref_gmg\": true, \"dim\": %d, \"degree\": %d, \"refinements\": %u, \"n_dofs\": %u, \"levels\": \"%s\", "
                "\"relative_tolerance\": %g, \"iterations\": %u, \"seconds\": %.6f, \"cores\": 1}\n",
                dim, degree, refinements, (unsigned)dof_handler.n_dofs(), sizeof(LevelNumber) == 4 ? "f32" : "f64", rel_tol,
                control.last_step(), best_seconds);
  dump.scalar("cg_iterations", control.last_step());
  dump.scalar("cg_residual", control.last_value());
  dump.scalar("rhs_l2", rhs.l2_norm());
  dump.scalar("solution_l2", solution.l2_norm());
  dump.vector("solution", solution);
  std::printf("ref_gmg: dim %d Q%d r%u levels %s coef %s: %u dofs, %u CG iterations, |x| = %.12e\n", dim, degree,
              refinements, sizeof(LevelNumber) == 4 ? "f32" : "f64", variable ? "step37" : "constant",
              (unsigned)dof_handler.n_dofs(), control.last_step(), solution.l2_norm());
}

int
main(int argc, char **argv)
{
  if (argc < 6)
    {
      std::fprintf(stderr, "usage: ref_gmg <dim> <refinements> <f32|f64> <constant|step37> <outdir>\n");
      return 2;
    }
  const int          dim         = std::atoi(argv[1]);
  const unsigned int refinements = std::atoi(argv[2]);
  const bool         f32         = std::string(argv[3]) == "f32";
  const bool         variable    = std::string(argv[4]) == "step37";
  const std::string  outdir      = argv[5];
  bool               timing      = false;
  double             deformation = 0.;
  for (int i = 6; i < argc; ++i)
    {
      if (std::string(argv[i]) == "timing")
        timing = true;
      if (std::string(argv[i]).rfind("deform=", 0) == 0)
        deformation = std::atof(argv[i] + 7);
    }
  constexpr int      degree      = REF_DEGREE;
  if (dim == 2)
    f32 ? run<2, degree, float>(refinements, variable, outdir, timing, deformation) : run<2, degree, double>(refinements, variable, outdir, timing, deformation);
  else
    f32 ? run<3, degree, float>(refinements, variable, outdir, timing, deformation) : run<3, degree, double>(refinements, variable, outdir, timing, deformation);
  return 0;
}
