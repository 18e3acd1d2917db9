// ref_bench -- TEST INFRASTRUCTURE / CPU BASELINE: times the reference's own vectorised CPU
// MatrixFree path on one core, linking the UNMODIFIED library of oracle/build_ref.sh.
//
// The operator is LaplaceOperator<dim, degree, Number, MemorySpace::Host> of
// tests/performance/timing_matrix_free_kokkos.cc:56-111 (read_dof_values_plain, evaluate(gradients),
// submit_gradient(get_gradient), integrate, distribute_local_to_global inside MatrixFree::cell_loop)
// on GridGenerator::hyper_cube + refine_global(r), MappingQ1, QGauss(degree+1), no constraints, src =
// interpolant of sum_d sin(x_d) -- the set-up of run() in that file (:217-300) on a serial
// Triangulation (there is no MPI in this image).  bench.py starts one instance per host core, pinned,
// at the same time; aggregate DoFs/s = sum over instances.
//
// With `cg` as mode the same mesh gets zero Dirichlet values and SolverCG + Jacobi
// (DiagonalMatrix of the inverse operator diagonal) runs a fixed number of iterations: the CG
// half of the metric on the CPU.
//
// usage: ref_bench <degree> <refinements> <steps> <warmup> [vmult|cg] [start_file]
//   start_file: if given, the timed region starts when that file exists (common start of all
//   instances).   Prints one JSON line.
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
#include <deal.II/lac/solver_cg.h>

#include <deal.II/matrix_free/fe_evaluation.h>
#include <deal.II/matrix_free/matrix_free.h>
#include <deal.II/matrix_free/tools.h>

#include <deal.II/numerics/vector_tools.h>

#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <string>

using namespace dealii;
using VectorType = LinearAlgebra::distributed::Vector<double, MemorySpace::Host>;

template <int dim, int degree>
class LaplaceOperator
{
public:
  void
  reinit(const Mapping<dim> &mapping, const DoFHandler<dim> &dof, const AffineConstraints<double> &constraints,
         const Quadrature<1> &quad)
  {
    typename MatrixFree<dim, double>::AdditionalData ad;
    ad.mapping_update_flags  = update_gradients;
    ad.tasks_parallel_scheme = MatrixFree<dim, double>::AdditionalData::none;
    matrix_free.reinit(mapping, dof, constraints, quad, ad);
  }
  void
  initialize_dof_vector(VectorType &v) const
  {
    matrix_free.initialize_dof_vector(v);
  }
  // dst = A src (dst zeroed inside the loop), the engine's vmult
  void
  vmult(VectorType &dst, const VectorType &src) const
  {
    matrix_free.cell_loop(&LaplaceOperator::local_apply, this, dst, src, /*zero_dst=*/true);
    for (const auto i : matrix_free.get_constrained_dofs())
      dst.local_element(i) = src.local_element(i);
  }
  void
  compute_inverse_diagonal(VectorType &inv) const
  {
    matrix_free.initialize_dof_vector(inv);
    MatrixFreeTools::compute_diagonal<dim, degree, degree + 1, 1, double, VectorizedArray<double>>(
      matrix_free, inv, [](auto &phi) {
        phi.evaluate(EvaluationFlags::gradients);
        for (unsigned int q = 0; q < phi.n_q_points; ++q)
          phi.submit_gradient(phi.get_gradient(q), q);
        phi.integrate(EvaluationFlags::gradients);
      });
    for (auto &v : inv)
      v = (v != 0.) ? 1. / v : 1.;
  }

private:
  void
  local_apply(const MatrixFree<dim, double> &data, VectorType &dst, const VectorType &src,
              const std::pair<unsigned int, unsigned int> &range) const
  {
    FEEvaluation<dim, degree, degree + 1, 1, double> phi(data);
    for (unsigned int cell = range.first; cell < range.second; ++cell)
      {
        phi.reinit(cell);
        phi.read_dof_values(src);
        phi.evaluate(EvaluationFlags::gradients);
        for (unsigned int q = 0; q < phi.n_q_points; ++q)
          phi.submit_gradient(phi.get_gradient(q), q);
        phi.integrate(EvaluationFlags::gradients);
        phi.distribute_local_to_global(dst);
      }
  }
  MatrixFree<dim, double> matrix_free;
};

template <int dim>
class AnalyticalFunction : public Function<dim>
{
public:
  double
  value(const Point<dim> &p, const unsigned int = 0) const override
  {
    double t = 0.;
    for (unsigned int d = 0; d < dim; ++d)
      t += std::sin(p[d]);
    return t;
  }
};

static double
now()
{
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

template <int degree>
int
run(unsigned int refinements, unsigned int steps, unsigned int warmup, const std::string &mode,
    const std::string &start_file)
{
  constexpr int      dim = 3;
  Triangulation<dim> tria;
  GridGenerator::hyper_cube(tria);
  tria.refine_global(refinements);
  const MappingQ1<dim> mapping;
  const FE_Q<dim>      fe(degree);
  const QGauss<1>      quad(degree + 1);
  DoFHandler<dim>      dof(tria);
  dof.distribute_dofs(fe);
  AffineConstraints<double> constraints;
  if (mode == "cg")
    VectorTools::interpolate_boundary_values(mapping, dof, 0, Functions::ZeroFunction<dim>(), constraints);
  constraints.close();
  LaplaceOperator<dim, degree> op;
  op.reinit(mapping, dof, constraints, quad);
  VectorType src, dst;
  op.initialize_dof_vector(src);
  op.initialize_dof_vector(dst);
  VectorTools::interpolate(mapping, dof, AnalyticalFunction<dim>(), src);
  constraints.set_zero(src);

  auto wait_for_start = [&]() {
    if (start_file.empty())
      return;
    struct stat st;
    while (stat(start_file.c_str(), &st) != 0)
      usleep(200);
  };
  double       seconds = 0., checksum = 0.;
  unsigned int iterations = 0;
  if (mode == "vmult")
    {
      for (unsigned int i = 0; i < warmup; ++i)
        op.vmult(dst, src);
      wait_for_start();
      const double t0 = now();
      for (unsigned int i = 0; i < steps; ++i)
        op.vmult(dst, src);
      seconds  = now() - t0;
      checksum = dst.l2_norm();
    }
  else
    {
      DiagonalMatrix<VectorType> jacobi;
      op.compute_inverse_diagonal(jacobi.get_vector());
      VectorType b, x;
      op.initialize_dof_vector(b);
      op.initialize_dof_vector(x);
      b = 1.;
      constraints.set_zero(b);
      for (unsigned int rep = 0; rep < 2; ++rep) // first solve = warm-up
        {
          IterationNumberControl control(steps, 1e-300);
          SolverCG<VectorType>   cg(control);
          x = 0.;
          if (rep == 1)
            wait_for_start();
          const double t0 = now();
          cg.solve(op, x, b, jacobi);
          seconds    = now() - t0;
          iterations = control.last_step();
        }
      checksum = x.l2_norm();
    }
  std::printf("{\"degree\": %d, \"refinements\": %u, \"n_dofs\": %llu, \"n_cells\": %llu, \"mode\": \"%s\", "
              "\"steps\": %u, \"warmup\": %u, \"iterations\": %u, \"seconds\": %.9g, \"checksum\": %.15g, "
              "\"vectorization_lanes\": %u}\n",
              degree, refinements, (unsigned long long)dof.n_dofs(), (unsigned long long)tria.n_active_cells(),
              mode.c_str(), steps, warmup, iterations, seconds, checksum,
              (unsigned int)VectorizedArray<double>::size());
  return 0;
}

int
main(int argc, char **argv)
{
  if (argc < 5)
    {
      std::fprintf(stderr, "usage: %s degree refinements steps warmup [vmult|cg] [start_file]\n", argv[0]);
      return 1;
    }
  const int          degree      = std::atoi(argv[1]);
  const unsigned int refinements = std::atoi(argv[2]), steps = std::atoi(argv[3]), warmup = std::atoi(argv[4]);
  const std::string  mode = argc > 5 ? argv[5] : "vmult", start_file = argc > 6 ? argv[6] : "";
  switch (degree)
    {
      case 1: return run<1>(refinements, steps, warmup, mode, start_file);
      case 2: return run<2>(refinements, steps, warmup, mode, start_file);
      case 3: return run<3>(refinements, steps, warmup, mode, start_file);
      case 4: return run<4>(refinements, steps, warmup, mode, start_file);
      case 5: return run<5>(refinements, steps, warmup, mode, start_file);
      case 6: return run<6>(refinements, steps, warmup, mode, start_file);
      case 8: return run<8>(refinements, steps, warmup, mode, start_file);
      default: std::fprintf(stderr, "degree %d not instantiated\n", degree); return 1;
    }
}
