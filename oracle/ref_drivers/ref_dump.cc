// ref_dump -- TEST INFRASTRUCTURE: links the UNMODIFIED reference library built by
// oracle/build_ref.sh (oracle/_ref/install) and dumps, for one small configuration of the hot
// path, the arrays the engine consumes next to the results the reference itself computes:
//
//   setup arrays    the PrecomputedData of Portable::MatrixFree on the Kokkos Serial backend
//                   (local_to_global with hanging-node redirection, ConstraintKinds masks, JxW,
//                   inv_jacobian; matrix_free/portable_matrix_free.h:275-370), cell vertices,
//                   constrained dofs, coefficient values in local_q_point_id order
//   results         dst of the reference's CPU MatrixFree (FEEvaluation in MatrixFree::cell_loop,
//                   matrix_free/matrix_free.h:5090) and of Portable::MatrixFree::cell_loop +
//                   copy_constrained_values, MatrixFreeTools::compute_diagonal, and SolverCG
//                   iteration counts / solution norms with PreconditionIdentity and Jacobi.
//
// The operator is the one of tests/matrix_free_kokkos/matrix_vector_device_mf.h and step-64:
//   (grad v, grad u) + (v, a u)   with a = 0 (Laplace), a = 10, or a = 10 / (0.05 + 2 |x|^2).
//
// usage: ref_dump <dim> <degree> <refinements> <mesh: cartesian|deformed|hanging|ball>
//                 <op: laplace|helmholtz|helmholtz_var> <dirichlet: 0|1> <outdir> [lexicographic]
// (lexicographic: DoFRenumbering::lexicographic after distribute_dofs)
// Compiled once per degree (-DREF_DEGREE=k).  Output: raw little-endian arrays + manifest.json.
#include <deal.II/base/function.h>
#include <deal.II/base/quadrature_lib.h>

#include <deal.II/dofs/dof_handler.h>
#include <deal.II/dofs/dof_renumbering.h>
#include <deal.II/dofs/dof_tools.h>

#include <deal.II/fe/fe_q.h>
#include <deal.II/fe/mapping_q1.h>

#include <deal.II/grid/grid_generator.h>
#include <deal.II/grid/grid_tools.h>
#include <deal.II/grid/tria.h>

#include <deal.II/lac/affine_constraints.h>
#include <deal.II/lac/diagonal_matrix.h>
#include <deal.II/lac/la_parallel_vector.h>
#include <deal.II/lac/precondition.h>
#include <deal.II/lac/solver_cg.h>

#include <deal.II/matrix_free/fe_evaluation.h>
#include <deal.II/matrix_free/matrix_free.h>
#include <deal.II/matrix_free/portable_fe_evaluation.h>
#include <deal.II/matrix_free/portable_matrix_free.h>
#include <deal.II/matrix_free/tools.h>

#include <deal.II/numerics/vector_tools.h>

#include <cstdint>
#include <cstdio>
#include <fstream>
#include <random>
#include <string>

#ifndef REF_DEGREE
#  error "compile with -DREF_DEGREE=<degree>"
#endif
#ifndef REF_NQ_EXTRA
#  define REF_NQ_EXTRA 0 // quadrature points per direction beyond fe_degree + 1 + REF_NQ_EXTRA (over-integration)
#endif

using namespace dealii;

using HostVector   = LinearAlgebra::distributed::Vector<double, MemorySpace::Host>;
using DeviceVector = LinearAlgebra::distributed::Vector<double, MemorySpace::Default>;

// ---------------------------------------------------------------------------- output
struct Dump
{
  std::string   dir;
  std::ofstream manifest;
  bool          first = true;
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
  void
  text(const std::string &k, const std::string &v)
  {
    key(k);
    manifest << "\"" << v << "\"";
  }
  template <typename T>
  void
  array(const std::string &k, const std::vector<T> &v, const char *dtype)
  {
    std::ofstream f(dir + "/" + k + ".bin", std::ios::binary);
    f.write(reinterpret_cast<const char *>(v.data()), sizeof(T) * v.size());
    key(k);
    manifest << "{\"file\": \"" << k << ".bin\", \"dtype\": \"" << dtype << "\", \"size\": " << v.size() << "}";
  }
};

// ---------------------------------------------------------------------------- coefficient
template <int dim>
double
coefficient_value(const std::string &op, const Point<dim> &p)
{
  if (op == "laplace")
    return 0.;
  if (op == "helmholtz")
    return 10.;
  return 10. / (0.05 + 2. * p.square()); // tests/matrix_free_kokkos/matrix_vector_device_common.h:60-65
}

// ---------------------------------------------------------------------------- CPU MatrixFree
template <int dim, int degree>
class CpuOperator
{
public:
  CpuOperator(const MatrixFree<dim, double> &mf, const AffineConstraints<double> &constraints,
              const std::string &op)
    : mf(mf)
    , constraints(constraints)
    , has_mass(op != "laplace")
  {
    FEEvaluation<dim, degree, degree + 1 + REF_NQ_EXTRA, 1, double> phi(mf);
    coef.reinit(mf.n_cell_batches(), phi.n_q_points);
    for (unsigned int cell = 0; cell < mf.n_cell_batches(); ++cell)
      {
        phi.reinit(cell);
        for (unsigned int q = 0; q < phi.n_q_points; ++q)
          {
            const auto                 pq = phi.quadrature_point(q);
            VectorizedArray<double>    c  = 0.;
            for (unsigned int v = 0; v < mf.n_active_entries_per_cell_batch(cell); ++v)
              {
                Point<dim> pt;
                for (unsigned int d = 0; d < dim; ++d)
                  pt[d] = pq[d][v];
                c[v] = coefficient_value<dim>(op, pt);
              }
            coef(cell, q) = c;
          }
      }
  }

  // MatrixFreeOperators::Base::vmult semantics (matrix_free/operators.h:1487-1630): constrained
  // entries of src are read as 0 by read_dof_values, and dst_c = src_c afterwards
  void
  vmult(HostVector &dst, const HostVector &src) const
  {
    mf.cell_loop(&CpuOperator::local_apply, this, dst, src, /*zero_dst=*/true);
    for (const auto i : mf.get_constrained_dofs())
      dst.local_element(i) = src.local_element(i);
  }

  types::global_dof_index
  m() const
  {
    return mf.get_dof_handler().n_dofs();
  }

private:
  void
  local_apply(const MatrixFree<dim, double> &data, HostVector &dst, const HostVector &src,
              const std::pair<unsigned int, unsigned int> &range) const
  {
    FEEvaluation<dim, degree, degree + 1 + REF_NQ_EXTRA, 1, double> phi(data);
    for (unsigned int cell = range.first; cell < range.second; ++cell)
      {
        phi.reinit(cell);
        phi.read_dof_values(src);
        phi.evaluate(has_mass ? (EvaluationFlags::values | EvaluationFlags::gradients) :
                                EvaluationFlags::gradients);
        for (unsigned int q = 0; q < phi.n_q_points; ++q)
          {
            if (has_mass)
              phi.submit_value(coef(cell, q) * phi.get_value(q), q);
            phi.submit_gradient(phi.get_gradient(q), q);
          }
        phi.integrate(has_mass ? (EvaluationFlags::values | EvaluationFlags::gradients) :
                                 EvaluationFlags::gradients);
        phi.distribute_local_to_global(dst);
      }
  }

  const MatrixFree<dim, double>         &mf;
  const AffineConstraints<double>       &constraints;
  const bool                             has_mass;
  Table<2, VectorizedArray<double>>      coef;
};

// ---------------------------------------------------------------------------- Portable::MatrixFree
template <int dim, int degree>
class PmfQuad
{
public:
  DEAL_II_HOST_DEVICE
  PmfQuad(const double *coef, bool has_mass)
    : coef(coef)
    , has_mass(has_mass)
  {}
  DEAL_II_HOST_DEVICE void
  operator()(Portable::FEEvaluation<dim, degree, degree + 1 + REF_NQ_EXTRA, 1, double> *fe_eval, const int q) const
  {
    if (has_mass)
      {
        const unsigned int pos =
          fe_eval->get_matrix_free_data()->local_q_point_id(fe_eval->get_current_cell_index(), q);
        fe_eval->submit_value(coef[pos] * fe_eval->get_value(q), q);
      }
    fe_eval->submit_gradient(fe_eval->get_gradient(q), q);
  }
  static const unsigned int n_q_points = Utilities::pow(degree + 1 + REF_NQ_EXTRA, dim);

private:
  const double *coef;
  bool          has_mass;
};

template <int dim, int degree>
class PmfLocal
{
public:
  static const unsigned int n_q_points = Utilities::pow(degree + 1 + REF_NQ_EXTRA, dim);
  PmfLocal(const double *coef, bool has_mass)
    : coef(coef)
    , has_mass(has_mass)
  {}
  DEAL_II_HOST_DEVICE void
  operator()(const typename Portable::MatrixFree<dim, double>::Data *data,
             const Portable::DeviceVector<double> &src, Portable::DeviceVector<double> &dst) const
  {
    Portable::FEEvaluation<dim, degree, degree + 1 + REF_NQ_EXTRA, 1, double> fe_eval(data);
    fe_eval.read_dof_values(src);
    fe_eval.evaluate(has_mass ? (EvaluationFlags::values | EvaluationFlags::gradients) :
                                EvaluationFlags::gradients);
    PmfQuad<dim, degree> quad(coef, has_mass);
    data->for_each_quad_point([&](const int &q) { quad(&fe_eval, q); });
    fe_eval.integrate(has_mass ? (EvaluationFlags::values | EvaluationFlags::gradients) :
                                 EvaluationFlags::gradients);
    fe_eval.distribute_local_to_global(dst);
  }

private:
  const double *coef;
  bool          has_mass;
};

// ---------------------------------------------------------------------------- one case
template <int dim, int degree>
int
run(const unsigned int refinements, const std::string &mesh, const std::string &op,
    const bool dirichlet, const std::string &outdir, const bool lexicographic)
{
  Triangulation<dim> tria;
  GridGenerator::hyper_cube(tria, 0., 1.);
  tria.refine_global(refinements);
  if (mesh == "hanging" || mesh == "ball")
    {
      // one extra level inside a ball (hanging nodes on its surface), like
      // tests/matrix_free/matrix_vector_03.cc
      Point<dim> centre;
      for (unsigned int d = 0; d < dim; ++d)
        centre[d] = mesh == "ball" ? 0.5 : 0.3;
      const double radius = mesh == "ball" ? 0.3 : 0.35;
      for (const auto &cell : tria.active_cell_iterators())
        if (cell->center().distance(centre) < radius)
          cell->set_refine_flag();
      tria.execute_coarsening_and_refinement();
    }
  if (mesh == "deformed")
    {
      // smooth displacement of the vertices: every cell becomes a general (non-affine) cell
      GridTools::transform(
        [](const Point<dim> &p) {
          double s = 0.08;
          for (unsigned int d = 0; d < dim; ++d)
            s *= std::sin(numbers::PI * p[d]);
          Point<dim> q = p;
          for (unsigned int d = 0; d < dim; ++d)
            q[d] += s * (1. + 0.3 * d);
          return q;
        },
        tria);
    }

  const FE_Q<dim>  fe(degree);
  const MappingQ1<dim> mapping;
  const QGauss<1>  quad(degree + 1 + REF_NQ_EXTRA);
  DoFHandler<dim>  dof(tria);
  dof.distribute_dofs(fe);
  if (lexicographic)
    DoFRenumbering::lexicographic(dof); // dofs/dof_renumbering.h:1327-1342
  const unsigned int n_dofs = dof.n_dofs();

  AffineConstraints<double> constraints;
  DoFTools::make_hanging_node_constraints(dof, constraints);
  if (dirichlet)
    VectorTools::interpolate_boundary_values(mapping, dof, 0, Functions::ZeroFunction<dim>(), constraints);
  constraints.close();

  // ---- reference CPU MatrixFree
  MatrixFree<dim, double> mf;
  {
    typename MatrixFree<dim, double>::AdditionalData ad;
    ad.tasks_parallel_scheme = MatrixFree<dim, double>::AdditionalData::none;
    ad.mapping_update_flags  = update_values | update_gradients | update_JxW_values | update_quadrature_points;
    mf.reinit(mapping, dof, constraints, quad, ad);
  }
  CpuOperator<dim, degree> cpu_op(mf, constraints, op);

  // ---- reference Portable::MatrixFree (Kokkos Serial: host-accessible views)
  Portable::MatrixFree<dim, double> pmf;
  {
    typename Portable::MatrixFree<dim, double>::AdditionalData ad;
    ad.mapping_update_flags = update_values | update_gradients | update_JxW_values | update_quadrature_points;
    pmf.reinit(mapping, dof, constraints, quad, ad);
  }
  const auto         data    = pmf.get_data(0);
  const unsigned int n_cells = data.n_cells;
  const unsigned int npc     = fe.n_dofs_per_cell();
  const unsigned int nq      = Utilities::pow(degree + 1 + REF_NQ_EXTRA, dim);
  if (n_cells != tria.n_active_cells())
    {
      std::fprintf(stderr, "unexpected colouring: %u cells in colour 0 of %u\n", n_cells, tria.n_active_cells());
      return 2;
    }

  Dump out(outdir);
  out.text("reference", "deal.II " DEAL_II_PACKAGE_VERSION " (oracle/_ref), CPU MatrixFree + Portable::MatrixFree on Kokkos Serial");
  out.scalar("dim", dim);
  out.scalar("degree", degree);
  out.scalar("refinements", refinements);
  out.text("mesh", mesh);
  out.text("op", op);
  out.scalar("dirichlet", dirichlet);
  out.scalar("n_dofs", n_dofs);
  out.scalar("n_cells", n_cells);

  {
    std::vector<uint32_t> l2g((size_t)n_cells * npc);
    std::vector<uint16_t> mask(n_cells);
    std::vector<double>   jxw((size_t)n_cells * nq), invj((size_t)n_cells * nq * dim * dim),
      qpts((size_t)n_cells * nq * dim);
    for (unsigned int c = 0; c < n_cells; ++c)
      {
        for (unsigned int i = 0; i < npc; ++i)
          l2g[(size_t)c * npc + i] = data.local_to_global(i, c);
        mask[c] = static_cast<uint16_t>(data.constraint_mask(c));
        for (unsigned int q = 0; q < nq; ++q)
          {
            jxw[(size_t)c * nq + q] = data.JxW(q, c);
            for (unsigned int d = 0; d < dim; ++d)
              {
                qpts[((size_t)c * nq + q) * dim + d] = data.q_points(q, c)[d];
                for (unsigned int e = 0; e < dim; ++e)
                  invj[(((size_t)c * nq + q) * dim + d) * dim + e] = data.inv_jacobian(q, c, d, e);
              }
          }
      }
    out.array("local_to_global", l2g, "uint32");
    out.array("constraint_mask", mask, "uint16");
    out.array("JxW", jxw, "float64");
    out.array("inv_jacobian", invj, "float64");
    out.array("q_points", qpts, "float64");
    std::vector<double> sv(data.shape_values.size()), cg(data.co_shape_gradients.size()),
      cw(data.constraint_weights.size());
    for (size_t i = 0; i < sv.size(); ++i)
      sv[i] = data.shape_values(i);
    for (size_t i = 0; i < cg.size(); ++i)
      cg[i] = data.co_shape_gradients(i);
    for (size_t i = 0; i < cw.size(); ++i)
      cw[i] = data.constraint_weights(i);
    out.array("shape_values", sv, "float64");
    out.array("co_shape_gradients", cg, "float64");
    out.array("constraint_weights", cw, "float64");
  }
  {
    std::vector<double> vert;
    std::vector<uint32_t> plain; // DoFHandler numbering, lexicographic, no hanging-node redirection
    std::vector<types::global_dof_index> idx(npc);
    const auto &lex = mf.get_shape_info().lexicographic_numbering;
    for (const auto &cell : dof.active_cell_iterators())
      {
        for (unsigned int v = 0; v < GeometryInfo<dim>::vertices_per_cell; ++v)
          for (unsigned int d = 0; d < dim; ++d)
            vert.push_back(cell->vertex(v)[d]);
        cell->get_dof_indices(idx);
        for (unsigned int i = 0; i < npc; ++i)
          plain.push_back(idx[lex[i]]);
      }
    out.array("cell_vertices", vert, "float64");
    out.array("dof_indices_lexicographic", plain, "uint32");
  }
  std::vector<uint32_t> constrained, hanging;
  for (unsigned int i = 0; i < n_dofs; ++i)
    if (constraints.is_constrained(i))
      {
        constrained.push_back(i);
        if (constraints.get_constraint_entries(i) != nullptr && !constraints.get_constraint_entries(i)->empty())
          hanging.push_back(i);
      }
  out.array("constrained_dofs", constrained, "uint32");
  out.array("hanging_dofs", hanging, "uint32");

  // coefficient in local_q_point_id order (cell * n_q + q for one colour without padding effects:
  // local_q_point_id = (row_start / padding_length + cell) * n_q + q, portable_matrix_free.h:400-415)
  DeviceVector coef(n_cells * nq);
  {
    std::vector<double> c((size_t)n_cells * nq);
    for (unsigned int cell = 0; cell < n_cells; ++cell)
      for (unsigned int q = 0; q < nq; ++q)
        {
          Point<dim> pt;
          for (unsigned int d = 0; d < dim; ++d)
            pt[d] = data.q_points(q, cell)[d];
          c[(size_t)cell * nq + q]          = coefficient_value<dim>(op, pt);
          coef.local_element(cell * nq + q) = c[(size_t)cell * nq + q];
        }
    out.array("coefficient", c, "float64");
  }
  const bool has_mass = op != "laplace";

  // ---- vmult: same src through both reference paths
  HostVector src(n_dofs), dst_cpu(n_dofs);
  std::mt19937_64 rng(42);
  std::uniform_real_distribution<double> uni(0., 1.);
  for (unsigned int i = 0; i < n_dofs; ++i)
    src.local_element(i) = constraints.is_constrained(i) ? 0. : uni(rng);
  cpu_op.vmult(dst_cpu, src);

  DeviceVector src_dev(n_dofs), dst_pmf(n_dofs);
  for (unsigned int i = 0; i < n_dofs; ++i)
    src_dev.local_element(i) = src.local_element(i);
  PmfLocal<dim, degree> pmf_local(coef.get_values(), has_mass);
  auto pmf_vmult = [&](DeviceVector &d, const DeviceVector &s) {
    d = 0.;
    pmf.cell_loop(pmf_local, s, d);
    pmf.copy_constrained_values(s, d);
  };
  pmf_vmult(dst_pmf, src_dev);
  Kokkos::fence();

  std::vector<double> v_src(n_dofs), v_cpu(n_dofs), v_pmf(n_dofs);
  double diff = 0., scale = 0.;
  for (unsigned int i = 0; i < n_dofs; ++i)
    {
      v_src[i] = src.local_element(i);
      v_cpu[i] = dst_cpu.local_element(i);
      v_pmf[i] = dst_pmf.local_element(i);
      diff     = std::max(diff, std::abs(v_cpu[i] - v_pmf[i]));
      scale    = std::max(scale, std::abs(v_cpu[i]));
    }
  out.array("src", v_src, "float64");
  out.array("dst_cpu_matrixfree", v_cpu, "float64");
  out.array("dst_portable_matrixfree", v_pmf, "float64");
  out.scalar("max_abs_diff_cpu_vs_portable", diff);
  out.scalar("dst_linfty", scale);

  // ---- compute_diagonal (matrix_free/tools.h:1392-1569), constrained entries = 1
  DeviceVector diag(n_dofs);
  {
    PmfQuad<dim, degree> quad_op(coef.get_values(), has_mass);
    MatrixFreeTools::compute_diagonal<dim, degree, degree + 1 + REF_NQ_EXTRA, 1, double>(
      pmf, diag, quad_op, has_mass ? (EvaluationFlags::values | EvaluationFlags::gradients) : EvaluationFlags::gradients,
      has_mass ? (EvaluationFlags::values | EvaluationFlags::gradients) : EvaluationFlags::gradients);
    Kokkos::fence();
    std::vector<double> v(n_dofs);
    for (unsigned int i = 0; i < n_dofs; ++i)
      v[i] = diag.local_element(i);
    out.array("diagonal_portable", v, "float64");
  }

  // ---- SolverCG on the CPU operator: rhs = 1 (0 at constrained dofs), tol = 1e-12 |b|
  if (dirichlet)
    {
      HostVector b(n_dofs), x(n_dofs);
      for (unsigned int i = 0; i < n_dofs; ++i)
        b.local_element(i) = constraints.is_constrained(i) ? 0. : 1.;
      const double tol = 1e-12 * b.l2_norm();
      {
        SolverControl        control(10000, tol);
        SolverCG<HostVector> cg(control);
        x = 0.;
        cg.solve(cpu_op, x, b, PreconditionIdentity());
        out.scalar("cg_identity_iterations", control.last_step());
        out.scalar("cg_identity_solution_l2", x.l2_norm());
      }
      {
        DiagonalMatrix<HostVector> jacobi;
        HostVector                &inv = jacobi.get_vector();
        inv.reinit(n_dofs);
        for (unsigned int i = 0; i < n_dofs; ++i)
          inv.local_element(i) = 1. / diag.local_element(i);
        SolverControl        control(10000, tol);
        SolverCG<HostVector> cg(control);
        x = 0.;
        cg.solve(cpu_op, x, b, jacobi);
        out.scalar("cg_jacobi_iterations", control.last_step());
        out.scalar("cg_jacobi_solution_l2", x.l2_norm());
        out.scalar("cg_tolerance", tol);
        std::vector<double> v(n_dofs);
        for (unsigned int i = 0; i < n_dofs; ++i)
          v[i] = x.local_element(i);
        out.array("cg_jacobi_solution", v, "float64");
      }
    }
  std::printf("ref_dump: dim %d Q%d %s %s: %u cells, %u dofs, |cpu - portable|_inf / |dst|_inf = %.3e\n", dim,
              degree, mesh.c_str(), op.c_str(), n_cells, n_dofs, diff / scale);
  return 0;
}

int
main(int argc, char **argv)
{
  if (argc != 8 && argc != 9)
    {
      std::fprintf(stderr, "usage: %s dim degree refinements mesh op dirichlet outdir\n", argv[0]);
      return 1;
    }
  Kokkos::initialize();
  int rc = 0;
  {
    const int          dim = std::atoi(argv[1]), degree = std::atoi(argv[2]);
    const unsigned int refinements = std::atoi(argv[3]);
    const std::string  mesh = argv[4], op = argv[5], outdir = argv[7];
    const bool         dirichlet = std::atoi(argv[6]) != 0;
    const bool         lexicographic = argc == 9 && std::string(argv[8]) == "lexicographic";
    if (degree != REF_DEGREE)
      {
        std::fprintf(stderr, "this binary was compiled for degree %d\n", REF_DEGREE);
        rc = 1;
      }
    else if (dim == 2)
      rc = run<2, REF_DEGREE>(refinements, mesh, op, dirichlet, outdir, lexicographic);
    else
      rc = run<3, REF_DEGREE>(refinements, mesh, op, dirichlet, outdir, lexicographic);
  }
  Kokkos::finalize();
  return rc;
}
