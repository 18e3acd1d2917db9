// b200mf_dealii.hpp -- the deal.II-side adapter: Portable::MatrixFree's surface for the hot path,
// filled from deal.II objects, running on libb200mf.so.  Needs deal.II headers (any build: the
// adapter only uses host-side classes) next to b200mf.h / b200mf_portable.hpp.
//
//   b200::dealii_adapter::MatrixFree<dim, Number>
//       reinit(mapping, dof_handler, constraints, Quadrature<1>, AdditionalData)
//                                   matrix_free/portable_matrix_free.h:480-527; extracts exactly what
//                                   ReinitHelper::fill_data extracts (portable_matrix_free.templates.h:
//                                   267-346): lexicographic index lists, HangingNodes::setup_constraints
//                                   masks + redirected indices, cell geometry, constrained dofs (:1366-1425)
//       initialize_dof_vector(vec)  :637         get_vector_partitioner()  :682
//       evaluate_coefficients(f)    :585  (f: Point<dim> -> Number on the host; returns the device array in
//                                   local_q_point_id order, :400-415)
//       copy_constrained_values / set_constrained_values / cell_loop  (inherited from b200::MatrixFree)
//   b200::dealii_adapter::Vector<Number>
//       device-resident vector with the vector-space interface deal.II's own solvers need
//       (lac/solver_cg.h:600-760 generic IterationWorker): the UNMODIFIED SolverCG runs on it, every
//       operation being one b200mf_vec_* kernel (lac/vector_operations_internal.h:2140-2660).
//   b200::dealii_adapter::DiagonalPreconditioner<Number>   DiagonalMatrix::vmult (lac/diagonal_matrix.h:435)
//   b200::dealii_adapter::PreconditionMG<dim, LevelNumber>
//       step-37's PreconditionMG / Multigrid / MGTransferMatrixFree / Chebyshev smoothers from a DoFHandler
//       with distribute_mg_dofs() + MGConstrainedDoFs, one V-cycle per vmult on the device
//       (examples/step37_dealii.cc)
//
// This image's deal.II has no MPI and Kokkos::Serial only, so LinearAlgebra::distributed::Vector<Number,
// MemorySpace::Default> lives in host memory there; on a CUDA-enabled deal.II its get_values() is a
// device pointer and can be handed to b200::Operator::vmult directly.
#ifndef B200MF_DEALII_HPP
#define B200MF_DEALII_HPP

#include <deal.II/base/partitioner.h>
#include <deal.II/base/quadrature_lib.h>

#include <deal.II/dofs/dof_handler.h>

#include <deal.II/fe/fe_values.h>
#include <deal.II/fe/mapping_q.h>

#include <deal.II/lac/affine_constraints.h>
#include <deal.II/lac/vector.h>
#include <deal.II/lac/vector_memory.templates.h> // GrowingVectorMemory for the adapter's vector type

#include <deal.II/matrix_free/hanging_nodes_internal.h>
#include <deal.II/matrix_free/shape_info.h>

#include <deal.II/multigrid/mg_constrained_dofs.h>

#include <cuda_runtime_api.h>

#include <functional>
#include <map>
#include <memory>

#include "b200mf_portable.hpp"

namespace b200 {
namespace dealii_adapter {

inline void cuda_check(cudaError_t e) {
  if (e != cudaSuccess) throw Exception(B200MF_ERR_CUDA, cudaGetErrorString(e));
}

// ------------------------------------------------------------------------------------------
template <typename Number>
class Vector {
public:
  using value_type = Number;
  using real_type  = Number;
  using size_type  = dealii::types::global_dof_index;

  Vector() = default;
  explicit Vector(size_type n) { reinit(n); }
  Vector(const Vector &v) { *this = v; }
  ~Vector() { cudaFree(data_); }

  void reinit(size_type n, bool omit_zeroing = false) {
    if (n != size_) {
      cudaFree(data_);
      data_ = nullptr;
      size_ = n;
      if (n) cuda_check(cudaMalloc(reinterpret_cast<void **>(&data_), n * sizeof(Number)));
    }
    if (!omit_zeroing && n) cuda_check(cudaMemset(data_, 0, n * sizeof(Number)));
  }
  void reinit(const Vector &v, bool omit_zeroing = false) { reinit(v.size_, omit_zeroing); }

  Vector &operator=(const Vector &v) {
    reinit(v.size_, true);
    if (size_) cuda_check(cudaMemcpy(data_, v.data_, size_ * sizeof(Number), cudaMemcpyDeviceToDevice));
    return *this;
  }
  Vector &operator=(Number s) {
    check(b200mf_vec_set(code(), data_, double(s), size_, nullptr));
    return *this;
  }
  void swap(Vector &v) noexcept { std::swap(data_, v.data_); std::swap(size_, v.size_); }
  size_type size() const { return size_; }
  size_type locally_owned_size() const { return size_; }
  Number *get_values() { return data_; }
  const Number *get_values() const { return data_; }

  void add(Number a, const Vector &x) { check(b200mf_vec_axpy(code(), data_, double(a), x.data_, size_, nullptr)); }
  void sadd(Number s, Number a, const Vector &x) { check(b200mf_vec_sadd(code(), data_, double(s), double(a), x.data_, size_, nullptr)); }
  void equ(Number a, const Vector &x) { check(b200mf_vec_equ(code(), data_, double(a), x.data_, 0.0, nullptr, size_, nullptr)); }
  void scale(const Vector &d) { check(b200mf_vec_scale(code(), data_, 1.0, d.data_, size_, nullptr)); }
  Vector &operator*=(Number a) { check(b200mf_vec_scale(code(), data_, double(a), nullptr, size_, nullptr)); return *this; }
  Vector &operator/=(Number a) { return *this *= Number(1) / a; }
  Vector &operator+=(const Vector &x) { add(Number(1), x); return *this; }
  Vector &operator-=(const Vector &x) { add(Number(-1), x); return *this; }
  Number operator*(const Vector &y) const {
    double r = 0;
    check(b200mf_vec_dot(code(), data_, y.data_, size_, &r, nullptr));
    return Number(r);
  }
  real_type l2_norm() const {
    double r = 0;
    check(b200mf_vec_norm_2(code(), data_, size_, &r, nullptr));
    return real_type(r);
  }
  real_type norm_sqr() const { const real_type n = l2_norm(); return n * n; }
  real_type l1_norm() const { double r = 0; check(b200mf_vec_norm_1(code(), data_, size_, &r, nullptr)); return real_type(r); }
  real_type linfty_norm() const { double r = 0; check(b200mf_vec_norm_linfty(code(), data_, size_, &r, nullptr)); return real_type(r); }
  Number add_and_dot(Number a, const Vector &x, const Vector &w) {
    double r = 0;
    check(b200mf_vec_add_and_dot(code(), data_, double(a), x.data_, w.data_, size_, &r, nullptr));
    return Number(r);
  }
  bool all_zero() const { return linfty_norm() == real_type(0); }
  std::size_t memory_consumption() const { return size_ * sizeof(Number) + sizeof(*this); }
  void update_ghost_values() const {}
  void compress(dealii::VectorOperation::values) {}
  void zero_out_ghost_values() const {}
  bool has_ghost_elements() const { return false; }

  void import_from_host(const dealii::Vector<Number> &h) {
    reinit(h.size(), true);
    if (size_) cuda_check(cudaMemcpy(data_, h.begin(), size_ * sizeof(Number), cudaMemcpyHostToDevice));
  }
  void export_to_host(dealii::Vector<Number> &h) const {
    h.reinit(size_);
    if (size_) cuda_check(cudaMemcpy(h.begin(), data_, size_ * sizeof(Number), cudaMemcpyDeviceToHost));
  }

private:
  static int code() { return number_code<Number>(); }
  Number   *data_ = nullptr;
  size_type size_ = 0;
};

// DiagonalMatrix<VectorType>::vmult with the (inverse) diagonal held on the device
template <typename Number>
class DiagonalPreconditioner {
public:
  Vector<Number> &get_vector() { return diag_; }
  const Vector<Number> &get_vector() const { return diag_; }
  void vmult(Vector<Number> &dst, const Vector<Number> &src) const {
    check(b200mf_vec_scale_by(number_code<Number>(), dst.get_values(), diag_.get_values(), src.get_values(),
                              src.size(), nullptr));
  }

private:
  Vector<Number> diag_;
};

// ------------------------------------------------------------------------------------------
template <int dim, typename Number>
class MatrixFree : public b200::MatrixFree<dim, Number> {
public:
  using Base = b200::MatrixFree<dim, Number>;
  struct AdditionalData : Base::AdditionalData {
    // Portable::MatrixFree::AdditionalData::mapping_update_flags is not needed: the engine derives
    // JxW / inverse Jacobians itself (compressed for Cartesian / affine cells)
    //
    // MatrixFree::AdditionalData::mg_level (matrix_free/matrix_free.h: "mg_level"): work on the cells and the
    // level dofs (DoFHandler::distribute_mg_dofs) of one multigrid level instead of the active cells
    unsigned int mg_level = dealii::numbers::invalid_unsigned_int;
    // the CPU MatrixFree's treatment of constrained dofs (MatrixFreeOperators::Base::vmult): they are not read
    // by the cell loop and their rows are the identity; Portable::MatrixFree (false) reads them
    bool eliminate_constrained_dofs = false;
  };

  void reinit(const dealii::Mapping<dim> &mapping, const dealii::DoFHandler<dim> &dof_handler,
              const dealii::AffineConstraints<Number> &constraints, const dealii::Quadrature<1> &quad,
              const AdditionalData &additional_data = AdditionalData()) {
    using namespace dealii;
    const FiniteElement<dim> &fe = dof_handler.get_fe();
    const unsigned int degree = fe.degree, n_q_1d = quad.size();
    // same requirement as AssertThrow(n_q_points_1d >= fe_degree + 1), portable_matrix_free.templates.h:1243
    if (n_q_1d < degree + 1)
      throw Exception(B200MF_ERR_INVALID, "n_q_points_1d >= fe_degree + 1 is required");
    const unsigned int dofs_per_cell = fe.n_dofs_per_cell(), nq = Utilities::pow(n_q_1d, dim);
    internal::MatrixFreeFunctions::ShapeInfo<Number> shape_info(quad, fe);
    const std::vector<unsigned int> &lexicographic_inv = shape_info.lexicographic_numbering;
    internal::MatrixFreeFunctions::HangingNodes<dim> hanging_nodes(dof_handler.get_triangulation());
    const std::shared_ptr<const Utilities::MPI::Partitioner> no_partitioner; // one process

    const auto &tria = dof_handler.get_triangulation();
    const bool on_level = additional_data.mg_level != numbers::invalid_unsigned_int;
    const unsigned int level = additional_data.mg_level;
    n_cells_ = on_level ? tria.n_cells(level) : tria.n_active_cells();
    const types::global_dof_index n_dofs = on_level ? dof_handler.n_dofs(level) : dof_handler.n_dofs();
    // d-linear geometry (MappingQ on cells without a curved manifold): hand over the vertices, the
    // engine classifies Cartesian / affine / general cells and compresses; else the arrays PMF stores
    bool flat = dynamic_cast<const MappingQ<dim> *>(&mapping) != nullptr;
    for (const auto id : tria.get_manifold_ids())
      flat = flat && (id == numbers::flat_manifold_id);
    std::vector<std::uint32_t> l2g((std::size_t)n_cells_ * dofs_per_cell);
    std::vector<std::uint16_t> mask(n_cells_, 0);
    std::vector<double> vertices, inv_jacobian, JxW;
    q_points_.assign((std::size_t)n_cells_ * nq * dim, 0.0);
    if (flat) vertices.resize((std::size_t)n_cells_ * GeometryInfo<dim>::vertices_per_cell * dim);
    else { inv_jacobian.resize((std::size_t)n_cells_ * nq * dim * dim); JxW.resize((std::size_t)n_cells_ * nq); }
    const QGauss<dim> quad_dim(n_q_1d);
    FEValues<dim> fe_values(mapping, fe, Quadrature<dim>(quad),
                            update_quadrature_points | (flat ? update_default : (update_inverse_jacobians | update_JxW_values)));
    std::vector<types::global_dof_index> local_dof_indices(dofs_per_cell), lexicographic(dofs_per_cell);
    std::size_t c = 0;
    // one body for active cells and for the cells of a level (level cells: no hanging-node masks)
    std::vector<typename Triangulation<dim>::cell_iterator> cells;
    if (on_level)
      for (const auto &cell : tria.cell_iterators_on_level(level)) cells.push_back(cell);
    else
      for (const auto &cell : tria.active_cell_iterators()) cells.push_back(cell);
    for (const auto &cell : cells) {
      if (on_level) {
        const typename DoFHandler<dim>::level_cell_iterator level_cell(&tria, cell->level(), cell->index(), &dof_handler);
        level_cell->get_mg_dof_indices(local_dof_indices);
        for (unsigned int i = 0; i < dofs_per_cell; ++i) lexicographic[i] = local_dof_indices[lexicographic_inv[i]];
      } else {
        const typename DoFHandler<dim>::active_cell_iterator active(&tria, cell->level(), cell->index(), &dof_handler);
        active->get_dof_indices(local_dof_indices);
        for (unsigned int i = 0; i < dofs_per_cell; ++i) lexicographic[i] = local_dof_indices[lexicographic_inv[i]];
        internal::MatrixFreeFunctions::ConstraintKinds kinds = internal::MatrixFreeFunctions::ConstraintKinds::unconstrained;
        const ArrayView<internal::MatrixFreeFunctions::ConstraintKinds> view(&kinds, 1);
        hanging_nodes.setup_constraints(active, no_partitioner, {lexicographic_inv}, lexicographic, view);
        mask[c] = static_cast<std::uint16_t>(kinds);
      }
      for (unsigned int i = 0; i < dofs_per_cell; ++i) l2g[c * dofs_per_cell + i] = lexicographic[i];
      if (additional_data.eliminate_constrained_dofs)
        for (unsigned int i = 0; i < dofs_per_cell; ++i)
          if (constraints.is_constrained(lexicographic[i])) l2g[c * dofs_per_cell + i] |= B200MF_L2G_CONSTRAINED;
      cell_index_[std::make_pair(cell->level(), cell->index())] = c;
      fe_values.reinit(cell);
      for (unsigned int q = 0; q < nq; ++q)
        for (unsigned int d = 0; d < dim; ++d) q_points_[(c * nq + q) * dim + d] = fe_values.quadrature_point(q)[d];
      if (flat) {
        const auto v = mapping.get_vertices(cell);
        for (unsigned int k = 0; k < GeometryInfo<dim>::vertices_per_cell; ++k)
          for (unsigned int d = 0; d < dim; ++d) vertices[(c * GeometryInfo<dim>::vertices_per_cell + k) * dim + d] = v[k][d];
      } else {
        for (unsigned int q = 0; q < nq; ++q) {
          JxW[c * nq + q] = fe_values.JxW(q);
          for (unsigned int d = 0; d < dim; ++d)
            for (unsigned int e = 0; e < dim; ++e)
              inv_jacobian[((c * nq + q) * dim + d) * dim + e] = fe_values.inverse_jacobian(q)[d][e];
        }
      }
      ++c;
    }
    std::vector<std::uint32_t> constrained;
    for (types::global_dof_index i = 0; i < n_dofs; ++i)
      if (constraints.is_constrained(i)) constrained.push_back(i);
    bool any_mask = false;
    for (auto m : mask) any_mask = any_mask || m != 0;

    ReinitData d;
    d.degree = degree;
    d.n_q_points_1d = n_q_1d;
    d.n_cells = n_cells_;
    d.n_owned_dofs = n_dofs;
    d.local_to_global = l2g.data();
    d.constraint_mask = any_mask ? mask.data() : nullptr;
    if (flat) d.cell_vertices = vertices.data();
    else { d.inv_jacobian = inv_jacobian.data(); d.JxW = JxW.data(); }
    d.constrained_dofs = constrained.data();
    d.n_constrained_dofs = constrained.size();
    Base::reinit(d, additional_data);
    n_q_ = nq;
    partitioner_ = std::make_shared<Utilities::MPI::Partitioner>(n_dofs);
  }

  // position of a cell (level, index) in this object's cell order
  std::size_t cell_position(int level, int index) const { return cell_index_.at(std::make_pair(level, index)); }
  std::size_t n_cells() const { return n_cells_; }

  void initialize_dof_vector(Vector<Number> &vec) const { vec.reinit(this->n_local_dofs()); }
  const std::shared_ptr<const dealii::Utilities::MPI::Partitioner> &get_vector_partitioner() const { return partitioner_; }

  // coefficient(q-point) evaluated for every (cell, q) in local_q_point_id order -> device array
  void evaluate_coefficients(const std::function<Number(const dealii::Point<dim> &)> &f, Vector<Number> &coef) const {
    dealii::Vector<Number> host((std::size_t)n_cells_ * n_q_);
    for (std::size_t i = 0; i < host.size(); ++i) {
      dealii::Point<dim> pt;
      for (unsigned int d = 0; d < dim; ++d) pt[d] = q_points_[i * dim + d];
      host[i] = f(pt);
    }
    coef.import_from_host(host);
  }
  unsigned int n_q_points_per_cell() const { return n_q_; }

private:
  std::size_t n_cells_ = 0;
  unsigned int n_q_ = 0;
  std::map<std::pair<int, int>, std::size_t> cell_index_;
  std::vector<double> q_points_;
  std::shared_ptr<const dealii::Utilities::MPI::Partitioner> partitioner_;
};

// ------------------------------------------------------------------------------------------
// PreconditionMG + Multigrid + MGTransferMatrixFree + mg::SmootherRelaxation<PreconditionChebyshev> +
// MGCoarseGridApplySmoother of step-37 (examples/step-37/step-37.cc:950-1060) in one object on the device:
//   reinit(mapping, dof_handler, mg_constrained_dofs, quad, coefficient)   levels 0 .. n_global_levels-1 from
//                               DoFHandler::distribute_mg_dofs + MGConstrainedDoFs::get_boundary_indices
//   vmult(dst, src)             one V-cycle (PreconditionMG::vmult), usable as the preconditioner of deal.II's
//                               own SolverCG on dealii_adapter::Vector
// LevelNumber = float under a double CG is step-37's configuration.
template <int dim, typename LevelNumber>
class PreconditionMG {
public:
  struct AdditionalData {
    unsigned int smoother_degree = 5;       // PreconditionChebyshev::AdditionalData of step-37.cc:965-975
    double smoothing_range = 15.;
    unsigned int eig_cg_n_iterations = 10;
    double coarse_tolerance = 1e-3;         // level 0: smoothing_range = 1e-3, degree = invalid_unsigned_int
  };
  PreconditionMG() = default;
  PreconditionMG(const PreconditionMG &) = delete;
  ~PreconditionMG() { clear(); }
  void clear() {
    if (mg_) b200mf_mg_destroy(mg_);
    mg_ = nullptr;
    levels_.clear();
    coefficients_.clear();
  }

  void reinit(const dealii::Mapping<dim> &mapping, const dealii::DoFHandler<dim> &dof_handler,
              const dealii::MGConstrainedDoFs &mg_constrained_dofs, const dealii::Quadrature<1> &quad,
              const std::function<LevelNumber(const dealii::Point<dim> &)> &coefficient = {},
              const AdditionalData &data = AdditionalData()) {
    using namespace dealii;
    clear();
    const auto &tria = dof_handler.get_triangulation();
    const unsigned int n_levels = tria.n_global_levels();
    std::vector<const b200mf_setup *> setups;
    std::vector<b200mf_operator> operators;
    std::vector<std::vector<std::uint32_t>> children(n_levels > 0 ? n_levels - 1 : 0);
    std::vector<const std::uint32_t *> child_ptrs;
    for (unsigned int level = 0; level < n_levels; ++level) {
      AffineConstraints<LevelNumber> level_constraints;
      for (const types::global_dof_index i : mg_constrained_dofs.get_boundary_indices(level))
        level_constraints.constrain_dof_to_zero(i);
      level_constraints.close();
      typename MatrixFree<dim, LevelNumber>::AdditionalData ad;
      ad.mg_level = level;
      ad.eliminate_constrained_dofs = true;
      levels_.emplace_back(new MatrixFree<dim, LevelNumber>());
      levels_.back()->reinit(mapping, dof_handler, level_constraints, quad, ad);
      coefficients_.emplace_back(new Vector<LevelNumber>());
      if (coefficient) levels_.back()->evaluate_coefficients(coefficient, *coefficients_.back());
      setups.push_back(levels_.back()->get_setup());
      operators.push_back(b200mf_operator{coefficient ? coefficients_.back()->get_values() : nullptr, nullptr, 1.0, 0.0});
    }
    // the children of every cell of level l, as positions in level l+1's cell order, x fastest
    // (GeometryInfo<dim>::child_cell_on_face ordering = the lexicographic child numbering of hypercubes)
    for (unsigned int level = 0; level + 1 < n_levels; ++level) {
      auto &table = children[level];
      table.assign(levels_[level]->n_cells() << dim, 0);
      for (const auto &cell : tria.cell_iterators_on_level(level)) {
        if (!cell->has_children())
          throw Exception(B200MF_ERR_UNSUPPORTED, "PreconditionMG: globally refined meshes only (a level cell without children)");
        const std::size_t c = levels_[level]->cell_position(cell->level(), cell->index());
        for (unsigned int k = 0; k < GeometryInfo<dim>::max_children_per_cell; ++k)
          table[(c << dim) + k] = levels_[level + 1]->cell_position(cell->child(k)->level(), cell->child(k)->index());
      }
      child_ptrs.push_back(table.data());
    }
    b200mf_mg_desc d{};
    d.n_levels = n_levels;
    d.levels = setups.data();
    d.operators = operators.data();
    d.child_cells = child_ptrs.empty() ? nullptr : child_ptrs.data();
    d.smoother_degree = data.smoother_degree;
    d.smoothing_range = data.smoothing_range;
    d.eig_cg_n_iterations = data.eig_cg_n_iterations;
    d.coarse_tolerance = data.coarse_tolerance;
    d.safety_factor = 1.2;
    check(b200mf_mg_create(&d, &mg_, nullptr));
  }

  template <typename Number>
  void vmult(Vector<Number> &dst, const Vector<Number> &src) const {
    check(b200mf_mg_vcycle(mg_, number_code<Number>(), dst.get_values(), src.get_values(), nullptr));
  }
  b200mf_mg_level_info level_info(unsigned int level) const {
    b200mf_mg_level_info info;
    check(b200mf_mg_get_level_info(mg_, level, &info));
    return info;
  }
  unsigned int n_levels() const { return levels_.size(); }

private:
  b200mf_mg *mg_ = nullptr;
  std::vector<std::unique_ptr<MatrixFree<dim, LevelNumber>>> levels_;
  std::vector<std::unique_ptr<Vector<LevelNumber>>> coefficients_;
};

} // namespace dealii_adapter
} // namespace b200
#endif
