"""Host-side mirror of the reference interface for the hot path, over the C ABI.

Names, argument meaning and error behaviour follow the reference so that the parity tests
read like deal.II's own:
  MatrixFree            Portable::MatrixFree<dim,Number>  (matrix_free/portable_matrix_free.h:185)
  LaplaceOperator /
  HelmholtzOperator     the operator classes of tests/performance/timing_matrix_free_kokkos.cc
                        and examples/step-64/step-64.cc:225-370 (vmult, compute_diagonal, ...)
  SolverCG, SolverControl, DiagonalMatrix, PreconditionChebyshev
                        lac/solver_cg.h, lac/solver_control.h, lac/diagonal_matrix.h,
                        lac/precondition.h
Vectors are torch CUDA tensors of length n_owned + n_ghost (LA::d::Vector layout).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L

_DT = {"f64": (L.F64, torch.float64, np.float64), "f32": (L.F32, torch.float32, np.float32)}


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _npptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class HyperCubeMesh:
    """Synthetic hyper_cube mesh + FE_Q numbering (b200mf_mesh_*), standing in for
    GridGenerator::hyper_cube + refine_global + DoFHandler::distribute_dofs."""

    def __init__(self, dim, degree, refinements=None, subdivisions=None, left=0.0, right=1.0,
                 deformation_amplitude=0.0, dirichlet_boundary=False, mark_constrained_l2g=False,
                 numbering="default"):
        lib = L.load()
        d = L.MeshDesc()
        d.dim, d.degree = dim, degree
        if refinements is not None:
            d.cells_per_direction, d.cell_order = 2 ** refinements, L.MESH_MORTON
        else:
            d.cells_per_direction, d.cell_order = subdivisions, L.MESH_LEXICOGRAPHIC
        d.left, d.right = left, right
        d.deformation = L.DEFORM_SINE if deformation_amplitude != 0.0 else L.DEFORM_NONE
        d.deformation_amplitude = deformation_amplitude
        d.dirichlet_boundary = int(dirichlet_boundary)
        d.mark_constrained_l2g = int(mark_constrained_l2g)
        d.dof_numbering = 1 if numbering == "lexicographic" else 0   # DoFRenumbering::lexicographic
        self._h = C.c_void_p()
        L.check(lib.b200mf_mesh_create(C.byref(d), C.byref(self._h)))
        v = L.MeshView()
        L.check(lib.b200mf_mesh_view_get(self._h, C.byref(v)))
        self.dim, self.degree = dim, degree
        self.n_cells, self.n_dofs = int(v.n_cells), int(v.n_dofs)
        self.dofs_per_cell = int(v.dofs_per_cell)
        self._view = v

    @property
    def l2g(self):
        return np.ctypeslib.as_array(self._view.local_to_global,
                                     shape=(self.n_cells, self.dofs_per_cell))

    @property
    def cell_vertices(self):
        return np.ctypeslib.as_array(self._view.cell_vertices,
                                     shape=(self.n_cells, 2 ** self.dim, self.dim))

    @property
    def boundary_dofs(self):
        n = int(self._view.n_boundary_dofs)
        if n == 0:
            return np.zeros(0, dtype=np.uint32)
        return np.ctypeslib.as_array(self._view.boundary_dofs, shape=(n,))

    def __del__(self):
        try:
            if self._h:
                L.load().b200mf_mesh_destroy(self._h)
                self._h = None
        except Exception:
            pass


class MatrixFree:
    """Portable::MatrixFree<dim,Number>: reinit / cell_loop / copy_constrained_values /
    set_constrained_values / initialize_dof_vector (portable_matrix_free.h:480-700)."""

    def __init__(self, number="f64", device="cuda:0"):
        self.number = number
        self._code, self.torch_dtype, self.np_dtype = _DT[number]
        self.device = torch.device(device)
        self._h = None
        self._lib = L.load()

    # -- reinit from explicit arrays (what a deal.II-side adapter would pass)
    def reinit(self, dim, degree, local_to_global, cell_vertices=None, constrained_dofs=None,
               constraint_mask=None, n_owned_dofs=None, n_ghost_dofs=0, inv_jacobian=None,
               JxW=None, n_q_points_1d=None, n_cells_interior=0):
        if not torch.cuda.is_available():
            raise L.B200MFError(L.ERR_CUDA, "no CUDA device: the engine has no CPU fallback")
        torch.cuda.set_device(self.device)
        self.clear()
        l2g = np.ascontiguousarray(local_to_global, dtype=np.uint32)
        d = L.SetupDesc()
        d.dim, d.degree = dim, degree
        d.n_q_points_1d = degree + 1 if n_q_points_1d is None else n_q_points_1d
        d.number = self._code
        d.n_cells = l2g.shape[0]
        if n_owned_dofs is None:
            n_owned_dofs = int((l2g & 0x7FFFFFFF).max()) + 1 if l2g.size else 0
        d.n_owned_dofs, d.n_ghost_dofs = n_owned_dofs, n_ghost_dofs
        d.local_to_global = _npptr(l2g)
        keep = [l2g]
        if constraint_mask is not None:
            cm = np.ascontiguousarray(constraint_mask, dtype=np.uint16)
            keep.append(cm)
            d.constraint_mask = _npptr(cm)
        if cell_vertices is not None:
            cv = np.ascontiguousarray(cell_vertices, dtype=np.float64)
            keep.append(cv)
            d.geometry, d.cell_vertices = L.GEOMETRY_Q1_VERTICES, _npptr(cv)
        else:
            ij = np.ascontiguousarray(inv_jacobian, dtype=np.float64)
            jw = np.ascontiguousarray(JxW, dtype=np.float64)
            keep += [ij, jw]
            d.geometry, d.inv_jacobian, d.JxW = L.GEOMETRY_JACOBIANS, _npptr(ij), _npptr(jw)
        if constrained_dofs is not None and len(constrained_dofs):
            cd = np.ascontiguousarray(constrained_dofs, dtype=np.uint32)
            keep.append(cd)
            d.constrained_dofs, d.n_constrained_dofs = _npptr(cd), len(cd)
        d.n_cells_interior = n_cells_interior
        h = C.c_void_p()
        L.check(self._lib.b200mf_setup_create(C.byref(d), C.byref(h)))
        self._h = h
        self._fill_info()
        return self

    def reinit_from_mesh(self, mesh):
        if not torch.cuda.is_available():
            raise L.B200MFError(L.ERR_CUDA, "no CUDA device: the engine has no CPU fallback")
        torch.cuda.set_device(self.device)
        self.clear()
        h = C.c_void_p()
        L.check(self._lib.b200mf_setup_create_from_mesh(mesh._h, self._code, C.byref(h)))
        self._h = h
        self._fill_info()
        return self

    def _fill_info(self):
        info = L.SetupInfo()
        L.check(self._lib.b200mf_setup_get_info(self._h, C.byref(info)))
        self.info = info
        self.dim, self.degree = info.dim, info.degree
        self.n_cells = int(info.n_cells)
        self.n_owned, self.n_ghost = int(info.n_owned_dofs), int(info.n_ghost_dofs)
        self.n_q_points = info.n_q_points_1d ** info.dim

    def clear(self):
        if self._h is not None:
            self._lib.b200mf_setup_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.clear()
        except Exception:
            pass

    # -- Portable::MatrixFree API
    def initialize_dof_vector(self):
        return torch.zeros(self.n_owned + self.n_ghost, dtype=self.torch_dtype, device=self.device)

    def get_quadrature_points(self):
        out = np.empty((self.n_cells, self.n_q_points, self.dim), dtype=np.float64)
        L.check(self._lib.b200mf_get_quadrature_points(self._h, _npptr(out)))
        return out

    def evaluate_coefficients(self, functor):
        """functor(points[(m, dim)]) -> (m,) evaluated at every quadrature point; returns the
        device coefficient array in local_q_point_id order
        (Portable::MatrixFree::evaluate_coefficients, portable_matrix_free.h:585)."""
        q = self.get_quadrature_points().reshape(-1, self.dim)
        vals = np.asarray(functor(q), dtype=self.np_dtype).reshape(-1)
        return torch.from_numpy(vals).to(self.device)

    def cell_loop(self, op, src, dst):
        L.check(self._lib.b200mf_cell_loop(self._h, C.byref(op), _ptr(dst), _ptr(src), _stream()))

    def vmult(self, op, dst, src):
        L.check(self._lib.b200mf_vmult(self._h, C.byref(op), _ptr(dst), _ptr(src), _stream()))

    def vmult_prepare(self, op, dst):
        L.check(self._lib.b200mf_vmult_prepare(self._h, C.byref(op), _ptr(dst), _stream()))

    def vmult_range(self, op, dst, src, cell_begin, cell_end, dot_ptr=None):
        """One piece of a vmult (dst zeroed by the caller before the first piece)."""
        L.check(self._lib.b200mf_vmult_range(self._h, C.byref(op), _ptr(dst), _ptr(src), cell_begin,
                                             cell_end, C.c_void_p(dot_ptr) if dot_ptr else None, _stream()))

    def enable_bulk(self, enable=True):
        """A/B switch between the bulk brick tables and the per-node index maps (tests, bench);
        returns whether the setup has bulk tables."""
        return bool(self._lib.b200mf_setup_enable_bulk(self._h, int(bool(enable))))

    def enable_strided(self, enable=True):
        """A/B switch for the computed-index bricks; returns whether the setup has them."""
        return bool(self._lib.b200mf_setup_enable_strided(self._h, int(bool(enable))))

    def select_brick_path(self, path):
        """0 = index maps + memset + atomics, 1 = coloured launches, 2 = bulk tables; returns the
        active path or -1 if the setup does not have the requested one."""
        return int(self._lib.b200mf_setup_select_brick_path(self._h, int(path)))

    def bulk_info(self):
        info = L.BulkInfo()
        L.check(self._lib.b200mf_setup_get_bulk_info(self._h, C.byref(info)))
        return {k: getattr(info, k) for k, _ in L.BulkInfo._fields_}

    def copy_constrained_values(self, src, dst):
        L.check(self._lib.b200mf_copy_constrained_values(self._h, _ptr(dst), _ptr(src), _stream()))

    def set_constrained_values(self, value, dst):
        L.check(self._lib.b200mf_set_constrained_values(self._h, _ptr(dst), float(value), _stream()))

    def compute_diagonal(self, op, diag):
        L.check(self._lib.b200mf_compute_diagonal(self._h, C.byref(op), _ptr(diag), _stream()))


class MatrixFreeOperator:
    """(c_grad grad u, grad v) + (c_mass u, v): vmult / compute_diagonal / m / initialize_dof_vector."""

    def __init__(self, matrix_free, grad_coefficient=None, mass_coefficient=None,
                 grad_constant=1.0, mass_constant=0.0):
        self.mf = matrix_free
        self._gc, self._mc = grad_coefficient, mass_coefficient   # keep tensors alive
        self.op = L.Operator(_ptr(grad_coefficient), _ptr(mass_coefficient),
                             float(grad_constant), float(mass_constant))
        self.inverse_diagonal = None

    def m(self):
        return self.mf.n_owned

    def initialize_dof_vector(self):
        return self.mf.initialize_dof_vector()

    def vmult(self, dst, src):
        self.mf.vmult(self.op, dst, src)

    def vmult_host(self, dst_host, src_host):
        L.check(self.mf._lib.b200mf_vmult_host(self.mf._h, C.byref(self.op), _npptr(dst_host),
                                               _npptr(src_host)))

    def vmult_host_batch(self, dst_hosts, src_hosts):
        """dst_hosts[k] = A src_hosts[k] for lists of (pinned) host numpy arrays, copies pipelined."""
        n = len(src_hosts)
        dp = (C.c_void_p * n)(*[d.ctypes.data for d in dst_hosts])
        sp = (C.c_void_p * n)(*[s.ctypes.data for s in src_hosts])
        L.check(self.mf._lib.b200mf_vmult_host_batch(self.mf._h, C.byref(self.op), n, dp, sp))

    def compute_diagonal(self):
        """examples/step-64/step-64.cc:339-368: diagonal by MatrixFreeTools::compute_diagonal,
        then inverted in place; returns the DiagonalMatrix."""
        diag = self.mf.initialize_dof_vector()
        self.mf.compute_diagonal(self.op, diag)
        self.diagonal = diag.clone()
        self.inverse_diagonal = DiagonalMatrix(1.0 / diag)
        return self.inverse_diagonal

    def get_matrix_diagonal_inverse(self):
        return self.inverse_diagonal


class LaplaceOperator(MatrixFreeOperator):
    def __init__(self, matrix_free, coefficient=None):
        super().__init__(matrix_free, grad_coefficient=coefficient)


class HelmholtzOperator(MatrixFreeOperator):
    def __init__(self, matrix_free, coefficient):
        if isinstance(coefficient, (int, float)):
            super().__init__(matrix_free, mass_constant=float(coefficient))
        else:
            super().__init__(matrix_free, mass_coefficient=coefficient)


class DiagonalMatrix:
    """DiagonalMatrix<VectorType> (lac/diagonal_matrix.h): holds the (inverse) diagonal."""

    def __init__(self, vector):
        self.vector = vector

    def get_vector(self):
        return self.vector

    def vmult(self, dst, src):
        lib = L.load()
        code = L.F64 if src.dtype == torch.float64 else L.F32
        L.check(lib.b200mf_vec_scale_by(code, _ptr(dst), _ptr(self.vector), _ptr(src),
                                        src.numel(), _stream()))


class PreconditionChebyshev:
    """PreconditionChebyshev<Operator, Vector, DiagonalMatrix>::AdditionalData holder
    (lac/precondition.h:2121-2175); the polynomial runs inside b200mf_cg_solve."""

    def __init__(self, degree=1, smoothing_range=0.0, eig_cg_n_iterations=8, preconditioner=None,
                 safety_factor=1.2, max_eigenvalue=1.0, constraints=True):
        """constraints=True: AdditionalData::constraints = the operator's constraints (their entries of the Lanczos
        start vector are zeroed); False: the reference's default, an empty AffineConstraints."""
        self.degree, self.smoothing_range = degree, smoothing_range
        self.eig_cg_n_iterations, self.preconditioner = eig_cg_n_iterations, preconditioner
        self.safety_factor, self.max_eigenvalue = safety_factor, max_eigenvalue
        self.constraints = constraints


class SolverControl:
    """SolverControl(max_steps, tolerance) (lac/solver_control.h)."""

    def __init__(self, max_steps=100, tolerance=1e-10):
        self.max_steps, self.tolerance = max_steps, tolerance
        self._last_step, self._last_value = 0, 0.0

    def last_step(self):
        return self._last_step

    def last_value(self):
        return self._last_value


class SolverCG:
    """SolverCG<VectorType>(SolverControl&).solve(A, x, b, preconditioner) (lac/solver_cg.h:1391);
    raises B200MFError(ERR_NOCONVERGENCE) like SolverControl::NoConvergence."""

    def __init__(self, control):
        self.control = control
        self.result = None

    def solve(self, A, x, b, preconditioner=None):
        mf = A.mf
        if hasattr(preconditioner, "solve") and hasattr(preconditioner, "level_operators"):
            # PreconditionMG: the V-cycle of dealii_b200.multigrid.GeometricMultigrid
            code, res = preconditioner.solve(A, x, b, self.control.tolerance, self.control.max_steps)
            self.result = res
            self.control._last_step, self.control._last_value = res.iterations, res.residual
            L.check(code)
            return res
        sd = L.SolverDesc()
        sd.tolerance, sd.max_iterations = self.control.tolerance, self.control.max_steps
        keep = None
        if preconditioner is None:
            sd.preconditioner = L.PRECOND_NONE
        elif isinstance(preconditioner, DiagonalMatrix):
            sd.preconditioner = L.PRECOND_JACOBI
            keep = preconditioner.vector
            sd.inverse_diagonal = _ptr(keep)
        elif isinstance(preconditioner, PreconditionChebyshev):
            sd.preconditioner = L.PRECOND_CHEBYSHEV
            keep = preconditioner.preconditioner.vector
            sd.inverse_diagonal = _ptr(keep)
            sd.chebyshev_degree = preconditioner.degree
            sd.smoothing_range = preconditioner.smoothing_range
            sd.eig_cg_n_iterations = preconditioner.eig_cg_n_iterations
            sd.safety_factor = preconditioner.safety_factor
            sd.max_eigenvalue = preconditioner.max_eigenvalue
            sd.eig_keep_constrained_entries = 0 if preconditioner.constraints else 1
        else:
            raise TypeError("unsupported preconditioner")
        res = L.SolverResult()
        code = mf._lib.b200mf_cg_solve(mf._h, C.byref(A.op), C.byref(sd), _ptr(x), _ptr(b),
                                       C.byref(res), _stream())
        self.result = res
        self.control._last_step, self.control._last_value = res.iterations, res.residual
        L.check(code)
        return res
