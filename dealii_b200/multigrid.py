"""Geometric multigrid over the C ABI (include/b200mf.h, b200mf_mg_*), named after the reference classes
it stands in for on the path of examples/step-37:

  GeometricMultigrid      Multigrid + PreconditionMG + MGTransferMatrixFree +
                          mg::SmootherRelaxation<PreconditionChebyshev> + MGCoarseGridApplySmoother
                          (multigrid/multigrid.h, mg_transfer_matrix_free.h, mg_smoother.h, mg_coarse.h)
  .vmult(dst, src)        PreconditionMG::vmult: one V-cycle
  .prolongate / .restrict_and_add   MGTransferMatrixFree
  SolverCG.solve(A, x, b, GeometricMultigrid)   (matrix_free.SolverCG dispatches here)

Levels are MatrixFreeOperator objects on globally refined HyperCubeMesh levels (Morton cell order, the order
refine_global produces), coarsest first.
"""
import ctypes as C

import torch

from . import _lib as L
from .matrix_free import HyperCubeMesh, LaplaceOperator, MatrixFree, _ptr, _stream


class GeometricMultigrid:
    def __init__(self, level_operators, smoother_degree=5, smoothing_range=15.0, eig_cg_n_iterations=10,
                 coarse_tolerance=1e-3, safety_factor=1.2):
        self.level_operators = list(level_operators)   # keeps setups and coefficient tensors alive
        n = len(self.level_operators)
        self._lib = L.load()
        d = L.MgDesc()
        d.n_levels = n
        self._levels = (C.c_void_p * n)(*[op.mf._h for op in self.level_operators])
        self._ops = (L.Operator * n)(*[op.op for op in self.level_operators])
        d.levels = C.cast(self._levels, C.POINTER(C.c_void_p))
        d.operators = C.cast(self._ops, C.POINTER(L.Operator))
        d.child_cells = None
        d.smoother_degree, d.smoothing_range = smoother_degree, smoothing_range
        d.eig_cg_n_iterations, d.coarse_tolerance = eig_cg_n_iterations, coarse_tolerance
        d.safety_factor = safety_factor
        self._h = C.c_void_p()
        L.check(self._lib.b200mf_mg_create(C.byref(d), C.byref(self._h), _stream()))
        self.number = self.level_operators[0].mf.number

    @classmethod
    def for_hyper_cube(cls, dim, degree, refinements, number="f64", coefficient=None, min_level=0,
                       left=0.0, right=1.0, device="cuda:0", deformation_amplitude=0.0, numbering="default",
                       **kwargs):
        """Levels min_level..refinements of GridGenerator::hyper_cube + refine_global with homogeneous
        Dirichlet boundary; coefficient(x) (torch, [n_points, dim] -> [n_points]) makes a variable
        gradient coefficient evaluated at each level's quadrature points (step-37's Coefficient).
        numbering="lexicographic": every level numbered by DoFRenumbering::lexicographic (the system operator must
        use the same numbering on the finest mesh); the level operators then run the strided brick kernel."""
        ops = []
        for level in range(min_level, refinements + 1):
            mesh = HyperCubeMesh(dim, degree, refinements=level, left=left, right=right, dirichlet_boundary=True,
                                 mark_constrained_l2g=True, deformation_amplitude=deformation_amplitude,
                                 numbering=numbering)
            mf = MatrixFree(number, device)
            mf.reinit_from_mesh(mesh)
            coef = mf.evaluate_coefficients(coefficient) if coefficient is not None else None
            ops.append(LaplaceOperator(mf, coef))
        return cls(ops, **kwargs)

    def n_levels(self):
        return len(self.level_operators)

    def level_info(self, level):
        info = L.MgLevelInfo()
        L.check(self._lib.b200mf_mg_get_level_info(self._h, level, C.byref(info)))
        return info

    def inverse_diagonal(self, level):
        info = self.level_info(level)
        mf = self.level_operators[level].mf
        out = mf.initialize_dof_vector()
        code = L.F64 if out.dtype == torch.float64 else L.F32
        L.check(self._lib.b200mf_vec_equ(code, _ptr(out), 1.0, C.c_void_p(info.inverse_diagonal), 0.0, None,
                                         out.numel(), _stream()))
        return out

    def prolongate(self, to_level, dst, src):
        L.check(self._lib.b200mf_mg_prolongate(self._h, to_level, _ptr(dst), _ptr(src), _stream()))

    def restrict_and_add(self, from_level, dst, src):
        L.check(self._lib.b200mf_mg_restrict_and_add(self._h, from_level, _ptr(dst), _ptr(src), _stream()))

    def vmult(self, dst, src):
        code = L.F64 if src.dtype == torch.float64 else L.F32
        L.check(self._lib.b200mf_mg_vcycle(self._h, code, _ptr(dst), _ptr(src), _stream()))

    def solve(self, A, x, b, tolerance, max_steps=100):
        """SolverCG(SolverControl(max_steps, tolerance)).solve(A, x, b, PreconditionMG)."""
        res = L.SolverResult()
        code = self._lib.b200mf_mg_cg_solve(self._h, A.mf._h, C.byref(A.op), float(tolerance), int(max_steps),
                                            _ptr(x), _ptr(b), C.byref(res), _stream())
        return code, res

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                self._lib.b200mf_mg_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass
