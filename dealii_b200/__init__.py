"""dealii_b200 -- host-side mirror of deal.II's Portable::MatrixFree hot path on top of
the C ABI of libb200mf.so (hand-written sm_100a CUDA).  PyTorch is used for device memory
and streams only."""
from .matrix_free import (DiagonalMatrix, HelmholtzOperator, HyperCubeMesh, LaplaceOperator,  # noqa: F401
                          MatrixFree, MatrixFreeOperator, PreconditionChebyshev, SolverCG,
                          SolverControl)
from ._lib import B200MFError  # noqa: F401
from .multigrid import GeometricMultigrid  # noqa: F401
