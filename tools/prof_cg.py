"""Tiny driver for ncu captures of the CG iteration kernels: a few Jacobi-CG iterations, one GPU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, dealii_b200
from dealii_b200.distributed import DistributedMatrixFree, PartitionedHyperCubeMesh, solve_cg
degree, ref = int(sys.argv[1]), int(sys.argv[2])
mesh = PartitionedHyperCubeMesh(3, degree, ref, 1, 0, dirichlet_boundary=True)
dmf = DistributedMatrixFree(mesh, "f64", "cuda:0")
A = dealii_b200.LaplaceOperator(dmf.mf)
inv = dmf.compute_diagonal(A.op)
b = dmf.initialize_dof_vector(); b[:mesh.n_owned] = 1.0
dmf.mf.set_constrained_values(0.0, b)
x = dmf.initialize_dof_vector()
solve_cg(dmf, A.op, x, b, inv, 1e-300, 6)
torch.cuda.synchronize()
