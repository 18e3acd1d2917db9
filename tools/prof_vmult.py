"""Tiny driver for ncu captures: a few vmults of one configuration."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, dealii_b200
dim, degree, ref, number = 3, int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
amp = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
mesh = dealii_b200.HyperCubeMesh(dim, degree, refinements=ref, deformation_amplitude=amp)
mf = dealii_b200.MatrixFree(number).reinit_from_mesh(mesh)
op = dealii_b200.LaplaceOperator(mf)
x = torch.rand(mf.n_owned, dtype=mf.torch_dtype, device="cuda")
y = mf.initialize_dof_vector()
for _ in range(4):
    op.vmult(y, x)
torch.cuda.synchronize()
