"""ncu driver: a few vmults on the lexicographically numbered mesh (strided brick path)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, dealii_b200
degree, ref, number = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
mesh = dealii_b200.HyperCubeMesh(3, degree, refinements=ref, numbering="lexicographic")
mf = dealii_b200.MatrixFree(number).reinit_from_mesh(mesh)
mf.select_brick_path(0)
op = dealii_b200.LaplaceOperator(mf)
x = torch.rand(mf.n_owned, dtype=mf.torch_dtype, device="cuda")
y = mf.initialize_dof_vector()
for _ in range(4):
    op.vmult(y, x)
torch.cuda.synchronize()
