import sys, time
import numpy as np, torch
sys.path.insert(0, "tests")
from test_multigrid_gpu import hierarchy, system_operator, unit_rhs
dim, p, r = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
number = sys.argv[4] if len(sys.argv) > 4 else "f64"
t0 = time.time()
mg = hierarchy(dim, p, r, number, True)
torch.cuda.synchronize(); print("create", time.time() - t0, flush=True)
for l in range(mg.n_levels()):
    i = mg.level_info(l)
    print(l, i.n_dofs, i.eig_min, i.eig_max, i.degree, i.eig_cg_iterations, flush=True)
mesh, mf, A = system_operator(dim, p, r, True)
b = unit_rhs(mesh, mf)
z = torch.zeros_like(b)
t0 = time.time()
mg.vmult(z, b)
torch.cuda.synchronize(); print("vcycle", time.time() - t0, float(z.norm()), flush=True)
