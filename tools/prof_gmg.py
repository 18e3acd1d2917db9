"""One multigrid-preconditioned CG solve for ncu launch lists:
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv \
       python tools/prof_gmg.py 4 6 f32
(without ncu it prints the solve time)."""
import sys
import time

import torch

import dealii_b200

degree, refinements = int(sys.argv[1]), int(sys.argv[2])
levels = sys.argv[3] if len(sys.argv) > 3 else "f32"
numbering = sys.argv[4] if len(sys.argv) > 4 else "default"
mg = dealii_b200.GeometricMultigrid.for_hyper_cube(3, degree, refinements, number=levels, numbering=numbering)
mesh = dealii_b200.HyperCubeMesh(3, degree, refinements=refinements, dirichlet_boundary=True, mark_constrained_l2g=True,
                                 numbering=numbering)
mf = dealii_b200.MatrixFree("f64")
mf.reinit_from_mesh(mesh)
A = dealii_b200.LaplaceOperator(mf)
b = mf.initialize_dof_vector()
b[:] = 1.0
mf.set_constrained_values(0.0, b)
tol = 1e-6 * float(b.norm())
for rep in range(3):
    x = mf.initialize_dof_vector()
    torch.cuda.synchronize()
    if rep == 2:
        torch.cuda.profiler.start()
    t0 = time.time()
    control = dealii_b200.SolverControl(100, tol)
    dealii_b200.SolverCG(control).solve(A, x, b, mg)
    torch.cuda.synchronize()
    dt = time.time() - t0
    if rep == 2:
        torch.cuda.profiler.stop()
print(f"Q{degree} r{refinements} levels {levels} numbering {numbering}: {mesh.n_dofs} dofs, {control.last_step()} iterations, {dt * 1e3:.2f} ms")
