"""Times MGTransfer prolongate / restrict_and_add between the two finest levels."""
import sys
import torch
import dealii_b200

degree, refinements = int(sys.argv[1]), int(sys.argv[2])
number = sys.argv[3] if len(sys.argv) > 3 else "f32"
mg = dealii_b200.GeometricMultigrid.for_hyper_cube(3, degree, refinements, number=number, min_level=refinements - 1)
top = mg.n_levels() - 1
fine, coarse = mg.level_operators[top].mf, mg.level_operators[top - 1].mf
u, v = coarse.initialize_dof_vector(), fine.initialize_dof_vector()
u.normal_(); v.normal_()
out_f, out_c = torch.zeros_like(v), torch.zeros_like(u)
for name, fn in (("prolongate", lambda: mg.prolongate(top, out_f, u)), ("restrict_and_add", lambda: mg.restrict_and_add(top, out_c, v))):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    nbytes = (v.numel() + u.numel()) * v.element_size()
    print(f"{name}: {ms:.4f} ms  ({nbytes / ms / 1e6:.0f} GB/s of vector bytes, {v.numel() / ms / 1e6:.1f} GDoF/s)")
