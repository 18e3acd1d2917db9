"""Device time of compute_diagonal for a few cases: degree:refinements:number[:amp]."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, dealii_b200
for spec in sys.argv[1:]:
    f = spec.split(":")
    degree, ref, number = int(f[0]), int(f[1]), f[2]
    amp = float(f[3]) if len(f) > 3 else 0.0
    mesh = dealii_b200.HyperCubeMesh(3, degree, refinements=ref, deformation_amplitude=amp)
    mf = dealii_b200.MatrixFree(number).reinit_from_mesh(mesh)
    op = dealii_b200.LaplaceOperator(mf)
    op.compute_diagonal(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); op.compute_diagonal(); e1.record(); torch.cuda.synchronize()
    print(json.dumps(dict(case=spec, n_dofs=mf.n_owned, ms=round(e0.elapsed_time(e1), 3))), flush=True)
