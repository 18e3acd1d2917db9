"""Device timing of the cell-loop kernel for a list of cases: degree:ref:number:amp ..."""
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
    x = torch.rand(mf.n_owned, dtype=mf.torch_dtype, device="cuda")
    y = mf.initialize_dof_vector()
    for _ in range(3):
        mf.cell_loop(op.op, x, y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        mf.cell_loop(op.op, x, y)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    for _ in range(3):
        op.vmult(y, x)
    e0.record()
    for _ in range(10):
        op.vmult(y, x)
    e1.record(); torch.cuda.synchronize()
    ms_v = e0.elapsed_time(e1) / 10
    print(json.dumps(dict(case=spec, n_dofs=mf.n_owned, kind=int(mf.info.cell_kind), bricks=int(mf.info.n_bricks),
                          ms_cell_loop=round(ms, 4), gdofs_cell_loop=round(mf.n_owned / ms / 1e6, 2),
                          ms_vmult=round(ms_v, 4), gdofs_vmult=round(mf.n_owned / ms_v / 1e6, 2))), flush=True)
    del mf, mesh, op, x, y
