"""Quick device timing of vmult for a few (degree, size, number, geometry) cases."""
import json, sys, time
import torch
import dealii_b200

def run(dim, degree, refinements, number, amp=0.0, reps=10):
    t0 = time.time()
    mesh = dealii_b200.HyperCubeMesh(dim, degree, refinements=refinements, deformation_amplitude=amp)
    t1 = time.time()
    mf = dealii_b200.MatrixFree(number).reinit_from_mesh(mesh)
    t2 = time.time()
    op = dealii_b200.LaplaceOperator(mf)
    x = torch.rand(mf.n_owned, dtype=mf.torch_dtype, device="cuda")
    y = mf.initialize_dof_vector()
    for _ in range(3):
        op.vmult(y, x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        op.vmult(y, x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    e0.record()
    for _ in range(reps):
        mf.cell_loop(op.op, x, y)
    e1.record(); torch.cuda.synchronize()
    ms_k = e0.elapsed_time(e1) / reps
    print(json.dumps(dict(dim=dim, degree=degree, n_dofs=mf.n_owned, number=number, kind=int(mf.info.cell_kind),
                          ms_vmult=round(ms, 4), ms_cell_kernel=round(ms_k, 4), gdofs=round(mf.n_owned / ms / 1e6, 2),
                          gdofs_kernel=round(mf.n_owned / ms_k / 1e6, 2),
                          mesh_s=round(t1 - t0, 2), setup_s=round(t2 - t1, 2))), flush=True)

if __name__ == "__main__":
    run(3, 4, 5, "f64")
    run(3, 4, 6, "f64")
    run(3, 4, 7, "f64")
    run(3, 4, 7, "f32")
    run(3, 4, 6, "f64", amp=0.05)
    run(3, 3, 7, "f64")
    run(3, 5, 6, "f64")
    run(3, 2, 7, "f64")
    run(3, 1, 8, "f64")
    run(3, 8, 5, "f64")
