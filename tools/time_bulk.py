"""Device timing of vmult through the bulk brick kernel vs the index-map brick kernel (A/B)."""
import json, sys, time
import torch
import dealii_b200

def run(degree, refinements, number, numbering="default", reps=20):
    mesh = dealii_b200.HyperCubeMesh(3, degree, refinements=refinements, numbering=numbering)
    t0 = time.time()
    mf = dealii_b200.MatrixFree(number).reinit_from_mesh(mesh)
    setup_s = time.time() - t0
    op = dealii_b200.LaplaceOperator(mf)
    x = torch.rand(mf.n_owned, dtype=mf.torch_dtype, device="cuda")
    y = mf.initialize_dof_vector()
    out = dict(degree=degree, n_dofs=mf.n_owned, number=number, numbering=numbering, setup_s=round(setup_s, 2), info=mf.bulk_info())
    ref = None
    for mode, path in (("bulk", 2), ("coloured", 1), ("map", 0)):
        if mf.select_brick_path(path) != path:
            continue
        for _ in range(3):
            op.vmult(y, x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            op.vmult(y, x)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        out[mode] = dict(ms=round(ms, 4), gdofs=round(mf.n_owned / ms / 1e6, 1),
                         frac=round((16 if number == "f64" else 8) * mf.n_owned / ms / 1e6 / 6551.4, 3))
        if ref is None:
            ref = y.clone()
        else:
            out["maxdiff_rel"] = float((y - ref).abs().max() / ref.abs().max())
    print(json.dumps(out), flush=True)

if __name__ == "__main__":
    cases = [(4, 5, "f64"), (4, 7, "f64"), (4, 7, "f32"), (3, 7, "f64"), (5, 6, "f64"), (2, 7, "f64"),
             (1, 8, "f64"), (6, 6, "f64"), (8, 5, "f64")]
    if len(sys.argv) > 1:
        cases = [(int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]) + tuple(sys.argv[4:5])]
    for c in cases:
        run(*c)
