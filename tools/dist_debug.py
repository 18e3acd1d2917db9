"""torchrun --nproc-per-node 2 tools/dist_debug.py : energy 1^T A 1 and CG iterations for ghost modes / brick paths."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import dealii_b200
from dealii_b200.distributed import DistributedMatrixFree, PartitionedHyperCubeMesh, solve_cg
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
if world > 1: dist.init_process_group("nccl", device_id=dev)
comm = None
for mode in ("relevant", "touched"):
    for path in (None, 0, 1):
        pm = PartitionedHyperCubeMesh(3, 4, 3, world, rank, dirichlet_boundary=True, ghost_mode=mode)
        dmf = DistributedMatrixFree(pm, "f64", dev, comm=comm); comm = dmf.comm
        if path is not None: dmf.mf.select_brick_path(path)
        A = dealii_b200.LaplaceOperator(dmf.mf)
        b = dmf.initialize_dof_vector(); b[:pm.n_owned] = 1.0; dmf.mf.set_constrained_values(0.0, b)
        y = dmf.initialize_dof_vector(); dmf.vmult(A.op, y, b)
        e = torch.dot(y[:pm.n_owned], b[:pm.n_owned]).reshape(1)
        if world > 1: dist.all_reduce(e)
        inv = dmf.compute_diagonal(A.op)
        x = dmf.initialize_dof_vector()
        its, res, ok = solve_cg(dmf, A.op, x, b, inv, 1e-8 * float(pm.n_global_dofs) ** 0.5, 3000)
        if rank == 0:
            print(mode, "path", path, "info path", dmf.mf.bulk_info()["path"], "n_ghost", pm.n_ghost, "n_int", pm.n_cells_interior, "bricks", int(dmf.mf.info.n_bricks), "energy %.10e" % float(e), "its", its, ok, flush=True)
if world > 1: dist.destroy_process_group()
