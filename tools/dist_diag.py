"""Where the distributed vmult spends its time: device timings of the pieces (run under torchrun)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import dealii_b200
from dealii_b200.distributed import PartitionedHyperCubeMesh, DistributedMatrixFree

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", lr); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
coarse = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[world]
ref = int(sys.argv[1]) if len(sys.argv) > 1 else 7
mesh = PartitionedHyperCubeMesh(3, 4, refinements=ref, coarse=coarse, n_ranks=world, rank=rank)
dmf = DistributedMatrixFree(mesh, "f64", dev)
op = dealii_b200.LaplaceOperator(dmf.mf)
src = dmf.initialize_dof_vector(); src[:mesh.n_owned] = torch.rand(mesh.n_owned, dtype=torch.float64, device=dev)
dst = dmf.initialize_dof_vector()
ex = dmf.exchange

def timed(f, reps=20):
    for _ in range(3): f()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / reps, 4), round((time.perf_counter() - t0) / reps * 1e3, 4)

out = {"rank": rank, "n_owned": mesh.n_owned, "n_ghost": mesh.n_ghost, "n_import": dmf.partitioner.n_import,
       "n_interior": dmf.n_interior, "n_cells": dmf.n_cells, "bricks": int(dmf.mf.info.n_bricks)}
out["update_ghost_values"] = timed(lambda: ex.update_ghost_values(src))
out["compress"] = timed(lambda: ex.compress(dst))
out["pack"] = timed(lambda: ex._pack(src))
out["unpack_add"] = timed(lambda: ex._unpack_add(dst))
out["vmult_overlap"] = timed(lambda: dmf.vmult(op.op, dst, src))
dmf.overlap = False
out["vmult_no_overlap_split"] = timed(lambda: dmf.vmult(op.op, dst, src))
def local_only():
    dmf.mf.vmult_prepare(op.op, dst); dmf.mf.vmult_range(op.op, dst, src, 0, dmf.n_cells)
out["local_cell_loop_one_launch"] = timed(local_only)
def local_three():
    dmf.mf.vmult_prepare(op.op, dst)
    ni, w = dmf.n_interior, dmf._brick_cells
    half = (ni // 2) // w * w
    dmf.mf.vmult_range(op.op, dst, src, 0, half); dmf.mf.vmult_range(op.op, dst, src, ni, dmf.n_cells)
    dmf.mf.vmult_range(op.op, dst, src, half, ni)
out["local_cell_loop_three_launches"] = timed(local_three)
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()
