"""The multigrid on a partitioned mesh: the same global problem (2 x 2 x 2 coarse cells, refined r times) solved
on 1, 2, 4 and 8 ranks must give the same level eigenvalue estimates, the same number of CG iterations and the
same solution -- the one-rank run is the serial multigrid that tests/test_multigrid_gpu.py pins on deal.II.
Multi-rank cases need that many GPUs and are skipped otherwise."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _run_rank(rank, world, port, degree, refinements, number, ret):
    import dealii_b200
    from dealii_b200.distributed import (DistributedGeometricMultigrid, DistributedMatrixFree,
                                         PartitionedHyperCubeMesh)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        coarse = (2, 2, 2)
        mg = DistributedGeometricMultigrid(3, degree, refinements, world, rank, coarse=coarse, number=number, device=dev)
        mesh = PartitionedHyperCubeMesh(3, degree, refinements, world, rank, coarse=coarse, dirichlet_boundary=True,
                                        mark_constrained_l2g=True, ghost_mode="touched")
        system = DistributedMatrixFree(mesh, "f64", dev, comm=mg.comm)
        A = dealii_b200.LaplaceOperator(system.mf)
        n = mesh.n_owned
        b = system.initialize_dof_vector()
        b[:n] = 1.0
        system.mf.set_constrained_values(0.0, b)
        scal = torch.tensor([float(torch.dot(b[:n], b[:n]))], device=dev, dtype=torch.float64)
        mg.comm.allreduce_sum(scal)
        bnorm = float(scal) ** 0.5
        # one V-cycle: a global functional of the result
        z = system.initialize_dof_vector()
        mg.vmult(z, b)
        scal = torch.tensor([float(torch.dot(z[:n], b[:n])), float(torch.dot(z[:n], z[:n]))], device=dev, dtype=torch.float64)
        mg.comm.allreduce_sum(scal)
        zb, zz = float(scal[0]), float(scal[1])
        ghost_clean = float(z[n:].abs().max()) if mesh.n_ghost else 0.0
        x = system.initialize_dof_vector()
        its, res, ok = mg.solve(system, A.op, x, b, 1e-10 * bnorm, 50)
        scal = torch.tensor([float(torch.dot(x[:n], x[:n]))], device=dev, dtype=torch.float64)
        mg.comm.allreduce_sum(scal)
        # true residual through the distributed operator
        r = system.initialize_dof_vector()
        system.vmult(A.op, r, x)
        rr = torch.tensor([float(torch.dot(b[:n] - r[:n], b[:n] - r[:n]))], device=dev, dtype=torch.float64)
        mg.comm.allreduce_sum(rr)
        infos = [(mg.level_info(l).eig_min, mg.level_info(l).eig_max, mg.level_info(l).degree) for l in range(mg.n_levels())]
        ret[rank] = dict(its=its, ok=ok, xx=float(scal), zb=zb, zz=zz, infos=infos, ghost_clean=ghost_clean,
                         true_res=float(rr) ** 0.5 / bnorm, n_global=mesh.n_global_dofs)
    finally:
        if world > 1:
            dist.destroy_process_group()


def _run(world, degree, refinements, number):
    port = 29700 + (os.getpid() % 2000) + world
    if world == 1:
        ret = {}
        _run_rank(0, 1, port, degree, refinements, number, ret)
        return ret[0]
    ret = mp.Manager().dict()
    mp.spawn(_run_rank, args=(world, port, degree, refinements, number, ret), nprocs=world, join=True)
    out = [ret[r] for r in range(world)]
    for o in out[1:]:
        assert o["its"] == out[0]["its"] and o["infos"] == out[0]["infos"]
    return out[0]


@pytest.mark.parametrize("degree,refinements,number", [(2, 3, "f64"), (4, 2, "f64"), (4, 3, "f32")])
def test_partitioned_multigrid_is_independent_of_the_rank_count(degree, refinements, number):
    base = _run(1, degree, refinements, number)
    assert base["ok"] and base["its"] <= 9 and base["true_res"] < 1e-9
    ran = 0
    for world in (2, 4, 8):
        if torch.cuda.device_count() < world:
            continue
        ran += 1
        o = _run(world, degree, refinements, number)
        tol = 1e-4 if number == "f32" else 1e-9
        assert o["ok"] and o["n_global"] == base["n_global"] and o["ghost_clean"] == 0.0
        assert abs(o["its"] - base["its"]) <= (1 if number == "f32" else 0), (o["its"], base["its"])
        assert o["true_res"] < 1e-9
        for (a0, a1, ad), (b0, b1, bd) in zip(o["infos"], base["infos"]):
            assert a0 == pytest.approx(b0, rel=1e-3 if number == "f32" else 1e-6)
            assert a1 == pytest.approx(b1, rel=1e-3 if number == "f32" else 1e-6)
            assert ad == bd
        assert o["zb"] == pytest.approx(base["zb"], rel=tol) and o["zz"] == pytest.approx(base["zz"], rel=tol)
        assert o["xx"] == pytest.approx(base["xx"], rel=1e-9)
    if ran == 0:
        pytest.skip("one GPU: only the one-rank run of the partitioned code path was checked")
