"""Distributed operator / CG on the GPU(s): DistributedMatrixFree + NCCL ghost exchange against
the serial oracle (vmult per entry through lattice ids) and against the single-rank solver
(iteration counts).  The multi-rank cases need >= 2 GPUs and are skipped otherwise."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import dealii_b200
from dealii_b200.distributed import DistributedMatrixFree, PartitionedHyperCubeMesh, solve_cg
from oracle.mesh import HyperCubeMesh as OracleMesh
from oracle.mf_oracle import MatrixFreeOracle

pytestmark = pytest.mark.gpu


def _value(lat):
    return np.sin(0.37 * lat.astype(np.float64)) + 1.5


def _run_rank(rank, world, port, dim, degree, refinements, amp, ret, path=None):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        pm = PartitionedHyperCubeMesh(dim, degree, refinements, world, rank, want_lattice_ids=True,
                                      dirichlet_boundary=True, deformation_amplitude=amp)
        dmf = DistributedMatrixFree(pm, "f64", dev)
        if path is not None:
            dmf.mf.select_brick_path(path)
        A = dealii_b200.LaplaceOperator(dmf.mf)
        src = dmf.initialize_dof_vector()
        src[:pm.n_owned] = torch.from_numpy(_value(pm.lattice_ids[:pm.n_owned])).to(dev)
        dmf.mf.set_constrained_values(0.0, src)
        dst = dmf.initialize_dof_vector()
        dmf.vmult(A.op, dst, src)
        torch.cuda.synchronize()
        ghost_clean = float(src[pm.n_owned:].abs().max()) if pm.n_ghost else 0.0
        # CG + Jacobi, rhs = 1 on unconstrained dofs
        inv = dmf.compute_diagonal(A.op)
        b = dmf.initialize_dof_vector()
        b[:pm.n_owned] = 1.0
        dmf.mf.set_constrained_values(0.0, b)
        x = dmf.initialize_dof_vector()
        bn = torch.tensor([float(torch.dot(b[:pm.n_owned], b[:pm.n_owned]))], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(bn)
        its, res, ok = solve_cg(dmf, A.op, x, b, inv, 1e-10 * float(bn) ** 0.5, 2000)
        ret[rank] = (pm.lattice_ids[:pm.n_owned].copy(), dst[:pm.n_owned].cpu().numpy(), ghost_clean,
                     its, ok, x[:pm.n_owned].cpu().numpy(), inv[:pm.n_owned].cpu().numpy())
    finally:
        if world > 1:
            dist.destroy_process_group()


def _reference(dim, degree, refinements, amp):
    defo = (lambda v: v + amp * np.prod(np.sin(np.pi * v), axis=1, keepdims=True)) if amp else None
    om = OracleMesh(dim, degree, refinements=refinements, deformation=defo)
    o = MatrixFreeOracle(om, constrained_dofs=om.boundary_dofs)
    lat_points = np.nonzero(om._number_of_lattice >= 0)[0]
    src = np.zeros(om.n_dofs)
    src[om._number_of_lattice[lat_points]] = _value(lat_points)
    src[om.boundary_dofs] = 0.0
    return om, o, o.vmult(src)


@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("dim,degree,refinements,amp,path", [(3, 4, 2, 0.0, None), (3, 4, 3, 0.0, 1), (3, 4, 3, 0.0, 0),
                                                             (3, 2, 3, 0.05, None), (2, 3, 4, 0.0, None)])
def test_distributed_vmult_and_cg(world, dim, degree, refinements, amp, path):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29600 + (os.getpid() % 2000)
    if world == 1:
        ret = {}
        _run_rank(0, 1, port, dim, degree, refinements, amp, ret, path)
    else:
        ret = mp.Manager().dict()
        mp.spawn(_run_rank, args=(world, port, dim, degree, refinements, amp, ret, path), nprocs=world, join=True)
    om, o, ref = _reference(dim, degree, refinements, amp)
    ref_diag = o.compute_diagonal()
    its = set()
    for rank in range(world):
        lat, dst, ghost_clean, it, ok, x, inv = ret[rank]
        ser = om._number_of_lattice[lat]
        assert np.abs(dst - ref[ser]).max() <= 1e-12 * np.abs(ref).max()
        assert np.abs(inv * ref_diag[ser] - 1.0).max() < 1e-12
        assert ghost_clean == 0.0 and ok
        its.add(it)
    assert len(its) == 1
    # iteration count of the single-rank engine solver on the same problem (+-1)
    mf = dealii_b200.MatrixFree("f64").reinit(dim, degree, om.l2g.astype(np.uint32), cell_vertices=om.cell_vertices,
                                              constrained_dofs=om.boundary_dofs, n_owned_dofs=om.n_dofs)
    A = dealii_b200.LaplaceOperator(mf)
    invd = A.compute_diagonal()
    b = torch.ones(om.n_dofs, dtype=torch.float64, device="cuda")
    mf.set_constrained_values(0.0, b)
    x = mf.initialize_dof_vector()
    control = dealii_b200.SolverControl(2000, 1e-10 * float(b.norm()))
    dealii_b200.SolverCG(control).solve(A, x, b, invd)
    assert abs(control.last_step() - its.pop()) <= 1
    xs = x.cpu().numpy()
    for rank in range(world):
        lat, _, _, _, _, xr, _ = ret[rank]
        assert np.abs(xr - xs[om._number_of_lattice[lat]]).max() < 1e-8 * np.abs(xs).max()


# ---------------------------------------------------------------------------------------------
# BASELINE configs[3]: the adaptive (hanging-node) partitioned mesh through the distributed loop
def _run_adaptive_rank(rank, world, port, degree, refinements, ret, path=None):
    from dealii_b200.distributed import AdaptiveHyperCubeMesh
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        coarse = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[world]
        mesh = AdaptiveHyperCubeMesh(3, degree, refinements, world, rank, coarse=coarse, ball_radius=0.33,
                                     want_coords=True)
        dmf = DistributedMatrixFree(mesh, "f64", dev)
        if path is not None:
            dmf.mf.select_brick_path(path)
        A = dealii_b200.LaplaceOperator(dmf.mf)
        n = mesh.n_owned
        cons = torch.from_numpy(mesh.constrained_dofs.astype(np.int64)).to(dev)     # hanging-node dofs

        def gsum(v):
            t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t)
            return float(t)

        def gmax(v):
            t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t)

        # (1) the Laplacian annihilates constants: needs every ghost value, every hanging-node
        # interpolation and its transpose, and the compress of the interface rows
        one = dmf.initialize_dof_vector()
        one[:n] = 1.0
        one[cons] = 0.0
        y = dmf.initialize_dof_vector()
        dmf.vmult(A.op, y, one)
        y[cons] = 0.0
        diag = 1.0 / dmf.compute_diagonal(A.op)[:n]
        defect = gmax((y[:n].abs() / diag.abs()).max())
        # (2) symmetry with partition-independent vectors (functions of the support point)
        xyz = torch.from_numpy(mesh.dof_coords[:n]).to(dev)
        u, v = dmf.initialize_dof_vector(), dmf.initialize_dof_vector()
        u[:n] = torch.sin(3.1 * xyz[:, 0]) * torch.cos(2.3 * xyz[:, 1]) + xyz[:, 2] ** 2
        v[:n] = torch.cos(1.7 * xyz[:, 0] + 0.4) * xyz[:, 1] + torch.sin(2.9 * xyz[:, 2])
        u[cons] = 0.0
        v[cons] = 0.0
        au, av = dmf.initialize_dof_vector(), dmf.initialize_dof_vector()
        dmf.vmult(A.op, au, u)
        dmf.vmult(A.op, av, v)
        vau, uav = gsum(torch.dot(v[:n], au[:n])), gsum(torch.dot(u[:n], av[:n]))
        uau = gsum(torch.dot(u[:n], au[:n]))
        torch.cuda.synchronize()
        ret[rank] = (defect, vau, uav, uau, mesh.n_global_dofs, mesh.n_masked_cells, int(dmf.mf.info.n_bricks))
    finally:
        if world > 1:
            dist.destroy_process_group()


@pytest.mark.parametrize("path", [0, 1])
@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_distributed_adaptive_mesh_properties(world, path):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29700 + (os.getpid() % 2000)
    if world == 1:
        ret = {}
        _run_adaptive_rank(0, 1, port, 3, 4, ret, path)
    else:
        ret = mp.Manager().dict()
        mp.spawn(_run_adaptive_rank, args=(world, port, 3, 4, ret, path), nprocs=world, join=True)
    defect, vau, uav, uau, n_global, n_masked, n_bricks = ret[0]
    assert n_masked > 0 and n_bricks > 0
    assert defect < 1e-11
    assert abs(vau - uav) < 1e-11 * abs(vau)
    for r in range(world):
        assert ret[r][1:5] == ret[0][1:5]
    # the energy u.Au of a function of the support points is the same on every partition: compare with
    # the single-cube value scaled... (domains differ with the rank count, so only symmetry and the
    # constant defect are partition independent) -- nothing more to assert here
