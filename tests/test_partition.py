"""Partitioned mesh generator and Partitioner index algebra (host logic, no GPU).

* the product's Partitioner reproduces tests/mpi/parallel_partitioner_03.mpirun=4.output;
* the product's partitioned mesh generator reproduces the oracle's restatement of
  parallel::distributed numbering bit-exactly (global numbers, owned ranges, ghost sets);
* a world_size-2 gloo run exchanges ghost values / compresses on CPU and reproduces the serial
  oracle's vmult with the oracle cell operator applied to each rank's local cells."""
import os

import re

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import dealii_b200
from dealii_b200.distributed import GhostExchange, PartitionedHyperCubeMesh, Partitioner
from oracle.mesh import HyperCubeMesh as OracleMesh
from oracle.mf_oracle import MatrixFreeOracle
from oracle.partition import distributed_numbering, relevant_dofs

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_partitioner_matches_reference_golden():
    nproc, s = 4, 200
    offsets, start = [0], 0
    for r in range(nproc):
        start += s - r
        offsets.append(start)
    ghosts = np.array([1, 2, 13, s - 2, s - 1, s, s + 1, 2 * s, 2 * s + 1, 2 * s + 3])
    per_rank = []
    for r in range(nproc):
        g = ghosts[(ghosts < offsets[r]) | (ghosts >= offsets[r + 1])]
        per_rank.append(np.unique(g))
    text = "".join(Partitioner(offsets, r, per_rank[r], all_ghosts=per_rank).format_like_reference_test()
                   for r in range(nproc))
    golden = open(os.path.join(GOLD, "parallel_partitioner_03.mpirun=4.output")).read()
    golden = golden[golden.index("**** proc 0"):]
    assert text.split() == golden.split()


@pytest.mark.parametrize("dim,degree,refinements,n_ranks", [
    (3, 2, 2, 2), (3, 2, 2, 4), (3, 3, 2, 8), (3, 1, 3, 8), (2, 3, 3, 4), (2, 2, 3, 2), (3, 4, 1, 8),
    (3, 2, 2, 1)])
@pytest.mark.parametrize("mode", ["relevant", "touched"])
def test_partitioned_numbering_matches_oracle(dim, degree, refinements, n_ranks, mode):
    om = OracleMesh(dim, degree, refinements=refinements)
    num = distributed_numbering(om, n_ranks)
    L = om.lattice_size
    serial_of_lat = om._number_of_lattice
    lat_points = np.nonzero(serial_of_lat >= 0)[0]
    lat_of_serial = np.zeros(om.n_dofs, dtype=np.int64)
    lat_of_serial[serial_of_lat[lat_points]] = lat_points
    seen = np.zeros(om.n_dofs, dtype=int)
    for rank in range(n_ranks):
        pm = PartitionedHyperCubeMesh(dim, degree, refinements, n_ranks, rank, ghost_mode=mode,
                                      want_lattice_ids=True, dirichlet_boundary=True)
        assert np.array_equal(pm.rank_offsets, num["rank_offsets"])
        assert pm.n_global_dofs == om.n_dofs
        # global number of every local dof through its lattice id
        serial = serial_of_lat[pm.lattice_ids]
        glob_expected = num["global_of_serial"][serial]
        glob_local = np.concatenate((pm.first_owned_global + np.arange(pm.n_owned), pm.ghost_global)).astype(np.int64)
        assert np.array_equal(glob_local, glob_expected)
        seen[serial[:pm.n_owned]] += 1
        # ghost set
        assert np.array_equal(pm.ghost_global.astype(np.int64), relevant_dofs(om, num, rank, mode))
        # index lists: same cells (as sets of lattice points), interior cells touch no ghost
        own_cells = np.nonzero(num["cell_rank"] == rank)[0]
        expect = np.sort(lat_of_serial[om.l2g[own_cells]], axis=1)
        got = np.sort(pm.lattice_ids[pm.l2g & 0x7FFFFFFF].astype(np.int64), axis=1)
        assert sorted(map(tuple, got)) == sorted(map(tuple, expect))
        assert (pm.l2g[:pm.n_cells_interior] < pm.n_owned).all()
        if pm.n_cells_interior < pm.n_cells and mode == "touched":
            # 2D: cell by cell; 3D: whole aligned blocks of cells (the bricks of the brick kernel)
            # are classified together, so every block behind n_cells_interior holds such a cell
            touches = (pm.l2g[pm.n_cells_interior:] >= pm.n_owned).any(axis=1)
            if dim == 2:
                assert touches.all()
            else:
                b = 16 if degree == 1 else 8 if degree == 2 else 4 if degree <= 4 else 2
                W = b ** 3 if pm.n_cells % b ** 3 == 0 else 1
                assert touches.reshape(-1, W).any(axis=1).all()
        # lexicographic order inside a cell is preserved: x fastest lattice ids
        lat = pm.lattice_ids[pm.l2g[0]]
        assert np.array_equal(np.sort(lat), lat) or dim > 1
        # Dirichlet list = owned boundary dofs
        coords = np.stack([(pm.lattice_ids[:pm.n_owned] // L ** d) % L for d in range(dim)], 1)
        bnd = np.nonzero(((coords == 0) | (coords == L - 1)).any(axis=1))[0]
        assert np.array_equal(pm.boundary_dofs, bnd)
    assert (seen == 1).all()


def test_single_rank_partition_equals_serial_generator():
    a = dealii_b200.HyperCubeMesh(3, 3, refinements=2)
    b = PartitionedHyperCubeMesh(3, 3, 2, 1, 0)
    assert np.array_equal(a.l2g, b.l2g) and b.n_ghost == 0 and b.n_cells_interior == b.n_cells


def test_coarse_grid_partition_one_cube_per_rank():
    # subdivided_hyper_rectangle(1,1,2) refined once, two ranks: each owns one cube
    tot = 0
    for rank in range(2):
        pm = PartitionedHyperCubeMesh(3, 2, 1, 2, rank, coarse=(1, 1, 2), want_lattice_ids=True)
        tot += pm.n_owned
        assert pm.n_cells == 8
    assert tot == 5 * 5 * 9 and pm.n_global_dofs == tot


# ----------------------------------------------------------------------------------------------
class _CpuExchange(GhostExchange):
    """Test-only plumbing: pack / unpack-add with torch CPU ops so the host-side message plan
    can run under gloo without a GPU (the product's kernels are CUDA only)."""

    def _pack(self, vec):
        self.buf[:self.part.n_import] = vec[self.import_idx.long()]

    def _unpack_add(self, vec):
        vec.index_add_(0, self.import_idx.long(), self.buf[:self.part.n_import])


def _worker(rank, world, port, dim, degree, refinements, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pm = PartitionedHyperCubeMesh(dim, degree, refinements, world, rank, want_lattice_ids=True)
        part = Partitioner(pm.rank_offsets, rank, pm.ghost_global)        # all_gather_object over gloo
        ex = _CpuExchange(part, "f64", "cpu")
        # src = f(lattice id): identical global vector on every partition
        val = lambda lat: np.sin(0.37 * lat.astype(np.float64)) + 1.5
        vec = torch.zeros(pm.n_owned + pm.n_ghost, dtype=torch.float64)
        vec[:pm.n_owned] = torch.from_numpy(val(pm.lattice_ids[:pm.n_owned]))
        ex.update_ghost_values(vec)
        ok_ghost = np.allclose(vec.numpy(), val(pm.lattice_ids))
        # local cell operator with the oracle on this rank's cells, then compress(add)
        class _M:  # minimal mesh view for the oracle
            pass
        m = _M()
        m.dim, m.degree, m.n_cells, m.n_dofs = dim, degree, pm.n_cells, pm.n_owned + pm.n_ghost
        m.l2g, m.cell_vertices = pm.l2g.astype(np.int64), pm.cell_vertices
        loc = MatrixFreeOracle(m).cell_loop(vec.numpy())
        dst = torch.from_numpy(loc.copy())
        ex.compress(dst)
        ret[rank] = (ok_ghost, pm.lattice_ids[:pm.n_owned].copy(), dst[:pm.n_owned].numpy().copy(),
                     float(dst[pm.n_owned:].abs().max()) if pm.n_ghost else 0.0)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("dim,degree,refinements", [(3, 2, 2), (2, 3, 3)])
def test_gloo_two_ranks_ghost_exchange_and_vmult(dim, degree, refinements):
    world, port = 2, 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, dim, degree, refinements, ret), nprocs=world, join=True)
    om = OracleMesh(dim, degree, refinements=refinements)
    lat_all = np.nonzero(om._number_of_lattice >= 0)[0]
    src = np.zeros(om.n_dofs)
    src[om._number_of_lattice[lat_all]] = np.sin(0.37 * lat_all.astype(np.float64)) + 1.5
    ref = MatrixFreeOracle(om).cell_loop(src)
    covered = 0
    for rank in range(world):
        ok_ghost, lat, dst, ghost_max = ret[rank]
        assert ok_ghost and ghost_max == 0.0
        expect = ref[om._number_of_lattice[lat]]
        assert np.abs(dst - expect).max() <= 1e-12 * np.abs(ref).max()
        covered += len(lat)
    assert covered == om.n_dofs


def _parse_index_set(txt):
    out = []
    for tok in re.findall(r"\[(\d+),(\d+)\]|(\d+)", txt):
        if tok[2]:
            out.append(int(tok[2]))
        else:
            out += list(range(int(tok[0]), int(tok[1]) + 1))
    return out


def test_owned_and_relevant_sets_match_the_references_two_rank_run():
    """tests/matrix_free_kokkos/matrix_free_device_initialize_vector (mpirun=2, p4est): 2D Q1 on
    hyper_cube refine_global(2): the locally owned range and the locally relevant dofs
    (= owned + the ghosts Portable::MatrixFree's partitioner holds) of both ranks, as printed by the
    reference, against the partitioned mesh generator -- a real multi-rank known answer for the
    "numbering and ghost index sets bit-exact" requirement."""
    import os
    path = os.path.join(os.path.dirname(__file__), "golden", "matrix_free_device_initialize_vector.mpirun=2.output")
    txt = open(path).read()
    blocks = re.findall(r"DEAL:(\d):\d::locally owned dofs :\n(\{.*\})\nDEAL:\d:\d::locally relevant dofs :\n(\{.*\})", txt)
    assert len(blocks) == 4                                  # two vector types x two ranks
    for rank_s, owned_s, relevant_s in blocks:
        rank = int(rank_s)
        owned, relevant = _parse_index_set(owned_s), _parse_index_set(relevant_s)
        pm = PartitionedHyperCubeMesh(2, 1, refinements=2, n_ranks=2, rank=rank, ghost_mode="relevant")
        assert owned == list(range(pm.first_owned_global, pm.first_owned_global + pm.n_owned))
        assert sorted(set(relevant) - set(owned)) == sorted(np.asarray(pm.ghost_global).tolist())


def test_global_numbering_matches_the_references_four_rank_run():
    """tests/mpi/p4est_2d_dofhandler_01 (mpirun=4, p4est): 2D Q2 on hyper_cube refine_global(2).
    The reference prints the owned dof counts of the four ranks and, for every cell rank 0 knows
    (its own and its ghost cells, active-cell = Morton order), the GLOBAL dof indices in
    hierarchical order.  Rebuilt here from the four ranks' partitioned meshes."""
    import os
    from oracle.mesh import hierarchic_to_lexicographic, morton_cell_coords
    path = os.path.join(os.path.dirname(__file__), "golden", "p4est_2d_dofhandler_01.mpirun=4.output")
    lines = [l.split("::", 1)[1].strip() for l in open(path) if "::" in l]
    counts = [int(t) for t in lines[1].split()[1:]]
    cell_lines = [[int(t) for t in l.split()] for l in lines[3:]]
    meshes = [PartitionedHyperCubeMesh(2, 2, refinements=2, n_ranks=4, rank=r, ghost_mode="relevant")
              for r in range(4)]
    assert [m.n_owned for m in meshes] == counts
    h2l = hierarchic_to_lexicographic(2, 2)
    # global hierarchical index list of every cell, keyed by its integer coordinates
    cells = {}
    for m in meshes:
        glob = np.concatenate((m.first_owned_global + np.arange(m.n_owned), np.asarray(m.ghost_global))).astype(np.int64)
        for c in range(m.n_cells):
            xy = tuple(np.round(m.cell_vertices[c, 0] * 4).astype(int))
            cells[xy] = glob[m.l2g[c] & 0x7FFFFFFF][h2l]
    assert len(cells) == 16
    # rank 0 owns the first 4 cells of the Morton curve and sees the cells touching them
    order = [tuple(c) for c in morton_cell_coords(2, 2)]
    own = set(order[:4])
    known = [c for c in order if c in own or any(abs(c[0] - o[0]) <= 1 and abs(c[1] - o[1]) <= 1 for o in own)]
    assert len(known) == len(cell_lines) == 9
    for c, expect in zip(known, cell_lines):
        assert cells[c].tolist() == expect, c


def _c_partitioner(n_ranks, rank, offsets, ghosts_of_all):
    """b200mf_partitioner_create_host: the index algebra of the C layer (no communicator, no GPU)."""
    import ctypes as C
    from dealii_b200 import _lib as L
    lib = L.load()
    ro = np.ascontiguousarray(offsets, dtype=np.uint64)
    gg = np.ascontiguousarray(ghosts_of_all[rank], dtype=np.uint64)
    ranks, counts, glob = [], [], []
    for r in range(n_ranks):
        if r == rank:
            continue
        g = np.asarray(ghosts_of_all[r], dtype=np.uint64)
        mine = g[(g >= offsets[rank]) & (g < offsets[rank + 1])]
        if len(mine):
            ranks.append(r); counts.append(len(mine)); glob.append(mine)
    ir = (C.c_int * max(len(ranks), 1))(*ranks)
    ic = (C.c_uint64 * max(len(counts), 1))(*counts)
    ig = np.ascontiguousarray(np.concatenate(glob) if glob else np.zeros(0), dtype=np.uint64)
    h = C.c_void_p()
    L.check(lib.b200mf_partitioner_create_host(n_ranks, rank, ro.ctypes.data_as(C.POINTER(C.c_uint64)),
                                               gg.ctypes.data_as(C.POINTER(C.c_uint64)), len(gg), L.F64, ir, ic,
                                               len(ranks), ig.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(h)))
    info = L.PartitionerInfo()
    L.check(lib.b200mf_partitioner_get_info(h, C.byref(info)))
    out = dict(
        ghost_targets=[(info.ghost_target_ranks[i], int(info.ghost_target_counts[i])) for i in range(info.n_ghost_targets)],
        import_targets=[(info.import_target_ranks[i], int(info.import_target_counts[i])) for i in range(info.n_import_targets)],
        import_indices=[int(info.import_indices[i]) for i in range(info.n_import)], n_import=int(info.n_import))
    lib.b200mf_partitioner_destroy(h)
    return out


def test_c_partitioner_matches_reference_golden_layout():
    """The C partitioner (csrc/comm.cu) on the index sets of tests/mpi/parallel_partitioner_03.cc:
    same ghost_targets / import_targets / import_indices as the Python restatement that reproduces
    the reference's golden output above."""
    nproc, s = 4, 200
    offsets, start = [0], 0
    for r in range(nproc):
        start += s - r
        offsets.append(start)
    ghosts = np.array([1, 2, 13, s - 2, s - 1, s, s + 1, 2 * s, 2 * s + 1, 2 * s + 3])
    per_rank = [np.unique(ghosts[(ghosts < offsets[r]) | (ghosts >= offsets[r + 1])]) for r in range(nproc)]
    for r in range(nproc):
        py = Partitioner(offsets, r, per_rank[r], all_ghosts=per_rank)
        c = _c_partitioner(nproc, r, offsets, per_rank)
        assert c["ghost_targets"] == [(int(a), int(b)) for a, b in py.ghost_targets]
        assert c["import_targets"] == [(int(a), int(b)) for a, b in py.import_targets]
        flat = np.concatenate(py.import_indices) if py.import_indices else np.zeros(0, dtype=np.int64)
        assert c["import_indices"] == [int(v) for v in flat]


def test_c_partitioner_on_a_partitioned_mesh():
    meshes = [PartitionedHyperCubeMesh(3, 2, 2, 8, r, ghost_mode="touched") for r in range(8)]
    ghosts = [m.ghost_global for m in meshes]
    for r, m in enumerate(meshes):
        py = Partitioner(m.rank_offsets, r, m.ghost_global, all_ghosts=ghosts)
        c = _c_partitioner(8, r, m.rank_offsets, ghosts)
        assert c["ghost_targets"] == [(int(a), int(b)) for a, b in py.ghost_targets]
        assert c["import_targets"] == [(int(a), int(b)) for a, b in py.import_targets]
        assert c["n_import"] == py.n_import
