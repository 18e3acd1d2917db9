"""Multi-GPU layer of the hot path: one process per GPU, torch.distributed for the plumbing.

Mirrors, by name and meaning:
  Partitioner            Utilities::MPI::Partitioner (base/partitioner.h:199; index algebra of
                         source/base/partitioner.cc:185-330): owned range + sorted ghost indices
                         -> ghost_targets, import_targets, import_indices
  GhostExchange          LA::d::Vector::update_ghost_values_start/finish, compress_start/finish,
                         zero_out_ghost_values (lac/la_parallel_vector.templates.h:1026-1332)
  DistributedMatrixFree  Portable::MatrixFree::distributed_cell_loop with
                         overlap_communication_computation
                         (matrix_free/portable_matrix_free.templates.h:1567-1690)
  solve_cg               SolverCG::solve (lac/solver_cg.h:1391) on distributed vectors: the fused
                         device kernels of the C ABI with one all-reduce of the scalar slots
                         where the reference calls MPI_Allreduce (lac/solver_cg.h:890-893)

Pack / unpack-add, the operator and the CG vector updates are CUDA kernels of libb200mf.so;
transport is NCCL point-to-point (ncclSend/ncclRecv grouped per exchange) over NVLink, issued
on a side stream so that interior cells run while the ghost values are in flight.

Overlap needs the transfer kernels to get SM slots while the cell loop has thousands of CTAs
queued: create the process group with TORCH_NCCL_HIGH_PRIORITY=1 in the environment (set below
as a default; it is read when the NCCL process group is created).  Measured with
tools/dist_diag.py on 2 x B200, Q4 FP64, 135 M DoFs per GPU: vmult 1.28 ms without, 1.18 ms with
(local cell loop alone: 1.13 ms).
"""
import ctypes as C
import os

os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")

import numpy as np
import torch
import torch.distributed as dist

from . import _lib as L
from .matrix_free import HyperCubeMesh, MatrixFree, _npptr, _ptr


class PartitionedHyperCubeMesh(HyperCubeMesh):
    """This rank's part of hyper_cube / subdivided_hyper_rectangle(coarse) refined globally,
    partitioned and numbered like parallel::distributed::Triangulation + DoFHandler
    (b200mf_mesh_create_partitioned)."""

    def __init__(self, dim, degree, refinements, n_ranks, rank, coarse=(1, 1, 1), left=0.0, right=1.0,
                 deformation_amplitude=0.0, dirichlet_boundary=False, mark_constrained_l2g=False,
                 ghost_mode="relevant", want_lattice_ids=False):
        lib = L.load()
        d = L.PartitionDesc()
        m = d.mesh
        m.dim, m.degree = dim, degree
        m.cells_per_direction, m.cell_order = 2 ** refinements, L.MESH_MORTON
        m.left, m.right = left, right
        m.deformation = L.DEFORM_SINE if deformation_amplitude != 0.0 else L.DEFORM_NONE
        m.deformation_amplitude = deformation_amplitude
        m.dirichlet_boundary = int(dirichlet_boundary)
        m.mark_constrained_l2g = int(mark_constrained_l2g)
        for k in range(3):
            d.coarse[k] = coarse[k] if k < len(coarse) else 1
        d.n_ranks, d.rank = n_ranks, rank
        d.ghost_mode = L.GHOSTS_RELEVANT if ghost_mode == "relevant" else L.GHOSTS_TOUCHED
        d.want_lattice_ids = int(want_lattice_ids)
        self._h = C.c_void_p()
        L.check(lib.b200mf_mesh_create_partitioned(C.byref(d), C.byref(self._h)))
        v = L.MeshView()
        L.check(lib.b200mf_mesh_view_get(self._h, C.byref(v)))
        pv = L.PartitionView()
        L.check(lib.b200mf_mesh_partition_view_get(self._h, C.byref(pv)))
        self.dim, self.degree = dim, degree
        self.n_cells, self.n_dofs = int(v.n_cells), int(v.n_dofs)
        self.dofs_per_cell = int(v.dofs_per_cell)
        self._view = v
        self.n_ranks, self.rank = n_ranks, rank
        self.n_owned, self.n_ghost = int(pv.n_owned), int(pv.n_ghost)
        self.n_global_dofs, self.n_global_cells = int(pv.n_global_dofs), int(pv.n_global_cells)
        self.first_owned_global = int(pv.first_owned_global)
        self.n_cells_interior = int(pv.n_cells_interior)
        self.rank_offsets = np.ctypeslib.as_array(pv.rank_offsets, shape=(n_ranks + 1,)).copy()
        self.ghost_global = (np.ctypeslib.as_array(pv.ghost_global, shape=(self.n_ghost,)).copy()
                             if self.n_ghost else np.zeros(0, dtype=np.uint64))
        self.lattice_ids = (np.ctypeslib.as_array(pv.lattice_ids, shape=(self.n_owned + self.n_ghost,)).copy()
                            if want_lattice_ids else None)
        # position of every local cell in this rank's Morton chunk (cells are ordered [no ghost | rest])
        self.cell_morton_position = np.ctypeslib.as_array(pv.cell_morton_position, shape=(self.n_cells,)).copy()


class AdaptiveHyperCubeMesh(PartitionedHyperCubeMesh):
    """This rank's cube of the partitioned mesh with one ball of refined cells per cube: hanging
    nodes, ConstraintKinds masks, redirected index lists (b200mf_mesh_create_adaptive)."""

    def __init__(self, dim, degree, refinements, n_ranks=1, rank=0, coarse=(1, 1, 1), ball_radius=0.3,
                 dirichlet_boundary=False, mark_constrained_l2g=False, brick_friendly_order=True,
                 want_coords=False):
        lib = L.load()
        a = L.AdaptiveDesc()
        d = a.part
        m = d.mesh
        m.dim, m.degree = dim, degree
        m.cells_per_direction, m.cell_order = 2 ** refinements, L.MESH_MORTON
        m.left, m.right = 0.0, 1.0
        m.deformation, m.deformation_amplitude = L.DEFORM_NONE, 0.0
        m.dirichlet_boundary = int(dirichlet_boundary)
        m.mark_constrained_l2g = int(mark_constrained_l2g)
        for k in range(3):
            d.coarse[k] = coarse[k] if k < len(coarse) else 1
        d.n_ranks, d.rank = n_ranks, rank
        d.ghost_mode = L.GHOSTS_TOUCHED
        d.want_lattice_ids = int(want_coords)
        a.ball_radius, a.brick_friendly_order = ball_radius, int(brick_friendly_order)
        self._h = C.c_void_p()
        L.check(lib.b200mf_mesh_create_adaptive(C.byref(a), C.byref(self._h)))
        v = L.MeshView()
        L.check(lib.b200mf_mesh_view_get(self._h, C.byref(v)))
        pv = L.PartitionView()
        L.check(lib.b200mf_mesh_partition_view_get(self._h, C.byref(pv)))
        av = L.AdaptiveView()
        L.check(lib.b200mf_mesh_adaptive_view_get(self._h, C.byref(av)))
        self.dim, self.degree = dim, degree
        self.n_cells, self.n_dofs = int(v.n_cells), int(v.n_dofs)
        self.dofs_per_cell = int(v.dofs_per_cell)
        self._view, self._aview = v, av
        self.n_ranks, self.rank = n_ranks, rank
        self.n_owned, self.n_ghost = int(pv.n_owned), int(pv.n_ghost)
        self.n_global_dofs, self.n_global_cells = int(pv.n_global_dofs), int(pv.n_global_cells)
        self.first_owned_global = int(pv.first_owned_global)
        self.n_cells_interior = int(pv.n_cells_interior)
        self.rank_offsets = np.ctypeslib.as_array(pv.rank_offsets, shape=(n_ranks + 1,)).copy()
        self.ghost_global = (np.ctypeslib.as_array(pv.ghost_global, shape=(self.n_ghost,)).copy()
                             if self.n_ghost else np.zeros(0, dtype=np.uint64))
        self.lattice_ids = None
        self.n_hanging_dofs, self.n_masked_cells = int(av.n_hanging_dofs), int(av.n_masked_cells)

    @property
    def constraint_mask(self):
        return np.ctypeslib.as_array(self._aview.constraint_mask, shape=(self.n_cells,))

    @property
    def active_cell_index(self):
        return np.ctypeslib.as_array(self._aview.active_cell_index, shape=(self.n_cells,))

    @property
    def constrained_dofs(self):
        return self.boundary_dofs

    @property
    def dof_coords(self):
        return np.ctypeslib.as_array(self._aview.dof_coords, shape=(self.n_owned + self.n_ghost, 3))


class Partitioner:
    """Utilities::MPI::Partitioner: who sends what to whom.

    ``rank_offsets[r] .. rank_offsets[r+1]`` is rank r's owned global range;
    ``ghost_global`` are this rank's ghost indices (sorted).  The import side needs the ghost
    sets of the other ranks: pass ``all_ghosts`` (list per rank) or a process ``group`` to
    exchange them (the reference runs a consensus algorithm, source/base/partitioner.cc:241-267).

    ghost_targets   [(rank, n)]  ranks owning my ghosts, ascending, with counts
                    (partitioner.cc:271-300); my ghost section is ordered accordingly
    import_targets  [(rank, n)]  ranks that ghost my owned dofs
    import_indices  per import target, the LOCAL owned indices in the order of that rank's
                    ghost list (compressed to half-open ranges by import_indices_ranges(),
                    the form tests/mpi/parallel_partitioner_03.cc prints)
    """

    def __init__(self, rank_offsets, rank, ghost_global, all_ghosts=None, group=None):
        self.rank = rank
        self.rank_offsets = np.asarray(rank_offsets, dtype=np.int64)
        self.n_ranks = len(self.rank_offsets) - 1
        self.first = int(self.rank_offsets[rank])
        self.n_owned = int(self.rank_offsets[rank + 1]) - self.first
        g = np.asarray(ghost_global, dtype=np.int64)
        assert np.all(np.diff(g) > 0), "ghost indices must be sorted and unique"
        assert not np.any((g >= self.first) & (g < self.first + self.n_owned)), "ghost index is owned"
        self.ghost_global = g
        self.n_ghost = len(g)
        owner = np.searchsorted(self.rank_offsets, g, side="right") - 1
        self.ghost_targets = [(int(r), int(c)) for r, c in zip(*np.unique(owner, return_counts=True))]
        if all_ghosts is None:
            if self.n_ranks == 1:
                all_ghosts = [g]
            else:
                all_ghosts = [None] * self.n_ranks
                dist.all_gather_object(all_ghosts, g, group=group)
        self.import_targets, self.import_indices = [], []
        for r in range(self.n_ranks):
            if r == rank:
                continue
            gr = np.asarray(all_ghosts[r], dtype=np.int64)
            mine = gr[(gr >= self.first) & (gr < self.first + self.n_owned)]
            if len(mine):
                self.import_targets.append((r, len(mine)))
                self.import_indices.append((mine - self.first).astype(np.int64))
        self.n_import = int(sum(c for _, c in self.import_targets))

    def global_to_local(self, gidx):
        gidx = np.asarray(gidx, dtype=np.int64)
        own = (gidx >= self.first) & (gidx < self.first + self.n_owned)
        pos = np.searchsorted(self.ghost_global, gidx)
        return np.where(own, gidx - self.first, self.n_owned + pos)

    def import_indices_ranges(self):
        """import_indices_data: half-open local ranges, per target consecutive runs merged."""
        out = []
        for idx in self.import_indices:
            if len(idx) == 0:
                continue
            breaks = np.nonzero(np.diff(idx) != 1)[0] + 1
            starts = np.concatenate(([0], breaks))
            ends = np.concatenate((breaks, [len(idx)]))
            out += [(int(idx[s]), int(idx[e - 1]) + 1) for s, e in zip(starts, ends)]
        return out

    def format_like_reference_test(self):
        """The text block tests/mpi/parallel_partitioner_03.cc:57-80 writes for this rank."""
        s = f"**** proc {self.rank}\n"
        s += "ghost targets: " + "".join(f"[{a}/{b}] " for a, b in self.ghost_targets) + "\n"
        s += "import targets: " + "".join(f"[{a}/{b}] " for a, b in self.import_targets) + "\n"
        s += "import indices:\n"
        s += "".join(f"[{a}/{b})\n" for a, b in self.import_indices_ranges())
        s += "****\n"
        return s


class GhostExchange:
    """update_ghost_values / compress(add) of LA::d::Vector over torch.distributed p2p.

    The pack and unpack-add steps are kernels of libb200mf.so; the transport is one grouped
    batch of isend/irecv (ncclGroupStart .. ncclSend/ncclRecv .. ncclGroupEnd under NCCL)."""

    def __init__(self, partitioner, number, device, group=None):
        self.part, self.group = partitioner, group
        self.number_code, self.dtype = (L.F64, torch.float64) if number == "f64" else (L.F32, torch.float32)
        self.device = torch.device(device)
        p = partitioner
        idx = (np.concatenate(p.import_indices) if p.import_indices else np.zeros(0, dtype=np.int64))
        self.import_idx = torch.from_numpy(idx.astype(np.int32)).to(self.device)
        self.buf = torch.empty(max(p.n_import, 1), dtype=self.dtype, device=self.device)
        # slices of the ghost section per owner / of the import buffer per importer
        self.ghost_slices, off = [], p.n_owned
        for r, c in p.ghost_targets:
            self.ghost_slices.append((r, off, off + c))
            off += c
        self.import_slices, off = [], 0
        for r, c in p.import_targets:
            self.import_slices.append((r, off, off + c))
            off += c
        self._lib = L.load() if self.device.type == "cuda" else None
        # NCCL: one all_to_all_single per exchange (a grouped ncclSend/ncclRecv issued from C++)
        # instead of 2 x n_neighbours Python-level P2POps -- the host cost of the latter (7
        # neighbours on a 2x2x2 box) was what bounded the 8-GPU vmult.  Ghosts are sorted by global
        # index = by owner rank, the import buffer by importing rank: both are split by rank.
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.ghost_splits, self.import_splits = [0] * world, [0] * world
        for r, c in p.ghost_targets:
            self.ghost_splits[r] = int(c)
        for r, c in p.import_targets:
            self.import_splits[r] = int(c)
        ranks_sorted = (sorted(r for r, _ in p.ghost_targets) == [r for r, _ in p.ghost_targets] and
                        sorted(r for r, _ in p.import_targets) == [r for r, _ in p.import_targets])
        self._use_a2a = (self.device.type == "cuda" and dist.is_initialized() and world > 1 and
                         dist.get_backend(group) == "nccl" and ranks_sorted)

    # -- kernels (CUDA only: the product has no CPU path)
    def _pack(self, vec):
        L.check(self._lib.b200mf_ghost_pack(self.number_code, _ptr(self.buf), _ptr(vec), _ptr(self.import_idx),
                                            self.part.n_import, C.c_void_p(torch.cuda.current_stream().cuda_stream)))

    def _unpack_add(self, vec):
        L.check(self._lib.b200mf_ghost_unpack_add(self.number_code, _ptr(vec), _ptr(self.buf), _ptr(self.import_idx),
                                                  self.part.n_import, C.c_void_p(torch.cuda.current_stream().cuda_stream)))

    def _transfer(self, vec, to_ghosts):
        if self._use_a2a:
            p = self.part
            ghost = vec[p.n_owned:p.n_owned + p.n_ghost]
            packed = self.buf[:p.n_import]
            if to_ghosts:
                w = dist.all_to_all_single(ghost, packed, self.ghost_splits, self.import_splits,
                                           group=self.group, async_op=True)
            else:
                w = dist.all_to_all_single(packed, ghost, self.import_splits, self.ghost_splits,
                                           group=self.group, async_op=True)
            return [w]
        ops = []
        for r, a, b in self.ghost_slices:          # my ghost section <-> its owner
            ops.append(dist.P2POp(dist.irecv if to_ghosts else dist.isend, vec[a:b], r, group=self.group))
        for r, a, b in self.import_slices:         # my packed owned values <-> the rank ghosting them
            ops.append(dist.P2POp(dist.isend if to_ghosts else dist.irecv, self.buf[a:b], r, group=self.group))
        return dist.batch_isend_irecv(ops) if ops else []

    def update_ghost_values_start(self, vec):
        self._pack(vec)
        return self._transfer(vec, True)

    def update_ghost_values_finish(self, works):
        for w in works:
            w.wait()

    def update_ghost_values(self, vec):
        self.update_ghost_values_finish(self.update_ghost_values_start(vec))

    def compress_start(self, vec):
        return self._transfer(vec, False)

    def compress_finish(self, vec, works):
        for w in works:
            w.wait()
        self._unpack_add(vec)
        self.zero_out_ghost_values(vec)

    def compress(self, vec):
        self.compress_finish(vec, self.compress_start(vec))

    def zero_out_ghost_values(self, vec):
        if self.part.n_ghost:
            vec[self.part.n_owned:].zero_()


class Communicator:
    """b200mf_comm: one NCCL communicator per process, created from a unique id that rank 0
    generates and torch.distributed (plumbing only) broadcasts -- a C++ host would MPI_Bcast it."""

    def __init__(self, device, group=None):
        self._lib = L.load()
        self.n_ranks = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        ident = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            buf = (C.c_ubyte * 128)()
            L.check(self._lib.b200mf_comm_get_unique_id(buf))
            ident = torch.tensor(list(buf), dtype=torch.uint8)
        if self.n_ranks > 1:
            ident = ident.to(device)
            dist.broadcast(ident, 0, group=group)
            ident = ident.cpu()
        raw = (C.c_ubyte * 128)(*ident.tolist())
        self._h = C.c_void_p()
        torch.cuda.set_device(device)
        L.check(self._lib.b200mf_comm_create(raw, self.n_ranks, self.rank, C.byref(self._h)))

    def allreduce_sum(self, t):
        L.check(self._lib.b200mf_comm_allreduce_sum(self._h, _ptr(t), t.numel(),
                                                     C.c_void_p(torch.cuda.current_stream().cuda_stream)))

    def close(self):
        if self._h:
            self._lib.b200mf_comm_destroy(self._h)
            self._h = None


class CPartitioner:
    """b200mf_partitioner: Utilities::MPI::Partitioner + the ghost exchange of LA::d::Vector, in C."""

    def __init__(self, comm, rank_offsets, ghost_global, number):
        self._lib, self.comm = L.load(), comm
        ro = np.ascontiguousarray(rank_offsets, dtype=np.uint64)
        gg = np.ascontiguousarray(ghost_global, dtype=np.uint64)
        self._h = C.c_void_p()
        L.check(self._lib.b200mf_partitioner_create(comm._h, ro.ctypes.data_as(C.POINTER(C.c_uint64)),
                                                    gg.ctypes.data_as(C.POINTER(C.c_uint64)), len(gg),
                                                    L.F64 if number == "f64" else L.F32, C.byref(self._h)))
        info = L.PartitionerInfo()
        L.check(self._lib.b200mf_partitioner_get_info(self._h, C.byref(info)))
        self.n_owned, self.n_ghost, self.n_import = int(info.n_owned), int(info.n_ghost), int(info.n_import)
        self.n_ranks, self.rank = comm.n_ranks, comm.rank
        self.ghost_targets = [(info.ghost_target_ranks[i], int(info.ghost_target_counts[i]))
                              for i in range(info.n_ghost_targets)]
        self.import_targets = [(info.import_target_ranks[i], int(info.import_target_counts[i]))
                               for i in range(info.n_import_targets)]

    def update_ghost_values(self, vec):
        L.check(self._lib.b200mf_update_ghost_values(self._h, _ptr(vec), C.c_void_p(torch.cuda.current_stream().cuda_stream)))

    def compress(self, vec):
        L.check(self._lib.b200mf_compress_add(self._h, _ptr(vec), C.c_void_p(torch.cuda.current_stream().cuda_stream)))

    def zero_out_ghost_values(self, vec):
        L.check(self._lib.b200mf_zero_out_ghost_values(self._h, _ptr(vec), C.c_void_p(torch.cuda.current_stream().cuda_stream)))


class DistributedMatrixFree:
    """Portable::MatrixFree on a partitioned mesh.  Everything on the data path -- partitioner, ghost
    exchange over NCCL, the overlapped cell loop, the distributed CG -- lives behind the C ABI
    (csrc/comm.cu: b200mf_comm_*, b200mf_partitioner_*, b200mf_dist_*); this class only binds it."""

    def __init__(self, mesh, number="f64", device="cuda:0", group=None, overlap=True, comm=None):
        self.mesh, self.group = mesh, group
        self.mf = MatrixFree(number, device).reinit_from_mesh(mesh)
        self.comm = comm if comm is not None else Communicator(self.mf.device, group)
        self.partitioner = CPartitioner(self.comm, mesh.rank_offsets, mesh.ghost_global, number)
        self.n_owned, self.n_ghost = mesh.n_owned, mesh.n_ghost
        self.n_cells, self.n_interior = mesh.n_cells, mesh.n_cells_interior
        self._lib = L.load()

    def initialize_dof_vector(self):
        return self.mf.initialize_dof_vector()

    def get_vector_partitioner(self):
        return self.partitioner

    def vmult(self, op, dst, src):
        """dst = A src on distributed vectors (b200mf_dist_vmult)."""
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        L.check(self._lib.b200mf_dist_vmult(self.mf._h, self.partitioner._h, C.byref(op), _ptr(dst), _ptr(src), st))

    def vmult_host_batch(self, op, dst_hosts, src_hosts):
        """dst_hosts[k] = A src_hosts[k] for lists of pinned host numpy arrays (owned part)."""
        n = len(src_hosts)
        dp = (C.c_void_p * n)(*[d.ctypes.data for d in dst_hosts])
        sp = (C.c_void_p * n)(*[s.ctypes.data for s in src_hosts])
        L.check(self._lib.b200mf_dist_vmult_host_batch(self.mf._h, self.partitioner._h, C.byref(op), n, dp, sp))

    def compute_diagonal(self, op):
        """MatrixFreeTools::compute_diagonal + compress(add); returns the inverse diagonal."""
        diag = self.initialize_dof_vector()
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        L.check(self._lib.b200mf_dist_compute_diagonal(self.mf._h, self.partitioner._h, C.byref(op), _ptr(diag), st))
        inv = torch.zeros_like(diag)
        inv[:self.n_owned] = 1.0 / diag[:self.n_owned]
        return inv


def solve_cg(dmf, op, x, b, inverse_diagonal, tolerance, max_iterations, check_every=1):
    """SolverCG with Jacobi (or no) preconditioner on distributed vectors (b200mf_dist_cg_solve);
    returns (iterations, residual, converged)."""
    lib = L.load()
    sd = L.SolverDesc()
    sd.tolerance, sd.max_iterations, sd.check_every = tolerance, max_iterations, check_every
    if inverse_diagonal is not None:
        sd.preconditioner, sd.inverse_diagonal = L.PRECOND_JACOBI, _ptr(inverse_diagonal)
    else:
        sd.preconditioner = L.PRECOND_NONE
    res = L.SolverResult()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    code = lib.b200mf_dist_cg_solve(dmf.mf._h, dmf.partitioner._h, C.byref(op), C.byref(sd), _ptr(x), _ptr(b),
                                    C.byref(res), st)
    if code not in (L.OK, L.ERR_NOCONVERGENCE):
        L.check(code)
    return res.iterations, res.residual, code == L.OK


class DistributedGeometricMultigrid:
    """step-37's multigrid on a partitioned mesh (b200mf_mg_* with per-level partitioners): every rank holds
    its coarse cell(s) of hyper_rectangle(coarse) refined l times for l = min_level..refinements; the levels'
    operators, smoothers and transfers exchange ghosts through the C partitioners, the dot products of the
    outer CG are all-reduced.  Names as in dealii_b200.multigrid.GeometricMultigrid."""

    def __init__(self, dim, degree, refinements, n_ranks, rank, coarse=(1, 1, 1), number="f32", device="cuda:0",
                 group=None, comm=None, min_level=0, smoother_degree=5, smoothing_range=15.0,
                 eig_cg_n_iterations=10, coarse_tolerance=1e-3, safety_factor=1.2):
        import dealii_b200
        self._lib = L.load()
        self.comm = comm if comm is not None else Communicator(torch.device(device), group)
        self.levels, self.level_operators, tables = [], [], []
        for level in range(min_level, refinements + 1):
            mesh = PartitionedHyperCubeMesh(dim, degree, level, n_ranks, rank, coarse=coarse, dirichlet_boundary=True,
                                            mark_constrained_l2g=True, ghost_mode="touched")
            dmf = DistributedMatrixFree(mesh, number, device, group, comm=self.comm)
            self.levels.append(dmf)
            self.level_operators.append(dealii_b200.LaplaceOperator(dmf.mf))
        for lc, lf in zip(self.levels[:-1], self.levels[1:]):
            pos_c, pos_f = lc.mesh.cell_morton_position, lf.mesh.cell_morton_position
            inv_f = np.empty(len(pos_f), dtype=np.int64)
            inv_f[pos_f.astype(np.int64)] = np.arange(len(pos_f))
            kids = (pos_c.astype(np.int64)[:, None] << dim) + np.arange(1 << dim)[None, :]
            tables.append(np.ascontiguousarray(inv_f[kids], dtype=np.uint32))
        self._tables = tables
        n = len(self.levels)
        desc = L.MgDesc()
        desc.n_levels = n
        self._setups = (C.c_void_p * n)(*[lv.mf._h for lv in self.levels])
        self._ops = (L.Operator * n)(*[op.op for op in self.level_operators])
        self._parts = (C.c_void_p * n)(*[lv.partitioner._h for lv in self.levels])
        self._children = (C.c_void_p * max(n - 1, 1))(*[t.ctypes.data for t in tables])
        desc.levels = C.cast(self._setups, C.POINTER(C.c_void_p))
        desc.operators = C.cast(self._ops, C.POINTER(L.Operator))
        desc.child_cells = C.cast(self._children, C.POINTER(C.c_void_p)) if n > 1 else None
        desc.partitioners = C.cast(self._parts, C.POINTER(C.c_void_p))
        desc.smoother_degree, desc.smoothing_range = smoother_degree, smoothing_range
        desc.eig_cg_n_iterations, desc.coarse_tolerance = eig_cg_n_iterations, coarse_tolerance
        desc.safety_factor = safety_factor
        self._h = C.c_void_p()
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        L.check(self._lib.b200mf_mg_create(C.byref(desc), C.byref(self._h), st))

    def n_levels(self):
        return len(self.levels)

    def level_info(self, level):
        info = L.MgLevelInfo()
        L.check(self._lib.b200mf_mg_get_level_info(self._h, level, C.byref(info)))
        return info

    def vmult(self, dst, src):
        code = L.F64 if src.dtype == torch.float64 else L.F32
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        L.check(self._lib.b200mf_mg_vcycle(self._h, code, _ptr(dst), _ptr(src), st))

    def solve(self, system, op, x, b, tolerance, max_steps=100):
        """SolverCG(SolverControl(max_steps, tolerance)).solve(A, x, b, PreconditionMG) with A = `op` on the
        DistributedMatrixFree `system` (finest level mesh); returns (iterations, residual, converged)."""
        res = L.SolverResult()
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        code = self._lib.b200mf_mg_dist_cg_solve(self._h, system.mf._h, system.partitioner._h, C.byref(op),
                                                 float(tolerance), int(max_steps), _ptr(x), _ptr(b), C.byref(res), st)
        if code not in (L.OK, L.ERR_NOCONVERGENCE):
            L.check(code)
        return res.iterations, res.residual, code == L.OK

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                self._lib.b200mf_mg_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass
