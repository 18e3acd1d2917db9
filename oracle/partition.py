"""Distributed DoF numbering of parallel::distributed::Triangulation + DoFHandler -- oracle
restatement in numpy (small meshes, all ranks at once).

Follows source/dofs/dof_handler_policy.cc:3644-3760 (ParallelDistributed::distribute_dofs):
  1. the active cells in Morton order are cut into n_ranks equal contiguous chunks (p4est
     partition of a uniformly refined forest);
  2. every rank runs the serial first-touch numbering (:1676-1719) over ITS cells;
  3. dofs on interfaces to ghost cells of a LOWER rank are invalidated (:3695-3704, the lowest
     subdomain id touching a dof owns it, helper :1739) and the rest renumbered compactly in
     the same order;
  4. the ranks' ranges are concatenated in rank order (exscan shift, :3720-3745).
Ghost sets: Portable::MatrixFree uses all locally relevant dofs = dofs of own + ghost cells
(matrix_free/portable_matrix_free.templates.h:1326-1343), ghost cell = cell of another rank
sharing at least a vertex with an own cell.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import numpy as np


def distributed_numbering(mesh, n_ranks):
    """mesh: oracle.mesh.HyperCubeMesh (serial, Morton order).  Returns dict with
    'global_of_serial' (serial dof number -> distributed global number), 'rank_offsets',
    'cell_rank'."""
    nc = mesh.n_cells
    assert nc % n_ranks == 0
    chunk = nc // n_ranks
    cell_rank = np.arange(nc) // chunk
    l2g_h = mesh.l2g_hier                                  # (nc, npc) serial numbers, hierarchical order
    owner = np.full(mesh.n_dofs, n_ranks, dtype=np.int64)
    np.minimum.at(owner, l2g_h.ravel(), np.repeat(cell_rank, l2g_h.shape[1]))
    glob = np.full(mesh.n_dofs, -1, dtype=np.int64)
    offsets = [0]
    for r in range(n_ranks):
        flat = l2g_h[cell_rank == r].ravel()               # first-touch order of rank r
        flat = flat[owner[flat] == r]
        _, first = np.unique(flat, return_index=True)
        order = flat[np.sort(first)]
        glob[order] = offsets[-1] + np.arange(len(order))
        offsets.append(offsets[-1] + len(order))
    assert (glob >= 0).all()
    return {"global_of_serial": glob, "rank_offsets": np.array(offsets), "cell_rank": cell_rank}


def relevant_dofs(mesh, numbering, rank, mode="relevant"):
    """Sorted global numbers of the ghost dofs of `rank` (owned ones removed)."""
    cell_rank, glob = numbering["cell_rank"], numbering["global_of_serial"]
    own_cells = np.nonzero(cell_rank == rank)[0]
    if mode == "touched":
        cells = own_cells
    else:
        # ghost cells: other ranks' cells sharing a vertex (= a lattice corner point) with an own cell
        ijk = mesh.cell_ijk
        own = set(map(tuple, ijk[own_cells]))
        keep = []
        for c in range(mesh.n_cells):
            t = ijk[c]
            nb = False
            for dx in (-1, 0, 1):
                for dy in (-1, 0, 1):
                    for dz in ((-1, 0, 1) if mesh.dim == 3 else (0,)):
                        q = (t[0] + dx, t[1] + dy) + ((t[2] + dz,) if mesh.dim == 3 else ())
                        if q in own:
                            nb = True
            if nb:
                keep.append(c)
        cells = np.array(keep)
    g = np.unique(glob[mesh.l2g[cells].ravel()])
    lo, hi = numbering["rank_offsets"][rank], numbering["rank_offsets"][rank + 1]
    return g[(g < lo) | (g >= hi)]
