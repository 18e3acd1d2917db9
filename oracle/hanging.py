"""Hanging-node meshes for the oracle: a geometric restatement of what deal.II's setup hands to
Portable::MatrixFree on an adaptively refined mesh, and the conforming operator to compare with.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The mesh is a grid of unit coarse cells, some refined once (2:1 balanced by construction).  Two
independent descriptions of the same finite element space are built:

 * the *algebraic* one used as the expected answer: every support point of every active cell is a
   node; a node is "hanging" iff it belongs to refined cells only and lies in an unrefined cell;
   its value is the interpolant of that (coarse) cell's basis functions
   (AffineConstraints from DoFTools::make_hanging_node_constraints,
   source/dofs/dof_tools_constraints.cc; FE_Q interpolation include/deal.II/fe/fe_q_base.h).
   A_conforming = C^T A_all-nodes C.

 * the *index + mask* one the device kernel consumes (matrix_free/hanging_nodes_internal.h:40-60,
   520-880: ConstraintKinds bits subcell_x..z | face_x..z | edge_x..z, and the dof indices of a
   refined cell's constrained faces / edges redirected to the coarse neighbour's dofs at the same
   local position), resolved per cell by Portable::internal::resolve_hanging_nodes
   (matrix_free/portable_hanging_nodes_internal.h:124-459).
"""
import itertools

import numpy as np
import scipy.sparse as sp

from .shape import gauss_lobatto_points, lagrange_values_and_derivatives


class HangingNodeMesh:
    def __init__(self, dim, degree, coarse_shape, refined):
        """coarse_shape: cells per direction; refined: iterable of coarse index tuples (x, y[, z])."""
        self.dim, self.degree = dim, degree
        n = degree + 1
        self.n = n
        shape = tuple(coarse_shape)
        refined = {tuple(r) for r in refined}
        gll = np.asarray(gauss_lobatto_points(n), dtype=np.float64)
        # active cells: (lo, h, parent coarse index, child position or None)
        cells = []
        for idx in itertools.product(*[range(s) for s in shape[::-1]]):
            c = idx[::-1]
            if c in refined:
                for child in itertools.product(*[range(2)] * dim):
                    cp = child[::-1]                      # x fastest
                    lo = np.array([c[d] + 0.5 * cp[d] for d in range(dim)])
                    cells.append((lo, 0.5, c, cp))
            else:
                cells.append((np.array(c, dtype=np.float64), 1.0, c, None))
        self.cells = cells
        self.n_cells = len(cells)
        self.refined, self.shape = refined, shape
        self.coarse_cell_id = {c[2]: i for i, c in enumerate(cells) if c[3] is None}

        # nodes = distinct support points of all active cells
        lex = np.array([[(i // n ** d) % n for d in range(dim)] for i in range(n ** dim)])
        self.lex = lex
        key_of, coords = {}, []
        cell_nodes = np.zeros((self.n_cells, n ** dim), dtype=np.int64)
        for ci, (lo, h, _, _) in enumerate(cells):
            pts = lo[None, :] + h * gll[lex]
            for li, x in enumerate(pts):
                key = tuple(np.round(x * 2 ** 24).astype(np.int64))
                if key not in key_of:
                    key_of[key] = len(coords)
                    coords.append(x)
                cell_nodes[ci, li] = key_of[key]
        self.cell_nodes = cell_nodes
        self.node_coords = np.array(coords)
        n_nodes = len(coords)

        # hanging nodes and their interpolation weights
        # (a node is hanging iff only refined cells own it and an unrefined cell contains it
        # without having it as a support point; the coarse side's own support points are dofs)
        hanging = {}
        tol = 1e-9
        owned_by_coarse = np.zeros(n_nodes, dtype=bool)
        for ci, (lo, h, _, cp) in enumerate(cells):
            if cp is None:
                owned_by_coarse[cell_nodes[ci]] = True
        for ci, (lo, h, _, cp) in enumerate(cells):
            if cp is not None:
                continue
            inside = np.all((self.node_coords >= lo - tol) & (self.node_coords <= lo + h + tol), axis=1)
            for nd in np.nonzero(inside & ~owned_by_coarse)[0]:
                xi = (self.node_coords[nd] - lo) / h
                w = np.ones(n ** dim)
                for d in range(dim):
                    v, _ = lagrange_values_and_derivatives(gll, np.array([xi[d]]))
                    w = w * np.asarray(v[:, 0], dtype=np.float64)[lex[:, d]]
                hanging[nd] = (cell_nodes[ci], w)
        self.hanging = hanging
        real = np.array([i for i in range(n_nodes) if i not in hanging], dtype=np.int64)
        self.n_dofs = len(real)
        dof_of_node = -np.ones(n_nodes, dtype=np.int64)
        dof_of_node[real] = np.arange(len(real))
        self.dof_of_node = dof_of_node
        rows, cols, vals = list(range(len(real))), list(range(len(real))), [1.0] * len(real)
        rows = list(real)
        for nd, (src_nodes, w) in hanging.items():
            for s, wt in zip(src_nodes, w):
                if abs(wt) > 1e-15:
                    assert dof_of_node[s] >= 0, "constraint chain: mesh is not 2:1 balanced"
                    rows.append(nd)
                    cols.append(dof_of_node[s])
                    vals.append(wt)
        self.C = sp.csr_matrix((vals, (rows, cols)), shape=(n_nodes, len(real)))

        # vertices of the active cells (Q1 geometry), lexicographic
        corner = np.array([[(v >> d) & 1 for d in range(dim)] for v in range(2 ** dim)], dtype=np.float64)
        self.cell_vertices = np.array([lo[None, :] + h * corner for lo, h, _, _ in cells])
        self._build_device_description()

    # the full-node mesh MatrixFreeOracle consumes (l2g = node numbers, no constraints)
    def all_nodes_mesh(self):
        class M:
            pass
        m = M()
        m.dim, m.degree, m.n_cells = self.dim, self.degree, self.n_cells
        m.cell_vertices, m.l2g, m.n_dofs = self.cell_vertices, self.cell_nodes, len(self.node_coords)
        return m

    def _coarse_neighbour(self, parent, offset):
        c = tuple(parent[d] + offset[d] for d in range(self.dim))
        if any(c[d] < 0 or c[d] >= self.shape[d] for d in range(self.dim)):
            return None
        return self.coarse_cell_id.get(c)      # None if that cell is refined

    def _build_device_description(self):
        dim, n, p = self.dim, self.n, self.degree
        l2g = np.zeros((self.n_cells, n ** dim), dtype=np.uint32)
        mask = np.zeros(self.n_cells, dtype=np.uint16)
        for ci, (lo, h, parent, cp) in enumerate(self.cells):
            if cp is None:
                assert (self.dof_of_node[self.cell_nodes[ci]] >= 0).all()
                l2g[ci] = self.dof_of_node[self.cell_nodes[ci]]
                continue
            side = [-1 if cp[d] == 0 else 1 for d in range(dim)]     # outer side of the parent
            m = 0
            for d in range(dim):
                if cp[d] == 0:
                    m |= 1 << d                                      # subcell bit: low child
            face = [None] * dim
            for d in range(dim):
                off = [0] * dim
                off[d] = side[d]
                face[d] = self._coarse_neighbour(parent, off)
                if face[d] is not None:
                    m |= 8 << d
            edge = [None] * dim
            if dim == 3:
                for e in range(3):
                    d1, d2 = (e + 1) % 3, (e + 2) % 3
                    if face[d1] is None and face[d2] is None:
                        off = [0, 0, 0]
                        off[d1], off[d2] = side[d1], side[d2]
                        edge[e] = self._coarse_neighbour(parent, off)
                        if edge[e] is not None:
                            m |= 64 << e
            # constraints exist only if something is constrained (a zero mask means "plain cell")
            if (m >> 3) == 0:
                m = 0
            mask[ci] = m
            for li, ijk in enumerate(self.lex):
                outer = [ijk[d] == (0 if cp[d] == 0 else p) for d in range(dim)]
                src_cell, flip = None, []
                for d in range(dim):
                    if face[d] is not None and outer[d]:
                        src_cell, flip = face[d], [d]
                        break
                if src_cell is None and dim == 3:
                    for e in range(3):
                        d1, d2 = (e + 1) % 3, (e + 2) % 3
                        if edge[e] is not None and outer[d1] and outer[d2]:
                            src_cell, flip = edge[e], [d1, d2]
                            break
                if src_cell is None:
                    dof = self.dof_of_node[self.cell_nodes[ci, li]]
                    assert dof >= 0, "hanging node outside a constrained face / edge"
                    l2g[ci, li] = dof
                else:
                    q = list(ijk)
                    for d in flip:
                        q[d] = p - q[d]        # the neighbour touches us with its opposite side
                    lj = sum(q[d] * n ** d for d in range(dim))
                    dof = self.dof_of_node[self.cell_nodes[src_cell, lj]]
                    assert dof >= 0
                    l2g[ci, li] = dof
        self.l2g, self.constraint_mask = l2g, mask
