"""Synthetic hyper-cube meshes with deal.II cell ordering and DoF numbering -- oracle.

Restates, for ``GridGenerator::hyper_cube`` + ``refine_global(r)`` (or
``subdivided_hyper_cube``) with ``FE_Q(p)``:
  * active-cell order: children are created parent by parent, child ``c`` sits at
    offset ``(c&1, c>>1&1, c>>2&1)`` (GeometryInfo<dim>::child_cell_on_face /
    unit_cell_vertex, include/deal.II/base/geometry_info.h) => Morton (z-order)
    for global refinement; ``subdivided_hyper_cube`` creates cells lexicographically
    (source/grid/grid_generator.cc, subdivided_hyper_rectangle).
  * DoF numbering: ``DoFHandler::distribute_dofs`` = first-touch numbering, cell by
    cell in active order, per cell vertices -> lines -> quads -> interior in the
    element's hierarchical order (source/dofs/dof_handler_policy.cc:1676-1719,
    process_dof_indices in include/deal.II/dofs/dof_accessor.templates.h).
  * hierarchical -> lexicographic map of FE_Q
    (include/deal.II/fe/fe_tools.templates.h:2987-3152).
  * cell-local index lists in lexicographic order, the layout
    Portable::MatrixFree stores (portable_matrix_free.templates.h:292-298).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import numpy as np

from .shape import gauss_lobatto_points


def hierarchic_local_offsets(dim, degree):
    """(n_dofs_per_cell, dim) integer offsets in [0,p]^dim of the FE_Q(p) dofs listed
    in deal.II's hierarchical order (vertices, lines, quads, hex).  Equivalent to
    FETools::hierarchic_to_lexicographic_numbering (fe_tools.templates.h:2987) with
    lex = ax + n*ay + n^2*az."""
    p = degree
    inner = range(1, p)
    out = []
    if dim == 1:
        out = [(0,), (p,)] + [(a,) for a in inner]
    elif dim == 2:
        for v in range(4):
            out.append(((v & 1) * p, (v >> 1 & 1) * p))
        for x in (0, p):                      # lines 0,1: x fixed, run along y
            out += [(x, a) for a in inner]
        for y in (0, p):                      # lines 2,3: y fixed, run along x
            out += [(a, y) for a in inner]
        out += [(a, b) for b in inner for a in inner]   # interior: x fastest
    elif dim == 3:
        for v in range(8):
            out.append(((v & 1) * p, (v >> 1 & 1) * p, (v >> 2 & 1) * p))
        for z in (0, p):                      # lines 0-3 (z=0), 4-7 (z=1)
            for x in (0, p):
                out += [(x, a, z) for a in inner]
            for y in (0, p):
                out += [(a, y, z) for a in inner]
        for (x, y) in ((0, 0), (p, 0), (0, p), (p, p)):   # lines 8-11 along z
            out += [(x, y, a) for a in inner]
        for x in (0, p):                      # faces 0,1: outer z, inner y
            out += [(x, b, a) for a in inner for b in inner]
        for y in (0, p):                      # faces 2,3: outer x, inner z
            out += [(a, y, b) for a in inner for b in inner]
        for z in (0, p):                      # faces 4,5: outer y, inner x
            out += [(b, a, z) for a in inner for b in inner]
        out += [(c, b, a) for a in inner for b in inner for c in inner]
    else:
        raise ValueError(dim)
    out = np.array(out, dtype=np.int64).reshape(-1, dim)
    assert out.shape[0] == (p + 1) ** dim
    return out


def hierarchic_to_lexicographic(dim, degree):
    off = hierarchic_local_offsets(dim, degree)
    n = degree + 1
    return sum(off[:, d] * n ** d for d in range(dim))


def morton_cell_coords(dim, refinements):
    """Integer cell coordinates of the active cells after refine_global(r), in active
    cell iterator order."""
    n_cells = (2 ** dim) ** refinements
    c = np.arange(n_cells, dtype=np.int64)
    ijk = np.zeros((n_cells, dim), dtype=np.int64)
    for level in range(refinements):
        child = (c >> (dim * level)) & (2 ** dim - 1)
        for d in range(dim):
            ijk[:, d] |= ((child >> d) & 1) << level
    return ijk


def lexicographic_cell_coords(dim, subdivisions):
    n_cells = subdivisions ** dim
    c = np.arange(n_cells, dtype=np.int64)
    ijk = np.zeros((n_cells, dim), dtype=np.int64)
    for d in range(dim):
        ijk[:, d] = (c // subdivisions ** d) % subdivisions
    return ijk


class HyperCubeMesh:
    """hyper_cube(left,right) refined globally ``refinements`` times (Morton order) or
    ``subdivided_hyper_cube(subdivisions)`` (lexicographic order), FE_Q(degree).

    ``deformation``: optional callable mapping an (m, dim) array of vertex
    coordinates to displaced coordinates (GridTools::transform); cells then use the
    multilinear MappingQ1 geometry of the displaced vertices.
    """

    def __init__(self, dim, degree, refinements=None, subdivisions=None,
                 left=0.0, right=1.0, deformation=None, cell_coords=None,
                 cells_per_dim=None):
        self.dim, self.degree = dim, degree
        p, n = degree, degree + 1
        if cell_coords is not None:
            self.N = int(cells_per_dim)
            self.cell_ijk = np.asarray(cell_coords, dtype=np.int64)
        elif refinements is not None:
            self.N = 2 ** refinements
            self.cell_ijk = morton_cell_coords(dim, refinements)
        else:
            self.N = subdivisions
            self.cell_ijk = lexicographic_cell_coords(dim, subdivisions)
        self.left, self.right = left, right
        self.n_cells = self.cell_ijk.shape[0]
        self.dofs_per_cell = n ** dim
        L = self.N * p + 1                      # lattice points per direction
        self.lattice_size = L

        # --- first-touch numbering on the lattice --------------------------------
        hier = hierarchic_local_offsets(dim, degree)          # (npc, dim)
        strides = np.array([L ** d for d in range(dim)], dtype=np.int64)
        base = (self.cell_ijk * p) @ strides                  # (n_cells,)
        hier_lin = hier @ strides                             # (npc,)
        flat = (base[:, None] + hier_lin[None, :]).ravel()    # cell-major, hierarchical
        uniq, first = np.unique(flat, return_index=True)
        order = np.argsort(first, kind="stable")
        number_of_lattice = np.full(L ** dim, -1, dtype=np.int64)
        number_of_lattice[uniq[order]] = np.arange(len(uniq))
        self.n_dofs = len(uniq)
        self._number_of_lattice = number_of_lattice

        # --- cell-local index lists, lexicographic (x fastest) -------------------
        lex = np.stack(np.meshgrid(*[np.arange(n)] * dim, indexing="ij"), -1)
        lex = lex.reshape(-1, dim)
        # meshgrid 'ij' makes the first axis slowest; we need x fastest
        lex = lex[:, ::-1] if dim > 1 else lex
        lex_lin = lex @ strides
        self.l2g = number_of_lattice[base[:, None] + lex_lin[None, :]]   # (n_cells, npc)
        self.l2g_hier = number_of_lattice[base[:, None] + hier_lin[None, :]]

        # --- geometry -------------------------------------------------------------
        h = (right - left) / self.N
        vert_off = np.array([[(v >> d) & 1 for d in range(dim)]
                             for v in range(2 ** dim)], dtype=np.int64)
        vijk = self.cell_ijk[:, None, :] + vert_off[None, :, :]
        verts = left + h * vijk.astype(np.float64)
        if deformation is not None:
            verts = deformation(verts.reshape(-1, dim)).reshape(verts.shape)
        self.cell_vertices = verts                              # (n_cells, 2^dim, dim)
        self.deformed = deformation is not None

        # --- lattice coordinates of every dof (support points) -------------------
        lat = np.nonzero(number_of_lattice >= 0)[0]
        coords = np.zeros((self.n_dofs, dim), dtype=np.int64)
        for d in range(dim):
            coords[number_of_lattice[lat], d] = (lat // strides[d]) % L
        self.dof_lattice = coords
        self.boundary_dofs = np.nonzero(
            np.any((coords == 0) | (coords == L - 1), axis=1))[0]

    def support_points(self):
        """Physical support point of every dof (undeformed meshes only need the 1D
        node positions; deformed meshes map through the cell's MappingQ1)."""
        p = self.degree
        nodes = gauss_lobatto_points(p + 1)
        pts = np.zeros((self.n_dofs, self.dim))
        n = p + 1
        lex = np.array([[(i // n ** d) % n for d in range(self.dim)]
                        for i in range(self.dofs_per_cell)])
        ref = nodes[lex]                                         # (npc, dim) in [0,1]
        phys = map_q1(self.cell_vertices, ref)                   # (n_cells, npc, dim)
        pts[self.l2g.ravel()] = phys.reshape(-1, self.dim)
        return pts


def q1_shape(ref):
    """Multilinear vertex shape functions and their gradients at reference points
    ``ref`` (m, dim); vertex v at ((v&1),(v>>1&1),(v>>2&1))  (MappingQ1 = MappingQ(1),
    include/deal.II/fe/mapping_q.h)."""
    m, dim = ref.shape
    nv = 2 ** dim
    N = np.ones((m, nv))
    dN = np.ones((m, nv, dim))
    for v in range(nv):
        for d in range(dim):
            b = (v >> d) & 1
            f = ref[:, d] if b else 1.0 - ref[:, d]
            df = np.full(m, 1.0 if b else -1.0)
            N[:, v] *= f
            for e in range(dim):
                dN[:, v, e] *= df if e == d else f
    return N, dN


def map_q1(cell_vertices, ref):
    N, _ = q1_shape(ref)
    return np.einsum("qv,cvd->cqd", N, cell_vertices)


def jacobians_q1(cell_vertices, ref):
    """J[c,q,d,e] = d x_d / d xi_e."""
    _, dN = q1_shape(ref)
    return np.einsum("qve,cvd->cqde", dN, cell_vertices)
