"""resolve_hanging_nodes -- numpy restatement of the reference's device kernel
(include/deal.II/matrix_free/portable_hanging_nodes_internal.h:124-459: interpolate_boundary_2d,
interpolate_boundary_3d, is_constrained_dof_2d/3d, resolve_hanging_nodes), driven by the
ConstraintKinds bit mask of matrix_free/hanging_nodes_internal.h:40-60.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Pinned on the reference's golden output
tests/matrix_free/hanging_node_kernels_01.output (tests/test_oracle_golden.py).
"""
import numpy as np

SUBCELL = (1 << 0, 1 << 1, 1 << 2)
FACE = (1 << 3, 1 << 4, 1 << 5)
EDGE = (1 << 6, 1 << 7, 1 << 8)


def resolve_hanging_nodes(values, mask, dim, degree, weights, transpose):
    """values: (degree+1)^dim local values, lexicographic (x fastest); weights[i, j] =
    subface_interpolation_matrices[0] (child node i <- parent basis j).  Returns the new values."""
    n = degree + 1
    v = np.array(values, dtype=np.float64).copy()
    W = np.asarray(weights, dtype=np.float64)
    for direction in range(dim):
        tmp = v.copy()
        for q in range(n ** dim):
            idx = [(q // n ** d) % n for d in range(dim)]
            if dim == 2:
                other = 1 - direction
                # portable_hanging_nodes_internal.h:140-149 (constrained_face), :50-72 (dof)
                constrained_face = bool(mask & FACE[other])
                on = (idx[other] == 0) if (mask & SUBCELL[other]) else (idx[other] == degree)
                constrained = constrained_face and on
            else:
                d1, d2 = (direction + 1) % 3, (direction + 2) % 3      # face1 / face2 directions
                constrained_face = bool(mask & (FACE[d1] | FACE[d2] | EDGE[direction]))
                on1 = (idx[d1] == 0) if (mask & SUBCELL[d1]) else (idx[d1] == degree)
                on2 = (idx[d2] == 0) if (mask & SUBCELL[d2]) else (idx[d2] == degree)
                dof = ((mask & FACE[d1]) and on1) or ((mask & FACE[d2]) and on2) or \
                      ((mask & EDGE[direction]) and on1 and on2)
                constrained = constrained_face and bool(dof)
            if not constrained:
                continue
            first = bool(mask & SUBCELL[direction])
            interp = idx[direction]
            s = 0.0
            for i in range(n):
                j = list(idx)
                j[direction] = i
                real = sum(j[d] * n ** d for d in range(dim))
                if first:
                    w = W[i, interp] if transpose else W[interp, i]
                else:
                    w = W[degree - i, degree - interp] if transpose else W[degree - interp, degree - i]
                s += w * v[real]
            tmp[q] = s
        v = tmp
    return v


def golden_cases():
    """(dim, degree, mask) in the order tests/matrix_free/hanging_node_kernels_01.cc:512-726
    calls test<dim>(degree, mask)."""
    sx, sy, sz = SUBCELL
    fx, fy, fz = FACE
    ex, ey, ez = EDGE
    out = []
    for degree in (1, 2, 3):
        out += [(2, degree, m) for m in (0, fx | sx | sy, fx | sx, fx | sy, fx,
                                         fy | sy | sx, fy | sy, fy | sx, fy)]
    for degree in (1, 2, 3):
        masks = [ex | sy | sz, ex | sy | sz | sx, ex | sz, ex | sz | sx, ex | sy, ex | sy | sx, ex, ex | sx,
                 ey | sx | sz, ey | sx | sz | sy, ey | sz, ey | sz | sy, ey | sx, ey | sx | sy, ey, ey | sy,
                 ez | sx | sy, ez | sx | sy | sz, ez | sy, ez | sy | sz, ez | sx, ez | sx | sz, ez, ez | sz,
                 fx | sx, fx, fy | sy, fy, fz | sz, fz]
        out += [(3, degree, m) for m in masks]
    return out


def parse_golden(path):
    """Blocks of (input, reference result, optimised result) in file order; per test<>() call one
    block for the interpolation and one for its transpose."""
    groups, cur = [], []
    for line in open(path):
        line = line.strip()
        if line.startswith("DEAL:0::"):
            line = line[len("DEAL:0::"):].strip()
        if not line:
            if cur:
                groups.append(cur)
                cur = []
            continue
        cur.append(np.array([float(t) for t in line.split()]))
    if cur:
        groups.append(cur)
    return groups
