"""Utilities::MPI::Partitioner index algebra -- oracle restatement (all ranks at once).

Follows ``source/base/partitioner.cc:185-330`` (Partitioner::set_ghost_indices): every
rank owns one contiguous global range; its ghost IndexSet minus the owned range is
sorted by global index; ``ghost_targets`` = [(owner rank, #ghosts)] sorted by rank;
``import_targets`` = [(requesting rank, #indices)] sorted by rank;
``import_indices`` = half-open *local* index ranges, grouped per requesting rank, one
range per interval of the requester's (sorted) ghost set.  Output format pinned by
``tests/mpi/parallel_partitioner_03.cc:57-80`` and its ``.mpirun=4.output``.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import numpy as np


def _intervals(sorted_idx):
    """Compress a sorted unique index array into [begin, end) intervals."""
    if len(sorted_idx) == 0:
        return []
    breaks = np.nonzero(np.diff(sorted_idx) != 1)[0]
    starts = np.concatenate(([0], breaks + 1))
    ends = np.concatenate((breaks, [len(sorted_idx) - 1]))
    return [(int(sorted_idx[s]), int(sorted_idx[e]) + 1) for s, e in zip(starts, ends)]


class PartitionerOracle:
    def __init__(self, owned_ranges, ghost_sets):
        """owned_ranges: [(begin, end)] per rank (contiguous, ascending with rank);
        ghost_sets: per rank an iterable of global indices (owned ones are dropped)."""
        self.n_ranks = len(owned_ranges)
        self.owned = [tuple(map(int, r)) for r in owned_ranges]
        starts = np.array([r[0] for r in self.owned])
        self.ghosts = []
        for rank, g in enumerate(ghost_sets):
            g = np.unique(np.asarray(list(g), dtype=np.int64))
            b, e = self.owned[rank]
            self.ghosts.append(g[(g < b) | (g >= e)])
        self.ghost_owner = [np.searchsorted(starts, g, side="right") - 1
                            for g in self.ghosts]
        self.ghost_targets, self.import_targets, self.import_indices = [], [], []
        for rank in range(self.n_ranks):
            owners = self.ghost_owner[rank]
            self.ghost_targets.append(
                [(int(o), int((owners == o).sum())) for o in np.unique(owners)])
        for rank in range(self.n_ranks):
            b, _ = self.owned[rank]
            it, ii = [], []
            for other in range(self.n_ranks):
                if other == rank:
                    continue
                req = self.ghosts[other][self.ghost_owner[other] == rank]
                if len(req) == 0:
                    continue
                it.append((other, len(req)))
                ii += [(s - b, e - b) for s, e in _intervals(req)]
            self.import_targets.append(it)
            self.import_indices.append(ii)

    def n_ghosts(self, rank):
        return len(self.ghosts[rank])

    def global_to_local(self, rank, idx):
        """Partitioner::global_to_local (include/deal.II/base/partitioner.h): owned
        -> idx - begin; ghost -> n_owned + position in the sorted ghost set."""
        idx = np.asarray(idx, dtype=np.int64)
        b, e = self.owned[rank]
        owned = (idx >= b) & (idx < e)
        pos = np.searchsorted(self.ghosts[rank], idx)
        return np.where(owned, idx - b, (e - b) + pos)

    def format_like_reference_test(self, rank):
        """The text block tests/mpi/parallel_partitioner_03.cc writes per rank."""
        s = f"**** proc {rank}\n"
        s += "ghost targets: " + "".join(f"[{a}/{b}] " for a, b in self.ghost_targets[rank]) + "\n"
        s += "import targets: " + "".join(f"[{a}/{b}] " for a, b in self.import_targets[rank]) + "\n"
        s += "import indices:\n"
        s += "".join(f"[{a}/{b})\n" for a, b in self.import_indices[rank])
        s += "****\n"
        return s
