"""CPU oracle for the deal.II matrix-free hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  The product path
(``dealii_b200`` + ``libb200mf.so``) never falls back to it.

It is a restatement (numpy + one C file) of the reference algorithm, every
function citing the deal.II file:line it follows (paths relative to the
deal.II source tree, version 9.9.0-pre).

Pinning status (see DESIGN.md "Oracle"): the reference needs its cmake build
system, generated headers (config.h, *.inst) and a 200 MB library, so under
the rules of this build it is treated as *unbuildable here*; the oracle is
pinned against the golden vectors the reference's own tests and tutorials
ship instead (tests/test_oracle_golden.py):
  * examples/step-64/doc/results.dox      DoF counts, CG iteration counts and
                                          solution norms of four refinement cycles
  * tests/lac/precondition_chebyshev_01   Chebyshev preconditioner output
  * tests/mpi/parallel_partitioner_03     ghost/import index algebra (4 ranks)
  * tests/matrix_free_kokkos/compute_diagonal_01 (mpirun=1) diagonal entries
  * tests/matrix_free_kokkos/matrix_free_device_matrix_vector_0x: the
    reference's own strategy -- matrix-free vmult vs an independently
    assembled sparse matrix -- repeated inside the oracle.
"""
