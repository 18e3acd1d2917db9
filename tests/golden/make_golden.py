"""Copies the golden OUTPUT files (not sources) the oracle is pinned against out of the
read-only reference tree into tests/golden/.  Run once in the build container:
    python tests/golden/make_golden.py /root/reference
The GPU box has no /root/reference; tests only read the committed copies."""
import re, shutil, sys, pathlib
ref = pathlib.Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
out = pathlib.Path(__file__).parent
files = {
    "tests/mpi/parallel_partitioner_03.mpirun=4.output": "parallel_partitioner_03.mpirun=4.output",
    "tests/lac/precondition_chebyshev_01.with_lapack=true.output": "precondition_chebyshev_01.output",
    "tests/matrix_free_kokkos/compute_diagonal_01.with_p4est=true.mpirun=1.output": "compute_diagonal_01.mpirun=1.output",
    "tests/matrix_free_kokkos/matrix_free_device_matrix_vector_01.output": "matrix_free_device_matrix_vector_01.output",
    "tests/matrix_free_kokkos/matrix_free_device_matrix_vector_02.output": "matrix_free_device_matrix_vector_02.output",
    "tests/matrix_free_kokkos/matrix_free_device_matrix_vector_03.output": "matrix_free_device_matrix_vector_03.output",
    "tests/matrix_free/step-37.with_lapack=true.output": "step-37.output",
    "tests/matrix_free/hanging_node_kernels_01.output": "hanging_node_kernels_01.output",
    "tests/mpi/p4est_2d_dofhandler_01.mpirun=4.with_p4est=true.output": "p4est_2d_dofhandler_01.mpirun=4.output",
    "tests/matrix_free_kokkos/matrix_free_device_initialize_vector.with_mpi=on.with_p4est=on.mpirun=2.output":
        "matrix_free_device_initialize_vector.mpirun=2.output",
    "tests/matrix_free/solver_cg_interleave.with_p4est=true.mpirun=3.output": "solver_cg_interleave.mpirun=3.output",
}
for src, dst in files.items():
    shutil.copyfile(ref / src, out / dst)
# step-64 expected output block of examples/step-64/doc/results.dox:6-30
txt = (ref / "examples/step-64/doc/results.dox").read_text()
block = txt[txt.index("@code") + 6: txt.index("@endcode")]
(out / "step-64.results.txt").write_text(block)
print("golden files written to", out)
