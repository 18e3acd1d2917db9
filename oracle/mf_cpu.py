"""ctypes binding of oracle/mf_cpu.c, the C restatement of the reference's vectorised CPU
MatrixFree Laplace operator (see the header of mf_cpu.c for the file:line map).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): tests/ use it as a second checker,
bench.py's cpu_baseline / --impl reference legs time it on the host cores.
"""
import ctypes as C
import os
import subprocess
import threading
import time

import numpy as np

from .shape import ShapeInfo

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libmfcpu.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            subprocess.check_call(["make", "-C", _HERE])
        lib = C.CDLL(LIB_PATH)
        lib.mfcpu_create.restype = C.c_void_p
        lib.mfcpu_create.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_uint64] + [C.c_void_p] * 6
        lib.mfcpu_destroy.argtypes = [C.c_void_p]
        lib.mfcpu_vmult.argtypes = [C.c_void_p] * 3
        lib.mfcpu_vmult_repeat.argtypes = [C.c_void_p] * 3 + [C.c_int]
        lib.mfcpu_is_cartesian.argtypes = [C.c_void_p]
        _lib = lib
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class MatrixFreeCPU:
    """CPU MatrixFree Laplace operator on the cells (l2g, cell_vertices) of one 'rank'."""

    def __init__(self, dim, degree, l2g, cell_vertices, n_dofs):
        lib = load()
        sh = ShapeInfo(degree)
        l2g = np.ascontiguousarray(l2g, dtype=np.uint32)
        cv = np.ascontiguousarray(cell_vertices, dtype=np.float64)
        sv = np.ascontiguousarray(sh.shape_values)
        sg = np.ascontiguousarray(sh.shape_gradients_collocation)
        qp, qw = np.ascontiguousarray(sh.q_points), np.ascontiguousarray(sh.q_weights)
        self.n_dofs, self.n_cells = int(n_dofs), l2g.shape[0]
        self._h = lib.mfcpu_create(dim, degree, self.n_cells, self.n_dofs, _p(l2g), _p(cv),
                                   _p(sv), _p(sg), _p(qp), _p(qw))
        if not self._h:
            raise ValueError("mfcpu_create failed")
        self.cartesian = bool(lib.mfcpu_is_cartesian(self._h))

    def vmult(self, src, repeat=1):
        src = np.ascontiguousarray(src, dtype=np.float64)
        dst = np.empty(self.n_dofs)
        load().mfcpu_vmult_repeat(self._h, _p(src), _p(dst), repeat)
        return dst

    def __del__(self):
        try:
            if self._h:
                load().mfcpu_destroy(self._h)
                self._h = None
        except Exception:
            pass


def time_vmult_on_cores(ops, srcs, repeat, warmup=1):
    """Run ops[i].vmult(srcs[i]) `repeat` times on one thread per operator, all started
    together (one independent single-rank instance per core, the reference's only way to
    use several cores without MPI/TBB); returns the wall time of the slowest thread."""
    lib = load()
    dsts = [np.empty(o.n_dofs) for o in ops]
    for o, s, d in zip(ops, srcs, dsts):          # first touch + warm-up
        lib.mfcpu_vmult_repeat(o._h, _p(s), _p(d), warmup)
    barrier = threading.Barrier(len(ops) + 1)
    times = [0.0] * len(ops)

    def work(i):
        barrier.wait()
        t0 = time.perf_counter()
        lib.mfcpu_vmult_repeat(ops[i]._h, _p(srcs[i]), _p(dsts[i]), repeat)
        times[i] = time.perf_counter() - t0

    th = [threading.Thread(target=work, args=(i,)) for i in range(len(ops))]
    for t in th:
        t.start()
    barrier.wait()
    for t in th:
        t.join()
    return max(times)
