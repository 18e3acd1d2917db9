"""The C-ABI library loads and exports every symbol include/b200mf.h declares; without a
GPU every compute entry point fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

import dealii_b200
from dealii_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200mf.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200mf_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    lib = L.load()
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in b200mf.h but not exported"
        assert name in L.SYMBOLS, f"{name} has no ctypes prototype"
    assert lib.b200mf_version() == 100


def test_struct_layout_matches_header_sizes(tmp_path):
    """sizeof() of every struct as gcc sees include/b200mf.h vs the ctypes mirrors."""
    import subprocess
    structs = {"b200mf_setup_desc": L.SetupDesc, "b200mf_setup_info": L.SetupInfo,
               "b200mf_operator": L.Operator, "b200mf_solver_desc": L.SolverDesc,
               "b200mf_solver_result": L.SolverResult, "b200mf_mesh_desc": L.MeshDesc,
               "b200mf_mesh_view": L.MeshView, "b200mf_partition_desc": L.PartitionDesc,
               "b200mf_partition_view": L.PartitionView, "b200mf_mg_desc": L.MgDesc,
               "b200mf_mg_level_info": L.MgLevelInfo, "b200mf_bulk_info": L.BulkInfo}
    src = tmp_path / "sizes.c"
    body = "".join(f'printf("{n} %zu\\n", sizeof({n}));' for n in structs)
    src.write_text('#include <stdio.h>\n#include "b200mf.h"\nint main(void){' + body + 'return 0;}\n')
    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).decode().split()
    sizes = dict(zip(out[0::2], map(int, out[1::2])))
    for name, cls in structs.items():
        assert sizes[name] == C.sizeof(cls), name


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    m = dealii_b200.HyperCubeMesh(3, 2, refinements=1)
    with pytest.raises(dealii_b200.B200MFError) as e:
        dealii_b200.MatrixFree().reinit_from_mesh(m)
    assert e.value.code == L.ERR_CUDA
    # the raw ABI call reports the same
    h = C.c_void_p()
    rc = L.load().b200mf_setup_create_from_mesh(m._h, L.F64, C.byref(h))
    assert rc == L.ERR_CUDA and b"no CPU fallback" in L.load().b200mf_last_error()


def _build_cxx_example(tmp_path, name="step64_like"):
    import subprocess
    exe = tmp_path / name
    subprocess.check_call(["g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-I/usr/local/cuda/include",
                           os.path.join(ROOT, "examples", name + ".cc"), "-L", os.path.join(ROOT, "dealii_b200"),
                           "-lb200mf", "-L/usr/local/cuda/lib64", "-lcudart", "-o", str(exe)])
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "dealii_b200") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    return subprocess.run([str(exe)], env=env, capture_output=True, text=True)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_cxx_shim_compiles_and_fails_loudly_without_gpu(tmp_path):
    """include/b200mf_portable.hpp (the Portable::MatrixFree-shaped C++ shim) builds against
    the C ABI with plain g++; without a device the program dies on b200::Exception."""
    r = _build_cxx_example(tmp_path)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_cxx_shim_step64_like_runs(tmp_path):
    r = _build_cxx_example(tmp_path)
    assert r.returncode == 0, r.stderr
    assert "Solved in" in r.stdout and "15625 DoFs" in r.stdout


@pytest.mark.gpu
def test_cxx_multigrid_over_the_c_abi(tmp_path):
    """examples/step37_like.cc: step-37's solver (CG + V-cycle, FP32 levels) from a C++ host that only includes
    b200mf.h -- mesh levels, setups, b200mf_mg_create, b200mf_mg_cg_solve."""
    r = _build_cxx_example(tmp_path, "step37_like")
    assert r.returncode == 0, r.stdout + r.stderr
    assert "35937 DoFs: CG + multigrid converged in" in r.stdout and "level 4" in r.stdout


def test_quadrature_point_count_is_validated_before_anything_else():
    """AssertThrow(n_q_points_1d >= fe_degree + 1) of portable_matrix_free.templates.h:1243: fewer points are an
    invalid argument (also without a device), more are accepted up to 12."""
    lib = L.load()
    l2g = np.zeros(27, dtype=np.uint32)
    for nq, expect_invalid in ((2, True), (13, True)):
        d = L.SetupDesc()
        d.dim, d.degree, d.n_q_points_1d, d.number = 3, 2, nq, L.F64
        d.n_cells, d.n_owned_dofs = 1, 27
        d.local_to_global = l2g.ctypes.data_as(C.c_void_p)
        h = C.c_void_p()
        rc = lib.b200mf_setup_create(C.byref(d), C.byref(h))
        assert rc == L.ERR_INVALID and b"n_q_points_1d" in lib.b200mf_last_error()
