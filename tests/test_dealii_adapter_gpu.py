"""deal.II host code on the engine: examples/step64_dealii.cc (deal.II's own Triangulation / DoFHandler /
AffineConstraints / SolverCG, operator through include/b200mf_dealii.hpp -> libb200mf.so) against the
same solve with deal.II's CPU MatrixFree operator, inside one executable that links the unmodified
reference (oracle/_ref) and the engine.  The binary is built by oracle/ref_drivers/build.sh where the
reference's headers exist; the GPU box runs the prebuilt file."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "bin", "step64_b200")

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not os.path.exists(EXE), reason="oracle/_ref/bin/step64_b200 not built (needs oracle/_ref)")
def test_step64_through_the_adapter_matches_deal_ii():
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "step-64 through libb200mf.so: OK" in r.stdout
    lines = [l for l in r.stdout.splitlines() if l.startswith("cycle")]
    assert len(lines) == 5 and any("hanging nodes" in l for l in lines)
    print(r.stdout)


EXE37 = os.path.join(ROOT, "oracle", "_ref", "bin", "step37_b200")


@pytest.mark.skipif(not os.path.exists(EXE37), reason="oracle/_ref/bin/step37_b200 not built (needs oracle/_ref)")
def test_step37_multigrid_through_the_adapter_matches_deal_ii():
    """examples/step37_dealii.cc: deal.II's SolverCG preconditioned by deal.II's own matrix-free GMG on the host
    against the same SolverCG with the engine's operator and dealii_adapter::PreconditionMG (hierarchy built
    from the same DoFHandler / MGConstrainedDoFs, V-cycle on the device)."""
    r = subprocess.run([EXE37], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "step-37 through libb200mf.so: OK" in r.stdout
    lines = [l for l in r.stdout.splitlines() if l.startswith("Q")]
    assert len(lines) == 5
    print(r.stdout)
