"""examples/dist_cg.cc: a C++ host on several GPUs through the C ABI alone (communicator, partitioner,
distributed vmult, distributed CG) -- no Python on the data path or the control path.  The global problem
(one cube cut into 2^k Morton chunks) does not depend on the rank count, so the energy src^T A src, the CG
iteration count and |x| must agree between 1 rank and N ranks."""
import os
import re
import subprocess
import tempfile

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def build(tmp):
    exe = os.path.join(tmp, "dist_cg")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), "-I/usr/local/cuda/include",
                           os.path.join(ROOT, "examples", "dist_cg.cc"), "-L", os.path.join(ROOT, "dealii_b200"),
                           "-lb200mf", "-L/usr/local/cuda/lib64", "-lcudart", "-o", exe])
    return exe


def run(exe, world, tmp):
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "dealii_b200") + ":" + os.environ.get("LD_LIBRARY_PATH", ""),
               WORLD_SIZE=str(world), B200MF_ID_FILE=os.path.join(tmp, f"id{world}"), B200MF_REFINEMENTS="3")
    # NCCL comes from the torch wheel in this image (a deal.II host would link the system libnccl)
    import nvidia.nccl
    env["LD_LIBRARY_PATH"] = os.path.join(list(nvidia.nccl.__path__)[0], "lib") + ":" + env["LD_LIBRARY_PATH"]
    procs = [subprocess.Popen([exe], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for r in range(world)]
    outs = [p.communicate(timeout=300) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, o + e
    line = [l for l in outs[0][0].splitlines() if l.startswith("dist_cg:")][0]
    m = re.search(r"(\d+) global dofs, 1\^T A 1 = (\S+), CG iterations (\d+), \|x\| = (\S+),", line)
    line = [l for l in outs[0][0].splitlines() if l.startswith("dist_gmg:")][0]
    g = re.search(r"CG \+ multigrid iterations (\d+), \|x\| = (\S+),", line)
    return int(m.group(1)), float(m.group(2)), int(m.group(3)), float(m.group(4)), int(g.group(1)), float(g.group(2))


def test_cxx_multi_rank_host():
    with tempfile.TemporaryDirectory() as tmp:
        exe = build(tmp)
        ref = run(exe, 1, tmp)
        assert ref[0] == (4 * 8 + 1) ** 3 and ref[2] > 10
        # the multigrid-preconditioned solve of the same system: few iterations, the same solution
        assert ref[4] <= 9 and abs(ref[5] - ref[3]) < 1e-6 * ref[3]
        n_gpus = torch.cuda.device_count()
        for world in (2, 4, 8):
            if n_gpus < world:
                continue
            got = run(exe, world, tmp)
            assert got[0] == ref[0]
            assert abs(got[1] - ref[1]) < 1e-10 * abs(ref[1])
            assert abs(got[2] - ref[2]) <= 1
            assert abs(got[3] - ref[3]) < 1e-7 * ref[3]
            assert abs(got[4] - ref[4]) <= 1 and abs(got[5] - ref[5]) < 1e-7 * ref[5]
