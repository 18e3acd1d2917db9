#!/usr/bin/env python
"""bench.py -- headline benchmark of the matrix-free operator engine (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W              # engine arm
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU reference arm

Workload (BASELINE.json configs[1]): 3D Q4 Laplace vmult, FP64, affine Cartesian hyper_cube
refined globally 7 times (128^3 cells, 513^3 = 135,005,697 DoFs) per GPU; a "step" is one
vmult (dst = 0; cell loop; copy_constrained_values) over the whole mesh.  At N > 1 the domain is
N such cubes (subdivided_hyper_rectangle, one cube per rank, p4est-style partition and DoF
numbering), vectors are [owned | ghosts], and every vmult does update_ghost_values (NCCL
send/recv over NVLink, overlapped with the interior cells) and compress(add): weak scaling.

One JSON line on stdout (rank 0), see the task contract: value = whole-job GDoF/s with the
vectors resident in HBM, e2e = the same through the C-ABI host entry point
(b200mf_vmult_host: pinned host src -> device -> vmult -> host dst inside the timed region),
roofline = algorithmic bytes (16 B/DoF) of the cell-loop kernel over its
CUDA-event duration against MEASURED_PEAKS.json, cpu_baseline = oracle/mf_cpu.c (the C
restatement of the reference's vectorised CPU MatrixFree path) on the host cores.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "vmult_throughput_3d_q4_laplace"
UNIT = "GDoF/s"
BYTES_PER_DOF = {"f64": 16.0, "f32": 8.0}            # read src + write dst
CG_BYTES_PER_DOF = {"f64": 72.0, "f32": 36.0}        # fused Jacobi-CG iteration


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--degree", type=int, default=None)
    ap.add_argument("--refinements", type=int, default=7)
    ap.add_argument("--number", default="f64", choices=["f64", "f32"])
    ap.add_argument("--deformation", type=float, default=0.0)
    ap.add_argument("--workload", default="c2", choices=["c2", "c4"],
                    help="c2: BASELINE configs[1] (3D Q4 Laplace, Cartesian hyper_cube); c4: configs[3] (3D Q3 "
                         "Poisson, one refinement ball per cube = hanging nodes)")
    ap.add_argument("--sweep", action="store_true",
                    help="BASELINE configs[4]: degrees 1-8, Cartesian and deformed (full-Jacobian) meshes, FP64 and "
                         "FP32 vmult with the HBM-roofline fraction per case; one JSON line with a 'sweep' array")
    ap.add_argument("--ball-radius", type=float, default=0.35)
    ap.add_argument("--no-cg", action="store_true")
    ap.add_argument("--no-converged-cg", action="store_true")
    ap.add_argument("--no-chebyshev", action="store_true")
    ap.add_argument("--no-renumbered", action="store_true")
    ap.add_argument("--no-gmg", action="store_true")
    ap.add_argument("--gmg-reference-refinements", type=int, default=5,
                    help="size of the same-problem comparison with the reference's own CPU multigrid (oracle/_ref ref_gmg)")
    ap.add_argument("--gmg-levels", default="f32", choices=["f32", "f64"],
                    help="number type of the multigrid levels under the FP64 CG (step-37 uses float)")
    ap.add_argument("--cg-rel-tol", type=float, default=1e-6)
    ap.add_argument("--cg-max-iterations", type=int, default=4000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ghosts", default="relevant", choices=["relevant", "touched"],
                    help="ghost set: Portable::MatrixFree's locally relevant dofs or the tight touched set")
    ap.add_argument("--cpu-refinements", type=int, default=6,
                    help="per-core sub-cube of the CPU baseline sample")
    a = ap.parse_args()
    if a.degree is None:
        a.degree = 3 if a.workload == "c4" else 4
    if a.workload == "c4":
        a.ghosts = "touched"
    return a


def kernel_name(args, n_bricks=0):
    """The kernel the dispatch policy of csrc/api.cu / cell_inst.cu picks for this workload."""
    n, f64, general = args.degree + 1, args.number == "f64", args.deformation != 0.0
    if n_bricks and not general and os.environ.get("B200MF_KERNEL") in (None, "brick"):
        b = 16 if args.degree == 1 else 8 if args.degree == 2 else 4 if args.degree <= 4 else 2
        return f"brick_cartesian_kernel<{args.degree},{b},{'double' if f64 else 'float'}>"
    plane = (n <= 4 if f64 else (n <= 3 or n == 5)) if general else n <= 5
    if os.environ.get("B200MF_KERNEL") == "v1":
        plane = False
    if os.environ.get("B200MF_KERNEL") == "plane" and n <= 6:
        plane = True
    t = "double" if f64 else "float"
    kind = "GENERAL" if general else "CARTESIAN"
    return (f"cell_loop_plane_kernel<{n},{t},{kind}>" if plane else f"cell_loop_kernel<3,{n},{t},{kind}>")


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def mark(self):
        """Only samples taken after this point are reported."""
        self.t_mark = time.perf_counter()

    def n_since_mark(self):
        return sum(1 for t, _ in self.lines if t >= getattr(self, "t_mark", 0.0))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.kill()          # the exact PID we started
        self.proc.wait()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t_ln, ln in self.lines:
            if t_ln < getattr(self, "t_mark", 0.0):
                continue
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- CPU reference
REF_BENCH = os.path.join(ROOT, "oracle", "_ref", "bin", "ref_bench")


def reference_available():
    return os.path.exists(REF_BENCH) and os.path.exists(
        os.path.join(ROOT, "oracle", "_ref", "install", "lib", "libdeal_II.so"))


def cpu_reference_sample(degree, refinements, steps, warmup, deformation=0.0, mode="vmult"):
    """The reference's own vectorised CPU MatrixFree path on every host core.

    With oracle/_ref present (deal.II built from /root/reference by oracle/build_ref.sh): the
    UNMODIFIED library through oracle/_ref/bin/ref_bench (operator of
    tests/performance/timing_matrix_free_kokkos.cc:56-111), one pinned single-rank process per
    core -- the image has no MPI/TBB, so this is how the reference uses several cores; zero
    communication cost, i.e. an upper bound for "MatrixFree with MPI" on the same cores -- all
    released together by a start file; kind = "reference".  Without it: the C restatement
    oracle/mf_cpu.c, kind = "port"."""
    cores = sorted(os.sched_getaffinity(0))
    if reference_available() and deformation == 0.0:
        import tempfile
        with tempfile.TemporaryDirectory() as tmp:
            start = os.path.join(tmp, "go")
            procs = [subprocess.Popen(["taskset", "-c", str(c), REF_BENCH, str(degree), str(refinements),
                                       str(steps), str(max(warmup, 1)), mode, start],
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
                     for c in cores]
            # every instance finishes its setup + warm-up, then polls for the start file
            # the instances poll for the start file once their setup + warm-up is done: give the
            # slowest setup a generous, size-scaled head start
            setup_wait = {1: 0.5, 2: 0.5, 3: 1.0, 4: 2.0, 5: 6.0, 6: 25.0, 7: 240.0}.get(refinements, 30.0) * (2.0 if mode == "cg" else 1.0)
            time.sleep(setup_wait)
            open(start, "w").close()
            outs = [p.communicate(timeout=900) for p in procs]
        res = []
        for p, (o, e) in zip(procs, outs):
            if p.returncode != 0:
                raise RuntimeError(f"ref_bench failed: {e[-400:]}")
            res.append(json.loads(o.strip().splitlines()[-1]))
        t = max(r["seconds"] for r in res)
        n_dofs = res[0]["n_dofs"]
        total = len(cores) * n_dofs
        units = steps if mode == "vmult" else res[0]["iterations"]
        return {"value": total * units / t / 1e9, "unit": UNIT if mode == "vmult" else "GDoF-iterations/s",
                "cores": len(cores), "kind": "reference",
                "sample": (f"deal.II 9.9.0-pre built from /root/reference (oracle/_ref), CPU MatrixFree "
                           f"{'vmult' if mode == 'vmult' else 'SolverCG + Jacobi'} "
                           f"(operator of tests/performance/timing_matrix_free_kokkos.cc:56-111, AVX-512 "
                           f"VectorizedArray<double,{res[0]['vectorization_lanes']}>): {len(cores)} pinned "
                           f"single-rank processes (no MPI/TBB in the image), each 3D Q{degree} hyper_cube "
                           f"refine_global({refinements}) = {n_dofs} DoFs, {units} "
                           f"{'vmults' if mode == 'vmult' else 'iterations'} after {max(warmup, 1)} warm-up, common "
                           f"start, slowest instance's time; the full-size mesh (refine_global(7)) does not fit "
                           f"the reference's host setup {len(cores)} times"),
                "seconds": t, "dofs_per_step": total, "sample_refinements": refinements,
                "per_core_mdofs": n_dofs * units / t / 1e6}
    import numpy as np
    import dealii_b200                       # host-side mesh generator only (no GPU work)
    from oracle.mf_cpu import MatrixFreeCPU, time_vmult_on_cores
    ncores = len(cores)
    mesh = dealii_b200.HyperCubeMesh(3, degree, refinements=refinements,
                                     deformation_amplitude=deformation)
    ops = [MatrixFreeCPU(3, degree, mesh.l2g, mesh.cell_vertices, mesh.n_dofs) for _ in range(ncores)]
    srcs = [np.random.default_rng(42 + i).random(mesh.n_dofs) for i in range(ncores)]
    t = time_vmult_on_cores(ops, srcs, steps, warmup=max(warmup, 1))
    total = ncores * mesh.n_dofs
    return {"value": total * steps / t / 1e9, "unit": UNIT, "cores": ncores, "kind": "port",
            "sample": (f"{ncores} concurrent single-rank instances (1 per core) of oracle/mf_cpu.c (C "
                       f"restatement of matrix_free/evaluation_kernels.h:1835-1916; oracle/_ref absent), each "
                       f"3D Q{degree} hyper_cube refine_global({refinements}) = {mesh.n_dofs} DoFs, "
                       f"{steps} vmults after {max(warmup, 1)} warm-up"),
            "seconds": t, "dofs_per_step": total, "sample_refinements": refinements}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 20))
    warm = max(min(args.warmup, 5), 1)
    r = cpu_reference_sample(args.degree, args.cpu_refinements, steps, warm, args.deformation)
    cfg = workload_config(args, r["dofs_per_step"], "host cores only")
    cfg["sample"] = (f"bounded sample: refine_global({r['sample_refinements']}) per core instead of "
                     f"refine_global({args.refinements}) per GPU (per-DoF CPU throughput out of cache does not "
                     f"depend on the size)")
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT,
            "n_gpus": args.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": r["seconds"] / steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    return line


def workload_config(args, n_dofs_total, parallelism):
    if getattr(args, "workload", "c2") == "c4":
        what = (f"3D Q{args.degree} Poisson vmult, {args.number}, hyper_cube refine_global({args.refinements}) + "
                f"cells within {args.ball_radius} of the cube centre refined once (hanging nodes) per GPU")
    else:
        what = (f"3D Q{args.degree} Laplace vmult, {args.number}, "
                f"{'deformed (general cells)' if args.deformation else 'affine Cartesian'} "
                f"hyper_cube refine_global({args.refinements}) per GPU")
    return {"workload": what,
            "degree": args.degree, "dim": 3, "n_dofs_total": int(n_dofs_total),
            "refinements": args.refinements, "parallelism": parallelism,
            "l2": "inputs larger than L2 (vectors >= 1 GB each; no flush needed)"}


# ----------------------------------------------------------------------------- engine arm
def make_mesh(args, world, rank, coarse, dirichlet):
    from dealii_b200.distributed import AdaptiveHyperCubeMesh, PartitionedHyperCubeMesh
    if args.workload == "c4":
        return AdaptiveHyperCubeMesh(3, args.degree, args.refinements, world, rank, coarse=coarse,
                                     ball_radius=args.ball_radius, dirichlet_boundary=dirichlet)
    return PartitionedHyperCubeMesh(3, args.degree, args.refinements, world, rank, coarse=coarse,
                                    deformation_amplitude=args.deformation, dirichlet_boundary=dirichlet,
                                    ghost_mode=args.ghosts)


def run_engine(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import dealii_b200
    from dealii_b200 import _lib as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL's version / debug banner goes to stdout by default: keep stdout for the JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        # torch.distributed is plumbing here (barrier, max over ranks, broadcast of the NCCL id of the
        # engine's own communicator); the data path is csrc/comm.cu
        dist.init_process_group("nccl", device_id=dev)
    lib = L.load()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_ranks(x, op="max"):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
        return float(t.item())

    max_over_ranks = reduce_ranks

    # ---- setup: one cube of the workload per rank, partitioned like p4est
    from dealii_b200.distributed import DistributedMatrixFree, solve_cg
    coarse = {1: (1, 1, 1), 2: (1, 1, 2), 4: (1, 2, 2), 8: (2, 2, 2)}.get(world)
    assert coarse is not None, "bench.py supports 1, 2, 4 or 8 GPUs"
    mesh = make_mesh(args, world, rank, coarse, dirichlet=False)
    dmf = DistributedMatrixFree(mesh, args.number, dev)
    mf = dmf.mf
    op = dealii_b200.LaplaceOperator(mf)
    n_dofs = mesh.n_owned
    n_total = mesh.n_global_dofs
    tdt = mf.torch_dtype
    gen = torch.Generator(device=dev).manual_seed(42 + rank)
    src = dmf.initialize_dof_vector()
    src[:n_dofs] = torch.rand(n_dofs, dtype=tdt, device=dev, generator=gen)
    cons = None
    if args.workload == "c4":      # Portable::MatrixFree semantics: hanging-node entries of src are not read
        cons = torch.from_numpy(mesh.constrained_dofs.astype(np.int64)).to(dev)
        src[cons] = 0.0
    dst = dmf.initialize_dof_vector()

    def step():
        dmf.vmult(op.op, dst, src)

    # ---- parity self-check inside the bench (every rank count the driver runs proves the NCCL path):
    # (1) the Laplacian annihilates constants -- needs every ghost value, every hanging-node
    # interpolation and its transpose, and the compress of the interface rows; (2) symmetry with
    # partition-independent global sums
    one = dmf.initialize_dof_vector()
    one[:n_dofs] = 1.0
    if cons is not None:
        one[cons] = 0.0
    dmf.vmult(op.op, dst, one)
    if cons is not None:
        dst[cons] = 0.0
    inv_diag0 = dmf.compute_diagonal(op.op)
    defect = reduce_ranks((dst[:n_dofs].abs() * inv_diag0[:n_dofs].abs()).max())
    v2 = dmf.initialize_dof_vector()
    v2[:n_dofs] = torch.rand(n_dofs, dtype=tdt, device=dev, generator=gen)
    if cons is not None:
        v2[cons] = 0.0
    w1, w2 = dmf.initialize_dof_vector(), dmf.initialize_dof_vector()
    dmf.vmult(op.op, w1, src)
    dmf.vmult(op.op, w2, v2)
    s12 = reduce_ranks(torch.dot(v2[:n_dofs].double(), w1[:n_dofs].double()), "sum")
    s21 = reduce_ranks(torch.dot(src[:n_dofs].double(), w2[:n_dofs].double()), "sum")
    tol_c = 1e-10 if args.number == "f64" else 1e-3
    parity = {"constants_defect": defect, "symmetry_rel": abs(s12 - s21) / max(abs(s12), 1e-300),
              "ok": bool(defect < tol_c and abs(s12 - s21) <= tol_c * abs(s12)),
              "what": "max |A 1| / diag over all rows (0 for the exact operator: ghosts, hanging-node "
                      "interpolation and compress all enter) and |v.Au - u.Av| / |v.Au| with global sums"}
    del one, v2, w1, w2, inv_diag0

    # ---- value: K vmults, vectors resident in HBM
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if rank == 0:
        sampler.mark()
    lc0 = lib.b200mf_kernel_launch_count()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    lc1 = lib.b200mf_kernel_launch_count()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    t_extra = time.perf_counter()
    while True:       # untimed continuation of the same loop for the clock samples (all ranks agree)
        enough = (rank != 0) or sampler.n_since_mark() >= 6 or time.perf_counter() - t_extra > 2.0
        if max_over_ranks(0.0 if enough else 1.0) == 0.0:
            break
        for _ in range(10):
            step()
        torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "timed region + identical untimed continuation until >= 6 samples"
    ms_per_step = ms / args.steps
    value = n_total / (ms_per_step * 1e-3) / 1e9

    # ---- roofline: the cell-loop kernel(s) alone (CUDA events on the launching stream).  The library
    # says which brick path its setup-time measurement chose: the bulk brick kernel IS the vmult (one
    # launch, no memset); the index-map brick kernel is the launch that follows vmult's memset.
    bulk = mf.bulk_info()
    reps = max(args.steps, 10)
    if bulk["enabled"] and world == 1:
        run_kernel = lambda: mf.vmult(op.op, dst, src)
    else:
        run_kernel = lambda: mf.vmult_range(op.op, dst, src, 0, mesh.n_cells)
    for _ in range(3):
        run_kernel()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        run_kernel()
    e1.record()
    torch.cuda.synchronize()
    ms_kernel = e0.elapsed_time(e1) / reps
    peak, peak_src = measured_peak()
    bpd = BYTES_PER_DOF[args.number]
    if args.deformation:
        p = args.degree
        bpd += 6 * (bpd / 2) * ((p + 1) / p) ** 3        # merged symmetric metric per q-point
    achieved = bpd * n_dofs / (ms_kernel * 1e-3) / 1e9
    traffic = ncu_traffic()
    cells_in_bricks = int(mf.info.n_bricks * mf.info.cells_per_brick)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak,
                "traffic": (traffic.get("dram_bytes_per_launch") if traffic and args.workload == "c2"
                            and args.degree == 4 and args.number == "f64" else None),
                "traffic_source": traffic.get("source") if traffic else None,
                "kernel": (kernel_name(args, int(mf.info.n_bricks)).replace("brick_cartesian_kernel", "bulk_brick_kernel")
                           if bulk["enabled"] else kernel_name(args, int(mf.info.n_bricks))),
                "brick_path": {"chosen": "bulk tables + first-toucher-stores" if bulk["enabled"] else
                               "index maps + memset + atomics", "chosen_by": "setup-time measurement",
                               "ms_index_map": bulk["tuned_ms_index_map"], "ms_bulk": bulk["tuned_ms_bulk"],
                               "patterns": bulk["n_patterns"]},
                "cells_in_bricks": cells_in_bricks, "cells": int(mesh.n_cells),
                "kernel_ms": ms_kernel, "algorithmic_bytes_per_dof": bpd, "peak_source": peak_src,
                "vmult_frac": bpd * n_dofs / (ms_per_step * 1e-3) / 1e9 / peak,
                "note": ("kernel_ms = the cell loop over all local cells (all its launches); vmult_frac = the "
                         "same bytes over the whole vmult (memset, ghost exchange, copy_constrained_values "
                         "included).  FP64 sum factorisation is co-bound by the FP64 pipe (37.1 TFLOP/s "
                         "measured, tools/fp64_peak.cu); see DESIGN.md")}
    if args.workload == "c4":
        roofline["hanging_node_cells"] = int(mesh.n_masked_cells)
        roofline["hanging_node_dofs"] = int(mesh.n_hanging_dofs)

    # ---- e2e: host buffers in, host buffers out, copies inside the timed region, through the C ABI:
    # b200mf_vmult_host_batch (N = 1) / b200mf_dist_vmult_host_batch (N > 1): per vector H2D of the
    # owned part, vmult (with its ghost exchange), D2H, the copies of neighbouring vectors overlapped
    nbytes = n_dofs * (8 if args.number == "f64" else 4)
    h_src = [torch.empty(n_dofs, dtype=tdt).pin_memory() for _ in range(2)]
    h_dst = [torch.empty(n_dofs, dtype=tdt).pin_memory() for _ in range(2)]
    h_src[0].copy_(src[:n_dofs])
    h_src[1].copy_(h_src[0])
    nb = 16 if world == 1 else 8
    srcs = [h_src[k % 2].numpy() for k in range(nb)]
    dsts = [h_dst[k % 2].numpy() for k in range(nb)]
    single = None
    if world == 1:
        op.vmult_host(dsts[0], srcs[0])
        t0 = time.perf_counter()
        for _ in range(3):
            op.vmult_host(dsts[0], srcs[0])
        single = {"value": n_total * 3 / (time.perf_counter() - t0) / 1e9, "api": "b200mf_vmult_host", "steps": 3}
    dmf.vmult_host_batch(op.op, dsts[:2], srcs[:2])
    barrier()
    t0 = time.perf_counter()
    dmf.vmult_host_batch(op.op, dsts, srcs)
    torch.cuda.synchronize()
    t_b = max_over_ranks(time.perf_counter() - t0)
    e2e = {"value": n_total * nb / t_b / 1e9, "unit": UNIT, "h2d_bytes_per_step": nbytes * world,
           "d2h_bytes_per_step": nbytes * world, "steps": nb,
           "api": ("b200mf_vmult_host_batch" if world == 1 else "b200mf_dist_vmult_host_batch") +
                  " (include/b200mf.h): per-vector H2D, vmult, D2H, pipelined over 2 slots"}
    if single:
        e2e["single_call"] = single
    step()
    torch.cuda.synchronize()
    # atomics make the summation order (hence the last bits) run-dependent: compare to 1e-12
    check = float((h_dst[1][:100000].to(dev) - dst[:100000]).abs().max() / dst[:100000].abs().max())
    del h_src, h_dst, srcs, dsts

    # ---- CG + Jacobi (the second half of the metric): DoF-iterations/s
    cg = None
    if not args.no_cg:
        del dst, src, dmf, mf, op
        torch.cuda.empty_cache()
        cg = run_cg(args, dev, rank, world, coarse, barrier, max_over_ranks, peak)

    # ---- the opt-in numbering DoFRenumbering::lexicographic (dofs/dof_renumbering.h:1327-1342): every
    # brick's lattice is then an affine image of the numbering, the brick kernel computes its indices and
    # streams no index map (strided bricks).  Reported beside the default numbering, never instead of it.
    renumbered = None
    if world == 1 and args.workload == "c2" and not args.deformation and not args.no_renumbered:
        renumbered = run_renumbered(args, dev, peak)

    launches = lc1 - lc0
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_sample(args.degree, args.cpu_refinements, 8, 1, args.deformation)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        cfg = workload_config(args, n_total,
                              "1 GPU" if world == 1 else
                              f"{world} ranks (1 per GPU), one cube per rank, {mesh.n_ghost} ghost dofs per rank, "
                              f"NCCL p2p ghost exchange behind the C ABI (b200mf_dist_vmult)")
        line = {"metric": METRIC if args.workload == "c2" else "vmult_throughput_3d_q3_poisson_hanging_nodes",
                "value": value, "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": args.number, "data": "synthetic", "config": cfg,
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
                "roofline": roofline, "cpu_baseline": cpu, "cg": cg, "renumbered": renumbered,
                "parity_check": parity["ok"], "parity": parity,
                "e2e_matches_device": bool(check < 1e-12 if args.number == "f64" else check < 1e-5)}
    result = line if rank == 0 else None
    if world > 1:
        dist.destroy_process_group()
    return result


def run_renumbered(args, dev, peak):
    """The same workload on the DoFHandler renumbered with DoFRenumbering::lexicographic: vmult, the
    cell-loop kernel alone, and CG + Jacobi (fixed iteration count)."""
    import torch
    import dealii_b200
    mesh = dealii_b200.HyperCubeMesh(3, args.degree, refinements=args.refinements, numbering="lexicographic")
    mf = dealii_b200.MatrixFree(args.number, dev).reinit_from_mesh(mesh)
    op = dealii_b200.LaplaceOperator(mf)
    info = mf.bulk_info()
    n = mf.n_owned
    x = torch.rand(n, dtype=mf.torch_dtype, device=dev)
    y = mf.initialize_dof_vector()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(f, reps):
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            f()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    reps = max(args.steps, 10)
    ms = timed(lambda: op.vmult(y, x), reps)
    ms_k = timed(lambda: mf.vmult_range(op.op, y, x, 0, mesh.n_cells), reps)
    bpd = BYTES_PER_DOF[args.number]
    out = {"numbering": "DoFRenumbering::lexicographic (opt-in; bit-identical to the reference's function, tests/test_mesh.py)",
           "strided_bricks": bool(info["strided"]), "brick_path": info["path"],
           "value": n / ms / 1e6, "unit": UNIT, "ms_per_step": ms, "kernel_ms": ms_k,
           "roofline_frac_kernel": bpd * n / (ms_k * 1e-3) / 1e9 / peak,
           "roofline_frac_vmult": bpd * n / (ms * 1e-3) / 1e9 / peak,
           "traffic": "profiles/r02_brick_strided_q4_f64_ncu.txt: 1.92 GB read + 1.06 GB written per launch"}
    del x, y
    if not args.no_cg:
        mesh_d = dealii_b200.HyperCubeMesh(3, args.degree, refinements=args.refinements, numbering="lexicographic",
                                           dirichlet_boundary=True)
        mfd = dealii_b200.MatrixFree(args.number, dev).reinit_from_mesh(mesh_d)
        A = dealii_b200.LaplaceOperator(mfd)
        inv = A.compute_diagonal()
        b = torch.ones(mfd.n_owned, dtype=mfd.torch_dtype, device=dev)
        mfd.set_constrained_values(0.0, b)
        iters = max(min(args.steps, 30), 10)
        best = None
        for rep in range(2):
            xs = mfd.initialize_dof_vector()
            control = dealii_b200.SolverControl(iters, 1e-300)
            e0.record()
            try:
                dealii_b200.SolverCG(control).solve(A, xs, b, inv)
            except dealii_b200.B200MFError:
                pass
            e1.record()
            torch.cuda.synchronize()
            best = e0.elapsed_time(e1)
        its = control.last_step()
        out["cg"] = {"value": mfd.n_owned * its / (best * 1e-3) / 1e9, "unit": "GDoF-iterations/s", "iterations": its,
                     "ms_per_iteration": best / max(its, 1),
                     "roofline_frac": CG_BYTES_PER_DOF[args.number] * mfd.n_owned * its / (best * 1e-3) / 1e9 / peak}
    return out


def run_cg(args, dev, rank, world, coarse, barrier, max_over_ranks, peak):
    """SolverCG + Jacobi on the same (partitioned) mesh with zero Dirichlet boundary, rhs = 1:
    (a) a fixed number of iterations (stopped by max_iterations, like IterationNumberControl) for the
    throughput, (b) a converged solve (time to solution), (c) on one GPU a CG + Chebyshev line."""
    import torch
    import dealii_b200
    from dealii_b200.distributed import DistributedMatrixFree, solve_cg
    mesh = make_mesh(args, world, rank, coarse, dirichlet=True)
    dmf = DistributedMatrixFree(mesh, args.number, dev)
    A = dealii_b200.LaplaceOperator(dmf.mf)
    inv_diag = dmf.compute_diagonal(A.op)
    b = dmf.initialize_dof_vector()
    b[:mesh.n_owned] = 1.0
    dmf.mf.set_constrained_values(0.0, b)
    iters = max(min(args.steps, 30), 10)
    best, its, res = None, 0, 0.0
    for rep in range(2):                      # first solve = warm-up
        x = dmf.initialize_dof_vector()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        its, res, _ = solve_cg(dmf, A.op, x, b, inv_diag, 1e-300, iters, check_every=10)
        e1.record()
        torch.cuda.synchronize()
        best = max_over_ranks(e0.elapsed_time(e1))
    n_total = mesh.n_global_dofs
    val = n_total * its / (best * 1e-3) / 1e9
    bpd = CG_BYTES_PER_DOF[args.number]
    out = {"metric": "cg_jacobi_throughput", "value": val, "unit": "GDoF-iterations/s",
           "iterations": its, "ms_per_iteration": best / max(its, 1), "residual": res,
           "roofline_frac": bpd * n_total / world * its / (best * 1e-3) / 1e9 / peak,
           "algorithmic_bytes_per_dof_iteration": bpd,
           "api": "b200mf_dist_cg_solve (include/b200mf.h), residual read back every 10 iterations"}
    if not args.no_converged_cg:
        # time to solution: ||r|| <= rel_tol ||b||
        bn = float(torch.dot(b[:mesh.n_owned].double(), b[:mesh.n_owned].double()))
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([bn], dtype=torch.float64, device=dev)
            dist.all_reduce(t)
            bn = float(t)
        rel = args.cg_rel_tol
        x = dmf.initialize_dof_vector()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        cits, cres, ok = solve_cg(dmf, A.op, x, b, inv_diag, rel * bn ** 0.5, args.cg_max_iterations, check_every=10)
        e1.record()
        torch.cuda.synchronize()
        tsol = max_over_ranks(e0.elapsed_time(e1))
        out["converged_solve"] = {"relative_tolerance": rel, "iterations": cits, "converged": bool(ok),
                                  "residual": cres, "seconds": tsol * 1e-3,
                                  "ms_per_iteration": tsol / max(cits, 1),
                                  "value": n_total * cits / (tsol * 1e-3) / 1e9, "unit": "GDoF-iterations/s"}
    if world == 1 and not args.no_chebyshev:
        # CG with PreconditionChebyshev(degree 3) over Jacobi: 3 operator applications per iteration
        cheb = dealii_b200.PreconditionChebyshev(degree=3, smoothing_range=20.0, eig_cg_n_iterations=10,
                                                 preconditioner=dealii_b200.DiagonalMatrix(inv_diag))
        best = None
        for rep in range(2):
            x = dmf.mf.initialize_dof_vector()
            control = dealii_b200.SolverControl(8, 1e-300)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            try:
                dealii_b200.SolverCG(control).solve(A, x, b, cheb)
            except dealii_b200.B200MFError:
                pass                           # max_iterations reached: that is the stopping rule here
            e1.record()
            torch.cuda.synchronize()
            best = e0.elapsed_time(e1)
        r = control.last_step()
        out["chebyshev"] = {"metric": "cg_chebyshev3_throughput", "iterations": r,
                            "ms_total_incl_eigenvalue_estimate": best,
                            "note": "SolverCG + PreconditionChebyshev(degree 3, 10 CG iterations for the "
                                    "eigenvalue estimate) through b200mf_cg_solve"}
    if args.workload == "c2" and args.deformation == 0.0 and not args.no_gmg:
        out["gmg"] = run_gmg(args, dev, rank, world, coarse, barrier, max_over_ranks, dmf.comm)
    return out


def run_gmg(args, dev, rank, world, coarse, barrier, max_over_ranks, comm):
    """Time to solution with the geometric multigrid of step-37 on the same problem
    as the converged Jacobi solve: FP64 CG preconditioned by one V-cycle (Chebyshev(5) smoothers on
    float levels, matrix-free transfer, Chebyshev coarse solver), through b200mf_mg_create +
    b200mf_mg_dist_cg_solve.  On N ranks every rank holds its cube on every level (weak scaling): the
    iteration count does not grow with N, unlike Jacobi-CG's."""
    import time
    import torch
    import dealii_b200
    from dealii_b200.distributed import (DistributedGeometricMultigrid, DistributedMatrixFree,
                                         PartitionedHyperCubeMesh)
    t0 = time.time()
    mg = DistributedGeometricMultigrid(3, args.degree, args.refinements, world, rank, coarse=coarse,
                                       number=args.gmg_levels, device=dev, comm=comm)
    torch.cuda.synchronize()
    t_setup = time.time() - t0
    mesh = PartitionedHyperCubeMesh(3, args.degree, args.refinements, world, rank, coarse=coarse,
                                    dirichlet_boundary=True, mark_constrained_l2g=True, ghost_mode="touched")
    system = DistributedMatrixFree(mesh, "f64", dev, comm=comm)
    A = dealii_b200.LaplaceOperator(system.mf)
    n = mesh.n_owned
    b = system.initialize_dof_vector()
    b[:n] = 1.0
    system.mf.set_constrained_values(0.0, b)
    scal = torch.tensor([float(torch.dot(b[:n], b[:n]))], device=dev, dtype=torch.float64)
    comm.allreduce_sum(scal)
    bnorm = float(scal) ** 0.5
    tol = args.cg_rel_tol * bnorm
    best, its, ok = None, 0, False
    for rep in range(2):                      # first solve = warm-up
        x = system.initialize_dof_vector()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        its, res, ok = mg.solve(system, A.op, x, b, tol, 100)
        e1.record()
        torch.cuda.synchronize()
        best = max_over_ranks(e0.elapsed_time(e1))
    # the answer is checked against the operator: ||b - A x|| / ||b||
    r = system.initialize_dof_vector()
    system.vmult(A.op, r, x)
    rr = torch.tensor([float(torch.dot(b[:n] - r[:n], b[:n] - r[:n]))], device=dev, dtype=torch.float64)
    comm.allreduce_sum(rr)
    same_size = None
    if world == 1 and not args.no_cpu_baseline:
        same_size = gmg_reference_comparison(args, dev)
    return {"metric": "cg_gmg_time_to_solution", "levels": mg.n_levels(), "level_number": args.gmg_levels,
            "same_problem_on_the_reference": same_size,
            "n_dofs": mesh.n_global_dofs, "relative_tolerance": args.cg_rel_tol, "iterations": its,
            "converged": bool(ok), "seconds": best * 1e-3, "true_relative_residual": float(rr) ** 0.5 / bnorm,
            "value": mesh.n_global_dofs / (best * 1e-3) / 1e9, "unit": "GDoF/s (unknowns solved per second)",
            "setup_seconds_host_incl_level_meshes": t_setup,
            "smoother": "Chebyshev degree 5, range 15, 10 Lanczos iterations (step-37.cc:965-975)",
            "api": "b200mf_mg_create + b200mf_mg_dist_cg_solve (include/b200mf.h)"}


def gmg_reference_comparison(args, dev):
    """The reference's own matrix-free multigrid (step-37's classes in the unmodified deal.II of oracle/_ref,
    driver oracle/ref_drivers/ref_gmg.cc, one host core: this build has neither MPI nor TBB) and the engine on
    the same, smaller problem with the same tolerance."""
    import subprocess
    import tempfile
    import torch
    import dealii_b200
    exe = os.path.join(ROOT, "oracle", "_ref", "bin", f"ref_gmg_q{args.degree}")
    r = args.gmg_reference_refinements
    if not os.path.exists(exe) or r > args.refinements:
        return None
    with tempfile.TemporaryDirectory() as tmp:
        try:
            out = subprocess.run([exe, "3", str(r), args.gmg_levels, "constant", tmp, "timing"], capture_output=True,
                                 text=True, timeout=600)
            ref = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
        except Exception as e:            # the checker is optional for the product's own numbers
            return {"unavailable": repr(e)[:200]}
    mg = dealii_b200.GeometricMultigrid.for_hyper_cube(3, args.degree, r, number=args.gmg_levels)
    mesh = dealii_b200.HyperCubeMesh(3, args.degree, refinements=r, dirichlet_boundary=True, mark_constrained_l2g=True)
    mf = dealii_b200.MatrixFree("f64", dev)
    mf.reinit_from_mesh(mesh)
    A = dealii_b200.LaplaceOperator(mf)
    plain = dealii_b200.MatrixFree("f64", dev)
    plain.reinit_from_mesh(dealii_b200.HyperCubeMesh(3, args.degree, refinements=r))
    b = mf.initialize_dof_vector()
    one = torch.ones_like(b)
    dealii_b200.MatrixFreeOperator(plain, grad_constant=0.0, mass_constant=1.0).vmult(b, one)   # rhs_i = (phi_i, 1)
    mf.set_constrained_values(0.0, b)
    best = None
    for rep in range(3):
        x = mf.initialize_dof_vector()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        control = dealii_b200.SolverControl(100, ref["relative_tolerance"] * float(b.norm()))
        dealii_b200.SolverCG(control).solve(A, x, b, mg)
        e1.record()
        torch.cuda.synchronize()
        best = e0.elapsed_time(e1) if best is None else min(best, e0.elapsed_time(e1))
    return {"n_dofs": ref["n_dofs"], "refinements": r, "relative_tolerance": ref["relative_tolerance"],
            "reference": {"seconds": ref["seconds"], "iterations": ref["iterations"], "cores": ref["cores"],
                          "what": "deal.II SolverCG + PreconditionMG (MGTransferMatrixFree, PreconditionChebyshev) on the host"},
            "engine": {"seconds": best * 1e-3, "iterations": control.last_step()}}


def run_sweep(args):
    """Degree sweep Q1-Q8, 3D, Cartesian (affine) and deformed (every cell general: full Jacobian,
    MappingQ1 on displaced vertices) meshes, FP64 and FP32: vmult GDoF/s and fraction of the HBM
    roofline with the algorithmic bytes of SURVEY.md 8(d) (2 s per dof + 6 s ((p+1)/p)^3 of merged
    metric on general cells).  Clocks are sampled over the whole sweep."""
    import torch
    import dealii_b200
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(0)
    peak, peak_src = measured_peak()
    sampler = ClockSampler(0)
    sampler.start()
    sampler.mark()
    refinements = {1: 8, 2: 7, 3: 7, 4: 6, 5: 6, 6: 6, 7: 5, 8: 5}
    rows = []
    for degree in range(1, 9):
        for amp in (0.0, 0.05):
            mesh = dealii_b200.HyperCubeMesh(3, degree, refinements=refinements[degree], deformation_amplitude=amp)
            for number in ("f64", "f32"):
                mf = dealii_b200.MatrixFree(number).reinit_from_mesh(mesh)
                op = dealii_b200.LaplaceOperator(mf)
                x = torch.rand(mf.n_owned, dtype=mf.torch_dtype, device="cuda")
                y = mf.initialize_dof_vector()
                for _ in range(3):
                    op.vmult(y, x)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 10
                e0.record()
                for _ in range(reps):
                    op.vmult(y, x)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                sz = 8 if number == "f64" else 4
                bpd = 2 * sz + (6 * sz * ((degree + 1) / degree) ** 3 if amp else 0.0)
                info = mf.bulk_info()
                rows.append({"degree": degree, "number": number, "mesh": "deformed" if amp else "cartesian",
                             "cell_kind": int(mf.info.cell_kind), "n_dofs": mf.n_owned, "ms_per_vmult": ms,
                             "gdofs": mf.n_owned / ms / 1e6, "algorithmic_bytes_per_dof": bpd,
                             "roofline_frac": bpd * mf.n_owned / (ms * 1e-3) / 1e9 / peak,
                             "cells_in_bricks": int(mf.info.n_bricks * mf.info.cells_per_brick),
                             "brick_path": info["path"] if mf.info.n_bricks else None})
                del mf, op, x, y
                torch.cuda.empty_cache()
            del mesh
    clocks = sampler.stop()
    clocks["window"] = "the whole sweep"
    return {"metric": "vmult_degree_sweep", "unit": UNIT, "n_gpus": 1, "higher_is_better": True, "data": "synthetic",
            "peak": peak, "peak_source": peak_src, "clocks": clocks, "sweep": rows,
            "config": {"workload": "3D Q1-Q8 Laplace vmult, Cartesian and deformed hyper_cube, f64 and f32 (BASELINE configs[4])",
                       "l2": "vectors of 17-57 M dofs (>= L2 except f32 at 17 M: 68 MB)"}}


class StdoutForJsonOnly:
    """Library banners (e.g. "NCCL version ..." printed by libnccl to fd 1) must not mix with the
    one JSON line of the contract: route fd 1 to stderr while the benchmark runs, restore it for
    the final print."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


if __name__ == "__main__":
    a = parse_args()
    with StdoutForJsonOnly() as guard:
        line = run_reference(a) if a.impl == "reference" else (run_sweep(a) if a.sweep else run_engine(a))
    if line is not None:
        print(json.dumps(line), flush=True)
