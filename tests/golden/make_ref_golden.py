"""Generates tests/golden/ref/*.npz by RUNNING THE REFERENCE ITSELF: oracle/_ref/bin/ref_dump_q<p>
(oracle/ref_drivers/ref_dump.cc linked against the unmodified deal.II built by oracle/build_ref.sh).

Every file holds, for one small configuration, the setup arrays Portable::MatrixFree built
(local_to_global with hanging-node redirection, ConstraintKinds masks, ...) and the results the
reference computed with them: dst of the CPU MatrixFree cell loop and of Portable::MatrixFree,
compute_diagonal, SolverCG iteration counts.  The GPU parity tests (tests/test_reference_parity.py)
feed the same arrays to the engine and compare PER ENTRY.

    python tests/golden/make_ref_golden.py          # needs oracle/_ref (this container only)
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
BIN = os.path.join(ROOT, "oracle", "_ref", "bin")
OUT = os.path.join(ROOT, "tests", "golden", "ref")
OUT_GMG = os.path.join(ROOT, "tests", "golden", "ref_gmg")

# name: (dim, degree, refinements, mesh, op, dirichlet, keep_jacobians)
CASES = {
    # configs[1]: 3D Laplace on the affine Cartesian hyper_cube, every degree
    "c2_q1": (3, 1, 3, "cartesian", "laplace", 0, False),
    "c2_q2": (3, 2, 3, "cartesian", "laplace", 0, False),
    "c2_q3": (3, 3, 2, "cartesian", "laplace", 0, False),
    "c2_q4": (3, 4, 2, "cartesian", "laplace", 0, False),
    "c2_q5": (3, 5, 1, "cartesian", "laplace", 0, False),
    "c2_q6": (3, 6, 1, "cartesian", "laplace", 0, False),
    "c2_q7": (3, 7, 1, "cartesian", "laplace", 0, False),
    "c2_q8": (3, 8, 1, "cartesian", "laplace", 0, False),
    "c2_q4_dirichlet": (3, 4, 2, "cartesian", "laplace", 1, False),
    # configs[2]: step-64 Helmholtz, variable coefficient, Q5 (and the shipped Q3), zero Dirichlet
    "c3_q5_step64": (3, 5, 2, "cartesian", "helmholtz_var", 1, False),
    "c3_q3_step64": (3, 3, 2, "cartesian", "helmholtz_var", 1, False),
    "c3_q3_hanging": (3, 3, 2, "hanging", "helmholtz_var", 1, False),
    "c3_q2_hanging_const": (3, 2, 2, "hanging", "helmholtz", 0, False),
    # configs[3]: Q3 Poisson with one refinement ball (hanging nodes), zero Dirichlet
    "c4_q3_ball": (3, 3, 3, "ball", "laplace", 1, False),
    "c4_q1_ball": (3, 1, 3, "ball", "laplace", 1, False),
    # configs[4]: deformed (full-Jacobian) cells
    "c5_q2_deformed": (3, 2, 2, "deformed", "laplace", 0, True),
    "c5_q4_deformed": (3, 4, 2, "deformed", "laplace", 0, False),
    "c5_q3_deformed_helmholtz": (3, 3, 2, "deformed", "helmholtz_var", 1, False),
    # opt-in numbering DoFRenumbering::lexicographic (the strided brick path)
    "lex_q4": (3, 4, 2, "cartesian", "laplace", 1, False, "lexicographic"),
    "lex_q2": (3, 2, 3, "cartesian", "helmholtz", 0, False, "lexicographic"),
    # 2D
    "d2_q2_hanging": (2, 2, 3, "hanging", "helmholtz_var", 1, True),
    "d2_q4_cartesian": (2, 4, 3, "cartesian", "laplace", 1, False),
}

# geometric multigrid of step-37 (oracle/ref_drivers/ref_gmg.cc): name: (dim, degree, refinements, level number,
# coefficient).  Big vectors are dropped for the larger cases (iteration counts / eigenvalues / norms stay).
GMG_CASES = {
    "gmg_q2_r2_f64": (3, 2, 2, "f64", "constant", True),
    "gmg_q2_r3_f64_step37": (3, 2, 3, "f64", "step37", True),
    "gmg_q4_r2_f64": (3, 4, 2, "f64", "constant", True),
    "gmg_q4_r2_f32_step37": (3, 4, 2, "f32", "step37", True),
    "gmg_q1_r3_f64": (3, 1, 3, "f64", "constant", True),
    "gmg_q3_r3_f32": (3, 3, 3, "f32", "constant", False),
    "gmg_q2_r5_f32_step37": (3, 2, 5, "f32", "step37", False),
    "gmg_q4_r4_f64": (3, 4, 4, "f64", "constant", False),
    "gmg_d2_q3_r4_f64": (2, 3, 4, "f64", "step37", True),
    # general cells on every level (the generator's sine displacement)
    "gmg_q2_r3_f64_deformed": (3, 2, 3, "f64", "step37", True, 0.05),
    "gmg_q3_r2_f32_deformed": (3, 3, 2, "f32", "constant", True, 0.04),
}

# over-integration (n_q_points_1d = degree + 1 + extra; ref_dump compiled with -DREF_NQ_EXTRA=extra):
# name: (dim, degree, extra, refinements, mesh, op, dirichlet)
NQ_CASES = {
    "nq_q2_e1_deformed_var": (3, 2, 1, 2, "deformed", "helmholtz_var", 1),
    "nq_q3_e1_cartesian": (3, 3, 1, 1, "cartesian", "helmholtz", 0),
    "nq_d2_q2_e1_cartesian_var": (2, 2, 1, 3, "cartesian", "helmholtz_var", 1),
    "nq_q2_e1_hanging_var": (3, 2, 1, 2, "hanging", "helmholtz_var", 1),
    "nq_d2_q3_e1_hanging": (2, 3, 1, 3, "hanging", "helmholtz", 0),
}
OUT_NQ = os.path.join(ROOT, "tests", "golden", "ref_nq")


def run_nq_case(name, spec):
    dim, degree, extra, ref, mesh, op, dirichlet = spec
    exe = os.path.join(BIN, f"ref_dump_q{degree}_nq{extra}")
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.check_call([exe, str(dim), str(degree), str(ref), mesh, op, str(dirichlet), tmp])
        man = json.load(open(os.path.join(tmp, "manifest.json")))
        out = {"n_q_points_1d": np.array(degree + 1 + extra)}
        for k, v in man.items():
            if isinstance(v, dict):
                if k in ("q_points",) or (k == "coefficient" and op != "helmholtz_var"):
                    continue
                out[k] = np.fromfile(os.path.join(tmp, v["file"]), dtype=v["dtype"])
            else:
                out[k] = np.array(v)
        os.makedirs(OUT_NQ, exist_ok=True)
        np.savez_compressed(os.path.join(OUT_NQ, name + ".npz"), **out)
        print(name, {k: v.item() for k, v in out.items() if v.ndim == 0})


def run_gmg_case(name, spec):
    dim, degree, ref, number, coef, keep_vectors = spec[:6]
    extra = [f"deform={spec[6]}"] if len(spec) > 6 else []
    exe = os.path.join(BIN, f"ref_gmg_q{degree}")
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.check_call([exe, str(dim), str(ref), number, coef, tmp] + extra)
        man = json.load(open(os.path.join(tmp, "manifest.json")))
        out = {}
        for k, v in man.items():
            if isinstance(v, dict):
                if not keep_vectors and not k.startswith("level_l2g_0"):
                    continue
                if k == "level_l2g_%d" % ref and not keep_vectors:
                    continue
                out[k] = np.fromfile(os.path.join(tmp, v["file"]), dtype=v["dtype"])
                assert out[k].size == v["size"]
            else:
                out[k] = np.array(v)
        os.makedirs(OUT_GMG, exist_ok=True)
        np.savez_compressed(os.path.join(OUT_GMG, name + ".npz"), **out)
        print(name, {k: v.item() for k, v in out.items() if v.ndim == 0})


def run_case(name, spec):
    dim, degree, ref, mesh, op, dirichlet, keep_jac = spec[:7]
    extra = list(spec[7:])
    exe = os.path.join(BIN, f"ref_dump_q{degree}")
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.check_call([exe, str(dim), str(degree), str(ref), mesh, op, str(dirichlet), tmp] + extra)
        man = json.load(open(os.path.join(tmp, "manifest.json")))
        out = {}
        for k, v in man.items():
            if isinstance(v, dict):
                if k in ("inv_jacobian", "JxW") and not keep_jac:
                    continue
                if k in ("q_points",) or (k == "coefficient" and op != "helmholtz_var"):
                    continue
                out[k] = np.fromfile(os.path.join(tmp, v["file"]), dtype=v["dtype"])
                assert out[k].size == v["size"]
            else:
                out[k] = np.array(v)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        print(name, {k: (v.shape if v.ndim else v.item()) for k, v in out.items() if v.ndim == 0 or k in ("src",)})


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    names = sys.argv[1:] or (list(CASES) + list(GMG_CASES) + list(NQ_CASES))
    for n in names:
        if n in NQ_CASES:
            run_nq_case(n, NQ_CASES[n])
        elif n in GMG_CASES:
            run_gmg_case(n, GMG_CASES[n])
        else:
            run_case(n, CASES[n])
