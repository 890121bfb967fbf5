"""GPU check of assembly variant 5 (fan scheme): parity against the CPU oracle and variant 4 on the golden meshes
in the three modes, for the three packet-staging paths, then timing on a refined mesh.
usage: python scripts/fan_check.py [timing_level [timing_mesh]]"""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from conftest import analytic_state, row_scaled_err
from test_gpu_parity import build, CASES
from oracle.oracle import Oracle
pkg = importlib.import_module("navier-stokes-dealii_b200")
L = int(sys.argv[1]) if len(sys.argv) > 1 else 4
MESH = sys.argv[2] if len(sys.argv) > 2 else "cmy"
worst = 0.0
for case in CASES:
    m, d, part, calls, neumann, inlet = build(pkg, case)
    dev, o = pkg.DeviceProblem(part, 0), Oracle(part)
    for mode in ("newton", "steady", "stokes"):
        kw = dict(nu=0.001, rho=1.3, p_out=10.0, deltat=0.05, forcing=(0.0, -0.7), neumann_id=neumann,
                  use_mass=0 if mode == "steady" else 1, stokes=1 if mode == "stokes" else 0)
        sol, old = analytic_state(d), analytic_state(d, 0.9)
        o.set_params(**kw); o.set_solution(sol); o.set_solution_old(old); o.assemble()
        Jo, Mo, Ro = o.get_matrix_values(), o.get_pm_values(), o.get_residual()
        dev.set_params(**kw); dev.set_solution(sol); dev.set_solution_old(old)
        for variant, stage in ((4, 0), (5, 0)):
            dev.set_tuning(1, variant)
            dev.assemble()
            J, M, R = dev.get_matrix_values(), dev.get_pm_values(), dev.get_residual()
            ej = row_scaled_err(J, Jo, part.jac_rowptr); em = row_scaled_err(M, Mo, part.pm_rowptr)
            er = float(np.abs(R - Ro).max() / np.abs(Ro).max())
            worst = max(worst, ej, em, er) if variant == 5 else worst
            flag = "" if max(ej, em, er) < 5e-12 else "   <-- MISMATCH"
            print(f"{case:7s} {mode:7s} variant {variant} stage {stage}: |dJ| {ej:.2e} |dMp| {em:.2e} |dR| {er:.2e}{flag}", flush=True)
            if flag:
                rows = np.repeat(np.arange(len(part.jac_rowptr) - 1), np.diff(part.jac_rowptr))
                badi = np.flatnonzero(np.abs(J - Jo) > 1e-9 * max(np.abs(Jo).max(), 1e-300))
                print("   bad J entries:", len(badi), "of", len(J), "nan:", int(np.isnan(J).sum()))
                for i in badi[:12]:
                    print(f"     row {rows[i]} col {part.jac_col[i]} dev {J[i]:.6e} oracle {Jo[i]:.6e}")
                badr = np.flatnonzero(np.abs(R - Ro) > 1e-9 * np.abs(Ro).max())
                print("   bad R entries:", len(badr), badr[:12])
        # bitwise run-to-run
        dev.assemble()
        assert np.array_equal(dev.get_matrix_values(), J) and np.array_equal(dev.get_residual(), R)
    dev.close()
print("worst variant-5 error", worst, flush=True)

t0 = time.perf_counter()
m, d, part, (ld, lv), neumann, sol = bench.build_problem(pkg, MESH, L, 1, 0)
print(f"timing mesh {MESH} L{L}: cells {m.n_cells} N {d.n} nnz {part.nnz_jac}  host build {time.perf_counter() - t0:.1f} s", flush=True)
t0 = time.perf_counter()
dev = pkg.DeviceProblem(part, 0)
print(f"device setup {time.perf_counter() - t0:.1f} s", flush=True)
dev.set_params(neumann_id=neumann)
dev.set_solution(sol); dev.set_solution_old(0.9 * sol)
ms = dev.time_kernel(5, 3)
sms = 148
print(f"FP64 pipe micro-benchmark: {ms:.3f} ms per launch -> {sms * 67108864 * 2 / ms / 1e9:.1f} TFLOP/s (if {sms} SMs)", flush=True)
ref = None
for variant, stage, conc in ((4, 0, 1), (5, 0, 1), (5, 0, 0), (5, 600, 1), (5, 1200, 1), (5, 2400, 1), (5, 1200 | (300 << 16), 1),
                            (5, 1200 | (600 << 16), 1), (5, 2400 | (600 << 16), 1), (5, 2400 | (1200 << 16), 1), (5, 4800 | (1200 << 16), 1)):
    os.environ["NSG_ASM_CONCURRENT"] = str(conc)
    dev.set_tuning(1, variant)
    dev.set_tuning(6, stage)
    dev.time_kernel(0, 3)
    ms = dev.time_kernel(0, 10)
    J, R = np.concatenate([dev.get_matrix_values(), dev.get_pm_values()]), dev.get_residual()
    if ref is None:
        ref = (J, R)
    ej = np.abs(J - ref[0]).max() / np.abs(ref[0]).max(); er = np.abs(R - ref[1]).max() / np.abs(ref[1]).max()
    print(f"variant {variant} prefetch rec {stage & 0xffff} pk {stage >> 16} concurrent {conc}: {ms:8.3f} ms  {d.n / ms / 1e3:9.1f} MDoF/s   |dJ| {ej:.2e} |dR| {er:.2e}", flush=True)
dev.close()
