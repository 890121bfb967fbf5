"""Assembly kernel study: variants 0/1/2 on a level-L mesh, sweeping the shared-memory image cap and the
register budget of the packet-based kernel (env NSG_ASM_STAGE_CAP / NSG_ASM3_MINB)."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module("navier-stokes-dealii_b200")
L = int(sys.argv[1]) if len(sys.argv) > 1 else 4
MESH = sys.argv[2] if len(sys.argv) > 2 else "cmy"
caps = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1 << 30, 5000]
m, d, part, (ld, lv), neumann, sol = bench.build_problem(pkg, MESH, L, 1, 0)
print("cells", m.n_cells, "N", d.n, "nnz", part.nnz_jac, flush=True)
ref = None
for cap in caps:
    os.environ["NSG_ASM_STAGE_CAP"] = str(cap)
    dev = pkg.DeviceProblem(part, 0)
    dev.set_params(neumann_id=neumann)
    dev.set_solution(sol); dev.set_solution_old(0.9 * sol)
    for av, minb, pf in ((4, 3, 0), (4, 4, 0), (4, 5, 0)):
        if True:
            os.environ["NSG_ASM3_MINB"] = str(minb)
            os.environ.pop("NSG_ASM_PF", None)
            os.environ["NSG_ASM_CONCURRENT"] = "0" if pf == -2 else "1"
            if pf >= 0:
                os.environ["NSG_ASM_PF"] = str(pf)
            dev.set_tuning(1, av)
            dev.time_kernel(0, 2)
            ms = dev.time_kernel(0, 5)
            J, R = np.concatenate([dev.get_matrix_values(), dev.get_pm_values()]), dev.get_residual()
            if ref is None:
                ref = (J, R)
            ej = np.abs(J - ref[0]).max() / np.abs(ref[0]).max()
            er = np.abs(R - ref[1]).max() / np.abs(ref[1]).max()
            print(f"cap {cap:>10d} variant {av} minb {minb} pf {pf:4d}: {ms:8.3f} ms  {d.n / ms / 1e3:9.1f} MDoF/s   |dJ| {ej:.2e} |dR| {er:.2e}", flush=True)
    dev.close()
