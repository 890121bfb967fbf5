"""Generates tests/golden/exact_cell.npz: the EXACT (symbolic, rational) integrals of the
reference's weak-form statements (src/NavierStokesSolver.cpp:249-311) on one generic triangle
with prescribed nodal data.  Every integrand is a polynomial of degree <= 5 on an affine triangle,
so sympy integrates it in closed form — a known-answer pin that does not depend on deal.II, on the
quadrature table or on this repo's code.  Run:  python tests/golden/make_exact_cell.py
"""
import os
import numpy as np
import sympy as sp

x, y = sp.symbols("x y")
R = sp.Rational
V = [(R(1, 10), R(1, 5)), (R(13, 10), R(2, 5)), (R(1, 2), R(11, 10))]   # CCW triangle
nu, rho, dt, p_out = R(1, 1000), R(1), R(1, 20), R(10)
f = (R(0), R(-3, 10))
# nodal data in FESystem local order (vertex v: ux,uy,p ; line l: ux,uy)
rng = np.random.default_rng(12345)
sol = [R(int(v), 64) for v in rng.integers(-64, 64, 15)]
old = [R(int(v), 64) for v in rng.integers(-64, 64, 15)]

l0, l1, l2 = 1 - x - y, x, y
psi = [l0 * (2 * l0 - 1), l1 * (2 * l1 - 1), l2 * (2 * l2 - 1), 4 * l0 * l1, 4 * l1 * l2, 4 * l2 * l0]
chi = [l0, l1, l2]
J = sp.Matrix([[V[1][0] - V[0][0], V[2][0] - V[0][0]], [V[1][1] - V[0][1], V[2][1] - V[0][1]]])
det = J.det()
JinvT = J.inv().T


def grad(fn):
    g = JinvT * sp.Matrix([sp.diff(fn, x), sp.diff(fn, y)])
    return [sp.expand(g[0]), sp.expand(g[1])]


def integ(e):
    return sp.integrate(sp.integrate(sp.expand(e), (y, 0, 1 - x)), (x, 0, 1)) * abs(det)


def local(i):
    if i < 9:
        v, r = divmod(i, 3)
        return (1, 2, v) if r == 2 else (0, r, v)
    return (0, (i - 9) % 2, 3 + (i - 9) // 2)


# vector-valued shape functions: value[a], gradient[a][b] = d_b phi_a, divergence, pressure value
val, gr, dv, pv = [], [], [], []
for i in range(15):
    isp, comp, k = local(i)
    v = [0, 0]
    g = [[0, 0], [0, 0]]
    if isp:
        val.append(v), gr.append(g), dv.append(0), pv.append(chi[k])
    else:
        gk = grad(psi[k])
        v[comp] = psi[k]
        g[comp] = gk
        val.append(v), gr.append(g), dv.append(gk[comp]), pv.append(0)

U = [sum(sol[i] * val[i][a] for i in range(15)) for a in range(2)]
Uo = [sum(old[i] * val[i][a] for i in range(15)) for a in range(2)]
G = [[sum(sol[i] * gr[i][a][b] for i in range(15)) for b in range(2)] for a in range(2)]
P = sum(sol[i] * pv[i] for i in range(15))

A = np.zeros((15, 15))
M = np.zeros((15, 15))
Rv = np.zeros(15)
for i in range(15):
    for j in range(15):
        e = sum(val[i][a] * val[j][a] for a in range(2)) / dt
        e += nu * rho * sum(gr[i][a][b] * gr[j][a][b] for a in range(2) for b in range(2))
        # (grad u * phi_j) . phi_i : contracts the LAST index of grad u (cpp:259-263)
        e += rho * sum(sum(G[a][b] * val[j][b] for b in range(2)) * val[i][a] for a in range(2))
        # (u * grad phi_j) . phi_i : contracts u with the FIRST index of grad phi_j (cpp:265-269)
        e += rho * sum(sum(U[a] * gr[j][a][b] for a in range(2)) * val[i][b] for b in range(2))
        e -= dv[i] * pv[j]
        e -= dv[j] * pv[i]
        A[i, j] = float(integ(e))
        M[i, j] = float(integ(pv[i] * pv[j] / nu))
    e = -rho * sum((U[a] - Uo[a]) / dt * val[i][a] for a in range(2))
    e -= nu * rho * sum(G[a][b] * gr[i][a][b] for a in range(2) for b in range(2))
    e -= rho * sum(sum(U[c] * G[c][a] for c in range(2)) * val[i][a] for a in range(2))
    e += P * dv[i]
    e += sum(f[a] * val[i][a] for a in range(2))
    Rv[i] = float(integ(e))
    print("row", i, flush=True)

# Neumann term on face 1 (v1 -> v2): -p_out * n . phi_i integrated exactly along the edge
s = sp.symbols("s")
ex, ey = V[2][0] - V[1][0], V[2][1] - V[1][1]
L = sp.sqrt(ex * ex + ey * ey)
n = (ey / L, -ex / L)
Nv = np.zeros(15)
for i in range(15):
    isp, comp, k = local(i)
    if isp:
        continue
    edge = psi[k].subs({x: 1 - s, y: s}, simultaneous=True)   # reference edge (1,0)->(0,1)
    Nv[i] = float(-p_out * n[comp] * sp.integrate(edge, (s, 0, 1)) * L)

out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "exact_cell.npz")
np.savez(out, vertices=np.array([[float(a), float(b)] for a, b in V]), sol=np.array([float(v) for v in sol]),
         old=np.array([float(v) for v in old]), nu=float(nu), rho=float(rho), deltat=float(dt), p_out=float(p_out),
         forcing=np.array([float(f[0]), float(f[1])]), A=A, M=M, R=Rv, neumann_face=1, N=Nv)
print("wrote", out)
