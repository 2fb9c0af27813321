"""Fit  erfc(t) ~= exp2(-t * Q(t)),  t in [0, TMAX]  (Q a polynomial; for t > TMAX the argument is
clamped -- erfc(TMAX) < 3e-9 is below fp32 resolution next to 1).  The fit minimises the ABSOLUTE
error of erfc (weight erfc(t) on the exponent error) with an LP minimax.  The exponent form keeps
fp32 evaluation error ~1 ulp of 1.0 because a rounding error d in the exponent only costs erfc(t)*d.
Used by csrc/nm_device.cuh::gelu_pair.  Run: python tools/fit_erfc.py"""
import numpy as np
from scipy.optimize import linprog
from scipy.special import erfc, erfcx

TMAX = 4.2
LOG2E = 1.4426950408889634

def g_of(t):                       # -log2(erfc(t)), stable via erfcx
    return (t * t - np.log(erfcx(t))) * LOG2E

def fit(n, grid):
    y = g_of(grid)
    w = erfc(grid) * np.log(2.0)   # d erfc = erfc * ln2 * d(exponent in base 2)
    A = np.stack([grid ** (k + 1) for k in range(n)], axis=1)
    m = len(grid)
    cvec = np.zeros(n + 1); cvec[-1] = 1.0
    Aw = A * w[:, None]
    Aub = np.block([[Aw, -np.ones((m, 1))], [-Aw, -np.ones((m, 1))]])
    bub = np.concatenate([y * w, -y * w])
    res = linprog(cvec, A_ub=Aub, b_ub=bub, bounds=[(None, None)] * n + [(0, None)], method="highs")
    return res.x[:n], res.x[-1]

grid = np.concatenate([np.linspace(1e-6, 1, 2000), np.linspace(1, TMAX, 3000)])
for n in (6, 7, 8, 9):
    q, eps = fit(n, grid)
    fine = np.linspace(0, TMAX, 400001)
    expo = sum(qk * fine ** (k + 1) for k, qk in enumerate(q))
    err = np.abs(np.exp2(-expo) - erfc(fine)).max()
    f32 = np.float32
    t = fine.astype(f32)
    acc = np.full_like(t, f32(q[-1]))
    for qk in q[-2::-1]:
        acc = (acc * t + f32(qk)).astype(f32)
    e = (acc * t).astype(f32)
    err32 = np.abs(np.exp2(-e.astype(np.float64)) - erfc(fine)).max()
    print(f"n={n} lp_eps={eps:.3e} maxerr={err:.3e} f32(no-fma)err={err32:.3e}")
    print("   q =", ", ".join(f"{qk:.9e}f" for qk in q))
