"""Gauss-Legendre quadrature on [x1, x2] -- restates ``quadpoints`` of the reference.

Reference: @egdstmodel/egdstmodel.m:1504-1529 (Newton iteration on the Legendre
polynomial, symmetric fill).  ``solve`` calls it as ``quadpoints(ny, 0, 1)`` and
stores ``[weights abscissas]`` column-wise (egdstmodel.m:1157-1160); the solver
gateway then maps the abscissas through the inverse normal cdf
(egdst_solver.c:162-164).
"""
import math

import numpy as np

_EPS = np.finfo(np.float64).eps


def quadpoints(n: int, x1: float = 0.0, x2: float = 1.0):
    """Return (x, w): n abscissas and weights of Gauss-Legendre quadrature on [x1, x2]."""
    n = int(n)
    x = np.zeros(n)
    w = np.zeros(n)
    m = int((n + 1) / 2)  # MATLAB `1:(n+1)/2` stops at floor
    xm = 0.5 * (x2 + x1)
    xl = 0.5 * (x2 - x1)
    for i in range(1, m + 1):
        z = math.cos(math.pi * (i - 0.25) / (n + 0.5))
        z1 = 2.0
        pp = 1.0
        guard = 0
        while abs(z - z1) > _EPS and guard < 200:
            p1 = 1.0
            p2 = 0.0
            for j in range(1, n + 1):
                p3 = p2
                p2 = p1
                p1 = ((2.0 * j - 1.0) * z * p2 - (j - 1.0) * p3) / j
            pp = n * (z * p1 - p2) / (z * z - 1.0)
            z1 = z
            z = z1 - p1 / pp
            guard += 1
        x[i - 1] = xm - xl * z
        x[n - i] = xm + xl * z
        w[i - 1] = 2.0 * xl / ((1.0 - z * z) * pp * pp)
        w[n - i] = w[i - 1]
    return x, w


def model_quadrature(ny: int) -> np.ndarray:
    """``model.quadrature`` as the class stores it: [weights(ny), abscissas(ny)] flattened
    column-major (egdstmodel.m:1159).  Abscissas are in (0,1); the solver boundary applies cdfni."""
    qx, qw = quadpoints(ny, 0.0, 1.0)
    return np.concatenate([qw, qx]).astype(np.float64)
