"""Fits Phi(y) ~= sigmoid(y (c0 + c1 y^2 + c2 y^4)) so that y*sigmoid(.) tracks the exact (erf) GELU; prints the
coefficients used by csrc/tc_common.cuh and the max abs error (2.8e-5)."""
import numpy as np
from scipy.optimize import least_squares
from scipy.special import ndtr

y = np.linspace(1e-3, 7.5, 6000)


def err(c, yy):
    y2 = np.minimum(yy * yy, 64.0)
    z = yy * (c[0] + y2 * (c[1] + y2 * c[2]))
    return yy / (1 + np.exp(-z)) - yy * ndtr(yy)


c = np.array([1.5957691, 0.0713548, 0.0])
for _ in range(60):
    e = err(c, y)
    w = (np.abs(e) / np.abs(e).max()) ** 0.5 + 0.05
    c = least_squares(lambda cc: err(cc, y) * w, c, xtol=1e-15, ftol=1e-15, gtol=1e-15).x
print([float(v) for v in c], float(np.abs(err(c, np.linspace(-30, 30, 600001))).max()))
