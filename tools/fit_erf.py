"""Reproduces the coefficients of gelu_erf() in csrc/gemm.cu: erfc(z) = 2^(z Q(z)) on [0, 4], Q of degree 4,
iteratively re-weighted least squares towards the minimax fit; prints the fp32 coefficients and the max errors."""
import numpy as np
from scipy.special import erf, erfc

T, DEG = 4.0, 4
t = np.linspace(1e-4, T, 20001)
target = np.log2(erfc(t)) / t
w = np.ones_like(t)
for _ in range(60):
    c = np.polyfit(t, target, DEG, w=w)
    err = np.abs(1 - np.exp2(t * np.polyval(c, t)) - erf(t))
    w = w * (1 + 3 * err / err.max())
c32 = c.astype(np.float32)
x = np.linspace(-8, 8, 800001)
z = np.minimum(np.abs(x) / np.sqrt(2), T)
e = 1 - np.exp2(z * np.polyval(c32.astype(np.float64), z))
gelu = 0.5 * x + 0.5 * np.abs(x) * e
ref = 0.5 * x * (1 + erf(x / np.sqrt(2)))
print("coefficients (highest degree first):", [float(v) for v in c32])
print("max |erf err|", np.abs(e - erf(np.abs(x) / np.sqrt(2))).max(), " max |gelu err|", np.abs(gelu - ref).max())
