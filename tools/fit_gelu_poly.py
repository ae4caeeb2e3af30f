"""MUFU-free erf-GELU: Phi(x) - 0.5 ~= x * Q(x^2) on |x| <= R (clamped), Q of degree d (minimax by reweighted LSQ)."""
import numpy as np
from scipy.special import erf

for R in (4.0, 4.25, 4.5):
    x = np.linspace(0, R, 40001)[1:]
    u = x * x
    target = (0.5 * (1 + erf(x / np.sqrt(2))) - 0.5) / x       # Q(u)
    for d in (6, 7, 8, 9):
        V = np.vander(u / (R * R), d + 1, increasing=True)
        w = x * x                                                # error in y = x*Phi is x * (x*Q err)
        for _ in range(80):
            c, *_ = np.linalg.lstsq(V * w[:, None], target * w, rcond=None)
            err = np.abs((V @ c - target) * x * x)
            w = w * (1 + 3 * err / err.max())
            w /= w.mean()
        # evaluate in float32 Horner on wide range with clamp
        xx = np.linspace(-10, 10, 200001).astype(np.float32)
        xc = np.clip(xx, -R, R).astype(np.float32)
        uu = (xc * xc / np.float32(R * R)).astype(np.float32)
        q = np.float32(c[-1]) * np.ones_like(uu)
        for k in range(d - 1, -1, -1):
            q = (q * uu + np.float32(c[k])).astype(np.float32)
        phi = (np.float32(0.5) + xc * q).astype(np.float32)
        y = (xx * phi).astype(np.float32)
        ref = xx.astype(np.float64) * 0.5 * (1 + erf(xx.astype(np.float64) / np.sqrt(2)))
        e = np.abs(y - ref)
        print("R=%.2f d=%d max abs err %.2e at x=%.2f   (|x|<=R: %.2e)" % (R, d, e.max(), xx[e.argmax()], e[np.abs(xx) <= R].max()))
        if d == 8 and R == 4.25:
            print("   coeffs (in u/R^2):", [float(np.float32(v)) for v in c])
