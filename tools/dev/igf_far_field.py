"""Accuracy of the far-field (Taylor) form of the integrated Green function against the exact
8-corner difference of the antiderivative (space_charge_kick.py:103-123) in 50-digit arithmetic.
Development aid for the threshold used by sc_green_* (csrc/space_charge.cu)."""
import itertools
import sys

import mpmath as mp
import numpy as np

mp.mp.dps = 50


def antiderivative(x, y, z):
    r = mp.sqrt(x * x + y * y + z * z)
    return (-z * z / 2 * mp.atan(x * y / (z * r)) - y * y / 2 * mp.atan(x * z / (y * r))
            - x * x / 2 * mp.atan(y * z / (x * r)) + y * z * mp.asinh(x / mp.sqrt(y * y + z * z))
            + x * z * mp.asinh(y / mp.sqrt(x * x + z * z)) + x * y * mp.asinh(z / mp.sqrt(x * x + y * y)))


def exact(i, j, k, h):
    total = mp.mpf(0)
    for a, b, c in itertools.product((0, 1), repeat=3):
        sign = (-1) ** (3 - a - b - c)
        total += sign * antiderivative((i - 0.5 + a) * h[0], (j - 0.5 + b) * h[1], (k - 0.5 + c) * h[2])
    return total


def far(i, j, k, h, dtype=np.float32):
    f = dtype
    x, y, z = f(i * h[0]), f(j * h[1]), f(k * h[2])
    hx2, hy2, hz2 = f(h[0] ** 2), f(h[1] ** 2), f(h[2] ** 2)
    volume = f(h[0] * h[1] * h[2])
    x2, y2, z2 = x * x, y * y, z * z
    r2 = x2 + y2 + z2
    inv_r2 = f(1) / r2
    inv_r = np.sqrt(inv_r2)
    tx, ty, tz = x2 * inv_r2, y2 * inv_r2, z2 * inv_r2
    ex, ey, ez = hx2 * inv_r2, hy2 * inv_r2, hz2 * inv_r2
    s2 = ex * (f(3) * tx - f(1)) + ey * (f(3) * ty - f(1)) + ez * (f(3) * tz - f(1))
    q = lambda t: (f(105) * t - f(90)) * t + f(9)  # noqa: E731
    s4a = ex * ex * q(tx) + ey * ey * q(ty) + ez * ez * q(tz)
    m = lambda ta, tb: f(105) * ta * tb - f(15) * (ta + tb) + f(3)  # noqa: E731
    s4b = ex * ey * m(tx, ty) + ex * ez * m(tx, tz) + ey * ez * m(ty, tz)
    return volume * inv_r * (f(1) + s2 * f(1 / 24) + s4a * f(1 / 1920) + s4b * f(1 / 576))


def main():
    ratio = float(sys.argv[1]) if len(sys.argv) > 1 else 8.0
    rng = np.random.default_rng(0)
    for h in ([16e-6, 16e-6, 147e-6], [1.0, 1.0, 1.0], [3e-5, 1e-5, 2e-4], [1e-4, 2e-5, 1e-5]):
        hmax = max(h)
        worst32 = worst64 = 0.0
        count = 0
        # sample lattice points just beyond the threshold and further out
        for _ in range(4000):
            i, j, k = (int(v) for v in rng.integers(0, 64, 3))
            r = np.sqrt((i * h[0]) ** 2 + (j * h[1]) ** 2 + (k * h[2]) ** 2)
            if r < ratio * hmax or r > 1.5 * ratio * hmax:
                continue
            e = exact(i, j, k, [mp.mpf(v) for v in h])
            worst32 = max(worst32, abs(float((mp.mpf(float(far(i, j, k, h))) - e) / e)))
            worst64 = max(worst64, abs(float((mp.mpf(float(far(i, j, k, h, np.float64))) - e) / e)))
            count += 1
        print(f"h = {h}: {count} points with r in [{ratio}, {1.5 * ratio}] h_max: worst relative "
              f"error float32 {worst32:.2e}, float64 {worst64:.2e}")


if __name__ == "__main__":
    main()
