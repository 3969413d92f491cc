"""Accuracy of the polynomial half-angle sincos used by the fast FK / IK paths (hrt_math.cuh
sincos_half_f), emulated in numpy fp32 over [-pi/2, pi/2]."""
import numpy as np

f = np.float32


def sincos_half(x):
    x = x.astype(f)
    ax = np.abs(x)
    fold = ax > f(0.78539816)
    y = np.where(fold, (f(1.5707963705062866) - ax) + f(-4.371139e-8), ax).astype(f)
    z = (y * y).astype(f)
    sp = (((f(-1.9515295891e-4) * z + f(8.3321608736e-3)) * z + f(-1.6666654611e-1)) * z * y + y).astype(f)
    cp = (((f(2.443315711809948e-5) * z + f(-1.388731625493765e-3)) * z + f(4.166664568298827e-2)) * z * z
          - f(0.5) * z + f(1)).astype(f)
    return np.copysign(np.where(fold, cp, sp), x), np.where(fold, sp, cp)


if __name__ == "__main__":
    x = np.linspace(-np.pi / 2, np.pi / 2, 2000001)
    s, c = sincos_half(x)
    xs = x.astype(f).astype(np.float64)
    print("max |sin err|", np.abs(s - np.sin(xs)).max(), "max |cos err|", np.abs(c - np.cos(xs)).max())
