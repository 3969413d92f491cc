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


def sincos_half_lim(x):
    """hrt_math.cuh sincos_half_lim: no fold, degree 11 / 10 (numpy has no FMA: the device result is a little better)."""
    x = x.astype(f)
    w = (x * x).astype(f)
    p = ((((f(-2.4735086867622158e-08) * w + f(2.7569451503950404e-06)) * w + f(-0.00019841609173454344)) * w
          + f(0.008333335630595684)) * w + f(-0.1666666716337204)).astype(f)
    q = ((((f(-2.629732875902846e-07) * w + f(2.4774593839538284e-05)) * w + f(-0.001388865290209651)) * w
          + f(0.0416666604578495)) * w + f(-0.5)).astype(f)
    return (p * w * x + x).astype(f), (q * w + f(1)).astype(f)


if __name__ == "__main__":
    x = np.linspace(-np.pi / 2, np.pi / 2, 2000001)
    s, c = sincos_half(x)
    xs = x.astype(f).astype(np.float64)
    print("folded pair: max |sin err|", np.abs(s - np.sin(xs)).max(), "max |cos err|", np.abs(c - np.cos(xs)).max())
    x = np.linspace(-1.5724, 1.5724, 2000001)
    s, c = sincos_half_lim(x)
    xs = x.astype(f).astype(np.float64)
    print("no-fold pair on [-1.5724, 1.5724]: max |sin err|", np.abs(s - np.sin(xs)).max(), "max |cos err|", np.abs(c - np.cos(xs)).max())
