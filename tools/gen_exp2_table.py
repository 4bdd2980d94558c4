#!/usr/bin/env python
"""Generates the 2^(j/32) table of ecckd_b200/csrc/trans_core.cuh (correctly rounded, via mpmath)
and sweeps the accuracy of the 1 - exp(-kd*tau) scheme emulated operation by operation."""
import struct
import sys

import mpmath as mp

mp.mp.prec = 400


def rnd(x):  # round an mpf to the nearest double, returned as mpf (exact)
    return mp.mpf(float(x))


def fma(a, b, c):
    return rnd(a * b + c)


def table():
    return [float(mp.power(2, mp.mpf(j) / 32)) for j in range(32)]


def consts(kd):
    kd = mp.mpf(kd)
    ln2 = mp.log(2)
    a = rnd(-kd * 32 / ln2)
    b = rnd(ln2 / (32 * kd))
    c = [rnd((-kd) ** (i + 1) / mp.factorial(i + 1)) for i in range(6)]
    return a, b, c


def one_minus_exp(tau, kd, T, K):
    a, b, c = K
    magic = mp.mpf(6755399441055744.0)
    tau = mp.mpf(tau)
    t = fma(tau, a, magic)
    nf = rnd(t - magic)
    n = int(nf)
    rho = fma(nf, b, tau)
    q = c[5]
    for i in (4, 3, 2, 1, 0):
        q = fma(q, rho, c[i])
    p = rnd(q * rho)
    j = n & 31
    k = (n - j) // 32
    Tj = mp.mpf(T[j]) * mp.power(2, k)
    om = rnd(1 - Tj)
    return float(fma(-Tj, p, om))


if __name__ == "__main__":
    T = table()
    if "--table" in sys.argv:
        for j in range(0, 32, 4):
            print("    " + ", ".join(float.hex(T[i]) for i in range(j, j + 4)) + ",")
        sys.exit(0)
    import random

    random.seed(1)
    for kd in (1.66, 3.32, 4.98, 16.6):
        K = consts(kd)
        worst_abs, worst_rel = 0.0, 0.0
        for _ in range(20000):
            e = random.uniform(-12, 2.7)
            tau = 10.0 ** e
            got = one_minus_exp(tau, kd, T, K)
            ref = 1 - mp.exp(-mp.mpf(kd) * mp.mpf(tau))
            err = abs(mp.mpf(got) - ref)
            worst_abs = max(worst_abs, float(err))
            worst_rel = max(worst_rel, float(err / ref))
        print(f"kd={kd}: max abs err {worst_abs:.3e}  max rel err {worst_rel:.3e}")
