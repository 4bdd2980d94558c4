"""Minimax polynomials q(x) ~ (exp(x) - 1)/x on [-h, h], h = ln2/(2N) (Remez exchange in 60-digit arithmetic),
for the table exponential of trans_core.cuh: exp(x) = 1 + x q(x) to within x * max|q - (e^x-1)/x|.
Prints the coefficients as C hex-float literals and the error bounds.  python tools/gen_expm1_minimax.py N degree"""
import sys

import mpmath as mp

mp.mp.dps = 60


def target(x):
    return mp.expm1(x) / x if x != 0 else mp.mpf(1)


def remez(h, deg, iters=30):
    n = deg + 2
    xs = [h * mp.cos(mp.pi * (n - 1 - i) / (n - 1)) for i in range(n)]  # Chebyshev extrema
    for _ in range(iters):
        A = mp.matrix(n, n)
        b = mp.matrix(n, 1)
        for i, x in enumerate(xs):
            for j in range(deg + 1):
                A[i, j] = x ** j
            A[i, deg + 1] = (-1) ** i
            b[i] = target(x)
        sol = mp.lu_solve(A, b)
        coef = [sol[j] for j in range(deg + 1)]
        err = lambda x: sum(c * x ** j for j, c in enumerate(coef)) - target(x)
        # new extrema: scan a fine grid between the roots
        grid = [(-h + 2 * h * k / 4000) for k in range(4001)]
        vals = [err(x) for x in grid]
        ext = []
        k0 = 0
        sign = mp.sign(vals[0]) or 1
        best = (abs(vals[0]), grid[0])
        for k in range(1, len(grid)):
            s = mp.sign(vals[k]) or sign
            if s != sign:
                ext.append(best[1])
                sign = s
                best = (abs(vals[k]), grid[k])
            elif abs(vals[k]) > best[0]:
                best = (abs(vals[k]), grid[k])
        ext.append(best[1])
        if len(ext) != n:
            break
        xs = ext
    e = max(abs(v) for v in vals)
    return coef, e


if __name__ == "__main__":
    N = int(sys.argv[1])
    deg = int(sys.argv[2])
    h = mp.log(2) / (2 * N)
    coef, e = remez(h, deg)
    print(f"// N = {N}: |x| <= ln2/{2 * N} = {float(h):.3e}; q degree {deg}; max |q - (e^x-1)/x| = {float(e):.2e}; "
          f"max error of exp(x) = {float(e * h):.2e}")
    for c in coef:
        print("  %s,  // %.17g" % (float(c).hex(), float(c)))
    tay = [1 / mp.factorial(j + 1) for j in range(deg + 1)]
    te = max(abs(sum(c * x ** j for j, c in enumerate(tay)) - target(x)) for x in [h, -h])
    print(f"// Taylor of the same degree: max error of exp(x) = {float(te * h):.2e}")
