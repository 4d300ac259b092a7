"""Near-minimax coefficients for exp(d) = 1 + d*(c1 + c2 d + c3 d^2 + c4 d^3 + c5 d^4) on |d| <= ln2/64
(used by cggp_b200/csrc/kmath.cuh::fast_exp).  Remez exchange on g(d) = (exp(d)-1)/d in mpmath, then the
double-rounded coefficients are verified against mpmath.exp on a dense grid."""
import mpmath as mp

mp.mp.dps = 60
a = mp.log(2) / 64
deg = 4  # degree of g's approximation


def g(d):
    return mp.expm1(d) / d if d != 0 else mp.mpf(1)


# Remez on [-a, a]
nodes = [a * mp.cos(mp.pi * (2 * k + 1) / (2 * (deg + 2))) for k in range(deg + 2)][::-1]
for it in range(30):
    A = mp.matrix(deg + 2, deg + 2)
    b = mp.matrix(deg + 2, 1)
    for i, x in enumerate(nodes):
        for j in range(deg + 1):
            A[i, j] = x ** j
        A[i, deg + 1] = (-1) ** i
        b[i] = g(x)
    sol = mp.lu_solve(A, b)
    coef = [sol[j] for j in range(deg + 1)]
    err = lambda x: sum(c * x ** j for j, c in enumerate(coef)) - g(x)  # noqa: E731
    # new extrema: scan
    grid = [-a + 2 * a * mp.mpf(k) / 4000 for k in range(4001)]
    vals = [err(x) for x in grid]
    ext = [grid[0]]
    for k in range(1, 4000):
        if (vals[k] - vals[k - 1]) * (vals[k + 1] - vals[k]) <= 0:
            ext.append(grid[k])
    ext.append(grid[-1])
    if len(ext) != deg + 2:
        break
    if max(abs(e1 - e0) for e0, e1 in zip(nodes, ext)) < a * 1e-6:
        nodes = ext
        break
    nodes = ext
dbl = [float(c) for c in coef]
print("c1..c5 =", ["%.17e" % c for c in dbl])
# verify the double coefficients in the exact Horner order the kernel uses
worst = 0
for k in range(20001):
    d = -a + 2 * a * mp.mpf(k) / 20000
    q = mp.mpf(dbl[4])
    for c in (dbl[3], dbl[2], dbl[1], dbl[0]):
        q = q * d + mp.mpf(c)
    p = q * d + 1
    worst = max(worst, abs(p / mp.exp(d) - 1))
print("max rel error of the polynomial (exact arithmetic, double coefficients): %.3e" % float(worst))
