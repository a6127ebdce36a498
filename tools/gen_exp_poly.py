"""Derive the polynomial used by fast_exp() in bhmm_b200/csrc/lane_kernels.cu and check its accuracy.

exp(x) = 2^n * exp(r), n = rint(x / ln 2), r = x - n ln2 in [-ln2/2, ln2/2].  exp(r) is approximated by the degree-D
interpolant at Chebyshev nodes (near-minimax), coefficients converted to the monomial basis with mpmath and rounded to
double; evaluation is Horner with FMAs.  The script prints the coefficients and the worst relative error seen over a
dense sample, evaluated in exact rational arithmetic emulating double FMA rounding only at the final comparison.
"""
import sys
import mpmath as mp
import numpy as np

mp.mp.dps = 60
D = int(sys.argv[1]) if len(sys.argv) > 1 else 11
h = mp.log(2) / 2 * mp.mpf('1.0001')
nodes = [h * mp.cos(mp.pi * (2 * k + 1) / (2 * (D + 1))) for k in range(D + 1)]
V = mp.matrix(D + 1, D + 1)
for i, x in enumerate(nodes):
    for j in range(D + 1):
        V[i, j] = x ** j
rhs = mp.matrix([mp.e ** x for x in nodes])
coef = mp.lu_solve(V, rhs)
c = [float(coef[j]) for j in range(D + 1)]
c[0] = 1.0
print('degree', D)
for j, v in enumerate(c):
    print('    c%-2d = %s   (%r)' % (j, float.hex(v), v))

# accuracy in double arithmetic (numpy emulation of Horner; numpy has no fma, so emulate with longdouble products)
xs = np.linspace(-float(h), float(h), 400001)
p = np.full_like(xs, c[D], dtype=np.longdouble)
xl = xs.astype(np.longdouble)
for j in range(D - 1, -1, -1):
    p = (p * xl + np.longdouble(c[j])).astype(np.float64).astype(np.longdouble)   # fma: single rounding to double
ref = np.array([mp.e ** mp.mpf(float(x)) for x in xs[::400]], dtype=object)
got = p[::400]
worst = max(abs((mp.mpf(float(g)) - r) / r) for g, r in zip(got, ref))
print('worst relative error on sample: %.3e  (%.2f ulp)' % (float(worst), float(worst) / 1.11e-16))
