"""CPU: how short can the polynomial of the angular distance be?  d(y) = (2/pi) acos(y) = sqrt(1 - y) P(y) on
[0, 0.9999] (qec_math.cuh one_minus_dist: degree 7, A&S 4.4.46, 2e-8).  Near-minimax fits (Chebyshev nodes)
and their worst absolute error in d -- every degree dropped saves one FFMA2 per evaluation and vote pair
(9 -> 7 for degree 5; the cluster kernels evaluate it 3x per pair in the forward and are fma-pipe bound)."""
import numpy as np
from numpy.polynomial import chebyshev as C, polynomial as Pn

y = np.clip(np.cos(np.linspace(0, np.pi, 20001)) * 0.5 + 0.5, 0, 0.9999)
f = lambda t: (2 / np.pi) * np.arccos(t) / np.sqrt(1 - t)  # noqa: E731
yy = np.linspace(0, 0.9999, 400001)
for deg in (3, 4, 5, 6, 7):
    p = C.Chebyshev.fit(y, f(y), deg, domain=[0, 1]).convert(kind=Pn.Polynomial)
    err = np.abs(np.sqrt(1 - yy) * (p(yy) - f(yy))).max()
    print("degree %d: max |error of d| = %.2e   coefficients (low -> high) %s" % (deg, err, np.array2string(p.coef, precision=10)))
