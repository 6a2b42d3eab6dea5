import numpy as np
from scipy.special import erfc, erf
# q(a) = log2(erfc(a/sqrt2)), a in [0, A]; gelu(x) = max(x,0) - 0.5*|x|*2^q(|x|)
A = 6.0
def fit(deg, A=A, n=4000):
    # Chebyshev nodes, weighted least squares; weight ~ a*erfc (error sensitivity of gelu)
    k = np.arange(n); a = 0.5*A*(1 - np.cos(np.pi*(k+0.5)/n))
    y = np.log2(erfc(a/np.sqrt(2)))
    wgt = a*erfc(a/np.sqrt(2)) + 1e-4
    # polynomial without constant term: q = a*(c1 + c2 a + ...)
    V = np.stack([a**i for i in range(1, deg+1)], 1)
    for it in range(30):
        c, *_ = np.linalg.lstsq(V*wgt[:,None], y*wgt, rcond=None)
        r = (V@c - y)
        err = 0.5*a*erfc(a/np.sqrt(2))*np.abs(2**r - 1)
        # reweight towards minimax
        wgt = wgt*(1 + 0.5*err/err.max())
    return c
for deg in (5,6,7,8):
    c = fit(deg)
    x = np.linspace(-8, 8, 400001)
    a = np.abs(x).astype(np.float32)
    # float32 emulation of Horner
    q = np.float32(c[-1])
    for ci in c[-2::-1]:
        q = (q*a + np.float32(ci)).astype(np.float32)
    q = (q*a).astype(np.float32)
    E = np.exp2(q.astype(np.float64)).astype(np.float32)
    g = (np.maximum(x,0).astype(np.float32) - np.float32(0.5)*a*E).astype(np.float32)
    ref = 0.5*x*(1+erf(x/np.sqrt(2)))
    e = np.abs(g-ref)
    print(deg, "max abs err", e.max(), "at", x[e.argmax()], "max rel err (|ref|>1e-3)", (e/np.maximum(np.abs(ref),1e-3)).max())
    print("  coeffs", [float(np.float32(v)) for v in c])
