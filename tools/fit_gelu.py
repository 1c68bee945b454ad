"""Generates the polynomial constants of the GELU / GELU' epilogues in csrc/gemm.cu (run offline; needs scipy).

forward :  gelu(h)  = max(h,0) - |h| * exp2(P(min(|h|,A)))          P ~= log2 Phi(-a), degree 7 on [0, 6]
backward:  gelu'(h) = [h>=0] - sign(h) * g * S(min(|h|,A)),  g = exp(-h^2/2),  S(a) = Phi(-a)/g - a/sqrt(2 pi)
Both are exact-erf GELU (nn.GELU default, the reference's timm Mlp) up to the fit error printed below;
the fits are weighted so the error of the *result* is uniform.  Evaluated here in float32 Horner form exactly as the
kernel does, against float64 scipy."""
import numpy as np
from scipy.special import log_ndtr, ndtr


def remez_like(x, y, deg, wfun, iters=200):
    w = np.ones_like(x)
    best = None
    for _ in range(iters):
        c = np.polynomial.polynomial.polyfit(x, y, deg, w=w * wfun)
        e = (np.polynomial.polynomial.polyval(x, c) - y) * wfun
        m = np.abs(e).max()
        if best is None or m < best[1]:
            best = (c, m)
        w = w * (1 + 1.5 * np.abs(e) / m)
        w /= w.max()
    return best


def horner32(c, a):
    c = np.asarray(c, np.float32)
    a = a.astype(np.float32)
    p = np.full_like(a, c[-1])
    for k in range(len(c) - 2, -1, -1):
        p = (p * a + c[k]).astype(np.float32)   # numpy has no fused fma; the kernel's FFMA is at least this accurate
    return p


def main():
    A = 6.0
    n = 20000
    x = np.cos(np.pi * (np.arange(n) + 0.5) / n) * A / 2 + A / 2
    # ---- forward: P(a) ~= log2 Phi(-a); error of q = q ln2 dP, result error |h| q ln2 dP: weight a*q (floored)
    y = log_ndtr(-x) / np.log(2)
    q = np.exp(log_ndtr(-x))
    cP, eP = remez_like(x, y, 7, np.maximum(x * q, 0.02) / 0.17)
    print("P coefficients (c0..c7):", ", ".join(f"{v:.9e}f" for v in cP), " weighted max err", eP)
    # ---- backward: S(a) = Phi(-a) exp(a^2/2) - a/sqrt(2pi); result error = g dS
    g = np.exp(-x * x / 2)
    S = np.exp(log_ndtr(-x) + x * x / 2) - x / np.sqrt(2 * np.pi)
    for deg in (6, 7, 8, 9):
        cS, eS = remez_like(x, S, deg, g)
        print(f"S deg {deg}: weighted max err {eS:.3e}")
    cS, eS = remez_like(x, S, 8, g)
    print("S coefficients (c0..c8):", ", ".join(f"{v:.9e}f" for v in cS))

    # ---- float32 check over a dense grid incl. far tails
    h = np.concatenate([np.linspace(-12, 12, 2_000_001), np.array([-65504., 65504., -100., 100., 0., -0.])]).astype(np.float32)
    a = np.minimum(np.abs(h), np.float32(A))
    h64 = h.astype(np.float64)
    gelu_ref = h64 * ndtr(h64)
    qf = np.exp2(horner32(cP, a))
    gelu = np.maximum(h, 0) - np.abs(h) * qf
    err = np.abs(gelu - gelu_ref)
    print("forward: max abs err", err.max(), "max rel-to-max(|ref|,1e-3)", (err / np.maximum(np.abs(gelu_ref), 1e-3)).max())
    dref = ndtr(h64) + h64 * np.exp(-h64 * h64 / 2) / np.sqrt(2 * np.pi)
    ap = (np.abs(h) * np.float32(np.sqrt(0.5 * np.log2(np.e)))).astype(np.float32)   # a' : g = exp2(-a'^2)
    gf = np.exp2(-(ap * ap).astype(np.float32))
    Sf = horner32(cS, a)
    d = np.where(h >= 0, 1.0, 0.0) - np.sign(h) * gf * Sf
    d = np.where(h == 0, 0.5, d)
    errd = np.abs(d - dref)
    print("backward: max abs err", errd.max())
    # far-tail sanity: P must stay very negative at the clamp, S*g must vanish
    print("P(A) =", horner32(cP, np.array([A]))[0], " g(A)*S(A) =", (np.exp(-A * A / 2) * horner32(cS, np.array([A])))[0])


if __name__ == "__main__":
    main()
