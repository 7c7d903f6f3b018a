"""numpy model of the TC32 pair kernel's arithmetic at any dimension: which error the planned d > 64
kernel would have.  Row side y_i = hi + lo (fp16 + fp16), column side fp16(y_j), exponent offsets exact, E = fp16(2^15 k),
v = hi + lo (fp16 + fp16), fp32 accumulation; optionally the lean variant's e5m2 correction terms.  Compares phi with the
FP64 oracle formula on the same particles.

    python scripts/tc32_numerics_model.py            # d = 64 (measured on B200: 2.8e-5 .. 1.1e-4) and d = 128, 256

Pure CPU; uses only numpy and the synthetic problems of svgdcpp_b200/synth.py."""
import sys

import numpy as np

sys.path.insert(0, ".")
from svgdcpp_b200 import synth


def f16(x):
    return x.astype(np.float16).astype(np.float64)


def split16(x):
    hi = f16(x)
    return hi, f16(x - hi)


def e5m2(x, truncate=False):
    """Round (to nearest even; or truncate, as taking the top byte of an fp16 value does) to the e5m2 grid: 2 mantissa bits,
    normal exponents -14 .. 15, subnormal step 2^-16, saturating at 57344."""
    x = np.asarray(x, dtype=np.float64)
    ax = np.abs(x)
    e = np.floor(np.log2(np.where(ax > 0, ax, 1.0)))
    q = np.exp2(np.maximum(e, -14.0) - 2.0)
    r = np.floor(ax / q) if truncate else np.rint(ax / q)
    return np.sign(x) * np.minimum(r * q, 57344.0)


def mixture_grad(X, means, covs):
    P = np.linalg.inv(covs)
    diff = X[:, None, :] - means[None, :, :]                 # n x C x d
    y = np.einsum("crk,nck->ncr", P, diff)
    h = -0.5 * np.einsum("ncr,ncr->nc", diff, y)
    w = np.exp(h - h.max(1, keepdims=True))
    w /= w.sum(1, keepdims=True)
    return -np.einsum("nc,ncr->nr", w, y)


def model(X, G, lean=False, one_term_v=False):
    """lean: the correction terms lo_i . y^_j and E . v_lo as e5m2 products (lo 2^10, y^ 2^-10; top byte of E, e5m2(v_lo));
    one_term_v: E . v_lo left out altogether (what the lean variant does from 32,768 particles)."""
    n, d = X.shape
    D2 = np.maximum((X ** 2).sum(1)[:, None] + (X ** 2).sum(1)[None, :] - 2 * X @ X.T, 0)
    np.fill_diagonal(D2, 0)
    a = np.log(n) / np.median(np.sqrt(D2)) ** 2
    K = np.exp(-a * D2)
    phi_ref = (K @ (G - 2 * a * X) + 2 * a * X * K.sum(1, keepdims=True)) / n
    # ---- the kernel's arithmetic ----
    Xc = X - X.mean(0)
    c = a * np.log2(np.e)
    y = np.sqrt(2 * c) * Xc
    yhi, ylo = split16(y)
    yb = f16(y)                                                # column side: one fp16 term
    S_lo = e5m2(ylo * 1024.0) @ e5m2(yb / 1024.0).T if lean else ylo @ yb.T
    S = ((yhi @ yb.T).astype(np.float32) + S_lo.astype(np.float32)).astype(np.float64)   # exact products, fp32 sums
    u = 15.0 - 0.5 * ((yhi + ylo) ** 2).sum(1)                 # three-term fp16 splits: exact to 2^-33
    w = -0.5 * (yb ** 2).sum(1)
    acc = (S + u[:, None] + w[None, :]).astype(np.float32).astype(np.float64)
    E = f16(np.exp2(np.minimum(acc, 15.0)))                    # 2^15 k, rounded to fp16
    V = (G - 2 * a * Xc).astype(np.float32).astype(np.float64)
    vhi, vlo = split16(V)
    if one_term_v:
        P_lo = np.zeros_like(E @ vhi)
    elif lean:
        P_lo = e5m2(E, truncate=True) @ e5m2(vlo)
    else:
        P_lo = E @ vlo
    Phi = ((E @ vhi).astype(np.float32) + P_lo.astype(np.float32)).astype(np.float64)
    rowsum = E.sum(1).astype(np.float32).astype(np.float64)
    phi = (Phi + 2 * a * Xc * rowsum[:, None]) * 2.0 ** -15 / n
    return np.max(np.abs(phi - phi_ref)) / np.max(np.abs(phi_ref)), a


if __name__ == "__main__":
    n = 2048
    for d in (64, 128, 256):
        x0, means, covs = synth.mvn_problem(n, d)
        X = np.ascontiguousarray(x0.T)
        err, a = model(X, mixture_grad(X, means, covs))
        print("one Gaussian  n=%d d=%3d: a = %.4f, phi max-rel error %.2e" % (n, d, a, err))
        if d == 64:
            G = mixture_grad(X, means, covs)
            print("   lean variant (e5m2 correction terms): %.2e; with v in one fp16 term: %.2e   (measured on B200 at this size: 2.4e-5 / 1.3e-4)"
                  % (model(X, G, lean=True)[0], model(X, G, lean=True, one_term_v=True)[0]))
    for d, C in ((64, 4), (256, 16)):
        x0, means, covs = synth.gmm_problem(n, d, C)
        X = np.ascontiguousarray(x0.T)
        err, a = model(X, mixture_grad(X, means, covs))
        print("%2d components n=%d d=%3d: a = %.4f, phi max-rel error %.2e" % (C, n, d, a, err))
