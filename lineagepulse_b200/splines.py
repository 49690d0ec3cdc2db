"""Natural cubic spline basis, the host-side equivalent of ``splines::ns(x, df, intercept=TRUE)``.

The reference builds its spline design matrices in R (initialiseMuModel.R:111-112,151-152;
initialiseDispModel.R:148-149; initialiseImpulseParameters.R:34-35); ``splines`` is an R base
package that is not part of /root/reference, so this restates the published construction
(SURVEY.md App. D): cubic B-splines on the knot vector ``c(rep(range(x),4), quantile knots)``,
projected on the null space of the two boundary second-derivative constraints via the full
Householder QR of ``t(const)`` with the first two columns dropped.
"""
from __future__ import annotations

import numpy as np


def _bspline_basis_all(knots: np.ndarray, x: np.ndarray, order: int, deriv: int = 0) -> np.ndarray:
    """Design matrix of all B-splines of ``order`` on ``knots`` at ``x`` (de Boor recursion).

    The right end is closed: x == max(knots) is evaluated in the last non-empty interval,
    as ``splines::splineDesign`` does.
    """
    t = np.asarray(knots, dtype=np.float64)
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    nb = len(t) - order
    # interval index: t[mu] <= x < t[mu+1], clamped to the last non-empty interval
    last = np.max(np.nonzero(t[:-1] < t[1:])[0])
    mu = np.searchsorted(t, x, side="right") - 1
    mu = np.clip(mu, order - 1, last)
    # order-1 (piecewise constant) basis
    B = np.zeros((len(x), len(t) - 1))
    B[np.arange(len(x)), mu] = 1.0
    for k in range(2, order + 1):
        n_k = len(t) - k
        Bn = np.zeros((len(x), n_k))
        lower_than_deriv = k > order - deriv  # this level is differentiated instead of evaluated
        for i in range(n_k):
            d1 = t[i + k - 1] - t[i]
            d2 = t[i + k] - t[i + 1]
            if lower_than_deriv:
                a = (k - 1) / d1 * B[:, i] if d1 > 0 else 0.0
                b = (k - 1) / d2 * B[:, i + 1] if d2 > 0 else 0.0
                Bn[:, i] = a - b
            else:
                a = (x - t[i]) / d1 * B[:, i] if d1 > 0 else 0.0
                b = (t[i + k] - x) / d2 * B[:, i + 1] if d2 > 0 else 0.0
                Bn[:, i] = a + b
        B = Bn
    return B[:, :nb]


def ns(x, df: int, intercept: bool = True) -> np.ndarray:
    """``splines::ns(x, df=df, intercept=intercept)`` as an N x df matrix."""
    x = np.asarray(x, dtype=np.float64)
    n_iknots = df - 1 - int(bool(intercept))
    if n_iknots < 0:
        raise ValueError("df too small for a natural spline")
    bk = np.array([x.min(), x.max()])
    if n_iknots > 0:
        probs = np.linspace(0.0, 1.0, n_iknots + 2)[1:-1]
        iknots = np.quantile(x, probs)  # type-7 quantiles, R's default
    else:
        iknots = np.array([])
    aknots = np.sort(np.concatenate([np.repeat(bk, 4), iknots]))
    basis = _bspline_basis_all(aknots, x, 4)
    const = _bspline_basis_all(aknots, bk, 4, deriv=2)
    if not intercept:
        basis = basis[:, 1:]
        const = const[:, 1:]
    q, _ = np.linalg.qr(const.T, mode="complete")
    return np.asfortranarray((basis @ q)[:, 2:])
