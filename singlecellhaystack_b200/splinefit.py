"""Host-side significance fit: log10 p-values from the GPU-produced D_KL values.

Mirrors get_log_p_D_KL_continuous / get_model_cv_ns / get_model_cv_bs
(R/haystack_continuous.R:376-430, :433-499, :501-582).  north_star keeps this step on the host
(it is milliseconds of work on S = 100 training genes); under R it is R's own splines::ns/bs +
lm + pnorm that run.  This module is the equivalent for the Python host layer, built on
scipy.interpolate.BSpline (the oracle restates the same fit independently with a hand-written
Cox-de Boor recursion, which is what the parity tests compare against).

Random inputs (the CV fold assignment, R: `sample(rep(1:10, length.out = n))`) come from the
caller: either a numpy Generator or an explicit array.
"""
from __future__ import annotations

import math
import sys

import numpy as np
from scipy.interpolate import BSpline
from scipy.special import log_ndtr

__all__ = ["ns_design", "bs_design", "SplineModel", "get_model_cv_ns", "get_model_cv_bs",
           "get_log_p_D_KL_continuous"]


def _message(verbose, *parts):
    if verbose:
        print("".join(str(p) for p in parts), file=sys.stderr)


def ns_design(x, knots, boundary):
    """Columns of splines::ns(x, knots = knots, Boundary.knots = boundary, intercept = FALSE)."""
    x = np.asarray(x, dtype=np.float64)
    b = np.asarray(boundary, dtype=np.float64)
    aknots = np.sort(np.concatenate([np.repeat(b, 4), np.asarray(knots, dtype=np.float64)]))
    nb = aknots.size - 4
    spl = BSpline(aknots, np.eye(nb), 3, extrapolate=True)
    basis = np.empty((x.size, nb))
    left = x < b[0]
    right = x > b[1]
    mid = ~(left | right)
    if mid.any():
        basis[mid] = spl(x[mid])
    for side, pt in ((left, b[0]), (right, b[1])):
        if side.any():
            v0 = spl(pt)
            v1 = spl(pt, nu=1)
            basis[side] = v0[None, :] + (x[side] - pt)[:, None] * v1[None, :]
    const = np.stack([spl(b[0], nu=2), spl(b[1], nu=2)])[:, 1:]          # 2 x (nb-1)
    qfull, _ = np.linalg.qr(const.T, mode="complete")
    return basis[:, 1:] @ qfull[:, 2:]


def bs_design(x, knots, boundary, degree):
    """Columns of splines::bs(x, knots = knots, degree = degree, Boundary.knots = boundary,
    intercept = FALSE).  Outside the boundary knots bs() continues the boundary polynomial
    (a degree-order Taylor expansion around a pivot inside the boundary interval, which for a
    polynomial piece is the piece itself), i.e. plain extrapolation of the end pieces."""
    x = np.asarray(x, dtype=np.float64)
    b = np.asarray(boundary, dtype=np.float64)
    ord_ = degree + 1
    aknots = np.sort(np.concatenate([np.repeat(b, ord_), np.asarray(knots, dtype=np.float64)]))
    nb = aknots.size - ord_
    spl = BSpline(aknots, np.eye(nb), degree, extrapolate=True)
    return spl(x)[:, 1:]


def _type7_quantiles(x, probs):
    return np.quantile(x, probs) if len(probs) else np.zeros(0)


def _interior_knots(x, n_iknots, boundary):
    """Interior knots as ns()/bs() place them when `df` is given: quantiles (type 7) of the x lying
    inside the boundary knots at seq(0, 1, length.out = n + 2)[-c(1, n + 2)]."""
    if n_iknots <= 0:
        return np.zeros(0)
    probs = np.linspace(0.0, 1.0, n_iknots + 2)[1:-1]
    inside = (x >= boundary[0]) & (x <= boundary[1])
    return _type7_quantiles(x[inside], probs)


class SplineModel:
    """lm(y ~ ns(x, df, Boundary.knots)) or lm(y ~ bs(x, df, degree, Boundary.knots)) with the
    knots frozen at fit time (R: makepredictcall) so predict() re-evaluates the same basis."""

    def __init__(self, x, y, df, boundary, method="ns", degree=3):
        x = np.asarray(x, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        self.method, self.df, self.degree = method, int(df), int(degree)
        self.boundary = np.asarray(boundary, dtype=np.float64)
        if method == "ns":
            self.knots = _interior_knots(x, self.df - 1, self.boundary)
        else:
            self.knots = _interior_knots(x, self.df - self.degree, self.boundary)
        a = np.column_stack([np.ones(x.size), self._basis(x)])
        self.coef, *_ = np.linalg.lstsq(a, y, rcond=None)

    def _basis(self, x):
        if self.method == "ns":
            return ns_design(x, self.knots, self.boundary)
        return bs_design(x, self.knots, self.boundary, self.degree)

    def predict(self, xnew):
        xnew = np.asarray(xnew, dtype=np.float64)
        return np.column_stack([np.ones(xnew.size), self._basis(xnew)]) @ self.coef


def _folds(n, n_cv, sets, rng):
    if sets is not None:
        sets = np.asarray(sets)
        if sets.shape != (n,):
            raise ValueError("fold assignment must have one entry per training gene")
        return sets
    base = (np.arange(n) % n_cv) + 1                     # rep(1:n_cv, length.out = n)
    return base[np.random.default_rng(rng).permutation(n)]   # sample(...)


def _cv_rmsd(x, y, sets, n_cv, df, boundary, method, degree):
    rmsds = np.full(n_cv, np.nan)
    for s in range(1, n_cv + 1):
        te = sets == s
        tr = ~te
        if not te.any():
            continue
        model = SplineModel(x[tr], y[tr], df, boundary, method, degree)
        rmsds[s - 1] = math.sqrt(float(np.mean((model.predict(x[te]) - y[te]) ** 2)))
    return float(np.mean(rmsds))     # mean() propagates NA like R's


def _boundary(x):
    lo, hi = float(np.min(x)), float(np.max(x))
    r = hi - lo
    return np.array([lo - 0.01 * r, hi + 0.01 * r]), (lo, hi)       # :440-442


def get_model_cv_ns(x, y, sets=None, rng=None, names=None, verbose=False):
    """R/haystack_continuous.R:433-499."""
    _message(verbose, "### using natural splines")
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    n_cv = 10
    sets = _folds(x.size, n_cv, sets, rng)
    boundary, (lo, hi) = _boundary(x)
    dfs = list(range(1, 11))
    best_rmsd, best_df = math.inf, None
    for df in dfs:
        rmsd = _cv_rmsd(x, y, sets, n_cv, df, boundary, "ns", 3)
        if rmsd < best_rmsd:
            best_rmsd, best_df = rmsd, df
    if best_df is None:
        raise ValueError("no spline model could be cross-validated (NaN in the randomized D_KL?)")
    _message(verbose, "### best RMSD  : ", round(best_rmsd, 3))
    _message(verbose, "### best df    : ", best_df)
    model = SplineModel(x, y, best_df, boundary, "ns")
    x_seq = np.linspace(lo, hi, 1000)
    info = dict(cv_parameters=dict(dfs=dfs), cv_selected=dict(df=best_df, rmsd=best_rmsd),
                observed=dict(feature=names, x=x, y=y), fitted=dict(x=x_seq, y=model.predict(x_seq)),
                spline_method="ns")
    return dict(model=model, info=info)


def get_model_cv_bs(x, y, sets=None, rng=None, names=None, verbose=False):
    """R/haystack_continuous.R:501-582."""
    _message(verbose, "### using B-splines")
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    n_cv = 10
    sets = _folds(x.size, n_cv, sets, rng)
    boundary, (lo, hi) = _boundary(x)
    degrees, dfs = list(range(1, 6)), list(range(1, 11))
    best = (math.inf, None, None)
    for degree in degrees:
        for df in dfs:
            if df - (degree + 1) + 1 < 0:        # nIknots < 0: df too small for this degree (:524-530)
                continue
            rmsd = _cv_rmsd(x, y, sets, n_cv, df, boundary, "bs", degree)
            if rmsd < best[0]:
                best = (rmsd, degree, df)
    best_rmsd, best_degree, best_df = best
    if best_df is None:
        raise ValueError("no spline model could be cross-validated (NaN in the randomized D_KL?)")
    _message(verbose, "### best RMSD  : ", round(best_rmsd, 3))
    _message(verbose, "### best degree: ", best_degree)
    _message(verbose, "### best df    : ", best_df)
    model = SplineModel(x, y, best_df, boundary, "bs", best_degree)
    x_seq = np.linspace(lo, hi, 1000)
    info = dict(cv_parameters=dict(degrees=degrees, dfs=dfs),
                cv_selected=dict(degree=best_degree, df=best_df, rmsd=best_rmsd),
                observed=dict(feature=names, x=x, y=y), fitted=dict(x=x_seq, y=model.predict(x_seq)),
                spline_method="bs")
    return dict(model=model, info=info)


def get_log_p_D_KL_continuous(D_KL_observed, D_KL_randomized, all_coeffVar, train_coeffVar,
                              spline_method="ns", sets_mean=None, sets_sd=None, rng=None, names=None,
                              verbose=False):
    """R/haystack_continuous.R:376-430.  Returns dict(fitted = log10 p-values, info = ...)."""
    rng = np.random.default_rng(rng)
    with np.errstate(divide="ignore", invalid="ignore"):
        log_dkl = np.log(np.asarray(D_KL_randomized, dtype=np.float64))     # :379
    mean_ = log_dkl.mean(axis=1)                                             # :380
    sd_ = log_dkl.std(axis=1, ddof=1)                                        # :381
    fit = get_model_cv_ns if spline_method == "ns" else get_model_cv_bs
    lx = np.log(np.asarray(train_coeffVar, dtype=np.float64))
    lall = np.log(np.asarray(all_coeffVar, dtype=np.float64))
    _message(verbose, "### picking model for mean D_KL...")
    m_mean = fit(lx, mean_, sets=sets_mean, rng=rng, names=names, verbose=verbose)
    fitted_mean = m_mean["model"].predict(lall)                              # :403
    _message(verbose, "### picking model for stdev D_KL...")
    m_sd = fit(lx, sd_, sets=sets_sd, rng=rng, names=names, verbose=verbose)
    fitted_sd = m_sd["model"].predict(lall)                                  # :422
    with np.errstate(divide="ignore", invalid="ignore"):
        z = (np.log(np.asarray(D_KL_observed, dtype=np.float64)) - fitted_mean) / fitted_sd
        logp = log_ndtr(-z) / math.log(10)                                   # :427
    logp = np.where(fitted_sd < 0, np.nan, logp)                             # pnorm: sd < 0 -> NaN
    info = dict(features=names, mean=m_mean["info"], sd=m_sd["info"])
    return dict(fitted=logp, info=info, fitted_mean=fitted_mean, fitted_sd=fitted_sd)
