"""fp64 CPU restatement of singleCellHaystack's continuous high-D scoring path.

TEST INFRASTRUCTURE ONLY.  This module is the parity oracle: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs may import
it.  Nothing under `singlecellhaystack_b200/` imports it, and the product path never falls back
to it.

PARITY STATUS: **pinned** against the only numbers the reference holds.  R is not installed in the build container
nor on the GPU boxes, and the reference's own tests assert no numbers (every numeric expectation is commented out:
tests/testthat/test-haystack_continuous.R:16-17,30-31; tests/testthat/test-haystack_sparse.R:20-21,36-37), so the pin
is the rendered vignette, docs/articles/a01_toy_example.html:129 and :168 (set.seed(1234); haystack(dat.tsne,
dat.expression)): with the parts of base R that call consumes restated in oracle/r_base.c (set.seed scrambling,
Mersenne-Twister, R_unif_index, sample.int, sample(prob=), Hartigan-Wong kmeans) and drawn in the reference's order,
this module reproduces "best df 3 / 5", the ten printed genes in order, their D_KL to all 7 printed digits and their
log.p to the 5 printed decimals (tests/test_r_parity_cpu.py; golden copy tests/golden/toy_r1234.npz).  Also pinned
(tests/test_oracle_cpu.py):
  * decoding of the reference's fixture data/toy.rda (dims, means, sds, sparsity),
  * the two exact-value reference tests (tests/testthat/test-extract_row.R:8-10,18-20),
  * the identity log.p.adj - log.p.vals = log10(500) from the same table (R/haystack_continuous.R:301-302),
  * dense-vs-sparse agreement and algebraic invariants of the restated formulas.

Every function cites the reference lines it follows (paths relative to /root/reference).
All arithmetic is IEEE double.  Random inputs (grid coordinates, permutation indices, CV
folds) are ARGUMENTS, never drawn here, exactly so that oracle and device path see identical
inputs (the reference accepts a user-supplied grid: R/haystack_continuous.R:30,37-46,157).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

PSEUDO = 1e-300  # R/haystack_continuous.R:184


# --------------------------------------------------------------------------------------------
# sparse row access (R/sparse.R)
# --------------------------------------------------------------------------------------------
@dataclass
class RowSparse:
    """dgRMatrix look-alike: slots p (int, M+1), j (0-based cell ids), x (fp64), shape (M, N)."""
    p: np.ndarray
    j: np.ndarray
    x: np.ndarray
    shape: tuple

    @staticmethod
    def from_dense(mat: np.ndarray) -> "RowSparse":
        """as(mat, "RsparseMatrix"): row-major walk keeps ascending j per row; explicit zeros
        are dropped, as Matrix does when coercing a dense matrix."""
        mat = np.asarray(mat, dtype=np.float64)
        nz = mat != 0
        p = np.concatenate([[0], np.cumsum(nz.sum(axis=1))]).astype(np.int64)
        rows, cols = np.nonzero(nz)
        return RowSparse(p=p, j=cols.astype(np.int32), x=mat[rows, cols], shape=mat.shape)

    def to_dense(self) -> np.ndarray:
        out = np.zeros(self.shape, dtype=np.float64)
        for i in range(self.shape[0]):
            out[i, self.j[self.p[i]:self.p[i + 1]]] = self.x[self.p[i]:self.p[i + 1]]
        return out


def extract_row_dgRMatrix(m: RowSparse, i: int) -> np.ndarray:
    """R/sparse.R:28-37 (1-based row i) -> dense numeric row."""
    r = np.zeros(m.shape[1], dtype=np.float64)
    lo, hi = int(m.p[i - 1]), int(m.p[i])
    r[m.j[lo:hi]] = m.x[lo:hi]
    return r


def extract_row_lgRMatrix(m: RowSparse, i: int) -> np.ndarray:
    """R/sparse.R:8-19 -> dense logical row."""
    return extract_row_dgRMatrix(m, i) != 0


def extract_row_dgRMatrix_as_sparse(m: RowSparse, i: int):
    """R/sparse.R:39-49 (1-based row i) -> (ind 1-based, val)."""
    lo, hi = int(m.p[i - 1]), int(m.p[i])
    if lo < hi:
        return m.j[lo:hi].astype(np.int64) + 1, m.x[lo:hi]
    return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.float64)


# --------------------------------------------------------------------------------------------
# row statistics and scaling
# --------------------------------------------------------------------------------------------
def row_mean_sd(expression):
    """R/haystack_continuous.R:57-83: Matrix::rowMeans and sd (n-1 denominator, zeros
    included) per gene.  Returns (mean, sd) WITHOUT the +1e-300 (added by the caller at :82-83)."""
    if isinstance(expression, RowSparse):
        m_, n_ = expression.shape
        mean = np.zeros(m_)
        sd = np.zeros(m_)
        for i in range(m_):
            v = expression.x[expression.p[i]:expression.p[i + 1]]
            mu = v.sum() / n_
            mean[i] = mu
            # two-pass sample variance with the implicit zeros counted
            ss = ((v - mu) ** 2).sum() + (n_ - v.size) * mu * mu
            sd[i] = math.sqrt(ss / (n_ - 1)) if n_ > 1 else float("nan")
        return mean, sd
    e = np.asarray(expression, dtype=np.float64)
    return e.mean(axis=1), e.std(axis=1, ddof=1)


def scale_coords(x: np.ndarray):
    """base::scale(x) as used at R/haystack_continuous.R:139-145: centre by column mean, divide
    by column sd (n-1).  Returns (scaled, center, scale)."""
    x = np.asarray(x, dtype=np.float64)
    center = x.mean(axis=0)
    xc = x - center
    sc = np.sqrt((xc ** 2).sum(axis=0) / (x.shape[0] - 1))
    return xc / sc, center, sc


# --------------------------------------------------------------------------------------------
# density stage (R/haystack_highD.R:7-24, R/haystack_continuous.R:168-186)
# --------------------------------------------------------------------------------------------
def get_euclidean_distance(x, y):
    """R/haystack_highD.R:7-9."""
    return math.sqrt(float(((x - y) ** 2).sum()))


def get_dist_two_sets(set1: np.ndarray, set2: np.ndarray) -> np.ndarray:
    """R/haystack_highD.R:17-24: one column per row of set2."""
    set1 = np.asarray(set1, dtype=np.float64)
    set2 = np.asarray(set2, dtype=np.float64)
    out = np.empty((set1.shape[0], set2.shape[0]), dtype=np.float64)
    for s2 in range(set2.shape[0]):
        diff = set1 - set2[s2]
        out[:, s2] = np.sqrt((diff * diff).sum(axis=1))
    return out


def density_stage(x_scaled: np.ndarray, grid_coord: np.ndarray, weights_advanced_q=None):
    """R/haystack_continuous.R:168-186.

    Returns dict(dist, nearest (0-based argmin per cell, first minimum), bandwidth, density, Q).
    """
    dist = get_dist_two_sets(x_scaled, grid_coord)                       # :168
    rowmin = dist.min(axis=1)
    bandwidth = float(np.median(rowmin))                                  # :174
    dn = dist / bandwidth                                                 # :175
    dens = np.exp(-dn * dn / 2)                                           # :176-177
    if weights_advanced_q is None:
        q = dens.sum(axis=0)                                              # :180
    else:
        w = np.asarray(weights_advanced_q, dtype=np.float64)
        q = (dens * w[:, None]).sum(axis=0)                               # :182 (recycled down rows)
    q = q + PSEUDO                                                        # :185
    q = q / q.sum()                                                       # :186
    return dict(dist=dist, nearest=dist.argmin(axis=1).astype(np.int32), bandwidth=bandwidth,
                density=dens, Q=q)


# --------------------------------------------------------------------------------------------
# D_KL (R/haystack_continuous.R:350-363 and :586-602)
# --------------------------------------------------------------------------------------------
def _kl_from_p(p_raw: np.ndarray, q: np.ndarray, pseudo: float) -> float:
    with np.errstate(divide="ignore", invalid="ignore"):
        p = p_raw + pseudo                    # :358 / :597
        p = p / p.sum()                       # :359 / :598
        return float((p * np.log(p / q)).sum())   # :360 / :599


def get_D_KL_continuous_highD(weights, density, q, pseudo=0.0) -> float:
    """R/haystack_continuous.R:350-363: P = colSums(density * weights) (weights recycled down
    rows => row c scaled by weights[c])."""
    p_raw = (density * np.asarray(weights, dtype=np.float64)[:, None]).sum(axis=0)   # :357
    return _kl_from_p(p_raw, q, pseudo)


def get_D_KL_continuous_highD_SPARSE(ind1, val, density, q, pseudo=0.0) -> float:
    """R/haystack_continuous.R:586-602: ind1 is 1-based; P = colSums(density[ind,] * val)."""
    rows = density[np.asarray(ind1, dtype=np.int64) - 1, :]                            # :596 gather
    p_raw = (rows * np.asarray(val, dtype=np.float64)[:, None]).sum(axis=0)
    return _kl_from_p(p_raw, q, pseudo)


def dkl_observed(expression, density, q, pseudo=PSEUDO) -> np.ndarray:
    """Hot loop #1, R/haystack_continuous.R:199-213 (per-gene loop kept on purpose)."""
    m_ = expression.shape[0]
    out = np.zeros(m_)
    if isinstance(expression, RowSparse):
        for i in range(1, m_ + 1):
            ind, val = extract_row_dgRMatrix_as_sparse(expression, i)                  # :209
            out[i - 1] = get_D_KL_continuous_highD_SPARSE(ind, val, density, q, pseudo)
    else:
        e = np.asarray(expression, dtype=np.float64)
        for i in range(m_):
            out[i] = get_D_KL_continuous_highD(e[i, :], density, q, pseudo)            # :204
    return out


# --------------------------------------------------------------------------------------------
# gene selection for randomization (R/haystack_continuous.R:222-236)
# --------------------------------------------------------------------------------------------
def r_seq_len_out(frm: float, to: float, length_out: int) -> np.ndarray:
    """seq(from, to, length.out=n) of base R (seq.default): ends exact, interior
    from + i*((to-from)/(n-1))."""
    n = int(length_out)
    if n <= 0:
        return np.zeros(0)
    if n == 1:
        return np.array([float(frm)])
    if n == 2:
        return np.array([float(frm), float(to)])
    if frm == to:
        return np.full(n, float(frm))
    by = (float(to) - float(frm)) / (n - 1)
    mid = float(frm) + np.arange(1, n - 1, dtype=np.float64) * by
    return np.concatenate([[float(frm)], mid, [float(to)]])


def select_genes_to_randomize(coeff_var: np.ndarray, n_genes_to_randomize: int,
                              method: str = "heavytails") -> np.ndarray:
    """R/haystack_continuous.R:227-236.  Returns 0-based gene indices (R's are 1-based).
    `order()` is stable for ties; NaN sorts last (na.last=TRUE)."""
    cv = np.asarray(coeff_var, dtype=np.float64)
    m_ = cv.size
    o = np.argsort(cv, kind="stable")          # NaN last in numpy as in R
    s = int(n_genes_to_randomize)
    if method == "uniform":
        pos = np.floor(r_seq_len_out(1, m_, s)).astype(np.int64)                       # :231
    elif method == "heavytails":
        mid = r_seq_len_out(10, m_ - 9, s - 18)
        mid = np.trunc(mid).astype(np.int64)                                           # as.integer
        pos = np.concatenate([np.arange(1, 10), mid, m_ - np.arange(8, -1, -1)])       # :233-235
    else:
        raise ValueError("selection.method.genes.to.randomize must be 'uniform' or 'heavytails'")
    return o[pos - 1]


# --------------------------------------------------------------------------------------------
# randomization (R/haystack_continuous.R:245-284); the random draws are INPUTS
# --------------------------------------------------------------------------------------------
def dkl_randomized_dense(expression_rows: np.ndarray, perms1: np.ndarray, density, q,
                         pseudo=PSEUDO) -> np.ndarray:
    """:250-265.  expression_rows: (S, N) rows of the genes to randomize;
    perms1: (S, R, N) 1-based permutations as sample.int(N) would return;
    sample(x=v) == v[sample.int(length(v))].  Returns (S, R)."""
    s_, r_ = perms1.shape[0], perms1.shape[1]
    out = np.full((s_, r_), np.nan)
    for i in range(s_):
        v = expression_rows[i]
        for r in range(r_):
            out[i, r] = get_D_KL_continuous_highD(v[perms1[i, r] - 1], density, q, pseudo)   # :258-261
    return out


def dkl_randomized_sparse(vals: list, inds1: list, density, q, pseudo=PSEUDO) -> np.ndarray:
    """:266-283.  vals[i]: nnz_i values in stored order; inds1[i]: (R, nnz_i) 1-based distinct
    cell ids as sample(x=count.cells, size=nnz_i) would return.  Returns (S, R)."""
    s_ = len(vals)
    r_ = inds1[0].shape[0] if s_ else 0
    out = np.full((s_, r_), np.nan)
    for i in range(s_):
        for r in range(r_):
            out[i, r] = get_D_KL_continuous_highD_SPARSE(inds1[i][r], vals[i], density, q, pseudo)  # :275-279
    return out


# --------------------------------------------------------------------------------------------
# significance fit (R/haystack_continuous.R:376-499, splines::ns restated)
# --------------------------------------------------------------------------------------------
def _bspline_design(knots: np.ndarray, x: np.ndarray, order: int = 4, deriv: int = 0) -> np.ndarray:
    """splines::splineDesign(knots, x, ord, derivs) by the Cox-de Boor recursion
        N_{i,1}(x) = [t_i <= x < t_{i+1}]
        N_{i,p}    = (x-t_i)/(t_{i+p-1}-t_i) N_{i,p-1} + (t_{i+p}-x)/(t_{i+p}-t_{i+1}) N_{i+1,p-1}
        N^{(k)}_{i,p} = (p-1) [ N^{(k-1)}_{i,p-1}/(t_{i+p-1}-t_i) - N^{(k-1)}_{i+1,p-1}/(t_{i+p}-t_{i+1}) ]
    (0/0 := 0).  x equal to the last knot belongs to the last non-empty interval, which is how
    splineDesign treats the right boundary.  Returns (len(x), len(knots)-order)."""
    t = np.asarray(knots, dtype=np.float64)
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))

    def rec(p: int, k: int) -> np.ndarray:
        nb = t.size - p
        if p == 1:
            out = np.zeros((x.size, nb))
            if k > 0:
                return out
            last = max(i for i in range(nb) if t[i] < t[i + 1])
            for i in range(nb):
                if t[i] < t[i + 1]:
                    hit = (x >= t[i]) & (x < t[i + 1])
                    if i == last:
                        hit |= x == t[i + 1]
                    out[hit, i] = 1.0
            return out
        lower = rec(p - 1, k - 1 if k > 0 else 0)
        out = np.zeros((x.size, nb))
        for i in range(nb):
            d1 = t[i + p - 1] - t[i]
            d2 = t[i + p] - t[i + 1]
            if k > 0:
                a = lower[:, i] / d1 if d1 > 0 else 0.0
                b = lower[:, i + 1] / d2 if d2 > 0 else 0.0
                out[:, i] = (p - 1) * (a - b)
            else:
                a = (x - t[i]) / d1 * lower[:, i] if d1 > 0 else 0.0
                b = (t[i + p] - x) / d2 * lower[:, i + 1] if d2 > 0 else 0.0
                out[:, i] = a + b
        return out

    return rec(order, deriv)


@dataclass
class NsBasis:
    """The pieces `predict` needs to re-evaluate an ns() term (R: makepredictcall.ns)."""
    knots: np.ndarray
    boundary: np.ndarray
    aknots: np.ndarray
    qmat: np.ndarray   # orthogonal basis of the null space of the boundary constraints


def _householder_null(const_t: np.ndarray) -> np.ndarray:
    """Columns 3.. of Q from qr(t(const)) -- what `qr.qty(qr.const, t(basis))[-(1:2)]` projects
    on (splines::ns).  Any orthonormal basis of the null space spans the same function space and
    gives the same least-squares fit; numpy's Householder QR (LAPACK) is the same construction."""
    qfull, _ = np.linalg.qr(const_t, mode="complete")
    return qfull[:, 2:]


def ns_fit_basis(x: np.ndarray, df: int, boundary_knots) -> NsBasis:
    """splines::ns(x, df=df, Boundary.knots=b) with intercept=FALSE: interior knots at the
    quantiles (type 7) of the x inside the boundary, df-1 of them."""
    x = np.asarray(x, dtype=np.float64)
    b = np.asarray(boundary_knots, dtype=np.float64)
    n_iknots = df - 1
    inside = (x >= b[0]) & (x <= b[1])
    if n_iknots > 0:
        probs = r_seq_len_out(0, 1, n_iknots + 2)[1:-1]
        knots = np.quantile(x[inside], probs)          # R default quantile type 7 == numpy 'linear'
    else:
        knots = np.zeros(0)
    aknots = np.sort(np.concatenate([np.repeat(b, 4), knots]))
    const = _bspline_design(aknots, b, 4, deriv=2)[:, 1:]           # drop first column (no intercept)
    qmat = _householder_null(const.T)
    return NsBasis(knots=knots, boundary=b, aknots=aknots, qmat=qmat)


def ns_eval(nb: NsBasis, x: np.ndarray) -> np.ndarray:
    """Evaluate the ns basis at x, linear extrapolation beyond the boundary knots as
    splines::ns does (value + slope at the boundary)."""
    x = np.asarray(x, dtype=np.float64)
    b = nb.boundary
    basis = np.zeros((x.size, nb.aknots.size - 4))
    left = x < b[0]
    right = x > b[1]
    mid = ~(left | right)
    if mid.any():
        basis[mid] = _bspline_design(nb.aknots, x[mid], 4, 0)
    if left.any():
        v0 = _bspline_design(nb.aknots, b[:1], 4, 0)
        v1 = _bspline_design(nb.aknots, b[:1], 4, 1)
        basis[left] = v0 + (x[left] - b[0])[:, None] * v1
    if right.any():
        v0 = _bspline_design(nb.aknots, b[1:], 4, 0)
        v1 = _bspline_design(nb.aknots, b[1:], 4, 1)
        basis[right] = v0 + (x[right] - b[1])[:, None] * v1
    return basis[:, 1:] @ nb.qmat


def _lm_fit_predict(x_train, y_train, x_pred, df, boundary):
    """lm(y.train ~ ns(x.train, df, Boundary.knots)) then predict(): R/haystack_continuous.R:457-459.
    Least squares with intercept; minimum-norm solution if rank deficient."""
    nb = ns_fit_basis(x_train, df, boundary)
    a = np.column_stack([np.ones(x_train.size), ns_eval(nb, x_train)])
    coef, *_ = np.linalg.lstsq(a, y_train, rcond=None)
    ap = np.column_stack([np.ones(np.size(x_pred)), ns_eval(nb, x_pred)])
    return ap @ coef


def get_model_cv_ns(x: np.ndarray, y: np.ndarray, sets: np.ndarray):
    """R/haystack_continuous.R:433-499.  `sets` is the fold assignment (values 1..10) that
    `sample(rep(1:10, length.out=length(x)))` would have produced (:437) -- an input here.
    Returns dict(best_df, best_rmsd, boundary, predict=callable)."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    n_cv = 10
    rng_x = (x.min(), x.max())
    r_x = rng_x[1] - rng_x[0]
    boundary = np.array([rng_x[0] - 0.01 * r_x, rng_x[1] + 0.01 * r_x])        # :440-442
    best_rmsd, best_df = math.inf, None
    for df in range(1, 11):                                                     # :444-448
        rmsds = np.full(n_cv, np.nan)
        for s in range(1, n_cv + 1):
            tr = sets != s
            te = sets == s
            pred = _lm_fit_predict(x[tr], y[tr], x[te], df, boundary)           # :457-459
            rmsds[s - 1] = math.sqrt(np.mean((pred - y[te]) ** 2)) if te.any() else float("nan")
        rmsd = rmsds.mean()                                                     # :462
        if rmsd < best_rmsd:                                                    # :467
            best_rmsd, best_df = rmsd, df

    def predict(xnew):
        return _lm_fit_predict(x, y, np.asarray(xnew, dtype=np.float64), best_df, boundary)   # :477

    return dict(best_df=best_df, best_rmsd=best_rmsd, boundary=boundary, predict=predict)


def _norm_logsf(z: np.ndarray) -> np.ndarray:
    """log of the upper normal tail, pnorm(lower.tail=FALSE, log.p=TRUE) on a standardised z."""
    from scipy.special import log_ndtr
    return log_ndtr(-z)


def get_log_p_D_KL_continuous(dkl_observed_, dkl_randomized, all_coeff_var, train_coeff_var,
                              sets_mean: np.ndarray, sets_sd: np.ndarray):
    """R/haystack_continuous.R:376-430 with spline.method="ns".  Returns dict(log_p (log10),
    mean_model, sd_model, fitted_mean, fitted_sd)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        log_dkl = np.log(dkl_randomized)                                        # :379
    mean_ = log_dkl.mean(axis=1)                                                # :380
    sd_ = log_dkl.std(axis=1, ddof=1)                                           # :381
    lx = np.log(train_coeff_var)
    m_mean = get_model_cv_ns(lx, mean_, sets_mean)                              # :391
    fitted_mean = m_mean["predict"](np.log(all_coeff_var))                      # :403
    m_sd = get_model_cv_ns(lx, sd_, sets_sd)                                    # :413
    fitted_sd = m_sd["predict"](np.log(all_coeff_var))                          # :422
    with np.errstate(divide="ignore", invalid="ignore"):
        z = (np.log(dkl_observed_) - fitted_mean) / fitted_sd
        logp = _norm_logsf(z) / math.log(10)                                    # :427
    logp = np.where(fitted_sd < 0, np.nan, logp)
    return dict(log_p=logp, mean_model=m_mean, sd_model=m_sd, fitted_mean=fitted_mean,
                fitted_sd=fitted_sd, train_mean=mean_, train_sd=sd_)


# --------------------------------------------------------------------------------------------
# the step after the path: prepare_clustering (R/haystack_continuous.R:604-666)
# --------------------------------------------------------------------------------------------
def prepare_clustering(x, expression, grid_coordinates, scale=True) -> np.ndarray:
    """Gene x grid densities, rows rescaled to sum to 1.  grid_coordinates are in the unscaled space (:612)."""
    x = np.asarray(x, dtype=np.float64)
    grid = np.asarray(grid_coordinates, dtype=np.float64)
    if scale:
        x, center, sc = scale_coords(x)                                          # :606-611
        grid = (grid - center) / sc                                              # :612
    dens = density_stage(x, grid)["density"]                                     # :624-632
    m_ = expression.shape[0]
    out = np.empty((m_, grid.shape[0]))
    for g in range(m_):
        if isinstance(expression, RowSparse):
            w = extract_row_dgRMatrix(expression, g + 1)                          # :655
        else:
            w = np.asarray(expression[g], dtype=np.float64)                       # :647
        out[g] = (dens * w[:, None]).sum(axis=0)
    return out / out.sum(axis=1, keepdims=True)                                   # :664


# --------------------------------------------------------------------------------------------
# the whole path (R/haystack_continuous.R:27-339) with the random inputs supplied
# --------------------------------------------------------------------------------------------
def haystack_continuous_highD(x, expression, grid_coord, perm_source, sets_mean, sets_sd,
                              weights_advanced_q=None, scale=True, randomization_count=100,
                              n_genes_to_randomize=100,
                              selection_method_genes_to_randomize="heavytails"):
    """Restates the orchestrator with `grid.coord` given (in the SCALED space, as the reference
    expects a supplied grid to be: :157-168 use it directly against scale(x)).

    perm_source(kind, gene_index0, n, size, count) -> (count, size) 1-based int array; it
    stands in for R's sample(): kind "dense" => full permutations of 1..n (size == n),
    kind "sparse" => `size` distinct ids from 1..n.
    """
    sparse = isinstance(expression, RowSparse)
    mean, sd = row_mean_sd(expression)                                           # :57-74
    if mean.min() < 0:
        raise ValueError("Some features have an average signal < 0. Expect average signal >= 0.")
    keep = ~(sd == 0)                                                            # :76-79
    if sparse:
        rows = np.nonzero(keep)[0]
        p_new = [0]
        js, xs = [], []
        for i in rows:
            js.append(expression.j[expression.p[i]:expression.p[i + 1]])
            xs.append(expression.x[expression.p[i]:expression.p[i + 1]])
            p_new.append(p_new[-1] + js[-1].size)
        expression = RowSparse(p=np.asarray(p_new, dtype=np.int64),
                               j=np.concatenate(js) if js else np.zeros(0, np.int32),
                               x=np.concatenate(xs) if xs else np.zeros(0),
                               shape=(rows.size, expression.shape[1]))
    else:
        expression = np.asarray(expression, dtype=np.float64)[keep]
    mean = mean[keep] + 1e-300                                                   # :82
    sd = sd[keep] + 1e-300                                                       # :83
    m_, n_ = expression.shape
    s_ = min(int(n_genes_to_randomize), m_)                                      # :120-124
    x = np.asarray(x, dtype=np.float64)
    if scale:
        x, center, sc = scale_coords(x)                                          # :139-145
    else:
        center = sc = None
    ds = density_stage(x, grid_coord, weights_advanced_q)                        # :168-186
    dens, q = ds["density"], ds["Q"]
    dkl_obs = dkl_observed(expression, dens, q, PSEUDO)                          # :199-213
    cv = sd / mean                                                               # :222
    genes = select_genes_to_randomize(cv, s_, selection_method_genes_to_randomize)   # :227-236
    if sparse:
        vals, inds = [], []
        for g in genes:
            _, val = extract_row_dgRMatrix_as_sparse(expression, int(g) + 1)     # :271
            vals.append(val)
            inds.append(perm_source("sparse", int(g), n_, val.size, randomization_count))   # :275
        dkl_rand = dkl_randomized_sparse(vals, inds, dens, q, PSEUDO)
    else:
        perms = np.stack([perm_source("dense", int(g), n_, n_, randomization_count) for g in genes])
        dkl_rand = dkl_randomized_dense(expression[genes], perms, dens, q, PSEUDO)
    fit = get_log_p_D_KL_continuous(dkl_obs, dkl_rand, cv, cv[genes], sets_mean, sets_sd)   # :288-293
    logp = fit["log_p"]
    padj = logp + math.log10(logp.size)                                          # :301
    padj = np.where(padj > 0, 0.0, padj)                                         # :302
    grid_out = grid_coord * sc + center if scale else grid_coord                 # :313-314
    return dict(D_KL=dkl_obs, log_p_vals=logp, log_p_adj=padj, keep=keep, cv=cv,
                genes_to_randomize=genes, KLD_rand=dkl_rand, fit=fit, grid_coordinates=grid_out,
                coord_mean=center, coord_std=sc, densities=dens, Q=q, bandwidth=ds["bandwidth"],
                nearest=ds["nearest"])
