# haystack_continuous_highD() on the B200.
#
# Same signature, same messages / warnings / errors for the same inputs, same "haystack" result object as
# singleCellHaystack::haystack_continuous_highD (reference R/haystack_continuous.R:27-339).  Three spans run on the
# device instead of in R:
#     :168-186  distances, bandwidth, densities, Q          -> b200_density()
#     :199-213  observed D_KL of every gene                  -> b200_dkl_dense() / b200_dkl_csr()
#     :245-284  randomized D_KL of the selected genes        -> b200_randomize()  (R draws the ids, in R's order)
# Everything that consumes R's random stream (kmeans in get_grid_points, sample() for the randomizations and for the
# cross-validation folds) is still R's own code, called in the reference's order, so set.seed() gives the reference's
# draws; the spline fit is the reference's get_log_p_D_KL_continuous.  WRITTEN WITHOUT AN R INTERPRETER (none in the
# build image); tests/test_rshim_cpu.py checks what can be checked statically (every .Call target is registered
# with the right arity, the C shim compiles and runs against a mock of R's C API).
#
# b200_enable() swaps this function into the singleCellHaystack namespace, after which
# singleCellHaystack::haystack(x, expression, ...) -- matrix, data.frame, Seurat and SingleCellExperiment
# methods alike (R/s3.R:21-87) -- runs on the B200 without any change in user code.

.b200_msg <- function(...) message("### ", ...)

# row means / sds exactly where the reference takes them from (:57-83)
.b200_row_stats <- function(expression) {
  use_sms <- requireNamespace("sparseMatrixStats", quietly = TRUE)
  if (use_sms) .b200_msg("Using package sparseMatrixStats to speed up statistics in sparse matrices.")
  else .b200_msg("Package sparseMatrixStats not found. Install for speed improvements.")
  .b200_msg("Calculating row-wise mean and SD... ")
  m <- Matrix::rowMeans(expression)
  s <- if (use_sms && !is.matrix(expression)) sparseMatrixStats::rowSds(expression, useNames = FALSE)
       else apply(expression, 1, stats::sd)
  names(s) <- names(m)
  list(mean = m, sd = s)
}

# the integer choice of genes whose rows are permuted (:227-236)
.b200_pick_genes <- function(coeffVar, n.pick, how) {
  ord <- order(coeffVar)
  n <- length(ord)
  if (how == "uniform") return(ord[floor(seq(1, n, length.out = n.pick))])
  if (how == "heavytails")
    return(ord[c(1:9, as.integer(seq(10, n - 9, length.out = n.pick - 18)), n - (8:0))])
  stop("'selection.method.genes.to.randomize' should be either 'uniform' or 'heavytails'")
}

haystack_continuous_highD <- function(x, expression, grid.points = 100, weights.advanced.Q = NULL,
                                      dir.randomization = NULL, scale = TRUE, grid.method = "centroid",
                                      randomization.count = 100, n.genes.to.randomize = 100,
                                      selection.method.genes.to.randomize = "heavytails", grid.coord = NULL,
                                      spline.method = "ns") {
  sch <- asNamespace("singleCellHaystack")
  .b200_msg("calling haystack_continuous_highD()...")

  if (!is.null(grid.coord)) {
    grid.points <- nrow(grid.coord)
    if (ncol(x) != ncol(grid.coord)) stop("Coordinates and grid points have different number of columns.")
    if ((!is.null(colnames(x)) || !is.null(colnames(grid.coord))) && !identical(colnames(x), colnames(grid.coord)))
      stop("Coordinates and grid points have different column names.")
  }

  # ---- per-gene statistics, zero-variance filter ----
  st <- .b200_row_stats(expression)
  if (min(st$mean) < 0) stop("Some features have an average signal < 0. Expect average signal >= 0.")
  flat <- st$sd == 0
  .b200_msg("Filtered ", sum(flat), " genes with zero variance...")
  expression <- expression[!flat, , drop = FALSE]
  expr.mean <- st$mean[!flat] + 1e-300
  expr.sd <- st$sd[!flat] + 1e-300
  .b200_msg("Using ", randomization.count, " randomizations...")
  .b200_msg("Using ", n.genes.to.randomize, " genes to randomize...")

  # ---- argument checks (messages as in the reference, :89-136) ----
  if (!is.numeric(x) && !is.matrix(x) && !is.data.frame(x)) stop("'x' must be a numeric matrix or data.frame")
  if (!is.matrix(expression) && !inherits(expression, "dgCMatrix") && !inherits(expression, "dgRMatrix"))
    stop("'expression' must be a matrix, dgCMatrix, or dgRMatrix")
  if (ncol(expression) != nrow(x)) stop("The number of columns in 'expression' must be the same as the rows in 'x'")
  if (!is.numeric(grid.points)) stop("The value of 'grid.points' must be a numeric")
  if (grid.points >= ncol(expression))
    stop("The number of grid points appears to be very high (higher than the number of cells). You can set the number of grid points using the 'grid.points' parameter.")
  if (!is.null(weights.advanced.Q)) {
    if (!is.numeric(weights.advanced.Q)) stop("'weights.advanced.Q' must either be NULL or a numeric vector")
    if (length(weights.advanced.Q) != nrow(x)) stop("The length of 'weights.advanced.Q' must be the same as the number of rows in 'x'")
  }
  if (!spline.method %in% c("ns", "bs")) stop("'spline.method' should be either 'ns' or 'bs'")
  if (inherits(expression, "dgCMatrix")) {
    .b200_msg("converting expression data from dgCMatrix to dgRMatrix")
    expression <- methods::as(expression, "RsparseMatrix")
  }
  count.cells <- ncol(expression)
  count.genes <- nrow(expression)
  if (n.genes.to.randomize > count.genes) {
    warning("Number of genes to randomize (", n.genes.to.randomize, ") is higher than the number of genes in the data.")
    warning("Setting number of genes to randomize to ", count.genes)
    n.genes.to.randomize <- count.genes
  }
  if (nrow(x) < 50) warning("The number of cells seems very low (", nrow(x), "). Check your input.")
  if (count.genes < 100) warning("The number of genes seems very low (", count.genes, "). Check your input.")
  if (grid.points < 10)
    warning("The value of 'grid.points' appears to be very low (<10). You can set the number of grid points using the 'grid.points' parameter.")
  if (grid.points > count.cells / 10)
    warning("The value of 'grid.points' appears to be very high (> No. of cells / 10). You can set the number of grid points using the 'grid.points' parameter.")

  if (scale) {
    .b200_msg("scaling input data...")
    x <- base::scale(x)
    x.scale.center <- attr(x, "scaled:center")
    x.scale.scale <- attr(x, "scaled:scale")
  }
  if (!is.null(dir.randomization) && !dir.exists(dir.randomization)) dir.create(dir.randomization)

  if (is.null(grid.coord)) {
    .b200_msg("deciding grid points...")
    grid.coord <- sch$get_grid_points(input = x, method = grid.method, grid.points = grid.points)   # R's kmeans / sample
    if (nrow(grid.coord) != grid.points) {
      warning("The number of grid points was changed from ", grid.points, " to ", nrow(grid.coord))
      grid.points <- nrow(grid.coord)
    }
  }

  # ---- device: density stage ----
  dev <- b200_context()
  dens <- b200_density(dev, as.matrix(x), as.matrix(grid.coord), weights.advanced.Q, densities = TRUE)
  density.contributions <- dens$densities

  # ---- device: observed D_KL ----
  .b200_msg("calculating Kullback-Leibler divergences...")
  D_KL.observed <- if (is.matrix(expression)) b200_dkl_dense(dev, expression) else b200_dkl_csr(dev, expression)

  # ---- host: which genes to permute; device: their randomized D_KL, ids drawn by R in the reference's order ----
  coeffVar <- expr.sd / expr.mean
  .b200_msg("performing randomizations...")
  genes.to.randomize <- .b200_pick_genes(coeffVar, n.genes.to.randomize, selection.method.genes.to.randomize)
  all.D_KL.randomized <- b200_randomize(dev, expression, genes.to.randomize, randomization.count, count.cells)

  # ---- host: the reference's own significance fit (draws the cross-validation folds with sample()) ----
  .b200_msg("estimating p-values...")
  fit <- sch$get_log_p_D_KL_continuous(D_KL.observed = D_KL.observed, D_KL.randomized = all.D_KL.randomized,
                                       all.coeffVar = coeffVar, train.coeffVar = coeffVar[genes.to.randomize],
                                       output.dir = dir.randomization, spline.method = spline.method)
  rand.info <- fit$info
  rand.info$method <- spline.method
  rand.info$genes_to_randomize <- genes.to.randomize
  rand.info$KLD_rand <- all.D_KL.randomized
  p.vals <- fit$fitted
  p.adjs <- pmin(0, p.vals + log10(length(p.vals)))          # Bonferroni on the log10 scale

  if (!is.null(dir.randomization)) {
    .b200_msg("writing randomized Kullback-Leibler divergences to file...")
    utils::write.csv(file = paste0(dir.randomization, "/random.D_KL.csv"), all.D_KL.randomized)
  }

  if (scale) grid.coord <- grid.coord * rep(x.scale.scale, each = nrow(grid.coord)) + rep(x.scale.center, each = nrow(grid.coord))
  .b200_msg("returning result...")
  res <- list(
    results = data.frame(D_KL = D_KL.observed, log.p.vals = p.vals, log.p.adj = p.adjs, row.names = row.names(expression)),
    info = list(method = "continuous_highD", randomization = rand.info, grid.coordinates = grid.coord,
                coord_mean = if (scale) x.scale.center else NULL, coord_std = if (scale) x.scale.scale else NULL,
                densities = density.contributions, cv = coeffVar))
  class(res) <- "haystack"
  res
}

# Make singleCellHaystack::haystack() and its methods use the B200 path / restore the original.
.b200_saved <- new.env(parent = emptyenv())

b200_enable <- function() {
  ns <- asNamespace("singleCellHaystack")
  if (is.null(.b200_saved$original)) .b200_saved$original <- get("haystack_continuous_highD", envir = ns)
  utils::assignInNamespace("haystack_continuous_highD", haystack_continuous_highD, ns = "singleCellHaystack")
  invisible(TRUE)
}

b200_disable <- function() {
  if (!is.null(.b200_saved$original))
    utils::assignInNamespace("haystack_continuous_highD", .b200_saved$original, ns = "singleCellHaystack")
  invisible(TRUE)
}
