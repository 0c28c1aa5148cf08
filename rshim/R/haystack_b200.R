# Drop-in replacement of the three hot spans of haystack_continuous_highD()
# (singleCellHaystack R/haystack_continuous.R:168-186, :199-213, :245-284) by device calls.
# Written without an R interpreter (none in the build image): see INTEGRATION.md and tests/test_rshim_cpu.py for what
# is checked statically and against a mock of the R C API.
#
# Usage inside the patched haystack_continuous_highD(), everything else of the function unchanged:
#
#   dev <- b200_context()
#   st  <- b200_density(dev, x, grid.coord, weights.advanced.Q)      # replaces :168-186
#   bandwidth <- st$bandwidth; Q <- st$Q; density.contributions <- st$densities
#   D_KL.observed <- if (is.matrix(expression)) b200_dkl_dense(dev, expression)
#                    else b200_dkl_csr(dev, expression)              # replaces :199-213
#   all.D_KL.randomized <- b200_randomize(dev, expression, genes.to.randomize,
#                                         randomization.count, count.cells)   # replaces :245-284
#
# The random draws stay in R, in the reference's order (gene-major, permutation-minor), so
# set.seed() reproducibility is untouched.

b200_context <- function(device = 0L) .Call("hsb_create", as.integer(device), PACKAGE = "haystackB200")

b200_density <- function(dev, x, grid.coord, weights.advanced.Q = NULL, densities = TRUE) {
  storage.mode(x) <- "double"; storage.mode(grid.coord) <- "double"
  res <- .Call("hsb_density", dev, x, grid.coord,
               if (is.null(weights.advanced.Q)) NULL else as.double(weights.advanced.Q),
               densities, PACKAGE = "haystackB200")
  names(res) <- c("bandwidth", "Q", "densities")
  res
}

b200_dkl_dense <- function(dev, expression) {
  storage.mode(expression) <- "double"
  .Call("hsb_dkl_dense", dev, expression, PACKAGE = "haystackB200")
}

b200_dkl_csr <- function(dev, expression) {          # expression: dgRMatrix
  .Call("hsb_dkl_csr", dev, expression@p, expression@j, expression@x, PACKAGE = "haystackB200")
}

b200_randomize <- function(dev, expression, genes.to.randomize, randomization.count, count.cells,
                           draw = c("R", "device"), seed = sample.int(.Machine$integer.max, 1), batch.genes = 10L) {
  draw <- match.arg(draw)
  n.sel <- length(genes.to.randomize)
  if (is.matrix(expression) && draw == "device") {
    # the whole loop :250-265 in one call; stream of (i, r) = (i - 1) * randomization.count + r - 1
    storage.mode(expression) <- "double"
    return(.Call("hsb_dkl_perm_dense_seeded_batch", dev, expression, as.integer(genes.to.randomize),
                 as.integer(randomization.count), as.double(seed), PACKAGE = "haystackB200"))
  }
  if (is.matrix(expression)) {
    out <- matrix(NA_real_, n.sel, randomization.count)
    for (i in seq_len(n.sel)) {
      v <- expression[genes.to.randomize[i], ]
      if (draw == "device") {
        out[i, ] <- .Call("hsb_dkl_perm_dense_seeded", dev, as.double(v), as.integer(randomization.count),
                          as.double(seed), as.double((i - 1) * randomization.count), PACKAGE = "haystackB200")
        next
      }
      # sample(x = v) == v[sample.int(length(v))]: same RNG consumption as the reference (:259)
      perm <- vapply(seq_len(randomization.count), function(r) sample.int(count.cells), integer(count.cells))
      out[i, ] <- .Call("hsb_dkl_perm_dense", dev, as.double(v), perm, PACKAGE = "haystackB200")
    }
    return(out)
  }
  if (draw == "device") {
    # the library draws sample(x = count.cells, size = nnz) itself (keyed permutation, stream =
    # (i - 1) * randomization.count + r - 1); b200_perm_sample() returns the ids of any one draw
    vals <- lapply(genes.to.randomize, function(g) {
      lo <- expression@p[g] + 1L; hi <- expression@p[g + 1L]
      if (hi >= lo) expression@x[lo:hi] else numeric(0)
    })
    gene.ptr <- c(0, cumsum(vapply(vals, length, 0)))
    return(.Call("hsb_dkl_perm_csr_seeded", dev, as.double(gene.ptr), unlist(vals),
                 as.integer(randomization.count), as.double(seed), PACKAGE = "haystackB200"))
  }
  # R draws the ids (in the reference's order: gene-major, randomization-minor) for one batch of genes while the
  # previous batch is uploaded and scored (hs_dkl_perm_csr_begin / _wait)
  out <- matrix(NA_real_, n.sel, randomization.count)
  batches <- split(seq_len(n.sel), ceiling(seq_len(n.sel) / batch.genes))
  in.flight <- NULL
  for (b in batches) {
    vals <- vector("list", length(b)); inds <- vector("list", length(b))
    for (k in seq_along(b)) {
      g <- genes.to.randomize[b[k]]
      lo <- expression@p[g] + 1L; hi <- expression@p[g + 1L]
      nnz <- hi - lo + 1L
      vals[[k]] <- if (nnz > 0) expression@x[lo:hi] else numeric(0)
      # column r = sample(x = count.cells, size = nnz), as R/haystack_continuous.R:275
      inds[[k]] <- as.integer(vapply(seq_len(randomization.count),
                                     function(r) sample(x = count.cells, size = nnz), integer(nnz)))
    }
    gene.ptr <- c(0, cumsum(vapply(vals, length, 0)))
    if (length(batches) == 1L)      # nothing to overlap with: the blocking call
      return(.Call("hsb_dkl_perm_csr", dev, as.double(gene.ptr), as.double(unlist(vals)), as.integer(unlist(inds)),
                   as.integer(randomization.count), PACKAGE = "haystackB200"))
    if (!is.null(in.flight)) out[in.flight, ] <- .Call("hsb_dkl_perm_csr_wait", dev, PACKAGE = "haystackB200")
    .Call("hsb_dkl_perm_csr_begin", dev, as.double(gene.ptr), as.double(unlist(vals)), as.integer(unlist(inds)),
          as.integer(randomization.count), PACKAGE = "haystackB200")
    in.flight <- b
  }
  if (!is.null(in.flight)) out[in.flight, ] <- .Call("hsb_dkl_perm_csr_wait", dev, PACKAGE = "haystackB200")
  out
}

b200_perm_sample <- function(seed, stream, n, size)
  .Call("hsb_perm_sample", as.double(seed), as.double(stream), as.double(n), as.double(size), PACKAGE = "haystackB200")

# D_KL of every gene and the seeded randomization of genes.to.randomize in one call (the randomization batch
# computes while the matrix uploads): replaces :199-213 and :245-284 once genes.to.randomize (:222-236) is known
b200_dkl_csr_randomized <- function(dev, expression, genes.to.randomize, randomization.count,
                                    seed = sample.int(.Machine$integer.max, 1)) {
  res <- .Call("hsb_dkl_csr_randomized", dev, expression@p, expression@j, expression@x,
               as.integer(genes.to.randomize), as.integer(randomization.count), as.double(seed),
               PACKAGE = "haystackB200")
  names(res) <- c("D_KL", "all.D_KL.randomized")
  res
}
