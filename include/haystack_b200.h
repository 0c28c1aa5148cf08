/*
 * haystack_b200.h -- C ABI of libhaystack_b200.so
 *
 * B200 (sm_100a) implementation of the continuous high-D scoring path of singleCellHaystack.
 * The reference (v1.0.3, pure interpreted R) has NO native interface (no src/, no useDynLib:
 * NAMESPACE:1-59), so this boundary is created at the three spans of
 * R/haystack_continuous.R that are replaced by device calls:
 *
 *   hs_density            <- R/haystack_continuous.R:168-186  (get_dist_two_sets R/haystack_highD.R:17-24,
 *                            bandwidth :174, density :175-177, reference distribution Q :179-186)
 *   hs_dkl_dense          <- R/haystack_continuous.R:202-206  (get_D_KL_continuous_highD :350-363)
 *   hs_dkl_csr            <- R/haystack_continuous.R:207-212  (extract_row_dgRMatrix_as_sparse R/sparse.R:39-49,
 *                            get_D_KL_continuous_highD_SPARSE :586-602)
 *   hs_dkl_perm_dense     <- R/haystack_continuous.R:250-265  (dense randomization loop)
 *   hs_dkl_perm_csr       <- R/haystack_continuous.R:266-283  (sparse randomization loop)
 *   hs_dkl_perm_csr_seeded<- the same loop including its sample() call (:275), drawn on the device
 *   hs_dkl_perm_dense_seeded <- the dense loop including its sample() call (:259), drawn on the device
 *                            (hs_dkl_perm_dense_seeded_batch: all selected genes in one call)
 *   hs_dkl_csr_randomized <- :199-213 and :245-284 in one call (the randomization batch computes under the upload)
 * and, around the path (optional; the host keeps them by default):
 *   hs_row_stats_csr      <- R/haystack_continuous.R:57-74    (row means / sds)
 *   hs_select_by_cv_dev   <- R/haystack_continuous.R:222, :227-236 (coeffVar, order(), the selected rows)
 *   hs_csc_to_csr         <- R/haystack_continuous.R:113-116  (as(expression, "RsparseMatrix"))
 *   hs_profile_csr/_dense <- R/haystack_continuous.R:604-666  (prepare_clustering)
 *   hs_density_shard_*    <- :168-186 sharded over the cells of several ranks (three all-gathers by the host layer)
 *   hs_r_*                <- base R's set.seed / unif_rand / sample.int / sample(prob=) / kmeans (Hartigan-Wong) on
 *                            the host, bit for bit: what get_grid_points (R/haystack_highD.R:72-134), :259, :275 and
 *                            :437 consume, for hosts without an R interpreter
 *
 * Conventions
 *   - Plain C, POD arguments only.  No R, torch or CUDA types in any signature.
 *   - Host matrices use R's layout: column-major double.  A dgRMatrix is passed as its three
 *     slots (@p int32 or int64, @j int32 0-based, @x double).
 *   - Every function returns 0 on success, non-zero on failure; the message is available from
 *     hs_last_error(ctx) (or hs_last_error(NULL) when no context exists).  Nothing throws or
 *     longjmps across this boundary.
 *   - The caller owns every host buffer for the duration of the call only.  The context owns
 *     all device memory and streams.  The ids-given entry points (hs_dkl_perm_dense,
 *     hs_dkl_perm_csr) draw NO random numbers: permutation indices are inputs (the R shim produces
 *     them with sample()/sample.int(), so set.seed() reproducibility is the host's).  The *_seeded
 *     entry points, hs_dkl_csr_randomized and hs_perm_sample draw from a keyed permutation that is
 *     a function of (seed, stream) only -- reproducible, but not R's Mersenne-Twister stream.
 *   - One context per process and device (one process per GPU; multi-GPU gene sharding is done
 *     by the host layer with torch.distributed/NCCL).
 *   - There is no CPU fallback: every entry point fails with HS_ERR_CUDA when no sm_100 device
 *     is usable.
 */
#ifndef HAYSTACK_B200_H
#define HAYSTACK_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HS_ABI_VERSION 1

enum {
  HS_OK = 0,
  HS_ERR_INVALID = 1,   /* bad argument (NULL pointer, negative size, shape mismatch) */
  HS_ERR_STATE = 2,     /* call order: no density stage resident yet */
  HS_ERR_CUDA = 3,      /* CUDA runtime error or no usable device */
  HS_ERR_NOMEM = 4      /* device or host allocation failed */
};

typedef struct hs_ctx hs_ctx;

/* ---- lifetime -------------------------------------------------------------------------- */
int hs_abi_version(void);
/* Create a context on CUDA device `device` (ordinal as seen by this process). */
int hs_create(int device, hs_ctx** out);
void hs_destroy(hs_ctx* ctx);
const char* hs_last_error(const hs_ctx* ctx);
/* Run all work of this context on an externally owned CUDA stream (a cudaStream_t passed as
 * void*; NULL restores the context's own stream).  Used by the host layer so that its timing
 * events and the library's kernels share one stream. */
int hs_set_stream(hs_ctx* ctx, void* cuda_stream);
int hs_synchronize(hs_ctx* ctx);

/* ---- density stage (K1) ---------------------------------------------------------------- */
/* x: n_cells x n_dims column-major, ALREADY scaled by the caller (scale() stays in R,
 * R/haystack_continuous.R:139-145).  grid: n_grid x n_dims column-major in the same space.
 * w_q: n_cells cell weights for Q (weights.advanced.Q, :182) or NULL.
 * Outputs (each may be NULL): bandwidth_out[1]; q_out[n_grid]; density_out[n_cells*n_grid]
 * column-major fp64 (info$densities, :331); nearest_out[n_cells] 0-based index of the nearest
 * grid point (first minimum).
 * Leaves the density matrix (fp32), Q and the bandwidth resident in the context. */
int hs_density(hs_ctx* ctx, const double* x, int64_t n_cells, int n_dims, const double* grid,
               int n_grid, const double* w_q, double* bandwidth_out, double* q_out,
               double* density_out, int32_t* nearest_out);

/* Multi-GPU (one process per GPU): the rank that ran hs_density exposes its resident state, the
 * other ranks reserve buffers of the same shape, receive into them with any collective (NCCL
 * broadcast over NVLink in the host layer) and commit.  d_dev: fp32 density matrix, d_bytes long;
 * q_dev: n_grid doubles of Q followed by one double holding the bandwidth, q_bytes long.
 * On a context with a resident state of exactly this shape the call only reports the pointers;
 * otherwise it (re)allocates and marks the state invalid until hs_density_commit. */
int hs_density_buffers(hs_ctx* ctx, int64_t n_cells, int n_grid, void** d_dev, int64_t* d_bytes,
                       void** q_dev, int64_t* q_bytes);
/* Declare the buffers filled: reads the bandwidth back from q_dev[n_grid] and validates it. */
int hs_density_commit(hs_ctx* ctx, int64_t n_cells, int n_grid);
/* The same exchange with the CELLS sharded instead (R/haystack_continuous.R:168-186 is per cell but for one scalar,
 * the bandwidth = median over all cells of the distance to the nearest grid point): rank r of `world` computes the
 * squared distances and row minima of its cell range (hs_density_shard_dist_dev), the host layer all-gathers the row
 * minima in place, every rank takes the exact median of all of them and fills the density rows and per-block column
 * sums of its range (hs_density_shard_dens_dev), those are all-gathered in place too, and hs_density_shard_finish
 * reduces Q in the fixed order of the single-GPU path -- so every rank ends with the bits hs_density_dev would have
 * produced, and no rank computes more than 1/world of the stage.  Ranges are whole 128-cell blocks;
 * hs_density_shard_buffers reports the three device buffers and the bytes each rank contributes (rank r's part starts
 * at r * bytes_per_rank), to be called after hs_density_shard_dist_dev. */
int hs_density_shard_dist_dev(hs_ctx* ctx, const double* x_dev, int64_t n_cells, int n_dims,
                              const double* grid_dev, int n_grid, int rank, int world);
int hs_density_shard_buffers(hs_ctx* ctx, int64_t n_cells, int n_grid, int world, void** rowmin_dev,
                             int64_t* rowmin_bytes_per_rank, void** d_dev, int64_t* d_bytes_per_rank,
                             void** qpart_dev, int64_t* qpart_bytes_per_rank);
int hs_density_shard_dens_dev(hs_ctx* ctx, int64_t n_cells, int n_grid, const double* w_q_dev, int rank,
                              int world);
int hs_density_shard_finish(hs_ctx* ctx, int64_t n_cells, int n_grid, double* bandwidth_out, double* q_out);
/* Shape of the resident state (0,0 when none). */
int hs_density_shape(const hs_ctx* ctx, int64_t* n_cells_out, int* n_grid_out);

/* ---- observed D_KL (K2 dense, K3 CSR) --------------------------------------------------- */
/* expr: n_genes x n_cells column-major double with leading dimension ld (>= n_genes), i.e. an
 * R matrix (rows = genes).  dkl_out[n_genes]. */
int hs_dkl_dense(hs_ctx* ctx, const double* expr, int64_t n_genes, int64_t ld, double* dkl_out);

/* dgRMatrix slots.  p has n_genes+1 entries (int32 when p_is_64 == 0, else int64); j are
 * 0-based cell ids (any order within a row; Matrix keeps them ascending); x the values.
 * Explicit zeros are treated like any other stored value, as the reference does (:272). */
int hs_dkl_csr(hs_ctx* ctx, const void* p, int p_is_64, const int32_t* j, const double* x,
               int64_t n_genes, double* dkl_out);

/* The same with single-precision values, for hosts that hold float32 matrices (AnnData / scanpy default):
 * the kernels multiply in fp32 either way, so nothing is lost and a third of the upload is saved.
 * (R has no single-precision vectors; the R shim uses hs_dkl_csr.) */
int hs_dkl_csr_f32(hs_ctx* ctx, const void* p, int p_is_64, const int32_t* j, const float* x,
                   int64_t n_genes, double* dkl_out);

/* ---- randomized D_KL (K4) --------------------------------------------------------------- */
/* Dense form: v[n_cells] is the gene's expression row; perm is n_cells x n_perm column-major,
 * column r being sample.int(n_cells) (values index_base .. n_cells-1+index_base; R passes
 * index_base = 1).  out[r] = D_KL of v[perm[, r]]  (sample(x = v) == v[sample.int(length(v))]). */
int hs_dkl_perm_dense(hs_ctx* ctx, const double* v, const int32_t* perm, int n_perm,
                      int index_base, double* out);

/* Sparse form, batched over genes so that several genes share a launch.
 * gene_ptr[n_sel+1]: prefix offsets into val (gene s owns val[gene_ptr[s] .. gene_ptr[s+1])).
 * ind: for gene s, an nnz_s x n_perm column-major block starting at element
 * gene_ptr[s]*n_perm; column r is sample(x = n_cells, size = nnz_s) -- distinct cell ids paired
 * in order with val (:275).  out is n_sel x n_perm column-major (all.D_KL.randomized, :245). */
int hs_dkl_perm_csr(hs_ctx* ctx, const int64_t* gene_ptr, int64_t n_sel, const double* val,
                    const int32_t* ind, int n_perm, int index_base, double* out);

/* Asynchronous pair of hs_dkl_perm_csr: _begin starts the call on a worker thread of the library and returns at once,
 * so the host can draw the ids of the next batch (R's sample()) while this one is uploaded and scored; _wait blocks
 * until it is done and writes the n_sel x n_perm result.  val and ind must stay valid until _wait returns (gene_ptr
 * is copied).  One call in flight per context; every other entry point fails with HS_ERR_STATE until it is waited for. */
int hs_dkl_perm_csr_begin(hs_ctx* ctx, const int64_t* gene_ptr, int64_t n_sel, const double* val,
                          const int32_t* ind, int n_perm, int index_base);
int hs_dkl_perm_csr_wait(hs_ctx* ctx, double* out);

int hs_dkl_perm_csr_f32(hs_ctx* ctx, const int64_t* gene_ptr, int64_t n_sel, const float* val,
                        const int32_t* ind, int n_perm, int index_base, double* out);

/* Sparse form with the draw made on the device.  Same scoring as hs_dkl_perm_csr, but instead of taking
 * the ids of sample(x = n_cells, size = nnz_s) (R/haystack_continuous.R:275) from the caller, gene s /
 * randomization r pairs val[gene_ptr[s] + i] with cell pi(i), i = 0 .. nnz_s-1, where pi is the keyed
 * permutation of [0, n_cells) for (seed, stream_base + s * n_perm + r) defined in csrc/hs_prp.cuh
 * (4-round Feistel network with cycle walking).  No ids cross PCIe, nothing is sorted, and repeated
 * calls give bit-identical results.  hs_perm_sample returns the very ids a (seed, stream) draws, so a
 * caller (or a test) can reproduce a randomization through hs_dkl_perm_csr.
 * out is n_sel x n_perm column-major. */
int hs_dkl_perm_csr_seeded(hs_ctx* ctx, const int64_t* gene_ptr, int64_t n_sel, const double* val,
                           int n_perm, uint64_t seed, uint64_t stream_base, double* out);
/* Dense form with the draw made on the device: randomization r scores v[pi_r(c)] at cell c, pi_r keyed by
 * (seed, stream_base + r) -- sample(x = v) == v[sample.int(n_cells)] (R/haystack_continuous.R:259) without the
 * n_cells x n_perm index matrix.  hs_perm_sample(seed, stream_base + r, n_cells, n_cells, ...) returns pi_r. */
int hs_dkl_perm_dense_seeded(hs_ctx* ctx, const double* v, int n_perm, uint64_t seed, uint64_t stream_base,
                             double* out);

/* The dense randomization of ALL selected genes in one call (the whole loop R/haystack_continuous.R:250-265): row
 * sel_rows[s] (0-based) of the n_genes x n_cells column-major matrix `expr` (leading dimension ld), n_perm seeded
 * randomizations each, stream of (s, r) = stream_base + s * n_perm + r -- the same draws as n_sel calls of
 * hs_dkl_perm_dense_seeded.  Nothing but the selected rows crosses the link; uploads overlap the scoring; one wait.
 * out: n_sel x n_perm column-major (all.D_KL.randomized). */
int hs_dkl_perm_dense_seeded_batch(hs_ctx* ctx, const double* expr, int64_t ld, const int64_t* sel_rows,
                                   int64_t n_sel, int n_perm, uint64_t seed, uint64_t stream_base, double* out);
/* hs_dkl_csr and hs_dkl_perm_csr_seeded for the rows sel_rows[0..n_sel) (0-based row numbers of the same
 * matrix, in the order of genes.to.randomize) in one call: the randomization batch is launched first and
 * computes while the matrix is still being uploaded.  Results are bit-identical to the two separate calls
 * (R/haystack_continuous.R:199-213 and :245-284; the selection :222-236 needs only the row statistics, so it
 * can precede both).  rand_out: n_sel x n_perm column-major. */
int hs_dkl_csr_randomized(hs_ctx* ctx, const void* p, int p_is_64, const int32_t* j, const double* x,
                          int64_t n_genes, const int64_t* sel_rows, int64_t n_sel, int n_perm,
                          uint64_t seed, uint64_t stream_base, double* dkl_out, double* rand_out);
/* Host helper (no GPU involved): ids_out[i] = pi(i) + index_base for i < size, pi keyed by
 * (seed, stream) over [0, n).  1 <= n < 2^31, 0 <= size <= n. */
int hs_perm_sample(uint64_t seed, uint64_t stream, int64_t n, int64_t size, int index_base,
                   int32_t* ids_out);

/* ---- device-resident variants (inputs already in HBM; used by bench.py and by the
 *      multi-GPU host layer).  Pointers are device pointers in this process. --------------- */
/* CSR with int64 row pointers, int32 0-based columns, fp32 values; dkl_dev_out fp64[n_genes]. */
int hs_dkl_csr_dev(hs_ctx* ctx, const int64_t* p_dev, const int32_t* j_dev, const float* x_dev,
                   int64_t n_genes, double* dkl_dev_out);
/* gene_ptr_dev int64[n_sel+1], val_dev fp32, ind_dev int32 laid out as in hs_dkl_perm_csr;
 * out_dev fp64 n_sel x n_perm column-major. */
int hs_dkl_perm_csr_seeded_dev(hs_ctx* ctx, const int64_t* gene_ptr_dev, int64_t n_sel,
                               const float* val_dev, int n_perm, uint64_t seed, uint64_t stream_base,
                               double* out_dev);
/* The same for units whose streams the caller names: unit u, randomization r draws from stream
 * unit_stream_dev[u] + r.  Lets several ranks split the (gene, permutation block) units of one job and still draw
 * exactly what a single rank would (stream of gene s, randomization r = s * n_perm_total + r). */
int hs_dkl_perm_csr_seeded_units_dev(hs_ctx* ctx, const int64_t* gene_ptr_dev, int64_t n_units,
                                     const float* val_dev, int n_perm, uint64_t seed,
                                     const uint64_t* unit_stream_dev, double* out_dev);
int hs_dkl_perm_dense_seeded_dev(hs_ctx* ctx, const float* v_dev, int n_perm, uint64_t seed,
                                 uint64_t stream_base, double* out_dev);
int hs_dkl_perm_csr_dev(hs_ctx* ctx, const int64_t* gene_ptr_dev, int64_t n_sel,
                        const float* val_dev, const int32_t* ind_dev, int n_perm, int index_base,
                        double* out_dev);
/* Dense: expr_dev fp32, n_genes x n_cells column-major with leading dimension ld. */
int hs_dkl_dense_dev(hs_ctx* ctx, const float* expr_dev, int64_t n_genes, int64_t ld,
                     double* dkl_dev_out);
/* v_dev fp32[n_cells]; perm_dev int32 n_cells x n_perm column-major; out_dev fp64[n_perm]. */
int hs_dkl_perm_dense_dev(hs_ctx* ctx, const float* v_dev, const int32_t* perm_dev, int n_perm,
                          int index_base, double* out_dev);
/* Density stage on device-resident inputs (x_dev, grid_dev fp64 column-major; w_q_dev fp64 or
 * NULL); scalars/vectors come back through host pointers as in hs_density. */
int hs_density_dev(hs_ctx* ctx, const double* x_dev, int64_t n_cells, int n_dims,
                   const double* grid_dev, int n_grid, const double* w_q_dev,
                   double* bandwidth_out, double* q_out);

/* ---- gene x grid profiles (the step after the path; optional) ------------------------------- */
/* prepare_clustering (R/haystack_continuous.R:604-666), the input of hclust_haystack / kmeans_haystack
 * (R/s3.R:213-253): per gene the column sums colSums(density * expression[g, ]) (:647 / :655), every row rescaled
 * to sum to 1 (:664).  Same contraction as hs_dkl_csr / hs_dkl_dense without the KL reduction; uses the resident
 * density state.  profile_out: n_genes x n_grid column-major doubles (an R matrix). */
int hs_profile_csr(hs_ctx* ctx, const void* p, int p_is_64, const int32_t* j, const double* x,
                   int64_t n_genes, double* profile_out);
int hs_profile_dense(hs_ctx* ctx, const double* expr, int64_t n_genes, int64_t ld, double* profile_out);

/* ---- row statistics (the step before the path; optional) ----------------------------------- */
/* Matrix::rowMeans and the per-row sample sd (n-1 denominator, implicit zeros counted) of a CSR matrix --
 * R/haystack_continuous.R:57-74, the inputs of coeffVar (:222) and of the gene selection (:227-236).  Two
 * passes in fp64 on the device; the doubles are uploaded as they are (the result feeds an integer decision).
 * Needs no density state.  mean_out / sd_out: n_genes doubles each, WITHOUT the +1e-300 of :82-83. */
int hs_row_stats_csr(hs_ctx* ctx, const void* p, int p_is_64, const double* x, int64_t n_genes,
                     int64_t n_cells, double* mean_out, double* sd_out);
/* Same on the fp32 values already resident for hs_dkl_csr_dev (statistics of the narrowed values:
 * relative differences of order 1e-8 from the above). */
int hs_row_stats_csr_dev(hs_ctx* ctx, const int64_t* p_dev, const float* x_dev, int64_t n_genes,
                         int64_t n_cells, double* mean_dev, double* sd_dev);

/* Gene selection on the device (R/haystack_continuous.R:222, :227-236): coeffVar = (sd + 1e-300) / (mean + 1e-300)
 * of device-resident row statistics (hs_row_stats_csr_dev), order(coeffVar) -- stable, NaN last -- and the rows at
 * `ranks` (0-based positions in that order; the positions themselves are the reference's integer arithmetic on seq()
 * and stay with the caller).  cv_dev_out: device fp64[n_genes] or NULL; rows_out: host int64[n_ranks], 0-based. */
int hs_select_by_cv_dev(hs_ctx* ctx, const double* mean_dev, const double* sd_dev, int64_t n_genes,
                        const int64_t* ranks, int64_t n_ranks, double* cv_dev_out, int64_t* rows_out);

/* ---- grid selection, "seeding" method (the step before the path; optional) ------------------------------------ */
/* get_grid_points(method = "seeding") (R/haystack_highD.R:87-129) is a chain of distance passes over all cells with a
 * weighted draw between them.  The passes run here; the draws -- sample(N, 1) (:95) and sample(N, 1, prob = probs)
 * (:110) -- stay with the caller, who owns the random stream:
 *   hs_grid_seed_begin(x)        uploads the (scaled) coordinates, N x d column-major; min.dist.to.grid := Inf
 *   hs_grid_seed_step(index, probs_out, nearest3_out)   the cell `index` (0-based) becomes the next grid point:
 *        tmp = distances of all cells to it (:113), min.dist.to.grid = pmin(min.dist.to.grid, tmp) (:116);
 *        probs_out (host, N doubles, or NULL) = min.dist.to.grid^2, the weights of the NEXT draw (:107);
 *        nearest3_out (host, 3 x int64, 0-based, -1 when N < 3) = order(tmp)[1:3], the cells whose weighted mean
 *        (weights 2, 1, 1) replaces the grid point at the end (:124-128).  Needs no density state. */
int hs_grid_seed_begin(hs_ctx* ctx, const double* x, int64_t n_cells, int n_dims);
int hs_grid_seed_step(hs_ctx* ctx, int64_t index, double* probs_out, int64_t* nearest3_out);

/* ---- CSC -> CSR (the step before the sparse path; optional) -------------------------------- */
/* as(expression, "RsparseMatrix") (R/haystack_continuous.R:113-116) on the device: a stable counting sort by
 * row id, column ids ascending inside every row, bit-reproducible.  cp: column pointers of the dgCMatrix
 * (@p, int32 or int64, n_cols + 1 entries), ri: @i (0-based row ids), x: @x.  Outputs: the dgRMatrix slots
 * p_out int64[n_rows + 1], j_out int32[nnz], x_out double[nnz].  Any number of rows (bands of 56 000 rows, one pass
 * over the entries per band). */
int hs_csc_to_csr(hs_ctx* ctx, const void* cp, int cp_is_64, const int32_t* ri, const double* x,
                  int64_t n_rows, int64_t n_cols, int64_t* p_out, int32_t* j_out, double* x_out);
/* Device-resident form (int64 column pointers, fp32 values), producing what hs_dkl_csr_dev consumes. */
int hs_csc_to_csr_dev(hs_ctx* ctx, const int64_t* cp_dev, const int32_t* ri_dev, const float* x_dev,
                      int64_t n_rows, int64_t n_cols, int64_t* p_out_dev, int32_t* j_out_dev,
                      float* x_out_dev);

/* ---- R-compatible host helpers (no GPU involved; optional) --------------------------------- */
/* For hosts WITHOUT R that want R's draws: set.seed(seed) followed by unif_rand() / sample.int() / kmeans() of base R
 * (Mersenne-Twister; sample.kind "Rejection" as in R >= 3.6.0, or "Rounding" when sample_kind_rounding != 0).
 * With them the Python host layer replays `set.seed(1234); haystack(dat.tsne, dat.expression)` and reproduces the
 * reference's printed vignette table on the B200 (tests/test_gpu_r_parity.py).  The R shim does not use them: under
 * R the draws are R's own (R/haystack_continuous.R:259, :275, :437; R/haystack_highD.R:83, :95, :110). */
typedef struct hs_rng hs_rng;
int hs_r_rng_create(uint32_t seed, int sample_kind_rounding, hs_rng** out);
void hs_r_rng_destroy(hs_rng* rng);
double hs_r_unif_rand(hs_rng* rng);
/* `count` consecutive sample.int(n, size) draws (1-based ids), one after the other in ids_out. */
int hs_r_sample_int(hs_rng* rng, int32_t n, int32_t size, int32_t count, int32_t* ids_out);
/* sample(n, 1, prob = prob): one 1-based id. */
int hs_r_sample_prob_one(hs_rng* rng, int32_t n, const double* prob, int32_t* id_out);
/* kmeans(x, centers = k, iter.max, nstart)$centers, Hartigan-Wong; x: m x d column-major, centers_out: k x d. */
int hs_r_kmeans(hs_rng* rng, const double* x, int32_t m, int32_t d, int32_t k, int32_t iter_max, int32_t nstart,
                double* centers_out, double* tot_withinss_out, int32_t* ifault_out);

/* ---- instrumentation -------------------------------------------------------------------- */
/* Number of kernels of this library launched on this context since creation. */
int64_t hs_launch_count(const hs_ctx* ctx);
/* CUDA-event time (ms) of the dominant kernel family in the most recent call:
 * slot 0 density kernels, 1 CSR contraction, 2 dense contraction, 3 permutation batch,
 * 4 host<->device copies.  Returns how many slots were written. */
int hs_last_timers(hs_ctx* ctx, float* ms_out, int n_slots);
/* Enable (1) / disable (0) the per-call event timers above (they add two event records per
 * kernel family; off by default). */
int hs_enable_timers(hs_ctx* ctx, int on);

#ifdef __cplusplus
}
#endif
#endif /* HAYSTACK_B200_H */
