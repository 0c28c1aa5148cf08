// Shared definitions for libhaystack_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/haystack_b200.h"

#define HS_PSEUDO 1e-300  // R/haystack_continuous.R:184

// Slots of hs_last_timers
enum { HS_T_DENSITY = 0, HS_T_CSR = 1, HS_T_DENSE = 2, HS_T_PERM = 3, HS_T_COPY = 4, HS_T_COUNT = 5 };

template <typename T>
__host__ __device__ __forceinline__ T hs_min(T a, T b) { return a < b ? a : b; }
template <typename T>
__host__ __device__ __forceinline__ T hs_max(T a, T b) { return a > b ? a : b; }

// Row stride (floats) of the resident density matrix: G rounded up to a float4.  (Padding rows to whole
// 128-byte lines was tried: fewer cells fit a slab and K3 got 5 % slower, 76 -> 81 ms.)
__host__ __device__ __forceinline__ int hs_ld_for_grid(int n_grid) { return (n_grid + 3) & ~3; }

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};

struct hs_ctx {
  int device = 0;
  int sm_count = 148;
  size_t smem_optin = 0;
  cudaStream_t own_stream = nullptr;   // created by hs_create
  cudaStream_t stream = nullptr;       // stream all kernels run on (own_stream or external)
  cudaStream_t copy_stream = nullptr;  // H2D staging for the host-pointer entry points
  cudaEvent_t ev_a = nullptr, ev_b = nullptr;   // timers
  cudaEvent_t ev_stage[2] = {nullptr, nullptr}; // staging buffer k has landed
  cudaEvent_t ev_done[2] = {nullptr, nullptr};  // compute on staging buffer k finished
  std::string err;
  int64_t launches = 0;
  bool timers = false;
  float t_ms[HS_T_COUNT] = {0, 0, 0, 0, 0};

  // ---- resident density state (written by the density stage) ----
  int64_t N = 0;   // cells
  int G = 0;       // grid points
  int ldD = 0;     // row stride (floats) of D32, multiple of 4
  double bw = 0.0;
  DevBuf D32;      // [N][ldD] fp32, cell-major
  DevBuf Q;        // [G] fp64 normalised reference distribution

  // ---- scratch (grow-only) ----
  DevBuf sq;        // [G][N] fp64 squared distances
  DevBuf rowmin;    // [N] fp64
  DevBuf nearest;   // [N] int32
  DevBuf sel;       // radix-select state + histograms
  DevBuf qpart;     // [blocks][G] fp64
  DevBuf dens64;    // [G][N] fp64 (only when density_out requested)
  DevBuf xin, gin, win;        // uploaded x / grid / w_q
  DevBuf part;      // [rows][ldP] fp64 column sums P
  DevBuf simt_part; // [ksplit][rows][ldP] fp64 split-K partials of the SIMT dense kernel
  DevBuf out;       // fp64 results
  DevBuf stage_j[2], stage_x[2];  // H2D staging (int32 ids, fp64 values)
  DevBuf bits[2];                 // cell-id bitmaps of a piece of rows (hs_api.cu: packed id upload), double-buffered
  cudaEvent_t ev_bits[2] = {};    // the bitmaps in bits[k] have been expanded
  bool bits_used[2] = {};
  DevBuf vals32;    // fp32 values
  DevBuf ptrs;      // int64 row pointers
  DevBuf spmm_ws;   // sparse contraction: flags, sort keys, cub workspace
  DevBuf spmm_slots;   // slab kernel: slot table
  DevBuf cols32;    // uploaded column / cell ids (host-pointer entry points)
  DevBuf dtile_hi, dtile_lo;   // fp16 hi / lo density operands pre-tiled for the tensor-core kernels
  bool tiles_valid = false;    // the tiles match the resident density matrix
  DevBuf umma_flags;
  DevBuf prof_x, prof_out;     // hs_profile_*: uploaded fp64 values, gene x grid result
  DevBuf tile_ent, tile_ws;    // tile path of the sparse contraction: tile-major entries, cuts / region / tile tables
  DevBuf dbg_times;            // HS_TILE_DEBUG: per-CTA timestamps
  DevBuf zeros;                // a block of zeros (hs_tiles.cu: operand tiles are cleared by bulk copies of it)
  DevBuf dense32;              // uploaded dense expression as fp32 (column-major, ld = rows)
  DevBuf perm_rp, perm_cols, perm_vals;   // randomization batch: row pointers, (id, value) pairs, values of the selected rows
  DevBuf conv_cp, conv_ri, conv_x, conv_p, conv_j, conv_xo;   // hs_csc_to_csr (host form): inputs and outputs
  DevBuf conv_sort;            // hs_csc_to_csr on large inputs: records, sort buffers, per-chunk row offsets
  DevBuf perm_gp, out_rand;               // hs_dkl_csr_randomized: gene pointers and results of the randomization batch
  bool skip_copy_fence = false;           // the caller guarantees that nothing in flight reads the upload buffers
  void* pinned = nullptr;      // small pinned host bounce buffer
  size_t pinned_cap = 0;
  // host-side fp64 -> fp32 narrowing of uploaded values (hs_api.cu): pinned ring + worker threads
  static constexpr int kRing = 4;
  void* ring[kRing] = {};
  size_t ring_cap = 0;                                   // bytes per ring buffer
  cudaEvent_t ev_ring[kRing] = {};   // the copy out of ring buffer k has finished
  bool ring_busy[kRing] = {};
  struct HostPool* pool = nullptr;
  int64_t seed_n = 0;      // hs_grid_seed_*: cells / dims of the coordinates resident in xin
  int seed_d = 0;
  struct HsAsync* async = nullptr;   // hs_dkl_perm_csr_begin / _wait: the call in flight on the library's worker thread
};

// ---- error helpers (host) -----------------------------------------------------------------
int hs_fail(hs_ctx* ctx, int code, const char* fmt, ...);
int hs_ensure(hs_ctx* ctx, DevBuf& b, size_t bytes);

#define HS_CUDA(ctx, expr)                                                                  \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess)                                                                  \
      return hs_fail((ctx), HS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                     __FILE__, __LINE__);                                                   \
  } while (0)

#define HS_TRY(expr)              \
  do {                            \
    int _rc = (expr);             \
    if (_rc != HS_OK) return _rc; \
  } while (0)

#define HS_LAUNCH_CHECK(ctx)                                                                \
  do {                                                                                      \
    (ctx)->launches++;                                                                      \
    cudaError_t _e = cudaGetLastError();                                                    \
    if (_e != cudaSuccess)                                                                  \
      return hs_fail((ctx), HS_ERR_CUDA, "kernel launch failed: %s (%s:%d)",                \
                     cudaGetErrorString(_e), __FILE__, __LINE__);                           \
  } while (0)

// ---- launchers implemented in the kernel translation units --------------------------------
// density stage (hs_density.cu): inputs device-resident fp64 column-major
int hs_run_density(hs_ctx* ctx, const double* x_dev, int64_t N, int d, const double* grid_dev, int G,
                   const double* wq_dev, bool want_dens64);
// grid selection, "seeding" method: the distance passes (hs_density.cu)
int hs_run_seed_begin(hs_ctx* ctx, int64_t N, int d);
int hs_run_seed_step(hs_ctx* ctx, const double* x_dev, int64_t N, int d, int64_t index, const double** probs_dev_out,
                     const void** nearest_dev_out);
// contraction kernels (hs_contract.cu)
// slab-staged sparse contraction (hs_spmm.cu); see there for the arguments
int hs_run_slab_spmm(hs_ctx* ctx, bool perm, const int64_t* ptr_dev, int64_t n_units, int n_perm,
                     const int32_t* cols_dev, const float* vals_dev, int col_base, double* out_dev,
                     int out_n_perm, int64_t out_n_sel, bool allow_tiles, const int** flag_dev_out,
                     const void* pairs_dev = nullptr, int64_t nnz_hint = -1, bool tile_only = false);
// tensor-core tile path of the sparse contraction (hs_tiles.cu)
bool hs_tile_path_enabled(const hs_ctx* ctx);
int hs_run_tile_spmm(hs_ctx* ctx, const int64_t* ptr_dev, const int32_t* sorted_rows_dev, int64_t n_rows,
                     int64_t nnz_total, const int32_t* cols_dev, const float* vals_dev, const void* pairs_dev,
                     int col_base, double* out_dev, int out_n_perm, int64_t out_n_sel, int* flags);
const int32_t* hs_tile_row_order(hs_ctx* ctx, const int32_t* sorted_rows_dev, int64_t n_rows, int32_t* out_dev);
int hs_ensure_density_tiles(hs_ctx* ctx);   // hs_umma.cu: pre-tiled fp16 hi / lo density operands
int hs_run_perm_sorted(hs_ctx* ctx, const int64_t* gene_ptr_dev, int64_t n_sel, int64_t n_vals_total,
                       const float* val_dev, const int32_t* ind_dev, int n_perm, int index_base, double* out_dev,
                       const int** fallback_flag_out);
int hs_run_csc_to_csr_f32(hs_ctx* ctx, const int64_t* cp, const int32_t* ri, const float* x, int64_t n_rows, int64_t n_cols,
                          int64_t nnz, int64_t* p_out, int32_t* j_out, float* x_out);
int hs_run_bitmap_to_ids(hs_ctx* ctx, const void* bits_dev, int64_t words, const int64_t* ptr_dev, int64_t n_rows,
                         int32_t* cols_dev, int* flags_dev, cudaStream_t st);
int hs_run_csc_to_csr_f64(hs_ctx* ctx, const int64_t* cp, const int32_t* ri, const double* x, int64_t n_rows, int64_t n_cols,
                          int64_t nnz, int64_t* p_out, int32_t* j_out, double* x_out);
int hs_run_row_stats_f64(hs_ctx* ctx, const int64_t* p_dev, const double* x_dev, int64_t n_rows, int64_t n_cells,
                         double* mean_dev, double* sd_dev);
int hs_run_row_stats_f32(hs_ctx* ctx, const int64_t* p_dev, const float* x_dev, int64_t n_rows, int64_t n_cells,
                         double* mean_dev, double* sd_dev);
int hs_run_perm_dense_seeded(hs_ctx* ctx, const float* v_dev, int n_perm, uint64_t seed, uint64_t stream_base,
                             double* out_dev);
int hs_run_perm_seeded(hs_ctx* ctx, const int64_t* gene_ptr_dev, int64_t n_sel, int64_t n_vals_total,
                       const float* val_dev, int n_perm, uint64_t seed, uint64_t stream_base, double* out_dev,
                       const unsigned long long* unit_stream_dev = nullptr);
int hs_run_order_by_cv(hs_ctx* ctx, const double* mean_dev, const double* sd_dev, int64_t n, const int64_t* ranks_dev,
                       int64_t n_ranks, double* cv_dev_out, int64_t* rows_dev_out, int* flag_dev);
int hs_run_perm_tiles(hs_ctx* ctx, const int64_t* gene_ptr_dev, int64_t n_sel, int64_t n_vals_total, const float* val_dev,
                      int n_perm, uint64_t seed, uint64_t stream_base, const unsigned long long* unit_stream_dev,
                      const int32_t* sorted_rows_dev, double* out_dev, int* flags);
int hs_run_gather_seeded(hs_ctx* ctx, const int64_t* gene_ptr_dev, const float* val_dev, int64_t n_sel, int n_perm,
                         uint64_t seed, uint64_t stream_base, const unsigned long long* unit_stream_dev,
                         const int* run_if_flag, double* out_dev);
// density stage in shardable steps (hs_density.cu)
int hs_density_step_dist(hs_ctx* ctx, const double* x_dev, int64_t N, int d, const double* grid_dev, int G,
                         int64_t c_lo, int64_t c_hi, int world);
int hs_density_step_dens(hs_ctx* ctx, int64_t N, int G, const double* wq_dev, int64_t c_lo, int64_t c_hi,
                         bool want_dens64);
int hs_density_step_q(hs_ctx* ctx, int64_t N, int G);
int64_t hs_density_cells_per_rank(int64_t N, int world);
int hs_run_csr(hs_ctx* ctx, const int64_t* p_dev, const int32_t* j_dev, const float* x_dev,
               int64_t n_rows, double* out_dev);
int hs_run_perm_csr(hs_ctx* ctx, const int64_t* gene_ptr_dev, int64_t n_sel, int64_t n_vals_total,
                    const float* val_dev, const int32_t* ind_dev, int n_perm, int index_base, double* out_dev);
// tensor-core dense contraction (hs_umma.cu); -1 = shape not supported
int hs_run_dense_umma(hs_ctx* ctx, bool perm, const float* e_dev, int64_t ld, const float* v_dev,
                      const int32_t* perm_dev, int index_base, int64_t n_rows, double* out_dev,
                      const int** flag_dev_out);
bool hs_umma_disabled();   // HS_DISABLE_UMMA=1 forces the SIMT kernels (A/B comparisons)
int hs_run_profile(hs_ctx* ctx, bool dense, const int64_t* p_dev, const int32_t* j_dev, const double* x_dev, int64_t ld,
                   int64_t n_rows, double* out_dev);
int hs_run_dense_f32(hs_ctx* ctx, const float* e_dev, int64_t n_genes, int64_t ld, double* out_dev);
int hs_run_perm_dense(hs_ctx* ctx, const float* v_dev, const int32_t* perm_dev, int n_perm,
                      int index_base, double* out_dev);
int hs_run_f64_to_f32(hs_ctx* ctx, const double* src, float* dst, int64_t n, cudaStream_t stream);

#ifdef __CUDACC__
// ---- device helpers -----------------------------------------------------------------------
__device__ __forceinline__ double hs_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Fused KL epilogue for one gene held by one warp (R/haystack_continuous.R:358-360, :597-599).
// Lane l holds the raw column sums P[l + 32*a], a < NACC; columns >= G are ignored.
// Returns D_KL on every lane.  All arithmetic fp64, same operation order as the reference:
// P + pseudo; P / sum(P); sum(P * log(P / Q)).
template <int NACC>
__device__ __forceinline__ double hs_kl_epilogue(double (&p)[NACC], const double* __restrict__ Q, int G,
                                                 int lane) {
  double s = 0.0;
#pragma unroll
  for (int a = 0; a < NACC; ++a) {
    const int j = lane + 32 * a;
    if (j < G) {
      p[a] += HS_PSEUDO;
      s += p[a];
    }
  }
  s = hs_warp_sum(s);
  double kl = 0.0;
#pragma unroll
  for (int a = 0; a < NACC; ++a) {
    const int j = lane + 32 * a;
    if (j < G) {
      const double pn = p[a] / s;
      kl += pn * log(pn / Q[j]);
    }
  }
  return hs_warp_sum(kl);
}
#endif  // __CUDACC__
