// K2/K3/K4: gene x grid contraction with the fused Kullback-Leibler epilogue.
//   P     = colSums(D * w)              R/haystack_continuous.R:357   (dense row w)
//   P     = colSums(D[ind,] * val)      R/haystack_continuous.R:596   (sparse row)
//   D_KL  = sum(Pn * log(Pn / Q)), Pn = (P + 1e-300) / sum(P + 1e-300)   :358-360, :597-599
// Accumulation is two-level: fp32 FMAs over a short run of cells, flushed into fp64 running
// sums, so that the error of a 1e5-term column sum stays ~1e-8 relative (D_KL of weakly
// structured genes is ~ 1/2 sum (P-Q)^2/Q and amplifies column-sum error ~30x).  The epilogue
// is fp64 and never writes P to HBM.
#include <math.h>
#include <stdlib.h>

#include "hs_common.cuh"
#include "hs_prp.cuh"

namespace {

// =================================================================================================
// K3 / K4-sparse: one warp per row; lanes span grid columns; D rows gathered through L2.
// =================================================================================================
// PERM == false: CSR. row r owns cols[p[r] .. p[r+1]) and vals at the same offsets.
// PERM == true : row = s * n_perm + r; gene s owns vals[gp[s] .. gp[s+1]) (nnz_s values) and the
//                index block starting at cols[gp[s] * n_perm], column r of an nnz_s x n_perm
//                column-major matrix (R/haystack_continuous.R:275); output is n_sel x n_perm
//                column-major: out[s + r * n_sel].
template <int NACC, bool PERM>
__global__ void __launch_bounds__(256)
k_rows_gather_kl(const int64_t* __restrict__ ptr, const int32_t* __restrict__ cols,
                 const float* __restrict__ vals, int64_t n_rows, int n_perm, int64_t n_sel, int col_base,
                 const float* __restrict__ D, int ldD, int G, int64_t n_cells, const double* __restrict__ Q,
                 const int* __restrict__ run_if_flag, double* __restrict__ out) {
  // chained behind the slab kernel: only runs when that kernel met rows it cannot handle
  if (run_if_flag && !(*run_if_flag & 1)) return;
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t row = warp0; row < n_rows; row += nwarps) {
    const int32_t* rc;
    const float* rv;
    int64_t nnz;
    int64_t out_idx;
    if (PERM) {
      const int64_t s = row / n_perm;
      const int r = (int)(row - s * n_perm);
      const int64_t g0 = ptr[s];
      nnz = ptr[s + 1] - g0;
      rc = cols + g0 * n_perm + (int64_t)r * nnz;
      rv = vals + g0;
      out_idx = s + (int64_t)r * n_sel;
    } else {
      const int64_t g0 = ptr[row];
      nnz = ptr[row + 1] - g0;
      rc = cols + g0;
      rv = vals + g0;
      out_idx = row;
    }
    double acc[NACC];
#pragma unroll
    for (int a = 0; a < NACC; ++a) acc[a] = 0.0;
    for (int64_t base = 0; base < nnz; base += 32) {
      const int cnt = (int)hs_min<int64_t>(32, nnz - base);
      int myc = 0;
      float myv = 0.f;
      if (lane < cnt) {
        myc = rc[base + lane] - col_base;
        myv = rv[base + lane];
        // ids outside [0, n_cells) cannot be dereferenced: poison the result instead
        if (myc < 0 || (int64_t)myc >= n_cells) {
          myc = 0;
          myv = __int_as_float(0x7fc00000);
        }
      }
      float part[NACC];
#pragma unroll
      for (int a = 0; a < NACC; ++a) part[a] = 0.f;
#pragma unroll 4
      for (int t = 0; t < cnt; ++t) {
        const int c = __shfl_sync(0xffffffffu, myc, t);
        const float v = __shfl_sync(0xffffffffu, myv, t);
        const float* drow = D + (int64_t)c * ldD;
#pragma unroll
        for (int a = 0; a < NACC; ++a) {
          const int j = lane + 32 * a;
          const float dv = (j < G) ? __ldg(drow + j) : 0.f;
          part[a] = fmaf(v, dv, part[a]);
        }
      }
#pragma unroll
      for (int a = 0; a < NACC; ++a) acc[a] += (double)part[a];
    }
    const double kl = hs_kl_epilogue<NACC>(acc, Q, G, lane);
    if (lane == 0) out[out_idx] = kl;
  }
}

// Seeded randomization, order-independent form: row (unit s, randomization r) pairs the unit's i-th value with cell
// pi(i) of the keyed permutation (hs_prp.cuh), evaluated on the fly.  Chained behind the tile path of the seeded entry
// points (values beyond fp16, an overflowing tile window) and used directly where that path does not apply.
template <int NACC>
__global__ void __launch_bounds__(256)
k_rows_gather_seeded(const int64_t* __restrict__ gp, const float* __restrict__ vals, int64_t n_sel, int n_perm,
                     unsigned long long seed, unsigned long long stream_base,
                     const unsigned long long* __restrict__ unit_stream, const float* __restrict__ D, int ldD, int G,
                     int64_t n_cells, const double* __restrict__ Q, const int* __restrict__ run_if_flag,
                     double* __restrict__ out) {
  if (run_if_flag && !(*run_if_flag & 1)) return;
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t n_rows = n_sel * n_perm;
  for (int64_t row = warp0; row < n_rows; row += nwarps) {
    const int64_t s = row / n_perm;
    const int r = (int)(row - s * n_perm);
    const int64_t g0 = gp[s];
    const int64_t nnz = hs_min<int64_t>(gp[s + 1] - g0, n_cells);
    const uint64_t stream = unit_stream ? (uint64_t)unit_stream[s] + (uint64_t)r : (uint64_t)stream_base + (uint64_t)row;
    const HsPrp prp = hs_prp_make(seed, stream, (uint32_t)n_cells);
    double acc[NACC];
#pragma unroll
    for (int a = 0; a < NACC; ++a) acc[a] = 0.0;
    for (int64_t base = 0; base < nnz; base += 32) {
      const int cnt = (int)hs_min<int64_t>(32, nnz - base);
      uint32_t myc = 0;
      float myv = 0.f;
      if (lane < cnt) {
        myc = hs_prp_forward(prp, (uint32_t)(base + lane));
        myv = vals[g0 + base + lane];
      }
      float part[NACC];
#pragma unroll
      for (int a = 0; a < NACC; ++a) part[a] = 0.f;
#pragma unroll 4
      for (int t = 0; t < cnt; ++t) {
        const uint32_t c = __shfl_sync(0xffffffffu, myc, t);
        const float v = __shfl_sync(0xffffffffu, myv, t);
        const float* drow = D + (int64_t)c * ldD;
#pragma unroll
        for (int a = 0; a < NACC; ++a) {
          const int j = lane + 32 * a;
          const float dv = (j < G) ? __ldg(drow + j) : 0.f;
          part[a] = fmaf(v, dv, part[a]);
        }
      }
#pragma unroll
      for (int a = 0; a < NACC; ++a) acc[a] += (double)part[a];
    }
    const double kl = hs_kl_epilogue<NACC>(acc, Q, G, lane);
    if (lane == 0) out[s + (int64_t)r * n_sel] = kl;
  }
}

// =================================================================================================
// prepare_clustering (R/haystack_continuous.R:604-666): the same per-gene column sums, returned as a gene x grid
// matrix with every row rescaled to sum to 1 (:664) instead of being reduced to D_KL.  One warp per row; the values
// stay fp64 (as the host hands them over), products are fp32 FMAs in partials of 32 folded into fp64.
//   CSR  : row r owns cols / vals [ptr[r], ptr[r+1])          (extract_row_dgRMatrix, R/sparse.R:28-37)
//   dense: E[row + cell * ld]                                  (:647)
// out is n_rows x G column-major (an R matrix).
// =================================================================================================
template <int NACC, bool DENSE>
__global__ void __launch_bounds__(256)
k_rows_profile(const int64_t* __restrict__ ptr, const int32_t* __restrict__ cols, const double* __restrict__ vals,
               int64_t ld, int64_t n_rows, const float* __restrict__ D, int ldD, int G, int64_t n_cells,
               double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t row = warp0; row < n_rows; row += nwarps) {
    const int64_t g0 = DENSE ? 0 : ptr[row];
    const int64_t nnz = DENSE ? n_cells : ptr[row + 1] - g0;
    double acc[NACC];
#pragma unroll
    for (int a = 0; a < NACC; ++a) acc[a] = 0.0;
    for (int64_t base = 0; base < nnz; base += 32) {
      const int cnt = (int)hs_min<int64_t>(32, nnz - base);
      int64_t myc = 0;
      float myv = 0.f;
      if (lane < cnt) {
        if (DENSE) {
          myc = base + lane;
          myv = (float)vals[row + myc * ld];
        } else {
          myc = cols[g0 + base + lane];
          myv = (float)vals[g0 + base + lane];
          if (myc < 0 || myc >= n_cells) {      // cannot be dereferenced: poison the row
            myc = 0;
            myv = __int_as_float(0x7fc00000);
          }
        }
      }
      float part[NACC];
#pragma unroll
      for (int a = 0; a < NACC; ++a) part[a] = 0.f;
      for (int t = 0; t < cnt; ++t) {
        const int64_t c = __shfl_sync(0xffffffffu, myc, t);
        const float v = __shfl_sync(0xffffffffu, myv, t);
        if (DENSE && v == 0.f) continue;       // warp-uniform: zeros of a dense row add nothing
        const float* drow = D + c * ldD;
#pragma unroll
        for (int a = 0; a < NACC; ++a) {
          const int j = lane + 32 * a;
          part[a] = fmaf(v, (j < G) ? __ldg(drow + j) : 0.f, part[a]);
        }
      }
#pragma unroll
      for (int a = 0; a < NACC; ++a) acc[a] += (double)part[a];
    }
    double s = 0.0;
#pragma unroll
    for (int a = 0; a < NACC; ++a)
      if (lane + 32 * a < G) s += acc[a];
    s = hs_warp_sum(s);
#pragma unroll
    for (int a = 0; a < NACC; ++a) {
      const int j = lane + 32 * a;
      if (j < G) out[row + (int64_t)j * n_rows] = acc[a] / s;      // densities / rowSums(densities), :664
    }
  }
}

// =================================================================================================
// K2 / K4-dense: register-tiled fp32 contraction over cells with split-K partials in fp64.
// A(row, cell): dense  -> E[row + cell * ld]                       (R matrix, genes x cells)
//               perm   -> v[perm[cell + row * N] - base]           (sample(x = v), :259)
// Tile: 64 rows x 128 grid columns per block, 16 cells per stage, 256 threads each owning
// 4 rows x 8 columns.
// =================================================================================================
constexpr int kTM = 64, kTN = 128, kTK = 16, kFlush = 16;  // flush fp32 -> fp64 every 16 stages (256 cells)

template <typename TE, bool PERM>
__global__ void __launch_bounds__(256)
k_dense_partial(const TE* __restrict__ E, int64_t ld, const float* __restrict__ v,
                const int32_t* __restrict__ perm, int index_base, int64_t n_rows, int64_t N,
                const float* __restrict__ D, int ldD, int G, int cells_per_split, int ldP,
                const int* __restrict__ run_if_flag, double* __restrict__ part) {
  if (run_if_flag && !(*run_if_flag & 1)) return;   // chained behind the tensor-core kernel
  __shared__ __align__(16) float As[kTK][kTM + 4];
  __shared__ __align__(16) float Bs[kTK][kTN];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t r0 = (int64_t)blockIdx.x * kTM;
  const int j0 = blockIdx.z * kTN;
  const int64_t c_begin = (int64_t)blockIdx.y * cells_per_split;
  const int64_t c_end = hs_min<int64_t>(N, c_begin + cells_per_split);

  double acc[4][8];
  float cur[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      acc[i][j] = 0.0;
      cur[i][j] = 0.f;
    }

  int stage = 0;
  for (int64_t c0 = c_begin; c0 < c_end; c0 += kTK, ++stage) {
    // ---- stage A tile (16 cells x 64 rows) ----
    if (!PERM) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int e = tid + 256 * i;       // 0..1023
        const int kk = e >> 6, rr = e & 63;  // consecutive threads -> consecutive genes (coalesced)
        const int64_t cell = c0 + kk, row = r0 + rr;
        float a = 0.f;
        if (cell < c_end && row < n_rows) a = (float)E[row + cell * ld];
        As[kk][rr] = a;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int e = tid + 256 * i;
        const int rr = e >> 4, kk = e & 15;  // consecutive threads -> consecutive cells of one permutation
        const int64_t cell = c0 + kk, row = r0 + rr;
        float a = 0.f;
        if (cell < c_end && row < n_rows) {
          const int64_t src = (int64_t)perm[cell + row * N] - index_base;
          a = (src >= 0 && src < N) ? v[src] : __int_as_float(0x7fc00000);
        }
        As[kk][rr] = a;
      }
    }
    // ---- stage B tile (16 cells x 128 columns of D) ----
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int e = tid + 256 * i;         // 0..2047
      const int kk = e >> 7, jj = e & 127;
      const int64_t cell = c0 + kk;
      const int j = j0 + jj;
      Bs[kk][jj] = (cell < c_end && j < G) ? D[cell * ldD + j] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kTK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      float b[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) b[j] = Bs[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) cur[i][j] = fmaf(a[i], b[j], cur[i][j]);
    }
    __syncthreads();
    if ((stage % kFlush) == kFlush - 1) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[i][j] += (double)cur[i][j];
          cur[i][j] = 0.f;
        }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t row = r0 + ty * 4 + i;
    if (row >= n_rows) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = j0 + tx + 16 * j;
      if (col < ldP) part[((int64_t)blockIdx.y * n_rows + row) * ldP + col] = acc[i][j] + (double)cur[i][j];
    }
  }
}

// Sum split-K partials in fixed order and apply the KL epilogue; one warp per row.
template <int NACC>
__global__ void __launch_bounds__(256)
k_rows_epilogue(const double* __restrict__ part, int n_split, int64_t n_rows, int ldP, int G,
                const double* __restrict__ Q, const int* __restrict__ run_if_flag, double* __restrict__ out) {
  if (run_if_flag && !(*run_if_flag & 1)) return;
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_rows) return;
  double p[NACC];
#pragma unroll
  for (int a = 0; a < NACC; ++a) p[a] = 0.0;
  for (int s = 0; s < n_split; ++s) {
    const double* pr = part + ((int64_t)s * n_rows + row) * ldP;
#pragma unroll
    for (int a = 0; a < NACC; ++a) {
      const int j = lane + 32 * a;
      if (j < G) p[a] += pr[j];
    }
  }
  const double kl = hs_kl_epilogue<NACC>(p, Q, G, lane);
  if (lane == 0) out[row] = kl;
}

template <bool PERM>
int launch_rows_gather(hs_ctx* ctx, const int64_t* ptr, const int32_t* cols, const float* vals,
                       int64_t n_rows, int n_perm, int64_t n_sel, int col_base, const int* run_if_flag, double* out) {
  if (n_rows <= 0) return HS_OK;
  const int G = ctx->G;
  const int nacc = (G + 31) / 32;
  const int64_t want_blocks = (n_rows + 7) / 8;  // 8 warps per block
  const int blocks = (int)hs_min<int64_t>(want_blocks, (int64_t)ctx->sm_count * 8);
  const float* D = (const float*)ctx->D32.p;
  const double* Q = (const double*)ctx->Q.p;
#define HS_GATHER_CASE(NA)                                                                          \
  case NA:                                                                                          \
    k_rows_gather_kl<NA, PERM><<<blocks, 256, 0, ctx->stream>>>(ptr, cols, vals, n_rows, n_perm, n_sel, \
                                                                col_base, D, ctx->ldD, G, ctx->N, Q, run_if_flag, out); \
    break;
  switch (nacc) {
    HS_GATHER_CASE(1)
    HS_GATHER_CASE(2)
    HS_GATHER_CASE(3)
    HS_GATHER_CASE(4)
    HS_GATHER_CASE(5)
    HS_GATHER_CASE(6)
    HS_GATHER_CASE(7)
    HS_GATHER_CASE(8)
    default:
      if (nacc <= 16) {
        k_rows_gather_kl<16, PERM><<<blocks, 256, 0, ctx->stream>>>(ptr, cols, vals, n_rows, n_perm, n_sel,
                                                                    col_base, D, ctx->ldD, G, ctx->N, Q, run_if_flag, out);
      } else if (nacc <= 32) {
        k_rows_gather_kl<32, PERM><<<blocks, 256, 0, ctx->stream>>>(ptr, cols, vals, n_rows, n_perm, n_sel,
                                                                    col_base, D, ctx->ldD, G, ctx->N, Q, run_if_flag, out);
      } else {
        return hs_fail(ctx, HS_ERR_INVALID, "grid.points=%d exceeds the supported maximum of 1024", G);
      }
  }
#undef HS_GATHER_CASE
  HS_LAUNCH_CHECK(ctx);
  return HS_OK;
}

int launch_rows_epilogue(hs_ctx* ctx, const double* part, int n_split, int64_t n_rows, int ldP, const int* run_if_flag,
                         double* out) {
  const int G = ctx->G;
  const int nacc = (G + 31) / 32;
  const int blocks = (int)((n_rows + 7) / 8);
  const double* Q = (const double*)ctx->Q.p;
#define HS_EPI_CASE(NA)                                                                               \
  k_rows_epilogue<NA><<<blocks, 256, 0, ctx->stream>>>(part, n_split, n_rows, ldP, G, Q, run_if_flag, out)
  if (nacc <= 1) HS_EPI_CASE(1);
  else if (nacc <= 2) HS_EPI_CASE(2);
  else if (nacc <= 3) HS_EPI_CASE(3);
  else if (nacc <= 4) HS_EPI_CASE(4);
  else if (nacc <= 8) HS_EPI_CASE(8);
  else if (nacc <= 16) HS_EPI_CASE(16);
  else if (nacc <= 32) HS_EPI_CASE(32);
  else return hs_fail(ctx, HS_ERR_INVALID, "grid.points=%d exceeds the supported maximum of 1024", G);
#undef HS_EPI_CASE
  HS_LAUNCH_CHECK(ctx);
  return HS_OK;
}

template <typename TE, bool PERM>
int launch_dense(hs_ctx* ctx, const TE* E, int64_t ld, const float* v, const int32_t* perm, int index_base,
                 int64_t n_rows, const int* run_if_flag, double* out) {
  if (n_rows <= 0) return HS_OK;
  const int G = ctx->G;
  const int64_t N = ctx->N;
  const int ldP = (G + 3) & ~3;
  const int row_tiles = (int)((n_rows + kTM - 1) / kTM);
  const int col_tiles = (G + kTN - 1) / kTN;
  // split the cell axis until the launch fills ~2 waves; every split covers a multiple of 256 cells
  int64_t max_split = (N + 255) / 256;
  int64_t want = ((int64_t)ctx->sm_count * 2 + (int64_t)row_tiles * col_tiles - 1) / ((int64_t)row_tiles * col_tiles);
  int n_split = (int)hs_max<int64_t>(1, hs_min<int64_t>(hs_min<int64_t>(want, max_split), 65535));
  int64_t cells_per_split = ((N + n_split - 1) / n_split + 255) / 256 * 256;
  n_split = (int)((N + cells_per_split - 1) / cells_per_split);
  HS_TRY(hs_ensure(ctx, ctx->simt_part, sizeof(double) * (size_t)n_split * n_rows * ldP));
  dim3 grid(row_tiles, n_split, col_tiles);
  k_dense_partial<TE, PERM><<<grid, 256, 0, ctx->stream>>>(E, ld, v, perm, index_base, n_rows, N,
                                                            (const float*)ctx->D32.p, ctx->ldD, G,
                                                            (int)cells_per_split, ldP, run_if_flag,
                                                            (double*)ctx->simt_part.p);
  HS_LAUNCH_CHECK(ctx);
  return launch_rows_epilogue(ctx, (const double*)ctx->simt_part.p, n_split, n_rows, ldP, run_if_flag, out);
}

}  // namespace

int hs_run_profile(hs_ctx* ctx, bool dense, const int64_t* p_dev, const int32_t* j_dev, const double* x_dev, int64_t ld,
                   int64_t n_rows, double* out_dev) {
  if (n_rows <= 0) return HS_OK;
  const int G = ctx->G;
  const int nacc = (G + 31) / 32;
  const int blocks = (int)hs_min<int64_t>((n_rows + 7) / 8, (int64_t)ctx->sm_count * 8);
  const float* D = (const float*)ctx->D32.p;
#define HS_PROFILE(NA)                                                                                              \
  do {                                                                                                              \
    if (dense) k_rows_profile<NA, true><<<blocks, 256, 0, ctx->stream>>>(p_dev, j_dev, x_dev, ld, n_rows, D, ctx->ldD, G, ctx->N, out_dev); \
    else k_rows_profile<NA, false><<<blocks, 256, 0, ctx->stream>>>(p_dev, j_dev, x_dev, ld, n_rows, D, ctx->ldD, G, ctx->N, out_dev);     \
  } while (0)
  if (nacc <= 1) HS_PROFILE(1);
  else if (nacc <= 2) HS_PROFILE(2);
  else if (nacc <= 4) HS_PROFILE(4);
  else if (nacc <= 8) HS_PROFILE(8);
  else if (nacc <= 16) HS_PROFILE(16);
  else if (nacc <= 32) HS_PROFILE(32);
  else return hs_fail(ctx, HS_ERR_INVALID, "grid.points=%d exceeds the supported maximum of 1024", G);
#undef HS_PROFILE
  HS_LAUNCH_CHECK(ctx);
  return HS_OK;
}

int hs_run_csr(hs_ctx* ctx, const int64_t* p_dev, const int32_t* j_dev, const float* x_dev, int64_t n_rows,
               double* out_dev) {
  if (n_rows <= 0) return HS_OK;
  // slab-staged kernel; rows it flags (ids not ascending / negative) are redone by the gather kernel,
  // chained on the device-side flag so no host round trip is needed
  const int* flag = nullptr;
  HS_TRY(hs_run_slab_spmm(ctx, false, p_dev, n_rows, 1, j_dev, x_dev, 0, out_dev, 0, 0, /*allow_tiles=*/true, &flag));
  return launch_rows_gather<false>(ctx, p_dev, j_dev, x_dev, n_rows, 1, n_rows, 0, flag, out_dev);
}

int hs_run_gather_seeded(hs_ctx* ctx, const int64_t* gene_ptr_dev, const float* val_dev, int64_t n_sel, int n_perm,
                         uint64_t seed, uint64_t stream_base, const unsigned long long* unit_stream_dev,
                         const int* run_if_flag, double* out_dev) {
  const int64_t n_rows = n_sel * (int64_t)n_perm;
  if (n_rows <= 0) return HS_OK;
  const int G = ctx->G;
  const int nacc = (G + 31) / 32;
  if (nacc > 4) return hs_fail(ctx, HS_ERR_INVALID, "internal: seeded gather is the fallback of the tile path (grid.points <= 128)");
  const int blocks = (int)hs_min<int64_t>((n_rows + 7) / 8, (int64_t)ctx->sm_count * 8);
  const float* D = (const float*)ctx->D32.p;
  const double* Q = (const double*)ctx->Q.p;
#define HS_SEEDED_CASE(NA)                                                                                              \
  case NA:                                                                                                              \
    k_rows_gather_seeded<NA><<<blocks, 256, 0, ctx->stream>>>(gene_ptr_dev, val_dev, n_sel, n_perm, seed, stream_base,    \
                                                              unit_stream_dev, D, ctx->ldD, G, ctx->N, Q, run_if_flag,   \
                                                              out_dev);                                                  \
    break;
  switch (nacc) {
    HS_SEEDED_CASE(1)
    HS_SEEDED_CASE(2)
    HS_SEEDED_CASE(3)
    HS_SEEDED_CASE(4)
  }
#undef HS_SEEDED_CASE
  HS_LAUNCH_CHECK(ctx);
  return HS_OK;
}

int hs_run_perm_csr(hs_ctx* ctx, const int64_t* gene_ptr_dev, int64_t n_sel, int64_t n_vals_total,
                    const float* val_dev, const int32_t* ind_dev, int n_perm, int index_base, double* out_dev) {
  // bucket the ids by slab and use the slab kernel; gather straight from L2/HBM only when the cell
  // axis is too long for the bucketing histogram
  const int* flag = nullptr;
  const int rc = hs_run_perm_sorted(ctx, gene_ptr_dev, n_sel, n_vals_total, val_dev, ind_dev, n_perm, index_base, out_dev,
                                    &flag);
  if (rc != -1) {
    if (rc != HS_OK || !flag) return rc;
    // tensor-core tile path: rows it could not take (ids outside [0, N), values beyond fp16) are gathered instead
    return launch_rows_gather<true>(ctx, gene_ptr_dev, ind_dev, val_dev, n_sel * (int64_t)n_perm, n_perm, n_sel,
                                    index_base, flag, out_dev);
  }
  return launch_rows_gather<true>(ctx, gene_ptr_dev, ind_dev, val_dev, n_sel * (int64_t)n_perm, n_perm, n_sel,
                                  index_base, nullptr, out_dev);
}

// Dense forms: tensor-core kernel first; the fp32 SIMT kernel is chained behind it and only does work
// when the tensor-core kernel flagged an fp16 overflow (or cannot take the shape / is disabled).
bool hs_umma_disabled() {
  const char* e = getenv("HS_DISABLE_UMMA");
  return e && *e && *e != '0';
}

int hs_run_dense_f32(hs_ctx* ctx, const float* e_dev, int64_t n_genes, int64_t ld, double* out_dev) {
  const int* flag = nullptr;
  if (!hs_umma_disabled()) {
    const int rc = hs_run_dense_umma(ctx, false, e_dev, ld, nullptr, nullptr, 0, n_genes, out_dev, &flag);
    if (rc != HS_OK && rc != -1) return rc;
  }
  return launch_dense<float, false>(ctx, e_dev, ld, nullptr, nullptr, 0, n_genes, flag, out_dev);
}

int hs_run_perm_dense(hs_ctx* ctx, const float* v_dev, const int32_t* perm_dev, int n_perm, int index_base,
                      double* out_dev) {
  const int* flag = nullptr;
  if (!hs_umma_disabled()) {
    const int rc = hs_run_dense_umma(ctx, true, nullptr, 0, v_dev, perm_dev, index_base, n_perm, out_dev, &flag);
    if (rc != HS_OK && rc != -1) return rc;
  }
  return launch_dense<float, true>(ctx, (const float*)nullptr, 0, v_dev, perm_dev, index_base, n_perm, flag, out_dev);
}


// ---- K4-dense, seeded form ---------------------------------------------------------------------------
// sample(x = row) == row[sample.int(N)] (R/haystack_continuous.R:259): randomization r reads cell pi_r(c) for
// cell c, pi_r the keyed permutation of (seed, stream_base + r) (hs_prp.cuh).  The ids are written once into
// device scratch (N x n_perm int32, 0-based) and scored by the ids-given kernels, so the tensor-core path,
// its overflow fallback and the G > 128 kernel are shared; what the seeded form removes is the 4 N n_perm
// bytes per gene that would otherwise cross PCIe (400 MB at 1M cells x 100 randomizations).
__global__ void __launch_bounds__(256)
k_perm_ids_dense(int64_t N, uint64_t seed, uint64_t stream_base, int32_t* __restrict__ perm) {
  const int r = blockIdx.y;
  const HsPrp prp = hs_prp_make(seed, stream_base + (uint64_t)r, (uint32_t)N);
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < N; c += (int64_t)gridDim.x * blockDim.x)
    perm[c + (int64_t)r * N] = (int32_t)hs_prp_forward(prp, (uint32_t)c);
}

int hs_run_perm_dense_seeded(hs_ctx* ctx, const float* v_dev, int n_perm, uint64_t seed, uint64_t stream_base,
                             double* out_dev) {
  const int64_t N = ctx->N;
  HS_TRY(hs_ensure(ctx, ctx->stage_j[1], sizeof(int32_t) * (size_t)N * (size_t)n_perm));
  int32_t* ids = (int32_t*)ctx->stage_j[1].p;
  const unsigned bx = (unsigned)hs_min<int64_t>((N + 255) / 256, 4096);
  k_perm_ids_dense<<<dim3(bx, (unsigned)n_perm), 256, 0, ctx->stream>>>(N, seed, stream_base, ids);
  HS_LAUNCH_CHECK(ctx);
  return hs_run_perm_dense(ctx, v_dev, ids, n_perm, 0, out_dev);
}


// ---- row statistics (R/haystack_continuous.R:57-74) --------------------------------------------------
// Matrix::rowMeans and the per-row sample standard deviation (n - 1 denominator, the implicit zeros of the
// sparse row counted), the inputs of coeffVar and of the integer gene selection (:222-236).  Two passes in
// fp64, one warp per row: mean = sum(x) / N, then ss = sum((x - mean)^2) + (N - nnz) * mean^2.
template <typename T>
__global__ void __launch_bounds__(256)
k_row_stats_csr(const int64_t* __restrict__ p, const T* __restrict__ x, int64_t n_rows, int64_t n_cells,
                double* __restrict__ mean_out, double* __restrict__ sd_out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int64_t lo = p[row], hi = p[row + 1];
  double s = 0.0;
  for (int64_t i = lo + lane; i < hi; i += 32) s += (double)x[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const double mean = s / (double)n_cells;
  double ss = 0.0;
  for (int64_t i = lo + lane; i < hi; i += 32) {
    const double d = (double)x[i] - mean;
    ss += d * d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if (lane == 0) {
    ss += (double)(n_cells - (hi - lo)) * mean * mean;
    mean_out[row] = mean;
    sd_out[row] = n_cells > 1 ? sqrt(ss / (double)(n_cells - 1)) : nan("");
  }
}

int hs_run_row_stats_f64(hs_ctx* ctx, const int64_t* p_dev, const double* x_dev, int64_t n_rows, int64_t n_cells,
                         double* mean_dev, double* sd_dev) {
  if (n_rows <= 0) return HS_OK;
  k_row_stats_csr<double><<<(unsigned)((n_rows + 7) / 8), 256, 0, ctx->stream>>>(p_dev, x_dev, n_rows, n_cells, mean_dev, sd_dev);
  HS_LAUNCH_CHECK(ctx);
  return HS_OK;
}

int hs_run_row_stats_f32(hs_ctx* ctx, const int64_t* p_dev, const float* x_dev, int64_t n_rows, int64_t n_cells,
                         double* mean_dev, double* sd_dev) {
  if (n_rows <= 0) return HS_OK;
  k_row_stats_csr<float><<<(unsigned)((n_rows + 7) / 8), 256, 0, ctx->stream>>>(p_dev, x_dev, n_rows, n_cells, mean_dev, sd_dev);
  HS_LAUNCH_CHECK(ctx);
  return HS_OK;
}
