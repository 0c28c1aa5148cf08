// K1: density stage of haystack_continuous_highD
//   dist  = get_dist_two_sets(x, grid)            R/haystack_highD.R:17-24, call R/haystack_continuous.R:168
//   bw    = median(apply(dist, 1, min))           R/haystack_continuous.R:174
//   D     = exp(-(dist/bw)^2 / 2)                 R/haystack_continuous.R:175-177
//   Q     = colSums(D [* w]) + 1e-300, / sum      R/haystack_continuous.R:179-186
// All arithmetic is fp64 (the bandwidth scales every exponent; see DESIGN.md "precision").
// The density matrix is left resident as fp32, cell-major [N][ldD], for the contraction kernels.
#include <math.h>

#include "hs_common.cuh"

namespace {

constexpr int kCellsPerBlock = 128;
constexpr int kJT = 8;  // grid points per register tile in the distance kernel: one x load feeds 8 independent
                        // (subtract, fused multiply-add) chains

// ---- squared distances, per-cell minimum and nearest grid point --------------------------------
// x: [N x d] column-major, grid: [G x d] column-major.  sq: [G][N] (column-major N x G, like R's
// dist.to.grid) so that both the write here and the read in k_density are coalesced over cells.
__global__ void __launch_bounds__(kCellsPerBlock)
k_sqdist(const double* __restrict__ x, int64_t N, int d, const double* __restrict__ grid, int G,
         int jchunk, int64_t c_lo, int64_t c_hi, double* __restrict__ sq, double* __restrict__ rowmin,
         int32_t* __restrict__ nearest) {
  // cells [c_lo, c_hi) of the N x d matrix; sq holds only this range, [G][c_hi - c_lo]
  extern __shared__ double gs[];  // [jchunk][d]
  const int64_t c = c_lo + (int64_t)blockIdx.x * kCellsPerBlock + threadIdx.x;
  const bool valid = c < c_hi;
  const int64_t cc = valid ? c : c_hi - 1;
  const int64_t n_loc = c_hi - c_lo;
  double best = INFINITY;
  int bestj = 0;
  for (int j0 = 0; j0 < G; j0 += jchunk) {
    const int jn = min(jchunk, G - j0);
    __syncthreads();
    for (int i = threadIdx.x; i < jn * d; i += kCellsPerBlock) {
      const int jj = i / d, k = i - jj * d;
      gs[i] = grid[(int64_t)(j0 + jj) + (int64_t)k * G];
    }
    __syncthreads();
    for (int jj = 0; jj < jn; jj += kJT) {
      double s[kJT];
      const double* g[kJT];
#pragma unroll
      for (int t = 0; t < kJT; ++t) {
        s[t] = 0.0;
        g[t] = gs + (size_t)min(jj + t, jn - 1) * d;
      }
      for (int k = 0; k < d; ++k) {
        const double xv = x[cc + (int64_t)k * N];
#pragma unroll
        for (int t = 0; t < kJT; ++t) {
          const double diff = xv - g[t][k];
          s[t] += diff * diff;  // (x-y)^2 summed in order, R/haystack_highD.R:8
        }
      }
#pragma unroll
      for (int t = 0; t < kJT; ++t) {
        if (jj + t < jn) {
          const int j = j0 + jj + t;
          if (valid) sq[(int64_t)j * n_loc + (c - c_lo)] = s[t];
          if (s[t] < best) {  // strict: first minimum wins
            best = s[t];
            bestj = j;
          }
        }
      }
    }
  }
  if (valid) {
    rowmin[c] = sqrt(best);
    nearest[c] = bestj;
  }
}

// ---- exact order statistics of rowmin by MSB-first radix select on the fp64 bit patterns ------
// (non-negative doubles order like their unsigned bit patterns).  Two ranks are tracked at once:
// the two middle order statistics R's median() averages for even N.
struct SelState {
  unsigned long long prefix[2];
  long long rank[2];
  unsigned int hist[2][256];
};

__global__ void k_sel_init(SelState* st, long long ra, long long rb) {
  const int t = threadIdx.x;
  if (t < 2) {
    st->prefix[t] = 0ull;
    st->rank[t] = t == 0 ? ra : rb;
  }
  for (int i = t; i < 512; i += blockDim.x) (&st->hist[0][0])[i] = 0u;
}

__global__ void __launch_bounds__(256)
k_sel_hist(const unsigned long long* __restrict__ keys, int64_t n, int shift, SelState* st) {
  __shared__ unsigned int h[2][256];
  h[0][threadIdx.x] = 0u;
  h[1][threadIdx.x] = 0u;
  __syncthreads();
  const unsigned long long himask = (shift + 8 >= 64) ? 0ull : (~0ull << (shift + 8));
  const unsigned long long pa = st->prefix[0], pb = st->prefix[1];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long k = keys[i];
    const unsigned int dig = (unsigned int)((k >> shift) & 0xFFull);
    if ((k & himask) == pa) atomicAdd(&h[0][dig], 1u);
    if ((k & himask) == pb) atomicAdd(&h[1][dig], 1u);
  }
  __syncthreads();
  if (h[0][threadIdx.x]) atomicAdd(&st->hist[0][threadIdx.x], h[0][threadIdx.x]);
  if (h[1][threadIdx.x]) atomicAdd(&st->hist[1][threadIdx.x], h[1][threadIdx.x]);
}

__global__ void k_sel_pick(SelState* st, int shift) {
  const int t = threadIdx.x;
  if (t < 2) {
    long long r = st->rank[t];
    int dig = 255;
    for (int b = 0; b < 256; ++b) {
      const long long cnt = st->hist[t][b];
      if (r < cnt) {
        dig = b;
        break;
      }
      r -= cnt;
    }
    st->rank[t] = r;
    st->prefix[t] |= ((unsigned long long)dig) << shift;
    for (int b = 0; b < 256; ++b) st->hist[t][b] = 0u;
  }
}

__global__ void k_sel_final(const SelState* st, double* bw) {
  const double a = __longlong_as_double((long long)st->prefix[0]);
  const double b = __longlong_as_double((long long)st->prefix[1]);
  *bw = (st->prefix[0] == st->prefix[1]) ? a : (a + b) / 2.0;  // mean of the two middle values
}

// ---- density + per-block column sums ------------------------------------------------------------
__global__ void __launch_bounds__(kCellsPerBlock)
k_density(const double* __restrict__ sq, int64_t N, int G, int ldD, int64_t c_lo, int64_t c_hi,
          const double* __restrict__ bw_dev, const double* __restrict__ wq, float* __restrict__ D32,
          double* __restrict__ dens64, double* __restrict__ qpart) {
  // 32 grid columns per pass: 17 KB of shared memory per block (a 128-column tile allowed only 3 blocks per SM,
  // too few warps to hide the latency of the fp64 exp / sqrt / divide chains)
  constexpr int kW = 32;
  __shared__ float tile[kCellsPerBlock][kW + 1];
  __shared__ double qs[4][kW];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t c = c_lo + (int64_t)blockIdx.x * kCellsPerBlock + tid;
  const bool valid = c < c_hi;
  const int64_t n_loc = c_hi - c_lo;
  const int64_t gblock = c_lo / kCellsPerBlock + blockIdx.x;      // block number among all cells
  const double bw = *bw_dev;
  const double w = (valid && wq) ? wq[c] : 1.0;
  for (int j0 = 0; j0 < ldD; j0 += kW) {
    const int jn = max(0, min(kW, G - j0));        // real columns in this chunk
    const int jpad = min(kW, ldD - j0);            // columns to store (zero padded)
    for (int jj = 0; jj < jn; ++jj) {
      double dv = 0.0;
      if (valid) {
        const double dist = sqrt(sq[(int64_t)(j0 + jj) * n_loc + (c - c_lo)]);
        const double dn = dist / bw;                // :175
        dv = exp(-dn * dn / 2);                     // :176-177
        if (dens64) dens64[(int64_t)(j0 + jj) * N + c] = dv;
      }
      tile[tid][jj] = (float)dv;
      const double qsum = hs_warp_sum(valid ? dv * w : 0.0);   // :180 / :182
      if (lane == 0) qs[warp][jj] = qsum;
    }
    for (int jj = jn; jj < jpad; ++jj) tile[tid][jj] = 0.f;
    __syncthreads();
    if (tid < jn) qpart[gblock * G + j0 + tid] = (qs[0][tid] + qs[1][tid]) + (qs[2][tid] + qs[3][tid]);
    for (int r = warp; r < kCellsPerBlock; r += 4) {
      const int64_t cr = c_lo + (int64_t)blockIdx.x * kCellsPerBlock + r;
      if (cr < c_hi && lane < jpad) D32[cr * ldD + j0 + lane] = tile[r][lane];
    }
    __syncthreads();
  }
}

// Q = (colsum + pseudo) / sum   (R/haystack_continuous.R:184-186); fixed summation order.
// Step 1: one block per grid column sums the per-block partials (strided, then a fixed tree).
__global__ void __launch_bounds__(256)
k_q_colsum(const double* __restrict__ qpart, int nblocks, int G, double* __restrict__ Q) {
  __shared__ double red[256];
  const int j = blockIdx.x;
  double local = 0.0;
  for (int b = threadIdx.x; b < nblocks; b += 256) local += qpart[(int64_t)b * G + j];
  red[threadIdx.x] = local;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) Q[j] = red[0] + HS_PSEUDO;
}

// Step 2: normalise.
__global__ void __launch_bounds__(256)
k_q_finalize(int G, double* __restrict__ Q) {
  __shared__ double red[256];
  double local = 0.0;
  for (int j = threadIdx.x; j < G; j += 256) local += Q[j];
  red[threadIdx.x] = local;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  const double total = red[0];
  for (int j = threadIdx.x; j < G; j += 256) Q[j] = Q[j] / total;
}

__global__ void k_f64_to_f32(const double* __restrict__ src, float* __restrict__ dst, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = (float)src[i];
}

}  // namespace

int hs_run_f64_to_f32(hs_ctx* ctx, const double* src, float* dst, int64_t n, cudaStream_t stream) {
  if (n <= 0) return HS_OK;
  const int blocks = (int)hs_min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 16);
  k_f64_to_f32<<<blocks, 256, 0, stream>>>(src, dst, n);
  HS_LAUNCH_CHECK(ctx);
  return HS_OK;
}

// The density stage in three steps, so that the cells can be sharded over ranks (one process per GPU): every rank
// runs steps 1 and 2 on its own cell range [c_lo, c_hi) (whole 128-cell blocks) and the host layer all-gathers the
// row minima after step 1 and the density rows + per-block column sums after step 2.  Every rank then holds exactly
// the buffers a single-GPU run would hold, and step 3 (fixed-order reductions) gives the same bits everywhere.
//   step 1  hs_density_step_dist:   squared distances, row minima and nearest grid point of [c_lo, c_hi)
//   step 2  hs_density_step_dens:   exact median of ALL row minima -> bandwidth; density rows + column sums of [c_lo, c_hi)
//   step 3  hs_density_step_q:      Q from all blocks' column sums; state becomes valid
int64_t hs_density_cells_per_rank(int64_t N, int world) {
  return ((N + world - 1) / world + kCellsPerBlock - 1) / kCellsPerBlock * kCellsPerBlock;
}
static int64_t padded_cells(int64_t N, int world) { return hs_density_cells_per_rank(N, world) * world; }

int hs_density_step_dist(hs_ctx* ctx, const double* x_dev, int64_t N, int d, const double* grid_dev, int G,
                         int64_t c_lo, int64_t c_hi, int world) {
  const int ldD = hs_ld_for_grid(G);
  const int64_t Npad = padded_cells(N, world);
  const int64_t n_loc = hs_max<int64_t>(c_hi - c_lo, 0);
  const int nblocks_all = (int)(Npad / kCellsPerBlock);
  HS_TRY(hs_ensure(ctx, ctx->sq, sizeof(double) * (size_t)hs_max<int64_t>(n_loc, 1) * G));
  HS_TRY(hs_ensure(ctx, ctx->rowmin, sizeof(double) * (size_t)Npad));
  HS_TRY(hs_ensure(ctx, ctx->nearest, sizeof(int32_t) * (size_t)Npad));
  HS_TRY(hs_ensure(ctx, ctx->sel, sizeof(SelState) + 64));
  HS_TRY(hs_ensure(ctx, ctx->qpart, sizeof(double) * (size_t)nblocks_all * G));
  HS_TRY(hs_ensure(ctx, ctx->D32, sizeof(float) * (size_t)Npad * ldD));
  HS_TRY(hs_ensure(ctx, ctx->Q, sizeof(double) * (size_t)G + 64));
  if (n_loc <= 0) return HS_OK;
  if (c_lo % kCellsPerBlock) return hs_fail(ctx, HS_ERR_INVALID, "internal: cell range must start on a 128-cell block");
  cudaStream_t st = ctx->stream;
  // grid chunk staged in shared memory: <= 64 KB
  int jchunk = (int)hs_min<int64_t>(G, hs_max<int64_t>(kJT, (64 * 1024) / (sizeof(double) * (size_t)d)));
  const size_t smem1 = sizeof(double) * (size_t)jchunk * d;
  if (smem1 > 200 * 1024) return hs_fail(ctx, HS_ERR_INVALID, "n_dims=%d too large for the distance kernel", d);
  HS_CUDA(ctx, cudaFuncSetAttribute(k_sqdist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
  const int nb = (int)((n_loc + kCellsPerBlock - 1) / kCellsPerBlock);
  k_sqdist<<<nb, kCellsPerBlock, smem1, st>>>(x_dev, N, d, grid_dev, G, jchunk, c_lo, c_hi, (double*)ctx->sq.p,
                                               (double*)ctx->rowmin.p, (int32_t*)ctx->nearest.p);
  HS_LAUNCH_CHECK(ctx);
  return HS_OK;
}

int hs_density_step_dens(hs_ctx* ctx, int64_t N, int G, const double* wq_dev, int64_t c_lo, int64_t c_hi,
                         bool want_dens64) {
  const int ldD = hs_ld_for_grid(G);
  cudaStream_t st = ctx->stream;
  // exact median of the row minima of all cells
  SelState* sel = (SelState*)ctx->sel.p;
  double* bw_dev = (double*)((char*)ctx->sel.p + sizeof(SelState));
  const long long ra = (N - 1) / 2, rb = N / 2;
  k_sel_init<<<1, 256, 0, st>>>(sel, ra, rb);
  HS_LAUNCH_CHECK(ctx);
  const int hblocks = (int)hs_min<int64_t>((N + 255) / 256, (int64_t)ctx->sm_count * 4);
  for (int shift = 56; shift >= 0; shift -= 8) {
    k_sel_hist<<<hblocks, 256, 0, st>>>((const unsigned long long*)ctx->rowmin.p, N, shift, sel);
    HS_LAUNCH_CHECK(ctx);
    k_sel_pick<<<1, 32, 0, st>>>(sel, shift);
    HS_LAUNCH_CHECK(ctx);
  }
  k_sel_final<<<1, 1, 0, st>>>(sel, bw_dev);
  HS_LAUNCH_CHECK(ctx);
  const int64_t n_loc = hs_max<int64_t>(c_hi - c_lo, 0);
  if (want_dens64) HS_TRY(hs_ensure(ctx, ctx->dens64, sizeof(double) * (size_t)N * G));
  if (n_loc > 0) {
    const int nb = (int)((n_loc + kCellsPerBlock - 1) / kCellsPerBlock);
    k_density<<<nb, kCellsPerBlock, 0, st>>>((const double*)ctx->sq.p, N, G, ldD, c_lo, c_hi, bw_dev, wq_dev,
                                                  (float*)ctx->D32.p, want_dens64 ? (double*)ctx->dens64.p : nullptr,
                                                  (double*)ctx->qpart.p);
    HS_LAUNCH_CHECK(ctx);
  }
  return HS_OK;
}

int hs_density_step_q(hs_ctx* ctx, int64_t N, int G) {
  cudaStream_t st = ctx->stream;
  const int nblocks = (int)((N + kCellsPerBlock - 1) / kCellsPerBlock);
  double* bw_dev = (double*)((char*)ctx->sel.p + sizeof(SelState));
  k_q_colsum<<<G, 256, 0, st>>>((const double*)ctx->qpart.p, nblocks, G, (double*)ctx->Q.p);
  HS_LAUNCH_CHECK(ctx);
  k_q_finalize<<<1, 256, 0, st>>>(G, (double*)ctx->Q.p);
  HS_LAUNCH_CHECK(ctx);
  // the bandwidth rides behind Q so that one collective moves both (hs_density_buffers)
  HS_CUDA(ctx, cudaMemcpyAsync((double*)ctx->Q.p + G, bw_dev, sizeof(double), cudaMemcpyDeviceToDevice, st));
  HS_CUDA(ctx, cudaMemcpyAsync(&ctx->bw, bw_dev, sizeof(double), cudaMemcpyDeviceToHost, st));
  ctx->N = N;
  ctx->G = G;
  ctx->ldD = hs_ld_for_grid(G);
  ctx->tiles_valid = false;
  return HS_OK;
}

int hs_run_density(hs_ctx* ctx, const double* x_dev, int64_t N, int d, const double* grid_dev, int G,
                   const double* wq_dev, bool want_dens64) {
  HS_TRY(hs_density_step_dist(ctx, x_dev, N, d, grid_dev, G, 0, N, 1));
  HS_TRY(hs_density_step_dens(ctx, N, G, wq_dev, 0, N, want_dens64));
  return hs_density_step_q(ctx, N, G);
}

// =================================================================================================================
// Grid selection, "seeding" method (R/haystack_highD.R:87-129): the K1-like passes on the device.
//   per new grid point (cell `index`):  tmp = get_dist_two_sets(input, grid[i, ])            :113
//                                       min.dist.to.grid = pmin(min.dist.to.grid, tmp)       :116
//                                       probs = min.dist.to.grid^2 (of the NEXT iteration)   :107
//                                       order(tmp)[1:3]: the three cells nearest to the new grid point, which the
//                                       final averaging (:124-128) replaces it by (ties: lower cell first)
// The weighted draw sample(N, 1, prob = probs) (:110) stays with the host: it consumes R's random stream.
// =================================================================================================================
namespace {

struct Near { double d; long long c; };
__device__ __forceinline__ bool near_less(const Near& a, const Near& b) { return a.d < b.d || (a.d == b.d && a.c < b.c); }
__device__ __forceinline__ Near near_min(Near a, Near b) { return near_less(b, a) ? b : a; }

// the three smallest (distance, cell) pairs of a block's values -> out[3]
__device__ void block_three_nearest(Near mine, Near* out) {
  __shared__ Near wbest[32];
  __shared__ Near pick;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  for (int r = 0; r < 3; ++r) {
    Near v = mine;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      Near u;
      u.d = __shfl_xor_sync(0xffffffffu, v.d, o);
      u.c = __shfl_xor_sync(0xffffffffu, v.c, o);
      v = near_min(v, u);
    }
    if (lane == 0) wbest[warp] = v;
    __syncthreads();
    if (warp == 0) {
      Near w = lane < nw ? wbest[lane] : Near{INFINITY, 0x7fffffffffffffffLL};
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        Near u;
        u.d = __shfl_xor_sync(0xffffffffu, w.d, o);
        u.c = __shfl_xor_sync(0xffffffffu, w.c, o);
        w = near_min(w, u);
      }
      if (lane == 0) {
        pick = w;
        out[r] = w;
      }
    }
    __syncthreads();
    if (mine.c == pick.c) mine = Near{INFINITY, 0x7fffffffffffffffLL};      // taken
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
k_seed_step(const double* __restrict__ x, int64_t N, int d, int64_t index, double* __restrict__ min_dist,
            double* __restrict__ probs, Near* __restrict__ cand) {
  extern __shared__ double xi[];      // coordinates of the new grid point
  for (int k = threadIdx.x; k < d; k += blockDim.x) xi[k] = x[index + (int64_t)k * N];
  __syncthreads();
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  Near mine{INFINITY, 0x7fffffffffffffffLL};
  if (c < N) {
    double acc = 0.0;
    for (int k = 0; k < d; ++k) {
      const double diff = x[c + (int64_t)k * N] - xi[k];
      acc = __dadd_rn(acc, __dmul_rn(diff, diff));      // (x - y)^2 summed in order, no contraction
    }
    const double dist = sqrt(acc);
    const double mn = fmin(min_dist[c], dist);
    min_dist[c] = mn;
    probs[c] = mn * mn;
    mine = Near{dist, (long long)c};
  }
  block_three_nearest(mine, cand + 3 * (size_t)blockIdx.x);
}

__global__ void __launch_bounds__(256)
k_seed_nearest(const Near* __restrict__ cand, int64_t n_cand, Near* __restrict__ out) {
  // the three smallest of all block candidates; candidates are distinct cells
  Near best[3] = {{INFINITY, 0x7fffffffffffffffLL}, {INFINITY, 0x7fffffffffffffffLL}, {INFINITY, 0x7fffffffffffffffLL}};
  __shared__ Near stage[3 * 256];
  for (int64_t i = threadIdx.x; i < n_cand; i += blockDim.x) {
    Near v = cand[i];
#pragma unroll
    for (int r = 0; r < 3; ++r)
      if (near_less(v, best[r])) {
        const Near t = best[r];
        best[r] = v;
        v = t;
      }
  }
  for (int r = 0; r < 3; ++r) stage[3 * threadIdx.x + r] = best[r];
  __syncthreads();
  if (threadIdx.x == 0) {
    Near top[3] = {{INFINITY, 0x7fffffffffffffffLL}, {INFINITY, 0x7fffffffffffffffLL}, {INFINITY, 0x7fffffffffffffffLL}};
    for (int i = 0; i < 3 * (int)blockDim.x; ++i) {
      Near v = stage[i];
#pragma unroll
      for (int r = 0; r < 3; ++r)
        if (near_less(v, top[r])) {
          const Near t = top[r];
          top[r] = v;
          v = t;
        }
    }
    for (int r = 0; r < 3; ++r) out[r] = top[r];
  }
}

__global__ void k_fill_inf(double* __restrict__ a, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = INFINITY;
}

}  // namespace

// x_dev: N x d fp64 column-major, resident for the whole selection (ctx->xin); min distances start at +Inf
int hs_run_seed_begin(hs_ctx* ctx, int64_t N, int d) {
  HS_TRY(hs_ensure(ctx, ctx->rowmin, sizeof(double) * (size_t)N));
  HS_TRY(hs_ensure(ctx, ctx->sq, sizeof(double) * (size_t)N));
  const int64_t n_blocks = (N + 255) / 256;
  HS_TRY(hs_ensure(ctx, ctx->qpart, sizeof(Near) * (size_t)(3 * n_blocks + 3)));
  k_fill_inf<<<(unsigned)n_blocks, 256, 0, ctx->stream>>>((double*)ctx->rowmin.p, N);
  HS_LAUNCH_CHECK(ctx);
  (void)d;
  return HS_OK;
}

// one new grid point = cell `index`: min distances and probs updated (probs_dev_out = ctx->sq), the three nearest cells
// of the new point written to nearest_dev (3 x {distance, cell})
int hs_run_seed_step(hs_ctx* ctx, const double* x_dev, int64_t N, int d, int64_t index, const double** probs_dev_out,
                     const void** nearest_dev_out) {
  const int64_t n_blocks = (N + 255) / 256;
  Near* cand = (Near*)ctx->qpart.p;
  k_seed_step<<<(unsigned)n_blocks, 256, sizeof(double) * (size_t)d, ctx->stream>>>(x_dev, N, d, index, (double*)ctx->rowmin.p,
                                                                                 (double*)ctx->sq.p, cand);
  HS_LAUNCH_CHECK(ctx);
  k_seed_nearest<<<1, 256, 0, ctx->stream>>>(cand, 3 * n_blocks, cand + 3 * n_blocks);
  HS_LAUNCH_CHECK(ctx);
  *probs_dev_out = (const double*)ctx->sq.p;
  *nearest_dev_out = cand + 3 * n_blocks;
  return HS_OK;
}
