// CSC -> CSR on the device: as(expression, "RsparseMatrix") (R/haystack_continuous.R:113-116), the step
// before the sparse scoring path (users hold a dgCMatrix; the path wants rows).
//
// A stable counting sort by row id that needs no atomics for the order: the columns are cut into blocks,
// one CTA per block.
//   1. k_csc_count     per block, how many entries every row has (shared-memory counters);
//   2. k_csc_offsets   per row, exclusive prefix over the blocks -> where each block's entries of that row
//                      start inside the row; the row totals give the CSR row pointers (cub exclusive sum);
//   3. k_csc_scatter   per block, the columns are walked IN ORDER; the entries of one column belong to
//                      distinct rows, so the threads of the CTA take them in parallel and bump the row's
//                      cursor with a plain read-modify-write.  Column ids therefore come out ascending in
//                      every row (what the slab kernel and R's dgRMatrix both want), bit-reproducibly.
// Row counters live in shared memory (4 bytes per row): rows are taken in bands of 56 000, one pass over the
// entries per band -- a single pass for any gene x cell matrix.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <stdlib.h>

#include <vector>
#include <stdint.h>

#include "hs_common.cuh"

namespace {

constexpr int kConvThreads = 512;

template <typename P>
__global__ void __launch_bounds__(kConvThreads)
k_csc_count(const P* __restrict__ cp, const int32_t* __restrict__ ri, int64_t n_rows, int64_t r0, int64_t r1, int64_t n_cols,
            int cols_per_block, uint32_t* __restrict__ hist /* [blocks][n_rows] */, int* __restrict__ flags) {
  // rows [r0, r1): the band of rows whose counters fit into shared memory (one launch per band)
  extern __shared__ uint32_t cnt[];
  const int64_t nb = r1 - r0;
  for (int64_t r = threadIdx.x; r < nb; r += kConvThreads) cnt[r] = 0;
  __syncthreads();
  const int64_t c0 = (int64_t)blockIdx.x * cols_per_block, c1 = hs_min<int64_t>(n_cols, c0 + cols_per_block);
  const int64_t e0 = (int64_t)cp[c0], e1 = (int64_t)cp[c1];
  bool bad = false;
  for (int64_t e = e0 + threadIdx.x; e < e1; e += kConvThreads) {
    const int64_t r = ri[e];
    if (r < 0 || r >= n_rows) bad = true; else if (r >= r0 && r < r1) atomicAdd(&cnt[r - r0], 1u);
  }
  if (bad) atomicOr(flags, 1);
  __syncthreads();
  uint32_t* out = hist + (int64_t)blockIdx.x * n_rows + r0;
  for (int64_t r = threadIdx.x; r < nb; r += kConvThreads) out[r] = cnt[r];
  // (a block past the last column cannot exist: n_blocks is derived from cols_per_block)
}

// hist[b][r] := number of entries of row r in blocks < b; total[r] := row length
__global__ void __launch_bounds__(256)
k_csc_offsets(uint32_t* __restrict__ hist, int n_blocks, int64_t n_rows, int64_t* __restrict__ total) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r > n_rows) return;
  if (r == n_rows) {
    total[r] = 0;   // slot of the exclusive sum's last element
    return;
  }
  uint32_t run = 0;
  for (int b = 0; b < n_blocks; ++b) {
    const uint32_t c = hist[(int64_t)b * n_rows + r];
    hist[(int64_t)b * n_rows + r] = run;
    run += c;
  }
  total[r] = (int64_t)run;
}

template <typename P, typename TI, typename TO>
__global__ void __launch_bounds__(kConvThreads)
k_csc_scatter(const P* __restrict__ cp, const int32_t* __restrict__ ri, const TI* __restrict__ x, int64_t n_rows,
              int64_t r0, int64_t r1, int64_t n_cols, int cols_per_block, const uint32_t* __restrict__ hist,
              const int64_t* __restrict__ p_out, int32_t* __restrict__ j_out, TO* __restrict__ x_out) {
  extern __shared__ uint32_t pos[];   // cursor of every row of the band inside the row, for this block
  const uint32_t* mine = hist + (int64_t)blockIdx.x * n_rows + r0;
  const int64_t nb = r1 - r0;
  for (int64_t r = threadIdx.x; r < nb; r += kConvThreads) pos[r] = mine[r];
  __syncthreads();
  const int64_t c0 = (int64_t)blockIdx.x * cols_per_block, c1 = hs_min<int64_t>(n_cols, c0 + cols_per_block);
  for (int64_t c = c0; c < c1; ++c) {
    const int64_t e0 = (int64_t)cp[c], e1 = (int64_t)cp[c + 1];
    for (int64_t e = e0 + threadIdx.x; e < e1; e += kConvThreads) {
      const int64_t r = ri[e];
      if (r < r0 || r >= r1) continue;               // another band's row (ids outside [0, n_rows): flagged by the counting pass)
      const uint32_t k = pos[r - r0];                // rows are distinct inside a column: no race
      pos[r - r0] = k + 1;
      const int64_t o = p_out[r] + k;
      j_out[o] = (int32_t)c;
      x_out[o] = (TO)x[e];
    }
    __syncthreads();                                 // the next column may revisit the same rows
  }
}

// ---- large matrices: placement by a stable radix sort -----------------------------------------------------------------
// k_csc_scatter writes 4-byte values to places scattered over the whole output; once the output is far larger than L2
// every such store costs a partial-sector round trip to DRAM (1M cells x 30k genes x 10 %: 240 ms, 3 % of the HBM
// roofline).  For large inputs the entries are therefore taken in chunks of whole column blocks: each chunk is zipped
// into (row id | column id, value) records, sorted by row id with cub's radix sort -- stable, so the columns stay
// ascending inside a row -- and copied row by row, in coalesced runs, to where the counting pass says the chunk's part
// of that row belongs.  Same output, bit for bit.
template <typename TI> struct ColVal { int32_t c; TI x; };

template <typename P, typename TI>
__global__ void __launch_bounds__(256)
k_csc_zip(const P* __restrict__ cp, const TI* __restrict__ x, int64_t c0, int64_t c1, int64_t e0,
          ColVal<TI>* __restrict__ rec) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t c = c0 + warp0; c < c1; c += nwarps) {
    const int64_t a = (int64_t)cp[c], b = (int64_t)cp[c + 1];
    for (int64_t e = a + lane; e < b; e += 32) {
      ColVal<TI> r;
      r.c = (int32_t)c;
      r.x = x[e];
      rec[e - e0] = r;
    }
  }
}

template <typename P>
__global__ void k_csc_block_starts(const P* __restrict__ cp, int64_t n_cols, int cols_per_block, int n_blocks,
                                   int64_t* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b <= n_blocks) out[b] = (int64_t)cp[hs_min<int64_t>(n_cols, (int64_t)b * cols_per_block)];
}

// cnt[r] = entries of row r inside column blocks [b0, b1)
__global__ void k_csc_chunk_count(const uint32_t* __restrict__ hist, const int64_t* __restrict__ total, int64_t n_rows,
                                  int b0, int b1, int n_blocks, int64_t* __restrict__ cnt) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r > n_rows) return;
  if (r == n_rows) {
    cnt[r] = 0;
    return;
  }
  const int64_t hi = b1 < n_blocks ? (int64_t)hist[(int64_t)b1 * n_rows + r] : total[r];
  cnt[r] = hi - (int64_t)hist[(int64_t)b0 * n_rows + r];
}

// row r of the chunk: sorted records [start[r], start[r + 1]) -> j_out / x_out at p_out[r] + hist[b0][r]
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
k_csc_place(const ColVal<TI>* __restrict__ rec, const int64_t* __restrict__ start, const uint32_t* __restrict__ hist_b0,
            const int64_t* __restrict__ p_out, int64_t n_rows, int32_t* __restrict__ j_out, TO* __restrict__ x_out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp0; r < n_rows; r += nwarps) {
    const int64_t a = start[r], b = start[r + 1];
    const int64_t dst = p_out[r] + (int64_t)hist_b0[r];
    for (int64_t i = a + lane; i < b; i += 32) {
      const ColVal<TI> v = rec[i];
      j_out[dst + (i - a)] = v.c;
      x_out[dst + (i - a)] = (TO)v.x;
    }
  }
}

// ---- cell ids packed as one bit per cell ------------------------------------------------------------------------------
// The upload of a sparse matrix is bound by PCIe, and at more than ~3 % stored values a row's ascending cell ids
// (dgRMatrix@j, 4 bytes each) are larger than the row as a bitmap over the cells (n_cells / 8 bytes: 1.25 bytes per
// id at 10 %).  hs_api.cu packs such rows on the host while earlier buffers are on the link; this kernel restores the
// ids: row r of the piece owns words [r * words, (r + 1) * words) and the ids go to cols[ptr[r] ...), ascending.
// One CTA per row: popcount per 64-bit word, block scan, every thread expands its word.
__global__ void __launch_bounds__(256)
k_bitmap_to_ids(const unsigned long long* __restrict__ bits, int64_t words, const int64_t* __restrict__ ptr,
                int32_t* __restrict__ cols, int* __restrict__ flags) {
  __shared__ unsigned int wsum[8];
  __shared__ unsigned int carry;
  const int row = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t b = ptr[row];
  const unsigned int nnz = (unsigned int)(ptr[row + 1] - b);
  const unsigned long long* w = bits + (int64_t)row * words;
  if (threadIdx.x == 0) carry = 0u;
  __syncthreads();
  for (int64_t w0 = 0; w0 < words; w0 += 256) {
    const int64_t wi = w0 + threadIdx.x;
    unsigned long long x = wi < words ? __ldg(w + wi) : 0ull;
    const unsigned int c = (unsigned int)__popcll(x);
    unsigned int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int u = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += u;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    unsigned int before = carry;
    for (int k = 0; k < warp; ++k) before += wsum[k];
    unsigned int at = before + incl - c;
    const int32_t id0 = (int32_t)(wi * 64);
    while (x) {
      const int bit = __ffsll((long long)x) - 1;
      x &= x - 1;
      if (at < nnz) cols[b + at] = id0 + bit;
      ++at;
    }
    __syncthreads();
    if (threadIdx.x == 255) carry = before + incl;
    __syncthreads();
  }
  if (flags && threadIdx.x == 0 && carry != nnz) atomicOr(flags, 1);      // a bitmap that does not match its row length
}

}  // namespace

// cp: column pointers (int32 or int64, n_cols + 1), ri: 0-based row ids (ascending inside a column, distinct),
// x: values.  p_out int64[n_rows + 1], j_out int32[nnz], x_out[nnz].  All device pointers.
template <typename P, typename TI, typename TO>
static int convert_impl(hs_ctx* ctx, const P* cp, const int32_t* ri, const TI* x, int64_t n_rows, int64_t n_cols,
                        int64_t nnz, int64_t* p_out, int32_t* j_out, TO* x_out) {
  cudaStream_t st = ctx->stream;
  // row counters live in shared memory: bands of 56 000 rows, one pass over the entries per band (a gene x cell matrix
  // has fewer genes than that; its transpose needs a band per 56 000 cells)
  const int64_t band = 56000;
  const int64_t n_bands = (hs_max<int64_t>(n_rows, 1) + band - 1) / band;
  const size_t smem = sizeof(uint32_t) * (size_t)hs_min<int64_t>(hs_max<int64_t>(n_rows, 1), band);
  // column blocks: enough CTAs to fill the machine several times over, at most 1024 histograms and 2 GB of them
  int64_t max_blocks = hs_min<int64_t>(1024, (int64_t)ctx->sm_count * 8);
  max_blocks = hs_max<int64_t>(1, hs_min<int64_t>(max_blocks, ((int64_t)2 << 30) / (4 * hs_max<int64_t>(n_rows, 1))));
  int n_blocks = (int)hs_min<int64_t>(hs_max<int64_t>(n_cols, 1), max_blocks);
  const int cols_per_block = (int)((hs_max<int64_t>(n_cols, 1) + n_blocks - 1) / n_blocks);
  n_blocks = (int)((hs_max<int64_t>(n_cols, 1) + cols_per_block - 1) / cols_per_block);
  if (n_rows + 1 > 0x7fffffffLL) return hs_fail(ctx, HS_ERR_INVALID, "hs_csc_to_csr: too many rows");
  size_t scan_ws = 0;
  HS_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, scan_ws, (const int64_t*)nullptr, (int64_t*)nullptr, (int)(n_rows + 1), st));
  auto up256 = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t hist_b = up256(sizeof(uint32_t) * (size_t)n_blocks * (size_t)hs_max<int64_t>(n_rows, 1));
  const size_t total_b = up256(sizeof(int64_t) * (size_t)(n_rows + 1));
  HS_TRY(hs_ensure(ctx, ctx->spmm_ws, 256 + hist_b + total_b + scan_ws + 256));
  char* w = (char*)ctx->spmm_ws.p;
  int* flags = (int*)w;                 w += 256;
  uint32_t* hist = (uint32_t*)w;        w += hist_b;
  int64_t* total = (int64_t*)w;         w += total_b;
  void* cub_ws = w;
  HS_CUDA(ctx, cudaMemsetAsync(flags, 0, 256, st));
  HS_CUDA(ctx, cudaFuncSetAttribute(k_csc_count<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  HS_CUDA(ctx, cudaFuncSetAttribute(k_csc_scatter<P, TI, TO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int64_t b = 0; b < n_bands; ++b) {
    const int64_t r0 = b * band, r1 = hs_min<int64_t>(n_rows, r0 + band);
    k_csc_count<P><<<n_blocks, kConvThreads, smem, st>>>(cp, ri, n_rows, r0, r1, n_cols, cols_per_block, hist, flags);
    HS_LAUNCH_CHECK(ctx);
  }
  k_csc_offsets<<<(unsigned)((n_rows + 256) / 256), 256, 0, st>>>(hist, n_blocks, n_rows, total);
  HS_LAUNCH_CHECK(ctx);
  HS_CUDA(ctx, cub::DeviceScan::ExclusiveSum(cub_ws, scan_ws, total, p_out, (int)(n_rows + 1), st));
  ctx->launches++;
  bool placed = false;
  bool use_sort = nnz >= ((int64_t)1 << 24);
  if (const char* e = getenv("HS_CONVERT_SORT")) use_sort = atoi(e) != 0 && nnz > 0;
  if (use_sort) {
    // the sort keys are the row ids: they must be inside [0, n_rows) before anything is placed by them
    int bad0 = 0;
    HS_CUDA(ctx, cudaMemcpyAsync(&bad0, flags, sizeof(int), cudaMemcpyDeviceToHost, st));
    HS_CUDA(ctx, cudaStreamSynchronize(st));
    if (bad0) return hs_fail(ctx, HS_ERR_INVALID, "hs_csc_to_csr: row ids outside [0, n_rows)");
    // chunks of whole column blocks, at most chunk_max entries each
    std::vector<int64_t> hcp((size_t)n_blocks + 1);
    {
      // first entry of every column block
      HS_TRY(hs_ensure(ctx, ctx->conv_sort, sizeof(int64_t) * (size_t)(n_blocks + 1)));
      int64_t* dcp = (int64_t*)ctx->conv_sort.p;
      k_csc_block_starts<P><<<(n_blocks + 256) / 256, 256, 0, st>>>(cp, n_cols, cols_per_block, n_blocks, dcp);
      HS_LAUNCH_CHECK(ctx);
      HS_CUDA(ctx, cudaMemcpyAsync(hcp.data(), dcp, sizeof(int64_t) * (size_t)(n_blocks + 1), cudaMemcpyDeviceToHost, st));
      HS_CUDA(ctx, cudaStreamSynchronize(st));
    }
    int64_t chunk_max = (int64_t)1 << 28;
    if (const char* e = getenv("HS_CONVERT_CHUNK")) chunk_max = hs_max<int64_t>(1024, atoll(e));
    int64_t biggest = 0;
    std::vector<int> cuts{0};
    for (int b = 0; b < n_blocks;) {
      int e = b + 1;
      while (e < n_blocks && hcp[(size_t)e + 1] - hcp[(size_t)b] <= chunk_max) ++e;
      biggest = hs_max(biggest, hcp[(size_t)e] - hcp[(size_t)b]);
      cuts.push_back(e);
      b = e;
    }
    int key_bits = 1;
    while (key_bits < 31 && ((int64_t)1 << key_bits) < n_rows) ++key_bits;
    typedef ColVal<TI> Rec;
    size_t sort_ws = 0, scan2_ws = 0;
    if (biggest > 0x7fffffffLL) use_sort = false;
    if (use_sort) {
      cub::DeviceRadixSort::SortPairs(nullptr, sort_ws, (const int32_t*)nullptr, (int32_t*)nullptr, (const Rec*)nullptr,
                                      (Rec*)nullptr, (int)biggest, 0, key_bits, st);
      cub::DeviceScan::ExclusiveSum(nullptr, scan2_ws, (const int64_t*)nullptr, (int64_t*)nullptr, (int)(n_rows + 1), st);
      const size_t rec_b = up256(sizeof(Rec) * (size_t)biggest), key_b = up256(sizeof(int32_t) * (size_t)biggest);
      const size_t row_b = up256(sizeof(int64_t) * (size_t)(n_rows + 1));
      const size_t need = 2 * rec_b + key_b + 2 * row_b + up256(sort_ws) + up256(scan2_ws);
      size_t free_b = 0, total_b = 0;
      HS_CUDA(ctx, cudaMemGetInfo(&free_b, &total_b));
      if (need > ctx->conv_sort.cap && need - ctx->conv_sort.cap + ((size_t)1 << 30) > free_b) use_sort = false;   // no room: scatter
      if (use_sort) {
        HS_TRY(hs_ensure(ctx, ctx->conv_sort, need));
        char* q = (char*)ctx->conv_sort.p;
        Rec* rec_in = (Rec*)q;             q += rec_b;
        Rec* rec_out = (Rec*)q;            q += rec_b;
        int32_t* keys_out = (int32_t*)q;   q += key_b;
        int64_t* cnt = (int64_t*)q;        q += row_b;
        int64_t* start = (int64_t*)q;      q += row_b;
        void* sort_tmp = q;                q += up256(sort_ws);
        void* scan_tmp = q;
        const int grid = ctx->sm_count * 16;
        for (size_t k = 0; k + 1 < cuts.size(); ++k) {
          const int b0 = cuts[k], b1 = cuts[k + 1];
          const int64_t c0 = hs_min<int64_t>(n_cols, (int64_t)b0 * cols_per_block), c1 = hs_min<int64_t>(n_cols, (int64_t)b1 * cols_per_block);
          const int64_t e0 = hcp[(size_t)b0], n = hcp[(size_t)b1] - e0;
          if (n <= 0) continue;
          k_csc_zip<P, TI><<<grid, 256, 0, st>>>(cp, x, c0, c1, e0, rec_in);
          HS_LAUNCH_CHECK(ctx);
          size_t ws = sort_ws;
          HS_CUDA(ctx, cub::DeviceRadixSort::SortPairs(sort_tmp, ws, ri + e0, keys_out, (const Rec*)rec_in, rec_out, (int)n, 0,
                                                       key_bits, st));
          ctx->launches += 3;
          k_csc_chunk_count<<<(unsigned)((n_rows + 256) / 256), 256, 0, st>>>(hist, total, n_rows, b0, b1, n_blocks, cnt);
          HS_LAUNCH_CHECK(ctx);
          size_t ws2 = scan2_ws;
          HS_CUDA(ctx, cub::DeviceScan::ExclusiveSum(scan_tmp, ws2, cnt, start, (int)(n_rows + 1), st));
          ctx->launches++;
          k_csc_place<TI, TO><<<grid, 256, 0, st>>>(rec_out, start, hist + (int64_t)b0 * n_rows, p_out, n_rows, j_out, x_out);
          HS_LAUNCH_CHECK(ctx);
        }
        placed = true;
      }
    }
  }
  if (nnz > 0 && !placed) {
    for (int64_t b = 0; b < n_bands; ++b) {
      const int64_t r0 = b * band, r1 = hs_min<int64_t>(n_rows, r0 + band);
      k_csc_scatter<P, TI, TO><<<n_blocks, kConvThreads, smem, st>>>(cp, ri, x, n_rows, r0, r1, n_cols, cols_per_block, hist,
                                                                    p_out, j_out, x_out);
      HS_LAUNCH_CHECK(ctx);
    }
  }
  int bad = 0;
  HS_CUDA(ctx, cudaMemcpyAsync(&bad, flags, sizeof(int), cudaMemcpyDeviceToHost, st));
  HS_CUDA(ctx, cudaStreamSynchronize(st));
  if (bad) return hs_fail(ctx, HS_ERR_INVALID, "hs_csc_to_csr: row ids outside [0, n_rows)");
  return HS_OK;
}

int hs_run_csc_to_csr_f32(hs_ctx* ctx, const int64_t* cp, const int32_t* ri, const float* x, int64_t n_rows, int64_t n_cols,
                          int64_t nnz, int64_t* p_out, int32_t* j_out, float* x_out) {
  return convert_impl<int64_t, float, float>(ctx, cp, ri, x, n_rows, n_cols, nnz, p_out, j_out, x_out);
}

int hs_run_csc_to_csr_f64(hs_ctx* ctx, const int64_t* cp, const int32_t* ri, const double* x, int64_t n_rows, int64_t n_cols,
                          int64_t nnz, int64_t* p_out, int32_t* j_out, double* x_out) {
  return convert_impl<int64_t, double, double>(ctx, cp, ri, x, n_rows, n_cols, nnz, p_out, j_out, x_out);
}

// ids of `n_rows` rows from their bitmaps (see k_bitmap_to_ids); ptr_dev: the rows' pointers into cols_dev
int hs_run_bitmap_to_ids(hs_ctx* ctx, const void* bits_dev, int64_t words, const int64_t* ptr_dev, int64_t n_rows,
                         int32_t* cols_dev, int* flags_dev, cudaStream_t st) {
  if (n_rows <= 0) return HS_OK;
  k_bitmap_to_ids<<<(unsigned)n_rows, 256, 0, st>>>((const unsigned long long*)bits_dev, words, ptr_dev, cols_dev, flags_dev);
  HS_LAUNCH_CHECK(ctx);
  return HS_OK;
}
