// K3 / K4-sparse on the tensor cores: sparse rows x density matrix as tile-bucketed dense MMAs.
//   P[row, :] = sum_i val[i] * D[col[i], :]        R/haystack_continuous.R:596 (colSums(D[ind,] * val))
//   D_KL      = sum(Pn * log(Pn / Q))              R/haystack_continuous.R:597-599
//
// A warp-per-row SIMT contraction needs 100 shared-memory operands per stored value (G = 100) and is bound
// by the shared-memory pipe at ~5 % of the HBM roofline (profiles/r1_k3_slab_spmm_v4_raw.csv).  Here the
// FMAs go to tcgen05 instead: rows are sorted by length and grouped into tiles of 128 rows; per 64-cell
// slab a tile's stored values are scattered into a dense fp16 hi/lo operand tile in shared memory and
// multiplied against the pre-tiled density operands (hs_umma.cu: k_density_tiles).  At 10 % fill the MMA
// does ten times the useful flops, and still costs a fraction of the gather.
//
// Two steps:
//  1. bucketing (k_tile_cuts, k_tile_region_size, k_tile_scan, k_tile_fill): CSR rows (ascending cell ids,
//     as Matrix keeps dgRMatrix@j) -> tile-major entries.  A tile row (128 rows) x chunk (64 slabs = 4096
//     cells) is a region; every row's run inside a chunk is found by binary search, counted per slab in
//     shared memory and written, slab by slab, as 8-byte entries (byte offset inside the swizzled operand
//     tile, fp16 hi | fp16 lo) -- already split and already addressed, so the MMA kernel's producers only
//     copy.  Reads and writes are coalesced row runs; a region's output window (~420 KB) stays in L2 until
//     its sectors are complete.  Every consumed entry is checked against its chunk and consecutive chunks
//     share their cut, so each stored value lands in exactly one tile whatever the order inside a chunk;
//     rows that are not ascending (or hold ids outside [0, N)) raise the flag that hands the call to the
//     order-independent gather kernel (hs_contract.cu).
//  2. k_tile_mma1 (default; k_tile_mma = the earlier form with two tile rows per CTA, HS_TILE_FORM=2): one CTA per
//     (tile row, cell segment), three 64 KB operand stages, three acc0 buffers in TMEM.  Warp roles: 0-7 drain TMEM;
//     8 streams the entry chunks through a small ring with cp.async.bulk, running ahead of the stages and prefetching
//     the stream into L2; 9-12 scatter the entries into the A tiles (every warp takes its share of every chunk);
//     13 issues the MMAs; 14 clears a released stage's A tiles and fetches its density tiles with bulk copies.
//     Precision as in hs_umma.cu: x = hi + lo * 2^-11 for both operands, acc0 += A_hi D_hi is drained after every
//     stage (the tensor core truncates each accumulating MMA; 4 MMAs per sum keep that bias at the fp32 SIMT level),
//     acc1 += A_hi D_lo + A_lo D_hi is 2^-11 smaller and stays in TMEM for a whole flush window.  fp32 register sums
//     go to fp64 (L2 resident, single writer) every 32 stages.  The last CTA to finish a tile row's segments runs the
//     fp64 KL epilogue for its 128 rows (segments added in fixed order: bit-identical repeats) -- no separate epilogue
//     launch, no memset of the sums.
//  3. The seeded randomization (hs_run_perm_tiles, end of this file) skips step 1: the draw is written straight into
//     fixed-capacity tile windows (k_perm_tile_draw) and the same MMA kernel scores them.
#include <cuda_fp16.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "hs_common.cuh"
#include "hs_prp.cuh"
#include "hs_ptx.cuh"

namespace {

using namespace hs_ptx;

constexpr int kTM = 128;                 // rows per tile
constexpr int kSlab = 64;                // cells per slab = one pipeline stage
constexpr int kChunkSlabs = 64;          // slabs per bucketing region
constexpr int kChunkCells = kChunkSlabs * kSlab;
constexpr int kTileBytes = kTM * kSlab * 2;      // 16 KB per fp16 operand term
constexpr int kStages = 2;
constexpr int kATileBytes = 2 * kTileBytes;      // A_hi + A_lo of one tile row
constexpr int kAStageBytes = 2 * kATileBytes;    // two tile rows
constexpr int kBStageBytes = 2 * kTileBytes;     // D_hi + D_lo
constexpr int kRingEntries = 512;                // entries per ring slot
constexpr int kRingSlotBytes = kRingEntries * 8;
constexpr int kRingSlots = 4;                    // per tile row
constexpr int kEntryLookahead = 64 * 1024;       // bytes of a tile row's entry stream kept ahead in L2
constexpr int kBLookahead = 4;                   // density operand stages prefetched into L2
constexpr int kFlushStages = 32;                 // fp32 registers -> fp64 every 32 stages (2048 cells)
constexpr int kMmaThreads = 640;                 // 20 warps: 96 registers per thread (warps are allocated in fours)
constexpr int kEpiWarps = 16;                    // warps 0..15 drain TMEM (TMEM lane quarter = warp % 4)
constexpr int kLoaderWarp = 16, kProducerWarp0 = 17, kMmaWarp = 19;
// The warp scheduler favours high warp ids: the MMA issuer, the producers and the loader sit above the sixteen
// epilogue warps, or their few instructions queue behind the epilogue's polling and arithmetic.
constexpr float kScaleD = 16384.f;               // must match hs_umma.cu (k_density_tiles)
constexpr float kLoScale = 2048.f;
constexpr uint32_t kLbo = 1024, kSbo = 2048;

constexpr size_t kSmemA = (size_t)kStages * kAStageBytes;                       // 128 KB
constexpr size_t kSmemB = (size_t)kStages * kBStageBytes;                       // 64 KB
constexpr size_t kSmemRing = 2 * (size_t)kRingSlots * kRingSlotBytes;           // 32 KB
constexpr size_t kSmemMma = 1024 + kSmemA + kSmemB + kSmemRing + 512;

// byte offset of element (row m, cell k) inside a 128 x 64 MN-major SW128 fp16 tile (hs_umma.cu layout); the kernels
// below assemble the same offset from its row and cell parts
[[maybe_unused]] __device__ __host__ __forceinline__ uint32_t elem_off(int m, int k) {
  const int mg = m >> 3, mi = mg >> 3, c8 = mg & 7, kg = k >> 3, k8 = k & 7;
  return (uint32_t)(((kg * 2 + mi) << 10) + (k8 << 7) + ((c8 ^ k8) << 4) + ((m & 7) << 1));
}
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)(kLbo >> 4) << 16) | ((uint64_t)(kSbo >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}
// D = F32, A = B = F16, both MN-major, M = 128, N = npad (multiple of 16)
__device__ __forceinline__ uint32_t make_idesc(int npad) {
  return (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(npad >> 3) << 17) | ((uint32_t)(kTM >> 4) << 24);
}

struct TileSrc {
  const int64_t* ptr;      // row r owns entries [ptr[r], ptr[r+1])
  const int32_t* cols;     // cell ids (separate arrays) ...
  const float* vals;
  const uint2* pairs;      // ... or interleaved (cell id, value bits)
  int col_base;            // 0 or 1
};
__device__ __forceinline__ int src_col(const TileSrc& s, int64_t i) {
  return (s.pairs ? (int)__ldg(&s.pairs[i].x) : __ldg(s.cols + i)) - s.col_base;
}

// ---- step 1a: cuts[t][c][r] = first entry of row (t, r) whose cell id is >= min(N, c * chunk cells) --------------
__global__ void __launch_bounds__(kTM)
k_tile_cuts(const TileSrc src, const int32_t* __restrict__ sorted_rows, int64_t n_rows, int64_t N, int n_chunks,
            uint32_t* __restrict__ cuts, int* __restrict__ flags) {
  const int c = blockIdx.x, t = blockIdx.y, r = threadIdx.x;
  const int64_t pos = (int64_t)t * kTM + r;
  uint32_t cut = 0;
  if (pos < n_rows) {
    const int row = sorted_rows[pos];
    const int64_t b = src.ptr[row];
    const int64_t len = src.ptr[row + 1] - b;
    const int64_t bound = hs_min<int64_t>(N, (int64_t)c * kChunkCells);
    int64_t lo = 0, hi = len;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if ((int64_t)src_col(src, b + mid) < bound) lo = mid + 1; else hi = mid;
    }
    cut = (uint32_t)lo;
    // ids below 0 (first cut not at the row start) or >= N (last cut not at the row end): not ours to score
    if ((c == 0 && lo != 0) || (c == n_chunks && lo != len)) atomicOr(flags, 1);
  }
  cuts[((size_t)t * (n_chunks + 1) + c) * kTM + r] = cut;
}

// ---- step 1b: capacity of every region = its entries + one pad entry per slab (tiles start 16-byte aligned) -------
__global__ void __launch_bounds__(256)
k_tile_region_size(const uint32_t* __restrict__ cuts, int n_tr, int n_chunks, unsigned long long* __restrict__ cap) {
  const int lane = threadIdx.x & 31;
  const int64_t reg = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (reg >= (int64_t)n_tr * n_chunks) return;
  const int t = (int)(reg / n_chunks), c = (int)(reg - (int64_t)t * n_chunks);
  const uint32_t* a = cuts + ((size_t)t * (n_chunks + 1) + c) * kTM;
  unsigned int s = 0;
#pragma unroll
  for (int k = 0; k < kTM / 32; ++k) s += a[kTM + lane + 32 * k] - a[lane + 32 * k];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) cap[reg] = (unsigned long long)((s + kChunkSlabs + 1) & ~1u);
}

// ---- step 1c: exclusive scan of the capacities (one block; a few 10^4 regions) ------------------------------------
__global__ void __launch_bounds__(1024)
k_tile_scan(const unsigned long long* __restrict__ cap, int64_t n, unsigned long long* __restrict__ base) {
  __shared__ unsigned long long wsum[32];
  __shared__ unsigned long long carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0ull;
  __syncthreads();
  for (int64_t i0 = 0; i0 < n; i0 += 1024) {
    const int64_t i = i0 + threadIdx.x;
    const unsigned long long v = i < n ? cap[i] : 0ull;
    unsigned long long incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long u = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += u;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const unsigned long long w = wsum[lane];
      unsigned long long wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long u = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += u;
      }
      wsum[lane] = wi - w;
    }
    __syncthreads();
    const unsigned long long c0 = carry;
    if (i < n) base[i] = c0 + wsum[warp] + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = c0 + wsum[31] + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) base[n] = carry;
}

// Lanes holding equal slab ids in a row (ids ascend along a row run, so equal ids are neighbours; if they are not,
// a slab simply gets several runs): head lane and length of this lane's run.  Cheaper than __match_any_sync.
__device__ __forceinline__ void lane_run(int s, int lane, int& head, int& len) {
  const int prev = __shfl_up_sync(0xffffffffu, s, 1);
  const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || s != prev);
  head = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
  const unsigned after = lane == 31 ? 0u : (heads >> (lane + 1)) << (lane + 1);
  const int next = after ? __ffs(after) - 1 : 32;
  len = next - head;
}

// ---- step 1d: fill the regions ----------------------------------------------------------------------------------------
// Instruction-lean on purpose (the first version spent 5.6 warp instructions per entry and was issue-bound at
// 73 % issue utilisation, profiles/r2_tile_fill_v1_raw.csv): 32-bit cell arithmetic, the source layout a template
// parameter, the row part of the operand offset hoisted, run detection by shuffle + ballot instead of match.
constexpr int kFillThreads = 512;

template <bool PAIRS>
__global__ void __launch_bounds__(kFillThreads)
k_tile_fill(const TileSrc src, const int32_t* __restrict__ sorted_rows, int64_t n_rows, int64_t N, int n_chunks,
            int n_slabs, const uint32_t* __restrict__ cuts, const unsigned long long* __restrict__ region_base,
            uint2* __restrict__ entries, unsigned long long* __restrict__ tile_start, uint32_t* __restrict__ tile_cnt,
            int* __restrict__ flags) {
  __shared__ unsigned int hist[kChunkSlabs];
  __shared__ unsigned int cursor[kChunkSlabs];
  __shared__ int64_t row_beg[kTM];
  __shared__ uint32_t row_a[kTM], row_b[kTM];
  const int c = blockIdx.x, t = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // cell ids relative to the chunk, as unsigned: anything outside [0, n_cell) -- ids before the chunk included --
  // compares as too large
  const int32_t base = (int32_t)((int64_t)c * kChunkCells) + src.col_base;
  const uint32_t n_cell = (uint32_t)hs_min<int64_t>(kChunkCells, N - (int64_t)c * kChunkCells);
  if (tid < kChunkSlabs) hist[tid] = 0u;
  if (tid < kTM) {
    const int64_t pos = (int64_t)t * kTM + tid;
    int64_t b = 0;
    uint32_t a0 = 0, a1 = 0;
    if (pos < n_rows) {
      b = src.ptr[sorted_rows[pos]];
      const uint32_t* cc = cuts + ((size_t)t * (n_chunks + 1) + c) * kTM;
      a0 = cc[tid];
      a1 = cc[kTM + tid];
    }
    row_beg[tid] = b;
    row_a[tid] = a0;
    row_b[tid] = a1;
  }
  __syncthreads();
  bool bad = false;
  auto load_rel = [&](const int64_t b, uint32_t i, uint32_t a1) -> uint32_t {
    if (i >= a1) return 0xffffffffu;
    const int32_t cid = PAIRS ? (int32_t)__ldg(&src.pairs[b + i].x) : __ldg(src.cols + b + i);
    return (uint32_t)(cid - base);
  };
  // pass 1: entries per slab; two batches of 32 in flight per warp
  // (a tile row's rows descend in length with their position: the warps take them in snake order, or warp 0 would get
  //  the longest row of every group of sixteen)
  for (int it = 0; it < kTM / (kFillThreads / 32); ++it) {
    const int r = it * (kFillThreads / 32) + ((it & 1) ? kFillThreads / 32 - 1 - warp : warp);
    const int64_t b = row_beg[r];
    const uint32_t a0 = row_a[r], a1 = row_b[r];
    for (uint32_t i0 = a0; i0 < a1; i0 += 64) {
      const uint32_t rel0 = load_rel(b, i0 + lane, a1), rel1 = load_rel(b, i0 + 32 + lane, a1);
      int head, len;
      {
        const bool ok = rel0 < n_cell;
        bad |= (i0 + lane < a1) && !ok;                 // the search assumed ascending ids
        const int s = ok ? (int)(rel0 >> 6) : -1;
        lane_run(s, lane, head, len);
        if (ok && lane == head) atomicAdd(&hist[s], (unsigned)len);
      }
      if (i0 + 32 < a1) {                               // warp-uniform
        const bool ok = rel1 < n_cell;
        bad |= (i0 + 32 + lane < a1) && !ok;
        const int s = ok ? (int)(rel1 >> 6) : -1;
        lane_run(s, lane, head, len);
        if (ok && lane == head) atomicAdd(&hist[s], (unsigned)len);
      }
    }
  }
  __syncthreads();
  // exclusive scan over the 64 slabs (tile starts are even: 16-byte aligned bulk copies)
  const unsigned long long rbase = region_base[(size_t)t * n_chunks + c];
  if (warp == 0) {
    const unsigned int h0 = hist[2 * lane], h1 = hist[2 * lane + 1];
    const unsigned int p0 = (h0 + 1) & ~1u, p1 = (h1 + 1) & ~1u;
    unsigned int incl = p0 + p1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int u = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += u;
    }
    const unsigned int e0 = incl - p0 - p1, e1 = e0 + p0;
    cursor[2 * lane] = e0;
    cursor[2 * lane + 1] = e1;
    const int s0 = c * kChunkSlabs + 2 * lane;
    if (s0 < n_slabs) {
      tile_start[(size_t)t * n_slabs + s0] = rbase + e0;
      tile_cnt[(size_t)t * n_slabs + s0] = h0;
    }
    if (s0 + 1 < n_slabs) {
      tile_start[(size_t)t * n_slabs + s0 + 1] = rbase + e1;
      tile_cnt[(size_t)t * n_slabs + s0 + 1] = h1;
    }
  }
  __syncthreads();
  // pass 2: place
  uint2* out = entries + rbase;
  auto place = [&](uint32_t rel, float xv, uint32_t roff, uint32_t c8) {
    const bool ok = rel < n_cell;
    const int s = ok ? (int)(rel >> 6) : -1;
    int head, len;
    lane_run(s, lane, head, len);
    unsigned int at = 0;
    if (ok && lane == head) at = atomicAdd(&cursor[s], (unsigned)len);
    at = __shfl_sync(0xffffffffu, at, head) + (unsigned)(lane - head);
    if (ok) {
      const uint32_t k = rel & 63u;
      const uint32_t off = roff | ((k >> 3) << 11) | ((k & 7u) << 7) | (((c8 ^ k) & 7u) << 4);
      const __half h = __float2half_rn(xv);
      const __half l = __float2half_rn((xv - __half2float(h)) * kLoScale);
      bad |= !(fabsf(xv) <= 65504.f);                   // fp16 overflow, Inf, NaN: the fp32 kernel scores the call
      out[at] = make_uint2(off, (uint32_t)__half_as_ushort(h) | ((uint32_t)__half_as_ushort(l) << 16));
    }
  };
  // (a tile row's rows descend in length with their position: the warps take them in snake order, or warp 0 would get
  //  the longest row of every group of sixteen)
  for (int it = 0; it < kTM / (kFillThreads / 32); ++it) {
    const int r = it * (kFillThreads / 32) + ((it & 1) ? kFillThreads / 32 - 1 - warp : warp);
    const int64_t b = row_beg[r];
    const uint32_t a0 = row_a[r], a1 = row_b[r];
    const uint32_t roff = ((uint32_t)(r >> 6) << 10) | ((uint32_t)(r & 7) << 1), c8 = (uint32_t)(r >> 3) & 7u;
    for (uint32_t i0 = a0; i0 < a1; i0 += 64) {
      uint32_t rel0 = 0xffffffffu, rel1 = 0xffffffffu;
      float x0 = 0.f, x1 = 0.f;
      if (i0 + lane < a1) {
        if (PAIRS) {
          const uint2 e = __ldg(src.pairs + b + i0 + lane);
          rel0 = (uint32_t)((int32_t)e.x - base);
          x0 = __uint_as_float(e.y);
        } else {
          rel0 = (uint32_t)(__ldg(src.cols + b + i0 + lane) - base);
          x0 = __ldg(src.vals + b + i0 + lane);
        }
      }
      if (i0 + 32 + lane < a1) {
        if (PAIRS) {
          const uint2 e = __ldg(src.pairs + b + i0 + 32 + lane);
          rel1 = (uint32_t)((int32_t)e.x - base);
          x1 = __uint_as_float(e.y);
        } else {
          rel1 = (uint32_t)(__ldg(src.cols + b + i0 + 32 + lane) - base);
          x1 = __ldg(src.vals + b + i0 + 32 + lane);
        }
      }
      place(rel0, x0, roff, c8);
      if (i0 + 32 < a1) place(rel1, x1, roff, c8);      // warp-uniform
    }
  }
  if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(flags, 1);
}

// ---- step 2: the MMA kernel -------------------------------------------------------------------------------------
struct TileMmaParams {
  const uint2* entries;
  const unsigned long long* tile_start;   // [n_tr][n_slabs] first entry of the tile
  const uint32_t* tile_cnt;               // [n_tr][n_slabs]
  const int32_t* sorted_rows;             // [n_rows] row ids by descending length
  int64_t n_rows;
  int n_tr, n_pairs, n_slabs, slabs_per_seg, n_seg;
  const unsigned char* Dh;
  const unsigned char* Dl;
  double* P64;                            // [n_seg][n_pairs * 256][ldP]
  int64_t seg_stride;
  int ldP, G, npad;
  const double* Q;
  unsigned int* pair_done;                // [n_pairs], zeroed by the caller
  double* out;
  int out_n_perm;                         // > 0: row = s * n_perm + r is written to out[s + r * n_sel]
  int64_t out_n_sel;
  const int* flags;                       // bit 0 set by the bucketing: the fallback kernel owns the output
  const unsigned char* zeros;             // kATileBytes of zeros (the copy engine clears the operand tiles with it)
  unsigned long long* dbg_times;          // development aid: [CTA][2] start / end of every CTA (globaltimer, ns)
  long long* dbg;                         // optional (HS_TILE_DEBUG): cycles the roles of one CTA spend waiting
  int dbg_block;                          // the CTA that reports (HS_TILE_DEBUG=<tile row>; default: mid-grid)
};

__device__ __forceinline__ void l2_prefetch(const void* gptr, uint32_t bytes) {     // bytes: multiple of 16
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sts_zero16(uint32_t a) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(a), "r"(0u) : "memory");
}
__device__ __forceinline__ uint2 lds_u2(uint32_t a) {
  uint2 r;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(a) : "memory");
  return r;
}
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((unsigned short)v) : "memory");
}

// volatile: the adds stay between the TMEM loads they consume, so only one group of loaded columns is live at a time
// (left to the scheduler, all loads of a stage are hoisted and the accumulators spill)
__device__ __forceinline__ void acc_add(float& a, uint32_t m) {
  asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a) : "f"(__uint_as_float(m)));
}
__device__ __forceinline__ void acc_fma(float& a, uint32_t m, float s) {
  asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a) : "f"(__uint_as_float(m)), "f"(s));
}

// fp32 sum -> fp64, unscaled, stored / added in place; one asm block per element so that the 64 conversions of a
// flush are not all computed (128 registers) before the first store
__device__ __forceinline__ void flush_store(double* addr, float a, double unscale) {
  asm volatile("{\n\t.reg .f64 d;\n\tcvt.f64.f32 d, %1;\n\tmul.f64 d, d, %2;\n\tst.global.f64 [%0], d;\n\t}"
               ::"l"(addr), "f"(a), "d"(unscale) : "memory");
}
__device__ __forceinline__ void flush_add(double* addr, float a, double unscale) {
  asm volatile("{\n\t.reg .f64 d;\n\tcvt.f64.f32 d, %1;\n\tmul.f64 d, d, %2;\n\tred.global.add.f64 [%0], d;\n\t}"
               ::"l"(addr), "f"(a), "d"(unscale) : "memory");
}

__global__ void __launch_bounds__(kMmaThreads, 1) k_tile_mma(const TileMmaParams p) {
  if (*p.flags & 1) return;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t a_base = smem_u32(smem);
  const uint32_t b_base = a_base + (uint32_t)kSmemA;
  const uint32_t ring_base = b_base + (uint32_t)kSmemB;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemA + kSmemB + kSmemRing);
  const uint32_t bar_base = smem_u32(bars);
  uint32_t* ring_meta = reinterpret_cast<uint32_t*>(bars + 32);     // [2 tiles][kRingSlots]
  uint32_t* tmem_slot = ring_meta + 2 * kRingSlots;
  uint32_t* last_flag = tmem_slot + 1;
  auto a_full = [&](int st, int t) { return bar_base + 8u * (st * 2 + t); };            // 0..3
  auto b_full = [&](int st) { return bar_base + 8u * (4 + st); };                       // 4..5
  auto stage_empty = [&](int st) { return bar_base + 8u * (6 + st); };                  // 6..7
  auto acc_full = [&](int t) { return bar_base + 8u * (8 + t); };                       // 8..9
  auto acc_empty = [&](int t) { return bar_base + 8u * (10 + t); };                     // 10..11
  auto ring_full = [&](int t, int sl) { return bar_base + 8u * (12 + t * kRingSlots + sl); };                     // 12..19
  auto ring_empty = [&](int t, int sl) { return bar_base + 8u * (12 + 2 * kRingSlots + t * kRingSlots + sl); };   // 20..27

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.x % p.n_pairs, seg = blockIdx.x / p.n_pairs;
  const int s0 = seg * p.slabs_per_seg;
  const int n_st = hs_max(0, hs_min(p.n_slabs, s0 + p.slabs_per_seg) - s0);
  const int n_tiles = (2 * pair + 1 < p.n_tr) ? 2 : 1;
  // development aid: cycles each non-epilogue role of one mid-grid CTA spends blocked on each barrier
  const bool dbg_on = p.dbg != nullptr && (int)blockIdx.x == p.dbg_block && warp >= kEpiWarps;
  long long w_acc[3] = {0, 0, 0};
  auto twait = [&](uint32_t bar, uint32_t parity, int slot) {
    if (dbg_on) {
      const long long t0 = clock64();
      mbar_wait(bar, parity);
      w_acc[slot] += clock64() - t0;
    } else {
      mbar_wait(bar, parity);
    }
  };
  const long long t_begin = clock64();

  if (threadIdx.x == 0) {
    for (int st = 0; st < kStages; ++st) {
      mbar_init(a_full(st, 0), 1);
      mbar_init(a_full(st, 1), 1);
      mbar_init(b_full(st), 1);
      mbar_init(stage_empty(st), 1);
    }
    for (int t = 0; t < 2; ++t) {
      for (int sl = 0; sl < kRingSlots; ++sl) {
        mbar_init(ring_full(t, sl), 1);
        mbar_init(ring_empty(t, sl), 1);
      }
      mbar_init(acc_full(t), 1);
      mbar_init(acc_empty(t), 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kLoaderWarp) {
    // ================= entry chunks: runs ahead of the stages as far as the rings allow =================
    // A tile row's entries of consecutive slabs are one contiguous stream (regions are laid out chunk after chunk),
    // so the stream is prefetched into L2 by address, kEntryLookahead bytes ahead of the stage being fetched: the
    // ring copies then see L2 latency, not HBM latency (four 4 KB slots in flight per tile row).
    unsigned long long m_start = 0ull;     // lane l: metadata of (stage base + l / 2, tile l & 1)
    uint32_t m_cnt = 0u;
    uint32_t seq[2] = {0u, 0u};
    unsigned long long pf_pos[2] = {0ull, 0ull}, pf_end[2] = {0ull, 0ull};
    if (n_st > 0)
      for (int t = 0; t < n_tiles; ++t) {
        const size_t i0 = (size_t)(2 * pair + t) * p.n_slabs + s0, i1 = i0 + (size_t)(n_st - 1);
        pf_pos[t] = p.tile_start[i0];
        pf_end[t] = p.tile_start[i1] + p.tile_cnt[i1];
      }
    for (int i = 0; i < n_st; ++i) {
      if ((i & 15) == 0) {
        const int si = i + (lane >> 1), t = lane & 1;
        m_start = 0ull;
        m_cnt = 0u;
        if (si < n_st && t < n_tiles) {
          const size_t idx = (size_t)(2 * pair + t) * p.n_slabs + (s0 + si);
          m_start = p.tile_start[idx];
          m_cnt = p.tile_cnt[idx];
        }
      }
      for (int t = 0; t < n_tiles; ++t) {
        const unsigned long long start = __shfl_sync(0xffffffffu, m_start, 2 * (i & 15) + t);
        const uint32_t cnt = __shfl_sync(0xffffffffu, m_cnt, 2 * (i & 15) + t);
        if (lane == 0) {
          const unsigned long long want = hs_min<unsigned long long>(pf_end[t], start + kEntryLookahead / 8);
          unsigned long long pos = hs_max<unsigned long long>(pf_pos[t], start) & ~511ull;     // 4 KB steps
          for (; pos < want; pos += 512) l2_prefetch(p.entries + pos, 4096);
          pf_pos[t] = pos;
        }
        uint32_t done = 0;
        do {
          const uint32_t n = hs_min<uint32_t>(kRingEntries, cnt - done);
          const int sl = (int)(seq[t] % kRingSlots);
          if (seq[t] >= kRingSlots) twait(ring_empty(t, sl), ((seq[t] / kRingSlots) - 1) & 1, 0);
          if (lane == 0) {
            const bool last = done + n >= cnt;
            ring_meta[t * kRingSlots + sl] = n | (last ? 0x80000000u : 0u);
            const uint32_t bytes = ((n + 1) & ~1u) * 8u;
            if (bytes) {
              mbar_arrive_expect_tx(ring_full(t, sl), bytes);
              bulk_g2s(ring_base + (uint32_t)(t * kRingSlots + sl) * kRingSlotBytes, p.entries + start + done, bytes,
                       ring_full(t, sl));
            } else {
              mbar_arrive(ring_full(t, sl));
            }
          }
          __syncwarp();
          ++seq[t];
          done += n;
        } while (done < cnt);
      }
    }
  } else if (warp == kMmaWarp) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(p.npad);
      for (int i = 0; i < n_st; ++i) {
        const int st = i & 1;
        const uint32_t par = (i >> 1) & 1;
        const bool window_start = (i % kFlushStages) == 0;
        for (int t = 0; t < n_tiles; ++t) {
          if (i >= 1) twait(acc_empty(t), (i - 1) & 1, 0);
          twait(a_full(st, t), par, 1);
          if (t == 0) twait(b_full(st), par, 2);
          tc_fence_after();
          const uint32_t a_st = a_base + (uint32_t)st * kAStageBytes + (uint32_t)t * kATileBytes;
          const uint32_t b_st = b_base + (uint32_t)st * kBStageBytes;
          const uint32_t acc0 = tmem_base + (uint32_t)(t * 256), acc1 = acc0 + 128;
#pragma unroll
          for (int k16 = 0; k16 < kSlab / 16; ++k16) {
            const uint32_t koff = (uint32_t)k16 * 2u * kSbo;
            const uint64_t a_hi = make_desc(a_st + koff), a_lo = make_desc(a_st + kTileBytes + koff);
            const uint64_t b_hi = make_desc(b_st + koff), b_lo = make_desc(b_st + kTileBytes + koff);
            umma_f16(acc0, a_hi, b_hi, idesc, k16 == 0 ? 0u : 1u);
            umma_f16(acc1, a_hi, b_lo, idesc, (window_start && k16 == 0) ? 0u : 1u);
            umma_f16(acc1, a_lo, b_hi, idesc, 1u);
          }
          umma_commit(acc_full(t));
        }
        umma_commit(stage_empty(st));
      }
    }
  } else if (warp >= kProducerWarp0) {
    // ================= A producers: one warp per tile row =================
    const int t = warp - kProducerWarp0;
    if (t < n_tiles) {
      uint32_t seq = 0;
      for (int i = 0; i < n_st; ++i) {
        const int st = i & 1;
        if (i >= kStages) twait(stage_empty(st), ((i >> 1) - 1) & 1, 0);
        if (t == 0 && lane == 0) {
          // the stage's density operands (B) travel with the first tile row's producer: same wait, no extra warp;
          // the tiles a few stages on are prefetched into L2 (the first CTA to reach a slab pays the HBM latency)
          const size_t pf = (size_t)(s0 + i + kBLookahead) * kTileBytes;
          if (i + kBLookahead < n_st) {
            l2_prefetch(p.Dh + pf, kTileBytes);
            l2_prefetch(p.Dl + pf, kTileBytes);
          }
          const uint32_t dst = b_base + (uint32_t)st * kBStageBytes;
          const size_t off = (size_t)(s0 + i) * kTileBytes;
          mbar_arrive_expect_tx(b_full(st), kBStageBytes);
          bulk_g2s(dst, p.Dh + off, kTileBytes, b_full(st));
          bulk_g2s(dst + kTileBytes, p.Dl + off, kTileBytes, b_full(st));
        }
        const uint32_t a_hi = a_base + (uint32_t)st * kAStageBytes + (uint32_t)t * kATileBytes;
        const uint32_t a_lo = a_hi + kTileBytes;
#pragma unroll 8
        for (int z = 0; z < kATileBytes / 512; ++z) sts_zero16(a_hi + (uint32_t)(z * 512 + lane * 16));
        __syncwarp();
        for (;;) {
          const int sl = (int)(seq % kRingSlots);
          twait(ring_full(t, sl), (seq / kRingSlots) & 1, 1);
          const uint32_t meta = ring_meta[t * kRingSlots + sl];
          const uint32_t n = meta & 0x7fffffffu;
          const uint32_t rb = ring_base + (uint32_t)(t * kRingSlots + sl) * kRingSlotBytes;
          uint32_t e = lane;
          for (; e + 96 < n; e += 128) {     // four entries in flight per lane
            const uint2 e0 = lds_u2(rb + e * 8u), e1 = lds_u2(rb + (e + 32) * 8u);
            const uint2 e2 = lds_u2(rb + (e + 64) * 8u), e3 = lds_u2(rb + (e + 96) * 8u);
            sts_u16(a_hi + e0.x, e0.y); sts_u16(a_lo + e0.x, e0.y >> 16);
            sts_u16(a_hi + e1.x, e1.y); sts_u16(a_lo + e1.x, e1.y >> 16);
            sts_u16(a_hi + e2.x, e2.y); sts_u16(a_lo + e2.x, e2.y >> 16);
            sts_u16(a_hi + e3.x, e3.y); sts_u16(a_lo + e3.x, e3.y >> 16);
          }
          for (; e < n; e += 32) {
            const uint2 e0 = lds_u2(rb + e * 8u);
            sts_u16(a_hi + e0.x, e0.y);
            sts_u16(a_lo + e0.x, e0.y >> 16);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(ring_empty(t, sl));
          ++seq;
          if (meta >> 31) break;
        }
        fence_proxy_async();      // generic-proxy stores -> visible to the tensor core's async proxy
        __syncwarp();
        if (lane == 0) mbar_arrive(a_full(st, t));
      }
    }
  } else {
    // ================= epilogue: TMEM -> fp32 registers -> fp64 =================
    const int ew = warp;
    const int q = warp & 3, half = (ew >> 2) & 1, t = ew >> 3;
    if (t < n_tiles) {
      const int64_t pos = ((int64_t)(2 * pair + t)) * kTM + 32 * q + lane;     // sorted position of this thread's row
      const bool live = pos < p.n_rows;
      const int col0 = 64 * half;
      const int n8 = hs_max(0, hs_min(8, (p.npad - col0) >> 3));
      const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(t * 256);
      double* prow = p.P64 + (int64_t)seg * p.seg_stride + pos * p.ldP;
      const double unscale = 1.0 / (double)kScaleD;
      // one flush window (<= kFlushStages stages) at a time: zero, accumulate, flush -- straight-line, so the 64
      // register sums are updated in place (a conditional reset inside one loop made the compiler keep two copies)
      for (int w0 = 0; w0 < n_st; w0 += kFlushStages) {
        const int wend = hs_min(n_st, w0 + kFlushStages);
        float acc2[64];
#pragma unroll
        for (int c = 0; c < 64; ++c) acc2[c] = 0.f;
        for (int i = w0; i < wend; ++i) {
          mbar_wait(acc_full(t), i & 1);
          tc_fence_after();
#pragma unroll
          for (int h = 0; h < 8; h += 2) {     // n8 is even; two loads in flight per wait
            if (h < n8) {
              uint32_t m0[8], m1[8];
              tmem_ld8_nowait(lane_addr + (uint32_t)(col0 + 8 * h), m0);
              tmem_ld8_nowait(lane_addr + (uint32_t)(col0 + 8 * h + 8), m1);
              tmem_ld_wait();
#pragma unroll
              for (int c = 0; c < 8; ++c) acc_add(acc2[8 * h + c], m0[c]);
#pragma unroll
              for (int c = 0; c < 8; ++c) acc_add(acc2[8 * h + 8 + c], m1[c]);
            }
          }
          if (i == wend - 1) {     // the cross terms of the whole window, scale 2^-11
#pragma unroll
            for (int h = 0; h < 8; h += 2) {
              if (h < n8) {
                uint32_t a0[8], a1[8];
                tmem_ld8_nowait(lane_addr + (uint32_t)(128 + col0 + 8 * h), a0);
                tmem_ld8_nowait(lane_addr + (uint32_t)(128 + col0 + 8 * h + 8), a1);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 8; ++c) acc_fma(acc2[8 * h + c], a0[c], 1.0f / kLoScale);
#pragma unroll
                for (int c = 0; c < 8; ++c) acc_fma(acc2[8 * h + 8 + c], a1[c], 1.0f / kLoScale);
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty(t));
        }
        if (live) {
#pragma unroll
          for (int c = 0; c < 64; ++c) {
            if (col0 + c < p.G) {
              // single writer per (segment, row, column): the order of the additions is fixed
              if (w0 == 0) flush_store(prow + col0 + c, acc2[c], unscale);
              else flush_add(prow + col0 + c, acc2[c], unscale);
            }
          }
        }
      }
      if (n_st == 0 && live) {
        for (int c = 0; c < 64; ++c)
          if (col0 + c < p.G) prow[col0 + c] = 0.0;
      }
    }
  }

  if (dbg_on && lane == 0) {
    // [0] total, [1] stages | loader: [2] ring_empty | mma: [3] acc_empty [4] a_full [5] b_full |
    // producer 0: [6] stage_empty [7] ring_full | producer 1: [8] stage_empty [9] ring_full
    if (warp == kLoaderWarp) { p.dbg[0] = clock64() - t_begin; p.dbg[1] = n_st; p.dbg[2] = w_acc[0]; }
    if (warp == kMmaWarp) { p.dbg[3] = w_acc[0]; p.dbg[4] = w_acc[1]; p.dbg[5] = w_acc[2]; }
    if (warp == kProducerWarp0) { p.dbg[6] = w_acc[0]; p.dbg[7] = w_acc[1]; }
    if (warp == kProducerWarp0 + 1) { p.dbg[8] = w_acc[0]; p.dbg[9] = w_acc[1]; }
  }
  // ---- teardown; the last CTA of this pair runs the KL epilogue ----
  tc_fence_before();
  __threadfence();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
  if (threadIdx.x == 0) {
    const unsigned int old = atomicAdd(p.pair_done + pair, 1u);
    *last_flag = (old == (unsigned int)(p.n_seg - 1)) ? 1u : 0u;
  }
  __syncthreads();
  if (*last_flag == 0u) return;
  __threadfence();
  for (int rr = warp; rr < 2 * kTM; rr += kMmaThreads / 32) {
    const int64_t pos = (int64_t)pair * 2 * kTM + rr;
    if (pos >= p.n_rows) break;
    double pr[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int j = lane + 32 * a;
      double sum = 0.0;
      if (j < p.G)
        for (int sg = 0; sg < p.n_seg; ++sg) sum += __ldcg(p.P64 + (int64_t)sg * p.seg_stride + pos * p.ldP + j);
      pr[a] = sum;
    }
    const double kl = hs_kl_epilogue<4>(pr, p.Q, p.G, lane);
    if (lane == 0) {
      const int64_t row = p.sorted_rows[pos];
      int64_t o = row;
      if (p.out_n_perm > 0) {
        const int64_t s = row / p.out_n_perm, r = row - s * p.out_n_perm;
        o = s + r * p.out_n_sel;
      }
      p.out[o] = kl;
    }
  }
}

// ---- step 2, second form: one tile row per CTA, three operand stages, three acc0 buffers -----------------------------
// Measured on the two-tile kernel above (HS_TILE_DEBUG): a tcgen05.commit reaches the waiting warps ~1000 clk after it
// is issued, so the round trip "MMAs of a stage done -> epilogue woken -> TMEM drained -> issuer released" is ~1800 clk
// while a tile row's stage is ~700-1200 clk of MMA work: with one acc0 buffer per tile row the tensor pipe idles, and
// with two operand stages the producers start late.  This form hides both hops: acc0 is triple-buffered
// (3 x 128 + 128 columns of acc1 = all of TMEM), the operand ring is three stages deep (3 x 64 KB), and the CTA owns a
// single tile row (the B operand is no longer shared between two tile rows; it comes out of L2).
constexpr int kS1 = 3;                                   // operand stages = acc0 buffers
// 32 KB of entries in flight; few large chunks: the loader spends ~500 clk per chunk whatever its size (measured:
// 13 chunks of 512 entries per stage made the densest tile rows loader-bound at 7300 clk per stage)
constexpr int kRing1Slots = 4, kRing1Entries = 1024, kRing1SlotBytes = kRing1Entries * 8;
constexpr int kProducers1 = 4;                           // A-producer warps (the scatter runs at ~0.5 entries/clk per warp)
constexpr int kMma1Threads = 32 * (11 + kProducers1);    // warps 0-7 epilogue, 8 loader, 9.. producers, then the MMA issuer and the stager
constexpr size_t kSmem1Stage = (size_t)kATileBytes + kBStageBytes;          // 64 KB: A_hi, A_lo, D_hi, D_lo
constexpr size_t kSmem1 = 1024 + kS1 * kSmem1Stage + kRing1Slots * kRing1SlotBytes + 512;

__global__ void __launch_bounds__(kMma1Threads, 1) k_tile_mma1(const TileMmaParams p) {
  if (*p.flags & 1) return;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t st_base = smem_u32(smem);
  const uint32_t ring_base = st_base + (uint32_t)(kS1 * kSmem1Stage);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kS1 * kSmem1Stage + kRing1Slots * kRing1SlotBytes);
  const uint32_t bar_base = smem_u32(bars);
  uint32_t* ring_meta = reinterpret_cast<uint32_t*>(bars + 32);
  uint32_t* tmem_slot = ring_meta + kRing1Slots;
  uint32_t* last_flag = tmem_slot + 1;
  auto a_full = [&](int st) { return bar_base + 8u * st; };                      // 0..2   (the producer warps)
  auto b_full = [&](int st) { return bar_base + 8u * (3 + st); };                // 3..5   (tx)
  auto stage_empty = [&](int st) { return bar_base + 8u * (6 + st); };           // 6..8   (commit)
  auto acc_full = [&](int b) { return bar_base + 8u * (9 + b); };                // 9..11  (commit)
  auto acc_empty = [&](int b) { return bar_base + 8u * (12 + b); };              // 12..14 (8 epilogue warps)
  auto ring_full = [&](int sl) { return bar_base + 8u * (15 + sl); };                    // 15..22 (tx)
  auto ring_empty = [&](int sl) { return bar_base + 8u * (15 + kRing1Slots + sl); };     // 19..22 (the producer warps)
  auto zero_full = [&](int st) { return bar_base + 8u * (15 + 2 * kRing1Slots + st); };   // 23..25 (tx): the A tiles are zeroed

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tr = blockIdx.x % p.n_tr, seg = blockIdx.x / p.n_tr;
  const int s0 = seg * p.slabs_per_seg;
  const int n_st = hs_max(0, hs_min(p.n_slabs, s0 + p.slabs_per_seg) - s0);
  constexpr int kLoader = 8, kProd0 = 9, kIssuer = 9 + kProducers1, kStager = 10 + kProducers1;
  // development aid (HS_TILE_DEBUG): cycles the non-epilogue roles of one mid-grid CTA spend blocked on each barrier
  const bool dbg_on = p.dbg != nullptr && (int)blockIdx.x == p.dbg_block && warp >= 8;
  long long w_acc[3] = {0, 0, 0};
  long long t_zero = 0, t_scatter = 0, t_fence = 0;
  auto twait = [&](uint32_t bar, uint32_t parity, int slot) {
    if (dbg_on) {
      const long long t0 = clock64();
      mbar_wait(bar, parity);
      w_acc[slot] += clock64() - t0;
    } else {
      mbar_wait(bar, parity);
    }
  };
  const long long t_begin = clock64();
  if (p.dbg_times && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    p.dbg_times[2 * blockIdx.x] = t;
  }

  if (threadIdx.x == 0) {
    for (int st = 0; st < kS1; ++st) {
      mbar_init(a_full(st), kProducers1);
      mbar_init(b_full(st), 1);
      mbar_init(stage_empty(st), 1);
      mbar_init(acc_full(st), 1);
      mbar_init(acc_empty(st), 8);
      mbar_init(zero_full(st), 1);
    }
    for (int sl = 0; sl < kRing1Slots; ++sl) {
      mbar_init(ring_full(sl), 1);
      mbar_init(ring_empty(sl), kProducers1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kIssuer) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  auto a_addr = [&](int st) { return st_base + (uint32_t)(st * kSmem1Stage); };
  auto b_addr = [&](int st) { return st_base + (uint32_t)(st * kSmem1Stage) + (uint32_t)kATileBytes; };

  if (warp == kLoader) {
    // ================= entry chunks through a ring, running ahead; the stream is prefetched into L2 =================
    unsigned long long m_start = 0ull;     // lane l: metadata of stage (base + l)
    uint32_t m_cnt = 0u;
    uint32_t seq = 0u;
    unsigned long long pf_pos = 0ull, pf_end = 0ull;
    if (n_st > 0) {
      const size_t i0 = (size_t)tr * p.n_slabs + s0, i1 = i0 + (size_t)(n_st - 1);
      pf_pos = p.tile_start[i0];
      pf_end = p.tile_start[i1] + p.tile_cnt[i1];
    }
    for (int i = 0; i < n_st; ++i) {
      if ((i & 31) == 0) {
        m_start = 0ull;
        m_cnt = 0u;
        if (i + lane < n_st) {
          const size_t idx = (size_t)tr * p.n_slabs + (s0 + i + lane);
          m_start = p.tile_start[idx];
          m_cnt = p.tile_cnt[idx];
        }
      }
      const unsigned long long start = __shfl_sync(0xffffffffu, m_start, i & 31);
      const uint32_t cnt = __shfl_sync(0xffffffffu, m_cnt, i & 31);
      if (lane == 0) {
        const unsigned long long want = hs_min<unsigned long long>(pf_end, start + kEntryLookahead / 8);
        unsigned long long pos = hs_max<unsigned long long>(pf_pos, start) & ~511ull;     // 4 KB steps
        for (; pos < want; pos += 512) l2_prefetch(p.entries + pos, 4096);
        pf_pos = pos;
      }
      uint32_t done = 0;
      do {
        const uint32_t n = hs_min<uint32_t>(kRing1Entries, cnt - done);
        const int sl = (int)(seq % kRing1Slots);
        if (seq >= kRing1Slots) twait(ring_empty(sl), ((seq / kRing1Slots) - 1) & 1, 0);
        if (lane == 0) {
          const bool last = done + n >= cnt;
          ring_meta[sl] = n | (last ? 0x80000000u : 0u);
          const uint32_t bytes = ((n + 1) & ~1u) * 8u;
          if (bytes) {
            mbar_arrive_expect_tx(ring_full(sl), bytes);
            bulk_g2s(ring_base + (uint32_t)sl * kRing1SlotBytes, p.entries + start + done, bytes, ring_full(sl));
          } else {
            mbar_arrive(ring_full(sl));
          }
        }
        __syncwarp();
        ++seq;
        done += n;
      } while (done < cnt);
    }
  } else if (warp == kIssuer) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(p.npad);
      const uint32_t acc1 = tmem_base + 384u;
      for (int i = 0; i < n_st; ++i) {
        const int st = i % kS1;
        const uint32_t use = (uint32_t)(i / kS1);
        const bool window_start = (i % kFlushStages) == 0;
        // acc1 restarts: the drain of the previous window's last stage (it reads acc1) must be over
        if (window_start && i > 0) twait(acc_empty((i - 1) % kS1), ((i - 1) / kS1) & 1, 0);
        if (i >= kS1) twait(acc_empty(st), (use - 1) & 1, 0);
        twait(a_full(st), use & 1, 1);
        twait(b_full(st), use & 1, 2);
        tc_fence_after();
        const uint32_t a_st = a_addr(st), b_st = b_addr(st);
        const uint32_t acc0 = tmem_base + (uint32_t)(st * 128);
#pragma unroll
        for (int k16 = 0; k16 < kSlab / 16; ++k16) {
          const uint32_t koff = (uint32_t)k16 * 2u * kSbo;
          const uint64_t a_hi = make_desc(a_st + koff), a_lo = make_desc(a_st + kTileBytes + koff);
          const uint64_t b_hi = make_desc(b_st + koff), b_lo = make_desc(b_st + kTileBytes + koff);
          umma_f16(acc0, a_hi, b_hi, idesc, k16 == 0 ? 0u : 1u);
          umma_f16(acc1, a_hi, b_lo, idesc, (window_start && k16 == 0) ? 0u : 1u);
          umma_f16(acc1, a_lo, b_hi, idesc, 1u);
        }
        umma_commit(acc_full(st));
        umma_commit(stage_empty(st));
      }
    }
  } else if (warp == kStager) {
    // ================= stager: as soon as the MMAs release a stage, its A tiles are zeroed and its density tiles fetched,
    //                   both by the copy engine (a 32 KB block of zeros sits in L2) =================
    if (lane == 0) {
      for (int i = 0; i < n_st; ++i) {
        const int st = i % kS1;
        const uint32_t use = (uint32_t)(i / kS1);
        if (i >= kS1) twait(stage_empty(st), (use - 1) & 1, 0);
        mbar_arrive_expect_tx(zero_full(st), kATileBytes);
        bulk_g2s(a_addr(st), p.zeros, kATileBytes, zero_full(st));
        const size_t pf = (size_t)(s0 + i + kBLookahead) * kTileBytes;
        if (i + kBLookahead < n_st) {
          l2_prefetch(p.Dh + pf, kTileBytes);
          l2_prefetch(p.Dl + pf, kTileBytes);
        }
        const size_t off = (size_t)(s0 + i) * kTileBytes;
        mbar_arrive_expect_tx(b_full(st), kBStageBytes);
        bulk_g2s(b_addr(st), p.Dh + off, kTileBytes, b_full(st));
        bulk_g2s(b_addr(st) + kTileBytes, p.Dl + off, kTileBytes, b_full(st));
      }
    }
  } else if (warp >= kProd0) {
    // ================= A producers: every warp scatters its share of every entry chunk into the zeroed tiles ===========
    const int pw = warp - kProd0;
    uint32_t seq = 0;
    for (int i = 0; i < n_st; ++i) {
      const int st = i % kS1;
      const uint32_t use = (uint32_t)(i / kS1);
      const long long tp0 = dbg_on ? clock64() : 0;
      mbar_wait(zero_full(st), use & 1);
      const uint32_t a_hi = a_addr(st), a_lo = a_hi + kTileBytes;
      const long long tp1 = dbg_on ? clock64() : 0;
      const long long wait_before = w_acc[1];
      for (;;) {
        const int sl = (int)(seq % kRing1Slots);
        twait(ring_full(sl), (seq / kRing1Slots) & 1, 1);
        const uint32_t meta = ring_meta[sl];
        {
          // 32 entries at a time, round robin over the warps; a chunk holds at most eight per lane: all of a lane's loads
          // are issued before its first store (one shared-memory round trip per chunk instead of one per entry)
          const uint32_t n = meta & 0x7fffffffu;
          const uint32_t rb = ring_base + (uint32_t)sl * kRing1SlotBytes;
          constexpr uint32_t kStep = 32u * kProducers1;
          constexpr int kPerLane = kRing1Entries / (int)kStep;
          const uint32_t e0 = (uint32_t)pw * 32u + lane;
          uint2 ent[kPerLane];
#pragma unroll
          for (int u = 0; u < kPerLane; ++u) {
            const uint32_t e = e0 + (uint32_t)u * kStep;
            ent[u] = e < n ? lds_u2(rb + e * 8u) : make_uint2(0xffffffffu, 0u);
          }
#pragma unroll
          for (int u = 0; u < kPerLane; ++u) {
            if (ent[u].x != 0xffffffffu) {
              sts_u16(a_hi + ent[u].x, ent[u].y);
              sts_u16(a_lo + ent[u].x, ent[u].y >> 16);
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(ring_empty(sl));
        ++seq;
        if (meta >> 31) break;
      }
      const long long tp2 = dbg_on ? clock64() : 0;
      fence_proxy_async();      // generic-proxy stores -> visible to the tensor core's async proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full(st));
      if (dbg_on) {
        t_zero += tp1 - tp0;
        t_scatter += (tp2 - tp1) - (w_acc[1] - wait_before);
        t_fence += clock64() - tp2;
      }
    }
  } else if (warp < 8) {
    // ================= epilogue: TMEM -> fp32 registers -> fp64 =================
    const int q = warp & 3, half = warp >> 2;
    const int64_t pos = (int64_t)tr * kTM + 32 * q + lane;     // sorted position of this thread's row
    const bool live = pos < p.n_rows;
    const int col0 = 64 * half;
    const int n8 = hs_max(0, hs_min(8, (p.npad - col0) >> 3));
    const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * q) << 16);
    double* prow = p.P64 + (int64_t)seg * p.seg_stride + pos * p.ldP;
    const double unscale = 1.0 / (double)kScaleD;
    for (int w0 = 0; w0 < n_st; w0 += kFlushStages) {
      const int wend = hs_min(n_st, w0 + kFlushStages);
      float acc2[64];
#pragma unroll
      for (int c = 0; c < 64; ++c) acc2[c] = 0.f;
      for (int i = w0; i < wend; ++i) {
        const int st = i % kS1;
        mbar_wait(acc_full(st), (i / kS1) & 1);
        tc_fence_after();
        const uint32_t acc0 = lane_addr + (uint32_t)(st * 128 + col0);
#pragma unroll
        for (int h = 0; h < 8; h += 2) {     // n8 is even; two loads in flight per wait
          if (h < n8) {
            uint32_t m0[8], m1[8];
            tmem_ld8_nowait(acc0 + (uint32_t)(8 * h), m0);
            tmem_ld8_nowait(acc0 + (uint32_t)(8 * h + 8), m1);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 8; ++c) acc_add(acc2[8 * h + c], m0[c]);
#pragma unroll
            for (int c = 0; c < 8; ++c) acc_add(acc2[8 * h + 8 + c], m1[c]);
          }
        }
        if (i == wend - 1) {     // the cross terms of the whole window, scale 2^-11
#pragma unroll
          for (int h = 0; h < 8; h += 2) {
            if (h < n8) {
              uint32_t a0[8], a1[8];
              tmem_ld8_nowait(lane_addr + (uint32_t)(384 + col0 + 8 * h), a0);
              tmem_ld8_nowait(lane_addr + (uint32_t)(384 + col0 + 8 * h + 8), a1);
              tmem_ld_wait();
#pragma unroll
              for (int c = 0; c < 8; ++c) acc_fma(acc2[8 * h + c], a0[c], 1.0f / kLoScale);
#pragma unroll
              for (int c = 0; c < 8; ++c) acc_fma(acc2[8 * h + 8 + c], a1[c], 1.0f / kLoScale);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty(st));
      }
      if (live) {
#pragma unroll
        for (int c = 0; c < 64; ++c) {
          if (col0 + c < p.G) {
            // single writer per (segment, row, column): the order of the additions is fixed
            if (w0 == 0) flush_store(prow + col0 + c, acc2[c], unscale);
            else flush_add(prow + col0 + c, acc2[c], unscale);
          }
        }
      }
    }
  }

  if (dbg_on && lane == 0) {
    if (warp == kLoader) { p.dbg[0] = clock64() - t_begin; p.dbg[1] = n_st; p.dbg[2] = w_acc[0]; }
    if (warp == kIssuer) { p.dbg[3] = w_acc[0]; p.dbg[4] = w_acc[1]; p.dbg[5] = w_acc[2]; }
    if (warp == kProd0) { p.dbg[6] = w_acc[0]; p.dbg[7] = w_acc[1]; p.dbg[10] = t_zero; p.dbg[11] = t_scatter; p.dbg[12] = t_fence; }
    if (warp == kProd0 + 1) { p.dbg[8] = w_acc[0]; p.dbg[9] = w_acc[1]; }
  }
  if (p.dbg_times && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    p.dbg_times[2 * blockIdx.x + 1] = t;
  }
  // ---- teardown; the last CTA of this tile row runs the KL epilogue ----
  tc_fence_before();
  __threadfence();
  __syncthreads();
  if (warp == kIssuer) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
  if (threadIdx.x == 0) {
    const unsigned int old = atomicAdd(p.pair_done + tr, 1u);
    *last_flag = (old == (unsigned int)(p.n_seg - 1)) ? 1u : 0u;
  }
  __syncthreads();
  if (*last_flag == 0u) return;
  __threadfence();
  for (int rr = warp; rr < kTM; rr += kMma1Threads / 32) {
    const int64_t pos = (int64_t)tr * kTM + rr;
    if (pos >= p.n_rows) break;
    double pr[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int j = lane + 32 * a;
      double sum = 0.0;
      if (j < p.G)
        for (int sg = 0; sg < p.n_seg; ++sg) sum += __ldcg(p.P64 + (int64_t)sg * p.seg_stride + pos * p.ldP + j);
      pr[a] = sum;
    }
    const double kl = hs_kl_epilogue<4>(pr, p.Q, p.G, lane);
    if (lane == 0) {
      const int64_t row = p.sorted_rows[pos];
      int64_t o = row;
      if (p.out_n_perm > 0) {
        const int64_t s = row / p.out_n_perm, r = row - s * p.out_n_perm;
        o = s + r * p.out_n_sel;
      }
      p.out[o] = kl;
    }
  }
}


// =================================================================================================================
// K4 seeded, drawn straight into the tiles (R/haystack_continuous.R:266-283 with the library's own keyed draw)
// =================================================================================================================
// Row (unit s, randomization r) pairs the unit's i-th stored value with cell pi(i), pi the keyed permutation of
// hs_prp.cuh.  Nothing about a tile depends on the order of its entries (distinct cells per row), so the draw needs no
// sort: every (tile row, slab) owns a fixed-capacity window of the entry buffer, a counter hands out the places, and
// the cost is one permutation evaluation per stored value -- proportional to nnz, not to the number of cells.
//   capacity of a window = mean + 8 sd + 16 of the entries it expects (binomial in the rows' lengths), at most the
//   128 x 64 a tile can hold; a window that still overflows raises the flag and the call is scored by the chained
//   gather kernel instead.
// The rows of one tile row run together (blockIdx order), so a window's sectors fill while they sit in L2.

// per tile row: window capacity and the base of its windows
__global__ void __launch_bounds__(256)
k_perm_tile_caps(const int64_t* __restrict__ gp, const int32_t* __restrict__ sorted_rows, int64_t n_rows, int n_perm,
                 int64_t N, int n_tr, int n_slabs, double sigmas, uint32_t* __restrict__ cap,
                 unsigned long long* __restrict__ span) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tr) return;
  double sum = 0.0;
  for (int r = 0; r < kTM; ++r) {
    const int64_t pos = (int64_t)t * kTM + r;
    if (pos >= n_rows) break;
    const int64_t s = sorted_rows[pos] / n_perm;
    sum += (double)(gp[s + 1] - gp[s]);
  }
  const double mean = sum * (double)kSlab / (double)N;
  double c = ceil(mean + sigmas * sqrt(mean) + 16.0);
  c = fmax(fmin(c, (double)(kTM * kSlab)), 2.0);
  const uint32_t ci = ((uint32_t)c + 1u) & ~1u;
  cap[t] = ci;
  span[t] = (unsigned long long)ci * (unsigned long long)n_slabs;
}

__global__ void __launch_bounds__(1024)
k_perm_tile_draw(const int64_t* __restrict__ gp, const float* __restrict__ vals, const int32_t* __restrict__ sorted_rows,
                 int64_t n_rows, int n_perm, uint32_t N, int n_slabs, int n_part, int slabs_per_part, uint64_t seed,
                 uint64_t stream_base, const unsigned long long* __restrict__ unit_stream,
                 const uint32_t* __restrict__ cap, const unsigned long long* __restrict__ tbase,
                 uint32_t* __restrict__ tile_cnt, uint2* __restrict__ entries, int* __restrict__ flags) {
  // blockIdx = ((tile row * n_part) + part) * 128 + row in tile: the CTAs in flight share one or two (tile row, part)
  const int rt = (int)(blockIdx.x % kTM);
  const int64_t grp = blockIdx.x / kTM;
  const int part = (int)(grp % n_part);
  const int t = (int)(grp / n_part);
  const int64_t pos = (int64_t)t * kTM + rt;
  if (pos >= n_rows) return;
  const int64_t row = sorted_rows[pos];
  const int64_t s = row / n_perm;
  const int64_t g0 = gp[s];
  const uint32_t nnz = (uint32_t)hs_min<int64_t>(gp[s + 1] - g0, (int64_t)N);
  const uint64_t stream = unit_stream ? (uint64_t)unit_stream[s] + (uint64_t)(row - s * n_perm) : stream_base + (uint64_t)row;
  const HsPrp prp = hs_prp_make(seed, stream, N);
  const uint32_t c_lo = (uint32_t)part * (uint32_t)slabs_per_part * kSlab;
  const uint32_t c_span = (uint32_t)slabs_per_part * kSlab;
  const uint32_t wcap = cap[t];
  uint2* const out = entries + tbase[t];
  uint32_t* const cnt = tile_cnt + (size_t)t * n_slabs;
  const uint32_t roff = ((uint32_t)(rt >> 6) << 10) | ((uint32_t)(rt & 7) << 1), c8 = (uint32_t)(rt >> 3) & 7u;
  const float* v = vals + g0;
  bool bad = false;
  // four draws in flight per thread: the counter round trips (an L2 atomic each) overlap
  constexpr int kU = 4;
  for (uint32_t i0 = threadIdx.x; i0 < nnz; i0 += kU * blockDim.x) {
    uint32_t cell[kU], at[kU];
    float xv[kU];
    bool ok[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const uint32_t i = i0 + (uint32_t)u * blockDim.x;
      ok[u] = i < nnz;
      cell[u] = hs_prp_enc1(prp, ok[u] ? i : 0u);
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      while (cell[u] >= N) cell[u] = hs_prp_enc1(prp, cell[u]);      // cycle walking: rare
      ok[u] = ok[u] && (cell[u] - c_lo < c_span);
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const uint32_t i = i0 + (uint32_t)u * blockDim.x;
      xv[u] = ok[u] ? __ldg(v + i) : 0.f;
      at[u] = ok[u] ? atomicAdd(cnt + (cell[u] >> 6), 1u) : 0u;
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      if (!ok[u]) continue;
      const uint32_t slab = cell[u] >> 6, k = cell[u] & 63u;
      const __half h = __float2half_rn(xv[u]);
      const __half l = __float2half_rn((xv[u] - __half2float(h)) * kLoScale);
      // beyond fp16, or the window is full: the gather kernel scores the call
      bad |= !(fabsf(xv[u]) <= 65504.f) || at[u] >= wcap;
      if (at[u] < wcap) {
        const uint32_t off = roff | ((k >> 3) << 11) | ((k & 7u) << 7) | (((c8 ^ k) & 7u) << 4);
        out[(size_t)slab * wcap + at[u]] =
            make_uint2(off, (uint32_t)__half_as_ushort(h) | ((uint32_t)__half_as_ushort(l) << 16));
      }
    }
  }
  if (bad) atomicOr(flags, 1);
}

// window bases for the MMA kernel; counters of overflowed windows back inside their window
__global__ void __launch_bounds__(256)
k_perm_tile_index(const uint32_t* __restrict__ cap, const unsigned long long* __restrict__ tbase, int n_tr, int n_slabs,
                  unsigned long long* __restrict__ tile_start, uint32_t* __restrict__ tile_cnt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)n_tr * n_slabs) return;
  const int t = (int)(i / n_slabs), sl = (int)(i - (int64_t)t * n_slabs);
  const uint32_t c = cap[t];
  tile_start[i] = tbase[t] + (unsigned long long)sl * c;
  tile_cnt[i] = hs_min(tile_cnt[i], c);
}

}  // namespace

bool hs_tile_path_enabled(const hs_ctx* ctx) {
  if (ctx->G > 128) return false;
  const char* e = getenv("HS_TILE_DISABLE");
  if (e && *e && *e != '0') return false;
  return !hs_umma_disabled();
}

// Launch the MMA kernel over prepared tiles (entries / tile_start / tile_cnt) and the fused KL epilogue.
static int launch_tile_mma(hs_ctx* ctx, const uint2* entries, const unsigned long long* tile_start, const uint32_t* tile_cnt,
                           unsigned int* pair_done, const int32_t* sorted_rows_dev, int64_t n_rows, int npad, double* out_dev,
                           int out_n_perm, int64_t out_n_sel, const int* flags, cudaEvent_t* ev, bool dbg) {
  cudaStream_t st = ctx->stream;
  const int G = ctx->G;
  const int n_tr = (int)((n_rows + kTM - 1) / kTM);
  const int n_pairs = (n_tr + 1) / 2;
  const int n_slabs = (int)((ctx->N + kSlab - 1) / kSlab);
  // ---- step 2: MMA ----
  // cell segments: enough CTAs for ~8 waves (the tail is then one short CTA), at least 64 stages each;
  // blockIdx runs over the pairs first, so the CTAs in flight share a segment's density tiles through L2
  int form = 1;      // 1: one tile row per CTA, three stages (k_tile_mma1); 2: two tile rows per CTA (k_tile_mma)
  if (const char* e = getenv("HS_TILE_FORM")) form = atoi(e) == 2 ? 2 : 1;
  const int n_units = form == 1 ? n_tr : n_pairs;
  // ... and among the segment counts around that, the one whose CTAs come out closest to whole waves (79 tile rows
  // x 15 segments is one CTA more than eight waves of 148: a ninth wave for a single CTA)
  int64_t want_seg = ((int64_t)ctx->sm_count * 8 + n_units - 1) / n_units;
  const int64_t max_seg = hs_max<int64_t>(1, n_slabs / 64);
  auto seg_len = [&](int64_t sg) {
    const int64_t per = (n_slabs + sg - 1) / sg;
    return (per + kFlushStages - 1) / kFlushStages * kFlushStages;
  };
  if (const char* e = getenv("HS_TILE_SEGMENTS")) {
    want_seg = hs_max(1, atoi(e));
  } else {
    int64_t best = hs_min<int64_t>(want_seg, max_seg), best_cost = -1;
    for (int64_t sg = hs_max<int64_t>(1, want_seg / 2); sg <= hs_min<int64_t>(max_seg, want_seg * 2); ++sg) {
      const int64_t len = seg_len(sg);
      const int64_t segs = (n_slabs + len - 1) / len;
      const int64_t waves = (segs * n_units + ctx->sm_count - 1) / ctx->sm_count;
      const int64_t cost = waves * (len + 24);          // + a CTA's fixed cost, in stages
      if (best_cost < 0 || cost < best_cost || (cost == best_cost && sg > best)) {
        best = sg;
        best_cost = cost;
      }
    }
    want_seg = best;
  }
  int n_seg = (int)hs_max<int64_t>(1, hs_min<int64_t>(want_seg, max_seg));
  int slabs_per_seg = (int)seg_len(n_seg);
  n_seg = (n_slabs + slabs_per_seg - 1) / slabs_per_seg;
  const int ldP = npad;
  const int64_t seg_stride = (int64_t)n_pairs * 2 * kTM * ldP;
  HS_TRY(hs_ensure(ctx, ctx->part, sizeof(double) * (size_t)n_seg * (size_t)seg_stride));
  HS_CUDA(ctx, cudaMemsetAsync(pair_done, 0, (size_t)n_tr * 4, st));

  TileMmaParams prm;
  prm.entries = entries;
  prm.tile_start = tile_start;
  prm.tile_cnt = tile_cnt;
  prm.sorted_rows = sorted_rows_dev;
  prm.n_rows = n_rows;
  prm.n_tr = n_tr;
  prm.n_pairs = n_pairs;
  prm.n_slabs = n_slabs;
  prm.slabs_per_seg = slabs_per_seg;
  prm.n_seg = n_seg;
  prm.Dh = (const unsigned char*)ctx->dtile_hi.p;
  prm.Dl = (const unsigned char*)ctx->dtile_lo.p;
  prm.P64 = (double*)ctx->part.p;
  prm.seg_stride = seg_stride;
  prm.ldP = ldP;
  prm.G = G;
  prm.npad = npad;
  prm.Q = (const double*)ctx->Q.p;
  prm.pair_done = pair_done;
  prm.out = out_dev;
  prm.out_n_perm = out_n_perm;
  prm.out_n_sel = out_n_sel;
  prm.flags = flags;
  HS_TRY(hs_ensure(ctx, ctx->zeros, kATileBytes));
  HS_CUDA(ctx, cudaMemsetAsync(ctx->zeros.p, 0, kATileBytes, st));
  prm.zeros = (const unsigned char*)ctx->zeros.p;
  prm.dbg = nullptr;
  prm.dbg_times = nullptr;
  if (dbg) {
    HS_TRY(hs_ensure(ctx, ctx->umma_flags, 256));
    HS_CUDA(ctx, cudaMemsetAsync(ctx->umma_flags.p, 0, 256, st));
    prm.dbg = (long long*)((char*)ctx->umma_flags.p + 64);
    if (form == 1) {
      HS_TRY(hs_ensure(ctx, ctx->dbg_times, sizeof(unsigned long long) * 2 * (size_t)n_tr * (size_t)n_seg));
      prm.dbg_times = (unsigned long long*)ctx->dbg_times.p;
    }
  }
  {
    const int units = form == 1 ? n_tr : n_pairs;
    int tr_dbg = dbg ? atoi(getenv("HS_TILE_DEBUG")) : 0;
    prm.dbg_block = (n_seg / 2) * units + hs_min(hs_max(tr_dbg, 0), units - 1);      // a middle segment of that tile row
  }
  if (form == 1) {
    HS_CUDA(ctx, cudaFuncSetAttribute(k_tile_mma1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem1));
    k_tile_mma1<<<(unsigned)(n_tr * n_seg), kMma1Threads, kSmem1, st>>>(prm);
  } else {
    HS_CUDA(ctx, cudaFuncSetAttribute(k_tile_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMma));
    k_tile_mma<<<(unsigned)(n_pairs * n_seg), kMmaThreads, kSmemMma, st>>>(prm);
  }
  HS_LAUNCH_CHECK(ctx);
  if (dbg) {
    long long h[16] = {0};
    cudaEventRecord(ev[3], st);
    cudaStreamSynchronize(st);
    float t_cut = 0, t_fill = 0, t_mma = 0;
    cudaEventElapsedTime(&t_cut, ev[0], ev[1]);
    cudaEventElapsedTime(&t_fill, ev[1], ev[2]);
    cudaEventElapsedTime(&t_mma, ev[2], ev[3]);
    for (int i = 0; i < 4; ++i) cudaEventDestroy(ev[i]);
    fprintf(stderr, "[tile dbg] prepare %.3f ms, fill / generate %.3f ms, mma %.3f ms\n", t_cut, t_fill, t_mma);
    if (prm.dbg_times) {
      const size_t n_cta = (size_t)n_tr * n_seg;
      std::vector<unsigned long long> t(2 * n_cta);
      cudaMemcpy(t.data(), prm.dbg_times, sizeof(unsigned long long) * t.size(), cudaMemcpyDeviceToHost);
      unsigned long long t0 = ~0ull, t1 = 0, dmin = ~0ull, dmax = 0;
      double dsum = 0;
      for (size_t c = 0; c < n_cta; ++c) {
        t0 = hs_min(t0, t[2 * c]);
        t1 = hs_max(t1, t[2 * c + 1]);
        const unsigned long long d = t[2 * c + 1] - t[2 * c];
        dmin = hs_min(dmin, d);
        dmax = hs_max(dmax, d);
        dsum += (double)d;
      }
      // first / last tile row of the middle segment, and the same tile row in the first and last segments
      auto dur = [&](int tr, int sg) { return (double)(t[2 * ((size_t)sg * n_tr + tr) + 1] - t[2 * ((size_t)sg * n_tr + tr)]) * 1e-3; };
      fprintf(stderr, "[tile dbg] %zu CTAs: span %.3f ms, CTA time min %.1f mean %.1f max %.1f us, busy SM share %.2f | tile row 0 / mid / last "
                      "in the middle segment: %.1f %.1f %.1f us | stages per CTA %d\n",
              n_cta, (double)(t1 - t0) * 1e-6, (double)dmin * 1e-3, dsum / (double)n_cta * 1e-3, (double)dmax * 1e-3,
              dsum / ((double)(t1 - t0) * ctx->sm_count), dur(0, n_seg / 2), dur(n_tr / 2, n_seg / 2), dur(n_tr - 1, n_seg / 2), slabs_per_seg);
    }
    cudaMemcpy(h, prm.dbg, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[tile dbg] rows %lld pairs %d seg %d | CTA: %lld clk over %lld stages (%.0f/stage) | loader wait ring_empty %lld | "
                    "mma wait acc_empty %lld a_full %lld b_full %lld | producer0 wait stage_empty %lld ring_full %lld | "
                    "producer1 wait stage_empty %lld ring_full %lld | producer0 zero+sync %lld scatter %lld fence %lld\n",
            (long long)n_rows, n_pairs, n_seg, h[0], h[1], h[1] ? (double)h[0] / (double)h[1] : 0.0, h[2], h[3], h[4], h[5], h[6], h[7],
            h[8], h[9], h[10], h[11], h[12]);
  }
  return HS_OK;
}

// Rows to tile rows.  The tile kernels take 128 consecutive rows of the list they are given as a tile row, and a tile
// row costs the same below ~20 % fill (the MMA floor) and ~1 clk per entry more above it: with the rows simply sorted by
// length the 128 longest rows make a tile row 2.4x as slow as the median one.  Dealing the sorted rows round robin over
// the full tile rows gives every tile row the same cut through the length spectrum (the short tail keeps its partial
// tile row).  A row's score does not depend on which rows share its tile.
namespace {
__global__ void k_stripe_rows(const int32_t* __restrict__ sorted, int64_t n_rows, int32_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows) return;
  const int64_t n_full = n_rows / kTM;
  out[i] = i < n_full * kTM ? sorted[(i % kTM) * n_full + i / kTM] : sorted[i];
}
}  // namespace

// out may not alias sorted; returns the list the tile path should use (sorted itself when striping is switched off)
const int32_t* hs_tile_row_order(hs_ctx* ctx, const int32_t* sorted_rows_dev, int64_t n_rows, int32_t* out_dev) {
  if (const char* e = getenv("HS_TILE_STRIPE"))
    if (atoi(e) == 0) return sorted_rows_dev;
  if (n_rows < 2 * kTM) return sorted_rows_dev;
  k_stripe_rows<<<(unsigned)((n_rows + 255) / 256), 256, 0, ctx->stream>>>(sorted_rows_dev, n_rows, out_dev);
  ctx->launches++;
  return out_dev;
}

// Rows (ascending cell ids) x resident density matrix -> D_KL on the tensor cores.
//   sorted_rows_dev: row ids by descending length (n_rows of them); entries of row r at [ptr[r], ptr[r+1]) of
//   (cols, vals) or of `pairs`; nnz_total: entries of all rows together.
//   flags: device word; bit 0 is set when the rows cannot be scored here (the caller chains the fallback).
int hs_run_tile_spmm(hs_ctx* ctx, const int64_t* ptr_dev, const int32_t* sorted_rows_dev, int64_t n_rows,
                     int64_t nnz_total, const int32_t* cols_dev, const float* vals_dev, const void* pairs_dev,
                     int col_base, double* out_dev, int out_n_perm, int64_t out_n_sel, int* flags) {
  if (n_rows <= 0) return HS_OK;
  const int64_t N = ctx->N;
  const int G = ctx->G;
  int npad = (G + 15) & ~15;
  if (const char* e = getenv("HS_TILE_N128")) {
    if (*e && *e != '0') npad = 128;
  }
  HS_TRY(hs_ensure_density_tiles(ctx));
  cudaStream_t st = ctx->stream;
  const int n_tr = (int)((n_rows + kTM - 1) / kTM);
  const int n_slabs = (int)((N + kSlab - 1) / kSlab);
  const int n_chunks = (n_slabs + kChunkSlabs - 1) / kChunkSlabs;
  const int64_t n_regions = (int64_t)n_tr * n_chunks;
  if (n_regions > 0x7fffffffLL || (int64_t)n_tr * n_slabs > ((int64_t)1 << 40))
    return hs_fail(ctx, HS_ERR_INVALID, "internal: too many tiles");

  // ---- workspace ----
  auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t cuts_b = up((size_t)n_tr * (n_chunks + 1) * kTM * sizeof(uint32_t));
  const size_t cap_b = up((size_t)(n_regions + 1) * 8);
  const size_t ts_b = up((size_t)n_tr * n_slabs * 8);
  const size_t tc_b = up((size_t)n_tr * n_slabs * 4);
  const size_t pd_b = up((size_t)n_tr * 4);      // completion counters: one per tile row (or pair)
  HS_TRY(hs_ensure(ctx, ctx->tile_ws, cuts_b + 2 * cap_b + ts_b + tc_b + pd_b));
  char* w = (char*)ctx->tile_ws.p;
  uint32_t* cuts = (uint32_t*)w;                               w += cuts_b;
  unsigned long long* cap = (unsigned long long*)w;            w += cap_b;
  unsigned long long* rbase = (unsigned long long*)w;          w += cap_b;
  unsigned long long* tile_start = (unsigned long long*)w;     w += ts_b;
  uint32_t* tile_cnt = (uint32_t*)w;                           w += tc_b;
  unsigned int* pair_done = (unsigned int*)w;
  const int64_t ent_cap = nnz_total + n_regions * (kChunkSlabs + 2) + 1024;   // + the L2 prefetcher's overshoot
  HS_TRY(hs_ensure(ctx, ctx->tile_ent, sizeof(uint2) * (size_t)ent_cap));

  TileSrc src;
  src.ptr = ptr_dev;
  src.cols = cols_dev;
  src.vals = vals_dev;
  src.pairs = (const uint2*)pairs_dev;
  src.col_base = col_base;

  const bool dbg = getenv("HS_TILE_DEBUG") != nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  if (dbg) {
    for (auto& e : ev) cudaEventCreate(&e);
    cudaEventRecord(ev[0], st);
  }
  // ---- step 1: bucketing ----
  k_tile_cuts<<<dim3(n_chunks + 1, n_tr), kTM, 0, st>>>(src, sorted_rows_dev, n_rows, N, n_chunks, cuts, flags);
  HS_LAUNCH_CHECK(ctx);
  k_tile_region_size<<<(unsigned)((n_regions + 7) / 8), 256, 0, st>>>(cuts, n_tr, n_chunks, cap);
  HS_LAUNCH_CHECK(ctx);
  k_tile_scan<<<1, 1024, 0, st>>>(cap, n_regions, rbase);
  HS_LAUNCH_CHECK(ctx);
  if (dbg) cudaEventRecord(ev[1], st);
  if (pairs_dev)
    k_tile_fill<true><<<dim3(n_chunks, n_tr), kFillThreads, 0, st>>>(src, sorted_rows_dev, n_rows, N, n_chunks, n_slabs, cuts,
                                                                    rbase, (uint2*)ctx->tile_ent.p, tile_start, tile_cnt, flags);
  else
    k_tile_fill<false><<<dim3(n_chunks, n_tr), kFillThreads, 0, st>>>(src, sorted_rows_dev, n_rows, N, n_chunks, n_slabs, cuts,
                                                                     rbase, (uint2*)ctx->tile_ent.p, tile_start, tile_cnt, flags);
  HS_LAUNCH_CHECK(ctx);
  if (dbg) cudaEventRecord(ev[2], st);

  return launch_tile_mma(ctx, (const uint2*)ctx->tile_ent.p, tile_start, tile_cnt, pair_done, sorted_rows_dev, n_rows, npad, out_dev,
                         out_n_perm, out_n_sel, flags, ev, dbg);
}

// Seeded randomization on the tile path: draw into the tiles, multiply.  sorted_rows_dev: the n_sel * n_perm rows
// (row = unit * n_perm + randomization) by descending length.  Returns -1 when the shape is not for this path; *flags is
// raised on the device when the call must be scored by the caller's chained fallback instead.
int hs_run_perm_tiles(hs_ctx* ctx, const int64_t* gene_ptr_dev, int64_t n_sel, int64_t n_vals_total, const float* val_dev,
                      int n_perm, uint64_t seed, uint64_t stream_base, const unsigned long long* unit_stream_dev,
                      const int32_t* sorted_rows_dev, double* out_dev, int* flags) {
  const int64_t n_rows = n_sel * (int64_t)n_perm;
  const int64_t N = ctx->N;
  if (n_rows <= 0) return HS_OK;
  const int n_tr = (int)((n_rows + kTM - 1) / kTM);
  const int n_slabs = (int)((N + kSlab - 1) / kSlab);
  if (!hs_tile_path_enabled(ctx) || N >= ((int64_t)1 << 31) || (int64_t)n_tr * n_slabs > ((int64_t)1 << 31)) return -1;
  int npad = (ctx->G + 15) & ~15;
  if (const char* e = getenv("HS_TILE_N128")) {
    if (*e && *e != '0') npad = 128;
  }
  double sigmas = 8.0;
  if (const char* e = getenv("HS_PERM_TILE_SIGMAS")) sigmas = atof(e);      // tests force the overflow fallback with a negative value
  // windows: sum_t n_slabs * (mean_t + sigmas * sqrt(mean_t) + 18), bounded with Cauchy-Schwarz from what the host knows
  const double n_ids = (double)n_vals_total * (double)n_perm;
  const double mean_sum = n_ids * (double)kSlab / (double)N;
  const double ent_bound = (double)n_slabs * (mean_sum + fmax(sigmas, 0.0) * sqrt((double)n_tr * mean_sum) + 18.0 * n_tr);
  const double ent_max = (double)n_tr * (double)n_slabs * (double)(kTM * kSlab);
  const int64_t ent_cap = (int64_t)(fmin(ent_bound, ent_max) * 1.001) + 4096;
  HS_TRY(hs_ensure_density_tiles(ctx));
  auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t cap_b = up((size_t)n_tr * 4), span_b = up((size_t)(n_tr + 1) * 8);
  const size_t ts_b = up((size_t)n_tr * n_slabs * 8), tc_b = up((size_t)n_tr * n_slabs * 4), pd_b = up((size_t)n_tr * 4);
  HS_TRY(hs_ensure(ctx, ctx->tile_ws, cap_b + 2 * span_b + ts_b + tc_b + pd_b));
  HS_TRY(hs_ensure(ctx, ctx->tile_ent, sizeof(uint2) * (size_t)ent_cap));
  char* w = (char*)ctx->tile_ws.p;
  uint32_t* cap = (uint32_t*)w;                                w += cap_b;
  unsigned long long* span = (unsigned long long*)w;           w += span_b;
  unsigned long long* tbase = (unsigned long long*)w;          w += span_b;
  unsigned long long* tile_start = (unsigned long long*)w;     w += ts_b;
  uint32_t* tile_cnt = (uint32_t*)w;                           w += tc_b;
  unsigned int* pair_done = (unsigned int*)w;
  cudaStream_t st = ctx->stream;

  // cell ranges (development knob): measured at 1M cells x 10 % fill, cutting the cell axis buys nothing -- the draw is
  // bound by the L2 request rate (one counter atomic + one 8-byte store per value, ~90 G requests/s), not by partial
  // sectors reaching DRAM (profiles/r2_perm_tile_draw_raw.csv: 1.5 GB read, 13.9 GB written for 8 GB of entries)
  int n_part = 1;
  if (const char* e = getenv("HS_PERM_TILE_PARTS")) n_part = atoi(e);
  n_part = hs_max(1, hs_min(n_part, n_slabs));
  const int slabs_per_part = (n_slabs + n_part - 1) / n_part;
  int threads = (double)n_vals_total / (double)hs_max<int64_t>(n_sel, 1) >= 16384.0 ? 1024 : 256;
  if (const char* e = getenv("HS_PERM_TILE_THREADS")) threads = hs_max(32, hs_min(1024, atoi(e) / 32 * 32));
  if ((int64_t)n_tr * n_part * kTM > 0x7fffffffLL) return -1;

  const bool dbg = getenv("HS_TILE_DEBUG") != nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  if (dbg) {
    for (auto& e : ev) cudaEventCreate(&e);
    cudaEventRecord(ev[0], st);
  }
  k_perm_tile_caps<<<(n_tr + 255) / 256, 256, 0, st>>>(gene_ptr_dev, sorted_rows_dev, n_rows, n_perm, N, n_tr, n_slabs, sigmas,
                                                     cap, span);
  HS_LAUNCH_CHECK(ctx);
  k_tile_scan<<<1, 1024, 0, st>>>(span, n_tr, tbase);
  HS_LAUNCH_CHECK(ctx);
  HS_CUDA(ctx, cudaMemsetAsync(tile_cnt, 0, (size_t)n_tr * n_slabs * 4, st));
  if (dbg) cudaEventRecord(ev[1], st);
  k_perm_tile_draw<<<(unsigned)((int64_t)n_tr * n_part * kTM), threads, 0, st>>>(
      gene_ptr_dev, val_dev, sorted_rows_dev, n_rows, n_perm, (uint32_t)N, n_slabs, n_part, slabs_per_part, seed, stream_base,
      unit_stream_dev, cap, tbase, tile_cnt, (uint2*)ctx->tile_ent.p, flags);
  HS_LAUNCH_CHECK(ctx);
  k_perm_tile_index<<<(unsigned)(((int64_t)n_tr * n_slabs + 255) / 256), 256, 0, st>>>(cap, tbase, n_tr, n_slabs, tile_start,
                                                                                     tile_cnt);
  HS_LAUNCH_CHECK(ctx);
  if (dbg) cudaEventRecord(ev[2], st);
  return launch_tile_mma(ctx, (const uint2*)ctx->tile_ent.p, tile_start, tile_cnt, pair_done, sorted_rows_dev, n_rows, npad, out_dev,
                         n_perm, n_sel, flags, ev, dbg);
}
