// K3 / K4-sparse: sparse-row x density-matrix contraction with shared-memory staged density slabs.
// Since round 2 this SIMT kernel is the FALLBACK: rows with grid.points <= 128 go through the tensor-core tile path
// (hs_tiles.cu) first, and this kernel is chained behind it on a device flag (values beyond fp16, ids not ascending), or
// runs alone for grid.points > 128.  This file also holds the row sort, the seeded draw (pair generators and the call
// into the tile path's forward draw), the ids-given bucketing, and the gene selection by coefficient of variation.
//   P[row, :] = sum_i val[i] * D[col[i], :]        R/haystack_continuous.R:596 (colSums(D[ind,] * val))
//
// Design (DESIGN.md "K3"): the density matrix D ([N][ldD] fp32, cell-major) is streamed through shared
// memory one slab of S cells at a time with cp.async.bulk + mbarriers (two buffers; warp 0 issues the
// requests between its own rows).
// Every consumer warp owns K rows for the whole kernel and keeps their fp32 partial column sums in
// registers (lane l holds columns 4l..4l+3 of each of its K rows).  For each slab a warp visits only
// those of its rows whose next stored column falls inside the slab, reads the row's (col, val) run in
// coalesced batches of 32, broadcasts them through a per-warp scratch line and gathers D rows from
// the slab with one LDS.128 per non-zero.  fp32 partials are flushed into fp64 sums (global, L2
// resident) every kFlushNnz non-zeros, so the accumulation error stays ~1e-7 relative.
// Rows are dealt to (CTA, warp) bins in snake order of descending length, so bins carry equal work
// and all CTAs walk the slabs at the same pace (the slab a CTA needs is then usually still in L2).
// The fused KL epilogue runs in a second tiny kernel over the fp64 sums (P never leaves L2 as fp32).
//
// Requirement: within a row the stored columns must be non-decreasing at slab granularity (Matrix
// keeps dgRMatrix@j ascending).  Violations and negative ids are detected and flagged; the caller
// then runs the order-independent gather kernel (hs_contract.cu) instead.
#include <cub/block/block_radix_sort.cuh>
#include <cub/device/device_radix_sort.cuh>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "hs_common.cuh"
#include "hs_prp.cuh"

namespace {

constexpr int kMaxWarps = 16;                  // warps per CTA: 16 (128 registers per thread) or 12 (168)
constexpr int kFlushNnz = 512;                 // fp32 -> fp64 flush interval (non-zeros per row)
constexpr int kMaxSmem = 232448;               // 227 KB opt-in limit per CTA on sm_100

// ---- PTX helpers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
// 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (TMA engine, no tensor map)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// shared-space accesses by 32-bit address (keeps the generic->shared conversion out of the hot loop)
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(a));
  return r;
}
__device__ __forceinline__ uint4 lds_u4(uint32_t a) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a) : "memory");
  return r;
}
__device__ __forceinline__ uint2 lds_u2(uint32_t a) {
  uint2 r;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(a) : "memory");
  return r;
}
__device__ __forceinline__ void sts_u2(uint32_t a, uint32_t x, uint32_t y) {
  asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}

// ---- slot table ------------------------------------------------------------------------------------
// One slot = one row assigned to (tile, warp, k).  Built on the device from the row pointers.
struct Slots {
  int64_t* coff;   // offset of the row's first column id in cols[]
  int64_t* voff;   // offset of the row's first value in vals[]
  int32_t* len;    // stored entries (0 for an empty slot)
  int32_t* row;    // row index (-1 for an empty slot)
};

// PERM == false: CSR row r: cols/vals at ptr[r].  PERM == true: row = s * n_perm + r, gene s owns
// vals[ptr[s] .. ptr[s+1]) and ids at cols[ptr[s] * n_perm + r * nnz_s ..) (R/haystack_continuous.R:275).
template <bool PERM>
__global__ void k_row_len(const int64_t* __restrict__ ptr, int64_t n_rows, int n_perm, uint32_t* __restrict__ keys,
                          int32_t* __restrict__ rows, int* __restrict__ flags) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows) return;
  const int64_t s = PERM ? i / n_perm : i;
  const int64_t l = ptr[s + 1] - ptr[s];
  if (l > 0x7fffffffLL || l < 0) atomicOr(flags, 2);
  keys[i] = (uint32_t)hs_min<int64_t>(hs_max<int64_t>(l, 0), 0x7fffffffLL);
  rows[i] = (int32_t)i;
}

// Deal the length-sorted rows to the (tile, warp, slot) grid of the slab kernel.  The kernel ends with its
// slowest CTA, and inside a CTA every slab ends with its slowest warp, so both levels are balanced:
//  * tiles take every n_tiles-th rank, snaking (tile t gets rank j * n_tiles + t for even j, the mirrored
//    one for odd j): every CTA sees the same cut through the length spectrum;
//  * inside a tile the rows go, longest first, to the warp with the least entries so far that still has a
//    free slot (greedy longest-processing-time).
// One warp per tile; lane w keeps the running total of warp w.
constexpr int kMaxTileRows = kMaxWarps * 17;
template <bool PERM>
__global__ void __launch_bounds__(32) k_build_slots(const int64_t* __restrict__ ptr, const int32_t* __restrict__ sorted_rows,
                              int64_t n_rows, int n_perm, int n_tiles, int w_per_tile, int K, Slots sl) {
  // sorted_rows / n_rows: the (sub)range of the length-sorted row list this launch covers
  __shared__ int64_t s_co[kMaxTileRows], s_vo[kMaxTileRows];
  __shared__ int32_t s_len[kMaxTileRows], s_row[kMaxTileRows];
  const int tile = blockIdx.x, lane = threadIdx.x;
  const int per_tile = w_per_tile * K;
  for (int j = lane; j < per_tile; j += 32) {
    const int64_t i = (int64_t)j * n_tiles + ((j & 1) ? (n_tiles - 1 - tile) : tile);
    int64_t co = 0, vo = 0;
    int32_t len = 0, row = -1;
    if (i < n_rows) {
      row = sorted_rows[i];
      if (PERM) {
        const int64_t s = row / n_perm, r = row - s * n_perm;
        const int64_t g0 = ptr[s], nnz = ptr[s + 1] - g0;
        co = g0 * n_perm + r * nnz;
        vo = g0;
        len = (int32_t)nnz;
      } else {
        co = vo = ptr[row];
        len = (int32_t)(ptr[row + 1] - ptr[row]);
      }
    }
    s_co[j] = co; s_vo[j] = vo; s_len[j] = len; s_row[j] = row;
  }
  __syncwarp();
  int64_t load = 0;
  int count = 0;
  for (int j = 0; j < per_tile; ++j) {
    if (s_row[j] < 0) continue;   // warp-uniform
    // least-loaded warp with a free slot; ties to the lower warp
    unsigned long long key = (lane < w_per_tile && count < K) ? (((unsigned long long)load << 5) | (unsigned)lane) : ~0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
      key = other < key ? other : key;
    }
    if ((int)(key & 31u) == lane) {
      const int64_t slot = ((int64_t)tile * w_per_tile + lane) * K + count;
      sl.coff[slot] = s_co[j];
      sl.voff[slot] = s_vo[j];
      sl.len[slot] = s_len[j];
      sl.row[slot] = s_row[j];
      load += s_len[j];
      ++count;
    }
  }
  if (lane < w_per_tile)
    for (int k = count; k < K; ++k) {
      const int64_t slot = ((int64_t)tile * w_per_tile + lane) * K + k;
      sl.coff[slot] = 0;
      sl.voff[slot] = 0;
      sl.len[slot] = 0;
      sl.row[slot] = -1;
    }
}

// ---- main kernel -----------------------------------------------------------------------------------
struct SpmmParams {
  Slots sl;
  const int32_t* cols;
  const float* vals;
  const uint2* pairs;      // when set: interleaved (col, val bits) entries at coff (cols / vals unused)
  int col_base;            // 0 or 1 (R indices)
  const float* D;          // [N][ldD]
  int ldD;                 // multiple of 4
  int64_t N;
  int slab_cells;          // S
  int64_t seg_cells;       // cells per grid.y segment (multiple of S)
  int n_seg;
  double* P64;             // [n_seg][n_rows][ldP] fp64 column sums (zeroed by the caller); a (segment, row)
                           // has exactly one writer, and the epilogue adds the segments in a fixed order
  int64_t seg_stride;      // doubles between the segments of P64
  int ldP;
  int* flags;              // bit 0: order violation / negative id -> caller falls back
  const int* run_if_flag;  // when set: the kernel only runs if bit 0 of this word is set (fallback of the tile path)
};

template <int NV, int K>
__device__ __forceinline__ void flush_row(float (&acc)[4 * NV], double* __restrict__ prow, int lane, int ldD) {
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int c0 = 4 * (lane + 32 * v);
    if (c0 < ldD) {
      double2 a = *reinterpret_cast<double2*>(prow + c0);
      double2 b = *reinterpret_cast<double2*>(prow + c0 + 2);
      a.x += (double)acc[4 * v + 0];
      a.y += (double)acc[4 * v + 1];
      b.x += (double)acc[4 * v + 2];
      b.y += (double)acc[4 * v + 3];
      *reinterpret_cast<double2*>(prow + c0) = a;
      *reinterpret_cast<double2*>(prow + c0 + 2) = b;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[4 * v + e] = 0.f;
  }
}

template <int NV, int K, int kWarps>
__global__ void __launch_bounds__(kWarps * 32, 1) k_slab_spmm(const SpmmParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  if (p.run_if_flag && !(*p.run_if_flag & 1)) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.slab_cells, ldD = p.ldD;
  const size_t slab_bytes = (size_t)S * ldD * sizeof(float);
  unsigned char* slab0 = smem;
  uint2* scratch = reinterpret_cast<uint2*>(smem + 2 * slab_bytes);                    // [kMaxWarps][32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * slab_bytes + kMaxWarps * 32 * sizeof(uint2));
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[2]);

  const int64_t c_begin = (int64_t)blockIdx.y * p.seg_cells;
  const int64_t c_end = hs_min<int64_t>(p.N, c_begin + p.seg_cells);
  const int n_slabs = (int)((c_end - c_begin + S - 1) / S);

  if (threadIdx.x == 0) {
    mbar_init(full0, 1);
    mbar_init(full0 + 8, 1);
    mbar_init(empty0, kWarps);
    mbar_init(empty0 + 8, kWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // ---------------- producer duty (warp 0, between its own rows) ----------------
  // Slab s goes into buffer s & 1 once every warp has released slab s - 2 and that fill has landed
  // (a warp that skipped a slab releases it without ever waiting for the data).
  int load_s = 0;   // next slab to request; only meaningful in warp 0
  auto issue = [&](int s) {
    if (lane == 0) {
      const int b = s & 1;
      const int64_t c0 = c_begin + (int64_t)s * S;
      const uint32_t bytes = (uint32_t)(hs_min<int64_t>(S, c_end - c0) * ldD * sizeof(float));
      mbar_arrive_expect_tx(full0 + 8 * b, bytes);
      bulk_g2s(smem_u32(slab0 + b * slab_bytes), p.D + c0 * ldD, bytes, full0 + 8 * b);
    }
  };
  auto produce = [&](int upto, bool block) {   // request slabs < upto; warp-uniform control flow
    while (load_s < upto && load_s < n_slabs) {
      if (load_s >= 2) {
        const int b = load_s & 1;
        const uint32_t par = ((load_s >> 1) - 1) & 1;
        if (block) {
          mbar_wait(empty0 + 8 * b, par);
          mbar_wait(full0 + 8 * b, par);
        } else {
          const bool ready = mbar_test(empty0 + 8 * b, par) && mbar_test(full0 + 8 * b, par);
          if (!__shfl_sync(0xffffffffu, (int)ready, 0)) break;
        }
      }
      issue(load_s);
      ++load_s;
    }
  };
  if (warp == 0) produce(2, true);

  // ---------------- consumers ----------------
  // lane k (< K) carries the cursor of this warp's k-th row
  const int64_t bin = (int64_t)blockIdx.x * kWarps + warp;
  int64_t coff = 0, voff = 0;
  int len = 0, row = -1, pos = 0, next_col = 0x7fffffff, cnt = 0;
  if (lane < K) {
    const int64_t slot = bin * K + lane;
    coff = p.sl.coff[slot];
    voff = p.sl.voff[slot];
    len = p.sl.len[slot];
    row = p.sl.row[slot];
    if (p.n_seg > 1 && c_begin > 0) {   // first stored column >= c_begin (rows are ascending)
      int lo = 0, hi = len;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const int cm = p.pairs ? (int)p.pairs[coff + mid].x : p.cols[coff + mid];
        if ((int64_t)cm - p.col_base < c_begin) lo = mid + 1; else hi = mid;
      }
      pos = lo;
    }
    if (pos < len) next_col = (p.pairs ? (int)p.pairs[coff + pos].x : p.cols[coff + pos]) - p.col_base;
  }
  float acc[K][4 * NV];
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int e = 0; e < 4 * NV; ++e) acc[k][e] = 0.f;
  const uint32_t scr = smem_u32(scratch + warp * 32);
  double* const P64 = p.P64 + (int64_t)blockIdx.y * p.seg_stride;   // this segment's sums
  const bool paired = p.pairs != nullptr;
  auto load_entry = [&](int64_t ci, int64_t vi, int& c, float& v) {
    if (paired) {
      const uint2 e = __ldg(p.pairs + ci);
      c = (int)e.x;
      v = __uint_as_float(e.y);
    } else {
      c = __ldg(p.cols + ci);
      v = __ldg(p.vals + vi);
    }
  };
  // move the fp32 sums of the rows in `mask` into the fp64 column sums (one copy of the code)
  auto flush_rows = [&](unsigned mask) {
    while (mask) {
      const int k = __ffs(mask) - 1;
      mask &= mask - 1;
      float tmp[4 * NV];
#pragma unroll
      for (int kk = 0; kk < K; ++kk)
        if (kk == k) {
#pragma unroll
          for (int e = 0; e < 4 * NV; ++e) {
            tmp[e] = acc[kk][e];
            acc[kk][e] = 0.f;
          }
        }
      const int row_k = __shfl_sync(0xffffffffu, row, k);
      flush_row<NV, K>(tmp, P64 + (int64_t)row_k * p.ldP, lane, ldD);
    }
  };
  bool violation = false;
  // byte offset of this lane's q-th float4 inside a density row; lanes past the row end re-read its
  // last float4 (a broadcast, no extra wavefront) into accumulators that are never flushed
  const int nq = ldD >> 2;
  uint32_t lane_off[NV];
#pragma unroll
  for (int q = 0; q < NV; ++q) lane_off[q] = (uint32_t)hs_min(lane + 32 * q, nq - 1) * 16u;
  constexpr int GROUP = NV == 1 ? 4 : (NV == 2 ? 2 : 1);   // non-zeros whose loads are issued together
  const unsigned row_bytes = (unsigned)ldD * 4u;

  for (int s = 0; s < n_slabs; ++s) {
    const int b = s & 1;
    const int64_t cbase64 = c_begin + (int64_t)s * S;
    const int cbase = (int)cbase64;
    const int cend = (int)hs_min<int64_t>(c_end, cbase64 + S);
    const unsigned act = __ballot_sync(0xffffffffu, lane < K && next_col < cend);
    // software prefetch: the batch of (col, val) the next visit starts with is always in flight
    int pf_c = 0;
    float pf_v = 0.f;
    if (act) {
      const int k0 = __ffs(act) - 1;
      const int64_t co = __shfl_sync(0xffffffffu, coff, k0), vo = __shfl_sync(0xffffffffu, voff, k0);
      const int idx = __shfl_sync(0xffffffffu, pos, k0) + lane;
      if (idx < __shfl_sync(0xffffffffu, len, k0)) {
        load_entry(co + idx, vo + idx, pf_c, pf_v);
      }
    }
    if (warp == 0) {
      produce(s + 1, true);     // the slab this iteration reads must have been requested
      produce(s + 2, false);    // the next one as soon as its buffer is free
    }
    // Every warp observes every fill, active rows or not: a warp that ran ahead of the data could
    // otherwise lap the two-phase parity of the barriers (and release a buffer twice in one phase).
    mbar_wait(full0 + 8 * b, (s >> 1) & 1);
    if (act) {
      const uint32_t slab = smem_u32(slab0) + (uint32_t)(b * slab_bytes);
      // One copy of the visit code for all rows (a K-fold unrolled copy overflows the instruction
      // cache: warps sit in different rows); only the register accumulators need a static index.
      unsigned todo = act;
      while (todo) {
        const int k = __ffs(todo) - 1;
        todo &= todo - 1;
        const int64_t co = __shfl_sync(0xffffffffu, coff, k);
        const int64_t vo = __shfl_sync(0xffffffffu, voff, k);
        const int len_k = __shfl_sync(0xffffffffu, len, k);
        int pos_k = __shfl_sync(0xffffffffu, pos, k);
        int nc = 0x7fffffff, visited = 0;
        float vis[4 * NV];            // this visit's products (fp32, two levels: <= 32 per partial)
#pragma unroll
        for (int e = 0; e < 4 * NV; ++e) vis[e] = 0.f;
        for (;;) {
          const int idx = pos_k + lane;
          const int mc = idx < len_k ? pf_c - p.col_base : 0x7fffffff;
          const float mv = pf_v;
          const unsigned in = __ballot_sync(0xffffffffu, mc < cend);
          const int n_in = __popc(in);
          // next batch: more of this row, or the first batch of the next active row of this slab
          if (n_in == 32) {
            if (idx + 32 < len_k) {
              load_entry(co + idx + 32, vo + idx + 32, pf_c, pf_v);
            }
          } else if (todo) {
            const int k2 = __ffs(todo) - 1;
            const int64_t co2 = __shfl_sync(0xffffffffu, coff, k2), vo2 = __shfl_sync(0xffffffffu, voff, k2);
            const int idx2 = __shfl_sync(0xffffffffu, pos, k2) + lane;
            if (idx2 < __shfl_sync(0xffffffffu, len, k2)) {
              load_entry(co2 + idx2, vo2 + idx2, pf_c, pf_v);
            }
          }
          // entries of this slab must form a prefix of the batch and none may lie before the slab
          if (mc < cbase || in != (n_in == 32 ? 0xffffffffu : ((1u << n_in) - 1u))) violation = true;
          const bool here = mc >= cbase && mc < cend;   // anything else must never index the slab
          sts_u2(scr + lane * 8, here ? (unsigned)(mc - cbase) * row_bytes : 0u, here ? __float_as_uint(mv) : 0u);
          __syncwarp();
          float part[4 * NV];
#pragma unroll
          for (int e = 0; e < 4 * NV; ++e) part[e] = 0.f;
          int t = 0;
          for (; t + GROUP <= n_in; t += GROUP) {
            uint32_t eo[GROUP];
            float ev[GROUP];
            if (GROUP == 4) {
              const uint4 e01 = lds_u4(scr + t * 8), e23 = lds_u4(scr + t * 8 + 16);
              eo[0] = e01.x; ev[0] = __uint_as_float(e01.y); eo[1 % GROUP] = e01.z; ev[1 % GROUP] = __uint_as_float(e01.w);
              eo[2 % GROUP] = e23.x; ev[2 % GROUP] = __uint_as_float(e23.y); eo[3 % GROUP] = e23.z; ev[3 % GROUP] = __uint_as_float(e23.w);
            } else if (GROUP == 2) {
              const uint4 e01 = lds_u4(scr + t * 8);
              eo[0] = e01.x; ev[0] = __uint_as_float(e01.y); eo[1 % GROUP] = e01.z; ev[1 % GROUP] = __uint_as_float(e01.w);
            } else {
              const uint2 e0 = lds_u2(scr + t * 8);
              eo[0] = e0.x; ev[0] = __uint_as_float(e0.y);
            }
            float4 d[GROUP][NV];
#pragma unroll
            for (int i = 0; i < GROUP; ++i)
#pragma unroll
              for (int q = 0; q < NV; ++q) d[i][q] = lds_f4(slab + eo[i] + lane_off[q]);
#pragma unroll
            for (int i = 0; i < GROUP; ++i)
#pragma unroll
              for (int q = 0; q < NV; ++q) {
                part[4 * q + 0] = fmaf(ev[i], d[i][q].x, part[4 * q + 0]);
                part[4 * q + 1] = fmaf(ev[i], d[i][q].y, part[4 * q + 1]);
                part[4 * q + 2] = fmaf(ev[i], d[i][q].z, part[4 * q + 2]);
                part[4 * q + 3] = fmaf(ev[i], d[i][q].w, part[4 * q + 3]);
              }
          }
          for (; t < n_in; ++t) {
            const uint2 e0 = lds_u2(scr + t * 8);
            const float v = __uint_as_float(e0.y);
#pragma unroll
            for (int q = 0; q < NV; ++q) {
              const float4 dd = lds_f4(slab + e0.x + lane_off[q]);
              part[4 * q + 0] = fmaf(v, dd.x, part[4 * q + 0]);
              part[4 * q + 1] = fmaf(v, dd.y, part[4 * q + 1]);
              part[4 * q + 2] = fmaf(v, dd.z, part[4 * q + 2]);
              part[4 * q + 3] = fmaf(v, dd.w, part[4 * q + 3]);
            }
          }
#pragma unroll
          for (int e = 0; e < 4 * NV; ++e) vis[e] += part[e];
          __syncwarp();
          pos_k += n_in;
          visited += n_in;
          if (n_in < 32) {
            nc = __shfl_sync(0xffffffffu, mc, n_in);
            break;
          }
        }
        // fold the visit into the row's register accumulators (static register index per case)
#pragma unroll
        for (int kk = 0; kk < K; ++kk)
          if (kk == k) {
#pragma unroll
            for (int e = 0; e < 4 * NV; ++e) acc[kk][e] += vis[e];
          }
        if (lane == k) {
          pos = pos_k;
          next_col = nc;
          cnt += visited;
        }
        if (warp == 0) produce(s + 2, false);
      }
      // fp32 -> fp64 for the rows that have accumulated enough products
      const unsigned fm = __ballot_sync(0xffffffffu, lane < K && cnt >= kFlushNnz);
      if (fm) {
        flush_rows(fm);
        if ((fm >> lane) & 1u) cnt = 0;
      }
    }
    // this warp is done with buffer b (whether or not it touched it)
    __syncwarp();
    if (lane == 0) mbar_arrive(empty0 + 8 * b);
  }

  // no copy may be in flight when the CTA retires
  if (warp == 0)
    for (int s = hs_max(0, n_slabs - 2); s < n_slabs; ++s) mbar_wait(full0 + 8 * (s & 1), (s >> 1) & 1);

  // ---------------- final flush, left-over (out of range) ids ----------------
  flush_rows(__ballot_sync(0xffffffffu, lane < K && row >= 0 && cnt > 0));
  // Cell segments (n_seg > 1) start at a binary search, which presumes ascending ids.  Every consumed entry was
  // checked against its slab, so the row is scored correctly iff the segments' runs tile it: this segment must
  // have stopped exactly where the next one's search starts.  Anything else (ids out of order) -> fallback.
  if (p.n_seg > 1 && blockIdx.y + 1 < gridDim.y && lane < K && row >= 0) {
    int lo = 0, hi = len;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      const int cm = p.pairs ? (int)p.pairs[coff + mid].x : p.cols[coff + mid];
      if ((int64_t)cm - p.col_base < c_end) lo = mid + 1; else hi = mid;
    }
    if (lo != pos) violation = true;
  }
  // ids >= N are never consumed: poison the row like the gather kernel does (last segment only)
  if (lane < K && row >= 0 && blockIdx.y == gridDim.y - 1 && pos < len) P64[(int64_t)row * p.ldP] = NAN;
  if (__any_sync(0xffffffffu, violation) && lane == 0) atomicOr(p.flags, 1);
}

// ---- KL epilogue over the fp64 column sums -----------------------------------------------------
template <int NACC>
__global__ void __launch_bounds__(256)
k_p64_epilogue(const double* __restrict__ P64, int n_seg, int64_t seg_stride, int64_t n_rows, int ldP, int G,
               const double* __restrict__ Q, int n_perm, int64_t n_sel, const int* __restrict__ flags,
               const int* __restrict__ run_if_flag, double* __restrict__ out) {
  if (*flags & 1) return;   // the fallback kernel produces the output
  if (run_if_flag && !(*run_if_flag & 1)) return;   // the tile path produced it
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_rows) return;
  double pr[NACC];
#pragma unroll
  for (int a = 0; a < NACC; ++a) {
    const int j = lane + 32 * a;
    double sum = 0.0;
    if (j < G)
      for (int sg = 0; sg < n_seg; ++sg) sum += P64[sg * seg_stride + row * ldP + j];   // fixed order
    pr[a] = sum;
  }
  const double kl = hs_kl_epilogue<NACC>(pr, Q, G, lane);
  if (lane == 0) {
    int64_t o = row;
    if (n_perm > 0) {
      const int64_t s = row / n_perm, r = row - s * n_perm;
      o = s + r * n_sel;
    }
    out[o] = kl;
  }
}

template <int NV, int K, int W>
int launch_spmm(hs_ctx* ctx, const SpmmParams& prm, int n_tiles, size_t smem) {
  HS_CUDA(ctx, cudaFuncSetAttribute(k_slab_spmm<NV, K, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(n_tiles, prm.n_seg);
  k_slab_spmm<NV, K, W><<<grid, W * 32, smem, ctx->stream>>>(prm);
  HS_LAUNCH_CHECK(ctx);
  return HS_OK;
}

// Tile shapes (rows per warp K, warps per CTA W) per float4-per-lane count NV, smallest first.
// 4*NV*K accumulator registers per thread: W = 16 allows 128 registers, W = 12 allows 168.
struct Shape { int k, w; };
const Shape kShapes1[] = {{1, 16}, {2, 16}, {4, 16}, {8, 16}, {13, 16}, {17, 12}};   // (11,20), (9,24) measured slower
const Shape kShapes2[] = {{1, 16}, {2, 16}, {4, 16}, {8, 12}};
const Shape kShapes4[] = {{1, 16}, {2, 16}, {4, 12}};
const Shape kShapes8[] = {{1, 16}, {2, 12}};

// smallest shape whose single wave of `sms` CTAs covers n_rows; the largest one otherwise
Shape pick_shape(int nv, int64_t n_rows, int sms) {
  const Shape* t = nv == 1 ? kShapes1 : nv == 2 ? kShapes2 : nv == 4 ? kShapes4 : kShapes8;
  const int n = nv == 1 ? 6 : nv == 2 ? 4 : nv == 4 ? 3 : 2;
  if (const char* e = getenv("HS_SLAB_SHAPE")) {      // "k,w": tuning experiments
    int k = 0, w = 0;
    if (sscanf(e, "%d,%d", &k, &w) == 2)
      for (int i = 0; i < n; ++i)
        if (t[i].k == k && t[i].w == w) return t[i];
  }
  for (int i = 0; i < n; ++i)
    if ((int64_t)t[i].k * t[i].w * sms >= n_rows) return t[i];
  return t[n - 1];
}

// ---- K4-sparse: bucket the permutation ids by slab ----------------------------------------------
// sample(x = count.cells, size = nnz) (R/haystack_continuous.R:275) yields ids in random order; the slab
// kernel needs them ascending at slab granularity.  One CTA per permuted row does a counting sort by
// slab id in shared memory (histogram, scan, scatter) and writes (id, value) pairs in slab order; the
// pairing id[i] <-> val[i] of the reference is carried along.  Ids outside [0, N) go to an overflow bin
// at the end of the row, where the slab kernel never consumes them and poisons the row (NaN), as the
// reference's D[ind, ] would fail.  The order inside a slab follows the atomics, i.e. the fp32 summation
// order of a row is not fixed run to run (differences ~1e-8 relative on D_KL).
constexpr int kSortThreads = 512;

__global__ void k_perm_row_ptr(const int64_t* __restrict__ gp, int64_t n_sel, int n_perm, int64_t* __restrict__ rp) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n_rows = n_sel * n_perm;
  if (row > n_rows) return;
  if (row == n_rows) {
    rp[row] = gp[n_sel] * n_perm;
    return;
  }
  const int64_t s = row / n_perm, r = row - s * n_perm;
  rp[row] = gp[s] * n_perm + r * (gp[s + 1] - gp[s]);
}

// exact id / S for 0 <= id < 2^31 and S <= 2048: magic = ceil(2^40 / S)
__device__ __forceinline__ int div_magic(uint32_t id, uint64_t magic) { return (int)(((uint64_t)id * magic) >> 40); }

constexpr int kSortUnroll = 8;   // independent loads in flight per thread
constexpr int kSortMaxParts = 8;      // output-range parts per row (grid.y)
constexpr int kSortPartItemsDefault = 262144; // target entries per part (measured best of 16k..2M at C5): 256 KB of pairs, so that the scatter
                                      // windows of all resident CTAs stay in L2 until their sectors fill

// Writes (id, value) pairs in slab order.  A row is split into up to kSortMaxParts contiguous bin ranges,
// one CTA each: every CTA streams the whole row twice (histogram, scatter) but scatters only its own
// range, into a window small enough to be L2-resident -- the single-range version wrote 16x the
// algorithmic bytes to DRAM (one 32-byte sector per 4-byte store; profiles/r1_k4_perm_slab_sort_raw.csv).
__global__ void __launch_bounds__(kSortThreads, 2)
k_perm_slab_sort(const int64_t* __restrict__ rp, const int64_t* __restrict__ gp, int n_perm,
                 const float* __restrict__ vals, const int32_t* __restrict__ ind, int index_base, int64_t N,
                 uint64_t magic, int n_bins, int kSortPartItems, uint2* __restrict__ pairs_out) {
  extern __shared__ unsigned int hist[];      // [bins of this part + 1]
  __shared__ unsigned int warp_tot[kSortThreads / 32];
  __shared__ unsigned int before_tot;
  // consecutive CTAs are the parts of one row: they run together and share the row's ids through L2
  const int64_t row = blockIdx.x / kSortMaxParts;
  const int64_t s = row / n_perm;
  const int64_t beg = rp[row], nnz = rp[row + 1] - beg;
  const int parts = (int)hs_max<int64_t>(1, hs_min<int64_t>(kSortMaxParts, (nnz + kSortPartItems - 1) / kSortPartItems));
  const int part = blockIdx.x % kSortMaxParts;
  if (part >= parts) return;
  // bins [b_lo, b_hi) belong to this part; the last bin (n_bins - 1) collects invalid ids
  const int b_lo = (int)((int64_t)n_bins * part / parts), b_hi = (int)((int64_t)n_bins * (part + 1) / parts);
  const int nb = b_hi - b_lo;
  const int32_t* in = ind + beg;
  const float* vin = vals + gp[s];
  const int tid = threadIdx.x;
  for (int i = tid; i <= nb; i += kSortThreads) hist[i] = 0u;
  if (tid == 0) before_tot = 0u;
  __syncthreads();
  auto bin_of = [&](int32_t raw) {
    const int64_t id = (int64_t)raw - index_base;
    return (id < 0 || id >= N) ? n_bins - 1 : div_magic((uint32_t)id, magic);
  };
  constexpr int64_t kStep = (int64_t)kSortThreads * kSortUnroll;
  unsigned int before = 0;
  for (int64_t i0 = 0; i0 < nnz; i0 += kStep) {
    int32_t raw[kSortUnroll];
#pragma unroll
    for (int u = 0; u < kSortUnroll; ++u) {
      const int64_t i = i0 + (int64_t)u * kSortThreads + tid;
      raw[u] = i < nnz ? __ldg(in + i) : 0;
    }
#pragma unroll
    for (int u = 0; u < kSortUnroll; ++u)
      if (i0 + (int64_t)u * kSortThreads + tid < nnz) {
        const int b = bin_of(raw[u]);
        if (b < b_lo) ++before;
        else if (b < b_hi) atomicAdd(&hist[b - b_lo], 1u);
      }
  }
  // entries in earlier parts
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
  if ((tid & 31) == 0 && before) atomicAdd(&before_tot, before);
  __syncthreads();
  // exclusive scan of hist[0 .. nb): each thread owns a contiguous run
  const int per = (nb + kSortThreads - 1) / kSortThreads;
  const int lo = hs_min(tid * per, nb), hi = hs_min(lo + per, nb);
  unsigned int sum = 0;
  for (int i = lo; i < hi; ++i) sum += hist[i];
  unsigned int incl = sum;
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  unsigned int base = before_tot;
  for (int w2 = 0; w2 < warp; ++w2) base += warp_tot[w2];
  unsigned int run = base + incl - sum;
  for (int i = lo; i < hi; ++i) {
    const unsigned int c = hist[i];
    hist[i] = run;
    run += c;
  }
  __syncthreads();
  uint2* out = pairs_out + beg;
  for (int64_t i0 = 0; i0 < nnz; i0 += kStep) {
    int32_t raw[kSortUnroll];
#pragma unroll
    for (int u = 0; u < kSortUnroll; ++u) {
      const int64_t i = i0 + (int64_t)u * kSortThreads + tid;
      raw[u] = i < nnz ? __ldg(in + i) : 0;
    }
#pragma unroll
    for (int u = 0; u < kSortUnroll; ++u) {
      const int64_t i = i0 + (int64_t)u * kSortThreads + tid;
      if (i < nnz) {
        const int b = bin_of(raw[u]);
        if (b >= b_lo && b < b_hi) {
          // (storing the rank from the histogram pass instead of a second atomic was tried twice: the
          // extra scattered 2-byte traffic costs more than the atomics save)
          const unsigned int at = atomicAdd(&hist[b - b_lo], 1u);
          const int64_t id = (int64_t)raw[u] - index_base;
          out[at] = make_uint2((id < 0 || id >= N) ? 0x7fffffffu : (uint32_t)id, __float_as_uint(__ldg(vin + i)));
        }
      }
    }
  }
}

}  // namespace

constexpr size_t kSlabFixed = kMaxWarps * 32 * sizeof(uint2) + 64;

// cells per density slab for the resident state
int hs_slab_cells(const hs_ctx* ctx) {
  int S = (int)((kMaxSmem - kSlabFixed) / (2 * (size_t)ctx->ldD * sizeof(float)));
  S = hs_min<int>(S & ~3, 2048);
  if ((int64_t)S > ctx->N) S = (int)((ctx->N + 3) & ~3LL);
  return S;
}

// Sparse rows x resident density matrix -> D_KL.  PERM selects the randomization layout (see k_row_len).
// out_dev: fp64, n_rows entries (CSR) or n_sel x n_perm column-major (PERM).
// flag_dev_out (optional) receives the device address of the violation flag so that the caller can
// chain the order-independent fallback without a host round trip.
int hs_run_slab_spmm(hs_ctx* ctx, bool perm, const int64_t* ptr_dev, int64_t n_units, int n_perm,
                     const int32_t* cols_dev, const float* vals_dev, int col_base, double* out_dev,
                     int out_n_perm, int64_t out_n_sel, bool allow_tiles, const int** flag_dev_out,
                     const void* pairs_dev, int64_t nnz_hint, bool tile_only) {
  const int64_t n_rows = perm ? n_units * (int64_t)n_perm : n_units;
  if (n_rows <= 0) return HS_OK;
  if (n_rows > 0x7fffffffLL) return hs_fail(ctx, HS_ERR_INVALID, "too many rows (%lld) for one call", (long long)n_rows);
  const int G = ctx->G, ldD = ctx->ldD;
  const int nvq = (ldD / 4 + 31) / 32;                      // float4 per lane
  const int NV = nvq <= 1 ? 1 : nvq <= 2 ? 2 : nvq <= 4 ? 4 : 8;
  if (nvq > 8) return hs_fail(ctx, HS_ERR_INVALID, "grid.points=%d exceeds the supported maximum of 1024", G);
  const int sms = ctx->sm_count;
  const size_t fixed = kSlabFixed;
  const int S = hs_slab_cells(ctx);
  if (S < 4) return hs_fail(ctx, HS_ERR_INVALID, "grid.points=%d leaves no room for a density slab", G);
  const size_t smem = 2 * (size_t)S * ldD * sizeof(float) + fixed;

  // ---- row lengths, sorted descending ----
  const int ldP = ldD;
  size_t cub_bytes = 0;
  cub::DeviceRadixSort::SortPairsDescending(nullptr, cub_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                            (const int32_t*)nullptr, (int32_t*)nullptr, (int)n_rows, 0, 32, ctx->stream);
  const size_t keys_b = ((size_t)n_rows * 4 + 255) & ~(size_t)255;
  HS_TRY(hs_ensure(ctx, ctx->spmm_ws, 256 + 4 * keys_b + ((cub_bytes + 255) & ~(size_t)255)));
  char* w = (char*)ctx->spmm_ws.p;
  int* flags = (int*)w;                    w += 256;
  uint32_t* keys_in = (uint32_t*)w;        w += keys_b;
  uint32_t* keys_out = (uint32_t*)w;       w += keys_b;
  int32_t* rows_in = (int32_t*)w;          w += keys_b;
  int32_t* rows_out = (int32_t*)w;         w += keys_b;
  void* cub_ws = w;
  cudaStream_t st = ctx->stream;
  HS_CUDA(ctx, cudaMemsetAsync(flags, 0, 256, st));
  const int tb = 256;
  if (perm)
    k_row_len<true><<<(unsigned)((n_rows + tb - 1) / tb), tb, 0, st>>>(ptr_dev, n_rows, n_perm, keys_in, rows_in, flags);
  else
    k_row_len<false><<<(unsigned)((n_rows + tb - 1) / tb), tb, 0, st>>>(ptr_dev, n_rows, n_perm, keys_in, rows_in, flags);
  HS_LAUNCH_CHECK(ctx);
  HS_CUDA(ctx, cub::DeviceRadixSort::SortPairsDescending(cub_ws, cub_bytes, keys_in, keys_out, rows_in, rows_out,
                                                         (int)n_rows, 0, 32, st));
  ctx->launches += 3;   // cub: histogram + onesweep passes (counted once as a family)

  // ---- rows with ascending ids and G <= 128: the tensor-core tile path (hs_tiles.cu) scores the call; the slab
  //      kernel below is then only its fallback (fp16 overflow of a value, ids not ascending), chained on a flag ----
  const int* run_if = nullptr;
  if (allow_tiles && !perm && hs_tile_path_enabled(ctx) && (n_rows + 127) / 128 <= 65535) {
    int64_t nnz_total = nnz_hint;
    if (nnz_total < 0) {
      int64_t ends[2] = {0, 0};
      HS_CUDA(ctx, cudaMemcpyAsync(&ends[0], ptr_dev, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
      HS_CUDA(ctx, cudaMemcpyAsync(&ends[1], ptr_dev + n_rows, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
      HS_CUDA(ctx, cudaStreamSynchronize(st));
      nnz_total = ends[1] - ends[0];
      if (nnz_total < 0) return hs_fail(ctx, HS_ERR_INVALID, "row pointers must be non-decreasing");
    }
    int* tile_flag = flags + 2;
    // (rows_in is free once the sort has run: it takes the tile path's row order)
    const int32_t* tile_rows = hs_tile_row_order(ctx, rows_out, n_rows, rows_in);
    HS_TRY(hs_run_tile_spmm(ctx, ptr_dev, tile_rows, n_rows, nnz_total, cols_dev, vals_dev, pairs_dev, col_base, out_dev,
                            out_n_perm, out_n_sel, tile_flag));
    run_if = tile_flag;
    if (tile_only) {          // the caller chains its own fallback on the tile flag
      if (flag_dev_out) *flag_dev_out = tile_flag;
      return HS_OK;
    }
  }
  const int64_t n_dense = 0;

  // ---- the remaining rows: slab kernel ----
  const int64_t n_sub = n_rows - n_dense;
  // Tiling.  A CTA streams its whole cell segment through shared memory whatever its rows need (a slab
  // fill takes ~5 us), so a tile must hold enough rows (~48 at 10 % density) for the contraction, not the
  // streaming, to set the pace.  With many rows: whole waves of CTAs, rows dealt evenly (slots may stay
  // empty).  With few rows: fewer, taller tiles, and the cell axis split over grid.y so that the machine
  // is still full (5000 rows: 74 tiles x 2 segments instead of 148 tiles of 34 rows).
  const int64_t total_slabs = (ctx->N + S - 1) / S;
  int min_tile_rows = 48;
  if (const char* e = getenv("HS_SLAB_MIN_TILE_ROWS")) min_tile_rows = hs_max(1, atoi(e));
  const Shape big = pick_shape(NV, (int64_t)1 << 40, sms);            // the largest shape of this NV
  min_tile_rows = hs_min(min_tile_rows, big.k * big.w);
  Shape shp;
  int n_tiles, n_seg = 1;
  if (n_sub >= (int64_t)sms * min_tile_rows) {
    shp = pick_shape(NV, hs_max<int64_t>(n_sub, 1), sms);
    const int64_t per_wave = (int64_t)shp.w * shp.k * sms;
    const int64_t waves = hs_max<int64_t>(1, (n_sub + per_wave - 1) / per_wave);
    n_tiles = (int)hs_max<int64_t>(1, hs_min<int64_t>(waves * sms, n_sub));
  } else {
    const int max_seg = (int)hs_max<int64_t>(1, hs_min<int64_t>(sms, total_slabs / 4));
    int best_seg = 1;
    for (int sg = 1; sg <= max_seg; ++sg) {
      const int64_t tiles = hs_max<int64_t>(1, hs_min<int64_t>(sms / sg, n_sub));
      const int64_t rows_per_tile = (hs_max<int64_t>(n_sub, 1) + tiles - 1) / tiles;
      if (rows_per_tile > (int64_t)big.k * big.w) break;
      best_seg = sg;
      if (rows_per_tile >= min_tile_rows || tiles == 1) break;
    }
    n_seg = best_seg;
    n_tiles = (int)hs_max<int64_t>(1, hs_min<int64_t>(sms / n_seg, n_sub));
    const int64_t rows_per_tile = (hs_max<int64_t>(n_sub, 1) + n_tiles - 1) / n_tiles;
    shp = pick_shape(NV, rows_per_tile, 1);                             // smallest shape that holds a tile
  }
  const int K = shp.k, kWarps = shp.w;
  const int n_bins = n_tiles * kWarps;
  const int64_t n_slots = (int64_t)n_bins * K;
  const int64_t seg_cells = ((total_slabs + n_seg - 1) / n_seg) * S;
  n_seg = (int)((ctx->N + seg_cells - 1) / seg_cells);
  // fp64 column sums: one [n_rows][ldP] block per cell segment of either kernel (single writer per block
  // and row; the epilogue adds the blocks in order, so repeated calls are bit-identical)
  const int n_seg_all = hs_max(n_seg, 1);
  const int64_t seg_stride = n_rows * (int64_t)ldP;
  HS_TRY(hs_ensure(ctx, ctx->part, sizeof(double) * (size_t)n_seg_all * (size_t)seg_stride));
  HS_CUDA(ctx, cudaMemsetAsync(ctx->part.p, 0, sizeof(double) * (size_t)n_seg_all * (size_t)seg_stride, st));
  const size_t slots_b = ((size_t)n_slots * 8 + 255) & ~(size_t)255;
  const size_t slots4_b = ((size_t)n_slots * 4 + 255) & ~(size_t)255;
  HS_TRY(hs_ensure(ctx, ctx->spmm_slots, 2 * slots_b + 2 * slots4_b));
  Slots sl;
  w = (char*)ctx->spmm_slots.p;
  sl.coff = (int64_t*)w;                   w += slots_b;
  sl.voff = (int64_t*)w;                   w += slots_b;
  sl.len = (int32_t*)w;                    w += slots4_b;
  sl.row = (int32_t*)w;                    w += slots4_b;
  if (perm)
    k_build_slots<true><<<(unsigned)n_tiles, 32, 0, st>>>(ptr_dev, rows_out + n_dense, n_sub, n_perm, n_tiles, kWarps, K, sl);
  else
    k_build_slots<false><<<(unsigned)n_tiles, 32, 0, st>>>(ptr_dev, rows_out + n_dense, n_sub, n_perm, n_tiles, kWarps, K, sl);
  HS_LAUNCH_CHECK(ctx);

  SpmmParams prm;
  prm.sl = sl;
  prm.cols = cols_dev;
  prm.vals = vals_dev;
  prm.pairs = (const uint2*)pairs_dev;
  prm.col_base = col_base;
  prm.D = (const float*)ctx->D32.p;
  prm.ldD = ldD;
  prm.N = ctx->N;
  prm.slab_cells = S;
  prm.seg_cells = seg_cells;
  prm.n_seg = n_seg;
  prm.P64 = (double*)ctx->part.p;
  prm.seg_stride = seg_stride;
  prm.ldP = ldP;
  prm.flags = flags;
  prm.run_if_flag = run_if;

#define HS_SPMM_CASE(NVv, Kv, Wv) \
  if (n_sub <= 0) {} else \
  if (NV == NVv && K == Kv && kWarps == Wv) { HS_TRY((launch_spmm<NVv, Kv, Wv>(ctx, prm, n_tiles, smem))); } else
  HS_SPMM_CASE(1, 1, 16) HS_SPMM_CASE(1, 2, 16) HS_SPMM_CASE(1, 4, 16) HS_SPMM_CASE(1, 8, 16) HS_SPMM_CASE(1, 13, 16) HS_SPMM_CASE(1, 17, 12)
  HS_SPMM_CASE(2, 1, 16) HS_SPMM_CASE(2, 2, 16) HS_SPMM_CASE(2, 4, 16) HS_SPMM_CASE(2, 8, 12)
  HS_SPMM_CASE(4, 1, 16) HS_SPMM_CASE(4, 2, 16) HS_SPMM_CASE(4, 4, 12)
  HS_SPMM_CASE(8, 1, 16) HS_SPMM_CASE(8, 2, 12)
  { return hs_fail(ctx, HS_ERR_INVALID, "internal: no slab kernel for NV=%d K=%d W=%d", NV, K, kWarps); }
#undef HS_SPMM_CASE

  const int nacc = (G + 31) / 32;
  const unsigned eb = (unsigned)((n_rows + 7) / 8);
  const double* Q = (const double*)ctx->Q.p;
  const int np = perm ? n_perm : out_n_perm;
  const int64_t nsel = perm ? n_units : out_n_sel;
#define HS_EPI(NA) \
  k_p64_epilogue<NA><<<eb, 256, 0, st>>>(prm.P64, n_seg_all, seg_stride, n_rows, ldP, G, Q, np, nsel, flags, run_if, out_dev)
  if (nacc <= 1) HS_EPI(1);
  else if (nacc <= 2) HS_EPI(2);
  else if (nacc <= 3) HS_EPI(3);
  else if (nacc <= 4) HS_EPI(4);
  else if (nacc <= 8) HS_EPI(8);
  else if (nacc <= 16) HS_EPI(16);
  else HS_EPI(32);
#undef HS_EPI
  HS_LAUNCH_CHECK(ctx);
  if (flag_dev_out) *flag_dev_out = flags;
  return HS_OK;
}

// Randomization batch, sparse form: bucket every permuted row by slab, then run the slab kernel over
// the n_sel * n_perm rows.  Returns HS_ERR_STATE-free HS_OK, or a negative "not applicable" marker (-1)
// when the cell axis has too many slabs for the shared-memory histogram (the caller then gathers).
// ---- K4-sparse, seeded form: draw the permutations on the device, already ordered by cell ---------------
// One CTA per permuted row walks the cells in order; a thread takes kGenPerThread consecutive cells, asks the
// row's keyed permutation for i = pi^-1(c) and keeps the cell when i < nnz, paired with the gene's i-th value
// (hs_prp.cuh).  A block-wide scan turns the keep counts into output positions, so the (cell, value) pairs
// come out ascending by cell: exactly what the slab kernel wants, with no ids uploaded and nothing sorted,
// and the fp32 summation order of a row is fixed (bit-identical repeats).
constexpr int kGenThreads = 256;
constexpr int kGenPerThread = 8;

__global__ void __launch_bounds__(kGenThreads)
k_perm_generate(const int64_t* __restrict__ rp, const int64_t* __restrict__ gp, int n_perm,
                const float* __restrict__ val, int64_t N, uint64_t seed, uint64_t stream_base,
                const unsigned long long* __restrict__ unit_stream, int small_max,
                uint2* __restrict__ out) {
  __shared__ unsigned int s_warp[kGenThreads / 32];
  __shared__ unsigned int s_total, s_cnt;
  __shared__ uint2 s_list[kGenThreads * kGenPerThread];      // (value to walk on, slot) of the cells that left the domain
  __shared__ uint32_t s_res[kGenThreads * kGenPerThread];
  const int64_t row = blockIdx.x;
  const int64_t s = row / n_perm;
  const int64_t g0 = gp[s];
  const uint32_t nnz = (uint32_t)(gp[s + 1] - g0);
  if (nnz <= (uint32_t)small_max) return;   // empty, or drawn forward by k_perm_generate_small
  // stream of (unit s, randomization r): stream_base + s * n_perm + r, or unit_stream[s] + r when the caller names
  // the units' streams (units dealt to ranks keep the streams of the whole job)
  const uint64_t stream = unit_stream ? (uint64_t)unit_stream[s] + (uint64_t)(row - s * n_perm) : stream_base + (uint64_t)row;
  const HsPrp prp = hs_prp_make(seed, stream, (uint32_t)N);
  uint2* dst = out + rp[row];
  const float* v = val + g0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  unsigned int running = 0;
  const uint32_t n32 = (uint32_t)N;
  for (uint32_t c0 = 0; c0 < n32; c0 += kGenThreads * kGenPerThread) {
    const uint32_t cb = c0 + threadIdx.x * kGenPerThread;
    uint32_t idx[kGenPerThread];
    // One decryption per cell.  The few results outside [0, N) (the network's domain is a power of four)
    // must be decrypted again until they fall inside; a per-thread loop would make every warp pay a second
    // pass for its one or two stragglers, so they are queued and finished by the first threads of the block.
    unsigned int walk = 0;
    if (c0 + kGenThreads * kGenPerThread <= n32) {   // whole chunk inside the population (block-uniform)
#pragma unroll
      for (int j = 0; j < kGenPerThread; ++j) {
        idx[j] = hs_prp_dec1(prp, cb + j);
        walk |= (idx[j] >= n32 ? 1u : 0u) << j;
      }
    } else {
#pragma unroll
      for (int j = 0; j < kGenPerThread; ++j) {
        idx[j] = 0xffffffffu;
        if (cb + j < n32) {
          idx[j] = hs_prp_dec1(prp, cb + j);
          walk |= (idx[j] >= n32 ? 1u : 0u) << j;
        }
      }
    }
    if (walk) {
      unsigned int q = atomicAdd(&s_cnt, (unsigned int)__popc(walk));
#pragma unroll
      for (int j = 0; j < kGenPerThread; ++j)
        if ((walk >> j) & 1u) s_list[q++] = make_uint2(idx[j], threadIdx.x * kGenPerThread + j);
    }
    __syncthreads();
    for (unsigned int t = threadIdx.x; t < s_cnt; t += kGenThreads) {
      const uint2 e = s_list[t];
      uint32_t x = e.x;
      do x = hs_prp_dec1(prp, x); while (x >= n32);
      s_res[e.y] = x;
    }
    __syncthreads();
    unsigned int keep = 0;
#pragma unroll
    for (int j = 0; j < kGenPerThread; ++j) {
      if ((walk >> j) & 1u) idx[j] = s_res[threadIdx.x * kGenPerThread + j];
      keep += idx[j] < nnz;
    }
    // exclusive scan of `keep` over the block, in thread order
    unsigned int incl = keep;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      unsigned int w = lane < kGenThreads / 32 ? s_warp[lane] : 0u;
      unsigned int wi = w;
#pragma unroll
      for (int o = 1; o < kGenThreads / 32; o <<= 1) {
        const unsigned int t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      if (lane < kGenThreads / 32) s_warp[lane] = wi - w;
      if (lane == kGenThreads / 32 - 1) s_total = wi;
      if (lane == 0) s_cnt = 0;
    }
    __syncthreads();
    unsigned int pos = running + s_warp[warp] + incl - keep;
    running += s_total;
#pragma unroll
    for (int j = 0; j < kGenPerThread; ++j)
      if (idx[j] < nnz) dst[pos++] = make_uint2(cb + j, __float_as_uint(__ldg(v + idx[j])));
    // (the next chunk rewrites s_warp / s_total only after two more block-wide barriers)
  }
}

// Rows with few stored values: walking all N cells for a handful of draws is wasteful, so these rows draw
// forward (pi(0 .. nnz-1), one thread per few draws) and a block-wide radix sort orders the pairs by cell.
// Same set of pairs in the same (ascending) order as k_perm_generate would emit.
constexpr int kSmallThreads = 512;
constexpr int kSmallItems = 16;
constexpr int kSmallRowMax = kSmallThreads * kSmallItems;   // 8192 stored values
using SmallSort = cub::BlockRadixSort<uint32_t, kSmallThreads, kSmallItems, uint32_t>;

__global__ void __launch_bounds__(kSmallThreads)
k_perm_generate_small(const int64_t* __restrict__ rp, const int64_t* __restrict__ gp, int n_perm,
                      const float* __restrict__ val, int64_t N, uint64_t seed, uint64_t stream_base,
                      const unsigned long long* __restrict__ unit_stream,
                      uint2* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char sort_smem[];
  SmallSort::TempStorage& tmp = *reinterpret_cast<SmallSort::TempStorage*>(sort_smem);
  const int64_t row = blockIdx.x;
  const int64_t s = row / n_perm;
  const int64_t g0 = gp[s];
  const int64_t nnz64 = gp[s + 1] - g0;
  if (nnz64 == 0 || nnz64 > kSmallRowMax) return;
  const uint32_t nnz = (uint32_t)nnz64;
  // stream of (unit s, randomization r): stream_base + s * n_perm + r, or unit_stream[s] + r when the caller names
  // the units' streams (units dealt to ranks keep the streams of the whole job)
  const uint64_t stream = unit_stream ? (uint64_t)unit_stream[s] + (uint64_t)(row - s * n_perm) : stream_base + (uint64_t)row;
  const HsPrp prp = hs_prp_make(seed, stream, (uint32_t)N);
  uint32_t keys[kSmallItems], vals[kSmallItems];
#pragma unroll
  for (int k = 0; k < kSmallItems; ++k) {
    const uint32_t i = threadIdx.x * kSmallItems + k;
    keys[k] = 0xffffffffu;   // padding sorts to the end
    vals[k] = 0u;
    if (i < nnz) {
      keys[k] = hs_prp_forward(prp, i);
      vals[k] = __float_as_uint(__ldg(val + g0 + i));
    }
  }
  SmallSort(tmp).Sort(keys, vals);
  uint2* dst = out + rp[row];
#pragma unroll
  for (int k = 0; k < kSmallItems; ++k) {
    const uint32_t i = threadIdx.x * kSmallItems + k;
    if (i < nnz) dst[i] = make_uint2(keys[k], vals[k]);
  }
}

int hs_run_perm_seeded(hs_ctx* ctx, const int64_t* gene_ptr_dev, int64_t n_sel, int64_t n_vals_total,
                       const float* val_dev, int n_perm, uint64_t seed, uint64_t stream_base, double* out_dev,
                       const unsigned long long* unit_stream_dev) {
  const int64_t n_rows = n_sel * (int64_t)n_perm;
  if (n_rows <= 0) return HS_OK;
  if (n_rows > 0x7fffffffLL) return hs_fail(ctx, HS_ERR_INVALID, "too many randomized rows in one call");
  if (hs_slab_cells(ctx) < 4)
    return hs_fail(ctx, HS_ERR_INVALID, "seeded randomization: the density rows do not fit the slab kernel");
  const int64_t n_ids = n_vals_total * n_perm;
  cudaStream_t st = ctx->stream;
  // ---- tile path: the draw goes straight into the tensor-core kernel's tiles, one permutation evaluation per stored
  //      value (hs_tiles.cu: hs_run_perm_tiles); the order-independent gather is chained behind it on the device flag ----
  bool fused = hs_tile_path_enabled(ctx) && (n_rows + 127) / 128 <= 65535;
  if (const char* e = getenv("HS_PERM_FUSED")) fused = fused && atoi(e) != 0;
  if (fused) {
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortPairsDescending(nullptr, cub_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                              (const int32_t*)nullptr, (int32_t*)nullptr, (int)n_rows, 0, 32, st);
    const size_t keys_b = ((size_t)n_rows * 4 + 255) & ~(size_t)255;
    HS_TRY(hs_ensure(ctx, ctx->spmm_ws, 256 + 4 * keys_b + ((cub_bytes + 255) & ~(size_t)255)));
    char* w = (char*)ctx->spmm_ws.p;
    int* flags = (int*)w;                    w += 256;
    uint32_t* keys_in = (uint32_t*)w;        w += keys_b;
    uint32_t* keys_out = (uint32_t*)w;       w += keys_b;
    int32_t* rows_in = (int32_t*)w;          w += keys_b;
    int32_t* rows_out = (int32_t*)w;         w += keys_b;
    HS_CUDA(ctx, cudaMemsetAsync(flags, 0, 256, st));
    k_row_len<true><<<(unsigned)((n_rows + 255) / 256), 256, 0, st>>>(gene_ptr_dev, n_rows, n_perm, keys_in, rows_in, flags);
    HS_LAUNCH_CHECK(ctx);
    HS_CUDA(ctx, cub::DeviceRadixSort::SortPairsDescending(w, cub_bytes, keys_in, keys_out, rows_in, rows_out, (int)n_rows,
                                                           0, 32, st));
    ctx->launches += 3;
    int* tile_flag = flags + 2;
    const int32_t* tile_rows = hs_tile_row_order(ctx, rows_out, n_rows, rows_in);
    const int rc = hs_run_perm_tiles(ctx, gene_ptr_dev, n_sel, n_vals_total, val_dev, n_perm, seed, stream_base,
                                     unit_stream_dev, tile_rows, out_dev, tile_flag);
    if (rc == HS_OK) {
      if (getenv("HS_PERM_CHECK")) {      // development aid: did the draw hand the call to the gather kernel?
        int h[4] = {0, 0, 0, 0};
        cudaStreamSynchronize(st);
        cudaMemcpy(h, flags, sizeof(h), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[perm check] rows %lld flags %d %d %d %d\n", (long long)n_rows, h[0], h[1], h[2], h[3]);
      }
      return hs_run_gather_seeded(ctx, gene_ptr_dev, val_dev, n_sel, n_perm, seed, stream_base, unit_stream_dev, tile_flag,
                                  out_dev);
    }
    if (rc != -1) return rc;
  }
  HS_TRY(hs_ensure(ctx, ctx->perm_rp, sizeof(int64_t) * (size_t)(n_rows + 1)));
  HS_TRY(hs_ensure(ctx, ctx->perm_cols, sizeof(uint2) * (size_t)hs_max<int64_t>(n_ids, 1)));   // (cell, value) pairs
  int64_t* rp = (int64_t*)ctx->perm_rp.p;
  k_perm_row_ptr<<<(unsigned)((n_rows + 256) / 256), 256, 0, st>>>(gene_ptr_dev, n_sel, n_perm, rp);
  HS_LAUNCH_CHECK(ctx);
  // every row is offered to both kernels; each takes the rows of its size class and leaves the others
  int small_max = kSmallRowMax;
  if (const char* e = getenv("HS_PERM_SMALL_ROW")) small_max = hs_max(0, hs_min(atoi(e), kSmallRowMax));
  if (small_max > 0) {
    const size_t sm = sizeof(SmallSort::TempStorage);
    HS_CUDA(ctx, cudaFuncSetAttribute(k_perm_generate_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_perm_generate_small<<<(unsigned)n_rows, kSmallThreads, sm, st>>>(rp, gene_ptr_dev, n_perm, val_dev, ctx->N, seed,
                                                                     stream_base, unit_stream_dev, (uint2*)ctx->perm_cols.p);
    HS_LAUNCH_CHECK(ctx);
  }
  k_perm_generate<<<(unsigned)n_rows, kGenThreads, 0, st>>>(rp, gene_ptr_dev, n_perm, val_dev, ctx->N, seed, stream_base,
                                                            unit_stream_dev, small_max, (uint2*)ctx->perm_cols.p);
  HS_LAUNCH_CHECK(ctx);
  const int* flag = nullptr;
  // the pairs are ascending by cell and in range: the tensor-core tile path takes them as they are (its slab-kernel
  // fallback covers values beyond fp16's range)
  HS_TRY(hs_run_slab_spmm(ctx, false, rp, n_rows, 1, nullptr, nullptr, 0, out_dev, n_perm, n_sel,
                          /*allow_tiles=*/true, &flag, ctx->perm_cols.p, n_ids));
  return HS_OK;
}

int hs_run_perm_sorted(hs_ctx* ctx, const int64_t* gene_ptr_dev, int64_t n_sel, int64_t n_vals_total,
                       const float* val_dev, const int32_t* ind_dev, int n_perm, int index_base, double* out_dev,
                       const int** fallback_flag_out) {
  const int64_t n_rows = n_sel * (int64_t)n_perm;
  if (fallback_flag_out) *fallback_flag_out = nullptr;
  if (n_rows <= 0) return HS_OK;
  int S = hs_slab_cells(ctx);
  // Tile path: bucket by a power-of-two number of 64-cell slabs instead, so that the pairs are ordered at the
  // granularity of the tile kernels' 4096-cell chunks (they need no order inside a chunk).
  bool tiles = fallback_flag_out != nullptr && hs_tile_path_enabled(ctx) && (n_rows + 127) / 128 <= 65535;
  if (tiles) {
    int s2 = 64;
    while (s2 < 4096 && ((ctx->N + s2 - 1) / s2 + 2) * 4 > 200 * 1024) s2 *= 2;
    if (((ctx->N + s2 - 1) / s2 + 2) * 4 > 200 * 1024) tiles = false; else S = s2;
  }
  if (S < 4) return -1;
  const int64_t n_slabs = (ctx->N + S - 1) / S;
  const int n_bins = (int)n_slabs + 1;                     // + overflow bin
  const size_t smem = sizeof(unsigned int) * (size_t)(n_bins + 1);
  if (smem > 200 * 1024) return -1;
  const int64_t n_ids = n_vals_total * n_perm;
  HS_TRY(hs_ensure(ctx, ctx->perm_rp, sizeof(int64_t) * (size_t)(n_rows + 1)));
  HS_TRY(hs_ensure(ctx, ctx->perm_cols, sizeof(uint2) * (size_t)hs_max<int64_t>(n_ids, 1)));   // (id, value) pairs
  cudaStream_t st = ctx->stream;
  int64_t* rp = (int64_t*)ctx->perm_rp.p;
  k_perm_row_ptr<<<(unsigned)((n_rows + 256) / 256), 256, 0, st>>>(gene_ptr_dev, n_sel, n_perm, rp);
  HS_LAUNCH_CHECK(ctx);
  HS_CUDA(ctx, cudaFuncSetAttribute(k_perm_slab_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const uint64_t magic = ((1ull << 40) + (uint64_t)S - 1) / (uint64_t)S;
  if (n_rows * kSortMaxParts > 0x7fffffffLL) return -1;
  int part_items = kSortPartItemsDefault;
  if (const char* e = getenv("HS_SORT_PART_ITEMS")) part_items = hs_max(1024, atoi(e));
  k_perm_slab_sort<<<(unsigned)(n_rows * kSortMaxParts), kSortThreads, smem, st>>>(
      rp, gene_ptr_dev, n_perm, val_dev, ind_dev, index_base, ctx->N, magic, n_bins, part_items, (uint2*)ctx->perm_cols.p);
  HS_LAUNCH_CHECK(ctx);
  if (tiles)      // ids outside [0, N) or values beyond fp16 raise *fallback_flag_out: the caller then gathers
    return hs_run_slab_spmm(ctx, false, rp, n_rows, 1, nullptr, nullptr, 0, out_dev, n_perm, n_sel,
                            /*allow_tiles=*/true, fallback_flag_out, ctx->perm_cols.p, n_ids, /*tile_only=*/true);
  return hs_run_slab_spmm(ctx, false, rp, n_rows, 1, nullptr, nullptr, 0, out_dev, n_perm, n_sel,
                          /*allow_tiles=*/false /* ids are only slab-sorted */, nullptr, ctx->perm_cols.p);
}

// =================================================================================================================
// Gene selection on the device (R/haystack_continuous.R:222, :227-236): coeffVar = sd / mean of the row statistics
// (each with the reference's + 1e-300, :82-83), order(coeffVar) -- stable, NaN last, as R's order() and numpy's stable
// argsort -- and the rows at the requested ranks.  The ranks themselves are the reference's integer arithmetic on
// seq() and stay with the host (haystack.py: select_genes_to_randomize).
// =================================================================================================================
namespace {
__global__ void k_cv_keys(const double* __restrict__ mean, const double* __restrict__ sd, int64_t n,
                          double* __restrict__ cv_out, unsigned long long* __restrict__ keys, int32_t* __restrict__ idx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double cv = (sd[i] + 1e-300) / (mean[i] + 1e-300);
  if (cv == 0.0) cv = 0.0;                                   // -0 and +0 compare equal in order(): one key
  if (cv_out) cv_out[i] = cv;
  unsigned long long b = (unsigned long long)__double_as_longlong(cv);
  b = (b >> 63) ? ~b : (b | 0x8000000000000000ull);          // ascending unsigned order = ascending double order
  if (cv != cv) b = ~0ull;                                   // NaN: last, in index order (the sort is stable)
  keys[i] = b;
  idx[i] = (int32_t)i;
}
__global__ void k_pick_ranks(const int32_t* __restrict__ order, const int64_t* __restrict__ ranks, int64_t n_ranks, int64_t n,
                             int64_t* __restrict__ rows, int* __restrict__ flags) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_ranks) return;
  const int64_t r = ranks[k];
  if (r < 0 || r >= n) {
    atomicOr(flags, 1);
    rows[k] = -1;
    return;
  }
  rows[k] = order[r];
}
}  // namespace

int hs_run_order_by_cv(hs_ctx* ctx, const double* mean_dev, const double* sd_dev, int64_t n, const int64_t* ranks_dev,
                       int64_t n_ranks, double* cv_dev_out, int64_t* rows_dev_out, int* flag_dev) {
  if (n <= 0 || n > 0x7fffffffLL) return hs_fail(ctx, HS_ERR_INVALID, "gene selection: bad number of rows");
  cudaStream_t st = ctx->stream;
  size_t cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, (int)n, 0, 64, st);
  const size_t k_b = ((size_t)n * 8 + 255) & ~(size_t)255, i_b = ((size_t)n * 4 + 255) & ~(size_t)255;
  HS_TRY(hs_ensure(ctx, ctx->spmm_ws, 2 * k_b + 2 * i_b + ((cub_bytes + 255) & ~(size_t)255)));
  char* w = (char*)ctx->spmm_ws.p;
  unsigned long long* keys_in = (unsigned long long*)w;   w += k_b;
  unsigned long long* keys_out = (unsigned long long*)w;  w += k_b;
  int32_t* idx_in = (int32_t*)w;                          w += i_b;
  int32_t* idx_out = (int32_t*)w;                         w += i_b;
  k_cv_keys<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(mean_dev, sd_dev, n, cv_dev_out, keys_in, idx_in);
  HS_LAUNCH_CHECK(ctx);
  HS_CUDA(ctx, cub::DeviceRadixSort::SortPairs(w, cub_bytes, keys_in, keys_out, idx_in, idx_out, (int)n, 0, 64, st));
  ctx->launches += 3;
  k_pick_ranks<<<(unsigned)((n_ranks + 255) / 256), 256, 0, st>>>(idx_out, ranks_dev, n_ranks, n, rows_dev_out, flag_dev);
  HS_LAUNCH_CHECK(ctx);
  return HS_OK;
}
