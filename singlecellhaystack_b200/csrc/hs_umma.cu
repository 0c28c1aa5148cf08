// K2 / K4-dense: gene x grid contraction on the 5th-generation tensor cores (tcgen05 + TMEM).
//   P[row, j] = sum_c A[row, c] * D[c, j]
//     dense:  A[row, c] = E[row + c * ld]                 R/haystack_continuous.R:357 (colSums(D * w))
//     perm :  A[row, c] = v[perm[c + row * N] - base]     R/haystack_continuous.R:259 (sample(x = v))
//
// Precision: a single 16-bit operand pass violates the 1e-5 D_KL tolerance, so both
// operands are split into two fp16 terms, x = hi + lo * 2^-11 (22 significant bits), and three MMAs per
// K step accumulate   acc0 += A_hi * D_hi,   acc1 += A_hi * D_lo + A_lo * D_hi   in fp32 TMEM.  The tensor
// core truncates (rounds toward zero) every accumulating MMA, a bias that grows with the number of MMAs
// summed into one TMEM value and does not cancel between grid columns (measured: 32 MMAs per sum give
// 4e-5 relative D_KL error, 4 MMAs give 5e-6 = the fp32 SIMT kernel's level).  TMEM is therefore drained
// after every stage (64 cells) into round-to-nearest fp32 register sums, and those every 1024 cells into
// fp64 (global atomics): the same three levels as the CSR kernel.  The operand scales (s_E = 2^-3 for A, 2^14 for D) keep realistic expression values and all
// densities >= 4e-9 in fp16's normal range; an A value that still overflows fp16 raises a device flag
// and the caller reruns the fp32 SIMT kernel instead (no silent loss).
//
// Layout: both operands are MN-major (E is gene-contiguous, D is grid-column-contiguous), 128B swizzle.
//   element (mn, k) of a 128 x 64 tile at byte  ((k/8)*2 + mn/64) * 1024 + (k%8) * 128
//                                             + (((mn%64)/8) ^ (k%8)) * 16 + (mn%8) * 2
//   => descriptor LBO = 1024 B (between the two 64-wide MN atoms), SBO = 2048 B (between groups of 8 k).
// D_hi / D_lo are pre-tiled in HBM in exactly this layout by k_density_tiles (one 16 KB block per 64
// cells), so a stage's B operand is one cp.async.bulk per term (1-D TMA, no tensor map).
// Warp roles (16 warps): 0-3 build the A tiles (load, split, swizzled st.shared, fence.proxy.async),
// 4 streams the B tiles, 5 owns TMEM and issues the MMAs, 8-15 drain TMEM and accumulate.
#include <cuda_fp16.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "hs_common.cuh"
#include "hs_ptx.cuh"

namespace {

constexpr int kTileM = 128;          // rows (genes / permutations) per CTA
constexpr int kTileN = 128;          // grid columns per CTA (G <= 128)
constexpr int kBK = 64;              // cells per pipeline stage
constexpr int kStagesDense = 3;   // dense / perm producers
constexpr int kTileBytes = kTileM * kBK * 2;          // 16 KB per fp16 operand term
constexpr int kStageBytes = 4 * kTileBytes;           // A_hi, A_lo, B_hi, B_lo
constexpr int kDrainBlocks = 1;      // TMEM -> registers after every stage (64 cells = 4 accumulating MMAs per sum)
constexpr int kFlushDrains = 16;     // registers -> fp64 every 16 drains (1024 cells)
constexpr int kUmmaThreads = 512;
constexpr float kScaleA = 1.0f;              // expression values go in unscaled (fp16 overflow -> fp32 kernel)
constexpr float kScaleD = 16384.f;           // 2^14
constexpr float kLoScale = 2048.f;           // 2^11
constexpr uint32_t kLbo = 1024, kSbo = 2048;

using namespace hs_ptx;

// shared-memory matrix descriptor: MN-major, 128B swizzle, LBO / SBO as above, sm_100 version bit
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)(kLbo >> 4) << 16) | ((uint64_t)(kSbo >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D = F32, A = B = F16, both MN-major, N = 128, M = 128
constexpr uint32_t kIdesc = (1u << 4) | (0u << 7) | (0u << 10) | (1u << 15) | (1u << 16) | ((kTileN >> 3) << 17) |
                            ((kTileM >> 4) << 24);

// byte offset of the 16-byte chunk holding elements (8*mg .. 8*mg+7, k) of a 128 x 64 MN-major SW128 tile
__device__ __host__ __forceinline__ uint32_t chunk_off(int mg, int k) {
  const int mi = mg >> 3, c8 = mg & 7, kg = k >> 3, k8 = k & 7;
  return (uint32_t)(((kg * 2 + mi) << 10) + (k8 << 7) + ((c8 ^ k8) << 4));
}

// split 8 scaled floats into fp16 hi / lo (lo pre-multiplied by 2^11) and pack each as one 16-byte chunk.
// Returns true when a finite value does not fit fp16 (|x| > 65504; Inf inputs take the same exit and are
// then handled by the fp32 kernel exactly like the reference's arithmetic would).
__device__ __forceinline__ bool split8(const float (&x)[8], uint4& hi, uint4& lo) {
  __half2 h[4], l[4];
  float amax = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float a = x[2 * i], b = x[2 * i + 1];
    amax = fmaxf(amax, fmaxf(fabsf(a), fabsf(b)));
    h[i] = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h[i]);
    l[i] = __floats2half2_rn((a - hf.x) * kLoScale, (b - hf.y) * kLoScale);
  }
  hi = make_uint4(*reinterpret_cast<uint32_t*>(&h[0]), *reinterpret_cast<uint32_t*>(&h[1]),
                  *reinterpret_cast<uint32_t*>(&h[2]), *reinterpret_cast<uint32_t*>(&h[3]));
  lo = make_uint4(*reinterpret_cast<uint32_t*>(&l[0]), *reinterpret_cast<uint32_t*>(&l[1]),
                  *reinterpret_cast<uint32_t*>(&l[2]), *reinterpret_cast<uint32_t*>(&l[3]));
  return amax > 65504.f;
}
__device__ __forceinline__ void sts_u4(uint32_t a, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---- pre-tiled fp16 density operands ---------------------------------------------------------------
// D32 [N][ldD] fp32  ->  Dh, Dl: one 16 KB tile per 64 cells, rows = grid columns (zero beyond G / N)
__global__ void __launch_bounds__(256)
k_density_tiles(const float* __restrict__ D32, int64_t N, int ldD, int G, unsigned char* __restrict__ Dh,
                unsigned char* __restrict__ Dl) {
  const int64_t kb = blockIdx.x;
  // 16 column groups x 64 cells = 1024 chunks per tile
  for (int task = threadIdx.x; task < 1024; task += 256) {
    const int mg = task & 15, k = task >> 4;
    const int64_t cell = kb * kBK + k;
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int j = 8 * mg + i;
      x[i] = (cell < N && j < G) ? D32[cell * ldD + j] * kScaleD : 0.f;
    }
    uint4 hi, lo;
    split8(x, hi, lo);
    const size_t off = (size_t)kb * kTileBytes + chunk_off(mg, k);
    *reinterpret_cast<uint4*>(Dh + off) = hi;
    *reinterpret_cast<uint4*>(Dl + off) = lo;
  }
}

// ---- main kernel -----------------------------------------------------------------------------------
struct UmmaParams {
  // A operand source
  const float* E;          // dense: [rows + cells * ld]
  int64_t ld;
  const float* v;          // perm: values [N]
  const int32_t* perm;     // perm: [N x n_rows] column-major
  int index_base;
  const int64_t* ptr;
  const int32_t* cols;
  const float* vals;
  int col_base;
  int64_t n_rows;          // rows covered by the tiles (dense / perm: matrix rows; csr: dense-ish rows)
  int64_t N;               // cells
  // B operand
  const unsigned char* Dh;
  const unsigned char* Dl;
  // split-K
  int64_t kblocks_per_seg;
  // output
  double* P64;             // [segments][rows][ldP] fp64 sums, zeroed by the caller; one writer per
                           // (segment, row, column), segments are added in order by the epilogue kernel
  int64_t seg_stride;      // doubles between segments
  int ldP;
  int G;
  long long* dbg;          // optional (HS_UMMA_DEBUG): per-role wait cycles of CTA (0,0), 16 slots
  int* flags;              // bit 0: caller must rerun on the fallback kernel (fp16 overflow of an A value;
                           //        csr: column ids not ascending)
};

enum { kModeDense = 0, kModePerm = 1, kModeDenseTma = 3 };
constexpr int kRawSlots = 3;                       // dense-tma: ring of raw fp32 tiles (64 cells x 128 rows)
constexpr int kRawBytes = kBK * kTileM * 4;        // 32 KB

__device__ __forceinline__ float4 lds_f4(uint32_t a) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(a) : "memory");
  return r;
}

template <int MODE>
__global__ void __launch_bounds__(kUmmaThreads, 1) k_dense_umma(const UmmaParams p) {
  constexpr bool PERM = MODE == kModePerm;
  constexpr bool TMA = MODE == kModeDenseTma;
  constexpr int kStages = TMA ? 2 : kStagesDense;
  // dense mode uses the two otherwise idle warps (6, 7) as extra producers; dense-tma uses warp 6 as a
  // producer and warp 7 to stream raw fp32 tiles of E into a shared-memory ring with cp.async.bulk
  constexpr int kProducerWarps = MODE == kModeDense ? 6 : (TMA ? 5 : 4);
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // 1024-byte aligned operand stages (swizzle atoms), then barriers
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  const uint32_t ring_base = smem_u32(smem + kStages * kStageBytes + 256);   // after stages + barriers
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_u32(bars);
  auto full_a = [&](int s) { return bar_base + 8u * s; };
  auto full_b = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto empty = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
  auto tmem_full = [&](int b) { return bar_base + 8u * (3 * kStages + b); };
  auto tmem_empty = [&](int b) { return bar_base + 8u * (3 * kStages + 2 + b); };
  auto raw_full = [&](int r) { return bar_base + 8u * (3 * kStages + 4 + r); };                 // dense-tma only
  auto raw_empty = [&](int r) { return bar_base + 8u * (3 * kStages + 4 + kRawSlots + r); };
  const uint32_t raw_base = ring_base;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // development aid: cycles each role spends blocked on each barrier (CTA (0,0) only)
  const bool dbg_on = p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0;
  long long w_acc[3] = {0, 0, 0};
  auto timed_wait = [&](uint32_t bar, uint32_t parity, int slot) {
    if (dbg_on) {
      const long long t0 = clock64();
      mbar_wait(bar, parity);
      w_acc[slot] += clock64() - t0;
    } else {
      mbar_wait(bar, parity);
    }
  };
  const long long t_begin = clock64();
  const int64_t row0 = (int64_t)blockIdx.x * kTileM;
  const int64_t total_kb = (p.N + kBK - 1) / kBK;
  const int64_t kb_begin = (int64_t)blockIdx.y * p.kblocks_per_seg;
  const int64_t kb_end = hs_min<int64_t>(total_kb, kb_begin + p.kblocks_per_seg);
  const int n_kb = (int)hs_max<int64_t>(0, kb_end - kb_begin);
  const int n_drains = (n_kb + kDrainBlocks - 1) / kDrainBlocks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_a(s), kProducerWarps);
      mbar_init(full_b(s), 1);
      mbar_init(empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tmem_full(b), 1);
      mbar_init(tmem_empty(b), 8);
    }
    if (TMA)
      for (int r = 0; r < kRawSlots; ++r) {
        mbar_init(raw_full(r), 1);
        mbar_init(raw_empty(r), kProducerWarps);
      }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (TMA && warp == 7) {
    // ================= raw E tiles: one 512-byte bulk copy per cell, three stages ahead =================
    {
      // all 32 lanes issue copies (two cells each): a single thread cannot issue 64 small bulk copies per
      // stage fast enough
      const uint32_t row_bytes = (uint32_t)hs_min<int64_t>(kTileM, p.n_rows - row0) * 4u;
      for (int i = 0; i < n_kb; ++i) {
        const int r = i % kRawSlots;
        if (i >= kRawSlots) timed_wait(raw_empty(r), ((i / kRawSlots) - 1) & 1, 0);
        const int64_t c0 = (kb_begin + i) * kBK;
        const int ncell = (int)hs_min<int64_t>(kBK, p.N - c0);
        if (lane == 0) mbar_arrive_expect_tx(raw_full(r), (uint32_t)ncell * row_bytes);
        __syncwarp();
        const uint32_t dst = raw_base + (uint32_t)r * kRawBytes;
        for (int k = lane; k < ncell; k += 32)
          bulk_g2s(dst + (uint32_t)k * (kTileM * 4), p.E + row0 + (c0 + k) * p.ld, row_bytes, raw_full(r));
      }
    }
  } else if (warp < 4 || (MODE == kModeDense && (warp == 6 || warp == 7)) || (TMA && warp == 6)) {
    // ================= A producers =================
    bool over = false;
    for (int i = 0; i < n_kb; ++i) {
      const int s = i % kStages;
      if (i >= kStages) timed_wait(empty(s), ((i / kStages) - 1) & 1, 0);
      const uint32_t a_hi = smem_base + s * kStageBytes, a_lo = a_hi + kTileBytes;
      const int64_t c0 = (kb_begin + i) * kBK;
      if (TMA) {
        // the raw tile of this stage sits in shared memory (warp 7 requested it three stages ago): read,
        // split, write the swizzled operand chunks.  160 producer threads, 6-7 tasks each.
        const int pt = (warp < 4 ? warp : 4) * 32 + lane;                    // 0..159
        const int r = i % kRawSlots;
        timed_wait(raw_full(r), (i / kRawSlots) & 1, 1);
        const uint32_t raw = raw_base + (uint32_t)r * kRawBytes;
        const bool full_tile = (c0 + kBK <= p.N) && (row0 + kTileM <= p.n_rows);   // no stale bytes to mask
#pragma unroll
        for (int q = 0; q < 7; ++q) {
          const int t = pt + 160 * q;
          if (t < 1024) {
            const int mg = t & 15, k = t >> 4;
            const float4 u = lds_f4(raw + (uint32_t)t * 32u);            // == k * 512 + mg * 32
            const float4 w = lds_f4(raw + (uint32_t)t * 32u + 16u);
            float x[8] = {u.x, u.y, u.z, u.w, w.x, w.y, w.z, w.w};
            if (!full_tile) {
              const bool cell_ok = c0 + k < p.N;
              const int64_t rr = row0 + 8 * mg;
#pragma unroll
              for (int e = 0; e < 8; ++e) x[e] = (cell_ok && rr + e < p.n_rows) ? x[e] : 0.f;
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) x[e] *= kScaleA;
            uint4 hi, lo;
            over |= split8(x, hi, lo);
            const uint32_t off = chunk_off(mg, k);
            sts_u4(a_hi + off, hi);
            sts_u4(a_lo + off, lo);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(raw_empty(r));
      } else if (!PERM) {
        // task = (row group of 8, cell): lanes 0-15 cover 128 rows of one cell, 16-31 the next cell.
        // 192 producer threads, 5-6 tasks each.  All of a stage's loads are issued before the first
        // conversion (one memory round trip per stage instead of one per task), and the lines the stage
        // four steps ahead will read are prefetched into L2 so that the round trip is an L2 hit.
        const int pt = (warp < 4 ? warp : warp - 2) * 32 + lane;            // 0..191
        const bool vec = ((p.ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.E) & 15) == 0);
        constexpr int kMaxTasks = (1024 + 191) / 192;                        // 6
        float4 bu[kMaxTasks], bw[kMaxTasks];
        {
          const int64_t pc0 = c0 + 4 * kBK;
#pragma unroll
          for (int q = 0; q < kMaxTasks; ++q) {
            const int t = pt + 192 * q;
            const int64_t cell = pc0 + (t >> 4), r = row0 + 8 * (t & 15);
            if (t < 1024 && cell < p.N && r < p.n_rows && (lane & 3) == 0)
              asm volatile("prefetch.global.L2 [%0];" ::"l"(p.E + r + cell * p.ld));
          }
        }
#pragma unroll
        for (int q = 0; q < kMaxTasks; ++q) {
          const int t = pt + 192 * q;
          const int mg = t & 15, k = t >> 4;
          const int64_t cell = c0 + k, r = row0 + 8 * mg;
          bu[q] = bw[q] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (t < 1024 && cell < p.N) {
            if (r + 8 <= p.n_rows && vec) {
              bu[q] = __ldg(reinterpret_cast<const float4*>(p.E + r + cell * p.ld));
              bw[q] = __ldg(reinterpret_cast<const float4*>(p.E + r + 4 + cell * p.ld));
            } else {
              float x[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) x[e] = (r + e < p.n_rows) ? __ldg(p.E + r + e + cell * p.ld) : 0.f;
              bu[q] = make_float4(x[0], x[1], x[2], x[3]);
              bw[q] = make_float4(x[4], x[5], x[6], x[7]);
            }
          }
        }
#pragma unroll
        for (int q = 0; q < kMaxTasks; ++q) {
          const int t = pt + 192 * q;
          if (t < 1024) {
            const float x[8] = {bu[q].x * kScaleA, bu[q].y * kScaleA, bu[q].z * kScaleA, bu[q].w * kScaleA,
                                bw[q].x * kScaleA, bw[q].y * kScaleA, bw[q].z * kScaleA, bw[q].w * kScaleA};
            uint4 hi, lo;
            over |= split8(x, hi, lo);
            const uint32_t off = chunk_off(t & 15, t >> 4);
            sts_u4(a_hi + off, hi);
            sts_u4(a_lo + off, lo);
          }
        }
      } else {
        // task = (row group of 8, cell): lanes run over 32 consecutive cells (coalesced index loads)
        for (int t = threadIdx.x; t < 1024; t += 128) {
          const int k = (t & 31) + 32 * ((t >> 5) & 1), mg = t >> 6;
          const int64_t cell = c0 + k, r = row0 + 8 * mg;
          float x[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float a = 0.f;
            if (cell < p.N && r + e < p.n_rows) {
              const int64_t src = (int64_t)__ldg(p.perm + cell + (r + e) * p.N) - p.index_base;
              a = (src >= 0 && src < p.N) ? __ldg(p.v + src) : __int_as_float(0x7fc00000);
            }
            x[e] = a * kScaleA;
          }
          uint4 hi, lo;
          over |= split8(x, hi, lo);
          const uint32_t off = chunk_off(mg, k);
          sts_u4(a_hi + off, hi);
          sts_u4(a_lo + off, lo);
        }
      }
      fence_proxy_async();      // generic-proxy stores -> visible to the tensor core's async proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(full_a(s));
    }
    if (__any_sync(0xffffffffu, over) && lane == 0) atomicOr(p.flags, 1);
  } else if (warp == 4) {
    // ================= B loader =================
    {
      for (int i = 0; i < n_kb; ++i) {
        const int s = i % kStages;
        if (i >= kStages) timed_wait(empty(s), ((i / kStages) - 1) & 1, 0);   // (L2 prefetching the tiles six stages ahead was tried: slower)
        if (lane == 0) {
          const uint32_t b_hi = smem_base + s * kStageBytes + 2 * kTileBytes;
          const size_t off = (size_t)(kb_begin + i) * kTileBytes;
          mbar_arrive_expect_tx(full_b(s), 2 * kTileBytes);
          bulk_g2s(b_hi, p.Dh + off, kTileBytes, full_b(s));
          bulk_g2s(b_hi + kTileBytes, p.Dl + off, kTileBytes, full_b(s));
        }
        __syncwarp();
      }
    }
  } else if (warp == 5) {
    // ================= MMA issuer =================
    if (lane == 0) {
      for (int i = 0; i < n_kb; ++i) {
        const int s = i % kStages;
        const int d = i / kDrainBlocks, buf = d & 1;
        const bool first = (i % kDrainBlocks) == 0;
        if (first && d >= 2) {
          timed_wait(tmem_empty(buf), ((d >> 1) - 1) & 1, 2);
          tc_fence_after();
        }
        timed_wait(full_a(s), (i / kStages) & 1, 0);
        timed_wait(full_b(s), (i / kStages) & 1, 1);
        tc_fence_after();
        const uint32_t st = smem_base + s * kStageBytes;
        const uint32_t acc0 = tmem_base + (uint32_t)(buf * 256), acc1 = acc0 + 128;
#pragma unroll
        for (int k16 = 0; k16 < kBK / 16; ++k16) {
          const uint32_t koff = (uint32_t)k16 * 2u * kSbo;
          const uint64_t a_hi = make_desc(st + koff), a_lo = make_desc(st + kTileBytes + koff);
          const uint64_t b_hi = make_desc(st + 2 * kTileBytes + koff), b_lo = make_desc(st + 3 * kTileBytes + koff);
          const uint32_t accum = (first && k16 == 0) ? 0u : 1u;
          umma_f16(acc0, a_hi, b_hi, kIdesc, accum);
          umma_f16(acc1, a_hi, b_lo, kIdesc, accum);
          umma_f16(acc1, a_lo, b_hi, kIdesc, 1u);
        }
        umma_commit(empty(s));                                   // stage free once these MMAs retire
        if ((i % kDrainBlocks) == kDrainBlocks - 1 || i == n_kb - 1) umma_commit(tmem_full(buf));
      }
    }
  } else if (warp >= 8) {
    // ================= epilogue: TMEM -> fp32 registers -> fp64 =================
    const int q = warp & 3, half = (warp - 8) >> 2;            // TMEM lane quarter, column half
    const int64_t trow = row0 + 32 * q + lane;
    const int64_t row = trow;
    const int col0 = 64 * half;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * q) << 16);
    float acc2[64];
#pragma unroll
    for (int c = 0; c < 64; ++c) acc2[c] = 0.f;
    const double unscale = 1.0 / ((double)kScaleA * (double)kScaleD);
    auto flush = [&]() {
      if (row >= 0 && trow < p.n_rows) {
        double* prow = p.P64 + blockIdx.y * p.seg_stride + row * p.ldP;
#pragma unroll
        for (int c = 0; c < 64; ++c)
          if (col0 + c < p.G) atomicAdd(prow + col0 + c, (double)acc2[c] * unscale);   // single writer: order fixed
      }
#pragma unroll
      for (int c = 0; c < 64; ++c) acc2[c] = 0.f;
    };
    for (int d = 0; d < n_drains; ++d) {
      const int buf = d & 1;
      timed_wait(tmem_full(buf), (d >> 1) & 1, 0);
      tc_fence_after();
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        if (col0 + 16 * h >= p.G) continue;                  // columns past the grid are zero padding
        uint32_t a[16], m[16];                               // both loads in flight, one wait
        tmem_ld16_nowait(lane_addr + (uint32_t)(buf * 256 + 128 + col0 + 16 * h), a);   // cross terms, scale 2^-11
        tmem_ld16_nowait(lane_addr + (uint32_t)(buf * 256 + col0 + 16 * h), m);         // hi * hi
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 16; ++c)
          acc2[16 * h + c] += fmaf(__uint_as_float(a[c]), 1.0f / kLoScale, __uint_as_float(m[c]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty(buf));
      if ((d % kFlushDrains) == kFlushDrains - 1) flush();
    }
    flush();
  }
  if (dbg_on && lane == 0) {
    // slots: [0] total cycles; producer warp 0: [1] empty/raw_empty, [2] raw_full; raw loader (warp 7): [3];
    // B loader: [4]; MMA: [5] full_a, [6] full_b, [7] tmem_empty; epilogue warp 8: [8] tmem_full
    if (warp == 0) { p.dbg[0] = clock64() - t_begin; p.dbg[1] = w_acc[0]; p.dbg[2] = w_acc[1]; }
    if (warp == 7) p.dbg[3] = w_acc[0];
    if (warp == 4) p.dbg[4] = w_acc[0];
    if (warp == 5) { p.dbg[5] = w_acc[0]; p.dbg[6] = w_acc[1]; p.dbg[7] = w_acc[2]; }
    if (warp == 8) p.dbg[8] = w_acc[0];
    if (warp == 0) p.dbg[9] = n_kb;
  }
  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

constexpr size_t kUmmaSmem = 1024 + (size_t)kStagesDense * kStageBytes + 256;
constexpr size_t kUmmaSmemTma = 1024 + 2 * (size_t)kStageBytes + 256 + (size_t)kRawSlots * kRawBytes;

template <int NACC>
__global__ void __launch_bounds__(256)
k_p64_kl(const double* __restrict__ P64, int n_seg, int64_t seg_stride, int64_t n_rows, int ldP, int G,
         const double* __restrict__ Q, const int* __restrict__ flags, double* __restrict__ out) {
  if (*flags & 1) return;   // the SIMT rerun produces the output
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
  if (lane == 0) out[row] = kl;
}

}  // namespace

// (Re)build the pre-tiled fp16 density operands for the resident state.
int hs_ensure_density_tiles(hs_ctx* ctx) {
  if (ctx->tiles_valid) return HS_OK;
  const int64_t total_kb = (ctx->N + kBK - 1) / kBK;
  HS_TRY(hs_ensure(ctx, ctx->dtile_hi, (size_t)total_kb * kTileBytes));
  HS_TRY(hs_ensure(ctx, ctx->dtile_lo, (size_t)total_kb * kTileBytes));
  k_density_tiles<<<(unsigned)total_kb, 256, 0, ctx->stream>>>((const float*)ctx->D32.p, ctx->N, ctx->ldD, ctx->G,
                                                               (unsigned char*)ctx->dtile_hi.p,
                                                               (unsigned char*)ctx->dtile_lo.p);
  HS_LAUNCH_CHECK(ctx);
  ctx->tiles_valid = true;
  return HS_OK;
}

// Tensor-core dense contraction + KL.  Returns -1 when the shape is not supported (G > 128): the caller
// then uses the SIMT kernel.  On success *flag_dev_out points at the device flag whose bit 0 asks for a
// SIMT rerun (fp16 overflow of an expression value).
int hs_run_dense_umma(hs_ctx* ctx, bool perm, const float* e_dev, int64_t ld, const float* v_dev,
                      const int32_t* perm_dev, int index_base, int64_t n_rows, double* out_dev,
                      const int** flag_dev_out) {
  if (n_rows <= 0) return HS_OK;
  if (ctx->G > kTileN) return -1;
  HS_TRY(hs_ensure_density_tiles(ctx));
  const int ldP = kTileN;
  HS_TRY(hs_ensure(ctx, ctx->umma_flags, 256));
  cudaStream_t st = ctx->stream;
  const int row_tiles = (int)((n_rows + kTileM - 1) / kTileM);
  const int64_t total_kb = (ctx->N + kBK - 1) / kBK;
  // split the cell axis until the grid covers the machine; segments are whole drain windows
  int64_t n_seg = hs_max<int64_t>(1, ctx->sm_count / row_tiles);
  int64_t kb_per_seg = (total_kb + n_seg - 1) / n_seg;
  kb_per_seg = (kb_per_seg + kDrainBlocks - 1) / kDrainBlocks * kDrainBlocks;
  n_seg = (total_kb + kb_per_seg - 1) / kb_per_seg;
  const int64_t seg_stride = n_rows * (int64_t)ldP;
  HS_TRY(hs_ensure(ctx, ctx->part, sizeof(double) * (size_t)n_seg * (size_t)seg_stride));
  HS_CUDA(ctx, cudaMemsetAsync(ctx->part.p, 0, sizeof(double) * (size_t)n_seg * (size_t)seg_stride, st));
  HS_CUDA(ctx, cudaMemsetAsync(ctx->umma_flags.p, 0, 256, st));
  UmmaParams prm = {};
  prm.E = e_dev;
  prm.ld = ld;
  prm.v = v_dev;
  prm.perm = perm_dev;
  prm.index_base = index_base;
  prm.n_rows = n_rows;
  prm.N = ctx->N;
  prm.Dh = (const unsigned char*)ctx->dtile_hi.p;
  prm.Dl = (const unsigned char*)ctx->dtile_lo.p;
  prm.kblocks_per_seg = kb_per_seg;
  prm.P64 = (double*)ctx->part.p;
  prm.seg_stride = seg_stride;
  prm.ldP = ldP;
  prm.G = ctx->G;
  prm.flags = (int*)ctx->umma_flags.p;
  const bool dbg = getenv("HS_UMMA_DEBUG") != nullptr;
  prm.dbg = dbg ? (long long*)((char*)ctx->umma_flags.p + 64) : nullptr;
  dim3 grid(row_tiles, (unsigned)n_seg);
  if (perm) {
    HS_CUDA(ctx, cudaFuncSetAttribute(k_dense_umma<kModePerm>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kUmmaSmem));
    k_dense_umma<kModePerm><<<grid, kUmmaThreads, kUmmaSmem, st>>>(prm);
  } else if ((ld & 3) == 0 && (n_rows & 3) == 0 && (reinterpret_cast<uintptr_t>(e_dev) & 15) == 0 && !getenv("HS_DENSE_NO_TMA")) {
    // every cell's run of rows is 16-byte aligned and sized: raw tiles can come through cp.async.bulk
    HS_CUDA(ctx, cudaFuncSetAttribute(k_dense_umma<kModeDenseTma>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kUmmaSmemTma));
    k_dense_umma<kModeDenseTma><<<grid, kUmmaThreads, kUmmaSmemTma, st>>>(prm);
  } else {
    HS_CUDA(ctx, cudaFuncSetAttribute(k_dense_umma<kModeDense>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kUmmaSmem));
    k_dense_umma<kModeDense><<<grid, kUmmaThreads, kUmmaSmem, st>>>(prm);
  }
  HS_LAUNCH_CHECK(ctx);
  if (dbg) {
    long long h[16] = {0};
    cudaStreamSynchronize(st);
    cudaMemcpy(h, prm.dbg, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[umma dbg] blocks %lld total %lld clk (%.0f/block) | producer: wait empty %lld, raw_full %lld | raw loader wait %lld | "
                    "B loader wait %lld | mma: full_a %lld full_b %lld tmem_empty %lld | epilogue wait tmem_full %lld\n",
            h[9], h[0], h[9] ? (double)h[0] / (double)h[9] : 0.0, h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8]);
  }
  const unsigned eb = (unsigned)((n_rows + 7) / 8);
  k_p64_kl<4><<<eb, 256, 0, st>>>((const double*)ctx->part.p, (int)n_seg, seg_stride, n_rows, ldP, ctx->G, (const double*)ctx->Q.p,
                                  (const int*)ctx->umma_flags.p, out_dev);
  HS_LAUNCH_CHECK(ctx);
  if (flag_dev_out) *flag_dev_out = (const int*)ctx->umma_flags.p;
  return HS_OK;
}

