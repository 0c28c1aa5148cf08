// C ABI of libhaystack_b200.so (declared in include/haystack_b200.h): context management,
// host<->device staging for the host-pointer entry points, and dispatch to the kernel launchers.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "hs_common.cuh"
#include "hs_prp.cuh"

static thread_local std::string g_global_err;

// No C++ exception may cross the C ABI (host allocations of the staging code can throw).
#define HS_ABI_CATCH(ctxp)                                                                       \
  catch (const std::bad_alloc&) { return hs_fail((ctxp), HS_ERR_NOMEM, "host allocation failed"); } \
  catch (const std::exception& e) { return hs_fail((ctxp), HS_ERR_INVALID, "internal error: %s", e.what()); } \
  catch (...) { return hs_fail((ctxp), HS_ERR_INVALID, "internal error"); }

int hs_fail(hs_ctx* ctx, int code, const char* fmt, ...) {
  try {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx)
    ctx->err = buf;
  else
    g_global_err = buf;
  return code;
  } HS_ABI_CATCH(ctx)
}

int hs_ensure(hs_ctx* ctx, DevBuf& b, size_t bytes) {
  try {
  if (bytes == 0) bytes = 16;
  if (b.cap >= bytes) return HS_OK;
  if (b.p) {
    // the old block may still be in use by queued work
    HS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    HS_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
    cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
  }
  const size_t want = bytes + bytes / 8 + 256;
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    e = cudaMalloc(&b.p, bytes);
    if (e != cudaSuccess) {
      cudaGetLastError();
      b.p = nullptr;
      return hs_fail(ctx, HS_ERR_NOMEM, "cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    }
    b.cap = bytes;
  } else {
    b.cap = want;
  }
  return HS_OK;
  } HS_ABI_CATCH(ctx)
}

// ---- host worker threads ---------------------------------------------------------------------------
// R hands the values over as doubles (dgRMatrix@x); the kernels read them as fp32.  Narrowing on the host
// halves the bytes of the values on PCIe (the contraction of a 1M-cell matrix is transfer-bound end to end),
// and is done by a few persistent threads into a pinned ring while earlier buffers are in flight.
struct HostPool {
  std::vector<std::thread> threads;
  std::mutex mu;
  std::condition_variable cv_work, cv_done;
  std::function<void(int)> job;   // job(worker index)
  uint64_t epoch = 0;
  int pending = 0;
  bool stop = false;

  explicit HostPool(int n) {
    for (int i = 0; i < n; ++i) threads.emplace_back([this, i] { loop(i); });
  }
  ~HostPool() {
    {
      std::lock_guard<std::mutex> lk(mu);
      stop = true;
    }
    cv_work.notify_all();
    for (auto& t : threads) t.join();
  }
  void loop(int idx) {
    uint64_t seen = 0;
    for (;;) {
      std::function<void(int)> fn;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_work.wait(lk, [&] { return stop || epoch != seen; });
        if (stop) return;
        seen = epoch;
        fn = job;
      }
      fn(idx);
      {
        std::lock_guard<std::mutex> lk(mu);
        if (--pending == 0) cv_done.notify_all();
      }
    }
  }
  // runs fn(0 .. n_workers-1) on the workers and returns when all are done
  void run(const std::function<void(int)>& fn) {
    std::unique_lock<std::mutex> lk(mu);
    job = fn;
    pending = (int)threads.size();
    ++epoch;
    cv_work.notify_all();
    cv_done.wait(lk, [&] { return pending == 0; });
  }
  int size() const { return (int)threads.size(); }
};

// One host-pointer call running on a worker thread of the library (hs_dkl_perm_csr_begin / hs_dkl_perm_csr_wait): the
// caller's thread is free to draw the next batch of ids meanwhile.  The context stays single-user: every other entry
// point refuses to run until the call has been waited for.
struct HsAsync {
  std::thread th;
  std::thread::id worker;
  bool pending = false;
  int rc = HS_OK;
  std::vector<int64_t> gene_ptr;
  std::vector<double> result;
};

// processes of this job on this node (torchrun exports LOCAL_WORLD_SIZE); 1 when not launched that way
static int local_ranks() {
  const char* e = getenv("LOCAL_WORLD_SIZE");
  const int lw = e ? atoi(e) : 1;
  return lw > 1 ? lw : 1;
}

static int host_threads() {
  int n = (int)std::thread::hardware_concurrency();
  if (n <= 0) n = 1;
  // one process per GPU: the ranks of a node share its cores (torchrun exports LOCAL_WORLD_SIZE); eight ranks
  // with sixteen threads each on a 32-core host spent their time context switching
  if (const char* e = getenv("LOCAL_WORLD_SIZE")) {
    const int lw = atoi(e);
    if (lw > 1) n = std::max(1, n / lw);
  }
  n = std::min(n, 16);
  if (const char* e = getenv("HS_HOST_THREADS")) {
    const int v = atoi(e);
    if (v > 0) n = std::min(v, 256);
  }
  return n;
}

// dst[i] = (float)src[i] (round to nearest even, as the plain cast).  With AVX2: four doubles per conversion and
// streaming stores -- the narrowed values are read next by the DMA engine, not by this core (measured on the GPU
// box's 16 hardware threads: 13.5 instead of 9.2 G values/s).
#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target("avx2"))) static void narrow_block_avx2(const double* src, float* dst, int64_t n) {
  int64_t i = 0;
  while (i < n && ((uintptr_t)(dst + i) & 31)) {
    dst[i] = (float)src[i];
    ++i;
  }
  for (; i + 8 <= n; i += 8) {
    const __m128 a = _mm256_cvtpd_ps(_mm256_loadu_pd(src + i)), b = _mm256_cvtpd_ps(_mm256_loadu_pd(src + i + 4));
    _mm256_stream_ps(dst + i, _mm256_set_m128(b, a));
  }
  for (; i < n; ++i) dst[i] = (float)src[i];
  _mm_sfence();
}
static const bool g_have_avx2 = __builtin_cpu_supports("avx2");
#else
static const bool g_have_avx2 = false;
static void narrow_block_avx2(const double*, float*, int64_t) {}
#endif
static void narrow_block(const double* src, float* dst, int64_t n) {
  if (g_have_avx2) {
    narrow_block_avx2(src, dst, n);
    return;
  }
  for (int64_t i = 0; i < n; ++i) dst[i] = (float)src[i];
}

// split over the pool when the block is large
static void narrow_f64(hs_ctx* ctx, const double* src, float* dst, int64_t n) {
  if (n < (int64_t)1 << 18 || host_threads() == 1) {
    narrow_block(src, dst, n);
    return;
  }
  if (!ctx->pool) ctx->pool = new HostPool(host_threads());
  const int nt = ctx->pool->size();
  const int64_t per = ((n + nt - 1) / nt + 15) & ~(int64_t)15;
  ctx->pool->run([=](int t) {
    const int64_t a = std::min<int64_t>(n, (int64_t)t * per), b = std::min<int64_t>(n, a + per);
    if (b > a) narrow_block(src + a, dst + a, b - a);
  });
}

// dst[i] = src[i] over the pool: moves a block of pageable caller memory into a pinned ring buffer at host memory
// bandwidth (a cudaMemcpyAsync from pageable memory is staged by the driver on one thread and blocks the caller)
static void copy_bytes(hs_ctx* ctx, const void* src, void* dst, size_t n) {
  if (n < ((size_t)1 << 20) || host_threads() == 1) {
    memcpy(dst, src, n);
    return;
  }
  if (!ctx->pool) ctx->pool = new HostPool(host_threads());
  const int nt = ctx->pool->size();
  const size_t per = ((n + nt - 1) / nt + 63) & ~(size_t)63;
  ctx->pool->run([=](int t) {
    const size_t a = std::min(n, (size_t)t * per), b = std::min(n, a + per);
    if (b > a) memcpy((char*)dst + a, (const char*)src + a, b - a);
  });
}

// Cell ids of rows [r0, r1) as bitmaps over the cells (`words` 64-bit words per row) into dst; the rows are split over
// the pool.  Returns false when a row's ids are not strictly ascending inside [0, n_cells): such input is uploaded as it
// is (the kernels score unsorted rows through their gather fallback).
static bool pack_ids(hs_ctx* ctx, const int32_t* ids, const int64_t* hp, int64_t r0, int64_t r1, int64_t n_cells,
                     int64_t words, unsigned long long* dst) {
  std::atomic<bool> ok{true};
  auto rows = [&](int t, int nt) {
    for (int64_t r = r0 + t; r < r1; r += nt) {
      unsigned long long* w = dst + (r - r0) * words;
      memset(w, 0, sizeof(unsigned long long) * (size_t)words);
      const int32_t* a = ids + hp[r];
      const int64_t n = hp[r + 1] - hp[r];
      // branch-free: the word under construction is kept in a register and stored after every id (ids ascend, so a
      // word's ids are consecutive; the stores of one word hit the same line)
      int64_t prev = -1, prev_w = -1;
      unsigned long long acc = 0ull;
      bool good = true;
      int64_t i = 0;
      for (; i < n; ++i) {
        const int64_t c = a[i];
        if (!(c > prev && c < n_cells)) { good = false; break; }      // never taken on valid input
        const int64_t wi = c >> 6;
        acc = (wi == prev_w ? acc : 0ull) | (1ull << (c & 63));
        w[wi] = acc;
        prev_w = wi;
        prev = c;
      }
      if (!good) {
        ok.store(false, std::memory_order_relaxed);
        return;
      }
    }
  };
  if (hp[r1] - hp[r0] < ((int64_t)1 << 18) || host_threads() == 1) {
    rows(0, 1);
  } else {
    if (!ctx->pool) ctx->pool = new HostPool(host_threads());
    const int nt = ctx->pool->size();
    ctx->pool->run([&](int t) { rows(t, nt); });
  }
  return ok.load();
}

// true when a host pointer is ordinary pageable memory (not cudaHostAlloc / cudaHostRegister memory)
static bool is_pageable(const void* p) {
  if (!p) return false;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return true;
  }
  return at.type == cudaMemoryTypeUnregistered;
}

static int ensure_ring(hs_ctx* ctx, size_t bytes) {
  if (ctx->ring_cap >= bytes) return HS_OK;
  for (int k = 0; k < hs_ctx::kRing; ++k) {
    if (ctx->ring_busy[k]) {
      HS_CUDA(ctx, cudaEventSynchronize(ctx->ev_ring[k]));
      ctx->ring_busy[k] = false;
    }
    if (ctx->ring[k]) cudaFreeHost(ctx->ring[k]);
    ctx->ring[k] = nullptr;
  }
  ctx->ring_cap = 0;
  for (int k = 0; k < hs_ctx::kRing; ++k) {
    HS_CUDA(ctx, cudaHostAlloc(&ctx->ring[k], bytes, cudaHostAllocDefault));
    if (!ctx->ev_ring[k]) HS_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_ring[k], cudaEventDisableTiming));
  }
  ctx->ring_cap = bytes;
  return HS_OK;
}

static void hs_free(DevBuf& b) {
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
}

namespace {

struct ScopedTimer {
  hs_ctx* c;
  int slot;
  bool on;
  ScopedTimer(hs_ctx* ctx, int s) : c(ctx), slot(s), on(ctx->timers) {
    if (on) cudaEventRecord(c->ev_a, c->stream);
  }
  // call after the work has been queued; synchronises only when timers are enabled
  void stop() {
    if (!on) return;
    cudaEventRecord(c->ev_b, c->stream);
    cudaEventSynchronize(c->ev_b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev_a, c->ev_b);
    c->t_ms[slot] = ms;
    on = false;
  }
};

int need_density(hs_ctx* ctx) {
  if (!ctx) return hs_fail(nullptr, HS_ERR_INVALID, "null context");
  if (ctx->N <= 0 || ctx->G <= 0 || !ctx->D32.p)
    return hs_fail(ctx, HS_ERR_STATE, "no density stage resident: call hs_density first");
  return HS_OK;
}

int set_device(hs_ctx* ctx) {
  if (ctx->async && ctx->async->pending && std::this_thread::get_id() != ctx->async->worker)
    return hs_fail(nullptr, HS_ERR_STATE, "an asynchronous call is in flight on this context: call hs_dkl_perm_csr_wait first");
  HS_CUDA(ctx, cudaSetDevice(ctx->device));
  return HS_OK;
}

size_t env_size(const char* name, size_t dflt) {
  const char* s = getenv(name);
  if (!s || !*s) return dflt;
  const long long v = atoll(s);
  return v > 0 ? (size_t)v : dflt;
}

}  // namespace

#pragma GCC visibility push(default)
extern "C" {

int hs_abi_version(void) { return HS_ABI_VERSION; }

const char* hs_last_error(const hs_ctx* ctx) { return ctx ? ctx->err.c_str() : g_global_err.c_str(); }

int hs_create(int device, hs_ctx** out) {
  if (!out) return hs_fail(nullptr, HS_ERR_INVALID, "hs_create: out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0) {
    cudaGetLastError();
    return hs_fail(nullptr, HS_ERR_CUDA, "no CUDA device available (%s); libhaystack_b200 has no CPU fallback",
                   e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  }
  if (device < 0 || device >= count)
    return hs_fail(nullptr, HS_ERR_INVALID, "device %d out of range (0..%d)", device, count - 1);
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return hs_fail(nullptr, HS_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    return hs_fail(nullptr, HS_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a (B200) only",
                   device, prop.major, prop.minor);
  hs_ctx* ctx = new (std::nothrow) hs_ctx();
  if (!ctx) return hs_fail(nullptr, HS_ERR_NOMEM, "out of host memory");
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->smem_optin = prop.sharedMemPerBlockOptin;
  bool ok = cudaSetDevice(device) == cudaSuccess &&
            cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreate(&ctx->ev_a) == cudaSuccess && cudaEventCreate(&ctx->ev_b) == cudaSuccess;
  for (int i = 0; ok && i < 2; ++i)
    ok = cudaEventCreateWithFlags(&ctx->ev_stage[i], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&ctx->ev_done[i], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&ctx->ev_bits[i], cudaEventDisableTiming) == cudaSuccess;
  if (!ok) {
    const char* msg = cudaGetErrorString(cudaGetLastError());
    hs_fail(nullptr, HS_ERR_CUDA, "context creation failed: %s", msg);
    hs_destroy(ctx);
    return HS_ERR_CUDA;
  }
  ctx->stream = ctx->own_stream;
  *out = ctx;
  return HS_OK;
}

void hs_destroy(hs_ctx* ctx) {
  if (!ctx) return;
  if (ctx->async) {
    if (ctx->async->th.joinable()) ctx->async->th.join();
    delete ctx->async;
    ctx->async = nullptr;
  }
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
  DevBuf* bufs[] = {&ctx->D32, &ctx->Q, &ctx->sq, &ctx->rowmin, &ctx->nearest, &ctx->sel, &ctx->qpart,
                    &ctx->dens64, &ctx->xin, &ctx->gin, &ctx->win, &ctx->part, &ctx->out,
                    &ctx->stage_j[0], &ctx->stage_j[1], &ctx->stage_x[0], &ctx->stage_x[1], &ctx->vals32,
                    &ctx->ptrs, &ctx->spmm_ws, &ctx->spmm_slots, &ctx->cols32, &ctx->dtile_hi, &ctx->dtile_lo, &ctx->umma_flags, &ctx->dense32, &ctx->simt_part, &ctx->perm_rp, &ctx->perm_cols, &ctx->perm_vals, &ctx->perm_gp, &ctx->out_rand,
                    &ctx->tile_ent, &ctx->tile_ws, &ctx->prof_x, &ctx->prof_out, &ctx->conv_cp, &ctx->conv_ri, &ctx->conv_x, &ctx->conv_p, &ctx->conv_j, &ctx->conv_xo,
                    &ctx->bits[0], &ctx->bits[1], &ctx->zeros, &ctx->dbg_times, &ctx->conv_sort};
  for (DevBuf* b : bufs) hs_free(*b);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  for (int k = 0; k < hs_ctx::kRing; ++k) {
    if (ctx->ring[k]) cudaFreeHost(ctx->ring[k]);
    if (ctx->ev_ring[k]) cudaEventDestroy(ctx->ev_ring[k]);
  }
  delete ctx->pool;
  if (ctx->ev_a) cudaEventDestroy(ctx->ev_a);
  if (ctx->ev_b) cudaEventDestroy(ctx->ev_b);
  for (int i = 0; i < 2; ++i) {
    if (ctx->ev_stage[i]) cudaEventDestroy(ctx->ev_stage[i]);
    if (ctx->ev_done[i]) cudaEventDestroy(ctx->ev_done[i]);
    if (ctx->ev_bits[i]) cudaEventDestroy(ctx->ev_bits[i]);
  }
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  delete ctx;
}

int hs_set_stream(hs_ctx* ctx, void* cuda_stream) {
  try {
  if (!ctx) return hs_fail(nullptr, HS_ERR_INVALID, "null context");
  HS_TRY(set_device(ctx));
  HS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return HS_OK;
  } HS_ABI_CATCH(ctx)
}

int hs_synchronize(hs_ctx* ctx) {
  try {
  if (!ctx) return hs_fail(nullptr, HS_ERR_INVALID, "null context");
  HS_TRY(set_device(ctx));
  HS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return HS_OK;
  } HS_ABI_CATCH(ctx)
}

int64_t hs_launch_count(const hs_ctx* ctx) { return ctx ? ctx->launches : 0; }

int hs_enable_timers(hs_ctx* ctx, int on) {
  try {
  if (!ctx) return hs_fail(nullptr, HS_ERR_INVALID, "null context");
  ctx->timers = on != 0;
  return HS_OK;
  } HS_ABI_CATCH(ctx)
}

int hs_last_timers(hs_ctx* ctx, float* ms_out, int n_slots) {
  try {
  if (!ctx || !ms_out) return 0;
  const int n = std::min(n_slots, (int)HS_T_COUNT);
  for (int i = 0; i < n; ++i) ms_out[i] = ctx->t_ms[i];
  return n;
  } HS_ABI_CATCH(ctx)
}

// ------------------------------------------------------------------------------------------------
// density stage
// ------------------------------------------------------------------------------------------------
static int density_common(hs_ctx* ctx, const double* x_dev, int64_t N, int d, const double* g_dev, int G,
                          const double* w_dev, double* bandwidth_out, double* q_out, double* density_out,
                          int32_t* nearest_out) {
  ScopedTimer tm(ctx, HS_T_DENSITY);
  HS_TRY(hs_run_density(ctx, x_dev, N, d, g_dev, G, w_dev, density_out != nullptr));
  tm.stop();
  cudaStream_t st = ctx->stream;
  if (q_out) HS_CUDA(ctx, cudaMemcpyAsync(q_out, ctx->Q.p, sizeof(double) * G, cudaMemcpyDeviceToHost, st));
  if (density_out)
    HS_CUDA(ctx, cudaMemcpyAsync(density_out, ctx->dens64.p, sizeof(double) * (size_t)N * G,
                                 cudaMemcpyDeviceToHost, st));
  if (nearest_out)
    HS_CUDA(ctx, cudaMemcpyAsync(nearest_out, ctx->nearest.p, sizeof(int32_t) * (size_t)N,
                                 cudaMemcpyDeviceToHost, st));
  HS_CUDA(ctx, cudaStreamSynchronize(st));
  if (bandwidth_out) *bandwidth_out = ctx->bw;
  return HS_OK;
}

static int check_density_args(hs_ctx* ctx, const void* x, int64_t N, int d, const void* grid, int G) {
  if (!ctx) return hs_fail(nullptr, HS_ERR_INVALID, "null context");
  if (!x || !grid) return hs_fail(ctx, HS_ERR_INVALID, "hs_density: x and grid must not be NULL");
  if (N <= 0 || d <= 0 || G <= 0)
    return hs_fail(ctx, HS_ERR_INVALID, "hs_density: n_cells=%lld n_dims=%d n_grid=%d must be positive",
                   (long long)N, d, G);
  if (G > 1024) return hs_fail(ctx, HS_ERR_INVALID, "hs_density: n_grid=%d exceeds the supported maximum of 1024", G);
  if (N > 2147483647LL)
    return hs_fail(ctx, HS_ERR_INVALID, "hs_density: n_cells=%lld exceeds int32 cell ids", (long long)N);
  return HS_OK;
}

int hs_density(hs_ctx* ctx, const double* x, int64_t n_cells, int n_dims, const double* grid, int n_grid,
               const double* w_q, double* bandwidth_out, double* q_out, double* density_out,
               int32_t* nearest_out) {
  try {
  HS_TRY(check_density_args(ctx, x, n_cells, n_dims, grid, n_grid));
  HS_TRY(set_device(ctx));
  ctx->N = 0;  // state is invalid until the stage completes
  HS_TRY(hs_ensure(ctx, ctx->xin, sizeof(double) * (size_t)n_cells * n_dims));
  HS_TRY(hs_ensure(ctx, ctx->gin, sizeof(double) * (size_t)n_grid * n_dims));
  cudaStream_t st = ctx->stream;
  HS_CUDA(ctx, cudaMemcpyAsync(ctx->xin.p, x, sizeof(double) * (size_t)n_cells * n_dims, cudaMemcpyHostToDevice, st));
  HS_CUDA(ctx, cudaMemcpyAsync(ctx->gin.p, grid, sizeof(double) * (size_t)n_grid * n_dims, cudaMemcpyHostToDevice, st));
  const double* w_dev = nullptr;
  if (w_q) {
    HS_TRY(hs_ensure(ctx, ctx->win, sizeof(double) * (size_t)n_cells));
    HS_CUDA(ctx, cudaMemcpyAsync(ctx->win.p, w_q, sizeof(double) * (size_t)n_cells, cudaMemcpyHostToDevice, st));
    w_dev = (const double*)ctx->win.p;
  }
  return density_common(ctx, (const double*)ctx->xin.p, n_cells, n_dims, (const double*)ctx->gin.p, n_grid, w_dev,
                        bandwidth_out, q_out, density_out, nearest_out);
  } HS_ABI_CATCH(ctx)
}

int hs_density_dev(hs_ctx* ctx, const double* x_dev, int64_t n_cells, int n_dims, const double* grid_dev,
                   int n_grid, const double* w_q_dev, double* bandwidth_out, double* q_out) {
  try {
  HS_TRY(check_density_args(ctx, x_dev, n_cells, n_dims, grid_dev, n_grid));
  HS_TRY(set_device(ctx));
  ctx->N = 0;
  return density_common(ctx, x_dev, n_cells, n_dims, grid_dev, n_grid, w_q_dev, bandwidth_out, q_out, nullptr,
                        nullptr);
  } HS_ABI_CATCH(ctx)
}

int hs_density_shape(const hs_ctx* ctx, int64_t* n_cells_out, int* n_grid_out) {
  try {
  if (!ctx) return hs_fail(nullptr, HS_ERR_INVALID, "null context");
  if (n_cells_out) *n_cells_out = ctx->N;
  if (n_grid_out) *n_grid_out = ctx->N > 0 ? ctx->G : 0;
  return HS_OK;
  } HS_ABI_CATCH(nullptr)
}

int hs_density_buffers(hs_ctx* ctx, int64_t n_cells, int n_grid, void** d_dev, int64_t* d_bytes, void** q_dev,
                       int64_t* q_bytes) {
  try {
  if (!ctx) return hs_fail(nullptr, HS_ERR_INVALID, "null context");
  if (n_cells <= 0 || n_grid <= 0 || n_grid > 1024 || n_cells > 2147483647LL || !d_dev || !q_dev)
    return hs_fail(ctx, HS_ERR_INVALID, "hs_density_buffers: bad arguments");
  HS_TRY(set_device(ctx));
  const int ldD = hs_ld_for_grid(n_grid);
  if (!(ctx->N == n_cells && ctx->G == n_grid && ctx->ldD == ldD && ctx->D32.p && ctx->Q.p)) {
    ctx->N = 0;
    HS_TRY(hs_ensure(ctx, ctx->Q, sizeof(double) * (size_t)n_grid + 64));
    HS_TRY(hs_ensure(ctx, ctx->D32, sizeof(float) * (size_t)n_cells * ldD));
  }
  *d_dev = ctx->D32.p;
  *q_dev = ctx->Q.p;
  if (d_bytes) *d_bytes = (int64_t)sizeof(float) * n_cells * ldD;
  if (q_bytes) *q_bytes = (int64_t)sizeof(double) * (n_grid + 1);
  return HS_OK;
  } HS_ABI_CATCH(ctx)
}

int hs_density_commit(hs_ctx* ctx, int64_t n_cells, int n_grid) {
  try {
  if (!ctx) return hs_fail(nullptr, HS_ERR_INVALID, "null context");
  HS_TRY(set_device(ctx));
  if (n_cells <= 0 || n_grid <= 0 || !ctx->D32.p || !ctx->Q.p)
    return hs_fail(ctx, HS_ERR_STATE, "hs_density_commit: call hs_density_buffers first");
  const int ldD = hs_ld_for_grid(n_grid);
  if (ctx->D32.cap < sizeof(float) * (size_t)n_cells * ldD || ctx->Q.cap < sizeof(double) * (size_t)(n_grid + 1))
    return hs_fail(ctx, HS_ERR_STATE, "hs_density_commit: buffers were reserved for a smaller shape");
  double bw = 0.0;
  HS_CUDA(ctx, cudaMemcpyAsync(&bw, (const double*)ctx->Q.p + n_grid, sizeof(double), cudaMemcpyDeviceToHost,
                               ctx->stream));
  HS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (!(bw > 0.0)) {
    ctx->N = 0;
    return hs_fail(ctx, HS_ERR_INVALID, "hs_density_commit: bandwidth %g in the received state is not positive", bw);
  }
  ctx->N = n_cells;
  ctx->G = n_grid;
  ctx->ldD = ldD;
  ctx->bw = bw;
  ctx->tiles_valid = false;
  return HS_OK;
  } HS_ABI_CATCH(ctx)
}

// ------------------------------------------------------------------------------------------------
// observed D_KL, host pointers
// ------------------------------------------------------------------------------------------------
// Host-pointer sparse entry points: upload the whole batch (ids as they are, values converted
// fp64 -> fp32 on the device through a double-buffered staging area fed by a second stream), then ONE
// launch over all rows -- the contraction kernels need many rows per launch to fill the machine.
// If the batch does not fit into free device memory the units are processed in several batches.
//   hp[u]  : prefix offsets in values per unit (row or selected gene); ids per unit = values * ipv
static int upload_and_run(hs_ctx* ctx, const int64_t* hp, int64_t n_units, const int32_t* ids, const void* vals_any,
                          bool vals_f32, int64_t ipv, bool perm, int n_perm, int index_base, double* dev_out,
                          int timer_slot) {
  const double* vals = (const double*)vals_any;       // when !vals_f32
  const float* vals32h = (const float*)vals_any;      // when vals_f32: straight into place, no conversion
  const int64_t v0 = hp[0], n_vals = hp[n_units] - v0, n_ids = n_vals * ipv;
  const size_t chunk = env_size("HS_STAGE_VALUES", (size_t)16 << 20);   // values per staging buffer
  HS_TRY(hs_ensure(ctx, ctx->ptrs, sizeof(int64_t) * (size_t)(n_units + 1)));
  HS_TRY(hs_ensure(ctx, ctx->cols32, sizeof(int32_t) * (size_t)hs_max<int64_t>(n_ids, 1)));
  HS_TRY(hs_ensure(ctx, ctx->vals32, sizeof(float) * (size_t)hs_max<int64_t>(n_vals, 1)));
  const size_t stage_vals = (size_t)hs_min<int64_t>((int64_t)chunk, hs_max<int64_t>(n_vals, 1));
  int ring_k = 0;
  // fp64 values: narrowed on the host into a pinned ring (default), or uploaded as doubles and narrowed by
  // a kernel (HS_NARROW_ON_DEVICE=1); both give the same fp32 bits
  // (with four host threads or fewer per rank the host is the bottleneck: the doubles cross the link and a kernel
  //  narrows them.  Measured with eight ranks on a 32-core host, 30 000 genes over all ranks: 363 ms per step narrowing
  //  on the host, 288 ms splitting the chunks between host and device, 238 ms narrowing on the device.  The ranks of a
  //  node also share its DRAM: host narrowing moves 20 bytes of host memory per value, the plain upload 12 -- four
  //  ranks on 32 cores: 398 ms narrowing on the host, 251 ms on the device; two ranks: 456 against 444 ms, a wash.)
  const bool narrow_on_device = getenv("HS_NARROW_ON_DEVICE") ? env_size("HS_NARROW_ON_DEVICE", 0) != 0
                                                              : (host_threads() <= 4 || local_ranks() >= 3);
  // R vectors (and numpy arrays) are pageable: their ids are bounced through the pinned ring by the host threads,
  // like the values, instead of being handed to cudaMemcpyAsync directly (HS_NO_ID_BOUNCE=1: old behaviour)
  const bool bounce_ids = !perm && !narrow_on_device && is_pageable(ids) && env_size("HS_NO_ID_BOUNCE", 0) == 0;
  const bool bounce_f32 = vals_f32 && bounce_ids;
  if (!vals_f32 && narrow_on_device) HS_TRY(hs_ensure(ctx, ctx->stage_x[0], sizeof(double) * stage_vals));
  // Rows with more than ~4 % stored values cross the link as bitmaps over the cells instead of 4-byte ids (pack_ids;
  // expanded again by hs_run_bitmap_to_ids): 1.25 instead of 4 bytes per id at 10 %.  Decided per piece of rows.
  const int64_t words = (ctx->N + 63) / 64;
  bool pack_ok = !perm && !narrow_on_device && env_size("HS_PACK_IDS", 0) != 0 && n_vals > 0;
  if ((!vals_f32 && !narrow_on_device) || bounce_ids || pack_ok) HS_TRY(ensure_ring(ctx, sizeof(float) * stage_vals));
  const int64_t rows_per_slot = (int64_t)(ctx->ring_cap / (sizeof(unsigned long long) * (size_t)hs_max<int64_t>(words, 1)));
  if (rows_per_slot < 1) pack_ok = false;
  // development aid (HS_UPLOAD_PROFILE=1): where the calling thread spends the upload -- waiting for a ring buffer
  // (the link is the limit), narrowing values, packing ids
  const bool prof = getenv("HS_UPLOAD_PROFILE") != nullptr;
  double t_wait = 0, t_narrow = 0, t_pack = 0;
  auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t_begin = prof ? now() : 0.0;
  auto ring_take = [&](int& k_out) -> int {
    const int k = ring_k;
    ring_k = (ring_k + 1) % hs_ctx::kRing;
    if (ctx->ring_busy[k]) {
      const double t0 = prof ? now() : 0.0;
      HS_CUDA(ctx, cudaEventSynchronize(ctx->ev_ring[k]));
      if (prof) t_wait += now() - t0;
      ctx->ring_busy[k] = false;
    }
    k_out = k;
    return HS_OK;
  };
  cudaStream_t st = ctx->stream, cs = ctx->copy_stream;

  // row pointers rebased to the uploaded arrays
  std::vector<int64_t> rel((size_t)n_units + 1);
  for (int64_t u = 0; u <= n_units; ++u) rel[u] = hp[u] - v0;
  // the copy stream must not overtake work of earlier calls that still reads the buffers
  if (!ctx->skip_copy_fence) {
    HS_CUDA(ctx, cudaEventRecord(ctx->ev_done[0], st));
    HS_CUDA(ctx, cudaStreamWaitEvent(cs, ctx->ev_done[0], 0));
  }
  // (on the copy stream: a pageable source makes this call wait for the copy, which on the compute stream
  // would queue behind whatever is running there; every piece's event covers it)
  HS_CUDA(ctx, cudaMemcpyAsync(ctx->ptrs.p, rel.data(), sizeof(int64_t) * (size_t)(n_units + 1), cudaMemcpyHostToDevice, cs));

  // Uploads (and the fp64 -> fp32 conversion of the values) run on the copy stream; the contraction of a
  // piece of rows starts on the compute stream as soon as that piece has landed, so PCIe transfer and
  // scoring overlap.  A piece holds enough rows to fill the machine; the permutation form is one piece
  // (its output block is batch-local).
  std::vector<int64_t> piece_end;
  if (!perm) {
    const int64_t min_rows = (int64_t)env_size("HS_PIECE_ROWS", (size_t)ctx->sm_count * 32);
    const int64_t target_vals = hs_max<int64_t>(n_vals / 8, (int64_t)env_size("HS_PIECE_VALUES", (size_t)32 << 20));
    int64_t u0 = 0;
    while (u0 < n_units) {
      int64_t u1 = hs_min<int64_t>(n_units, u0 + min_rows);
      while (u1 < n_units && rel[u1] - rel[u0] < target_vals) ++u1;
      if (n_units - u1 < min_rows) u1 = n_units;
      piece_end.push_back(u1);
      u0 = u1;
    }
  } else {
    piece_end.push_back(n_units);
  }
  ScopedTimer tm(ctx, timer_slot);
  const int64_t id_piece = (int64_t)64 << 20;
  int64_t u0 = 0;
  for (size_t pc = 0; pc < piece_end.size(); ++pc) {
    const int64_t u1 = piece_end[pc];
    const int64_t a0 = rel[u0], a1 = rel[u1];
    // Ids go straight into place; values are narrowed (fp64 sources) and follow.  The two are interleaved
    // chunk by chunk so that the link always has an id chunk to move while the host narrows the matching
    // values (the permutation form has few values and many ids: its ids go first, in pieces, so that
    // pageable sources do not serialise one huge copy).
    if (perm)
      for (int64_t off = a0 * ipv; off < a1 * ipv; off += id_piece) {
        const int64_t n = hs_min<int64_t>(id_piece, a1 * ipv - off);
        HS_CUDA(ctx, cudaMemcpyAsync((int32_t*)ctx->cols32.p + off, ids + v0 * ipv + off, sizeof(int32_t) * (size_t)n,
                                     cudaMemcpyHostToDevice, cs));
      }
    // ids [o, o + m) of the uploaded arrays as they are: bounced through the ring (pageable sources) or straight
    auto upload_ids_raw = [&](int64_t o, int64_t m) -> int {
      if (!bounce_ids) {
        HS_CUDA(ctx, cudaMemcpyAsync((int32_t*)ctx->cols32.p + o, ids + v0 + o, sizeof(int32_t) * (size_t)m,
                                     cudaMemcpyHostToDevice, cs));
        return HS_OK;
      }
      for (int64_t q = 0; q < m; q += (int64_t)stage_vals) {
        const int64_t nq = hs_min<int64_t>((int64_t)stage_vals, m - q);
        int k = 0;
        HS_TRY(ring_take(k));
        copy_bytes(ctx, ids + v0 + o + q, ctx->ring[k], sizeof(int32_t) * (size_t)nq);
        HS_CUDA(ctx, cudaMemcpyAsync((int32_t*)ctx->cols32.p + o + q, ctx->ring[k], sizeof(int32_t) * (size_t)nq,
                                     cudaMemcpyHostToDevice, cs));
        HS_CUDA(ctx, cudaEventRecord(ctx->ev_ring[k], cs));
        ctx->ring_busy[k] = true;
      }
      return HS_OK;
    };
    // this piece's ids as bitmaps?
    const int par = (int)(pc & 1);
    bool packed = pack_ok && (double)(u1 - u0) * (double)words * 8.0 <= 0.75 * 4.0 * (double)(a1 - a0);
    if (packed) {
      HS_TRY(hs_ensure(ctx, ctx->bits[par], sizeof(unsigned long long) * (size_t)((u1 - u0) * words)));
      if (ctx->bits_used[par]) HS_CUDA(ctx, cudaStreamWaitEvent(cs, ctx->ev_bits[par], 0));   // expanded two pieces ago
    }
    int64_t next_row = u0;
    for (int64_t off = a0; off < a1; off += (int64_t)stage_vals) {
      const int64_t n = hs_min<int64_t>((int64_t)stage_vals, a1 - off);
      if (!perm && packed) {
        // the rows that end inside this chunk of values (the last chunk takes what is left), a ring buffer at a time
        const int64_t lim = off + n >= a1 ? a1 : off + n;
        int64_t r_end = next_row;
        while (r_end < u1 && rel[r_end + 1] <= lim) ++r_end;
        if (off + n >= a1) r_end = u1;
        while (next_row < r_end) {
          const int64_t g1 = hs_min<int64_t>(r_end, next_row + rows_per_slot);
          int k = 0;
          HS_TRY(ring_take(k));
          const double t0 = prof ? now() : 0.0;
          const bool ok = pack_ids(ctx, ids, hp, next_row, g1, ctx->N, words, (unsigned long long*)ctx->ring[k]);
          if (prof) t_pack += now() - t0;
          if (!ok) {
            packed = false;      // unsorted or out-of-range ids: this piece goes up as it is
            break;
          }
          HS_CUDA(ctx, cudaMemcpyAsync((unsigned long long*)ctx->bits[par].p + (next_row - u0) * words, ctx->ring[k],
                                       sizeof(unsigned long long) * (size_t)((g1 - next_row) * words), cudaMemcpyHostToDevice, cs));
          HS_CUDA(ctx, cudaEventRecord(ctx->ev_ring[k], cs));
          ctx->ring_busy[k] = true;
          next_row = g1;
        }
        if (!packed) HS_TRY(upload_ids_raw(a0, off + n - a0));
      } else if (!perm) {
        HS_TRY(upload_ids_raw(off, n));
      }
      if (bounce_f32) {
        int k = 0;
        HS_TRY(ring_take(k));
        copy_bytes(ctx, vals32h + v0 + off, ctx->ring[k], sizeof(float) * (size_t)n);
        HS_CUDA(ctx, cudaMemcpyAsync((float*)ctx->vals32.p + off, ctx->ring[k], sizeof(float) * (size_t)n,
                                     cudaMemcpyHostToDevice, cs));
        HS_CUDA(ctx, cudaEventRecord(ctx->ev_ring[k], cs));
        ctx->ring_busy[k] = true;
      } else if (vals_f32) {
        HS_CUDA(ctx, cudaMemcpyAsync((float*)ctx->vals32.p + off, vals32h + v0 + off, sizeof(float) * (size_t)n,
                                     cudaMemcpyHostToDevice, cs));
      } else if (narrow_on_device) {
        HS_CUDA(ctx, cudaMemcpyAsync(ctx->stage_x[0].p, vals + v0 + off, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, cs));
        HS_TRY(hs_run_f64_to_f32(ctx, (const double*)ctx->stage_x[0].p, (float*)ctx->vals32.p + off, n, cs));
      } else {
        int k = 0;
        HS_TRY(ring_take(k));
        const double t0 = prof ? now() : 0.0;
        narrow_f64(ctx, vals + v0 + off, (float*)ctx->ring[k], n);
        if (prof) t_narrow += now() - t0;
        HS_CUDA(ctx, cudaMemcpyAsync((float*)ctx->vals32.p + off, ctx->ring[k], sizeof(float) * (size_t)n,
                                     cudaMemcpyHostToDevice, cs));
        HS_CUDA(ctx, cudaEventRecord(ctx->ev_ring[k], cs));
        ctx->ring_busy[k] = true;
      }
    }
    HS_CUDA(ctx, cudaEventRecord(ctx->ev_stage[0], cs));
    HS_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_stage[0], 0));
    if (packed) {
      HS_TRY(hs_run_bitmap_to_ids(ctx, ctx->bits[par].p, words, (const int64_t*)ctx->ptrs.p + u0, u1 - u0,
                                  (int32_t*)ctx->cols32.p, nullptr, st));
      HS_CUDA(ctx, cudaEventRecord(ctx->ev_bits[par], st));
      ctx->bits_used[par] = true;
    }
    if (perm)
      HS_TRY(hs_run_perm_csr(ctx, (const int64_t*)ctx->ptrs.p, n_units, n_vals, (const float*)ctx->vals32.p,
                             (const int32_t*)ctx->cols32.p, n_perm, index_base, dev_out));
    else
      HS_TRY(hs_run_csr(ctx, (const int64_t*)ctx->ptrs.p + u0, (const int32_t*)ctx->cols32.p,
                        (const float*)ctx->vals32.p, u1 - u0, dev_out + u0));
    u0 = u1;
  }
  tm.stop();
  if (prof)
    fprintf(stderr, "[upload] %lld values in %.1f ms on the calling thread: ring wait %.1f, narrow %.1f, pack %.1f ms (%d threads)\n",
            (long long)n_vals, 1e3 * (now() - t_begin), 1e3 * t_wait, 1e3 * t_narrow, 1e3 * t_pack, host_threads());
  return HS_OK;
}

// Split the units into batches that fit into device memory and run them one after the other.
// Each batch b covers units [ub, ue) and writes `out_per_unit` doubles per unit: contiguous for CSR rows;
// for the permutation form a batch-local (ue-ub) x n_perm column-major block that is re-laid out on the host.
static int staged_rows(hs_ctx* ctx, const std::vector<int64_t>& hp, const int32_t* ids, const void* vals, bool vals_f32,
                       int64_t ipv, bool perm, int n_perm, int index_base, double* host_out, int timer_slot) {
  const int64_t n_units = (int64_t)hp.size() - 1;
  const int64_t per_unit = perm ? n_perm : 1;
  HS_TRY(hs_ensure(ctx, ctx->out, sizeof(double) * (size_t)hs_max<int64_t>(n_units * per_unit, 1)));
  size_t free_b = 0, total_b = 0;
  HS_CUDA(ctx, cudaMemGetInfo(&free_b, &total_b));
  // bytes already held by the grow-only upload buffers count as available
  free_b += ctx->cols32.cap + ctx->vals32.cap + ctx->perm_cols.cap + ctx->perm_vals.cap;
  const double budget = (double)env_size("HS_UPLOAD_BUDGET", (size_t)(0.80 * (double)free_b));
  // uploaded ids + fp32 values (+ the slab-bucketed copy of both for the permutation form)
  const double bytes_per_value = 4.0 * (double)ipv + 4.0 + (perm ? 8.0 * (double)ipv : 0.0);
  std::vector<double> tmp;
  cudaStream_t st = ctx->stream;
  int64_t ub = 0;
  while (ub < n_units) {
    int64_t ue = ub + 1;
    while (ue < n_units && (double)(hp[ue + 1] - hp[ub]) * bytes_per_value <= budget) ++ue;
    double* dev_out = (double*)ctx->out.p + ub * per_unit;
    HS_TRY(upload_and_run(ctx, hp.data() + ub, ue - ub, ids, vals, vals_f32, ipv, perm, n_perm, index_base,
                          dev_out, timer_slot));
    if (!perm || (ub == 0 && ue == n_units)) {
      HS_CUDA(ctx, cudaMemcpyAsync(host_out + ub * per_unit, dev_out, sizeof(double) * (size_t)((ue - ub) * per_unit),
                                   cudaMemcpyDeviceToHost, st));
      HS_CUDA(ctx, cudaStreamSynchronize(st));
    } else {
      tmp.resize((size_t)((ue - ub) * per_unit));
      HS_CUDA(ctx, cudaMemcpyAsync(tmp.data(), dev_out, sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost, st));
      HS_CUDA(ctx, cudaStreamSynchronize(st));
      for (int r = 0; r < n_perm; ++r)
        for (int64_t s = 0; s < ue - ub; ++s)
          host_out[(ub + s) + (int64_t)r * n_units] = tmp[(size_t)s + (size_t)r * (size_t)(ue - ub)];
    }
    ub = ue;
  }
  return HS_OK;
}

static int dkl_csr_any(hs_ctx* ctx, const void* p, int p_is_64, const int32_t* j, const void* x, bool x_f32,
                       int64_t n_genes, double* dkl_out) {
  HS_TRY(need_density(ctx));
  HS_TRY(set_device(ctx));
  if (n_genes < 0 || !p || !dkl_out) return hs_fail(ctx, HS_ERR_INVALID, "hs_dkl_csr: bad arguments");
  if (n_genes == 0) return HS_OK;
  std::vector<int64_t> hp((size_t)n_genes + 1);
  for (int64_t i = 0; i <= n_genes; ++i)
    hp[i] = p_is_64 ? ((const int64_t*)p)[i] : (int64_t)((const int32_t*)p)[i];
  for (int64_t i = 0; i < n_genes; ++i)
    if (hp[i + 1] < hp[i]) return hs_fail(ctx, HS_ERR_INVALID, "hs_dkl_csr: row pointers must be non-decreasing");
  if (hp[n_genes] > hp[0] && (!j || !x)) return hs_fail(ctx, HS_ERR_INVALID, "hs_dkl_csr: j/x are NULL");
  // j and x are the caller's whole slots; rows start at element hp[0]
  return staged_rows(ctx, hp, j, x, x_f32, 1, false, 1, 0, dkl_out, HS_T_CSR);
}

int hs_dkl_csr(hs_ctx* ctx, const void* p, int p_is_64, const int32_t* j, const double* x, int64_t n_genes,
               double* dkl_out) {
  try {
  return dkl_csr_any(ctx, p, p_is_64, j, x, false, n_genes, dkl_out);
  } HS_ABI_CATCH(ctx)
}

// Observed D_KL of every gene and the seeded randomization batch of the selected ones in ONE call: the
// batch (values of ~100 rows, a few MB) is uploaded and launched first and computes while the CSR matrix
// is still crossing PCIe, instead of after it (hs_dkl_csr followed by hs_dkl_perm_csr_seeded).
int hs_dkl_csr_randomized(hs_ctx* ctx, const void* p, int p_is_64, const int32_t* j, const double* x, int64_t n_genes,
                          const int64_t* sel_rows, int64_t n_sel, int n_perm, uint64_t seed, uint64_t stream_base,
                          double* dkl_out, double* rand_out) {
  try {
  HS_TRY(need_density(ctx));
  HS_TRY(set_device(ctx));
  if (n_genes < 0 || n_sel < 0 || n_perm <= 0 || !p || !dkl_out || (n_sel > 0 && (!sel_rows || !rand_out)))
    return hs_fail(ctx, HS_ERR_INVALID, "hs_dkl_csr_randomized: bad arguments");
  auto ptr_at = [&](int64_t i) { return p_is_64 ? ((const int64_t*)p)[i] : (int64_t)((const int32_t*)p)[i]; };
  std::vector<int64_t> gp((size_t)n_sel + 1, 0);
  for (int64_t s = 0; s < n_sel; ++s) {
    const int64_t r = sel_rows[s];
    if (r < 0 || r >= n_genes) return hs_fail(ctx, HS_ERR_INVALID, "hs_dkl_csr_randomized: selected row out of range");
    const int64_t nnz = ptr_at(r + 1) - ptr_at(r);
    if (nnz < 0 || nnz > ctx->N) return hs_fail(ctx, HS_ERR_INVALID, "hs_dkl_csr_randomized: bad row length");
    gp[(size_t)s + 1] = gp[(size_t)s] + nnz;
  }
  const int64_t nv = gp[(size_t)n_sel];
  if (nv > 0 && !x) return hs_fail(ctx, HS_ERR_INVALID, "hs_dkl_csr_randomized: x is NULL");
  size_t free_b = 0, total_b = 0;
  HS_CUDA(ctx, cudaMemGetInfo(&free_b, &total_b));
  free_b += ctx->perm_cols.cap;
  const bool fits = (double)nv * (8.0 * n_perm + 16.0) < 0.25 * (double)free_b;   // leave the CSR upload its room
  if (n_sel > 0 && fits) {
    cudaStream_t st = ctx->stream;
    std::unique_ptr<float[]> v32(new float[(size_t)hs_max<int64_t>(nv, 1)]);
    {
      float* base = v32.get();
      auto rows = [&, base](int t, int nt) {
        for (int64_t s = t; s < n_sel; s += nt) {
          const double* src = x + ptr_at(sel_rows[s]);
          float* dst = base + gp[(size_t)s];
          const int64_t n = gp[(size_t)s + 1] - gp[(size_t)s];
          for (int64_t i = 0; i < n; ++i) dst[i] = (float)src[i];
        }
      };
      if (nv < ((int64_t)1 << 18) || host_threads() == 1) {
        rows(0, 1);
      } else {
        if (!ctx->pool) ctx->pool = new HostPool(host_threads());
        const int nt = ctx->pool->size();
        ctx->pool->run([&](int t) { rows(t, nt); });
      }
    }
    HS_TRY(hs_ensure(ctx, ctx->perm_gp, sizeof(int64_t) * (size_t)(n_sel + 1)));
    HS_TRY(hs_ensure(ctx, ctx->perm_vals, sizeof(float) * (size_t)hs_max<int64_t>(nv, 1)));
    HS_TRY(hs_ensure(ctx, ctx->out_rand, sizeof(double) * (size_t)(n_sel * n_perm)));
    HS_CUDA(ctx, cudaMemcpyAsync(ctx->perm_gp.p, gp.data(), sizeof(int64_t) * (size_t)(n_sel + 1), cudaMemcpyHostToDevice, st));
    HS_CUDA(ctx, cudaMemcpyAsync(ctx->perm_vals.p, v32.get(), sizeof(float) * (size_t)nv, cudaMemcpyHostToDevice, st));
    HS_CUDA(ctx, cudaStreamSynchronize(st));     // pageable sources; nothing else is queued yet
    HS_TRY(hs_run_perm_seeded(ctx, (const int64_t*)ctx->perm_gp.p, n_sel, nv, (const float*)ctx->perm_vals.p, n_perm, seed,
                              stream_base, (double*)ctx->out_rand.p));
    // the batch reads none of the CSR upload buffers: the uploads need not wait for it
    struct FenceGuard {   // cleared on every exit path, exceptions included
      hs_ctx* c;
      explicit FenceGuard(hs_ctx* cc) : c(cc) { c->skip_copy_fence = true; }
      ~FenceGuard() { c->skip_copy_fence = false; }
    };
    {
      FenceGuard guard(ctx);
      HS_TRY(dkl_csr_any(ctx, p, p_is_64, j, x, false, n_genes, dkl_out));
    }
    HS_CUDA(ctx, cudaMemcpyAsync(rand_out, ctx->out_rand.p, sizeof(double) * (size_t)(n_sel * n_perm), cudaMemcpyDeviceToHost, st));
    HS_CUDA(ctx, cudaStreamSynchronize(st));
    return HS_OK;
  }
  // randomization batch too large to sit beside the upload: one after the other
  HS_TRY(dkl_csr_any(ctx, p, p_is_64, j, x, false, n_genes, dkl_out));
  if (n_sel == 0) return HS_OK;
  std::vector<double> vsel((size_t)hs_max<int64_t>(nv, 1));
  for (int64_t s = 0; s < n_sel; ++s)
    memcpy(vsel.data() + gp[(size_t)s], x + ptr_at(sel_rows[s]), sizeof(double) * (size_t)(gp[(size_t)s + 1] - gp[(size_t)s]));
  return hs_dkl_perm_csr_seeded(ctx, gp.data(), n_sel, vsel.data(), n_perm, seed, stream_base, rand_out);
  } HS_ABI_CATCH(ctx)
}

int hs_dkl_csr_f32(hs_ctx* ctx, const void* p, int p_is_64, const int32_t* j, const float* x, int64_t n_genes,
                   double* dkl_out) {
  try {
  return dkl_csr_any(ctx, p, p_is_64, j, x, true, n_genes, dkl_out);
  } HS_ABI_CATCH(ctx)
}

static int dkl_perm_csr_any(hs_ctx* ctx, const int64_t* gene_ptr, int64_t n_sel, const void* val, bool val_f32,
                            const int32_t* ind, int n_perm, int index_base, double* out) {
  HS_TRY(need_density(ctx));
  HS_TRY(set_device(ctx));
  if (n_sel < 0 || n_perm <= 0 || !gene_ptr || !out)
    return hs_fail(ctx, HS_ERR_INVALID, "hs_dkl_perm_csr: bad arguments");
  if (n_sel == 0) return HS_OK;
  std::vector<int64_t> hp(gene_ptr, gene_ptr + n_sel + 1);
  if (hp[0] != 0) return hs_fail(ctx, HS_ERR_INVALID, "hs_dkl_perm_csr: gene_ptr[0] must be 0");
  for (int64_t i = 0; i < n_sel; ++i)
    if (hp[i + 1] < hp[i]) return hs_fail(ctx, HS_ERR_INVALID, "hs_dkl_perm_csr: gene_ptr must be non-decreasing");
  if (hp[n_sel] > 0 && (!val || !ind)) return hs_fail(ctx, HS_ERR_INVALID, "hs_dkl_perm_csr: val/ind are NULL");
  return staged_rows(ctx, hp, ind, val, val_f32, n_perm, true, n_perm, index_base, out, HS_T_PERM);
}

int hs_dkl_perm_csr(hs_ctx* ctx, const int64_t* gene_ptr, int64_t n_sel, const double* val, const int32_t* ind,
                    int n_perm, int index_base, double* out) {
  try {
  return dkl_perm_csr_any(ctx, gene_ptr, n_sel, val, false, ind, n_perm, index_base, out);
  } HS_ABI_CATCH(ctx)
}

// Asynchronous pair of hs_dkl_perm_csr: _begin starts the call on a worker thread and returns at once; _wait blocks
// until it is done and delivers the result.  val and ind must stay valid until _wait returns (gene_ptr is copied).
int hs_dkl_perm_csr_begin(hs_ctx* ctx, const int64_t* gene_ptr, int64_t n_sel, const double* val, const int32_t* ind,
                          int n_perm, int index_base) {
  try {
  HS_TRY(need_density(ctx));
  HS_TRY(set_device(ctx));
  if (n_sel < 0 || n_perm <= 0 || !gene_ptr)
    return hs_fail(ctx, HS_ERR_INVALID, "hs_dkl_perm_csr_begin: bad arguments");
  if (!ctx->async) ctx->async = new HsAsync();
  HsAsync* a = ctx->async;
  if (a->th.joinable()) a->th.join();        // a finished call nobody waited for
  a->gene_ptr.assign(gene_ptr, gene_ptr + n_sel + 1);
  a->result.assign((size_t)(n_sel * (int64_t)n_perm), 0.0);
  a->rc = HS_OK;
  a->pending = true;
  a->th = std::thread([ctx, a, n_sel, val, ind, n_perm, index_base] {
    a->worker = std::this_thread::get_id();
    try {
      a->rc = dkl_perm_csr_any(ctx, a->gene_ptr.data(), n_sel, val, false, ind, n_perm, index_base, a->result.data());
    } catch (const std::bad_alloc&) {
      a->rc = hs_fail(ctx, HS_ERR_NOMEM, "out of host memory");
    } catch (...) {
      a->rc = hs_fail(ctx, HS_ERR_INVALID, "unexpected C++ exception in the asynchronous call");
    }
  });
  return HS_OK;
  } HS_ABI_CATCH(ctx)
}

int hs_dkl_perm_csr_wait(hs_ctx* ctx, double* out) {
  try {
  if (!ctx) return hs_fail(nullptr, HS_ERR_INVALID, "null context");
  HsAsync* a = ctx->async;
  if (!a || !a->pending) return hs_fail(ctx, HS_ERR_STATE, "hs_dkl_perm_csr_wait: no asynchronous call in flight");
  if (a->th.joinable()) a->th.join();
  a->pending = false;
  if (a->rc != HS_OK) return a->rc;          // the worker left its message in the context
  if (!out && !a->result.empty()) return hs_fail(ctx, HS_ERR_INVALID, "hs_dkl_perm_csr_wait: out is NULL");
  if (!a->result.empty()) memcpy(out, a->result.data(), sizeof(double) * a->result.size());
  return HS_OK;
  } HS_ABI_CATCH(ctx)
}

int hs_dkl_perm_csr_f32(hs_ctx* ctx, const int64_t* gene_ptr, int64_t n_sel, const float* val, const int32_t* ind,
                        int n_perm, int index_base, double* out) {
  try {
  return dkl_perm_csr_any(ctx, gene_ptr, n_sel, val, true, ind, n_perm, index_base, out);
  } HS_ABI_CATCH(ctx)
}

int hs_dkl_dense(hs_ctx* ctx, const double* expr, int64_t n_genes, int64_t ld, double* dkl_out) {
  try {
  HS_TRY(need_density(ctx));
  HS_TRY(set_device(ctx));
  if (n_genes < 0 || ld < n_genes || !dkl_out || (n_genes > 0 && !expr))
    return hs_fail(ctx, HS_ERR_INVALID, "hs_dkl_dense: bad arguments");
  if (n_genes == 0) return HS_OK;
  const int64_t N = ctx->N;
  // upload in column blocks through the fp64 staging buffers, converting to fp32 on the device
  const int64_t cols_per_chunk = hs_max<int64_t>(1, ((int64_t)16 << 20) / n_genes);
  const size_t stage_elems = (size_t)hs_min<int64_t>(cols_per_chunk, N) * n_genes;
  for (int b = 0; b < 2; ++b) HS_TRY(hs_ensure(ctx, ctx->stage_x[b], sizeof(double) * stage_elems));
  HS_TRY(hs_ensure(ctx, ctx->dense32, sizeof(float) * (size_t)n_genes * N));
  HS_TRY(hs_ensure(ctx, ctx->out, sizeof(double) * (size_t)n_genes));
  cudaStream_t st = ctx->stream, cs = ctx->copy_stream;
  HS_CUDA(ctx, cudaEventRecord(ctx->ev_done[0], st));
  HS_CUDA(ctx, cudaStreamWaitEvent(cs, ctx->ev_done[0], 0));
  int k = 0;
  for (int64_t c0 = 0; c0 < N; c0 += cols_per_chunk, ++k) {
    const int b = k & 1;
    const int64_t nc = hs_min<int64_t>(cols_per_chunk, N - c0);
    if (k >= 2) HS_CUDA(ctx, cudaStreamWaitEvent(cs, ctx->ev_done[b], 0));
    HS_CUDA(ctx, cudaMemcpy2DAsync(ctx->stage_x[b].p, sizeof(double) * n_genes, expr + c0 * ld, sizeof(double) * ld,
                                   sizeof(double) * n_genes, (size_t)nc, cudaMemcpyHostToDevice, cs));
    HS_CUDA(ctx, cudaEventRecord(ctx->ev_stage[b], cs));
    HS_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_stage[b], 0));
    HS_TRY(hs_run_f64_to_f32(ctx, (const double*)ctx->stage_x[b].p, (float*)ctx->dense32.p + c0 * n_genes, nc * n_genes, st));
    HS_CUDA(ctx, cudaEventRecord(ctx->ev_done[b], st));
  }
  ScopedTimer tm(ctx, HS_T_DENSE);
  HS_TRY(hs_run_dense_f32(ctx, (const float*)ctx->dense32.p, n_genes, n_genes, (double*)ctx->out.p));
  tm.stop();
  HS_CUDA(ctx, cudaMemcpyAsync(dkl_out, ctx->out.p, sizeof(double) * (size_t)n_genes, cudaMemcpyDeviceToHost, st));
  HS_CUDA(ctx, cudaStreamSynchronize(st));
  return HS_OK;
  } HS_ABI_CATCH(ctx)
}

int hs_dkl_perm_dense(hs_ctx* ctx, const double* v, const int32_t* perm, int n_perm, int index_base, double* out) {
  try {
  HS_TRY(need_density(ctx));
  HS_TRY(set_device(ctx));
  if (!v || !perm || !out || n_perm <= 0) return hs_fail(ctx, HS_ERR_INVALID, "hs_dkl_perm_dense: bad arguments");
  const int64_t N = ctx->N;
  HS_TRY(hs_ensure(ctx, ctx->stage_x[0], sizeof(double) * (size_t)N));
  HS_TRY(hs_ensure(ctx, ctx->vals32, sizeof(float) * (size_t)N));
  HS_TRY(hs_ensure(ctx, ctx->stage_j[0], sizeof(int32_t) * (size_t)N * n_perm));
  HS_TRY(hs_ensure(ctx, ctx->out, sizeof(double) * (size_t)n_perm));
  cudaStream_t st = ctx->stream;
  HS_CUDA(ctx, cudaMemcpyAsync(ctx->stage_x[0].p, v, sizeof(double) * (size_t)N, cudaMemcpyHostToDevice, st));
  HS_CUDA(ctx, cudaMemcpyAsync(ctx->stage_j[0].p, perm, sizeof(int32_t) * (size_t)N * n_perm,
                               cudaMemcpyHostToDevice, st));
  HS_TRY(hs_run_f64_to_f32(ctx, (const double*)ctx->stage_x[0].p, (float*)ctx->vals32.p, N, st));
  ScopedTimer tm(ctx, HS_T_PERM);
  HS_TRY(hs_run_perm_dense(ctx, (const float*)ctx->vals32.p, (const int32_t*)ctx->stage_j[0].p, n_perm,
                           index_base, (double*)ctx->out.p));
  tm.stop();
  HS_CUDA(ctx, cudaMemcpyAsync(out, ctx->out.p, sizeof(double) * (size_t)n_perm, cudaMemcpyDeviceToHost, st));
  HS_CUDA(ctx, cudaStreamSynchronize(st));
  return HS_OK;
  } HS_ABI_CATCH(ctx)
}

// ------------------------------------------------------------------------------------------------
// device-resident variants
// ------------------------------------------------------------------------------------------------
int hs_dkl_csr_dev(hs_ctx* ctx, const int64_t* p_dev, const int32_t* j_dev, const float* x_dev, int64_t n_genes,
                   double* dkl_dev_out) {
  try {
  HS_TRY(need_density(ctx));
  HS_TRY(set_device(ctx));
  if (n_genes < 0 || !p_dev || !dkl_dev_out) return hs_fail(ctx, HS_ERR_INVALID, "hs_dkl_csr_dev: bad arguments");
  ScopedTimer tm(ctx, HS_T_CSR);
  HS_TRY(hs_run_csr(ctx, p_dev, j_dev, x_dev, n_genes, dkl_dev_out));
  tm.stop();
  return HS_OK;
  } HS_ABI_CATCH(ctx)
}

int hs_dkl_perm_csr_seeded_dev(hs_ctx* ctx, const int64_t* gene_ptr_dev, int64_t n_sel, const float* val_dev,
                               int n_perm, uint64_t seed, uint64_t stream_base, double* out_dev) {
  try {
  HS_TRY(need_density(ctx));
  HS_TRY(set_device(ctx));
  if (n_sel < 0 || n_perm <= 0 || !gene_ptr_dev || !out_dev)
    return hs_fail(ctx, HS_ERR_INVALID, "hs_dkl_perm_csr_seeded_dev: bad arguments");
  if (n_sel == 0) return HS_OK;
  int64_t n_vals = 0;   // sizes the (cell, value) pair buffer
  HS_CUDA(ctx, cudaMemcpyAsync(&n_vals, gene_ptr_dev + n_sel, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
  HS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (n_vals < 0) return hs_fail(ctx, HS_ERR_INVALID, "hs_dkl_perm_csr_seeded_dev: gene_ptr is not a prefix sum");
  if (n_vals > 0 && !val_dev) return hs_fail(ctx, HS_ERR_INVALID, "hs_dkl_perm_csr_seeded_dev: val is NULL");
  ScopedTimer tm(ctx, HS_T_PERM);
  HS_TRY(hs_run_perm_seeded(ctx, gene_ptr_dev, n_sel, n_vals, val_dev, n_perm, seed, stream_base, out_dev));
  tm.stop();
  return HS_OK;
  } HS_ABI_CATCH(ctx)
}

int hs_dkl_perm_csr_seeded_units_dev(hs_ctx* ctx, const int64_t* gene_ptr_dev, int64_t n_units, const float* val_dev,
                                     int n_perm, uint64_t seed, const uint64_t* unit_stream_dev, double* out_dev) {
  try {
  HS_TRY(need_density(ctx));
  HS_TRY(set_device(ctx));
  if (n_units < 0 || n_perm <= 0 || !gene_ptr_dev || !out_dev || (n_units > 0 && !unit_stream_dev))
    return hs_fail(ctx, HS_ERR_INVALID, "hs_dkl_perm_csr_seeded_units_dev: bad arguments");
  if (n_units == 0) return HS_OK;
  int64_t n_vals = 0;
  HS_CUDA(ctx, cudaMemcpyAsync(&n_vals, gene_ptr_dev + n_units, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
  HS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (n_vals < 0) return hs_fail(ctx, HS_ERR_INVALID, "hs_dkl_perm_csr_seeded_units_dev: gene_ptr is not a prefix sum");
  if (n_vals > 0 && !val_dev) return hs_fail(ctx, HS_ERR_INVALID, "hs_dkl_perm_csr_seeded_units_dev: val is NULL");
  ScopedTimer tm(ctx, HS_T_PERM);
  HS_TRY(hs_run_perm_seeded(ctx, gene_ptr_dev, n_units, n_vals, val_dev, n_perm, seed, 0, out_dev,
                            (const unsigned long long*)unit_stream_dev));
  tm.stop();
  return HS_OK;
  } HS_ABI_CATCH(ctx)
}

// ---- density stage with the cells sharded over ranks ----
static int shard_range(hs_ctx* ctx, int64_t n_cells, int rank, int world, int64_t* lo, int64_t* hi) {
  if (world < 1 || rank < 0 || rank >= world) return hs_fail(ctx, HS_ERR_INVALID, "bad rank / world");
  const int64_t per = hs_density_cells_per_rank(n_cells, world);
  *lo = hs_min<int64_t>(n_cells, (int64_t)rank * per);
  *hi = hs_min<int64_t>(n_cells, *lo + per);
  return HS_OK;
}

int hs_density_shard_dist_dev(hs_ctx* ctx, const double* x_dev, int64_t n_cells, int n_dims, const double* grid_dev,
                              int n_grid, int rank, int world) {
  try {
  HS_TRY(check_density_args(ctx, x_dev, n_cells, n_dims, grid_dev, n_grid));
  HS_TRY(set_device(ctx));
  int64_t lo = 0, hi = 0;
  HS_TRY(shard_range(ctx, n_cells, rank, world, &lo, &hi));
  ctx->N = 0;
  return hs_density_step_dist(ctx, x_dev, n_cells, n_dims, grid_dev, n_grid, lo, hi, world);
  } HS_ABI_CATCH(ctx)
}

int hs_density_shard_buffers(hs_ctx* ctx, int64_t n_cells, int n_grid, int world, void** rowmin_dev,
                             int64_t* rowmin_bytes_per_rank, void** d_dev, int64_t* d_bytes_per_rank, void** qpart_dev,
                             int64_t* qpart_bytes_per_rank) {
  try {
  if (!ctx) return hs_fail(nullptr, HS_ERR_INVALID, "null context");
  if (n_cells <= 0 || n_grid <= 0 || world < 1 || !ctx->rowmin.p || !ctx->D32.p || !ctx->qpart.p)
    return hs_fail(ctx, HS_ERR_STATE, "hs_density_shard_buffers: call hs_density_shard_dist_dev first");
  const int64_t per = hs_density_cells_per_rank(n_cells, world);
  const int ldD = hs_ld_for_grid(n_grid);
  if (rowmin_dev) *rowmin_dev = ctx->rowmin.p;
  if (d_dev) *d_dev = ctx->D32.p;
  if (qpart_dev) *qpart_dev = ctx->qpart.p;
  if (rowmin_bytes_per_rank) *rowmin_bytes_per_rank = (int64_t)sizeof(double) * per;
  if (d_bytes_per_rank) *d_bytes_per_rank = (int64_t)sizeof(float) * per * ldD;
  if (qpart_bytes_per_rank) *qpart_bytes_per_rank = (int64_t)sizeof(double) * (per / 128) * n_grid;
  return HS_OK;
  } HS_ABI_CATCH(ctx)
}

int hs_density_shard_dens_dev(hs_ctx* ctx, int64_t n_cells, int n_grid, const double* w_q_dev, int rank, int world) {
  try {
  if (!ctx) return hs_fail(nullptr, HS_ERR_INVALID, "null context");
  HS_TRY(set_device(ctx));
  int64_t lo = 0, hi = 0;
  HS_TRY(shard_range(ctx, n_cells, rank, world, &lo, &hi));
  return hs_density_step_dens(ctx, n_cells, n_grid, w_q_dev, lo, hi, false);
  } HS_ABI_CATCH(ctx)
}

int hs_density_shard_finish(hs_ctx* ctx, int64_t n_cells, int n_grid, double* bandwidth_out, double* q_out) {
  try {
  if (!ctx) return hs_fail(nullptr, HS_ERR_INVALID, "null context");
  HS_TRY(set_device(ctx));
  HS_TRY(hs_density_step_q(ctx, n_cells, n_grid));
  cudaStream_t st = ctx->stream;
  if (q_out) HS_CUDA(ctx, cudaMemcpyAsync(q_out, ctx->Q.p, sizeof(double) * n_grid, cudaMemcpyDeviceToHost, st));
  HS_CUDA(ctx, cudaStreamSynchronize(st));
  if (bandwidth_out) *bandwidth_out = ctx->bw;
  if (!(ctx->bw > 0.0)) {
    ctx->N = 0;
    return hs_fail(ctx, HS_ERR_INVALID, "hs_density_shard_finish: bandwidth %g is not positive", ctx->bw);
  }
  return HS_OK;
  } HS_ABI_CATCH(ctx)
}

int hs_dkl_perm_csr_seeded(hs_ctx* ctx, const int64_t* gene_ptr, int64_t n_sel, const double* val, int n_perm,
                           uint64_t seed, uint64_t stream_base, double* out) {
  try {
  HS_TRY(need_density(ctx));
  HS_TRY(set_device(ctx));
  if (n_sel < 0 || n_perm <= 0 || !gene_ptr || !out)
    return hs_fail(ctx, HS_ERR_INVALID, "hs_dkl_perm_csr_seeded: bad arguments");
  if (n_sel == 0) return HS_OK;
  if (gene_ptr[0] != 0) return hs_fail(ctx, HS_ERR_INVALID, "hs_dkl_perm_csr_seeded: gene_ptr[0] must be 0");
  for (int64_t i = 0; i < n_sel; ++i) {
    if (gene_ptr[i + 1] < gene_ptr[i])
      return hs_fail(ctx, HS_ERR_INVALID, "hs_dkl_perm_csr_seeded: gene_ptr must be non-decreasing");
    if (gene_ptr[i + 1] - gene_ptr[i] > ctx->N)
      return hs_fail(ctx, HS_ERR_INVALID, "hs_dkl_perm_csr_seeded: cannot take a sample larger than the population");
  }
  const int64_t n_vals = gene_ptr[n_sel];
  if (n_vals > 0 && !val) return hs_fail(ctx, HS_ERR_INVALID, "hs_dkl_perm_csr_seeded: val is NULL");
  // Only the values travel (n_vals floats), so memory is bounded by the pair buffer: split the genes into
  // batches whose n_perm * nnz pairs fit.
  size_t free_b = 0, total_b = 0;
  HS_CUDA(ctx, cudaMemGetInfo(&free_b, &total_b));
  free_b += ctx->perm_cols.cap + ctx->vals32.cap;
  const double budget = (double)env_size("HS_UPLOAD_BUDGET", (size_t)(0.80 * (double)free_b));
  HS_TRY(hs_ensure(ctx, ctx->out, sizeof(double) * (size_t)(n_sel * n_perm)));
  cudaStream_t st = ctx->stream;
  std::vector<float> v32;
  std::vector<double> tmp;
  int64_t ub = 0;
  while (ub < n_sel) {
    int64_t ue = ub + 1;
    while (ue < n_sel && (double)(gene_ptr[ue + 1] - gene_ptr[ub]) * (8.0 * n_perm + 4.0) <= budget) ++ue;
    const int64_t v0 = gene_ptr[ub], nv = gene_ptr[ue] - v0, nu = ue - ub;
    std::vector<int64_t> rel((size_t)nu + 1);
    for (int64_t u = 0; u <= nu; ++u) rel[u] = gene_ptr[ub + u] - v0;
    v32.resize((size_t)hs_max<int64_t>(nv, 1));
    for (int64_t i = 0; i < nv; ++i) v32[i] = (float)val[v0 + i];
    HS_TRY(hs_ensure(ctx, ctx->ptrs, sizeof(int64_t) * (size_t)(nu + 1)));
    HS_TRY(hs_ensure(ctx, ctx->vals32, sizeof(float) * (size_t)hs_max<int64_t>(nv, 1)));
    HS_CUDA(ctx, cudaMemcpyAsync(ctx->ptrs.p, rel.data(), sizeof(int64_t) * (size_t)(nu + 1), cudaMemcpyHostToDevice, st));
    HS_CUDA(ctx, cudaMemcpyAsync(ctx->vals32.p, v32.data(), sizeof(float) * (size_t)nv, cudaMemcpyHostToDevice, st));
    double* dev_out = (double*)ctx->out.p + ub * n_perm;
    ScopedTimer tm(ctx, HS_T_PERM);
    HS_TRY(hs_run_perm_seeded(ctx, (const int64_t*)ctx->ptrs.p, nu, nv, (const float*)ctx->vals32.p, n_perm, seed,
                              stream_base + (uint64_t)ub * (uint64_t)n_perm, dev_out));
    tm.stop();
    if (ub == 0 && ue == n_sel) {
      HS_CUDA(ctx, cudaMemcpyAsync(out, dev_out, sizeof(double) * (size_t)(nu * n_perm), cudaMemcpyDeviceToHost, st));
      HS_CUDA(ctx, cudaStreamSynchronize(st));
    } else {   // batch-local nu x n_perm block -> its rows of the n_sel x n_perm result
      tmp.resize((size_t)(nu * n_perm));
      HS_CUDA(ctx, cudaMemcpyAsync(tmp.data(), dev_out, sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost, st));
      HS_CUDA(ctx, cudaStreamSynchronize(st));
      for (int r = 0; r < n_perm; ++r)
        for (int64_t s = 0; s < nu; ++s) out[(ub + s) + (int64_t)r * n_sel] = tmp[(size_t)s + (size_t)r * (size_t)nu];
    }
    ub = ue;
  }
  return HS_OK;
  } HS_ABI_CATCH(ctx)
}

int hs_perm_sample(uint64_t seed, uint64_t stream, int64_t n, int64_t size, int index_base, int32_t* ids_out) {
  try {
  if (n < 1 || n >= ((int64_t)1 << 31) || size < 0 || size > n || (size > 0 && !ids_out))
    return hs_fail(nullptr, HS_ERR_INVALID, "hs_perm_sample: need 1 <= n < 2^31 and 0 <= size <= n");
  const HsPrp prp = hs_prp_make(seed, stream, (uint32_t)n);
  for (int64_t i = 0; i < size; ++i) ids_out[i] = (int32_t)((int64_t)hs_prp_forward(prp, (uint32_t)i) + index_base);
  return HS_OK;
  } HS_ABI_CATCH(nullptr)
}

int hs_dkl_perm_csr_dev(hs_ctx* ctx, const int64_t* gene_ptr_dev, int64_t n_sel, const float* val_dev,
                        const int32_t* ind_dev, int n_perm, int index_base, double* out_dev) {
  try {
  HS_TRY(need_density(ctx));
  HS_TRY(set_device(ctx));
  if (n_sel < 0 || n_perm <= 0 || !gene_ptr_dev || !out_dev)
    return hs_fail(ctx, HS_ERR_INVALID, "hs_dkl_perm_csr_dev: bad arguments");
  if (n_sel == 0) return HS_OK;
  int64_t n_vals = 0;   // the slab bucketing sizes its buffers from the total number of stored values
  HS_CUDA(ctx, cudaMemcpyAsync(&n_vals, gene_ptr_dev + n_sel, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
  HS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (n_vals < 0) return hs_fail(ctx, HS_ERR_INVALID, "hs_dkl_perm_csr_dev: gene_ptr is not a prefix sum");
  ScopedTimer tm(ctx, HS_T_PERM);
  HS_TRY(hs_run_perm_csr(ctx, gene_ptr_dev, n_sel, n_vals, val_dev, ind_dev, n_perm, index_base, out_dev));
  tm.stop();
  return HS_OK;
  } HS_ABI_CATCH(ctx)
}

int hs_dkl_dense_dev(hs_ctx* ctx, const float* expr_dev, int64_t n_genes, int64_t ld, double* dkl_dev_out) {
  try {
  HS_TRY(need_density(ctx));
  HS_TRY(set_device(ctx));
  if (n_genes < 0 || ld < n_genes || !expr_dev || !dkl_dev_out)
    return hs_fail(ctx, HS_ERR_INVALID, "hs_dkl_dense_dev: bad arguments");
  ScopedTimer tm(ctx, HS_T_DENSE);
  HS_TRY(hs_run_dense_f32(ctx, expr_dev, n_genes, ld, dkl_dev_out));
  tm.stop();
  return HS_OK;
  } HS_ABI_CATCH(ctx)
}

// ---- prepare_clustering (R/haystack_continuous.R:604-666): gene x grid profiles, rows summing to 1 ----
// Not a hot path (a few hundred selected genes): plain synchronous uploads.
int hs_profile_csr(hs_ctx* ctx, const void* p, int p_is_64, const int32_t* j, const double* x, int64_t n_genes,
                   double* profile_out) {
  try {
  HS_TRY(need_density(ctx));
  HS_TRY(set_device(ctx));
  if (n_genes < 0 || !p || !profile_out) return hs_fail(ctx, HS_ERR_INVALID, "hs_profile_csr: bad arguments");
  if (n_genes == 0) return HS_OK;
  std::vector<int64_t> hp((size_t)n_genes + 1);
  for (int64_t i = 0; i <= n_genes; ++i) hp[i] = p_is_64 ? ((const int64_t*)p)[i] : (int64_t)((const int32_t*)p)[i];
  for (int64_t i = 0; i < n_genes; ++i)
    if (hp[i + 1] < hp[i]) return hs_fail(ctx, HS_ERR_INVALID, "hs_profile_csr: row pointers must be non-decreasing");
  const int64_t v0 = hp[0], nnz = hp[n_genes] - v0;
  if (nnz > 0 && (!j || !x)) return hs_fail(ctx, HS_ERR_INVALID, "hs_profile_csr: j/x are NULL");
  for (auto& v : hp) v -= v0;
  cudaStream_t st = ctx->stream;
  HS_TRY(hs_ensure(ctx, ctx->ptrs, sizeof(int64_t) * (size_t)(n_genes + 1)));
  HS_TRY(hs_ensure(ctx, ctx->cols32, sizeof(int32_t) * (size_t)hs_max<int64_t>(nnz, 1)));
  HS_TRY(hs_ensure(ctx, ctx->prof_x, sizeof(double) * (size_t)hs_max<int64_t>(nnz, 1)));
  HS_TRY(hs_ensure(ctx, ctx->prof_out, sizeof(double) * (size_t)n_genes * ctx->G));
  HS_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
  HS_CUDA(ctx, cudaMemcpyAsync(ctx->ptrs.p, hp.data(), sizeof(int64_t) * (size_t)(n_genes + 1), cudaMemcpyHostToDevice, st));
  if (nnz > 0) {
    HS_CUDA(ctx, cudaMemcpyAsync(ctx->cols32.p, j + v0, sizeof(int32_t) * (size_t)nnz, cudaMemcpyHostToDevice, st));
    HS_CUDA(ctx, cudaMemcpyAsync(ctx->prof_x.p, x + v0, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, st));
  }
  HS_TRY(hs_run_profile(ctx, false, (const int64_t*)ctx->ptrs.p, (const int32_t*)ctx->cols32.p, (const double*)ctx->prof_x.p, 0,
                        n_genes, (double*)ctx->prof_out.p));
  HS_CUDA(ctx, cudaMemcpyAsync(profile_out, ctx->prof_out.p, sizeof(double) * (size_t)n_genes * ctx->G, cudaMemcpyDeviceToHost, st));
  HS_CUDA(ctx, cudaStreamSynchronize(st));
  return HS_OK;
  } HS_ABI_CATCH(ctx)
}

int hs_profile_dense(hs_ctx* ctx, const double* expr, int64_t n_genes, int64_t ld, double* profile_out) {
  try {
  HS_TRY(need_density(ctx));
  HS_TRY(set_device(ctx));
  if (n_genes < 0 || ld < n_genes || !profile_out || (n_genes > 0 && !expr))
    return hs_fail(ctx, HS_ERR_INVALID, "hs_profile_dense: bad arguments");
  if (n_genes == 0) return HS_OK;
  const int64_t N = ctx->N;
  cudaStream_t st = ctx->stream;
  HS_TRY(hs_ensure(ctx, ctx->prof_x, sizeof(double) * (size_t)n_genes * N));
  HS_TRY(hs_ensure(ctx, ctx->prof_out, sizeof(double) * (size_t)n_genes * ctx->G));
  HS_CUDA(ctx, cudaMemcpy2DAsync(ctx->prof_x.p, sizeof(double) * n_genes, expr, sizeof(double) * ld, sizeof(double) * n_genes,
                                 (size_t)N, cudaMemcpyHostToDevice, st));
  HS_TRY(hs_run_profile(ctx, true, nullptr, nullptr, (const double*)ctx->prof_x.p, n_genes, n_genes, (double*)ctx->prof_out.p));
  HS_CUDA(ctx, cudaMemcpyAsync(profile_out, ctx->prof_out.p, sizeof(double) * (size_t)n_genes * ctx->G, cudaMemcpyDeviceToHost, st));
  HS_CUDA(ctx, cudaStreamSynchronize(st));
  return HS_OK;
  } HS_ABI_CATCH(ctx)
}

int hs_dkl_perm_dense_dev(hs_ctx* ctx, const float* v_dev, const int32_t* perm_dev, int n_perm, int index_base,
                          double* out_dev) {
  try {
  HS_TRY(need_density(ctx));
  HS_TRY(set_device(ctx));
  if (!v_dev || !perm_dev || !out_dev || n_perm <= 0)
    return hs_fail(ctx, HS_ERR_INVALID, "hs_dkl_perm_dense_dev: bad arguments");
  ScopedTimer tm(ctx, HS_T_PERM);
  HS_TRY(hs_run_perm_dense(ctx, v_dev, perm_dev, n_perm, index_base, out_dev));
  tm.stop();
  return HS_OK;
  } HS_ABI_CATCH(ctx)
}

int hs_dkl_perm_dense_seeded_dev(hs_ctx* ctx, const float* v_dev, int n_perm, uint64_t seed, uint64_t stream_base,
                                 double* out_dev) {
  try {
  HS_TRY(need_density(ctx));
  HS_TRY(set_device(ctx));
  if (!v_dev || !out_dev || n_perm <= 0)
    return hs_fail(ctx, HS_ERR_INVALID, "hs_dkl_perm_dense_seeded_dev: bad arguments");
  ScopedTimer tm(ctx, HS_T_PERM);
  HS_TRY(hs_run_perm_dense_seeded(ctx, v_dev, n_perm, seed, stream_base, out_dev));
  tm.stop();
  return HS_OK;
  } HS_ABI_CATCH(ctx)
}

int hs_dkl_perm_dense_seeded(hs_ctx* ctx, const double* v, int n_perm, uint64_t seed, uint64_t stream_base,
                             double* out) {
  try {
  HS_TRY(need_density(ctx));
  HS_TRY(set_device(ctx));
  if (!v || !out || n_perm <= 0) return hs_fail(ctx, HS_ERR_INVALID, "hs_dkl_perm_dense_seeded: bad arguments");
  const int64_t N = ctx->N;
  HS_TRY(hs_ensure(ctx, ctx->stage_x[0], sizeof(double) * (size_t)N));
  HS_TRY(hs_ensure(ctx, ctx->vals32, sizeof(float) * (size_t)N));
  HS_TRY(hs_ensure(ctx, ctx->out, sizeof(double) * (size_t)n_perm));
  cudaStream_t st = ctx->stream;
  HS_CUDA(ctx, cudaMemcpyAsync(ctx->stage_x[0].p, v, sizeof(double) * (size_t)N, cudaMemcpyHostToDevice, st));
  HS_TRY(hs_run_f64_to_f32(ctx, (const double*)ctx->stage_x[0].p, (float*)ctx->vals32.p, N, st));
  ScopedTimer tm(ctx, HS_T_PERM);
  HS_TRY(hs_run_perm_dense_seeded(ctx, (const float*)ctx->vals32.p, n_perm, seed, stream_base, (double*)ctx->out.p));
  tm.stop();
  HS_CUDA(ctx, cudaMemcpyAsync(out, ctx->out.p, sizeof(double) * (size_t)n_perm, cudaMemcpyDeviceToHost, st));
  HS_CUDA(ctx, cudaStreamSynchronize(st));
  return HS_OK;
  } HS_ABI_CATCH(ctx)
}

// The dense randomization of ALL selected genes in one call (R/haystack_continuous.R:250-265, the whole loop over
// genes.to.randomize): row sel_rows[s] of the column-major expression matrix, n_perm seeded randomizations each,
// streams stream_base + s * n_perm + r as in the sparse form.  No ids cross the link (the reference's loop would
// need an n_cells x n_perm index matrix per gene); the rows are gathered by the host threads, two staging buffers
// keep the upload of gene s + 1 under the scoring of gene s, and the caller waits once.  out: n_sel x n_perm
// column-major, like all.D_KL.randomized.
int hs_dkl_perm_dense_seeded_batch(hs_ctx* ctx, const double* expr, int64_t ld, const int64_t* sel_rows, int64_t n_sel,
                                   int n_perm, uint64_t seed, uint64_t stream_base, double* out) {
  try {
  HS_TRY(need_density(ctx));
  HS_TRY(set_device(ctx));
  if (!expr || !out || !sel_rows || n_sel < 0 || n_perm <= 0 || ld <= 0)
    return hs_fail(ctx, HS_ERR_INVALID, "hs_dkl_perm_dense_seeded_batch: bad arguments");
  for (int64_t s = 0; s < n_sel; ++s)
    if (sel_rows[s] < 0 || sel_rows[s] >= ld)
      return hs_fail(ctx, HS_ERR_INVALID, "hs_dkl_perm_dense_seeded_batch: selected row out of range");
  if (n_sel == 0) return HS_OK;
  const int64_t N = ctx->N;
  HS_TRY(ensure_ring(ctx, sizeof(float) * (size_t)N));
  HS_TRY(hs_ensure(ctx, ctx->stage_x[0], sizeof(float) * (size_t)N));
  HS_TRY(hs_ensure(ctx, ctx->stage_x[1], sizeof(float) * (size_t)N));
  HS_TRY(hs_ensure(ctx, ctx->out_rand, sizeof(double) * (size_t)(n_sel * n_perm)));
  cudaStream_t st = ctx->stream, cs = ctx->copy_stream;
  HS_CUDA(ctx, cudaEventRecord(ctx->ev_done[0], st));
  HS_CUDA(ctx, cudaStreamWaitEvent(cs, ctx->ev_done[0], 0));
  ScopedTimer tm(ctx, HS_T_PERM);
  int ring_k = 0;
  for (int64_t s = 0; s < n_sel; ++s) {
    const int b = (int)(s & 1);
    // gather the strided row as fp32 into a pinned buffer (host threads), then one contiguous copy
    const int k = ring_k;
    ring_k = (ring_k + 1) % hs_ctx::kRing;
    if (ctx->ring_busy[k]) {
      HS_CUDA(ctx, cudaEventSynchronize(ctx->ev_ring[k]));
      ctx->ring_busy[k] = false;
    }
    float* dst = (float*)ctx->ring[k];
    const double* src = expr + sel_rows[s];
    auto part = [=](int t, int nt) {
      const int64_t per = (N + nt - 1) / nt, a = hs_min<int64_t>(N, (int64_t)t * per), e = hs_min<int64_t>(N, a + per);
      for (int64_t c = a; c < e; ++c) dst[c] = (float)src[c * ld];
    };
    if (N < ((int64_t)1 << 16) || host_threads() == 1) {
      part(0, 1);
    } else {
      if (!ctx->pool) ctx->pool = new HostPool(host_threads());
      const int nt = ctx->pool->size();
      ctx->pool->run([&](int t) { part(t, nt); });
    }
    if (s >= 2) HS_CUDA(ctx, cudaStreamWaitEvent(cs, ctx->ev_done[b], 0));      // gene s - 2 has been scored
    HS_CUDA(ctx, cudaMemcpyAsync(ctx->stage_x[b].p, dst, sizeof(float) * (size_t)N, cudaMemcpyHostToDevice, cs));
    HS_CUDA(ctx, cudaEventRecord(ctx->ev_ring[k], cs));
    ctx->ring_busy[k] = true;
    HS_CUDA(ctx, cudaEventRecord(ctx->ev_stage[b], cs));
    HS_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_stage[b], 0));
    HS_TRY(hs_run_perm_dense_seeded(ctx, (const float*)ctx->stage_x[b].p, n_perm, seed,
                                    stream_base + (uint64_t)s * (uint64_t)n_perm,
                                    (double*)ctx->out_rand.p + s * (int64_t)n_perm));
    HS_CUDA(ctx, cudaEventRecord(ctx->ev_done[b], st));
  }
  tm.stop();
  std::vector<double> tmp((size_t)(n_sel * n_perm));
  HS_CUDA(ctx, cudaMemcpyAsync(tmp.data(), ctx->out_rand.p, sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost, st));
  HS_CUDA(ctx, cudaStreamSynchronize(st));
  for (int64_t s = 0; s < n_sel; ++s)
    for (int r = 0; r < n_perm; ++r) out[s + (int64_t)r * n_sel] = tmp[(size_t)(s * n_perm + r)];
  return HS_OK;
  } HS_ABI_CATCH(ctx)
}

int hs_row_stats_csr_dev(hs_ctx* ctx, const int64_t* p_dev, const float* x_dev, int64_t n_genes, int64_t n_cells,
                         double* mean_dev, double* sd_dev) {
  try {
  if (!ctx) return hs_fail(nullptr, HS_ERR_INVALID, "null context");
  HS_TRY(set_device(ctx));
  if (n_genes < 0 || n_cells < 1 || !p_dev || !mean_dev || !sd_dev)
    return hs_fail(ctx, HS_ERR_INVALID, "hs_row_stats_csr_dev: bad arguments");
  return hs_run_row_stats_f32(ctx, p_dev, x_dev, n_genes, n_cells, mean_dev, sd_dev);
  } HS_ABI_CATCH(ctx)
}

// ---- grid selection, "seeding" method (R/haystack_highD.R:87-129) ------------------------------------------------
int hs_grid_seed_begin(hs_ctx* ctx, const double* x, int64_t n_cells, int n_dims) {
  try {
  if (!ctx) return hs_fail(nullptr, HS_ERR_INVALID, "null context");
  HS_TRY(set_device(ctx));
  if (!x || n_cells <= 0 || n_dims <= 0 || n_dims > 4096)
    return hs_fail(ctx, HS_ERR_INVALID, "hs_grid_seed_begin: bad arguments");
  HS_TRY(hs_ensure(ctx, ctx->xin, sizeof(double) * (size_t)n_cells * (size_t)n_dims));
  HS_CUDA(ctx, cudaMemcpyAsync(ctx->xin.p, x, sizeof(double) * (size_t)n_cells * (size_t)n_dims, cudaMemcpyHostToDevice,
                               ctx->stream));
  HS_TRY(hs_run_seed_begin(ctx, n_cells, n_dims));
  HS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->seed_n = n_cells;
  ctx->seed_d = n_dims;
  return HS_OK;
  } HS_ABI_CATCH(ctx)
}

int hs_grid_seed_step(hs_ctx* ctx, int64_t index, double* probs_out, int64_t* nearest3_out) {
  try {
  if (!ctx) return hs_fail(nullptr, HS_ERR_INVALID, "null context");
  HS_TRY(set_device(ctx));
  if (ctx->seed_n <= 0) return hs_fail(ctx, HS_ERR_STATE, "hs_grid_seed_step: call hs_grid_seed_begin first");
  if (index < 0 || index >= ctx->seed_n || !nearest3_out)
    return hs_fail(ctx, HS_ERR_INVALID, "hs_grid_seed_step: bad arguments");
  const double* probs_dev = nullptr;
  const void* near_dev = nullptr;
  HS_TRY(hs_run_seed_step(ctx, (const double*)ctx->xin.p, ctx->seed_n, ctx->seed_d, index, &probs_dev, &near_dev));
  struct { double d; long long c; } h[3];
  HS_CUDA(ctx, cudaMemcpyAsync(h, near_dev, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  if (probs_out)
    HS_CUDA(ctx, cudaMemcpyAsync(probs_out, probs_dev, sizeof(double) * (size_t)ctx->seed_n, cudaMemcpyDeviceToHost, ctx->stream));
  HS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (int r = 0; r < 3; ++r) nearest3_out[r] = (r < ctx->seed_n) ? (int64_t)h[r].c : -1;
  return HS_OK;
  } HS_ABI_CATCH(ctx)
}

// coeffVar and the rows at the given ranks of order(coeffVar) from device-resident row statistics; see hs_spmm.cu
int hs_select_by_cv_dev(hs_ctx* ctx, const double* mean_dev, const double* sd_dev, int64_t n_genes, const int64_t* ranks,
                        int64_t n_ranks, double* cv_dev_out, int64_t* rows_out) {
  try {
  HS_TRY(set_device(ctx));
  if (!mean_dev || !sd_dev || n_genes <= 0 || n_ranks < 0 || (n_ranks > 0 && (!ranks || !rows_out)))
    return hs_fail(ctx, HS_ERR_INVALID, "hs_select_by_cv_dev: bad arguments");
  cudaStream_t st = ctx->stream;
  HS_TRY(hs_ensure(ctx, ctx->out, sizeof(int64_t) * (size_t)(2 * hs_max<int64_t>(n_ranks, 1) + 1)));
  int64_t* ranks_dev = (int64_t*)ctx->out.p;
  int64_t* rows_dev = ranks_dev + hs_max<int64_t>(n_ranks, 1);
  int* flag = (int*)(rows_dev + hs_max<int64_t>(n_ranks, 1));
  HS_CUDA(ctx, cudaMemsetAsync(flag, 0, sizeof(int), st));
  if (n_ranks > 0)
    HS_CUDA(ctx, cudaMemcpyAsync(ranks_dev, ranks, sizeof(int64_t) * (size_t)n_ranks, cudaMemcpyHostToDevice, st));
  HS_TRY(hs_run_order_by_cv(ctx, mean_dev, sd_dev, n_genes, ranks_dev, n_ranks, cv_dev_out, rows_dev, flag));
  int h_flag = 0;
  if (n_ranks > 0)
    HS_CUDA(ctx, cudaMemcpyAsync(rows_out, rows_dev, sizeof(int64_t) * (size_t)n_ranks, cudaMemcpyDeviceToHost, st));
  HS_CUDA(ctx, cudaMemcpyAsync(&h_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
  HS_CUDA(ctx, cudaStreamSynchronize(st));
  if (h_flag) return hs_fail(ctx, HS_ERR_INVALID, "hs_select_by_cv_dev: a rank lies outside [0, n_genes)");
  return HS_OK;
  } HS_ABI_CATCH(ctx)
}

int hs_row_stats_csr(hs_ctx* ctx, const void* p, int p_is_64, const double* x, int64_t n_genes, int64_t n_cells,
                     double* mean_out, double* sd_out) {
  try {
  if (!ctx) return hs_fail(nullptr, HS_ERR_INVALID, "null context");
  HS_TRY(set_device(ctx));
  if (n_genes < 0 || n_cells < 1 || !p || !mean_out || !sd_out)
    return hs_fail(ctx, HS_ERR_INVALID, "hs_row_stats_csr: bad arguments");
  if (n_genes == 0) return HS_OK;
  std::vector<int64_t> hp((size_t)n_genes + 1);
  for (int64_t i = 0; i <= n_genes; ++i) hp[i] = p_is_64 ? ((const int64_t*)p)[i] : (int64_t)((const int32_t*)p)[i];
  int64_t longest = 0;
  for (int64_t i = 0; i < n_genes; ++i) {
    if (hp[i + 1] < hp[i]) return hs_fail(ctx, HS_ERR_INVALID, "hs_row_stats_csr: row pointers must be non-decreasing");
    longest = hs_max(longest, hp[i + 1] - hp[i]);
  }
  if (hp[n_genes] > hp[0] && !x) return hs_fail(ctx, HS_ERR_INVALID, "hs_row_stats_csr: x is NULL");
  // The doubles travel as they are (the statistics feed an integer decision, the gene selection): pieces of
  // whole rows through two staging buffers, so the upload of one piece overlaps the passes over the other.
  const int64_t piece_vals = hs_max<int64_t>(longest, (int64_t)env_size("HS_STAGE_VALUES", (size_t)16 << 20));
  std::vector<int64_t> first;          // first row of every piece (+ n_genes at the end)
  for (int64_t u0 = 0; u0 < n_genes;) {
    int64_t u1 = u0 + 1;
    while (u1 < n_genes && hp[u1 + 1] - hp[u0] <= piece_vals) ++u1;
    first.push_back(u0);
    u0 = u1;
  }
  first.push_back(n_genes);
  const size_t n_pieces = first.size() - 1;
  // row pointers relative to each piece's first value; piece pc owns slots [first[pc] + pc, first[pc+1] + pc]
  std::vector<int64_t> pall((size_t)n_genes + n_pieces);
  for (size_t pc = 0; pc < n_pieces; ++pc)
    for (int64_t u = first[pc]; u <= first[pc + 1]; ++u) pall[(size_t)u + pc] = hp[u] - hp[first[pc]];
  for (int b = 0; b < 2; ++b) HS_TRY(hs_ensure(ctx, ctx->stage_x[b], sizeof(double) * (size_t)hs_max<int64_t>(piece_vals, 1)));
  HS_TRY(hs_ensure(ctx, ctx->ptrs, sizeof(int64_t) * pall.size()));
  HS_TRY(hs_ensure(ctx, ctx->out, sizeof(double) * (size_t)(2 * n_genes)));
  cudaStream_t st = ctx->stream, cs = ctx->copy_stream;
  double* mean_dev = (double*)ctx->out.p;
  double* sd_dev = mean_dev + n_genes;
  HS_CUDA(ctx, cudaMemcpyAsync(ctx->ptrs.p, pall.data(), sizeof(int64_t) * pall.size(), cudaMemcpyHostToDevice, st));
  HS_CUDA(ctx, cudaEventRecord(ctx->ev_done[0], st));
  HS_CUDA(ctx, cudaStreamWaitEvent(cs, ctx->ev_done[0], 0));   // earlier work on the staging buffers is done
  for (size_t pc = 0; pc < n_pieces; ++pc) {
    const int64_t r0 = first[pc], r1 = first[pc + 1];
    const int bsel = (int)(pc & 1);
    const int64_t nv = hp[r1] - hp[r0];
    if (pc >= 2) HS_CUDA(ctx, cudaStreamWaitEvent(cs, ctx->ev_done[bsel], 0));
    if (nv > 0)
      HS_CUDA(ctx, cudaMemcpyAsync(ctx->stage_x[bsel].p, x + hp[r0], sizeof(double) * (size_t)nv, cudaMemcpyHostToDevice, cs));
    HS_CUDA(ctx, cudaEventRecord(ctx->ev_stage[bsel], cs));
    HS_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_stage[bsel], 0));
    HS_TRY(hs_run_row_stats_f64(ctx, (const int64_t*)ctx->ptrs.p + r0 + (int64_t)pc, (const double*)ctx->stage_x[bsel].p,
                                r1 - r0, n_cells, mean_dev + r0, sd_dev + r0));
    HS_CUDA(ctx, cudaEventRecord(ctx->ev_done[bsel], st));
  }
  HS_CUDA(ctx, cudaMemcpyAsync(mean_out, mean_dev, sizeof(double) * (size_t)n_genes, cudaMemcpyDeviceToHost, st));
  HS_CUDA(ctx, cudaMemcpyAsync(sd_out, sd_dev, sizeof(double) * (size_t)n_genes, cudaMemcpyDeviceToHost, st));
  HS_CUDA(ctx, cudaStreamSynchronize(st));
  return HS_OK;
  } HS_ABI_CATCH(ctx)
}

int hs_csc_to_csr_dev(hs_ctx* ctx, const int64_t* cp_dev, const int32_t* ri_dev, const float* x_dev, int64_t n_rows,
                      int64_t n_cols, int64_t* p_out_dev, int32_t* j_out_dev, float* x_out_dev) {
  try {
  if (!ctx) return hs_fail(nullptr, HS_ERR_INVALID, "null context");
  HS_TRY(set_device(ctx));
  if (n_rows < 0 || n_cols < 0 || !cp_dev || !p_out_dev)
    return hs_fail(ctx, HS_ERR_INVALID, "hs_csc_to_csr_dev: bad arguments");
  int64_t nnz = 0;
  HS_CUDA(ctx, cudaMemcpyAsync(&nnz, cp_dev + n_cols, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
  HS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (nnz < 0 || (nnz > 0 && (!ri_dev || !x_dev || !j_out_dev || !x_out_dev)))
    return hs_fail(ctx, HS_ERR_INVALID, "hs_csc_to_csr_dev: bad arguments");
  return hs_run_csc_to_csr_f32(ctx, cp_dev, ri_dev, x_dev, n_rows, n_cols, nnz, p_out_dev, j_out_dev, x_out_dev);
  } HS_ABI_CATCH(ctx)
}

int hs_csc_to_csr(hs_ctx* ctx, const void* cp, int cp_is_64, const int32_t* ri, const double* x, int64_t n_rows,
                  int64_t n_cols, int64_t* p_out, int32_t* j_out, double* x_out) {
  try {
  if (!ctx) return hs_fail(nullptr, HS_ERR_INVALID, "null context");
  HS_TRY(set_device(ctx));
  if (n_rows < 0 || n_cols < 0 || !cp || !p_out) return hs_fail(ctx, HS_ERR_INVALID, "hs_csc_to_csr: bad arguments");
  std::vector<int64_t> hp((size_t)n_cols + 1);
  for (int64_t c = 0; c <= n_cols; ++c) hp[c] = cp_is_64 ? ((const int64_t*)cp)[c] : (int64_t)((const int32_t*)cp)[c];
  if (hp[0] != 0) return hs_fail(ctx, HS_ERR_INVALID, "hs_csc_to_csr: column pointers must start at 0");
  for (int64_t c = 0; c < n_cols; ++c)
    if (hp[c + 1] < hp[c]) return hs_fail(ctx, HS_ERR_INVALID, "hs_csc_to_csr: column pointers must be non-decreasing");
  const int64_t nnz = hp[n_cols];
  if (nnz > 0 && (!ri || !x || !j_out || !x_out)) return hs_fail(ctx, HS_ERR_INVALID, "hs_csc_to_csr: NULL array");
  const size_t nz = (size_t)hs_max<int64_t>(nnz, 1);
  HS_TRY(hs_ensure(ctx, ctx->conv_cp, sizeof(int64_t) * (size_t)(n_cols + 1)));
  HS_TRY(hs_ensure(ctx, ctx->conv_ri, sizeof(int32_t) * nz));
  HS_TRY(hs_ensure(ctx, ctx->conv_x, sizeof(double) * nz));
  HS_TRY(hs_ensure(ctx, ctx->conv_p, sizeof(int64_t) * (size_t)(n_rows + 1)));
  HS_TRY(hs_ensure(ctx, ctx->conv_j, sizeof(int32_t) * nz));
  HS_TRY(hs_ensure(ctx, ctx->conv_xo, sizeof(double) * nz));
  cudaStream_t st = ctx->stream;
  HS_CUDA(ctx, cudaMemcpyAsync(ctx->conv_cp.p, hp.data(), sizeof(int64_t) * (size_t)(n_cols + 1), cudaMemcpyHostToDevice, st));
  if (nnz > 0) {
    HS_CUDA(ctx, cudaMemcpyAsync(ctx->conv_ri.p, ri, sizeof(int32_t) * (size_t)nnz, cudaMemcpyHostToDevice, st));
    HS_CUDA(ctx, cudaMemcpyAsync(ctx->conv_x.p, x, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, st));
  }
  HS_TRY(hs_run_csc_to_csr_f64(ctx, (const int64_t*)ctx->conv_cp.p, (const int32_t*)ctx->conv_ri.p, (const double*)ctx->conv_x.p,
                               n_rows, n_cols, nnz, (int64_t*)ctx->conv_p.p, (int32_t*)ctx->conv_j.p, (double*)ctx->conv_xo.p));
  HS_CUDA(ctx, cudaMemcpyAsync(p_out, ctx->conv_p.p, sizeof(int64_t) * (size_t)(n_rows + 1), cudaMemcpyDeviceToHost, st));
  if (nnz > 0) {
    HS_CUDA(ctx, cudaMemcpyAsync(j_out, ctx->conv_j.p, sizeof(int32_t) * (size_t)nnz, cudaMemcpyDeviceToHost, st));
    HS_CUDA(ctx, cudaMemcpyAsync(x_out, ctx->conv_xo.p, sizeof(double) * (size_t)nnz, cudaMemcpyDeviceToHost, st));
  }
  HS_CUDA(ctx, cudaStreamSynchronize(st));
  return HS_OK;
  } HS_ABI_CATCH(ctx)
}

}  // extern "C"
#pragma GCC visibility pop
