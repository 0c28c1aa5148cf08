// development probe: host-side narrowing (fp64 -> fp32) and id packing throughput of the GPU box's CPU
// g++ -O3 -pthread tools/host_probe.cpp -o tools/host_probe.bin
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>
#include <immintrin.h>

static void narrow_sse(const double* s, float* d, int64_t n) { for (int64_t i = 0; i < n; ++i) d[i] = (float)s[i]; }
__attribute__((target("avx2"))) static void narrow_avx2(const double* s, float* d, int64_t n) {
  int64_t i = 0;
  for (; i + 8 <= n; i += 8) {
    __m128 a = _mm256_cvtpd_ps(_mm256_loadu_pd(s + i)), b = _mm256_cvtpd_ps(_mm256_loadu_pd(s + i + 4));
    _mm256_stream_ps(d + i, _mm256_set_m128(b, a));
  }
  for (; i < n; ++i) d[i] = (float)s[i];
}
__attribute__((target("avx512f,avx512dq"))) static void narrow_avx512(const double* s, float* d, int64_t n) {
  int64_t i = 0;
  for (; i + 16 <= n; i += 16) {
    __m256 a = _mm512_cvtpd_ps(_mm512_loadu_pd(s + i)), b = _mm512_cvtpd_ps(_mm512_loadu_pd(s + i + 8));
    _mm512_stream_ps(d + i, _mm512_insertf32x8(_mm512_castps256_ps512(a), b, 1));
  }
  for (; i < n; ++i) d[i] = (float)s[i];
}
static void pack(const int32_t* a, int64_t n, unsigned long long* w, int64_t words) {
  memset(w, 0, 8 * words);
  int64_t prev_w = -1; unsigned long long acc = 0;
  for (int64_t i = 0; i < n; ++i) { const int64_t c = a[i], wi = c >> 6; acc = (wi == prev_w ? acc : 0ull) | (1ull << (c & 63)); w[wi] = acc; prev_w = wi; }
}
template <class F> static double par(int nt, F f) {
  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> th;
  for (int t = 0; t < nt; ++t) th.emplace_back([=] { f(t); });
  for (auto& x : th) x.join();
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
int main() {
  const int64_t n = (int64_t)1 << 29;     // 512M values: 4 GB of doubles
  std::vector<double> src(n, 1.25);
  float* dst = (float*)aligned_alloc(64, n * 4);
  memset(dst, 0, n * 4);
  printf("avx2 %d avx512f %d hw threads %u\n", __builtin_cpu_supports("avx2"), __builtin_cpu_supports("avx512f"), std::thread::hardware_concurrency());
  for (int nt : {8, 16, 32}) {
    const int64_t per = n / nt;
    double t = par(nt, [&](int k) { narrow_sse(src.data() + k * per, dst + k * per, per); });
    printf("narrow sse    %2d threads: %6.2f G values/s\n", nt, n / t / 1e9);
    if (__builtin_cpu_supports("avx2")) { t = par(nt, [&](int k) { narrow_avx2(src.data() + k * per, dst + k * per, per); }); printf("narrow avx2   %2d threads: %6.2f G values/s\n", nt, n / t / 1e9); }
    if (__builtin_cpu_supports("avx512f")) { t = par(nt, [&](int k) { narrow_avx512(src.data() + k * per, dst + k * per, per); }); printf("narrow avx512 %2d threads: %6.2f G values/s\n", nt, n / t / 1e9); }
  }
  // ids: rows of 100000 ascending ids in [0, 1e6)
  const int64_t rows = 2048, nnz = 100000, N = 1000000, words = (N + 63) / 64;
  std::vector<int32_t> ids(rows * nnz);
  for (int64_t r = 0; r < rows; ++r) for (int64_t i = 0; i < nnz; ++i) ids[r * nnz + i] = (int32_t)(i * 10 + (r * 7 + i * 3) % 10);
  std::vector<unsigned long long> bits(rows * words);
  for (int nt : {8, 16, 32}) {
    double t = par(nt, [&](int k) { for (int64_t r = k; r < rows; r += nt) pack(ids.data() + r * nnz, nnz, bits.data() + r * words, words); });
    printf("pack ids      %2d threads: %6.2f G ids/s\n", nt, rows * nnz / t / 1e9);
  }
  return 0;
}
