/*
 * r_base.c -- TEST INFRASTRUCTURE (oracle).  Plain-C restatement of the pieces of base R that the
 * reference's continuous path calls and that are NOT vendored under /root/reference:
 *
 *   set.seed(seed)                    initial scrambling + Mersenne-Twister seeding   (R: src/main/RNG.c)
 *   unif_rand()                       MT19937 + fixup into (0, 1)                      (R: src/main/RNG.c)
 *   R_unif_index(dn)                  sample.kind = "Rejection" (R >= 3.6.0) or "Rounding"
 *   sample.int(n, size)               partial Fisher-Yates of do_sample                (R: src/main/random.c)
 *   stats::kmeans(algorithm = "Hartigan-Wong")   AS 136 (Hartigan & Wong 1979) as in R: src/library/stats/src/kmns.c
 *
 * Call sites in the reference: sample() at R/haystack_continuous.R:259, :275, :437, :505 and
 * R/haystack_highD.R:95, :110; kmeans() at R/haystack_highD.R:83.  R itself cannot run in this image; these
 * functions are pinned by reproducing the reference's only printed results (docs/articles/a01_toy_example.html:129,
 * :168, rendered 2023-08 with set.seed(1234)) -- tests/test_r_parity_cpu.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this library.  Written from the published
 * algorithms (Matsumoto & Nishimura 1998; AS 136) and R's documented behaviour; no R source is copied.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MT_N 624
#define MT_M 397

typedef struct {
  uint32_t mt[MT_N];
  int mti;
  int sample_rounding; /* 0: "Rejection" (R >= 3.6.0 default), 1: "Rounding" */
} r_rng;

/* set.seed(): the seed is scrambled 50 times by the LCG 69069 x + 1, the next 625 values fill the generator's
 * seed vector; its first word is the position counter, which the fix-up sets to N (= "regenerate on first use"). */
void r_set_seed(r_rng* g, uint32_t seed, int sample_rounding) {
  for (int j = 0; j < 50; ++j) seed = 69069u * seed + 1u;
  seed = 69069u * seed + 1u; /* i_seed[0]: overwritten by the position counter */
  for (int j = 0; j < MT_N; ++j) {
    seed = 69069u * seed + 1u;
    g->mt[j] = seed;
  }
  g->mti = MT_N;
  g->sample_rounding = sample_rounding;
}

static uint32_t mt_next(r_rng* g) {
  static const uint32_t mag01[2] = {0x0u, 0x9908b0dfu};
  uint32_t y;
  if (g->mti >= MT_N) {
    int kk;
    for (kk = 0; kk < MT_N - MT_M; ++kk) {
      y = (g->mt[kk] & 0x80000000u) | (g->mt[kk + 1] & 0x7fffffffu);
      g->mt[kk] = g->mt[kk + MT_M] ^ (y >> 1) ^ mag01[y & 1u];
    }
    for (; kk < MT_N - 1; ++kk) {
      y = (g->mt[kk] & 0x80000000u) | (g->mt[kk + 1] & 0x7fffffffu);
      g->mt[kk] = g->mt[kk + (MT_M - MT_N)] ^ (y >> 1) ^ mag01[y & 1u];
    }
    y = (g->mt[MT_N - 1] & 0x80000000u) | (g->mt[0] & 0x7fffffffu);
    g->mt[MT_N - 1] = g->mt[MT_M - 1] ^ (y >> 1) ^ mag01[y & 1u];
    g->mti = 0;
  }
  y = g->mt[g->mti++];
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}

/* unif_rand(): [0,1) value of the generator moved into the open interval */
double r_unif_rand(r_rng* g) {
  const double i2_32m1 = 2.328306437080797e-10; /* 1 / (2^32 - 1) */
  double v = (double)mt_next(g) * 2.3283064365386963e-10; /* 2^-32 */
  if (v <= 0.0) return 0.5 * i2_32m1;
  if (1.0 - v <= 0.0) return 1.0 - 0.5 * i2_32m1;
  return v;
}

/* random bits, 16 per uniform */
static double r_rbits(r_rng* g, int bits) {
  int64_t v = 0;
  for (int n = 0; n <= bits; n += 16) {
    int v1 = (int)floor(r_unif_rand(g) * 65536.0);
    v = 65536 * v + v1;
  }
  if (bits < 64) v &= ((int64_t)1 << bits) - 1;
  return (double)v;
}

/* uniform index in [0, dn) */
double r_unif_index(r_rng* g, double dn) {
  if (g->sample_rounding) return floor(dn * r_unif_rand(g));
  if (dn <= 0) return 0.0;
  int bits = (int)ceil(log2(dn));
  double dv;
  do {
    dv = r_rbits(g, bits);
  } while (dn <= dv);
  return dv;
}

/* sample.int(n, k) without replacement, no prob: 1-based ids into out[k]; scratch must hold n ints.
 * (k < 2 draws directly, as R does.) */
void r_sample_int(r_rng* g, int n, int k, int32_t* out, int32_t* scratch) {
  if (k < 2) {
    for (int i = 0; i < k; ++i) out[i] = (int32_t)(r_unif_index(g, (double)n) + 1);
    return;
  }
  for (int i = 0; i < n; ++i) scratch[i] = i;
  for (int i = 0; i < k; ++i) {
    int j = (int)r_unif_index(g, (double)n);
    out[i] = scratch[j] + 1;
    scratch[j] = scratch[--n];
  }
}

/* count x sample.int(n, k): out is count rows of k (row-major) */
void r_sample_int_many(r_rng* g, int n, int k, int count, int32_t* out) {
  int32_t* scratch = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
  for (int c = 0; c < count; ++c) r_sample_int(g, n, k, out + (size_t)c * k, scratch);
  free(scratch);
}

/* sample(n, 1, prob = p): one id (1-based) with the given (unnormalised, non-negative) weights, as do_sample
 * handles size < 2: weights normalised, Walker alias when more than 200 of them are "large", else the sorted
 * cumulative search. */
static void revsort_idx(double* a, int* ib, int n) { /* sort a descending, carrying ib (heapsort as R's revsort) */
  int l, j, ir, i;
  double ra;
  int ii;
  if (n <= 1) return;
  a--; ib--;
  l = (n >> 1) + 1;
  ir = n;
  for (;;) {
    if (l > 1) {
      l = l - 1;
      ra = a[l];
      ii = ib[l];
    } else {
      ra = a[ir];
      ii = ib[ir];
      a[ir] = a[1];
      ib[ir] = ib[1];
      if (--ir == 1) {
        a[1] = ra;
        ib[1] = ii;
        return;
      }
    }
    i = l;
    j = l << 1;
    while (j <= ir) {
      if (j < ir && a[j] > a[j + 1]) ++j;
      if (ra > a[j]) {
        a[i] = a[j];
        ib[i] = ib[j];
        j += (i = j);
      } else
        j = ir + 1;
    }
    a[i] = ra;
    ib[i] = ii;
  }
}

int r_sample_prob_one(r_rng* g, int n, const double* prob) {
  double* p = (double*)malloc(sizeof(double) * (size_t)n);
  int* perm = (int*)malloc(sizeof(int) * (size_t)n);
  double sum = 0.0;
  int nc = 0, ans = 1;
  for (int i = 0; i < n; ++i) sum += prob[i];
  for (int i = 0; i < n; ++i) p[i] = prob[i] / sum;
  for (int i = 0; i < n; ++i)
    if (n * p[i] > 0.1) ++nc;
  if (nc > 200) {
    /* Walker alias method */
    double* q = (double*)malloc(sizeof(double) * (size_t)n);
    int* a = (int*)malloc(sizeof(int) * (size_t)n);
    int* HL = (int*)malloc(sizeof(int) * (size_t)n);
    int *H = HL - 1, *L = HL + n;
    for (int i = 0; i < n; ++i) {
      q[i] = p[i] * n;
      if (q[i] < 1.0) *++H = i; else *--L = i;
    }
    if (H >= HL && L < HL + n) {
      for (int k = 0; k < n - 1; ++k) {
        int i = HL[k], j = *L;
        a[i] = j;
        q[j] += q[i] - 1;
        if (q[j] < 1.0) L++;
        if (L >= HL + n) break;
      }
    }
    for (int i = 0; i < n; ++i) q[i] += i;
    double rU = r_unif_rand(g) * n;
    int k = (int)rU;
    ans = (rU < q[k]) ? k + 1 : a[k] + 1;
    free(q); free(a); free(HL);
  } else {
    for (int i = 0; i < n; ++i) perm[i] = i + 1;
    revsort_idx(p, perm, n);
    for (int i = 1; i < n; ++i) p[i] += p[i - 1];
    double rU = r_unif_rand(g);
    int j;
    for (j = 0; j < n - 1; ++j)
      if (rU <= p[j]) break;
    ans = perm[j];
  }
  free(p); free(perm);
  return ans;
}

/* ---------------------------------------------------------------------------------------------------------
 * Hartigan & Wong (1979), Algorithm AS 136.  a: m x n data, column-major (R matrix); c: k x n centres,
 * column-major, initial centres in, final centres out.  ic1: closest-centre index (1-based) per point;
 * nc: cluster sizes; wss: within-cluster sums of squares.  iter: in = iter.max, out = iterations used.
 * ifault: 0 ok, 1 empty cluster, 2 no convergence in iter.max, 4 quick-transfer steps exceeded.
 * --------------------------------------------------------------------------------------------------------- */
#define A(i, j) a[(size_t)(i) + (size_t)(j) * m]
#define C(l, j) c[(size_t)(l) + (size_t)(j) * k]

static void hw_optra(const double* a, int m, int n, double* c, int k, int* ic1, int* ic2, int* nc, double* an1,
                     double* an2, int* ncp, double* d, int* itran, int* live, int* indx) {
  const double big = 1e30;
  for (int l = 0; l < k; ++l)
    if (itran[l] == 1) live[l] = m; /* clusters updated in the last quick-transfer stage are live throughout */
  for (int i = 0; i < m; ++i) {
    ++(*indx);
    int l1 = ic1[i];
    if (nc[l1] != 1) { /* a point that is alone in its cluster stays */
      if (ncp[l1] != 0) { /* cluster updated since: recompute the point's cost of staying */
        double de = 0.0;
        for (int j = 0; j < n; ++j) {
          double df = A(i, j) - C(l1, j);
          de += df * df;
        }
        d[i] = de * an1[l1];
      }
      double da = 0.0;
      int l2 = ic2[i], ll = l2;
      for (int j = 0; j < n; ++j) {
        double db = A(i, j) - C(l2, j);
        da += db * db;
      }
      double r2 = da * an2[l2];
      for (int l = 0; l < k; ++l) {
        /* if neither the point's cluster nor l is live (step numbers are 1-based), l cannot have improved */
        if ((i + 1 >= live[l1] && i + 1 >= live[l]) || l == l1 || l == ll) continue;
        double rr = r2 / an2[l];
        double dc = 0.0;
        int j;
        for (j = 0; j < n; ++j) {
          double dd = A(i, j) - C(l, j);
          dc += dd * dd;
          if (dc >= rr) break;
        }
        if (j < n) continue;
        r2 = dc * an2[l];
        l2 = l;
      }
      if (r2 >= d[i]) {
        ic2[i] = l2; /* no transfer; l2 is the new second-closest */
      } else {
        *indx = 0;
        live[l1] = m + i + 1;
        live[l2] = m + i + 1;
        ncp[l1] = i + 1;
        ncp[l2] = i + 1;
        double al1 = nc[l1], alw = al1 - 1.0, al2 = nc[l2], alt = al2 + 1.0;
        for (int j = 0; j < n; ++j) {
          C(l1, j) = (C(l1, j) * al1 - A(i, j)) / alw;
          C(l2, j) = (C(l2, j) * al2 + A(i, j)) / alt;
        }
        --nc[l1];
        ++nc[l2];
        an2[l1] = alw / al1;
        an1[l1] = (alw > 1.0) ? alw / (alw - 1.0) : big;
        an1[l2] = alt / al2;
        an2[l2] = alt / (alt + 1.0);
        ic1[i] = l2;
        ic2[i] = l1;
      }
    }
    if (*indx == m) return;
  }
  for (int l = 0; l < k; ++l) {
    itran[l] = 0;
    live[l] -= m;
  }
}

static void hw_qtran(const double* a, int m, int n, double* c, int k, int* ic1, int* ic2, int* nc, double* an1,
                     double* an2, int* ncp, double* d, int* itran, int* indx, int* imaxqtr) {
  const double big = 1e30;
  int icoun = 0, istep = 0;
  for (;;) {
    for (int i = 0; i < m; ++i) {
      ++icoun;
      ++istep;
      if (istep >= *imaxqtr) {
        *imaxqtr = -1;
        return;
      }
      int l1 = ic1[i], l2 = ic2[i];
      if (nc[l1] != 1) {
        if (istep <= ncp[l1]) {
          double da = 0.0;
          for (int j = 0; j < n; ++j) {
            double db = A(i, j) - C(l1, j);
            da += db * db;
          }
          d[i] = da * an1[l1];
        }
        if (istep < ncp[l1] || istep < ncp[l2]) {
          double r2 = d[i] / an2[l2];
          double dd = 0.0;
          int j;
          for (j = 0; j < n; ++j) {
            double de = A(i, j) - C(l2, j);
            dd += de * de;
            if (dd >= r2) break;
          }
          if (j == n) {
            icoun = 0;
            *indx = 0;
            itran[l1] = 1;
            itran[l2] = 1;
            ncp[l1] = istep + m;
            ncp[l2] = istep + m;
            double al1 = nc[l1], alw = al1 - 1.0, al2 = nc[l2], alt = al2 + 1.0;
            for (int jj = 0; jj < n; ++jj) {
              C(l1, jj) = (C(l1, jj) * al1 - A(i, jj)) / alw;
              C(l2, jj) = (C(l2, jj) * al2 + A(i, jj)) / alt;
            }
            --nc[l1];
            ++nc[l2];
            an2[l1] = alw / al1;
            an1[l1] = (alw > 1.0) ? alw / (alw - 1.0) : big;
            an1[l2] = alt / al2;
            an2[l2] = alt / (alt + 1.0);
            ic1[i] = l2;
            ic2[i] = l1;
          }
        }
      }
      if (icoun == m) return;
    }
  }
}

void r_kmeans_hartigan_wong(const double* a, int m, int n, double* c, int k, int* ic1, int* nc, int* iter,
                            double* wss, int* ifault, int imaxqtr) {
  const double big = 1e30;
  int* ic2 = (int*)malloc(sizeof(int) * (size_t)m);
  double* an1 = (double*)malloc(sizeof(double) * (size_t)k);
  double* an2 = (double*)malloc(sizeof(double) * (size_t)k);
  int* ncp = (int*)malloc(sizeof(int) * (size_t)k);
  double* d = (double*)calloc((size_t)m, sizeof(double));
  int* itran = (int*)malloc(sizeof(int) * (size_t)k);
  int* live = (int*)calloc((size_t)k, sizeof(int));
  int ij = 0;
  *ifault = 3;
  if (k <= 1 || k >= m) goto done;

  /* the two closest centres of every point */
  for (int i = 0; i < m; ++i) {
    double dt[2];
    ic1[i] = 0;
    ic2[i] = 1;
    for (int il = 0; il < 2; ++il) {
      dt[il] = 0.0;
      for (int j = 0; j < n; ++j) {
        double da = A(i, j) - C(il, j);
        dt[il] += da * da;
      }
    }
    if (dt[0] > dt[1]) {
      ic1[i] = 1;
      ic2[i] = 0;
      double t = dt[0];
      dt[0] = dt[1];
      dt[1] = t;
    }
    for (int l = 2; l < k; ++l) {
      double db = 0.0;
      int j;
      for (j = 0; j < n; ++j) {
        double dc = A(i, j) - C(l, j);
        db += dc * dc;
        if (db >= dt[1]) break;
      }
      if (j < n) continue;
      if (db < dt[0]) {
        dt[1] = dt[0];
        ic2[i] = ic1[i];
        dt[0] = db;
        ic1[i] = l;
      } else {
        dt[1] = db;
        ic2[i] = l;
      }
    }
  }
  /* centres = means of their points */
  for (int l = 0; l < k; ++l) {
    nc[l] = 0;
    for (int j = 0; j < n; ++j) C(l, j) = 0.0;
  }
  for (int i = 0; i < m; ++i) {
    int l = ic1[i];
    ++nc[l];
    for (int j = 0; j < n; ++j) C(l, j) += A(i, j);
  }
  for (int l = 0; l < k; ++l)
    if (nc[l] == 0) {
      *ifault = 1;
      goto done;
    }
  for (int l = 0; l < k; ++l) {
    double aa = nc[l];
    for (int j = 0; j < n; ++j) C(l, j) /= aa;
    an2[l] = aa / (aa + 1.0);
    an1[l] = (aa > 1.0) ? aa / (aa - 1.0) : big;
    itran[l] = 1;
    ncp[l] = -1;
  }
  {
    int indx = 0;
    *ifault = 2;
    for (ij = 1; ij <= *iter; ++ij) {
      hw_optra(a, m, n, c, k, ic1, ic2, nc, an1, an2, ncp, d, itran, live, &indx);
      if (indx == m) { /* no transfer in the last m optimal-transfer steps */
        *ifault = 0;
        break;
      }
      hw_qtran(a, m, n, c, k, ic1, ic2, nc, an1, an2, ncp, d, itran, &indx, &imaxqtr);
      if (imaxqtr < 0) {
        *ifault = 4;
        break;
      }
      if (k == 2) {
        *ifault = 0;
        break;
      }
      for (int l = 0; l < k; ++l) ncp[l] = 0;
    }
    *iter = ij;
  }
  /* final centres and within-cluster sums of squares */
  for (int l = 0; l < k; ++l) {
    wss[l] = 0.0;
    for (int j = 0; j < n; ++j) C(l, j) = 0.0;
  }
  for (int i = 0; i < m; ++i) {
    int ii = ic1[i];
    for (int j = 0; j < n; ++j) C(ii, j) += A(i, j);
  }
  for (int j = 0; j < n; ++j) {
    for (int l = 0; l < k; ++l) C(l, j) /= (double)nc[l];
    for (int i = 0; i < m; ++i) {
      int ii = ic1[i];
      double da = A(i, j) - C(ii, j);
      wss[ii] += da * da;
    }
  }
done:
  for (int i = 0; i < m; ++i) ic1[i] += 1; /* 1-based cluster ids like R */
  free(ic2); free(an1); free(an2); free(ncp); free(d); free(itran); free(live);
}

size_t r_rng_sizeof(void) { return sizeof(r_rng); }
