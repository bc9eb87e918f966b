/* orx_oracle.c -- C/OpenMP port of oracle/openrec_oracle.py's pairwise training step.
 *
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY (never linked into liborx, never on the product path).
 * It restates, op for op, what the reference's TF2 graph does for one BPR/UCML step
 * (openrec/tf2/recommenders/bpr.py:21-37, ucml.py:21-42, modules/pairwise_log_loss.py:15-34,
 * tf2_examples/bpr_citeulike.py:33-39): gather rows -> score -> loss -> per-lookup gradient rows
 * (IndexedSlices) -> dedup (unique + segment sum, batch order) -> optimizer apply per unique row.
 * PARITY UNPINNED for the TensorFlow-internal parts -- see oracle/__init__.py; it is pinned to
 * the numpy oracle by tests/test_oracle_c.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* scratch buffers are kept across calls (a fresh 100 MB malloc per step would bill page faults
 * to the CPU baseline); single caller at a time. */
static void* scratch(int slot, size_t bytes) {
  static void* buf[8];
  static size_t cap[8];
  if (bytes > cap[slot]) {
    free(buf[slot]);
    buf[slot] = malloc(bytes);
    cap[slot] = bytes;
  }
  return buf[slot];
}

/* stable LSD radix sort of (id << 32 | position) keys by id, 11 bits per pass: the dedup of TF's IndexedSlices
 * aggregation (unique + segment sum, batch order inside a segment).  qsort with a comparison callback took a third
 * of the step on its own and does not scale with threads -- the CPU baseline should not be a strawman. */
static void radix_sort_by_id(uint64_t* key, uint64_t* tmp, int n, int64_t max_id) {
  int bits = 1;
  while (bits < 32 && ((int64_t)1 << bits) <= max_id) ++bits;
  uint64_t *src = key, *dst = tmp;
  for (int shift = 0; shift < bits; shift += 11) {
    size_t hist[2049];
    memset(hist, 0, sizeof(hist));
    for (int i = 0; i < n; ++i) ++hist[((src[i] >> (32 + shift)) & 2047u) + 1];
    for (int b = 0; b < 2048; ++b) hist[b + 1] += hist[b];
    for (int i = 0; i < n; ++i) dst[hist[(src[i] >> (32 + shift)) & 2047u]++] = src[i];
    uint64_t* t = src; src = dst; dst = t;
  }
  if (src != key) memcpy(key, src, sizeof(uint64_t) * (size_t)n);
}

/* dedup + apply: indices[n], values[n,D] -> per unique row: G = sum (batch order); opt: 0 SGD, 1 Adagrad */
static void sparse_apply(float* var, float* acc, int D, const int32_t* idx, const float* val, int n, int opt,
                         float lr, float eps) {
  uint64_t* key = (uint64_t*)scratch(0, sizeof(uint64_t) * (size_t)n);
  uint64_t* tmp = (uint64_t*)scratch(6, sizeof(uint64_t) * (size_t)n);
  int32_t max_id = 0;
  for (int i = 0; i < n; ++i) {
    key[i] = ((uint64_t)(uint32_t)idx[i] << 32) | (uint32_t)i;
    if (idx[i] > max_id) max_id = idx[i];
  }
  radix_sort_by_id(key, tmp, n, max_id);
  int* seg = (int*)scratch(1, sizeof(int) * (size_t)(n + 1));
  int ns = 0;
  for (int i = 0; i < n; ++i)
    if (i == 0 || (key[i] >> 32) != (key[i - 1] >> 32)) seg[ns++] = i;
  seg[ns] = n;
#pragma omp parallel
  {
    float* g = (float*)malloc(sizeof(float) * (size_t)D);
#pragma omp for schedule(static)
    for (int s = 0; s < ns; ++s) {
      memset(g, 0, sizeof(float) * (size_t)D);
      for (int i = seg[s]; i < seg[s + 1]; ++i) {
        const float* v = val + (size_t)(uint32_t)key[i] * D;
        for (int d = 0; d < D; ++d) g[d] += v[d];
      }
      float* w = var + (size_t)(key[seg[s]] >> 32) * D;
      if (opt == 0) {
        for (int d = 0; d < D; ++d) w[d] -= lr * g[d];
      } else {
        float* a = acc + (size_t)(key[seg[s]] >> 32) * D;
        for (int d = 0; d < D; ++d) {
          a[d] += g[d] * g[d];
          w[d] -= lr * g[d] / (sqrtf(a[d]) + eps);
        }
      }
    }
    free(g);
  }
}

/* kind 0 BPR / 1 UCML ; opt 0 SGD / 1 Adagrad.  out2 = {loss, l2_loss}.  Returns 0, or -1 on bad args. */
int orx_oracle_pairwise_step(int kind, float* U, float* Ua, float* I, float* Ia, float* Bv, float* Ba, int64_t rowsU,
                             int64_t rowsI, int D, const int32_t* uid, const int32_t* pid, const int32_t* nid, int B,
                             float margin, float c_loss, float c_l2, int opt, float lr, float eps, int nthreads,
                             double* out2) {
  if (B <= 0 || D <= 0 || (opt != 0 && opt != 1) || (kind != 0 && kind != 1)) return -1;
  for (int b = 0; b < B; ++b)
    if (uid[b] < 0 || uid[b] >= rowsU || pid[b] < 0 || pid[b] >= rowsI || nid[b] < 0 || nid[b] >= rowsI) return -1;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  float* gu = (float*)scratch(2, sizeof(float) * (size_t)B * D);         /* IndexedSlices values, user table */
  float* gi = (float*)scratch(3, sizeof(float) * (size_t)2 * B * D);     /* item table: p || n */
  float* gb = (float*)scratch(4, sizeof(float) * (size_t)2 * B);         /* item_bias: p || n */
  int32_t* ii = (int32_t*)scratch(5, sizeof(int32_t) * (size_t)2 * B);
  double loss = 0.0, l2 = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : loss, l2)
  for (int b = 0; b < B; ++b) {
    const float* u = U + (size_t)uid[b] * D;
    const float* p = I + (size_t)pid[b] * D;
    const float* n = I + (size_t)nid[b] * D;
    float s1 = 0.f, s2 = 0.f, sq = 0.f;
    for (int d = 0; d < D; ++d) {
      if (kind == 0) {
        s1 += u[d] * p[d];
        s2 += u[d] * n[d];
      } else {
        s1 += (u[d] - p[d]) * (u[d] - p[d]);
        s2 += (u[d] - n[d]) * (u[d] - n[d]);
      }
      sq += u[d] * u[d] + p[d] * p[d] + n[d] * n[d];
    }
    l2 += 0.5 * (double)sq;
    const float bp = Bv[pid[b]], bn = Bv[nid[b]];
    float g, gbias;
    float* ru = gu + (size_t)b * D;
    float* rp = gi + (size_t)b * D;
    float* rn = gi + (size_t)(B + b) * D;
    if (kind == 0) {
      const float x = (s1 + bp) - (s2 + bn);
      const float y = x > -30.f ? x : -30.f;
      const float e = expf(-fabsf(y));
      loss += -(double)((y < 0.f ? y : 0.f) - log1pf(e)) / (double)B;
      const float sn = y >= 0.f ? e / (1.f + e) : 1.f / (1.f + e);
      g = x >= -30.f ? -(c_loss / (float)B) * sn : 0.f;
      gbias = g;
      for (int d = 0; d < D; ++d) {
        ru[d] = g * (p[d] - n[d]) + c_l2 * u[d];
        rp[d] = g * u[d] + c_l2 * p[d];
        rn[d] = -g * u[d] + c_l2 * n[d];
      }
    } else {
      const float h = margin - (((-s1) + bp) - ((-s2) + bn));
      loss += (double)(h > 0.f ? h : 0.f);
      g = h >= 0.f ? c_loss : 0.f;
      gbias = -g;
      const float t = 2.f * g;
      for (int d = 0; d < D; ++d) {
        ru[d] = t * (n[d] - p[d]) + c_l2 * u[d];
        rp[d] = t * (p[d] - u[d]) + c_l2 * p[d];
        rn[d] = t * (u[d] - n[d]) + c_l2 * n[d];
      }
    }
    gb[b] = gbias;
    gb[B + b] = -gbias;
    ii[b] = pid[b];
    ii[B + b] = nid[b];
  }
  sparse_apply(U, Ua, D, uid, gu, B, opt, lr, eps);
  sparse_apply(I, Ia, D, ii, gi, 2 * B, opt, lr, eps);
  sparse_apply(Bv, Ba, 1, ii, gb, 2 * B, opt, lr, eps);
  out2[0] = loss;
  out2[1] = l2;
  return 0;
}

int orx_oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
