// orx_misc.cu -- LatentFactor init / gather / censor, dense optimizer apply, full-catalogue scoring,
// ranking metrics.
#include "orx_common.cuh"

// ---------------------------------------------------------------------------------------
// LatentFactor.__init__ initializer (latent_factor.py:8-15): U(lo,hi) from a counter-based hash
// (TF's RNG stream is not reproducible across frameworks; only the distribution matters).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__global__ void k_fill_uniform(float* dst, int64_t n, float lo, float hi, uint64_t seed) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint64_t r = splitmix64(seed * 0xD1342543DE82EF95ull + (uint64_t)i);
    const float u = (float)(r >> 40) * (1.0f / 16777216.0f);  // [0,1)
    dst[i] = lo + (hi - lo) * u;
  }
}

extern "C" int orx_fill_uniform(orx_handle_t h, float* dst, int64_t n, float lo, float hi, uint64_t seed,
                                orx_stream_t s) {
  ORX_REQUIRE(h != nullptr && dst != nullptr && n >= 0, "null handle/dst or negative n");
  if (n == 0) return ORX_OK;
  ORX_CUDA(cudaSetDevice(h->device));
  int64_t blocks = (n + 1023) / 1024;
  if (blocks > (int64_t)h->num_sms * 16) blocks = (int64_t)h->num_sms * 16;
  k_fill_uniform<<<(int)blocks, 256, 0, (cudaStream_t)s>>>(dst, n, lo, hi, seed);
  ORX_LAUNCH_CHECK();
  return ORX_OK;
}

// ---------------------------------------------------------------------------------------
// LatentFactor.__call__ (Embedding.call): out[b,:] = tab[ids[b],:]
// ---------------------------------------------------------------------------------------
template <typename IdT>
__global__ void __launch_bounds__(256) k_gather(const float* __restrict__ tab, int64_t rows, int D,
                                                const IdT* __restrict__ ids, int64_t n, float* __restrict__ out,
                                                int32_t* n_bad) {
  const int lane = threadIdx.x & 31;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const bool vec = (D & 3) == 0;
  for (int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < n; b += nw) {
    const int64_t id = (int64_t)ids[b];
    const bool ok = id >= 0 && id < rows;
    if (!ok && lane == 0 && n_bad) atomicAdd(n_bad, 1);
    if (vec) {
      const float4* src = reinterpret_cast<const float4*>(tab + id * D);
      float4* dst = reinterpret_cast<float4*>(out + b * D);
      for (int e = lane; e < D / 4; e += 32) dst[e] = ok ? __ldg(src + e) : make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      for (int e = lane; e < D; e += 32) out[b * D + e] = ok ? __ldg(tab + id * D + e) : 0.f;
    }
  }
}

extern "C" int orx_gather(orx_handle_t h, const float* tab, int64_t rows, int32_t dim, const void* ids,
                          int32_t id_is_i64, int64_t n, float* out, int32_t* n_bad, orx_stream_t s) {
  ORX_REQUIRE(h != nullptr && tab && ids && out, "null pointer");
  ORX_REQUIRE(rows > 0 && dim > 0 && n >= 0, "bad sizes");
  if (n == 0) return ORX_OK;
  ORX_CUDA(cudaSetDevice(h->device));
  int64_t blocks = (n + 7) / 8;
  if (blocks > (int64_t)h->num_sms * 32) blocks = (int64_t)h->num_sms * 32;
  if (id_is_i64)
    k_gather<int64_t><<<(int)blocks, 256, 0, (cudaStream_t)s>>>(tab, rows, dim, (const int64_t*)ids, n, out, n_bad);
  else
    k_gather<int32_t><<<(int)blocks, 256, 0, (cudaStream_t)s>>>(tab, rows, dim, (const int32_t*)ids, n, out, n_bad);
  ORX_LAUNCH_CHECK();
  return ORX_OK;
}

// ---------------------------------------------------------------------------------------
// LatentFactor.censor (latent_factor.py:17-23), "K10": unique ids via the batch hash (the first
// warp to insert an id owns the row), row <- row / max(||row||, min_norm).
// ---------------------------------------------------------------------------------------
// A warp takes 8 ids per iteration: lanes 0..7 load the ids and claim the rows in the hash in parallel, then (128-bit path)
// all 8 rows are loaded before the first norm is reduced -- one row per warp left the kernel latency-bound (id -> hash ->
// row -> reduce -> divide -> store: 23 us for 65 536 ids at D = 128, three of them per UCML step).
__global__ void __launch_bounds__(256) k_censor(float* tab, int64_t rows, int D, const int32_t* __restrict__ ids,
                                                int n, float min_norm, OrxHash hsh) {
  const int lane = threadIdx.x & 31;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  const bool vec = (D & 3) == 0 && D <= 128;
  for (int b0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 8; b0 < n; b0 += nw * 8) {
    int32_t my_id = -1;
    if (lane < 8 && b0 + lane < n) {
      const int32_t id = ids[b0 + lane];
      if (id >= 0 && (int64_t)id < rows && orx_hash_insert(hsh, id, 2) == 0u) my_id = id;   // first claim owns the row
    }
    if (vec) {
      const int nq = D >> 2;
      float4 v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int32_t id = __shfl_sync(ORX_FULL, my_id, k);
        v[k] = (id >= 0 && lane < nq) ? __ldcg(reinterpret_cast<const float4*>(tab + (int64_t)id * D) + lane)
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int32_t id = __shfl_sync(ORX_FULL, my_id, k);
        if (id < 0) continue;
        float sq = v[k].x * v[k].x + v[k].y * v[k].y + v[k].z * v[k].z + v[k].w * v[k].w;
        sq = orx_group_sum<32>(sq);
        const float den = fmaxf(sqrtf(sq), min_norm);
        if (lane < nq)
          __stcg(reinterpret_cast<float4*>(tab + (int64_t)id * D) + lane,
                 make_float4(v[k].x / den, v[k].y / den, v[k].z / den, v[k].w / den));
      }
    } else {
      for (int k = 0; k < 8; ++k) {
        const int32_t id = __shfl_sync(ORX_FULL, my_id, k);
        if (id < 0) continue;
        float* row = tab + (int64_t)id * D;
        float sq = 0.f;
        for (int e = lane; e < D; e += 32) sq += row[e] * row[e];
        sq = orx_group_sum<32>(sq);
        const float den = fmaxf(sqrtf(sq), min_norm);
        for (int e = lane; e < D; e += 32) row[e] = row[e] / den;
      }
    }
  }
}

extern "C" int orx_censor(orx_handle_t h, float* tab, int64_t rows, int32_t dim, const int32_t* ids, int32_t n,
                          float min_norm, orx_stream_t s) {
  ORX_REQUIRE(h != nullptr && tab && ids, "null pointer");
  ORX_REQUIRE(rows > 0 && dim > 0 && n >= 0, "bad sizes");
  if (n == 0) return ORX_OK;
  ORX_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)s;
  int rc = orx_ensure_workspace(h, n, h->g_dim > 0 ? h->g_dim : 1, false);
  if (rc) return rc;
  if ((rc = orx_next_epoch(h, st))) return rc;   // the dedup hash needs no clearing: a new epoch empties it
  int blocks = (n + 63) / 64;                 // 8 warps x 8 ids per block and iteration
  if (blocks > h->num_sms * 8) blocks = h->num_sms * 8;
  k_censor<<<blocks, 256, 0, st>>>(tab, rows, dim, ids, n, min_norm, h->hu);
  ORX_LAUNCH_CHECK();
  return ORX_OK;
}

// ---------------------------------------------------------------------------------------
// Keras dense apply for dense variables (GMF w, MLP kernels / biases)
// ---------------------------------------------------------------------------------------
template <int OPT>
__global__ void k_dense_apply(float* var, float* s0, float* s1, const float* __restrict__ grad, int64_t n,
                              OrxOptDev o) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float a = (OPT != ORX_OPT_SGD) ? s0[i] : 0.f, b = (OPT == ORX_OPT_ADAM_LAZY) ? s1[i] : 0.f;
    var[i] = orx_apply<OPT>(var[i], grad[i], a, b, o);
    if (OPT != ORX_OPT_SGD) s0[i] = a;
    if (OPT == ORX_OPT_ADAM_LAZY) s1[i] = b;
  }
}

extern "C" int orx_dense_apply(orx_handle_t h, float* var, float* s0, float* s1, const float* grad, int64_t n,
                               const orx_opt_t* opt, orx_stream_t s) {
  ORX_REQUIRE(h != nullptr && var && grad && opt, "null pointer");
  ORX_REQUIRE(n >= 0, "negative n");
  if (n == 0) return ORX_OK;
  ORX_CUDA(cudaSetDevice(h->device));
  const OrxOptDev o = orx_opt_to_dev(opt);
  int64_t blocks = (n + 255) / 256;
  if (blocks > (int64_t)h->num_sms * 16) blocks = (int64_t)h->num_sms * 16;
  cudaStream_t st = (cudaStream_t)s;
  switch (opt->kind) {
    case ORX_OPT_SGD: k_dense_apply<ORX_OPT_SGD><<<(int)blocks, 256, 0, st>>>(var, s0, s1, grad, n, o); break;
    case ORX_OPT_ADAGRAD:
      ORX_REQUIRE(s0, "Adagrad needs s0");
      k_dense_apply<ORX_OPT_ADAGRAD><<<(int)blocks, 256, 0, st>>>(var, s0, s1, grad, n, o);
      break;
    case ORX_OPT_ADAM_LAZY:
    case ORX_OPT_ADAM_DENSE:  // identical on a dense variable
      ORX_REQUIRE(s0 && s1, "Adam needs s0 and s1");
      k_dense_apply<ORX_OPT_ADAM_LAZY><<<(int)blocks, 256, 0, st>>>(var, s0, s1, grad, n, o);
      break;
    default: orx_set_error("unknown optimizer kind %d", opt->kind); return ORX_ERR_INVALID;
  }
  ORX_LAUNCH_CHECK();
  return ORX_OK;
}

// ---------------------------------------------------------------------------------------
// inference: scores[Bu, I]  (bpr.py:39-43, wrmf.py:36-40, ucml.py:50-53, gmf.py:36-41), "K11".
// 64 users x 64 items per block, D consumed in chunks of 16 through shared memory; 4x4 per thread.
// ---------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(256) k_score_all(const float* __restrict__ user_tab, int64_t U,
                                                   const int32_t* __restrict__ uid, int Bu,
                                                   const float* __restrict__ scale, const float* __restrict__ item_tab,
                                                   const float* __restrict__ bias, int64_t I, int D,
                                                   float* __restrict__ scores) {
  constexpr int T = 64, KC = 16;
  __shared__ float su[KC][T + 1], si[KC][T + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t i0 = (int64_t)blockIdx.x * T;
  const int u0 = blockIdx.y * T;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  for (int k0 = 0; k0 < D; k0 += KC) {
    for (int e = threadIdx.x; e < T * KC; e += 256) {
      const int r = e / KC, k = e % KC;
      float uv = 0.f, iv = 0.f;
      if (k0 + k < D) {
        if (u0 + r < Bu) {
          const int32_t id = uid[u0 + r];
          if (id >= 0 && (int64_t)id < U) {
            uv = user_tab[(int64_t)id * D + k0 + k];
            if (scale) uv *= scale[k0 + k];
          }
        }
        if (i0 + r < I) iv = item_tab[(i0 + r) * D + k0 + k];
      }
      su[k][r] = uv;
      si[k][r] = iv;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      if (k0 + k < D) {
        float uu[4], ii[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) uu[a] = su[k][ty * 4 + a];
#pragma unroll
        for (int b = 0; b < 4; ++b) ii[b] = si[k][tx + 16 * b];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            if (KIND == ORX_SCORE_DOT) acc[a][b] += uu[a] * ii[b];
            else acc[a][b] -= (uu[a] - ii[b]) * (uu[a] - ii[b]);
          }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int u = u0 + ty * 4 + a;
    if (u >= Bu) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int64_t i = i0 + tx + 16 * b;
      if (i < I) scores[(int64_t)u * I + i] = acc[a][b] + (bias ? bias[i] : 0.f);
    }
  }
}

extern "C" int orx_score_all(orx_handle_t h, int32_t kind, const float* user_tab, int64_t U, const int32_t* uid,
                             int32_t Bu, const float* scale, const float* item_tab, const float* item_bias, int64_t I,
                             int32_t dim, float* scores, orx_stream_t s) {
  ORX_REQUIRE(h != nullptr && user_tab && uid && item_tab && scores, "null pointer");
  ORX_REQUIRE(kind == ORX_SCORE_DOT || kind == ORX_SCORE_NEG_SQDIST, "unknown score kind");
  ORX_REQUIRE(U > 0 && I > 0 && dim > 0 && Bu >= 0, "bad sizes");
  if (Bu == 0) return ORX_OK;
  ORX_CUDA(cudaSetDevice(h->device));
  dim3 grid((unsigned)((I + 63) / 64), (unsigned)((Bu + 63) / 64));
  if (kind == ORX_SCORE_DOT)
    k_score_all<ORX_SCORE_DOT><<<grid, 256, 0, (cudaStream_t)s>>>(user_tab, U, uid, Bu, scale, item_tab, item_bias, I, dim, scores);
  else
    k_score_all<ORX_SCORE_NEG_SQDIST><<<grid, 256, 0, (cudaStream_t)s>>>(user_tab, U, uid, Bu, scale, item_tab, item_bias, I, dim, scores);
  ORX_LAUNCH_CHECK();
  return ORX_OK;
}

// ---------------------------------------------------------------------------------------
// ranking metrics (openrec/tf2/metrics/ranking_metrics.py:8-69), one block per user row.
//   AUC    = #{(e,p): pred_e <= pred_p, e in eval, p in pos} / (n_pos*n_eval),  eval = !(pos|excl)
//   s      = exp(pred) * !excl ;  rank_p = #{i: s_i > s_p}
//   NDCG@k = sum_p [rank_p<k] / log2(rank_p+2)   (no ideal-DCG normaliser)
//   Recall@k = #{p: rank_p<k} / n_pos
// ---------------------------------------------------------------------------------------
#define ORX_MAX_AT 8
struct RankArgs {
  const float* pred;
  const uint8_t *pos, *excl;
  int R;
  int64_t I;
  int at[ORX_MAX_AT];
  int n_at;
  float *auc, *ndcg, *recall;
};

__global__ void __launch_bounds__(256) k_rank_metrics(const RankArgs a) {
  constexpr int TILE = 1024;
  __shared__ int s_idx[TILE];
  __shared__ int s_n, s_npos, s_neval;
  __shared__ unsigned long long s_auc;
  __shared__ double s_dcg[ORX_MAX_AT];
  __shared__ int s_hit[ORX_MAX_AT];
  const int r = blockIdx.x;
  const float* pred = a.pred + (int64_t)r * a.I;
  const uint8_t* pos = a.pos + (int64_t)r * a.I;
  const uint8_t* excl = a.excl + (int64_t)r * a.I;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  if (threadIdx.x == 0) {
    s_npos = 0;
    s_neval = 0;
    s_auc = 0ull;
  }
  if (threadIdx.x < ORX_MAX_AT) {
    s_dcg[threadIdx.x] = 0.0;
    s_hit[threadIdx.x] = 0;
  }
  __syncthreads();
  int np = 0, ne = 0;
  for (int64_t i = threadIdx.x; i < a.I; i += blockDim.x) {
    np += pos[i] ? 1 : 0;
    ne += (pos[i] || excl[i]) ? 0 : 1;
  }
  atomicAdd(&s_npos, np);
  atomicAdd(&s_neval, ne);
  for (int64_t t0 = 0; t0 < a.I; t0 += TILE) {
    __syncthreads();
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    for (int64_t i = t0 + threadIdx.x; i < t0 + TILE && i < a.I; i += blockDim.x)
      if (pos[i]) s_idx[atomicAdd(&s_n, 1)] = (int)i;
    __syncthreads();
    for (int q = warp; q < s_n; q += nwarp) {
      const int p = s_idx[q];
      const float pp = pred[p];
      const float sp = expf(pp) * (excl[p] ? 0.f : 1.f);
      unsigned int c_auc = 0, c_rank = 0;
      for (int64_t i = lane; i < a.I; i += 32) {
        const float pi = pred[i];
        const bool ex = excl[i] != 0;
        if (!(pos[i] || ex) && pi <= pp) ++c_auc;
        const float si = expf(pi) * (ex ? 0.f : 1.f);
        if (si > sp) ++c_rank;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        c_auc += __shfl_xor_sync(ORX_FULL, c_auc, o);
        c_rank += __shfl_xor_sync(ORX_FULL, c_rank, o);
      }
      if (lane == 0) {
        atomicAdd(&s_auc, (unsigned long long)c_auc);
        const float ra = (float)c_rank;
        const float rec = 1.f / (logf(ra + 2.f) / logf(2.0f));
        for (int k = 0; k < a.n_at; ++k)
          if (ra < (float)a.at[k]) {
            atomicAdd(&s_dcg[k], (double)rec);
            atomicAdd(&s_hit[k], 1);
          }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && a.auc) a.auc[r] = (float)s_auc / (float)((long long)s_npos * (long long)s_neval);
  if (threadIdx.x < a.n_at) {
    if (a.ndcg) a.ndcg[(int64_t)r * a.n_at + threadIdx.x] = (float)s_dcg[threadIdx.x];
    if (a.recall) a.recall[(int64_t)r * a.n_at + threadIdx.x] = (float)s_hit[threadIdx.x] / (float)s_npos;
  }
}

extern "C" int orx_rank_metrics(orx_handle_t h, const float* pred, const uint8_t* pos, const uint8_t* excl, int32_t R,
                                int64_t I, const int32_t* at_host, int32_t n_at, float* auc, float* ndcg,
                                float* recall, orx_stream_t s) {
  ORX_REQUIRE(h != nullptr && pred && pos && excl, "null pointer");
  ORX_REQUIRE(R >= 0 && I > 0 && n_at >= 0 && n_at <= ORX_MAX_AT, "bad sizes (at most 8 cut-offs)");
  ORX_REQUIRE(n_at == 0 || at_host, "null cut-offs");
  if (R == 0) return ORX_OK;
  ORX_CUDA(cudaSetDevice(h->device));
  RankArgs a;
  a.pred = pred; a.pos = pos; a.excl = excl; a.R = R; a.I = I; a.n_at = n_at;
  for (int k = 0; k < ORX_MAX_AT; ++k) a.at[k] = k < n_at ? at_host[k] : 0;
  a.auc = auc; a.ndcg = ndcg; a.recall = recall;
  k_rank_metrics<<<R, 256, 0, (cudaStream_t)s>>>(a);
  ORX_LAUNCH_CHECK();
  return ORX_OK;
}
