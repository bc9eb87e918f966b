// orx_peer.cu -- the row-sharded BPR / UCML step over NVLink PEER MEMORY (one-sided; no NCCL in the data path).
//
// SURVEY 8(e) "B200-native fused design".  Every rank maps every other rank's shard and inbox with CUDA IPC
// (NVSwitch gives each GPU full bandwidth to every peer).  One step on rank `me`:
//   [tiny barrier]  all shards final, all inboxes consumed
//   k_owner_hist/scan/scatter : position of each of my 3B lookups inside its owner's bucket (local, no atomics
//                               per lookup on global memory); my per-owner counts are written straight into the
//                               owners' inbox_cnt[me] (peer store)
//   k_peer_step     : ONE kernel per rank: reads u, p, n rows (+ item bias) from the OWNERS' shards with direct peer
//                     loads (ld.global on mapped pointers), scores, computes the per-lookup gradient rows and pushes
//                     them, with the combined local row id, into the owners' inboxes with peer stores.  Nothing has
//                     been written to any shard yet, so every gather sees pre-step values.
//   [tiny barrier]  all pushes have landed
//   k_inbox_index / k_inbox_apply / k_sparse_tail : each owner deduplicates ALL ranks' lookups of its rows (batch
//                     hash), applies the optimizer once per unique row (rows hit once: straight from the inbox row).
// Local layout per rank: combined table emb[user rows | item rows][D] + bias[user rows | item rows] (user part 0).
#include <stdlib.h>

#include "orx_common.cuh"

struct PeerDev {
  int world, rank, D;
  int64_t U, cap;
  float* const* emb;         // [world] peer pointers: owner's combined table
  float* const* bias;        // [world] owner's bias column
  float* const* inbox_emb;   // [world] owner's inbox rows  [world(src)][cap][D]
  float* const* inbox_bias;  // [world] owner's inbox bias gradients [world][cap]
  int32_t* const* inbox_ids; // [world] owner's inbox combined local row ids [world][cap]
  int32_t* const* inbox_cnt; // [world] owner's per-source counts [world]
};

// ---------------------------------------------------------------------------------------
// CUDA IPC helpers
// ---------------------------------------------------------------------------------------
extern "C" int orx_peer_alloc(orx_handle_t h, int64_t bytes, void** dev_ptr_out, uint8_t* handle_out64) {
  ORX_REQUIRE(h != nullptr && dev_ptr_out && handle_out64 && bytes > 0, "bad arguments");
  ORX_CUDA(cudaSetDevice(h->device));
  void* p = nullptr;
  ORX_CUDA(cudaMalloc(&p, (size_t)bytes));
  ORX_CUDA(cudaMemset(p, 0, (size_t)bytes));
  cudaIpcMemHandle_t hd;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  ORX_CUDA(cudaIpcGetMemHandle(&hd, p));
  memcpy(handle_out64, &hd, 64);
  *dev_ptr_out = p;
  return ORX_OK;
}

extern "C" int orx_peer_open(orx_handle_t h, const uint8_t* handle64, void** dev_ptr_out) {
  ORX_REQUIRE(h != nullptr && dev_ptr_out && handle64, "bad arguments");
  ORX_CUDA(cudaSetDevice(h->device));
  cudaIpcMemHandle_t hd;
  memcpy(&hd, handle64, 64);
  void* p = nullptr;
  ORX_CUDA(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
  *dev_ptr_out = p;
  return ORX_OK;
}

extern "C" int orx_peer_close(orx_handle_t h, void* dev_ptr) {
  ORX_REQUIRE(h != nullptr, "null handle");
  if (dev_ptr) ORX_CUDA(cudaIpcCloseMemHandle(dev_ptr));
  return ORX_OK;
}

extern "C" int orx_peer_free(orx_handle_t h, void* dev_ptr) {
  ORX_REQUIRE(h != nullptr, "null handle");
  if (dev_ptr) ORX_CUDA(cudaFree(dev_ptr));
  return ORX_OK;
}

// ---------------------------------------------------------------------------------------
// bucket positions (shared-memory aggregated) + publish my counts to the owners
// ---------------------------------------------------------------------------------------
__global__ void k_peer_hist(const int32_t* __restrict__ uid, const int32_t* __restrict__ pid,
                            const int32_t* __restrict__ nid, int B, int R, int32_t* counts) {
  extern __shared__ int32_t sh[];
  for (int r = threadIdx.x; r < R; r += blockDim.x) sh[r] = 0;
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 3 * B; i += gridDim.x * blockDim.x) {
    const int32_t id = i < B ? uid[i] : (i < 2 * B ? pid[i - B] : nid[i - 2 * B]);
    if (id >= 0) atomicAdd(&sh[id % R], 1);
  }
  __syncthreads();
  for (int r = threadIdx.x; r < R; r += blockDim.x)
    if (sh[r]) atomicAdd(counts + r, sh[r]);
}

// cursor[r] = 0 (positions are relative to the owner's bucket), publish counts[r] to owner r's inbox_cnt[me]
__global__ void k_peer_publish(const int32_t* counts, int32_t* cursor, PeerDev pd) {
  const int r = threadIdx.x;
  if (r < pd.world) {
    cursor[r] = 0;
    pd.inbox_cnt[r][pd.rank] = counts[r];
  }
  __threadfence_system();
}

__global__ void __launch_bounds__(256) k_peer_positions(const int32_t* __restrict__ uid, const int32_t* __restrict__ pid,
                                                        const int32_t* __restrict__ nid, int B, int R, int32_t* cursor,
                                                        int32_t* __restrict__ pos) {
  extern __shared__ int32_t sh[];   // [R] counts, [R] bases
  int32_t* cnt = sh;
  int32_t* base = sh + R;
  for (int r = threadIdx.x; r < R; r += blockDim.x) cnt[r] = 0;
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int r = 0, rk = 0;
  bool ok = false;
  if (i < 3 * B) {
    const int32_t id = i < B ? uid[i] : (i < 2 * B ? pid[i - B] : nid[i - 2 * B]);
    ok = id >= 0;
    if (ok) {
      r = id % R;
      rk = atomicAdd(&cnt[r], 1);
    }
  }
  __syncthreads();
  for (int q = threadIdx.x; q < R; q += blockDim.x) base[q] = cnt[q] ? atomicAdd(cursor + q, cnt[q]) : 0;
  __syncthreads();
  if (i < 3 * B) pos[i] = ok ? base[r] + rk : -1;
}

// ---------------------------------------------------------------------------------------
// the fused peer kernel: gather over NVLink -> score -> gradient rows -> push over NVLink
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float pdot4(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
__device__ __forceinline__ float psqd4(float4 a, float4 b) {
  const float x = a.x - b.x, y = a.y - b.y, z = a.z - b.z, w = a.w - b.w;
  return x * x + y * y + z * z + w * w;
}

template <int KIND, int K>   // K float4 per lane per row: D = 128*K
__global__ void __launch_bounds__(256) k_peer_step(PeerDev pd, const int32_t* __restrict__ uid,
                                                   const int32_t* __restrict__ pid, const int32_t* __restrict__ nid,
                                                   const int32_t* __restrict__ pos, int B, int64_t total_items,
                                                   float margin, float c_loss, float c_l2, float inv_B,
                                                   float* partials) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  const int R = pd.world, D = pd.D;
  float loss_acc = 0.f, l2_acc = 0.f;
  for (int t = warp; t < B; t += nwarps) {
    const int32_t u = uid[t], p = pid[t], n = nid[t];
    const bool ok = u >= 0 && (int64_t)u < pd.U && p >= 0 && (int64_t)p < total_items && n >= 0 && (int64_t)n < total_items;
    if (!ok) continue;   // warp-uniform
    const int ou = u % R, op = p % R, on = n % R;
    // combined local rows on the owners: users first, then items
    const int64_t lu = u / R;
    const int64_t lp = p / R + (pd.U - op + R - 1) / R;
    const int64_t ln = n / R + (pd.U - on + R - 1) / R;
    const float* ru = pd.emb[ou] + lu * D;
    const float* rp = pd.emb[op] + lp * D;
    const float* rn = pd.emb[on] + ln * D;
    float4 uv[K], pv[K], nv[K];
#pragma unroll
    // peer loads over NVLink (or local HBM when the owner is this rank); non-coherent path, the shards are read-only
    // during the push.  NOTE (profiles/r1k_p2p_probe.txt): random 512 B rows from a peer mapping run at ~600 GB/s while
    // the mapped shard is <= 2 GB but collapse to ~35 GB/s at 6.6 GB (translation reach of peer mappings); bulk peer
    // copies are unaffected.  Large shards therefore want an owner-side gather into a compact outbox first.
    for (int k = 0; k < K; ++k) {
      uv[k] = __ldg(reinterpret_cast<const float4*>(ru + (k * 32 + lane) * 4));
      pv[k] = __ldg(reinterpret_cast<const float4*>(rp + (k * 32 + lane) * 4));
      nv[k] = __ldg(reinterpret_cast<const float4*>(rn + (k * 32 + lane) * 4));
    }
    const float bp = __ldg(pd.bias[op] + lp), bn = __ldg(pd.bias[on] + ln);
    float s1 = 0.f, s2 = 0.f, sq = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      if (KIND == ORX_PAIR_BPR) {
        s1 += pdot4(uv[k], pv[k]);
        s2 += pdot4(uv[k], nv[k]);
      } else {
        s1 += psqd4(uv[k], pv[k]);
        s2 += psqd4(uv[k], nv[k]);
      }
      sq += pdot4(uv[k], uv[k]) + pdot4(pv[k], pv[k]) + pdot4(nv[k], nv[k]);
    }
    l2_acc += sq;
    s1 = orx_group_sum<32>(s1);
    s2 = orx_group_sum<32>(s2);
    float lt, g;
    if (KIND == ORX_PAIR_BPR) {
      const float x = (s1 + bp) - (s2 + bn);
      const float y = fmaxf(x, -30.f);
      float ls, sn;
      orx_logsig(y, &ls, &sn);
      lt = -ls;
      g = (x >= -30.f) ? -(c_loss * inv_B) * sn : 0.f;
    } else {
      const float hgn = margin - (((-s1) + bp) - ((-s2) + bn));
      lt = fmaxf(hgn, 0.f);
      g = (hgn >= 0.f) ? c_loss : 0.f;
    }
    if (lane == 0) loss_acc += lt;
    // push the three gradient rows into the owners' inboxes (segment of source `rank`)
    const int64_t qu = (int64_t)pd.rank * pd.cap + pos[t];
    const int64_t qp = (int64_t)pd.rank * pd.cap + pos[B + t];
    const int64_t qn = (int64_t)pd.rank * pd.cap + pos[2 * B + t];
    float* du = pd.inbox_emb[ou] + qu * D;
    float* dp = pd.inbox_emb[op] + qp * D;
    float* dn = pd.inbox_emb[on] + qn * D;
    const float t2 = 2.f * g, c2 = c_l2;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      float4 gu, gp, gn;
      const float4 a = uv[k], b = pv[k], c = nv[k];
      if (KIND == ORX_PAIR_BPR) {
        gu = make_float4(g * (b.x - c.x) + c2 * a.x, g * (b.y - c.y) + c2 * a.y, g * (b.z - c.z) + c2 * a.z, g * (b.w - c.w) + c2 * a.w);
        gp = make_float4(g * a.x + c2 * b.x, g * a.y + c2 * b.y, g * a.z + c2 * b.z, g * a.w + c2 * b.w);
        gn = make_float4(-g * a.x + c2 * c.x, -g * a.y + c2 * c.y, -g * a.z + c2 * c.z, -g * a.w + c2 * c.w);
      } else {
        gu = make_float4(t2 * (c.x - b.x) + c2 * a.x, t2 * (c.y - b.y) + c2 * a.y, t2 * (c.z - b.z) + c2 * a.z, t2 * (c.w - b.w) + c2 * a.w);
        gp = make_float4(t2 * (b.x - a.x) + c2 * b.x, t2 * (b.y - a.y) + c2 * b.y, t2 * (b.z - a.z) + c2 * b.z, t2 * (b.w - a.w) + c2 * b.w);
        gn = make_float4(t2 * (a.x - c.x) + c2 * c.x, t2 * (a.y - c.y) + c2 * c.y, t2 * (a.z - c.z) + c2 * c.z, t2 * (a.w - c.w) + c2 * c.w);
      }
      *reinterpret_cast<float4*>(du + (k * 32 + lane) * 4) = gu;
      *reinterpret_cast<float4*>(dp + (k * 32 + lane) * 4) = gp;
      *reinterpret_cast<float4*>(dn + (k * 32 + lane) * 4) = gn;
    }
    if (lane == 0) {
      const float gbias = (KIND == ORX_PAIR_BPR) ? g : -g;
      pd.inbox_ids[ou][qu] = (int32_t)lu;
      pd.inbox_ids[op][qp] = (int32_t)lp;
      pd.inbox_ids[on][qn] = (int32_t)ln;
      pd.inbox_bias[ou][qu] = 0.f;
      pd.inbox_bias[op][qp] = gbias;
      pd.inbox_bias[on][qn] = -gbias;
    }
  }
  // (no per-thread system fence here: 300k MEMBAR.SYS cost 1.3 ms in profile r1k; the peer stores are complete when
  //  the grid completes, and the barrier that follows is stream-ordered after this kernel)
  __shared__ float sred[8][2];
  loss_acc = orx_group_sum<32>(loss_acc);
  l2_acc = orx_group_sum<32>(l2_acc);
  if (lane == 0) {
    sred[threadIdx.x >> 5][0] = loss_acc;
    sred[threadIdx.x >> 5][1] = l2_acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float l = 0.f, q = 0.f;
    for (int w = 0; w < 8; ++w) {
      l += sred[w][0];
      q += sred[w][1];
    }
    partials[2 * blockIdx.x] = l;
    partials[2 * blockIdx.x + 1] = q;
  }
}

struct PeerHost {   // mirrors the Python-side ctypes struct orx_peer_t
  int32_t world, rank, dim, _pad;
  int64_t total_users, total_items, cap;
  void *emb, *bias, *inbox_emb, *inbox_bias, *inbox_ids, *inbox_cnt;   // DEVICE arrays of `world` pointers each
};

static PeerDev to_dev(const PeerHost* p) {
  PeerDev d;
  d.world = p->world; d.rank = p->rank; d.D = p->dim; d.U = p->total_users; d.cap = p->cap;
  d.emb = (float* const*)p->emb; d.bias = (float* const*)p->bias;
  d.inbox_emb = (float* const*)p->inbox_emb; d.inbox_bias = (float* const*)p->inbox_bias;
  d.inbox_ids = (int32_t* const*)p->inbox_ids; d.inbox_cnt = (int32_t* const*)p->inbox_cnt;
  return d;
}

extern "C" int orx_peer_pairwise_push(orx_handle_t h, int32_t kind, const void* peer_host, const int32_t* uid,
                                      const int32_t* pid, const int32_t* nid, int32_t B, int32_t* pos_scratch,
                                      float margin, float c_loss, float c_l2, float inv_B, float* out4,
                                      orx_stream_t s) {
  ORX_REQUIRE(h != nullptr && peer_host && uid && pid && nid && pos_scratch && out4, "null pointer");
  ORX_REQUIRE(kind == ORX_PAIR_BPR || kind == ORX_PAIR_UCML, "unknown pairwise kind");
  const PeerHost* ph = (const PeerHost*)peer_host;
  ORX_REQUIRE(B > 0 && ph->world >= 1 && ph->world <= 64 && (ph->dim == 128 || ph->dim == 256),
              "peer step supports dim 128 / 256 and world <= 64");
  ORX_REQUIRE(3 * (int64_t)B <= ph->cap, "inbox capacity must be >= 3*B");
  ORX_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)s;
  const PeerDev pd = to_dev(ph);
  const int R = ph->world;
  if (!h->bucket_cursor) ORX_CUDA(cudaMalloc(&h->bucket_cursor, sizeof(int32_t) * 1024));
  int32_t* counts = h->bucket_cursor + 512;
  int32_t* cursor = h->bucket_cursor;
  static int dbg = -1;
  static cudaEvent_t ev[6];
  static int dbg_left = 0;
  if (dbg < 0) {
    dbg = getenv("ORX_PEER_DBG") ? 1 : 0;
    if (dbg) {
      for (int i = 0; i < 6; ++i) cudaEventCreate(&ev[i]);
      dbg_left = 30;
    }
  }
  const bool timing = dbg && dbg_left > 0;
#define ORX_EV(i) do { if (timing) cudaEventRecord(ev[i], st); } while (0)
  ORX_EV(0);
  ORX_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * R, st));
  int hb = (3 * B + 255) / 256;
  if (hb > h->num_sms * 4) hb = h->num_sms * 4;
  k_peer_hist<<<hb, 256, sizeof(int32_t) * R, st>>>(uid, pid, nid, B, R, counts);
  ORX_LAUNCH_CHECK();
  ORX_EV(1);
  k_peer_publish<<<1, 64, 0, st>>>(counts, cursor, pd);
  ORX_LAUNCH_CHECK();
  ORX_EV(2);
  k_peer_positions<<<(3 * B + 255) / 256, 256, 2 * sizeof(int32_t) * R, st>>>(uid, pid, nid, B, R, cursor, pos_scratch);
  ORX_LAUNCH_CHECK();
  ORX_EV(3);
  const int blocks = h->num_sms * 8;
  int rc = orx_ensure_partials(h, blocks, st);
  if (rc) return rc;
#define ORX_PEER(KIND, KK) k_peer_step<KIND, KK><<<blocks, 256, 0, st>>>(pd, uid, pid, nid, pos_scratch, B, ph->total_items, margin, c_loss, c_l2, inv_B, h->partials)
  if (kind == ORX_PAIR_BPR) { if (ph->dim == 128) ORX_PEER(ORX_PAIR_BPR, 1); else ORX_PEER(ORX_PAIR_BPR, 2); }
  else { if (ph->dim == 128) ORX_PEER(ORX_PAIR_UCML, 1); else ORX_PEER(ORX_PAIR_UCML, 2); }
#undef ORX_PEER
  ORX_LAUNCH_CHECK();
  ORX_EV(4);
  rc = orx_launch_reduce_partials(h->partials, blocks, kind == ORX_PAIR_BPR ? inv_B : 1.f, out4, st);
  ORX_EV(5);
  if (timing) {
    cudaEventSynchronize(ev[5]);
    float ms[5];
    for (int i = 0; i < 5; ++i) cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]);
    if (--dbg_left < 4 && ph->rank == 0)
      fprintf(stderr, "[orx peer dbg] hist %.1f us | publish %.1f | positions %.1f | peer_step %.1f | reduce %.1f\n",
              ms[0] * 1e3, ms[1] * 1e3, ms[2] * 1e3, ms[3] * 1e3, ms[4] * 1e3);
  }
#undef ORX_EV
  return rc;
}

// ---------------------------------------------------------------------------------------
// owner side: dedup over every source's inbox segment, optimizer once per unique row
// ---------------------------------------------------------------------------------------
// flat lookup index q in [0, sum cnt) -> (src, pos)
__device__ __forceinline__ bool inbox_locate(const int32_t* __restrict__ cnt, int R, int64_t q, int* src, int* pos) {
  int64_t acc = 0;
  for (int r = 0; r < R; ++r) {
    const int c = cnt[r];
    if (q < acc + c) {
      *src = r;
      *pos = (int)(q - acc);
      return true;
    }
    acc += c;
  }
  return false;
}

__global__ void __launch_bounds__(256) k_inbox_index(OrxHash hi, const int32_t* __restrict__ ids,
                                                     const int32_t* __restrict__ cnt, int R, int64_t cap, int64_t rows,
                                                     int32_t* bad) {
  const int64_t nth = (int64_t)gridDim.x * blockDim.x;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;; q += nth) {
    int src, pos;
    if (!inbox_locate(cnt, R, q, &src, &pos)) break;
    const int32_t id = ids[(int64_t)src * cap + pos];
    if (id >= 0 && (int64_t)id < rows) orx_hash_insert(hi, id, 0);
    else atomicAdd(bad, 1);
  }
}

template <int OPT>
__global__ void __launch_bounds__(256) k_inbox_apply(float* emb, float* s0, float* s1, float* bias, float* bs0,
                                                     float* bs1, int64_t rows, int D, const int32_t* __restrict__ ids,
                                                     const float* __restrict__ vals, const float* __restrict__ bvals,
                                                     const int32_t* __restrict__ cnt, int R, int64_t cap, OrxHash hi,
                                                     float* gstage, float* gbstage, OrxOptDev o) {
  constexpr bool S0 = (OPT == ORX_OPT_ADAGRAD || OPT == ORX_OPT_ADAM_LAZY);
  constexpr bool S1 = (OPT == ORX_OPT_ADAM_LAZY);
  const int lane = threadIdx.x & 31;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;; q += nw) {
    int src, pos;
    if (!inbox_locate(cnt, R, q, &src, &pos)) break;   // warp-uniform
    const int64_t slot = (int64_t)src * cap + pos;
    const int32_t id = ids[slot];
    if (id < 0 || (int64_t)id >= rows) continue;
    int d = -1;
    uint32_t c = 0;
    if (lane == 0) c = orx_hash_find(hi, id, &d);
    c = __shfl_sync(ORX_FULL, c, 0);
    d = __shfl_sync(ORX_FULL, d, 0);
    const float* v = vals + slot * D;
    if (c == 1u) {   // the only lookup of this row in the whole global batch: apply straight from the inbox
      for (int e = lane * 4; e < D; e += 128) {
        const int64_t off = (int64_t)id * D + e;
        const float4 g = __ldcg(reinterpret_cast<const float4*>(v + e));
        float4 w = __ldcg(reinterpret_cast<const float4*>(emb + off));
        float4 a = S0 ? __ldcg(reinterpret_cast<const float4*>(s0 + off)) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 b = S1 ? __ldcg(reinterpret_cast<const float4*>(s1 + off)) : make_float4(0.f, 0.f, 0.f, 0.f);
        __stcg(reinterpret_cast<float4*>(emb + off), orx_apply4<OPT>(w, g, a, b, o));
        if (S0) __stcg(reinterpret_cast<float4*>(s0 + off), a);
        if (S1) __stcg(reinterpret_cast<float4*>(s1 + off), b);
      }
      if (lane == 0) {
        float a = S0 ? bs0[id] : 0.f, b = S1 ? bs1[id] : 0.f;
        bias[id] = orx_apply<OPT>(bias[id], bvals[slot], a, b, o);
        if (S0) bs0[id] = a;
        if (S1) bs1[id] = b;
      }
    } else {
      for (int e = lane * 4; e < D; e += 128)
        orx_red4(gstage + (int64_t)d * D + e, __ldcg(reinterpret_cast<const float4*>(v + e)));
      if (lane == 0) atomicAdd(gbstage + d, bvals[slot]);
    }
  }
}

// emb / bias: this rank's combined shard (+ optimizer slots); inbox_*: this rank's inbox (LOCAL pointers).
extern "C" int orx_peer_apply(orx_handle_t h, const orx_table_t* emb, const orx_table_t* bias, const int32_t* inbox_ids,
                              const float* inbox_emb, const float* inbox_bias, const int32_t* inbox_cnt, int32_t world,
                              int64_t cap, const orx_opt_t* opt, orx_stream_t s) {
  ORX_REQUIRE(h != nullptr && emb && bias && emb->var && bias->var && inbox_ids && inbox_emb && inbox_bias && inbox_cnt && opt,
              "null pointer");
  ORX_REQUIRE(opt->kind == ORX_OPT_SGD || opt->kind == ORX_OPT_ADAGRAD || opt->kind == ORX_OPT_ADAM_LAZY,
              "peer apply supports SGD / Adagrad / lazy Adam");
  ORX_REQUIRE((emb->dim & 3) == 0 && bias->rows == emb->rows && world >= 1 && cap > 0, "bad shapes");
  if (opt->kind != ORX_OPT_SGD) ORX_REQUIRE(emb->s0 && bias->s0, "optimizer slot s0 missing");
  if (opt->kind == ORX_OPT_ADAM_LAZY) ORX_REQUIRE(emb->s1 && bias->s1, "optimizer slot s1 missing");
  ORX_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)s;
  const int D = emb->dim;
  // worst case every lookup of the global batch lands here: size the item-side hash / staging for world*cap/2 lookups
  int rc = orx_ensure_workspace(h, (int64_t)world * cap / 2 + 1, D, false);
  if (rc) return rc;
  orx_new_epoch(h);
  const OrxOptDev o = orx_opt_to_dev(opt);
  const int grid = h->num_sms * 8;
  k_inbox_index<<<grid, 256, 0, st>>>(h->hi, inbox_ids, inbox_cnt, world, cap, emb->rows, h->counters + 3);
  ORX_LAUNCH_CHECK();
#define ORX_IA(OPT) k_inbox_apply<OPT><<<grid, 256, 0, st>>>(emb->var, emb->s0, emb->s1, bias->var, bias->s0, bias->s1, emb->rows, D, inbox_ids, inbox_emb, inbox_bias, inbox_cnt, world, cap, h->hi, h->gi, h->gb, o)
  if (opt->kind == ORX_OPT_SGD) ORX_IA(ORX_OPT_SGD);
  else if (opt->kind == ORX_OPT_ADAGRAD) ORX_IA(ORX_OPT_ADAGRAD);
  else ORX_IA(ORX_OPT_ADAM_LAZY);
#undef ORX_IA
  ORX_LAUNCH_CHECK();
  // staged rows: the shared tail with the combined table in the "item" role (no user-side rows)
  TailArgs ta;
  ta.U = emb->var; ta.Us0 = emb->s0; ta.Us1 = emb->s1;
  ta.I = emb->var; ta.Is0 = emb->s0; ta.Is1 = emb->s1;
  ta.Bv = bias->var; ta.Bs0 = bias->s0; ta.Bs1 = bias->s1;
  ta.D = D; ta.opt = o; ta.hu = h->hu; ta.hi = h->hi;
  ta.gu = h->gu; ta.gi = h->gi; ta.gb = h->gb;
  ta.partials = h->partials; ta.n_partials = 0; ta.loss_scale = 0.f;
  ta.counters = h->counters; ta.out4 = h->out_stage[0];
  ta.W = ta.Ws0 = ta.Ws1 = ta.gw = nullptr; ta.c_l2 = 0.f;
  return orx_launch_tail(h, ta, opt->kind, st);
}
