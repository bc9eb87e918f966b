// orx_xchg.cu -- "mailbox" exchange of the row-sharded BPR / UCML step: the three variable-size
// all-to-alls of the step (ids -> owners, rows -> requesters, gradient rows -> owners) done by liborx kernels with
// PEER STORES into small per-rank mailboxes mapped through CUDA IPC -- no NCCL call in the data path, no count ever
// visits the host, so the launching thread runs ahead of the GPU for the whole step.
//
// Why stores into mailboxes and not loads from the shards: random 512 B peer LOADS from a multi-GB mapped shard fall
// off a cliff (35 GB/s at 6.6 GB, profiles/r1k_p2p_probe.txt); every access to a table here is local, and everything
// that crosses NVLink is a contiguous 528 B row written into a region of a few hundred MB.
//
// One step on rank `me` (R ranks, B triplets per rank, W = D + 4 floats per combined row):
//   orx_owner_bucket_combined : counts[R], send_local[3B] (lookups sorted by owner), slot[3B]      (orx_sharded.cu)
//   orx_xchg_push_ids         : bucket o of send_local -> owner o's idbox[me][*]; (count, my bucket offset) -> its meta[me]
//   -- barrier A --
//   orx_xchg_gather_push      : owner: prefix over meta counts = where each source's segment starts in my gradient
//                               inbox (published to the source: its meta[me].base); for every requested id: read the
//                               LOCAL table row, store it into the requester's `got` at the requester's own position;
//                               compact the ids into req[] and their number into *n_dev
//   -- barrier B --
//   orx_xchg_grad_push        : requester: score + per-lookup gradient rows from `got` (all pre-step values), each row
//                               stored straight into its owner's gradient inbox `gin` at base + index-in-bucket
//   -- barrier C --
//   orx_sparse_apply_devn     : owner: dedup ALL ranks' lookups (req[0..*n_dev)), apply the optimizer once per row
// Barriers: orx_xchg_barrier (flags in peer memory, release/acquire at system scope, bounded spin) or any stream-ordered
// collective of the caller.  Buffer reuse across steps is ordered by the same three barriers (see sharded.py).
#include "orx_common.cuh"
#include "orx_pair.cuh"

struct XchgDev {
  int world, rank, W, cap;        // cap: idbox entries per source rank (>= 3B)
  int32_t* const* idbox;          // [world] owner's int32 idbox[world][cap]
  int32_t* const* meta;           // [world] peer's  int32 meta[world][8]: {count, sender's bucket offset, base in my gin, -, loss, l2 (float bits), -, -}
  float* const* got;              // [world] requester's float got[cap][W]
  float* const* gin;              // [world] owner's float gin[gin_rows][W]
  int32_t* const* flags;          // [world] peer's int32 flags[world + 1]  (last entry: sticky error)
};

struct XchgHost {   // mirrors orx_xchg_t in include/orx.h
  int32_t world, rank, width, cap;
  void *idbox, *meta, *got, *gin, *flags;
};

static XchgDev to_dev(const XchgHost* x) {
  XchgDev d;
  d.world = x->world; d.rank = x->rank; d.W = x->width; d.cap = x->cap;
  d.idbox = (int32_t* const*)x->idbox;
  d.meta = (int32_t* const*)x->meta;
  d.got = (float* const*)x->got;
  d.gin = (float* const*)x->gin;
  d.flags = (int32_t* const*)x->flags;
  return d;
}

#define XCHG_MAX_R 64
#define XCHG_META 8   // int32 words per peer in a meta mailbox

// first r with off[r+1] > p   (off[0] = 0 <= p < off[R])
__device__ __forceinline__ int xchg_bucket_of(const int32_t* off, int R, int p) {
  int lo = 0, hi = R - 1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (off[mid + 1] > p) hi = mid; else lo = mid + 1;
  }
  return lo;
}

// ---------------------------------------------------------------------------------------
// ids -> owners
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_xchg_push_ids(XchgDev x, const int32_t* __restrict__ counts,
                                                       const int32_t* __restrict__ send_local, int n) {
  __shared__ int32_t off[XCHG_MAX_R + 1];
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int r = 0; r < x.world; ++r) { off[r] = acc; acc += counts[r]; }
    off[x.world] = acc;
  }
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x < x.world) {
    const int o = threadIdx.x;
    int32_t* m = x.meta[o] + XCHG_META * x.rank;
    m[0] = counts[o];
    m[1] = off[o];
  }
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
    const int o = xchg_bucket_of(off, x.world, p);
    x.idbox[o][(int64_t)x.rank * x.cap + (p - off[o])] = send_local[p];
  }
}

// ---------------------------------------------------------------------------------------
// fused form of (owner bucketing + ids -> owners): a stable two-pass counting sort over 1024-lookup chunks.
//   k_xchg_hist         : chunk c's per-owner histogram -> bc[c][R]                       (no global atomics)
//   k_xchg_scatter_push : chunk c sums bc[*][r] (totals) and bc[<c][r] (its base), then stores each lookup's combined
//                         local row id straight into the owner's idbox (peer store) and its position into slot[]
// Lookup i: i < B user, < 2B positive item, else negative item; items' local rows are offset by the owner's user rows.
// ---------------------------------------------------------------------------------------
#define XCHG_CHUNK 1024

__device__ __forceinline__ int32_t xchg_lookup(const int32_t* __restrict__ uid, const int32_t* __restrict__ pid,
                                               const int32_t* __restrict__ nid, int B, int i) {
  return i < B ? uid[i] : (i < 2 * B ? pid[i - B] : nid[i - 2 * B]);
}

__global__ void __launch_bounds__(256) k_xchg_hist(const int32_t* __restrict__ uid, const int32_t* __restrict__ pid,
                                                   const int32_t* __restrict__ nid, int B, int R,
                                                   int32_t* __restrict__ bc) {
  __shared__ int32_t cnt[XCHG_MAX_R];
  if (threadIdx.x < R) cnt[threadIdx.x] = 0;
  __syncthreads();
  const int n = 3 * B, i0 = blockIdx.x * XCHG_CHUNK;
  for (int k = threadIdx.x; k < XCHG_CHUNK && i0 + k < n; k += 256) {
    const int32_t id = xchg_lookup(uid, pid, nid, B, i0 + k);
    atomicAdd(&cnt[id >= 0 ? id % R : 0], 1);
  }
  __syncthreads();
  if (threadIdx.x < R) bc[blockIdx.x * R + threadIdx.x] = cnt[threadIdx.x];
}

__global__ void __launch_bounds__(256) k_xchg_scatter_push(XchgDev x, const int32_t* __restrict__ uid,
                                                           const int32_t* __restrict__ pid,
                                                           const int32_t* __restrict__ nid, int B, int64_t U,
                                                           const int32_t* __restrict__ bc, int nchunks,
                                                           int32_t* __restrict__ counts_out,
                                                           int32_t* __restrict__ slot) {
  __shared__ int32_t tot[XCHG_MAX_R], before[XCHG_MAX_R], off[XCHG_MAX_R + 1], cur[XCHG_MAX_R];
  const int R = x.world;
  if (threadIdx.x < R) { tot[threadIdx.x] = 0; before[threadIdx.x] = 0; cur[threadIdx.x] = 0; }
  __syncthreads();
  for (int e = threadIdx.x; e < nchunks * R; e += 256) {
    const int c = e / R, r = e - c * R;
    const int v = bc[e];
    if (v) {
      atomicAdd(&tot[r], v);
      if (c < (int)blockIdx.x) atomicAdd(&before[r], v);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int r = 0; r < R; ++r) { off[r] = acc; acc += tot[r]; }
    off[R] = acc;
  }
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x < R) {
    const int o = threadIdx.x;
    counts_out[o] = tot[o];
    int32_t* m = x.meta[o] + XCHG_META * x.rank;
    m[0] = tot[o];
    m[1] = off[o];
  }
  const int n = 3 * B, i0 = blockIdx.x * XCHG_CHUNK;
  for (int k = threadIdx.x; k < XCHG_CHUNK && i0 + k < n; k += 256) {
    const int i = i0 + k;
    const int32_t id = xchg_lookup(uid, pid, nid, B, i);
    const int r = id >= 0 ? id % R : 0;
    const int idx = before[r] + atomicAdd(&cur[r], 1);             // index inside my bucket for owner r
    const int32_t user_rows_on_r = (i >= B) ? (int32_t)((U - r + R - 1) / R) : 0;
    x.idbox[r][(int64_t)x.rank * x.cap + idx] = id >= 0 ? id / R + user_rows_on_r : -1;
    slot[i] = off[r] + idx;
  }
}

// ---------------------------------------------------------------------------------------
// owner: requested rows -> requesters' `got`; compact ids; gradient-inbox bases -> requesters
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_xchg_gather_push(XchgDev x, const float* __restrict__ table, int64_t rows,
                                                          int gin_rows, int32_t* __restrict__ req,
                                                          int32_t* __restrict__ n_dev, int32_t* n_bad) {
  __shared__ int32_t roff[XCHG_MAX_R + 1], soff[XCHG_MAX_R];
  __shared__ int overflow;
  const int R = x.world, me = x.rank, W = x.W;
  if (threadIdx.x == 0) {
    const int32_t* m = x.meta[me];
    int acc = 0;
    for (int s = 0; s < R; ++s) {
      int c = m[XCHG_META * s];
      c = c < 0 ? 0 : (c > x.cap ? x.cap : c);
      roff[s] = acc;
      soff[s] = m[XCHG_META * s + 1];
      acc += c;
    }
    roff[R] = acc;
    overflow = acc > gin_rows;
  }
  __syncthreads();
  if (overflow) {   // gradient inbox too small for this batch: the step becomes a no-op everywhere, error is sticky
    if (blockIdx.x == 0 && threadIdx.x < R) x.meta[threadIdx.x][XCHG_META * me + 2] = -1;
    if (blockIdx.x == 0 && threadIdx.x == 0) { *n_dev = 0; x.flags[me][R] = 2; }
    return;
  }
  if (blockIdx.x == 0 && threadIdx.x < R) x.meta[threadIdx.x][XCHG_META * me + 2] = roff[threadIdx.x];
  if (blockIdx.x == 0 && threadIdx.x == 0) *n_dev = roff[R];
  const int total = roff[R];
  const int lane = threadIdx.x & 31;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  const int32_t* box = x.idbox[me];
  const int nq = W / 4;   // float4 per row (33 at D = 128): lane e takes float4 e, lanes < nq - 32 also float4 32 + e
  // a warp moves 8 consecutive rows per iteration: lanes 0..7 resolve (source rank, id, destination), then all 8 rows'
  // loads are issued before the first (peer) store -- one dependent chain per row would leave the kernel latency-bound
  for (int j0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 8; j0 < total; j0 += nw * 8) {
    int32_t my_id = -1;
    const float* my_src = table;
    float* my_dst = nullptr;
    if (lane < 8 && j0 + lane < total) {
      const int j = j0 + lane;
      const int s = xchg_bucket_of(roff, R, j);
      const int idx = j - roff[s];
      const int32_t id = __ldcg(box + (int64_t)s * x.cap + idx);
      const bool ok = id >= 0 && (int64_t)id < rows;
      req[j] = ok ? id : -1;
      if (!ok && n_bad) atomicAdd(n_bad, 1);
      my_id = ok ? id : -1;
      my_src = table + (int64_t)(ok ? id : 0) * W;
      my_dst = x.got[s] + ((int64_t)soff[s] + idx) * W;
    }
    if (nq <= 64) {
      float4 v0[8], v1[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int32_t id = __shfl_sync(ORX_FULL, my_id, k);
        const float4* src = reinterpret_cast<const float4*>(__shfl_sync(ORX_FULL, (unsigned long long)my_src, k));
        v0[k] = v1[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (id >= 0) {
          if (lane < nq) v0[k] = __ldcg(src + lane);
          if (lane + 32 < nq) v1[k] = __ldcg(src + 32 + lane);
        }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float4* dst = reinterpret_cast<float4*>(__shfl_sync(ORX_FULL, (unsigned long long)my_dst, k));
        if (dst) {
          if (lane < nq) dst[lane] = v0[k];
          if (lane + 32 < nq) dst[32 + lane] = v1[k];
        }
      }
    } else {
      for (int k = 0; k < 8; ++k) {
        const int32_t id = __shfl_sync(ORX_FULL, my_id, k);
        const float4* src = reinterpret_cast<const float4*>(__shfl_sync(ORX_FULL, (unsigned long long)my_src, k));
        float4* dst = reinterpret_cast<float4*>(__shfl_sync(ORX_FULL, (unsigned long long)my_dst, k));
        if (!dst) continue;
        for (int e = lane; e < nq; e += 32) dst[e] = id >= 0 ? __ldcg(src + e) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// requester: score, gradient rows -> owners' gradient inboxes
// ---------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(256) k_xchg_grad_push(XchgDev x, const int32_t* __restrict__ counts,
                                                        const int32_t* __restrict__ slot, int B, int D, float margin,
                                                        float c_loss, float c_l2, float inv_B, float* partials) {
  __shared__ int32_t off[XCHG_MAX_R + 1], base[XCHG_MAX_R];
  const int R = x.world, me = x.rank, W = x.W;
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int r = 0; r < R; ++r) { off[r] = acc; acc += counts[r]; }
    off[R] = acc;
  }
  if (threadIdx.x < R) base[threadIdx.x] = x.meta[me][XCHG_META * threadIdx.x + 2];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const float* got = x.got[me];
  PairArgs sa;
  sa.margin = margin; sa.c_loss = c_loss; sa.inv_B = inv_B;
  float loss_acc = 0.f, l2_acc = 0.f;
  for (int j = 0; j < 8; ++j) {
    const int t = warp * 8 + j;
    if (t >= B) break;
    const int pu = slot[t], pp = slot[B + t], pn = slot[2 * B + t];
    const float* ur = got + (int64_t)pu * W;
    const float* pr = got + (int64_t)pp * W;
    const float* nr = got + (int64_t)pn * W;
    float s1 = 0.f, s2 = 0.f, sq = 0.f;
    for (int e = lane * 4; e < D; e += 128) {
      const float4 u = *reinterpret_cast<const float4*>(ur + e), p = *reinterpret_cast<const float4*>(pr + e),
                   n = *reinterpret_cast<const float4*>(nr + e);
      if (KIND == ORX_PAIR_BPR) {
        s1 += u.x * p.x + u.y * p.y + u.z * p.z + u.w * p.w;
        s2 += u.x * n.x + u.y * n.y + u.z * n.z + u.w * n.w;
      } else {
        s1 += (u.x - p.x) * (u.x - p.x) + (u.y - p.y) * (u.y - p.y) + (u.z - p.z) * (u.z - p.z) + (u.w - p.w) * (u.w - p.w);
        s2 += (u.x - n.x) * (u.x - n.x) + (u.y - n.y) * (u.y - n.y) + (u.z - n.z) * (u.z - n.z) + (u.w - n.w) * (u.w - n.w);
      }
      sq += u.x * u.x + u.y * u.y + u.z * u.z + u.w * u.w + p.x * p.x + p.y * p.y + p.z * p.z + p.w * p.w +
            n.x * n.x + n.y * n.y + n.z * n.z + n.w * n.w;
    }
    l2_acc += sq;
    s1 = orx_group_sum<32>(s1);
    s2 = orx_group_sum<32>(s2);
    float lt = 0.f, g = 0.f;
    pair_score<KIND>(s1, s2, pr[D], nr[D], sa, &lt, &g);
    if (lane == 0) loss_acc += lt;
    // destinations: owner o of position p, row base[o] + (p - off[o]) of o's gradient inbox
    const int ou = xchg_bucket_of(off, R, pu), op = xchg_bucket_of(off, R, pp), on = xchg_bucket_of(off, R, pn);
    const bool live = base[ou] >= 0 && base[op] >= 0 && base[on] >= 0;   // -1: an owner declared overflow
    float* du = x.gin[ou] + ((int64_t)base[ou] + (pu - off[ou])) * W;
    float* dp = x.gin[op] + ((int64_t)base[op] + (pp - off[op])) * W;
    float* dn = x.gin[on] + ((int64_t)base[on] + (pn - off[on])) * W;
    const float t2 = 2.f * g, c2 = c_l2;
    if (live) {
      for (int e = lane * 4; e < D; e += 128) {
        const float4 u = *reinterpret_cast<const float4*>(ur + e), p = *reinterpret_cast<const float4*>(pr + e),
                     n = *reinterpret_cast<const float4*>(nr + e);
        float4 gu, gp, gn;
        if (KIND == ORX_PAIR_BPR) {
          gu = make_float4(g * (p.x - n.x) + c2 * u.x, g * (p.y - n.y) + c2 * u.y, g * (p.z - n.z) + c2 * u.z, g * (p.w - n.w) + c2 * u.w);
          gp = make_float4(g * u.x + c2 * p.x, g * u.y + c2 * p.y, g * u.z + c2 * p.z, g * u.w + c2 * p.w);
          gn = make_float4(-g * u.x + c2 * n.x, -g * u.y + c2 * n.y, -g * u.z + c2 * n.z, -g * u.w + c2 * n.w);
        } else {
          gu = make_float4(t2 * (n.x - p.x) + c2 * u.x, t2 * (n.y - p.y) + c2 * u.y, t2 * (n.z - p.z) + c2 * u.z, t2 * (n.w - p.w) + c2 * u.w);
          gp = make_float4(t2 * (p.x - u.x) + c2 * p.x, t2 * (p.y - u.y) + c2 * p.y, t2 * (p.z - u.z) + c2 * p.z, t2 * (p.w - u.w) + c2 * p.w);
          gn = make_float4(t2 * (u.x - n.x) + c2 * n.x, t2 * (u.y - n.y) + c2 * n.y, t2 * (u.z - n.z) + c2 * n.z, t2 * (u.w - n.w) + c2 * n.w);
        }
        *reinterpret_cast<float4*>(du + e) = gu;
        *reinterpret_cast<float4*>(dp + e) = gp;
        *reinterpret_cast<float4*>(dn + e) = gn;
      }
      if (lane == 0) {   // column D: item-bias gradient (0 for users); the padding columns stay 0
        const float gb = (KIND == ORX_PAIR_BPR) ? g : -g;
        *reinterpret_cast<float4*>(du + D) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(dp + D) = make_float4(gb, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(dn + D) = make_float4(-gb, 0.f, 0.f, 0.f);
      }
    }
  }
  loss_acc = orx_group_sum<32>(loss_acc);
  l2_acc = orx_group_sum<32>(l2_acc);
  if (lane == 0) {
    partials[2 * warp] = loss_acc;
    partials[2 * warp + 1] = l2_acc;
  }
}

// D = 128 (one float4 per lane and row): a warp owns 4 triplets and issues all 12 row loads before the first reduction;
// lanes 0..3 resolve the slots, owners and destination rows of their triplet.
template <int KIND>
__global__ void __launch_bounds__(256) k_xchg_grad_push128(XchgDev x, const int32_t* __restrict__ counts,
                                                           const int32_t* __restrict__ slot, int B, float margin,
                                                           float c_loss, float c_l2, float inv_B, float* partials) {
  constexpr int D = 128, T = 4;
  __shared__ int32_t off[XCHG_MAX_R + 1], base[XCHG_MAX_R];
  const int R = x.world, me = x.rank, W = x.W;
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int r = 0; r < R; ++r) { off[r] = acc; acc += counts[r]; }
    off[R] = acc;
  }
  if (threadIdx.x < R) base[threadIdx.x] = x.meta[me][XCHG_META * threadIdx.x + 2];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const float* got = x.got[me];
  PairArgs sa;
  sa.margin = margin; sa.c_loss = c_loss; sa.inv_B = inv_B;
  const int t0 = warp * T;
  // lanes 0..T-1: positions and destinations of triplet t0 + lane
  int my_p[3] = {0, 0, 0};
  float* my_d[3] = {nullptr, nullptr, nullptr};
  if (lane < T && t0 + lane < B) {
    bool live = true;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const int p = slot[q * B + t0 + lane];
      const int o = xchg_bucket_of(off, R, p);
      my_p[q] = p;
      my_d[q] = x.gin[o] + ((int64_t)base[o] + (p - off[o])) * W;
      live = live && base[o] >= 0;
    }
    if (!live) my_d[0] = my_d[1] = my_d[2] = nullptr;     // an owner declared overflow: nothing is pushed
  }
  float4 u[T], p[T], n[T];
  float bp[T], bn[T];
#pragma unroll
  for (int k = 0; k < T; ++k) {
    const int pu = __shfl_sync(ORX_FULL, my_p[0], k), pp = __shfl_sync(ORX_FULL, my_p[1], k), pn = __shfl_sync(ORX_FULL, my_p[2], k);
    u[k] = p[k] = n[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    bp[k] = bn[k] = 0.f;
    if (t0 + k < B) {
      u[k] = __ldcg(reinterpret_cast<const float4*>(got + (int64_t)pu * W) + lane);
      p[k] = __ldcg(reinterpret_cast<const float4*>(got + (int64_t)pp * W) + lane);
      n[k] = __ldcg(reinterpret_cast<const float4*>(got + (int64_t)pn * W) + lane);
      bp[k] = __ldcg(got + (int64_t)pp * W + D);
      bn[k] = __ldcg(got + (int64_t)pn * W + D);
    }
  }
  float loss_acc = 0.f, l2_acc = 0.f;
#pragma unroll
  for (int k = 0; k < T; ++k) {
    if (t0 + k >= B) break;
    float s1, s2;
    if (KIND == ORX_PAIR_BPR) {
      s1 = u[k].x * p[k].x + u[k].y * p[k].y + u[k].z * p[k].z + u[k].w * p[k].w;
      s2 = u[k].x * n[k].x + u[k].y * n[k].y + u[k].z * n[k].z + u[k].w * n[k].w;
    } else {
      s1 = (u[k].x - p[k].x) * (u[k].x - p[k].x) + (u[k].y - p[k].y) * (u[k].y - p[k].y) + (u[k].z - p[k].z) * (u[k].z - p[k].z) + (u[k].w - p[k].w) * (u[k].w - p[k].w);
      s2 = (u[k].x - n[k].x) * (u[k].x - n[k].x) + (u[k].y - n[k].y) * (u[k].y - n[k].y) + (u[k].z - n[k].z) * (u[k].z - n[k].z) + (u[k].w - n[k].w) * (u[k].w - n[k].w);
    }
    l2_acc += u[k].x * u[k].x + u[k].y * u[k].y + u[k].z * u[k].z + u[k].w * u[k].w + p[k].x * p[k].x + p[k].y * p[k].y +
              p[k].z * p[k].z + p[k].w * p[k].w + n[k].x * n[k].x + n[k].y * n[k].y + n[k].z * n[k].z + n[k].w * n[k].w;
    s1 = orx_group_sum<32>(s1);
    s2 = orx_group_sum<32>(s2);
    float lt = 0.f, g = 0.f;
    pair_score<KIND>(s1, s2, bp[k], bn[k], sa, &lt, &g);
    if (lane == 0) loss_acc += lt;
    float* du = reinterpret_cast<float*>(__shfl_sync(ORX_FULL, (unsigned long long)my_d[0], k));
    float* dp = reinterpret_cast<float*>(__shfl_sync(ORX_FULL, (unsigned long long)my_d[1], k));
    float* dn = reinterpret_cast<float*>(__shfl_sync(ORX_FULL, (unsigned long long)my_d[2], k));
    if (!du) continue;
    float4 gu, gp, gn;
    pair_row_grads<KIND>(g, c_l2, u[k], p[k], n[k], &gu, &gp, &gn);
    reinterpret_cast<float4*>(du)[lane] = gu;
    reinterpret_cast<float4*>(dp)[lane] = gp;
    reinterpret_cast<float4*>(dn)[lane] = gn;
    if (lane == 0) {   // column D: item-bias gradient (0 for users); the padding columns stay 0
      const float gb = (KIND == ORX_PAIR_BPR) ? g : -g;
      *reinterpret_cast<float4*>(du + D) = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(dp + D) = make_float4(gb, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(dn + D) = make_float4(-gb, 0.f, 0.f, 0.f);
    }
  }
  loss_acc = orx_group_sum<32>(loss_acc);
  l2_acc = orx_group_sum<32>(l2_acc);
  if (lane == 0) {
    partials[2 * warp] = loss_acc;
    partials[2 * warp + 1] = l2_acc;
  }
}

// launches the gradient kernel (D = 128 fast path or the generic one); returns the number of partial pairs written
static int xchg_launch_grad(orx_handle_t h, int kind, const XchgDev& xd, const int32_t* counts, const int32_t* slot, int B,
                            int dim, float margin, float c_loss, float c_l2, float inv_B, cudaStream_t st, int* n_partials) {
  const bool fast = dim == 128;
  const int per_warp = fast ? 4 : 8;
  const int nw = (B + per_warp - 1) / per_warp, blocks = (nw + 7) / 8;
  int rc = orx_ensure_partials(h, blocks * 8, st);
  if (rc) return rc;
  *n_partials = blocks * 8;
  if (fast) {
    if (kind == ORX_PAIR_BPR) k_xchg_grad_push128<ORX_PAIR_BPR><<<blocks, 256, 0, st>>>(xd, counts, slot, B, margin, c_loss, c_l2, inv_B, h->partials);
    else k_xchg_grad_push128<ORX_PAIR_UCML><<<blocks, 256, 0, st>>>(xd, counts, slot, B, margin, c_loss, c_l2, inv_B, h->partials);
  } else {
    if (kind == ORX_PAIR_BPR) k_xchg_grad_push<ORX_PAIR_BPR><<<blocks, 256, 0, st>>>(xd, counts, slot, B, dim, margin, c_loss, c_l2, inv_B, h->partials);
    else k_xchg_grad_push<ORX_PAIR_UCML><<<blocks, 256, 0, st>>>(xd, counts, slot, B, dim, margin, c_loss, c_l2, inv_B, h->partials);
  }
  ORX_LAUNCH_CHECK();
  return ORX_OK;
}

// (loss, l2) of the GLOBAL batch without a collective: every rank reduces its partials and stores the pair into every
// peer's meta[me][4..5] (before barrier C); after the barrier each rank adds the R pairs it holds, in rank order, so all
// ranks get bit-identical totals.
__global__ void k_xchg_loss_push(XchgDev x, const float* partials, int n, float loss_scale) {
  __shared__ double sh[2][256];
  double l = 0.0, q = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    l += (double)partials[2 * i];
    q += (double)partials[2 * i + 1];
  }
  sh[0][threadIdx.x] = l;
  sh[1][threadIdx.x] = q;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + s];
      sh[1][threadIdx.x] += sh[1][threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x < x.world) {
    int32_t* m = x.meta[threadIdx.x] + XCHG_META * x.rank;
    m[4] = __float_as_int((float)(sh[0][0] * (double)loss_scale));
    m[5] = __float_as_int((float)(0.5 * sh[1][0]));
  }
}

__global__ void k_xchg_loss_sum(XchgDev x, float* out4) {
  if (threadIdx.x == 0) {
    const int32_t* m = x.meta[x.rank];
    float l = 0.f, q = 0.f;
    for (int r = 0; r < x.world; ++r) {
      l += __int_as_float(m[XCHG_META * r + 4]);
      q += __int_as_float(m[XCHG_META * r + 5]);
    }
    out4[0] = l; out4[1] = q; out4[2] = 0.f; out4[3] = 0.f;
  }
}

// ---------------------------------------------------------------------------------------
// barrier over the ranks' streams: flags[r][me] = epoch at every peer, then wait for every peer's epoch in mine
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long xchg_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__global__ void k_xchg_barrier(XchgDev x, int epoch, unsigned long long timeout_ns) {
  const int r = threadIdx.x;
  if (r >= x.world) return;
  __threadfence_system();
  int32_t* remote = x.flags[r] + x.rank;
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(remote), "r"(epoch) : "memory");
  const int32_t* mine = x.flags[x.rank] + r;
  const unsigned long long t0 = xchg_now();
  while (true) {
    int32_t v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
    if (v >= epoch) break;
    if (xchg_now() - t0 > timeout_ns) {   // a peer never arrived: do not hang the GPU, leave a sticky error
      x.flags[x.rank][x.world] = 1;
      break;
    }
  }
  __threadfence_system();
}

// ---------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------
static int xchg_check(orx_handle_t h, const XchgHost* x) {
  ORX_REQUIRE(h != nullptr && x != nullptr, "null handle / exchange descriptor");
  ORX_REQUIRE(x->world >= 1 && x->world <= XCHG_MAX_R && x->rank >= 0 && x->rank < x->world, "bad world / rank");
  ORX_REQUIRE(x->width >= 8 && (x->width & 3) == 0 && x->cap > 0, "row width must be a multiple of 4 floats");
  ORX_REQUIRE(x->idbox && x->meta && x->got && x->gin && x->flags, "null mailbox pointer table");
  return ORX_OK;
}

extern "C" int orx_xchg_push_ids(orx_handle_t h, const void* xchg_host, const int32_t* counts,
                                 const int32_t* send_local, int32_t n, orx_stream_t s) {
  const XchgHost* x = (const XchgHost*)xchg_host;
  int rc = xchg_check(h, x);
  if (rc) return rc;
  ORX_REQUIRE(counts && send_local && n >= 0 && n <= x->cap, "bad lookups (n must not exceed the idbox capacity)");
  ORX_CUDA(cudaSetDevice(h->device));
  int blocks = (n + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > h->num_sms * 8) blocks = h->num_sms * 8;
  k_xchg_push_ids<<<blocks, 256, 0, (cudaStream_t)s>>>(to_dev(x), counts, send_local, n);
  ORX_LAUNCH_CHECK();
  return ORX_OK;
}

extern "C" int orx_xchg_gather_push(orx_handle_t h, const void* xchg_host, const float* table, int64_t rows,
                                    int32_t gin_rows, int32_t* req, int32_t* n_dev, int32_t* n_bad, orx_stream_t s) {
  const XchgHost* x = (const XchgHost*)xchg_host;
  int rc = xchg_check(h, x);
  if (rc) return rc;
  ORX_REQUIRE(table && rows > 0 && gin_rows > 0 && req && n_dev, "bad arguments");
  ORX_CUDA(cudaSetDevice(h->device));
  k_xchg_gather_push<<<h->num_sms * 8, 256, 0, (cudaStream_t)s>>>(to_dev(x), table, rows, gin_rows, req, n_dev, n_bad);
  ORX_LAUNCH_CHECK();
  return ORX_OK;
}

extern "C" int orx_xchg_grad_push(orx_handle_t h, int32_t kind, const void* xchg_host, const int32_t* counts,
                                  const int32_t* slot, int32_t B, int32_t dim, float margin, float c_loss, float c_l2,
                                  float inv_B, float* out4, orx_stream_t s) {
  const XchgHost* x = (const XchgHost*)xchg_host;
  int rc = xchg_check(h, x);
  if (rc) return rc;
  ORX_REQUIRE(kind == ORX_PAIR_BPR || kind == ORX_PAIR_UCML, "unknown pairwise kind");
  ORX_REQUIRE(counts && slot && out4 && B > 0 && 3 * (int64_t)B <= x->cap, "bad batch");
  ORX_REQUIRE(dim > 0 && (dim & 3) == 0 && dim + 4 <= x->width, "dim must be a multiple of 4 and leave the bias column");
  ORX_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)s;
  int np = 0;
  if ((rc = xchg_launch_grad(h, kind, to_dev(x), counts, slot, B, dim, margin, c_loss, c_l2, inv_B, st, &np))) return rc;
  return orx_launch_reduce_partials(h->partials, np, kind == ORX_PAIR_BPR ? inv_B : 1.f, out4, st);
}

extern "C" int orx_xchg_barrier(orx_handle_t h, const void* xchg_host, int32_t epoch, int32_t timeout_ms,
                                orx_stream_t s) {
  const XchgHost* x = (const XchgHost*)xchg_host;
  int rc = xchg_check(h, x);
  if (rc) return rc;
  ORX_REQUIRE(epoch > 0 && timeout_ms > 0, "epoch and timeout must be positive");
  ORX_CUDA(cudaSetDevice(h->device));
  k_xchg_barrier<<<1, XCHG_MAX_R, 0, (cudaStream_t)s>>>(to_dev(x), epoch, (unsigned long long)timeout_ms * 1000000ull);
  ORX_LAUNCH_CHECK();
  return ORX_OK;
}

// The whole step in one call (nine launches + three barriers, nothing returns to the host): see the file header.
// out4[0..1] = (loss, l2_loss) of the GLOBAL batch on every rank -- exchanged through the meta mailboxes, no collective.
// work: int32[world + 1 + ceil(3B/1024) * world] scratch; slot: int32[3B]; req: int32[gin_rows]; epoch_base: the
// barriers use epochs epoch_base+1..+3 (the caller advances it by 3 per step); gin_local: this rank's own gradient
// inbox (the local address of gin[rank]).
extern "C" int orx_xchg_step(orx_handle_t h, int32_t kind, const void* xchg_host, const orx_table_t* tab,
                             const int32_t* uid, const int32_t* pid, const int32_t* nid, int32_t B, int64_t total_users,
                             int32_t dim, const float* gin_local, int32_t gin_rows, int32_t* work, int32_t* slot, int32_t* req,
                             float margin,
                             float c_loss, float c_l2, float inv_B, const orx_opt_t* opt, int32_t epoch_base,
                             int32_t timeout_ms, float* out4, orx_stream_t s) {
  const XchgHost* x = (const XchgHost*)xchg_host;
  int rc = xchg_check(h, x);
  if (rc) return rc;
  ORX_REQUIRE(kind == ORX_PAIR_BPR || kind == ORX_PAIR_UCML, "unknown pairwise kind");
  ORX_REQUIRE(tab && tab->var && uid && pid && nid && work && slot && req && opt && out4 && gin_local, "null pointer");
  ORX_REQUIRE(B > 0 && 3 * (int64_t)B <= x->cap && total_users > 0 && gin_rows > 0, "bad batch / sizes");
  ORX_REQUIRE(dim > 0 && (dim & 3) == 0 && dim + 4 <= x->width && tab->dim == x->width, "bad dim / row width");
  ORX_REQUIRE(epoch_base >= 0 && timeout_ms > 0, "bad epoch / timeout");
  ORX_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)s;
  if ((rc = orx_ensure_workspace(h, gin_rows, x->width, opt->kind == ORX_OPT_ADAM_DENSE))) return rc;
  const XchgDev xd = to_dev(x);
  const int R = x->world, nchunks = (3 * B + XCHG_CHUNK - 1) / XCHG_CHUNK;
  int32_t* counts = work;            // [R]
  int32_t* n_dev = work + R;         // [1]
  int32_t* bc = work + R + 1;        // [nchunks][R]
  const unsigned long long tmo = (unsigned long long)timeout_ms * 1000000ull;
  k_xchg_hist<<<nchunks, 256, 0, st>>>(uid, pid, nid, B, R, bc);
  k_xchg_scatter_push<<<nchunks, 256, 0, st>>>(xd, uid, pid, nid, B, total_users, bc, nchunks, counts, slot);
  k_xchg_barrier<<<1, XCHG_MAX_R, 0, st>>>(xd, epoch_base + 1, tmo);
  k_xchg_gather_push<<<h->num_sms * 8, 256, 0, st>>>(xd, tab->var, tab->rows, gin_rows, req, n_dev, nullptr);
  // the owner's ids are final: build their batch index on the side stream while the gradient rows travel
  if (!h->side_stream) {
    ORX_CUDA(cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking));
    ORX_CUDA(cudaEventCreateWithFlags(&h->side_ev[0], cudaEventDisableTiming));
    ORX_CUDA(cudaEventCreateWithFlags(&h->side_ev[1], cudaEventDisableTiming));
  }
  ORX_CUDA(cudaEventRecord(h->side_ev[0], st));
  ORX_CUDA(cudaStreamWaitEvent(h->side_stream, h->side_ev[0], 0));
  if ((rc = orx_launch_index_build_strided(h, req, 1, tab->rows, gin_rows, n_dev, opt->kind == ORX_OPT_ADAM_DENSE, h->side_stream))) return rc;
  ORX_CUDA(cudaEventRecord(h->side_ev[1], h->side_stream));
  k_xchg_barrier<<<1, XCHG_MAX_R, 0, st>>>(xd, epoch_base + 2, tmo);
  int np = 0;
  if ((rc = xchg_launch_grad(h, kind, xd, counts, slot, B, dim, margin, c_loss, c_l2, inv_B, st, &np))) return rc;
  k_xchg_loss_push<<<1, 256, 0, st>>>(xd, h->partials, np, kind == ORX_PAIR_BPR ? inv_B : 1.f);
  k_xchg_barrier<<<1, XCHG_MAX_R, 0, st>>>(xd, epoch_base + 3, tmo);
  k_xchg_loss_sum<<<1, 32, 0, st>>>(xd, out4);   // out4 = the GLOBAL (loss, l2_loss), identical on every rank
  ORX_LAUNCH_CHECK();
  ORX_CUDA(cudaStreamWaitEvent(st, h->side_ev[1], 0));
  return orx_sparse_apply_prebuilt(h, tab, req, gin_local, x->width, gin_rows, n_dev, opt, st);
}
