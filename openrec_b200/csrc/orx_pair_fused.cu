// orx_pair_fused.cu -- the whole BPR / UCML training step as ONE persistent cooperative kernel.
//
// Profiles r1b-r1d showed (a) ~35 % of the step in the two small latency-chain kernels around the main one and
// (b) the main kernel structure-bound at ~4.4 TB/s while random 512 B read-modify-write reaches 5.6-6.0 TB/s on
// this part (tools/randrow_bw.cu): with 8 triplets per short-lived warp the load pipeline never left its
// prologue.  Here one CTA per SM stays resident for the whole step:
//   phase A  batch index (hash inserts, several independent ids per thread)            | grid.sync
//   phase B  every warp owns an equal, static range of triplets; the rows of its next STAGES triplets are always
//            in flight in a per-warp shared-memory ring (cp.async / LDGSTS); the first rows are issued before the sync
//            (equal static partition of the triplets over all warps of the grid)                        | grid.sync
//   phase C  optimizer for the staged (shared) rows, staging re-zeroed, deterministic loss reduction, counters reset
// Semantics are exactly those of orx_pairwise.cu's three-launch path (same staging rule, same math).
#include <cooperative_groups.h>
#include <stdlib.h>

#include "orx_common.cuh"
#include "orx_pair.cuh"

namespace cg = cooperative_groups;

struct FusedArgs {
  PairArgs p;
  int32_t* counters;  // [0] staged user rows [1] staged item rows [2] ticket (unused here) [3] bad ids [5] work counter
  float* out4;
  float loss_scale;
  int n_chunks;
  unsigned long long* dbg;  // optional [16]: globaltimer stamps of block 0 / last-finishing warp (ORX_FUSED_DBG)
};

__device__ __forceinline__ unsigned long long orx_gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define ORX_STAMP(i) do { if (fa.dbg && blockIdx.x == 0 && threadIdx.x == 0) fa.dbg[i] = orx_gtime(); } while (0)

__device__ __forceinline__ void cp16(void* smem_dst, const float* gsrc, bool pred) {
  const unsigned saddr = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int bytes = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(saddr), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp4(void* smem_dst, const float* gsrc, bool pred) {
  const unsigned saddr = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int bytes = pred ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(saddr), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct ChunkMeta {  // per-lane: triplet (chunk*CH + lane), lanes < CH
  int u, p, n, du, dp, dn, fl;
};

template <int KIND, int OPT, int D, int STAGES, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) k_pair_fused(const FusedArgs fa) {
  constexpr int K = D / 128;  // float4 per lane per row (D = 128 or 256)
  constexpr bool S0 = (OPT == ORX_OPT_ADAGRAD || OPT == ORX_OPT_ADAM_LAZY);
  constexpr bool S1 = (OPT == ORX_OPT_ADAM_LAZY);
  constexpr int NR = 3 + (S0 ? 3 : 0) + (S1 ? 3 : 0);
  constexpr int STAGE_BYTES = NR * K * 512 + 32;  // rows + {bp, bn, bps0, bns0, bps1, bns1, -, -}
  static_assert(STAGES >= 2 && STAGES <= 8, "bad ring depth");
  extern __shared__ __align__(16) unsigned char orx_smem[];
  const PairArgs& a = fa.p;
  cg::grid_group grid = cg::this_grid();

  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;

  unsigned char* wring = orx_smem + (size_t)wib * STAGES * STAGE_BYTES;
  auto row_slot = [&](int stage, int row, int k) -> float4* {
    return reinterpret_cast<float4*>(wring + (size_t)stage * STAGE_BYTES) + (row * K + k) * 32 + lane;
  };
  auto mini = [&](int stage) -> float* {
    return reinterpret_cast<float*>(wring + (size_t)stage * STAGE_BYTES + NR * K * 512);
  };
  auto load_ids_at = [&](int c0, int cnt, ChunkMeta& m) {   // triplets c0 .. c0+cnt-1, one per lane
    m.u = m.p = m.n = 0;
    m.du = m.dp = m.dn = -1;
    m.fl = 0;
    if (lane < cnt) {
      const int t = c0 + lane;
      m.u = a.uid[t];
      m.p = a.pid[t];
      m.n = a.nid[t];
      m.fl = (m.u >= 0 && m.u < a.rowsU && m.p >= 0 && m.p < a.rowsI && m.n >= 0 && m.n < a.rowsI) ? 1 : 0;
    }
  };
  // put the variable rows of triplet j of metadata m in flight into ring slot `stage` (needs ids only)
  auto issue_var = [&](const ChunkMeta& m, int j, int stage) {
    const bool v = __shfl_sync(ORX_FULL, m.fl, j) & 1;
    const int uu = __shfl_sync(ORX_FULL, m.u, j), pp = __shfl_sync(ORX_FULL, m.p, j), nn = __shfl_sync(ORX_FULL, m.n, j);
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int off = (k * 32 + lane) * 4;
      cp16(row_slot(stage, 0, k), v ? a.U + (int64_t)uu * D + off : a.U, v);
      cp16(row_slot(stage, 1, k), v ? a.I + (int64_t)pp * D + off : a.I, v);
      cp16(row_slot(stage, 2, k), v ? a.I + (int64_t)nn * D + off : a.I, v);
    }
    float* ms = mini(stage);
    if (lane == 0) cp4(ms + 0, v ? a.Bv + pp : a.Bv, v);
    if (lane == 1) cp4(ms + 1, v ? a.Bv + nn : a.Bv, v);
  };
  ORX_STAMP(0);
  // ------------------------------------------------------------------ phase A: batch index
  // Four independent ids per thread: first-slot loads, then claims (CAS) are issued back to back so their L2
  // round trips overlap; anything that did not claim an empty slot at once takes the general insert.
  {
    const int total = 3 * a.B;
    for (int base = gtid; base < total; base += 4 * nth) {
      int32_t id[4];
      uint32_t h[4];
      unsigned long long w[4], mine[4];
      bool live[4], isu[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int i = base + q * nth;
        live[q] = i < total;
        isu[q] = i < a.B;
        id[q] = !live[q] ? 0 : (isu[q] ? a.uid[i] : (i < 2 * a.B ? a.pid[i - a.B] : a.nid[i - 2 * a.B]));
        if (live[q] && !(id[q] >= 0 && (int64_t)id[q] < (isu[q] ? a.rowsU : a.rowsI))) {
          atomicAdd(fa.counters + 3, 1);
          live[q] = false;
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const OrxHash& t = isu[q] ? a.hu : a.hi;
        h[q] = orx_hash32((uint32_t)id[q], t.shift);
        mine[q] = orx_slot_word(t.epoch, id[q]);
        w[q] = live[q] ? __ldcg(t.slots + h[q]) : 0ull;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const OrxHash& t = isu[q] ? a.hu : a.hi;
        if (live[q] && (uint32_t)(w[q] >> 33) != t.epoch) {
          if (atomicCAS(t.slots + h[q], w[q], mine[q]) == w[q]) live[q] = false;   // claimed: first occurrence
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (live[q]) orx_hash_insert(isu[q] ? a.hu : a.hi, id[q], 0);
    }
  }
  // this warp's own triplet range is static, so its first rows can already travel while the grid synchronises
  const int nw_total = gridDim.x * WARPS;
  const int T = (a.B + nw_total - 1) / nw_total;
  const int t_begin = min(a.B, (blockIdx.x * WARPS + wib) * T), t_end = min(a.B, t_begin + T);
  ChunkMeta m_first;
  load_ids_at(t_begin, min(32, t_end - t_begin), m_first);
#pragma unroll
  for (int j = 0; j < STAGES; ++j) issue_var(m_first, j, j);
  ORX_STAMP(1);
  grid.sync();
  ORX_STAMP(2);
  const int n_su = fa.counters[0], n_si = fa.counters[1], n_bad = fa.counters[3];  // final after phase A

  // ------------------------------------------------------------------ phase B: persistent gather-score-update
  auto probe = [&](ChunkMeta& m) {
    if (m.fl & 1) {
      const uint32_t cu = orx_hash_find(a.hu, m.u, &m.du);
      const uint32_t cp = orx_hash_find(a.hi, m.p, &m.dp);
      const uint32_t cn = orx_hash_find(a.hi, m.n, &m.dn);
      m.fl |= (cu == 1u ? 2 : 0) | (cp == 1u ? 4 : 0) | (cn == 1u ? 8 : 0);
    }
  };
  // optimizer-slot rows (+ bias slots) of the rows this triplet owns: needs the probe results
  auto issue_slots = [&](const ChunkMeta& m, int j, int stage) {
    if (!S0) return;
    const int fl = __shfl_sync(ORX_FULL, m.fl, j);
    const int uu = __shfl_sync(ORX_FULL, m.u, j), pp = __shfl_sync(ORX_FULL, m.p, j), nn = __shfl_sync(ORX_FULL, m.n, j);
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int off = (k * 32 + lane) * 4;
      cp16(row_slot(stage, 3, k), (fl & 2) ? a.Us0 + (int64_t)uu * D + off : a.Us0, fl & 2);
      cp16(row_slot(stage, 4, k), (fl & 4) ? a.Is0 + (int64_t)pp * D + off : a.Is0, fl & 4);
      cp16(row_slot(stage, 5, k), (fl & 8) ? a.Is0 + (int64_t)nn * D + off : a.Is0, fl & 8);
      if (S1) {
        cp16(row_slot(stage, 6, k), (fl & 2) ? a.Us1 + (int64_t)uu * D + off : a.Us1, fl & 2);
        cp16(row_slot(stage, 7, k), (fl & 4) ? a.Is1 + (int64_t)pp * D + off : a.Is1, fl & 4);
        cp16(row_slot(stage, 8, k), (fl & 8) ? a.Is1 + (int64_t)nn * D + off : a.Is1, fl & 8);
      }
    }
    float* ms = mini(stage);
    if (lane == 2) cp4(ms + 2, (fl & 4) ? a.Bs0 + pp : a.Bs0, fl & 4);
    if (lane == 3) cp4(ms + 3, (fl & 8) ? a.Bs0 + nn : a.Bs0, fl & 8);
    if (S1 && lane == 4) cp4(ms + 4, (fl & 4) ? a.Bs1 + pp : a.Bs1, fl & 4);
    if (S1 && lane == 5) cp4(ms + 5, (fl & 8) ? a.Bs1 + nn : a.Bs1, fl & 8);
  };

  float loss_acc = 0.f, l2_acc = 0.f;
  auto process = [&](const ChunkMeta& m, int j, int stage) {
    const int fl = __shfl_sync(ORX_FULL, m.fl, j);
    const int uu = __shfl_sync(ORX_FULL, m.u, j), pp = __shfl_sync(ORX_FULL, m.p, j), nn = __shfl_sync(ORX_FULL, m.n, j);
    const int duj = __shfl_sync(ORX_FULL, m.du, j), dpj = __shfl_sync(ORX_FULL, m.dp, j),
              dnj = __shfl_sync(ORX_FULL, m.dn, j);
    __syncwarp();  // the 4-byte bias copies were issued by lanes 0..5; their wait_group has completed above
    const float* ms = mini(stage);
    const float bp = ms[0], bn = ms[1];
    float4 u[K], p[K], n[K];
    float s1 = 0.f, s2 = 0.f, sq = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      u[k] = *row_slot(stage, 0, k);
      p[k] = *row_slot(stage, 1, k);
      n[k] = *row_slot(stage, 2, k);
      if (KIND == ORX_PAIR_BPR) {
        s1 += dot4(u[k], p[k]);
        s2 += dot4(u[k], n[k]);
      } else {
        s1 += sqd4(u[k], p[k]);
        s2 += sqd4(u[k], n[k]);
      }
      sq += dot4(u[k], u[k]) + dot4(p[k], p[k]) + dot4(n[k], n[k]);
    }
    l2_acc += sq;
    s1 = orx_group_sum<32>(s1);
    s2 = orx_group_sum<32>(s2);
    float lt, g;
    pair_score<KIND>(s1, s2, bp, bn, a, &lt, &g);
    const bool v = fl & 1;
    if (!v) { lt = 0.f; g = 0.f; }
    if (lane == 0) loss_acc += lt;
    if (v) {
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const int off = (k * 32 + lane) * 4;
        float4 gu, gp, gn;
        pair_row_grads<KIND>(g, a.c_l2, u[k], p[k], n[k], &gu, &gp, &gn);
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        if (fl & 2) {
          const int64_t o = (int64_t)uu * D + off;
          float4 a0 = S0 ? *row_slot(stage, 3, k) : z, a1 = S1 ? *row_slot(stage, 6, k) : z;
          __stcg(reinterpret_cast<float4*>(a.U + o), orx_apply4<OPT>(u[k], gu, a0, a1, a.opt));
          if (S0) __stcg(reinterpret_cast<float4*>(a.Us0 + o), a0);
          if (S1) __stcg(reinterpret_cast<float4*>(a.Us1 + o), a1);
        } else {
          orx_red4(a.gu + (int64_t)duj * D + off, gu);
        }
        if (fl & 4) {
          const int64_t o = (int64_t)pp * D + off;
          float4 a0 = S0 ? *row_slot(stage, 4, k) : z, a1 = S1 ? *row_slot(stage, 7, k) : z;
          __stcg(reinterpret_cast<float4*>(a.I + o), orx_apply4<OPT>(p[k], gp, a0, a1, a.opt));
          if (S0) __stcg(reinterpret_cast<float4*>(a.Is0 + o), a0);
          if (S1) __stcg(reinterpret_cast<float4*>(a.Is1 + o), a1);
        } else {
          orx_red4(a.gi + (int64_t)dpj * D + off, gp);
        }
        if (fl & 8) {
          const int64_t o = (int64_t)nn * D + off;
          float4 a0 = S0 ? *row_slot(stage, 5, k) : z, a1 = S1 ? *row_slot(stage, 8, k) : z;
          __stcg(reinterpret_cast<float4*>(a.I + o), orx_apply4<OPT>(n[k], gn, a0, a1, a.opt));
          if (S0) __stcg(reinterpret_cast<float4*>(a.Is0 + o), a0);
          if (S1) __stcg(reinterpret_cast<float4*>(a.Is1 + o), a1);
        } else {
          orx_red4(a.gi + (int64_t)dnj * D + off, gn);
        }
      }
      // item_bias: lane 0 updates the positive item's bias, lane 1 the negative's (BPR +g/-g, UCML -a/+a)
      const float gb = (KIND == ORX_PAIR_BPR) ? g : -g;
      if (lane < 2) {
        const bool pos = lane == 0;
        const int id = pos ? pp : nn, dj = pos ? dpj : dnj;
        const float gbias = pos ? gb : -gb, b0 = pos ? bp : bn;
        if (fl & (pos ? 4 : 8)) {
          float a0 = S0 ? ms[pos ? 2 : 3] : 0.f, a1 = S1 ? ms[pos ? 4 : 5] : 0.f;
          __stcg(a.Bv + id, orx_apply<OPT>(b0, gbias, a0, a1, a.opt));
          if (S0) __stcg(a.Bs0 + id, a0);
          if (S1) __stcg(a.Bs1 + id, a1);
        } else {
          atomicAdd(a.gb + dj, gbias);
        }
      }
    }
    __syncwarp();  // every lane has read this stage (incl. the shared bias scalars) before it is refilled
  };

  // static partition: every warp of the grid owns T consecutive triplets (equal work, no work counter);
  // metadata for up to 32 of them lives in the lanes.
  {
    bool first = true;
    for (int c0 = t_begin; c0 < t_end; c0 += 32) {
      const int cnt = min(32, t_end - c0);
      ChunkMeta mc;
      if (first) mc = m_first;               // ids loaded (and the first rows put in flight) before the grid sync
      else load_ids_at(c0, cnt, mc);
      if (!first) {
#pragma unroll
        for (int j = 0; j < STAGES; ++j) issue_var(mc, j, j);
      }
      probe(mc);
#pragma unroll
      for (int j = 0; j < STAGES; ++j) {
        issue_slots(mc, j, j);
        cp_commit();
      }
      first = false;
      int ring = 0;
#pragma unroll 1
      for (int j = 0; j < cnt; ++j) {
        cp_wait<STAGES - 1>();
        process(mc, j, ring);
        if (j + STAGES < cnt) {
          issue_var(mc, j + STAGES, ring);
          issue_slots(mc, j + STAGES, ring);
        }
        cp_commit();
        ring = ring + 1 == STAGES ? 0 : ring + 1;
      }
      cp_wait<0>();
    }
    cp_wait<0>();   // (empty range: the zero-filled copies issued before the sync)
  }
  ORX_STAMP(3);
  if (fa.dbg && lane == 0) atomicMax(fa.dbg + 8, orx_gtime());   // when the LAST warp of the grid left phase B

  __shared__ float sred[WARPS][2];
  __shared__ double sdbl[2][256];
  loss_acc = orx_group_sum<32>(loss_acc);
  l2_acc = orx_group_sum<32>(l2_acc);
  if (lane == 0) {
    sred[wib][0] = loss_acc;
    sred[wib][1] = l2_acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float l = 0.f, q = 0.f;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) {
      l += sred[w][0];
      q += sred[w][1];
    }
    a.partials[2 * blockIdx.x] = l;
    a.partials[2 * blockIdx.x + 1] = q;
  }
  grid.sync();
  ORX_STAMP(4);

  // ------------------------------------------------------------------ phase C: staged rows, loss, reset
  {
    const int gwarp = gtid >> 5, nwarps = nth >> 5;
    for (int r = gwarp; r < n_su + n_si; r += nwarps) {
      const bool is_u = r < n_su;
      const int d = is_u ? r : r - n_su;
      const int id = is_u ? a.hu.did[d] : a.hi.did[d];
      float* G = (is_u ? a.gu : a.gi) + (int64_t)d * D;
      float* W = (is_u ? a.U : a.I) + (int64_t)id * D;
      float* P0 = (is_u ? a.Us0 : a.Is0) + (int64_t)id * D;
      float* P1 = (is_u ? a.Us1 : a.Is1) + (int64_t)id * D;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const int e = (k * 32 + lane) * 4;
        const float4 g = __ldcg(reinterpret_cast<const float4*>(G + e));
        float4 w = *reinterpret_cast<const float4*>(W + e);
        float4 s0v = S0 ? *reinterpret_cast<const float4*>(P0 + e) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 s1v = S1 ? *reinterpret_cast<const float4*>(P1 + e) : make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(W + e) = orx_apply4<OPT>(w, g, s0v, s1v, a.opt);
        if (S0) *reinterpret_cast<float4*>(P0 + e) = s0v;
        if (S1) *reinterpret_cast<float4*>(P1 + e) = s1v;
        *reinterpret_cast<float4*>(G + e) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (!is_u && lane == 0) {
        float s0v = S0 ? a.Bs0[id] : 0.f, s1v = S1 ? a.Bs1[id] : 0.f;
        a.Bv[id] = orx_apply<OPT>(a.Bv[id], __ldcg(a.gb + d), s0v, s1v, a.opt);
        if (S0) a.Bs0[id] = s0v;
        if (S1) a.Bs1[id] = s1v;
        a.gb[d] = 0.f;
      }
    }
    if (blockIdx.x == 0) {
      double l = 0.0, q = 0.0;
      if (threadIdx.x < 256) {
        for (int i = threadIdx.x; i < (int)gridDim.x; i += 256) {
          l += (double)__ldcg(a.partials + 2 * i);
          q += (double)__ldcg(a.partials + 2 * i + 1);
        }
        sdbl[0][threadIdx.x] = l;
        sdbl[1][threadIdx.x] = q;
      }
      __syncthreads();
      for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
          sdbl[0][threadIdx.x] += sdbl[0][threadIdx.x + s];
          sdbl[1][threadIdx.x] += sdbl[1][threadIdx.x + s];
        }
        __syncthreads();
      }
      if (threadIdx.x == 0) {
        fa.out4[0] = (float)(sdbl[0][0] * (double)fa.loss_scale);
        fa.out4[1] = (float)(0.5 * sdbl[1][0]);
        fa.out4[2] = (float)n_bad;
        fa.out4[3] = (float)(n_su + n_si);
        fa.counters[0] = 0;   // every block read these right after the first grid.sync
        fa.counters[1] = 0;
        fa.counters[3] = 0;
        fa.counters[5] = 0;   // work counter: all fetches happened before the second grid.sync
      }
    }
  }
  ORX_STAMP(5);
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
template <int KIND, int OPT, int D, int STAGES, int WARPS>
static int launch_fused(orx_ctx* c, FusedArgs& fa, cudaStream_t st) {
  constexpr bool S0 = (OPT == ORX_OPT_ADAGRAD || OPT == ORX_OPT_ADAM_LAZY);
  constexpr bool S1 = (OPT == ORX_OPT_ADAM_LAZY);
  constexpr int NR = 3 + (S0 ? 3 : 0) + (S1 ? 3 : 0);
  const size_t smem = (size_t)WARPS * STAGES * (NR * (D / 128) * 512 + 32);
  auto kern = k_pair_fused<KIND, OPT, D, STAGES, WARPS>;
  static bool configured = false;
  if (!configured) {
    ORX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    ORX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem));
    if (per_sm < 1) {
      orx_set_error("k_pair_fused does not fit on an SM (smem %zu)", smem);
      return ORX_ERR_UNSUPPORTED;
    }
    configured = true;
  }
  static unsigned long long* dbg = nullptr;
  static int dbg_on = -1, dbg_left = 0;
  if (dbg_on < 0) {
    dbg_on = getenv("ORX_FUSED_DBG") ? 1 : 0;
    if (dbg_on) {
      ORX_CUDA(cudaMalloc(&dbg, 16 * sizeof(unsigned long long)));
      dbg_left = 40;
    }
  }
  fa.dbg = (dbg_on && dbg_left > 0) ? dbg : nullptr;
  if (fa.dbg) ORX_CUDA(cudaMemsetAsync(dbg, 0, 16 * sizeof(unsigned long long), st));
  void* args[] = {(void*)&fa};
  ORX_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3(c->num_sms), dim3(WARPS * 32), args, smem, st));
  if (fa.dbg) {
    unsigned long long h[16];
    ORX_CUDA(cudaStreamSynchronize(st));
    ORX_CUDA(cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost));
    if (--dbg_left < 5)
      fprintf(stderr, "[orx fused dbg] A %.1f us | sync %.1f | B(block0) %.1f  B(last warp) %.1f | sync %.1f | C %.1f | total %.1f\n",
              (h[1] - h[0]) * 1e-3, (h[2] - h[1]) * 1e-3, (h[3] - h[2]) * 1e-3, (h[8] - h[2]) * 1e-3, (h[4] - h[3]) * 1e-3,
              (h[5] - h[4]) * 1e-3, (h[5] - h[0]) * 1e-3);
  }
  return ORX_OK;
}

// Returns ORX_ERR_UNSUPPORTED when this (kind, optimizer, dim) has no fused instance: the caller then uses
// the three-launch path.  ORX_FUSED_CFG = "<warps>x<stages>" picks the tuning point (default 16x4).
int orx_launch_pair_fused(orx_ctx* c, int kind, int opt_kind, PairArgs& pa, float loss_scale, float* out4,
                          cudaStream_t st) {
  if (pa.D != 128 || opt_kind == ORX_OPT_ADAM_DENSE || opt_kind == ORX_OPT_ADAM_LAZY) return ORX_ERR_UNSUPPORTED;
  FusedArgs fa;
  fa.p = pa;
  fa.counters = c->counters;
  fa.out4 = out4;
  fa.loss_scale = loss_scale;
  fa.n_chunks = (pa.B + 15) / 16;
  fa.p.partials = c->partials;
  static int cfg = -1;
  if (cfg < 0) {
    const char* e = getenv("ORX_FUSED_CFG");
    int w = 16, s = 4;
    if (e) sscanf(e, "%dx%d", &w, &s);
    cfg = w * 100 + s;
  }
#define ORX_FUSED_CASE(KIND, OPT)                                                         \
  switch (cfg) {                                                                           \
    case 808: return launch_fused<KIND, OPT, 128, 8, 8>(c, fa, st);                        \
    case 1603: return launch_fused<KIND, OPT, 128, 3, 16>(c, fa, st);                      \
    case 2402: return launch_fused<KIND, OPT, 128, 2, 24>(c, fa, st);                      \
    case 2403: return launch_fused<KIND, OPT, 128, 3, 24>(c, fa, st);                      \
    case 3202: return launch_fused<KIND, OPT, 128, 2, 32>(c, fa, st);                      \
    default: return launch_fused<KIND, OPT, 128, 4, 16>(c, fa, st);                        \
  }
  if (kind == ORX_PAIR_BPR) {
    if (opt_kind == ORX_OPT_SGD) { ORX_FUSED_CASE(ORX_PAIR_BPR, ORX_OPT_SGD) }
    ORX_FUSED_CASE(ORX_PAIR_BPR, ORX_OPT_ADAGRAD)
  } else {
    if (opt_kind == ORX_OPT_SGD) { ORX_FUSED_CASE(ORX_PAIR_UCML, ORX_OPT_SGD) }
    ORX_FUSED_CASE(ORX_PAIR_UCML, ORX_OPT_ADAGRAD)
  }
#undef ORX_FUSED_CASE
}
