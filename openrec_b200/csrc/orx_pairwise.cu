// orx_pairwise.cu -- BPR / UCML fused training step (K1/K2), its batch index (K9) and tail.
//
// Reference path replaced (paths relative to the reference repo):
//   openrec/tf2/recommenders/bpr.py:21-37, ucml.py:21-42, modules/pairwise_log_loss.py:15-34
//   + tape.gradient + optimizer.apply_gradients (tf2_examples/bpr_citeulike.py:33-39).
//
// Synchronous-batch semantics in three launches on one stream:
//   1. k_index_build : hash every id of the batch; rows hit more than once get a "staging" slot.
//   2. k_pair_step   : per triplet gather u,p,n (128-bit loads), score, loss, per-sample gradient.
//        * a row referenced exactly once in the batch is owned by its triplet: optimizer applied
//          in registers, row + slots written back once (read once, written once == algorithmic bytes);
//        * a row referenced more than once is NEVER written here: its per-sample gradient is
//          red.global.add'ed into the compact staging buffer (so every gather sees pre-step values).
//   3. k_pair_tail   : optimizer for the staged rows (once per unique row), staging re-zeroed,
//                      hash cleared, deterministic loss reduction.
#include <stdlib.h>

#include "orx_common.cuh"
#include "orx_pair.cuh"

// ---------------------------------------------------------------------------------------
// K9: batch index
// ---------------------------------------------------------------------------------------
__global__ void k_index_build(OrxHash hu, OrxHash hi, const int32_t* __restrict__ a, int64_t rows_a, int na,
                              const int32_t* __restrict__ b0, const int32_t* __restrict__ b1, int64_t rows_b, int nb,
                              int stage_all, int32_t* bad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = na + (b1 ? 2 * nb : nb);
  if (i >= total) return;
  if (stage_all == 3) {
    // reference-counting mode (pairwise only: na == nb, b1 != 0): a triplet with ANY id out of range is skipped as a
    // whole by k_pair_step, so none of its ids may be counted -- the counts must equal the decrements
    const int tq = i < na ? i : (i - na < nb ? i - na : i - na - nb);
    const int32_t x = a[tq], y = b0[tq], z = b1[tq];
    const bool ok = x >= 0 && (int64_t)x < rows_a && y >= 0 && (int64_t)y < rows_b && z >= 0 && (int64_t)z < rows_b;
    const int32_t id = i < na ? x : (i - na < nb ? y : z);
    const bool mine_ok = id >= 0 && (int64_t)id < (i < na ? rows_a : rows_b);
    if (!mine_ok) atomicAdd(bad, 1);
    if (ok) orx_hash_insert(i < na ? hu : hi, id, 3);
    return;
  }
  if (i < na) {
    const int32_t id = a[i];
    if (id >= 0 && (int64_t)id < rows_a) orx_hash_insert(hu, id, stage_all);
    else atomicAdd(bad, 1);
  } else {
    const int j = i - na;
    const int32_t id = j < nb ? b0[j] : b1[j - nb];
    if (id >= 0 && (int64_t)id < rows_b) orx_hash_insert(hi, id, stage_all);
    else atomicAdd(bad, 1);
  }
}

__global__ void k_index_build_strided(OrxHash hu, const int32_t* __restrict__ a, int64_t stride, int64_t rows, int n,
                                       const int32_t* __restrict__ n_dev, int stage_all, int32_t* bad) {
  if (n_dev) n = min(n, *n_dev);   // count produced on the device (mailbox exchange): grid-stride over it
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int32_t id = a[(int64_t)i * stride];
    if (id >= 0 && (int64_t)id < rows) orx_hash_insert(hu, id, stage_all);
    else atomicAdd(bad, 1);
  }
}

int orx_launch_index_build_strided(orx_ctx* c, const int32_t* a, int64_t stride, int64_t rows, int32_t n,
                                   const int32_t* n_dev, bool stage_all, cudaStream_t st) {
  if (n <= 0) return ORX_OK;
  orx_new_epoch(c);
  int blocks = (n + 255) / 256;
  if (n_dev && blocks > c->num_sms * 8) blocks = c->num_sms * 8;
  k_index_build_strided<<<blocks, 256, 0, st>>>(c->hu, a, stride, rows, n, n_dev, stage_all ? 1 : 0, c->counters + 3);
  ORX_LAUNCH_CHECK();
  return ORX_OK;
}

int orx_launch_index_build(orx_ctx* c, const int32_t* a, int64_t rows_a, int32_t na, const int32_t* b0,
                           const int32_t* b1, int64_t rows_b, int32_t nb, int mode, cudaStream_t st) {
  return orx_launch_index_build_on(c, c->hu, c->hi, c->counters, a, rows_a, na, b0, b1, rows_b, nb, mode, st);
}

// index build into an explicit (hash pair, counter block): the context's first set, or the second one of the
// experimental index / step overlap.  Epochs come from the one context counter, so each set sees increasing values.
int orx_launch_index_build_on(orx_ctx* c, OrxHash& hu, OrxHash& hi, int32_t* counters, const int32_t* a, int64_t rows_a,
                              int32_t na, const int32_t* b0, const int32_t* b1, int64_t rows_b, int32_t nb, int mode,
                              cudaStream_t st) {
  const int total = na + (b1 ? 2 * nb : nb);
  if (total <= 0) return ORX_OK;
  c->epoch++;
  hu.epoch = hi.epoch = c->epoch;
  k_index_build<<<(total + 255) / 256, 256, 0, st>>>(hu, hi, a, rows_a, na, b0, b1, rows_b, nb, mode, counters + 3);
  ORX_LAUNCH_CHECK();
  return ORX_OK;
}

// ---------------------------------------------------------------------------------------
// K1/K2: fused step
// ---------------------------------------------------------------------------------------
template <int K, bool S0, bool S1>
struct TripRegs {
  float4 u[K], p[K], n[K];
  float4 us0[K], ps0[K], ns0[K];  // dead arrays are eliminated when the optimizer has no such slot
  float4 us1[K], ps1[K], ns1[K];
  int fl, uu, pp, nn, du, dp, dn;
  int su, sp, sn;  // hash slot of the row (reference counter), LA variant only
  float bp, bn;
};

// flags: bit0 triplet valid, bit1/2/3 user/pos/neg row owned by this triplet (fast path)
//
// One warp owns CH consecutive triplets.  Order of issue inside a warp (latency first):
//   ids (coalesced, lanes < CH) -> variable rows of the first one/two triplet groups (they need only
//   the ids) -> hash probes + bias loads (lanes < CH, overlap the row loads) -> slot rows of the first
//   groups -> steady state: process one register buffer while the other's 128-bit loads are in flight.
// LA ("last arriver applies", ORX_PAIR_VARIANT=7/8, experimental): shared rows carry a reference count built by
// k_index_build (mode 3).  After a triplet has RED-added its gradient for a shared row it decrements the count; the
// contributor that takes it to zero applies the optimizer to that row on the spot.  Every other reader's load of the
// row precedes that reader's own decrement, so all gathers still see pre-step values, and the tail launch disappears
// (loss reduction + counter reset move to a last-block-done epilogue).  DESIGN.md section 10.1.
template <int KIND, int OPT, int D, int CH, int MINB, bool PIPE, bool LA = false>
__global__ void __launch_bounds__(256, MINB) k_pair_step(const PairArgs a) {
  constexpr int G = (D / 4 < 32) ? D / 4 : 32;  // lanes per triplet
  constexpr int K = D / (4 * G);                // float4 per lane per row
  constexpr int TPW = 32 / G;                   // triplets in flight per warp
  constexpr bool S0 = (OPT == ORX_OPT_ADAGRAD || OPT == ORX_OPT_ADAM_LAZY);
  constexpr bool S1 = (OPT == ORX_OPT_ADAM_LAZY);
  constexpr bool STAGE_ONLY = (OPT == ORX_OPT_ADAM_DENSE);
  static_assert(CH % TPW == 0, "chunk must be a multiple of the triplets per warp");
  typedef TripRegs<K, S0, S1> Regs;

  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int grp = lane / G, gl = lane % G;
  const int t = warp * CH + lane;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);

  // ---- ids of triplet t (lanes < CH)
  int u_id = 0, p_id = 0, n_id = 0, du = -1, dp = -1, dn = -1, flags = 0;
  if (lane < CH && t < a.B) {
    u_id = a.uid[t];
    p_id = a.pid[t];
    n_id = a.nid[t];
    flags = (u_id >= 0 && u_id < a.rowsU && p_id >= 0 && p_id < a.rowsI && n_id >= 0 && n_id < a.rowsI) ? 1 : 0;
  }

  // variable rows: need ids + the valid bit only
  auto load_var = [&](int j, Regs& r) {
    const int src = j + grp;
    r.fl = __shfl_sync(ORX_FULL, flags, src) & 1;
    r.uu = __shfl_sync(ORX_FULL, u_id, src);
    r.pp = __shfl_sync(ORX_FULL, p_id, src);
    r.nn = __shfl_sync(ORX_FULL, n_id, src);
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int off = (k * G + gl) * 4;
      r.u[k] = r.fl ? __ldcg(reinterpret_cast<const float4*>(a.U + (int64_t)r.uu * D + off)) : z4;
      r.p[k] = r.fl ? __ldcg(reinterpret_cast<const float4*>(a.I + (int64_t)r.pp * D + off)) : z4;
      r.n[k] = r.fl ? __ldcg(reinterpret_cast<const float4*>(a.I + (int64_t)r.nn * D + off)) : z4;
    }
  };
  Regs ra, rb;
  load_var(0, ra);
  if (PIPE && TPW < CH) load_var(TPW, rb);

  // ---- hash probes + item_bias (lanes < CH), overlapping the row loads above
  float bp = 0.f, bn = 0.f, bps0 = 0.f, bps1 = 0.f, bns0 = 0.f, bns1 = 0.f;
  int hsu = -1, hsp = -1, hsn = -1;
  if (flags & 1) {
    uint32_t cu, cp, cn;
    if constexpr (LA) {
      cu = orx_hash_find_slot(a.hu, u_id, &du, &hsu);
      cp = orx_hash_find_slot(a.hi, p_id, &dp, &hsp);
      cn = orx_hash_find_slot(a.hi, n_id, &dn, &hsn);
    } else {
      cu = orx_hash_find(a.hu, u_id, &du);
      cp = orx_hash_find(a.hi, p_id, &dp);
      cn = orx_hash_find(a.hi, n_id, &dn);
    }
    bp = __ldcg(a.Bv + p_id);
    bn = __ldcg(a.Bv + n_id);
    if (!STAGE_ONLY) {
      flags |= (cu == 1u ? 2 : 0) | (cp == 1u ? 4 : 0) | (cn == 1u ? 8 : 0);
      if (S0) {
        if (flags & 4) bps0 = __ldcg(a.Bs0 + p_id);
        if (flags & 8) bns0 = __ldcg(a.Bs0 + n_id);
      }
      if (S1) {
        if (flags & 4) bps1 = __ldcg(a.Bs1 + p_id);
        if (flags & 8) bns1 = __ldcg(a.Bs1 + n_id);
      }
    }
  }

  // optimizer-slot rows + staging indices: need the probe results
  auto load_slots = [&](int j, Regs& r) {
    const int src = j + grp;
    r.fl = __shfl_sync(ORX_FULL, flags, src);
    r.du = __shfl_sync(ORX_FULL, du, src);
    r.dp = __shfl_sync(ORX_FULL, dp, src);
    r.dn = __shfl_sync(ORX_FULL, dn, src);
    r.bp = __shfl_sync(ORX_FULL, bp, src);
    r.bn = __shfl_sync(ORX_FULL, bn, src);
    if constexpr (LA) {
      r.su = __shfl_sync(ORX_FULL, hsu, src);
      r.sp = __shfl_sync(ORX_FULL, hsp, src);
      r.sn = __shfl_sync(ORX_FULL, hsn, src);
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int off = (k * G + gl) * 4;
      if (S0) {
        r.us0[k] = (r.fl & 2) ? __ldcg(reinterpret_cast<const float4*>(a.Us0 + (int64_t)r.uu * D + off)) : z4;
        r.ps0[k] = (r.fl & 4) ? __ldcg(reinterpret_cast<const float4*>(a.Is0 + (int64_t)r.pp * D + off)) : z4;
        r.ns0[k] = (r.fl & 8) ? __ldcg(reinterpret_cast<const float4*>(a.Is0 + (int64_t)r.nn * D + off)) : z4;
      }
      if (S1) {
        r.us1[k] = (r.fl & 2) ? __ldcg(reinterpret_cast<const float4*>(a.Us1 + (int64_t)r.uu * D + off)) : z4;
        r.ps1[k] = (r.fl & 4) ? __ldcg(reinterpret_cast<const float4*>(a.Is1 + (int64_t)r.pp * D + off)) : z4;
        r.ns1[k] = (r.fl & 8) ? __ldcg(reinterpret_cast<const float4*>(a.Is1 + (int64_t)r.nn * D + off)) : z4;
      }
    }
  };
  load_slots(0, ra);
  if (PIPE && TPW < CH) load_slots(TPW, rb);

  float loss_acc = 0.f, l2_acc = 0.f, g_own = 0.f;

  auto process = [&](int j, Regs& r) {
    float s1 = 0.f, s2 = 0.f, sq = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      if (KIND == ORX_PAIR_BPR) {
        s1 += dot4(r.u[k], r.p[k]);
        s2 += dot4(r.u[k], r.n[k]);
      } else {
        s1 += sqd4(r.u[k], r.p[k]);
        s2 += sqd4(r.u[k], r.n[k]);
      }
      sq += dot4(r.u[k], r.u[k]) + dot4(r.p[k], r.p[k]) + dot4(r.n[k], r.n[k]);
    }
    l2_acc += sq;  // invalid triplets contribute exact zeros
    s1 = orx_group_sum<G>(s1);
    s2 = orx_group_sum<G>(s2);
    float lt, g;
    pair_score<KIND>(s1, s2, r.bp, r.bn, a, &lt, &g);
    const bool v = r.fl & 1;
    if (!v) { lt = 0.f; g = 0.f; }
    if (gl == 0) loss_acc += lt;
    // bias gradient of the positive item: BPR +g, UCML -a  (negative item gets the opposite sign)
    const float gbias = (KIND == ORX_PAIR_BPR) ? g : -g;
#pragma unroll
    for (int q = 0; q < TPW; ++q) {
      const float val = __shfl_sync(ORX_FULL, gbias, q * G);
      if (lane == j + q) g_own = val;
    }
    if (v) {
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const int off = (k * G + gl) * 4;
        float4 gu, gp, gn;
        pair_row_grads<KIND>(g, a.c_l2, r.u[k], r.p[k], r.n[k], &gu, &gp, &gn);
        if (!STAGE_ONLY && (r.fl & 2)) {
          const int64_t o = (int64_t)r.uu * D + off;
          __stcg(reinterpret_cast<float4*>(a.U + o), orx_apply4<OPT>(r.u[k], gu, r.us0[k], r.us1[k], a.opt));
          if (S0) __stcg(reinterpret_cast<float4*>(a.Us0 + o), r.us0[k]);
          if (S1) __stcg(reinterpret_cast<float4*>(a.Us1 + o), r.us1[k]);
        } else {
          orx_red4(a.gu + (int64_t)r.du * D + off, gu);
        }
        if (!STAGE_ONLY && (r.fl & 4)) {
          const int64_t o = (int64_t)r.pp * D + off;
          __stcg(reinterpret_cast<float4*>(a.I + o), orx_apply4<OPT>(r.p[k], gp, r.ps0[k], r.ps1[k], a.opt));
          if (S0) __stcg(reinterpret_cast<float4*>(a.Is0 + o), r.ps0[k]);
          if (S1) __stcg(reinterpret_cast<float4*>(a.Is1 + o), r.ps1[k]);
        } else {
          orx_red4(a.gi + (int64_t)r.dp * D + off, gp);
        }
        if (!STAGE_ONLY && (r.fl & 8)) {
          const int64_t o = (int64_t)r.nn * D + off;
          __stcg(reinterpret_cast<float4*>(a.I + o), orx_apply4<OPT>(r.n[k], gn, r.ns0[k], r.ns1[k], a.opt));
          if (S0) __stcg(reinterpret_cast<float4*>(a.Is0 + o), r.ns0[k]);
          if (S1) __stcg(reinterpret_cast<float4*>(a.Is1 + o), r.ns1[k]);
        } else {
          orx_red4(a.gi + (int64_t)r.dn * D + off, gn);
        }
      }
      if constexpr (LA && !STAGE_ONLY) {
        const int shared = (~r.fl) & 14;   // rows of this triplet that went to the staging buffers
        if (shared) {
          // bias contributions of shared item rows must be in before the reference is given back
          if (gl == 0) {
            if (shared & 4) atomicAdd(a.gb + r.dp, gbias);
            if (shared & 8) atomicAdd(a.gb + r.dn, -gbias);
          }
          __threadfence();                 // my REDs / atomics are visible before my decrements
          int last = 0;
          if (gl == 0) {
            if ((shared & 2) && (uint32_t)atomicAdd(a.hu.cnt + r.su, ~0ull) == 1u) last |= 2;
            if ((shared & 4) && (uint32_t)atomicAdd(a.hi.cnt + r.sp, ~0ull) == 1u) last |= 4;
            if ((shared & 8) && (uint32_t)atomicAdd(a.hi.cnt + r.sn, ~0ull) == 1u) last |= 8;
          }
          last = __shfl_sync(ORX_FULL, last, grp * G);
          if (last) {
            __threadfence();               // every contributor's REDs happened before its decrement, hence before mine
            auto apply_row = [&](float* W, float* P0, float* P1, float* stage, int id, int d, const float4* cur) {
#pragma unroll
              for (int k = 0; k < K; ++k) {
                const int off = (k * G + gl) * 4;
                const int64_t o = (int64_t)id * D + off;
                float* sp_ = stage + (int64_t)d * D + off;
                const float4 gs = __ldcg(reinterpret_cast<const float4*>(sp_));
                float4 s0v = S0 ? __ldcg(reinterpret_cast<const float4*>(P0 + o)) : z4;
                float4 s1v = S1 ? __ldcg(reinterpret_cast<const float4*>(P1 + o)) : z4;
                __stcg(reinterpret_cast<float4*>(W + o), orx_apply4<OPT>(cur[k], gs, s0v, s1v, a.opt));
                if (S0) __stcg(reinterpret_cast<float4*>(P0 + o), s0v);
                if (S1) __stcg(reinterpret_cast<float4*>(P1 + o), s1v);
                __stcg(reinterpret_cast<float4*>(sp_), z4);
              }
            };
            auto apply_bias = [&](int id, int d) {   // item rows only; one lane
              const float gbv = __ldcg(a.gb + d);
              float s0v = S0 ? __ldcg(a.Bs0 + id) : 0.f, s1v = S1 ? __ldcg(a.Bs1 + id) : 0.f;
              __stcg(a.Bv + id, orx_apply<OPT>(__ldcg(a.Bv + id), gbv, s0v, s1v, a.opt));
              if (S0) __stcg(a.Bs0 + id, s0v);
              if (S1) __stcg(a.Bs1 + id, s1v);
              __stcg(a.gb + d, 0.f);
            };
            if (last & 2) apply_row(a.U, a.Us0, a.Us1, a.gu, r.uu, r.du, r.u);
            if (last & 4) {
              apply_row(a.I, a.Is0, a.Is1, a.gi, r.pp, r.dp, r.p);
              if (gl == 0) apply_bias(r.pp, r.dp);
            }
            if (last & 8) {
              apply_row(a.I, a.Is0, a.Is1, a.gi, r.nn, r.dn, r.n);
              if (gl == 0) apply_bias(r.nn, r.dn);
            }
          }
        }
      }
    }
  };

  if (PIPE) {
#pragma unroll 1
    for (int j = 0; j < CH; j += 2 * TPW) {
      process(j, ra);
      if (j + 2 * TPW < CH) {
        load_var(j + 2 * TPW, ra);
        load_slots(j + 2 * TPW, ra);
      }
      if (j + TPW < CH) {
        process(j + TPW, rb);
        if (j + 3 * TPW < CH) {
          load_var(j + 3 * TPW, rb);
          load_slots(j + 3 * TPW, rb);
        }
      }
    }
  } else {
#pragma unroll 1
    for (int j = 0; j < CH; j += TPW) {
      if (j > 0) {
        load_var(j, ra);
        load_slots(j, ra);
      }
      process(j, ra);
    }
  }

  // ---- item_bias: lane-parallel, one lane per triplet of the chunk
  if (flags & 1) {
    if (flags & 4) {
      __stcg(a.Bv + p_id, orx_apply<OPT>(bp, g_own, bps0, bps1, a.opt));
      if (S0) __stcg(a.Bs0 + p_id, bps0);
      if (S1) __stcg(a.Bs1 + p_id, bps1);
    } else if (!LA || STAGE_ONLY) {
      atomicAdd(a.gb + dp, g_own);
    }
    if (flags & 8) {
      __stcg(a.Bv + n_id, orx_apply<OPT>(bn, -g_own, bns0, bns1, a.opt));
      if (S0) __stcg(a.Bs0 + n_id, bns0);
      if (S1) __stcg(a.Bs1 + n_id, bns1);
    } else if (!LA || STAGE_ONLY) {
      atomicAdd(a.gb + dn, -g_own);
    }
    if (a.g_out) a.g_out[t] = (KIND == ORX_PAIR_BPR) ? g_own : -g_own;
  } else if (a.g_out && lane < CH && t < a.B) {
    a.g_out[t] = 0.f;
  }

  // ---- one (loss, l2) partial per block, fixed order => deterministic
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
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      l += sred[w][0];
      q += sred[w][1];
    }
    a.partials[2 * blockIdx.x] = l;
    a.partials[2 * blockIdx.x + 1] = q;
  }
  if constexpr (LA) {
    // last block done: deterministic (fixed order, double) reduction of the per-block partials, out4, counter reset
    __shared__ bool is_last;
    __shared__ double dred[2][256];
    if (threadIdx.x == 0) {
      __threadfence();
      is_last = (atomicAdd(a.counters + 2, 1) == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
      __threadfence();
      double l = 0.0, q = 0.0;
      for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
        l += (double)__ldcg(a.partials + 2 * i);
        q += (double)__ldcg(a.partials + 2 * i + 1);
      }
      dred[0][threadIdx.x] = l;
      dred[1][threadIdx.x] = q;
      __syncthreads();
      for (int sft = 128; sft > 0; sft >>= 1) {
        if (threadIdx.x < sft) {
          dred[0][threadIdx.x] += dred[0][threadIdx.x + sft];
          dred[1][threadIdx.x] += dred[1][threadIdx.x + sft];
        }
        __syncthreads();
      }
      if (threadIdx.x == 0) {
        a.out4[0] = (float)(dred[0][0] * (double)a.loss_scale);
        a.out4[1] = (float)(0.5 * dred[1][0]);
        a.out4[2] = (float)a.counters[3];
        a.out4[3] = (float)(a.counters[0] + a.counters[1]);
        a.counters[0] = a.counters[1] = a.counters[2] = a.counters[3] = 0;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// cp.async variant: the rows of the next STAGES triplet groups are in flight in a per-warp shared-memory
// ring (LDGSTS.128, no registers held while in flight), so memory-level parallelism is set by STAGES x
// resident warps instead of by the register file.  Each lane copies and later reads only ITS OWN 16 bytes of
// every row, so no cross-lane synchronisation is needed: cp.async.wait_group orders a thread's own copies.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void orx_cp_async16(float4* smem_dst, const float* gsrc, bool pred) {
  const unsigned saddr = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int bytes = pred ? 16 : 0;   // src-size 0 => the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(saddr), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void orx_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void orx_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int KIND, int OPT, int D, int CH, int MINB, int STAGES>
__global__ void __launch_bounds__(256, MINB) k_pair_step_async(const PairArgs a) {
  constexpr int G = (D / 4 < 32) ? D / 4 : 32;
  constexpr int K = D / (4 * G);
  constexpr int TPW = 32 / G;
  constexpr int NG = CH / TPW;  // triplet groups per warp
  constexpr bool S0 = (OPT == ORX_OPT_ADAGRAD || OPT == ORX_OPT_ADAM_LAZY);
  constexpr bool S1 = (OPT == ORX_OPT_ADAM_LAZY);
  constexpr bool STAGE_ONLY = (OPT == ORX_OPT_ADAM_DENSE);
  constexpr int NR = 3 + (S0 ? 3 : 0) + (S1 ? 3 : 0);  // rows per triplet in the ring
  static_assert(CH % TPW == 0 && STAGES >= 1 && STAGES <= NG, "bad staging");
  extern __shared__ float4 orx_ring[];

  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int grp = lane / G, gl = lane % G;
  const int t = warp * CH + lane;
  float4* ring = orx_ring + (size_t)wib * STAGES * NR * K * 32 + lane;   // this lane's column of the warp's ring
  auto slot = [&](int stage, int row, int k) -> float4* { return ring + ((stage * NR + row) * K + k) * 32; };

  int u_id = 0, p_id = 0, n_id = 0, du = -1, dp = -1, dn = -1, flags = 0;
  if (lane < CH && t < a.B) {
    u_id = a.uid[t];
    p_id = a.pid[t];
    n_id = a.nid[t];
    flags = (u_id >= 0 && u_id < a.rowsU && p_id >= 0 && p_id < a.rowsI && n_id >= 0 && n_id < a.rowsI) ? 1 : 0;
  }

  auto issue_var = [&](int g, int stage) {
    const int src = g * TPW + grp;
    const bool v = __shfl_sync(ORX_FULL, flags, src) & 1;
    const int uu = __shfl_sync(ORX_FULL, u_id, src), pp = __shfl_sync(ORX_FULL, p_id, src),
              nn = __shfl_sync(ORX_FULL, n_id, src);
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int off = (k * G + gl) * 4;
      orx_cp_async16(slot(stage, 0, k), v ? a.U + (int64_t)uu * D + off : a.U, v);
      orx_cp_async16(slot(stage, 1, k), v ? a.I + (int64_t)pp * D + off : a.I, v);
      orx_cp_async16(slot(stage, 2, k), v ? a.I + (int64_t)nn * D + off : a.I, v);
    }
  };
  auto issue_slots = [&](int g, int stage) {
    if (!S0) return;
    const int src = g * TPW + grp;
    const int fl = __shfl_sync(ORX_FULL, flags, src);
    const int uu = __shfl_sync(ORX_FULL, u_id, src), pp = __shfl_sync(ORX_FULL, p_id, src),
              nn = __shfl_sync(ORX_FULL, n_id, src);
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int off = (k * G + gl) * 4;
      orx_cp_async16(slot(stage, 3, k), (fl & 2) ? a.Us0 + (int64_t)uu * D + off : a.Us0, fl & 2);
      orx_cp_async16(slot(stage, 4, k), (fl & 4) ? a.Is0 + (int64_t)pp * D + off : a.Is0, fl & 4);
      orx_cp_async16(slot(stage, 5, k), (fl & 8) ? a.Is0 + (int64_t)nn * D + off : a.Is0, fl & 8);
      if (S1) {
        orx_cp_async16(slot(stage, 6, k), (fl & 2) ? a.Us1 + (int64_t)uu * D + off : a.Us1, fl & 2);
        orx_cp_async16(slot(stage, 7, k), (fl & 4) ? a.Is1 + (int64_t)pp * D + off : a.Is1, fl & 4);
        orx_cp_async16(slot(stage, 8, k), (fl & 8) ? a.Is1 + (int64_t)nn * D + off : a.Is1, fl & 8);
      }
    }
  };

  // variable rows of the first STAGES groups go out before the hash probes
#pragma unroll
  for (int g = 0; g < STAGES; ++g) issue_var(g, g);

  float bp = 0.f, bn = 0.f, bps0 = 0.f, bps1 = 0.f, bns0 = 0.f, bns1 = 0.f;
  if (flags & 1) {
    const uint32_t cu = orx_hash_find(a.hu, u_id, &du);
    const uint32_t cp = orx_hash_find(a.hi, p_id, &dp);
    const uint32_t cn = orx_hash_find(a.hi, n_id, &dn);
    bp = __ldcg(a.Bv + p_id);
    bn = __ldcg(a.Bv + n_id);
    if (!STAGE_ONLY) {
      flags |= (cu == 1u ? 2 : 0) | (cp == 1u ? 4 : 0) | (cn == 1u ? 8 : 0);
      if (S0) {
        if (flags & 4) bps0 = __ldcg(a.Bs0 + p_id);
        if (flags & 8) bns0 = __ldcg(a.Bs0 + n_id);
      }
      if (S1) {
        if (flags & 4) bps1 = __ldcg(a.Bs1 + p_id);
        if (flags & 8) bns1 = __ldcg(a.Bs1 + n_id);
      }
    }
  }
#pragma unroll
  for (int g = 0; g < STAGES; ++g) {
    issue_slots(g, g);
    orx_cp_commit();   // group g = slots(g) (+ every var row issued so far for g == 0)
  }

  float loss_acc = 0.f, l2_acc = 0.f, g_own = 0.f;
#pragma unroll 1
  for (int g = 0; g < NG; ++g) {
    const int stage = g % STAGES;
    orx_cp_wait<STAGES - 1>();
    const int src = g * TPW + grp;
    const int fl = __shfl_sync(ORX_FULL, flags, src);
    const int uu = __shfl_sync(ORX_FULL, u_id, src), pp = __shfl_sync(ORX_FULL, p_id, src),
              nn = __shfl_sync(ORX_FULL, n_id, src);
    const int duj = __shfl_sync(ORX_FULL, du, src), dpj = __shfl_sync(ORX_FULL, dp, src),
              dnj = __shfl_sync(ORX_FULL, dn, src);
    const float bpj = __shfl_sync(ORX_FULL, bp, src), bnj = __shfl_sync(ORX_FULL, bn, src);
    float4 u[K], p[K], n[K];
    float s1 = 0.f, s2 = 0.f, sq = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      u[k] = *slot(stage, 0, k);
      p[k] = *slot(stage, 1, k);
      n[k] = *slot(stage, 2, k);
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
    s1 = orx_group_sum<G>(s1);
    s2 = orx_group_sum<G>(s2);
    float lt, gsc;
    pair_score<KIND>(s1, s2, bpj, bnj, a, &lt, &gsc);
    const bool v = fl & 1;
    if (!v) { lt = 0.f; gsc = 0.f; }
    if (gl == 0) loss_acc += lt;
    const float gbias = (KIND == ORX_PAIR_BPR) ? gsc : -gsc;
#pragma unroll
    for (int q = 0; q < TPW; ++q) {
      const float val = __shfl_sync(ORX_FULL, gbias, q * G);
      if (lane == g * TPW + q) g_own = val;
    }
    if (v) {
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const int off = (k * G + gl) * 4;
        float4 gu, gp, gn;
        pair_row_grads<KIND>(gsc, a.c_l2, u[k], p[k], n[k], &gu, &gp, &gn);
        float4 z0 = make_float4(0.f, 0.f, 0.f, 0.f), z1 = z0;
        if (!STAGE_ONLY && (fl & 2)) {
          const int64_t o = (int64_t)uu * D + off;
          float4 a0 = S0 ? *slot(stage, 3, k) : z0, a1 = S1 ? *slot(stage, 6, k) : z1;
          __stcg(reinterpret_cast<float4*>(a.U + o), orx_apply4<OPT>(u[k], gu, a0, a1, a.opt));
          if (S0) __stcg(reinterpret_cast<float4*>(a.Us0 + o), a0);
          if (S1) __stcg(reinterpret_cast<float4*>(a.Us1 + o), a1);
        } else {
          orx_red4(a.gu + (int64_t)duj * D + off, gu);
        }
        if (!STAGE_ONLY && (fl & 4)) {
          const int64_t o = (int64_t)pp * D + off;
          float4 a0 = S0 ? *slot(stage, 4, k) : z0, a1 = S1 ? *slot(stage, 7, k) : z1;
          __stcg(reinterpret_cast<float4*>(a.I + o), orx_apply4<OPT>(p[k], gp, a0, a1, a.opt));
          if (S0) __stcg(reinterpret_cast<float4*>(a.Is0 + o), a0);
          if (S1) __stcg(reinterpret_cast<float4*>(a.Is1 + o), a1);
        } else {
          orx_red4(a.gi + (int64_t)dpj * D + off, gp);
        }
        if (!STAGE_ONLY && (fl & 8)) {
          const int64_t o = (int64_t)nn * D + off;
          float4 a0 = S0 ? *slot(stage, 5, k) : z0, a1 = S1 ? *slot(stage, 8, k) : z1;
          __stcg(reinterpret_cast<float4*>(a.I + o), orx_apply4<OPT>(n[k], gn, a0, a1, a.opt));
          if (S0) __stcg(reinterpret_cast<float4*>(a.Is0 + o), a0);
          if (S1) __stcg(reinterpret_cast<float4*>(a.Is1 + o), a1);
        } else {
          orx_red4(a.gi + (int64_t)dnj * D + off, gn);
        }
      }
    }
    if (g + STAGES < NG) {   // refill the slot just consumed
      issue_var(g + STAGES, stage);
      issue_slots(g + STAGES, stage);
    }
    orx_cp_commit();         // one group per iteration (possibly empty) keeps wait_group<STAGES-1> exact
  }

  if (flags & 1) {
    if (flags & 4) {
      __stcg(a.Bv + p_id, orx_apply<OPT>(bp, g_own, bps0, bps1, a.opt));
      if (S0) __stcg(a.Bs0 + p_id, bps0);
      if (S1) __stcg(a.Bs1 + p_id, bps1);
    } else {
      atomicAdd(a.gb + dp, g_own);
    }
    if (flags & 8) {
      __stcg(a.Bv + n_id, orx_apply<OPT>(bn, -g_own, bns0, bns1, a.opt));
      if (S0) __stcg(a.Bs0 + n_id, bns0);
      if (S1) __stcg(a.Bs1 + n_id, bns1);
    } else {
      atomicAdd(a.gb + dn, -g_own);
    }
    if (a.g_out) a.g_out[t] = (KIND == ORX_PAIR_BPR) ? g_own : -g_own;
  } else if (a.g_out && lane < CH && t < a.B) {
    a.g_out[t] = 0.f;
  }

  __shared__ float sred[8][2];
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
    for (int w = 0; w < 8; ++w) {
      l += sred[w][0];
      q += sred[w][1];
    }
    a.partials[2 * blockIdx.x] = l;
    a.partials[2 * blockIdx.x + 1] = q;
  }
}

// Any dim (e.g. the example's D=50): one triplet per warp-iteration, lanes stride the row.
template <int KIND, int OPT>
__global__ void __launch_bounds__(256) k_pair_step_generic(const PairArgs a) {
  constexpr bool S0 = (OPT == ORX_OPT_ADAGRAD || OPT == ORX_OPT_ADAM_LAZY);
  constexpr bool S1 = (OPT == ORX_OPT_ADAM_LAZY);
  constexpr bool STAGE_ONLY = (OPT == ORX_OPT_ADAM_DENSE);
  constexpr int CH = 8;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int D = a.D;
  float loss_acc = 0.f, l2_acc = 0.f;
  for (int j = 0; j < CH; ++j) {
    const int t = warp * CH + j;
    if (t >= a.B) break;
    const int uu = a.uid[t], pp = a.pid[t], nn = a.nid[t];
    const bool ok = uu >= 0 && uu < a.rowsU && pp >= 0 && pp < a.rowsI && nn >= 0 && nn < a.rowsI;
    if (!ok) {
      if (a.g_out && lane == 0) a.g_out[t] = 0.f;
      continue;
    }
    int du, dp, dn;
    const uint32_t cu = orx_hash_find(a.hu, uu, &du);
    const uint32_t cp = orx_hash_find(a.hi, pp, &dp);
    const uint32_t cn = orx_hash_find(a.hi, nn, &dn);
    const bool fu = !STAGE_ONLY && cu == 1u, fp = !STAGE_ONLY && cp == 1u, fn = !STAGE_ONLY && cn == 1u;
    float* ur = a.U + (int64_t)uu * D;
    float* pr = a.I + (int64_t)pp * D;
    float* nr = a.I + (int64_t)nn * D;
    float s1 = 0.f, s2 = 0.f, sq = 0.f;
    for (int d = lane; d < D; d += 32) {
      const float u = ur[d], p = pr[d], n = nr[d];
      if (KIND == ORX_PAIR_BPR) {
        s1 += u * p;
        s2 += u * n;
      } else {
        s1 += (u - p) * (u - p);
        s2 += (u - n) * (u - n);
      }
      sq += u * u + p * p + n * n;
    }
    l2_acc += sq;
    s1 = orx_group_sum<32>(s1);
    s2 = orx_group_sum<32>(s2);
    const float bp = a.Bv[pp], bn = a.Bv[nn];
    float lt, g;
    pair_score<KIND>(s1, s2, bp, bn, a, &lt, &g);
    if (lane == 0) loss_acc += lt;
    const float t2 = 2.f * g, c2 = a.c_l2;
    for (int d = lane; d < D; d += 32) {
      const float u = ur[d], p = pr[d], n = nr[d];
      float gu, gp, gn;
      if (KIND == ORX_PAIR_BPR) {
        gu = g * (p - n) + c2 * u;
        gp = g * u + c2 * p;
        gn = -g * u + c2 * n;
      } else {
        gu = t2 * (n - p) + c2 * u;
        gp = t2 * (p - u) + c2 * p;
        gn = t2 * (u - n) + c2 * n;
      }
      float s0v = 0.f, s1v = 0.f;
      if (fu) {
        const int64_t o = (int64_t)uu * D + d;
        if (S0) s0v = a.Us0[o];
        if (S1) s1v = a.Us1[o];
        ur[d] = orx_apply<OPT>(u, gu, s0v, s1v, a.opt);
        if (S0) a.Us0[o] = s0v;
        if (S1) a.Us1[o] = s1v;
      } else {
        atomicAdd(a.gu + (int64_t)du * D + d, gu);
      }
      if (fp) {
        const int64_t o = (int64_t)pp * D + d;
        if (S0) s0v = a.Is0[o];
        if (S1) s1v = a.Is1[o];
        pr[d] = orx_apply<OPT>(p, gp, s0v, s1v, a.opt);
        if (S0) a.Is0[o] = s0v;
        if (S1) a.Is1[o] = s1v;
      } else {
        atomicAdd(a.gi + (int64_t)dp * D + d, gp);
      }
      if (fn) {
        const int64_t o = (int64_t)nn * D + d;
        if (S0) s0v = a.Is0[o];
        if (S1) s1v = a.Is1[o];
        nr[d] = orx_apply<OPT>(n, gn, s0v, s1v, a.opt);
        if (S0) a.Is0[o] = s0v;
        if (S1) a.Is1[o] = s1v;
      } else {
        atomicAdd(a.gi + (int64_t)dn * D + d, gn);
      }
    }
    if (lane == 0) {
      const float gbias = (KIND == ORX_PAIR_BPR) ? g : -g;
      float s0v = 0.f, s1v = 0.f;
      if (fp) {
        if (S0) s0v = a.Bs0[pp];
        if (S1) s1v = a.Bs1[pp];
        a.Bv[pp] = orx_apply<OPT>(bp, gbias, s0v, s1v, a.opt);
        if (S0) a.Bs0[pp] = s0v;
        if (S1) a.Bs1[pp] = s1v;
      } else {
        atomicAdd(a.gb + dp, gbias);
      }
      if (fn) {
        if (S0) s0v = a.Bs0[nn];
        if (S1) s1v = a.Bs1[nn];
        a.Bv[nn] = orx_apply<OPT>(bn, -gbias, s0v, s1v, a.opt);
        if (S0) a.Bs0[nn] = s0v;
        if (S1) a.Bs1[nn] = s1v;
      } else {
        atomicAdd(a.gb + dn, -gbias);
      }
      if (a.g_out) a.g_out[t] = g;
    }
  }
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
    a.partials[2 * blockIdx.x] = l;
    a.partials[2 * blockIdx.x + 1] = q;
  }
}

// ---------------------------------------------------------------------------------------
// ADAM_DENSE sweep: Keras-2.0 Adam on IndexedSlices touches EVERY row ("K12").
// One warp per table row; the row's summed gradient comes from the staging buffer via the hash.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_adam_sweep(float* var, float* m, float* v, int64_t rows, int D, OrxHash h,
                                                    const float* gstage, OrxOptDev o) {
  const int lane = threadIdx.x & 31;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += nw) {
    int d = -1;
    uint32_t c = 0;
    if (lane == 0) c = orx_hash_find(h, (int32_t)r, &d);
    c = __shfl_sync(ORX_FULL, c, 0);
    d = __shfl_sync(ORX_FULL, d, 0);
    for (int e = lane; e < D; e += 32) {
      const int64_t off = r * D + e;
      const float g = c ? gstage[(int64_t)d * D + e] : 0.f;
      const float mm = o.beta1 * m[off] + (1.f - o.beta1) * g;
      const float vv = o.beta2 * v[off] + (1.f - o.beta2) * g * g;
      m[off] = mm;
      v[off] = vv;
      var[off] = var[off] - o.lr * mm / (sqrtf(vv) + o.eps);
    }
  }
}

// ---------------------------------------------------------------------------------------
// tail: staged rows -> optimizer (once per unique row), zero staging, clear hash, reduce loss
// ---------------------------------------------------------------------------------------

template <int OPT>
__global__ void __launch_bounds__(256) k_sparse_tail(const TailArgs a) {
  constexpr bool S0 = (OPT == ORX_OPT_ADAGRAD || OPT == ORX_OPT_ADAM_LAZY);
  constexpr bool S1 = (OPT == ORX_OPT_ADAM_LAZY);
  constexpr bool ZERO_ONLY = (OPT == ORX_OPT_ADAM_DENSE);
  const int lane = threadIdx.x & 31;
  const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int nu = a.counters[0], ni = a.counters[1], nbad = a.counters[3];
  const int D = a.D;
  for (int r = gwarp; r < nu + ni; r += nwarps) {
    const bool is_u = r < nu;
    const int d = is_u ? r : r - nu;
    const int id = is_u ? a.hu.did[d] : a.hi.did[d];
    float* G = (is_u ? a.gu : a.gi) + (int64_t)d * D;
    float* W = (is_u ? a.U : a.I) + (int64_t)id * D;
    float* P0 = (is_u ? a.Us0 : a.Is0) + (int64_t)id * D;
    float* P1 = (is_u ? a.Us1 : a.Is1) + (int64_t)id * D;
    if ((D & 3) == 0) {  // 128-bit path
      for (int e = lane * 4; e < D; e += 128) {
        const float4 g = *reinterpret_cast<const float4*>(G + e);
        if (!ZERO_ONLY) {
          float4 w = *reinterpret_cast<const float4*>(W + e);
          float4 s0v = S0 ? *reinterpret_cast<const float4*>(P0 + e) : make_float4(0.f, 0.f, 0.f, 0.f);
          float4 s1v = S1 ? *reinterpret_cast<const float4*>(P1 + e) : make_float4(0.f, 0.f, 0.f, 0.f);
          *reinterpret_cast<float4*>(W + e) = orx_apply4<OPT>(w, g, s0v, s1v, a.opt);
          if (S0) *reinterpret_cast<float4*>(P0 + e) = s0v;
          if (S1) *reinterpret_cast<float4*>(P1 + e) = s1v;
        }
        *reinterpret_cast<float4*>(G + e) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
      for (int e = lane; e < D; e += 32) {
        if (!ZERO_ONLY) {
          float s0v = S0 ? P0[e] : 0.f, s1v = S1 ? P1[e] : 0.f;
          W[e] = orx_apply<OPT>(W[e], G[e], s0v, s1v, a.opt);
          if (S0) P0[e] = s0v;
          if (S1) P1[e] = s1v;
        }
        G[e] = 0.f;
      }
    }
    if (!is_u && lane == 0) {
      if (!ZERO_ONLY) {
        float s0v = S0 ? a.Bs0[id] : 0.f, s1v = S1 ? a.Bs1[id] : 0.f;
        a.Bv[id] = orx_apply<OPT>(a.Bv[id], a.gb[d], s0v, s1v, a.opt);
        if (S0) a.Bs0[id] = s0v;
        if (S1) a.Bs1[id] = s1v;
      }
      a.gb[d] = 0.f;
    }
  }
  // (the hash tables are not cleared: the next step uses a new epoch)

  __shared__ double sh[2][256];
  __shared__ bool last;
  if (blockIdx.x == 0) {
    // deterministic loss reduction (fixed order, double accumulation)
    double l = 0.0, q = 0.0;
    for (int i = threadIdx.x; i < a.n_partials; i += blockDim.x) {
      l += (double)a.partials[2 * i];
      q += (double)a.partials[2 * i + 1];
    }
    if (a.W)  // GMF: l2_loss also holds 0.5*sum(w^2) of the PRE-step weight (gmf.py:31-32)
      for (int e = threadIdx.x; e < D; e += blockDim.x) q += (double)a.W[e] * (double)a.W[e];
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
    if (threadIdx.x == 0) {
      a.out4[0] = (float)(sh[0][0] * (double)a.loss_scale);
      a.out4[1] = (float)(0.5 * sh[1][0]);
      a.out4[2] = (float)nbad;
      a.out4[3] = (float)(nu + ni);
    }
    // GMF dense weight: grad = gw + c_l2*w  (gmf.py:31-32), Keras dense apply
    if (a.W) {
      for (int e = threadIdx.x; e < D; e += blockDim.x) {
        const float g = a.gw[e] + a.c_l2 * a.W[e];
        if (ZERO_ONLY) {  // ADAM_DENSE: dense Adam
          const float mm = a.opt.beta1 * a.Ws0[e] + (1.f - a.opt.beta1) * g;
          const float vv = a.opt.beta2 * a.Ws1[e] + (1.f - a.opt.beta2) * g * g;
          a.Ws0[e] = mm;
          a.Ws1[e] = vv;
          a.W[e] = a.W[e] - a.opt.lr * mm / (sqrtf(vv) + a.opt.eps);
        } else {
          float s0v = S0 ? a.Ws0[e] : 0.f, s1v = S1 ? a.Ws1[e] : 0.f;
          a.W[e] = orx_apply<OPT>(a.W[e], g, s0v, s1v, a.opt);
          if (S0) a.Ws0[e] = s0v;
          if (S1) a.Ws1[e] = s1v;
        }
        a.gw[e] = 0.f;
      }
    }
  }
  // last block to finish resets the counters (every block has read them by then)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const int tk = atomicAdd(a.counters + 2, 1);
    last = (tk == (int)gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x < 4) a.counters[threadIdx.x] = 0;
}

int orx_launch_tail(orx_ctx* c, const TailArgs& ta, int opt_kind, cudaStream_t st) {
  const int grid = c->num_sms * 4;  // ~1-2 staged rows per warp: the tail is a latency chain, not bandwidth
  switch (opt_kind) {
    case ORX_OPT_SGD: k_sparse_tail<ORX_OPT_SGD><<<grid, 256, 0, st>>>(ta); break;
    case ORX_OPT_ADAGRAD: k_sparse_tail<ORX_OPT_ADAGRAD><<<grid, 256, 0, st>>>(ta); break;
    case ORX_OPT_ADAM_LAZY: k_sparse_tail<ORX_OPT_ADAM_LAZY><<<grid, 256, 0, st>>>(ta); break;
    default: k_sparse_tail<ORX_OPT_ADAM_DENSE><<<grid, 256, 0, st>>>(ta); break;
  }
  ORX_LAUNCH_CHECK();
  return ORX_OK;
}

int orx_launch_adam_sweep(orx_ctx* c, float* var, float* m, float* v, int64_t rows, int D, const OrxHash& h,
                          const float* gstage, const OrxOptDev& o, cudaStream_t st) {
  int64_t blocks = (rows + 7) / 8;
  const int64_t cap = (int64_t)c->num_sms * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  k_adam_sweep<<<(int)blocks, 256, 0, st>>>(var, m, v, rows, D, h, gstage, o);
  ORX_LAUNCH_CHECK();
  return ORX_OK;
}

// ---------------------------------------------------------------------------------------
// host dispatch
// ---------------------------------------------------------------------------------------
// Tuning variants of the D=128 kernel, selected with ORX_PAIR_VARIANT (A/B on the GPU):
//   0: 2 CTAs/SM, register double-buffer             1: 3 CTAs/SM, double-buffer
//   2 (default, fastest in profiles r1b/r1c): 4 CTAs/SM, single buffer   3: 3 CTAs/SM, single buffer
//   4/5/6: cp.async shared-memory ring, (stages, CTAs/SM) = (4,2) / (3,3) / (2,4)
static int pair_variant() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("ORX_PAIR_VARIANT");
    v = e ? atoi(e) : 2;
  }
  return v;
}

template <int KIND, int OPT>
static int launch_pair_step_kind_opt(const PairArgs& pa, int n_warps_hint, cudaStream_t st, int* n_partials) {
  const int B = pa.B;
  constexpr bool LAZY = (OPT == ORX_OPT_ADAM_LAZY);  // 9 rows per triplet: no register double-buffer
  auto go = [&](auto kern, int ch) {
    const int nw = (B + ch - 1) / ch;
    const int blocks = (nw + 7) / 8;
    *n_partials = blocks;
    kern<<<blocks, 256, 0, st>>>(pa);
  };
  auto go_async = [&](auto kern, int ch, int stages) {
    constexpr int NRr = 3 + ((OPT == ORX_OPT_ADAGRAD || OPT == ORX_OPT_ADAM_LAZY) ? 3 : 0) + (LAZY ? 3 : 0);
    const size_t smem = (size_t)8 * stages * NRr * (pa.D / 128 > 0 ? pa.D / 128 : 1) * 512;
    const int nw = (B + ch - 1) / ch;
    const int blocks = (nw + 7) / 8;
    *n_partials = blocks;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<blocks, 256, smem, st>>>(pa);
  };
  (void)n_warps_hint;
  switch (pa.D) {
    case 32: go(k_pair_step<KIND, OPT, 32, 8, 2, !LAZY>, 8); break;
    case 64: go(k_pair_step<KIND, OPT, 64, 8, 2, !LAZY>, 8); break;
    case 128:
      switch (LAZY ? 3 : pair_variant()) {
        case 0: go(k_pair_step<KIND, OPT, 128, 8, 2, !LAZY>, 8); break;
        case 1: go(k_pair_step<KIND, OPT, 128, 8, 3, !LAZY>, 8); break;
        case 3: go(k_pair_step<KIND, OPT, 128, 8, 3, false>, 8); break;
        case 4: go_async(k_pair_step_async<KIND, OPT, 128, 8, 2, 4>, 8, 4); break;
        case 5: go_async(k_pair_step_async<KIND, OPT, 128, 8, 3, 3>, 8, 3); break;
        case 6: go_async(k_pair_step_async<KIND, OPT, 128, 8, 4, 2>, 8, 2); break;
        case 7:   // last arriver applies (pa.out4 is set only when the caller built the counting index), 4 CTAs/SM
          if (pa.out4) go(k_pair_step<KIND, OPT, 128, 8, 4, false, true>, 8);
          else go(k_pair_step<KIND, OPT, 128, 8, 4, false>, 8);
          break;
        case 8:   // same, 3 CTAs/SM (85 registers)
          if (pa.out4) go(k_pair_step<KIND, OPT, 128, 8, 3, false, true>, 8);
          else go(k_pair_step<KIND, OPT, 128, 8, 4, false>, 8);
          break;
        default: go(k_pair_step<KIND, OPT, 128, 8, 4, false>, 8); break;
      }
      break;
    case 256: go(k_pair_step<KIND, OPT, 256, 8, 2, !LAZY>, 8); break;
    default: go(k_pair_step_generic<KIND, OPT>, 8); break;
  }
  ORX_LAUNCH_CHECK();
  return ORX_OK;
}

template <int KIND>
static int launch_pair_step_kind(const PairArgs& pa, int opt_kind, cudaStream_t st, int* n_partials) {
  switch (opt_kind) {
    case ORX_OPT_SGD: return launch_pair_step_kind_opt<KIND, ORX_OPT_SGD>(pa, 0, st, n_partials);
    case ORX_OPT_ADAGRAD: return launch_pair_step_kind_opt<KIND, ORX_OPT_ADAGRAD>(pa, 0, st, n_partials);
    case ORX_OPT_ADAM_LAZY: return launch_pair_step_kind_opt<KIND, ORX_OPT_ADAM_LAZY>(pa, 0, st, n_partials);
    case ORX_OPT_ADAM_DENSE: return launch_pair_step_kind_opt<KIND, ORX_OPT_ADAM_DENSE>(pa, 0, st, n_partials);
  }
  orx_set_error("unknown optimizer kind %d", opt_kind);
  return ORX_ERR_INVALID;
}

static int check_tables(const orx_table_t* user, const orx_table_t* item, const orx_table_t* bias, int opt_kind) {
  ORX_REQUIRE(user && item && bias, "null table");
  ORX_REQUIRE(user->var && item->var && bias->var, "null table storage");
  ORX_REQUIRE(user->dim == item->dim && user->dim > 0, "user/item dims must match and be positive");
  ORX_REQUIRE(bias->dim == 1 && bias->rows == item->rows, "item_bias must be [item.rows, 1]");
  ORX_REQUIRE(user->rows > 0 && item->rows > 0 && user->rows <= 0x7fffffffLL && item->rows <= 0x7fffffffLL,
              "row counts must fit int32 ids");
  if (opt_kind != ORX_OPT_SGD) ORX_REQUIRE(user->s0 && item->s0 && bias->s0, "optimizer slot s0 missing");
  if (opt_kind == ORX_OPT_ADAM_LAZY || opt_kind == ORX_OPT_ADAM_DENSE)
    ORX_REQUIRE(user->s1 && item->s1 && bias->s1, "optimizer slot s1 missing");
  return ORX_OK;
}

// ORX_FUSED=0 selects the three-launch path (index / step / tail); default: fused persistent kernel where an
// instance exists (D = 128, SGD / Adagrad).
static bool pair_fused_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("ORX_FUSED");
    v = e ? atoi(e) : 0;   // opt-in: steady-state it is no faster than the 3-launch path (profiles/README.md r1g)
  }
  return v != 0;
}

// index_stream != nullptr (experimental, LA variants only): the index is built into the context's SECOND hash set on
// that stream (behind whatever the caller queued there, e.g. the id upload), `st` waits for it, and the previous step --
// which uses the first set -- may still be running on `st` meanwhile.  `set` alternates per call.
static int pairwise_step_impl(orx_ctx* c, int kind, const orx_table_t* user, const orx_table_t* item,
                              const orx_table_t* bias, const int32_t* uid, const int32_t* pid, const int32_t* nid,
                              int B, float margin, float c_loss, float c_l2, const orx_opt_t* opt, float* out4,
                              cudaStream_t st, int set = 0, cudaStream_t index_stream = nullptr,
                              cudaEvent_t index_done = nullptr) {
  ORX_REQUIRE(kind == ORX_PAIR_BPR || kind == ORX_PAIR_UCML, "unknown pairwise kind");
  ORX_REQUIRE(opt != nullptr && out4 != nullptr, "null opt/out");
  ORX_REQUIRE(opt->kind >= ORX_OPT_SGD && opt->kind <= ORX_OPT_ADAM_DENSE, "unknown optimizer kind");
  ORX_REQUIRE(B > 0 && uid && pid && nid, "empty batch or null ids");
  int rc = check_tables(user, item, bias, opt->kind);
  if (rc) return rc;
  const int D = user->dim;
  if ((rc = orx_ensure_workspace(c, B, D, opt->kind == ORX_OPT_ADAM_DENSE))) return rc;
  const bool dense = opt->kind == ORX_OPT_ADAM_DENSE;
  PairArgs pa;
  pa.U = user->var; pa.Us0 = user->s0; pa.Us1 = user->s1;
  pa.I = item->var; pa.Is0 = item->s0; pa.Is1 = item->s1;
  pa.Bv = bias->var; pa.Bs0 = bias->s0; pa.Bs1 = bias->s1;
  pa.rowsU = user->rows; pa.rowsI = item->rows; pa.D = D;
  pa.uid = uid; pa.pid = pid; pa.nid = nid; pa.B = B;
  pa.margin = margin; pa.c_loss = c_loss; pa.c_l2 = c_l2; pa.inv_B = 1.0f / (float)B;
  pa.opt = orx_opt_to_dev(opt);
  pa.gu = c->gu; pa.gi = c->gi; pa.gb = c->gb;
  pa.g_out = nullptr;
  if ((rc = orx_ensure_partials(c, (B + 63) / 64 + 8 > c->num_sms ? (B + 63) / 64 + 8 : c->num_sms, st))) return rc;
  pa.partials = c->partials;
  orx_prof_mark(c, 0, st);
  if (pair_fused_enabled()) {   // one persistent cooperative launch (orx_pair_fused.cu)
    orx_new_epoch(c);
    pa.hu = c->hu; pa.hi = c->hi;
    orx_prof_mark(c, 1, st);     // all phases live inside one kernel: it is reported as the "step" phase
    rc = orx_launch_pair_fused(c, kind, opt->kind, pa, kind == ORX_PAIR_BPR ? pa.inv_B : 1.0f, out4, st);
    if (rc == ORX_OK) {
      orx_prof_mark(c, 2, st);
      orx_prof_mark(c, 3, st);
      orx_prof_next(c);
      return ORX_OK;
    }
    if (rc != ORX_ERR_UNSUPPORTED) return rc;
  }
  // experimental "last arriver applies" variants (D = 128, not Keras-dense Adam): reference-counted index, no tail
  const bool la = D == 128 && !dense && opt->kind != ORX_OPT_ADAM_LAZY && (pair_variant() == 7 || pair_variant() == 8);
  const bool second = la && set == 1;
  if (second && (rc = orx_ensure_second_index(c))) return rc;
  OrxHash& HU = second ? c->hu_b : c->hu;
  OrxHash& HI = second ? c->hi_b : c->hi;
  int32_t* ctr = second ? c->counters + 4 : c->counters;
  pa.out4 = la ? out4 : nullptr; pa.counters = ctr; pa.loss_scale = (kind == ORX_PAIR_BPR) ? pa.inv_B : 1.0f;
  cudaStream_t ist = (la && index_stream) ? index_stream : st;
  if (index_stream && !la) {   // no overlap without the LA variants: just order `st` behind the caller's id upload
    ORX_CUDA(cudaEventRecord(index_done, index_stream));
    ORX_CUDA(cudaStreamWaitEvent(st, index_done, 0));
  }
  if ((rc = orx_launch_index_build_on(c, HU, HI, ctr, uid, user->rows, B, pid, nid, item->rows, B,
                                      la ? 3 : (dense ? 1 : 0), ist)))
    return rc;
  if (index_stream && la) {    // upload + index ran on index_stream: `st` (the step kernel) waits for both
    ORX_CUDA(cudaEventRecord(index_done, index_stream));
    ORX_CUDA(cudaStreamWaitEvent(st, index_done, 0));
  }
  orx_prof_mark(c, 1, st);
  pa.hu = HU; pa.hi = HI;
  int n_partials = 0;
  rc = (kind == ORX_PAIR_BPR) ? launch_pair_step_kind<ORX_PAIR_BPR>(pa, opt->kind, st, &n_partials)
                              : launch_pair_step_kind<ORX_PAIR_UCML>(pa, opt->kind, st, &n_partials);
  if (rc) return rc;
  orx_prof_mark(c, 2, st);
  if (la) {   // the step kernel's own epilogue wrote out4 and reset the counters
    orx_prof_mark(c, 3, st);
    orx_prof_next(c);
    return ORX_OK;
  }
  if (dense) {
    if ((rc = orx_launch_adam_sweep(c, user->var, user->s0, user->s1, user->rows, D, c->hu, c->gu, pa.opt, st))) return rc;
    if ((rc = orx_launch_adam_sweep(c, item->var, item->s0, item->s1, item->rows, D, c->hi, c->gi, pa.opt, st))) return rc;
    if ((rc = orx_launch_adam_sweep(c, bias->var, bias->s0, bias->s1, bias->rows, 1, c->hi, c->gb, pa.opt, st))) return rc;
  }
  TailArgs ta;
  ta.U = user->var; ta.Us0 = user->s0; ta.Us1 = user->s1;
  ta.I = item->var; ta.Is0 = item->s0; ta.Is1 = item->s1;
  ta.Bv = bias->var; ta.Bs0 = bias->s0; ta.Bs1 = bias->s1;
  ta.D = D; ta.opt = pa.opt; ta.hu = c->hu; ta.hi = c->hi;
  ta.gu = c->gu; ta.gi = c->gi; ta.gb = c->gb;
  ta.partials = c->partials; ta.n_partials = n_partials;
  ta.loss_scale = (kind == ORX_PAIR_BPR) ? pa.inv_B : 1.0f;
  ta.counters = c->counters; ta.out4 = out4;
  ta.W = ta.Ws0 = ta.Ws1 = ta.gw = nullptr; ta.c_l2 = c_l2;
  rc = orx_launch_tail(c, ta, opt->kind, st);
  orx_prof_mark(c, 3, st);
  orx_prof_next(c);
  return rc;
}

extern "C" int orx_pairwise_step(orx_handle_t h, int32_t kind, const orx_table_t* user, const orx_table_t* item,
                                 const orx_table_t* item_bias, const int32_t* uid, const int32_t* pid,
                                 const int32_t* nid, int32_t B, float margin, float c_loss, float c_l2,
                                 const orx_opt_t* opt, float* out4, orx_stream_t s) {
  ORX_REQUIRE(h != nullptr, "null handle");
  ORX_CUDA(cudaSetDevice(h->device));
  return pairwise_step_impl(h, kind, user, item, item_bias, uid, pid, nid, B, margin, c_loss, c_l2, opt, out4,
                            (cudaStream_t)s);
}

extern "C" int orx_pairwise_step_host(orx_handle_t h, int32_t kind, const orx_table_t* user, const orx_table_t* item,
                                      const orx_table_t* item_bias, const int32_t* uid_host, const int32_t* pid_host,
                                      const int32_t* nid_host, int32_t B, float margin, float c_loss, float c_l2,
                                      const orx_opt_t* opt, float* out4_host, orx_stream_t s) {
  ORX_REQUIRE(h != nullptr, "null handle");
  ORX_REQUIRE(B > 0 && uid_host && pid_host && nid_host && out4_host, "empty batch or null host buffers");
  ORX_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)s;
  int rc = orx_ensure_stage(h, 3 * (int64_t)B);
  if (rc) return rc;
  const uint32_t f = (h->stage_flip++) & 1u;
  int32_t* ids = h->ids_stage[f];
  static int use_copy_stream = -1;
  if (use_copy_stream < 0) {
    const char* e = getenv("ORX_HOST_COPY_STREAM");
    use_copy_stream = (e && atoi(e)) ? 1 : 0;
  }
  if (use_copy_stream) {
    // experimental: the upload of this batch runs on a copy stream, i.e. under the previous step's kernels when the
    // caller enqueues ahead.  ids_stage[f] was last read by the step two calls ago: wait for that step first.
    if (!h->copy_stream) {
      ORX_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
      for (int i = 0; i < 2; ++i) {
        ORX_CUDA(cudaEventCreateWithFlags(&h->copy_done[i], cudaEventDisableTiming));
        ORX_CUDA(cudaEventCreateWithFlags(&h->stage_free[i], cudaEventDisableTiming));
        h->stage_free_valid[i] = 0;
      }
    }
    cudaStream_t cs = h->copy_stream;
    if (h->stage_free_valid[f]) ORX_CUDA(cudaStreamWaitEvent(cs, h->stage_free[f], 0));
    if (pid_host == uid_host + B && nid_host == pid_host + B) {   // one contiguous (uid | pid | nid) host block
      ORX_CUDA(cudaMemcpyAsync(ids, uid_host, sizeof(int32_t) * 3 * (size_t)B, cudaMemcpyHostToDevice, cs));
    } else {
      ORX_CUDA(cudaMemcpyAsync(ids, uid_host, sizeof(int32_t) * B, cudaMemcpyHostToDevice, cs));
      ORX_CUDA(cudaMemcpyAsync(ids + B, pid_host, sizeof(int32_t) * B, cudaMemcpyHostToDevice, cs));
      ORX_CUDA(cudaMemcpyAsync(ids + 2 * (int64_t)B, nid_host, sizeof(int32_t) * B, cudaMemcpyHostToDevice, cs));
    }
    static int overlap_index = -1;
    if (overlap_index < 0) {
      const char* e = getenv("ORX_OVERLAP_INDEX");
      overlap_index = (e && atoi(e)) ? 1 : 0;
    }
    if (overlap_index) {
      // experimental: the index build follows the upload on the copy stream into hash set f, so with the LA variants
      // (no tail) it runs beside the previous step's k_pair_step, which uses the other set
      rc = pairwise_step_impl(h, kind, user, item, item_bias, ids, ids + B, ids + 2 * (int64_t)B, B, margin, c_loss,
                              c_l2, opt, h->out_stage[f], st, (int)f, cs, h->copy_done[f]);
      if (rc) return rc;
      ORX_CUDA(cudaEventRecord(h->stage_free[f], st));
      h->stage_free_valid[f] = 1;
      ORX_CUDA(cudaMemcpyAsync(out4_host, h->out_stage[f], sizeof(float) * 4, cudaMemcpyDeviceToHost, st));
      return ORX_OK;
    }
    ORX_CUDA(cudaEventRecord(h->copy_done[f], cs));
    ORX_CUDA(cudaStreamWaitEvent(st, h->copy_done[f], 0));
  } else {
    ORX_CUDA(cudaMemcpyAsync(ids, uid_host, sizeof(int32_t) * B, cudaMemcpyHostToDevice, st));
    ORX_CUDA(cudaMemcpyAsync(ids + B, pid_host, sizeof(int32_t) * B, cudaMemcpyHostToDevice, st));
    ORX_CUDA(cudaMemcpyAsync(ids + 2 * (int64_t)B, nid_host, sizeof(int32_t) * B, cudaMemcpyHostToDevice, st));
  }
  rc = pairwise_step_impl(h, kind, user, item, item_bias, ids, ids + B, ids + 2 * (int64_t)B, B, margin, c_loss, c_l2,
                          opt, h->out_stage[f], st);
  if (rc) return rc;
  if (use_copy_stream) {
    ORX_CUDA(cudaEventRecord(h->stage_free[f], st));
    h->stage_free_valid[f] = 1;
  }
  ORX_CUDA(cudaMemcpyAsync(out4_host, h->out_stage[f], sizeof(float) * 4, cudaMemcpyDeviceToHost, st));
  return ORX_OK;
}

// ---------------------------------------------------------------------------------------
// forward only / explicit gradients (un-fused; parity tests and the lazy-handle fallback)
// ---------------------------------------------------------------------------------------
struct PairGradArgs {
  const float *U, *I, *Bv;
  int64_t rowsU, rowsI;
  int D;
  const int32_t *uid, *pid, *nid;
  int B;
  float margin, c_loss, c_l2, inv_B;
  float *d_user, *d_pos, *d_neg, *d_bp, *d_bn, *g_out;
  float* partials;
  int slots;  // 1: outputs are indexed by the lookup's row (compact sharded form) instead of by triplet
  int64_t ld;  // row stride of U / I / outputs in floats (0 => D); ld > D: item bias lives in column D of the row
};

template <int KIND>
__global__ void __launch_bounds__(256) k_pair_fwd_grad(const PairGradArgs a) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int D = a.D;
  const int64_t ld = a.ld ? a.ld : D;
  const bool bias_in_row = a.ld > D;
  float loss_acc = 0.f, l2_acc = 0.f;
  PairArgs sa;  // only the scalar fields pair_score reads
  sa.margin = a.margin;
  sa.c_loss = a.c_loss;
  sa.inv_B = a.inv_B;
  for (int j = 0; j < 8; ++j) {
    const int t = warp * 8 + j;
    if (t >= a.B) break;
    const int uu = a.uid[t], pp = a.pid[t], nn = a.nid[t];
    const bool ok = uu >= 0 && uu < a.rowsU && pp >= 0 && pp < a.rowsI && nn >= 0 && nn < a.rowsI;
    float s1 = 0.f, s2 = 0.f, sq = 0.f, bp = 0.f, bn = 0.f, lt = 0.f, g = 0.f;
    const float* ur = a.U + (int64_t)uu * ld;
    const float* pr = a.I + (int64_t)pp * ld;
    const float* nr = a.I + (int64_t)nn * ld;
    if (ok) {
      for (int d = lane; d < D; d += 32) {
        const float u = ur[d], p = pr[d], n = nr[d];
        if (KIND == ORX_PAIR_BPR) {
          s1 += u * p;
          s2 += u * n;
        } else {
          s1 += (u - p) * (u - p);
          s2 += (u - n) * (u - n);
        }
        sq += u * u + p * p + n * n;
      }
      bp = bias_in_row ? pr[D] : a.Bv[pp];
      bn = bias_in_row ? nr[D] : a.Bv[nn];
    }
    l2_acc += sq;
    s1 = orx_group_sum<32>(s1);
    s2 = orx_group_sum<32>(s2);
    if (ok) pair_score<KIND>(s1, s2, bp, bn, sa, &lt, &g);
    if (lane == 0) loss_acc += lt;
    const float t2 = 2.f * g, c2 = a.c_l2;
    if (a.d_user || a.d_pos || a.d_neg) {
      for (int d = lane; d < D; d += 32) {
        float gu = 0.f, gp = 0.f, gn = 0.f;
        if (ok) {
          const float u = ur[d], p = pr[d], n = nr[d];
          if (KIND == ORX_PAIR_BPR) {
            gu = g * (p - n) + c2 * u;
            gp = g * u + c2 * p;
            gn = -g * u + c2 * n;
          } else {
            gu = t2 * (n - p) + c2 * u;
            gp = t2 * (p - u) + c2 * p;
            gn = t2 * (u - n) + c2 * n;
          }
        }
        if (a.slots) {
          if (ok) {
            a.d_user[(int64_t)uu * ld + d] = gu;
            a.d_pos[(int64_t)pp * ld + d] = gp;
            a.d_neg[(int64_t)nn * ld + d] = gn;
          }
        } else {
          const int64_t o = (int64_t)t * D + d;
          if (a.d_user) a.d_user[o] = gu;
          if (a.d_pos) a.d_pos[o] = gp;
          if (a.d_neg) a.d_neg[o] = gn;
        }
      }
    }
    if (lane == 0) {
      const float gbias = (KIND == ORX_PAIR_BPR) ? g : -g;
      if (a.slots && bias_in_row) {
        if (ok) {   // column D = bias gradient (items) / 0 (users); remaining padding columns = 0
          for (int64_t c = D; c < ld; ++c) {
            a.d_user[(int64_t)uu * ld + c] = 0.f;
            a.d_pos[(int64_t)pp * ld + c] = c == D ? gbias : 0.f;
            a.d_neg[(int64_t)nn * ld + c] = c == D ? -gbias : 0.f;
          }
        }
      } else if (a.slots) {
        if (ok) {
          a.d_bp[pp] = gbias;
          a.d_bn[nn] = -gbias;
        }
      } else {
        if (a.d_bp) a.d_bp[t] = gbias;
        if (a.d_bn) a.d_bn[t] = -gbias;
      }
      if (a.g_out) a.g_out[t] = g;
    }
  }
  loss_acc = orx_group_sum<32>(loss_acc);
  l2_acc = orx_group_sum<32>(l2_acc);
  if (lane == 0) {
    a.partials[2 * warp] = loss_acc;
    a.partials[2 * warp + 1] = l2_acc;
  }
}

__global__ void k_reduce_partials(const float* partials, int n, float loss_scale, float* out4) {
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
  if (threadIdx.x == 0) {
    out4[0] = (float)(sh[0][0] * (double)loss_scale);
    out4[1] = (float)(0.5 * sh[1][0]);
    out4[2] = 0.f;
    out4[3] = 0.f;
  }
}

int orx_launch_reduce_partials(const float* partials, int n, float loss_scale, float* out4, cudaStream_t st) {
  k_reduce_partials<<<1, 256, 0, st>>>(partials, n, loss_scale, out4);
  ORX_LAUNCH_CHECK();
  return ORX_OK;
}

int orx_ensure_partials(orx_ctx* c, int need, cudaStream_t st) {
  if (need <= c->cap_partials) return ORX_OK;
  ORX_CUDA(cudaStreamSynchronize(st));
  cudaFree(c->partials);
  c->partials = nullptr;
  c->cap_partials = need * 2;
  ORX_CUDA(cudaMalloc(&c->partials, sizeof(float) * 2 * (size_t)c->cap_partials));
  return ORX_OK;
}

static int pair_fwd_grad(orx_ctx* c, int kind, const orx_table_t* user, const orx_table_t* item,
                         const orx_table_t* bias, const int32_t* uid, const int32_t* pid, const int32_t* nid, int B,
                         float margin, float c_loss, float c_l2, float* d_user, float* d_pos, float* d_neg,
                         float* d_bp, float* d_bn, float* g_out, float* out4, cudaStream_t st) {
  ORX_REQUIRE(kind == ORX_PAIR_BPR || kind == ORX_PAIR_UCML, "unknown pairwise kind");
  ORX_REQUIRE(B > 0 && uid && pid && nid, "empty batch or null ids");
  int rc = check_tables(user, item, bias, ORX_OPT_SGD);
  if (rc) return rc;
  const int nw = (B + 7) / 8, blocks = (nw + 7) / 8;
  if ((rc = orx_ensure_partials(c, blocks * 8, st))) return rc;
  PairGradArgs a;
  a.U = user->var; a.I = item->var; a.Bv = bias->var;
  a.rowsU = user->rows; a.rowsI = item->rows; a.D = user->dim;
  a.uid = uid; a.pid = pid; a.nid = nid; a.B = B;
  a.margin = margin; a.c_loss = c_loss; a.c_l2 = c_l2; a.inv_B = 1.0f / (float)B;
  a.d_user = d_user; a.d_pos = d_pos; a.d_neg = d_neg; a.d_bp = d_bp; a.d_bn = d_bn; a.g_out = g_out;
  a.partials = c->partials; a.slots = 0; a.ld = 0;
  if (kind == ORX_PAIR_BPR) k_pair_fwd_grad<ORX_PAIR_BPR><<<blocks, 256, 0, st>>>(a);
  else k_pair_fwd_grad<ORX_PAIR_UCML><<<blocks, 256, 0, st>>>(a);
  ORX_LAUNCH_CHECK();
  if (out4) {
    return orx_launch_reduce_partials(c->partials, blocks * 8, kind == ORX_PAIR_BPR ? a.inv_B : 1.f, out4, st);
  }
  return ORX_OK;
}

extern "C" int orx_pairwise_fwd(orx_handle_t h, int32_t kind, const orx_table_t* user, const orx_table_t* item,
                                const orx_table_t* item_bias, const int32_t* uid, const int32_t* pid,
                                const int32_t* nid, int32_t B, float margin, float* out4, orx_stream_t s) {
  ORX_REQUIRE(h != nullptr && out4 != nullptr, "null handle/out");
  ORX_CUDA(cudaSetDevice(h->device));
  return pair_fwd_grad(h, kind, user, item, item_bias, uid, pid, nid, B, margin, 1.f, 1.f, nullptr, nullptr, nullptr,
                       nullptr, nullptr, nullptr, out4, (cudaStream_t)s);
}

extern "C" int orx_pairwise_grad(orx_handle_t h, int32_t kind, const orx_table_t* user, const orx_table_t* item,
                                 const orx_table_t* item_bias, const int32_t* uid, const int32_t* pid,
                                 const int32_t* nid, int32_t B, float margin, float c_loss, float c_l2, float* d_user,
                                 float* d_pos, float* d_neg, float* d_bp, float* d_bn, float* g_out, orx_stream_t s) {
  ORX_REQUIRE(h != nullptr, "null handle");
  ORX_CUDA(cudaSetDevice(h->device));
  return pair_fwd_grad(h, kind, user, item, item_bias, uid, pid, nid, B, margin, c_loss, c_l2, d_user, d_pos, d_neg,
                       d_bp, d_bn, g_out, nullptr, (cudaStream_t)s);
}

extern "C" int orx_pairwise_grad_slots(orx_handle_t h, int32_t kind, const float* user_rows, const float* item_rows,
                                       const float* bias_rows, int32_t dim, const int32_t* uslot, const int32_t* pslot,
                                       const int32_t* nslot, int32_t B, float margin, float c_loss, float c_l2,
                                       float inv_B, float* d_user_rows, float* d_item_rows, float* d_bias_rows,
                                       float* out4, orx_stream_t s) {
  ORX_REQUIRE(h != nullptr && user_rows && item_rows && bias_rows && uslot && pslot && nslot, "null input");
  ORX_REQUIRE(d_user_rows && d_item_rows && d_bias_rows && out4, "null output");
  ORX_REQUIRE(kind == ORX_PAIR_BPR || kind == ORX_PAIR_UCML, "unknown pairwise kind");
  ORX_REQUIRE(B > 0 && dim > 0, "bad sizes");
  ORX_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)s;
  const int nw = (B + 7) / 8, blocks = (nw + 7) / 8;
  int rc = orx_ensure_partials(h, blocks * 8, st);
  if (rc) return rc;
  PairGradArgs a;
  a.U = user_rows; a.I = item_rows; a.Bv = bias_rows;
  a.rowsU = B; a.rowsI = 2 * (int64_t)B; a.D = dim;
  a.uid = uslot; a.pid = pslot; a.nid = nslot; a.B = B;
  a.margin = margin; a.c_loss = c_loss; a.c_l2 = c_l2; a.inv_B = inv_B;
  a.d_user = d_user_rows; a.d_pos = d_item_rows; a.d_neg = d_item_rows; a.d_bp = d_bias_rows; a.d_bn = d_bias_rows;
  a.g_out = nullptr; a.partials = h->partials; a.slots = 1; a.ld = 0;
  if (kind == ORX_PAIR_BPR) k_pair_fwd_grad<ORX_PAIR_BPR><<<blocks, 256, 0, st>>>(a);
  else k_pair_fwd_grad<ORX_PAIR_UCML><<<blocks, 256, 0, st>>>(a);
  ORX_LAUNCH_CHECK();
  return orx_launch_reduce_partials(h->partials, blocks * 8, kind == ORX_PAIR_BPR ? inv_B : 1.f, out4, st);
}

extern "C" int orx_pairwise_grad_rows(orx_handle_t h, int32_t kind, const float* rows, int64_t ld, int32_t dim,
                                      const int32_t* uslot, const int32_t* pslot, const int32_t* nslot, int32_t B,
                                      float margin, float c_loss, float c_l2, float inv_B, float* d_rows, float* out4,
                                      orx_stream_t s) {
  ORX_REQUIRE(h != nullptr && rows && uslot && pslot && nslot && d_rows && out4, "null pointer");
  ORX_REQUIRE(kind == ORX_PAIR_BPR || kind == ORX_PAIR_UCML, "unknown pairwise kind");
  ORX_REQUIRE(B > 0 && dim > 0 && ld > dim, "bad sizes (ld must exceed dim: the bias lives in column dim)");
  ORX_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)s;
  const int nw = (B + 7) / 8, blocks = (nw + 7) / 8;
  int rc = orx_ensure_partials(h, blocks * 8, st);
  if (rc) return rc;
  PairGradArgs a;
  a.U = rows; a.I = rows; a.Bv = nullptr;
  a.rowsU = 3 * (int64_t)B; a.rowsI = 3 * (int64_t)B; a.D = dim;
  a.uid = uslot; a.pid = pslot; a.nid = nslot; a.B = B;
  a.margin = margin; a.c_loss = c_loss; a.c_l2 = c_l2; a.inv_B = inv_B;
  a.d_user = d_rows; a.d_pos = d_rows; a.d_neg = d_rows; a.d_bp = nullptr; a.d_bn = nullptr;
  a.g_out = nullptr; a.partials = h->partials; a.slots = 1; a.ld = ld;
  if (kind == ORX_PAIR_BPR) k_pair_fwd_grad<ORX_PAIR_BPR><<<blocks, 256, 0, st>>>(a);
  else k_pair_fwd_grad<ORX_PAIR_UCML><<<blocks, 256, 0, st>>>(a);
  ORX_LAUNCH_CHECK();
  return orx_launch_reduce_partials(h->partials, blocks * 8, kind == ORX_PAIR_BPR ? inv_B : 1.f, out4, st);
}
