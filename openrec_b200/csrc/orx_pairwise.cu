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
  int rc = orx_next_epoch(c, st);
  if (rc) return rc;
  int blocks = (n + 255) / 256;
  if (n_dev && blocks > c->num_sms * 8) blocks = c->num_sms * 8;
  k_index_build_strided<<<blocks, 256, 0, st>>>(c->hu, a, stride, rows, n, n_dev, stage_all ? 1 : 0, c->counters + 3);
  ORX_LAUNCH_CHECK();
  return ORX_OK;
}

int orx_launch_index_build(orx_ctx* c, const int32_t* a, int64_t rows_a, int32_t na, const int32_t* b0,
                           const int32_t* b1, int64_t rows_b, int32_t nb, int mode, cudaStream_t st) {
  const int total = na + (b1 ? 2 * nb : nb);
  if (total <= 0) return ORX_OK;
  int rc = orx_next_epoch(c, st);
  if (rc) return rc;
  k_index_build<<<(total + 255) / 256, 256, 0, st>>>(c->hu, c->hi, a, rows_a, na, b0, b1, rows_b, nb, mode, c->counters + 3);
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
  float bp, bn;
};

// flags: bit0 triplet valid, bit1/2/3 user/pos/neg row owned by this triplet (fast path)
//
// One warp owns CH consecutive triplets.  Order of issue inside a warp (latency first):
//   ids (coalesced, lanes < CH) -> variable rows of the first one/two triplet groups (they need only
//   the ids) -> hash probes + bias loads (lanes < CH, overlap the row loads) -> slot rows of the first
//   groups -> steady state: process one register buffer while the other's 128-bit loads are in flight.
template <int KIND, int OPT, int D, int CH, int MINB, bool PIPE>
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
  orx_pdl_wait();

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

  // optimizer-slot rows + staging indices: need the probe results
  auto load_slots = [&](int j, Regs& r) {
    const int src = j + grp;
    r.fl = __shfl_sync(ORX_FULL, flags, src);
    r.du = __shfl_sync(ORX_FULL, du, src);
    r.dp = __shfl_sync(ORX_FULL, dp, src);
    r.dn = __shfl_sync(ORX_FULL, dn, src);
    r.bp = __shfl_sync(ORX_FULL, bp, src);
    r.bn = __shfl_sync(ORX_FULL, bn, src);
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

  orx_pdl_trigger();
  // ---- item_bias: lane-parallel, one lane per triplet of the chunk
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
  orx_pdl_wait();
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
  orx_pdl_wait();
  const int nu = a.counters[0], ni = a.counters[1], nbad = a.counters[3];
  const int D = a.D;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  // The tail is a chain of dependent round trips (counters -> row id -> rows) over a few thousand rows, i.e. latency,
  // not bandwidth.  A warp therefore takes FOUR staged rows at once, eight lanes per row (a quarter-warp still covers
  // 128 contiguous bytes per access), and issues all of a row's loads before the first use: 12 independent 128-bit
  // loads per lane in flight at D = 128.
  const int sub = lane >> 3, sl = lane & 7;
  for (int r0 = gwarp * 4; r0 < nu + ni; r0 += nwarps * 4) {
    const int r = r0 + sub;
    const bool on = r < nu + ni;
    const bool is_u = on && r < nu;
    const int d = on ? (is_u ? r : r - nu) : 0;
    const int id = on ? (is_u ? a.hu.did[d] : a.hi.did[d]) : 0;
    float* G = (is_u ? a.gu : a.gi) + (int64_t)d * D;
    float* W = (is_u ? a.U : a.I) + (int64_t)id * D;
    float* P0 = (is_u ? a.Us0 : a.Is0) + (int64_t)id * D;
    float* P1 = (is_u ? a.Us1 : a.Is1) + (int64_t)id * D;
    if ((D & 3) == 0) {  // 128-bit path: float4 index sl + 8k
      const int nq = D >> 2;
      for (int e0 = 0; e0 < nq; e0 += 32) {
        float4 g[4], w[4], s0v[4], s1v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int e = e0 + sl + 8 * k;
          const bool ld = on && e < nq;
          g[k] = ld ? __ldcg(reinterpret_cast<const float4*>(G) + e) : z4;
          w[k] = (ld && !ZERO_ONLY) ? __ldcg(reinterpret_cast<const float4*>(W) + e) : z4;
          s0v[k] = (ld && S0 && !ZERO_ONLY) ? __ldcg(reinterpret_cast<const float4*>(P0) + e) : z4;
          s1v[k] = (ld && S1 && !ZERO_ONLY) ? __ldcg(reinterpret_cast<const float4*>(P1) + e) : z4;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int e = e0 + sl + 8 * k;
          if (!on || e >= nq) continue;
          if (!ZERO_ONLY) {
            __stcg(reinterpret_cast<float4*>(W) + e, orx_apply4<OPT>(w[k], g[k], s0v[k], s1v[k], a.opt));
            if (S0) __stcg(reinterpret_cast<float4*>(P0) + e, s0v[k]);
            if (S1) __stcg(reinterpret_cast<float4*>(P1) + e, s1v[k]);
          }
          __stcg(reinterpret_cast<float4*>(G) + e, z4);
        }
      }
    } else if (on) {
      for (int e = sl; e < D; e += 8) {
        if (!ZERO_ONLY) {
          float s0v = S0 ? P0[e] : 0.f, s1v = S1 ? P1[e] : 0.f;
          W[e] = orx_apply<OPT>(W[e], G[e], s0v, s1v, a.opt);
          if (S0) P0[e] = s0v;
          if (S1) P1[e] = s1v;
        }
        G[e] = 0.f;
      }
    }
    if (on && !is_u && sl == 0) {     // the item bias of the staged row
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
    case ORX_OPT_SGD: orx_launch_pdl(k_sparse_tail<ORX_OPT_SGD>, dim3(grid), dim3(256), 0, st, ta); break;
    case ORX_OPT_ADAGRAD: orx_launch_pdl(k_sparse_tail<ORX_OPT_ADAGRAD>, dim3(grid), dim3(256), 0, st, ta); break;
    case ORX_OPT_ADAM_LAZY: orx_launch_pdl(k_sparse_tail<ORX_OPT_ADAM_LAZY>, dim3(grid), dim3(256), 0, st, ta); break;
    default: orx_launch_pdl(k_sparse_tail<ORX_OPT_ADAM_DENSE>, dim3(grid), dim3(256), 0, st, ta); break;
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
// D = 128 runs 4 CTAs/SM with a single register buffer (64 registers): the fastest of the variants A/B-tested in
// profiles r1b/r1c/r2a (2-3 CTAs/SM with a register double-buffer, a cp.async shared-memory ring, and a
// "last arriver applies" form without the tail launch were all slower and are gone).
template <int KIND, int OPT>
static int launch_pair_step_kind_opt(const PairArgs& pa, cudaStream_t st, int* n_partials) {
  const int B = pa.B;
  constexpr bool LAZY = (OPT == ORX_OPT_ADAM_LAZY);  // 9 rows per triplet: no register double-buffer
  auto go = [&](auto kern, int ch) {
    const int nw = (B + ch - 1) / ch;
    const int blocks = (nw + 7) / 8;
    *n_partials = blocks;
    orx_launch_pdl(kern, dim3(blocks), dim3(256), 0, st, pa);
  };
  switch (pa.D) {
    case 32: go(k_pair_step<KIND, OPT, 32, 8, 2, !LAZY>, 8); break;
    case 64: go(k_pair_step<KIND, OPT, 64, 8, 2, !LAZY>, 8); break;
    case 128:
      if (LAZY) go(k_pair_step<KIND, OPT, 128, 8, 3, false>, 8);
      else go(k_pair_step<KIND, OPT, 128, 8, 4, false>, 8);
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
    case ORX_OPT_SGD: return launch_pair_step_kind_opt<KIND, ORX_OPT_SGD>(pa, st, n_partials);
    case ORX_OPT_ADAGRAD: return launch_pair_step_kind_opt<KIND, ORX_OPT_ADAGRAD>(pa, st, n_partials);
    case ORX_OPT_ADAM_LAZY: return launch_pair_step_kind_opt<KIND, ORX_OPT_ADAM_LAZY>(pa, st, n_partials);
    case ORX_OPT_ADAM_DENSE: return launch_pair_step_kind_opt<KIND, ORX_OPT_ADAM_DENSE>(pa, st, n_partials);
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

// ---- pipelined batch index ------------------------------------------------------------------------------------------
// The index of a batch depends only on its ids, so it can be built while the PREVIOUS step's kernels still run: on the
// context's side stream, into one of two "prefetch" index sets (hash tables + counters) that alternate -- set 0 stays
// with everything that builds its index on the caller's stream (pointwise / DLRM / censor / sharded / un-prefetched
// pairwise steps).  The step that consumes a prefetched index waits for it with an event; the set is handed back with
// an event recorded behind that step's tail.
static int side_stream_ensure(orx_ctx* c) {
  if (c->side_stream) return ORX_OK;
  ORX_CUDA(cudaStreamCreateWithFlags(&c->side_stream, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    ORX_CUDA(cudaEventCreateWithFlags(&c->side_ev[i], cudaEventDisableTiming));
    ORX_CUDA(cudaEventCreateWithFlags(&c->pf_done[i], cudaEventDisableTiming));
    ORX_CUDA(cudaEventCreateWithFlags(&c->pf_free[i], cudaEventDisableTiming));
    ORX_CUDA(cudaEventCreateWithFlags(&c->stage_free[i], cudaEventDisableTiming));
    c->pf_free_valid[i] = c->stage_free_valid[i] = 0;
  }
  return ORX_OK;
}

// a prefetched index nobody consumed: wait for it, reset its counters (the hash itself dies with its epoch)
static int prefetch_drop(orx_ctx* c, cudaStream_t st) {
  if (!c->pf_valid) return ORX_OK;
  const int k = c->pf_set;
  ORX_CUDA(cudaStreamWaitEvent(st, c->pf_done[k], 0));
  ORX_CUDA(cudaMemsetAsync(c->counters + 4 * (1 + k), 0, sizeof(int32_t) * 4, st));
  ORX_CUDA(cudaEventRecord(c->pf_free[k], st));
  c->pf_free_valid[k] = 1;
  c->pf_valid = 0;
  return ORX_OK;
}

// Build the index of (uid, pid, nid) on the side stream.  `after` (may be null) is an event the build must wait for
// (the caller's "ids are final" point); the ids themselves may also be produced on the side stream (host upload).
static int prefetch_issue(orx_ctx* c, const int32_t* uid, const int32_t* pid, const int32_t* nid, int B, int64_t rows_u,
                          int64_t rows_i, int mode) {
  const int k = c->pf_next;
  c->pf_next ^= 1;
  cudaStream_t ss = c->side_stream;
  if (c->pf_free_valid[k]) ORX_CUDA(cudaStreamWaitEvent(ss, c->pf_free[k], 0));   // the step that last used set k is done
  int rc = orx_next_epoch(c, ss);
  if (rc) return rc;
  c->pf_u[k].epoch = c->pf_i[k].epoch = c->epoch;
  k_index_build<<<(3 * B + 255) / 256, 256, 0, ss>>>(c->pf_u[k], c->pf_i[k], uid, rows_u, B, pid, nid, rows_i, B, mode,
                                                     c->counters + 4 * (1 + k) + 3);
  ORX_LAUNCH_CHECK();
  ORX_CUDA(cudaEventRecord(c->pf_done[k], ss));
  c->pf_valid = 1; c->pf_set = k; c->pf_uid = uid; c->pf_pid = pid; c->pf_nid = nid; c->pf_B = B;
  c->pf_rows_u = rows_u; c->pf_rows_i = rows_i; c->pf_mode = mode;
  return ORX_OK;
}

extern "C" int orx_pairwise_prefetch(orx_handle_t h, const orx_table_t* user, const orx_table_t* item, const int32_t* uid,
                                     const int32_t* pid, const int32_t* nid, int32_t B, int32_t opt_kind, int32_t ids_ready,
                                     orx_stream_t ids_stream) {
  ORX_REQUIRE(h != nullptr && user && item && uid && pid && nid && B > 0, "bad arguments");
  ORX_REQUIRE(opt_kind >= ORX_OPT_SGD && opt_kind <= ORX_OPT_ADAM_DENSE, "unknown optimizer kind");
  ORX_CUDA(cudaSetDevice(h->device));
  cudaStream_t is = (cudaStream_t)ids_stream;
  int rc = orx_ensure_workspace(h, B, user->dim, false);
  if (rc) return rc;
  if ((rc = side_stream_ensure(h))) return rc;
  if ((rc = prefetch_drop(h, is))) return rc;
  if (!ids_ready) {   // the ids are final once everything queued on ids_stream so far has run
    ORX_CUDA(cudaEventRecord(h->side_ev[0], is));
    ORX_CUDA(cudaStreamWaitEvent(h->side_stream, h->side_ev[0], 0));
  }
  return prefetch_issue(h, uid, pid, nid, B, user->rows, item->rows, opt_kind == ORX_OPT_ADAM_DENSE ? 1 : 0);
}

static int pairwise_step_impl(orx_ctx* c, int kind, const orx_table_t* user, const orx_table_t* item,
                              const orx_table_t* bias, const int32_t* uid, const int32_t* pid, const int32_t* nid,
                              int B, float margin, float c_loss, float c_l2, const orx_opt_t* opt, float* out4,
                              cudaStream_t st) {
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
  // index: a matching prefetched one (side stream, possibly still running), else built here on the caller's stream
  int set = 0;
  if (c->pf_valid && c->pf_uid == uid && c->pf_pid == pid && c->pf_nid == nid && c->pf_B == B &&
      c->pf_rows_u == user->rows && c->pf_rows_i == item->rows && c->pf_mode == (dense ? 1 : 0)) {
    set = 1 + c->pf_set;
    ORX_CUDA(cudaStreamWaitEvent(st, c->pf_done[c->pf_set], 0));
    c->pf_valid = 0;
  } else {
    if ((rc = prefetch_drop(c, st))) return rc;
    if ((rc = orx_launch_index_build(c, uid, user->rows, B, pid, nid, item->rows, B, dense ? 1 : 0, st))) return rc;
  }
  const OrxHash& HU = set ? c->pf_u[set - 1] : c->hu;
  const OrxHash& HI = set ? c->pf_i[set - 1] : c->hi;
  int32_t* ctr = c->counters + 4 * set;
  orx_prof_mark(c, 1, st);
  pa.hu = HU; pa.hi = HI;
  int n_partials = 0;
  rc = (kind == ORX_PAIR_BPR) ? launch_pair_step_kind<ORX_PAIR_BPR>(pa, opt->kind, st, &n_partials)
                              : launch_pair_step_kind<ORX_PAIR_UCML>(pa, opt->kind, st, &n_partials);
  if (rc) return rc;
  orx_prof_mark(c, 2, st);
  if (dense) {
    if ((rc = orx_launch_adam_sweep(c, user->var, user->s0, user->s1, user->rows, D, HU, c->gu, pa.opt, st))) return rc;
    if ((rc = orx_launch_adam_sweep(c, item->var, item->s0, item->s1, item->rows, D, HI, c->gi, pa.opt, st))) return rc;
    if ((rc = orx_launch_adam_sweep(c, bias->var, bias->s0, bias->s1, bias->rows, 1, HI, c->gb, pa.opt, st))) return rc;
  }
  TailArgs ta;
  ta.U = user->var; ta.Us0 = user->s0; ta.Us1 = user->s1;
  ta.I = item->var; ta.Is0 = item->s0; ta.Is1 = item->s1;
  ta.Bv = bias->var; ta.Bs0 = bias->s0; ta.Bs1 = bias->s1;
  ta.D = D; ta.opt = pa.opt; ta.hu = HU; ta.hi = HI;
  ta.gu = c->gu; ta.gi = c->gi; ta.gb = c->gb;
  ta.partials = c->partials; ta.n_partials = n_partials;
  ta.loss_scale = (kind == ORX_PAIR_BPR) ? pa.inv_B : 1.0f;
  ta.counters = ctr; ta.out4 = out4;
  ta.W = ta.Ws0 = ta.Ws1 = ta.gw = nullptr; ta.c_l2 = c_l2;
  rc = orx_launch_tail(c, ta, opt->kind, st);
  if (set) {   // the prefetch set is free again once this tail has run
    ORX_CUDA(cudaEventRecord(c->pf_free[set - 1], st));
    c->pf_free_valid[set - 1] = 1;
  }
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

// Host-buffer form: the upload of this batch's ids and its index build run on the side stream, i.e. under the previous
// step's kernels whenever the caller enqueues ahead of the GPU; the step kernels and the read-back of out4 stay on `s`.
// The id staging buffers alternate; buffer f is reused only after the step that read it has finished (stage_free[f]).
extern "C" int orx_pairwise_step_host(orx_handle_t h, int32_t kind, const orx_table_t* user, const orx_table_t* item,
                                      const orx_table_t* item_bias, const int32_t* uid_host, const int32_t* pid_host,
                                      const int32_t* nid_host, int32_t B, float margin, float c_loss, float c_l2,
                                      const orx_opt_t* opt, float* out4_host, orx_stream_t s) {
  ORX_REQUIRE(h != nullptr, "null handle");
  ORX_REQUIRE(B > 0 && uid_host && pid_host && nid_host && out4_host && user && item && opt, "empty batch or null host buffers");
  ORX_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)s;
  int rc = orx_ensure_stage(h, 3 * (int64_t)B);
  if (rc) return rc;
  if ((rc = orx_ensure_workspace(h, B, user->dim, false))) return rc;
  if ((rc = side_stream_ensure(h))) return rc;
  if ((rc = prefetch_drop(h, st))) return rc;
  const uint32_t f = (h->stage_flip++) & 1u;
  int32_t* ids = h->ids_stage[f];
  cudaStream_t ss = h->side_stream;
  if (h->stage_free_valid[f]) ORX_CUDA(cudaStreamWaitEvent(ss, h->stage_free[f], 0));
  // three copies: the caller's arrays are separate (pinned) allocations even when they happen to be adjacent
  ORX_CUDA(cudaMemcpyAsync(ids, uid_host, sizeof(int32_t) * B, cudaMemcpyHostToDevice, ss));
  ORX_CUDA(cudaMemcpyAsync(ids + B, pid_host, sizeof(int32_t) * B, cudaMemcpyHostToDevice, ss));
  ORX_CUDA(cudaMemcpyAsync(ids + 2 * (int64_t)B, nid_host, sizeof(int32_t) * B, cudaMemcpyHostToDevice, ss));
  if ((rc = prefetch_issue(h, ids, ids + B, ids + 2 * (int64_t)B, B, user->rows, item->rows,
                           opt->kind == ORX_OPT_ADAM_DENSE ? 1 : 0)))
    return rc;
  rc = pairwise_step_impl(h, kind, user, item, item_bias, ids, ids + B, ids + 2 * (int64_t)B, B, margin, c_loss, c_l2,
                          opt, h->out_stage[f], st);
  if (rc) return rc;
  ORX_CUDA(cudaEventRecord(h->stage_free[f], st));
  h->stage_free_valid[f] = 1;
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
