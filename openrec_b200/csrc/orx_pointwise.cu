// orx_pointwise.cu -- GMF / WRMF fused training step (K3/K4).
//
// Reference path replaced: openrec/tf2/recommenders/gmf.py:22-34, wrmf.py:21-34,
// modules/pointwise_mse_loss.py:18-31 + tape.gradient + apply_gradients.
// Same three-launch structure as orx_pairwise.cu (index -> step -> tail); GMF's dense [D] weight
// gets its batch-summed gradient through shared-memory + one global reduction per block and is
// updated by the tail.
#include "orx_common.cuh"

struct PointArgs {
  float *U, *Us0, *Us1;
  float *I, *Is0, *Is1;
  float *Bv, *Bs0, *Bs1;
  const float* W;  // GMF weight [D] (pre-step) or null
  float* gw;       // staged dense gradient of W
  int64_t rowsU, rowsI;
  int D;
  const int32_t *uid, *iid;
  const float* label;
  int B;
  float wa, wb, c_loss, c_l2, inv_B;
  int use_sigmoid;
  OrxOptDev opt;
  OrxHash hu, hi;
  float *gu, *gi, *gb;
  float* partials;
  // un-fused outputs (grad kernel only)
  float *d_user, *d_item, *d_bias, *g_out;
};

// (loss term, dloss/dscore scalar).  GMF: BCE-with-logits mean (gmf.py:28-29) ; WRMF: weighted SSE
// (pointwise_mse_loss.py:22-31).
template <int KIND>
__device__ __forceinline__ void point_score(float s, float bias, float label, const PointArgs& a, float* lt,
                                            float* g) {
  if (KIND == ORX_POINT_GMF) {
    const float z = s + bias;
    *lt = fmaxf(z, 0.f) - z * label + log1pf(expf(-fabsf(z)));
    *g = a.c_loss * (orx_sigmoid(z) - label) * a.inv_B;
  } else {
    float pred = s + bias;
    if (a.use_sigmoid) pred = orx_sigmoid(pred);
    const float wgt = (a.wa - a.wb) * label + a.wb;
    const float diff = label - pred;
    *lt = wgt * diff * diff;
    float d = a.c_loss * -2.f * wgt * diff;
    if (a.use_sigmoid) d = d * pred * (1.f - pred);
    *g = d;
  }
}

template <int KIND, int OPT, int D, int CH>
__global__ void __launch_bounds__(256) k_point_step(const PointArgs a) {
  constexpr int G = (D / 4 < 32) ? D / 4 : 32;
  constexpr int K = D / (4 * G);
  constexpr int TPW = 32 / G;
  constexpr bool S0 = (OPT == ORX_OPT_ADAGRAD || OPT == ORX_OPT_ADAM_LAZY);
  constexpr bool S1 = (OPT == ORX_OPT_ADAM_LAZY);
  constexpr bool STAGE_ONLY = (OPT == ORX_OPT_ADAM_DENSE);
  constexpr bool GMF = (KIND == ORX_POINT_GMF);
  __shared__ float sgw[GMF ? D : 1];

  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int grp = lane / G, gl = lane % G;
  const int t = warp * CH + lane;
  if (GMF) {
    for (int e = threadIdx.x; e < D; e += blockDim.x) sgw[e] = 0.f;
    __syncthreads();
  }

  int u_id = 0, i_id = 0, du = -1, di = -1, flags = 0;
  float bi = 0.f, bs0 = 0.f, bs1 = 0.f, lab = 0.f;
  if (lane < CH && t < a.B) {
    u_id = a.uid[t];
    i_id = a.iid[t];
    lab = a.label[t];
    if (u_id >= 0 && u_id < a.rowsU && i_id >= 0 && i_id < a.rowsI) {
      const uint32_t cu = orx_hash_find(a.hu, u_id, &du);
      const uint32_t ci = orx_hash_find(a.hi, i_id, &di);
      bi = __ldcg(a.Bv + i_id);
      flags = 1;
      if (!STAGE_ONLY) {
        flags |= (cu == 1u ? 2 : 0) | (ci == 1u ? 4 : 0);
        if (S0 && (flags & 4)) bs0 = __ldcg(a.Bs0 + i_id);
        if (S1 && (flags & 4)) bs1 = __ldcg(a.Bs1 + i_id);
      }
    }
  }
  float4 w[K], gwacc[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    w[k] = GMF ? *reinterpret_cast<const float4*>(a.W + (k * G + gl) * 4) : make_float4(1.f, 1.f, 1.f, 1.f);
    gwacc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  }

  float loss_acc = 0.f, l2_acc = 0.f, g_own = 0.f;
#pragma unroll 1
  for (int j = 0; j < CH; j += TPW) {
    const int src = j + grp;
    const int fl = __shfl_sync(ORX_FULL, flags, src);
    const int uu = __shfl_sync(ORX_FULL, u_id, src), ii = __shfl_sync(ORX_FULL, i_id, src);
    const int duj = __shfl_sync(ORX_FULL, du, src), dij = __shfl_sync(ORX_FULL, di, src);
    const float bj = __shfl_sync(ORX_FULL, bi, src), lj = __shfl_sync(ORX_FULL, lab, src);
    const bool v = fl & 1;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 u[K], it[K], us0[K], is0[K], us1[K], is1[K];
    float s = 0.f, sq = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int off = (k * G + gl) * 4;
      u[k] = v ? __ldcg(reinterpret_cast<const float4*>(a.U + (int64_t)uu * D + off)) : z4;
      it[k] = v ? __ldcg(reinterpret_cast<const float4*>(a.I + (int64_t)ii * D + off)) : z4;
      if (S0) {
        us0[k] = (fl & 2) ? __ldcg(reinterpret_cast<const float4*>(a.Us0 + (int64_t)uu * D + off)) : z4;
        is0[k] = (fl & 4) ? __ldcg(reinterpret_cast<const float4*>(a.Is0 + (int64_t)ii * D + off)) : z4;
      }
      if (S1) {
        us1[k] = (fl & 2) ? __ldcg(reinterpret_cast<const float4*>(a.Us1 + (int64_t)uu * D + off)) : z4;
        is1[k] = (fl & 4) ? __ldcg(reinterpret_cast<const float4*>(a.Is1 + (int64_t)ii * D + off)) : z4;
      }
      s += w[k].x * u[k].x * it[k].x + w[k].y * u[k].y * it[k].y + w[k].z * u[k].z * it[k].z +
           w[k].w * u[k].w * it[k].w;
      sq += u[k].x * u[k].x + u[k].y * u[k].y + u[k].z * u[k].z + u[k].w * u[k].w + it[k].x * it[k].x +
            it[k].y * it[k].y + it[k].z * it[k].z + it[k].w * it[k].w;
    }
    l2_acc += sq;
    s = orx_group_sum<G>(s);
    float lt = 0.f, g = 0.f;
    point_score<KIND>(s, bj, lj, a, &lt, &g);
    if (!v) { lt = 0.f; g = 0.f; }
    if (gl == 0) loss_acc += lt;
#pragma unroll
    for (int q = 0; q < TPW; ++q) {
      const float val = __shfl_sync(ORX_FULL, g, q * G);
      if (lane == j + q) g_own = val;
    }
    if (v) {
      const float c2 = a.c_l2;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const int off = (k * G + gl) * 4;
        float4 gu, gi;
        gu.x = g * w[k].x * it[k].x + c2 * u[k].x; gu.y = g * w[k].y * it[k].y + c2 * u[k].y;
        gu.z = g * w[k].z * it[k].z + c2 * u[k].z; gu.w = g * w[k].w * it[k].w + c2 * u[k].w;
        gi.x = g * w[k].x * u[k].x + c2 * it[k].x; gi.y = g * w[k].y * u[k].y + c2 * it[k].y;
        gi.z = g * w[k].z * u[k].z + c2 * it[k].z; gi.w = g * w[k].w * u[k].w + c2 * it[k].w;
        if (GMF) {
          gwacc[k].x += g * u[k].x * it[k].x; gwacc[k].y += g * u[k].y * it[k].y;
          gwacc[k].z += g * u[k].z * it[k].z; gwacc[k].w += g * u[k].w * it[k].w;
        }
        if (!STAGE_ONLY && (fl & 2)) {
          const int64_t o = (int64_t)uu * D + off;
          __stcg(reinterpret_cast<float4*>(a.U + o), orx_apply4<OPT>(u[k], gu, us0[k], us1[k], a.opt));
          if (S0) __stcg(reinterpret_cast<float4*>(a.Us0 + o), us0[k]);
          if (S1) __stcg(reinterpret_cast<float4*>(a.Us1 + o), us1[k]);
        } else {
          orx_red4(a.gu + (int64_t)duj * D + off, gu);
        }
        if (!STAGE_ONLY && (fl & 4)) {
          const int64_t o = (int64_t)ii * D + off;
          __stcg(reinterpret_cast<float4*>(a.I + o), orx_apply4<OPT>(it[k], gi, is0[k], is1[k], a.opt));
          if (S0) __stcg(reinterpret_cast<float4*>(a.Is0 + o), is0[k]);
          if (S1) __stcg(reinterpret_cast<float4*>(a.Is1 + o), is1[k]);
        } else {
          orx_red4(a.gi + (int64_t)dij * D + off, gi);
        }
      }
    }
  }
  if (flags & 1) {
    if (flags & 4) {
      __stcg(a.Bv + i_id, orx_apply<OPT>(bi, g_own, bs0, bs1, a.opt));
      if (S0) __stcg(a.Bs0 + i_id, bs0);
      if (S1) __stcg(a.Bs1 + i_id, bs1);
    } else {
      atomicAdd(a.gb + di, g_own);
    }
  }
  if (GMF) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int off = (k * G + gl) * 4;
      atomicAdd(&sgw[off + 0], gwacc[k].x);
      atomicAdd(&sgw[off + 1], gwacc[k].y);
      atomicAdd(&sgw[off + 2], gwacc[k].z);
      atomicAdd(&sgw[off + 3], gwacc[k].w);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < D; e += blockDim.x) atomicAdd(a.gw + e, sgw[e]);
  }
  loss_acc = orx_group_sum<32>(loss_acc);
  l2_acc = orx_group_sum<32>(l2_acc);
  if (lane == 0) {
    a.partials[2 * warp] = loss_acc;
    a.partials[2 * warp + 1] = l2_acc;
  }
}

// Any dim; MODE 0 = fused step, 1 = forward / explicit (un-fused) gradients.
template <int KIND, int OPT, int MODE>
__global__ void __launch_bounds__(256) k_point_generic(const PointArgs a) {
  constexpr bool S0 = (OPT == ORX_OPT_ADAGRAD || OPT == ORX_OPT_ADAM_LAZY);
  constexpr bool S1 = (OPT == ORX_OPT_ADAM_LAZY);
  constexpr bool STAGE_ONLY = (OPT == ORX_OPT_ADAM_DENSE);
  constexpr bool GMF = (KIND == ORX_POINT_GMF);
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int D = a.D;
  float loss_acc = 0.f, l2_acc = 0.f;
  for (int j = 0; j < 8; ++j) {
    const int t = warp * 8 + j;
    if (t >= a.B) break;
    const int uu = a.uid[t], ii = a.iid[t];
    const bool ok = uu >= 0 && uu < a.rowsU && ii >= 0 && ii < a.rowsI;
    float* ur = a.U + (int64_t)uu * D;
    float* ir = a.I + (int64_t)ii * D;
    float s = 0.f, sq = 0.f;
    if (ok) {
      for (int d = lane; d < D; d += 32) {
        const float u = ur[d], it = ir[d], w = GMF ? a.W[d] : 1.f;
        s += w * u * it;
        sq += u * u + it * it;
      }
    }
    l2_acc += sq;
    s = orx_group_sum<32>(s);
    float lt = 0.f, g = 0.f;
    const float bi = ok ? a.Bv[ii] : 0.f;
    if (ok) point_score<KIND>(s, bi, a.label[t], a, &lt, &g);
    if (lane == 0) loss_acc += lt;
    int du = -1, di = -1;
    bool fu = false, fi = false;
    if (MODE == 0 && ok) {
      const uint32_t cu = orx_hash_find(a.hu, uu, &du);
      const uint32_t ci = orx_hash_find(a.hi, ii, &di);
      fu = !STAGE_ONLY && cu == 1u;
      fi = !STAGE_ONLY && ci == 1u;
    }
    const float c2 = a.c_l2;
    if (MODE == 0 ? ok : (a.d_user || a.d_item || a.gw)) {
      for (int d = lane; d < D; d += 32) {
        float gu = 0.f, gi = 0.f;
        if (ok) {
          const float u = ur[d], it = ir[d], w = GMF ? a.W[d] : 1.f;
          gu = g * w * it + c2 * u;
          gi = g * w * u + c2 * it;
          if (GMF && a.gw) atomicAdd(a.gw + d, g * u * it);
          if (MODE == 0) {
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
            if (fi) {
              const int64_t o = (int64_t)ii * D + d;
              if (S0) s0v = a.Is0[o];
              if (S1) s1v = a.Is1[o];
              ir[d] = orx_apply<OPT>(it, gi, s0v, s1v, a.opt);
              if (S0) a.Is0[o] = s0v;
              if (S1) a.Is1[o] = s1v;
            } else {
              atomicAdd(a.gi + (int64_t)di * D + d, gi);
            }
          }
        }
        if (MODE == 1) {
          const int64_t o = (int64_t)t * D + d;
          if (a.d_user) a.d_user[o] = gu;
          if (a.d_item) a.d_item[o] = gi;
        }
      }
    }
    if (lane == 0) {
      if (MODE == 0 && ok) {
        float s0v = 0.f, s1v = 0.f;
        if (fi) {
          if (S0) s0v = a.Bs0[ii];
          if (S1) s1v = a.Bs1[ii];
          a.Bv[ii] = orx_apply<OPT>(bi, g, s0v, s1v, a.opt);
          if (S0) a.Bs0[ii] = s0v;
          if (S1) a.Bs1[ii] = s1v;
        } else {
          atomicAdd(a.gb + di, g);
        }
      }
      if (MODE == 1) {
        if (a.d_bias) a.d_bias[t] = g;
        if (a.g_out) a.g_out[t] = g;
      }
    }
  }
  loss_acc = orx_group_sum<32>(loss_acc);
  l2_acc = orx_group_sum<32>(l2_acc);
  if (lane == 0) {
    a.partials[2 * warp] = loss_acc;
    a.partials[2 * warp + 1] = l2_acc;
  }
}

// adds 0.5*sum(w^2) to out4[1] (gmf.py:31-32) and, for the explicit-gradient path, c_l2*w to d_w
__global__ void k_gmf_w_terms(const float* W, int D, float* out4, float* d_w, float c_l2) {
  __shared__ float sh[256];
  float q = 0.f;
  for (int e = threadIdx.x; e < D; e += blockDim.x) {
    q += W[e] * W[e];
    if (d_w) d_w[e] += c_l2 * W[e];
  }
  sh[threadIdx.x] = q;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0 && out4) out4[1] += 0.5f * sh[0];
}

template <int KIND, int OPT>
static int launch_point_kind_opt(const PointArgs& pa, cudaStream_t st, int* n_partials) {
  const int nw = (pa.B + 7) / 8, blocks = (nw + 7) / 8;
  *n_partials = blocks * 8;
  switch (pa.D) {
    case 32: k_point_step<KIND, OPT, 32, 8><<<blocks, 256, 0, st>>>(pa); break;
    case 64: k_point_step<KIND, OPT, 64, 8><<<blocks, 256, 0, st>>>(pa); break;
    case 128: k_point_step<KIND, OPT, 128, 8><<<blocks, 256, 0, st>>>(pa); break;
    case 256: k_point_step<KIND, OPT, 256, 8><<<blocks, 256, 0, st>>>(pa); break;
    default: k_point_generic<KIND, OPT, 0><<<blocks, 256, 0, st>>>(pa); break;
  }
  ORX_LAUNCH_CHECK();
  return ORX_OK;
}

template <int KIND>
static int launch_point_kind(const PointArgs& pa, int opt_kind, cudaStream_t st, int* n_partials) {
  switch (opt_kind) {
    case ORX_OPT_SGD: return launch_point_kind_opt<KIND, ORX_OPT_SGD>(pa, st, n_partials);
    case ORX_OPT_ADAGRAD: return launch_point_kind_opt<KIND, ORX_OPT_ADAGRAD>(pa, st, n_partials);
    case ORX_OPT_ADAM_LAZY: return launch_point_kind_opt<KIND, ORX_OPT_ADAM_LAZY>(pa, st, n_partials);
    case ORX_OPT_ADAM_DENSE: return launch_point_kind_opt<KIND, ORX_OPT_ADAM_DENSE>(pa, st, n_partials);
  }
  orx_set_error("unknown optimizer kind %d", opt_kind);
  return ORX_ERR_INVALID;
}

static int check_point(int kind, const orx_table_t* user, const orx_table_t* item, const orx_table_t* bias,
                       const orx_table_t* w, int opt_kind) {
  ORX_REQUIRE(kind == ORX_POINT_GMF || kind == ORX_POINT_WRMF, "unknown pointwise kind");
  ORX_REQUIRE(user && item && bias && user->var && item->var && bias->var, "null table");
  ORX_REQUIRE(user->dim == item->dim && user->dim > 0, "user/item dims must match and be positive");
  ORX_REQUIRE(bias->dim == 1 && bias->rows == item->rows, "item_bias must be [item.rows, 1]");
  ORX_REQUIRE(user->rows > 0 && item->rows > 0 && user->rows <= 0x7fffffffLL && item->rows <= 0x7fffffffLL,
              "row counts must fit int32 ids");
  if (kind == ORX_POINT_GMF) ORX_REQUIRE(w && w->var && w->dim == user->dim, "GMF needs w with dim == D");
  const bool has0 = opt_kind != ORX_OPT_SGD, has1 = opt_kind == ORX_OPT_ADAM_LAZY || opt_kind == ORX_OPT_ADAM_DENSE;
  if (has0) ORX_REQUIRE(user->s0 && item->s0 && bias->s0 && (kind != ORX_POINT_GMF || w->s0), "slot s0 missing");
  if (has1) ORX_REQUIRE(user->s1 && item->s1 && bias->s1 && (kind != ORX_POINT_GMF || w->s1), "slot s1 missing");
  return ORX_OK;
}

static void fill_point_args(PointArgs& pa, orx_ctx* c, int kind, const orx_table_t* user, const orx_table_t* item,
                            const orx_table_t* bias, const orx_table_t* w, const int32_t* uid, const int32_t* iid,
                            const float* label, int B, float a, float b, int use_sigmoid, float c_loss, float c_l2) {
  pa.U = user->var; pa.Us0 = user->s0; pa.Us1 = user->s1;
  pa.I = item->var; pa.Is0 = item->s0; pa.Is1 = item->s1;
  pa.Bv = bias->var; pa.Bs0 = bias->s0; pa.Bs1 = bias->s1;
  pa.W = (kind == ORX_POINT_GMF) ? w->var : nullptr;
  pa.gw = nullptr;
  pa.rowsU = user->rows; pa.rowsI = item->rows; pa.D = user->dim;
  pa.uid = uid; pa.iid = iid; pa.label = label; pa.B = B;
  pa.wa = a; pa.wb = b; pa.c_loss = c_loss; pa.c_l2 = c_l2; pa.inv_B = 1.0f / (float)B;
  pa.use_sigmoid = use_sigmoid;
  pa.hu = c->hu; pa.hi = c->hi; pa.gu = c->gu; pa.gi = c->gi; pa.gb = c->gb;
  pa.partials = c->partials;
  pa.d_user = pa.d_item = pa.d_bias = pa.g_out = nullptr;
  pa.opt.kind = 0; pa.opt.lr = 0.f; pa.opt.eps = 0.f; pa.opt.beta1 = 0.f; pa.opt.beta2 = 0.f;
}

extern "C" int orx_pointwise_step(orx_handle_t h, int32_t kind, const orx_table_t* user, const orx_table_t* item,
                                  const orx_table_t* item_bias, const orx_table_t* w, const int32_t* uid,
                                  const int32_t* iid, const float* label, int32_t B, float a, float b,
                                  int32_t use_sigmoid, float c_loss, float c_l2, const orx_opt_t* opt, float* out4,
                                  orx_stream_t s) {
  ORX_REQUIRE(h != nullptr && opt != nullptr && out4 != nullptr, "null handle/opt/out");
  ORX_REQUIRE(opt->kind >= ORX_OPT_SGD && opt->kind <= ORX_OPT_ADAM_DENSE, "unknown optimizer kind");
  ORX_REQUIRE(B > 0 && uid && iid && label, "empty batch or null inputs");
  ORX_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)s;
  int rc = check_point(kind, user, item, item_bias, w, opt->kind);
  if (rc) return rc;
  const int D = user->dim;
  const bool dense = opt->kind == ORX_OPT_ADAM_DENSE;
  if ((rc = orx_ensure_workspace(h, B, D, dense))) return rc;
  if ((rc = orx_ensure_partials(h, (B + 7) / 8 + 8, st))) return rc;
  if ((rc = orx_launch_index_build(h, uid, user->rows, B, iid, nullptr, item->rows, B, dense, st))) return rc;
  PointArgs pa;
  fill_point_args(pa, h, kind, user, item, item_bias, w, uid, iid, label, B, a, b, use_sigmoid, c_loss, c_l2);
  pa.opt = orx_opt_to_dev(opt);
  pa.gw = (kind == ORX_POINT_GMF) ? h->gw : nullptr;
  int n_partials = 0;
  rc = (kind == ORX_POINT_GMF) ? launch_point_kind<ORX_POINT_GMF>(pa, opt->kind, st, &n_partials)
                               : launch_point_kind<ORX_POINT_WRMF>(pa, opt->kind, st, &n_partials);
  if (rc) return rc;
  if (dense) {
    if ((rc = orx_launch_adam_sweep(h, user->var, user->s0, user->s1, user->rows, D, h->hu, h->gu, pa.opt, st))) return rc;
    if ((rc = orx_launch_adam_sweep(h, item->var, item->s0, item->s1, item->rows, D, h->hi, h->gi, pa.opt, st))) return rc;
    if ((rc = orx_launch_adam_sweep(h, item_bias->var, item_bias->s0, item_bias->s1, item_bias->rows, 1, h->hi, h->gb, pa.opt, st))) return rc;
  }
  TailArgs ta;
  ta.U = user->var; ta.Us0 = user->s0; ta.Us1 = user->s1;
  ta.I = item->var; ta.Is0 = item->s0; ta.Is1 = item->s1;
  ta.Bv = item_bias->var; ta.Bs0 = item_bias->s0; ta.Bs1 = item_bias->s1;
  ta.D = D; ta.opt = pa.opt; ta.hu = h->hu; ta.hi = h->hi;
  ta.gu = h->gu; ta.gi = h->gi; ta.gb = h->gb;
  ta.partials = h->partials; ta.n_partials = n_partials;
  ta.loss_scale = (kind == ORX_POINT_GMF) ? pa.inv_B : 1.0f;
  ta.counters = h->counters; ta.out4 = out4;
  ta.W = ta.Ws0 = ta.Ws1 = ta.gw = nullptr; ta.c_l2 = c_l2;
  if (kind == ORX_POINT_GMF) {
    ta.W = w->var; ta.Ws0 = w->s0; ta.Ws1 = w->s1; ta.gw = h->gw;
  }
  return orx_launch_tail(h, ta, opt->kind, st);
}

static int point_fwd_grad(orx_ctx* h, int kind, const orx_table_t* user, const orx_table_t* item,
                          const orx_table_t* bias, const orx_table_t* w, const int32_t* uid, const int32_t* iid,
                          const float* label, int B, float a, float b, int use_sigmoid, float c_loss, float c_l2,
                          float* d_user, float* d_item, float* d_bias, float* d_w, float* g_out, float* out4,
                          cudaStream_t st) {
  ORX_REQUIRE(B > 0 && uid && iid && label, "empty batch or null inputs");
  int rc = check_point(kind, user, item, bias, w, ORX_OPT_SGD);
  if (rc) return rc;
  const int nw = (B + 7) / 8, blocks = (nw + 7) / 8;
  if ((rc = orx_ensure_partials(h, blocks * 8, st))) return rc;
  PointArgs pa;
  fill_point_args(pa, h, kind, user, item, bias, w, uid, iid, label, B, a, b, use_sigmoid, c_loss, c_l2);
  pa.d_user = d_user; pa.d_item = d_item; pa.d_bias = d_bias; pa.g_out = g_out;
  if (kind == ORX_POINT_GMF && d_w) {
    ORX_CUDA(cudaMemsetAsync(d_w, 0, sizeof(float) * user->dim, st));
    pa.gw = d_w;
  }
  if (kind == ORX_POINT_GMF) k_point_generic<ORX_POINT_GMF, ORX_OPT_SGD, 1><<<blocks, 256, 0, st>>>(pa);
  else k_point_generic<ORX_POINT_WRMF, ORX_OPT_SGD, 1><<<blocks, 256, 0, st>>>(pa);
  ORX_LAUNCH_CHECK();
  if (out4) {
    rc = orx_launch_reduce_partials(h->partials, blocks * 8, kind == ORX_POINT_GMF ? pa.inv_B : 1.f, out4, st);
    if (rc) return rc;
  }
  if (kind == ORX_POINT_GMF && (out4 || d_w)) {
    k_gmf_w_terms<<<1, 256, 0, st>>>(w->var, user->dim, out4, d_w, c_l2);
    ORX_LAUNCH_CHECK();
  }
  return ORX_OK;
}

extern "C" int orx_pointwise_fwd(orx_handle_t h, int32_t kind, const orx_table_t* user, const orx_table_t* item,
                                 const orx_table_t* item_bias, const orx_table_t* w, const int32_t* uid,
                                 const int32_t* iid, const float* label, int32_t B, float a, float b,
                                 int32_t use_sigmoid, float* out4, orx_stream_t s) {
  ORX_REQUIRE(h != nullptr && out4 != nullptr, "null handle/out");
  ORX_CUDA(cudaSetDevice(h->device));
  return point_fwd_grad(h, kind, user, item, item_bias, w, uid, iid, label, B, a, b, use_sigmoid, 1.f, 1.f, nullptr,
                        nullptr, nullptr, nullptr, nullptr, out4, (cudaStream_t)s);
}

extern "C" int orx_pointwise_grad(orx_handle_t h, int32_t kind, const orx_table_t* user, const orx_table_t* item,
                                  const orx_table_t* item_bias, const orx_table_t* w, const int32_t* uid,
                                  const int32_t* iid, const float* label, int32_t B, float a, float b,
                                  int32_t use_sigmoid, float c_loss, float c_l2, float* d_user, float* d_item,
                                  float* d_bias, float* d_w, float* g_out, orx_stream_t s) {
  ORX_REQUIRE(h != nullptr, "null handle");
  ORX_CUDA(cudaSetDevice(h->device));
  return point_fwd_grad(h, kind, user, item, item_bias, w, uid, iid, label, B, a, b, use_sigmoid, c_loss, c_l2, d_user,
                        d_item, d_bias, d_w, g_out, nullptr, (cudaStream_t)s);
}
