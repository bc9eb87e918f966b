// orx_common.cuh -- shared host/device helpers for liborx (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/orx.h"

// ---------------------------------------------------------------------------------------
// host: errors + context
// ---------------------------------------------------------------------------------------
void orx_set_error(const char* fmt, ...);

#define ORX_CUDA(call)                                                                         \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      orx_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));     \
      return ORX_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

#define ORX_REQUIRE(cond, msg)                                  \
  do {                                                          \
    if (!(cond)) {                                              \
      orx_set_error("%s: %s", __func__, msg);                   \
      return ORX_ERR_INVALID;                                   \
    }                                                           \
  } while (0)

#define ORX_LAUNCH_CHECK() ORX_CUDA(cudaGetLastError())

// Open-addressing hash index over the ids of one table for one batch ("K9").
// slot word: [63:33] epoch | [32] "seen more than once" | [31:0] id.  A slot whose epoch differs from the
// current one is EMPTY, so the table is never cleared: every step just uses a new epoch (31 bits).
struct OrxHash {
  unsigned long long* slots;  // [cap]
  int32_t* didx;              // [cap]  compact staging index of a staged row
  int32_t* did;               // [cap_rows] row id of staging index d
  int32_t* counter;           // number of staged rows
  uint32_t mask;
  int32_t shift;  // 32 - log2(cap)
  uint32_t epoch;  // current epoch, >= 1
};

struct orx_ctx {
  int device;
  int num_sms;
  // index workspace (sized for cap_B lookups per table side)
  int64_t cap_B;  // largest batch the workspace is sized for
  OrxHash hu, hi;      // index set 0: everything that builds its index on the caller's stream
  OrxHash pf_u[2], pf_i[2];  // index sets 1, 2: pairwise batches indexed ahead on the side stream (orx_pairwise.cu)
  int32_t* counters;  // [16]: per index set k at 4k: staged_u, staged_i, ticket, bad ids; 12.. spare
  // staged-row gradient buffers
  float *gu, *gi, *gb, *gw;
  int64_t g_rows_u, g_rows_i;
  int32_t g_dim;
  // loss partials
  float* partials;  // [cap_partials*2]
  int32_t cap_partials;
  // id staging for the *_host entry points (double buffered)
  int32_t* ids_stage[2];
  float* out_stage[2];
  int64_t stage_cap;
  uint32_t stage_flip;
  // measurement hook (orx_profile_*)
  int prof_on, prof_n, prof_cap;
  int prof_step;  // steps seen since orx_profile_enable: every 8th one carries the phase events
  cudaEvent_t* prof_ev;  // [prof_cap * ORX_PROF_EV]
  int32_t* bucket_cursor;  // owner-bucket scratch
  cudaStream_t side_stream;  // id upload + index build of the NEXT pairwise batch, beside the running step
  cudaEvent_t side_ev[2];
  cudaEvent_t pf_done[2], pf_free[2], stage_free[2];   // prefetched index k built / handed back; id staging f free
  int pf_free_valid[2], stage_free_valid[2];
  int pf_valid, pf_set, pf_next, pf_B, pf_mode;        // the one outstanding prefetched index and what it was built for
  const int32_t *pf_uid, *pf_pid, *pf_nid;
  int64_t pf_rows_u, pf_rows_i;
  uint32_t epoch;          // hash epoch of the last step, in [1, 2^31)
  void* shard_ws;          // orx_shard.cu: local scratch of the row-sharded step (orx_shard_ws*)
};

// Start a new hash epoch (once per step, before the index build; `st` = the stream the step runs on).
// A slot word holds 31 epoch bits: epochs live in [1, 2^31) and on wrap every table is zeroed on `st`, so a stale
// slot can never alias the current epoch (about 65 h of back-to-back steps between wraps).
int orx_next_epoch(orx_ctx* c, cudaStream_t st);
void orx_shard_ws_release(orx_ctx* c);

#define ORX_PROF_EV 8   // event slots per sampled step: up to 7 phases (the sharded step has seven launches)
// record phase boundary k (0..7) of the current step on `st` when profiling is enabled
// Only every 8th step is instrumented: four timing-event records per step sit between the kernels of the step that is
// being timed.  Suspected cost (not yet isolated): bench.py's un-instrumented UCML loop ran at 612 M/s against 555 M/s
// for the instrumented BPR loop in r1w although both step kernels take 71 us under ncu.
static inline bool orx_prof_sampled(const orx_ctx* c) {
  return c->prof_on && (c->prof_step & 7) == 0 && c->prof_n < c->prof_cap;
}
static inline void orx_prof_mark(orx_ctx* c, int k, cudaStream_t st) {
  if (orx_prof_sampled(c)) cudaEventRecord(c->prof_ev[c->prof_n * ORX_PROF_EV + k], st);
}
static inline void orx_prof_next(orx_ctx* c) {
  if (!c->prof_on) return;
  if (orx_prof_sampled(c)) c->prof_n++;
  c->prof_step++;
}

int orx_ensure_workspace(orx_ctx* c, int64_t B, int32_t dim, bool full_staging);
int orx_ensure_stage(orx_ctx* c, int64_t n_ints);

// ---------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------
struct OrxOptDev {
  int32_t kind;
  float lr;    // SGD/Adagrad: lr ; Adam: bias-corrected lr_t
  float eps, beta1, beta2;
};

#ifdef __CUDACC__

#define ORX_FULL 0xffffffffu

// Programmatic dependent launch (sm_90+): a kernel launched with orx_launch_pdl may become resident while its
// predecessor on the stream is still draining; it must execute orx_pdl_wait() before it touches anything the predecessor
// wrote (a no-op when launched normally).  The predecessor calls orx_pdl_trigger() once a block has finished its main
// loop: the dependent grid is released when every block has triggered or exited, so launch latency and grid drain overlap.
__device__ __forceinline__ void orx_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void orx_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool orx_pdl_enabled();   // ORX_PDL=0 turns the attribute off (A/B measurements)

template <typename... KArgs, typename... Args>
static inline cudaError_t orx_launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                         Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = orx_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

__device__ __forceinline__ uint32_t orx_hash32(uint32_t id, int shift) { return (id * 2654435769u) >> shift; }

#define ORX_DUP_BIT (1ull << 32)
__device__ __forceinline__ unsigned long long orx_slot_word(uint32_t epoch, int32_t id) {
  return ((unsigned long long)epoch << 33) | (unsigned long long)(uint32_t)id;
}

// Insert one id.  mode 0: rows get a staging index when they are seen the SECOND time (duplicates only);
// mode 1: on the FIRST occurrence (ADAM_DENSE stages every row); mode 2: never (pure dedup, censor).
// Returns 0 if this call was the first occurrence of the id in this epoch, else 1.
__device__ __forceinline__ uint32_t orx_hash_insert(const OrxHash& t, int32_t id, int mode) {
  const unsigned long long mine = orx_slot_word(t.epoch, id);
  uint32_t h = orx_hash32((uint32_t)id, t.shift);
  while (true) {
    unsigned long long w = __ldcg(t.slots + h);
    if ((uint32_t)(w >> 33) != t.epoch) {   // empty or stale: try to claim it
      const unsigned long long old = atomicCAS(t.slots + h, w, mine);
      if (old == w) {
        if (mode == 1) {
          const int d = atomicAdd(t.counter, 1);
          t.didx[h] = d;
          t.did[d] = id;
        }
        return 0u;
      }
      w = old;                              // somebody else claimed it meanwhile (same epoch by construction)
    }
    if ((w & ~ORX_DUP_BIT) == mine) {
      if (!(w & ORX_DUP_BIT)) {
        const unsigned long long old = atomicOr(t.slots + h, ORX_DUP_BIT);
        if (mode == 0 && !(old & ORX_DUP_BIT)) {   // this call made the row "shared": give it a staging slot
          const int d = atomicAdd(t.counter, 1);
          t.didx[h] = d;
          t.did[d] = id;
        }
      }
      return 1u;
    }
    h = (h + 1) & t.mask;
  }
}

// Lookup.  Returns 0 = absent, 1 = present once, 2 = present more than once; *d = staging index if the row
// has one (duplicates in mode 0, every row in mode 1).
__device__ __forceinline__ uint32_t orx_hash_find(const OrxHash& t, int32_t id, int32_t* d) {
  const unsigned long long mine = orx_slot_word(t.epoch, id);
  uint32_t h = orx_hash32((uint32_t)id, t.shift);
  while (true) {
    const unsigned long long w = __ldg(t.slots + h);
    if ((w & ~ORX_DUP_BIT) == mine) {
      *d = __ldg(t.didx + h);
      return (w & ORX_DUP_BIT) ? 2u : 1u;
    }
    if ((uint32_t)(w >> 33) != t.epoch) {
      *d = -1;
      return 0u;
    }
    h = (h + 1) & t.mask;
  }
}

template <int W>
__device__ __forceinline__ float orx_group_sum(float v) {
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(ORX_FULL, v, o);
  return v;
}

__device__ __forceinline__ float4 orx_ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void orx_st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
// one 128-bit fire-and-forget reduction (REDG.E.ADD.F32x4 on sm_90+)
__device__ __forceinline__ void orx_red4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// log(sigmoid(y)) and sigmoid(-y), the stable forms TF uses.
__device__ __forceinline__ void orx_logsig(float y, float* logsig, float* sig_neg) {
  float e = expf(-fabsf(y));
  *logsig = fminf(y, 0.f) - log1pf(e);
  *sig_neg = (y >= 0.f) ? e / (1.f + e) : 1.f / (1.f + e);
}
__device__ __forceinline__ float orx_sigmoid(float y) {
  float e = expf(-fabsf(y));
  return (y >= 0.f) ? 1.f / (1.f + e) : e / (1.f + e);
}


// MUFU approximations (max rel. error ~2^-22): the IEEE sqrtf + division of Adagrad/Adam cost ~25
// instructions per element and made the fused step issue-bound (ncu r1a: 553 warp-inst per triplet).
// The induced error on an updated parameter is < 1e-9 absolute at the reference's value ranges.
__device__ __forceinline__ float orx_sqrt_fast(float x) {
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float orx_rcp_fast(float x) {
  float r;
  asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// One optimizer update of one scalar.  OPT is an orx_opt_kind (ADAM_DENSE never reaches here:
// its rows are staged and swept).
template <int OPT>
__device__ __forceinline__ float orx_apply(float w, float g, float& s0, float& s1, const OrxOptDev& o) {
  if (OPT == ORX_OPT_SGD) {
    return w - o.lr * g;
  } else if (OPT == ORX_OPT_ADAGRAD) {
    s0 = s0 + g * g;
    return w - o.lr * g * orx_rcp_fast(orx_sqrt_fast(s0) + o.eps);
  } else {
    s0 = o.beta1 * s0 + (1.f - o.beta1) * g;
    s1 = o.beta2 * s1 + (1.f - o.beta2) * g * g;
    return w - o.lr * s0 * orx_rcp_fast(orx_sqrt_fast(s1) + o.eps);
  }
}

template <int OPT>
__device__ __forceinline__ float4 orx_apply4(float4 w, float4 g, float4& s0, float4& s1, const OrxOptDev& o) {
  float4 r;
  r.x = orx_apply<OPT>(w.x, g.x, s0.x, s1.x, o);
  r.y = orx_apply<OPT>(w.y, g.y, s0.y, s1.y, o);
  r.z = orx_apply<OPT>(w.z, g.z, s0.z, s1.z, o);
  r.w = orx_apply<OPT>(w.w, g.w, s0.w, s1.w, o);
  return r;
}

#endif  // __CUDACC__

// Arguments of the shared tail kernel (staged rows -> optimizer, hash clear, loss reduction).
struct TailArgs {
  float *U, *Us0, *Us1;
  float *I, *Is0, *Is1;
  float *Bv, *Bs0, *Bs1;
  int D;
  OrxOptDev opt;
  OrxHash hu, hi;
  float *gu, *gi, *gb;
  const float* partials;
  int n_partials;
  float loss_scale;  // BPR: 1/B (mean), UCML: 1 (sum)
  int32_t* counters;
  float* out4;
  // GMF dense weight (pointwise tail only)
  float *W, *Ws0, *Ws1, *gw;
  float c_l2;
};

OrxOptDev orx_opt_to_dev(const orx_opt_t* o);
int orx_launch_index_build_strided(orx_ctx* c, const int32_t* a, int64_t stride, int64_t rows, int32_t n,
                                   const int32_t* n_dev, bool stage_all, cudaStream_t st);
int orx_launch_tail(orx_ctx* c, const TailArgs& ta, int opt_kind, cudaStream_t st);
int orx_launch_adam_sweep(orx_ctx* c, float* var, float* m, float* v, int64_t rows, int D, const OrxHash& h,
                          const float* gstage, const OrxOptDev& o, cudaStream_t st);
int orx_ensure_partials(orx_ctx* c, int need, cudaStream_t st);
int orx_launch_reduce_partials(const float* partials, int n, float loss_scale, float* out4, cudaStream_t st);
int orx_launch_index_build(orx_ctx* c, const int32_t* a, int64_t rows_a, int32_t na, const int32_t* b0,
                           const int32_t* b1, int64_t rows_b, int32_t nb, int mode /* orx_hash_insert mode: 0 | 1 */,
                           cudaStream_t st);
