// orx_ctx.cu -- context, workspace and error plumbing of liborx.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "orx_common.cuh"

static thread_local char g_err[512] = "";

void orx_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* orx_last_error_string(void) { return g_err; }

bool orx_pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("ORX_PDL");
    v = (e && atoi(e) == 0) ? 0 : 1;
  }
  return v != 0;
}
extern "C" int orx_abi_version(void) { return ORX_ABI_VERSION; }

extern "C" int orx_device_count(int* n) {
  ORX_REQUIRE(n != nullptr, "null output");
  *n = 0;
  ORX_CUDA(cudaGetDeviceCount(n));
  return ORX_OK;
}

static void free_hash(OrxHash& t) {
  cudaFree(t.slots);
  cudaFree(t.didx);
  cudaFree(t.did);
  t.slots = nullptr;
  t.didx = nullptr;
  t.did = nullptr;
}

static void free_workspace(orx_ctx* c) {
  free_hash(c->hu);
  free_hash(c->hi);
  for (int k = 0; k < 2; ++k) {
    free_hash(c->pf_u[k]);
    free_hash(c->pf_i[k]);
  }
  cudaFree(c->gu);
  cudaFree(c->gi);
  cudaFree(c->gb);
  cudaFree(c->gw);
  c->gu = c->gi = c->gb = c->gw = nullptr;
  c->cap_B = 0;
  c->g_dim = 0;
}

static uint32_t pow2_at_least(int64_t n) {
  uint32_t p = 1024;
  while ((int64_t)p < n) p <<= 1;
  return p;
}

static int alloc_hash(OrxHash& t, int64_t lookups, int32_t* counter) {
  // load factor <= 0.25: with linear probing the slowest of a warp's 32 inserts/probes sets the pace
  // (profile r1b: ~7 serialized L2 round trips per warp at 0.5)
  uint32_t cap = pow2_at_least(4 * lookups);
  int lg = 0;
  while ((1u << lg) < cap) ++lg;
  t.mask = cap - 1;
  t.shift = 32 - lg;
  t.counter = counter;
  ORX_CUDA(cudaMalloc(&t.slots, sizeof(unsigned long long) * cap));
  ORX_CUDA(cudaMalloc(&t.didx, sizeof(int32_t) * cap));
  ORX_CUDA(cudaMalloc(&t.did, sizeof(int32_t) * (lookups + 1)));
  ORX_CUDA(cudaMemset(t.slots, 0, sizeof(unsigned long long) * cap));
  return ORX_OK;
}

// Workspace is sized for B lookups on the user side and 2B on the item side; every lookup may be
// staged (ADAM_DENSE stages all rows), so the staging buffers hold B resp. 2B rows of `dim` floats.
int orx_ensure_workspace(orx_ctx* c, int64_t B, int32_t dim, bool /*full_staging*/) {
  if (B <= c->cap_B && dim <= c->g_dim) return ORX_OK;
  int64_t nb = B > c->cap_B ? B : c->cap_B;
  int32_t nd = dim > c->g_dim ? dim : c->g_dim;
  ORX_CUDA(cudaDeviceSynchronize());
  free_workspace(c);
  int rc;
  if ((rc = alloc_hash(c->hu, nb, c->counters + 0)) != ORX_OK) return rc;
  if ((rc = alloc_hash(c->hi, 2 * nb, c->counters + 1)) != ORX_OK) return rc;
  for (int k = 0; k < 2; ++k) {
    if ((rc = alloc_hash(c->pf_u[k], nb, c->counters + 4 * (1 + k))) != ORX_OK) return rc;
    if ((rc = alloc_hash(c->pf_i[k], 2 * nb, c->counters + 4 * (1 + k) + 1)) != ORX_OK) return rc;
  }
  c->pf_valid = 0;
  c->pf_free_valid[0] = c->pf_free_valid[1] = 0;
  c->g_rows_u = nb;
  c->g_rows_i = 2 * nb;
  size_t bu = sizeof(float) * (size_t)c->g_rows_u * nd, bi = sizeof(float) * (size_t)c->g_rows_i * nd;
  ORX_CUDA(cudaMalloc(&c->gu, bu));
  ORX_CUDA(cudaMalloc(&c->gi, bi));
  ORX_CUDA(cudaMalloc(&c->gb, sizeof(float) * (size_t)c->g_rows_i));
  ORX_CUDA(cudaMalloc(&c->gw, sizeof(float) * (size_t)nd));
  ORX_CUDA(cudaMemset(c->gu, 0, bu));
  ORX_CUDA(cudaMemset(c->gi, 0, bi));
  ORX_CUDA(cudaMemset(c->gb, 0, sizeof(float) * (size_t)c->g_rows_i));
  ORX_CUDA(cudaMemset(c->gw, 0, sizeof(float) * (size_t)nd));
  ORX_CUDA(cudaMemset(c->counters, 0, sizeof(int32_t) * 16));
  // the memsets above ran on the legacy default stream; the caller's stream may be a non-blocking one
  ORX_CUDA(cudaDeviceSynchronize());
  c->cap_B = nb;
  c->g_dim = nd;
  c->epoch = 0;  // fresh (zeroed) tables: epochs restart at 1
  return ORX_OK;
}

static int zero_hash(OrxHash& t) {
  if (!t.slots) return ORX_OK;
  ORX_CUDA(cudaMemset(t.slots, 0, sizeof(unsigned long long) * ((size_t)t.mask + 1)));
  return ORX_OK;
}

int orx_next_epoch(orx_ctx* c, cudaStream_t /*st*/) {
  c->epoch = (c->epoch + 1) & 0x7fffffffu;
  if (c->epoch == 0) {
    // 31-bit wrap (once per 2^31 index builds): a stale slot could alias the epochs to come, so every table is emptied.
    // Index builds may be in flight on two streams: drain the device around the memsets.
    ORX_CUDA(cudaDeviceSynchronize());
    int rc;
    if ((rc = zero_hash(c->hu)) || (rc = zero_hash(c->hi))) return rc;
    for (int k = 0; k < 2; ++k)
      if ((rc = zero_hash(c->pf_u[k])) || (rc = zero_hash(c->pf_i[k]))) return rc;
    ORX_CUDA(cudaDeviceSynchronize());
    c->epoch = 1;
  }
  c->hu.epoch = c->hi.epoch = c->epoch;
  return ORX_OK;
}

// test hook: place the epoch counter (tests/test_gpu_kernels.py::test_epoch_wrap starts it just below 2^31)
extern "C" int orx_debug_set_epoch(orx_handle_t h, uint32_t epoch) {
  ORX_REQUIRE(h != nullptr && epoch < 0x80000000u, "null handle / epoch must be < 2^31");
  h->epoch = epoch;
  return ORX_OK;
}

int orx_ensure_stage(orx_ctx* c, int64_t n_ints) {
  if (n_ints <= c->stage_cap) return ORX_OK;
  ORX_CUDA(cudaDeviceSynchronize());
  for (int i = 0; i < 2; ++i) {
    cudaFree(c->ids_stage[i]);
    c->ids_stage[i] = nullptr;
    ORX_CUDA(cudaMalloc(&c->ids_stage[i], sizeof(int32_t) * (size_t)n_ints));
  }
  c->stage_cap = n_ints;
  return ORX_OK;
}

extern "C" int orx_create(int device, orx_handle_t* out) {
  ORX_REQUIRE(out != nullptr, "null output handle");
  *out = nullptr;
  int n = 0;
  ORX_CUDA(cudaGetDeviceCount(&n));
  ORX_REQUIRE(device >= 0 && device < n, "device ordinal out of range");
  ORX_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  ORX_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    orx_set_error("orx_create: liborx is built for sm_100a only; device %d is sm_%d%d", device, prop.major,
                  prop.minor);
    return ORX_ERR_UNSUPPORTED;
  }
  orx_ctx* c = new orx_ctx();
  memset(c, 0, sizeof(*c));
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  c->cap_partials = c->num_sms * 64;
  if (cudaMalloc(&c->counters, sizeof(int32_t) * 16) != cudaSuccess ||
      cudaMalloc(&c->partials, sizeof(float) * 2 * (size_t)c->cap_partials) != cudaSuccess ||
      cudaMalloc(&c->out_stage[0], sizeof(float) * 8) != cudaSuccess ||
      cudaMalloc(&c->out_stage[1], sizeof(float) * 8) != cudaSuccess) {
    orx_set_error("orx_create: workspace allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
    delete c;
    return ORX_ERR_NOMEM;
  }
  cudaMemset(c->counters, 0, sizeof(int32_t) * 16);
  *out = c;
  return ORX_OK;
}

extern "C" int orx_destroy(orx_handle_t h) {
  if (!h) return ORX_OK;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  free_workspace(h);
  for (int i = 0; i < 2; ++i) {
    cudaFree(h->ids_stage[i]);
    cudaFree(h->out_stage[i]);
  }
  cudaFree(h->counters);
  cudaFree(h->partials);
  cudaFree(h->bucket_cursor);
  orx_shard_ws_release(h);
  if (h->side_stream) {
    cudaStreamDestroy(h->side_stream);
    for (int i = 0; i < 2; ++i) {
      cudaEventDestroy(h->side_ev[i]);
      cudaEventDestroy(h->pf_done[i]);
      cudaEventDestroy(h->pf_free[i]);
      cudaEventDestroy(h->stage_free[i]);
    }
  }
  if (h->prof_ev) {
    for (int i = 0; i < h->prof_cap * ORX_PROF_EV; ++i) cudaEventDestroy(h->prof_ev[i]);
    delete[] h->prof_ev;
  }
  delete h;
  return ORX_OK;
}

extern "C" int orx_profile_enable(orx_handle_t h, int32_t on) {
  ORX_REQUIRE(h != nullptr, "null handle");
  ORX_CUDA(cudaSetDevice(h->device));
  if (on && !h->prof_ev) {
    h->prof_cap = 1024;
    h->prof_ev = new cudaEvent_t[h->prof_cap * ORX_PROF_EV];
    for (int i = 0; i < h->prof_cap * ORX_PROF_EV; ++i) ORX_CUDA(cudaEventCreate(&h->prof_ev[i]));
  }
  h->prof_on = on ? 1 : 0;
  h->prof_n = 0;
  h->prof_step = 0;
  return ORX_OK;
}

extern "C" int orx_profile_read(orx_handle_t h, float* ms, int32_t n_phases, int32_t* n_steps) {
  ORX_REQUIRE(h != nullptr && ms && n_steps && n_phases >= 1 && n_phases < ORX_PROF_EV, "bad arguments");
  ORX_CUDA(cudaSetDevice(h->device));
  for (int k = 0; k < n_phases; ++k) ms[k] = 0.f;
  *n_steps = h->prof_n;
  for (int i = 0; i < h->prof_n; ++i) {
    ORX_CUDA(cudaEventSynchronize(h->prof_ev[i * ORX_PROF_EV + n_phases]));
    for (int k = 0; k < n_phases; ++k) {
      float t = 0.f;
      ORX_CUDA(cudaEventElapsedTime(&t, h->prof_ev[i * ORX_PROF_EV + k], h->prof_ev[i * ORX_PROF_EV + k + 1]));
      ms[k] += t;
    }
  }
  h->prof_n = 0;
  return ORX_OK;
}

extern "C" int orx_stream_synchronize(orx_handle_t h, orx_stream_t s) {
  ORX_REQUIRE(h != nullptr, "null handle");
  ORX_CUDA(cudaStreamSynchronize((cudaStream_t)s));
  return ORX_OK;
}

OrxOptDev orx_opt_to_dev(const orx_opt_t* o) {
  OrxOptDev d;
  d.kind = o->kind;
  d.lr = o->lr;
  d.eps = o->eps;
  d.beta1 = o->beta1;
  d.beta2 = o->beta2;
  if (o->kind == ORX_OPT_ADAM_LAZY || o->kind == ORX_OPT_ADAM_DENSE) {
    // lr_t = lr*sqrt(1-b2^t)/(1-b1^t), evaluated in double like the oracle's adam_lr_t
    double t = (double)(o->step < 1 ? 1 : o->step);
    d.lr = (float)((double)o->lr * sqrt(1.0 - pow((double)o->beta2, t)) / (1.0 - pow((double)o->beta1, t)));
  }
  return d;
}
