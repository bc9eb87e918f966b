// orx_sampler.cu -- device-side samplers: the reference's Python generators
// (openrec/tf2/data/dataset.py:7-58 + data/utils.py:82-87,102-116) yield ~2.5e5 samples/s/process, three to four orders
// of magnitude below the training kernels.  Same semantics on the device:
//   * records are consumed in the order of a per-epoch random permutation and NEVER dropped: a batch that crosses the
//     end of an epoch takes the tail of the current permutation and the head of the next one (utils.py:82-87);
//   * pairwise (dataset.py:7-16): one uniform negative per record, rejected while it is one of the user's positives
//     (binary search in the user's sorted CSR row, utils.py:110-116);
//   * stratified_pointwise (dataset.py:18-34): a coin per sample -- the next record (label 1) with probability pos_ratio,
//     else a uniform (user, item) pair rejected while observed (label 0); the positives of a batch take consecutive
//     records, so the kernel is one block that scans the coins chunk by chunk;
//   * per_pos_stratified_pointwise (dataset.py:36-58): every record followed by int((1-r)/r) distinct items != the
//     positive (random.sample without replacement); the stream is cut into batches at arbitrary positions, so a sample
//     is a pure function of (seed, position in the stream).
// Counter-based RNG: every draw is smix64 of (seed, position, attempt); nothing is carried between launches.
#include "orx_common.cuh"

__device__ __forceinline__ uint64_t smix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__device__ __forceinline__ uint64_t srand3(uint64_t seed, uint64_t a, uint64_t b) { return smix64(seed ^ smix64((a << 24) ^ b)); }

struct SamplerData {
  const int32_t *rec_user, *rec_item;   // [n_records]
  const int64_t *perm_cur, *perm_next;  // permutations of the current and of the following epoch
  int64_t cursor, n_records;            // position inside perm_cur
  const int64_t* csr_off;               // [U+1]
  const int32_t* csr_items;             // positives of user u: csr_items[csr_off[u] .. csr_off[u+1]) sorted
  int32_t total_users, total_items;
};

// k-th record consumed from now on (k >= 0): the tail of the current permutation, then the next one
__device__ __forceinline__ int64_t sampler_record(const SamplerData& d, int64_t k) {
  const int64_t left = d.n_records - d.cursor;
  return k < left ? d.perm_cur[d.cursor + k] : d.perm_next[(k - left) % d.n_records];
}
__device__ __forceinline__ bool sampler_is_positive(const SamplerData& d, int32_t u, int32_t i) {
  int64_t lo = d.csr_off[u];
  const int64_t hi0 = d.csr_off[u + 1];
  int64_t hi = hi0;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (d.csr_items[mid] < i) lo = mid + 1;
    else hi = mid;
  }
  return lo < hi0 && d.csr_items[lo] == i;
}

__global__ void k_sample_pairwise(SamplerData d, uint64_t seed, int64_t pos0, int32_t B, int32_t* __restrict__ uid,
                                  int32_t* __restrict__ pid, int32_t* __restrict__ nid) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int64_t r = sampler_record(d, b);
  const int32_t u = d.rec_user[r], p = d.rec_item[r];
  int32_t n = 0;
  for (uint32_t attempt = 0;; ++attempt) {
    n = (int32_t)(srand3(seed, (uint64_t)(pos0 + b), attempt) % (uint64_t)d.total_items);
    if (!sampler_is_positive(d, u, n)) break;
    if (attempt > 1000000u) break;       // a user positive on the whole catalogue: give up like a bounded reference
  }
  uid[b] = u;
  pid[b] = p;
  nid[b] = n;
}

// one block: chunks of 1024 slots; the positives of the batch take consecutive records (exclusive scan of the coins).
// n_pos_out[0] = records consumed by this batch.
__global__ void __launch_bounds__(1024) k_sample_stratified(SamplerData d, uint64_t seed, int64_t pos0, int32_t B, float pos_ratio,
                                                            int32_t* __restrict__ uid, int32_t* __restrict__ iid,
                                                            float* __restrict__ label, int32_t* n_pos_out) {
  __shared__ int32_t warp_sum[32];
  __shared__ int64_t base;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int c0 = 0; c0 < B; c0 += 1024) {
    const int b = c0 + threadIdx.x;
    const bool live = b < B;
    // random.random() <= pos_ratio  (dataset.py:23): 53-bit uniform in [0, 1)
    const double coin = (double)(srand3(seed, (uint64_t)(pos0 + b), 0xC01Full) >> 11) * (1.0 / 9007199254740992.0);
    const bool pos = live && coin <= (double)pos_ratio;
    const unsigned m = __ballot_sync(0xffffffffu, pos);
    if (lane == 0) warp_sum[wid] = __popc(m);
    __syncthreads();
    int before = __popc(m & ((1u << lane) - 1u)), total = 0;
    for (int w = 0; w < 32; ++w) {
      const int s = warp_sum[w];
      if (w < wid) before += s;
      total += s;
    }
    if (live) {
      int32_t u, i;
      if (pos) {
        const int64_t r = sampler_record(d, base + before);
        u = d.rec_user[r];
        i = d.rec_item[r];
      } else {
        for (uint32_t attempt = 1;; ++attempt) {
          const uint64_t x = srand3(seed, (uint64_t)(pos0 + b), attempt);
          u = (int32_t)((x >> 32) % (uint64_t)d.total_users);
          i = (int32_t)((x & 0xffffffffull) % (uint64_t)d.total_items);
          if (!sampler_is_positive(d, u, i) || attempt > 1000000u) break;
        }
      }
      uid[b] = u;
      iid[b] = i;
      label[b] = pos ? 1.f : 0.f;
    }
    __syncthreads();
    if (threadIdx.x == 0) base += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_pos_out = (int32_t)base;
}

// stream position s = pos0 + b: group s / (quota + 1); member 0 = the group's record, member m >= 1 = its m-th negative.
// The negatives of a group: quota + 1 distinct uniform items, the positive removed, the first `quota` kept
// (random.sample(range(I), quota + 1) minus the positive, dataset.py:47-55).  rec0 = records consumed before pos0's group.
#define ORX_MAX_QUOTA 64
__global__ void k_sample_per_positive(SamplerData d, uint64_t seed, int64_t pos0, int32_t B, int32_t quota,
                                      int32_t* __restrict__ uid, int32_t* __restrict__ iid, float* __restrict__ label) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int64_t s = pos0 + b, g = quota + 1;
  const int64_t grp = s / g, grp0 = pos0 / g;
  const int m = (int)(s - grp * g);
  const int64_t r = sampler_record(d, grp - grp0);
  const int32_t u = d.rec_user[r], p = d.rec_item[r];
  int32_t item = p;
  if (m > 0) {
    int32_t cand[ORX_MAX_QUOTA + 1];
    int n = 0, kept = 0;
    uint32_t attempt = 0;
    while (n < quota + 1) {              // distinct candidates, in draw order
      const int32_t c = (int32_t)(srand3(seed, (uint64_t)grp, attempt++) % (uint64_t)d.total_items);
      bool dup = false;
      for (int k = 0; k < n; ++k) dup = dup || cand[k] == c;
      if (dup && attempt < 100000u) continue;
      cand[n++] = c;
    }
    for (int k = 0; k < quota + 1; ++k) {
      if (cand[k] == p) continue;
      if (++kept == m) { item = cand[k]; break; }
    }
  }
  uid[b] = u;
  iid[b] = item;
  label[b] = m == 0 ? 1.f : 0.f;
}

static int sampler_check(orx_handle_t h, const orx_sampler_t* sd) {
  ORX_REQUIRE(h != nullptr && sd != nullptr, "null handle / sampler data");
  ORX_REQUIRE(sd->rec_user && sd->rec_item && sd->perm_cur && sd->perm_next && sd->csr_off && sd->csr_items, "null pointer");
  ORX_REQUIRE(sd->n_records > 0 && sd->cursor >= 0 && sd->cursor <= sd->n_records && sd->total_users > 0 && sd->total_items > 0,
              "bad sizes");
  return ORX_OK;
}
static SamplerData sampler_dev(const orx_sampler_t* s) {
  SamplerData d;
  d.rec_user = s->rec_user; d.rec_item = s->rec_item; d.perm_cur = s->perm_cur; d.perm_next = s->perm_next;
  d.cursor = s->cursor; d.n_records = s->n_records; d.csr_off = s->csr_off; d.csr_items = s->csr_items;
  d.total_users = s->total_users; d.total_items = s->total_items;
  return d;
}

extern "C" int orx_sample_pairwise(orx_handle_t h, const orx_sampler_t* sd, uint64_t seed, int64_t stream_pos, int32_t B,
                                   int32_t* uid, int32_t* pid, int32_t* nid, orx_stream_t s) {
  int rc = sampler_check(h, sd);
  if (rc) return rc;
  ORX_REQUIRE(uid && pid && nid && B >= 0 && (int64_t)B <= sd->n_records, "bad batch (at most one epoch per batch)");
  if (B == 0) return ORX_OK;
  ORX_CUDA(cudaSetDevice(h->device));
  k_sample_pairwise<<<(B + 255) / 256, 256, 0, (cudaStream_t)s>>>(sampler_dev(sd), seed, stream_pos, B, uid, pid, nid);
  ORX_LAUNCH_CHECK();
  return ORX_OK;
}

extern "C" int orx_sample_stratified(orx_handle_t h, const orx_sampler_t* sd, uint64_t seed, int64_t stream_pos, int32_t B,
                                     float pos_ratio, int32_t* uid, int32_t* iid, float* label, int32_t* n_pos_out,
                                     orx_stream_t s) {
  int rc = sampler_check(h, sd);
  if (rc) return rc;
  ORX_REQUIRE(uid && iid && label && n_pos_out && B >= 0 && (int64_t)B <= sd->n_records, "bad batch (at most one epoch per batch)");
  ORX_REQUIRE(pos_ratio >= 0.f && pos_ratio <= 1.f, "pos_ratio must be in [0, 1]");
  if (B == 0) return ORX_OK;
  ORX_CUDA(cudaSetDevice(h->device));
  k_sample_stratified<<<1, 1024, 0, (cudaStream_t)s>>>(sampler_dev(sd), seed, stream_pos, B, pos_ratio, uid, iid, label, n_pos_out);
  ORX_LAUNCH_CHECK();
  return ORX_OK;
}

extern "C" int orx_sample_per_positive(orx_handle_t h, const orx_sampler_t* sd, uint64_t seed, int64_t stream_pos, int32_t B,
                                       int32_t quota, int32_t* uid, int32_t* iid, float* label, orx_stream_t s) {
  int rc = sampler_check(h, sd);
  if (rc) return rc;
  ORX_REQUIRE(uid && iid && label && B >= 0 && quota >= 0 && quota <= ORX_MAX_QUOTA && quota < sd->total_items, "bad batch / quota");
  ORX_REQUIRE((int64_t)B / (quota + 1) + 2 <= sd->n_records, "at most one epoch of records per batch");
  if (B == 0) return ORX_OK;
  ORX_CUDA(cudaSetDevice(h->device));
  k_sample_per_positive<<<(B + 127) / 128, 128, 0, (cudaStream_t)s>>>(sampler_dev(sd), seed, stream_pos, B, quota, uid, iid, label);
  ORX_LAUNCH_CHECK();
  return ORX_OK;
}
