// orx_sampler.cu -- device-side pairwise sampler: the reference's Python generator
// (openrec/tf2/data/dataset.py:7-16 + data/utils.py:82-87,102-116) yields ~2.5e5 triplets/s/process, three to four
// orders of magnitude below the training kernel.  Same semantics on the device: records are consumed in the order of a
// per-epoch permutation; the negative is drawn uniformly over the catalogue and rejected while it is one of the user's
// positives (binary search in the user's sorted CSR row).  Counter-based RNG: (seed, slot, attempt) -> item.
#include "orx_common.cuh"

__device__ __forceinline__ uint64_t smix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__global__ void k_sample_pairwise(const int32_t* __restrict__ rec_user, const int32_t* __restrict__ rec_item,
                                  const int64_t* __restrict__ perm, int64_t cursor, int64_t n_records,
                                  const int64_t* __restrict__ csr_off, const int32_t* __restrict__ csr_items,
                                  int32_t total_items, uint64_t seed, int32_t B, int32_t* __restrict__ uid,
                                  int32_t* __restrict__ pid, int32_t* __restrict__ nid) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int64_t r = perm[(cursor + b) % n_records];
  const int32_t u = rec_user[r], p = rec_item[r];
  const int64_t lo0 = csr_off[u], hi0 = csr_off[u + 1];
  int32_t n = 0;
  for (uint32_t attempt = 0;; ++attempt) {
    n = (int32_t)(smix64(seed ^ smix64(((uint64_t)(cursor + b) << 20) | attempt)) % (uint64_t)total_items);
    int64_t lo = lo0, hi = hi0;          // is n one of u's positives?
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (csr_items[mid] < n) lo = mid + 1;
      else hi = mid;
    }
    if (!(lo < hi0 && csr_items[lo] == n)) break;
    if (attempt > 1000000u) break;       // a user positive on the whole catalogue: give up like a bounded reference
  }
  uid[b] = u;
  pid[b] = p;
  nid[b] = n;
}

extern "C" int orx_sample_pairwise(orx_handle_t h, const int32_t* rec_user, const int32_t* rec_item, const int64_t* perm,
                                   int64_t cursor, int64_t n_records, const int64_t* csr_off, const int32_t* csr_items,
                                   int32_t total_items, uint64_t seed, int32_t B, int32_t* uid, int32_t* pid,
                                   int32_t* nid, orx_stream_t s) {
  ORX_REQUIRE(h != nullptr && rec_user && rec_item && perm && csr_off && csr_items && uid && pid && nid, "null pointer");
  ORX_REQUIRE(n_records > 0 && total_items > 0 && B >= 0 && cursor >= 0, "bad sizes");
  if (B == 0) return ORX_OK;
  ORX_CUDA(cudaSetDevice(h->device));
  k_sample_pairwise<<<(B + 255) / 256, 256, 0, (cudaStream_t)s>>>(rec_user, rec_item, perm, cursor, n_records, csr_off,
                                                                   csr_items, total_items, seed, B, uid, pid, nid);
  ORX_LAUNCH_CHECK();
  return ORX_OK;
}
