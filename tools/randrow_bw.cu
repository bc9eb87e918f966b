// randrow_bw.cu -- what can B200 HBM deliver for RANDOM 512-byte rows?  (ceiling for gather/scatter kernels)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o randrow_bw tools/randrow_bw.cu && ./randrow_bw
// modes: stream copy | random row read | random row read-modify-write (distinct rows) | 6-row RMW mix like Adagrad BPR
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__global__ void k_copy(const float4* __restrict__ a, float4* __restrict__ b, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) b[i] = a[i];
}
// one warp per row, UNROLL rows in flight per warp; rows of 128 floats
template <int UNROLL, int MODE>  // MODE 0: read (reduce), 1: rmw
__global__ void __launch_bounds__(256) k_rows(float* __restrict__ tab, const int* __restrict__ ids, int n, float* sink) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  float acc = 0.f;
  for (int r = warp * UNROLL; r < n; r += nw * UNROLL) {
    float4 v[UNROLL];
    int id[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) id[u] = r + u < n ? __ldg(ids + r + u) : 0;
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) v[u] = __ldcg(reinterpret_cast<const float4*>(tab + (size_t)id[u] * 128) + lane);
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      if (MODE == 0) acc += v[u].x + v[u].y + v[u].z + v[u].w;
      else {
        v[u].x += 1.f; v[u].y += 1.f; v[u].z += 1.f; v[u].w += 1.f;
        if (r + u < n) __stcg(reinterpret_cast<float4*>(tab + (size_t)id[u] * 128) + lane, v[u]);
      }
    }
  }
  if (MODE == 0 && acc == 123.456f) *sink = acc;
}

int main() {
  const size_t ROWS = 8u << 20;  // 8M rows x 512 B = 4 GB
  const int N = 6 * 65536 * 8;   // rows touched per launch (8 BPR-Adagrad steps' worth)
  float *tab, *tab2, *sink;
  int* ids;
  CK(cudaMalloc(&tab, ROWS * 512)); CK(cudaMalloc(&tab2, ROWS * 512 / 4)); CK(cudaMalloc(&ids, sizeof(int) * N)); CK(cudaMalloc(&sink, 4));
  CK(cudaMemset(tab, 0, ROWS * 512));
  std::vector<int> h(N);
  // distinct random rows (a random permutation prefix) so RMW has no conflicts
  std::vector<int> perm(ROWS);
  for (size_t i = 0; i < ROWS; ++i) perm[i] = (int)i;
  uint64_t s = 88172645463325252ull;
  for (int i = 0; i < N; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; size_t j = i + s % (ROWS - i); std::swap(perm[i], perm[j]); h[i] = perm[i]; }
  CK(cudaMemcpy(ids, h.data(), sizeof(int) * N, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto time = [&](auto f, const char* name, double bytes) {
    for (int i = 0; i < 3; ++i) f();
    float best = 1e9;
    for (int i = 0; i < 10; ++i) { CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms; }
    printf("%-44s %8.3f ms  %8.1f GB/s\n", name, best, bytes / best / 1e6);
  };
  const size_t n4 = ROWS * 512 / 4 / 16;  // 1 GB copy
  time([&] { k_copy<<<148 * 16, 256>>>((const float4*)tab, (float4*)tab2, n4); }, "stream copy 1 GB (read+write bytes)", 2.0 * n4 * 16);
  const int grid = 148 * 8;
  time([&] { k_rows<4, 0><<<grid, 256>>>(tab, ids, N, sink); }, "random 512B row read, 4 rows/warp in flight", (double)N * 512);
  time([&] { k_rows<8, 0><<<grid, 256>>>(tab, ids, N, sink); }, "random 512B row read, 8 rows/warp in flight", (double)N * 512);
  time([&] { k_rows<8, 0><<<grid * 2, 256>>>(tab, ids, N, sink); }, "random 512B row read, 8/warp, 2x grid", (double)N * 512);
  time([&] { k_rows<4, 1><<<grid, 256>>>(tab, ids, N, sink); }, "random 512B row read-modify-write, 4/warp", 2.0 * N * 512);
  time([&] { k_rows<8, 1><<<grid, 256>>>(tab, ids, N, sink); }, "random 512B row read-modify-write, 8/warp", 2.0 * N * 512);
  time([&] { k_rows<8, 1><<<grid * 2, 256>>>(tab, ids, N, sink); }, "random 512B row RMW, 8/warp, 2x grid", 2.0 * N * 512);
  return 0;
}
