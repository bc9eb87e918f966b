// Cost of the fences / flag handshakes used by the sharded step (1 or 2 GPUs).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/fence_cost tools/fence_cost.cu && ./tools/fence_cost
#include <cuda_runtime.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__global__ void k(float4* dst, int n_stores, int mode, long long* out) {
  // every thread stores n_stores float4 (block-contiguous), then thread 0 fences; clock around the fence
  for (int i = 0; i < n_stores; ++i) dst[((size_t)blockIdx.x * n_stores + i) * blockDim.x + threadIdx.x] = make_float4(1.f, 2.f, 3.f, 4.f);
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t0 = clock64();
    if (mode == 1) __threadfence();
    if (mode == 2) __threadfence_system();
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
}
int main() {
  int nd = 0; CK(cudaGetDeviceCount(&nd));
  CK(cudaSetDevice(0));
  float4 *loc, *peer = nullptr; long long* out; long long h[592];
  CK(cudaMalloc(&loc, 592ull * 256 * 64 * 16)); CK(cudaMalloc(&out, sizeof(h)));
  if (nd > 1) { CK(cudaDeviceEnablePeerAccess(1, 0)); CK(cudaSetDevice(1)); CK(cudaMalloc(&peer, 592ull * 256 * 64 * 16)); CK(cudaSetDevice(0)); }
  const char* names[3] = {"no fence", "__threadfence()", "__threadfence_system()"};
  for (int target = 0; target < (nd > 1 ? 2 : 1); ++target)
    for (int stores = 0; stores <= 64; stores += 32)
      for (int mode = 0; mode < 3; ++mode) {
        for (int rep = 0; rep < 3; ++rep) k<<<592, 256>>>(target ? peer : loc, stores, mode, out);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost));
        double s = 0; long long mx = 0; for (int i = 0; i < 592; ++i) { s += h[i]; if (h[i] > mx) mx = h[i]; }
        printf("%-5s stores/thread %2d  %-24s mean %8.0f clk  max %8lld clk\n", target ? "PEER" : "local", stores, names[mode], s / 592, mx);
      }
  return 0;
}
