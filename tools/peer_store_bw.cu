// Peer-STORE bandwidth by access pattern, 2 GPUs, single process (cudaDeviceEnablePeerAccess), both directions at once.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/peer_store_bw tools/peer_store_bw.cu && ./tools/peer_store_bw
// Each GPU's kernel reads LOCAL source rows and stores them into the OTHER GPU's buffer; rows = 196608.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// one warp per row (grid-stride); W floats per row (multiple of 4); perm: destination row of source row j (or null = j)
// split_bias: the row's last float4 is NOT written as a row tail; a single float goes to bias[dest] instead
__global__ void __launch_bounds__(256) k_rows(const float* __restrict__ src, float* __restrict__ dst, int n, int Wsrc, int Wdst,
                                              int nq, const int32_t* __restrict__ perm, float* __restrict__ bias, int unroll8) {
  const int lane = threadIdx.x & 31;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  for (int j0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 8; j0 < n; j0 += nw * 8) {
    float4 v0[8], v1[8];
    int d[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int j = j0 + k;
      d[k] = -1;
      if (j < n) {
        d[k] = perm ? perm[j] : j;
        const float4* s = reinterpret_cast<const float4*>(src + (int64_t)j * Wsrc);
        if (lane < nq) v0[k] = __ldcg(s + lane);
        if (lane + 32 < nq) v1[k] = __ldcg(s + 32 + lane);
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (d[k] < 0) continue;
      float4* t = reinterpret_cast<float4*>(dst + (int64_t)d[k] * Wdst);
      if (lane < nq) t[lane] = v0[k];
      if (lane + 32 < nq) t[32 + lane] = v1[k];
      if (bias && lane == 0) bias[d[k]] = v0[k].x;
    }
  }
}

__global__ void k_copy(const float4* __restrict__ s, float4* __restrict__ d, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) d[i] = __ldcg(s + i);
}

int main() {
  int nd = 0;
  CK(cudaGetDeviceCount(&nd));
  if (nd < 2) { printf("needs 2 GPUs\n"); return 0; }
  const int n = 196608, WMAX = 160;
  float *src[2], *dst[2], *bias[2];
  int32_t *perm[2];
  cudaStream_t st[2];
  cudaEvent_t e0[2], e1[2];
  std::vector<int32_t> hp(n);
  for (int i = 0; i < n; ++i) hp[i] = i;
  srand(1);
  for (int i = n - 1; i > 0; --i) { int j = rand() % (i + 1); int t = hp[i]; hp[i] = hp[j]; hp[j] = t; }
  for (int g = 0; g < 2; ++g) {
    CK(cudaSetDevice(g));
    CK(cudaDeviceEnablePeerAccess(1 - g, 0));
    CK(cudaMalloc(&src[g], sizeof(float) * (size_t)n * WMAX));
    CK(cudaMalloc(&dst[g], sizeof(float) * (size_t)n * WMAX));
    CK(cudaMalloc(&bias[g], sizeof(float) * n));
    CK(cudaMalloc(&perm[g], sizeof(int32_t) * n));
    CK(cudaMemset(src[g], 1, sizeof(float) * (size_t)n * WMAX));
    CK(cudaMemcpy(perm[g], hp.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice));
    CK(cudaStreamCreate(&st[g]));
    CK(cudaEventCreate(&e0[g]));
    CK(cudaEventCreate(&e1[g]));
  }
  struct Case { const char* name; int Wsrc, Wdst, nq; bool random, split; bool local; };
  const Case cases[] = {
      {"rows 512B aligned, sequential dst        ", 128, 128, 32, false, false, false},
      {"rows 528B (W=132), sequential dst        ", 132, 132, 33, false, false, false},
      {"rows 512B aligned, RANDOM dst            ", 128, 128, 32, true, false, false},
      {"rows 528B (W=132), RANDOM dst            ", 132, 132, 33, true, false, false},
      {"rows 512B + 4B bias array, sequential dst", 128, 128, 32, false, true, false},
      {"rows 512B + 4B bias array, RANDOM dst    ", 128, 128, 32, true, true, false},
      {"rows 640B (W=160) aligned, RANDOM dst    ", 160, 160, 40, true, false, false},
      {"rows 576B (W=144) RANDOM dst             ", 144, 144, 36, true, false, false},
      {"LOCAL rows 528B RANDOM dst (no NVLink)   ", 132, 132, 33, true, false, true},
  };
  const int iters = 20;
  for (const Case& c : cases) {
    for (int both = 0; both < 2; ++both) {
      float ms[2] = {0, 0};
      for (int rep = 0; rep < 2; ++rep) {   // rep 0 = warm-up
        for (int g = 0; g < (both ? 2 : 1); ++g) {
          CK(cudaSetDevice(g));
          CK(cudaEventRecord(e0[g], st[g]));
          for (int it = 0; it < iters; ++it)
            k_rows<<<148 * 8, 256, 0, st[g]>>>(src[g], c.local ? dst[g] : dst[1 - g], n, c.Wsrc, c.Wdst, c.nq, c.random ? perm[g] : nullptr,
                                              c.split ? (c.local ? bias[g] : bias[1 - g]) : nullptr, 1);
          CK(cudaEventRecord(e1[g], st[g]));
        }
        for (int g = 0; g < (both ? 2 : 1); ++g) {
          CK(cudaSetDevice(g));
          CK(cudaEventSynchronize(e1[g]));
          CK(cudaEventElapsedTime(&ms[g], e0[g], e1[g]));
        }
      }
      const double bytes = (double)n * c.nq * 16 + (c.split ? 4.0 * n : 0.0);
      printf("%s %s : %7.1f us/launch  %6.1f GB/s per direction\n", c.name, both ? "both dirs" : "one dir  ", ms[0] / iters * 1e3,
             bytes / (ms[0] / iters * 1e-3) / 1e9);
    }
  }
  // bulk copy kernel for reference
  for (int both = 0; both < 2; ++both) {
    float ms[2];
    for (int rep = 0; rep < 2; ++rep) {
      for (int g = 0; g < (both ? 2 : 1); ++g) {
        CK(cudaSetDevice(g));
        CK(cudaEventRecord(e0[g], st[g]));
        for (int it = 0; it < iters; ++it) k_copy<<<148 * 8, 256, 0, st[g]>>>((const float4*)src[g], (float4*)dst[1 - g], (int64_t)n * 32);
        CK(cudaEventRecord(e1[g], st[g]));
      }
      for (int g = 0; g < (both ? 2 : 1); ++g) {
        CK(cudaSetDevice(g));
        CK(cudaEventSynchronize(e1[g]));
        CK(cudaEventElapsedTime(&ms[g], e0[g], e1[g]));
      }
    }
    printf("bulk float4 copy kernel (100 MB)            %s : %7.1f us/launch  %6.1f GB/s per direction\n", both ? "both dirs" : "one dir  ",
           ms[0] / iters * 1e3, (double)n * 512 / (ms[0] / iters * 1e-3) / 1e9);
  }
  return 0;
}
