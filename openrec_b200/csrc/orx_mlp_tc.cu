// orx_mlp_tc.cu -- Dense-layer GEMMs on the 5th-gen tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM).
//
// DLRM's MLPs (openrec/tf2/modules/multi_layer_perceptron.py:5-18, recommenders/dlrm.py:34-37,87,90-95) are the one
// dense contraction on the path.  The parity bar is 1e-5 against an fp32 reference, which plain TF32 (10-bit
// mantissa) cannot meet, so every fp32 operand is split on the fly into two TF32 terms (hi = top 19 bits, lo = the
// TF32 of the exact remainder) and  C += Ahi*Bhi + Ahi*Blo + Alo*Bhi  is accumulated in fp32 in TMEM (3xTF32, relative
// error ~2^-21 per product).
//
// One CTA (128 threads) computes a 128 x 128 tile of  C[M,N] = op(A)[M,K] * op(B)[K,N]:
//   * K is consumed in blocks of 32: the four warps copy the A and B blocks from global memory (any layout: TA/TB as
//     in orx_dlrm.cu) into shared memory in the canonical K-major, no-swizzle UMMA layout
//     (8-row x 16-byte core matrices; LBO = 128 B between K-adjacent cores, SBO = 1 KB between 8-row groups),
//     splitting into hi / lo tiles as they go; double-buffered, the global loads of block k+1 overlap the MMAs of k;
//   * one elected thread issues 4 k-steps x 3 tcgen05.mma (M=128, N=128, K=8) per block, then tcgen05.commit to the
//     stage's mbarrier (which frees that stage for the copy of block k+2);
//   * epilogue: tcgen05.ld 32x32b (each warp owns 32 TMEM lanes = 32 rows), + bias, activation, store.
#include <stdlib.h>

#include "orx_common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 32;
constexpr int TILE_BYTES = BM * BK * 4;        // 16 KB: one operand tile (hi or lo)
constexpr int STAGE_BYTES = 4 * TILE_BYTES;    // A_hi, A_lo, B_hi, B_lo
constexpr int NSTAGE = 2;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp: SmemDescriptor)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);                  // start address        bits [0,14)
  d |= (uint64_t)(128 >> 4) << 16;                          // leading byte offset  bits [16,30): K-adjacent core matrices
  d |= (uint64_t)((BK / 4) * 128 >> 4) << 32;               // stride byte offset   bits [32,46): next 8-row group
  d |= (uint64_t)1 << 46;                                   // descriptor version (sm_100)
  return d;                                                 // base offset 0, layout type SWIZZLE_NONE (bits 61-63 = 0)
}

// kind::tf32, D = F32, A/B = TF32, both K-major, M = 128, N = 128 (InstrDescriptor bit layout)
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// smem byte offset of element (row r, k) inside one K-major no-swizzle operand tile
__device__ __forceinline__ int tile_off(int r, int k) { return (r >> 3) * ((BK / 4) * 128) + (k >> 2) * 128 + (r & 7) * 16 + (k & 3) * 4; }

// hi = TF32(v) and lo = TF32(v - hi), both rounded to nearest (cvt.rna): truncation would bias every product
// the same way and the error would grow linearly with K instead of with sqrt(K).
__device__ __forceinline__ float to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}
__device__ __forceinline__ void split_tf32(float v, float* hi, float* lo) {
  const float h = to_tf32(v);
  *hi = h;
  *lo = to_tf32(v - h);
}

// The tensor core accumulates in fp32 with truncation; over hundreds of accumulations that bias grows linearly
// with K (measured 3e-4 abs at K=1024).  So a TMEM accumulator only ever sums GROUP_KB k-blocks (64 K-elements,
// 24 MMAs); it is then drained with tcgen05.ld into fp32 REGISTERS (round-to-nearest adds, one output row per
// thread).  Two TMEM accumulators alternate, so the drain of group g-1 overlaps the MMAs of group g.
constexpr int GROUP_KB = 2;

// ---------------------------------------------------------------------------------------
// Version 2 of the tile kernel (default): same tcgen05 / TMEM / descriptor code as k_gemm_tc above, different division
// of labour, after the r1n ncu capture showed v1 latency-bound in its staging loop (7 % SM busy, 6 % of a wave's warps):
//   * 256 threads: both warpgroups stage (half the loads per thread, all of a k-block's global loads -- A and B tile --
//     are issued before the first conversion), warps 0-3 drain accumulator columns 0..63, warps 4-7 columns 64..127
//     (a warp reaches TMEM lanes 32*(warp%4)..+31);
//   * [K, rows] sources (w in the forward pass, x and dz in dw = x^T dz) are read as 4 coalesced scalars per 16-byte
//     core-matrix row and stored with one conflict-free float4 (v1: 4-way conflicted scalar stores);
//   * epilogue through shared memory: coalesced 128-byte row segments instead of one row per thread;
//   * split-K: blockIdx.z takes a contiguous range of k-blocks and writes a raw partial tile; dw = x^T dz has
//     K = batch (32768) and only (in/128) x (out/128) tiles, i.e. 8-64 CTAs for 148 SMs without it.
// ---------------------------------------------------------------------------------------
constexpr int NT2 = 256;

// [rows, K] row-major source, k contiguous: 1024 float4 per tile, 4 per thread
__device__ __forceinline__ void load_rowmajor(const float* __restrict__ src, int64_t ld, int r0, int k0, int R, int K, float4 (&v)[4]) {
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int e = it * NT2 + threadIdx.x;
    const int r = e >> 3, kq = (e & 7) * 4;
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r0 + r < R) {
      const float* p = src + (int64_t)(r0 + r) * ld + k0 + kq;
      if (k0 + kq + 3 < K && ((((uintptr_t)p) & 15) == 0)) {
        q = __ldg(reinterpret_cast<const float4*>(p));
      } else {
        if (k0 + kq + 0 < K) q.x = __ldg(p + 0);
        if (k0 + kq + 1 < K) q.y = __ldg(p + 1);
        if (k0 + kq + 2 < K) q.z = __ldg(p + 2);
        if (k0 + kq + 3 < K) q.w = __ldg(p + 3);
      }
    }
    v[it] = q;
  }
}
// [K, rows] row-major source, rows contiguous: item (r, kq) = 4 scalars src[(k0+4kq+i)*ld + r0+r]; 1024 items, 4 per thread;
// consecutive threads walk r (coalesced), a warp shares kq
__device__ __forceinline__ void load_colmajor(const float* __restrict__ src, int64_t ld, int r0, int k0, int R, int K, float4 (&v)[4]) {
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int e = it * NT2 + threadIdx.x;
    const int r = e & 127, kq = (e >> 7) * 4;
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r0 + r < R) {
      const float* p = src + (int64_t)(k0 + kq) * ld + r0 + r;
      if (k0 + kq + 0 < K) q.x = __ldg(p);
      if (k0 + kq + 1 < K) q.y = __ldg(p + ld);
      if (k0 + kq + 2 < K) q.z = __ldg(p + 2 * ld);
      if (k0 + kq + 3 < K) q.w = __ldg(p + 3 * ld);
    }
    v[it] = q;
  }
}
template <int T>
__device__ __forceinline__ void store_split(const float4 (&v)[4], unsigned char* hi_tile, unsigned char* lo_tile) {
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int e = it * NT2 + threadIdx.x;
    const int r = T == 0 ? (e >> 3) : (e & 127);
    const int kq = T == 0 ? (e & 7) * 4 : (e >> 7) * 4;
    float4 h, l;
    split_tf32(v[it].x, &h.x, &l.x); split_tf32(v[it].y, &h.y, &l.y); split_tf32(v[it].z, &h.z, &l.z); split_tf32(v[it].w, &h.w, &l.w);
    const int off = tile_off(r, kq);
    *reinterpret_cast<float4*>(hi_tile + off) = h;
    *reinterpret_cast<float4*>(lo_tile + off) = l;
  }
}

template <int TA, int TB>
__global__ void __launch_bounds__(NT2, 1) k_gemm_tc2(const float* __restrict__ A, int64_t lda, const float* __restrict__ Bm,
                                                     int64_t ldb, float* __restrict__ C, int64_t ldc, int M, int N, int K,
                                                     const float* __restrict__ bias, int act, float* __restrict__ part) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t mbar[NSTAGE];
  __shared__ __align__(8) uint64_t accbar[2];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  constexpr int HN = BN / 2;                       // accumulator columns per warpgroup
  const int chalf = (warp >> 2) * HN;              // this warp's column half
  const int lane_base = (warp & 3) * 32;           // this warp's TMEM lanes

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) mbar_init(&mbar[s], 1);
    mbar_init(&accbar[0], 1);
    mbar_init(&accbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_s;

  float acc[HN];
#pragma unroll
  for (int j = 0; j < HN; ++j) acc[j] = 0.f;

  auto drain = [&](int g) {
    const int buf = g & 1;
    mbar_wait(&accbar[buf], (g >> 1) & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
    for (int c0 = 0; c0 < HN; c0 += 32) {
      uint32_t v[32];
      const uint32_t taddr = tmem_d + ((uint32_t)lane_base << 16) + (uint32_t)(buf * BN + chalf + c0);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
            "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
            "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
            "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[c0 + j] += __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  };

  // this CTA's range of k-blocks (split-K over blockIdx.z)
  const int nkb_all = (K + BK - 1) / BK;
  const int per = (nkb_all + (int)gridDim.z - 1) / (int)gridDim.z;
  const int kb_lo = blockIdx.z * per;
  const int kb_hi = min(nkb_all, kb_lo + per);
  const int nkb = max(0, kb_hi - kb_lo);
  const int ngroups = (nkb + GROUP_KB - 1) / GROUP_KB;
  float4 va[4], vb[4];   // register stage: the global loads of k-block kb+1 are in flight while kb is converted, issued
                         // to the tensor core and the previous accumulation group is drained
  auto load_block = [&](int kb) {
    const int k0 = (kb_lo + kb) * BK;
    if (TA == 0) load_rowmajor(A, lda, m0, k0, M, K, va); else load_colmajor(A, lda, m0, k0, M, K, va);
    if (TB == 1) load_rowmajor(Bm, ldb, n0, k0, N, K, vb); else load_colmajor(Bm, ldb, n0, k0, N, K, vb);
  };
  if (nkb > 0) load_block(0);
  for (int kb = 0; kb < nkb; ++kb) {
    const int s = kb % NSTAGE;
    const int g = kb / GROUP_KB;
    const bool g_first = (kb % GROUP_KB) == 0, g_last = (kb % GROUP_KB) == GROUP_KB - 1 || kb == nkb - 1;
    unsigned char* st = smem + (size_t)s * STAGE_BYTES;
    if (kb >= NSTAGE) mbar_wait(&mbar[s], ((kb / NSTAGE) - 1) & 1);   // the MMAs of block kb-NSTAGE are done with this stage
    store_split<TA>(va, st, st + TILE_BYTES);
    store_split<(TB == 1 ? 0 : 1)>(vb, st + 2 * TILE_BYTES, st + 3 * TILE_BYTES);
    if (kb + 1 < nkb) load_block(kb + 1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_hi = smem_u32(st), a_lo = a_hi + TILE_BYTES, b_hi = a_hi + 2 * TILE_BYTES, b_lo = a_hi + 3 * TILE_BYTES;
      const uint32_t d = tmem_d + (uint32_t)((g & 1) * BN);
#pragma unroll
      for (int ks = 0; ks < BK / 8; ++ks) {
        const uint32_t o = ks * 256;
        mma_tf32(d, make_desc(a_hi + o), make_desc(b_hi + o), (g_first && ks == 0) ? 0u : 1u);
        mma_tf32(d, make_desc(a_hi + o), make_desc(b_lo + o), 1u);
        mma_tf32(d, make_desc(a_lo + o), make_desc(b_hi + o), 1u);
      }
      umma_commit(&mbar[s]);
      if (g_last) umma_commit(&accbar[g & 1]);
    }
    if (g_first && g >= 1) drain(g - 1);
  }
  if (ngroups > 0) drain(ngroups - 1);

  // epilogue through shared memory (every MMA has completed: the last drain waited for the last group, and groups
  // complete in order): tile[128][BN + 1] floats, then coalesced row segments
  __syncthreads();
  float* tile = reinterpret_cast<float*>(smem);
  {
    const int r = lane_base + lane;
#pragma unroll
    for (int j = 0; j < HN; ++j) tile[r * (BN + 1) + chalf + j] = acc[j];
  }
  __syncthreads();
  const bool split = gridDim.z > 1;
  float* out = split ? part + (size_t)blockIdx.z * (size_t)M * (size_t)N : C;
  const int64_t ldo = split ? (int64_t)N : ldc;
  for (int r = warp; r < BM; r += NT2 / 32) {
    const int m = m0 + r;
    if (m >= M) break;
#pragma unroll
    for (int c = 0; c < BN; c += 32) {
      const int n = n0 + c + lane;
      if (n < N) {
        float x = tile[r * (BN + 1) + c + lane];
        if (!split) {
          x += bias ? bias[n] : 0.f;
          if (act == 1) x = fmaxf(x, 0.f);
          else if (act == 2) x = orx_sigmoid(x);
        }
        out[(int64_t)m * ldo + n] = x;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_d) : "memory");
}

// sum of split-K partials (deterministic: fixed order), + bias, activation
__global__ void __launch_bounds__(256) k_splitk_reduce(const float* __restrict__ part, int S, int M, int N, float* __restrict__ C,
                                                       int64_t ldc, const float* __restrict__ bias, int act) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)M * N) return;
  const int m = (int)(i / N), n = (int)(i % N);
  float x = 0.f;
  for (int z = 0; z < S; ++z) x += part[(size_t)z * (size_t)M * (size_t)N + (size_t)i];
  x += bias ? bias[n] : 0.f;
  if (act == 1) x = fmaxf(x, 0.f);
  else if (act == 2) x = orx_sigmoid(x);
  C[(int64_t)m * ldc + n] = x;
}

}  // namespace

// split-K workspace: one buffer per device (indexed by the current device), grown on demand, reused by every GEMM /
// column sum of that device's stream in order
static float* g_part[64] = {nullptr};
static size_t g_part_floats[64] = {0};
float* orx_splitk_workspace(size_t floats) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (floats > g_part_floats[dev]) {
    cudaDeviceSynchronize();
    cudaFree(g_part[dev]);
    g_part[dev] = nullptr;
    g_part_floats[dev] = 0;
    if (cudaMalloc(&g_part[dev], sizeof(float) * floats) != cudaSuccess) return nullptr;
    g_part_floats[dev] = floats;
  }
  return g_part[dev];
}
int orx_launch_splitk_reduce(const float* part, int S, int M, int N, float* C, int64_t ldc, const float* bias, int act, cudaStream_t st) {
  const int64_t n = (int64_t)M * N;
  k_splitk_reduce<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(part, S, M, N, C, ldc, bias, act);
  ORX_LAUNCH_CHECK();
  return ORX_OK;
}

// C[M,N] = op(A) * op(B) (+bias, act) on tcgen05; same operand conventions as launch_gemm in orx_dlrm.cu.
// Returns ORX_ERR_UNSUPPORTED for shapes that are better left to the SIMT kernel (tiny N or K).
int orx_launch_gemm_tc(int TA, int TB, const float* A, int64_t lda, const float* Bm, int64_t ldb, float* C, int64_t ldc,
                       int M, int N, int K, const float* bias, int act, cudaStream_t st) {
  if (N < 16 || K < 8 || M < 64) return ORX_ERR_UNSUPPORTED;
  const size_t smem = (size_t)NSTAGE * STAGE_BYTES + 1024;
  {
    const int tiles = ((N + BN - 1) / BN) * ((M + BM - 1) / BM);
    const int nkb = (K + BK - 1) / BK;
    int S = 1;
    if (tiles < 148 && nkb >= 16) {           // too few tiles for the machine and a long K: split it
      S = (2 * 148 + tiles - 1) / tiles;
      if (S > nkb / 4) S = nkb / 4;           // at least 4 k-blocks (two accumulation groups) per split
      if (S < 1) S = 1;
    }
    float* part = nullptr;
    if (S > 1) {
      part = orx_splitk_workspace((size_t)S * (size_t)M * (size_t)N);
      if (!part) { orx_set_error("split-K workspace allocation failed"); return ORX_ERR_CUDA; }
    }
    dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, S);
#define ORX_TC2(ta, tb)                                                                                           \
  {                                                                                                               \
    static bool done = false;                                                                                     \
    if (!done) {                                                                                                  \
      ORX_CUDA(cudaFuncSetAttribute(k_gemm_tc2<ta, tb>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      done = true;                                                                                                \
    }                                                                                                             \
    k_gemm_tc2<ta, tb><<<grid, NT2, smem, st>>>(A, lda, Bm, ldb, C, ldc, M, N, K, bias, act, part);                \
  }
    if (TA == 0 && TB == 0) ORX_TC2(0, 0)
    else if (TA == 0 && TB == 1) ORX_TC2(0, 1)
    else if (TA == 1 && TB == 0) ORX_TC2(1, 0)
    else ORX_TC2(1, 1)
#undef ORX_TC2
    ORX_LAUNCH_CHECK();
    if (S > 1) return orx_launch_splitk_reduce(part, S, M, N, C, ldc, bias, act, st);
    return ORX_OK;
  }
}
