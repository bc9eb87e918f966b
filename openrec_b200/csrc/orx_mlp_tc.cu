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

// A tile source: TA=0 -> A[m*lda + k] (row-major [M,K]); TA=1 -> A[k*lda + m] ([K,M] row-major).
// B tile source: TB=0 -> B[k*ldb + n] ([K,N] row-major); TB=1 -> B[n*ldb + k] ([N,K] row-major).
// `rows` counts along M (A) or N (B); both tiles are staged as [128 rows][32 k] K-major.
template <int T>  // T=0: source is [rows, K] row-major (k contiguous); T=1: source is [K, rows] row-major
__device__ __forceinline__ void stage_tile(const float* __restrict__ src, int64_t ld, int r0, int k0, int R, int K,
                                           unsigned char* hi_tile, unsigned char* lo_tile) {
  if (T == 0) {
    // 128 rows x 8 float4 per row = 1024 float4; 128 threads -> 8 each; consecutive threads walk k (coalesced)
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int e = it * 128 + threadIdx.x;
      const int r = e >> 3, kq = (e & 7) * 4;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (r0 + r < R) {
        const float* p = src + (int64_t)(r0 + r) * ld + k0 + kq;
        if (k0 + kq + 3 < K && ((((uintptr_t)p) & 15) == 0)) {
          const float4 q = *reinterpret_cast<const float4*>(p);
          v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        } else {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (k0 + kq + c < K) v[c] = p[c];
        }
      }
      float4 h, l;
      split_tf32(v[0], &h.x, &l.x); split_tf32(v[1], &h.y, &l.y); split_tf32(v[2], &h.z, &l.z); split_tf32(v[3], &h.w, &l.w);
      const int off = tile_off(r, kq);
      *reinterpret_cast<float4*>(hi_tile + off) = h;
      *reinterpret_cast<float4*>(lo_tile + off) = l;
    }
  } else {
    // source [K, rows]: consecutive threads walk rows (coalesced), 32 k x 128 rows = 4096 scalars, 32 per thread
#pragma unroll 8
    for (int it = 0; it < 32; ++it) {
      const int k = it, r = threadIdx.x;
      float v = 0.f;
      if (k0 + k < K && r0 + r < R) v = src[(int64_t)(k0 + k) * ld + r0 + r];
      float h, l;
      split_tf32(v, &h, &l);
      const int off = tile_off(r, k);
      *reinterpret_cast<float*>(hi_tile + off) = h;
      *reinterpret_cast<float*>(lo_tile + off) = l;
    }
  }
}

// The tensor core accumulates in fp32 with truncation; over hundreds of accumulations that bias grows linearly
// with K (measured 3e-4 abs at K=1024).  So a TMEM accumulator only ever sums GROUP_KB k-blocks (64 K-elements,
// 24 MMAs); it is then drained with tcgen05.ld into fp32 REGISTERS (round-to-nearest adds, one output row per
// thread).  Two TMEM accumulators alternate, so the drain of group g-1 overlaps the MMAs of group g.
constexpr int GROUP_KB = 2;

template <int TA, int TB>
__global__ void __launch_bounds__(128, 1) k_gemm_tc(const float* __restrict__ A, int64_t lda, const float* __restrict__ Bm,
                                                    int64_t ldb, float* __restrict__ C, int64_t ldc, int M, int N, int K,
                                                    const float* __restrict__ bias, int act) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t mbar[NSTAGE];   // smem stage free again (its MMAs finished)
  __shared__ __align__(8) uint64_t accbar[2];      // TMEM accumulator buffer complete (its group's MMAs finished)
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

  if (warp == 0) {   // TMEM: 2 x 128 fp32 accumulator columns
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

  float acc[BN];   // this thread's output row (m0 + 32*warp + lane), fp32 accumulation across groups
#pragma unroll
  for (int j = 0; j < BN; ++j) acc[j] = 0.f;

  // drain TMEM accumulator `buf` (group number g) into the register accumulators
  auto drain = [&](int g) {
    const int buf = g & 1;
    mbar_wait(&accbar[buf], (g >> 1) & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[32];
      const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * BN + c0);
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

  const int nkb = (K + BK - 1) / BK;
  const int ngroups = (nkb + GROUP_KB - 1) / GROUP_KB;
  for (int kb = 0; kb < nkb; ++kb) {
    const int s = kb % NSTAGE;
    const int g = kb / GROUP_KB;                 // accumulation group and its TMEM buffer
    const bool g_first = (kb % GROUP_KB) == 0, g_last = (kb % GROUP_KB) == GROUP_KB - 1 || kb == nkb - 1;
    unsigned char* st = smem + (size_t)s * STAGE_BYTES;
    if (kb >= NSTAGE) mbar_wait(&mbar[s], ((kb / NSTAGE) - 1) & 1);   // MMAs of block kb-NSTAGE are done with this stage
    stage_tile<TA>(A, lda, m0, kb * BK, M, K, st, st + TILE_BYTES);
    stage_tile<(TB == 1 ? 0 : 1)>(Bm, ldb, n0, kb * BK, N, K, st + 2 * TILE_BYTES, st + 3 * TILE_BYTES);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the MMA (async proxy)
    // group g reuses the TMEM buffer of group g-2: every thread must have drained g-2 before its first MMA.
    // (the drain of g-2 ran after group g-1's first block was issued, i.e. earlier in program order)
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_hi = smem_u32(st), a_lo = a_hi + TILE_BYTES, b_hi = a_hi + 2 * TILE_BYTES, b_lo = a_hi + 3 * TILE_BYTES;
      const uint32_t d = tmem_d + (uint32_t)((g & 1) * BN);
#pragma unroll
      for (int ks = 0; ks < BK / 8; ++ks) {           // one MMA consumes K = 8 (two 16-byte core matrices = 256 B)
        const uint32_t o = ks * 256;
        mma_tf32(d, make_desc(a_hi + o), make_desc(b_hi + o), (g_first && ks == 0) ? 0u : 1u);
        mma_tf32(d, make_desc(a_hi + o), make_desc(b_lo + o), 1u);
        mma_tf32(d, make_desc(a_lo + o), make_desc(b_hi + o), 1u);
      }
      umma_commit(&mbar[s]);                 // stage s may be overwritten once these MMAs are done
      if (g_last) umma_commit(&accbar[g & 1]);   // ... and the group's accumulator is complete
    }
    // overlap: while group g's MMAs run, fold the previous group's accumulator into the registers
    if (g_first && g >= 1) drain(g - 1);
  }
  drain(ngroups - 1);

  // epilogue from registers: thread = one row
  const int m = m0 + warp * 32 + lane;
  if (m < M) {
#pragma unroll
    for (int j = 0; j < BN; ++j) {
      const int n = n0 + j;
      if (n < N) {
        float x = acc[j] + (bias ? bias[n] : 0.f);
        if (act == 1) x = fmaxf(x, 0.f);
        else if (act == 2) x = orx_sigmoid(x);
        C[(int64_t)m * ldc + n] = x;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_d) : "memory");
}

}  // namespace

// C[M,N] = op(A) * op(B) (+bias, act) on tcgen05; same operand conventions as launch_gemm in orx_dlrm.cu.
// Returns ORX_ERR_UNSUPPORTED for shapes that are better left to the SIMT kernel (tiny N or K).
int orx_launch_gemm_tc(int TA, int TB, const float* A, int64_t lda, const float* Bm, int64_t ldb, float* C, int64_t ldc,
                       int M, int N, int K, const float* bias, int act, cudaStream_t st) {
  if (N < 16 || K < 8 || M < 64) return ORX_ERR_UNSUPPORTED;
  const size_t smem = (size_t)NSTAGE * STAGE_BYTES + 1024;
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
#define ORX_TC(ta, tb)                                                                                          \
  {                                                                                                             \
    static bool done = false;                                                                                   \
    if (!done) {                                                                                                \
      ORX_CUDA(cudaFuncSetAttribute(k_gemm_tc<ta, tb>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      done = true;                                                                                              \
    }                                                                                                           \
    k_gemm_tc<ta, tb><<<grid, 128, smem, st>>>(A, lda, Bm, ldb, C, ldc, M, N, K, bias, act);                     \
  }
  if (TA == 0 && TB == 0) ORX_TC(0, 0)
  else if (TA == 0 && TB == 1) ORX_TC(0, 1)
  else if (TA == 1 && TB == 0) ORX_TC(1, 0)
  else ORX_TC(1, 1)
#undef ORX_TC
  ORX_LAUNCH_CHECK();
  return ORX_OK;
}
