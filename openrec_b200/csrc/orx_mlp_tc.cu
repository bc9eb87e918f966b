// orx_mlp_tc.cu -- Dense-layer GEMMs on the 5th-gen tensor cores: TMA-fed tcgen05.mma kind::tf32, accumulators in TMEM.
//
// DLRM's MLPs (openrec/tf2/modules/multi_layer_perceptron.py:5-18, recommenders/dlrm.py:34-37,87,90-95) are the one
// dense contraction on the path.  The parity bar is 1e-5 against an fp32 reference, which plain TF32 (10-bit mantissa)
// cannot meet, so every fp32 operand is split into two TF32 terms (hi = TF32(v) rounded to nearest, lo = v - hi)
// and  C += Ahi*Bhi + Ahi*Blo + Alo*Bhi  is accumulated in fp32 (3xTF32, relative error ~2^-21 per product).
//
// One CTA computes a 128 x TN tile (TN = 256, or 128 for narrow outputs) of  C[M,N] = op(A)[M,K] * op(B)[K,N], K in
// blocks of 16, with three kinds of warps and six rings of mbarriers between them:
//   warp 0 (one lane)  TMA producer: cp.async.bulk.tensor.2d loads the RAW fp32 A and B blocks of a k-block straight from
//                      the operands' own layouts (row stride = ld; out-of-range rows / k are zero-filled by the TMA) into
//                      a 3-stage raw ring.  K-contiguous sources arrive as [rows][16] with the 64-byte swizzle,
//                      M/N-contiguous ones ([k][rows], e.g. w in the forward pass, x and dz in dw = x^T dz) as [16][rows].
//   warps 2..          converters: read a raw stage (conflict-free: swizzled 128-bit loads, or four coalesced scalars for
//                      the transposing case), split into hi / lo and store both tiles in the canonical K-major,
//                      no-swizzle UMMA layout (8-row x 16-byte core matrices, LBO 128 B, SBO 512 B) of a 3-stage operand
//                      ring -- conflict-free 128-bit stores in both cases.
//   warp 1 (one lane)  MMA issuer: per k-block 2 k-steps x 3 tcgen05.mma (M = 128, N = TN, K = 8), tcgen05.commit frees the
//                      operand stage.
// The tensor core's fp32 accumulation truncates, and over hundreds of accumulations that bias grows linearly with K
// (measured 3e-4 abs at K = 1024), so a TMEM accumulator only ever sums 4 k-blocks (64 K-elements, 24 MMAs); the
// converter warps then drain it with tcgen05.ld into fp32 REGISTERS (round-to-nearest adds; a warp owns 32 rows x 64
// columns) while the MMAs of the next group fill the other TMEM accumulator.
// Epilogue through shared memory (coalesced rows, + bias, activation); split-K (blockIdx.z) for dw = x^T dz, whose K is
// the batch.  Shapes TMA cannot describe (row stride not a multiple of 16 bytes) or that are too small for a tile are
// left to the fp32 SIMT kernel of orx_dlrm.cu (ORX_ERR_UNSUPPORTED).
#include <cuda.h>
#include <stdlib.h>

#include "orx_common.cuh"

namespace {

constexpr int TM = 128, TK = 16;
constexpr int RAW_STAGES = 3, OP_STAGES = 3, GROUP_KB = 4;
constexpr int OP_SBO = (TK / 4) * 128;           // bytes between 8-row groups of an operand tile

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp: SmemDescriptor)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);                  // start address        bits [0,14)
  d |= (uint64_t)(128 >> 4) << 16;                          // leading byte offset  bits [16,30): K-adjacent core matrices
  d |= (uint64_t)(OP_SBO >> 4) << 32;                       // stride byte offset   bits [32,46): next 8-row group
  d |= (uint64_t)1 << 46;                                   // descriptor version (sm_100)
  return d;                                                 // base offset 0, layout type SWIZZLE_NONE (bits 61-63 = 0)
}

// kind::tf32, D = F32, A/B = TF32, both K-major, M = 128, N = TN (InstrDescriptor bit layout)
template <int TN>
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
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
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// one 2-D TMA load: box at (c0 = inner coordinate, c1 = outer coordinate) -> shared memory, completion on `bar`
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(smem_dst)),
      "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}

// hi = TF32(v) rounded to nearest (cvt.rna: truncation would bias every product the same way and the error would grow
// linearly with K instead of with sqrt(K)); lo = v - hi exactly (|lo| <= 2^-11 |v|).  lo is handed to the tensor core as
// it is: kind::tf32 reads the top 19 bits, i.e. truncates lo by at most 2^-10 |lo| <= 2^-21 |v| -- the size of the
// lo*lo term 3xTF32 drops anyway, and of random sign because hi was rounded to nearest.
__device__ __forceinline__ float to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}
__device__ __forceinline__ void split4(const float4 v, float4* hi, float4* lo) {
  hi->x = to_tf32(v.x); lo->x = v.x - hi->x;
  hi->y = to_tf32(v.y); lo->y = v.y - hi->y;
  hi->z = to_tf32(v.z); lo->z = v.z - hi->z;
  hi->w = to_tf32(v.w); lo->w = v.w - hi->w;
}

// One warp-item of the conversion: 32 (row, 4-k group) pairs of a raw tile -> hi / lo operand tiles.
//  KC = true : raw is [R rows][16 k] with the TMA 64-byte swizzle (16-byte chunk c of row r sits at chunk c ^ ((r >> 1) & 3));
//              item `wi` covers rows 8*wi .. 8*wi+7: lane -> row 8*wi + (lane & 7), k-core lane >> 3.  A quarter-warp reads
//              eight rows of one k-core (8 distinct bank groups thanks to the swizzle) and writes 128 contiguous bytes.
//  KC = false: raw is [16 k][R rows]; item `wi` covers k-core wi & 3 of rows 32*(wi >> 2) .. +31: four coalesced scalar
//              loads per lane, one 128-bit store per tile (a quarter-warp again writes 128 contiguous bytes).
// The byte offsets of an item inside a raw stage / an operand tile do not depend on the k-block: computed once per thread.
template <bool KC>
__device__ __forceinline__ void item_offsets(int R, int wi, int lane, int* src_off, int* dst_off) {
  int row, kcore;
  if (KC) {
    row = 8 * wi + (lane & 7);
    kcore = lane >> 3;
    *src_off = row * 64 + ((kcore ^ ((row >> 1) & 3)) << 4);
  } else {
    row = 32 * (wi >> 2) + lane;
    kcore = wi & 3;
    *src_off = (kcore * 4 * R + row) * 4;
  }
  *dst_off = (row >> 3) * OP_SBO + kcore * 128 + (row & 7) * 16;
}
template <bool KC, int R>
__device__ __forceinline__ void convert_item(const unsigned char* raw, int src_off, int dst_off, unsigned char* hi_tile,
                                             unsigned char* lo_tile) {
  float4 v;
  if (KC) {
    v = *reinterpret_cast<const float4*>(raw + src_off);
  } else {
    const float* p = reinterpret_cast<const float*>(raw + src_off);
    v.x = p[0];
    v.y = p[R];
    v.z = p[2 * R];
    v.w = p[3 * R];
  }
  float4 h, l;
  split4(v, &h, &l);
  *reinterpret_cast<float4*>(hi_tile + dst_off) = h;
  *reinterpret_cast<float4*>(lo_tile + dst_off) = l;
}

template <int TN>
struct Cfg {
  static constexpr int NCW = 4 * (TN / 64);              // converter / drain / epilogue warps: 32 rows x 64 columns each
  static constexpr int THREADS = (2 + NCW) * 32;
  static constexpr int RAW_A = TM * TK * 4, RAW_B = TN * TK * 4, RAW_STAGE = RAW_A + RAW_B;
  static constexpr int OP_A = TM * TK * 4, OP_B = TN * TK * 4, OP_STAGE = 2 * OP_A + 2 * OP_B;
  static constexpr int SMEM = RAW_STAGES * RAW_STAGE + OP_STAGES * OP_STAGE;   // 216 KB (TN = 256) / 144 KB (TN = 128)
  static constexpr int TMEM_COLS = 2 * TN;
  static_assert(TM * (TN + 1) * 4 <= SMEM, "the epilogue tile must fit in the rings");
};

// TA / TB as in orx_dlrm.cu: TA = 0: A[m*lda + k]; TA = 1: A[k*lda + m]; TB = 0: B[k*ldb + n]; TB = 1: B[n*ldb + k].
template <int TA, int TB, int TN>
__global__ void __launch_bounds__(Cfg<TN>::THREADS, 1)
k_gemm_tma(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* __restrict__ C,
           int64_t ldc, int M, int N, int K, const float* __restrict__ bias, int act, float* __restrict__ part) {
  using G = Cfg<TN>;
  constexpr bool A_KC = TA == 0, B_KC = TB == 1;   // K-contiguous sources
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t raw_full[RAW_STAGES], raw_empty[RAW_STAGES], op_full[OP_STAGES], op_empty[OP_STAGES],
      acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  unsigned char* raw_ring = smem;
  unsigned char* op_ring = smem + RAW_STAGES * G::RAW_STAGE;

  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "n"(G::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < RAW_STAGES; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], G::NCW); }
    for (int s = 0; s < OP_STAGES; ++s) { mbar_init(&op_full[s], G::NCW); mbar_init(&op_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], G::NCW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_s;

  // this CTA's range of k-blocks (split-K over blockIdx.z)
  const int nkb_all = (K + TK - 1) / TK;
  const int per = (nkb_all + (int)gridDim.z - 1) / (int)gridDim.z;
  const int kb_lo = blockIdx.z * per;
  const int nkb = max(0, min(nkb_all, kb_lo + per) - kb_lo);
  const int ngroups = (nkb + GROUP_KB - 1) / GROUP_KB;

  float acc[64];
#pragma unroll
  for (int j = 0; j < 64; ++j) acc[j] = 0.f;

  if (warp == 0) {
    // ------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int rs = kb % RAW_STAGES;
        if (kb >= RAW_STAGES) mbar_wait(&raw_empty[rs], ((kb / RAW_STAGES) - 1) & 1);
        unsigned char* st = raw_ring + (size_t)rs * G::RAW_STAGE;
        const int k0 = (kb_lo + kb) * TK;
        mbar_expect_tx(&raw_full[rs], G::RAW_STAGE);          // a box is always delivered whole (out of range = zeros)
        if (A_KC) tma_load_2d(st, &tmA, k0, m0, &raw_full[rs]); else tma_load_2d(st, &tmA, m0, k0, &raw_full[rs]);
        if (B_KC) tma_load_2d(st + G::RAW_A, &tmB, k0, n0, &raw_full[rs]); else tma_load_2d(st + G::RAW_A, &tmB, n0, k0, &raw_full[rs]);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint64_t desc0 = make_desc(smem_u32(op_ring));
      for (int kb = 0; kb < nkb; ++kb) {
        const int os = kb % OP_STAGES, g = kb / GROUP_KB;
        const bool g_first = (kb % GROUP_KB) == 0, g_last = (kb % GROUP_KB) == GROUP_KB - 1 || kb == nkb - 1;
        if (g_first && g >= 2) mbar_wait(&acc_empty[g & 1], ((g >> 1) - 1) & 1);   // group g-2 has been drained from this buffer
        mbar_wait(&op_full[os], (kb / OP_STAGES) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // descriptors: the address field is the low 14 bits (16-byte units), everything else is constant
        const uint64_t a_hi = desc0 + (uint64_t)(os * (G::OP_STAGE >> 4)), a_lo = a_hi + (G::OP_A >> 4),
                       b_hi = a_lo + (G::OP_A >> 4), b_lo = b_hi + (G::OP_B >> 4);
        const uint32_t d = tmem_d + (uint32_t)((g & 1) * TN);
#pragma unroll
        for (int ks = 0; ks < TK / 8; ++ks) {
          const uint64_t o = (uint64_t)(ks * (256 >> 4));     // two k-cores of 128 bytes
          mma_tf32<TN>(d, a_hi + o, b_hi + o, (g_first && ks == 0) ? 0u : 1u);
          mma_tf32<TN>(d, a_hi + o, b_lo + o, 1u);
          mma_tf32<TN>(d, a_lo + o, b_hi + o, 1u);
        }
        umma_commit(&op_empty[os]);
        if (g_last) umma_commit(&acc_full[g & 1]);
      }
    }
  } else {
    // ------------------------------------------------ converters + accumulator drain
    const int cw = warp - 2;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32);   // the TMEM lanes a warp may touch: 32 * (warp % 4)
    const int cchunk = (cw >> 2) * 64;                         // its 64 accumulator columns
    auto drain = [&](int g) {
      const int buf = g & 1;
      mbar_wait(&acc_full[buf], (g >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem_d + (lane_base << 16) + (uint32_t)(buf * TN + cchunk + c0);
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
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    };
    constexpr int ITEMS_A = TM / 8, ITEMS_B = TN / 8;          // warp-items per tile (32 (row, k-core) pairs each)
    constexpr int NIT = (ITEMS_A + ITEMS_B) / G::NCW;          // items per warp and k-block (3 or 4); item i of warp cw is
    static_assert((ITEMS_A + ITEMS_B) % G::NCW == 0 && ITEMS_A % G::NCW == 0, "items must divide evenly");   // cw + NCW*i
    constexpr int NIT_A = ITEMS_A / G::NCW;                    // the first NIT_A items of a warp belong to A, the rest to B
    int src_off[NIT], dst_off[NIT];
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
      const int wi = cw + G::NCW * i;
      if (i < NIT_A) item_offsets<A_KC>(TM, wi, lane, &src_off[i], &dst_off[i]);
      else item_offsets<B_KC>(TN, wi - ITEMS_A, lane, &src_off[i], &dst_off[i]);
    }
    int drained = 0;
    for (int kb = 0; kb < nkb; ++kb) {
      const int rs = kb % RAW_STAGES, os = kb % OP_STAGES, g = kb / GROUP_KB;
      mbar_wait(&raw_full[rs], (kb / RAW_STAGES) & 1);
      if (kb >= OP_STAGES) mbar_wait(&op_empty[os], ((kb / OP_STAGES) - 1) & 1);   // the MMAs of block kb-3 left this stage
      const unsigned char* ra = raw_ring + (size_t)rs * G::RAW_STAGE;
      const unsigned char* rb = ra + G::RAW_A;
      unsigned char* oa = op_ring + (size_t)os * G::OP_STAGE;
      unsigned char* ob = oa + 2 * G::OP_A;
#pragma unroll
      for (int i = 0; i < NIT; ++i) {
        if (i < NIT_A) convert_item<A_KC, TM>(ra, src_off[i], dst_off[i], oa, oa + G::OP_A);
        else convert_item<B_KC, TN>(rb, src_off[i], dst_off[i], ob, ob + G::OP_B);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&op_full[os]);
        mbar_arrive(&raw_empty[rs]);
      }
      // drain group g-1 one k-block AFTER group g has started: by then its last MMAs have retired, so the converters do
      // not sit in the accumulator wait while the tensor core runs out of converted operands
      if ((kb % GROUP_KB) == 1 && g >= 1) { drain(g - 1); drained = g; }
    }
    for (int d = drained; d < ngroups; ++d) drain(d);
  }

  // ---- epilogue through shared memory: every MMA has completed (the last drain waited for the last group, and groups
  // complete in order) and every TMA load has been consumed, so the rings are free: tile[128][TN + 1] floats
  __syncthreads();
  float* tile = reinterpret_cast<float*>(smem);
  if (warp >= 2) {
    const int cw = warp - 2;
    const int r = (warp & 3) * 32 + lane, c0 = (cw >> 2) * 64;
#pragma unroll
    for (int j = 0; j < 64; ++j) tile[r * (TN + 1) + c0 + j] = acc[j];
  }
  __syncthreads();
  const bool split = gridDim.z > 1;
  float* out = split ? part + (size_t)blockIdx.z * (size_t)M * (size_t)N : C;
  const int64_t ldo = split ? (int64_t)N : ldc;
  for (int r = warp; r < TM; r += G::THREADS / 32) {
    const int m = m0 + r;
    if (m >= M) break;
#pragma unroll
    for (int c = 0; c < TN; c += 32) {
      const int n = n0 + c + lane;
      if (n < N) {
        float x = tile[r * (TN + 1) + c + lane];
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
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(G::TMEM_COLS) : "memory");
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

// ---- host: TMA descriptors (cuTensorMapEncodeTiled through the runtime's driver entry point: no -lcuda needed)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// map of one operand: `kc` = its K dimension is the contiguous one ([rows, K] with row stride ld), else [K, rows]
int make_map(CUtensorMap* tm, const float* base, int rows, int K, int64_t ld, bool kc, int tile_rows) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) { orx_set_error("cuTensorMapEncodeTiled is not available in this driver"); return ORX_ERR_CUDA; }
  cuuint64_t dims[2], strides[1];
  cuuint32_t box[2], estr[2] = {1, 1};
  if (kc) { dims[0] = (cuuint64_t)K; dims[1] = (cuuint64_t)rows; box[0] = TK; box[1] = (cuuint32_t)tile_rows; }
  else { dims[0] = (cuuint64_t)rows; dims[1] = (cuuint64_t)K; box[0] = (cuuint32_t)tile_rows; box[1] = TK; }
  strides[0] = (cuuint64_t)ld * 4;
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, kc ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { orx_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return ORX_ERR_CUDA; }
  return ORX_OK;
}

template <int TA, int TB, int TN>
int launch_tma(const CUtensorMap& ta, const CUtensorMap& tb, float* C, int64_t ldc, int M, int N, int K, const float* bias,
               int act, float* part, int S, cudaStream_t st) {
  using G = Cfg<TN>;
  static bool done = false;
  if (!done) {
    ORX_CUDA(cudaFuncSetAttribute(k_gemm_tma<TA, TB, TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM + 1024));
    done = true;
  }
  dim3 grid((N + TN - 1) / TN, (M + TM - 1) / TM, S);
  k_gemm_tma<TA, TB, TN><<<grid, G::THREADS, G::SMEM + 1024, st>>>(ta, tb, C, ldc, M, N, K, bias, act, part);
  ORX_LAUNCH_CHECK();
  return ORX_OK;
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
// Returns ORX_ERR_UNSUPPORTED for shapes that are left to the SIMT kernel: tiny N / K / M, or operands the TMA cannot
// describe (base not 16-byte aligned, row stride not a multiple of 16 bytes).
int orx_launch_gemm_tc(int TA, int TB, const float* A, int64_t lda, const float* Bm, int64_t ldb, float* C, int64_t ldc,
                       int M, int N, int K, const float* bias, int act, cudaStream_t st) {
  if (N < 16 || K < 8 || M < 64) return ORX_ERR_UNSUPPORTED;
  if ((lda & 3) || (ldb & 3) || (((uintptr_t)A | (uintptr_t)Bm) & 15)) return ORX_ERR_UNSUPPORTED;
  const int TN = N > 128 ? 256 : 128;
  const int tiles = ((N + TN - 1) / TN) * ((M + TM - 1) / TM);
  const int nkb = (K + TK - 1) / TK;
  int S = 1;
  if (tiles < 148 && nkb >= 32) {             // too few tiles for the machine and a long K: split it
    S = (2 * 148 + tiles - 1) / tiles;
    if (S > nkb / 8) S = nkb / 8;             // at least 8 k-blocks (two accumulation groups) per split
    if (S < 1) S = 1;
  }
  float* part = nullptr;
  if (S > 1) {
    part = orx_splitk_workspace((size_t)S * (size_t)M * (size_t)N);
    if (!part) { orx_set_error("split-K workspace allocation failed"); return ORX_ERR_CUDA; }
  }
  CUtensorMap ta, tb;
  int rc;
  if ((rc = make_map(&ta, A, M, K, lda, TA == 0, TM))) return rc;
  if ((rc = make_map(&tb, Bm, N, K, ldb, TB == 1, TN))) return rc;
#define ORX_TMA(a, b)                                                                                     \
  rc = TN == 256 ? launch_tma<a, b, 256>(ta, tb, C, ldc, M, N, K, bias, act, part, S, st)                 \
                 : launch_tma<a, b, 128>(ta, tb, C, ldc, M, N, K, bias, act, part, S, st)
  if (TA == 0 && TB == 0) ORX_TMA(0, 0);
  else if (TA == 0 && TB == 1) ORX_TMA(0, 1);
  else if (TA == 1 && TB == 0) ORX_TMA(1, 0);
  else ORX_TMA(1, 1);
#undef ORX_TMA
  if (rc) return rc;
  if (S > 1) return orx_launch_splitk_reduce(part, S, M, N, C, ldc, bias, act, st);
  return ORX_OK;
}
