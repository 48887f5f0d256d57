// K8 (hand-written): fp32-accurate dense transform  D[M,N] = A[M,K] * Wt[N,K]^T (+ C) (+ bias) (ReLU)
// on the 5th-generation tensor cores, written directly against PTX (no CUTLASS):
//
//   * operands staged by TMA (cp.async.bulk.tensor.2d, 128-byte swizzle) into a multi-stage shared-memory ring,
//   * tcgen05.mma.kind::tf32 issued by ONE thread, fp32 accumulators in TMEM (2 x BN columns, double buffered),
//   * "3xTF32" split for fp32 accuracy without an extra HBM pass: transform warps rewrite the fp32 A tile in
//     shared memory as A_hi = round_tf32(A) (in place) and A_lo = round_tf32(A - A_hi) (second tile); the
//     (tiny) weight matrix is split once per call into Wt_hi / Wt_lo in global memory;
//     each k-step issues  A_lo*B_hi + A_hi*B_lo + A_hi*B_hi  (the dropped lo*lo term is ~2^-22),
//   * warp-specialised persistent CTAs (one per SM): warp 0 TMA producer, warp 1 MMA issuer + TMEM owner,
//     warps 2-5 transform, warps 6-9 epilogue (tcgen05.ld -> registers -> +C/+bias/ReLU -> global),
//     connected by mbarriers (full / lo-ready / empty per stage, tmem-full / tmem-empty per accumulator).
//
// M is the node dimension (millions), N <= 256 and K are feature widths: the A stream is read once from HBM,
// W stays in L2.  SASS: UTCHMMA-class UTCMMA (tf32), UTMALDG, LDTM.
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"

namespace kgb {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;  // 32 fp32 = 128 bytes = one swizzle row
constexpr int TC_THREADS = 320;   // tc_dw_kernel: TMA, MMA, 4 transform, 4 epilogue warps
#ifndef KGB_TC_NT
#define KGB_TC_NT 4
#endif
constexpr int TC_NT = KGB_TC_NT;                         // transform (tf32 split) warps of the forward kernels
constexpr int TC_GEMM_THREADS = 64 + 32 * TC_NT + 128;   // TMA, MMA, TC_NT transform, 4 epilogue warps

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// bulk tensor store / add-reduce of one {32 cols, 32 rows, 1} box from shared memory into a 3-D fp32 tensor
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_c),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// fp32 -> (hi, lo) with hi = a rounded to tf32 (10 explicit mantissa bits) and lo = tf32-rounded remainder; rounding
// (instead of the truncation the tensor core applies to raw fp32) keeps the split error unbiased.  inf/nan: lo = 0.
__device__ __forceinline__ void split_tf32(float a, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(a) + 0x1000u) & 0xFFFFE000u);
  float r = a - hi;
  r = __uint_as_float((__float_as_uint(r) + 0x1000u) & 0xFFFFE000u);
  lo = (fabsf(r) <= 3.0e38f) ? r : 0.f;
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// start address >> 4 [0,14), LBO [16,30) (unused for swizzled K-major, 1), SBO [32,46) = 1024 B between 8-row groups,
// version [46,48) = 1 (Blackwell), layout type [61,64) = 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ float4 ld_stream4(const float* p) {  // read-once data: do not allocate in L1
  float4 v;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

struct TcParams {
  int M, N, K;
  const float* C; int64_t ldc;   // optional addend
  const float* bias;             // optional [N]
  int relu;
  float* D; int64_t ldd;
  int n_tiles;
  int kblocks;  // k-blocks in total
  int kb1;      // k-blocks read through map_a; the rest comes from map_a2 (second A operand, K-concatenated)
};

// TMEM -> registers: 32 consecutive fp32 columns of this thread's accumulator row (asynchronous until tcgen05.wait::ld)
#define KGB_TMEM_LD32(r, addr)                                                                                       \
  asm volatile(                                                                                                      \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                      \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                      \
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                      \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),      \
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),     \
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                   \
      : "r"(addr))

// Epilogue of one 128 x BN accumulator, one 32-column chunk at a time: TMEM -> registers -> padded shared tile
// (transpose) -> (+C, +bias, ReLU) -> global.  Each thread holds 32 consecutive columns of ITS row; storing that
// directly would touch 32 different rows per instruction, so the chunk goes through a padded shared tile and every
// global access covers whole 128-byte row segments (4 rows x 128 B per warp instruction).
// (Software-pipelining the TMEM reads over two register sets was measured slower: 168 registers, 0.70 -> 0.83 ms.
//  Storing each thread's 128 contiguous bytes straight from registers, without the shared-memory transpose, was also
//  measured slower: 256->256 1.59 -> 1.89 ms, and 5.4 ms with an addend - its row-per-thread loads serialise.
//  Eight instead of four transform warps change nothing (1.59 -> 1.60 ms): the split is not throughput-limited, it
//  lengthens the TMA -> split -> MMA chain that three 64 KB stages have to cover.)
template <int BN, int EPI_LD>
__device__ __forceinline__ void tc_epilogue_tile(uint32_t taddr, float* stg, int lane, int64_t tile_row0, int M, int N,
                                                 const float* C, int64_t ldc, const float* bias, int relu, float* D,
                                                 int64_t ldd) {
  const int cv = (lane & 7) * 4;        // column (within the chunk) of this lane's float4
#pragma unroll 1
  for (int c0 = 0; c0 < BN; c0 += 32) {
    if (c0 >= N) break;
    const bool col_ok = (c0 + cv) < N;  // N is a multiple of 4
    // addend rows of this chunk: issued before the TMEM read so their DRAM latency hides behind it (loading them
    // next to the store would serialise 8 dependent round trips per chunk: D may alias C as far as the compiler knows)
    float4 cc[8];
    if (C) {
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int64_t grow = tile_row0 + it * 4 + (lane >> 3);
        cc[it] = (grow < M && col_ok) ? ld_stream4(C + grow * ldc + c0 + cv) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    uint32_t r[32];
    KGB_TMEM_LD32(r, taddr + c0);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int v = 0; v < 8; ++v)
      *reinterpret_cast<float4*>(stg + lane * EPI_LD + v * 4) =
          make_float4(__uint_as_float(r[4 * v]), __uint_as_float(r[4 * v + 1]), __uint_as_float(r[4 * v + 2]),
                      __uint_as_float(r[4 * v + 3]));
    __syncwarp();
    float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias && col_ok) bb = __ldg(reinterpret_cast<const float4*>(bias + c0 + cv));
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int rr = it * 4 + (lane >> 3);
      const int64_t grow = tile_row0 + rr;
      float4 o = *reinterpret_cast<const float4*>(stg + rr * EPI_LD + cv);
      if (grow < M && col_ok) {
        if (C) { o.x += cc[it].x; o.y += cc[it].y; o.z += cc[it].z; o.w += cc[it].w; }
        o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
        if (relu) {
          o.x = o.x <= 0.f ? 0.f : o.x; o.y = o.y <= 0.f ? 0.f : o.y;
          o.z = o.z <= 0.f ? 0.f : o.z; o.w = o.w <= 0.f ? 0.f : o.w;
        }
        *reinterpret_cast<float4*>(D + grow * ldd + c0 + cv) = o;
      }
    }
    __syncwarp();  // the staging tile is reused by the next column chunk
  }
}

template <int BN>
struct TcCfg {
  static constexpr int STAGES = BN == 256 ? 2 : (BN == 128 ? 3 : 4);
  static constexpr int A_BYTES = TC_BM * TC_BK * 4;  // 16 KB
  static constexpr int B_BYTES = BN * TC_BK * 4;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int EPI_LD = 36;                                   // padded row (floats) of the store staging tile
  static constexpr int EPI_BYTES = 4 * 32 * EPI_LD * 4;               // one 32x32 tile per epilogue warp
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
};

template <int BN>
__global__ void __launch_bounds__(TC_GEMM_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_a2,
               const __grid_constant__ CUtensorMap map_bhi, const __grid_constant__ CUtensorMap map_blo,
               const TcParams p) {
  using Cfg = TcCfg<BN>;
  constexpr int S = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* epi_stage = reinterpret_cast<float*>(smem + S * Cfg::STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * Cfg::STAGE_BYTES + Cfg::EPI_BYTES);
  uint64_t* full = bars;            // [S]  TMA bytes landed
  uint64_t* lo_rdy = bars + S;      // [S]  A_lo written
  uint64_t* empty = bars + 2 * S;   // [S]  MMAs that read the stage retired
  uint64_t* tfull = bars + 3 * S;   // [2]  accumulator complete
  uint64_t* tempty = tfull + 2;     // [2]  accumulator drained
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kblocks = p.kblocks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full + s, 1);
      mbar_init(lo_rdy + s, TC_NT);   // one arrival per transform warp
      mbar_init(empty + s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull + b, 1);
      mbar_init(tempty + b, 4);   // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM: 2 accumulators of BN fp32 columns (power of two >= 32)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)),
                 "r"(2 * BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  auto sA = [&](int s) { return smem + s * Cfg::STAGE_BYTES; };
  auto sAlo = [&](int s) { return smem + s * Cfg::STAGE_BYTES + Cfg::A_BYTES; };
  auto sBhi = [&](int s) { return smem + s * Cfg::STAGE_BYTES + 2 * Cfg::A_BYTES; };
  auto sBlo = [&](int s) { return smem + s * Cfg::STAGE_BYTES + 2 * Cfg::A_BYTES + Cfg::B_BYTES; };

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          mbar_wait(empty + s, ph ^ 1);
          mbar_expect_tx(full + s, Cfg::A_BYTES + 2 * Cfg::B_BYTES);
          if (kb < p.kb1) tma_load_2d(sA(s), &map_a, full + s, kb * TC_BK, tile * TC_BM);
          else tma_load_2d(sA(s), &map_a2, full + s, (kb - p.kb1) * TC_BK, tile * TC_BM);
          tma_load_2d(sBhi(s), &map_bhi, full + s, kb * TC_BK, 0);
          tma_load_2d(sBlo(s), &map_blo, full + s, kb * TC_BK, 0);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (single thread) =====
    if (lane == 0) {
      // instruction descriptor: D fp32 (1<<4), A/B tf32 (2<<7, 2<<10), both K-major, N>>3 at [17,23), M>>4 at [24,29)
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      uint32_t it = 0;
      int j = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++j) {
        const int b = j & 1;
        mbar_wait(tempty + b, ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_c = tmem_base + b * BN;
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          mbar_wait(full + s, ph);
          mbar_wait(lo_rdy + s, ph);
          tc_fence_after();
          const uint64_t da_hi = make_desc(smem_u32(sA(s))), da_lo = make_desc(smem_u32(sAlo(s)));
          const uint64_t db_hi = make_desc(smem_u32(sBhi(s))), db_lo = make_desc(smem_u32(sBlo(s)));
#pragma unroll
          for (int k4 = 0; k4 < TC_BK / 8; ++k4) {  // UMMA_K = 8 tf32 = 32 bytes: advance the start address
            const uint64_t adv = (uint64_t)(k4 * 2);
            umma_tf32(tmem_c, da_lo + adv, db_hi + adv, idesc, (kb | k4) != 0);
            umma_tf32(tmem_c, da_hi + adv, db_lo + adv, idesc, 1);
            umma_tf32(tmem_c, da_hi + adv, db_hi + adv, idesc, 1);
          }
          umma_commit(empty + s);  // frees the stage when these MMAs retire
        }
        umma_commit(tfull + b);    // accumulator ready for the epilogue
      }
    }
  } else if (warp < 2 + TC_NT) {
    // ===== transform warps: A -> (A_hi, A_lo) tf32 split (element-wise, so the swizzled layout is preserved) =====
    const int t = threadIdx.x - 64;  // 0 .. 32 * TC_NT - 1
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < kblocks; ++kb, ++it) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1;
        mbar_wait(full + s, ph);
        float4* src = reinterpret_cast<float4*>(sA(s));
        float4* dst = reinterpret_cast<float4*>(sAlo(s));
#pragma unroll
        for (int i = 0; i < Cfg::A_BYTES / 16 / (32 * TC_NT); ++i) {
          const float4 v = src[t + i * (32 * TC_NT)];
          float4 h, r;
          split_tf32(v.x, h.x, r.x);
          split_tf32(v.y, h.y, r.y);
          split_tf32(v.z, h.z, r.z);
          split_tf32(v.w, h.w, r.w);
          src[t + i * (32 * TC_NT)] = h;   // the TMA tile becomes the (rounded) high operand in place
          dst[t + i * (32 * TC_NT)] = r;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor-core reads
        __syncwarp();
        if (lane == 0) mbar_arrive(lo_rdy + s);
      }
    }
  } else {
    // ===== epilogue warps: TMEM -> registers -> (+C, +bias, ReLU) -> global =====
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    int j = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++j) {
      const int b = j & 1;
      mbar_wait(tfull + b, (j >> 1) & 1);
      tc_fence_after();
      const int64_t tile_row0 = (int64_t)tile * TC_BM + q * 32;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + b * BN;
      tc_epilogue_tile<BN, Cfg::EPI_LD>(taddr, epi_stage + q * 32 * Cfg::EPI_LD, lane, tile_row0, p.M, p.N, p.C, p.ldc,
                                        p.bias, p.relu, p.D, p.ldd);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty + b);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN));
  }
}

// --------------------------------------------------------------------------------------------------
// CTA-pair variant of the kernel above (tcgen05 cta_group::2): a cluster of two CTAs on one TPC owns a 256-row tile
// pair.  Each CTA stages ITS 128 rows of A (and splits them) plus ITS half of the weight rows; one thread of the
// leader CTA issues M=256 MMAs that read both CTAs' shared memory, each CTA's TMEM receives its 128 rows.  The weight
// tile is therefore fetched from L2 once per 256 rows instead of once per 128 (the 1-CTA kernel moves 5x more weight
// bytes than A bytes through the L2->SM path and is bound by it), and a stage shrinks to 64 KB (3 stages instead of 2).
//   full[s]   (local)   this CTA's TMA bytes landed                       -> its transform warps
//   ready[s]  (leader)  8 arrivals: 4 transform warps x 2 CTAs            -> MMA issuer
//   empty[s]  (local)   multicast tcgen05.commit: stage consumed          -> this CTA's TMA producer
//   tfull[b]  (local)   multicast tcgen05.commit: accumulator complete    -> this CTA's epilogue warps
//   tempty[b] (leader)  8 arrivals: 4 epilogue warps x 2 CTAs             -> MMA issuer
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(bar)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void umma_commit2(uint64_t* bar) {  // arrives on the same barrier offset in both CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void umma2_tf32(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_c),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

template <int BN>
struct Tc2Cfg {
  static constexpr int STAGES = BN == 256 ? 3 : 4;
  static constexpr int A_BYTES = TC_BM * TC_BK * 4;         // this CTA's 128 rows
  static constexpr int B_BYTES = (BN / 2) * TC_BK * 4;      // this CTA's half of the weight rows
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int EPI_LD = 36;
  static constexpr int EPI_BYTES = 4 * 32 * EPI_LD * 4;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 + 256;
};

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_GEMM_THREADS, 1)
tc_gemm2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_a2,
                const __grid_constant__ CUtensorMap map_bhi, const __grid_constant__ CUtensorMap map_blo,
                const TcParams p) {
  using Cfg = Tc2Cfg<BN>;
  constexpr int S = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* epi_stage = reinterpret_cast<float*>(smem + S * Cfg::STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * Cfg::STAGE_BYTES + Cfg::EPI_BYTES);
  uint64_t* full = bars;
  uint64_t* ready = bars + S;
  uint64_t* empty = bars + 2 * S;
  uint64_t* tfull = bars + 3 * S;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int kblocks = p.kblocks;
  const int n_pairs = (p.n_tiles + 1) / 2;
  const int pair0 = blockIdx.x >> 1, pair_stride = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full + s, 1);
      mbar_init(ready + s, 2 * TC_NT);
      mbar_init(empty + s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull + b, 1);
      mbar_init(tempty + b, 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)),
                 "r"(2 * BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
  }
  tc_fence_before();
  cluster_sync_all();   // barriers of both CTAs are initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  auto sA = [&](int s) { return smem + s * Cfg::STAGE_BYTES; };
  auto sAlo = [&](int s) { return smem + s * Cfg::STAGE_BYTES + Cfg::A_BYTES; };
  auto sBhi = [&](int s) { return smem + s * Cfg::STAGE_BYTES + 2 * Cfg::A_BYTES; };
  auto sBlo = [&](int s) { return smem + s * Cfg::STAGE_BYTES + 2 * Cfg::A_BYTES + Cfg::B_BYTES; };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int pair = pair0; pair < n_pairs; pair += pair_stride) {
        const int tile = pair * 2 + (int)rank;
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          mbar_wait(empty + s, ph ^ 1);
          mbar_expect_tx(full + s, Cfg::A_BYTES + 2 * Cfg::B_BYTES);
          if (kb < p.kb1) tma_load_2d(sA(s), &map_a, full + s, kb * TC_BK, tile * TC_BM);
          else tma_load_2d(sA(s), &map_a2, full + s, (kb - p.kb1) * TC_BK, tile * TC_BM);
          tma_load_2d(sBhi(s), &map_bhi, full + s, kb * TC_BK, (int)rank * (BN / 2));
          tma_load_2d(sBlo(s), &map_blo, full + s, kb * TC_BK, (int)rank * (BN / 2));
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // D fp32, A/B tf32, K-major, N = BN (both halves), M = 256 (128 rows per CTA)
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      uint32_t it = 0;
      int j = 0;
      for (int pair = pair0; pair < n_pairs; pair += pair_stride, ++j) {
        const int b = j & 1;
        mbar_wait(tempty + b, ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_c = tmem_base + b * BN;
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          mbar_wait(ready + s, ph);
          tc_fence_after();
          const uint64_t da_hi = make_desc(smem_u32(sA(s))), da_lo = make_desc(smem_u32(sAlo(s)));
          const uint64_t db_hi = make_desc(smem_u32(sBhi(s))), db_lo = make_desc(smem_u32(sBlo(s)));
#pragma unroll
          for (int k4 = 0; k4 < TC_BK / 8; ++k4) {
            const uint64_t adv = (uint64_t)(k4 * 2);
            umma2_tf32(tmem_c, da_lo + adv, db_hi + adv, idesc, (kb | k4) != 0);
            umma2_tf32(tmem_c, da_hi + adv, db_lo + adv, idesc, 1);
            umma2_tf32(tmem_c, da_hi + adv, db_hi + adv, idesc, 1);
          }
          umma_commit2(empty + s);
        }
        umma_commit2(tfull + b);
      }
    }
  } else if (warp < 2 + TC_NT) {
    const int t = threadIdx.x - 64;
    uint32_t it = 0;
    for (int pair = pair0; pair < n_pairs; pair += pair_stride) {
      for (int kb = 0; kb < kblocks; ++kb, ++it) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1;
        mbar_wait(full + s, ph);
        float4* src = reinterpret_cast<float4*>(sA(s));
        float4* dst = reinterpret_cast<float4*>(sAlo(s));
#pragma unroll
        for (int i = 0; i < Cfg::A_BYTES / 16 / (32 * TC_NT); ++i) {
          const float4 v = src[t + i * (32 * TC_NT)];
          float4 h, r;
          split_tf32(v.x, h.x, r.x);
          split_tf32(v.y, h.y, r.y);
          split_tf32(v.z, h.z, r.z);
          split_tf32(v.w, h.w, r.w);
          src[t + i * (32 * TC_NT)] = h;
          dst[t + i * (32 * TC_NT)] = r;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(ready + s, 0);
      }
    }
  } else {
    const int q = warp & 3;
    int j = 0;
    for (int pair = pair0; pair < n_pairs; pair += pair_stride, ++j) {
      const int b = j & 1;
      mbar_wait(tfull + b, (j >> 1) & 1);
      tc_fence_after();
      const int64_t tile_row0 = (int64_t)(pair * 2 + (int)rank) * TC_BM + q * 32;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + b * BN;
      tc_epilogue_tile<BN, Cfg::EPI_LD>(taddr, epi_stage + q * 32 * Cfg::EPI_LD, lane, tile_row0, p.M, p.N, p.C, p.ldc,
                                        p.bias, p.relu, p.D, p.ldd);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty + b, 0);
    }
  }

  tc_fence_before();
  cluster_sync_all();   // no CTA leaves (or frees TMEM) while its peer can still signal it or read its shared memory
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN));
  }
}

// --------------------------------------------------------------------------------------------------
// dW[Kx,N] = X[M,Kx]^T * G[M,N]: the reduction runs over the M nodes, so BOTH operands are MN-major in memory
// (their feature index is contiguous).  TMA boxes of {32 features x 16 nodes} with the 128B/32B-atom swizzle land as
// canonical MN-major SWIZZLE_128B_BASE32B atoms (4 node-rows x 128 B); feature groups are LBO = 2 KB apart, 4-node
// groups SBO = 512 B apart, one k-step (8 nodes) advances the start address by 1 KB.  Both tiles are split into tf32 hi/lo in shared memory (both operands are large here).
// Every CTA reduces a contiguous slice of nodes into its own [Kx,N] partial (TMEM: up to 2 x 256 columns);
// the partials are added in CTA order by kgb_reduce_parts (deterministic split-K, no atomics).
// The tensor core adds into its fp32 accumulator with truncation, which biases long chains (1e-4 relative after
// 16 k nodes), so the TMEM chain is cut every `chain_blocks` k-blocks (256 nodes): the epilogue warps add the chain's
// result to the CTA's partial with IEEE fp32 adds (the partial is L2-resident) and the next chain starts from zero.
constexpr int DW_R = 16;  // nodes per pipeline stage (2 k-steps of 8)

struct DwParams {
  int M, Kx, N;
  int nodes_per_cta;
  int chain_blocks;  // k-blocks (of DW_R nodes) accumulated in TMEM before the sum is promoted to fp32 in memory
  int g2_box0;       // first 32-column box of the G tile that is loaded from the SECOND gradient operand (>= BN/32: none)
  int x2_box0;       // first 32-feature box of the X tile that is loaded from the SECOND feature operand (>= MT*4: none)
  float* partial;    // [gridDim.x, Kx, N]
};

template <int BN, int MT>
struct DwCfg {
  static constexpr int X_BYTES = MT * 128 * DW_R * 4;   // MT * 8 KB
  static constexpr int G_BYTES = BN * DW_R * 4;         // BN * 64 B
  static constexpr int HALF = X_BYTES + G_BYTES;        // raw (hi) tiles; the lo tiles follow
  static constexpr int STAGE_BYTES = 2 * HALF;
  static constexpr int STAGES = (192 * 1024) / STAGE_BYTES > 6 ? 6 : (192 * 1024) / STAGE_BYTES;
  static constexpr int EPI_TILE = 32 * 32 * 4;            // one 32 x 32 fp32 tile of the drain (128-byte swizzled rows)
  static constexpr int EPI_BYTES = 4 * 2 * EPI_TILE;      // two tiles per epilogue warp: one fills while one drains
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 + 256;
};

// MN-major descriptor for 32-bit operands: the only legal swizzle is SWIZZLE_128B_BASE32B (layout type 1): atoms of
// 4 K-rows x 128 B with 32-byte chunks XOR-ed by the row index (TMA: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B).
// LBO = bytes between 32-element MN atoms, SBO = bytes between 4-row K groups.
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (1ull << 61);
}

template <int BN, int MT>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_dw_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_x2,
             const __grid_constant__ CUtensorMap map_g, const __grid_constant__ CUtensorMap map_g2,
             const __grid_constant__ CUtensorMap map_p, const DwParams p) {
  using Cfg = DwCfg<BN, MT>;
  constexpr int S = Cfg::STAGES;
  constexpr int BOX = 32 * DW_R * 4;  // 2 KB per TMA box
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* stage_all = reinterpret_cast<float*>(smem + S * Cfg::STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * Cfg::STAGE_BYTES + Cfg::EPI_BYTES);
  uint64_t* full = bars;
  uint64_t* lo_rdy = bars + S;
  uint64_t* empty = bars + 2 * S;
  uint64_t* tfull = bars + 3 * S;
  uint64_t* tempty = tfull + 1;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tempty + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t node0 = (int64_t)blockIdx.x * p.nodes_per_cta;
  int64_t node1 = node0 + p.nodes_per_cta;
  if (node1 > p.M) node1 = p.M;
  const int kblocks = node1 > node0 ? (int)((node1 - node0 + DW_R - 1) / DW_R) : 0;
  constexpr int TMEM_COLS = MT * BN < 32 ? 32 : MT * BN;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full + s, 1);
      mbar_init(lo_rdy + s, 4);
      mbar_init(empty + s, 1);
    }
    mbar_init(tfull, 1);
    mbar_init(tempty, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // The chain boundaries are staggered across the CTAs (the first chain of CTA b is shorter by chain_off k-blocks):
  // with identical boundaries all 148 CTAs drain 256 KB each into the L2 in the same few microseconds and the drain
  // runs at the L2's aggregate fp32-add rate (38 MB per burst); spread over the chain period each CTA's drain only
  // sees its own share.  The offsets are a function of blockIdx: the result stays deterministic.
  const int chain_off = p.chain_blocks > 1 ? (int)((blockIdx.x * 5u) % (unsigned)p.chain_blocks) : 0;
  const int n_chains = kblocks > 0 ? (kblocks + chain_off + p.chain_blocks - 1) / p.chain_blocks : 0;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)),
                 "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < kblocks; ++kb) {
        const int s = kb % S;
        const uint32_t ph = (kb / S) & 1;
        mbar_wait(empty + s, ph ^ 1);
        mbar_expect_tx(full + s, Cfg::HALF);
        uint8_t* base = smem + s * Cfg::STAGE_BYTES;
        const int row = (int)(node0 + (int64_t)kb * DW_R);
#pragma unroll
        for (int b = 0; b < MT * 4; ++b) {
          // two feature operands that share G ([X1 | X2]^T G, e.g. SAGE's dW_neigh = agg^T g and dW_self = x^T g when
          // both are at most 128 wide): their feature boxes sit one after the other in ONE X tile, G is loaded and
          // split once
          if (b < p.x2_box0) tma_load_2d(base + b * BOX, &map_x, full + s, b * 32, row);
          else tma_load_2d(base + b * BOX, &map_x2, full + s, (b - p.x2_box0) * 32, row);
        }
#pragma unroll
        for (int b = 0; b < BN / 32; ++b) {
          // two gradient operands that share X (dW_a = X^T G1, dW_b = X^T G2): their column boxes sit side by side
          // in ONE G tile, so X is loaded and split once for both products
          if (b < p.g2_box0) tma_load_2d(base + Cfg::X_BYTES + b * BOX, &map_g, full + s, b * 32, row);
          else tma_load_2d(base + Cfg::X_BYTES + b * BOX, &map_g2, full + s, (b - p.g2_box0) * 32, row);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // D fp32, A/B tf32, both MN-major (bits 15, 16)
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      for (int kb = 0; kb < kblocks; ++kb) {
        const int s = kb % S;
        const uint32_t ph = (kb / S) & 1;
        const int chain = (kb + chain_off) / p.chain_blocks, kc = (kb + chain_off) % p.chain_blocks;
        const bool first_kb = (kb == 0) || (kc == 0);
        if (first_kb) {  // new accumulation chain: the epilogue must have drained the previous one
          mbar_wait(tempty, (chain & 1) ^ 1);
          tc_fence_after();
        }
        mbar_wait(full + s, ph);
        mbar_wait(lo_rdy + s, ph);
        tc_fence_after();
        const uint32_t hi = smem_u32(smem + s * Cfg::STAGE_BYTES), lo = hi + Cfg::HALF;
#pragma unroll
        for (int ks = 0; ks < DW_R / 8; ++ks) {
          const uint64_t gb_hi = make_desc_mn(hi + Cfg::X_BYTES + ks * 1024, BOX, 512);
          const uint64_t gb_lo = make_desc_mn(lo + Cfg::X_BYTES + ks * 1024, BOX, 512);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const uint64_t xa_hi = make_desc_mn(hi + mt * 4 * BOX + ks * 1024, BOX, 512);
            const uint64_t xa_lo = make_desc_mn(lo + mt * 4 * BOX + ks * 1024, BOX, 512);
            const uint32_t tc = tmem_base + mt * BN;
            umma_tf32(tc, xa_lo, gb_hi, idesc, !(first_kb && ks == 0));
            umma_tf32(tc, xa_hi, gb_lo, idesc, 1);
            umma_tf32(tc, xa_hi, gb_hi, idesc, 1);
          }
        }
        umma_commit(empty + s);
        if (kc == p.chain_blocks - 1 || kb == kblocks - 1) umma_commit(tfull);  // chain complete
      }
    }
  } else if (warp < 6) {
    const int t = threadIdx.x - 64;
    for (int kb = 0; kb < kblocks; ++kb) {
      const int s = kb % S;
      const uint32_t ph = (kb / S) & 1;
      mbar_wait(full + s, ph);
      float4* src = reinterpret_cast<float4*>(smem + s * Cfg::STAGE_BYTES);
      float4* dst = reinterpret_cast<float4*>(smem + s * Cfg::STAGE_BYTES + Cfg::HALF);
#pragma unroll 4
      for (int i = t; i < Cfg::HALF / 16; i += 128) {
        const float4 v = src[i];
        float4 h, r;
        split_tf32(v.x, h.x, r.x);
        split_tf32(v.y, h.y, r.y);
        split_tf32(v.z, h.z, r.z);
        split_tf32(v.w, h.w, r.w);
        src[i] = h;
        dst[i] = r;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(lo_rdy + s);
    }
  } else {
    // epilogue: after every chain, partial (+)= accumulator with IEEE fp32 adds; a CTA without nodes writes zeros.
    // TMEM is full at 256 x 256, so the MMAs of the next chain wait for this drain.  Draining with per-thread
    // red.global.add.v4 cost 10.7 us per chain (16 k half-filled sector transactions through the LSU; 34 % of the
    // kernel), through a shared transpose with coalesced reds 7.7 us.  Here the epilogue warps only move TMEM ->
    // registers -> a 128-byte-swizzled shared tile and hand each 32 x 32 tile to the TMA engine as ONE bulk tensor
    // add-reduce (cp.reduce.async.bulk.tensor, performed at the L2: IEEE fp32 adds, one writer per address); the
    // accumulator is released as soon as it has been read, the bulk operations finish behind the next chain's MMAs.
    // Order of the adds = chain order: every drain first waits for the complete retirement of the previous one.
    const int q = warp & 3;
    uint8_t* stg_base = reinterpret_cast<uint8_t*>(stage_all) + q * 2 * Cfg::EPI_TILE;
    const int passes = n_chains > 0 ? n_chains : 1;
    uint32_t n_issued = 0;   // tiles this warp has handed to the TMA so far (selects the staging tile)
#pragma unroll 1
    for (int chain = 0; chain < passes; ++chain) {
      if (n_chains > 0) {
        mbar_wait(tfull, chain & 1);
        tc_fence_after();
      }
      if (chain > 0) {   // the previous chain's stores / adds to the same addresses are complete and visible
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        __syncwarp();
      }
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
        const int row0 = mt * 128 + q * 32;  // first feature row (kx) of this warp's TMEM lanes
        if (row0 >= p.Kx) break;             // rows past Kx do not exist in the partial (their TMEM lanes are padding)
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          if (c0 >= p.N) break;
          uint32_t r[32];
          if (n_chains > 0) {
            KGB_TMEM_LD32(r, tmem_base + ((uint32_t)(q * 32) << 16) + mt * BN + c0);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = 0u;
          }
          float* stg = reinterpret_cast<float*>(stg_base + (n_issued & 1u) * Cfg::EPI_TILE);
          if (n_issued >= 2) {   // the bulk operation that last read this tile has finished reading it
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
          }
          // row `lane` of the tile, 16-byte chunk v at position v ^ (lane & 7): the TMA's 128-byte swizzle (and
          // conflict-free for the eight lanes of a quarter warp)
#pragma unroll
          for (int v = 0; v < 8; ++v)
            *reinterpret_cast<float4*>(stg + lane * 32 + ((v ^ (lane & 7)) << 2)) =
                make_float4(__uint_as_float(r[4 * v]), __uint_as_float(r[4 * v + 1]), __uint_as_float(r[4 * v + 2]),
                            __uint_as_float(r[4 * v + 3]));
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> bulk-copy reads
          __syncwarp();
          if (lane == 0) {
            // columns past N and rows past Kx of the box are clipped by the tensor map
            if (chain > 0) tma_reduce_add_3d(&map_p, stg, c0, row0, (int)blockIdx.x);
            else tma_store_3d(&map_p, stg, c0, row0, (int)blockIdx.x);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          ++n_issued;
        }
      }
      if (n_chains > 0) {   // the accumulator has been read out: the next chain may start
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty);
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // nothing in flight at kernel exit
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

// W [rows, cols] (row-major, ld) -> hi/lo tf32 parts, optionally transposed: out is [cols, rows] when transpose
__global__ void split_tf32_kernel(const float* __restrict__ w, int rows, int cols, int64_t ld, int transpose,
                                  float* __restrict__ hi, float* __restrict__ lo, int64_t ldo) {
  const int64_t n = (int64_t)rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i % cols);
    const float v = w[(int64_t)r * ld + c];
    float h, l;
    split_tf32(v, h, l);
    const int64_t o = transpose ? (int64_t)c * ldo + r : (int64_t)r * ldo + c;
    hi[o] = h;
    lo[o] = l;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D fp32 tensor [rows, cols] row-major with leading dimension ld, box = {32 cols, box_rows rows}, 128B swizzle
static int make_map(CUtensorMap* m, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                    CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return KGB_ERR_CUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld", (int)r, (long long)rows,
              (long long)cols, (long long)ld);
    return KGB_ERR_CUDA;
  }
  return KGB_OK;
}

template <int BN>
static int tc_launch(int device, const CUtensorMap& ma, const CUtensorMap& ma2, const CUtensorMap& mh,
                     const CUtensorMap& ml, const TcParams& p, cudaStream_t st) {
  using Cfg = TcCfg<BN>;
  static bool attr_done[64] = {};
  if (device < 64 && !attr_done[device]) {
    KGB_CHECK_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_done[device] = true;
  }
  int grid = sm_count(device);
  if (grid > p.n_tiles) grid = p.n_tiles;
  tc_gemm_kernel<BN><<<grid, TC_GEMM_THREADS, Cfg::SMEM_BYTES, st>>>(ma, ma2, mh, ml, p);
  KGB_CHECK_LAUNCH();
  return KGB_OK;
}


template <int BN>
static int tc2_launch(int device, const CUtensorMap& ma, const CUtensorMap& ma2, const CUtensorMap& mh,
                      const CUtensorMap& ml, const TcParams& p, cudaStream_t st) {
  using Cfg = Tc2Cfg<BN>;
  static int clusters[64] = {};
  if (device < 64 && clusters[device] == 0) {
    KGB_CHECK_CUDA(cudaFuncSetAttribute(tc_gemm2_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(sm_count(device) / 2 * 2);
    cfg.blockDim = dim3(TC_GEMM_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    int n = 0;
    KGB_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&n, tc_gemm2_kernel<BN>, &cfg));  // co-resident CTA pairs
    clusters[device] = n > 0 ? n : -1;
  }
  const int resident = device < 64 ? clusters[device] : sm_count(device) / 2;
  if (resident <= 0) return KGB_ERR_UNSUPPORTED;
  int pairs = (p.n_tiles + 1) / 2;
  if (pairs > resident) pairs = resident;
  tc_gemm2_kernel<BN><<<2 * pairs, TC_GEMM_THREADS, Cfg::SMEM_BYTES, st>>>(ma, ma2, mh, ml, p);
  KGB_CHECK_LAUNCH();
  return KGB_OK;
}

// the per-CTA partials [n_parts, Kx, N] as a 3-D fp32 tensor, box = {32 cols, 32 rows, 1}, 128B swizzle (the drain's
// bulk stores / add-reduces; columns past N and rows past Kx of a box are clipped)
static int make_map_partials(CUtensorMap* m, float* base, int64_t n_parts, int64_t Kx, int64_t N) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return KGB_ERR_CUDA;
  }
  cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)Kx, (cuuint64_t)n_parts};
  cuuint64_t strides[2] = {(cuuint64_t)N * sizeof(float), (cuuint64_t)Kx * (cuuint64_t)N * sizeof(float)};
  cuuint32_t box[3] = {32, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (partials) failed (%d) parts=%lld Kx=%lld N=%lld", (int)r, (long long)n_parts,
              (long long)Kx, (long long)N);
    return KGB_ERR_CUDA;
  }
  return KGB_OK;
}

template <int BN, int MT>
static int dw_launch(int device, const CUtensorMap& mx, const CUtensorMap& mx2, const CUtensorMap& mg, const CUtensorMap& mg2,
                     const CUtensorMap& mp, const DwParams& p, int grid,
                     cudaStream_t st) {
  using Cfg = DwCfg<BN, MT>;
  static bool attr_done[64] = {};
  if (device < 64 && !attr_done[device]) {
    KGB_CHECK_CUDA(cudaFuncSetAttribute(tc_dw_kernel<BN, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_done[device] = true;
  }
  tc_dw_kernel<BN, MT><<<grid, TC_THREADS, Cfg::SMEM_BYTES, st>>>(mx, mx2, mg, mg2, mp, p);
  KGB_CHECK_LAUNCH();
  return KGB_OK;
}

}  // namespace kgb

using namespace kgb;

extern "C" {

int32_t kgb_linear_tc_dw_parts(int device, int64_t M) {
  if (kgb::use_device(device) != KGB_OK) return -1;
  int64_t g = sm_count(device);
  const int64_t blocks = (M + DW_R - 1) / DW_R;
  if (g > blocks) g = blocks;
  return (int32_t)(g < 1 ? 1 : g);
}

// dW = X^T [G1 | G2]: G2 (optional) starts at column N1 rounded up to a whole 32-column box of the virtual G tile
static int dw_impl(int device, const float* X, int64_t ldx, const float* G, int64_t ldg, int32_t N1, const float* G2,
                   int64_t ldg2, int32_t N2, int32_t M, int32_t Kx1, float* partials, int32_t n_parts, cudaStream_t st,
                   const float* X2 = nullptr, int64_t ldx2 = 0, int32_t Kx2 = 0) {
  const int n1_boxes = (N1 + 31) / 32;
  const int N = G2 ? n1_boxes * 32 + N2 : N1;   // columns of the partial (the gap columns are exact zeros)
  const int k1_boxes = (Kx1 + 31) / 32;
  const int Kx = X2 ? k1_boxes * 32 + Kx2 : Kx1;   // rows of the partial (the gap rows are exact zeros)
  CUtensorMap mx, mx2, mg, mg2, mp;
  int rc = make_map(&mx, X, M, Kx1, ldx, DW_R, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (rc != KGB_OK) return rc;
  if (X2) rc = make_map(&mx2, X2, M, Kx2, ldx2, DW_R, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  else mx2 = mx;
  if (rc != KGB_OK) return rc;
  rc = make_map(&mg, G, M, N1, ldg, DW_R, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (rc != KGB_OK) return rc;
  if (G2) rc = make_map(&mg2, G2, M, N2, ldg2, DW_R, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  else mg2 = mg;
  if (rc != KGB_OK) return rc;
  rc = make_map_partials(&mp, partials, n_parts, Kx, N);
  if (rc != KGB_OK) return rc;
  DwParams p;
  p.M = M; p.Kx = Kx; p.N = N; p.partial = partials;
  const int64_t blocks = (M + DW_R - 1) / DW_R;
  p.nodes_per_cta = (int)(((blocks + n_parts - 1) / n_parts) * DW_R);
  p.chain_blocks = 256 / DW_R;  // promote the TMEM sum to fp32 memory every 256 nodes
  {
    // timing experiments only (tools/exp_tc.py): longer chains lose fp32 parity, see the kernel's header comment
    static const int chain_nodes = [] { const char* e = getenv("KGB200_DW_CHAIN"); return e ? atoi(e) : 0; }();
    if (chain_nodes >= DW_R) p.chain_blocks = chain_nodes / DW_R;
  }
  const int BN = N <= 64 ? 64 : (N <= 128 ? 128 : 256);
  p.g2_box0 = G2 ? n1_boxes : BN / 32;
  const int MT = Kx <= 128 ? 1 : 2;
  p.x2_box0 = X2 ? k1_boxes : MT * 4;
#define KGB_DW_CASE(BN_, MT_) if (BN == BN_ && MT == MT_) return dw_launch<BN_, MT_>(device, mx, mx2, mg, mg2, mp, p, n_parts, st);
  KGB_DW_CASE(64, 1) KGB_DW_CASE(64, 2) KGB_DW_CASE(128, 1) KGB_DW_CASE(128, 2) KGB_DW_CASE(256, 1) KGB_DW_CASE(256, 2)
#undef KGB_DW_CASE
  return KGB_ERR_UNSUPPORTED;
}

int kgb_linear_tc_dw(int device, const float* X, int64_t ldx, const float* G, int64_t ldg, int32_t M, int32_t Kx,
                     int32_t N, float* partials, int32_t n_parts, kgb_stream_t stream) {
  KGB_USE_DEVICE(device);
  KGB_REQUIRE(M >= 1 && Kx > 0 && N > 0, "kgb_linear_tc_dw needs M >= 1 and positive widths");  // rows past M: TMA zero fill
  KGB_REQUIRE(X && G && partials, "NULL operand");
  KGB_REQUIRE(Kx <= 256 && N <= 256 && Kx % 4 == 0 && N % 4 == 0, "needs Kx, N <= 256 and multiples of 4");
  KGB_REQUIRE(aligned16(X) && aligned16(G) && aligned16(partials) && ldx % 4 == 0 && ldg % 4 == 0, "alignment");
  KGB_REQUIRE(n_parts == kgb_linear_tc_dw_parts(device, M), "n_parts must come from kgb_linear_tc_dw_parts");
  return dw_impl(device, X, ldx, G, ldg, N, nullptr, 0, 0, M, Kx, partials, n_parts, (cudaStream_t)stream);
}

int32_t kgb_linear_tc_dw2_cols(int32_t N1, int32_t N2) { return (N1 + 31) / 32 * 32 + N2; }

int kgb_linear_tc_dw_x2(int device, const float* X1, int64_t ldx1, int32_t Kx1, const float* X2, int64_t ldx2, int32_t Kx2,
                        const float* G, int64_t ldg, int32_t N, int32_t M, float* partials, int32_t n_parts,
                        kgb_stream_t stream) {
  KGB_USE_DEVICE(device);
  KGB_REQUIRE(M >= 1 && Kx1 > 0 && Kx2 > 0 && N > 0, "kgb_linear_tc_dw_x2 needs M >= 1 and positive widths");
  KGB_REQUIRE(X1 && X2 && G && partials, "NULL operand");
  KGB_REQUIRE(N <= 256 && N % 4 == 0 && Kx1 % 4 == 0 && Kx2 % 4 == 0 && kgb_linear_tc_dw2_cols(Kx1, Kx2) <= 256,
              "needs N <= 256, widths multiples of 4 and ceil32(Kx1) + Kx2 <= 256");
  KGB_REQUIRE(aligned16(X1) && aligned16(X2) && aligned16(G) && aligned16(partials) && ldx1 % 4 == 0 && ldx2 % 4 == 0 &&
                  ldg % 4 == 0,
              "alignment");
  KGB_REQUIRE(n_parts == kgb_linear_tc_dw_parts(device, M), "n_parts must come from kgb_linear_tc_dw_parts");
  return dw_impl(device, X1, ldx1, G, ldg, N, nullptr, 0, 0, M, Kx1, partials, n_parts, (cudaStream_t)stream, X2, ldx2,
                 Kx2);
}

int kgb_linear_tc_dw2(int device, const float* X, int64_t ldx, const float* G1, int64_t ldg1, int32_t N1, const float* G2,
                      int64_t ldg2, int32_t N2, int32_t M, int32_t Kx, float* partials, int32_t n_parts,
                      kgb_stream_t stream) {
  KGB_USE_DEVICE(device);
  KGB_REQUIRE(M >= 1 && Kx > 0 && N1 > 0 && N2 > 0, "kgb_linear_tc_dw2 needs M >= 1 and positive widths");
  KGB_REQUIRE(X && G1 && G2 && partials, "NULL operand");
  KGB_REQUIRE(Kx <= 256 && Kx % 4 == 0 && N1 % 4 == 0 && N2 % 4 == 0 && kgb_linear_tc_dw2_cols(N1, N2) <= 256,
              "needs Kx <= 256, widths multiples of 4 and ceil32(N1) + N2 <= 256");
  KGB_REQUIRE(aligned16(X) && aligned16(G1) && aligned16(G2) && aligned16(partials) && ldx % 4 == 0 && ldg1 % 4 == 0 &&
                  ldg2 % 4 == 0,
              "alignment");
  KGB_REQUIRE(n_parts == kgb_linear_tc_dw_parts(device, M), "n_parts must come from kgb_linear_tc_dw_parts");
  return dw_impl(device, X, ldx, G1, ldg1, N1, G2, ldg2, N2, M, Kx, partials, n_parts, (cudaStream_t)stream);
}

int32_t kgb_linear_tc_rows(int32_t N) { return N <= 64 ? 64 : (N <= 128 ? 128 : 256); }

int kgb_split_tf32_ld(int device, const float* w, int32_t rows, int32_t cols, int64_t ld, int32_t transpose, float* hi,
                      float* lo, int64_t ldo, kgb_stream_t stream) {
  KGB_USE_DEVICE(device);
  KGB_REQUIRE(w && hi && lo && rows > 0 && cols > 0 && ld >= cols, "bad arguments");
  KGB_REQUIRE(ldo >= (transpose ? rows : cols), "output leading dimension too small");
  const int64_t n = (int64_t)rows * cols;
  int grid = (int)((n + 255) / 256);
  if (grid > 1184) grid = 1184;
  split_tf32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(w, rows, cols, ld, transpose, hi, lo, ldo);
  KGB_CHECK_LAUNCH();
  return KGB_OK;
}

int kgb_split_tf32(int device, const float* w, int32_t rows, int32_t cols, int64_t ld, int32_t transpose, float* hi,
                   float* lo, kgb_stream_t stream) {
  return kgb_split_tf32_ld(device, w, rows, cols, ld, transpose, hi, lo, transpose ? rows : cols, stream);
}

// row length of the split weights: [W1 ; W2] with W1's block padded to whole k-blocks, and a multiple of 4 floats
// overall (TMA rows are 16-byte multiples); the A operands themselves may have any width (TMA zero-fills past K)
int32_t kgb_linear_tc2_k(int32_t K1, int32_t K2) {
  return K2 > 0 ? (K1 + TC_BK - 1) / TC_BK * TC_BK + (K2 + 3) / 4 * 4 : (K1 + 3) / 4 * 4;
}

int kgb_linear_tc2(int device, const float* A1, int64_t lda1, int32_t K1, const float* A2, int64_t lda2, int32_t K2,
                   int32_t M, const float* wt_hi, const float* wt_lo, int32_t N, const float* C, int64_t ldc,
                   const float* bias, int32_t act, float* D, int64_t ldd, kgb_stream_t stream) {
  KGB_USE_DEVICE(device);
  KGB_REQUIRE(M >= 0 && N > 0 && K1 > 0 && K2 >= 0, "bad sizes");
  if (M == 0) return KGB_OK;
  KGB_REQUIRE(A1 && wt_hi && wt_lo && D && (K2 == 0 || A2), "NULL operand");
  KGB_REQUIRE(N <= 256 && N % 4 == 0, "kgb_linear_tc needs N <= 256 and a multiple of 4 (pad the output)");
  KGB_REQUIRE(aligned16(A1) && (!A2 || aligned16(A2)) && aligned16(wt_hi) && aligned16(wt_lo) && aligned16(D) &&
                  (!C || aligned16(C)) && (!bias || aligned16(bias)) && lda1 % 4 == 0 && (!A2 || lda2 % 4 == 0) &&
                  ldd % 4 == 0 && (!C || ldc % 4 == 0),
              "operands must be 16-byte aligned with leading dimensions multiple of 4");
  const int BN = N <= 64 ? 64 : (N <= 128 ? 128 : 256);
  const int kb1 = (K1 + TC_BK - 1) / TC_BK;
  const int Kcat = kgb_linear_tc2_k(K1, K2);          // the split weights are [BN, Kcat]
  CUtensorMap ma, ma2, mh, ml;
  int rc = make_map(&ma, A1, M, K1, lda1, TC_BM);
  if (rc != KGB_OK) return rc;
  if (K2 > 0) {
    rc = make_map(&ma2, A2, M, K2, lda2, TC_BM);
    if (rc != KGB_OK) return rc;
  } else {
    ma2 = ma;
  }
  // CTA pairs (cta_group::2) once there are enough 256-row tile pairs to fill the machine twice over
  static const int pair_env = [] { const char* e = getenv("KGB_TC_PAIR"); return e ? atoi(e) : 1; }();
  // (only the 256-wide tile gains: 1.63 -> 1.58 ms at 2.45 M x 256 x 256; narrower tiles are slower paired)
  const bool use_pair = pair_env != 0 && BN == 256 && M >= 4 * TC_BM * sm_count(device);
  rc = make_map(&mh, wt_hi, BN, Kcat, Kcat, use_pair ? BN / 2 : BN);  // zero-padded to BN rows (kgb_linear_tc_rows)
  if (rc != KGB_OK) return rc;
  rc = make_map(&ml, wt_lo, BN, Kcat, Kcat, use_pair ? BN / 2 : BN);
  if (rc != KGB_OK) return rc;
  TcParams p;
  p.M = M; p.N = N; p.K = Kcat; p.C = C; p.ldc = ldc; p.bias = bias; p.relu = (act == KGB_ACT_RELU);
  p.D = D; p.ldd = ldd; p.n_tiles = (M + TC_BM - 1) / TC_BM;
  p.kb1 = kb1;
  p.kblocks = kb1 + (K2 + TC_BK - 1) / TC_BK;
  cudaStream_t st = (cudaStream_t)stream;
  if (use_pair) return tc2_launch<256>(device, ma, ma2, mh, ml, p, st);
  if (BN == 64) return tc_launch<64>(device, ma, ma2, mh, ml, p, st);
  if (BN == 128) return tc_launch<128>(device, ma, ma2, mh, ml, p, st);
  return tc_launch<256>(device, ma, ma2, mh, ml, p, st);
}

int kgb_linear_tc(int device, const float* A, int64_t lda, int32_t M, int32_t K, const float* wt_hi, const float* wt_lo,
                  int32_t N, const float* C, int64_t ldc, const float* bias, int32_t act, float* D, int64_t ldd,
                  kgb_stream_t stream) {
  return kgb_linear_tc2(device, A, lda, K, nullptr, 0, 0, M, wt_hi, wt_lo, N, C, ldc, bias, act, D, ldd, stream);
}

}  // extern "C"
