// fp32 GEMM  C = beta*C + alpha * A * B^T  on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), with fp32-level
// accuracy through the 3xTF32 split:   a = a_hi + a_lo  (a_hi = the 19 bits the TF32 datapath keeps, a_lo = a - a_hi)
//     a b  ~=  a_hi b_hi + a_hi b_lo + a_lo b_hi        (relative error ~2^-21, the fp32 FFMA kernel's is 2^-24)
//
// One 128 x 128 or 128 x 256 output tile per CTA, fp32 accumulator in TMEM (128 lanes x 128 / 256 columns), 6 warps:
//   warp 0      TMA producer: 128 x 32 fp32 tiles of A and B per k-block (cp.async.bulk.tensor.3d, SWIZZLE_128B,
//               mbarrier complete_tx) into a 3-stage ring
//   warps 2-5   splitter: as soon as a stage lands they write the low parts a - trunc_tf32(a) of both tiles next to it
//               (element-wise, so the swizzled layout does not matter), fence.proxy.async, arrive on the stage's barrier;
//               after the main loop the same warps are the epilogue: tcgen05.ld the accumulator, C read-modify-write
//   warp 1      MMA issuer: one elected lane issues 12 x tcgen05.mma.kind::tf32 (M=128, N=128, K=8) per k-block
//               -- (A, B), (A, B_lo), (A_lo, B) for each of the four 32-byte K slices -- straight from shared-memory
//               descriptors, then tcgen05.commit frees the stage for the producer
// Used by the fp32 (batched) Cholesky for its trailing / panel updates (BASELINE config 3).  The fp64 default path
// cannot use this unit (no .kind::f64); see DESIGN.md section 8.
//
// SASS evidence: UTCHMMA (tcgen05.mma), UTMALDG (TMA), LDTM (tcgen05.ld), UTCBAR (tcgen05.commit).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace gpk {

constexpr int TC_BM = 128, TC_BK = 32;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;  // 16 KB per A tile
constexpr int TC_GEMM_THREADS = 192;
// BN = 128: stage = A, A_lo, B, B_lo = 64 KB, 3 stages.  BN = 256: stage = 16 + 16 + 32 + 32 = 96 KB, 2 stages -- the
// same bytes in flight, but every staged byte feeds twice the MMA work (the kernel is bound by bytes-in-flight / latency).
// ST = 1 (BN = 128): a single 64 KB stage per CTA, so that THREE CTAs share an SM (3 x 128 TMEM columns, 3 x 65 KB) -- the
// pipelining then happens ACROSS CTAs, including each tile's prologue and epilogue, which a one-CTA-per-SM ring cannot hide.
// For the many small K = 512 tiles of batched problems (config 3: tensor pipe 15 % busy with the 2-stage 128 x 256 kernel).
template <int BN, int ST = ((BN == 128) ? 3 : 2)>
struct TcCfg {
  static constexpr int B_BYTES = BN * TC_BK * 4;
  static constexpr int STAGE_BYTES = 2 * TC_A_BYTES + 2 * B_BYTES;
  static constexpr int STAGES = ST;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;
};

struct TcParams {
  float alpha, beta;
  void* C;  // float* or double* (CT of the kernel)
  int64_t ldc, c_bs;
  int32_t K, lower, tiles_m, tiles_n;
};

__device__ __forceinline__ void mbar_wait_parity(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "TCW_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra TCW_DONE;\n"
      "bra TCW_LOOP;\n"
      "TCW_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4, LBO = 1,
// SBO = 1024 B (8 rows x 128 B) >> 4, version 1 (Blackwell), layout type 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  const uint64_t hi = (uint64_t)(64u | (1u << 14) | (2u << 29)) << 32;
  return hi | (uint64_t)(((smem_addr >> 4) & 0x3FFFu) | (1u << 16));
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n"
      "}\n" ::"r"(tmem_c),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0));
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ float4 tf32_low_part(float4 v) {
  float4 lo;
  lo.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
  lo.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
  lo.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
  lo.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
  return lo;
}

// CT = type of C: float (fp32 problems) or double (opt-in mixed precision: fp64 matrix, fp32 operands -- the
// "tf32 where the user opts in" trailing update of the fp64 Cholesky).
template <typename CT, int TC_BN, int TC_ST = ((TC_BN == 128) ? 3 : 2)>
__global__ void __launch_bounds__(TC_GEMM_THREADS, 1)
gemm_nt_f32_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                      const TcParams p) {
  using Cfg = TcCfg<TC_BN, TC_ST>;
  constexpr int TC_STAGES = Cfg::STAGES, TC_STAGE_BYTES = Cfg::STAGE_BYTES, TC_TILE_BYTES = TC_A_BYTES;
  constexpr int TC_B_BYTES = Cfg::B_BYTES;
  int tm, tn;
  {
    const int GROUP = 8, per_group = GROUP * p.tiles_n, id = blockIdx.x;
    const int group = id / per_group, first_m = group * GROUP, gsize = min(p.tiles_m - first_m, GROUP);
    const int r = id - group * per_group;
    tm = first_m + r % gsize;
    tn = r / gsize;
  }
  if (p.lower && tn * TC_BN >= (tm + 1) * TC_BM) return;
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  extern __shared__ uint8_t tc_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full_bar[TC_STAGES], split_bar[TC_STAGES], empty_bar[TC_STAGES], accum_bar;
  __shared__ uint32_t tmem_base_holder;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_holder)),
                 "r"(TC_BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  if (threadIdx.x == 32) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&split_bar[s], 128);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&accum_bar, 1);
    fence_mbar_init();
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
  const uint32_t tmem = tmem_base_holder;
  const int KB = p.K / TC_BK;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % TC_STAGES, it = kb / TC_STAGES;
        if (it > 0) mbar_wait_parity(&empty_bar[s], (it - 1) & 1);
        uint8_t* st = smem + s * TC_STAGE_BYTES;
        mbar_arrive_expect_tx(&full_bar[s], TC_TILE_BYTES + TC_B_BYTES);
        tma_load_3d(st, &mapA, kb * TC_BK, tm * TC_BM, b, &full_bar[s]);
        tma_load_3d(st + 2 * TC_TILE_BYTES, &mapB, kb * TC_BK, tn * TC_BN, b, &full_bar[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor: D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10), K-major, N >> 3 at bit 17, M >> 4 at 24
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BN >> 3) << 17) |
                             ((uint32_t)(TC_BM >> 4) << 24);
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % TC_STAGES, it = kb / TC_STAGES;
        mbar_wait_parity(&split_bar[s], it & 1);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
        const uint32_t a_hi = smem_u32(smem + s * TC_STAGE_BYTES), a_lo = a_hi + TC_TILE_BYTES;
        const uint32_t b_hi = a_hi + 2 * TC_TILE_BYTES, b_lo = b_hi + TC_B_BYTES;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {  // UMMA K = 8 tf32 = 32 bytes inside the 128-byte swizzle row
          const uint32_t off = ks * 32;
          umma_tf32(tmem, umma_desc(a_hi + off), umma_desc(b_hi + off), idesc, (kb > 0 || ks > 0) ? 1u : 0u);
          umma_tf32(tmem, umma_desc(a_hi + off), umma_desc(b_lo + off), idesc, 1u);
          umma_tf32(tmem, umma_desc(a_lo + off), umma_desc(b_hi + off), idesc, 1u);
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(&accum_bar);
    }
  } else {
    // ---- splitter (main loop) ----
    const int t = threadIdx.x - 64;  // 0..127
    for (int kb = 0; kb < KB; ++kb) {
      const int s = kb % TC_STAGES, it = kb / TC_STAGES;
      mbar_wait_parity(&full_bar[s], it & 1);
      float4* hiA = reinterpret_cast<float4*>(smem + s * TC_STAGE_BYTES);
      float4* loA = hiA + TC_TILE_BYTES / 16;
      float4* hiB = hiA + 2 * TC_TILE_BYTES / 16;
      float4* loB = hiB + TC_B_BYTES / 16;
#pragma unroll 4
      for (int i = t; i < TC_TILE_BYTES / 16; i += 128) loA[i] = tf32_low_part(hiA[i]);
#pragma unroll 4
      for (int i = t; i < TC_B_BYTES / 16; i += 128) loB[i] = tf32_low_part(hiB[i]);
      fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
      mbar_arrive(&split_bar[s]);
    }
    // ---- epilogue: TMEM -> registers -> C (read-modify-write) ----
    mbar_wait_parity(&accum_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
    const int lane_group = warp & 3;  // a warp may only touch TMEM lanes 32 * (warp % 4) .. + 31
    const int row = lane_group * 32 + lane;
    CT* Crow = static_cast<CT*>(p.C) + (int64_t)b * p.c_bs + ((int64_t)tm * TC_BM + row) * p.ldc + (int64_t)tn * TC_BN;
    const float alpha = p.alpha, beta = p.beta;
#pragma unroll 1
    for (int c = 0; c < TC_BN / 32; ++c) {
      uint32_t r[32];
      const uint32_t taddr = tmem + ((uint32_t)(lane_group * 32) << 16) + c * 32;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
            "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
            "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::);
      if (sizeof(CT) == 4) {
        float4* cp = reinterpret_cast<float4*>(Crow + c * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 v;
          v.x = alpha * __uint_as_float(r[4 * j + 0]);
          v.y = alpha * __uint_as_float(r[4 * j + 1]);
          v.z = alpha * __uint_as_float(r[4 * j + 2]);
          v.w = alpha * __uint_as_float(r[4 * j + 3]);
          if (beta != 0.f) {
            const float4 o = cp[j];
            v.x = fmaf(beta, o.x, v.x);
            v.y = fmaf(beta, o.y, v.y);
            v.z = fmaf(beta, o.z, v.z);
            v.w = fmaf(beta, o.w, v.w);
          }
          cp[j] = v;
        }
      } else {
        double2* cp = reinterpret_cast<double2*>(Crow + c * 32);
        const double da = (double)alpha, db = (double)beta;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          double2 v;
          v.x = da * (double)__uint_as_float(r[2 * j + 0]);
          v.y = da * (double)__uint_as_float(r[2 * j + 1]);
          if (beta != 0.f) {
            const double2 o = cp[j];
            v.x = fma(db, o.x, v.x);
            v.y = fma(db, o.y, v.y);
          }
          cp[j] = v;
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(TC_BN));
}

// ---- host ------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) return (EncodeTiledFn) nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}

static bool make_map(CUtensorMap* m, const float* base, int64_t K, int64_t rows, int64_t ld, int64_t bs, int32_t batch,
                     int box_rows) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)((batch > 1) ? bs : rows * ld) * 4};
  cuuint32_t box[3] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows, 1};
  cuuint32_t es[3] = {1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <typename CT, int BN, int ST = ((BN == 128) ? 3 : 2)>
static int launch_tc_bn(int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t lda, int64_t a_bs,
                        const float* B, int64_t ldb, int64_t b_bs, float beta, CT* C, int64_t ldc, int64_t c_bs,
                        int32_t lower, int32_t batch, cudaStream_t stream) {
  CUtensorMap mA, mB;
  if (!make_map(&mA, A, K, M, lda, a_bs, batch, TC_BM) || !make_map(&mB, B, K, N, ldb, b_bs, batch, BN)) return 0;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_nt_f32_tc_kernel<CT, BN, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         TcCfg<BN, ST>::SMEM_BYTES);
    if (e != cudaSuccess) return -1000 - (int)e;
    attr_set = true;
  }
  TcParams p{alpha, beta, C, ldc, c_bs, (int32_t)K, lower, (int32_t)(M / TC_BM), (int32_t)(N / BN)};
  dim3 grid((unsigned)(p.tiles_m * p.tiles_n), (unsigned)batch);
  gemm_nt_f32_tc_kernel<CT, BN, ST><<<grid, TC_GEMM_THREADS, TcCfg<BN, ST>::SMEM_BYTES, stream>>>(mA, mB, p);
  GPK_COUNT_LAUNCH();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return -1000 - (int)e;
  return 1;
}

template <typename CT>
static int launch_tc(int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t lda, int64_t a_bs,
                     const float* B, int64_t ldb, int64_t b_bs, float beta, CT* C, int64_t ldc, int64_t c_bs, int32_t lower,
                     int32_t batch, cudaStream_t stream) {
  // 128 x 256 tiles when there are enough of them to fill the machine (each staged byte feeds twice the MMA work)
  static const int force_bn = getenv("GPK_TC_BN") ? atoi(getenv("GPK_TC_BN")) : 0;
  // batched problems (many small tiles): the one-stage 128 x 128 kernel, three CTAs per SM -- measured on config 3 (round 2):
  // 64 x 2048: 4.28 -> 4.09 ms, 512 x 2048: 30.8 -> 28.3 ms, bit-identical results.  GPK_TC_STAGES1: 0 = never, 1 = always,
  // 2 = batched (default), 3 = batched with 128 x 256 tiles / two CTAs per SM (28.7 ms)
  static const int one_stage = getenv("GPK_TC_STAGES1") ? atoi(getenv("GPK_TC_STAGES1")) : 2;
  if (one_stage == 3 && batch >= 16 && N % 256 == 0)  // 128 x 256 tiles, one 96 KB stage: two CTAs per SM
    return launch_tc_bn<CT, 256, 1>(M, N, K, alpha, A, lda, a_bs, B, ldb, b_bs, beta, C, ldc, c_bs, lower, batch, stream);
  if (one_stage == 1 || ((one_stage == 2 || one_stage == 3) && batch >= 16))
    return launch_tc_bn<CT, 128, 1>(M, N, K, alpha, A, lda, a_bs, B, ldb, b_bs, beta, C, ldc, c_bs, lower, batch, stream);
  const bool wide = (force_bn == 256) || (force_bn == 0 && N % 256 == 0 && (M / TC_BM) * (N / 256) * batch >= 148);
  if (wide && N % 256 == 0)
    return launch_tc_bn<CT, 256>(M, N, K, alpha, A, lda, a_bs, B, ldb, b_bs, beta, C, ldc, c_bs, lower, batch, stream);
  return launch_tc_bn<CT, 128>(M, N, K, alpha, A, lda, a_bs, B, ldb, b_bs, beta, C, ldc, c_bs, lower, batch, stream);
}

// returns 1 if the problem was launched on the tcgen05 path, 0 if the caller should use the FFMA kernel, < 0 on error
int gemm_nt_f32_tc(int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t lda, int64_t a_bs, const float* B,
                   int64_t ldb, int64_t b_bs, float beta, float* C, int64_t ldc, int64_t c_bs, int32_t lower,
                   int32_t batch, cudaStream_t stream) {
  static const bool disabled = getenv("GPK_F32_FFMA") != nullptr;
  if (disabled || K < 128 || K % TC_BK || M % TC_BM || N % 128) return 0;
  if (lda % 4 || ldb % 4 || ldc % 4 || (batch > 1 && (a_bs % 4 || b_bs % 4))) return 0;
  if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(C)) % 16) return 0;
  return launch_tc<float>(M, N, K, alpha, A, lda, a_bs, B, ldb, b_bs, beta, C, ldc, c_bs, lower, batch, stream);
}

// ---- opt-in mixed precision for the fp64 Cholesky: trailing update C(fp64) -= P P^T with P rounded to fp32 and the
// product formed by the 3xTF32 tensor-core kernel above (north_star: "tf32/bf16 where the user opts in") ----------------
__global__ void f64_to_f32_panel_kernel(const double* __restrict__ src, int64_t lds, float* __restrict__ dst, int64_t rows,
                                        int64_t K) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * K) return;
  const int64_t r = idx / K, k = idx - r * K;
  dst[idx] = (float)src[r * lds + k];
}

// C[M x N] (lower tiles) -= P[0:M] P[0:N]^T, P = fp64 panel (M x K, ld = ldp) converted into ws (M x K floats).
int syrk_f64_tf32x3(int64_t M, int64_t N, int64_t K, const float* ws_rows, double* C, int64_t ldc, cudaStream_t stream) {
  if (K % TC_BK || M % TC_BM || N % 128 || K < 128) return GPK_ERR_ARG;
  const int rc = launch_tc<double>(M, N, K, -1.0f, ws_rows, K, 0, ws_rows, K, 0, 1.0f, C, ldc, 0, 1, 1, stream);
  return rc == 1 ? 0 : (rc < 0 ? rc : GPK_ERR_UNSUPPORTED);
}

int convert_panel_f32(const double* P, int64_t ldp, int64_t rows, int64_t K, float* ws, cudaStream_t stream) {
  const int64_t total = rows * K;
  if (total == 0) return 0;
  f64_to_f32_panel_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(P, ldp, ws, rows, K);
  GPK_COUNT_LAUNCH();
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -1000 - (int)e;
}

}  // namespace gpk
