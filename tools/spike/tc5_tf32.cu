// Research spike (not product code): minimal tcgen05 + TMEM + TMA GEMM on sm_100a, to pin down descriptor encodings
// for the round-2 cluster-tile kernel.   D[128 x 128] = A[128 x K] * B[128 x K]^T, fp32 in (tf32 MMA), fp32 out.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                             float* D, int K) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float* As = reinterpret_cast<float*>(smem);            // 128 x 32 floats, SWIZZLE_128B
  float* Bs = reinterpret_cast<float*>(smem + 16384);
  __shared__ __align__(8) uint64_t full_bar, mma_bar;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_base)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full_bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&mma_bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::);
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::);
  const uint32_t tmem = tmem_base;

  // instruction descriptor: D = F32, A = B = TF32, K-major both, N = 128, M = 128
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
  const int KB = K / 32;
  if (threadIdx.x == 0) {
    uint32_t phase = 0;
    for (int kb = 0; kb < KB; ++kb) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full_bar)), "r"(32768));
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(s32(As)), "l"(&mapA), "r"(kb * 32), "r"(0), "r"(s32(&full_bar)) : "memory");
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(s32(Bs)), "l"(&mapB), "r"(kb * 32), "r"(0), "r"(s32(&full_bar)) : "memory");
      asm volatile("{\n.reg .pred P1;\nWAIT_A:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE_A;\nbra WAIT_A;\nDONE_A:\n}" ::"r"(s32(&full_bar)), "r"(phase));
      asm volatile("tcgen05.fence::after_thread_sync;" ::);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {  // UMMA_K = 8 tf32 = 32 bytes
        const uint64_t hi = (uint64_t)(64u | (1u << 14) | (2u << 29)) << 32;
        const uint64_t adesc = hi | (uint64_t)((((s32(As) + ks * 32) >> 4) & 0x3FFF) | (1u << 16));
        const uint64_t bdesc = hi | (uint64_t)((((s32(Bs) + ks * 32) >> 4) & 0x3FFF) | (1u << 16));
        const uint32_t accum = (kb > 0 || ks > 0) ? 1u : 0u;
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n}"
                     ::"r"(tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum), "r"(0), "r"(0), "r"(0), "r"(0));
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&mma_bar)) : "memory");
      asm volatile("{\n.reg .pred P1;\nWAIT_B:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE_B;\nbra WAIT_B;\nDONE_B:\n}" ::"r"(s32(&mma_bar)), "r"(phase));
      phase ^= 1;
    }
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::);
  // epilogue: warp w reads TMEM lanes 32w .. 32w+31 (= rows of D), 4 x 32 columns
  for (int c = 0; c < 4; ++c) {
    uint32_t r[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c * 32;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
                   "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
                   "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::);
    for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * 128 + c * 32 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::);
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int K = 256;
  std::vector<float> A(128 * K), B(128 * K), D(128 * 128);
  for (int i = 0; i < 128; ++i) for (int k = 0; k < K; ++k) { A[i * K + k] = (float)((i * 7 + k * 3) % 11 - 5) * 0.25f; B[i * K + k] = (float)((i * 5 + k * 2) % 13 - 6) * 0.125f; }
  float *dA, *dB, *dD;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0, D.size() * 4));
  void* fn = nullptr; cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  EncodeFn encode = (EncodeFn)fn;
  CUtensorMap mA, mB;
  cuuint64_t dims[2] = {(cuuint64_t)K, 128}; cuuint64_t strides[1] = {(cuuint64_t)K * 4};
  cuuint32_t box[2] = {32, 128}; cuuint32_t es[2] = {1, 1};
  CUresult r1 = encode(&mA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dA, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUresult r2 = encode(&mB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dB, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d %d\n", (int)r1, (int)r2);
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 + 1024));
  k<<<1, 128, 32768 + 1024>>>(mA, mB, dD, K);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0; int bad = 0;
  for (int i = 0; i < 128; ++i) for (int j = 0; j < 128; ++j) {
    double s = 0; for (int kk = 0; kk < K; ++kk) s += (double)A[i * K + kk] * B[j * K + kk];
    double e = fabs(s - D[i * 128 + j]); if (e > maxerr) maxerr = e; if (e > 1e-3) ++bad;
  }
  printf("tcgen05 tf32 GEMM 128x128x%d: max abs err %.3e, mismatches %d / 16384; D[0][0]=%f D[5][7]=%f D[127][127]=%f\n", K, maxerr, bad, D[0], D[5 * 128 + 7], D[127 * 128 + 127]);
  return 0;
}
