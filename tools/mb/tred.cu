// Does the TMA tensor reduce (cp.reduce.async.bulk.tensor .add) accept FLOAT64 tensor maps on sm_100a?
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap m) {
  __shared__ __align__(1024) double buf[128 * 16];
  for (int i = threadIdx.x; i < 128 * 16; i += blockDim.x) buf[i] = 1.0 + i;
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned s = (unsigned)__cvta_generic_to_shared(buf);
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2}], [%3];\n" ::"l"(&m),
                 "r"(16), "r"(128), "r"(s)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
  }
}
int main() {
  const int R = 512, Cc = 256;
  double* d;
  cudaMalloc(&d, sizeof(double) * R * Cc);
  double* h = new double[R * Cc];
  for (int i = 0; i < R * Cc; ++i) h[i] = 1000.0;
  cudaMemcpy(d, h, sizeof(double) * R * Cc, cudaMemcpyHostToDevice);
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)Cc, (cuuint64_t)R};
  cuuint64_t strides[1] = {(cuuint64_t)Cc * 8};
  cuuint32_t box[2] = {16, 128}, es[2] = {1, 1};
  CUresult r = ((EncFn)p)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d\n", (int)r);
  k<<<1, 128>>>(m);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  cudaMemcpy(h, d, sizeof(double) * R * Cc, cudaMemcpyDeviceToHost);
  printf("C[128][16]=%.1f (expect 1001) C[129][17]=%.1f (expect 1018) C[255][31]=%.1f (expect %.1f) C[0][0]=%.1f\n",
         h[128 * Cc + 16], h[129 * Cc + 17], h[255 * Cc + 31], 1000.0 + 1 + 127 * 16 + 15, h[0]);
  return 0;
}
