// DMMA pipe experiments: what separates the 37 TF/s register-only peak from the 30.5 TF/s GEMM main loop?
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// MODE 0: regs only, 8x4 accumulators, operands fixed.  1: + LDS.128 fragment loads each k8.  2: + __syncthreads per 2 k8
// 3: + cp.async 16B x 8 per k-tile (to smem scratch from global)
template <int MODE, int WM, int WN>
__global__ void __launch_bounds__(512, 1) k(double* out, const double* gsrc, int iters) {
  extern __shared__ __align__(16) double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double acc[WM][WN][2];
  for (int i = 0; i < WM; i++) for (int j = 0; j < WN; j++) { acc[i][j][0] = lane; acc[i][j][1] = i + j; }
  for (int i = threadIdx.x; i < 16384; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i;
  __syncthreads();
  double2 a[WM], b[WN];
  for (int i = 0; i < WM; i++) a[i] = make_double2(1.0 + i * 1e-9, 1.0 - i * 1e-9);
  for (int j = 0; j < WN; j++) b[j] = make_double2(1.0 + j * 1e-9, 1.0 - j * 1e-9);
  for (int it = 0; it < iters; it++) {
    if (MODE >= 2) __syncthreads();
    if (MODE >= 3) {
      for (int q = 0; q < 8; q++) {
        unsigned dst = (unsigned)__cvta_generic_to_shared(sm + 8192 + ((it & 3) * 2048 + q * 256 + (threadIdx.x & 255)) * 2 % 8192);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(gsrc + ((size_t)blockIdx.x * 4096 + (it & 63) * 64 + q * 512 + (threadIdx.x & 255) * 2) % (1 << 20)));
      }
      asm volatile("cp.async.commit_group;\n" ::);
      asm volatile("cp.async.wait_group 2;\n" ::);
    }
#pragma unroll
    for (int k8 = 0; k8 < 2; k8++) {
      if (MODE >= 1) {
        const double* base = sm + ((it * 2 + k8) & 3) * 2048 + (warp & 3) * 512 + lane * 2;
#pragma unroll
        for (int i = 0; i < WM; i++) a[i] = *reinterpret_cast<const double2*>(base + i * 64);
#pragma unroll
        for (int j = 0; j < WN; j++) b[j] = *reinterpret_cast<const double2*>(base + 1024 + j * 64);
      }
#pragma unroll
      for (int i = 0; i < WM; i++)
#pragma unroll
        for (int j = 0; j < WN; j++) dmma(acc[i][j][0], acc[i][j][1], a[i].x, b[j].x);
#pragma unroll
      for (int i = 0; i < WM; i++)
#pragma unroll
        for (int j = 0; j < WN; j++) dmma(acc[i][j][0], acc[i][j][1], a[i].y, b[j].y);
    }
  }
  double s = 0;
  for (int i = 0; i < WM; i++) for (int j = 0; j < WN; j++) s += acc[i][j][0] + acc[i][j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE, int WM, int WN>
void run(const char* name, int threads, double* out, const double* g) {
  int iters = 4000;
  auto kern = k<MODE, WM, WN>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  kern<<<148, threads, 131072>>>(out, g, iters); cudaDeviceSynchronize();
  cudaEventRecord(e0); kern<<<148, threads, 131072>>>(out, g, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fl = 512.0 * WM * WN * 4 * iters * (threads / 32) * 148;
  printf("%-40s threads=%3d  %.2f TF/s  (%s)\n", name, threads, fl / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  double *out, *g; cudaMalloc(&out, 148 * 512 * 8); cudaMalloc(&g, (1 << 20) * 8 + 65536); cudaMemset(g, 0, (1 << 20) * 8);
  run<0, 8, 4>("regs only 8x4", 256, out, g);
  run<1, 8, 4>("+LDS 8x4", 256, out, g);
  run<2, 8, 4>("+LDS+sync 8x4", 256, out, g);
  run<3, 8, 4>("+LDS+sync+cpasync 8x4", 256, out, g);
  run<0, 4, 4>("regs only 4x4", 512, out, g);
  run<1, 4, 4>("+LDS 4x4", 512, out, g);
  run<2, 4, 4>("+LDS+sync 4x4", 512, out, g);
  run<3, 4, 4>("+LDS+sync+cpasync 4x4", 512, out, g);
  run<1, 4, 4>("+LDS 4x4 8 warps", 256, out, g);
  run<1, 4, 8>("+LDS 4x8 8 warps", 256, out, g);
  run<1, 8, 8>("+LDS 8x8 4 warps", 128, out, g);
  return 0;
}
