#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
    : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
    : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
template<int ILP>
__global__ void k884(double* out, int iters, double a, double b) {
  double c[ILP][2];
  for (int i = 0; i < ILP; i++) { c[i][0] = threadIdx.x; c[i][1] = i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0; for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template<int ILP>
__global__ void k1688(double* out, int iters, double a, double b) {
  double c[ILP][4]; double av[4] = {a, a+1, a+2, a+3}; double bv[2] = {b, b+1};
  for (int i = 0; i < ILP; i++) for (int j = 0; j < 4; j++) c[i][j] = threadIdx.x + j;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) dmma1688(c[i], av, bv);
  }
  double s = 0; for (int i = 0; i < ILP; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template<int ILP>
__global__ void k16816(double* out, int iters, double a, double b) {
  double c[ILP][4]; double av[8]; double bv[4];
  for (int j = 0; j < 8; j++) av[j] = a + j; for (int j = 0; j < 4; j++) bv[j] = b + j;
  for (int i = 0; i < ILP; i++) for (int j = 0; j < 4; j++) c[i][j] = threadIdx.x + j;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) dmma16816(c[i], av, bv);
  }
  double s = 0; for (int i = 0; i < ILP; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template<int ILP>
__global__ void kdfma(double* out, int iters, double a, double b) {
  double c[ILP];
  for (int i = 0; i < ILP; i++) c[i] = threadIdx.x + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) c[i] = fma(c[i], a, b);
  }
  double s = 0; for (int i = 0; i < ILP; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template<typename F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  double* out; cudaMalloc(&out, 148 * 8 * 1024 * sizeof(double));
  int iters = 20000;
  for (int warps : {4, 8, 16}) {
    int threads = warps * 32; int blocks = 148;
    float ms;
    ms = timeit([&]{ k884<8><<<blocks, threads>>>(out, iters, 1.0000001, 0.9999999); });
    printf("m8n8k4   warps/SM=%2d: %.2f TF/s\n", warps, 512.0 * 8 * iters * warps * blocks / ms / 1e9);
    ms = timeit([&]{ k1688<8><<<blocks, threads>>>(out, iters, 1.0000001, 0.9999999); });
    printf("m16n8k8  warps/SM=%2d: %.2f TF/s\n", warps, 2048.0 * 8 * iters * warps * blocks / ms / 1e9);
    ms = timeit([&]{ k16816<8><<<blocks, threads>>>(out, iters, 1.0000001, 0.9999999); });
    printf("m16n8k16 warps/SM=%2d: %.2f TF/s\n", warps, 4096.0 * 8 * iters * warps * blocks / ms / 1e9);
    ms = timeit([&]{ kdfma<16><<<blocks, threads>>>(out, iters, 1.0000001, 0.9999999); });
    printf("DFMA     warps/SM=%2d: %.2f TF/s\n", warps, 2.0 * 32 * 16 * iters * warps * blocks / ms / 1e9);
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
