// Where does the leaf Cholesky's per-column latency go?  Variants of potrf_leaf_kernel (timing only for V >= 3).
//   0 baseline   1 fast rsqrt (f32 seed + 2 Newton steps)   2 = 1 + no info/diag bookkeeping in the loop
//   3 no barrier (wrong results)   4 inv = const (no rsqrt; wrong results)   5 no FMA update (wrong results)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
constexpr int NB = 128;
template <int V>
__global__ void __launch_bounds__(256, 1) leaf(double* __restrict__ A, int64_t lda, double* __restrict__ logdet, int* __restrict__ info) {
  const int tid = threadIdx.x, ti = tid >> 4, tk = tid & 15;
  __shared__ double colbuf[2][NB];
  __shared__ double diag[NB];
  double acc[8][8];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[a][b] = A[(int64_t)(ti + 16 * a) * lda + tk + 16 * b];
  int buf = 0;
#pragma unroll
  for (int jb = 0; jb < 8; ++jb) {
    for (int jj = 0; jj < 16; ++jj) {
      const int j = jj + 16 * jb;
      if (tk == jj) {
#pragma unroll
        for (int a = 0; a < 8; ++a) colbuf[buf][ti + 16 * a] = acc[a][jb];
      }
      if (V != 3) __syncthreads();
      const double djj = colbuf[buf][j];
      if (V < 2) { if (tid == 0 && !(djj > 0.0)) atomicCAS(info, 0, j + 1); }
      double inv;
      if (V == 0 || V == 3 || V == 5) inv = rsqrt(djj);
      else if (V == 4) inv = 0.75;
      else {
        float s = rsqrtf((float)djj);
        double y = (double)s;
        const double h = 0.5 * djj;
        y = y * (1.5 - h * y * y);
        y = y * (1.5 - h * y * y);
        inv = y;
      }
      const double dsq = djj * inv;
      double li[8], lk[8];
#pragma unroll
      for (int a = 0; a < 8; ++a) li[a] = colbuf[buf][ti + 16 * a] * inv;
#pragma unroll
      for (int b = 0; b < 8; ++b) lk[b] = colbuf[buf][tk + 16 * b] * inv;
      if (V != 5) {
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        if (b < jb) continue;
        const bool col_ok = (b > jb) || (tk > jj);
#pragma unroll
        for (int a = 0; a < 8; ++a) {
          if (a < b) continue;
          const bool ok = col_ok && ((a > b) || (ti >= tk));
          if (ok) acc[a][b] -= li[a] * lk[b];
        }
      }
      }
      if (tk == jj) {
#pragma unroll
        for (int a = 0; a < 8; ++a) {
          const int i = ti + 16 * a;
          if (i > j) acc[a][jb] = li[a];
          else if (i == j) acc[a][jb] = dsq;
        }
      }
      if (V < 2) { if (tid == 0) diag[j] = dsq; }
      buf ^= 1;
    }
  }
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const int i = ti + 16 * a, k = tk + 16 * b;
      if (i >= k) A[(int64_t)i * lda + k] = acc[a][b];
    }
  if (V < 2 && tid == 0) logdet[0] = diag[5];
}
template <int V>
void run(double* dA, const std::vector<double>& h, double* dl, int* di, const std::vector<double>& ref) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9;
  for (int r = 0; r < 6; ++r) {
    cudaMemcpy(dA, h.data(), sizeof(double) * NB * NB, cudaMemcpyHostToDevice);
    cudaEventRecord(e0);
    leaf<V><<<1, 256>>>(dA, NB, dl, di);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  std::vector<double> out(NB * NB);
  cudaMemcpy(out.data(), dA, sizeof(double) * NB * NB, cudaMemcpyDeviceToHost);
  double err = 0;
  for (int i = 0; i < NB; ++i) for (int k = 0; k <= i; ++k) err = fmax(err, fabs(out[i * NB + k] - ref[i * NB + k]));
  printf("variant %d: %.2f us   max|L - Lref| = %.2e  (%s)\n", V, best * 1e3, err, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  std::vector<double> h(NB * NB), ref(NB * NB, 0.0);
  srand(1);
  std::vector<double> G(NB * NB);
  for (auto& g : G) g = rand() / (double)RAND_MAX - 0.5;
  for (int i = 0; i < NB; ++i) for (int k = 0; k < NB; ++k) { double s = (i == k) ? 8.0 : 0.0; for (int q = 0; q < NB; ++q) s += G[i * NB + q] * G[k * NB + q]; h[i * NB + k] = s; }
  ref = h;
  for (int j = 0; j < NB; ++j) { double d = sqrt(ref[j * NB + j]); ref[j * NB + j] = d; for (int i = j + 1; i < NB; ++i) ref[i * NB + j] /= d;
    for (int k = j + 1; k < NB; ++k) for (int i = k; i < NB; ++i) ref[i * NB + k] -= ref[i * NB + j] * ref[k * NB + j]; }
  double *dA, *dl; int* di; cudaMalloc(&dA, sizeof(double) * NB * NB); cudaMalloc(&dl, 8); cudaMalloc(&di, 4); cudaMemset(di, 0, 4);
  run<0>(dA, h, dl, di, ref); run<1>(dA, h, dl, di, ref); run<2>(dA, h, dl, di, ref); run<3>(dA, h, dl, di, ref); run<4>(dA, h, dl, di, ref); run<5>(dA, h, dl, di, ref);
  return 0;
}
