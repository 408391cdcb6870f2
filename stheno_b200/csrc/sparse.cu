// Streamed sparse (inducing-point) accumulation: one chunk of data points per call, K_zx is never held.
//
//   reference: AbstractPseudoObservations._compute, stheno/model/observations.py:279-336
//       :285  K_zx = k(z, x)              :301  W = L_z^-1 K_zx            :304-306 corr = diag K_x - colsum(W o W)
//       :308-313 trace part / FITC noise  :322  A = I + W K_n^-1 W^T       :327  prod = W K_n^-1 ybar
//       :334-336 the scalars of the ELBO
//   Round 1 materialised K_xz (8.6 GB at n = 262144, m = 4096), its transpose (8.6 GB) and one more copy.  Here the caller
//   walks the data in chunks of c points; per chunk (all stream-ordered, caller-provided workspace, no allocation):
//       K1 rows k(x_c, z) -> [c_pad, m_pad]  ->  right TRSM against L_z (rows are independent: the same tensor-core solve the
//       posterior uses)  ->  per-row |w_i|^2 (Q_ii)  ->  per-row scalars (corr, trace, FITC noise, 1/sqrt(K_n), ELBO sums)  ->
//       scaled transpose [m_pad, c_pad]  ->  A += W_s W_s^T (tensor-core SYRK on the lower tiles, K = c)  ->  prod += W_s ybar_s
//   so the m^2 n flops of the solve and of A both stay on the tensor cores and device memory is O(c m + m^2).
#include "common.cuh"

namespace gpk {

// per data point of the chunk: corr, method-specific noise, the scale 1/sqrt(K_n'), ybar scaled, and the three ELBO sums
template <typename T>
__global__ void sparse_rows_kernel(int64_t c, int64_t c_pad, const T* __restrict__ kdiag, const T* __restrict__ q,
                                   const T* __restrict__ kn, const T* __restrict__ ybar, int32_t method,
                                   T* __restrict__ rs, T* __restrict__ ybs, T* __restrict__ scalars) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double s_log = 0.0, s_yy = 0.0, s_tr = 0.0;
  if (i < c_pad) {
    if (i < c) {
      double k = (double)kn[i];
      if (method != 2) {
        const double corr = (double)kdiag[i] - (double)q[i];  // :306
        if (method == 0) s_tr = corr / k;                     // :308-310  B.ratio(Diagonal(corr), K_n)
        else k += corr;                                       // :311-313
      }
      const double r = rsqrt(k);
      const double yb = (double)ybar[i];
      rs[i] = (T)r;
      ybs[i] = (T)(yb * r);
      s_log = log(6.283185307179586476925286766559 * k);      // :334
      s_yy = yb * yb / k;                                     // :335
    } else {
      rs[i] = T(0);
      ybs[i] = T(0);
    }
  }
  __shared__ double red[3][8];
  s_log = warp_sum(s_log);
  s_yy = warp_sum(s_yy);
  s_tr = warp_sum(s_tr);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) {
    red[0][w] = s_log;
    red[1][w] = s_yy;
    red[2][w] = s_tr;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += red[threadIdx.x][k];
    atomicAdd(scalars + threadIdx.x, (T)s);
  }
}

// dst[j][i] = src[i][j] * rs[i]   (src: rows x cols, dst: cols x rows), 32 x 32 tiles through shared memory
template <typename T>
__global__ void transpose_scaled_kernel(const T* __restrict__ src, int64_t lds, int64_t rows, int64_t cols,
                                        const T* __restrict__ rs, T* __restrict__ dst, int64_t ldd) {
  __shared__ T tile[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int64_t r0 = (int64_t)blockIdx.y * 32, c0 = (int64_t)blockIdx.x * 32;
  for (int i = ty; i < 32; i += 8) {
    const int64_t r = r0 + i, cc = c0 + tx;
    tile[i][tx] = (r < rows && cc < cols) ? src[r * lds + cc] * rs[r] : T(0);
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int64_t r = c0 + i, cc = r0 + tx;
    if (r < cols && cc < rows) dst[r * ldd + cc] = tile[tx][i];
  }
}

// acc[r] += <V[r, :n_cols], b>   (one warp per row)
template <typename T>
__global__ void row_dot_acc_kernel(const T* __restrict__ V, int64_t ldv, int64_t rows, int64_t n_cols,
                                   const T* __restrict__ b, T* __restrict__ acc) {
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const T* row = V + r * ldv;
  double s = 0.0;
  for (int64_t j = lane; j < n_cols; j += 32) s = fma((double)row[j], (double)b[j], s);
  s = warp_sum(s);
  if (lane == 0) acc[r] += (T)s;
}

template <typename T>
struct SparseAbi;
template <>
struct SparseAbi<double> {
  static int km(const gpk_kernel_desc* d, const double* x, int64_t xg, int64_t n, const double* y, int64_t yg, int64_t n2, int32_t dim,
                double* out, int64_t ldo, void* s) {
    return gpk_kernel_matrix_f64(d, x, xg, 0, n, y, yg, 0, n2, dim, 0.0, nullptr, 0, 0.0, GPK_KM_PAD_ZERO, out, ldo, 0, 1, s);
  }
  static int trsm(const double* L, int64_t ldl, int64_t n, double* B, int64_t ldb, int64_t rows, void* s) {
    return gpk_trsm_right_f64(L, ldl, 0, n, B, ldb, 0, rows, 1, s);
  }
  static int sq(const double* V, int64_t ldv, int64_t rows, int64_t nc, double* out, void* s) {
    return gpk_row_dot_sq_f64(V, ldv, 0, rows, nc, nullptr, 0, nullptr, out, 0, 1, s);
  }
  static int syrk(int64_t M, int64_t K, const double* A, int64_t lda, double* C, int64_t ldc, void* s) {
    return gpk_gemm_nt_f64(M, M, K, 1.0, A, lda, 0, A, lda, 0, 1.0, C, ldc, 0, 1, 1, s);
  }
};
template <>
struct SparseAbi<float> {
  static int km(const gpk_kernel_desc* d, const float* x, int64_t xg, int64_t n, const float* y, int64_t yg, int64_t n2, int32_t dim,
                float* out, int64_t ldo, void* s) {
    return gpk_kernel_matrix_f32(d, x, xg, 0, n, y, yg, 0, n2, dim, 0.0, nullptr, 0, 0.0, GPK_KM_PAD_ZERO, out, ldo, 0, 1, s);
  }
  static int trsm(const float* L, int64_t ldl, int64_t n, float* B, int64_t ldb, int64_t rows, void* s) {
    return gpk_trsm_right_f32(L, ldl, 0, n, B, ldb, 0, rows, 1, s);
  }
  static int sq(const float* V, int64_t ldv, int64_t rows, int64_t nc, float* out, void* s) {
    return gpk_row_dot_sq_f32(V, ldv, 0, rows, nc, nullptr, 0, nullptr, out, 0, 1, s);
  }
  static int syrk(int64_t M, int64_t K, const float* A, int64_t lda, float* C, int64_t ldc, void* s) {
    return gpk_gemm_nt_f32(M, M, K, 1.0f, A, lda, 0, A, lda, 0, 1.0f, C, ldc, 0, 1, 1, s);
  }
};

static inline int64_t pad128(int64_t v) { return (v + 127) / 128 * 128; }

template <typename T>
static int sparse_accumulate(const gpk_kernel_desc* desc, const T* xg, int64_t xg_gstride, int64_t c, const T* zg,
                             int64_t zg_gstride, int64_t m, int32_t d, const T* Lz, int64_t ldl, int64_t m_pad,
                             const T* kdiag, const T* kn, const T* ybar, int32_t method, T* A, int64_t lda, T* prod,
                             T* scalars, T* ws, int64_t ws_elems, void* stream) {
  if (!desc || !xg || !zg || !Lz || !kn || !ybar || !A || !prod || !scalars || !ws) return GPK_ERR_ARG;
  if (c < 1 || m < 1 || d < 1 || m_pad % 128 || m_pad < m || ldl < m_pad || lda < m_pad) return GPK_ERR_ARG;
  if (method < 0 || method > 2 || (method != 2 && !kdiag)) return GPK_ERR_ARG;
  const int64_t c_pad = pad128(c);
  if (ws_elems < gpk_sparse_ws_elems(c, m_pad)) return GPK_ERR_ARG;
  if (reinterpret_cast<uintptr_t>(ws) % 16) return GPK_ERR_ALIGN;
  T* Wc = ws;                      // [c_pad][m_pad]
  T* WcT = Wc + c_pad * m_pad;     // [m_pad][c_pad]
  T* q = WcT + m_pad * c_pad;      // [c_pad]
  T* rs = q + c_pad;
  T* ybs = rs + c_pad;
  cudaStream_t s = (cudaStream_t)stream;
  int rc;
  if ((rc = SparseAbi<T>::km(desc, xg, xg_gstride, c, zg, zg_gstride, m, d, Wc, m_pad, stream))) return rc;   // :285
  if ((rc = SparseAbi<T>::trsm(Lz, ldl, m_pad, Wc, m_pad, c_pad, stream))) return rc;                          // :301
  if (method != 2 && (rc = SparseAbi<T>::sq(Wc, m_pad, c, m_pad, q, stream))) return rc;                       // :305
  sparse_rows_kernel<T><<<(unsigned)((c_pad + 255) / 256), 256, 0, s>>>(c, c_pad, kdiag, q, kn, ybar, method, rs, ybs, scalars);
  GPK_COUNT_LAUNCH();
  GPK_CHECK_LAUNCH();
  dim3 grid((unsigned)(m_pad / 32), (unsigned)(c_pad / 32)), block(32, 8);
  transpose_scaled_kernel<T><<<grid, block, 0, s>>>(Wc, m_pad, c_pad, m_pad, rs, WcT, c_pad);
  GPK_COUNT_LAUNCH();
  GPK_CHECK_LAUNCH();
  if ((rc = SparseAbi<T>::syrk(m_pad, c_pad, WcT, c_pad, A, lda, stream))) return rc;                           // :322
  row_dot_acc_kernel<T><<<(unsigned)((m_pad + 7) / 8), 256, 0, s>>>(WcT, c_pad, m_pad, c_pad, ybs, prod);       // :327
  GPK_COUNT_LAUNCH();
  GPK_CHECK_LAUNCH();
  return 0;
}

}  // namespace gpk

extern "C" {

int64_t gpk_sparse_ws_elems(int64_t c, int64_t m_pad) {
  const int64_t c_pad = gpk::pad128(c);
  return 2 * c_pad * m_pad + 3 * c_pad;
}

int gpk_sparse_accumulate_f64(const gpk_kernel_desc* desc_host, const double* xg, int64_t xg_gstride, int64_t c,
                              const double* zg, int64_t zg_gstride, int64_t m, int32_t d, const double* Lz, int64_t ldl,
                              int64_t m_pad, const double* kdiag, const double* kn, const double* ybar, int32_t method,
                              double* A, int64_t lda, double* prod, double* scalars, double* ws, int64_t ws_elems,
                              void* stream) {
  return gpk::sparse_accumulate<double>(desc_host, xg, xg_gstride, c, zg, zg_gstride, m, d, Lz, ldl, m_pad, kdiag, kn, ybar,
                                        method, A, lda, prod, scalars, ws, ws_elems, stream);
}
int gpk_sparse_accumulate_f32(const gpk_kernel_desc* desc_host, const float* xg, int64_t xg_gstride, int64_t c,
                              const float* zg, int64_t zg_gstride, int64_t m, int32_t d, const float* Lz, int64_t ldl,
                              int64_t m_pad, const float* kdiag, const float* kn, const float* ybar, int32_t method, float* A,
                              int64_t lda, float* prod, float* scalars, float* ws, int64_t ws_elems, void* stream) {
  return gpk::sparse_accumulate<float>(desc_host, xg, xg_gstride, c, zg, zg_gstride, m, d, Lz, ldl, m_pad, kdiag, kn, ybar,
                                       method, A, lda, prod, scalars, ws, ws_elems, stream);
}
}
