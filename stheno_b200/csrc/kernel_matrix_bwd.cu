// K1-backward: contraction of an upstream gradient G = d(loss)/dK (symmetric, n x n) with dK/d(theta) WITHOUT
// materialising any n x n gradient tensor per hyper-parameter.
//
// For K_ij = sum_t c_t prod_{f in t} phi_f(x_i^(g_f), x_j^(g_f))   (same points on both sides, x^(g) = x / l_g) it returns
//   term_sum[t]   = sum_ij G_ij prod_f phi_f(i, j)                       -> d loss / d c_t
//   grad_xg[g][i] = 2 sum_j G_ij sum_{t, f: g_f = g} c_t (prod_{f' != f} phi_f') d phi_f(x_i, x_j) / d x_i
//                                                                        -> d loss / d x^(g)_i  (chain rule to l_g and x
//                                                                           is left to torch autograd on x / l_g)
//   diag[i]       = G_ii                                                 -> d loss / d noise_i ; its sum for a scalar noise
// One CTA owns 64 rows and sweeps all column tiles; every thread keeps private partial sums for its 4 rows in shared
// memory and the 16 threads sharing a row are reduced once at the end.
//
// Reference: torch autograd through exp / cholesky / triangular_solve in readme_example13_optimisation_torch.py:46-53.
#include "common.cuh"

namespace gpk {

constexpr int KB_TILE = 64;
constexpr int KB_THREADS = 256;
constexpr int KB_MAXF = 4;  // factors per term supported by the backward pass

struct KbParams {
  gpk_kernel_desc desc;
  const void* xg;
  int64_t xg_gstride, x_bstride;
  int64_t n;
  int32_t d;
  const void* G;
  int64_t ldg, g_bstride;
  void* term_sum;  // [batch][GPK_MAX_TERMS]
  void* grad_xg;   // [groups][batch][n][d]  (same strides as xg)
  void* diag;      // [batch][n]
};

template <typename T>
__device__ __forceinline__ T kb_exp(T v) {
  return sizeof(T) == 8 ? (T)exp((double)v) : (T)expf((float)v);
}
template <typename T>
__device__ __forceinline__ T kb_sqrt(T v) {
  return sizeof(T) == 8 ? (T)sqrt((double)v) : (T)sqrtf((float)v);
}

// value and derivative w.r.t. the squared distance (for LINEAR: value = dot, dval = 1 marks d/d(dot))
template <typename T>
__device__ __forceinline__ void eval_factor_grad(int kind, T d2, T dot, bool same_pt, int d, T& val, T& dval, double param = 0.0) {
  switch (kind) {
    case GPK_RQ: {  // v = (1 + d2 / (2 a))^-a ;  dv / d(d2) = -v / (2 (1 + d2 / (2 a)))
      const double a = param, u = 1.0 + (double)d2 / (2.0 * a);
      const double v = exp(-a * log(u));
      val = (T)v;
      dval = (T)(-0.5 * v / u);
      return;
    }
    case GPK_EQ: {
      val = kb_exp<T>(T(-0.5) * d2);
      dval = T(-0.5) * val;
      return;
    }
    case GPK_MATERN12: {
      T r = (d == 1) ? kb_sqrt<T>(d2) : kb_sqrt<T>(d2 > T(1e-30) ? d2 : T(1e-30));
      val = kb_exp<T>(-r);
      dval = same_pt ? T(0) : -val / (T(2) * (r > T(1e-300) ? r : T(1e-300)));
      return;
    }
    case GPK_MATERN32: {
      T r = (d == 1) ? kb_sqrt<T>(d2) : kb_sqrt<T>(d2 > T(1e-30) ? d2 : T(1e-30));
      T s = T(1.7320508075688772) * r;
      T e = kb_exp<T>(-s);
      val = (T(1) + s) * e;
      dval = T(-1.5) * e;
      return;
    }
    case GPK_MATERN52: {
      T r = (d == 1) ? kb_sqrt<T>(d2) : kb_sqrt<T>(d2 > T(1e-30) ? d2 : T(1e-30));
      T s = T(2.23606797749979) * r;
      T e = kb_exp<T>(-s);
      val = (T(1) + s + T(1.6666666666666667) * d2) * e;
      dval = T(-0.8333333333333334) * (T(1) + s) * e;
      return;
    }
    case GPK_LINEAR:
      val = dot;
      dval = T(1);
      return;
    case GPK_DELTA:
      val = same_pt ? T(1) : T(0);
      dval = T(0);
      return;
    default:
      val = T(1);
      dval = T(0);
      return;
  }
}

template <typename T>
__global__ void __launch_bounds__(KB_THREADS) kernel_matrix_bwd_kernel(const KbParams p) {
  const int tile_r = blockIdx.x, b = blockIdx.y;
  const int d = p.d, G = p.desc.n_groups, nt = p.desc.n_terms;
  const int64_t r0 = (int64_t)tile_r * KB_TILE;
  extern __shared__ __align__(16) unsigned char kb_smem[];
  T* xs = reinterpret_cast<T*>(kb_smem);            // [G][64][d]   rows of this CTA
  T* yt = xs + (size_t)G * KB_TILE * d;              // [G][d][65]   current column tile, transposed
  T* part = yt + (size_t)G * d * (KB_TILE + 1);      // [256 threads][G][4 rows][d + 1]: (sum_j c_ij x_jk ..., sum_j w_ij)
  const T* xg = static_cast<const T*>(p.xg) + (int64_t)b * p.x_bstride;
  const T* Gm = static_cast<const T*>(p.G) + (int64_t)b * p.g_bstride;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int pstride = G * 4 * (d + 1);
  T* mypart = part + (size_t)tid * pstride;
  for (int i = 0; i < pstride; ++i) mypart[i] = T(0);

  const int xr = (int)max((int64_t)0, min((int64_t)KB_TILE, p.n - r0));
  for (int g = 0; g < G; ++g)
    for (int i = tid; i < KB_TILE * d; i += KB_THREADS)
      xs[(size_t)g * KB_TILE * d + i] = (i < xr * d) ? xg[g * p.xg_gstride + r0 * d + i] : T(0);

  T tsum[GPK_MAX_TERMS];
#pragma unroll
  for (int t = 0; t < GPK_MAX_TERMS; ++t) tsum[t] = T(0);

  const int n_ctiles = (int)((p.n + KB_TILE - 1) / KB_TILE);
  for (int tc = 0; tc < n_ctiles; ++tc) {
    const int64_t c0 = (int64_t)tc * KB_TILE;
    const int yr = (int)min((int64_t)KB_TILE, p.n - c0);
    __syncthreads();
    for (int idx = tid; idx < G * KB_TILE * d; idx += KB_THREADS) {
      const int g = idx / (KB_TILE * d), rem = idx - g * KB_TILE * d;
      const int c = rem / d, k = rem - c * d;
      yt[((size_t)g * d + k) * (KB_TILE + 1) + c] = (c < yr) ? xg[g * p.xg_gstride + (c0 + c) * d + k] : T(0);
    }
    __syncthreads();

#pragma unroll 1
    for (int i = 0; i < 4; ++i) {
      const int64_t r = r0 + ty * 4 + i;
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {
        const int64_t c = c0 + tx + 16 * j;
        if (r >= p.n || c >= p.n) continue;
        const T gij = Gm[r * p.ldg + c];
        const bool same_pt = (r == c);
        for (int t = 0; t < nt; ++t) {
          const int f0 = p.desc.term_begin[t], f1 = p.desc.term_begin[t + 1];
          T val[KB_MAXF], dval[KB_MAXF];
          T prod = T(1);
#pragma unroll
          for (int q = 0; q < KB_MAXF; ++q) {
            val[q] = T(1);
            dval[q] = T(0);
            if (f0 + q < f1) {
              const int g = p.desc.fac_group[f0 + q];
              const T* xr_ = xs + ((size_t)g * KB_TILE + ty * 4 + i) * d;
              const T* yc_ = yt + (size_t)g * d * (KB_TILE + 1) + tx + 16 * j;
              T d2 = T(0), dot = T(0);
              for (int k = 0; k < d; ++k) {
                const T xv = xr_[k], yv = yc_[(size_t)k * (KB_TILE + 1)];
                const T df = xv - yv;
                d2 = fma(df, df, d2);
                dot = fma(xv, yv, dot);
              }
              eval_factor_grad<T>(p.desc.fac_kind[f0 + q], d2, dot, same_pt, d, val[q], dval[q], p.desc.fac_param[f0 + q]);
              prod *= val[q];
            }
          }
#pragma unroll
          for (int tt = 0; tt < GPK_MAX_TERMS; ++tt)
            if (tt == t) tsum[tt] = fma(gij, prod, tsum[tt]);
          const T ct = (T)p.desc.coef[t];
#pragma unroll
          for (int q = 0; q < KB_MAXF; ++q) {
            if (f0 + q >= f1 || dval[q] == T(0)) continue;
            T others = T(1);
#pragma unroll
            for (int q2 = 0; q2 < KB_MAXF; ++q2)
              if (q2 != q) others *= val[q2];
            const T wgt = gij * ct * others * dval[q];
            const int g = p.desc.fac_group[f0 + q];
            const int kind = p.desc.fac_kind[f0 + q];
            T* pp = mypart + ((size_t)g * 4 + i) * (d + 1);
            const T* yc_ = yt + (size_t)g * d * (KB_TILE + 1) + tx + 16 * j;
            if (kind == GPK_LINEAR) {
              // d(dot)/dx_ik = x_jk ; factor 2 for the symmetric counterpart
              for (int k = 0; k < d; ++k) pp[k] = fma(T(2) * wgt, yc_[(size_t)k * (KB_TILE + 1)], pp[k]);
            } else {
              // d(d2)/dx_ik = 2 (x_ik - x_jk) ; factor 2 for the symmetric counterpart:
              //   grad_ik += 4 wgt x_ik - 4 wgt x_jk  -> keep sum_j wgt in slot d, sum_j wgt x_jk in slot k
              for (int k = 0; k < d; ++k) pp[k] = fma(T(-4) * wgt, yc_[(size_t)k * (KB_TILE + 1)], pp[k]);
              pp[d] += T(4) * wgt;
            }
          }
        }
        if (same_pt && p.diag) static_cast<T*>(p.diag)[(int64_t)b * p.n + r] = gij;
      }
    }
  }
  __syncthreads();
  // reduce the 16 column-threads of every row and write grad_xg
  T* gout = static_cast<T*>(p.grad_xg) + (int64_t)b * p.x_bstride;
  for (int idx = tid; idx < G * KB_TILE * d; idx += KB_THREADS) {
    const int g = idx / (KB_TILE * d), rem = idx - g * KB_TILE * d;
    const int row = rem / d, k = rem - row * d;
    if (r0 + row >= p.n) continue;
    const int rty = row >> 2, ri = row & 3;
    T s = T(0), sw = T(0);
    for (int t16 = 0; t16 < 16; ++t16) {
      const T* pp = part + (size_t)(rty * 16 + t16) * pstride + ((size_t)g * 4 + ri) * (d + 1);
      s += pp[k];
      sw += pp[d];
    }
    gout[g * p.xg_gstride + (r0 + row) * d + k] = s + sw * xs[(size_t)g * KB_TILE * d + row * d + k];
  }
  // term sums: block reduction + one atomic per term
  __shared__ T red[8][GPK_MAX_TERMS];
#pragma unroll
  for (int t = 0; t < GPK_MAX_TERMS; ++t) {
    T v = warp_sum(tsum[t]);
    if ((tid & 31) == 0) red[tid >> 5][t] = v;
  }
  __syncthreads();
  if (tid < nt) {
    T s = T(0);
    for (int w = 0; w < 8; ++w) s += red[w][tid];
    atomicAdd(static_cast<T*>(p.term_sum) + (int64_t)b * GPK_MAX_TERMS + tid, s);
  }
}

template <typename T>
static int launch_kernel_matrix_bwd(const gpk_kernel_desc* desc, const T* xg, int64_t xg_gstride, int64_t x_bstride,
                                    int64_t n, int32_t d, const T* G, int64_t ldg, int64_t g_bstride, T* term_sum,
                                    T* grad_xg, T* diag, int32_t batch, void* stream) {
  if (!desc || !xg || !G || !term_sum || !grad_xg || n < 0 || d < 1 || batch < 1) return GPK_ERR_ARG;
  if (desc->n_terms < 0 || desc->n_terms > GPK_MAX_TERMS || desc->n_groups < 1 || desc->n_groups > GPK_MAX_GROUPS)
    return GPK_ERR_ARG;
  for (int t = 0; t < desc->n_terms; ++t)
    if (desc->term_begin[t + 1] - desc->term_begin[t] > KB_MAXF) return GPK_ERR_UNSUPPORTED;
  if (n == 0) return 0;
  KbParams p;
  p.desc = *desc;
  p.xg = xg;
  p.xg_gstride = xg_gstride;
  p.x_bstride = x_bstride;
  p.n = n;
  p.d = d;
  p.G = G;
  p.ldg = ldg;
  p.g_bstride = g_bstride;
  p.term_sum = term_sum;
  p.grad_xg = grad_xg;
  p.diag = diag;
  const int Gn = desc->n_groups;
  const size_t smem =
      ((size_t)Gn * KB_TILE * d + (size_t)Gn * d * (KB_TILE + 1) + (size_t)KB_THREADS * Gn * 4 * (d + 1)) * sizeof(T);
  if (smem > 200 * 1024) return GPK_ERR_UNSUPPORTED;
  auto kern = kernel_matrix_bwd_kernel<T>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return -1000 - (int)e;
  }
  dim3 grid((unsigned)((n + KB_TILE - 1) / KB_TILE), (unsigned)batch);
  kern<<<grid, KB_THREADS, smem, (cudaStream_t)stream>>>(p);
  GPK_COUNT_LAUNCH();
  GPK_CHECK_LAUNCH();
  return 0;
}

}  // namespace gpk

extern "C" {
int gpk_kernel_matrix_bwd_f64(const gpk_kernel_desc* desc_host, const double* xg, int64_t xg_gstride,
                              int64_t x_bstride, int64_t n, int32_t d, const double* G, int64_t ldg, int64_t g_bstride,
                              double* term_sum, double* grad_xg, double* diag, int32_t batch, void* stream) {
  return gpk::launch_kernel_matrix_bwd<double>(desc_host, xg, xg_gstride, x_bstride, n, d, G, ldg, g_bstride, term_sum,
                                               grad_xg, diag, batch, stream);
}
int gpk_kernel_matrix_bwd_f32(const gpk_kernel_desc* desc_host, const float* xg, int64_t xg_gstride, int64_t x_bstride,
                              int64_t n, int32_t d, const float* G, int64_t ldg, int64_t g_bstride, float* term_sum,
                              float* grad_xg, float* diag, int32_t batch, void* stream) {
  return gpk::launch_kernel_matrix_bwd<float>(desc_host, xg, xg_gstride, x_bstride, n, d, G, ldg, g_bstride, term_sum,
                                              grad_xg, diag, batch, stream);
}
}
