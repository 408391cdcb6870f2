/*
 * gpk.h -- C-ABI of libgpk: the B200 (sm_100a) kernels behind the Stheno GP-inference hot path.
 *
 * The reference (wesselb/stheno) has no FFI: its operator boundary is Python multiple dispatch into
 * `lab.B.*` / `matrix` / `mlkernels`.  Each entry point below names the reference
 * call site(s) whose arithmetic it replaces (paths relative to the reference tree).
 *
 * Conventions
 *   - All pointers are DEVICE pointers unless the parameter name ends in `_host`.
 *   - Matrices are ROW-MAJOR with an explicit leading dimension `ld` (in elements).
 *   - "Padded" matrices have their dimensions rounded up to a multiple of GPK_TILE (128); the padding of a
 *     matrix that will be factorised is the identity (1 on the diagonal, 0 elsewhere), padding of right-hand
 *     sides is 0.  `gpk_round_up(n)` gives the padded size.
 *   - Only the LOWER triangle of symmetric matrices / Cholesky factors is read or written.
 *   - Functions are stateless, re-entrant and stream-ordered on `stream` (a cudaStream_t passed as void*);
 *     they never allocate device memory and never synchronise the host.
 *   - Return value: 0 = launched OK; < 0 = bad argument (GPK_ERR_*) or a CUDA launch error (-1000 - cudaError).
 *     Numerical failure (non-positive pivot) is reported LAPACK-style through the device-side `info` word
 *     (index of the first bad pivot, 1-based; 0 = success) so that no host sync is forced.
 *   - `_f64` / `_f32` suffix = arithmetic type (double / float); everything is computed in that type.
 */
#ifndef GPK_H_
#define GPK_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPK_TILE 128
#define GPK_VERSION 100

#define GPK_ERR_ARG (-1)
#define GPK_ERR_ALIGN (-2)
#define GPK_ERR_UNSUPPORTED (-3)

/* ---- kernel-expression descriptor -------------------------------------------------------------------
 * A kernel is flattened by the host into a sum of products:
 *     k(x, y) = sum_t coef[t] * prod_{f in term t} phi_{kind[f]}( x^(group[f]), y^(group[f]) )
 * where x^(g) = x / lengthscale_g is a pre-stretched copy of the inputs ("group" g), exactly as the
 * reference evaluates `k.stretch(l)` by dividing the inputs (mlkernels Stretched; call sites
 * stheno/model/fdd.py:79, stheno/model/observations.py:139,285,286).
 */
enum gpk_kind {
  GPK_EQ = 0,       /* exp(-r^2/2)                                  */
  GPK_MATERN12 = 1, /* exp(-r)                                      */
  GPK_MATERN32 = 2, /* (1 + sqrt3 r) exp(-sqrt3 r)                  */
  GPK_MATERN52 = 3, /* (1 + sqrt5 r + 5 r^2 / 3) exp(-sqrt5 r)      */
  GPK_LINEAR = 4,   /* <x, y>                                       */
  GPK_DELTA = 5,    /* same inputs: [i == j]; else [r^2 < 1e-10]    */
  GPK_ONE = 6,      /* 1                                            */
  GPK_RQ = 7        /* (1 + r^2 / (2 alpha))^-alpha, alpha = fac_param[f] (mlkernels RQ; README.md:1076-1088) */
};

#define GPK_MAX_TERMS 8
#define GPK_MAX_FACTORS 16
#define GPK_MAX_GROUPS 8

typedef struct gpk_kernel_desc {
  int32_t n_terms;
  int32_t n_groups;
  int32_t term_begin[GPK_MAX_TERMS + 1]; /* factors of term t: [term_begin[t], term_begin[t+1]) */
  int32_t fac_kind[GPK_MAX_FACTORS];
  int32_t fac_group[GPK_MAX_FACTORS];
  double coef[GPK_MAX_TERMS];
  double fac_param[GPK_MAX_FACTORS]; /* per-factor shape parameter (GPK_RQ: alpha); unused by the other kinds */
} gpk_kernel_desc;

/* flags for gpk_kernel_matrix_* */
#define GPK_KM_LOWER 1        /* x and y are the same points: write only tiles on/below the diagonal      */
#define GPK_KM_SAME 2         /* x and y are the same object (Delta -> [i == j]; diagonal terms apply)     */
#define GPK_KM_PAD_IDENTITY 4 /* fill rows/cols >= n up to the padded size with the identity               */
#define GPK_KM_PAD_ZERO 8     /* fill rows/cols >= n up to the padded size with zeros                      */

int gpk_version(void);
int64_t gpk_round_up(int64_t n);

/* K1: fused pairwise-distance + kernel evaluation (+ diagonal noise + Cholesky jitter).
 * Replaces `p.kernel(x)` / `B.add(K, noise)` / `B.reg` : stheno/model/fdd.py:79,
 * stheno/model/observations.py:139,285,286 and the `+ B.epsilon I` of every B.cholesky (README.md:820-830).
 *   xg: [n_groups][batch][n][d]  pre-stretched inputs (strides: xg_gstride, x_bstride, d)
 *   yg: same for the second argument (n2 points); may equal xg.
 *   out[b][i][j] (ld = ldo, batch stride = o_bstride), i < rows_out, j < cols_out where rows_out/cols_out are
 *   n / n2 rounded up to GPK_TILE when a PAD flag is given, else n / n2.
 *   Diagonal (only with GPK_KM_SAME): out[i][i] += noise_scalar (+ noise_vec[b][i] if non-NULL), then += jitter. */
int gpk_kernel_matrix_f64(const gpk_kernel_desc* desc_host, const double* xg, int64_t xg_gstride, int64_t x_bstride,
                          int64_t n, const double* yg, int64_t yg_gstride, int64_t y_bstride, int64_t n2, int32_t d,
                          double noise_scalar, const double* noise_vec, int64_t nv_bstride, double jitter,
                          int32_t flags, double* out, int64_t ldo, int64_t o_bstride, int32_t batch, void* stream);
int gpk_kernel_matrix_f32(const gpk_kernel_desc* desc_host, const float* xg, int64_t xg_gstride, int64_t x_bstride,
                          int64_t n, const float* yg, int64_t yg_gstride, int64_t y_bstride, int64_t n2, int32_t d,
                          double noise_scalar, const float* noise_vec, int64_t nv_bstride, double jitter,
                          int32_t flags, float* out, int64_t ldo, int64_t o_bstride, int32_t batch, void* stream);

/* elwise: out[b][i] = k(x_i, y_i)  -- `k.elwise(x)` at stheno/model/fdd.py:66, observations.py:304. */
int gpk_kernel_diag_f64(const gpk_kernel_desc* desc_host, const double* xg, int64_t xg_gstride, int64_t x_bstride,
                        const double* yg, int64_t yg_gstride, int64_t y_bstride, int64_t n, int32_t d, int32_t same,
                        double* out, int64_t o_bstride, int32_t batch, void* stream);
int gpk_kernel_diag_f32(const gpk_kernel_desc* desc_host, const float* xg, int64_t xg_gstride, int64_t x_bstride,
                        const float* yg, int64_t yg_gstride, int64_t y_bstride, int64_t n, int32_t d, int32_t same,
                        float* out, int64_t o_bstride, int32_t batch, void* stream);

/* K1-backward: contraction of an upstream gradient G = d(loss)/dK (symmetric n x n, ld = ldg) with dK/d(theta) for
 * K = k(x, x) (same points): term_sum[b][t] += sum_ij G_ij prod_f phi_f  (d loss / d coef_t; caller zeroes it; stride
 * GPK_MAX_TERMS), grad_xg[g][b][i][:] = d loss / d x^(g)_i (layout and strides of xg), diag[b][i] = G_ii (noise
 * gradient).  No n x n tensor per hyper-parameter is ever formed.  Replaces torch autograd through
 * exp / pw_dists2 in the reference's optimisation loop (readme_example13_optimisation_torch.py:46-53). */
int gpk_kernel_matrix_bwd_f64(const gpk_kernel_desc* desc_host, const double* xg, int64_t xg_gstride,
                              int64_t x_bstride, int64_t n, int32_t d, const double* G, int64_t ldg, int64_t g_bstride,
                              double* term_sum, double* grad_xg, double* diag, int32_t batch, void* stream);
int gpk_kernel_matrix_bwd_f32(const gpk_kernel_desc* desc_host, const float* xg, int64_t xg_gstride, int64_t x_bstride,
                              int64_t n, int32_t d, const float* G, int64_t ldg, int64_t g_bstride, float* term_sum,
                              float* grad_xg, float* diag, int32_t batch, void* stream);

/* GEMM  C = beta*C + alpha * A * B^T   (A: M x K, B: N x K, both K-contiguous; C: M x N).
 * M, N multiples of 128; K a multiple of 16; pointers 16-byte aligned; ld multiples of 2.
 * lower != 0: only tiles with (row tile >= col tile) are touched (SYRK-style trailing update, M >= N).
 * Replaces the BLAS-3 inside B.cholesky / B.iqf / B.mm (stheno/random.py:274-276,
 * stheno/model/observations.py:301,322,323). */
int gpk_gemm_nt_f64(int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t lda, int64_t a_bstride,
                    const double* B, int64_t ldb, int64_t b_bstride, double beta, double* C, int64_t ldc,
                    int64_t c_bstride, int32_t lower, int32_t batch, void* stream);
int gpk_gemm_nt_f32(int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t lda, int64_t a_bstride,
                    const float* B, int64_t ldb, int64_t b_bstride, float beta, float* C, int64_t ldc,
                    int64_t c_bstride, int32_t lower, int32_t batch, void* stream);

/* K2: blocked right-looking Cholesky, in place, lower, row-major.
 *   A: [(n_pad + extra_rows) x n_pad] (ld = lda): the first n_pad rows hold the (padded) SPD matrix; the
 *   `extra_rows` (multiple of 128, may be 0) rows below it hold right-hand sides b^T, one per row, which come out
 *   as (L^-1 b)^T -- the triangular solve of B.iqf_diag fused into the factorisation (stheno/random.py:276).
 *   logdet[b] += 2 sum log diag(L) (caller zeroes it); info[b] = first non-positive pivot (1-based) or 0.
 * Replaces B.cholesky + B.logdet: stheno/random.py:274, stheno/model/observations.py:300,334. */
int gpk_potrf_f64(double* A, int64_t lda, int64_t a_bstride, int64_t n_pad, int64_t extra_rows, double* logdet,
                  int32_t* info, int32_t batch, void* stream);
int gpk_potrf_f32(float* A, int64_t lda, int64_t a_bstride, int64_t n_pad, int64_t extra_rows, float* logdet,
                  int32_t* info, int32_t batch, void* stream);

/* Opt-in mixed precision (BASELINE north_star: "tf32/bf16 where the user opts in"): same factorisation, but the
 * K >= 128 trailing updates of the fp64 matrix are formed from an fp32 copy of the panel by the tcgen05 3xTF32 kernel
 * (fp32-level products, fp64 accumulation into the matrix).  `ws`: float workspace of >= (n_pad + extra_rows) * 512
 * elements.  Results agree with gpk_potrf_f64 to ~1e-6 relative, NOT to the 1e-10 parity bar: never the default. */
int gpk_potrf_f64_tf32x3(double* A, int64_t lda, int64_t a_bstride, int64_t n_pad, int64_t extra_rows, double* logdet,
                         int32_t* info, int32_t batch, float* ws, int64_t ws_elems, void* stream);

/* fp64 emulation on the INT8 tensor cores (tcgen05.mma.kind::i8; Ozaki splitting): every row of the panel is scaled by
 * a power of two and split error-free into `slices` signed 7-bit integers; the slice products are EXACT in int32 and are
 * recombined in fp64.  6 slices: 21 int8 GEMMs, product error ~2^-40 |a||b| (zero-mean); 7 slices: 28 GEMMs, ~2^-47.
 * `ws`: 1024-byte aligned workspace of gpk_potrf_oz_ws_bytes(n_pad, extra_rows, slices) bytes.  Same contract as
 * gpk_potrf_f64 otherwise (batch must be 1 for the emulated update; batched problems fall back to the DMMA update). */
int64_t gpk_potrf_oz_ws_bytes(int64_t n_pad, int64_t extra_rows, int32_t slices);
int gpk_potrf_f64_oz(double* A, int64_t lda, int64_t a_bstride, int64_t n_pad, int64_t extra_rows, double* logdet,
                     int32_t* info, int32_t batch, int32_t slices, void* ws, int64_t ws_bytes, void* stream);
/* Library-wide switch (like cublasSetMathMode), per host thread and device: with slices = 5..8 and a caller-owned, 1024-byte
 * aligned scratch buffer, the LARGE fp64 GEMM-shaped updates inside gpk_potrf_f64 (trailing updates, n_pad >= 2048),
 * gpk_trsm_right_f64 and gpk_gemm_nt_f64 (batch 1, M, N >= 256, M N K >= 1.5e9, K <= 65536) run on the int8 tensor cores
 * whenever the scratch is large enough (gpk_f64_emulation_scratch_bytes for a GEMM, gpk_potrf_oz_ws_bytes for a
 * factorisation); everything else stays on the fp64 tensor cores.  slices = 0 switches it off.  Calls that use the scratch
 * must be stream-ordered with respect to each other. */
int gpk_set_f64_emulation(int32_t slices, void* scratch, int64_t scratch_bytes);
int64_t gpk_f64_emulation_scratch_bytes(int64_t M, int64_t N, int64_t K, int32_t slices);

/* The emulated GEMM on its own: C = beta C + alpha A B^T (M % 128 == 0, N % 64 == 0, K % 128 == 0, K <= 65536).
 * `ws`: 1024-byte aligned, >= round_up(gpk_oz_ws_bytes(M, K, slices), 1024) + gpk_oz_ws_bytes(N, K, slices) bytes. */
int64_t gpk_oz_ws_bytes(int64_t rows, int64_t K, int32_t slices);
int gpk_gemm_nt_f64_oz(int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t lda, const double* B,
                       int64_t ldb, double beta, double* C, int64_t ldc, int32_t lower, int32_t slices, void* ws,
                       int64_t ws_bytes, void* stream);

/* K3: X * L^T = B  in place (B: rows x n_pad, rows a multiple of 64; L: n_pad x n_pad lower).
 * Row r of the result is (L^-1 b_r)^T.  Replaces B.solve(L, .) / B.iqf:  stheno/model/observations.py:301 and
 * the PosteriorMean / PosteriorKernel evaluation behind observations.py:143-168. */
int gpk_trsm_right_f64(const double* L, int64_t ldl, int64_t l_bstride, int64_t n_pad, double* B, int64_t ldb,
                       int64_t b_bstride, int64_t rows, int32_t batch, void* stream);
int gpk_trsm_right_f32(const float* L, int64_t ldl, int64_t l_bstride, int64_t n_pad, float* B, int64_t ldb,
                       int64_t b_bstride, int64_t rows, int32_t batch, void* stream);

/* X * L = B in place (B: rows x n_pad): row r of the result is (L^-T b_r)^T.  Backward substitution used for
 * K^-1 b = L^-T L^-1 b (autograd, sampling from the posterior precision, B.solve with transposes). */
int gpk_trsm_right_t_f64(const double* L, int64_t ldl, int64_t l_bstride, int64_t n_pad, double* B, int64_t ldb,
                         int64_t b_bstride, int64_t rows, int32_t batch, void* stream);
int gpk_trsm_right_t_f32(const float* L, int64_t ldl, int64_t l_bstride, int64_t n_pad, float* B, int64_t ldb,
                         int64_t b_bstride, int64_t rows, int32_t batch, void* stream);

/* K4: log-marginal finish: out[b][c] = -0.5 * (logdet[b] + n * log(2 pi) + sum_j a[b][c][j]^2), c < k, where row c
 * of `a` (ld = lda) is (L^-1 (y_c - mu))^T.  stheno/random.py:272-279. */
int gpk_logpdf_finish_f64(const double* a, int64_t lda, int64_t a_bstride, int64_t n, int64_t n_cols, int32_t k,
                          const double* logdet, double* out, int32_t batch, void* stream);
int gpk_logpdf_finish_f32(const float* a, int64_t lda, int64_t a_bstride, int64_t n, int64_t n_cols, int32_t k,
                          const float* logdet, float* out, int32_t batch, void* stream);

/* Row reductions over V (rows x n_cols, ld = ldv):  dot[r] = sum_j V[r][j] * b[j] (b may be NULL),
 * sq[r] = sum_j V[r][j]^2 (sq may be NULL).  Posterior mean  m(x*) + V b  and marginal variance
 * k(x*,x*) - sum V^2 (mlkernels mean_var_diag via stheno/model/fdd.py:72-74); B.matmul_diag at observations.py:305. */
int gpk_row_dot_sq_f64(const double* V, int64_t ldv, int64_t v_bstride, int64_t rows, int64_t n_cols,
                       const double* b, int64_t b_bstride, double* dot, double* sq, int64_t o_bstride,
                       int32_t batch, void* stream);
int gpk_row_dot_sq_f32(const float* V, int64_t ldv, int64_t v_bstride, int64_t rows, int64_t n_cols, const float* b,
                       int64_t b_bstride, float* dot, float* sq, int64_t o_bstride, int32_t batch, void* stream);

/* Layout helpers (padding, symmetrisation, transposition) -- the `B.dense`, `B.transpose`, `B.reg` plumbing.
 * gpk_pad_copy: dst[(rows_pad) x (cols_pad)] = src[rows x cols] (+ diag_add on the diagonal), padding = identity
 *   (pad_identity != 0) or zero.  gpk_symmetrize: mirror the lower triangle into the upper one (n x n).
 * gpk_transpose: dst[cols x rows] = src[rows x cols]^T. */
int gpk_pad_copy_f64(const double* src, int64_t lds, int64_t s_bstride, int64_t rows, int64_t cols, double* dst,
                     int64_t ldd, int64_t d_bstride, int64_t rows_pad, int64_t cols_pad, double diag_add,
                     int32_t pad_identity, int32_t batch, void* stream);
int gpk_pad_copy_f32(const float* src, int64_t lds, int64_t s_bstride, int64_t rows, int64_t cols, float* dst,
                     int64_t ldd, int64_t d_bstride, int64_t rows_pad, int64_t cols_pad, double diag_add,
                     int32_t pad_identity, int32_t batch, void* stream);
int gpk_symmetrize_f64(double* A, int64_t lda, int64_t a_bstride, int64_t n, int32_t batch, void* stream);
int gpk_symmetrize_f32(float* A, int64_t lda, int64_t a_bstride, int64_t n, int32_t batch, void* stream);
int gpk_transpose_f64(const double* src, int64_t lds, int64_t s_bstride, int64_t rows, int64_t cols, double* dst,
                      int64_t ldd, int64_t d_bstride, int32_t batch, void* stream);
int gpk_transpose_f32(const float* src, int64_t lds, int64_t s_bstride, int64_t rows, int64_t cols, float* dst,
                      int64_t ldd, int64_t d_bstride, int32_t batch, void* stream);

/* K3 in one call: posterior mean and marginal variance terms at m test points (PosteriorMean / PosteriorKernel behind
 * stheno/model/observations.py:143-168, mlkernels.mean_var_diag via stheno/model/fdd.py:72-74), batch 1:
 *   V^T = k(x*, x) L^-T (K1 rows + right TRSM);  dot[i] = <v_i, half_y>;  sq[i] = |v_i|^2
 * half_y = (L^-1 (y - m(x)))^T zero-padded to n_pad (dot may be NULL: variance only; sq may be NULL: mean only).
 * The caller adds the prior mean / subtracts sq from the prior variance.  Test points are processed in chunks of `chunk`
 * rows (multiple of 128) through `ws` (>= chunk * n_pad elements, 16-byte aligned): K(x*, x) is never held whole. */
int gpk_posterior_marginals_f64(const gpk_kernel_desc* desc_host, const double* xsg, int64_t xsg_gstride, int64_t m,
                                const double* xg, int64_t xg_gstride, int64_t n, int32_t d, const double* L, int64_t ldl,
                                int64_t n_pad, const double* half_y, double* dot, double* sq, int64_t chunk, double* ws,
                                int64_t ws_elems, void* stream);
int gpk_posterior_marginals_f32(const gpk_kernel_desc* desc_host, const float* xsg, int64_t xsg_gstride, int64_t m,
                                const float* xg, int64_t xg_gstride, int64_t n, int32_t d, const float* L, int64_t ldl,
                                int64_t n_pad, const float* half_y, float* dot, float* sq, int64_t chunk, float* ws,
                                int64_t ws_elems, void* stream);

/* Streamed sparse (inducing-point) accumulation -- AbstractPseudoObservations._compute, stheno/model/observations.py:279-336,
 * one chunk of `c` data points per call; K_zx (8.6 GB at n = 262144, m = 4096) is never held.  Per chunk, stream-ordered:
 *   W_c^T = k(x_c, z) L_z^-T  (:285, :301; rows = data points, [c_pad x m_pad], c_pad = round_up(c))
 *   corr_i = kdiag_i - |w_i|^2 (:304-306);  method 0 (VFE): scalars[2] += sum corr_i / kn_i (:308-310);
 *   method 1 (FITC): kn_i += corr_i (:311-313);  method 2 (DTC): neither (kdiag may be NULL)
 *   A    += W diag(1/kn) W^T   (:322; lower 128-tiles of the m_pad x m_pad accumulator, the caller starts it at I)
 *   prod += W diag(1/kn) ybar  (:327);  scalars[0] += sum log(2 pi kn_i) (:334);  scalars[1] += sum ybar_i^2 / kn_i (:335)
 * xg / zg: pre-stretched inputs [n_groups][c or m][d] (group strides given); Lz: padded lower factor of K_z + eps I
 * (gpk_potrf); ws: 16-byte aligned workspace of gpk_sparse_ws_elems(c, m_pad) elements.  The m^2 c flops of the solve and of
 * the accumulation run on the tensor cores (the int8 emulation when gpk_set_f64_emulation is on and its scratch fits
 * gpk_f64_emulation_scratch_bytes(m_pad, m_pad, c_pad) and the solve's largest product). */
int64_t gpk_sparse_ws_elems(int64_t c, int64_t m_pad);
int gpk_sparse_accumulate_f64(const gpk_kernel_desc* desc_host, const double* xg, int64_t xg_gstride, int64_t c,
                              const double* zg, int64_t zg_gstride, int64_t m, int32_t d, const double* Lz, int64_t ldl,
                              int64_t m_pad, const double* kdiag, const double* kn, const double* ybar, int32_t method,
                              double* A, int64_t lda, double* prod, double* scalars, double* ws, int64_t ws_elems,
                              void* stream);
int gpk_sparse_accumulate_f32(const gpk_kernel_desc* desc_host, const float* xg, int64_t xg_gstride, int64_t c,
                              const float* zg, int64_t zg_gstride, int64_t m, int32_t d, const float* Lz, int64_t ldl,
                              int64_t m_pad, const float* kdiag, const float* kn, const float* ybar, int32_t method, float* A,
                              int64_t lda, float* prod, float* scalars, float* ws, int64_t ws_elems, void* stream);

/* Measurement helper (bench.py): runs a register-resident fp64 tensor-core (DMMA) loop on every SM and returns the
 * achieved TFLOP/s -- the denominator of the fp64 roofline -- or a negative error code.  Synchronises the device. */
double gpk_probe_dmma_tflops(void);

/* In-situ timing of the dominant kernel (the fp64 tensor-core GEMM) for bench.py's roofline leg: while enabled every
 * launch is bracketed by CUDA events on its own stream; read() (after the caller synchronised the device) returns the
 * summed durations, the summed ALGORITHMIC flops (2*128*128*K per computed tile) and the launch count.
 * enable(0/1) also clears the record. */
void gpk_gemm_profile_enable(int32_t on);
int gpk_gemm_profile_read_kind(int32_t kind, double* total_ms, double* total_flops, int64_t* launches); /* 0 DMMA, 1 int8 emulation, -1 all */
int gpk_gemm_profile_read(double* total_ms_host, double* total_flops_host, int64_t* launches_host);

/* Test helper (no GPU needed): the tile order of the emulated GEMM evaluated on the host -- tile index t of a launch with
 * tiles_m x tiles_n tiles (128 x 64; `lower`: only the tiles touching the lower triangle; `band`: tile rows per band) ->
 * (*tm, *tn).  Returns the number of tiles of the launch. */
int32_t gpk_debug_oz_tile(int32_t lower, int32_t tiles_m, int32_t tiles_n, int32_t band, int32_t t, int32_t* tm, int32_t* tn);

/* Measurement helper (tools/time_leaf_phases.py): while `buf` (>= 16 int64, device memory) is set, every leaf-Cholesky launch
 * records clock64() at its phase boundaries there; NULL switches it off.  Synchronises the device. */
int gpk_debug_leaf_phase_clock(void* buf16_int64);

/* Number of kernels this library has launched since load / the last reset (bench.py's `gpu_launches`). */
int64_t gpk_launch_count(void);
void gpk_launch_count_reset(void);

#ifdef __cplusplus
}
#endif
#endif /* GPK_H_ */
