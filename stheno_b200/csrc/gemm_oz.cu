// fp64 GEMM  C = beta*C + alpha * A * B^T  emulated on the 5th-generation tensor cores' INT8 path
// (tcgen05.mma.kind::i8, int32 accumulators in TMEM) with the Ozaki splitting: every row of A and B is scaled by a power
// of two (its largest magnitude) and cut into S signed 7-bit slices
//        a = 2^e * sum_s  q_s 2^-(6+7s),   q_s = round-to-nearest of the running remainder, |q_s| <= 64   (error-free)
// so that   a . b = 2^(e_a+e_b) * sum_{s,t} 2^-(12+7(s+t)) * (q_s . r_t),   and every integer dot product q_s . r_t is
// EXACT in int32 (|q r| <= 2^12, K <= 65536, up to S products per accumulator).  Products with s + t >= S are dropped
// (they are below the slicing error), leaving S(S+1)/2 int8 GEMMs -- 21 for S = 6 (error ~2^-40 of |a||b|, zero-mean), 28
// for S = 7 (~2^-47) -- against the 8192 MAC/clk/SM of the int8 pipe instead of the 64 FMA/clk/SM of DMMA.
//
// 128 x 64 output tiles, a few consecutive tiles per CTA, enumerated in bands of tile rows, column-major inside a band (tiles
// in flight share the band's A slices in L2; oz_tile); products with the same s + t share an accumulator: S accumulators
// x 64 TMEM columns.  The producer and the issuer run ahead into the next tile while the epilogue warps finish the previous
// one (TMEM full / empty mbarriers); only the accumulator drain itself (S x 64 columns at the 64 B/clk TMEM read rate,
// ~1.9 us of 11.7 us per tile for S = 7) is serial.
//   warp 0      TMA producer: per 64-byte k-block ONE box {64 B, 128 rows, S slices} of A and one {64 B, 64 rows, S} of B
//               (cp.async.bulk.tensor.3d, SWIZZLE_64B, mbarrier complete_tx) into a 2-stage ring
//   warp 1      MMA issuer: per K = 32 step slice s of A against slices 0..S-1-s of B as ONE tcgen05.mma.kind::i8 with
//               N = 64 (S - s) (B slices adjacent in shared memory, accumulators adjacent in TMEM), tcgen05.commit per stage
//   warps 2-5   epilogue: tcgen05.ld the S accumulators 16 columns at a time, Horner-combine them in fp64 (exact int32 ->
//               double through the 2^52 trick), scale by 2^(e_row + e_col), stage 128 x 16 boxes in shared memory and add
//               them into C with TMA tensor reduces (cp.reduce.async.bulk.tensor .add, fp64 tensor map: the
//               read-modify-write of C happens in L2, not in the SM)
// The slicing kernel (oz_slice_kernel) is O(rows*K) and runs once per panel; in the Cholesky its output is shared by
// every tile of the trailing update.  Every inexact step is ONE correctly rounded fp64 operation, so the result equals a
// NumPy integer model of the algorithm bit for bit (tests/_oz_model.py, tests/test_emulation.py).
//
// SASS evidence: UTCIMMA / UTCHMMA (tcgen05.mma), UTMALDG (TMA), LDTM (tcgen05.ld), UTCBAR (tcgen05.commit).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace gpk {
bool prof_enabled();  // gemm.cu: in-situ event profile (bench.py's roofline)
void prof_begin(cudaStream_t, double flops, int kind);
void prof_end(cudaStream_t);

namespace {

constexpr int OZ_BM = 128, OZ_BN = 64, OZ_BK = 64;  // BK in bytes = int8 elements
constexpr int OZ_THREADS = 192;

template <int S>
struct OzCfg {
  static constexpr int A_SLICE = OZ_BM * OZ_BK;  // 8 KB
  static constexpr int B_SLICE = OZ_BN * OZ_BK;  // 4 KB
  static constexpr int A_BYTES = S * A_SLICE;
  static constexpr int B_BYTES = S * B_SLICE;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = 2;
  // epilogue staging: one 128 x 16 fp64 box (16 KB, SWIZZLE_128B so that row-per-thread 16-byte stores are conflict-free)
  // per TMA tensor reduce; ping-pong when it fits next to the ring
  static constexpr int OUT_BYTES = OZ_BM * 128;
  static constexpr int SMEM_MAX = 226 * 1024;  // 227 KB per CTA minus the static barriers / alignment slack
  static constexpr int OUT_BUFS = (STAGES * STAGE_BYTES + 2 * OUT_BYTES + 1024 <= SMEM_MAX) ? 2 : 1;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + OUT_BUFS * OUT_BYTES + 1024;
  static constexpr int TMEM_COLS = 512;  // S * 64 rounded up to a power of two (S = 5..8)
};

struct OzParams {
  double alpha;
  double* C;
  const double* sc_a;  // 2^e per A row (already offset to the first row of the problem)
  const double* sc_b;
  int64_t ldc;
  int32_t a_row0, b_row0;  // first plane row of A / B
  int32_t KB, lower, tiles_m, tiles_n;
  int32_t total_tiles, tiles_per_cta, tri_rows;  // tri_rows: tile rows in the triangular part (lower mode)
  int32_t accumulate;                            // 1: C += alpha A B^T (reduce-add), 0: C = alpha A B^T (store)
  int32_t band;                                  // tile rows per band of the tile order
};

__device__ __forceinline__ void oz_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "OZW_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra OZW_DONE;\n"
      "bra OZW_LOOP;\n"
      "OZW_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void oz_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void oz_tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}
// K-major SWIZZLE_64B shared-memory matrix descriptor: start address >> 4, LBO = 1 (ignored for swizzled K-major),
// SBO = 512 B (8 rows x 64 B) >> 4 = 32, version 1 (Blackwell), layout type 4 (SWIZZLE_64B).
__device__ __forceinline__ uint64_t oz_desc(uint32_t smem_addr) {
  const uint64_t hi = (uint64_t)(32u | (1u << 14) | (4u << 29)) << 32;
  return hi | (uint64_t)(((smem_addr >> 4) & 0x3FFFu) | (1u << 16));
}
__device__ __forceinline__ void oz_umma_i8(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n"
      "}\n" ::"r"(tmem_c),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0));
}
__device__ __forceinline__ void oz_umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void oz_tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
// exact int32 -> double without the conversion pipe: bits(2^52 + 2^31 + x) = 0x43300000 : (x ^ 0x80000000)
__device__ __forceinline__ double oz_i2d(uint32_t x) {
  return __hiloint2double(0x43300000, (int)(x ^ 0x80000000u)) - 4503601774854144.0;  // 2^52 + 2^31
}
// tile index -> (tile row, tile column).  Lower mode enumerates only the tiles that touch the lower triangle: row tm holds
// nc(tm) = min(tiles_n, 2 (tm + 1)) tiles (128 x 64 tiles), so f(r) = r (r + 1) tiles precede row r while r <= tri_rows.
// Order: BANDS of `band` (16) tile rows, column-major inside a band.  Tiles that run at the same time then share the band's A rows
// (16 x 128 rows of slices: 15 MB at K = 1024, S = 7) and sweep the B rows once per band, instead of once per tile ROW
// as the plain row-major order did -- whose K = 1024 launches read 2.1x their algorithmic bytes from HBM (ncu, round 2: the
// 104 MB of slices no longer fit L2 next to the C traffic).
__host__ __device__ __forceinline__ int oz_rows_before(const OzParams& p, int r) {  // tiles in tile rows < r
  if (!p.lower) return r * p.tiles_n;
  return r <= p.tri_rows ? r * (r + 1) : p.tri_rows * (p.tri_rows + 1) + (r - p.tri_rows) * p.tiles_n;
}
__host__ __device__ __forceinline__ void oz_tile(const OzParams& p, int t, int& tm, int& tn) {
  // (1) band
  int r0 = 0;
  while (r0 + p.band < p.tiles_m && oz_rows_before(p, r0 + p.band) <= t) r0 += p.band;
  const int r1 = (r0 + p.band < p.tiles_m) ? r0 + p.band : p.tiles_m;
  int u = t - oz_rows_before(p, r0);
  // (2) column-major inside the band: row tm has nc(tm) columns, non-decreasing in tm
  const int nc0 = (p.lower && 2 * (r0 + 1) < p.tiles_n) ? 2 * (r0 + 1) : p.tiles_n;  // columns that every row of the band has
  const int h = r1 - r0;
  if (u < nc0 * h) {
    tn = u / h;
    tm = r0 + (u - tn * h);
    return;
  }
  u -= nc0 * h;
  // ragged part (lower mode only): column c >= nc0 exists in the rows tm with 2 (tm + 1) > c, i.e. tm = c / 2 ... r1 - 1
  for (int c = nc0; c < p.tiles_n; ++c) {
    const int first = (c >> 1) > r0 ? (c >> 1) : r0;  // first row of the band that has column c (c < tiles_n is implied by u's range)
    const int cnt = r1 - first;
    if (u < cnt) {
      tn = c;
      tm = first + u;
      return;
    }
    u -= cnt;
  }
  tm = r1 - 1;  // not reached for t < total_tiles (bijection checked on the host for 1047 shapes); keeps the loop bounded
  tn = p.tiles_n - 1;
}

// One CTA works through `tiles_per_cta` consecutive tiles: the TMA producer and the MMA issuer run ahead into the next tile
// while the epilogue warps drain the accumulators of the previous one (TMEM full / empty barriers), and the update of C
// itself leaves the SM asynchronously (bulk reduce-add from the staging buffer), overlapping the next main loop.  CTAs stay
// short-lived (a few tiles) on purpose: the Cholesky's look-ahead stream needs SMs to come free every few tens of us.
template <int S>
__global__ void __launch_bounds__(OZ_THREADS, 1)
oz_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
               const __grid_constant__ CUtensorMap mapC, const OzParams p) {
  using Cfg = OzCfg<S>;
  constexpr int STAGES = Cfg::STAGES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t_begin = blockIdx.x * p.tiles_per_cta;
  const int t_end = min(t_begin + p.tiles_per_cta, p.total_tiles);

  extern __shared__ uint8_t oz_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(oz_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* out_smem = smem + STAGES * Cfg::STAGE_BYTES;
  __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar, tmem_empty_bar;
  __shared__ uint32_t tmem_base_holder;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_holder)),
                 "r"(Cfg::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::);
  }
  if (threadIdx.x == 32) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    mbar_init(&tmem_empty_bar, 128);
    fence_mbar_init();
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
  const uint32_t tmem = tmem_base_holder;
  const int KB = p.KB;

  if (warp == 0) {
    if (lane == 0) {
      int g = 0;  // k-blocks issued so far (all tiles)
      for (int t = t_begin; t < t_end; ++t) {
        int tm, tn;
        oz_tile(p, t, tm, tn);
        for (int kb = 0; kb < KB; ++kb, ++g) {
          const int s = g % STAGES, it = g / STAGES;
          if (it > 0) oz_mbar_wait(&empty_bar[s], (it - 1) & 1);
          uint8_t* st = smem + s * Cfg::STAGE_BYTES;
          mbar_arrive_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
          oz_tma_load_3d(st, &mapA, kb * OZ_BK, p.a_row0 + tm * OZ_BM, 0, &full_bar[s]);
          oz_tma_load_3d(st + Cfg::A_BYTES, &mapB, kb * OZ_BK, p.b_row0 + tn * OZ_BN, 0, &full_bar[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor: D = S32 (2 << 4), A = B = signed int8 (1 << 7, 1 << 10), K-major, N >> 3 at 17, M >> 4 at 24
      const uint32_t idesc0 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OZ_BM >> 4) << 24);
      int g = 0;
      for (int t = t_begin, ti = 0; t < t_end; ++t, ++ti) {
        if (ti > 0) {  // the epilogue must have drained the previous tile's accumulators
          oz_mbar_wait(&tmem_empty_bar, (ti - 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
        }
        for (int kb = 0; kb < KB; ++kb, ++g) {
          const int s = g % STAGES, it = g / STAGES;
          oz_mbar_wait(&full_bar[s], it & 1);
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
          const uint32_t a0 = smem_u32(smem + s * Cfg::STAGE_BYTES), b0 = a0 + Cfg::A_BYTES;
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {  // UMMA K = 32 int8 = 32 bytes inside the 64-byte swizzle row
            // Slice sa of A meets slices 0 .. S-1-sa of B, whose products go to the ADJACENT accumulators sa .. S-1: the B
            // slices are contiguous in shared memory (64 rows each) and the accumulators contiguous in TMEM (64 columns
            // each), so they are ONE MMA with N = 64 (S - sa) (split at the N = 256 limit) -- A is read from shared memory
            // 8 times per K step instead of 21 (S = 6), which takes the kernel off the shared-memory bandwidth limit.
#pragma unroll
            for (int sa = 0; sa < S; ++sa) {
#pragma unroll
              for (int sb = 0; sb + sa < S; sb += 4) {
                const int nsl = (S - sa - sb) < 4 ? (S - sa - sb) : 4;  // B slices in this MMA
                const uint32_t idesc = idesc0 | ((uint32_t)((nsl * OZ_BN) >> 3) << 17);
                oz_umma_i8(tmem + (uint32_t)((sa + sb) * OZ_BN), oz_desc(a0 + sa * Cfg::A_SLICE + ks * 32),
                           oz_desc(b0 + sb * Cfg::B_SLICE + ks * 32), idesc, (kb > 0 || ks > 0 || sa > 0) ? 1u : 0u);
              }
            }
          }
          oz_umma_commit(&empty_bar[s]);
        }
        oz_umma_commit(&tmem_full_bar);
      }
    }
  } else {
    // ---- epilogue: TMEM -> registers -> fp64 combine -> staging row in shared memory -> bulk reduce-add into C ----
    const int lane_group = warp & 3;  // a warp may only touch TMEM lanes 32 * (warp % 4) .. + 31
    const int row = lane_group * 32 + lane;
    // 2^-(12 + 7 (S-1)): weight of the last kept diagonal; Horner runs from diagonal 0 (largest weight) down
    const double w_last = __hiloint2double((1023 - (12 + 7 * (S - 1))) << 20, 0);
    const bool elected = threadIdx.x == 64;
    int chunk = 0;  // boxes staged so far
    for (int t = t_begin, ti = 0; t < t_end; ++t, ++ti) {
      int tm, tn;
      oz_tile(p, t, tm, tn);
      const double rs = p.alpha * w_last * __ldg(p.sc_a + (int64_t)tm * OZ_BM + row);
      const double* cs = p.sc_b + (int64_t)tn * OZ_BN;
      oz_mbar_wait(&tmem_full_bar, ti & 1);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::);
      // (1) drain: all 64 columns of this row, Horner-combined on the fly (diagonal 0 carries the largest weight), then hand
      //     TMEM back so that the next tile's MMAs overlap everything below
      double v[OZ_BN];
#pragma unroll
      for (int c = 0; c < OZ_BN / 16; ++c) {
        const uint32_t taddr = tmem + ((uint32_t)(lane_group * 32) << 16) + c * 16;
#pragma unroll
        for (int d0 = 0; d0 < S; d0 += 4) {
          constexpr int G = 4;
          uint32_t r[G][16];
#pragma unroll
          for (int d = d0; d < d0 + G && d < S; ++d) oz_tmem_ld16(taddr + d * OZ_BN, r[d - d0]);
          asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
          for (int d = d0; d < d0 + G && d < S; ++d) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              v[c * 16 + j] = (d == 0) ? oz_i2d(r[0][j]) : fma(v[c * 16 + j], 128.0, oz_i2d(r[d - d0][j]));
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
      oz_mbar_arrive(&tmem_empty_bar);
      // (2) scale and stage 16 columns of all 128 rows (one 16 KB box, 128-byte swizzle), then ONE thread adds the box into C
      //     with a TMA tensor reduce (cp.reduce.async.bulk.tensor .add on an fp64 tensor map: the read-modify-write of C
      //     happens in L2).  All of this overlaps the next tile's main loop.
#pragma unroll
      for (int c = 0; c < OZ_BN / 16; ++c, ++chunk) {
        uint8_t* buf = out_smem + (Cfg::OUT_BUFS == 2 ? (chunk & 1) * Cfg::OUT_BYTES : 0);
        if (elected) {  // the reduce that last used this buffer has finished reading it
          if (Cfg::OUT_BUFS == 2)
            asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory");
          else
            asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
        }
        asm volatile("bar.sync 1, 128;\n" ::: "memory");
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int col = c * 16 + 2 * j;
          double2 out;
          out.x = v[col] * (rs * __ldg(cs + col));
          out.y = v[col + 1] * (rs * __ldg(cs + col + 1));
          *reinterpret_cast<double2*>(buf + row * 128 + ((j ^ (row & 7)) << 4)) = out;
        }
        fence_proxy_async();  // generic-proxy writes -> visible to the TMA (async-proxy) read
        asm volatile("bar.sync 1, 128;\n" ::: "memory");
        if (elected) {
          if (p.accumulate)
            asm volatile(
                "cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2}], [%3];\n" ::"l"(&mapC),
                "r"(tn * OZ_BN + c * 16), "r"(tm * OZ_BM), "r"(smem_u32(buf))
                : "memory");
          else
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];\n" ::"l"(&mapC),
                         "r"(tn * OZ_BN + c * 16), "r"(tm * OZ_BM), "r"(smem_u32(buf))
                         : "memory");
          asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
        }
      }
    }
    asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::);
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(Cfg::TMEM_COLS));
}

// ---- slicing: one warp per row.  planes[s][row][k] (int8), sc[row] = 2^e -----------------------------------------------
template <int S>
__global__ void __launch_bounds__(256)
oz_slice_kernel(const double* __restrict__ P, int64_t ldp, int64_t rows, int32_t K, int8_t* __restrict__ planes,
                int64_t plane_stride, double* __restrict__ sc) {
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const double* src = P + row * ldp;
  double m = 0.0;
  for (int k = lane * 4; k < K; k += 128) {
    const double2 a = *reinterpret_cast<const double2*>(src + k), b = *reinterpret_cast<const double2*>(src + k + 2);
    m = fmax(fmax(fabs(a.x), fabs(a.y)), fmax(m, fmax(fabs(b.x), fabs(b.y))));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  int e = 0;
  if (m > 0.0 && m < 1e300) {  // (inf / nan rows: the factorisation reports them through `info` anyway)
    e = ilogb(m) + 1;
    e = max(-1000, min(1000, e));
  }
  const double inv = __hiloint2double((1023 - e) << 20, 0);  // 2^-e
  if (lane == 0) sc[row] = __hiloint2double((1023 + e) << 20, 0);
  int8_t* dst = planes + row * K;
  for (int k = lane * 4; k < K; k += 128) {
    const double2 a = *reinterpret_cast<const double2*>(src + k), b = *reinterpret_cast<const double2*>(src + k + 2);
    double r[4] = {a.x * inv, a.y * inv, b.x * inv, b.y * inv};
    double pw = 64.0, ipw = 1.0 / 64.0;
#pragma unroll
    for (int s = 0; s < S; ++s) {
      char4 q;
      double t;
      t = rint(r[0] * pw); r[0] = fma(-t, ipw, r[0]); q.x = (signed char)(int)t;
      t = rint(r[1] * pw); r[1] = fma(-t, ipw, r[1]); q.y = (signed char)(int)t;
      t = rint(r[2] * pw); r[2] = fma(-t, ipw, r[2]); q.z = (signed char)(int)t;
      t = rint(r[3] * pw); r[3] = fma(-t, ipw, r[3]); q.w = (signed char)(int)t;
      *reinterpret_cast<char4*>(dst + (int64_t)s * plane_stride + k) = q;
      pw *= 128.0;
      ipw *= 1.0 / 128.0;
    }
  }
}

typedef CUresult (*OzEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
OzEncodeTiledFn oz_encode_fn() {
  static OzEncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess)
      return (OzEncodeTiledFn) nullptr;
    return (OzEncodeTiledFn)p;
  }();
  return fn;
}

bool oz_make_map(CUtensorMap* m, const int8_t* planes, int64_t K, int64_t rows_cap, int64_t plane_stride, int S,
                 int box_rows) {
  OzEncodeTiledFn enc = oz_encode_fn();
  if (!enc) return false;
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows_cap, (cuuint64_t)S};
  cuuint64_t strides[2] = {(cuuint64_t)K, (cuuint64_t)plane_stride};
  cuuint32_t box[3] = {(cuuint32_t)OZ_BK, (cuuint32_t)box_rows, (cuuint32_t)S};
  cuuint32_t es[3] = {1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<int8_t*>(planes), dims, strides, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int S>
int oz_launch_slice(const double* P, int64_t ldp, int64_t rows, int64_t K, int8_t* planes, int64_t plane_stride, double* sc,
                    cudaStream_t stream) {
  if (rows == 0) return 0;
  oz_slice_kernel<S><<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(P, ldp, rows, (int32_t)K, planes, plane_stride, sc);
  GPK_COUNT_LAUNCH();
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -1000 - (int)e;
}

__global__ void oz_scale_kernel(double* C, int64_t ldc, int64_t M, int64_t N, double beta) {
  const int64_t i = (int64_t)blockIdx.y, j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < M && j < N) C[i * ldc + j] *= beta;
}

template <int S>
int oz_launch_gemm(int64_t M, int64_t N, int64_t K, double alpha, const int8_t* planesA, int64_t capA, int64_t strideA,
                   const double* scA, int64_t rowA, const int8_t* planesB, int64_t capB, int64_t strideB,
                   const double* scB, int64_t rowB, double beta, double* C, int64_t ldc, int32_t lower,
                   cudaStream_t stream) {
  CUtensorMap mA, mB, mC;
  if (!oz_make_map(&mA, planesA, K, capA, strideA, S, OZ_BM) || !oz_make_map(&mB, planesB, K, capB, strideB, S, OZ_BN))
    return GPK_ERR_UNSUPPORTED;
  {
    cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M};
    cuuint64_t strides[1] = {(cuuint64_t)ldc * 8};
    cuuint32_t box[2] = {16, (cuuint32_t)OZ_BM}, es[2] = {1, 1};
    if (oz_encode_fn()(&mC, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, C, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return GPK_ERR_UNSUPPORTED;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e =
        cudaFuncSetAttribute(oz_gemm_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, OzCfg<S>::SMEM_BYTES);
    if (e != cudaSuccess) return -1000 - (int)e;
    attr_set = true;
  }
  if (beta != 0.0 && beta != 1.0) {  // the kernel adds into C (or overwrites it): apply any other beta first
    if (M > 65535) return GPK_ERR_UNSUPPORTED;
    oz_scale_kernel<<<dim3((unsigned)((N + 255) / 256), (unsigned)M), 256, 0, stream>>>(C, ldc, M, N, beta);
    GPK_COUNT_LAUNCH();
  }
  const int32_t tiles_m = (int32_t)(M / OZ_BM), tiles_n = (int32_t)(N / OZ_BN);
  const int32_t tri_rows = lower ? (tiles_m < tiles_n / 2 ? tiles_m : tiles_n / 2) : 0;
  const int32_t total = lower ? tri_rows * (tri_rows + 1) + (tiles_m - tri_rows) * tiles_n : tiles_m * tiles_n;
  static const int force_tpc = getenv("GPK_OZ_TPC") ? atoi(getenv("GPK_OZ_TPC")) : 0;
  int32_t tpc = force_tpc > 0 ? force_tpc : total / 296;  // >= 2 waves of CTAs over 148 SMs before CTAs grow
  static const int env_cap = getenv("GPK_OZ_TPC_CAP") ? atoi(getenv("GPK_OZ_TPC_CAP")) : 0;  // experiments
  static const int env_band = getenv("GPK_OZ_BAND") ? atoi(getenv("GPK_OZ_BAND")) : 0;
  // band height of the tile order: as many tile rows as keep the band's A slices (128 K S bytes per tile row) within ~24 MB
  // of L2, at most 16 (K = 1024, S = 7: 16 rows = 15 MB; the K = 8192 products of the triangular solves: 2-3 rows)
  int32_t band = (int32_t)((24ll << 20) / (128ll * K * S));
  band = env_band > 0 ? env_band : (band < 1 ? 1 : (band > 16 ? 16 : band));
  const int32_t tpc_cap = env_cap > 0 ? env_cap : (K > 512 ? 2 : 4);  // CTAs stay short-lived (look-ahead streams need SMs every few tens of us)
  tpc = tpc < 1 ? 1 : (tpc > tpc_cap && force_tpc <= 0 ? tpc_cap : tpc);
  OzParams p{alpha, C, scA + rowA, scB + rowB, ldc, (int32_t)rowA, (int32_t)rowB, (int32_t)(K / OZ_BK), lower,
             tiles_m, tiles_n, total, tpc, tri_rows, beta != 0.0 ? 1 : 0, band};
  // profile: algorithmic (fp64-equivalent) flops of the tiles computed; the int8 work is S (S + 1) / 2 times that
  if (prof_enabled()) prof_begin(stream, (double)total * 2.0 * OZ_BM * OZ_BN * (double)K, 1);
  oz_gemm_kernel<S><<<(unsigned)((total + tpc - 1) / tpc), OZ_THREADS, OzCfg<S>::SMEM_BYTES, stream>>>(mA, mB, mC, p);
  if (prof_enabled()) prof_end(stream);
  GPK_COUNT_LAUNCH();
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -1000 - (int)e;
}

}  // namespace

// ---- entry points used by potrf.cu and the C-ABI ----------------------------------------------------------------------
// workspace for `rows` rows of K columns: S int8 planes [S][rows][K] followed by `rows` doubles (the row scales)
int64_t oz_ws_bytes(int64_t rows, int64_t K, int32_t S) { return (int64_t)S * rows * K + rows * 8 + 256; }

static inline double* oz_scales(void* ws, int64_t rows, int64_t K, int32_t S) {
  const uintptr_t p = reinterpret_cast<uintptr_t>(ws) + (uintptr_t)((int64_t)S * rows * K);
  return reinterpret_cast<double*>((p + 255) & ~uintptr_t(255));
}

int oz_slice_panel(const double* P, int64_t ldp, int64_t rows, int64_t K, void* ws, int64_t cap_rows, int32_t S,
                   cudaStream_t stream) {
  if (K % 128 || K > 65536 || rows > cap_rows || ldp % 2 || reinterpret_cast<uintptr_t>(P) % 16) return GPK_ERR_ARG;
  int8_t* planes = static_cast<int8_t*>(ws);
  double* sc = oz_scales(ws, cap_rows, K, S);
  switch (S) {
    case 5: return oz_launch_slice<5>(P, ldp, rows, K, planes, cap_rows * K, sc, stream);
    case 6: return oz_launch_slice<6>(P, ldp, rows, K, planes, cap_rows * K, sc, stream);
    case 7: return oz_launch_slice<7>(P, ldp, rows, K, planes, cap_rows * K, sc, stream);
    case 8: return oz_launch_slice<8>(P, ldp, rows, K, planes, cap_rows * K, sc, stream);
  }
  return GPK_ERR_ARG;
}

// C[M x N] = beta C + alpha A B^T with A = sliced rows [rowA, rowA + M) of wsA, B = sliced rows [rowB, rowB + N) of wsB
int oz_gemm_sliced(int64_t M, int64_t N, int64_t K, double alpha, const void* wsA, int64_t capA, int64_t rowA,
                   const void* wsB, int64_t capB, int64_t rowB, double beta, double* C, int64_t ldc, int32_t lower,
                   int32_t S, cudaStream_t stream) {
  if (M % OZ_BM || N % OZ_BN || K % 128 || K > 65536 || ldc % 2 || reinterpret_cast<uintptr_t>(C) % 16) return GPK_ERR_ARG;
  if (M == 0 || N == 0) return 0;
  const int8_t* pa = static_cast<const int8_t*>(wsA);
  const int8_t* pb = static_cast<const int8_t*>(wsB);
  const double* sa = oz_scales(const_cast<void*>(wsA), capA, K, S);
  const double* sb = oz_scales(const_cast<void*>(wsB), capB, K, S);
#define OZ_CASE(SS)                                                                                                   \
  case SS:                                                                                                            \
    return oz_launch_gemm<SS>(M, N, K, alpha, pa, capA, capA * K, sa, rowA, pb, capB, capB * K, sb, rowB, beta, C, ldc, \
                              lower, stream);
  switch (S) {
    OZ_CASE(5)
    OZ_CASE(6)
    OZ_CASE(7)
    OZ_CASE(8)
  }
#undef OZ_CASE
  return GPK_ERR_ARG;
}

// ---- library-wide fp64 emulation mode (like cublasSetMathMode): set per host thread and device by the caller, who also
// owns the scratch buffer.  Work that uses the scratch must be stream-ordered (one stream at a time per host thread). ----
Emulation& emulation() {
  static thread_local Emulation em[64];
  int dev = 0;
  cudaGetDevice(&dev);
  return em[(dev < 0 || dev >= 64) ? 0 : dev];
}

static inline int64_t round_up_1k(int64_t x) { return (x + 1023) & ~int64_t(1023); }

int64_t oz_gemm_scratch_bytes(int64_t M, int64_t N, int64_t K, int32_t S) {
  return round_up_1k(oz_ws_bytes(M, K, S)) + oz_ws_bytes(N, K, S);
}

// 1 = done on the emulated path, 0 = not applicable (caller uses DMMA), < 0 = error
int gemm_nt_f64_emulated(int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t lda, const double* B,
                         int64_t ldb, double beta, double* C, int64_t ldc, int32_t lower, cudaStream_t stream) {
  const Emulation& em = emulation();
  if (em.slices < 5 || em.slices > 8 || !em.scratch) return 0;
  if (M % OZ_BM || N % OZ_BN || K % 128 || lda % 2 || ldb % 2 || ldc % 2) return 0;
  if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(C)) % 16) return 0;
  // worth it only when the product dwarfs the slicing passes and fills the machine
  if (M < 256 || N < 256 || (double)M * (double)N * (double)K < 1.5e9) return 0;
  const bool same = (A == B && lda == ldb && M >= N);
  // int32 accumulation is exact for K <= 65536 per pass: longer reductions (the sparse path's n = 262144) run as several
  // passes over K chunks, each adding into C
  constexpr int64_t KC_MAX = 65536;
  const int64_t passes = (K + KC_MAX - 1) / KC_MAX;
  const int64_t kc = ((K / passes + 127) / 128) * 128;  // chunk length (multiple of 128), last chunk = remainder
  const int64_t off_b = same ? 0 : round_up_1k(oz_ws_bytes(M, kc, em.slices));
  if (em.bytes < off_b + oz_ws_bytes(same ? M : N, kc, em.slices)) return 0;
  int rc;
  void* wsa = em.scratch;
  void* wsb = static_cast<char*>(em.scratch) + off_b;
  for (int64_t k0 = 0; k0 < K; k0 += kc) {
    const int64_t kk = (K - k0 < kc) ? K - k0 : kc;
    if ((rc = oz_slice_panel(A + k0, lda, M, kk, wsa, M, em.slices, stream))) return rc;
    if (!same && (rc = oz_slice_panel(B + k0, ldb, N, kk, wsb, N, em.slices, stream))) return rc;
    if ((rc = oz_gemm_sliced(M, N, kk, alpha, wsa, M, 0, same ? wsa : wsb, same ? M : N, 0, k0 == 0 ? beta : 1.0, C, ldc, lower,
                             em.slices, stream)))
      return rc;
  }
  return 1;
}

}  // namespace gpk

extern "C" {

int gpk_set_f64_emulation(int32_t slices, void* scratch, int64_t scratch_bytes) {
  if (slices != 0 && (slices < 5 || slices > 8)) return GPK_ERR_ARG;
  if (slices != 0 && (!scratch || reinterpret_cast<uintptr_t>(scratch) % 1024 || scratch_bytes <= 0)) return GPK_ERR_ARG;
  gpk::Emulation& em = gpk::emulation();
  em.slices = slices;
  em.scratch = slices ? scratch : nullptr;
  em.bytes = slices ? scratch_bytes : 0;
  return 0;
}
int64_t gpk_f64_emulation_scratch_bytes(int64_t M, int64_t N, int64_t K, int32_t slices) {
  return gpk::oz_gemm_scratch_bytes(M, N, K, slices);
}

int64_t gpk_oz_ws_bytes(int64_t rows, int64_t K, int32_t slices) { return gpk::oz_ws_bytes(rows, K, slices); }

// Host-side evaluation of the emulation GEMM's tile order (the SAME function the kernel runs): tile index t of a launch with
// tiles_m x tiles_n tiles of 128 x 64 -> (tile row, tile column).  Returns the number of tiles of the launch.  Lets the
// CPU test-suite prove that the order is a bijection for every shape without a GPU (an unbounded / wrong order would hang
// or corrupt a launch on the device).
int32_t gpk_debug_oz_tile(int32_t lower, int32_t tiles_m, int32_t tiles_n, int32_t band, int32_t t, int32_t* tm, int32_t* tn) {
  gpk::OzParams p{};
  p.lower = lower;
  p.tiles_m = tiles_m;
  p.tiles_n = tiles_n;
  p.band = band;
  p.tri_rows = lower ? (tiles_m < tiles_n / 2 ? tiles_m : tiles_n / 2) : 0;
  p.total_tiles = lower ? p.tri_rows * (p.tri_rows + 1) + (tiles_m - p.tri_rows) * tiles_n : tiles_m * tiles_n;
  if (t >= 0 && t < p.total_tiles && tm && tn) {
    int a = 0, b = 0;
    gpk::oz_tile(p, t, a, b);
    *tm = a;
    *tn = b;
  }
  return p.total_tiles;
}

int gpk_gemm_nt_f64_oz(int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t lda, const double* B,
                       int64_t ldb, double beta, double* C, int64_t ldc, int32_t lower, int32_t slices, void* ws,
                       int64_t ws_bytes, void* stream) {
  if (slices < 5 || slices > 8 || !ws) return GPK_ERR_ARG;
  const int64_t need_a = gpk::oz_ws_bytes(M, K, slices), need_b = gpk::oz_ws_bytes(N, K, slices);
  const bool same = (A == B && lda == ldb && M >= N);
  const int64_t off_b = same ? 0 : ((need_a + 1023) & ~int64_t(1023));
  if (ws_bytes < off_b + (same ? need_a : need_b) || reinterpret_cast<uintptr_t>(ws) % 1024) return GPK_ERR_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  int rc;
  void* wsb = static_cast<char*>(ws) + off_b;
  if ((rc = gpk::oz_slice_panel(A, lda, M, K, ws, M, slices, s))) return rc;
  if (!same && (rc = gpk::oz_slice_panel(B, ldb, N, K, wsb, N, slices, s))) return rc;
  return gpk::oz_gemm_sliced(M, N, K, alpha, ws, M, 0, same ? ws : wsb, same ? M : N, 0, beta, C, ldc, lower, slices, s);
}

}  // extern "C"
