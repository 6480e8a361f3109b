// dmma_gemm.cuh — FP64 tensor-core (DMMA.8x8x4) GEMM engine for sm_100a.
//
// tcgen05 has no f64 kind, so FP64 tensor math is warp-level
// mma.sync.m8n8k4 with register accumulators and shared-memory staged operands:
// a cp.async (LDGSTS, 16-byte, zero-filling) multi-stage pipeline feeds padded,
// bank-conflict-free tiles.  Everything is column-major (R layout).
//
//   C[m,n] (op)= alpha * sum_k A[m,k] * B[k,n]
//
// Operand layouts (template flags) describe which index is contiguous in global memory:
//   A_KC = false : A[m,k] at A[m + k*lda]   ("m-contiguous", a column-major M x K matrix)
//   A_KC = true  : A[m,k] at A[k + m*lda]   (the transpose of a column-major K x M matrix)
//   B_KC = false : B[k,n] at B[n + k*ldb]   (the transpose of a column-major N x K matrix)
//   B_KC = true  : B[k,n] at B[k + n*ldb]   (a column-major K x N matrix)
// Shared-memory tiles keep the global contiguous index contiguous (so the 16-byte async
// copies are legal) and pad the leading dimension to == 4 (mod 16) doubles, which makes the
// per-thread 8-byte fragment loads of a half-warp hit 16 distinct 8-byte banks.
//
// MMA mapping: the instruction computes D(8x8) = Amma(8x4,row) * Bmma(4x8,col).  We put our
// n index on the MMA row and our m index on the MMA column, so that the two accumulator
// values a thread owns are two *consecutive m* of one column n — a 16-byte vector in the
// column-major C.  lane = 4*g + t:  Amma(row g, k t) <- B[k0+t, n0+g];  Bmma(k t, col g) <-
// A[m0+g, k0+t];  acc[0/1] <-> C[m0 + 2t + {0,1}, n0 + g].
//
// Requirements: lda/ldb/ldc even, base pointers 16-byte aligned, tile origins even (they
// are multiples of the tile sizes).  Edges are handled by zero-filling loads and predicated
// stores; no padding of M, N, K is needed.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gps {

struct GemmParams {
  const double* A;
  const double* B;
  int M, N, K;
  int lda, ldb;
  long long strideA, strideB;   // per batch entry (blockIdx.z / ksplit)
  // tile skipping: with lower != 0 a tile is skipped when it lies entirely above the
  // diagonal  row + diag_off == col  (tile rows m0..m0+BM-1, tile cols n0..)
  int lower;
  int diag_off;
  // triangular operands: restrict the k range per tile (local coordinates)
  int klo_n;   // k >= n0          (B lower-triangular in (k,n): B[k,n] = 0 for k < n)
  int klo_m;   // k >= m0          (A upper-triangular in (m,k): A[m,k] = 0 for k < m)
  int khi_m;   // k <  m0 + BM     (A lower-triangular in (m,k): A[m,k] = 0 for k > m)
  int khi_n;   // k <  n0 + BN     (B upper-triangular in (k,n): B[k,n] = 0 for k > n)
  int ksplit;  // split-K factor (>= 1); gridDim.z = batch * nodes * ksplit
  // "nodes": several independent sub-problems of one matrix in a single launch (the merges of
  // one level of the triangular-inverse tree).  Node i adds i*nodeStride{A,B} to the operand
  // pointers; its row count is clipped to the matrix edge: M_i = min(M, clip_total - i*clip_step)
  // (nodes with M_i <= 0 exit) and, with k_eq_m, K_i = M_i.  With clip_n the clipped extent is N instead
  // (N_i = min(N, ...), k_eq_m then means K_i = N_i).
  int nodes;
  long long nodeStrideA, nodeStrideB;
  int clip_total, clip_step, k_eq_m, clip_n;
  // weighted product C = A diag(w) B (warp-specialised kernel only, template flag KS): w[k] per batch entry,
  // readable in whole 16-element groups up to K rounded up to 16 (zeros beyond K)
  const double* kscale;
  long long strideScale;
};

// per-node clipping of the problem extents (see GemmParams); false when the node has nothing to do
__device__ __forceinline__ bool apply_clip(GemmParams& p, int node) {
  if (p.clip_step) {
    const int lim = p.clip_total - node * p.clip_step;
    if (p.clip_n) {
      p.N = min(p.N, lim);
      if (p.k_eq_m) p.K = p.N;
    } else {
      p.M = min(p.M, lim);
      if (p.k_eq_m) p.K = p.M;
    }
  }
  return p.M > 0 && p.N > 0;
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a_mma, double b_mma) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a_mma), "d"(b_mma));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem_src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

template <int BM_, int BN_, int BK_, int WARPS_M_, int WARPS_N_, int STAGES_>
struct TileCfg {
  static constexpr int BM = BM_, BN = BN_, BK = BK_, WARPS_M = WARPS_M_, WARPS_N = WARPS_N_, STAGES = STAGES_;
  static constexpr int NT = WARPS_M * WARPS_N * 32;
  static constexpr int WM = BM / WARPS_M, WN = BN / WARPS_N;
  static constexpr int FM = WM / 8, FN = WN / 8;
  static_assert(BM % (8 * WARPS_M) == 0 && BN % (8 * WARPS_N) == 0, "warp tiling");
  static_assert(BK % 4 == 0 && BM % BK == 0 && BK % 16 == 0, "k tiling");
  static __device__ __forceinline__ void sync() { __syncthreads(); }   // all threads that run the epilogue
};

// Accumulator view handed to epilogues: acc[mi][ni][0..1] <-> element
// (m = m_base + 8*mi + {0,1}, n = n_base + 8*ni) where m_base/n_base already include the
// per-thread offsets (2t and g).
template <class Cfg>
struct AccTile {
  double v[Cfg::FM][Cfg::FN][2];
  int m_base, n_base;      // global (tile + warp + lane) coordinates of v[0][0][0]
  int m0, n0;              // tile origin
  int batch, split, node;
  bool skip;               // this warp's sub-tile lies entirely above the diagonal of a lower-triangular product's diagonal
                           // tile: its accumulators were not computed (zeros) and nothing of it may be stored
};

// Epilogues that tolerate (and honour) AccTile::skip declare `static constexpr bool kSkipUpper = true`.
template <class Epi, class = void>
struct epi_skips_upper { static constexpr bool value = false; };
template <class Epi>
struct epi_skips_upper<Epi, decltype((void)Epi::kSkipUpper)> { static constexpr bool value = Epi::kSkipUpper; };

// Epilogues that never touch the shared scratch area declare `static constexpr bool kNoScratch = true`: the persistent
// kernel then needs no barrier among its consumer warps between tiles.
template <class Epi, class = void>
struct epi_no_scratch { static constexpr bool value = false; };
template <class Epi>
struct epi_no_scratch<Epi, decltype((void)Epi::kNoScratch)> { static constexpr bool value = Epi::kNoScratch; };

template <class Cfg, bool A_KC, bool B_KC>
struct SmemLayout {
  static constexpr int LDA = (A_KC ? Cfg::BK : Cfg::BM) + 4;
  static constexpr int LDB = (B_KC ? Cfg::BK : Cfg::BN) + 4;
  static constexpr int A_STAGE = (A_KC ? Cfg::BM : Cfg::BK) * LDA;
  static constexpr int B_STAGE = (B_KC ? Cfg::BN : Cfg::BK) * LDB;
  static constexpr size_t BYTES = (size_t)Cfg::STAGES * (A_STAGE + B_STAGE) * sizeof(double);
};

// Load one operand tile (R rows of the "other" index x BK of k) into a stage.
//   KC == false: tile is BK "rows" of k, each CONT contiguous elements of the m/n index
//   KC == true : tile is CONT "rows" of the m/n index, each BK contiguous elements of k
template <int CONT, int BK, int LDS, int NT, bool KC>
__device__ __forceinline__ void load_tile(double* __restrict__ s, const double* __restrict__ g, int ld,
                                          int x0, int X, int kbase, int K, int tid) {
  if (!KC) {
    constexpr int CH = CONT / 2;   // 16-byte chunks per k row
#pragma unroll
    for (int c = tid; c < BK * CH; c += NT) {
      int kk = c / CH, xc = (c % CH) * 2;
      int gx = x0 + xc, gk = kbase + kk;
      int valid = (gk < K) ? min(max(X - gx, 0), 2) : 0;
      const double* src = valid ? (g + (size_t)gx + (size_t)gk * ld) : g;
      cp_async16(s + kk * LDS + xc, src, valid * 8);
    }
  } else {
    constexpr int CH = BK / 2;
#pragma unroll
    for (int c = tid; c < CONT * CH; c += NT) {
      int xx = c / CH, kc = (c % CH) * 2;
      int gx = x0 + xx, gk = kbase + kc;
      int valid = (gx < X) ? min(max(K - gk, 0), 2) : 0;
      const double* src = valid ? (g + (size_t)gk + (size_t)gx * ld) : g;
      cp_async16(s + xx * LDS + kc, src, valid * 8);
    }
  }
}

// The kernel.  Epi must provide:
//   __device__ void operator()(const GemmParams&, AccTile<Cfg>&, double* smem_scratch) const
// called by every thread of the CTA after the main loop (smem is free to reuse after a
// __syncthreads()).
template <class Cfg, bool A_KC, bool B_KC, class Epi>
__global__ void __launch_bounds__(Cfg::NT) dmma_gemm_kernel(GemmParams p_in, Epi epi) {
  GemmParams p = p_in;
  using SL = SmemLayout<Cfg, A_KC, B_KC>;
  constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, STAGES = Cfg::STAGES, NT = Cfg::NT;
  constexpr int FM = Cfg::FM, FN = Cfg::FN;
  extern __shared__ __align__(16) double smem[];
  double* sA = smem;
  double* sB = smem + STAGES * SL::A_STAGE;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int wm0 = (warp % Cfg::WARPS_M) * Cfg::WM;
  const int wn0 = (warp / Cfg::WARPS_M) * Cfg::WN;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int zz = blockIdx.z / p.ksplit, split = blockIdx.z % p.ksplit;
  const int batch = zz / p.nodes, node = zz % p.nodes;

  if (!apply_clip(p, node)) return;
  if (m0 >= p.M || n0 >= p.N) return;
  if (p.lower && (m0 + BM - 1 + p.diag_off < n0)) return;

  const double* __restrict__ A = p.A + (size_t)batch * p.strideA + (size_t)node * p.nodeStrideA;
  const double* __restrict__ B = p.B + (size_t)batch * p.strideB + (size_t)node * p.nodeStrideB;

  int k_lo = 0, k_hi = p.K;
  if (p.klo_n) k_lo = max(k_lo, n0);
  if (p.klo_m) k_lo = max(k_lo, m0);
  if (p.khi_m) k_hi = min(k_hi, m0 + BM);
  if (p.khi_n) k_hi = min(k_hi, n0 + BN);
  if (p.ksplit > 1) {
    int kc = ((p.K + BK - 1) / BK + p.ksplit - 1) / p.ksplit * BK;
    k_lo = max(k_lo, split * kc);
    k_hi = min(k_hi, split * kc + kc);
  }
  k_lo = (k_lo / BK) * BK;
  const int nkt = (k_hi > k_lo) ? (k_hi - k_lo + BK - 1) / BK : 0;

  AccTile<Cfg> acc;
#pragma unroll
  for (int i = 0; i < FM; i++)
#pragma unroll
    for (int j = 0; j < FN; j++) acc.v[i][j][0] = acc.v[i][j][1] = 0.0;

  // diagonal tile of a lower-triangular product: a warp whose sub-tile is entirely above the diagonal has nothing to do
  const bool skip = epi_skips_upper<Epi>::value && p.lower && p.diag_off == 0 && m0 == n0 && (wm0 + Cfg::WM - 1 < wn0);

  auto issue = [&](int stage, int kt) {
    const int kbase = k_lo + kt * BK;
    load_tile<BM, BK, SL::LDA, NT, A_KC>(sA + stage * SL::A_STAGE, A, p.lda, m0, p.M, kbase, p.K, tid);
    load_tile<BN, BK, SL::LDB, NT, B_KC>(sB + stage * SL::B_STAGE, B, p.ldb, n0, p.N, kbase, p.K, tid);
  };

#pragma unroll
  for (int s = 0; s < STAGES - 1; s++) {
    if (s < nkt) issue(s, s);
    cp_async_commit();
  }

  for (int kt = 0; kt < nkt; kt++) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      int nx = kt + STAGES - 1;
      if (nx < nkt) issue(nx % STAGES, nx);
      cp_async_commit();
    }
    const double* a_s = sA + (kt % STAGES) * SL::A_STAGE;
    const double* b_s = sB + (kt % STAGES) * SL::B_STAGE;
    if (skip) continue;
#pragma unroll
    for (int ks = 0; ks < BK / 4; ks++) {
      double af[FM], bf[FN];
#pragma unroll
      for (int i = 0; i < FM; i++) {
        int m = wm0 + 8 * i + g, k = 4 * ks + t;
        af[i] = A_KC ? a_s[m * SL::LDA + k] : a_s[k * SL::LDA + m];
      }
#pragma unroll
      for (int j = 0; j < FN; j++) {
        int n = wn0 + 8 * j + g, k = 4 * ks + t;
        bf[j] = B_KC ? b_s[n * SL::LDB + k] : b_s[k * SL::LDB + n];
      }
#pragma unroll
      for (int i = 0; i < FM; i++)
#pragma unroll
        for (int j = 0; j < FN; j++) dmma884(acc.v[i][j][0], acc.v[i][j][1], bf[j], af[i]);
    }
  }
  cp_async_wait<0>();
  __syncthreads();   // smem reusable by the epilogue

  acc.m0 = m0;
  acc.n0 = n0;
  acc.m_base = m0 + wm0 + 2 * t;
  acc.n_base = n0 + wn0 + g;
  acc.batch = batch;
  acc.split = split;
  acc.node = node;
  acc.skip = skip;
  epi(p, acc, smem);
}

// ---------------------------------------------------------------------------------------
// Epilogues
// ---------------------------------------------------------------------------------------

// C = alpha*acc + beta*C.  With ksplit > 1 the split index selects a workspace slice
// (strideSplit) and beta must be 0.
struct EpiPlain {
  double* C;
  int ldc;
  long long strideC;       // per batch
  long long strideSplit;   // per split slice (0 if ksplit == 1)
  long long nodeStrideC;   // per node (0 if nodes == 1)
  double alpha, beta;
  int lower_strict_mask;   // 1: do not write elements with row + diag_off < col (keeps upper triangle untouched)
  int diag_off;
  // optional second, TRANSPOSED copy of alpha*acc (no beta): Ct[n + m*ldct]; same batch stride as C
  double* Ct;
  int ldct;
  long long nodeStrideCt;
  // lower-triangular products (p.lower): what lies above the diagonal inside a diagonal tile is never read by any consumer
  // of the Gram slices / the Cholesky's trailing matrix / Kinv, so warps that own only such entries skip the tile
  static constexpr bool kSkipUpper = true;
  static constexpr bool kNoScratch = true;
  template <class Cfg>
  __device__ __forceinline__ void operator()(const GemmParams& p, AccTile<Cfg>& a, double*) const {
    if (a.skip) return;
    double* __restrict__ Cb = C + (size_t)a.batch * strideC + (size_t)a.split * strideSplit + (size_t)a.node * nodeStrideC;
    if (Ct) {
      double* __restrict__ Tb = Ct + (size_t)a.batch * strideC + (size_t)a.node * nodeStrideCt;
#pragma unroll
      for (int i = 0; i < Cfg::FM; i++) {
        int m = a.m_base + 8 * i;
#pragma unroll
        for (int j = 0; j < Cfg::FN; j++) {
          int n = a.n_base + 8 * j;
          if (n >= p.N) continue;
          if (m < p.M) Tb[(size_t)n + (size_t)m * ldct] = alpha * a.v[i][j][0];
          if (m + 1 < p.M) Tb[(size_t)n + (size_t)(m + 1) * ldct] = alpha * a.v[i][j][1];
        }
      }
    }
#pragma unroll
    for (int j = 0; j < Cfg::FN; j++) {
      int n = a.n_base + 8 * j;
      if (n >= p.N) continue;
#pragma unroll
      for (int i = 0; i < Cfg::FM; i++) {
        int m = a.m_base + 8 * i;
        if (m >= p.M) continue;
        double* c = Cb + (size_t)m + (size_t)n * ldc;
        bool w0 = !lower_strict_mask || (m + diag_off >= n);
        bool w1 = (m + 1 < p.M) && (!lower_strict_mask || (m + 1 + diag_off >= n));
        double r0 = alpha * a.v[i][j][0], r1 = alpha * a.v[i][j][1];
        if (w0 && w1) {
          double2 o = make_double2(0.0, 0.0);
          if (beta != 0.0) o = *reinterpret_cast<const double2*>(c);
          *reinterpret_cast<double2*>(c) = make_double2(r0 + beta * o.x, r1 + beta * o.y);
        } else {
          if (w0) c[0] = r0 + (beta != 0.0 ? beta * c[0] : 0.0);
          if (w1) c[1] = r1 + (beta != 0.0 ? beta * c[1] : 0.0);
        }
      }
    }
  }
};

// Covariance as a function of the scaled squared distance (CovFunc.R:33-39).  kind: 0 = SqExp / ARD
// sf2 exp(-r2/2); 1, 3, 5 = Matern nu = 1/2, 3/2, 5/2 in r = sqrt(r2) (extraParam of the reference).
__device__ __forceinline__ double cov_kernel_fn(int kind, double r2, double s2) {
  if (kind == 0) return s2 * exp(-0.5 * r2);
  const double r = sqrt(r2);
  if (kind == 1) return s2 * exp(-r);
  if (kind == 3) {
    const double a = 1.7320508075688772 * r;
    return s2 * (1.0 + a) * exp(-a);
  }
  const double a = 2.23606797749979 * r;
  return s2 * (1.0 + a + (5.0 * (r * r)) / 3.0) * exp(-a);
}

// (used by cov_finish_kernel only: the fused epilogue below stays lean)
// Kd = -sf2 f'(r)/r for K = sf2 f(r): dK_ij / dlog(ell_k) = Kd_ij (x_ik - x_jk)^2 / ell_k^2.  SqExp: Kd = K (never
// asked for).  Matern gradients are derived here — the reference has none (OptimisationGradientLog.R:6).
// At r = 0 the distance factor vanishes; nu = 1/2 (1/r) is defined as 0 there.
__device__ __forceinline__ double cov_deriv_fn(int kind, double r2, double s2) {
  if (kind == 0) return s2 * exp(-0.5 * r2);
  const double r = sqrt(r2);
  if (kind == 1) return r > 0.0 ? s2 * exp(-r) / r : 0.0;
  if (kind == 3) return 3.0 * s2 * exp(-1.7320508075688772 * r);
  return (5.0 / 3.0) * s2 * (1.0 + 2.23606797749979 * r) * exp(-2.23606797749979 * r);
}

// Covariance epilogue (CovFunc.R:33-39 on the Gram form):
//   r2 = max(0, sA[m] + sB[n] - 2*G[m,n]);  K = sf2 * exp(-r2/2);
// symmetric case (sym != 0): the diagonal is written as exactly sf2 and Ky additionally gets
// noise_diag[m] on the diagonal (ForOptimisationLog.R:35-40).  Hyper-parameters are read
// from device memory so that the launch is CUDA-graph static.
struct EpiCov {
  const double* normA;     // [batch][ldn] squared row norms of the scaled A rows
  const double* normB;
  const double* hyp;       // [batch][4]: sn2, sf2, sc2, -
  const double* noise_diag;   // [batch][ldn] or nullptr
  double* Kout;            // noise-free covariance (may be nullptr)
  double* Kyout;           // covariance + noise on the diagonal (may be nullptr)
  int ldk, ldn, ldnB;      // ldn / ldnB: per-batch strides of normA (and noise_diag) / normB
  long long strideK;
  int sym;
  int kind;                // cov_kernel_fn selector
  template <class Cfg>
  __device__ __forceinline__ void operator()(const GemmParams& p, AccTile<Cfg>& a, double*) const {
    const double s2 = hyp[a.batch * 4 + 1];
    const double* nA = normA + (size_t)a.batch * ldn;
    const double* nB = normB + (size_t)a.batch * ldnB;
    const double* nd = noise_diag ? noise_diag + (size_t)a.batch * ldn : nullptr;
    double* Kb = Kout ? Kout + (size_t)a.batch * strideK : nullptr;
    double* Kyb = Kyout ? Kyout + (size_t)a.batch * strideK : nullptr;
#pragma unroll
    for (int j = 0; j < Cfg::FN; j++) {
      int n = a.n_base + 8 * j;
      if (n >= p.N) continue;
      double sn = nB[n];
#pragma unroll
      for (int i = 0; i < Cfg::FM; i++) {
        int m = a.m_base + 8 * i;
        if (m >= p.M) continue;
        double kk[2], ky[2];
#pragma unroll
        for (int e = 0; e < 2; e++) {
          int mm = m + e;
          double r2 = (mm < p.M) ? fmax(0.0, nA[mm] + sn - 2.0 * a.v[i][j][e]) : 0.0;
          kk[e] = cov_kernel_fn(kind, r2, s2);
          ky[e] = kk[e];
          if (sym && mm == n) {
            kk[e] = s2;
            ky[e] = nd ? s2 + nd[mm] : s2;
          }
        }
        // the thread's two values are consecutive rows of one column: one 16-byte store per matrix
        size_t off = (size_t)m + (size_t)n * ldk;
        if (m + 1 < p.M) {
          if (Kb) *reinterpret_cast<double2*>(Kb + off) = make_double2(kk[0], kk[1]);
          if (Kyb) *reinterpret_cast<double2*>(Kyb + off) = make_double2(ky[0], ky[1]);
        } else {
          if (Kb) Kb[off] = kk[0];
          if (Kyb) Kyb[off] = ky[0];
        }
      }
    }
  }
};

// Gradient contraction epilogue (north_star item 4): P = M * Xaug is never written.
//   colpart[mtile][split][n] = sum_{m in tile} Xaug[m,n] * P[m,n]          (deterministic per-tile partials)
//   tpart[mtile][n]          = sum_{m in tile} rs[m] * Xaug[m,n]^2         (rs = rowsum(M), from wk_kernel; split 0 only)
// so that 2 rowsum(M) . x_k^2 - 2 x_k' M x_k needs no further pass over Xaug.
struct EpiColDot {
  const double* X;   // Xaug, column-major, ldx
  int ldx;
  long long strideX;
  double* colpart;   // [batch][mtiles * ksplit][ldp]   (slice = mtile * ksplit + split)
  int ldp, mtiles;
  const double* rs;  // [batch][ldn] rowsum(M)
  double* tpart;     // [batch][mtiles][ldp]
  int ldn;
  template <class Cfg>
  __device__ __forceinline__ void operator()(const GemmParams& p, AccTile<Cfg>& a, double* smem) const {
    const double* __restrict__ Xb = X + (size_t)a.batch * strideX;
    const double* __restrict__ rb = rs + (size_t)a.batch * ldn;
    const bool want_t = (a.split == 0);
    double rsm[Cfg::FM][2];
#pragma unroll
    for (int i = 0; i < Cfg::FM; i++) {
      int m = a.m_base + 8 * i;
      rsm[i][0] = (want_t && m < p.M) ? rb[m] : 0.0;
      rsm[i][1] = (want_t && m + 1 < p.M) ? rb[m + 1] : 0.0;
    }
    double part[Cfg::FN], tp[Cfg::FN];
#pragma unroll
    for (int j = 0; j < Cfg::FN; j++) {
      int n = a.n_base + 8 * j;
      double s = 0.0, ts = 0.0;
      if (n < p.N) {
#pragma unroll
        for (int i = 0; i < Cfg::FM; i++) {
          int m = a.m_base + 8 * i;
          if (m < p.M) {
            const double x0 = Xb[(size_t)m + (size_t)n * ldx];
            s += x0 * a.v[i][j][0];
            ts += rsm[i][0] * (x0 * x0);
          }
          if (m + 1 < p.M) {
            const double x1 = Xb[(size_t)m + 1 + (size_t)n * ldx];
            s += x1 * a.v[i][j][1];
            ts += rsm[i][1] * (x1 * x1);
          }
        }
      }
      // fixed-order reduction over the 4 lanes (t) that share column n
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      ts += __shfl_xor_sync(0xffffffffu, ts, 1);
      ts += __shfl_xor_sync(0xffffffffu, ts, 2);
      part[j] = s;
      tp[j] = ts;
    }
    // across the WARPS_M warps that share the same columns: shared memory, fixed order
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wmi = warp % Cfg::WARPS_M;
    const int wn0 = (warp / Cfg::WARPS_M) * Cfg::WN;
    double* red = smem;                              // [WARPS_M][BN]
    double* redt = smem + Cfg::WARPS_M * Cfg::BN;    // [WARPS_M][BN]
    if ((lane & 3) == 0) {
#pragma unroll
      for (int j = 0; j < Cfg::FN; j++) {
        red[wmi * Cfg::BN + wn0 + 8 * j + (lane >> 2)] = part[j];
        redt[wmi * Cfg::BN + wn0 + 8 * j + (lane >> 2)] = tp[j];
      }
    }
    Cfg::sync();
    for (int c = threadIdx.x; c < Cfg::BN; c += Cfg::NT) {
      int n = a.n0 + c;
      if (n < p.N) {
        double s = 0.0, ts = 0.0;
#pragma unroll
        for (int w = 0; w < Cfg::WARPS_M; w++) {
          s += red[w * Cfg::BN + c];
          ts += redt[w * Cfg::BN + c];
        }
        colpart[(((size_t)a.batch * mtiles + a.m0 / Cfg::BM) * p.ksplit + a.split) * ldp + n] = s;
        if (want_t) tpart[((size_t)a.batch * mtiles + a.m0 / Cfg::BM) * ldp + n] = ts;
      }
    }
  }
};

// W o K fused into the epilogue of Kinv = U U' (OptimisationGradientLog.R:51-53 in W-form; north_star item 4):
//   W = a a' - Kinv,  M = W o K,  wdiag = diag(W),  rowsum(M) as fixed-order per-column-block partials.
// Kinv itself is never written (it was only ever re-read to form M): one n x n write, one n x n read and one launch
// per evaluation less than the separate wk_kernel pass.  The product runs over LOWER tiles only (BM == BN):
//   * an off-diagonal tile (m0 > n0) holds M[m, n] for its rows/columns and, by symmetry, M[n, m]: both are stored,
//     its row sums go to rs_part[col block][m] and its column sums to rs_part[row block][n];
//   * a diagonal tile was computed in full (both halves): every entry is stored once, row sums only.
// Every (row block, column block) pair of rs_part is therefore written exactly once, without atomics.
struct EpiWK {
  const double* K;        // noise-free covariance, lower triangle valid, ld = ldk
  double* Mout;           // full symmetric M (both halves), same layout
  int ldk;
  long long strideK;
  const double* alpha;    // [batch][2 * ldn]: the right-hand side the reference's quadratic forms use
  double* wdiag;          // [batch][ldn]
  double* rs_part;        // [batch][nblk][ldn]
  int ldn, nblk, nvalid;  // nvalid = n (rows/columns >= n are identity padding and are skipped)
  static constexpr bool kSkipUpper = true;
  template <class Cfg>
  __device__ __forceinline__ void operator()(const GemmParams& p, AccTile<Cfg>& a, double* smem) const {
    // (lower-tile products only ever run on square tiles; the non-square configurations merely have to compile)
    const double* __restrict__ Kb = K + (size_t)a.batch * strideK;
    double* __restrict__ Mb = Mout + (size_t)a.batch * strideK;
    const double* __restrict__ al = alpha + (size_t)a.batch * 2 * ldn;
    const bool diag = (a.m0 == a.n0);
    double rsum[Cfg::FM][2], csum[Cfg::FN];
#pragma unroll
    for (int i = 0; i < Cfg::FM; i++) rsum[i][0] = rsum[i][1] = 0.0;
#pragma unroll
    for (int j = 0; j < Cfg::FN; j++) {
      const int n = a.n_base + 8 * j;
      const double an = (n < nvalid) ? al[n] : 0.0;
      double cs = 0.0;
#pragma unroll
      for (int i = 0; i < Cfg::FM; i++) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int m = a.m_base + 8 * i + e;
          if (m >= nvalid || n >= nvalid || m < n) continue;   // (m < n only inside a diagonal tile)
          const double w = al[m] * an - a.v[i][j][e];
          const double mv = w * Kb[(size_t)m + (size_t)n * ldk];
          Mb[(size_t)m + (size_t)n * ldk] = mv;
          rsum[i][e] += mv;
          if (m == n) {
            wdiag[(size_t)a.batch * ldn + m] = w;
          } else {
            Mb[(size_t)n + (size_t)m * ldk] = mv;
            cs += mv;
          }
        }
      }
      csum[j] = cs;
    }
    // row sums: over the 8 lanes (g) that hold the other columns of the same rows, then over the WARPS_N warps
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wmi = warp % Cfg::WARPS_M, wni = warp / Cfg::WARPS_M;
    double* rowp = smem;                               // [WARPS_N][BM]
    double* colp = smem + Cfg::WARPS_N * Cfg::BM;      // [WARPS_M][BN]
#pragma unroll
    for (int i = 0; i < Cfg::FM; i++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        double v = rsum[i][e];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        if ((lane >> 2) == 0) rowp[wni * Cfg::BM + wmi * Cfg::WM + 8 * i + 2 * (lane & 3) + e] = v;
      }
#pragma unroll
    for (int j = 0; j < Cfg::FN; j++) {
      double v = csum[j];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      if ((lane & 3) == 0) colp[wmi * Cfg::BN + wni * Cfg::WN + 8 * j + (lane >> 2)] = v;
    }
    Cfg::sync();
    double* rp = rs_part + (size_t)a.batch * nblk * ldn;
    const int ti = a.m0 / Cfg::BM, tj = a.n0 / Cfg::BN;
    for (int c = threadIdx.x; c < Cfg::BM; c += Cfg::NT) {
      const int m = a.m0 + c, n = a.n0 + c;
      double sr = 0.0, sc = 0.0;
#pragma unroll
      for (int w = 0; w < Cfg::WARPS_N; w++) sr += rowp[w * Cfg::BM + c];
#pragma unroll
      for (int w = 0; w < Cfg::WARPS_M; w++) sc += colp[w * Cfg::BN + c];
      if (diag) {
        if (m < nvalid) rp[(size_t)tj * ldn + m] = sr + sc;
      } else {
        if (m < nvalid) rp[(size_t)tj * ldn + m] = sr;
        if (n < nvalid) rp[(size_t)ti * ldn + n] = sc;
      }
    }
  }
};

// ---------------------------------------------------------------------------------------
// Launch helper
// ---------------------------------------------------------------------------------------
// Measured on B200 (profiles/README.md): the 4-warp 64 x 64 x 16 tile with a 3-stage pipeline (4 CTAs/SM)
// wins at every size, alone and batched (31.9 TF/s at 8192^3 = 90 % of cublasDgemm).  Variants that lost and
// were removed: 4 stages (30.5), 8 warps on the same tile, BK = 32 (31.0).  The 128 x 128 tile (29.9) is
// kept behind GPS_TILE=2 / GPS_LARGE_MIN_DIM for experiments only.
using CfgLarge = TileCfg<128, 128, 16, 2, 4, 4>;
using CfgS3 = TileCfg<64, 64, 16, 2, 2, 3>;
// covariance from very few features (K <= 32, e.g. C2's d = 20): the k-loop is two steps and the kernel is all epilogue
// (exp per element), latency-bound at the 16 warps/SM the 64 x 64 tile's 124 registers allow; 32 x 32 tiles of four
// 16 x 16 warps need ~60 registers and run 8+ CTAs per SM
using CfgTiny = TileCfg<32, 32, 16, 2, 2, 2>;

// Launch priority of the kernels issued by the calling host thread (0 = stream default).  The latency-bound kernels of
// the Cholesky's diagonal blocks are launched with the device's greatest priority: when another stream group's GEMM is
// draining, the block scheduler hands freed SM slots to them first.
inline int& launch_priority() {
  static thread_local int prio = 0;
  return prio;
}
template <class... KArgs, class... Args>
cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributePriority;
  at[0].val.priority = launch_priority();
  cfg.attrs = at;
  cfg.numAttrs = launch_priority() != 0 ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

template <class Cfg, bool A_KC, bool B_KC, class Epi>
cudaError_t launch_gemm(const GemmParams& p, const Epi& epi, int batch, cudaStream_t st) {
  using SL = SmemLayout<Cfg, A_KC, B_KC>;
  auto kern = dmma_gemm_kernel<Cfg, A_KC, B_KC, Epi>;
  static bool attr_set[64] = {false};   // per instantiation, per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SL::BYTES);
    if (e != cudaSuccess) return e;
    attr_set[dev & 63] = true;
  }
  if (p.M <= 0 || p.N <= 0) return cudaSuccess;
  dim3 grid((p.M + Cfg::BM - 1) / Cfg::BM, (p.N + Cfg::BN - 1) / Cfg::BN, batch * p.nodes * p.ksplit);
  return launch_kernel(kern, grid, dim3(Cfg::NT), SL::BYTES, st, p, epi);
}

}  // namespace gps
