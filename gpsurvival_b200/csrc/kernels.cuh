// kernels.cuh — the non-GEMM kernels of the GP hot path (all FP64, column-major, batched
// over blockIdx.z or blockIdx.y as noted).  Every hyper-parameter is read from device memory
// so that an evaluation is a static CUDA graph.  All reductions use a fixed order
// (per-thread strided serial sum, then a shared-memory tree) => run-to-run reproducible.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace gps {

constexpr int NB = 64;   // leaf block of the recursive Cholesky / triangular inverse

// Options mirrored from gpsurv.h (kept as ints on the device)
struct ModelOpts {
  int cov_form;     // 0 SqExp, 1 ARD
  int mean_form;    // 0 Zero, 1 Linear
  int noise_corr;   // 0 none, 1 learned (GPS2), 2 vec (GPS3)
  int ard_mode;     // 0 intended, 1 as written
};

// Block-wide deterministic sum (blockDim.x threads, power of two, <= 1024).
__device__ __forceinline__ double block_sum(double v, double* sh) {
  int tid = threadIdx.x;
  sh[tid] = v;
  __syncthreads();
  for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
    if (tid < s) sh[tid] += sh[tid + s];
    __syncthreads();
  }
  double r = sh[0];
  __syncthreads();
  return r;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------
// set_data helpers
// ---------------------------------------------------------------------------------------
// Column means of X (n x d, ldx); one warp per column.  grid (ceil(d/8), batch), block 256.
__global__ void colmean_kernel(const double* __restrict__ X, int n, int d, int ldx, long long strideX,
                               double* __restrict__ mu, int ldmu) {
  int col = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31, b = blockIdx.y;
  if (col >= d) return;
  const double* x = X + (size_t)b * strideX + (size_t)col * ldx;
  double s = 0.0;
  for (int i = lane; i < n; i += 32) s += x[i];
  s = warp_sum(s);
  if (lane == 0) mu[(size_t)b * ldmu + col] = s / (double)n;
}

// Xaug[:, k] = X[:, k] - mu[k] (k < d);  Xaug[:, d] = 1.   grid (ceil(n/256), d+1, batch)
__global__ void center_kernel(double* __restrict__ X, int n, int d, int ldx, long long strideX,
                              const double* __restrict__ mu, int ldmu) {
  int i = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y, b = blockIdx.z;
  if (i >= n) return;
  double* x = X + (size_t)b * strideX + (size_t)k * ldx + i;
  if (k < d) *x = *x - mu[(size_t)b * ldmu + k];
  else *x = 1.0;
}

// cens_rank[i] = index of row i among the censored rows (events == 0), else -1. One thread
// per batch entry (n is small enough; runs once per gps_set_data).
__global__ void cens_rank_kernel(const double* __restrict__ events, int n, int ldn, int* __restrict__ rank,
                                 int* __restrict__ ncens, int batch) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  int c = 0;
  for (int i = 0; i < n; i++) {
    bool cens = events[(size_t)b * ldn + i] == 0.0;
    rank[(size_t)b * ldn + i] = cens ? c : -1;
    c += cens ? 1 : 0;
  }
  ncens[b] = c;
}

// ---------------------------------------------------------------------------------------
// Per-evaluation hyper-parameter preparation (ForOptimisationLog.R:9-40).
//   par layout: [log sn2, log sf2, log ell(p), (w(d), b), (log sc2)]
// Outputs: hyp[b][0]=sn2, [1]=sf2, [2]=sc2 (GPS2) ; ell[b][p] ; meanw[b][d+1] (w, b' with
// b' = b + mu.w because the stored X is centred) ; noise_diag[b][i] = sn2 + corr_i.
// grid (batch), block 256.
// ---------------------------------------------------------------------------------------
__global__ void hyp_prep_kernel(const double* __restrict__ par, int len, ModelOpts o, int n, int d, int p,
                                const double* __restrict__ events, const int* __restrict__ cens_rank,
                                const double* __restrict__ log_sc2_vec, int ldc, int ldn,
                                const double* __restrict__ mu, int ldmu,
                                double* __restrict__ hyp, double* __restrict__ ell, int ldell,
                                double* __restrict__ meanw, double* __restrict__ noise_diag) {
  __shared__ double sh[256];
  int b = blockIdx.x, tid = threadIdx.x;
  const double* pr = par + (size_t)b * len;
  int plen = len - (o.noise_corr == 1 ? 1 : 0);
  double sn2 = exp(pr[0]), sf2 = exp(pr[1]);
  double sc2 = (o.noise_corr == 1) ? exp(pr[len - 1]) : 0.0;
  if (tid == 0) {
    hyp[b * 4 + 0] = sn2;
    hyp[b * 4 + 1] = sf2;
    hyp[b * 4 + 2] = sc2;
    hyp[b * 4 + 3] = 0.0;
  }
  for (int k = tid; k < p; k += blockDim.x) ell[(size_t)b * ldell + k] = exp(pr[2 + k]);
  // mean weights
  double dot = 0.0;
  if (o.mean_form == 1) {
    const double* w = pr + (plen - d - 1);
    for (int k = tid; k < d; k += blockDim.x) {
      meanw[(size_t)b * ldmu + k] = w[k];
      dot += w[k] * mu[(size_t)b * ldmu + k];
    }
    dot = block_sum(dot, sh);
    if (tid == 0) meanw[(size_t)b * ldmu + d] = w[d] + dot;
  } else {
    for (int k = tid; k <= d; k += blockDim.x) meanw[(size_t)b * ldmu + k] = 0.0;
  }
  for (int i = tid; i < n; i += blockDim.x) {
    double v = sn2;
    int r = cens_rank[(size_t)b * ldn + i];
    if (r >= 0) {
      if (o.noise_corr == 1) v += sc2;
      else if (o.noise_corr == 2) v += exp(log_sc2_vec[(size_t)b * ldc + r]);
    }
    noise_diag[(size_t)b * ldn + i] = v;
  }
}

// ---------------------------------------------------------------------------------------
// Pre-scale (CovFunc.R:25,27):  Z = X (/) ell, plus partial squared row norms.
//   intended : Z[i,k] = Xc[i,k] / ell[k]            (Xc centred; distances are invariant)
//   as written: Z[i,k] = (Xc[i,k] + mu[k]) / ell[(i + k*nrec) mod p]
// grid (ceil(n/128), ksplit, batch), block 128; thread = one row, serial over its k chunk.
// ---------------------------------------------------------------------------------------
__global__ void prescale_kernel(const double* __restrict__ X, int n, int d, int ldx, long long strideX,
                                const double* __restrict__ mu, int ldmu, const double* __restrict__ ell,
                                int ldell, int p, int as_written, int nrec, double* __restrict__ Z, int ldz,
                                long long strideZ, double* __restrict__ part, int ldn, int ksplit,
                                int ell_batch_stride_on, int mu_batch_stride_on) {
  int i = blockIdx.x * blockDim.x + threadIdx.x, ks = blockIdx.y, b = blockIdx.z;
  if (i >= n) return;
  int kc = (d + ksplit - 1) / ksplit;
  int k0 = ks * kc, k1 = min(d, k0 + kc);
  const double* x = X + (size_t)b * strideX + i;
  double* z = Z + (size_t)b * strideZ + i;
  const double* el = ell + (ell_batch_stride_on ? (size_t)b * ldell : 0);
  const double* m = mu + (mu_batch_stride_on ? (size_t)b * ldmu : 0);
  double acc = 0.0;
  for (int k = k0; k < k1; k++) {
    double xv = x[(size_t)k * ldx];
    double v;
    if (p == 1) v = xv / el[0];
    else if (!as_written) v = xv / el[k];
    else v = (xv + m[k]) / el[(int)(((long long)i + (long long)k * nrec) % p)];
    z[(size_t)k * ldz] = v;
    acc += v * v;
  }
  part[((size_t)b * ksplit + ks) * ldn + i] = acc;
}

// s[i] = sum_ks part[ks][i] (fixed order). grid (ceil(n/256), batch)
__global__ void rownorm_finish_kernel(const double* __restrict__ part, int n, int ldn, int ksplit,
                                      double* __restrict__ s) {
  int i = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (i >= n) return;
  double a = 0.0;
  for (int ks = 0; ks < ksplit; ks++) a += part[((size_t)b * ksplit + ks) * ldn + i];
  s[(size_t)b * ldn + i] = a;
}

// Split-K Gram finish: G = sum_s Gpart[s] (fixed order), then the SE epilogue (see EpiCov).
// grid (ceil(n1/32), ceil(n2/8), batch), block (32, 8).
__global__ void cov_finish_kernel(const double* __restrict__ Gpart, int n1, int n2, int ldg, long long strideSplit,
                                  long long strideG, int ksplit, const double* __restrict__ normA,
                                  const double* __restrict__ normB, int ldn, const double* __restrict__ hyp,
                                  const double* __restrict__ noise_diag, double* __restrict__ Kout,
                                  double* __restrict__ Kyout, int ldk, long long strideK, int sym) {
  int i = blockIdx.x * 32 + threadIdx.x, j = blockIdx.y * 8 + threadIdx.y, b = blockIdx.z;
  if (i >= n1 || j >= n2) return;
  if (sym && i < j) return;   // lower triangle only, like the fused path
  double gsum = 0.0;
  for (int s = 0; s < ksplit; s++)
    gsum += Gpart[(size_t)b * strideG + (size_t)s * strideSplit + (size_t)i + (size_t)j * ldg];
  double s2 = hyp[b * 4 + 1];
  double r2 = fmax(0.0, normA[(size_t)b * ldn + i] + normB[(size_t)b * ldn + j] - 2.0 * gsum);
  double k = s2 * exp(-0.5 * r2), ky = k;
  if (sym && i == j) {
    k = s2;
    ky = noise_diag ? s2 + noise_diag[(size_t)b * ldn + i] : s2;
  }
  size_t off = (size_t)b * strideK + (size_t)i + (size_t)j * ldk;
  if (Kout) Kout[off] = k;
  if (Kyout) Kyout[off] = ky;
}

// ---------------------------------------------------------------------------------------
// Mean function + augmented right-hand-side rows (MeanFunc.R:8-13; ForOptimisationLog.R:34,55).
//   meanv[i] = Xc[i,:] . w + b'   (0 for a Zero mean)
//   Ky[ne + 0, i] = y[i] - meanv[i]       (appended row of the augmented matrix: forward solve for free;
//                                          ne = n rounded up to even, padded with an identity row/column)
//   Ky[ne + 1, i] = y[i]                  (only when nrhs == 2: the reference's W uses the
//                                          un-centred targets, OptimisationGradientLog.R:51-53)
// grid (ceil(n/128), batch), block 128.
// ---------------------------------------------------------------------------------------
__global__ void fill_aug_kernel(const double* __restrict__ X, int n, int ne, int d, int ldx, long long strideX,
                                const double* __restrict__ meanw, int ldmu, int linear,
                                const double* __restrict__ y, int ldn, double* __restrict__ meanv,
                                double* __restrict__ Ky, int ld, long long strideK, int nrhs) {
  int i = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (i >= ne) return;
  double* col = Ky + (size_t)b * strideK + (size_t)i * ld;
  if (i >= n) {   // identity padding column (n odd): keeps every sub-block origin even
    for (int r = 0; r < ne + nrhs; r++) col[r] = (r == i) ? 1.0 : 0.0;
    return;
  }
  double m = 0.0;
  if (linear) {
    const double* x = X + (size_t)b * strideX + i;
    const double* w = meanw + (size_t)b * ldmu;
    for (int k = 0; k < d; k++) m += x[(size_t)k * ldx] * w[k];
    m += w[d];
  }
  meanv[(size_t)b * ldn + i] = m;
  double yi = y[(size_t)b * ldn + i];
  for (int r = n; r < ne; r++) col[r] = 0.0;   // padding row
  col[ne] = yi - m;
  if (nrhs == 2) col[ne + 1] = yi;
}

// ---------------------------------------------------------------------------------------
// Leaf of the recursive Cholesky: factor one NB x NB (64 x 64) diagonal block and invert the
// factor (chol(): ForOptimisationLog.R:41-42).  Reads the lower triangle of A[e0.., e0..],
// writes L (lower, zeros above) and Linv (lower, zeros above) at the same coordinates.
//
// The pivot chain (n sequential sqrt/scale steps) is the latency floor of the whole
// factorisation, so the leaf is built from warp-synchronous 32 x 32 primitives whose working
// row lives in REGISTERS (fully unrolled, static indices): per column the critical path is
// shuffle(pivot) -> rsqrt -> scale -> one FMA on the next pivot (~90 cycles), with the rest of
// the rank-1 update fed by shared-memory broadcasts off the critical path.
//   64x64 = [A11 0; A21 A22]:  potf2(A11) | trsm(A21) + inv(L11) | syrk(A22) | potf2(A22) |
//   inv(L22) | T = L21*X11 | X21 = -X22*T.
// A non-positive pivot records info[b] = failing column (1-based, first wins) and the block
// is completed with NaNs (the host turns info into GPS_ERR_NOT_PD).
// grid (batch), block 256, dynamic smem LEAF_SMEM.
// ---------------------------------------------------------------------------------------
constexpr int LS = NB + 1;    // row stride of the 64 x 64 shared tiles (odd => conflict-free rows)
constexpr int LT = 33;
constexpr size_t LEAF_SMEM = (2 * NB * LS + 32 * LT + NB) * sizeof(double);

// The 32x32 warp primitives keep the working row/column in registers but stay COMPACT: the
// outer loop over the pivot is rolled and the register window rotates (the shift is fused
// into the FMA destination: a[c-1] = a[c] - l*lc), so indices are static without unrolling
// 32 x 32 steps.  A fully unrolled version is ~200 KB of straight-line SASS executed once and
// is instruction-fetch bound (measured: 20 K warp-instructions in 120 K cycles).

// One warp, lane = row.  a[] = row `lane` of a 32x32 SPD block (lower part valid).  On exit
// sL[r*LS + c] holds the factor (zeros above the diagonal), sInvD[j] = 1 / L_jj.
// Returns the first non-positive pivot column (1-based) or 0.
__device__ __forceinline__ int warp_potf2_32(double (&a)[32], double* sL, double* sInvD, int lane) {
  int bad = 0;
#pragma unroll 1
  for (int j = 0; j < 32; j++) {
    double d = __shfl_sync(0xffffffffu, a[0], j);       // pivot: column j of row j
    if (!(d > 0.0) && bad == 0) bad = j + 1;
    double rs = rsqrt(d);
    double lr = (lane > j) ? a[0] * rs : ((lane == j) ? d * rs : 0.0);
    sL[lane * LS + j] = lr;
    if (lane == j) sInvD[j] = rs;
    __syncwarp();
    const double* colj = sL + (j + 1) * LS + j;         // L[j+1.., j]: broadcast reads
#pragma unroll
    for (int c = 1; c < 32; c++) {
      double lc = (j + c < 32) ? colj[(c - 1) * LS] : 0.0;
      a[c - 1] = a[c] - lr * lc;                        // rank-1 update + window shift
    }
  }
  return bad;
}

// One warp, lane = column c of X = L^-1 for a 32x32 lower-triangular L in sL (inverse
// diagonal in sInvD).  Writes sX[r*LS + c] (zeros above the diagonal).
__device__ __forceinline__ void warp_trinv_32(const double* sL, const double* sInvD, double* sX, int lane) {
  double x[32];
#pragma unroll
  for (int r = 0; r < 32; r++) x[r] = (r == lane) ? 1.0 : 0.0;
#pragma unroll 1
  for (int k = 0; k < 32; k++) {
    double xk = x[0] * sInvD[k];
    sX[k * LS + lane] = xk;
    const double* colk = sL + (k + 1) * LS + k;
#pragma unroll
    for (int r = 1; r < 32; r++) {
      double lv = (k + r < 32) ? colk[(r - 1) * LS] : 0.0;
      x[r - 1] = x[r] - lv * xk;
    }
  }
}

// One warp, lane = row of a 32x32 block B: B <- B * L^-T (L lower 32x32 in sL) by forward
// substitution; result written to sOut[lane*LS + j].
__device__ __forceinline__ void warp_trsm_32(double (&a)[32], const double* sL, const double* sInvD, double* sOut,
                                             int lane) {
#pragma unroll 1
  for (int j = 0; j < 32; j++) {
    double x = a[0] * sInvD[j];
    sOut[lane * LS + j] = x;
    const double* colj = sL + (j + 1) * LS + j;
#pragma unroll
    for (int c = 1; c < 32; c++) {
      double lc = (j + c < 32) ? colj[(c - 1) * LS] : 0.0;
      a[c - 1] = a[c] - x * lc;
    }
  }
}

__global__ void __launch_bounds__(256) potf2_inv_kernel(const double* __restrict__ A, double* __restrict__ L,
                                                        double* __restrict__ Linv, int ld, long long stride,
                                                        int e0, int nb, int* __restrict__ info) {
  extern __shared__ __align__(16) double leaf_smem[];
  double* S = leaf_smem;                 // 64 x LS : A, then L
  double* X = leaf_smem + NB * LS;       // 64 x LS : L^-1
  double* T = leaf_smem + 2 * NB * LS;   // 32 x LT
  double* invd = T + 32 * LT;            // 64
  __shared__ int sbad;
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const double* Ab = A + (size_t)b * stride + (size_t)e0 + (size_t)e0 * ld;
  for (int idx = tid; idx < NB * NB; idx += 256) {
    int r = idx & 63, c = idx >> 6;
    double v = (r == c) ? 1.0 : 0.0;     // identity padding for a partial block
    if (r < nb && c < nb) v = (r >= c) ? Ab[(size_t)r + (size_t)c * ld] : 0.0;
    S[r * LS + c] = v;
    X[r * LS + c] = 0.0;
  }
  if (tid == 0) sbad = 0;
  __syncthreads();
  if (warp == 0) {                        // potf2(A11)
    double a[32];
#pragma unroll
    for (int c = 0; c < 32; c++) a[c] = S[lane * LS + c];
    int bad = warp_potf2_32(a, S, invd, lane);
    if (bad && lane == 0) sbad = bad;
  }
  __syncthreads();
  if (warp == 1) {                        // L21 = A21 * L11^-T by substitution, lane = row
    double a[32];
#pragma unroll
    for (int c = 0; c < 32; c++) a[c] = S[(32 + lane) * LS + c];
    warp_trsm_32(a, S, invd, S + 32 * LS, lane);
  } else if (warp == 0) {
    warp_trinv_32(S, invd, X, lane);      // X11 = L11^-1
  }
  __syncthreads();
  if (warp < 4) {                         // A22 -= L21 * L21'  (warp w: columns 8w..8w+7; lane = row)
    double xr[32];
#pragma unroll
    for (int k = 0; k < 32; k++) xr[k] = S[(32 + lane) * LS + k];
#pragma unroll 1
    for (int cc = 0; cc < 8; cc++) {
      int c = warp * 8 + cc;
      double acc = S[(32 + lane) * LS + 32 + c];
#pragma unroll
      for (int k = 0; k < 32; k++) acc -= xr[k] * S[(32 + c) * LS + k];
      S[(32 + lane) * LS + 32 + c] = acc;
    }
  }
  __syncthreads();
  if (warp == 0) {                        // potf2(A22), then X22 = L22^-1
    double a[32];
#pragma unroll
    for (int c = 0; c < 32; c++) a[c] = S[(32 + lane) * LS + 32 + c];
    int bad = warp_potf2_32(a, S + 32 * LS + 32, invd + 32, lane);
    if (bad && lane == 0 && sbad == 0) sbad = 32 + bad;
    __syncwarp();
    warp_trinv_32(S + 32 * LS + 32, invd + 32, X + 32 * LS + 32, lane);
  }
  __syncthreads();
  {                                       // T = L21 * X11   (warp w: columns 4w..4w+3; lane = row)
    double xr[32];
#pragma unroll
    for (int k = 0; k < 32; k++) xr[k] = S[(32 + lane) * LS + k];
#pragma unroll 1
    for (int cc = 0; cc < 4; cc++) {
      int c = warp * 4 + cc;
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < 32; k++) acc += xr[k] * X[k * LS + c];
      T[lane * LT + c] = acc;
    }
  }
  __syncthreads();
  {                                       // X21 = -X22 * T
    double xr[32];
#pragma unroll
    for (int k = 0; k < 32; k++) xr[k] = X[(32 + lane) * LS + 32 + k];
#pragma unroll 1
    for (int cc = 0; cc < 4; cc++) {
      int c = warp * 4 + cc;
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < 32; k++) acc -= xr[k] * T[k * LT + c];
      X[(32 + lane) * LS + c] = acc;
    }
  }
  __syncthreads();
  if (tid == 0 && sbad != 0 && sbad <= nb) atomicCAS(&info[b], 0, e0 + sbad);
  double* Lb = L + (size_t)b * stride + (size_t)e0 + (size_t)e0 * ld;
  double* Ib = Linv + (size_t)b * stride + (size_t)e0 + (size_t)e0 * ld;
  for (int idx = tid; idx < NB * NB; idx += 256) {
    int r = idx & 63, c = idx >> 6;
    if (r < nb && c < nb) {
      Lb[(size_t)r + (size_t)c * ld] = (r >= c) ? S[r * LS + c] : 0.0;
      Ib[(size_t)r + (size_t)c * ld] = (r >= c) ? X[r * LS + c] : 0.0;
    }
  }
}

// ---------------------------------------------------------------------------------------
// NLML assembly (ForOptimisationLog.R:63):  nlml = 1/2 z'z + sum log L_ii + n/2 log(2 pi),
// z = L^-1 (y - m) = augmented row n of L.  Also gathers the augmented rows into contiguous
// vectors zvec[r][k] for the alpha kernel.  grid (batch), block 256.
// ---------------------------------------------------------------------------------------
__global__ void nlml_finish_kernel(const double* __restrict__ L, int n, int ne, int ld, long long stride, int nrhs,
                                   double* __restrict__ zvec, int ldn, double* __restrict__ out,
                                   double* __restrict__ logdet_out) {
  __shared__ double sh[256];
  int b = blockIdx.x, tid = threadIdx.x;
  const double* Lb = L + (size_t)b * stride;
  double q = 0.0, ld_sum = 0.0;
  for (int k = tid; k < n; k += blockDim.x) {
    double z0 = Lb[(size_t)ne + (size_t)k * ld];
    zvec[((size_t)b * 2 + 0) * ldn + k] = z0;
    if (nrhs == 2) zvec[((size_t)b * 2 + 1) * ldn + k] = Lb[(size_t)ne + 1 + (size_t)k * ld];
    q += z0 * z0;
    ld_sum += log(Lb[(size_t)k + (size_t)k * ld]);
  }
  q = block_sum(q, sh);
  ld_sum = block_sum(ld_sum, sh);
  if (tid == 0) {
    out[b] = 0.5 * q + ld_sum + 0.5 * (double)n * 1.8378770664093454835606594728112;   // log(2 pi)
    if (logdet_out) logdet_out[b] = ld_sum;
  }
}

// alpha_r[i] = sum_{k >= i} Linv[k,i] * z_r[k]   (alpha = L^-T z; ForOptimisationLog.R:55).
// One warp per column i. grid (ceil(n/8), nrhs, batch), block 256.
__global__ void alpha_kernel(const double* __restrict__ Linv, int n, int ld, long long stride,
                             const double* __restrict__ zvec, int ldn, double* __restrict__ alpha) {
  int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31, r = blockIdx.y, b = blockIdx.z;
  if (i >= n) return;
  const double* col = Linv + (size_t)b * stride + (size_t)i * ld;
  const double* z = zvec + ((size_t)b * 2 + r) * ldn;
  double s = 0.0;
  for (int k = i + lane; k < n; k += 32) s += col[k] * z[k];
  s = warp_sum(s);
  if (lane == 0) alpha[((size_t)b * 2 + r) * ldn + i] = s;
}

// ---------------------------------------------------------------------------------------
// W o K (OptimisationGradientLog.R:51-53 in W-form):  W = a a' - Ky^-1,  M = W o K, from the
// lower triangles of Kinv and K; writes the full symmetric M (both halves) and wdiag = diag(W).
// grid (ceil(n/32), ceil(n/32), batch) — tiles with ti < tj exit; block (32, 8).
// ---------------------------------------------------------------------------------------
__global__ void wk_kernel(const double* __restrict__ Kinv, const double* __restrict__ K, int n, int ld,
                          long long stride, const double* __restrict__ alpha_w, int ldn,
                          double* __restrict__ M, double* __restrict__ wdiag) {
  __shared__ double tile[32][33];
  int ti = blockIdx.x, tj = blockIdx.y, b = blockIdx.z;
  if (ti < tj) return;
  const double* Ki = Kinv + (size_t)b * stride;
  const double* Kb = K + (size_t)b * stride;
  double* Mb = M + (size_t)b * stride;
  const double* a = alpha_w + (size_t)b * 2 * ldn;
  int i = ti * 32 + threadIdx.x;
  for (int rr = 0; rr < 4; rr++) {
    int jl = threadIdx.y + 8 * rr;
    int j = tj * 32 + jl;
    double m = 0.0;
    if (i < n && j < n) {
      int ii = i, jj = j;
      if (ti == tj && i < j) { ii = j; jj = i; }   // diagonal tile: read the lower element
      size_t off = (size_t)ii + (size_t)jj * ld;
      double w = a[ii] * a[jj] - Ki[off];
      m = w * Kb[off];
      Mb[(size_t)i + (size_t)j * ld] = m;
      if (i == j) wdiag[(size_t)b * ldn + i] = w;
    }
    tile[jl][threadIdx.x] = m;
  }
  if (ti == tj) return;
  __syncthreads();
  // mirrored write: M[j, i] for the tile (tj, ti); coalesced along j
  int jo = tj * 32 + threadIdx.x;
  for (int rr = 0; rr < 4; rr++) {
    int il = threadIdx.y + 8 * rr;
    int io = ti * 32 + il;
    if (jo < n && io < n) Mb[(size_t)jo + (size_t)io * ld] = tile[threadIdx.x][il];
  }
}

// ---------------------------------------------------------------------------------------
// Gradient assembly.  grad_cols: one warp per column k of Xaug (k = 0..d):
//   q[k] = sum_mt colpart[mt][k]          (= sum_i Xaug[i,k] (M Xaug)[i,k];  q[d] = sum(M))
//   t[k] = sum_i rs[i] * Xaug[i,k]^2
//   u[k] = sum_i Xaug[i,k] * alpha_c[i]   (mean-function gradients; u[d] = sum alpha_c)
// grid (ceil((d+1)/8), batch), block 256.
// ---------------------------------------------------------------------------------------
__global__ void grad_cols_kernel(const double* __restrict__ X, int n, int d, int ldx, long long strideX,
                                 const double* __restrict__ colpart, int ldp, int mtiles,
                                 const double* __restrict__ rs, const double* __restrict__ alpha, int ldn,
                                 double* __restrict__ qtu) {
  int k = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31, b = blockIdx.y;
  if (k > d) return;
  const double* x = X + (size_t)b * strideX + (size_t)k * ldx;
  const double* r = rs + (size_t)b * ldn;
  const double* ac = alpha + (size_t)b * 2 * ldn;   // rhs 0 = centred targets
  double t = 0.0, u = 0.0;
  for (int i = lane; i < n; i += 32) {
    double xv = x[i];
    t += r[i] * xv * xv;
    u += xv * ac[i];
  }
  t = warp_sum(t);
  u = warp_sum(u);
  if (lane == 0) {
    double q = 0.0;
    for (int mt = 0; mt < mtiles; mt++) q += colpart[((size_t)b * mtiles + mt) * ldp + k];
    double* o = qtu + (size_t)b * 3 * ldp;
    o[k] = q;
    o[ldp + k] = t;
    o[2 * ldp + k] = u;
  }
}

// grad_finish (OptimisationGradientLog.R:51-57,88-89 in W-form). grid (batch), block 256.
__global__ void grad_finish_kernel(const double* __restrict__ qtu, int ldp, int n, int d, int p, ModelOpts o,
                                   const double* __restrict__ wdiag, const int* __restrict__ cens_rank,
                                   int ldn, const double* __restrict__ hyp, const double* __restrict__ ell,
                                   int ldell, const double* __restrict__ mu, int ldmu, int len,
                                   double* __restrict__ grad) {
  __shared__ double sh[256];
  int b = blockIdx.x, tid = threadIdx.x;
  const double* q = qtu + (size_t)b * 3 * ldp;
  const double* t = q + ldp;
  const double* u = q + 2 * ldp;
  double* g = grad + (size_t)b * len;
  double tr = 0.0, trc = 0.0;
  for (int i = tid; i < n; i += blockDim.x) {
    double w = wdiag[(size_t)b * ldn + i];
    tr += w;
    if (cens_rank[(size_t)b * ldn + i] >= 0) trc += w;
  }
  tr = block_sum(tr, sh);
  trc = block_sum(trc, sh);
  const double* el = ell + (size_t)b * ldell;
  if (o.cov_form == 1) {
    for (int k = tid; k < d; k += blockDim.x) g[2 + k] = -(t[k] - q[k]) / (el[k] * el[k]);
  } else {
    double s = 0.0;
    for (int k = tid; k < d; k += blockDim.x) s += t[k] - q[k];
    s = block_sum(s, sh);
    if (tid == 0) g[2] = -s / (el[0] * el[0]);
  }
  if (o.mean_form == 1) {
    // stored X is centred: x_ik = xc_ik + mu_k  =>  d/dw_k = -(u_k + mu_k * u_d)
    for (int k = tid; k < d; k += blockDim.x) g[2 + p + k] = -(u[k] + mu[(size_t)b * ldmu + k] * u[d]);
    if (tid == 0) g[2 + p + d] = -u[d];
  }
  if (tid == 0) {
    g[0] = -0.5 * tr * hyp[b * 4 + 0];
    g[1] = -0.5 * q[d];
    if (o.noise_corr == 1) g[len - 1] = -0.5 * trc * hyp[b * 4 + 2];
  }
}

// ---------------------------------------------------------------------------------------
// Prediction assembly (GPPredict.R:49-56), one warp per test column j:
//   mu[j]   = mstar[j] + sum_i Ks[i,j] * alpha_c[i]
//   varf[j] = sf2 - sum_i V[i,j]^2 ;  vard[j] = varf[j] + sn2
// grid (ceil(m/8), batch), block 256.
// ---------------------------------------------------------------------------------------
__global__ void predict_finish_kernel(const double* __restrict__ Ks, const double* __restrict__ V, int n, int m,
                                      int ldk, long long strideK, const double* __restrict__ alpha, int ldn,
                                      const double* __restrict__ mstar, int ldm, const double* __restrict__ hyp,
                                      double* __restrict__ mu_out, double* __restrict__ varf, double* __restrict__ vard) {
  int j = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31, b = blockIdx.y;
  if (j >= m) return;
  const double* ks = Ks + (size_t)b * strideK + (size_t)j * ldk;
  const double* v = V + (size_t)b * strideK + (size_t)j * ldk;
  const double* a = alpha + (size_t)b * 2 * ldn;
  double s = 0.0, q = 0.0;
  for (int i = lane; i < n; i += 32) {
    s += ks[i] * a[i];
    double vv = v[i];
    q += vv * vv;
  }
  s = warp_sum(s);
  q = warp_sum(q);
  if (lane == 0) {
    double vf = hyp[b * 4 + 1] - q;
    mu_out[(size_t)b * ldm + j] = mstar[(size_t)b * ldm + j] + s;
    varf[(size_t)b * ldm + j] = vf;
    vard[(size_t)b * ldm + j] = vf + hyp[b * 4 + 0];
  }
}

// mstar[j] = Xc*[j,:] . w + b'  (MeanFunc.R on testData). grid (ceil(m/128), batch), block 128
__global__ void mean_kernel(const double* __restrict__ X, int m, int d, int ldx, long long strideX,
                            const double* __restrict__ meanw, int ldmu, int linear, double* __restrict__ out,
                            int ldm) {
  int j = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (j >= m) return;
  double s = 0.0;
  if (linear) {
    const double* x = X + (size_t)b * strideX + j;
    const double* w = meanw + (size_t)b * ldmu;
    for (int k = 0; k < d; k++) s += x[(size_t)k * ldx] * w[k];
    s += w[d];
  }
  out[(size_t)b * ldm + j] = s;
}

// X*c[:, k] = X*[:, k] - mu[k]  (test rows centred with the *training* means). grid (ceil(m/256), d, batch)
__global__ void center_test_kernel(double* __restrict__ X, int m, int d, int ldx, long long strideX,
                                   const double* __restrict__ mu, int ldmu) {
  int i = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y, b = blockIdx.z;
  if (i >= m) return;
  X[(size_t)b * strideX + (size_t)k * ldx + i] -= mu[(size_t)b * ldmu + k];
}

// Mirror the lower triangle into the upper one (for the K output of gps_regress_finalize /
// gps_cov). grid (ceil(n/32), ceil(n/32), batch), block (32, 8).
__global__ void symmetrize_kernel(double* __restrict__ K, int n, int ld, long long stride) {
  __shared__ double tile[32][33];
  int ti = blockIdx.x, tj = blockIdx.y, b = blockIdx.z;
  if (ti <= tj && !(ti == tj)) return;
  double* Kb = K + (size_t)b * stride;
  int i = ti * 32 + threadIdx.x;
  for (int rr = 0; rr < 4; rr++) {
    int jl = threadIdx.y + 8 * rr, j = tj * 32 + jl;
    tile[jl][threadIdx.x] = (i < n && j < n) ? Kb[(size_t)i + (size_t)j * ld] : 0.0;
  }
  __syncthreads();
  int jo = tj * 32 + threadIdx.x;
  for (int rr = 0; rr < 4; rr++) {
    int il = threadIdx.y + 8 * rr, io = ti * 32 + il;
    if (jo < n && io < n && io > jo) Kb[(size_t)jo + (size_t)io * ld] = tile[threadIdx.x][il];
  }
}

// ---------------------------------------------------------------------------------------
// AdjustTrainingSurvivalMeanVariance.R:10,21-31 — truncated-normal moments, one thread per
// censored sample, coalesced.  pnorm/dnorm follow the algorithm R uses (W. J. Cody 1969;
// R nmath pnorm_both / dnorm4) so that `1 - pnorm(alpha)` loses precision exactly the way
// the reference does; compiled with contraction off for this function (see build flags:
// the arithmetic below uses explicit __dmul_rn/__dadd_rn where the order matters).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double cody_pnorm(double x) {
  const double a[5] = {2.2352520354606839287, 161.02823106855587881, 1067.6894854603709582,
                       18154.981253343561249, 0.065682337918207449113};
  const double bb[4] = {47.20258190468824187, 976.09855173777669322, 10260.932208618978205,
                        45507.789335026729956};
  const double c[9] = {0.39894151208813466764, 8.8831497943883759412, 93.506656132177855979,
                       597.27027639480026226, 2494.5375852903726711, 6848.1904505362823326,
                       11602.651437647350124, 9842.7148383839780218, 1.0765576773720192317e-8};
  const double dd[8] = {22.266688044328115691, 235.38790178262499861, 1519.377599407554805,
                        6485.558298266760755, 18615.571640885098091, 34900.952721145977266,
                        38912.003286093271411, 19685.429676859990727};
  const double pp[6] = {0.21589853405795699, 0.1274011611602473639, 0.022235277870649807,
                        0.001421619193227893466, 2.9112874951168792e-5, 0.02307344176494017303};
  const double qq[5] = {1.28426009614491121, 0.468238212480865118, 0.0659881378689285515,
                        0.00378239633202758244, 7.29751555083966205e-5};
  const double M_1_SQRT_2PI_ = 0.398942280401432677939946059934;
  if (isnan(x)) return x;
  double y = fabs(x), xnum, xden, temp, xsq, del, cum, ccum;
  if (y <= 0.67448975) {
    if (y > 2.220446049250313e-16 * 0.5) {
      xsq = __dmul_rn(x, x);
      xnum = __dmul_rn(a[4], xsq);
      xden = xsq;
#pragma unroll
      for (int i = 0; i < 3; i++) {
        xnum = __dmul_rn(__dadd_rn(xnum, a[i]), xsq);
        xden = __dmul_rn(__dadd_rn(xden, bb[i]), xsq);
      }
    } else {
      xnum = xden = 0.0;
    }
    temp = __ddiv_rn(__dmul_rn(x, __dadd_rn(xnum, a[3])), __dadd_rn(xden, bb[3]));
    return __dadd_rn(0.5, temp);
  }
  if (y <= 5.656854249492380195206754896838) {
    xnum = __dmul_rn(c[8], y);
    xden = y;
#pragma unroll
    for (int i = 0; i < 7; i++) {
      xnum = __dmul_rn(__dadd_rn(xnum, c[i]), y);
      xden = __dmul_rn(__dadd_rn(xden, dd[i]), y);
    }
    temp = __ddiv_rn(__dadd_rn(xnum, c[7]), __dadd_rn(xden, dd[7]));
    xsq = trunc(__dmul_rn(y, 16.0)) / 16.0;
    del = __dmul_rn(__dadd_rn(y, -xsq), __dadd_rn(y, xsq));
    cum = __dmul_rn(__dmul_rn(exp(__dmul_rn(__dmul_rn(-xsq, xsq), 0.5)), exp(__dmul_rn(-del, 0.5))), temp);
    ccum = __dadd_rn(1.0, -cum);
    return x > 0.0 ? ccum : cum;
  }
  if (x > -37.5193 && x < 8.2924) {
    xsq = __ddiv_rn(1.0, __dmul_rn(x, x));
    xnum = __dmul_rn(pp[5], xsq);
    xden = xsq;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      xnum = __dmul_rn(__dadd_rn(xnum, pp[i]), xsq);
      xden = __dmul_rn(__dadd_rn(xden, qq[i]), xsq);
    }
    temp = __ddiv_rn(__dmul_rn(xsq, __dadd_rn(xnum, pp[4])), __dadd_rn(xden, qq[4]));
    temp = __ddiv_rn(__dadd_rn(M_1_SQRT_2PI_, -temp), y);
    xsq = trunc(__dmul_rn(x, 16.0)) / 16.0;
    del = __dmul_rn(__dadd_rn(x, -xsq), __dadd_rn(x, xsq));
    cum = __dmul_rn(__dmul_rn(exp(__dmul_rn(__dmul_rn(-xsq, xsq), 0.5)), exp(__dmul_rn(-del, 0.5))), temp);
    ccum = __dadd_rn(1.0, -cum);
    return x > 0.0 ? ccum : cum;
  }
  return x > 0.0 ? 1.0 : 0.0;
}

__device__ __forceinline__ double r_dnorm_std(double x) {
  const double M_1_SQRT_2PI_ = 0.398942280401432677939946059934;
  if (isnan(x)) return x;
  x = fabs(x);
  if (x >= 2.0 * 1.3407807929942596e154) return 0.0;   // 2*sqrt(DBL_MAX)
  if (x < 5.0) return __dmul_rn(M_1_SQRT_2PI_, exp(__dmul_rn(__dmul_rn(-0.5, x), x)));
  if (x > 38.56804181549334) return 0.0;   // sqrt(-2*ln2*(DBL_MIN_EXP + 1 - DBL_MANT_DIG))
  double x1 = ldexp(rint(ldexp(x, 16)), -16);
  double x2 = __dadd_rn(x, -x1);
  double e1 = exp(__dmul_rn(__dmul_rn(-0.5, x1), x1));
  double e2 = exp(__dmul_rn(__dadd_rn(__dmul_rn(-0.5, x2), -x1), x2));
  return __dmul_rn(M_1_SQRT_2PI_, __dmul_rn(e1, e2));
}

__global__ void adjust_kernel(const double* __restrict__ a, const double* __restrict__ mu,
                              const double* __restrict__ s, int c, double* __restrict__ mean_out,
                              double* __restrict__ var_out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  double ai = a[i], mi = mu[i], si = s[i];
  double alpha = __ddiv_rn(__dadd_rn(ai, -mi), si);                 // :10
  double phi = r_dnorm_std(alpha);                                  // :21
  double Phi = cody_pnorm(alpha);                                   // :22
  double om = __dadd_rn(1.0, -Phi);                                 // the lossy 1 - pnorm
  double em, ev;
  if (om < 1e-12) {                                                 // :28-31
    em = ai;
    ev = 0.0;
  } else {
    double lam = __ddiv_rn(phi, om);
    em = __dadd_rn(mi, __dmul_rn(si, lam));                                                      // :25
    ev = __dmul_rn(__dmul_rn(si, si), __dadd_rn(1.0, -__dmul_rn(lam, __dadd_rn(lam, -alpha))));  // :26
  }
  mean_out[i] = em;
  var_out[i] = ev;
}

}  // namespace gps
