// kernels.cuh — the non-GEMM kernels of the GP hot path (all FP64, column-major, batched
// over blockIdx.z or blockIdx.y as noted).  Every hyper-parameter is read from device memory
// so that an evaluation is a static CUDA graph.  All reductions use a fixed order
// (per-thread strided serial sum, then a shared-memory tree) => run-to-run reproducible.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "dmma_gemm.cuh"

namespace gps {

constexpr int NB = 64;   // leaf block of the recursive Cholesky / triangular inverse

// Options mirrored from gpsurv.h (kept as ints on the device)
struct ModelOpts {
  int cov_form;     // 0 SqExp, 1 ARD, 2/3/4 Matern nu = 1/2, 3/2, 5/2
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

// cens_rank[i] = index of row i among the censored rows (events == 0), else -1 (once per gps_set_data).
__global__ void cens_rank_kernel(const double* __restrict__ events, int n, int ldn, int* __restrict__ rank,
                                 int* __restrict__ ncens, int batch) {
  // one warp per fit: 32 rows per step, ranks by ballot + popcount (grid = batch, block = 32)
  const int b = blockIdx.x, lane = threadIdx.x;
  if (b >= batch) return;
  int c = 0;
  for (int i0 = 0; i0 < n; i0 += 32) {
    const int i = i0 + lane;
    const bool cens = (i < n) && events[(size_t)b * ldn + i] == 0.0;
    const unsigned mask = __ballot_sync(0xffffffffu, cens);
    if (i < n) rank[(size_t)b * ldn + i] = cens ? c + __popc(mask & ((1u << lane) - 1u)) : -1;
    c += __popc(mask);
  }
  if (lane == 0) ncens[b] = c;
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
                                double* __restrict__ meanw, double* __restrict__ noise_diag,
                                double* __restrict__ wsc) {
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
  // k-weights of the weighted Gram product X diag(1/ell^2) X' (entries beyond d stay zero)
  if (wsc)
    for (int k = tid; k < d; k += blockDim.x) {
      const double e = exp(pr[2 + (p == 1 ? 0 : k)]);
      wsc[(size_t)b * ldell + k] = 1.0 / (e * e);
    }
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
  // 8 independent loads in flight per thread: with one, the loop is bound by DRAM latency, not bandwidth
  // (measured 643 us for 1.46 GB at C3 x 32); the sum stays in column order.
  constexpr int UN = 8;
  int k = k0;
  for (; k + UN <= k1; k += UN) {
    double xv[UN], dv[UN];
#pragma unroll
    for (int u = 0; u < UN; u++) xv[u] = x[(size_t)(k + u) * ldx];
#pragma unroll
    for (int u = 0; u < UN; u++) {
      if (p == 1) dv[u] = el[0];
      else if (!as_written) dv[u] = el[k + u];
      else {
        xv[u] += m[k + u];
        dv[u] = el[(int)(((long long)i + (long long)(k + u) * nrec) % p)];
      }
    }
#pragma unroll
    for (int u = 0; u < UN; u++) {
      const double v = xv[u] / dv[u];
      z[(size_t)(k + u) * ldz] = v;
      acc += v * v;
    }
  }
  for (; k < k1; k++) {
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

// Squared row norms of the scaled rows = diagonal of the (split) Gram matrix, slices summed in fixed order.
// grid (ceil(n/256), batch)
__global__ void gram_diag_kernel(const double* __restrict__ Gpart, int n, int ldg, long long strideSplit,
                                 long long strideG, int ksplit, double* __restrict__ s, int ldn) {
  int i = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (i >= n) return;
  double a = 0.0;
  for (int k = 0; k < ksplit; k++) a += Gpart[(size_t)b * strideG + (size_t)k * strideSplit + (size_t)i + (size_t)i * ldg];
  s[(size_t)b * ldn + i] = a;
}

// Split-K Gram finish: G = sum_s Gpart[s] (fixed order), then the SE epilogue (see EpiCov).
// grid (ceil(n1/32), ceil(n2/8), batch), block (32, 8).
__global__ void cov_finish_kernel(const double* __restrict__ Gpart, int n1, int n2, int ldg, long long strideSplit,
                                  long long strideG, int ksplit, const double* __restrict__ normA,
                                  const double* __restrict__ normB, int ldn, const double* __restrict__ hyp,
                                  const double* __restrict__ noise_diag, double* __restrict__ Kout,
                                  double* __restrict__ Kyout, int ldk, long long strideK, int sym, int kind,
                                  double* __restrict__ Kdout) {
  int i = blockIdx.x * 32 + threadIdx.x, j = blockIdx.y * 8 + threadIdx.y, b = blockIdx.z;
  if (i >= n1 || j >= n2) return;
  if (sym && i < j) return;   // lower triangle only, like the fused path
  double gsum = 0.0;
  for (int s = 0; s < ksplit; s++)
    gsum += Gpart[(size_t)b * strideG + (size_t)s * strideSplit + (size_t)i + (size_t)j * ldg];
  double s2 = hyp[b * 4 + 1];
  double r2 = fmax(0.0, normA[(size_t)b * ldn + i] + normB[(size_t)b * ldn + j] - 2.0 * gsum);
  if (sym && i == j) r2 = 0.0;
  double k = cov_kernel_fn(kind, r2, s2), ky = k;
  if (sym && i == j) {
    k = s2;
    ky = noise_diag ? s2 + noise_diag[(size_t)b * ldn + i] : s2;
  }
  size_t off = (size_t)b * strideK + (size_t)i + (size_t)j * ldk;
  if (Kout) Kout[off] = k;
  if (Kyout) Kyout[off] = ky;
  if (Kdout) Kdout[off] = cov_deriv_fn(kind, r2, s2);
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
// Fused front end for few features (d <= FRONT_DMAX): hyper-parameter preparation, covariance by DIRECT
// differences, noise diagonal, mean function and the augmented right-hand-side rows in ONE launch
// (replaces memset + hyp_prep + prescale + rownorm + Gram GEMM + fill_aug).  With d of a few dozen the
// Gram GEMM has a 2-step k-loop and its traffic is all epilogue; here each thread forms
//   r2 = sum_k ((x_ik - x_jk) / ell_k)^2      (the reference's own order of operations: rdist of the scaled
//                                              inputs, CovFunc.R:25-29 — no Gram-trick cancellation at all)
// from rows staged in shared memory.  The reference's drivers all live here (n = 100..500, d <= 6), where it
// removes five of the nine launches of an evaluation.
// grid (ceil(ne/32), ceil(ne/32), batch), block (32, 8); tiles above the diagonal exit.
// ---------------------------------------------------------------------------------------
constexpr int FRONT_DMAX = 32;
__global__ void __launch_bounds__(256) small_front_kernel(
    const double* __restrict__ par, int len, ModelOpts o, int n, int ne, int d, int p, int nrhs,
    const double* __restrict__ X, int ldx, long long strideX, const double* __restrict__ mu, int ldmu,
    const double* __restrict__ events, const int* __restrict__ cens_rank, const double* __restrict__ log_sc2_vec,
    int ldc, int ldn, const double* __restrict__ y, double* __restrict__ hyp, double* __restrict__ ell,
    double* __restrict__ meanw, double* __restrict__ noise_diag, double* __restrict__ meanv, double* __restrict__ K,
    double* __restrict__ Ky, double* __restrict__ Kd, int ld, long long strideK, int* __restrict__ info, int kind) {
  const int ti = blockIdx.x, tj = blockIdx.y, b = blockIdx.z;
  if (ti < tj) return;
  __shared__ double zi[32][FRONT_DMAX + 1], zj[32][FRONT_DMAX + 1];
  __shared__ double s_ell[FRONT_DMAX], s_w[FRONT_DMAX + 1];
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;
  const double* pr = par + (size_t)b * len;
  const int plen = len - (o.noise_corr == 1 ? 1 : 0);
  const double sn2 = exp(pr[0]), s2 = exp(pr[1]);
  const double sc2 = (o.noise_corr == 1) ? exp(pr[len - 1]) : 0.0;
  const bool as_written = (o.ard_mode == 1 && p > 1);
  if (tid < d) s_ell[tid] = exp(pr[2 + (p == 1 ? 0 : tid)]);
  if (tid <= d) {
    double w = 0.0;
    if (o.mean_form == 1) {
      const double* wp = pr + (plen - d - 1);
      w = wp[tid];
      if (tid == d)   // stored X is centred: b' = b + mu . w
        for (int k = 0; k < d; k++) w += wp[k] * mu[(size_t)b * ldmu + k];
    }
    s_w[tid] = w;
  }
  __syncthreads();
  const double* Xb = X + (size_t)b * strideX;
  // stage the scaled rows of both tiles: z = x / ell (intended) or (x + mu) / ell[(i + k n) mod p] (as written)
  for (int c = tid; c < 32 * d; c += 256) {
    const int r = c & 31, k = c >> 5;
    const int i = ti * 32 + r, j = tj * 32 + r;
    double vi = 0.0, vj = 0.0;
    if (i < n) {
      const double xv = Xb[(size_t)i + (size_t)k * ldx];
      vi = as_written ? (xv + mu[(size_t)b * ldmu + k]) / s_ell[(int)(((long long)i + (long long)k * n) % p)] : xv / s_ell[k];
    }
    if (j < n) {
      const double xv = Xb[(size_t)j + (size_t)k * ldx];
      vj = as_written ? (xv + mu[(size_t)b * ldmu + k]) / s_ell[(int)(((long long)j + (long long)k * n) % p)] : xv / s_ell[k];
    }
    zi[r][k] = vi;
    zj[r][k] = vj;
  }
  __syncthreads();
  double* Kb = K + (size_t)b * strideK;
  double* Kyb = Ky + (size_t)b * strideK;
  double* Kdb = Kd ? Kd + (size_t)b * strideK : nullptr;
  const int i = ti * 32 + tx;
  const bool mirror = (ti != tj) && ((ti >> 1) == (tj >> 1));   // upper half of a diagonal 64-block: kept symmetric
  double nd_i = sn2;
  if (i < n) {
    const int r = cens_rank[(size_t)b * ldn + i];
    if (r >= 0) nd_i += (o.noise_corr == 1) ? sc2 : (o.noise_corr == 2 ? exp(log_sc2_vec[(size_t)b * ldc + r]) : 0.0);
  }
#pragma unroll
  for (int rr = 0; rr < 4; rr++) {
    const int jl = ty + 8 * rr, j = tj * 32 + jl;
    if (i >= n || j >= n) continue;
    double r2 = 0.0;
    for (int k = 0; k < d; k++) {
      const double df = zi[tx][k] - zj[jl][k];
      r2 = fma(df, df, r2);
    }
    double kv = cov_kernel_fn(kind, r2, s2), kyv = kv;
    if (i == j) {
      r2 = 0.0;
      kv = s2;
      kyv = s2 + nd_i;
    }
    const size_t off = (size_t)i + (size_t)j * ld, offt = (size_t)j + (size_t)i * ld;
    Kb[off] = kv;
    Kyb[off] = kyv;
    if (Kdb) Kdb[off] = cov_deriv_fn(kind, r2, s2);
    if (mirror || (ti == tj && i != j)) {
      Kb[offt] = kv;
      Kyb[offt] = kyv;
      if (Kdb) Kdb[offt] = cov_deriv_fn(kind, r2, s2);
    }
  }
  if (tj != 0) return;
  // ---- per-row work, once per row block (tiles of the first block column) --------------------------------
  if (ty == 0 && i < ne) {
    double* col = Kyb + (size_t)i * ld;
    if (i >= n) {   // identity padding column (n odd)
      for (int r = 0; r < ne + nrhs; r++) col[r] = (r == i) ? 1.0 : 0.0;
      Kb[(size_t)i + (size_t)i * ld] = 1.0;
    } else {
      double m = 0.0;
      if (o.mean_form == 1) {
        for (int k = 0; k < d; k++) m += Xb[(size_t)i + (size_t)k * ldx] * s_w[k];
        m += s_w[d];
      }
      meanv[(size_t)b * ldn + i] = m;
      noise_diag[(size_t)b * ldn + i] = nd_i;
      const double yi = y[(size_t)b * ldn + i];
      for (int r = n; r < ne; r++) col[r] = 0.0;   // padding row
      col[ne] = yi - m;
      if (nrhs == 2) col[ne + 1] = yi;
    }
  }
  if (ti == 0) {   // once per fit
    if (tid == 0) {
      hyp[b * 4 + 0] = sn2;
      hyp[b * 4 + 1] = s2;
      hyp[b * 4 + 2] = sc2;
      hyp[b * 4 + 3] = 0.0;
      info[b] = 0;
    }
    if (tid < p) ell[(size_t)b * ldmu + tid] = s_ell[tid];
    if (tid <= d) meanw[(size_t)b * ldmu + tid] = s_w[tid];
  }
}

// ---------------------------------------------------------------------------------------
// Fused panel kernel of the blocked Cholesky (chol(): ForOptimisationLog.R:41-42): for the
// block column [e0, e0+nb) it
//   (1) factors the NB x NB diagonal block          L11 = chol(A11)
//   (2) inverts the factor                          X11 = L11^-1          (kept for the panel
//       solve here, for the triangular inverse of the gradient and for prediction)
//   (3) solves the panel below                      L21 = A21 * X11'      (M rows, incl. the
//       appended right-hand-side rows)
// in ONE launch.  Every CTA redoes (1)+(2) for itself — redundant work, but no extra launch
// and no inter-CTA dependency — and owns 128 rows of (3).
//
// (1) is the sequential pivot chain that bounds the whole factorisation at n <~ 4k, so it is
// organised for latency (measured history in DESIGN.md):
//   * blocked: four 16-column sub-panels.  Inside a sub-panel ONE warp (lane = rows lane and
//     lane+32) keeps the 16 live columns of its rows in a rotating register window
//     (p[i-1] = p[i] - l*lc: static indices, rolled loop, compact code, no predicates);
//     the pivot and the next pivot travel by warp shuffle, so the per-column critical path is
//     rsqrt -> mul -> shfl -> fma (~115 cycles) and the 15-element rank-1 update runs in its
//     shadow from shared-memory broadcasts.
//   * after each sub-panel the rank-16 update of the trailing block runs on DMMA from shared
//     memory on all 8 warps.
// (2) is a blocked inverse: four 16 x 16 diagonal blocks inverted in registers (lane = column),
//     then two merge levels  X21 = -X22 (L21 X11)  on DMMA.
// (3) is a 128 x 64 x 64 DMMA GEMM from shared memory (8 warps, 32 x 32 per warp).
// A non-positive pivot records info[b] = failing column (1-based, first wins); the block is
// completed with NaNs and the host turns info into GPS_ERR_NOT_PD.
// grid (max(1, ceil(M/128)), batch), block 256, dynamic smem PANEL_SMEM.
// ---------------------------------------------------------------------------------------
constexpr int PS = NB + 4;     // S  column-major: S[c*PS + r] = A11/L11[r][c]   (== 4 mod 16, even)
constexpr int PX = NB + 4;     // sX row-major:    sX[k*PX + c] = X11[k][c]
constexpr int PA = 128 + 4;    // sA: sA[k*PA + m] = A21[m][k]
constexpr int PT = 32 + 4;     // sT column-major: sT[n*PT + m]
constexpr size_t PANEL_SMEM = (size_t)(NB * PS + NB * PX + NB * PA + 32 * PT + NB) * sizeof(double);
// + the 64 x 64 block of the previous block column, for the fused rank-64 pre-update (panel_kernel<32>, kprev = 64)
constexpr size_t PANEL_SMEM_FUSED = PANEL_SMEM + (size_t)(NB * PS) * sizeof(double);

__device__ __forceinline__ void dmma_acc(double& d0, double& d1, double a_mma, double b_mma) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a_mma), "d"(b_mma));
}

// Branch-free 1/sqrt(d) for the pivot chain: hardware seed (rsqrt.approx.f64, 2^-26) + two Newton steps
// y <- y + y (1/2 - (d/2) y^2) -> ~1 ulp, like rsqrt().  The library routine carries special-case branches, i.e.
// basic-block boundaries the scheduler cannot move the rank-1 update's independent DFMAs across; this one is
// straight-line code, so they interleave with its dependent chain.  d <= 0 / NaN give NaN (the caller flags it).
__device__ __forceinline__ double pivot_rsqrt(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double hd = 0.5 * d;
  double e = fma(-hd * y, y, 0.5);
  y = fma(y, e, y);
  e = fma(-hd * y, y, 0.5);
  y = fma(y, e, y);
  return y;
}

// Sub-panel Q (columns 16Q..16Q+15) of the 64 x 64 block in S, executed by warp 0.
// SLOTS = 2: rows lane and lane+32 are both live (Q < 2); SLOTS = 1: only rows lane+32 (Q >= 2).
template <int Q>
__device__ __forceinline__ void potf2_subpanel(double* S, double* invd, int lane, int& bad) {
  constexpr int S0 = (Q < 2) ? 0 : 1;            // first live slot; the pivot row is in slot Q/2
  constexpr int PSLOT = Q / 2;
  double p[2][16];
#pragma unroll
  for (int s = S0; s < 2; s++)
#pragma unroll
    for (int i = 0; i < 16; i++) p[s][i] = S[(16 * Q + i) * PS + lane + 32 * s];
  double d = __shfl_sync(0xffffffffu, p[PSLOT][0], (16 * Q) & 31);
  double rs = pivot_rsqrt(d);
#pragma unroll 1
  for (int jj = 0; jj < 16; jj++) {
    const int j = 16 * Q + jj;
    if (!(d > 0.0) && bad == 0) bad = j + 1;
    double lr[2];
#pragma unroll
    for (int s = S0; s < 2; s++) {
      const int r = lane + 32 * s;
      lr[s] = (r > j) ? p[s][0] * rs : ((r == j) ? d * rs : 0.0);
      S[j * PS + r] = lr[s];
    }
    if (Q >= 2) S[j * PS + lane] = 0.0;          // rows above the sub-panel: zeros above the diagonal
    if (lane == 0) invd[j] = rs;
    // next pivot = a_{j+1,j+1} - l_{j+1,j}^2, straight from the owner's registers — and its rsqrt is issued
    // HERE, ahead of this column's rank-1 update in program order: a warp issues in order, so started after
    // the update's 30 DFMAs the ~120-cycle rsqrt chain would sit fully exposed behind them (measured 280
    // cycles per pivot); started first, the update's independent DFMAs fill its latency.
    {
      const int o = (j + 1) & 31;
      double a11 = __shfl_sync(0xffffffffu, p[PSLOT][1], o);
      double l10 = __shfl_sync(0xffffffffu, lr[PSLOT], o);
      d = a11 - l10 * l10;
      rs = pivot_rsqrt(d);
    }
    __syncwarp();
    // (the L[j+i][j] broadcasts stay in shared memory: fetching them by warp shuffle from the owners' registers
    //  was measured SLOWER — 7.2 K vs 5.0 K cycles per 16-column sub-panel: two SHFL.32 + a select per value)
    const double* col = S + j * PS + j;
#pragma unroll
    for (int i = 1; i < 16; i++) {
      const double lc = col[i];
#pragma unroll
      for (int s = S0; s < 2; s++) p[s][i - 1] = p[s][i] - lr[s] * lc;
    }
  }
}

// Trailing update after sub-panel Q:  S[R.., R..] -= L[R.., 16Q..16Q+15] * L[R.., 16Q..16Q+15]',
// R = 16(Q+1); lower 8x8 tiles distributed over the 8 warps.
template <int Q>
__device__ __forceinline__ void potf2_trailing(double* S, int warp, int g, int t) {
  constexpr int R = 16 * (Q + 1), NT8 = (NB - R) / 8;
  int tile = 0;
#pragma unroll 1
  for (int mi = 0; mi < NT8; mi++)
#pragma unroll 1
    for (int ni = 0; ni <= mi; ni++, tile++) {
      if ((tile & 7) != warp) continue;
      const int m0 = R + 8 * mi, n0 = R + 8 * ni;
      double2 c = *reinterpret_cast<double2*>(S + (n0 + g) * PS + m0 + 2 * t);
#pragma unroll
      for (int ks = 0; ks < 4; ks++) {
        const int k = 16 * Q + 4 * ks + t;
        dmma_acc(c.x, c.y, -S[k * PS + n0 + g], S[k * PS + m0 + g]);
      }
      *reinterpret_cast<double2*>(S + (n0 + g) * PS + m0 + 2 * t) = c;
    }
}

// One merge of the blocked inverse: rows [R1, R1+SZ), cols [C0, C0+SZ) of X:
//   pass 0:  T   = L21 * X11      (T in sT, column-major)
//   pass 1:  X21 = -X22 * T
// 8x8 output tiles are dealt round-robin to `nw` warps (this warp has index `w`).
template <int SZ>
__device__ __forceinline__ void trinv_merge(const double* S, double* sX, double* sT, int R1, int C0, int pass, int w,
                                            int nw, int g, int t) {
  constexpr int NT8 = SZ / 8;
#pragma unroll 1
  for (int tile = w; tile < NT8 * NT8; tile += nw) {
    const int m0 = 8 * (tile % NT8), n0 = 8 * (tile / NT8);
    double c0 = 0.0, c1 = 0.0;
    if (pass == 0) {
#pragma unroll
      for (int ks = 0; ks < SZ / 4; ks++) {
        const int k = 4 * ks + t;
        dmma_acc(c0, c1, sX[(C0 + k) * PX + C0 + n0 + g], S[(C0 + k) * PS + R1 + m0 + g]);
      }
      *reinterpret_cast<double2*>(sT + (n0 + g) * PT + m0 + 2 * t) = make_double2(c0, c1);
    } else {
#pragma unroll
      for (int ks = 0; ks < SZ / 4; ks++) {
        const int k = 4 * ks + t;
        dmma_acc(c0, c1, -sT[(n0 + g) * PT + k], sX[(R1 + m0 + g) * PX + R1 + k]);
      }
      sX[(R1 + m0 + 2 * t) * PX + C0 + n0 + g] = c0;
      sX[(R1 + m0 + 2 * t + 1) * PX + C0 + n0 + g] = c1;
    }
  }
}

// RT = rows of one panel tile (32: many small CTAs for a lone fit; 128: throughput for batches),
// T = consecutive tiles per CTA.  The LAST CTA along x only publishes L11 / X11.
template <int RT>
__global__ void __launch_bounds__(256) panel_kernel(const double* __restrict__ A, double* __restrict__ L,
                                                    double* __restrict__ Linv, double* __restrict__ Uinv, int ld,
                                                    long long stride, int e0,
                                                    int nb, int M, int T, int* __restrict__ info,
                                                    long long* __restrict__ dbg, int kprev, int merged) {
  // kprev = 64 (RT = 32 only): the rank-64 update from the PREVIOUS block column is still pending for this one
  // (the second leaf of a leaf pair of the recursion).  It is applied here — to the diagonal block and to every
  // row tile, on DMMA from shared memory, before they are used — instead of by a separate K = 64 GEMM launch
  // over the same rows: a lone fit's launches are latency-bound and this removes one in four of its Cholesky's.
  extern __shared__ __align__(16) double panel_smem[];
  int dbg_i = 0;
#define PANEL_STAMP() do { if (dbg && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) dbg[dbg_i++] = clock64(); } while (0)
  PANEL_STAMP();
  double* S = panel_smem;
  double* sX = S + NB * PS;
  double* sA = sX + NB * PX;
  double* sT = sA + NB * PA;
  double* invd = sT + 32 * PT;
  double* sLd = invd + NB;                       // [k][r] = L[e0 + r][e0 - 64 + k]   (kprev only; PS stride)
  double* sLt = sA + 64;                         // [k][m] = L[row0 + m][e0 - 64 + k], beside the A21 tile in sA's rows (RT = 32)
  const int tid = threadIdx.x, cta = blockIdx.x, b = blockIdx.y;
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const double* Ab = A + (size_t)b * stride;
  const double* Lprev = L + (size_t)b * stride + (size_t)(e0 - kprev) * ld;   // columns e0-64 .. e0-1 of L (kprev only)
  const int e1 = e0 + nb;
  // merged != 0: no extra CTA — CTA 0 publishes and then takes its row tiles like the others (large batches: the launch is
  // throughput-bound, every CTA of a fit repeats the pivot chain, and one CTA less per fit can be one wave less)
  const bool publisher = merged ? (cta == 0) : (cta == (int)gridDim.x - 1);
  int row0 = e1 + cta * T * RT;                 // first panel row of this CTA
  int rows = (publisher && !merged) ? 0 : min(RT, e1 + M - row0);

  // stage A11 (zero-filled to 64 x 64) and the A21 tile with coalesced 16-byte async copies
  for (int c = tid; c < NB * 32; c += 256) {
    int col = c >> 5, rc = (c & 31) * 2;
    int valid = (col < nb) ? min(max(nb - rc, 0), 2) : 0;
    const double* src = valid ? (Ab + (size_t)(e0 + rc) + (size_t)(e0 + col) * ld) : Ab;
    unsigned dst = (unsigned)__cvta_generic_to_shared(S + col * PS + rc);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(valid * 8));
  }
  if (RT == 32 && kprev) {
    for (int c = tid; c < NB * 32; c += 256) {
      int k = c >> 5, rc = (c & 31) * 2;
      int valid = min(max(nb - rc, 0), 2);
      const double* src = valid ? (Lprev + (size_t)(e0 + rc) + (size_t)k * ld) : Lprev;
      unsigned dst = (unsigned)__cvta_generic_to_shared(sLd + k * PS + rc);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(valid * 8));
    }
  }
  asm volatile("cp.async.commit_group;\n" ::);
  auto load_a21 = [&]() {
    for (int c = tid; c < NB * (RT / 2); c += 256) {
      int k = c / (RT / 2), mc = (c % (RT / 2)) * 2;
      int valid = (k < nb) ? min(max(rows - mc, 0), 2) : 0;
      const double* src = valid ? (Ab + (size_t)(row0 + mc) + (size_t)(e0 + k) * ld) : Ab;
      unsigned dst = (unsigned)__cvta_generic_to_shared(sA + k * PA + mc);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(valid * 8));
      if (RT == 32 && kprev) {   // the same rows of the previous block column (all 64 columns exist)
        int v2 = min(max(rows - mc, 0), 2);
        const double* src2 = v2 ? (Lprev + (size_t)(row0 + mc) + (size_t)k * ld) : Lprev;
        unsigned dst2 = (unsigned)__cvta_generic_to_shared(sLt + k * PA + mc);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst2), "l"(src2), "r"(v2 * 8));
      }
    }
    asm volatile("cp.async.commit_group;\n" ::);
  };
  load_a21();
  for (int i = tid; i < NB * PX; i += 256) sX[i] = 0.0;
  asm volatile("cp.async.wait_group 1;\n" ::);     // A11 has landed (A21 may still be in flight)
  __syncthreads();
  if (RT == 32 && kprev) {   // A11 -= Ld Ld'  (lower 8 x 8 tiles over the 8 warps; rows >= nb of Ld are zero)
    int tile = 0;
#pragma unroll 1
    for (int mi = 0; mi < NB / 8; mi++)
#pragma unroll 1
      for (int ni = 0; ni <= mi; ni++, tile++) {
        if ((tile & 7) != warp) continue;
        const int m0 = 8 * mi, n0 = 8 * ni;
        double2 c = *reinterpret_cast<double2*>(S + (n0 + g) * PS + m0 + 2 * t);
#pragma unroll
        for (int ks = 0; ks < NB / 4; ks++) {
          const int k = 4 * ks + t;
          dmma_acc(c.x, c.y, -sLd[k * PS + n0 + g], sLd[k * PS + m0 + g]);
        }
        *reinterpret_cast<double2*>(S + (n0 + g) * PS + m0 + 2 * t) = c;
      }
    __syncthreads();
  }
  if (tid >= nb && tid < NB) S[tid * PS + tid] = 1.0;   // identity padding of a partial block
  __syncthreads();
  PANEL_STAMP();

  // ---- (1) blocked factorisation -----------------------------------------------------------
  int bad = 0;
  if (warp == 0) potf2_subpanel<0>(S, invd, lane, bad);
  __syncthreads();
  PANEL_STAMP();
  potf2_trailing<0>(S, warp, g, t);
  __syncthreads();
  PANEL_STAMP();
  if (warp == 0) potf2_subpanel<1>(S, invd, lane, bad);
  __syncthreads();
  PANEL_STAMP();
  potf2_trailing<1>(S, warp, g, t);
  __syncthreads();
  PANEL_STAMP();
  if (warp == 0) potf2_subpanel<2>(S, invd, lane, bad);
  __syncthreads();
  PANEL_STAMP();
  potf2_trailing<2>(S, warp, g, t);
  __syncthreads();
  PANEL_STAMP();
  if (warp == 0) potf2_subpanel<3>(S, invd, lane, bad);
  __syncthreads();
  PANEL_STAMP();

  // ---- (2) blocked inverse ------------------------------------------------------------------
  if (warp < 4 && lane < 16) {                  // 16 x 16 diagonal blocks: lane = column
    const int o = 16 * warp, c = lane;
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = (i == c) ? 1.0 : 0.0;
#pragma unroll 1
    for (int k = 0; k < 16; k++) {
      const double xk = x[0] * invd[o + k];
      sX[(o + k) * PX + o + c] = xk;
      const double* col = S + (o + k) * PS + o + k;
#pragma unroll
      for (int i = 1; i < 16; i++) x[i - 1] = x[i] - col[i] * xk;
    }
  }
  __syncthreads();
  PANEL_STAMP();
  trinv_merge<16>(S, sX, sT + (warp >> 2) * 16 * PT, (warp >> 2) * 32 + 16, (warp >> 2) * 32, 0, warp & 3, 4, g, t);
  __syncthreads();
  trinv_merge<16>(S, sX, sT + (warp >> 2) * 16 * PT, (warp >> 2) * 32 + 16, (warp >> 2) * 32, 1, warp & 3, 4, g, t);
  __syncthreads();
  trinv_merge<32>(S, sX, sT, 32, 0, 0, warp, 8, g, t);
  __syncthreads();
  trinv_merge<32>(S, sX, sT, 32, 0, 1, warp, 8, g, t);
  asm volatile("cp.async.wait_group 0;\n" ::);
  __syncthreads();
  PANEL_STAMP();

  if (publisher) {                              // publish L11 and X11 (lower, zeros above) and X11' (upper)
    if (tid == 0 && bad != 0 && bad <= nb) atomicCAS(&info[b], 0, e0 + bad);
    double* Lb = L + (size_t)b * stride + (size_t)e0 + (size_t)e0 * ld;
    double* Ib = Linv + (size_t)b * stride + (size_t)e0 + (size_t)e0 * ld;
    double* Ub = Uinv + (size_t)b * stride + (size_t)e0 + (size_t)e0 * ld;
    for (int idx = tid; idx < NB * NB; idx += 256) {
      int r = idx & 63, c = idx >> 6;
      if (r < nb && c < nb) {
        Lb[(size_t)r + (size_t)c * ld] = (r >= c) ? S[c * PS + r] : 0.0;
        Ib[(size_t)r + (size_t)c * ld] = (r >= c) ? sX[r * PX + c] : 0.0;
        Ub[(size_t)r + (size_t)c * ld] = (c >= r) ? sX[c * PX + r] : 0.0;   // U = Linv' (trtri_levels works on both)
      }
    }
    PANEL_STAMP();
    if (!merged) return;
  }
  PANEL_STAMP();

  // ---- (3) L21 = A21 * X11'  : C[m][n] = sum_k A21[m][k] * X[n][k]   (DMMA; warps 4 (m) x 2 (n))
  constexpr int FMT = RT / 32;                   // 8-row fragments per warp along m
  const int wm0 = (warp & 3) * (RT / 4), wn0 = (warp >> 2) * 32;
#pragma unroll 1
  for (int tile = 0; tile < T; tile++) {
    if (tile > 0) {
      row0 += RT;
      rows = min(RT, e1 + M - row0);
      if (rows <= 0) break;
      __syncthreads();                           // everyone is done reading the previous tile
      load_a21();
      asm volatile("cp.async.wait_group 0;\n" ::);
      __syncthreads();
    } else if (rows <= 0) {
      break;
    }
    if (RT == 32 && kprev) {   // A21 tile -= Lt Ld'  (32 x 64 outputs = 32 8 x 8 tiles, four per warp)
#pragma unroll 1
      for (int q = 0; q < 4; q++) {
        const int tl = warp * 4 + q, m0 = 8 * (tl & 3), n0 = 8 * (tl >> 2);
        double2 c = *reinterpret_cast<double2*>(sA + (n0 + g) * PA + m0 + 2 * t);
#pragma unroll
        for (int ks = 0; ks < NB / 4; ks++) {
          const int k = 4 * ks + t;
          dmma_acc(c.x, c.y, -sLd[k * PS + n0 + g], sLt[k * PA + m0 + g]);
        }
        *reinterpret_cast<double2*>(sA + (n0 + g) * PA + m0 + 2 * t) = c;
      }
      __syncthreads();
    }
    double acc[FMT][4][2];
#pragma unroll
    for (int i = 0; i < FMT; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll 4
    for (int ks = 0; ks < NB / 4; ks++) {
      double af[FMT], bf[4];
#pragma unroll
      for (int i = 0; i < FMT; i++) af[i] = sA[(4 * ks + t) * PA + wm0 + 8 * i + g];
#pragma unroll
      for (int j = 0; j < 4; j++) bf[j] = sX[(wn0 + 8 * j + g) * PX + 4 * ks + t];
#pragma unroll
      for (int i = 0; i < FMT; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) dmma_acc(acc[i][j][0], acc[i][j][1], bf[j], af[i]);
    }
    double* Lp = L + (size_t)b * stride + (size_t)row0 + (size_t)e0 * ld;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      int n = wn0 + 8 * j + g;
      if (n >= nb) continue;
#pragma unroll
      for (int i = 0; i < FMT; i++) {
        int m = wm0 + 8 * i + 2 * t;
        double* c = Lp + (size_t)m + (size_t)n * ld;
        if (m + 1 < rows) *reinterpret_cast<double2*>(c) = make_double2(acc[i][j][0], acc[i][j][1]);
        else if (m < rows) c[0] = acc[i][j][0];
      }
    }
  }
  PANEL_STAMP();
#undef PANEL_STAMP
}

// ---------------------------------------------------------------------------------------
// Out-of-place transpose of a batch of column-major matrices: T[c + r*ldt] = A[r + c*lda].
// grid (ceil(rows/32), ceil(cols/32), batch), block (32, 8).  Used once per gps_set_data to keep
// Xaug' next to Xaug, so that the gradient contraction M * Xaug is an "NT" GEMM whose operands
// both stream as long contiguous rows.
// ---------------------------------------------------------------------------------------
__global__ void transpose_kernel(const double* __restrict__ A, int rows, int cols, int lda, long long strideA,
                                 double* __restrict__ T, int ldt, long long strideT) {
  __shared__ double tile[32][33];
  const double* Ab = A + (size_t)blockIdx.z * strideA;
  double* Tb = T + (size_t)blockIdx.z * strideT;
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int k = threadIdx.y; k < 32; k += 8) {
    int r = r0 + threadIdx.x, c = c0 + k;
    tile[k][threadIdx.x] = (r < rows && c < cols) ? Ab[(size_t)r + (size_t)c * lda] : 0.0;
  }
  __syncthreads();
  for (int k = threadIdx.y; k < 32; k += 8) {
    int c = c0 + threadIdx.x, r = r0 + k;
    if (r < rows && c < cols) Tb[(size_t)c + (size_t)r * ldt] = tile[threadIdx.x][k];
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
                          double* __restrict__ M, double* __restrict__ wdiag, double* __restrict__ rs_part, int nt32,
                          const double* __restrict__ Kd, double* __restrict__ rsK_part) {
  // Kd != nullptr (Matern): M = W o Kd feeds the length-scale contraction, and the row sums of W o K (needed for
  // d/dlog sf2 = -1/2 sum(W o K)) go to rsK_part; for SqExp / ARD Kd = K and sum(M) serves both.
  __shared__ double tile[32][33];
  __shared__ double tileK[32][33];
  const double* Kdb = Kd ? Kd + (size_t)blockIdx.z * stride : nullptr;
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
      if (Kdb) {
        tileK[jl][threadIdx.x] = m;
        m = w * Kdb[off];
      }
      Mb[(size_t)i + (size_t)j * ld] = m;
      if (i == j) wdiag[(size_t)b * ldn + i] = w;
    } else if (Kdb) {
      tileK[jl][threadIdx.x] = 0.0;
    }
    tile[jl][threadIdx.x] = m;
  }
  __syncthreads();
  if (Kdb) {   // the same partial row sums for W o K
    double* rk = rsK_part + (size_t)b * nt32 * ldn;
    if (threadIdx.y == 2 && i < n) {
      double sum = 0.0;
#pragma unroll 8
      for (int jl = 0; jl < 32; jl++) sum += tileK[jl][threadIdx.x];
      rk[(size_t)tj * ldn + i] = sum;
    }
    if (ti != tj && threadIdx.y == 3) {
      int jo = tj * 32 + threadIdx.x;
      if (jo < n) {
        double sum = 0.0;
#pragma unroll 8
        for (int il = 0; il < 32; il++) sum += tileK[threadIdx.x][il];
        rk[(size_t)ti * ldn + jo] = sum;
      }
    }
  }
  // partial row sums of M over this tile's 32 columns (fixed order): rs_part[b][tj][i]; the mirrored block
  // gives rs_part[b][ti][j].  Every (row block, column block) pair is written exactly once.
  double* rp = rs_part + (size_t)b * nt32 * ldn;
  if (threadIdx.y == 0 && i < n) {
    double sum = 0.0;
#pragma unroll 8
    for (int jl = 0; jl < 32; jl++) sum += tile[jl][threadIdx.x];
    rp[(size_t)tj * ldn + i] = sum;
  }
  if (ti == tj) return;
  if (threadIdx.y == 1) {
    int jo = tj * 32 + threadIdx.x;
    if (jo < n) {
      double sum = 0.0;
#pragma unroll 8
      for (int il = 0; il < 32; il++) sum += tile[threadIdx.x][il];
      rp[(size_t)ti * ldn + jo] = sum;
    }
  }
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
//   t[k] = sum_mt tpart[mt][k]            (= sum_i rowsum(M)[i] * Xaug[i,k]^2, contracted in the GEMM epilogue)
//   u[k] = sum_i Xaug[i,k] * alpha_c[i]   (mean-function gradients; u[d] = sum alpha_c)
// grid (ceil((d+1)/8), batch), block 256.
// ---------------------------------------------------------------------------------------
__global__ void grad_cols_kernel(const double* __restrict__ X, int n, int d, int ldx, long long strideX,
                                 const double* __restrict__ colpart, int ldp, int mtiles, int ksplit,
                                 const double* __restrict__ tpart, const double* __restrict__ alpha, int ldn,
                                 double* __restrict__ qtu, const double* __restrict__ ell, int ldell, int ard,
                                 int linear, int len, double* __restrict__ grad) {
  // colpart holds mtiles * ksplit slices per fit (m-tile major), tpart mtiles slices (written by split 0)
  int k = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31, b = blockIdx.y;
  if (k > d) return;
  double u = 0.0;
  if (linear) {   // mean-function gradients only (MeanFunc.R 'Linear')
    const double* x = X + (size_t)b * strideX + (size_t)k * ldx;
    const double* ac = alpha + (size_t)b * 2 * ldn;   // rhs 0 = centred targets
    for (int i = lane; i < n; i += 32) u += x[i] * ac[i];
    u = warp_sum(u);
  }
  // fixed-order sums of the per-tile partials, spread over the lanes then combined in lane order
  double q = 0.0, t = 0.0;
  const int nq = mtiles * ksplit;
  for (int mt = lane; mt < nq; mt += 32) q += colpart[((size_t)b * nq + mt) * ldp + k];
  for (int mt = lane; mt < mtiles; mt += 32) t += tpart[((size_t)b * mtiles + mt) * ldp + k];
  q = warp_sum(q);
  t = warp_sum(t);
  if (lane == 0) {
    double* o = qtu + (size_t)b * 3 * ldp;
    o[k] = q;
    o[ldp + k] = t;
    o[2 * ldp + k] = u;
    if (ard && k < d) {       // d/dlog ell_k = -(t_k - q_k) / ell_k^2, one length-scale per column
      double e = ell[(size_t)b * ldell + k];
      grad[(size_t)b * len + 2 + k] = -(t - q) / (e * e);
    }
  }
}

// rowsum(M)[i] = sum over the split-K slices of the gradient GEMM's last column (fixed order).
// grid (ceil(n/256), batch)
__global__ void rs_reduce_kernel(const double* __restrict__ rs_part, int n, int ldn, int ksplit, double* __restrict__ rs) {
  int i = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (i >= n) return;
  double s = 0.0;
  for (int k = 0; k < ksplit; k++) s += rs_part[((size_t)b * ksplit + k) * ldn + i];
  rs[(size_t)b * ldn + i] = s;
}

// grad_finish (OptimisationGradientLog.R:51-57,88-89 in W-form). grid (batch), block 256.
__global__ void grad_finish_kernel(const double* __restrict__ qtu, int ldp, int n, int d, int p, ModelOpts o,
                                   const double* __restrict__ wdiag, const int* __restrict__ cens_rank,
                                   int ldn, const double* __restrict__ hyp, const double* __restrict__ ell,
                                   int ldell, const double* __restrict__ mu, int ldmu, int len,
                                   double* __restrict__ grad, const double* __restrict__ rsK) {
  __shared__ double sh[256];
  int b = blockIdx.x, tid = threadIdx.x;
  const double* q = qtu + (size_t)b * 3 * ldp;
  double sumWK = 0.0;   // sum(W o K): q[d] unless the contraction ran on W o Kd (Matern), then from its own row sums
  if (rsK) {
    for (int i = tid; i < n; i += blockDim.x) sumWK += rsK[(size_t)b * ldn + i];
    sumWK = block_sum(sumWK, sh);
  }
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
    // ARD length-scale gradients were written by grad_cols_kernel (one warp per column)
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
    g[1] = -0.5 * (rsK ? sumWK : q[d]);
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

// out[i] = sum_{k <= i} L[i,k] * v[k]   (t(chol(K)) %*% d1 of MakeSyntheticData.R:148; L lower, column-major).
// Block (32, 8): 32 consecutive rows (every load of a column is one coalesced 256-byte run), the k range dealt to the
// eight warps (k = ty, ty + 8, ...: each a fixed-order serial sum), partials added in warp order -> reproducible.
// grid (ceil(n/32), batch).
__global__ void __launch_bounds__(256) trmv_lower_kernel(const double* __restrict__ L, int n, int ld, long long stride,
                                                         const double* __restrict__ v, int ldn, double* __restrict__ out) {
  __shared__ double part[8][33];
  const int tx = threadIdx.x, ty = threadIdx.y, b = blockIdx.y;
  const int i = blockIdx.x * 32 + tx;
  double s = 0.0;
  if (i < n) {
    const double* Lb = L + (size_t)b * stride + i;
    const double* vb = v + (size_t)b * ldn;
    for (int k = ty; k <= i; k += 8) s += Lb[(size_t)k * ld] * vb[k];
  }
  part[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && i < n) {
    double t = part[0][tx];
#pragma unroll
    for (int w = 1; w < 8; w++) t += part[w][tx];
    out[(size_t)b * ldn + i] = t;
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

// ---------------------------------------------------------------------------------------
// CalculateMetrics.R:18,21 — concordance index and RMSE of test predictions.
// survcomp::concordance.index(x, surv.time, surv.event) (third-party, absent from the reference tree; defaults
// method = "noether", outx = TRUE) restated from its published definition: a pair (i, j) is USABLE when the shorter
// observed time is an event (time_i < time_j and event_i = 1; pairs tied in time are not usable); it is CONCORDANT when the
// higher score x belongs to the shorter time (x is a risk score in survcomp's convention — the reference passes the
// predicted log survival time as is), DISCORDANT when the lower one does; pairs tied in x are left out (outx = TRUE;
// CalculateMetrics.R:12-16 jitters tied predictions beforehand anyway, with rnorm: that jitter is NOT restated, ties are
// reported).  c.index = concordant / (concordant + discordant).  Integer pair counts: exact and order-independent.
// grid (ceil(m/128), batch), block 128.  counts[b][3] must be zero on entry; sq_part[b][gridDim.x] gets the per-block
// sums of (time - x)^2 (reduced in block order by metrics_finish_kernel).
// ---------------------------------------------------------------------------------------
__global__ void cindex_kernel(const double* __restrict__ x, const double* __restrict__ time, const double* __restrict__ event,
                              int m, unsigned long long* __restrict__ counts, double* __restrict__ sq_part) {
  __shared__ double sx[128], st[128];
  __shared__ double sh[128];
  __shared__ unsigned long long sc[3];
  const int b = blockIdx.y, tid = threadIdx.x, i = blockIdx.x * 128 + tid;
  const double* xb = x + (size_t)b * m;
  const double* tb = time + (size_t)b * m;
  const double* eb = event + (size_t)b * m;
  const bool live = i < m;
  const double xi = live ? xb[i] : 0.0, ti = live ? tb[i] : 0.0;
  const bool ev_i = live && eb[i] == 1.0;
  if (tid < 3) sc[tid] = 0ull;
  unsigned long long conc = 0, disc = 0, tied = 0;
  for (int j0 = 0; j0 < m; j0 += 128) {
    __syncthreads();
    const int j = j0 + tid;
    sx[tid] = j < m ? xb[j] : 0.0;
    st[tid] = j < m ? tb[j] : -INFINITY;      // never later than any time: padding is never the longer time of a pair
    __syncthreads();
    if (ev_i) {
      const int lim = min(128, m - j0);
      for (int q = 0; q < lim; q++) {
        if (ti < st[q]) {
          conc += xi > sx[q];
          disc += xi < sx[q];
          tied += xi == sx[q];
        }
      }
    }
  }
  atomicAdd(&sc[0], conc);
  atomicAdd(&sc[1], disc);
  atomicAdd(&sc[2], tied);
  const double dv = live ? (ti - xi) : 0.0;
  sh[tid] = dv * dv;
  __syncthreads();
  for (int s2 = 64; s2 > 0; s2 >>= 1) {
    if (tid < s2) sh[tid] += sh[tid + s2];
    __syncthreads();
  }
  if (tid == 0) sq_part[(size_t)b * gridDim.x + blockIdx.x] = sh[0];
  if (tid < 3) atomicAdd(&counts[(size_t)b * 3 + tid], sc[tid]);
}

// c.index = concordant / (concordant + discordant) (NaN without usable untied pairs), rmse = sqrt(sum / m).  grid (batch), block 32
__global__ void metrics_finish_kernel(const unsigned long long* __restrict__ counts, const double* __restrict__ sq_part, int nblk,
                                      int m, double* __restrict__ cindex, double* __restrict__ rmse) {
  const int b = blockIdx.x;
  if (threadIdx.x != 0) return;
  const double c = (double)counts[(size_t)b * 3], d = (double)counts[(size_t)b * 3 + 1];
  cindex[b] = (c + d > 0.0) ? c / (c + d) : NAN;
  double s = 0.0;
  for (int k = 0; k < nblk; k++) s += sq_part[(size_t)b * nblk + k];
  rmse[b] = sqrt(s / (double)m);
}

}  // namespace gps
