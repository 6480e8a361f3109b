// nearpd.cuh — Matrix::nearPD on the device: the fallback ForOptimisationLog.R:43-51 takes when chol(Ky)
// fails.  Matrix is a third-party package (not under /root/reference, version unpinned); the algorithm restated here is
// the published one its nearPD implements with the defaults the reference uses (Higham 2002, alternating projections
// with Dykstra's correction; corr = FALSE, keepDiag = FALSE, do2eigen = TRUE, conv.norm.type = "I", eig.tol = 1e-6,
// conv.tol = 1e-7, posd.tol = 1e-8, maxit = 100); the tests hold it against a line-cited CPU restatement.
//
// The symmetric eigen-decompositions it needs are a cyclic two-sided JACOBI method in the parallel (round-robin) ordering:
// every step rotates n/2 disjoint index pairs at once,
//     A <- J' A J,  V <- V J,
// as one pass over column pairs (one CTA per pair: computes the rotation from a_pp, a_qq, a_pq, rotates the two columns of
// A and of V — contiguous streams) and one pass over rows (one CTA per column: the column is staged in shared memory, all
// n/2 row rotations applied, written back contiguously).  For n <= 2048 both matrices (2 x 8 n^2 <= 67 MB) stay in the
// 126 MB L2.  Jacobi is slow next to tridiagonalisation + QR but has no sequential bulge chase, is accurate to the last
// bits of the small eigenvalues (which is what decides the projection here) and is a fallback path: n = 2000 takes
// ~9 sweeps x 1999 steps of ~35 us.
// Included by gpsurv.cu (uses the handle internals).
#pragma once

namespace {

constexpr int NEARPD_MAX_N = 4096;     // the row pass stages one column (8 n bytes) in shared memory

// round-robin pairing of step s: indices 0..ne-2 sit on a circle, ne-1 is fixed (ne even)
__device__ __forceinline__ void jacobi_pair(int s, int k, int ne, int& p, int& q) {
  const int m = ne - 1;
  if (k == 0) {
    p = s % m;
    q = ne - 1;
  } else {
    p = (s + k) % m;
    q = (s - k + m) % m;
  }
  if (p > q) {
    const int t = p;
    p = q;
    q = t;
  }
}

// one CTA per pair: rotation + column pass on A and V.  rot[k] = (c, s); identity for padded / already-diagonal pairs
__global__ void jacobi_cols_kernel(double* __restrict__ A, double* __restrict__ V, int n, int ne, int ld, int step,
                                   double2* __restrict__ rot) {
  int p, q;
  jacobi_pair(step, blockIdx.x, ne, p, q);
  if (q >= n) {                              // padding index (n odd)
    if (threadIdx.x == 0) rot[blockIdx.x] = make_double2(1.0, 0.0);
    return;
  }
  const double app = A[(size_t)p + (size_t)p * ld], aqq = A[(size_t)q + (size_t)q * ld], apq = A[(size_t)p + (size_t)q * ld];
  __syncthreads();                           // every thread has the three values before any column element changes
  double c = 1.0, s = 0.0;
  if (fabs(apq) > 1e-300 && fabs(apq) > 2.5e-17 * sqrt(fabs(app) * fabs(aqq))) {
    const double tau = (aqq - app) / (2.0 * apq);
    const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
    c = 1.0 / sqrt(1.0 + t * t);
    s = t * c;
  }
  if (threadIdx.x == 0) rot[blockIdx.x] = make_double2(c, s);
  if (s == 0.0) return;
  double* ap = A + (size_t)p * ld;
  double* aq = A + (size_t)q * ld;
  double* vp = V + (size_t)p * ld;
  double* vq = V + (size_t)q * ld;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double x = ap[i], y = aq[i];
    ap[i] = c * x - s * y;
    aq[i] = s * x + c * y;
    const double u = vp[i], w = vq[i];
    vp[i] = c * u - s * w;
    vq[i] = s * u + c * w;
  }
}

// one CTA per column j: rows p, q of every pair rotated in shared memory
__global__ void jacobi_rows_kernel(double* __restrict__ A, int n, int ne, int ld, int step, const double2* __restrict__ rot) {
  extern __shared__ double col[];
  const int j = blockIdx.x;
  double* a = A + (size_t)j * ld;
  for (int i = threadIdx.x; i < n; i += blockDim.x) col[i] = a[i];
  __syncthreads();
  for (int k = threadIdx.x; k < ne / 2; k += blockDim.x) {
    int p, q;
    jacobi_pair(step, k, ne, p, q);
    if (q >= n) continue;
    const double2 cs = rot[k];
    if (cs.y == 0.0) continue;
    const double x = col[p], y = col[q];
    col[p] = cs.x * x - cs.y * y;
    col[q] = cs.y * x + cs.x * y;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) a[i] = col[i];
}

// out[0] = sum of squared off-diagonal entries, out[1] = sum of squared diagonal entries (fixed-order block partials)
__global__ void offdiag_norm_kernel(const double* __restrict__ A, int n, int ld, double* __restrict__ part) {
  __shared__ double sh[256];
  const int j = blockIdx.x;
  double off = 0.0, dg = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = A[(size_t)i + (size_t)j * ld];
    if (i == j) dg += v * v;
    else off += v * v;
  }
  off = block_sum(off, sh);
  dg = block_sum(dg, sh);
  if (threadIdx.x == 0) {
    part[2 * j] = off;
    part[2 * j + 1] = dg;
  }
}

__global__ void set_identity_kernel(double* __restrict__ V, int n, int ld) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i < n) V[(size_t)i + (size_t)j * ld] = (i == j) ? 1.0 : 0.0;
}

// full symmetric Ky = K (lower triangle valid) + diag(noise)
__global__ void ky_full_kernel(const double* __restrict__ K, const double* __restrict__ noise, int n, int ld, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i >= n) return;
  double v = (i >= j) ? K[(size_t)i + (size_t)j * ld] : K[(size_t)j + (size_t)i * ld];
  if (i == j) v += noise[i];
  out[(size_t)i + (size_t)j * ld] = v;
}

// C = A - B (n x n)
__global__ void mat_sub_kernel(const double* A, const double* B, int n, int ld, double* C) {   // C may alias A or B
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i < n) C[(size_t)i + (size_t)j * ld] = A[(size_t)i + (size_t)j * ld] - B[(size_t)i + (size_t)j * ld];
}

// row-sum norms for conv = ||Y - X||_inf / ||Y||_inf: rsum[i] = sum_j |Y_ij - X_ij|, rsum[n + i] = sum_j |Y_ij| (symmetric: column sums)
__global__ void inf_norm_rows_kernel(const double* __restrict__ Y, const double* __restrict__ X, int n, int ld, double* __restrict__ rsum) {
  __shared__ double sh[256];
  const int j = blockIdx.x;
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double y = Y[(size_t)i + (size_t)j * ld];
    a += fabs(y - X[(size_t)i + (size_t)j * ld]);
    b += fabs(y);
  }
  a = block_sum(a, sh);
  b = block_sum(b, sh);
  if (threadIdx.x == 0) {
    rsum[j] = a;
    rsum[n + j] = b;
  }
}

// eigenvalues = diag(A) -> d[]; and the diagonal of X for do2eigen's rescaling
__global__ void take_diag_kernel(const double* __restrict__ A, int n, int ld, double* __restrict__ d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) d[i] = A[(size_t)i + (size_t)i * ld];
}

// X <- D X D with D = sqrt(max(eps, odiag) / diag(X)) (do2eigen: keeps the diagonal of the projected matrix)
__global__ void rescale_diag_kernel(double* __restrict__ X, int n, int ld, const double* __restrict__ odiag, const double* __restrict__ xdiag,
                                    double eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i >= n) return;
  const double di = sqrt(fmax(eps, odiag[i]) / xdiag[i]), dj = sqrt(fmax(eps, odiag[j]) / xdiag[j]);
  X[(size_t)i + (size_t)j * ld] *= di * dj;
}

// Eigen-decomposition of the symmetric n x n matrix in A (destroyed: its diagonal ends as the eigenvalues), V = eigenvectors
// (columns).  `rot`: ne/2 double2, `part`: 2 n doubles of device scratch.
int jacobi_eigh(H* h, double* A, double* V, int n, int ld, double2* rot, double* part) {
  const int ne = (n + 1) & ~1;
  if (n == 1) {
    set_identity_kernel<<<dim3(1, 1), 32, 0, h->st>>>(V, 1, ld);
    h->launches++;
    return GPS_OK;
  }
  set_identity_kernel<<<dim3(ceil_div(n, 256), n), 256, 0, h->st>>>(V, n, ld);
  h->launches++;
  std::vector<double> hp((size_t)2 * n);
  const size_t smem = (size_t)n * sizeof(double);
  double prev_ratio = 1.0;
  for (int sweep = 0; sweep < 60; sweep++) {
    for (int step = 0; step < ne - 1; step++) {
      jacobi_cols_kernel<<<ne / 2, 256, 0, h->st>>>(A, V, n, ne, ld, step, rot);
      jacobi_rows_kernel<<<n, 256, smem, h->st>>>(A, n, ne, ld, step, rot);
    }
    h->launches += 2LL * (ne - 1);
    offdiag_norm_kernel<<<n, 256, 0, h->st>>>(A, n, ld, part);
    h->launches++;
    CK(h, cudaMemcpyAsync(hp.data(), part, sizeof(double) * 2 * n, cudaMemcpyDeviceToHost, h->st));
    CK(h, cudaStreamSynchronize(h->st));
    double off = 0.0, dg = 0.0;
    for (int j = 0; j < n; j++) {
      off += hp[2 * j];
      dg += hp[2 * j + 1];
    }
    // converged: off-diagonal mass below (n eps)^2 of the squared Frobenius norm (never tighter than 1e-15 of the norm) — or
    // stagnated at the rounding noise the rotations themselves leave behind (n = 2000 with a thousand-fold zero eigenvalue
    // can stall above the first bound): the spectrum is then as accurate as it gets, which is far inside nearPD's own eig.tol = 1e-6
    const double ratio = off / (off + dg);
    const double tol = std::max(1e-30, (double)n * (double)n * 4.93e-32);
    if (!(ratio > tol)) return GPS_OK;
    if (sweep >= 6 && ratio <= 1e-20 && ratio > 0.25 * prev_ratio) return GPS_OK;
    prev_ratio = ratio;
  }
  // no usable spectrum: the caller treats the point like any other that cannot be repaired (ForOptimisationLog.R:50, stop())
  {
    char buf[256];
    snprintf(buf, sizeof buf, "covariance matrix is not positive definite; nearPD: the Jacobi eigen-solver did not converge in 60 sweeps "
                              "(off-diagonal share of the squared Frobenius norm %.3g)", prev_ratio);
    return fail(h, GPS_ERR_NOT_PD, buf);
  }
}

// X = V diag(w) V' (full symmetric, n x n) on the DMMA engine: lower tiles of the weighted product, then mirrored
void spectral_rebuild(H* v, const double* V, const double* w, double* X, int n) {
  GemmParams p = gp(V, v->ld, 0, V, v->ld, 0, n, n, n);
  p.lower = 1;
  p.kscale = w;
  p.strideScale = 0;
  gemm_ks(v, p, plain(X, v->ld, 0, 1.0, 0.0));
  symmetrize_kernel<<<dim3(ceil_div(n, 32), ceil_div(n, 32), 1), dim3(32, 8), 0, v->st>>>(X, n, v->ld, 0);
  v->seq_launches++;
}

// ForOptimisationLog.R:43-51 for fit b of the batch, after its Cholesky failed: Ky <- nearPD(Ky), chol, alpha, NLML
// (and, when asked, the gradient contraction with the inverse of the repaired matrix).  On success d_nlml[b] (d_grad[b])
// hold the values and the fit's pivot is cleared.
int nearpd_recover(H* h, int b, bool want_grad) {
  const int n = h->n, ld = h->ld;
  if (n > NEARPD_MAX_N) return fail(h, GPS_ERR_NOT_PD, "covariance matrix is not positive definite (nearPD on the device is limited to n <= 4096)");
  H v;
  make_view(h, b, 1, h->st, &v);
  v.groups = 1;
  // the failed fit's own work matrices are free: R / the Jacobi iterate in Ky, eigenvectors in L, X in U, D_S in Kinv, Y in Linv
  double *R = v.Ky, *V = v.L, *X = v.U, *DS = v.Kinv, *Y = v.Linv;
  const size_t mat_bytes = sizeof(double) * (size_t)h->strideM;
  double *vec = nullptr;         // [0, ldw): eigenvalue weights (zero padded), [ldw, 2 ldw): odiag, [2 ldw, 3 ldw): xdiag, then 2 n norms
  double2* rot = nullptr;
  const int ldw = round_up(n + 16, 16);
  auto cleanup = [&]() { cudaStreamSynchronize(h->st); cudaFree(vec); cudaFree(rot); };
#define NP_CK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { cleanup(); return fail(h, GPS_ERR_CUDA, std::string("nearPD: ") + cudaGetErrorString(e__)); } } while (0)
  NP_CK(dalloc(&vec, (size_t)3 * ldw + 2 * n));
  NP_CK(dalloc(&rot, (size_t)(n / 2 + 1)));
  double* part = vec + 3 * ldw;
  const dim3 g2(ceil_div(n, 256), n);
  std::vector<double> hd(n), hw(ldw, 0.0), hn((size_t)2 * n);
  h->cu_err = cudaSuccess;
  // x = Ky (full, symmetric by construction: ensureSymmetry is a no-op); X <- x, D_S <- 0
  ky_full_kernel<<<g2, 256, 0, h->st>>>(v.K, v.noise_diag, n, ld, X);
  NP_CK(cudaMemsetAsync(DS, 0, mat_bytes, h->st));
  h->launches++;
  bool converged = false;
  int iter = 0;
  double lam_max = 0.0;
  auto eig_weights = [&](double* Amat, bool clip, double* eps_out) -> int {   // eigen-decompose Amat (destroyed) -> V, weights in vec
    int rc = jacobi_eigh(h, Amat, V, n, ld, rot, part);
    if (rc) return rc;
    take_diag_kernel<<<ceil_div(n, 256), 256, 0, h->st>>>(Amat, n, ld, vec);
    h->launches++;
    if (cudaMemcpyAsync(hd.data(), vec, sizeof(double) * n, cudaMemcpyDeviceToHost, h->st) != cudaSuccess ||
        cudaStreamSynchronize(h->st) != cudaSuccess)
      return fail(h, GPS_ERR_CUDA, "nearPD: copy of the eigenvalues failed");
    lam_max = hd[0];
    double lam_min = hd[0];
    for (int i = 1; i < n; i++) {
      lam_max = std::max(lam_max, hd[i]);
      lam_min = std::min(lam_min, hd[i]);
    }
    std::fill(hw.begin(), hw.end(), 0.0);
    if (!clip) {                      // Higham step: keep eigenvalues > eig.tol * lambda_max, DROP the others
      bool any = false;
      for (int i = 0; i < n; i++)
        if (hd[i] > 1e-6 * lam_max) { hw[i] = hd[i]; any = true; }
      if (!any) return fail(h, GPS_ERR_NOT_PD, "nearPD: matrix seems negative semi-definite");
    } else {                          // do2eigen: raise eigenvalues below posd.tol * lambda_max to that value
      const double eps = 1e-8 * fabs(lam_max);
      *eps_out = eps;
      if (!(lam_min < eps)) return -1;                  // nothing to do
      for (int i = 0; i < n; i++) hw[i] = hd[i] < eps ? eps : hd[i];
    }
    if (cudaMemcpyAsync(vec, hw.data(), sizeof(double) * ldw, cudaMemcpyHostToDevice, h->st) != cudaSuccess)
      return fail(h, GPS_ERR_CUDA, "nearPD: upload of the weights failed");
    return GPS_OK;
  };
  while (iter < 100 && !converged) {
    NP_CK(cudaMemcpyAsync(Y, X, mat_bytes, cudaMemcpyDeviceToDevice, h->st));                 // Y <- X
    mat_sub_kernel<<<g2, 256, 0, h->st>>>(Y, DS, n, ld, R);                                   // R <- Y - D_S
    NP_CK(cudaMemcpyAsync(DS, R, mat_bytes, cudaMemcpyDeviceToDevice, h->st));                // keep R (the Jacobi iteration destroys it)
    h->launches++;
    int rc = eig_weights(R, false, nullptr);
    if (rc) { cleanup(); return rc; }
    v.seq_launches = 0;
    spectral_rebuild(&v, V, vec, X, n);                                                       // X <- Q_p diag(d_p) Q_p'
    h->launches += v.seq_launches;
    mat_sub_kernel<<<g2, 256, 0, h->st>>>(X, DS, n, ld, DS);                                  // D_S <- X - R
    inf_norm_rows_kernel<<<n, 256, 0, h->st>>>(Y, X, n, ld, part);
    h->launches += 2;
    NP_CK(cudaMemcpyAsync(hn.data(), part, sizeof(double) * 2 * n, cudaMemcpyDeviceToHost, h->st));
    NP_CK(cudaStreamSynchronize(h->st));
    double num = 0.0, den = 0.0;
    for (int i = 0; i < n; i++) {
      num = std::max(num, hn[i]);
      den = std::max(den, hn[n + i]);
    }
    iter++;
    converged = (num / den) <= 1e-7;
  }
  if (!converged) { cleanup(); return fail(h, GPS_ERR_NOT_PD, "covariance matrix is not positive definite; generation of a nearby p.d. matrix failed (nearPD did not converge)"); }
  {  // do2eigen
    take_diag_kernel<<<ceil_div(n, 256), 256, 0, h->st>>>(X, n, ld, vec + ldw);               // o.diag
    NP_CK(cudaMemcpyAsync(R, X, mat_bytes, cudaMemcpyDeviceToDevice, h->st));
    h->launches++;
    double eps = 0.0;
    int rc = eig_weights(R, true, &eps);
    if (rc > 0) { cleanup(); return rc; }
    if (rc == GPS_OK) {
      v.seq_launches = 0;
      spectral_rebuild(&v, V, vec, X, n);
      h->launches += v.seq_launches;
      take_diag_kernel<<<ceil_div(n, 256), 256, 0, h->st>>>(X, n, ld, vec + 2 * ldw);
      rescale_diag_kernel<<<g2, 256, 0, h->st>>>(X, n, ld, vec + ldw, vec + 2 * ldw, eps);
      h->launches += 2;
    }
  }
  // Ky <- the repaired matrix (+ identity padding and the appended right-hand-side rows), then the usual sequence
  NP_CK(cudaMemsetAsync(v.Ky, 0, mat_bytes, h->st));
  NP_CK(cudaMemcpy2DAsync(v.Ky, (size_t)ld * sizeof(double), X, (size_t)ld * sizeof(double), (size_t)n * sizeof(double), n,
                          cudaMemcpyDeviceToDevice, h->st));
  NP_CK(cudaMemsetAsync(v.L, 0, mat_bytes, h->st));
  NP_CK(cudaMemsetAsync(v.Linv, 0, mat_bytes, h->st));
  NP_CK(cudaMemsetAsync(v.U, 0, mat_bytes, h->st));
  NP_CK(cudaMemsetAsync(v.info, 0, sizeof(int), h->st));
  v.seq_launches = 0;
  fill_aug_kernel<<<dim3(ceil_div(v.ne, 128), 1), 128, 0, h->st>>>(v.Xaug, n, v.ne, v.d, v.ldx, v.strideX, v.meanw, v.ldmu,
                                                                   v.o.mean_form == 1, v.y, v.ld, v.meanv, v.Ky, v.ld, v.strideM, v.nrhs);
  v.seq_launches++;
  enqueue_potrf(&v);
  nlml_finish_kernel<<<1, 256, 0, h->st>>>(v.L, n, v.ne, v.ld, v.strideM, v.nrhs, v.zvec, v.ld, v.d_nlml, v.d_logdet);
  v.seq_launches++;
  if (want_grad) {
    enqueue_inverse(&v);
    enqueue_grad(&v);
  }
  h->launches += v.seq_launches;
  int info = 0;
  NP_CK(cudaMemcpyAsync(&info, v.info, sizeof(int), cudaMemcpyDeviceToHost, h->st));
  NP_CK(cudaStreamSynchronize(h->st));
#undef NP_CK
  cleanup();
  if (v.cu_err != cudaSuccess || h->cu_err != cudaSuccess) return fail(h, GPS_ERR_CUDA, v.err.empty() ? h->err : v.err);
  if (info != 0) return fail(h, GPS_ERR_NOT_PD, "covariance matrix is not positive definite even after nearPD");
  return GPS_OK;
}

}  // namespace
