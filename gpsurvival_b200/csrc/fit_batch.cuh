// fit_batch.cuh — C-side whole fits: ApplyGP.R:136-282 (GPS outer loop, final predictions) with
// stats::optim restated in C++ (Nash 1990: Algorithm 19 = R's nmmin, Algorithm 21 = R's vmmin; R's
// defaults alpha=1, beta=0.5, gamma=2, abstol=-Inf, reltol=sqrt(DBL_EPSILON)), for ALL fits of a
// batched handle in lock-step: every tick serves the next evaluation point of every fit with ONE
// batched NLML(+gradient) evaluation.  Nelder-Mead keeps one small state machine per fit on the
// host; BFGS keeps the dense inverse-Hessian approximations of all fits on the DEVICE (matvec and
// rank-3 update kernels below) and vmmin's control flow as per-fit masks on the host.
// This file is included by gpsurv.cu (it uses the handle internals); it is the C++ twin of
// gpsurvival_b200/{optim,batch_bfgs,batchfit}.py, against which the tests compare it.
#pragma once

namespace {

constexpr double NM_BIG = 1.0e35;
const double OPT_RELTOL = 1.4901161193847656e-08;   // sqrt(.Machine$double.eps)

inline double fin_or_big(double f) { return std::isfinite(f) ? f : NM_BIG; }

// ---------------------------------------------------------------------------------------------
// Nelder-Mead (nmmin) as a resumable state machine: point() is the next point to evaluate,
// feed(f) consumes its value.  Mirrors optim.py:nelder_mead_steps line by line.
// ---------------------------------------------------------------------------------------------
struct NelderMead {
  int n = 0, maxit = 0, evals = 0, lo = 0, hi = 0, jbuild = 0, fail = 0;
  double convtol = 0, size = 0, oldsize = 0, vl = 0, vh = 0, fr = 0;
  bool done = false, need_vertices = false;
  enum St { INIT, BUILD, REFLECT, EXPAND, CONTRACT } st = INIT;
  std::vector<double> V, F, x, cen, xr, xe, xc, par;
  double value = 0;

  void start(const double* p, int n_, int maxit_) {
    n = n_; maxit = maxit_;
    x.assign(p, p + n);
    V.assign((size_t)(n + 1) * n, 0.0);
    F.assign(n + 1, 0.0);
    cen.assign(n, 0.0); xr = xe = xc = cen;
    st = INIT; done = false; evals = 0; fail = 0;
  }
  const double* point() const { return x.data(); }
  double* vert(int j) { return V.data() + (size_t)j * n; }

  void finish() {
    value = F[lo];
    par.assign(vert(lo), vert(lo) + n);
    if (evals > maxit) fail = 1;
    done = true;
  }
  bool next_build() {           // request the next vertex that needs a value; false when none is left
    while (jbuild <= n && jbuild == lo) jbuild++;
    if (jbuild > n) return false;
    x.assign(vert(jbuild), vert(jbuild) + n);
    st = BUILD;
    return true;
  }
  void loop_top() {
    for (;;) {
      if (need_vertices) {
        jbuild = 0;
        if (next_build()) return;
        need_vertices = false;
      }
      vl = F[lo]; vh = vl; hi = lo;
      for (int j = 0; j <= n; j++)
        if (j != lo) {
          if (F[j] < vl) { lo = j; vl = F[j]; }
          if (F[j] > vh) { hi = j; vh = F[j]; }
        }
      if (vh <= vl + convtol || vl <= -INFINITY) { finish(); return; }
      for (int i = 0; i < n; i++) {
        double t = -vert(hi)[i];
        for (int j = 0; j <= n; j++) t += vert(j)[i];
        cen[i] = t / n;
      }
      for (int i = 0; i < n; i++) xr[i] = 2.0 * cen[i] - vert(hi)[i];     // (1 + alpha) c - alpha h, alpha = 1
      x = xr;
      st = REFLECT;
      return;
    }
  }
  void iter_end() {
    if (evals > maxit) { finish(); return; }
    loop_top();
  }
  void feed(double f) {
    switch (st) {
      case INIT: {
        if (maxit <= 0) { value = f; par = x; done = true; return; }
        if (!std::isfinite(f)) { value = f; par = x; fail = 2; done = true; return; }   // R: error()
        evals = 1;
        convtol = OPT_RELTOL * (fabs(f) + OPT_RELTOL);
        double step = 0.0;
        for (int i = 0; i < n; i++) step = std::max(step, 0.1 * fabs(x[i]));
        if (step == 0.0) step = 0.1;
        for (int j = 0; j <= n; j++) std::copy(x.begin(), x.end(), vert(j));
        F[0] = f;
        size = 0.0;
        for (int j = 1; j <= n; j++) {
          double trystep = step;
          while (vert(j)[j - 1] == x[j - 1]) { vert(j)[j - 1] = x[j - 1] + trystep; trystep *= 10.0; }
          size += trystep;
        }
        oldsize = size;
        lo = 0;
        need_vertices = true;
        loop_top();
        return;
      }
      case BUILD:
        F[jbuild] = fin_or_big(f);
        evals++;
        jbuild++;
        if (next_build()) return;
        need_vertices = false;
        loop_top();
        return;
      case REFLECT:
        fr = fin_or_big(f);
        evals++;
        if (fr < vl) {
          for (int i = 0; i < n; i++) xe[i] = 2.0 * xr[i] + (1.0 - 2.0) * cen[i];     // gamma = 2
          x = xe;
          st = EXPAND;
        } else {
          if (fr < vh) { std::copy(xr.begin(), xr.end(), vert(hi)); F[hi] = fr; }
          for (int i = 0; i < n; i++) xc[i] = (1.0 - 0.5) * vert(hi)[i] + 0.5 * cen[i];   // beta = 0.5
          x = xc;
          st = CONTRACT;
        }
        return;
      case EXPAND: {
        double fe = fin_or_big(f);
        evals++;
        if (fe < fr) { std::copy(xe.begin(), xe.end(), vert(hi)); F[hi] = fe; }
        else { std::copy(xr.begin(), xr.end(), vert(hi)); F[hi] = fr; }
        iter_end();
        return;
      }
      case CONTRACT: {
        double fc = fin_or_big(f);
        evals++;
        if (fc < F[hi]) {
          std::copy(xc.begin(), xc.end(), vert(hi));
          F[hi] = fc;
        } else if (fr >= vh) {
          need_vertices = true;
          size = 0.0;
          for (int j = 0; j <= n; j++)
            if (j != lo)
              for (int i = 0; i < n; i++) {
                vert(j)[i] = 0.5 * (vert(j)[i] - vert(lo)[i]) + vert(lo)[i];
                size += fabs(vert(j)[i] - vert(lo)[i]);
              }
          if (size < oldsize) oldsize = size;
          else { fail = 10; finish(); return; }
        }
        iter_end();
        return;
      }
    }
  }
};

// ---------------------------------------------------------------------------------------------
// Device side of the batched BFGS: B[b] (len x len, symmetric) for every fit.
// ---------------------------------------------------------------------------------------------
__global__ void bfgs_matvec_kernel(const double* __restrict__ Bm, const double* __restrict__ v, double* __restrict__ out,
                                   int len, long long strideB) {
  int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31, b = blockIdx.y;
  if (i >= len) return;
  const double* row = Bm + (size_t)b * strideB + (size_t)i * len;
  const double* vb = v + (size_t)b * len;
  double s = 0.0;
  for (int j = lane; j < len; j += 32) s += row[j] * vb[j];
  s = warp_sum(s);
  if (lane == 0) out[(size_t)b * len + i] = s;
}

__global__ void bfgs_identity_kernel(double* __restrict__ Bm, int len, long long strideB, const int* __restrict__ mask) {
  int b = blockIdx.y;
  if (!mask[b]) return;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)len * len) return;
  Bm[(size_t)b * strideB + idx] = (idx / len == idx % len) ? 1.0 : 0.0;
}

// B += U V' with U, V = [3][len] per fit (zero coefficients for fits that do not update)
__global__ void bfgs_rank3_kernel(double* __restrict__ Bm, int len, long long strideB, const double* __restrict__ U,
                                  const double* __restrict__ V, const int* __restrict__ mask) {
  int b = blockIdx.y;
  if (!mask[b]) return;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)len * len) return;
  int i = (int)(idx / len), j = (int)(idx % len);
  const double* u = U + (size_t)b * 3 * len;
  const double* w = V + (size_t)b * 3 * len;
  Bm[(size_t)b * strideB + idx] += u[i] * w[j] + u[len + i] * w[len + j] + u[2 * len + i] * w[2 * len + j];
}

struct FitEvaluator {          // one batched evaluation of all fits at P [B][len]
  H* h;
  int len;
  long long ticks = 0;
  std::vector<double> f, g;
  int eval(const double* P, bool want_grad) {
    const int B = h->batch;
    f.assign(B, 0.0);
    int rc;
    if (want_grad) {
      g.assign((size_t)B * len, 0.0);
      rc = gps_nlml_grad(h, P, len, f.data(), g.data());
    } else {
      rc = gps_nlml(h, P, len, f.data());
    }
    ticks++;
    if (rc == GPS_ERR_NOT_PD) rc = GPS_OK;    // +Inf marks the rejected fits (the reference would try nearPD)
    return rc;
  }
};

struct OptResult {
  std::vector<double> par;
  double value = 0;
  int fail = 0;
  bool ran = false;
};

// All fits' Nelder-Mead runs in lock-step.
int run_nelder_mead(FitEvaluator& ev, const std::vector<double>& starts, int len, int maxit, const std::vector<char>& active,
                    std::vector<OptResult>& out) {
  const int B = ev.h->batch;
  std::vector<NelderMead> nm(B);
  std::vector<double> P(starts);
  out.assign(B, OptResult());
  int live = 0;
  for (int b = 0; b < B; b++)
    if (active[b]) { nm[b].start(&starts[(size_t)b * len], len, maxit); live++; }
  while (live > 0) {
    for (int b = 0; b < B; b++)
      if (active[b] && !nm[b].done) std::copy(nm[b].point(), nm[b].point() + len, &P[(size_t)b * len]);
    int rc = ev.eval(P.data(), false);
    if (rc) return rc;
    for (int b = 0; b < B; b++)
      if (active[b] && !nm[b].done) {
        nm[b].feed(ev.f[b]);
        if (nm[b].done) {
          live--;
          out[b].par = nm[b].par; out[b].value = nm[b].value; out[b].fail = nm[b].fail; out[b].ran = true;
          std::copy(nm[b].par.begin(), nm[b].par.end(), &P[(size_t)b * len]);
        }
      }
  }
  return GPS_OK;
}

// All fits' BFGS runs in lock-step (twin of batch_bfgs.py): dense state on the device.
int run_bfgs(FitEvaluator& ev, const std::vector<double>& starts, int len, int maxit, const std::vector<char>& active,
             std::vector<OptResult>& out) {
  H* h = ev.h;
  const int B = h->batch, n = len;
  const double stepredn = 0.2, acctol = 1.0e-4, reltest = 10.0;
  enum { NEED_DIR = 0, LINESEARCH = 1, DONE = 2 };
  out.assign(B, OptResult());
  std::vector<double> b(starts), X(starts), g((size_t)B * n), c((size_t)B * n), t((size_t)B * n, 0.0), tmp((size_t)B * n);
  std::vector<double> f(B), fmin(B), step(B, 1.0), gradproj(B, 0.0), U((size_t)B * 3 * n), Vv((size_t)B * 3 * n);
  std::vector<int> nf(B, 1), ng(B, 1), it(B, 1), ilast(B, 1), count(B, 0), phase(B), mask(B);
  const long long strideB = (long long)n * n;
  double *dB = nullptr, *dv = nullptr, *dout = nullptr, *dU = nullptr, *dV = nullptr;
  int* dmask = nullptr;
  auto cleanup = [&]() { cudaStreamSynchronize(h->st); cudaFree(dB); cudaFree(dv); cudaFree(dout); cudaFree(dU); cudaFree(dV); cudaFree(dmask); };
#define FB_CK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { cleanup(); return fail(h, GPS_ERR_CUDA, std::string("gps_fit_batch: ") + cudaGetErrorString(e__)); } } while (0)
  FB_CK(dalloc(&dB, (size_t)strideB * B, false));
  FB_CK(dalloc(&dv, (size_t)B * n));
  FB_CK(dalloc(&dout, (size_t)B * n));
  FB_CK(dalloc(&dU, (size_t)B * 3 * n));
  FB_CK(dalloc(&dV, (size_t)B * 3 * n));
  FB_CK(dalloc(&dmask, (size_t)B));
  auto matvec = [&](const std::vector<double>& v, std::vector<double>& o) -> cudaError_t {   // o = B v for every fit
    cudaError_t e = cudaMemcpyAsync(dv, v.data(), sizeof(double) * B * n, cudaMemcpyHostToDevice, h->st);
    if (e != cudaSuccess) return e;
    bfgs_matvec_kernel<<<dim3(ceil_div(n, 8), B), 256, 0, h->st>>>(dB, dv, dout, n, strideB);
    h->launches++;
    e = cudaMemcpyAsync(o.data(), dout, sizeof(double) * B * n, cudaMemcpyDeviceToHost, h->st);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(h->st);
  };
  auto set_identity = [&]() -> cudaError_t {       // for the fits flagged in mask
    cudaError_t e = cudaMemcpyAsync(dmask, mask.data(), sizeof(int) * B, cudaMemcpyHostToDevice, h->st);
    if (e != cudaSuccess) return e;
    bfgs_identity_kernel<<<dim3((unsigned)((strideB + 255) / 256), B), 256, 0, h->st>>>(dB, n, strideB, dmask);
    h->launches++;
    return cudaStreamSynchronize(h->st);
  };

  if (maxit <= 0) {
    int rc = ev.eval(starts.data(), false);
    cleanup();
    if (rc) return rc;
    for (int i = 0; i < B; i++)
      if (active[i]) { out[i].par.assign(&starts[(size_t)i * n], &starts[(size_t)i * n] + n); out[i].value = ev.f[i]; out[i].ran = true; }
    return GPS_OK;
  }
  int rc = ev.eval(b.data(), true);
  if (rc) { cleanup(); return rc; }
  for (int i = 0; i < B; i++) {
    f[i] = ev.f[i];
    fmin[i] = f[i];
    phase[i] = active[i] ? NEED_DIR : DONE;
    if (active[i] && !std::isfinite(f[i])) phase[i] = DONE;     // R: "initial value in 'vmmin' is not finite"
  }
  g = ev.g;
  c = g;
  auto end_of_iteration = [&](int i) {
    if (it[i] >= maxit) { phase[i] = DONE; return; }
    if (ng[i] - ilast[i] > 2 * n) ilast[i] = ng[i];
    phase[i] = (count[i] == n && ilast[i] == ng[i]) ? DONE : NEED_DIR;
  };
  auto any_phase = [&](int p) { for (int i = 0; i < B; i++) if (phase[i] == p) return true; return false; };
  auto any_not_done = [&]() { for (int i = 0; i < B; i++) if (phase[i] != DONE) return true; return false; };

  while (any_not_done()) {
    int guard = 0;
    while (any_phase(NEED_DIR) && guard++ < 5) {
      bool any_reset = false;
      for (int i = 0; i < B; i++) { mask[i] = (phase[i] == NEED_DIR && ilast[i] == ng[i]) ? 1 : 0; any_reset |= mask[i] != 0; }
      if (any_reset) FB_CK(set_identity());
      FB_CK(matvec(g, tmp));
      for (int i = 0; i < B; i++) {
        if (phase[i] != NEED_DIR) continue;
        double gp = 0.0;
        for (int k = 0; k < n; k++) {
          X[(size_t)i * n + k] = b[(size_t)i * n + k];
          c[(size_t)i * n + k] = g[(size_t)i * n + k];
          t[(size_t)i * n + k] = -tmp[(size_t)i * n + k];
          gp += t[(size_t)i * n + k] * g[(size_t)i * n + k];
        }
        gradproj[i] = gp;
        if (gp < 0.0) {
          step[i] = 1.0;
          phase[i] = LINESEARCH;
        } else {
          count[i] = 0;
          if (ilast[i] == ng[i]) count[i] = n;
          else ilast[i] = ng[i];
          end_of_iteration(i);
        }
      }
    }
    if (!any_phase(LINESEARCH)) continue;
    std::vector<int> finished, need;
    for (int i = 0; i < B; i++) {
      if (phase[i] != LINESEARCH) continue;
      int same = 0;
      for (int k = 0; k < n; k++) {
        double xo = X[(size_t)i * n + k];
        double nb = xo + step[i] * t[(size_t)i * n + k];
        b[(size_t)i * n + k] = nb;
        if (reltest + xo == reltest + nb) same++;
      }
      count[i] = same;
      if (same < n) need.push_back(i);
      else finished.push_back(i);
    }
    if (!need.empty()) {
      rc = ev.eval(b.data(), true);
      if (rc) { cleanup(); return rc; }
      for (int i : need) {
        f[i] = ev.f[i];
        nf[i]++;
        bool acc = std::isfinite(f[i]) && f[i] <= fmin[i] + gradproj[i] * step[i] * acctol;
        if (acc) finished.push_back(i);
        else step[i] *= stepredn;
      }
    }
    std::vector<int> upd;
    for (int i : finished) {
      bool enough = (f[i] > -INFINITY) && fabs(f[i] - fmin[i]) > OPT_RELTOL * (fabs(fmin[i]) + OPT_RELTOL);
      if (!enough) { count[i] = n; fmin[i] = f[i]; }
      if (count[i] < n) { fmin[i] = f[i]; upd.push_back(i); }
      else if (ilast[i] < ng[i]) { count[i] = 0; ilast[i] = ng[i]; }
    }
    if (!upd.empty()) {
      bool any_pos = false;
      std::fill(mask.begin(), mask.end(), 0);
      for (int i : upd) {
        ng[i]++; it[i]++;
        double D1 = 0.0;
        for (int k = 0; k < n; k++) {
          size_t q = (size_t)i * n + k;
          g[q] = ev.g[q];
          t[q] = step[i] * t[q];
          c[q] = g[q] - c[q];
          D1 += t[q] * c[q];
        }
        if (D1 > 0) { mask[i] = 1; any_pos = true; }
        else ilast[i] = ng[i];
      }
      if (any_pos) {
        FB_CK(matvec(c, tmp));                                      // X = B c
        std::fill(U.begin(), U.end(), 0.0);
        std::fill(Vv.begin(), Vv.end(), 0.0);
        for (int i = 0; i < B; i++) {
          if (!mask[i]) continue;
          double D1 = 0.0, xc_ = 0.0;
          for (int k = 0; k < n; k++) {
            size_t q = (size_t)i * n + k;
            D1 += t[q] * c[q];
            xc_ += tmp[q] * c[q];
          }
          double D2 = 1.0 + xc_ / D1;
          for (int k = 0; k < n; k++) {
            size_t q = (size_t)i * n + k, u0 = (size_t)i * 3 * n + k;
            U[u0] = D2 / D1 * t[q];        Vv[u0] = t[q];
            U[u0 + n] = -tmp[q] / D1;      Vv[u0 + n] = t[q];
            U[u0 + 2 * n] = -t[q] / D1;    Vv[u0 + 2 * n] = tmp[q];
            X[q] = tmp[q];
          }
        }
        FB_CK(cudaMemcpyAsync(dU, U.data(), sizeof(double) * B * 3 * n, cudaMemcpyHostToDevice, h->st));
        FB_CK(cudaMemcpyAsync(dV, Vv.data(), sizeof(double) * B * 3 * n, cudaMemcpyHostToDevice, h->st));
        FB_CK(cudaMemcpyAsync(dmask, mask.data(), sizeof(int) * B, cudaMemcpyHostToDevice, h->st));
        bfgs_rank3_kernel<<<dim3((unsigned)((strideB + 255) / 256), B), 256, 0, h->st>>>(dB, n, strideB, dU, dV, dmask);
        h->launches++;
        FB_CK(cudaStreamSynchronize(h->st));
      }
    }
    for (int i : finished) end_of_iteration(i);
  }
#undef FB_CK
  cleanup();
  for (int i = 0; i < B; i++)
    if (active[i]) {
      out[i].par.assign(&b[(size_t)i * n], &b[(size_t)i * n] + n);
      out[i].value = fmin[i];
      out[i].fail = it[i] < maxit ? 0 : 1;
      out[i].ran = true;
    }
  return GPS_OK;
}

}  // namespace
