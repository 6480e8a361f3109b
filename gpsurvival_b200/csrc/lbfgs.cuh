// lbfgs.cuh — device-resident limited-memory BFGS for all fits of a batched handle ("L-BFGS using the
// new gradient"; call site in the reference: optim() at GPRegression.R:28-31).
//
// vmmin (R's 'BFGS', restated in fit_batch.cuh) keeps a dense len x len inverse-Hessian per fit: 8.2 GB of state at
// C4 x 256 fits, streamed twice per tick, plus a host round trip per matrix-vector product.  Here the state is the
// last LBFGS_M (s, y) pairs per fit — m x len doubles — and EVERYTHING an iteration needs lives on the device: points,
// gradients, directions, step lengths, phases.  A tick is
//     lbfgs_dir -> lbfgs_trial (writes the trial points straight into the handle's `par` array)
//     -> the NLML graph (no host upload)  -> lbfgs_accept
//     -> [if any fit accepted its point: the inverse + gradient graphs -> lbfgs_update]
// and the host reads back four phase counters (16 bytes).  Line-search trials are fn-only, as in vmmin; the
// gradient is computed on acceptance, from the factorisation the accepted trial left behind.
// The state machine is optim.lbfgs_steps (gpsurvival_b200/optim.py) statement for statement; tests compare the two.
// Included by gpsurv.cu after fit_batch.cuh.
#pragma once

namespace {

constexpr int LBFGS_M = 10;
enum { LB_NEED_DIR = 0, LB_LINESEARCH = 1, LB_ACCEPTED = 2, LB_DONE = 3 };

struct LbfgsState {      // device arrays; [B], [B][len] or [B][LBFGS_M][len]
  double *x, *g, *t, *q, *S, *Y, *rho, *al, *f, *fnew, *step, *gradproj, *gamma;
  int *phase, *hist_n, *hist_head, *iters, *nf, *fail, *flags;
};

// x = par, g = grad, f = nlml (+Inf when the start did not factorise): vmmin's "initial value is not finite" ends the fit
__global__ void lbfgs_init_kernel(LbfgsState st, const double* __restrict__ par, const double* __restrict__ grad,
                                  const double* __restrict__ nlml, const int* __restrict__ info, const int* __restrict__ active,
                                  int len) {
  const int b = blockIdx.x;
  for (int k = threadIdx.x; k < len; k += blockDim.x) {
    st.x[(size_t)b * len + k] = par[(size_t)b * len + k];
    st.g[(size_t)b * len + k] = grad[(size_t)b * len + k];
  }
  if (threadIdx.x == 0) {
    const double f = info[b] ? INFINITY : nlml[b];
    st.f[b] = f;
    st.hist_n[b] = st.hist_head[b] = 0;
    st.iters[b] = 1;
    st.nf[b] = 1;
    st.gamma[b] = 1.0;
    st.fail[b] = isfinite(f) ? 0 : 2;
    st.phase[b] = (active[b] && isfinite(f)) ? LB_NEED_DIR : LB_DONE;
  }
}

// two-loop recursion t = -H g over the stored pairs (newest first, then oldest first), H0 = gamma I
__global__ void lbfgs_dir_kernel(LbfgsState st, int len) {
  __shared__ double sh[256];
  const int b = blockIdx.x, tid = threadIdx.x;
  if (st.phase[b] != LB_NEED_DIR) return;
  const double* g = st.g + (size_t)b * len;
  double* q = st.q + (size_t)b * len;
  double* t = st.t + (size_t)b * len;
  int hn = st.hist_n[b];
  const int head = st.hist_head[b];
  double gp = 0.0;
  for (int attempt = 0; attempt < 2; attempt++) {
    for (int k = tid; k < len; k += blockDim.x) q[k] = g[k];
    __syncthreads();
    for (int i = 0; i < hn; i++) {                       // newest -> oldest
      const int slot = (head - 1 - i + 2 * LBFGS_M) % LBFGS_M;
      const double* s = st.S + ((size_t)b * LBFGS_M + slot) * len;
      const double* y = st.Y + ((size_t)b * LBFGS_M + slot) * len;
      double d = 0.0;
      for (int k = tid; k < len; k += blockDim.x) d += s[k] * q[k];
      d = block_sum(d, sh);
      const double a = st.rho[b * LBFGS_M + slot] * d;
      if (tid == 0) st.al[b * LBFGS_M + slot] = a;
      for (int k = tid; k < len; k += blockDim.x) q[k] -= a * y[k];
      __syncthreads();
    }
    const double gam = hn > 0 ? st.gamma[b] : 1.0;
    for (int k = tid; k < len; k += blockDim.x) q[k] *= gam;
    __syncthreads();
    for (int i = hn - 1; i >= 0; i--) {                  // oldest -> newest
      const int slot = (head - 1 - i + 2 * LBFGS_M) % LBFGS_M;
      const double* s = st.S + ((size_t)b * LBFGS_M + slot) * len;
      const double* y = st.Y + ((size_t)b * LBFGS_M + slot) * len;
      double d = 0.0;
      for (int k = tid; k < len; k += blockDim.x) d += y[k] * q[k];
      d = block_sum(d, sh);
      const double c = st.al[b * LBFGS_M + slot] - st.rho[b * LBFGS_M + slot] * d;
      for (int k = tid; k < len; k += blockDim.x) q[k] += c * s[k];
      __syncthreads();
    }
    double d = 0.0;
    for (int k = tid; k < len; k += blockDim.x) {
      t[k] = -q[k];
      d += -q[k] * g[k];
    }
    gp = block_sum(d, sh);
    if (gp < 0.0 || hn == 0) break;
    hn = 0;                                              // not a descent direction: forget the history, steepest descent
  }
  if (tid == 0) {
    st.hist_n[b] = hn;
    st.gradproj[b] = gp;
    if (gp < 0.0) {
      st.step[b] = 1.0;
      st.phase[b] = LB_LINESEARCH;
    } else {
      st.phase[b] = LB_DONE;
    }
  }
}

// trial point par = x + step t of every fit in its line search (the others re-evaluate x: harmless, keeps par finite)
__global__ void lbfgs_trial_kernel(LbfgsState st, double* __restrict__ par, int len) {
  __shared__ double sh[256];
  const int b = blockIdx.x, tid = threadIdx.x;
  const double* x = st.x + (size_t)b * len;
  double* p = par + (size_t)b * len;
  const int ph = st.phase[b];
  if (ph != LB_LINESEARCH) {
    for (int k = tid; k < len; k += blockDim.x) p[k] = x[k];
    return;
  }
  const double* t = st.t + (size_t)b * len;
  const double step = st.step[b];
  double same = 0.0;
  for (int k = tid; k < len; k += blockDim.x) {
    const double nb = x[k] + step * t[k];
    p[k] = nb;
    if (10.0 + x[k] == 10.0 + nb) same += 1.0;           // vmmin's reltest
  }
  same = block_sum(same, sh);
  if ((int)same == len) {                                // no coordinate moves any more
    for (int k = tid; k < len; k += blockDim.x) p[k] = x[k];
    if (tid == 0) {
      if (st.hist_n[b] > 0) {                            // restart once from steepest descent (vmmin: reset B = I)
        st.hist_n[b] = 0;
        st.phase[b] = LB_NEED_DIR;
      } else {
        st.phase[b] = LB_DONE;
      }
    }
  }
}

// Armijo test of the trial values; flags[phase]++ tells the host what the next tick needs
__global__ void lbfgs_accept_kernel(LbfgsState st, const double* __restrict__ nlml, const int* __restrict__ info, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int ph = st.phase[b];
  if (ph == LB_LINESEARCH) {
    const double fn = info[b] ? INFINITY : nlml[b];
    st.nf[b]++;
    if (isfinite(fn) && fn <= st.f[b] + st.gradproj[b] * st.step[b] * 1.0e-4) {
      st.fnew[b] = fn;
      st.phase[b] = ph = LB_ACCEPTED;
    } else {
      st.step[b] *= 0.2;
    }
  }
  atomicAdd(&st.flags[ph], 1);
}

// after the gradient at the accepted points: s, y, history, stopping rules
__global__ void lbfgs_update_kernel(LbfgsState st, const double* __restrict__ par, const double* __restrict__ grad, int len,
                                    int maxit) {
  __shared__ double sh[256];
  const int b = blockIdx.x, tid = threadIdx.x;
  if (st.phase[b] != LB_ACCEPTED) return;
  double* x = st.x + (size_t)b * len;
  double* g = st.g + (size_t)b * len;
  const int head = st.hist_head[b];
  double* s = st.S + ((size_t)b * LBFGS_M + head) * len;
  double* y = st.Y + ((size_t)b * LBFGS_M + head) * len;
  double sy = 0.0, yy = 0.0;
  for (int k = tid; k < len; k += blockDim.x) {
    const double pk = par[(size_t)b * len + k], gk = grad[(size_t)b * len + k];
    const double sk = pk - x[k], yk = gk - g[k];
    s[k] = sk;
    y[k] = yk;
    x[k] = pk;
    g[k] = gk;
    sy += sk * yk;
    yy += yk * yk;
  }
  sy = block_sum(sy, sh);
  yy = block_sum(yy, sh);
  if (tid == 0) {
    const double f = st.f[b], fb = st.fnew[b];
    const bool enough = fabs(fb - f) > 1.4901161193847656e-08 * (fabs(f) + 1.4901161193847656e-08);
    st.f[b] = fb;
    const int it = ++st.iters[b];
    int hn = st.hist_n[b];
    const bool had = hn > 0;
    if (sy > 0.0) {
      st.rho[b * LBFGS_M + head] = 1.0 / sy;
      st.gamma[b] = sy / yy;
      st.hist_head[b] = (head + 1) % LBFGS_M;
      hn = hn < LBFGS_M ? hn + 1 : LBFGS_M;
    } else {
      hn = 0;
    }
    int ph = LB_NEED_DIR;
    if (it >= maxit) {
      st.fail[b] = 1;
      ph = LB_DONE;
    } else if (!enough) {
      if (had) hn = 0;
      else ph = LB_DONE;
    }
    st.hist_n[b] = hn;
    st.phase[b] = ph;
  }
}

// All fits' L-BFGS runs in lock-step; `ev` only counts ticks (the evaluations run from device-resident points).
int run_lbfgs(FitEvaluator& ev, const std::vector<double>& starts, int len, int maxit, const std::vector<char>& active,
              std::vector<OptResult>& out) {
  H* h = ev.h;
  const int B = h->batch;
  out.assign(B, OptResult());
  if (maxit <= 0) {
    int rc = ev.eval(starts.data(), false);
    if (rc) return rc;
    for (int i = 0; i < B; i++)
      if (active[i]) { out[i].par.assign(&starts[(size_t)i * len], &starts[(size_t)i * len] + len); out[i].value = ev.f[i]; out[i].ran = true; }
    return GPS_OK;
  }
  {
    int rc0 = grad_precheck(h);
    if (rc0) return rc0;
  }
  LbfgsState st;
  memset(&st, 0, sizeof st);
  int* d_active = nullptr;
  std::vector<void*> owned;
  auto cleanup = [&]() {
    cudaStreamSynchronize(h->st);
    for (void* p : owned) cudaFree(p);
    h->fact_level = 0;            // the points were set on the device: nothing the host-side caches know about
    h->fact_par.clear();
  };
#define LB_CK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { cleanup(); return fail(h, GPS_ERR_CUDA, std::string("gps_fit_batch (L-BFGS): ") + cudaGetErrorString(e__)); } } while (0)
#define LB_ALLOC(ptr, elems) do { LB_CK(dalloc(&(ptr), (size_t)(elems))); owned.push_back((void*)(ptr)); } while (0)
  const size_t BL = (size_t)B * len;
  LB_ALLOC(st.x, BL); LB_ALLOC(st.g, BL); LB_ALLOC(st.t, BL); LB_ALLOC(st.q, BL);
  LB_ALLOC(st.S, BL * LBFGS_M); LB_ALLOC(st.Y, BL * LBFGS_M);
  LB_ALLOC(st.rho, (size_t)B * LBFGS_M); LB_ALLOC(st.al, (size_t)B * LBFGS_M);
  LB_ALLOC(st.f, B); LB_ALLOC(st.fnew, B); LB_ALLOC(st.step, B); LB_ALLOC(st.gradproj, B); LB_ALLOC(st.gamma, B);
  LB_ALLOC(st.phase, B); LB_ALLOC(st.hist_n, B); LB_ALLOC(st.hist_head, B); LB_ALLOC(st.iters, B); LB_ALLOC(st.nf, B);
  LB_ALLOC(st.fail, B); LB_ALLOC(st.flags, 4); LB_ALLOC(d_active, B);
  {
    std::vector<int> act(B);
    for (int i = 0; i < B; i++) act[i] = active[i] ? 1 : 0;
    LB_CK(cudaMemcpyAsync(d_active, act.data(), sizeof(int) * B, cudaMemcpyHostToDevice, h->st));
    LB_CK(cudaStreamSynchronize(h->st));
  }
  // start: one full evaluation at the start points (the only host -> device upload of points)
  h->fact_level = 0;
  int rc = upload_par(h, starts.data(), len);
  if (rc) { cleanup(); return rc; }
  rc = run_full(h);
  if (rc) { cleanup(); return rc; }
  ev.ticks++;
  lbfgs_init_kernel<<<B, 256, 0, h->st>>>(st, h->par, h->d_grad, h->d_nlml, h->info, d_active, len);
  h->launches++;
  int* flags = h->pin_info;       // pinned, >= B ints... only 4 are needed; allocate separately when the batch is tiny
  int* flags_own = nullptr;
  if (B < 4) {
    LB_CK(cudaMallocHost((void**)&flags_own, 4 * sizeof(int)));
    flags = flags_own;
  }
  auto free_flags = [&]() { if (flags_own) cudaFreeHost(flags_own); };
  for (long long guard = 0; guard < 200LL * (maxit + 2); guard++) {
    for (int round = 0; round < 2; round++) {          // a "no move" restart gets its steepest-descent trial in the same tick
      lbfgs_dir_kernel<<<B, 256, 0, h->st>>>(st, len);
      lbfgs_trial_kernel<<<B, 256, 0, h->st>>>(st, h->par, len);
      h->launches += 2;
    }
    LB_CK(cudaMemsetAsync(st.flags, 0, 4 * sizeof(int), h->st));
    rc = run_factor(h);                                // NLML of every trial point, points already on the device
    if (rc) { free_flags(); cleanup(); return rc; }
    ev.ticks++;
    lbfgs_accept_kernel<<<ceil_div(B, 128), 128, 0, h->st>>>(st, h->d_nlml, h->info, B);
    h->launches++;
    LB_CK(cudaMemcpyAsync(flags, st.flags, 4 * sizeof(int), cudaMemcpyDeviceToHost, h->st));
    LB_CK(cudaStreamSynchronize(h->st));
    if (flags[LB_ACCEPTED] > 0) {
      rc = run_inverse(h);
      if (!rc) rc = run_grad(h);
      if (rc) { free_flags(); cleanup(); return rc; }
      lbfgs_update_kernel<<<B, 256, 0, h->st>>>(st, h->par, h->d_grad, len, maxit);
      h->launches++;
    } else if (flags[LB_LINESEARCH] == 0 && flags[LB_NEED_DIR] == 0) {
      break;                                           // every fit is done
    }
  }
  free_flags();
  std::vector<double> xs(BL), fs(B);
  std::vector<int> fl(B);
  LB_CK(cudaMemcpyAsync(xs.data(), st.x, sizeof(double) * BL, cudaMemcpyDeviceToHost, h->st));
  LB_CK(cudaMemcpyAsync(fs.data(), st.f, sizeof(double) * B, cudaMemcpyDeviceToHost, h->st));
  LB_CK(cudaMemcpyAsync(fl.data(), st.fail, sizeof(int) * B, cudaMemcpyDeviceToHost, h->st));
  LB_CK(cudaStreamSynchronize(h->st));
#undef LB_ALLOC
#undef LB_CK
  cleanup();
  for (int i = 0; i < B; i++)
    if (active[i]) {
      out[i].par.assign(&xs[(size_t)i * len], &xs[(size_t)i * len] + len);
      out[i].value = fs[i];
      out[i].fail = fl[i];
      out[i].ran = true;
    }
  return GPS_OK;
}

}  // namespace
