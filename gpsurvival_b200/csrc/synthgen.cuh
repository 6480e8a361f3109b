// synthgen.cuh — the reference's synthetic-data recipe on the device:
//   MakeSyntheticData.R:40-41,145-149 ('likeRealData'):  x ~ N(0,1),  K = CovFunc(x, x),  y = exp(t(chol(K)) %*% d1 + sqrt(noise) d2)
//   CensorData.R:36-48 ('NormalLoopSample'):             censor randomly picked samples at |N(mean, sd)| times until nCensored
// so that Exp1-style repetitions (runSyntheticExp1.R:208-213: 20-30 data sets per setting) need no host data preparation.
// The reference draws from R's Mersenne-Twister seeded from the wall clock (runSyntheticExp1.R:64) — a stream nobody can
// reproduce — so the recipe defines its own: Philox4x32-10 (Salmon et al. 2011), counter = (block index, stream, data set,
// 0), key = seed; normals by Box-Muller from 53-bit uniforms.  K's Cholesky runs on the library's own factorisation.
// Included by gpsurv.cu.
#pragma once

namespace {

__device__ __forceinline__ void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1, unsigned* out) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ double u53(unsigned hi, unsigned lo) {   // (0, 1): 53 bits, never 0 or 1
  return ((double)(((unsigned long long)(hi >> 5) << 26) + (lo >> 6)) + 0.5) * (1.0 / 9007199254740992.0);
}

// out[b][i] (leading dimension ldo per data set... laid out by the caller through `stride`) = N(0,1) draws of `stream`
// grid (ceil(count/2 / 256), batch), block 256; one Philox block = two normals
__global__ void synth_normals_kernel(double* __restrict__ out, long long stride, int count, int rows, int ld, int stream,
                                     unsigned long long seed) {
  const int pair = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (2 * pair >= count) return;
  unsigned w[4];
  philox4x32_10((unsigned)pair, (unsigned)stream, (unsigned)b, 0u, (unsigned)seed, (unsigned)(seed >> 32), w);
  const double u1 = u53(w[0], w[1]), u2 = u53(w[2], w[3]);
  const double r = sqrt(-2.0 * log(u1));
  double sn, cs;
  sincos(2.0 * 3.141592653589793238462643383279502884 * u2, &sn, &cs);
  // element e of the flat (column-major, `rows` per column) array lands at column e / rows, row e % rows of an ld-strided matrix
  for (int q = 0; q < 2; q++) {
    const int e = 2 * pair + q;
    if (e < count) out[(size_t)b * stride + (size_t)(e / rows) * ld + (e % rows)] = q == 0 ? r * cs : r * sn;
  }
}

// y0 = exp(f + sqrt(sn2) d2)   (MakeSyntheticData.R:148-149, zero mean)
__global__ void synth_targets_kernel(const double* __restrict__ f, const double* __restrict__ d2, int n, int ldn, double sn2,
                                     double* __restrict__ y0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (i < n) y0[(size_t)b * ldn + i] = exp(f[(size_t)b * 2 * ldn + i] + sqrt(sn2) * d2[(size_t)b * ldn + i]);
}

// CensorData.R:36-48 'NormalLoopSample': sequential by nature (each draw depends on which samples are still uncensored):
// one thread per data set.  remaining: [batch][n] int scratch.
__global__ void synth_censor_kernel(const double* __restrict__ y0, int n, int ldn, int n_censor, double cmean, double csd,
                                    unsigned long long seed, double* __restrict__ y, double* __restrict__ events,
                                    int* __restrict__ remaining, int* __restrict__ done_out) {
  const int b = blockIdx.x;
  if (threadIdx.x != 0) return;
  int* rem = remaining + (size_t)b * n;
  double* yb = y + (size_t)b * ldn;
  double* eb = events + (size_t)b * ldn;
  for (int i = 0; i < n; i++) {
    rem[i] = i;
    yb[i] = y0[(size_t)b * ldn + i];
    eb[i] = 1.0;
  }
  int nrem = n, done = 0;
  const long long max_draws = 4000LL * n;
  for (long long draw = 0; done < n_censor && draw < max_draws; draw++) {
    unsigned w[4];
    philox4x32_10((unsigned)draw, 3u, (unsigned)b, 0u, (unsigned)seed, (unsigned)(seed >> 32), w);
    int slot = (int)(u53(w[0], w[1]) * (double)nrem);
    if (slot > nrem - 1) slot = nrem - 1;
    const int pick = rem[slot];                                        // sample(indicesToCensor, 1)
    const double u1 = u53(w[2], w[3]), u2 = u53(w[0] ^ 0x5BD1E995u, w[1] ^ 0x9E3779B9u);
    const double z = sqrt(-2.0 * log(u1)) * cos(2.0 * 3.141592653589793238462643383279502884 * u2);
    const double t = fabs(cmean + csd * z);                             // abs(rnorm(1, mean, sd))
    if (yb[pick] > t) {
      yb[pick] = t;
      eb[pick] = 0.0;
      for (int k = slot; k < nrem - 1; k++) rem[k] = rem[k + 1];        // indicesToCensor[-which(...)]: order kept
      nrem--;
      done++;
    }
  }
  done_out[b] = done;
}

}  // namespace
