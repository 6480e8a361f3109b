// dmma_lab.cu — DMMA microbenchmarks for sm_100a (developer tool; not part of the library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o lab/dmma_lab lab/dmma_lab.cu -lcublas
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>
#include "../gpsurvival_b200/csrc/dmma_ws.cuh"
using namespace gps;

#define CHECK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e__), __LINE__); exit(1);} } while (0)

// ---- 1. register-only DMMA issue rate -------------------------------------------------
template <int ILP>
__global__ void reg_dmma(double* out, int iters) {
  double acc[ILP][2];
#pragma unroll
  for (int i = 0; i < ILP; i++) acc[i][0] = acc[i][1] = 0.0;
  double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3 + 1.0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) dmma884(acc[i][0], acc[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += acc[i][0] + acc[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- 2. smem-fed DMMA loop without global traffic (same fragment pattern as the engine) ----
template <int FM, int FN, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) smem_dmma(double* out, int iters) {
  constexpr int BM = 64, BN = 64, BK = 16, LD = 68;
  __shared__ double sA[BK * LD], sB[BK * LD];
  for (int i = threadIdx.x; i < BK * LD; i += blockDim.x) { sA[i] = i * 1e-4; sB[i] = 1.0 - i * 1e-4; }
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  double acc[FM][FN][2];
#pragma unroll
  for (int i = 0; i < FM; i++)
#pragma unroll
    for (int j = 0; j < FN; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int ks = 0; ks < BK / 4; ks++) {
      double af[FM], bf[FN];
#pragma unroll
      for (int i = 0; i < FM; i++) af[i] = sA[(4 * ks + t) * LD + (8 * i + g) % BM];
#pragma unroll
      for (int j = 0; j < FN; j++) bf[j] = sB[(4 * ks + t) * LD + (8 * j + g) % BN];
#pragma unroll
      for (int i = 0; i < FM; i++)
#pragma unroll
        for (int j = 0; j < FN; j++) dmma884(acc[i][j][0], acc[i][j][1], bf[j], af[i]);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < FM; i++)
#pragma unroll
    for (int j = 0; j < FN; j++) s += acc[i][j][0] + acc[i][j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static float time_ms(cudaEvent_t e0, cudaEvent_t e1) { float ms; cudaEventElapsedTime(&ms, e0, e1); return ms; }

template <int ILP>
void run_reg(double* out, int warps_per_sm, int sms) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  // one CTA per SM with warps_per_sm warps
  reg_dmma<ILP><<<sms, warps_per_sm * 32>>>(out, 100);
  cudaEventRecord(e0);
  reg_dmma<ILP><<<sms, warps_per_sm * 32>>>(out, iters);
  cudaEventRecord(e1);
  CHECK(cudaDeviceSynchronize());
  double fl = 2.0 * 256 * ILP * (double)iters * warps_per_sm * sms;
  printf("reg  ILP=%2d warps/SM=%2d : %7.3f ms  %6.2f TF/s\n", ILP, warps_per_sm, time_ms(e0, e1), fl / time_ms(e0, e1) / 1e9);
}

template <int FM, int FN, int WARPS>
void run_smem(double* out, int ctas_per_sm, int sms) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 4000;
  smem_dmma<FM, FN, WARPS><<<sms * ctas_per_sm, WARPS * 32>>>(out, 10);
  cudaEventRecord(e0);
  smem_dmma<FM, FN, WARPS><<<sms * ctas_per_sm, WARPS * 32>>>(out, iters);
  cudaEventRecord(e1);
  CHECK(cudaDeviceSynchronize());
  double fl = 2.0 * 256 * FM * FN * 4.0 * iters * WARPS * ctas_per_sm * sms;
  printf("smem FM=%d FN=%d warps/CTA=%d CTAs/SM=%d : %7.3f ms  %6.2f TF/s\n", FM, FN, WARPS, ctas_per_sm, time_ms(e0, e1),
         fl / time_ms(e0, e1) / 1e9);
}

template <class Cfg, bool AKC, bool BKC>
void run_gemm(const char* name, double* A, double* B, double* C, int M, int N, int K, int batch = 1) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  GemmParams p; memset(&p, 0, sizeof p);
  p.A = A; p.B = B; p.M = M; p.N = N; p.K = K; p.lda = AKC ? K : M; p.ldb = BKC ? K : N; p.ksplit = 1; p.nodes = 1;
  p.strideA = (long long)M * K; p.strideB = (long long)N * K;
  EpiPlain e{}; e.C = C; e.ldc = M; e.strideC = (long long)M * N; e.strideSplit = 0; e.nodeStrideC = 0; e.alpha = 1; e.beta = 0;
  e.lower_strict_mask = 0; e.diag_off = 0;
  for (int w = 0; w < 2; w++) CHECK((launch_gemm<Cfg, AKC, BKC, EpiPlain>(p, e, batch, 0)));
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dmma_gemm_kernel<Cfg, AKC, BKC, EpiPlain>, Cfg::NT, SmemLayout<Cfg, AKC, BKC>::BYTES);
  const int reps = 5;
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    cudaEventRecord(e0);
    launch_gemm<Cfg, AKC, BKC, EpiPlain>(p, e, batch, 0);
    cudaEventRecord(e1);
    CHECK(cudaDeviceSynchronize());
    float ms = time_ms(e0, e1); if (ms < best) best = ms;
  }
  printf("gemm %-28s %dx%dx%d b=%d occ=%d : %8.3f ms  %6.2f TF/s\n", name, M, N, K, batch, occ, best, 2.0 * M * N * K * batch / best / 1e9);
}


static int* g_sched = nullptr;
// ---- warp-specialised kernel: timing + bitwise comparison with the 64x64 kernel -------------------
template <class Cfg, bool AKC, bool BKC>
float time_ws(GemmParams p, EpiPlain e, int batch, int reps = 4) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  CHECK((launch_gemm_ws<Cfg, AKC, BKC, EpiPlain>(p, e, batch, 148, g_sched, 0)));
  CHECK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    cudaEventRecord(e0);
    launch_gemm_ws<Cfg, AKC, BKC, EpiPlain>(p, e, batch, 148, g_sched, 0);
    cudaEventRecord(e1);
    CHECK(cudaDeviceSynchronize());
    float ms = time_ms(e0, e1); if (ms < best) best = ms;
  }
  return best;
}
template <bool AKC, bool BKC>
float time_old(GemmParams p, EpiPlain e, int batch, int reps = 4) {
  using C3 = TileCfg<64, 64, 16, 2, 2, 3>;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  CHECK((launch_gemm<C3, AKC, BKC, EpiPlain>(p, e, batch, 0)));
  CHECK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    cudaEventRecord(e0);
    launch_gemm<C3, AKC, BKC, EpiPlain>(p, e, batch, 0);
    cudaEventRecord(e1);
    CHECK(cudaDeviceSynchronize());
    float ms = time_ms(e0, e1); if (ms < best) best = ms;
  }
  return best;
}
// flops counted per 64x64 tile actually computed are hard to state generically: report times and the ratio
template <bool AKC, bool BKC>
void compare(const char* name, GemmParams p, EpiPlain e, int batch, double* C1, double* C2, size_t celems, double flops, bool lower_only = false) {
  // beta != 0 accumulates: reset both outputs to the same pattern before the single checked launch
  std::vector<double> h1(celems), h2(celems);
  for (size_t i = 0; i < celems; i++) h1[i] = (double)((i * 40503u) % 997) / 997.0;
  CHECK(cudaMemcpy(C1, h1.data(), celems * 8, cudaMemcpyHostToDevice));
  CHECK(cudaMemcpy(C2, h1.data(), celems * 8, cudaMemcpyHostToDevice));
  using C3 = TileCfg<64, 64, 16, 2, 2, 3>;
  EpiPlain e1 = e, e2 = e; e1.C = C1 + (e.C - (double*)nullptr); e2.C = C2 + (e.C - (double*)nullptr);
  CHECK((launch_gemm<C3, AKC, BKC, EpiPlain>(p, e1, batch, 0)));
  CHECK((launch_gemm_ws<WsCfg128, AKC, BKC, EpiPlain>(p, e2, batch, 148, g_sched, 0)));
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(err)); exit(1); }
  CHECK(cudaMemcpy(h1.data(), C1, celems * 8, cudaMemcpyDeviceToHost));
  CHECK(cudaMemcpy(h2.data(), C2, celems * 8, cudaMemcpyDeviceToHost));
  double md = 0; size_t nd = 0; double mx = 0;
  for (size_t i = 0; i < celems; i++) { if (lower_only) { size_t w = i % ((size_t)e.ldc * e.ldc); if (w % e.ldc < w / e.ldc) continue; } double d = fabs(h1[i] - h2[i]); if (d > 0 || (h1[i] != h1[i]) != (h2[i] != h2[i])) nd++; if (d > md) md = d; if (fabs(h1[i]) > mx) mx = fabs(h1[i]); }
  float t_old = 0, t_ws = 0;
  if (e.beta == 0.0) { t_old = time_old<AKC, BKC>(p, e1, batch); t_ws = time_ws<WsCfg128, AKC, BKC>(p, e2, batch); }
  printf("cmp %-34s maxdiff=%.3e ndiff=%zu (max|C|=%.3g)  old %.3f ms (%.2f TF/s)  ws %.3f ms (%.2f TF/s)\n", name, md, nd, mx, t_old,
         t_old > 0 ? flops / t_old / 1e9 : 0.0, t_ws, t_ws > 0 ? flops / t_ws / 1e9 : 0.0);
}

int main(int argc, char** argv) {
  int dev = 0; cudaSetDevice(dev);
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, dev);
  int sms = pr.multiProcessorCount;
  printf("device %s, %d SMs, clock %d kHz\n", pr.name, sms, pr.clockRate);
  CHECK(cudaMalloc(&g_sched, 8)); CHECK(cudaMemset(g_sched, 0, 8));
  double* out; CHECK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
  if (argc <= 1) {
  for (int w : {4, 8, 12, 16}) run_reg<16>(out, w, sms);
  for (int w : {4, 8, 16}) run_reg<8>(out, w, sms);
  for (int w : {4, 8}) run_reg<32>(out, w, sms);
  run_smem<4, 4, 4>(out, 1, sms); run_smem<4, 4, 4>(out, 2, sms); run_smem<4, 4, 4>(out, 3, sms); run_smem<4, 4, 4>(out, 4, sms);
  run_smem<8, 4, 4>(out, 1, sms); run_smem<8, 4, 4>(out, 2, sms);
  run_smem<4, 4, 8>(out, 1, sms); run_smem<4, 4, 8>(out, 2, sms);
  }

  const int n = 8192;
  double *A, *B, *C;
  CHECK(cudaMalloc(&A, sizeof(double) * n * n)); CHECK(cudaMalloc(&B, sizeof(double) * n * n)); CHECK(cudaMalloc(&C, sizeof(double) * n * n));
  {
    std::vector<double> hbuf((size_t)n * n);
    for (size_t i = 0; i < hbuf.size(); i++) hbuf[i] = (double)((i * 2654435761u) % 1000) / 1000.0 - 0.5;
    CHECK(cudaMemcpy(A, hbuf.data(), sizeof(double) * n * n, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(B, hbuf.data(), sizeof(double) * n * n, cudaMemcpyHostToDevice));
  }
  cublasHandle_t cb; cublasCreate(&cb);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int sz : {2048, 4096, 8192}) {
    for (int tr = 0; tr < 2; tr++) {
      double one = 1.0, zero = 0.0;
      float best = 1e30f;
      for (int r = 0; r < 6; r++) {
        cudaEventRecord(e0);
        cublasDgemm(cb, CUBLAS_OP_N, tr ? CUBLAS_OP_T : CUBLAS_OP_N, sz, sz, sz, &one, A, sz, B, sz, &zero, C, sz);
        cudaEventRecord(e1);
        CHECK(cudaDeviceSynchronize());
        float ms = time_ms(e0, e1); if (r > 0 && ms < best) best = ms;
      }
      printf("cublasDgemm N%c %d^3 : %8.3f ms  %6.2f TF/s\n", tr ? 'T' : 'N', sz, best, 2.0 * sz * sz * sz / best / 1e9);
    }
  }

  {
    double *C1, *C2; size_t cel = (size_t)8192 * 8192;
    CHECK(cudaMalloc(&C1, cel * 8)); CHECK(cudaMalloc(&C2, cel * 8));
    auto mk = [&](int M, int N, int K, int lda, int ldb) { GemmParams p; memset(&p, 0, sizeof p); p.A = A; p.B = B; p.M = M; p.N = N; p.K = K; p.lda = lda; p.ldb = ldb; p.ksplit = 1; p.nodes = 1; return p; };
    auto ep = [&](int ldc, long long sC, double alpha, double beta) { EpiPlain e{}; e.C = nullptr; e.ldc = ldc; e.strideC = sC; e.strideSplit = 0; e.nodeStrideC = 0; e.alpha = alpha; e.beta = beta; e.lower_strict_mask = 0; e.diag_off = 0; return e; };
    { GemmParams p = mk(2048, 2048, 2048, 2048, 2048); compare<false, false>("NT 2048^3", p, ep(2048, 0, 1, 0), 1, C1, C2, 2048 * 2048, 2.0 * 2048 * 2048 * 2048); }
    { GemmParams p = mk(8192, 8192, 8192, 8192, 8192); compare<false, false>("NT 8192^3", p, ep(8192, 0, 1, 0), 1, C1, C2, cel, 2.0 * 8192 * 8192 * 8192.0); }
    { GemmParams p = mk(2048, 2048, 2048, 2048, 2048); p.strideA = 2048 * 2048; p.strideB = 2048 * 2048; compare<false, false>("NT 2048^3 batch 4", p, ep(2048, 2048 * 2048, 1, 0), 4, C1, C2, (size_t)4 * 2048 * 2048, 4 * 2.0 * 2048 * 2048 * 2048); }
    { GemmParams p = mk(285, 285, 1000, 288, 288); compare<false, false>("NT 285x285x1000 (odd M,N)", p, ep(288, 0, 1, 0), 1, C1, C2, 288 * 288, 2.0 * 285 * 285 * 1000); }
    { GemmParams p = mk(1938, 976, 1024, 2016, 2016); p.lower = 1; p.strideA = 2016 * 2016; p.strideB = 2016 * 2016; compare<false, false>("NT lower upd beta=1 batch 3", p, ep(2016, 2016 * 2016, -1, 1), 3, C1, C2, (size_t)3 * 2016 * 2016, 0, true); }
    { GemmParams p = mk(1938, 976, 1024, 2016, 2016); p.lower = 1; p.strideA = 2016 * 2016; p.strideB = 2016 * 2016; compare<false, false>("NT lower upd beta=0 batch 8", p, ep(2016, 2016 * 2016, -1, 0), 8, C1, C2, (size_t)8 * 2016 * 2016, 8 * 2.0 * (1938.0 * 976 - 976.0 * 976 / 2) * 1024, true); }
    // NT-only forms used by the library (U = upper-triangular operand with ld 2016, batch 8)
    { double* Ut; size_t ue = (size_t)8 * 2016 * 2016; CHECK(cudaMalloc(&Ut, ue * 8));
      std::vector<double> hu(ue, 0.0);
      for (int b = 0; b < 8; b++) for (int c = 0; c < 2000; c++) for (int r = 0; r <= c; r++) hu[(size_t)b * 2016 * 2016 + r + (size_t)c * 2016] = (double)(((r * 31 + c * 17 + b) * 2654435761u) % 1000) / 1000.0 - 0.5;
      CHECK(cudaMemcpy(Ut, hu.data(), ue * 8, cudaMemcpyHostToDevice));
      { GemmParams p = mk(2000, 2000, 2000, 2016, 2016); p.A = Ut; p.B = Ut; p.lower = 1; p.klo_m = 1; p.strideA = p.strideB = 2016 * 2016;
        compare<false, false>("NT lower klo_m (Kinv=UU') b8", p, ep(2016, 2016 * 2016, 1, 0), 8, C1, C2, (size_t)8 * 2016 * 2016, 8 * 2000.0 * 2000 * 2000 / 3, true); }
      { int bs = 512, ne = 2000, ld = 2016; GemmParams p = mk(bs, bs, bs, ld, ld); p.A = Ut; p.B = B + bs; p.klo_m = 1; p.nodes = 2; p.nodeStrideA = p.nodeStrideB = (long long)2 * bs * (1 + ld);
        p.clip_total = ne - bs; p.clip_step = 2 * bs; p.k_eq_m = 0; p.clip_n = 1; p.strideA = p.strideB = ld * ld;
        EpiPlain e = ep(ld, ld * ld, 1, 0); e.nodeStrideC = p.nodeStrideA; compare<false, false>("NT klo_m nodes clip_n (trtri 1) b8", p, e, 8, C1, C2, (size_t)8 * ld * ld, 8 * 2 * 512.0 * 512 * 512); }
      cudaFree(Ut); }
    { GemmParams p = mk(1000, 1000, 1000, 1008, 1008); compare<true, true>("TN 1000^3 (K tail)", p, ep(1008, 0, 1, 0), 1, C1, C2, 1008 * 1008, 2.0 * 1000 * 1000 * 1000); }
    { GemmParams p = mk(1000, 70, 1000, 1008, 1008); compare<false, true>("NN 1000x70x1000", p, ep(1008, 0, 1, 0), 1, C1, C2, 1008 * 1008, 2.0 * 1000 * 70 * 1000); }

    {
      using W128 = WsCfg128;
      using W64 = WsCfg<64, 64, 16, 2, 2, 4, 3>;      // 4 consumer warps of 32x32, 3 CTAs/SM
      using W64s3 = WsCfg<64, 64, 16, 2, 2, 3, 3>;
      using W64s6 = WsCfg<64, 64, 16, 2, 2, 6, 3>;
      using W64k32 = WsCfg<64, 64, 32, 2, 2, 3, 3>;
      using W96 = WsCfg<96, 96, 16, 2, 2, 4, 2>;      // 4 consumer warps of 48x48, 2 CTAs/SM
      using W80 = WsCfg<80, 80, 16, 2, 2, 4, 2>;      // 4 consumer warps of 40x40
      using W128x64w8 = WsCfg<128, 64, 16, 4, 2, 4, 2>;   // 8 consumer warps of 32x32, 2 CTAs/SM
      struct Case { const char* nm; int M, N, K, batch, lower, klo_m; };
      Case cases[] = {{"NT 8192^3", 8192, 8192, 8192, 1, 0, 0}, {"NT 2048^3 b8", 2048, 2048, 2048, 8, 0, 0}, {"NT 2000 lower klo_m b16(Kinv)", 2000, 2000, 2000, 16, 1, 1},
                      {"NT upd 978x976x1024 lower b32", 978, 976, 1024, 32, 1, 0}, {"NT upd 1746x256x256 lower b32", 1746, 256, 256, 32, 1, 0}, {"NT upd 1874x128x128 lower b32", 1874, 128, 128, 32, 1, 0},
                      {"NT upd 1938x64x64 b32", 1938, 64, 64, 32, 1, 0},
                      {"NT 400x2001x400 b64 (C4 grad)", 400, 2001, 400, 64, 0, 0}, {"NT 400x400x2000 lower b64 (C4 gram)", 400, 400, 2000, 64, 1, 0},
                      {"NT 285x10001x285 b8 (C3 grad)", 285, 10001, 285, 8, 0, 0}, {"NT 285x285x10000 lower b8 (C3 gram)", 285, 285, 10000, 8, 1, 0}};
      for (auto& c : cases) {
        int ld = ((std::max(c.M, c.N) + 15) / 16) * 16;
        GemmParams p = mk(c.M, c.N, c.K, ld, ld); p.lower = c.lower; p.klo_m = c.klo_m;
        long long st = (long long)ld * c.K; p.strideA = p.strideB = st;
        if ((size_t)st * c.batch > cel || (size_t)ld * c.N * c.batch > cel) { printf("skip %s\n", c.nm); continue; }
        EpiPlain e = ep(ld, (long long)ld * c.N, 1, 0); e.C = C1;
        double fl = 2.0 * c.M * c.N * c.K * c.batch; if (c.lower) fl = 2.0 * (c.M * (double)c.N - (double)c.N * c.N / 2) * c.K * c.batch; if (c.klo_m) fl = 2.0 * c.M * (double)c.M * c.M / 6 * c.batch * 1.0;
        float t0 = time_old<false, false>(p, e, c.batch);
        float t1 = time_ws<W128, false, false>(p, e, c.batch);
        float t2 = time_ws<W64, false, false>(p, e, c.batch);
        float t3 = time_ws<W64s3, false, false>(p, e, c.batch);
        float t4 = time_ws<W64s6, false, false>(p, e, c.batch);
        float t5 = time_ws<W64k32, false, false>(p, e, c.batch);
        float t6 = time_ws<W96, false, false>(p, e, c.batch);
        float t7 = time_ws<W80, false, false>(p, e, c.batch);
        float t8 = c.lower ? 1e30f : time_ws<W128x64w8, false, false>(p, e, c.batch);
        printf("cfg %-36s alg TF/s: old64 %.2f | ws128 %.2f | ws64 %.2f | ws64s3 %.2f | ws64s6 %.2f | ws64k32 %.2f | ws96 %.2f | ws80 %.2f | ws128x64w8 %.2f\n", c.nm, fl / t0 / 1e9, fl / t1 / 1e9,
               fl / t2 / 1e9, fl / t3 / 1e9, fl / t4 / 1e9, fl / t5 / 1e9, fl / t6 / 1e9, fl / t7 / 1e9, fl / t8 / 1e9);
      }
    }
    cudaFree(C1); cudaFree(C2);
  }
  if (argc > 1) return 0;
  using C3 = TileCfg<64, 64, 16, 2, 2, 3>;
  using C4 = TileCfg<64, 64, 16, 2, 2, 4>;
  using C128x64 = TileCfg<128, 64, 16, 4, 2, 3>;
  using C128x64w4 = TileCfg<128, 64, 16, 2, 2, 3>;
  using C64x128w4 = TileCfg<64, 128, 16, 2, 2, 3>;
  using C128w8 = TileCfg<128, 128, 16, 2, 4, 3>;
  for (int sz : {2048, 8192}) {
    run_gemm<C3, false, false>("64x64x16 w4 s3 NT", A, B, C, sz, sz, sz);
    run_gemm<C4, false, false>("64x64x16 w4 s4 NT", A, B, C, sz, sz, sz);
    run_gemm<C128x64, false, false>("128x64x16 w8 s3 NT", A, B, C, sz, sz, sz);
    run_gemm<C128x64w4, false, false>("128x64x16 w4(64x32) s3 NT", A, B, C, sz, sz, sz);
    run_gemm<C64x128w4, false, false>("64x128x16 w4(32x64) s3 NT", A, B, C, sz, sz, sz);
    run_gemm<C128w8, false, false>("128x128x16 w8(64x32) s3 NT", A, B, C, sz, sz, sz);
    run_gemm<C3, false, true>("64x64x16 w4 s3 NN", A, B, C, sz, sz, sz);
    run_gemm<C3, true, true>("64x64x16 w4 s3 TN", A, B, C, sz, sz, sz);
  }
  run_gemm<C3, false, false>("64x64x16 w4 s3 NT batch", A, B, C, 2048, 2048, 2048, 8);
  run_gemm<C128x64, false, false>("128x64 w8 s3 NT batch", A, B, C, 2048, 2048, 2048, 8);
  return 0;
}
