// dmma_dag.cuh — the blocked Cholesky (and the inverse levels inside its diagonal blocks) as ONE persistent kernel.
//
// The multi-launch Cholesky (gpsurv.cu: potrf_big / potrf_range) is a chain of ~100 launches per group of fits in which
// every 256-wide block column starts with a latency-bound phase — four 64-column pivot chains, short-k products — that
// leaves the tensor pipes idle: measured on the bench workload (C2 x 64), 1.7 ms of a 21.1 ms step (GPS_DEBUG_SKIPDIAG).
// Stream groups, launch priorities and oversubscribed grids do not hide it (measured neutral): a persistent GEMM holds
// every SM until its last tile.  Here the same work is a LIST OF TASKS consumed in order by one persistent grid:
//   * a SEGMENT is what used to be one launch restricted to one group of fits: a batch of DMMA tiles with the
//     warp-specialised engine's decoding (dmma_ws.cuh), or one LEAF task per fit (factor + invert a 64 x 64 diagonal tile);
//   * a task may start when all tasks of its segment's PREDECESSOR (the previous segment of the same group) are complete
//     FOR ITS OWN FIT — a per-(segment, fit) counter in global memory, incremented with release semantics after the tile's
//     stores, polled with acquire semantics by the producer warp before it issues the tile's first copy.  Fits are
//     independent, so nothing ever waits for another fit: there is no launch boundary, no grid-wide barrier;
//   * the groups' segment lists are merged on the host by a discrete-event LIST SCHEDULER (gpsurv.cu, enqueue_potrf_dag): a
//     machine clock advanced by the bulk slices, a ready time per group advanced by the latency of its dependent steps; bulk
//     products are cut into slices of a few dozen tiles, so that while one group's fits sit in their pivot chains the list
//     offers the CTAs other groups' trailing-update tiles, and a dependent step is listed about when its predecessor is done.
// Status: EXPERIMENTAL, off by default (GPS_DAG=1; =2 appends the triangular-inverse levels to the same list).  Measured
// parity with the multi-launch path, not a win: a dependent hop through global counters costs 17 - 20 us where
// a launch boundary costs 2 - 3, and the default path's two stream groups already overlap most of what the list overlaps.
// GPS_DAG_TRACE=1 records per-task time stamps (scripts/dag_trace.py).
// Deadlock freedom: a task only ever waits for tasks that precede it in the list, tasks are handed out in list order,
// and a CTA that holds a task never waits for a later one.
// Determinism: which CTA runs a tile is dynamic, what a tile computes is not (same k order as the multi-launch path).
#pragma once
#include "dmma_ws.cuh"

namespace gps {

struct DagSeg {
  GemmParams p;          // GEMM segment: operands already advanced to the group's first fit
  EpiPlain epi;
  int kind;              // 0 = GEMM tiles, 1 = LEAF (one task per fit)
  int zcount;            // fits * nodes * ksplit
  int rev_m, rev_n;
  int tile_begin;        // first global task index
  int local0;            // first task of this SLICE within its product (a bulk product is listed as several slices)
  int ntasks;
  int dep_cnt;           // index of the predecessor's counters in done[] (-1: none); counter of fit f = done[dep_cnt + f]
  int dep_need;          // completed tasks per fit that release this segment
  int cnt;               // index of this segment's counters in done[]
  // LEAF: factor the nb x nb block at (e0, e0) of A, publish L11 -> L, L11^-1 -> Linv, its transpose -> Uinv
  const double* A;
  double *L, *Linv, *Uinv;
  int ld, e0, nb;
  long long stride;
  int* info;
};

namespace dag {

constexpr int LPS = NB + 4;    // leaf tile in shared memory, column-major: S[c * LPS + r]
constexpr int LPT = 32 + 4;    // merge temporary, column-major: sT[n * LPT + m]
constexpr int LEAF_SMEM_DOUBLES = NB * LPS + 32 * LPT + NB;

__device__ __forceinline__ long long gtime() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
  return t;
}
__device__ __forceinline__ int smid() {
  int r;
  asm volatile("mov.u32 %0, %%smid;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// trailing update after sub-panel Q on `nw` warps (potf2_trailing of kernels.cuh is written for 8)
template <int Q>
__device__ __forceinline__ void leaf_trailing(double* S, int warp, int nw, int g, int t) {
  constexpr int R = 16 * (Q + 1), NT8 = (NB - R) / 8;
  int tile = 0;
#pragma unroll 1
  for (int mi = 0; mi < NT8; mi++)
#pragma unroll 1
    for (int ni = 0; ni <= mi; ni++, tile++) {
      if ((tile % nw) != warp) continue;
      const int m0 = R + 8 * mi, n0 = R + 8 * ni;
      double2 c = *reinterpret_cast<double2*>(S + (n0 + g) * LPS + m0 + 2 * t);
#pragma unroll
      for (int ks = 0; ks < 4; ks++) {
        const int k = 16 * Q + 4 * ks + t;
        dmma_acc(c.x, c.y, -S[k * LPS + n0 + g], S[k * LPS + m0 + g]);
      }
      *reinterpret_cast<double2*>(S + (n0 + g) * LPS + m0 + 2 * t) = c;
    }
}

// One merge of the IN-PLACE blocked inverse (X overwrites L inside S, same column-major layout):
//   pass 0:  T   = L21 * X11          (sT, column-major)        L21 = S rows [R1, R1+SZ) x cols [C0, C0+SZ)
//   pass 1:  X21 = -X22 * T  -> written over L21
template <int SZ>
__device__ __forceinline__ void leaf_merge(double* S, double* sT, int R1, int C0, int pass, int w, int nw, int g, int t) {
  constexpr int NT8 = SZ / 8;
#pragma unroll 1
  for (int tile = w; tile < NT8 * NT8; tile += nw) {
    const int m0 = 8 * (tile % NT8), n0 = 8 * (tile / NT8);
    double c0 = 0.0, c1 = 0.0;
    if (pass == 0) {
#pragma unroll
      for (int ks = 0; ks < SZ / 4; ks++) {
        const int k = 4 * ks + t;
        dmma_acc(c0, c1, S[(C0 + n0 + g) * LPS + C0 + k], S[(C0 + k) * LPS + R1 + m0 + g]);     // X11[k][n], L21[m][k]
      }
      *reinterpret_cast<double2*>(sT + (n0 + g) * LPT + m0 + 2 * t) = make_double2(c0, c1);
    } else {
#pragma unroll
      for (int ks = 0; ks < SZ / 4; ks++) {
        const int k = 4 * ks + t;
        dmma_acc(c0, c1, -sT[(n0 + g) * LPT + k], S[(R1 + k) * LPS + R1 + m0 + g]);             // T[k][n], X22[m][k]
      }
      *reinterpret_cast<double2*>(S + (C0 + n0 + g) * LPS + R1 + m0 + 2 * t) = make_double2(c0, c1);
    }
  }
}

// LEAF task on all NTHR threads of the CTA (named barrier 2).  smem: LEAF_SMEM_DOUBLES doubles at `sm`.
template <int NTHR>
__device__ void leaf_task(const DagSeg& sg, int fit, double* sm) {
  auto bar = [] { asm volatile("bar.sync 2, %0;\n" ::"n"(NTHR) : "memory"); };
  constexpr int NW = NTHR / 32;
  double* S = sm;
  double* sT = S + NB * LPS;
  double* invd = sT + 32 * LPT;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int ld = sg.ld, e0 = sg.e0, nb = sg.nb;
  const double* Ab = sg.A + (size_t)fit * sg.stride + (size_t)e0 + (size_t)e0 * ld;
  for (int c = tid; c < NB * 32; c += NTHR) {            // A11, zero-filled to 64 x 64
    const int col = c >> 5, r = (c & 31) * 2;
    double2 v = make_double2(0.0, 0.0);
    if (col < nb && r < nb) v = *reinterpret_cast<const double2*>(Ab + (size_t)r + (size_t)col * ld);   // nb even
    *reinterpret_cast<double2*>(S + col * LPS + r) = v;
  }
  bar();
  if (tid >= nb && tid < NB) S[tid * LPS + tid] = 1.0;   // identity padding of a partial block
  bar();
  int bad = 0;
  if (warp == 0) potf2_subpanel<0>(S, invd, lane, bad);
  bar();
  leaf_trailing<0>(S, warp, NW, g, t);
  bar();
  if (warp == 0) potf2_subpanel<1>(S, invd, lane, bad);
  bar();
  leaf_trailing<1>(S, warp, NW, g, t);
  bar();
  if (warp == 0) potf2_subpanel<2>(S, invd, lane, bad);
  bar();
  leaf_trailing<2>(S, warp, NW, g, t);
  bar();
  if (warp == 0) potf2_subpanel<3>(S, invd, lane, bad);
  bar();
  if (tid == 0 && bad != 0 && bad <= nb) atomicCAS(&sg.info[fit], 0, e0 + bad);
  {                                                      // publish L11 before the inverse overwrites it
    double* Lb = sg.L + (size_t)fit * sg.stride + (size_t)e0 + (size_t)e0 * ld;
    for (int idx = tid; idx < NB * NB; idx += NTHR) {
      const int r = idx & 63, c = idx >> 6;
      if (r < nb && c < nb) Lb[(size_t)r + (size_t)c * ld] = (r >= c) ? S[c * LPS + r] : 0.0;
    }
  }
  bar();
  if (warp < 4 && lane < 16) {                           // 16 x 16 diagonal blocks inverted in place, lane = column
    const int o = 16 * warp, c = lane;
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = (i == c) ? 1.0 : 0.0;
#pragma unroll 1
    for (int k = 0; k < 16; k++) {
      const double xk = x[0] * invd[o + k];
      const double* col = S + (o + k) * LPS + o + k;
      double lc[16];
#pragma unroll
      for (int i = 1; i < 16; i++) lc[i] = col[i];       // column k of L below the diagonal (rows > k: not yet overwritten)
      __syncwarp(0x0000ffffu);                           // every lane has read before row k is overwritten
      S[(o + c) * LPS + o + k] = xk;                     // X[k][c]  (zero above the diagonal)
#pragma unroll
      for (int i = 1; i < 16; i++) x[i - 1] = x[i] - lc[i] * xk;
    }
  }
  bar();
  for (int node = 0; node < 2; node++) {                 // level 16: (0,1) and (2,3)
    leaf_merge<16>(S, sT, node * 32 + 16, node * 32, 0, warp, NW, g, t);
    bar();
    leaf_merge<16>(S, sT, node * 32 + 16, node * 32, 1, warp, NW, g, t);
    bar();
  }
  leaf_merge<32>(S, sT, 32, 0, 0, warp, NW, g, t);
  bar();
  leaf_merge<32>(S, sT, 32, 0, 1, warp, NW, g, t);
  bar();
  {
    double* Ib = sg.Linv + (size_t)fit * sg.stride + (size_t)e0 + (size_t)e0 * ld;
    double* Ub = sg.Uinv + (size_t)fit * sg.stride + (size_t)e0 + (size_t)e0 * ld;
    for (int idx = tid; idx < NB * NB; idx += NTHR) {
      const int r = idx & 63, c = idx >> 6;
      if (r < nb && c < nb) {
        Ib[(size_t)r + (size_t)c * ld] = (r >= c) ? S[c * LPS + r] : 0.0;
        Ub[(size_t)r + (size_t)c * ld] = (c >= r) ? S[r * LPS + c] : 0.0;     // U = X'
      }
    }
  }
  bar();
}

}  // namespace dag

template <class Cfg>
__global__ void __launch_bounds__(Cfg::NTHREADS, Cfg::CTAS_PER_SM)
    dmma_dag_kernel(const DagSeg* __restrict__ segs, int nseg, int total, int* __restrict__ sched, int* __restrict__ done,
                    long long* __restrict__ trace) {   // trace (developer): per task [picked, ready, done (ns), sm | leaf << 8 | K/16 << 9 | fit << 24 | segment << 32]
  using SL = WsSmem<Cfg, false, false>;
  constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, STAGES = Cfg::STAGES;
  constexpr int FM = Cfg::FM, FN = Cfg::FN;
  static_assert((size_t)STAGES * (SL::A_STAGE + SL::B_STAGE) >= (size_t)dag::LEAF_SMEM_DOUBLES, "the leaf works inside the stage buffers");
  extern __shared__ __align__(16) double smem[];
  double* sA = smem;
  double* sB = smem + STAGES * SL::A_STAGE;
  double* scratch = smem + STAGES * (SL::A_STAGE + SL::B_STAGE);
  double* s_w = scratch + SL::SCRATCH;
  unsigned long long* full = reinterpret_cast<unsigned long long*>(s_w + SL::WSC);
  unsigned long long* empty = full + STAGES;
  volatile int* stage_tile = reinterpret_cast<volatile int*>(empty + STAGES);
  __shared__ volatile int stage_seg[STAGES];

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const bool producer = (warp == Cfg::NT / 32);
  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) {
      ws::mbar_init(full + s, 32);
      ws::mbar_init(empty + s, Cfg::NT / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  unsigned it = 0;
  ws::TileWork w;
  int cur = 0;

  if (producer) {
    for (;;) {
      int t = 0;
      if (lane == 0) t = atomicAdd(&sched[0], 1);
      t = __shfl_sync(0xffffffffu, t, 0);
      if (t >= total) break;
      if (trace && lane == 0) trace[4 * (size_t)t] = dag::gtime();
      while (t >= segs[cur].tile_begin + segs[cur].ntasks) cur++;
      const DagSeg& sg = segs[cur];
      const int local = t - sg.tile_begin + sg.local0;
      if (sg.kind == 1) {                                 // LEAF: one task per fit, run by the whole CTA
        const int fit = local;
        if (sg.dep_cnt >= 0 && lane == 0)
          while (dag::ld_acquire(done + sg.dep_cnt + fit) < sg.dep_need) __nanosleep(100);
        __syncwarp();
        if (trace && lane == 0) trace[4 * (size_t)t + 1] = dag::gtime();
        const int s = it % STAGES;
        ws::mbar_wait(empty + s, ((it / STAGES) & 1) ^ 1);
        if (lane == 0) {
          stage_tile[s] = t;
          stage_seg[s] = cur;
        }
        ws::mbar_arrive_tx(full + s, 0);
        it++;
        // every earlier stage has been consumed when the consumers see this marker (they take stages in order) and
        // nothing is issued after it: the stage buffers are free for the leaf
        asm volatile("bar.sync 2, %0;\n" ::"n"(Cfg::NTHREADS) : "memory");
        dag::leaf_task<Cfg::NTHREADS>(sg, fit, smem);
        if (lane == 0) {
          __threadfence();
          atomicAdd(done + sg.cnt + fit, 1);
          if (trace) {
            trace[4 * (size_t)t + 2] = dag::gtime();
            trace[4 * (size_t)t + 3] = dag::smid() | (1 << 8) | ((long long)fit << 24) | ((long long)cur << 32);
          }
        }
        continue;
      }
      ws::TileEnum te;
      te.init(sg.p.M, sg.p.N, BM, BN, sg.p.lower, sg.rev_m, sg.rev_n);
      ws::decode_tile<BM, BN, BK>(sg.p, te, local, w);
      if (!w.valid) {                                     // outside its node's clip: counts as done
        if (lane == 0) atomicAdd(done + sg.cnt + w.batch, 1);
        continue;
      }
      if (sg.dep_cnt >= 0) {
        if (lane == 0)
          while (dag::ld_acquire(done + sg.dep_cnt + w.batch) < sg.dep_need) __nanosleep(100);
        __syncwarp();
        asm volatile("fence.proxy.async;\n" ::: "memory");   // the bulk copies below read what other CTAs' generic stores wrote
      }
      if (trace && lane == 0) trace[4 * (size_t)t + 1] = dag::gtime();
      const GemmParams& p = w.p;
      const int m0 = w.m0, n0 = w.n0;
      const double* __restrict__ A = p.A + (size_t)w.batch * p.strideA + (size_t)w.node * p.nodeStrideA;
      const double* __restrict__ B = p.B + (size_t)w.batch * p.strideB + (size_t)w.node * p.nodeStrideB;
      const int vm = min(p.M - m0, BM), vn = min(p.N - n0, BN);
      const int nst = w.nkt > 0 ? w.nkt : 1;
      for (int kt = 0; kt < nst; kt++, it++) {
        const int s = it % STAGES;
        ws::mbar_wait(empty + s, ((it / STAGES) & 1) ^ 1);
        if (lane == 0) {
          stage_tile[s] = t;
          stage_seg[s] = cur;
        }
        if (w.nkt == 0) {
          ws::mbar_arrive_tx(full + s, 0);
          continue;
        }
        const int kbase = w.k_lo + kt * BK;
        const int vk = min(max(p.K - kbase, 0), BK);
        double* a_s = sA + s * SL::A_STAGE;
        double* b_s = sB + s * SL::B_STAGE;
        unsigned bytes = 0;
        for (int r = lane; r < BK; r += 32) {
          if (r < vk) bytes += ((vm + 1) & ~1) * 8 + ((vn + 1) & ~1) * 8;
          else {
            ws::zero_row(a_s + r * SL::LDA, 0, BM);
            ws::zero_row(b_s + r * SL::LDB, 0, BN);
          }
        }
        if (vk < BK) asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        ws::mbar_arrive_tx(full + s, bytes);
        const unsigned nba = ((vm + 1) & ~1) * 8, nbb = ((vn + 1) & ~1) * 8;
        for (int r = lane; r < vk; r += 32) {
          ws::bulk_g2s(a_s + r * SL::LDA, A + (size_t)m0 + (size_t)(kbase + r) * p.lda, nba, full + s);
          ws::bulk_g2s(b_s + r * SL::LDB, B + (size_t)n0 + (size_t)(kbase + r) * p.ldb, nbb, full + s);
        }
      }
    }
    {
      const int s = it % STAGES;
      ws::mbar_wait(empty + s, ((it / STAGES) & 1) ^ 1);
      if (lane == 0) stage_tile[s] = -1;
      ws::mbar_arrive_tx(full + s, 0);
    }
    if (lane == 0) {
      __threadfence();
      const int fin = atomicAdd(&sched[1], 1);
      if (fin == (int)gridDim.x - 1) {
        sched[0] = 0;
        sched[1] = 0;
        __threadfence();
      }
    }
  } else {
    const int g = lane >> 2, t4 = lane & 3;
    const int wm0 = (warp % Cfg::WARPS_M) * Cfg::WM;
    const int wn0 = (warp / Cfg::WARPS_M) * Cfg::WN;
    for (;;) {
      int s = it % STAGES;
      ws::mbar_wait(full + s, (it / STAGES) & 1);
      const int t = stage_tile[s];
      if (t < 0) break;
      const int segi = stage_seg[s];
      const DagSeg& sg = segs[segi];
      if (sg.kind == 1) {                                 // LEAF marker: the whole CTA factors the tile
        asm volatile("bar.sync 2, %0;\n" ::"n"(Cfg::NTHREADS) : "memory");
        dag::leaf_task<Cfg::NTHREADS>(sg, t - sg.tile_begin + sg.local0, smem);
        __syncwarp();
        if (lane == 0) ws::mbar_arrive(empty + s);
        it++;
        continue;
      }
      ws::TileEnum te;
      te.init(sg.p.M, sg.p.N, BM, BN, sg.p.lower, sg.rev_m, sg.rev_n);
      ws::decode_tile<BM, BN, BK>(sg.p, te, t - sg.tile_begin + sg.local0, w);
      AccTile<Cfg> acc;
#pragma unroll
      for (int i = 0; i < FM; i++)
#pragma unroll
        for (int j = 0; j < FN; j++) acc.v[i][j][0] = acc.v[i][j][1] = 0.0;
      const bool skip = w.p.lower && w.p.diag_off == 0 && w.m0 == w.n0 && (wm0 + Cfg::WM - 1 < wn0);
      if (w.nkt == 0) {
        __syncwarp();
        if (lane == 0) ws::mbar_arrive(empty + s);
        it++;
      }
      int wk_lo = 0, wk_hi = 0x7fffffff;                  // this warp's own k-range (see dmma_ws_kernel)
      if (w.p.klo_m) wk_lo = max(wk_lo, w.m0 + wm0);
      if (w.p.klo_n) wk_lo = max(wk_lo, w.n0 + wn0);
      if (w.p.khi_m) wk_hi = min(wk_hi, w.m0 + wm0 + Cfg::WM);
      if (w.p.khi_n) wk_hi = min(wk_hi, w.n0 + wn0 + Cfg::WN);
      for (int kt = 0; kt < w.nkt; kt++, it++) {
        if (kt > 0) {
          s = it % STAGES;
          ws::mbar_wait(full + s, (it / STAGES) & 1);
        }
        const double* a_s = sA + s * SL::A_STAGE;
        const double* b_s = sB + s * SL::B_STAGE;
        const int kb = w.k_lo + kt * BK;
        if (!skip && kb + BK > wk_lo && kb < wk_hi) {
#pragma unroll
          for (int ks = 0; ks < BK / 4; ks++) {
            double af[FM], bf[FN];
#pragma unroll
            for (int i = 0; i < FM; i++) af[i] = a_s[(4 * ks + t4) * SL::LDA + wm0 + 8 * i + g];
#pragma unroll
            for (int j = 0; j < FN; j++) bf[j] = b_s[(4 * ks + t4) * SL::LDB + wn0 + 8 * j + g];
#pragma unroll
            for (int i = 0; i < FM; i++)
#pragma unroll
              for (int j = 0; j < FN; j++) dmma884(acc.v[i][j][0], acc.v[i][j][1], bf[j], af[i]);
          }
        }
        __syncwarp();
        if (lane == 0) ws::mbar_arrive(empty + s);
      }
      acc.skip = skip;
      acc.m0 = w.m0;
      acc.n0 = w.n0;
      acc.m_base = w.m0 + wm0 + 2 * t4;
      acc.n_base = w.n0 + wn0 + g;
      acc.batch = w.batch;
      acc.split = w.split;
      acc.node = w.node;
      sg.epi(w.p, acc, scratch);
      Cfg::sync();                                        // every consumer warp's stores are issued ...
      if (tid == 0) {
        __threadfence();                                  // ... and visible before the tile counts as done
        atomicAdd(done + sg.cnt + w.batch, 1);
        if (trace) {
          trace[4 * (size_t)t + 2] = dag::gtime();
          trace[4 * (size_t)t + 3] = dag::smid() | ((long long)(sg.p.K >> 4) << 9) | ((long long)w.batch << 24) | ((long long)segi << 32);
        }
      }
    }
  }
}

template <class Cfg>
cudaError_t launch_dag(const DagSeg* d_segs, int nseg, int total, int num_sms, int* sched, int* done, long long* trace, cudaStream_t st) {
  using SL = WsSmem<Cfg, false, false>;
  auto kern = dmma_dag_kernel<Cfg>;
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SL::BYTES);
    if (e != cudaSuccess) return e;
    attr_set[dev & 63] = true;
  }
  if (total <= 0) return cudaSuccess;
  const long long slots = (long long)num_sms * Cfg::CTAS_PER_SM;
  const int grid = (int)(total < slots ? total : slots);
  kern<<<grid, Cfg::NTHREADS, SL::BYTES, st>>>(d_segs, nseg, total, sched, done, trace);
  return cudaGetLastError();
}

}  // namespace gps
