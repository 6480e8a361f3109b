// dmma_ws.cuh — warp-specialised, persistent FP64 tensor-core GEMM for sm_100a (large tiles).
//
// Same contract as dmma_gemm_kernel (dmma_gemm.cuh: GemmParams, operand layout flags, epilogues), different
// machinery.  The 64 x 64 cp.async kernel tops out at 86 % of the DMMA pipe (31.9 of 37.1 TF/s measured
// register-only on B200, lab/dmma_lab.cu): every k-step ends in a CTA-wide barrier and the small tile asks
// L2 for 16 KB per 131 Kflop.  Here
//   * one PRODUCER warp moves operand rows with bulk async copies (cp.async.bulk global -> shared, the TMA
//     engine's 1-D form; rows land in the padded, bank-conflict-free layout the fragment loads need, which
//     a dense 2-D tensor-map box could not give) and signals an mbarrier per pipeline stage (complete_tx);
//   * eight CONSUMER warps (64 x 32 accumulators each, 128 x 128 per CTA) wait on the stage's "full"
//     barrier, issue DMMA.8x8x4 from shared-memory fragments, and release the stage on its "empty" barrier.
//     No __syncthreads in the main loop: consumers never wait for each other, only for data;
//   * the kernel is persistent (one CTA per SM).  Tiles are handed out DYNAMICALLY, longest k-range first:
//     the producer takes the next tile index from a global counter and tells the consumers through the
//     pipeline itself (each stage carries its tile index), so it runs ahead into the next tile while the
//     consumers are still in the epilogue of the current one.  Triangular operands give tiles whose k-ranges
//     differ by 10x; dealing them round-robin left 12-16 % of the machine idle (measured), the counter does not.
//     The counter resets itself (the last CTA to finish zeroes it), so a captured CUDA graph can replay.
// Determinism: a tile's k-loop order is fixed and tiles never share an output, so results do not depend on
// which CTA runs a tile; they are bit-identical to dmma_gemm_kernel's (same BK, same k-order).
//
// Requirements beyond dmma_gemm.cuh's: K even for k-contiguous operands (A_KC / B_KC) so that every bulk
// copy is a multiple of 16 bytes; rows may be over-read by one element up to the (even) leading dimension.
#pragma once
#include <stdlib.h>

#include "dmma_gemm.cuh"

namespace gps {

template <int BM_, int BN_, int BK_, int WARPS_M_, int WARPS_N_, int STAGES_, int CTAS_PER_SM_ = 1>
struct WsCfg {
  static constexpr int BM = BM_, BN = BN_, BK = BK_, WARPS_M = WARPS_M_, WARPS_N = WARPS_N_, STAGES = STAGES_;
  static constexpr int CTAS_PER_SM = CTAS_PER_SM_;
  static constexpr int NT = WARPS_M * WARPS_N * 32;   // consumer threads (epilogues see only these)
  static constexpr int NTHREADS = NT + 32;            // + the producer warp (last warp of the CTA)
  static constexpr int WM = BM / WARPS_M, WN = BN / WARPS_N;
  static constexpr int FM = WM / 8, FN = WN / 8;
  static_assert(BM % (8 * WARPS_M) == 0 && BN % (8 * WARPS_N) == 0, "warp tiling");
  static_assert(BK % 16 == 0, "k tiling");
  // barrier among the consumer warps only (named barrier 1); the producer never joins it
  static __device__ __forceinline__ void sync() { asm volatile("bar.sync 1, %0;\n" ::"n"(NT) : "memory"); }
};

namespace ws {

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ unsigned mbar_try_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// 1-D bulk async copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void zero_row(double* row, int from, int to) {   // [from, to) doubles, both even
  for (int c = from; c < to; c += 2) *reinterpret_cast<double2*>(row + c) = make_double2(0.0, 0.0);
}

// Valid tiles of one z-slice (batch entry x node x k-split), longest k-range first where it is known.
// lower (BM == BN, diag_off == 0): tile (i, j) is kept iff j <= i; rows i < nt hold i + 1 tiles, the rest nt.
struct TileEnum {
  int mt, nt, lower, per_z, rev_m, rev_n;
  __host__ __device__ void init(int M, int N, int BM, int BN, int lower_, int rev_m_, int rev_n_ = 0) {
    mt = (M + BM - 1) / BM;
    nt = (N + BN - 1) / BN;
    lower = lower_;
    rev_m = rev_m_;
    rev_n = rev_n_;
    if (lower) {
      int tri = mt < nt ? mt : nt;
      per_z = tri * (tri + 1) / 2 + (mt > nt ? (mt - nt) * nt : 0);
    } else {
      per_z = mt * nt;
    }
  }
  __device__ void decode(int r, int& i, int& j) const {
    if (!lower) {
      i = r % mt;
      j = r / mt;
    } else {
      const int tri = mt < nt ? mt : nt;
      const int ntri = tri * (tri + 1) / 2;
      if (r < ntri) {
        i = (int)((sqrt(8.0 * r + 1.0) - 1.0) * 0.5);
        while ((i + 1) * (i + 2) / 2 <= r) i++;
        while (i * (i + 1) / 2 > r) i--;
        j = r - i * (i + 1) / 2;
      } else {
        r -= ntri;
        i = nt + r / nt;
        j = r % nt;
      }
    }
    if (rev_m) i = mt - 1 - i;
    if (rev_n) j = nt - 1 - j;
  }
};

// One tile's work, derived identically by the producer and the consumers from the tile index.
struct TileWork {
  GemmParams p;   // extents after per-node clipping
  int m0, n0, batch, node, split, k_lo, nkt;
  bool valid;
};
template <int BM, int BN, int BK>
__device__ __forceinline__ void decode_tile(const GemmParams& p_in, const TileEnum& te, long long tile, TileWork& w) {
  w.p = p_in;
  GemmParams& p = w.p;
  const int zz_all = (int)(tile / te.per_z);
  int ti, tj;
  te.decode((int)(tile % te.per_z), ti, tj);
  w.m0 = ti * BM;
  w.n0 = tj * BN;
  const int zz = zz_all / p.ksplit;
  w.split = zz_all % p.ksplit;
  w.batch = zz / p.nodes;
  w.node = zz % p.nodes;
  w.valid = apply_clip(p, w.node) && w.m0 < p.M && w.n0 < p.N && !(p.lower && (w.m0 + BM - 1 + p.diag_off < w.n0));
  int k_lo = 0, k_hi = p.K;
  if (p.klo_n) k_lo = max(k_lo, w.n0);
  if (p.klo_m) k_lo = max(k_lo, w.m0);
  if (p.khi_m) k_hi = min(k_hi, w.m0 + BM);
  if (p.khi_n) k_hi = min(k_hi, w.n0 + BN);
  if (p.ksplit > 1) {
    int kc = ((p.K + BK - 1) / BK + p.ksplit - 1) / p.ksplit * BK;
    k_lo = max(k_lo, w.split * kc);
    k_hi = min(k_hi, w.split * kc + kc);
  }
  k_lo = (k_lo / BK) * BK;
  w.k_lo = k_lo;
  w.nkt = (k_hi > k_lo) ? (k_hi - k_lo + BK - 1) / BK : 0;
}

}  // namespace ws

template <class Cfg, bool A_KC, bool B_KC>
struct WsSmem {
  static constexpr int LDA = (A_KC ? Cfg::BK : Cfg::BM) + 4;
  static constexpr int LDB = (B_KC ? Cfg::BK : Cfg::BN) + 4;
  static constexpr int A_ROWS = A_KC ? Cfg::BM : Cfg::BK;
  static constexpr int B_ROWS = B_KC ? Cfg::BN : Cfg::BK;
  static constexpr int A_STAGE = A_ROWS * LDA;
  static constexpr int B_STAGE = B_ROWS * LDB;
  // doubles, for epilogues (EpiColDot: two [WARPS_M][BN] arrays; EpiWK: [WARPS_N][BM] + [WARPS_M][BN])
  static constexpr int SCRATCH_A = 2 * Cfg::WARPS_M * Cfg::BN, SCRATCH_B = Cfg::WARPS_N * Cfg::BM + Cfg::WARPS_M * Cfg::BN;
  static constexpr int SCRATCH = SCRATCH_A > SCRATCH_B ? SCRATCH_A : SCRATCH_B;
  static constexpr int WSC = Cfg::STAGES * Cfg::BK;        // doubles, per-stage k-weights (KS kernels)
  static constexpr size_t BYTES = ((size_t)Cfg::STAGES * (A_STAGE + B_STAGE) + SCRATCH + WSC) * sizeof(double) +
                                  2 * Cfg::STAGES * 8 + Cfg::STAGES * sizeof(int);
};

// KS: weighted product A diag(w) B — the producer also stages w[kbase .. kbase+BK) and the consumers scale their
// A fragments by it (CovFunc's x / ell without ever materialising the scaled copy of X).
template <class Cfg, bool A_KC, bool B_KC, class Epi, bool KS = false>
__global__ void __launch_bounds__(Cfg::NTHREADS, Cfg::CTAS_PER_SM)
    dmma_ws_kernel(GemmParams p_in, Epi epi, int zcount, int rev_m, int rev_n, int* __restrict__ sched, int quota) {
  using SL = WsSmem<Cfg, A_KC, B_KC>;
  constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, STAGES = Cfg::STAGES;
  constexpr int FM = Cfg::FM, FN = Cfg::FN;
  extern __shared__ __align__(16) double smem[];
  double* sA = smem;
  double* sB = smem + STAGES * SL::A_STAGE;
  double* scratch = smem + STAGES * (SL::A_STAGE + SL::B_STAGE);
  double* s_w = scratch + SL::SCRATCH;
  unsigned long long* full = reinterpret_cast<unsigned long long*>(s_w + SL::WSC);
  unsigned long long* empty = full + STAGES;
  volatile int* stage_tile = reinterpret_cast<volatile int*>(empty + STAGES);   // tile index carried by each stage

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const bool producer = (warp == Cfg::NT / 32);

  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) {
      ws::mbar_init(full + s, 32);                 // every producer lane arrives (with its byte count)
      ws::mbar_init(empty + s, Cfg::NT / 32);      // one arrival per consumer warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  ws::TileEnum te;
  te.init(p_in.M, p_in.N, BM, BN, p_in.lower, rev_m, rev_n);
  const long long total = (long long)te.per_z * zcount;

  unsigned it = 0;   // pipeline slot counter, runs across tiles (same sequence in the producer and the consumers)
  ws::TileWork w;

  if (producer) {
    // quota: tiles this CTA may take before it retires.  One wave (quota = all) is the plain persistent kernel; with the
    // grid oversubscribed (ws_waves() > 1) CTAs retire early and hand their SM slot back to the block scheduler, which
    // lets another stream's higher-priority latency-bound kernels in while this launch is still running.
    for (int taken = 0; taken < quota; taken++) {
      int t = 0;
      if (lane == 0) t = atomicAdd(&sched[0], 1);
      t = __shfl_sync(0xffffffffu, t, 0);
      if (t >= total) break;
      ws::decode_tile<BM, BN, BK>(p_in, te, t, w);
      if (!w.valid) {
        taken--;
        continue;
      }
      const GemmParams& p = w.p;
      const int m0 = w.m0, n0 = w.n0;
      const double* __restrict__ A = p.A + (size_t)w.batch * p.strideA + (size_t)w.node * p.nodeStrideA;
      const double* __restrict__ B = p.B + (size_t)w.batch * p.strideB + (size_t)w.node * p.nodeStrideB;
      const int vm = min(p.M - m0, BM), vn = min(p.N - n0, BN);   // valid extent of the tile (> 0)
      const int nst = w.nkt > 0 ? w.nkt : 1;                      // an empty k-range still sends one (data-less) stage
      for (int kt = 0; kt < nst; kt++, it++) {
        const int s = it % STAGES;
        ws::mbar_wait(empty + s, ((it / STAGES) & 1) ^ 1);
        if (lane == 0) stage_tile[s] = t;
        if (w.nkt == 0) {
          ws::mbar_arrive_tx(full + s, 0);
          continue;
        }
        const int kbase = w.k_lo + kt * BK;
        const int vk = min(max(p.K - kbase, 0), BK);    // valid k of this stage (even for KC operands)
        double* a_s = sA + s * SL::A_STAGE;
        double* b_s = sB + s * SL::B_STAGE;
        unsigned bytes = 0;
        // pass 1: byte count of this lane's rows, zero fill of what no copy will write
        if (!A_KC) {
          for (int r = lane; r < BK; r += 32) {
            if (r < vk) bytes += ((vm + 1) & ~1) * 8;
            else ws::zero_row(a_s + r * SL::LDA, 0, BM);
          }
        } else {
          for (int r = lane; r < BM; r += 32) {
            if (r < vm) {
              bytes += vk * 8;
              if (vk < BK) ws::zero_row(a_s + r * SL::LDA, vk, BK);
            } else if (vk < BK) {
              ws::zero_row(a_s + r * SL::LDA, 0, BK);
            }
          }
        }
        if (!B_KC) {
          for (int r = lane; r < BK; r += 32) {
            if (r < vk) bytes += ((vn + 1) & ~1) * 8;
            else ws::zero_row(b_s + r * SL::LDB, 0, BN);
          }
        } else {
          for (int r = lane; r < BN; r += 32) {
            if (r < vn) {
              bytes += vk * 8;
              if (vk < BK) ws::zero_row(b_s + r * SL::LDB, vk, BK);
            } else if (vk < BK) {
              ws::zero_row(b_s + r * SL::LDB, 0, BK);
            }
          }
        }
        if (KS && lane == 0) bytes += BK * 8;
        if (vk < BK) asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        ws::mbar_arrive_tx(full + s, bytes);
        if (KS && lane == 0)
          ws::bulk_g2s(s_w + s * BK, p.kscale + (size_t)w.batch * p.strideScale + (size_t)kbase, BK * 8, full + s);
        // pass 2: the copies
        if (!A_KC) {
          const unsigned nb = ((vm + 1) & ~1) * 8;
          for (int r = lane; r < vk; r += 32)
            ws::bulk_g2s(a_s + r * SL::LDA, A + (size_t)m0 + (size_t)(kbase + r) * p.lda, nb, full + s);
        } else if (vk > 0) {
          for (int r = lane; r < vm; r += 32)
            ws::bulk_g2s(a_s + r * SL::LDA, A + (size_t)kbase + (size_t)(m0 + r) * p.lda, vk * 8, full + s);
        }
        if (!B_KC) {
          const unsigned nb = ((vn + 1) & ~1) * 8;
          for (int r = lane; r < vk; r += 32)
            ws::bulk_g2s(b_s + r * SL::LDB, B + (size_t)n0 + (size_t)(kbase + r) * p.ldb, nb, full + s);
        } else if (vk > 0) {
          for (int r = lane; r < vn; r += 32)
            ws::bulk_g2s(b_s + r * SL::LDB, B + (size_t)kbase + (size_t)(n0 + r) * p.ldb, vk * 8, full + s);
        }
      }
    }
    {   // end-of-work marker for the consumers
      const int s = it % STAGES;
      ws::mbar_wait(empty + s, ((it / STAGES) & 1) ^ 1);
      if (lane == 0) stage_tile[s] = -1;
      ws::mbar_arrive_tx(full + s, 0);
    }
    if (lane == 0) {   // the last CTA to run out of tiles re-arms the counter for the next launch / graph replay
      __threadfence();
      const int done = atomicAdd(&sched[1], 1);
      if (done == (int)gridDim.x - 1) {
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
      ws::decode_tile<BM, BN, BK>(p_in, te, t, w);
      AccTile<Cfg> acc;
#pragma unroll
      for (int i = 0; i < FM; i++)
#pragma unroll
        for (int j = 0; j < FN; j++) acc.v[i][j][0] = acc.v[i][j][1] = 0.0;
      // diagonal tile of a lower-triangular product: a warp whose sub-tile lies entirely above the diagonal only keeps the
      // pipeline's barriers in step (it waits for each stage and releases it without touching the data)
      const bool skip = epi_skips_upper<Epi>::value && w.p.lower && w.p.diag_off == 0 && w.m0 == w.n0 && (wm0 + Cfg::WM - 1 < wn0);

      if (w.nkt == 0) {
        __syncwarp();
        if (lane == 0) ws::mbar_arrive(empty + s);
        it++;
      }
      // triangular operands: the tile's k-range [k_lo, k_hi) is the union over its rows / columns; THIS warp's rows and
      // columns need less, and the k-steps outside its own range only multiply the operand's stored zeros — it keeps the
      // barriers in step for them and skips the math (same sums bit for bit: the skipped terms are exact zeros)
      int wk_lo = 0, wk_hi = 0x7fffffff;
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
          for (int i = 0; i < FM; i++) {
            int m = wm0 + 8 * i + g, k = 4 * ks + t4;
            af[i] = A_KC ? a_s[m * SL::LDA + k] : a_s[k * SL::LDA + m];
          }
          if (KS) {
            const double wk = s_w[s * BK + 4 * ks + t4];
#pragma unroll
            for (int i = 0; i < FM; i++) af[i] *= wk;
          }
#pragma unroll
          for (int j = 0; j < FN; j++) {
            int n = wn0 + 8 * j + g, k = 4 * ks + t4;
            bf[j] = B_KC ? b_s[n * SL::LDB + k] : b_s[k * SL::LDB + n];
          }
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
      if (!epi_no_scratch<Epi>::value) Cfg::sync();   // the epilogue scratch of the previous tile is no longer being read
      epi(w.p, acc, scratch);
    }
  }
}

// Measured on B200 (lab/dmma_lab.cu, "NT" operands, TF/s; cublasDgemm 35.9, register-only DMMA issue 37.1):
//                               8192^3   2048^3 x 8   Kinv n=2000 x 16   update 978 x 976 x 1024 x 32   Gram 400 x 400 x 2000 x 64
//   64 x 64 cp.async (4 CTAs/SM)  31.2      30.7           26.7                  25.9                         19.8
//   WsCfg128 (1 CTA/SM)           35.3      34.8           26.0                  27.3                         14.9
//   WsCfg64  (3 CTAs/SM)          34.6      34.2           29.7                  28.7                         21.9
//   WsCfg80  (2 CTAs/SM)          34.8      33.5           29.1                  27.7                         26.1
// 128 x 128: least L2 traffic, best for large problems; 64 x 64: half the over-compute on triangular / ragged
// problems and small CTAs that share an SM with other kernels; 80 x 80: for row counts that 80 divides better.
using WsCfg128 = WsCfg<128, 128, 16, 2, 4, 5>;        // 8 consumer warps of 64 x 32, 5 stages, 170 KB smem
using WsCfg64 = WsCfg<64, 64, 16, 2, 2, 4, 3>;        // 4 consumer warps of 32 x 32, 4 stages, 69 KB smem (3 stages: 20.89 -> 20.84 ms at C2 x 64; BK = 32 x 2 stages: slower)
using WsCfg48 = WsCfg<48, 48, 16, 2, 2, 4, 4>;        // 4 consumer warps of 24 x 24, 4 CTAs/SM: lower-triangular products of a few hundred rows (least diagonal-tile over-compute)
using WsCfg80 = WsCfg<80, 80, 16, 2, 2, 4, 2>;        // 4 consumer warps of 40 x 40, 4 stages, 86 KB smem
using WsCfg96 = WsCfg<96, 96, 16, 2, 4, 6, 1>;         // 8 consumer warps of 48 x 24: row counts just under a multiple of 96 (n = 285)
using WsCfg96x64 = WsCfg<96, 64, 16, 2, 2, 3, 3>;     // 4 consumer warps of 48 x 32, 3 CTAs/SM: the same rows against a wide N (M * Xaug, d = 10^4)
using WsCfgN32 = WsCfg<128, 32, 16, 4, 1, 4, 3>;      // narrow outputs (N <= 32): 4 warps of 32 x 32
using WsCfgN24 = WsCfg<128, 24, 16, 4, 1, 4, 3>;      // N <= 24 (M * Xaug with d = 20: 21 columns): 4 warps of 32 x 24

inline int& ws_waves() {   // GPS_WAVES: grid oversubscription of the persistent kernel (1 = plain persistent)
  static int waves = [] {
    const char* ev = getenv("GPS_WAVES");
    const int v = ev ? atoi(ev) : 1;
    return v < 1 ? 1 : (v > 16 ? 16 : v);
  }();
  return waves;
}

// `sched`: two ints in device memory, zero before the first launch that uses them (the kernel re-arms them);
// launches that may overlap in time must not share them.
template <class Cfg, bool A_KC, bool B_KC, class Epi, bool KS = false>
cudaError_t launch_gemm_ws(const GemmParams& p, const Epi& epi, int batch, int num_sms, int* sched, cudaStream_t st) {
  using SL = WsSmem<Cfg, A_KC, B_KC>;
  auto kern = dmma_ws_kernel<Cfg, A_KC, B_KC, Epi, KS>;
  static bool attr_set[64] = {false};   // per instantiation, per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SL::BYTES);
    if (e != cudaSuccess) return e;
    attr_set[dev & 63] = true;
  }
  if (p.M <= 0 || p.N <= 0) return cudaSuccess;
  ws::TileEnum te;
  // longest tiles first: with A upper-triangular in (m, k) (klo_m) the k-range shrinks as m grows -> natural
  // order; with A lower-triangular (khi_m) or B upper-triangular (khi_n) it grows -> reverse that index
  const int rev_m = (p.khi_m && !p.lower) ? 1 : 0;
  const int rev_n = (p.khi_n && !p.lower) ? 1 : 0;
  te.init(p.M, p.N, Cfg::BM, Cfg::BN, p.lower, rev_m, rev_n);
  const int zcount = batch * p.nodes * p.ksplit;
  const long long total = (long long)te.per_z * zcount;
  const long long slots = (long long)num_sms * Cfg::CTAS_PER_SM * ws_waves();
  const int grid = (int)(total < slots ? total : slots);
  const int quota = ws_waves() > 1 ? (int)((total + grid - 1) / grid) : 0x7fffffff;
  return launch_kernel(kern, dim3(grid), dim3(Cfg::NTHREADS), SL::BYTES, st, p, epi, zcount, rev_m, rev_n, sched, quota);
}

}  // namespace gps
