// lstm_tc.cu — persistent BLSTM recurrence on tcgen05 tensor cores for H = 392 (BSRNN_baseline, N = 196).
//
// Replaces the recurrent half of nn.LSTM(N, 2N, bidirectional) [reference bsrnn_flowse.py:226-238, called at
// :296-297 (time axis: 2176 sequences x 1001 steps at BASELINE config 2) and :303-304 (band axis: 64064 x 34)].
//
// Decomposition.  A work unit is (direction d, tile j of 128 sequences).  A thread-block CLUSTER of 8 CTAs owns a
// GROUP of up to NS units of one direction and advances them in lock step, slot after slot ("interleaved"): the
// dependency chain of one unit (h published -> peers wake -> 100 KB h tile fetched from L2 -> 25 MMAs -> gates ->
// h stored -> published) is ~10 us long but occupies each resource (copy ring, tensor pipe, epilogue warps) for a
// fraction of that, so NS independent chains overlap on the same CTAs.  CTA q of the cluster owns hidden units
// [49q, 49q+49) i.e. 196 of the 1568 gate columns, and keeps that slice of W_hh (208 x 400 fp16, 166 KB, UMMA KB8
// layout) resident in shared memory for the whole launch.  Per (step, slot) every CTA
//   (1) producer warp: acquires the slot's cluster-scope h_ready mbarrier, bulk-copies the full h_{t-1} tile
//       (128 x 400 fp16 = 100 KB) from the y buffer in L2 through a 3-stage ring,
//   (2) MMA warp: 25 tcgen05.mma (M=128, N=208, K=16) into one of two TMEM accumulators,
//   (3) 8 epilogue warps: tcgen05.ld, add the precomputed input projection (fp16, prefetched into L2 one round
//       ahead), gates with MUFU tanh, c kept in REGISTERS (25 f32 per thread per slot), h_t slice stored to y (which
//       is at once the layer output consumed by the Linear GEMM and the exchange buffer for the other 7 CTAs),
//       then a non-blocking named-barrier arrive,
//   (4) publisher warp: completes that named barrier, one fence, 8 relaxed remote arrives on the slot's h_ready
//       barrier of every CTA of the cluster.  The epilogue warps never wait for the publication.
// No grid-wide synchronisation exists: clusters are independent, rows never mix.
//
// y layout (fp16): [step][seq_tile][dir][50 k-cores][128 rows][8]   (k-core 49 = zero padding, K = 400 per dir);
//                  a (step, seq_tile) block is therefore a 128 x 800 KB8 operand tile for the Linear(4N->N) GEMM.
// gates_x (fp16) : [token][dir][q][208]  column c = 4*u_local + gate (i,f,g,o), 196 real + 12 pad
// w_hh pack      : [dir][q][50 k-cores][208][8]
#include "common.cuh"
#include "umma.cuh"
#include <cuda_fp16.h>

namespace bsrnn {
using namespace umma;

constexpr int LH = 392;            // hidden size
constexpr int LCL = 8;             // cluster size
constexpr int LBN = 208;           // gate columns per CTA (4*49 = 196, padded to a multiple of 16)
constexpr int LKC = 50;            // k-cores of the recurrent operand (K = 400)
constexpr int LKS = 10;            // k-cores per A stage
constexpr int LNST = LKC / LKS;    // 5 stages per (step, slot)
constexpr int LSTAGES = 3;
constexpr int LMAXS = 4;           // most slots any instantiation uses
constexpr int LACC = 256;          // TMEM columns per accumulator buffer
constexpr uint32_t L_W_BYTES = LKC * LBN * 16;          // 166400
constexpr uint32_t L_A_STAGE = LKS * 128 * 16;          // 20480
constexpr int L_NBARS = 2 * LSTAGES + 2 + 2 + LMAXS + 2;
constexpr size_t L_SMEM = L_W_BYTES + LSTAGES * L_A_STAGE + L_NBARS * 8 + 16;
static_assert(L_SMEM <= 232448, "exceeds the 227 KB per-CTA shared memory limit");

struct LstmTcArgs {
  const __half* gates_x;
  const __half* w_pack;
  __half* y;
  int R, steps, seq_tiles;
  int gpd;               // groups per direction (each group = up to NS consecutive sequence tiles)
  long seq_inner, seq_outer, seq_inner_stride, step_stride;
  long long* trace;      // optional (debug): [step][8] SM-clock stamps written by cluster 0 / CTA 0 / slot 0
};

struct Group {
  int d, j0, nact;
};
__device__ __forceinline__ Group group_of(const LstmTcArgs& a, int g) {
  Group r;
  r.d = g / a.gpd;
  const int gi = g - r.d * a.gpd;
  r.j0 = (int)(((long)gi * a.seq_tiles) / a.gpd);
  r.nact = (int)(((long)(gi + 1) * a.seq_tiles) / a.gpd) - r.j0;
  return r;
}

#define LSTM_TRACE(slot, step)                                                                   \
  do {                                                                                           \
    if (a.trace && cid == 0 && q == 0 && (step) < 64) a.trace[(step) * 8 + (slot)] = clock64();  \
  } while (0)

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigm_fast(float x) { return fmaf(tanh_fast(0.5f * x), 0.5f, 0.5f); }

__device__ __forceinline__ void gate_update(float pi, float pf, float pg, float po, float& c, float& h) {
  const float ig = sigm_fast(pi), fg = sigm_fast(pf), gg = tanh_fast(pg), og = sigm_fast(po);
  c = fmaf(fg, c, ig * gg);
  h = og * tanh_fast(c);
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Epilogue work split: 8 warps; warp (quadrant, half) owns 32 rows x units [U0, U0+NU) with
//   half 0: units [0,24)  = accumulator columns [0,96)    (3 chunks of 32 columns)
//   half 1: units [24,49) = accumulator columns [96,196)  (3 chunks of 32 + one of 4)
// Unit i of CTA Q is h column k = 49Q + i -> k-core 6Q + (Q+i)/8, slot (Q+i)%8 of the y tile.
// store_ready writes every 16-byte core (or the part of it this thread owns) that became complete when the local
// units [D0, D1) were produced, so h values live in registers only until their core is full.
template <int Q, int U0, int NU, int D0, int D1>
__device__ __forceinline__ void store_ready(__half* ytile_row /* &y[...][kc=0][r][0] */, const float (&h)[25]) {
#pragma unroll
  for (int jj = 0; jj < 7; ++jj) {
    const int lo = 8 * jj - Q;                       // unit index sitting in slot 0 of this core
    const int a0 = lo > U0 ? lo : U0;
    const int b0 = (lo + 8) < (U0 + NU) ? (lo + 8) : (U0 + NU);
    if (a0 >= b0) continue;                          // core holds none of our units
    const int bl = b0 - U0;                          // local index one past the last of our units in the core
    if (!(bl > D0 && bl <= D1)) continue;            // completed earlier / not complete yet
    __half* dst = ytile_row + (size_t)(6 * Q + jj) * 128 * 8;
    if (a0 == lo && b0 == lo + 8) {
      const int b = lo - U0;
      __half2 p0 = __floats2half2_rn(h[b], h[b + 1]), p1 = __floats2half2_rn(h[b + 2], h[b + 3]);
      __half2 p2 = __floats2half2_rn(h[b + 4], h[b + 5]), p3 = __floats2half2_rn(h[b + 6], h[b + 7]);
      uint4 pk = make_uint4(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1),
                            *reinterpret_cast<uint32_t*>(&p2), *reinterpret_cast<uint32_t*>(&p3));
      *reinterpret_cast<uint4*>(dst) = pk;
    } else {
#pragma unroll
      for (int sl = 0; sl < 8; ++sl) {
        const int i = lo + sl;
        if (i >= a0 && i < b0) dst[sl] = __float2half_rn(h[i - U0]);
      }
    }
  }
}

__device__ __forceinline__ void load_g4(const uint4* p, bool ok, uint4 (&g)[4]) {
  if (ok) {
#pragma unroll
    for (int i = 0; i < 4; ++i) g[i] = __ldg(p + i);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) g[i] = make_uint4(0, 0, 0, 0);
  }
}

// One (step, slot) item of the epilogue for a thread: row r of the tile, units [U0, U0+NU) of CTA Q.
template <int Q, int HALF>
__device__ __forceinline__ void epilogue_item(uint32_t t_addr, bool have_acc, uint64_t* acc_full, uint32_t acc_phase,
                                              const __half* gx, bool row_ok, __half* ytile_row, float (&c)[25]) {
  constexpr int U0 = HALF == 0 ? 0 : 24;
  constexpr int NU = HALF == 0 ? 24 : 25;
  const uint4* gp = reinterpret_cast<const uint4*>(gx + HALF * 96);
  uint4 g[2][4];
  uint2 gt = make_uint2(0, 0);
  load_g4(gp, row_ok, g[0]);                         // issued before the wait
  if (have_acc) {
    mbar_wait(acc_full, acc_phase);
    tc_fence_after();
  }
  float h[25];
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    if (ch < 2) load_g4(gp + 4 * (ch + 1), row_ok, g[(ch + 1) & 1]);
    else if (HALF == 1 && row_ok) gt = __ldg(reinterpret_cast<const uint2*>(gx + 192));
    uint32_t acc[32];
    if (have_acc) {
      tmem_ld_x32(t_addr + HALF * 96 + ch * 32, acc);
      tmem_ld_wait();
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] = 0u;
    }
    const __half2* gh = reinterpret_cast<const __half2*>(&g[ch & 1][0]);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float2 g01 = __half22float2(gh[2 * u]), g23 = __half22float2(gh[2 * u + 1]);
      gate_update(__uint_as_float(acc[4 * u]) + g01.x, __uint_as_float(acc[4 * u + 1]) + g01.y,
                  __uint_as_float(acc[4 * u + 2]) + g23.x, __uint_as_float(acc[4 * u + 3]) + g23.y, c[ch * 8 + u],
                  h[ch * 8 + u]);
    }
    if (ch == 0) store_ready<Q, U0, NU, 0, 8>(ytile_row, h);
    if (ch == 1) store_ready<Q, U0, NU, 8, 16>(ytile_row, h);
    if (ch == 2) store_ready<Q, U0, NU, 16, 24>(ytile_row, h);
  }
  if (HALF == 1) {
    uint32_t acc[4];
    if (have_acc) {
      tmem_ld_x4(t_addr + 192, acc);
      tmem_ld_wait();
    } else {
      acc[0] = acc[1] = acc[2] = acc[3] = 0u;
    }
    const __half2* gh = reinterpret_cast<const __half2*>(&gt);
    const float2 g01 = __half22float2(gh[0]), g23 = __half22float2(gh[1]);
    gate_update(__uint_as_float(acc[0]) + g01.x, __uint_as_float(acc[1]) + g01.y, __uint_as_float(acc[2]) + g23.x,
                __uint_as_float(acc[3]) + g23.y, c[24], h[24]);
    store_ready<Q, U0, NU, 24, 25>(ytile_row, h);
  }
}

// The epilogue role for one (Q, HALF): loops over this cluster's groups, steps and slots.
template <int Q, int HALF, int NS, int NTHR_PUB>
__device__ __forceinline__ void epilogue_role(const LstmTcArgs& a, uint32_t tmem_base, int warp, int lane, int cid, int ncl,
                                              bool tracer, uint64_t* acc_full, uint64_t* acc_empty, uint64_t* w_free) {
  const int quad = warp & 3;
  const int q = Q;
  const int r = quad * 32 + lane;
  const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
  const int ngroups = 2 * a.gpd;
  const size_t tile_elems = (size_t)LKC * 128 * 8;
  const long gx_step = a.step_stride * 2 * (LCL * LBN);      // halves between consecutive positions of a sequence
  uint32_t it = 0;                                            // accumulator items consumed so far (all groups)
  float c[NS][25];
  for (int g = cid; g < ngroups; g += ncl) {
    const Group G = group_of(a, g);
    long tok0[NS];
    bool row_ok[NS];
#pragma unroll
    for (int k = 0; k < NS; ++k) {
      const long seq = (long)(G.j0 + k) * 128 + r;
      row_ok[k] = k < G.nact && seq < a.R;
      tok0[k] = row_ok[k] ? (seq / a.seq_inner) * a.seq_outer + (seq % a.seq_inner) * a.seq_inner_stride : 0;
#pragma unroll
      for (int i = 0; i < 25; ++i) c[k][i] = 0.f;
    }
    for (int s = 0; s < a.steps; ++s) {
      const int p = G.d == 0 ? s : a.steps - 1 - s;
      const bool have_acc = s > 0;
      const bool more = s + 1 < a.steps;
#pragma unroll
      for (int k = 0; k < NS; ++k) {
        if (k < G.nact) {
          const long token = tok0[k] + (long)p * a.step_stride;
          const __half* gx = a.gates_x + (token * 2 + G.d) * (LCL * LBN) + Q * LBN;
          __half* ytile_row = a.y + (((size_t)p * a.seq_tiles + (G.j0 + k)) * 2 + G.d) * tile_elems + (size_t)r * 8;
          if (more && row_ok[k]) {                   // next step's input projection -> L2, one full round ahead
            const char* nx = reinterpret_cast<const char*>(gx + (G.d == 0 ? gx_step : -gx_step) + HALF * 96);
            prefetch_l2(nx);
            prefetch_l2(nx + 100);
            prefetch_l2(nx + 199);
          }
          const uint32_t buf = it & 1;
          epilogue_item<Q, HALF>(t_lane + buf * LACC, have_acc, acc_full + buf, (it >> 1) & 1, gx, row_ok[k], ytile_row,
                                 c[k]);
          if (tracer && k == 0) LSTM_TRACE(6, s);
          if (have_acc) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty + buf);
            ++it;
          }
          // h_t slice of this warp is stored: tell the publisher (non-blocking)
          if (more) asm volatile("bar.arrive %0, %1;" ::"r"(1 + k), "n"(256 + NTHR_PUB) : "memory");
        }
      }
    }
    const int gn = g + ncl;                          // next group of this cluster switches direction?
    if (gn < ngroups && gn / a.gpd != G.d) {
      __syncwarp();
      if (lane == 0) mbar_arrive(w_free);
    }
  }
}

// NS = interleaved slots; REGSPLIT: 12 warps with setmaxnreg (producer/MMA/publisher/idle give registers to the 8
// epilogue warps) instead of 11 warps with a uniform budget.
template <int NS, bool REGSPLIT>
__global__ void __cluster_dims__(LCL, 1, 1) __launch_bounds__(REGSPLIT ? 384 : 352, 1) lstm_tc_kernel(const LstmTcArgs a) {
  constexpr int EW0 = REGSPLIT ? 4 : 3;              // first epilogue warp
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sW = smem;
  uint8_t* sA = smem + L_W_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + LSTAGES * L_A_STAGE);
  uint64_t* full = bars;                       // [3]
  uint64_t* empty = bars + LSTAGES;            // [3]
  uint64_t* acc_full = bars + 2 * LSTAGES;     // [2]
  uint64_t* acc_empty = acc_full + 2;          // [2]
  uint64_t* h_ready = acc_empty + 2;           // [LMAXS]
  uint64_t* w_full = h_ready + LMAXS;
  uint64_t* w_free = w_full + 1;               // epilogue -> producer: the resident W slice may be overwritten
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_free + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t q = cluster_ctarank();
  const int cid = cluster_id_x(), ncl = num_clusters_x();

  if (threadIdx.x == 0) {
    for (int i = 0; i < LSTAGES; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, 8); }
    for (int i = 0; i < LMAXS; ++i) mbar_init(h_ready + i, LCL);
    mbar_init(w_full, 1);
    mbar_init(w_free, 8);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 2 * LACC);
  tc_fence_before();
  __syncthreads();
  cluster_sync();                       // every CTA's barriers are initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int ngroups = 2 * a.gpd;
  const size_t tile_elems = (size_t)LKC * 128 * 8;     // halves per (step, tile, dir)

  if (warp == 0) {
    // ------------------------------------------------------------------ producer: W slice + h tiles
    if (REGSPLIT) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, hphase = 0, wfphase = 0;
      int cur_dir = -1;
      for (int g = cid; g < ngroups; g += ncl) {
        const Group G = group_of(a, g);
        if (G.d != cur_dir) {
          // The epilogue warps signal w_free after consuming the last accumulator of the previous direction, i.e.
          // after every MMA that read the old slice has retired.
          if (cur_dir >= 0) {
            mbar_wait(w_free, wfphase);
            wfphase ^= 1;
          }
          mbar_expect_tx(w_full, L_W_BYTES);
          const uint8_t* src = reinterpret_cast<const uint8_t*>(a.w_pack) + ((size_t)G.d * LCL + q) * L_W_BYTES;
          for (uint32_t off = 0; off < L_W_BYTES; off += 33280) bulk_g2s(sW + off, src + off, 33280, w_full);
          cur_dir = G.d;
        }
        for (int s = 1; s < a.steps; ++s) {
          const int p_prev = G.d == 0 ? s - 1 : a.steps - s;        // position whose h feeds this step
          for (int k = 0; k < G.nact; ++k) {
            mbar_wait_cluster(h_ready + k, (hphase >> k) & 1);
            hphase ^= 1u << k;
            if (k == 0) LSTM_TRACE(0, s);
            fence_proxy_async_global();
            const uint8_t* src = reinterpret_cast<const uint8_t*>(
                a.y + (((size_t)p_prev * a.seq_tiles + (G.j0 + k)) * 2 + G.d) * tile_elems);
            for (int ks = 0; ks < LNST; ++ks) {
              mbar_wait(empty + stage, phase ^ 1);
              mbar_expect_tx(full + stage, L_A_STAGE);
              bulk_g2s(sA + stage * L_A_STAGE, src + (size_t)ks * L_A_STAGE, L_A_STAGE, full + stage);
              if (++stage == LSTAGES) { stage = 0; phase ^= 1; }
            }
            if (k == 0) LSTM_TRACE(1, s);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (REGSPLIT) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (lane == 0) {
      const uint32_t idesc = idesc_f16_f32(128, LBN);
      uint32_t stage = 0, phase = 0, it = 0, wphase = 0;
      int cur_dir = -1;
      const uint32_t sw = smem_u32(sW);
      for (int g = cid; g < ngroups; g += ncl) {
        const Group G = group_of(a, g);
        if (G.d != cur_dir) {
          mbar_wait(w_full, wphase);
          wphase ^= 1;
          cur_dir = G.d;
        }
        for (int s = 1; s < a.steps; ++s) {
          for (int k = 0; k < G.nact; ++k, ++it) {
            const uint32_t buf = it & 1;
            mbar_wait(acc_empty + buf, ((it >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + buf * LACC;
            for (int ks = 0; ks < LNST; ++ks) {
              mbar_wait(full + stage, phase);
              if (k == 0 && ks == 0) LSTM_TRACE(2, s);
              if (k == 0 && ks == LNST - 1) LSTM_TRACE(3, s);
              tc_fence_after();
              const uint32_t sa = smem_u32(sA + stage * L_A_STAGE);
#pragma unroll
              for (int jk = 0; jk < LKS / 2; ++jk) {
                const uint64_t da = smem_desc_kb8(sa + jk * 2 * 2048, 2048, 128);
                const uint64_t db = smem_desc_kb8(sw + (ks * LKS + jk * 2) * (LBN * 16), LBN * 16, 128);
                mma_f16_ss(d_tmem, da, db, idesc, (ks | jk) != 0);
              }
              mma_commit(empty + stage);
              if (++stage == LSTAGES) { stage = 0; phase ^= 1; }
            }
            mma_commit(acc_full + buf);
            if (k == 0) LSTM_TRACE(4, s);
          }
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ publisher
    if (REGSPLIT) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    for (int g = cid; g < ngroups; g += ncl) {
      const Group G = group_of(a, g);
      for (int s = 0; s + 1 < a.steps; ++s) {
        for (int k = 0; k < G.nact; ++k) {
          // completes once the 8 epilogue warps have stored their h_t slices of slot k
          asm volatile("bar.sync %0, %1;" ::"r"(1 + k), "n"(256 + 32) : "memory");
          if (lane < LCL) {
            fence_proxy_async_global();
            fence_acq_rel_cluster();
            mbar_arrive_cluster_relaxed(h_ready + k, lane);
            if (lane == 0 && k == 0) LSTM_TRACE(7, s);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp >= EW0) {
    // ------------------------------------------------------------------ epilogue: 8 warps (see epilogue_role)
    if (REGSPLIT) asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int half = (warp - EW0) >> 2;
    const bool tracer = warp == EW0 && lane == 0;
#define BSRNN_EPI_CASE(QQ)                                                                                               \
  case QQ:                                                                                                               \
    if (half == 0) epilogue_role<QQ, 0, NS, 32>(a, tmem_base, warp, lane, cid, ncl, tracer, acc_full, acc_empty, w_free); \
    else epilogue_role<QQ, 1, NS, 32>(a, tmem_base, warp, lane, cid, ncl, tracer, acc_full, acc_empty, w_free);          \
    break;
    switch (q) {
      BSRNN_EPI_CASE(0) BSRNN_EPI_CASE(1) BSRNN_EPI_CASE(2) BSRNN_EPI_CASE(3)
      BSRNN_EPI_CASE(4) BSRNN_EPI_CASE(5) BSRNN_EPI_CASE(6) BSRNN_EPI_CASE(7)
    }
#undef BSRNN_EPI_CASE
  } else {
    if (REGSPLIT) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");      // idle 4th warp of warpgroup 0
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();                       // no CTA exits while peers may still arrive on its barriers
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * LACC);
  }
}

template <int NS, bool REGSPLIT>
static int max_active_clusters() {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(LCL * 64);
  cfg.blockDim = dim3(REGSPLIT ? 384 : 352);
  cfg.dynamicSmemBytes = L_SMEM;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = LCL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int n = 0;
  if (cudaFuncSetAttribute(lstm_tc_kernel<NS, REGSPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L_SMEM) != cudaSuccess)
    return -1;
  if (cudaOccupancyMaxActiveClusters(&n, lstm_tc_kernel<NS, REGSPLIT>, &cfg) != cudaSuccess) return -1;
  return n;
}

template <int NS, bool REGSPLIT>
static int launch_lstm(LstmTcArgs a, int max_clusters, cudaStream_t st) {
  static int max_active = -1;
  if (max_active < 0) {
    int n = max_active_clusters<NS, REGSPLIT>();
    if (n <= 0) {
      cudaGetLastError();
      set_error("blstm_recurrence_tc: no co-resident 8-CTA cluster for slots=%d regsplit=%d", NS, (int)REGSPLIT);
      return 2;
    }
    max_active = n;
  }
  a.gpd = (a.seq_tiles + NS - 1) / NS;
  int ncl = 2 * a.gpd;
  if (ncl > max_active) ncl = max_active;
  if (max_clusters > 0 && ncl > max_clusters) ncl = max_clusters;
  lstm_tc_kernel<NS, REGSPLIT><<<ncl * LCL, REGSPLIT ? 384 : 352, L_SMEM, st>>>(a);
  BSRNN_LAUNCH_OK();
  return 0;
}

}  // namespace bsrnn
using namespace bsrnn;

static long long* g_lstm_trace = nullptr;
extern "C" void bsrnn_debug_set_lstm_trace(void* p) { g_lstm_trace = reinterpret_cast<long long*>(p); }

// slots: interleaved sequence tiles per cluster (1..4; <= 0 = automatic); variant 0 = uniform register budget
// (slots <= 3), 1 = setmaxnreg register split (slots <= 4).
extern "C" int bsrnn_blstm_recurrence_tc_ex(const void* gates_x, const void* w_pack, void* y, int R, int steps,
                                            int seq_tiles, long seq_inner, long seq_outer, long seq_inner_stride,
                                            long step_stride, int max_clusters, int slots, int variant, void* stream) {
  BSRNN_CHECK_ARG(gates_x && w_pack && y, "blstm_recurrence_tc: null pointer");
  BSRNN_CHECK_ARG(R > 0 && steps > 0 && seq_tiles * 128 >= R && seq_inner > 0, "blstm_recurrence_tc: bad dims");
  LstmTcArgs a{reinterpret_cast<const __half*>(gates_x), reinterpret_cast<const __half*>(w_pack),
               reinterpret_cast<__half*>(y), R, steps, seq_tiles, 0, seq_inner, seq_outer, seq_inner_stride, step_stride,
               g_lstm_trace};
  cudaStream_t st = (cudaStream_t)stream;
  if (slots <= 0) slots = 3;
  if (slots > seq_tiles) slots = seq_tiles;
  if (variant == 0) {
    switch (slots) {
      case 1: return launch_lstm<1, false>(a, max_clusters, st);
      case 2: return launch_lstm<2, false>(a, max_clusters, st);
      case 3: return launch_lstm<3, false>(a, max_clusters, st);
    }
  } else if (variant == 1) {
    switch (slots) {
      case 3: return launch_lstm<3, true>(a, max_clusters, st);
      case 4: return launch_lstm<4, true>(a, max_clusters, st);
    }
  }
  set_error("blstm_recurrence_tc: unsupported slots=%d variant=%d", slots, variant);
  return 1;
}

static int g_slots = 0, g_variant = 0;
extern "C" int bsrnn_blstm_tc_configure(int slots, int variant) {
  g_slots = slots;
  g_variant = variant;
  return 0;
}

extern "C" int bsrnn_blstm_recurrence_tc(const void* gates_x, const void* w_pack, void* y, int R, int steps,
                                         int seq_tiles, long seq_inner, long seq_outer, long seq_inner_stride,
                                         long step_stride, int max_clusters, void* stream) {
  return bsrnn_blstm_recurrence_tc_ex(gates_x, w_pack, y, R, steps, seq_tiles, seq_inner, seq_outer, seq_inner_stride,
                                      step_stride, max_clusters, g_slots, g_variant, stream);
}

extern "C" int bsrnn_blstm_tc_max_clusters(void) { return max_active_clusters<3, false>(); }
