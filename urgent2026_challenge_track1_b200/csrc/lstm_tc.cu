// lstm_tc.cu — persistent BLSTM recurrence on tcgen05 tensor cores for H = 392 (BSRNN_baseline, N = 196).
//
// Replaces the recurrent half of nn.LSTM(N, 2N, bidirectional) [reference bsrnn_flowse.py:226-238, called at
// :296-297 (time axis: 2176 sequences x 1001 steps at BASELINE config 2) and :303-304 (band axis: 64064 x 34)].
//
// Decomposition.  A work unit is (direction d, tile j of 128 sequences).  A thread-block CLUSTER of 8 CTAs owns a
// GROUP of up to 3 units of one direction and advances them in lock step, slot after slot: the dependency chain of
// one unit (h published -> peers wake -> 100 KB h tile fetched from L2 -> 25 MMAs -> gates -> h stored -> published)
// is ~10 us long but occupies each resource (copy ring, tensor pipe, MUFU) for a fraction of that, so the chains of
// the slots overlap on the same CTAs.  CTA q of the cluster owns hidden units [49q, 49q+49) i.e. 196 of the 1568
// gate columns, and keeps that slice of W_hh (208 x 400 fp16, 166 KB, UMMA KB8 layout) resident in shared memory
// for the whole launch.  16 warps per CTA:
//   warp 0      producer : acquires the slot's cluster-scope h_ready mbarrier, bulk-copies the h_{t-1} tile
//                          (128 x 400 fp16 = 100 KB, from the y buffer in L2; a zero tile at step 0) through a
//                          3-stage ring,
//   warp 1      MMA      : 25 tcgen05.mma (M=128, N=208, K=16) per (step, slot) into one of two TMEM accumulators,
//   warp 2      publisher: completes the slot's named barrier, one fence, 8 relaxed remote arrives on the slot's
//                          h_ready barrier of every CTA of the cluster,
//   warps 4..15 epilogue : SLOT-SPECIALISED — slot k is served by warps 4+4k..7+4k (one per TMEM lane quadrant);
//                          a thread owns one sequence row of its slot for the whole launch: tcgen05.ld, add the
//                          precomputed input projection (fp16 tiles, coalesced, prefetched into L2 a step ahead),
//                          gates with MUFU tanh, c in 49 REGISTERS, h_t stored to y (at once the layer output the
//                          Linear GEMM consumes and the exchange buffer for the other 7 CTAs), then a
//                          non-blocking named-barrier arrive.  All slots execute the SAME code (the slot is a
//                          run-time offset), which keeps the hot loop inside the instruction cache: the previous
//                          version unrolled the slots and lost 43 % of its issue slots to instruction-fetch
//                          stalls (profiles/r01/call10_*).
// Control warps give their registers to the epilogue warps (setmaxnreg 40 / 152).
// No grid-wide synchronisation exists: clusters are independent, rows never mix.
//
// y layout (fp16): [step][seq_tile][dir][50 k-cores][128 rows][8]   (k-core 49 = zero padding, K = 400 per dir);
//                  a (step, seq_tile) block is therefore a 128 x 800 KB8 operand tile for the Linear(4N->N) GEMM.
// gates_x (fp16) : [step][seq_tile][dir][q][26 cores][128 rows][8]: column c = 4*u_local + gate (i,f,g,o) of CTA q,
//                  196 real + 12 pad — the KB8 tile the input-projection GEMM writes (epilogue 4).
// w_hh pack      : [dir][q][50 k-cores][208][8]
// The i, f, o gate rows of W_ih, W_hh and the bias are pre-multiplied by 0.5 at pack time: sigmoid(x) =
// 0.5*tanh(x/2) + 0.5 then costs one MUFU and one FMA.
#include "lstm_tc_common.cuh"
#include <stdlib.h>

namespace bsrnn {
using namespace umma;

constexpr uint32_t L_W_BYTES = LKC * LBN * 16;          // 166400
constexpr uint32_t L_A_STAGE = LKS * 128 * 16;          // 20480
constexpr int L_NBARS = 2 * LSTAGES + LNS + 2 + LNS + 2;
constexpr size_t L_SMEM = L_W_BYTES + LSTAGES * L_A_STAGE + L_NBARS * 8 + 16;
static_assert(L_SMEM <= 232448, "exceeds the 227 KB per-CTA shared memory limit");
// Ring geometry of lstm_tc_kernel (v4..v8): FKS k-cores per stage, FSTAGES stages.  Measured (profiles/r01/call31): a
// FINE ring (FKS = 2: one 4 KB stage per MMA, 15 stages, meant to keep ~56 KB in flight instead of ~40) is much
// SLOWER -- 13.3 / 10.4 ms against 7.2 / 6.3 ms: the producer then issues 25 bulk copies per item and each
// wait + expect_tx + cp.async.bulk round costs it ~400 cycles, i.e. 4 KB copies are issue-bound at ~10 B/clk.  The
// coarse ring (10 k-cores = 20 KB per copy, 3 stages) stays.
constexpr int FKS = 10;
constexpr int FNST = LKC / FKS;           // 5 stages per (step, slot)
constexpr int FSTAGES = 3;
constexpr uint32_t F_A_STAGE = FKS * 128 * 16;          // 20480
constexpr int F_NBARS = 2 * FSTAGES + LNS + 2 + LNS + 2;
constexpr size_t F_SMEM = L_W_BYTES + FSTAGES * F_A_STAGE + F_NBARS * 8 + 16;
static_assert(F_SMEM <= 232448, "exceeds the 227 KB per-CTA shared memory limit");

struct LstmTcArgs {
  const __half* gates_x;
  const __half* w_pack;
  const __half* zero_tile;   // 50*128*8 zeros: h_{-1}
  __half* y;
  int R, steps, seq_tiles;
  int gpd;                   // groups per direction (each group = up to `slots` consecutive sequence tiles)
  long long* probe;          // optional (debug): per-role wait/busy cycle totals of cluster probe_cid / CTA 0
  int probe_cid;
  unsigned* sync;            // flag-group schedule only: [0] ticket counter, [32 + 32*(3*group + slot)] h_ready counters
};

// ---- flag groups (lstm_tc_flag_kernel): the 8 CTAs that share a work unit need not be a hardware cluster.  The only
// cluster feature the recurrence uses is the remote mbarrier arrive that announces "h_t of slot k is in L2"; a
// gpu-scope release/acquire counter in global memory does the same job between ANY 8 co-resident CTAs.  That lifts
// the 15-cluster co-residency limit of 8-CTA clusters (GPC granularity: 120 of 148 SMs) to 18 groups = 144 SMs, and
// the time axis of BASELINE config 2 (34 units) fits with 2 interleaved tiles per group instead of 3.
constexpr int L_SYNC_WORDS = 32 + 32 * 3 * 32;     // ticket line + (up to 32 groups) x 3 slots, one 128-byte line each
__device__ __forceinline__ unsigned* flag_of(const LstmTcArgs& a, int cid, int k) { return a.sync + 32 + 32 * (3 * cid + k); }

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

// whole-launch accumulators (debug): P_MARK(v) adds the cycles since the previous mark to v.  Compiled in only with
// -DBSRNN_LSTM_PROBE (tools/prof_lstm.py --trace needs such a build): even predicated off, every mark left a clock
// read and a dependent add in the hot loops, and the ncu source page of v8 (profiles/r01/call41) showed ~15 % of the
// epilogue warps' stall samples on exactly those instructions.
#ifdef BSRNN_LSTM_PROBE
#define P_DECL(cond) const bool prb_ = a.probe && cid == a.probe_cid && q == 0 && (cond); long long pt_ = prb_ ? clock64() : 0
#define P_MARK(v)                      \
  do {                                 \
    if (prb_) {                        \
      const long long n_ = clock64();  \
      (v) += n_ - pt_;                 \
      pt_ = n_;                        \
    }                                  \
  } while (0)
#else
#define P_DECL(cond) constexpr bool prb_ = false
#define P_MARK(v) do { (void)(v); } while (0)
#endif


// Unit i of CTA Q is h column k = 49Q + i -> k-core 6Q + (Q+i)/8, slot (Q+i)%8 of the y tile.
// store_ready writes every 16-byte core (or the part of it this CTA owns) that became complete when the local units
// [D0, D1) were produced, so h values live in registers only until their core is full.
template <int Q, int D0, int D1>
__device__ __forceinline__ void store_ready(__half* ytile_row /* &y[...][kc=0][r][0] */, const float (&h)[LUN]) {
#pragma unroll
  for (int jj = 0; jj < 7; ++jj) {
    const int lo = 8 * jj - Q;                       // unit index sitting in slot 0 of this core
    const int a0 = lo > 0 ? lo : 0;
    const int b0 = (lo + 8) < LUN ? (lo + 8) : LUN;
    if (a0 >= b0) continue;                          // core holds none of our units
    if (!(b0 > D0 && b0 <= D1)) continue;            // completed earlier / not complete yet
    __half* dst = ytile_row + (size_t)(6 * Q + jj) * 128 * 8;
    if (a0 == lo && b0 == lo + 8) {
      __half2 p0 = __floats2half2_rn(h[lo], h[lo + 1]), p1 = __floats2half2_rn(h[lo + 2], h[lo + 3]);
      __half2 p2 = __floats2half2_rn(h[lo + 4], h[lo + 5]), p3 = __floats2half2_rn(h[lo + 6], h[lo + 7]);
      uint4 pk = make_uint4(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1),
                            *reinterpret_cast<uint32_t*>(&p2), *reinterpret_cast<uint32_t*>(&p3));
      *reinterpret_cast<uint4*>(dst) = pk;
    } else {
#pragma unroll
      for (int sl = 0; sl < 8; ++sl) {
        const int i = lo + sl;
        if (i >= a0 && i < b0) dst[sl] = __float2half_rn(h[i]);
      }
    }
  }
}

__device__ __forceinline__ void load_g4(const uint4* p, uint4 (&g)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) g[i] = __ldg(p + i * 128);             // cores are 128 rows x 16 B apart
}

template <int Q, int CH>
__device__ __forceinline__ void epilogue_chunk(uint32_t t_addr, const uint4 (&g)[4], __half* ytile_row, float (&c)[LUN],
                                               float (&h)[LUN]) {
  uint32_t acc[32];
  tmem_ld_x32(t_addr + CH * 32, acc);
  tmem_ld_wait();
  const __half2* gh = reinterpret_cast<const __half2*>(&g[0]);
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const float2 g01 = __half22float2(gh[2 * u]), g23 = __half22float2(gh[2 * u + 1]);
    gate_update(__uint_as_float(acc[4 * u]) + g01.x, __uint_as_float(acc[4 * u + 1]) + g01.y,
                __uint_as_float(acc[4 * u + 2]) + g23.x, __uint_as_float(acc[4 * u + 3]) + g23.y, c[CH * 8 + u],
                h[CH * 8 + u]);
  }
  store_ready<Q, CH * 8, CH * 8 + 8>(ytile_row, h);
}

// One (step, slot) item for a thread: row r of the tile, all 49 units of CTA Q.  gp -> this row's 16 bytes of core 0.
// Rows beyond R (padding of the last sequence tile) are computed like any other: their gates_x rows hold the bias
// (the GEMM ran on zero operand rows), rows never mix, and nothing downstream reads them.
template <int Q>
__device__ __forceinline__ void epilogue_item(uint32_t t_addr, const uint4* gp, __half* ytile_row, float (&c)[LUN]) {
  float h[LUN];
  uint4 ga[4], gb[4];
  load_g4(gp, ga);
  load_g4(gp + 4 * 128, gb);
  epilogue_chunk<Q, 0>(t_addr, ga, ytile_row, c, h);
  load_g4(gp + 8 * 128, ga);
  epilogue_chunk<Q, 1>(t_addr, gb, ytile_row, c, h);
  load_g4(gp + 12 * 128, gb);
  epilogue_chunk<Q, 2>(t_addr, ga, ytile_row, c, h);
  load_g4(gp + 16 * 128, ga);
  epilogue_chunk<Q, 3>(t_addr, gb, ytile_row, c, h);
  load_g4(gp + 20 * 128, gb);
  epilogue_chunk<Q, 4>(t_addr, ga, ytile_row, c, h);
  const uint2 gt = __ldg(reinterpret_cast<const uint2*>(gp + 24 * 128));
  epilogue_chunk<Q, 5>(t_addr, gb, ytile_row, c, h);
  {
    uint32_t acc[4];
    tmem_ld_x4(t_addr + 192, acc);
    tmem_ld_wait();
    const __half2* gh = reinterpret_cast<const __half2*>(&gt);
    const float2 g01 = __half22float2(gh[0]), g23 = __half22float2(gh[1]);
    gate_update(__uint_as_float(acc[0]) + g01.x, __uint_as_float(acc[1]) + g01.y, __uint_as_float(acc[2]) + g23.x,
                __uint_as_float(acc[3]) + g23.y, c[48], h[48]);
    store_ready<Q, 48, 49>(ytile_row, h);
  }
}

// The epilogue role of one warp: slot k, TMEM lane quadrant `quad`; loops over this cluster's groups and steps.
template <int Q>
__device__ __forceinline__ void epilogue_role(const LstmTcArgs& a, uint32_t tmem_base, int k, int quad, int lane, int cid,
                                              int ncl, uint64_t* acc_full, uint64_t* acc_empty, uint64_t* w_free) {
  const int q = Q;
  const int r = quad * 32 + lane;
  const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
  const int ngroups = 2 * a.gpd;
  const size_t y_tile = (size_t)LKC * 128 * 8;                 // halves per (step, tile, dir) of y
  const size_t g_tile = (size_t)2 * LCL * LGC * 128 * 8;       // halves per (step, tile) of gates_x
  uint32_t it0 = 0;                                            // accumulator items of the groups done so far
  uint32_t nfull = 0;                                          // items this slot has consumed (phase of acc_full[k])
  long long w_acc = 0, w_busy = 0;
  P_DECL(k == 0 && quad == 0 && lane == 0);
  float c[LUN];
  for (int g = cid; g < ngroups; g += ncl) {
    const Group G = group_of(a, g);
    if (k < G.nact) {
      const int j = G.j0 + k;
#pragma unroll
      for (int i = 0; i < LUN; ++i) c[i] = 0.f;
      for (int s = 0; s < a.steps; ++s) {
        const int p = G.d == 0 ? s : a.steps - 1 - s;
        const size_t tile = (size_t)p * a.seq_tiles + j;
        const __half* gbase = a.gates_x + tile * g_tile + (size_t)(G.d * LCL + Q) * (LGC * 128 * 8);
        __half* ytile_row = a.y + (tile * 2 + G.d) * y_tile + (size_t)r * 8;
        if (s + 1 < a.steps) {                       // next step's input projection (53 KB) -> L2
          const long step_off = (G.d == 0 ? 1 : -1) * (long)a.seq_tiles * (long)g_tile;
          const char* nx = reinterpret_cast<const char*>(gbase + step_off) + r * 128;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (i * 128 + r < LGC * 16) prefetch_l2(nx + i * 128 * 128);
        }
        const uint32_t it = it0 + (uint32_t)(s * G.nact + k);
        const uint32_t buf = it & 1;
        P_MARK(w_busy);
        mbar_wait(acc_full + k, nfull & 1);
        ++nfull;
        tc_fence_after();
        P_MARK(w_acc);
        epilogue_item<Q>(t_lane + buf * LACC, reinterpret_cast<const uint4*>(gbase) + r, ytile_row, c);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty + buf);
        // h_t slice of this warp is stored: tell the publisher (non-blocking)
        if (s + 1 < a.steps) asm volatile("bar.arrive %0, %1;" ::"r"(1 + k), "n"(128 + 32) : "memory");
      }
      const int gn = g + ncl;                        // next group of this cluster switches direction?
      if (k == G.nact - 1 && gn < ngroups && gn / a.gpd != G.d) {
        __syncwarp();
        if (lane == 0) mbar_arrive(w_free);          // the last slot's last accumulator is consumed: W may go
      }
    }
    it0 += (uint32_t)(a.steps * G.nact);
  }
  P_MARK(w_busy);
  if (prb_) { a.probe[12] = w_acc; a.probe[13] = w_busy; }
}


// ------------------------------------------------------------------------------------------------ v5 epilogue
// All 12 epilogue warps serve EVERY accumulator item: thread = (row r, third T); third T owns the local units
// [16T, 16T+16) (third 2 also unit 48).  One item then costs a warp 16-17 units instead of 49, the three warps of
// an SM sub-partition run the same instructions at the same time (one instruction-cache stream instead of three),
// and the item's input-projection row segment (8 x 16 B) is already in registers when the accumulator arrives
// because the loads are issued BEFORE the wait.  profiles/r01/call19 (v4): 34 % of the epilogue's issue slots
// were lost to instruction fetch and 23 % to L2 latency of those loads.
//
// Unit j of third T is h column 49Q + 16T + j -> k-core 6Q + 2T + (Q+j)/8, slot (Q+j)%8: the store pattern of a
// third depends on Q only (compile time); T is a run-time offset.

// One item for one thread.  t_col: TMEM address of this thread's lane, accumulator column 64T.  g: the row's 8
// gates_x cores of this third.  ycore: &y[..][k-core 6Q + 2T][r][0].
template <int Q>
__device__ __forceinline__ void epi5_item(uint32_t t_col, bool last_third, uint32_t t_col48, const uint4 (&g)[8],
                                          const uint2 g48, __half* ycore, float (&c)[17], bool st = true) {
  constexpr size_t CORE = 128 * 8;                  // halves between consecutive k-cores of a y tile
  float h[16];
  uint32_t acc[32];
  const uint32_t* gw = reinterpret_cast<const uint32_t*>(&g[0]);
  tmem_ld_x32(t_col, acc);
  tmem_ld_wait();
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const float2 p01 = add_h2_f32(gw[2 * u], acc[4 * u], acc[4 * u + 1]);
    const float2 p23 = add_h2_f32(gw[2 * u + 1], acc[4 * u + 2], acc[4 * u + 3]);
    gate_update(p01.x, p01.y, p23.x, p23.y, c[u], h[u]);
  }
  tmem_ld_x32(t_col + 32, acc);
  if (st) {
    if (Q == 0) store_full<0>(ycore, h);
    else store_partial<Q, 8, 0>(ycore, h);          // core A: slots Q..7 <- j = 0..7-Q
  }
  tmem_ld_wait();
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const float2 p01 = add_h2_f32(gw[16 + 2 * u], acc[4 * u], acc[4 * u + 1]);
    const float2 p23 = add_h2_f32(gw[16 + 2 * u + 1], acc[4 * u + 2], acc[4 * u + 3]);
    gate_update(p01.x, p01.y, p23.x, p23.y, c[8 + u], h[8 + u]);
  }
  if (st) {
    store_full<8 - Q>(ycore + CORE, h);             // core B: j = 8-Q .. 15-Q
    if (Q > 0) store_partial<0, Q, 16 - Q>(ycore + 2 * CORE, h);   // core C: slots 0..Q-1 <- j = 16-Q .. 15
  }
  if (last_third) {                                 // local unit 48 -> slot Q of core C
    uint32_t a4[4];
    tmem_ld_x4(t_col48, a4);
    tmem_ld_wait();
    const float2 p01 = add_h2_f32(g48.x, a4[0], a4[1]), p23 = add_h2_f32(g48.y, a4[2], a4[3]);
    float h48;
    gate_update(p01.x, p01.y, p23.x, p23.y, c[16], h48);
    if (st) ycore[2 * CORE + Q] = __float2half_rn(h48);
  }
}

// Pipelined variant (v6): 4 chunks of 4 units (16 accumulator columns each, so 16 instead of 32 registers hold the
// accumulator), and the gates registers of a chunk are refilled with the NEXT item's values as soon as they are consumed.
template <int Q>
__device__ __forceinline__ void epi6_item(uint32_t t_col, bool last_third, uint32_t t_col48, uint4 (&g)[8], uint2& g48,
                                          const uint4* gnext, bool have_next, __half* ycore, float (&c)[17]) {
  constexpr size_t CORE = 128 * 8;
  float h[16];
  uint32_t acc[16];
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    tmem_ld_x16(t_col + 16 * ch, acc);
    tmem_ld_wait();
    const __half2* gh = reinterpret_cast<const __half2*>(&g[2 * ch]);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float2 g01 = __half22float2(gh[2 * u]), g23 = __half22float2(gh[2 * u + 1]);
      gate_update(__uint_as_float(acc[4 * u]) + g01.x, __uint_as_float(acc[4 * u + 1]) + g01.y,
                  __uint_as_float(acc[4 * u + 2]) + g23.x, __uint_as_float(acc[4 * u + 3]) + g23.y, c[4 * ch + u],
                  h[4 * ch + u]);
    }
    if (have_next) {
      g[2 * ch] = __ldg(gnext + (2 * ch) * 128);
      g[2 * ch + 1] = __ldg(gnext + (2 * ch + 1) * 128);
    }
    if (ch == 1) {
      if (Q == 0) store_full<0>(ycore, h);
      else store_partial<Q, 8, 0>(ycore, h);          // core A: slots Q..7 <- j = 0..7-Q
    }
  }
  store_full<8 - Q>(ycore + CORE, h);                 // core B: j = 8-Q .. 15-Q
  if (Q > 0) store_partial<0, Q, 16 - Q>(ycore + 2 * CORE, h);   // core C: slots 0..Q-1 <- j = 16-Q .. 15
  if (last_third) {                                   // local unit 48 -> slot Q of core C
    uint32_t a4[4];
    tmem_ld_x4(t_col48, a4);
    tmem_ld_wait();
    const __half2* g4 = reinterpret_cast<const __half2*>(&g48);
    const float2 g01 = __half22float2(g4[0]), g23 = __half22float2(g4[1]);
    float h48;
    gate_update(__uint_as_float(a4[0]) + g01.x, __uint_as_float(a4[1]) + g01.y, __uint_as_float(a4[2]) + g23.x,
                __uint_as_float(a4[3]) + g23.y, c[16], h48);
    if (have_next) g48 = __ldg(reinterpret_cast<const uint2*>(gnext + (24 - 8 * 2) * 128));   // core 24 (T = 2: gnext -> core 16)
    ycore[2 * CORE + Q] = __float2half_rn(h48);
  }
}

// v8 item: v6's register-refilled gates_x (the NEXT item's input projection replaces each pair of cores as soon as it
// is consumed, so no global-load latency is exposed after the accumulator wait) + FHADD gate adds + software-pipelined
// TMEM loads: 4 chunks of 4 units (16 columns), the load of chunk j+1 in flight while chunk j runs through the MUFU
// pipe.  profiles/r01/call25 (v5, ncu source page): the epilogue paces both axes (~4 700 cycles per item against 2 600
// of MMA and a 1 960-cycle MUFU floor = 5 tanh x 49 units x 128 rows / 16 per clock); its MUFU pipe idles while all 12
// warps sit in the same TMEM-load / gate-load / store phases.
template <int CH>
__device__ __forceinline__ void epi8_chunk(const uint32_t (&acc)[16], uint4 (&g)[8], const uint4* gnext, bool have_next,
                                           float (&c)[17], float (&h)[16]) {
  const uint32_t* gw = reinterpret_cast<const uint32_t*>(&g[2 * CH]);
  float2 p01[4], p23[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    p01[u] = add_h2_f32(gw[2 * u], acc[4 * u], acc[4 * u + 1]);
    p23[u] = add_h2_f32(gw[2 * u + 1], acc[4 * u + 2], acc[4 * u + 3]);
  }
  if (have_next) {                                   // the two cores are consumed: fetch the next item's
    g[2 * CH] = __ldg(gnext + (2 * CH) * 128);
    g[2 * CH + 1] = __ldg(gnext + (2 * CH + 1) * 128);
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) gate_update(p01[u].x, p01[u].y, p23[u].x, p23[u].y, c[4 * CH + u], h[4 * CH + u]);
}
template <int Q>
__device__ __forceinline__ void epi8_item(uint32_t t_col, bool last_third, uint32_t t_col48, uint4 (&g)[8], uint2& g48,
                                          const uint4* gnext, bool have_next, __half* ycore, float (&c)[17]) {
  constexpr size_t CORE = 128 * 8;
  float h[16];
  uint32_t accA[16], accB[16];
  tmem_ld_x16(t_col, accA);
  tmem_ld_wait();
  tmem_ld_x16(t_col + 16, accB);
  tmem_ld_pin16(accA);
  epi8_chunk<0>(accA, g, gnext, have_next, c, h);
  tmem_ld_wait();
  tmem_ld_x16(t_col + 32, accA);
  tmem_ld_pin16(accB);
  epi8_chunk<1>(accB, g, gnext, have_next, c, h);
  if (Q == 0) store_full<0>(ycore, h);
  else store_partial<Q, 8, 0>(ycore, h);              // core A: slots Q..7 <- j = 0..7-Q
  tmem_ld_wait();
  tmem_ld_x16(t_col + 48, accB);
  tmem_ld_pin16(accA);
  epi8_chunk<2>(accA, g, gnext, have_next, c, h);
  uint32_t a4[4] = {0u, 0u, 0u, 0u};
  tmem_ld_wait();
  if (last_third) tmem_ld_x4(t_col48, a4);
  tmem_ld_pin16(accB);
  epi8_chunk<3>(accB, g, gnext, have_next, c, h);
  store_full<8 - Q>(ycore + CORE, h);                 // core B: j = 8-Q .. 15-Q
  if (Q > 0) store_partial<0, Q, 16 - Q>(ycore + 2 * CORE, h);   // core C: slots 0..Q-1 <- j = 16-Q .. 15
  if (last_third) {                                   // local unit 48 -> slot Q of core C
    tmem_ld_wait();
    asm volatile("" : "+r"(a4[0]), "+r"(a4[1]), "+r"(a4[2]), "+r"(a4[3]));
    const float2 p01 = add_h2_f32(g48.x, a4[0], a4[1]), p23 = add_h2_f32(g48.y, a4[2], a4[3]);
    float h48;
    gate_update(p01.x, p01.y, p23.x, p23.y, c[16], h48);
    if (have_next) g48 = __ldg(reinterpret_cast<const uint2*>(gnext + (24 - 8 * 2) * 128));   // core 24 (T = 2: gnext -> core 16)
    ycore[2 * CORE + Q] = __float2half_rn(h48);
  }
}

template <int Q, int MODE>   // 0: v5 item, 1: v6 item (gates refilled in registers), 2: v8 item
__device__ __forceinline__ void epilogue5_role(const LstmTcArgs& a, uint32_t tmem_base, int T, int quad, int lane, int cid,
                                               int ncl, uint64_t* acc_full, uint64_t* acc_empty, uint64_t* w_free) {
  constexpr bool PIPE = MODE != 0;
  const int q = Q;
  const int r = quad * 32 + lane;
  const bool last_third = T == 2;
  const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16) + 64 * T;
  const uint32_t t_lane48 = tmem_base + ((uint32_t)(quad * 32) << 16) + 192;
  const int ngroups = 2 * a.gpd;
  const size_t y_tile = (size_t)LKC * 128 * 8;                 // halves per (step, tile, dir) of y
  const size_t g_tile = (size_t)2 * LCL * LGC * 128 * 8;       // halves per (step, tile) of gates_x
  const int pf_idx = (T * 4 + quad) * 32 + lane;               // 0..383: share of the next step's L2 prefetch
  uint32_t it0 = 0;                                            // accumulator items of the groups done so far
  uint32_t nfull = 0;                                          // bit k: parity of the next wait on acc_full[k]
  long long w_acc = 0, w_busy = 0;
  P_DECL(T == 0 && quad == 0 && lane == 0);
  float c0[17], c1[17], c2[17];
  for (int g = cid; g < ngroups; g += ncl) {
    const Group G = group_of(a, g);
#pragma unroll
    for (int i = 0; i < 17; ++i) c0[i] = c1[i] = c2[i] = 0.f;
    uint4 gg[8];
    uint2 g48 = make_uint2(0u, 0u);
    if (PIPE) {                                      // first item of the group: (step 0, slot 0)
      const size_t tile = (size_t)(G.d == 0 ? 0 : a.steps - 1) * a.seq_tiles + G.j0;
      const uint4* gb4 = reinterpret_cast<const uint4*>(a.gates_x + tile * g_tile + (size_t)(G.d * LCL + Q) * (LGC * 128 * 8));
#pragma unroll
      for (int i = 0; i < 8; ++i) gg[i] = __ldg(gb4 + (size_t)(8 * T + i) * 128 + r);
      if (last_third) g48 = __ldg(reinterpret_cast<const uint2*>(gb4 + 24 * 128 + r));
    }
    for (int s = 0; s < a.steps; ++s) {
      const int p = G.d == 0 ? s : a.steps - 1 - s;
#pragma unroll
      for (int k = 0; k < LNS; ++k) {
        if (k < G.nact) {
          const size_t tile = (size_t)p * a.seq_tiles + (G.j0 + k);
          const __half* gbase = a.gates_x + tile * g_tile + (size_t)(G.d * LCL + Q) * (LGC * 128 * 8);
          __half* ycore = a.y + (tile * 2 + G.d) * y_tile + (size_t)(6 * Q + 2 * T) * (128 * 8) + (size_t)r * 8;
          const uint4* gp = reinterpret_cast<const uint4*>(gbase) + (size_t)(8 * T) * 128 + r;
          const uint4* gnext = gp;                   // PIPE: this thread's segment of the next item it will serve
          bool have_next = false;
          if (PIPE) {
            const bool wrap = k + 1 >= G.nact;
            have_next = !wrap || s + 1 < a.steps;
            // tiles of one step are consecutive; the next step is seq_tiles tiles further (forward) or back (reverse)
            const long dt = wrap ? (long)(G.d == 0 ? a.seq_tiles : -a.seq_tiles) - (G.nact - 1) : 1;
            gnext = gp + dt * (long)(g_tile / 8);
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) gg[i] = __ldg(gp + i * 128);
            if (last_third) g48 = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint4*>(gbase) + 24 * 128 + r));
          }
          // next step's input projection (53 KB = 416 lines) -> L2.  Not in v8: its register refill already runs a
          // whole item (> 4 000 cycles) ahead of use, which covers a DRAM round trip without the prefetch instructions.
          if (MODE != 2 && s + 1 < a.steps) {
            const long step_off = (G.d == 0 ? 1 : -1) * (long)a.seq_tiles * (long)g_tile;
            const char* nx = reinterpret_cast<const char*>(gbase + step_off);
            prefetch_l2(nx + pf_idx * 128);
            if (pf_idx < LGC * 16 - 384) prefetch_l2(nx + (384 + pf_idx) * 128);
          }
          const uint32_t it = it0 + (uint32_t)(s * G.nact + k);
          const uint32_t buf = it & 1;
          P_MARK(w_busy);
          mbar_wait(acc_full + k, (nfull >> k) & 1);
          nfull ^= 1u << k;
          tc_fence_after();
          P_MARK(w_acc);
          if (MODE == 2) {
            if (k == 0) epi8_item<Q>(t_lane + buf * LACC, last_third, t_lane48 + buf * LACC, gg, g48, gnext, have_next, ycore, c0);
            else if (k == 1) epi8_item<Q>(t_lane + buf * LACC, last_third, t_lane48 + buf * LACC, gg, g48, gnext, have_next, ycore, c1);
            else epi8_item<Q>(t_lane + buf * LACC, last_third, t_lane48 + buf * LACC, gg, g48, gnext, have_next, ycore, c2);
          } else if (PIPE) {
            if (k == 0) epi6_item<Q>(t_lane + buf * LACC, last_third, t_lane48 + buf * LACC, gg, g48, gnext, have_next, ycore, c0);
            else if (k == 1) epi6_item<Q>(t_lane + buf * LACC, last_third, t_lane48 + buf * LACC, gg, g48, gnext, have_next, ycore, c1);
            else epi6_item<Q>(t_lane + buf * LACC, last_third, t_lane48 + buf * LACC, gg, g48, gnext, have_next, ycore, c2);
          } else {
            if (k == 0) epi5_item<Q>(t_lane + buf * LACC, last_third, t_lane48 + buf * LACC, gg, g48, ycore, c0);
            else if (k == 1) epi5_item<Q>(t_lane + buf * LACC, last_third, t_lane48 + buf * LACC, gg, g48, ycore, c1);
            else epi5_item<Q>(t_lane + buf * LACC, last_third, t_lane48 + buf * LACC, gg, g48, ycore, c2);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty + buf);
          // h_t slice of this warp is stored: tell the publisher (non-blocking)
          if (s + 1 < a.steps) asm volatile("bar.arrive %0, %1;" ::"r"(1 + k), "n"(384 + 32) : "memory");
        }
      }
    }
    const int gn = g + ncl;                          // next group of this cluster switches direction?
    if (gn < ngroups && gn / a.gpd != G.d) {
      __syncwarp();
      if (lane == 0) mbar_arrive(w_free);            // this warp consumed the last accumulator: W may go
    }
    it0 += (uint32_t)(a.steps * G.nact);
  }
  P_MARK(w_busy);
  if (prb_) { a.probe[12] = w_acc; a.probe[13] = w_busy; }
  (void)q;
}

// VER 4: slot-specialised epilogue; 5: all-warps epilogue; 6: 5 + multicast h loads + pipelined gate loads;
//     8: 5 + register-refilled gates, FHADD, software-pipelined TMEM loads (epi8_item)
template <int VER, bool GF>   // GF: flag groups instead of clusters (see above)
__device__ __forceinline__ void lstm_tc_body(const LstmTcArgs& a) {
  constexpr bool V5 = VER >= 5;
  constexpr bool MC = VER == 6 || VER == 9;          // 9: v8 epilogue + multicast h loads
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sW = smem;
  uint8_t* sA = smem + L_W_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + FSTAGES * F_A_STAGE);
  uint64_t* full = bars;                       // [3]
  uint64_t* empty = bars + FSTAGES;            // [3]
  uint64_t* acc_full = bars + 2 * FSTAGES;     // [LNS]: one per SLOT — each slot's warps see every phase of theirs
  uint64_t* acc_empty = acc_full + LNS;        // [2]: one per TMEM buffer — the MMA thread sees every phase
  uint64_t* h_ready = acc_empty + 2;           // [LNS]
  uint64_t* w_full = h_ready + LNS;
  uint64_t* w_free = w_full + 1;               // epilogue -> producer: the resident W slice may be overwritten
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_free + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t q;
  int cid, ncl;
  if (GF) {
    // rank within the launch by arrival order: group = ticket / 8, unit slice = ticket % 8 (any placement works)
    if (threadIdx.x == 0) tmem_slot[1] = atomicAdd(a.sync, 1u);
    __syncthreads();
    const uint32_t ticket = tmem_slot[1];
    q = ticket % LCL;
    cid = (int)(ticket / LCL);
    ncl = (int)(gridDim.x / LCL);
  } else {
    q = cluster_ctarank();
    cid = cluster_id_x();
    ncl = num_clusters_x();
  }

  if (threadIdx.x == 0) {
    for (int i = 0; i < FSTAGES; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, MC ? LCL : 1); }
    for (int i = 0; i < LNS; ++i) mbar_init(acc_full + i, 1);
    for (int i = 0; i < 2; ++i) mbar_init(acc_empty + i, V5 ? 12 : 4);
    for (int i = 0; i < LNS; ++i) mbar_init(h_ready + i, LCL);
    mbar_init(w_full, 1);
    mbar_init(w_free, V5 ? 12 : 4);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 2 * LACC);
  tc_fence_before();
  __syncthreads();
  if (!GF) cluster_sync();              // every CTA's barriers are initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int ngroups = 2 * a.gpd;
  const size_t y_tile = (size_t)LKC * 128 * 8;         // halves per (step, tile, dir)

  // The producer and MMA warps run CONVERGED: all 32 lanes walk the schedule and wait on the barriers, the bulk
  // copies / MMAs / commits are issued inside `if (elect_one())`.  Under `if (lane == 0)` the compiler wraps each
  // UTCHMMA / UBLKCP / UTCBAR in an elect-and-retry loop with descriptors rebuilt in vector registers (~15
  // dependent instructions per MMA): the issue thread, not the tensor pipe, then paces the kernel
  // (profiles/r01/call26, same pattern in the GEMM).
  if (warp == 0) {
    // ------------------------------------------------------------------ producer: W slice + h tiles
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    {
      uint32_t stage = 0, phase = 0, hphase = 0, wfphase = 0;
      uint32_t npub0 = 0u, npub1 = 0u, npub2 = 0u;   // GF: h tiles of slot k published so far by each CTA of the group
      int cur_dir = -1;
      long long w_h = 0, w_e = 0, w_o = 0;
      P_DECL(true);
      for (int g = cid; g < ngroups; g += ncl) {
        const Group G = group_of(a, g);
        if (G.d != cur_dir) {
          // the epilogue warps of the last slot signal w_free after consuming the last accumulator of the previous
          // direction, i.e. after every MMA that read the old slice has retired.
          if (cur_dir >= 0) {
            mbar_wait(w_free, wfphase);
            wfphase ^= 1;
          }
          if (elect_one()) {
            mbar_expect_tx(w_full, L_W_BYTES);
            const uint8_t* src = reinterpret_cast<const uint8_t*>(a.w_pack) + ((size_t)G.d * LCL + q) * L_W_BYTES;
            for (uint32_t off = 0; off < L_W_BYTES; off += 33280) bulk_g2s(sW + off, src + off, 33280, w_full);
          }
          __syncwarp();
          cur_dir = G.d;
        }
        for (int s = 0; s < a.steps; ++s) {
          const int p_prev = G.d == 0 ? s - 1 : a.steps - s;        // position whose h feeds this step
          for (int k = 0; k < G.nact; ++k) {
            const uint8_t* src = reinterpret_cast<const uint8_t*>(a.zero_tile);
            if (s > 0) {
              P_MARK(w_o);
              if (GF) {
                // all 8 CTAs of the group have released their slice of h_{t-1} (gpu-scope acquire on the counter)
                uint32_t& np = k == 0 ? npub0 : (k == 1 ? npub1 : npub2);
                const uint32_t want = LCL * (++np);
                const unsigned* fl = flag_of(a, cid, k);
                uint32_t spins = 0;
                while ((int32_t)(ld_acquire_gpu(fl) - want) < 0) {
                  if (++spins > (1u << 22)) { __trap(); }
                }
              } else {
                mbar_wait_cluster(h_ready + k, (hphase >> k) & 1);
              }
              P_MARK(w_h);
              hphase ^= 1u << k;
              src = reinterpret_cast<const uint8_t*>(a.y + (((size_t)p_prev * a.seq_tiles + (G.j0 + k)) * 2 + G.d) * y_tile);
            }
            for (int ks = 0; ks < FNST; ++ks) {
              P_MARK(w_o);
              mbar_wait(empty + stage, phase ^ 1);
              P_MARK(w_e);
              if (elect_one()) {
                if (s > 0 && ks == 0) fence_proxy_async_global();      // peers' generic-proxy h stores -> this thread's async-proxy reads
                mbar_expect_tx(full + stage, F_A_STAGE);
                if (MC) {
                  // every CTA of the cluster needs the same h tile: each fetches 1/8 of the stage and multicasts it into
                  // the same ring slot of all 8 (empty[stage] counts the 8 MMA commits, so the slot is free everywhere)
                  constexpr uint32_t SL = F_A_STAGE / LCL;
                  bulk_g2s_multicast(sA + stage * F_A_STAGE + q * SL, src + (size_t)ks * F_A_STAGE + q * SL, SL, full + stage,
                                     (uint16_t)((1u << LCL) - 1));
                } else {
                  bulk_g2s(sA + stage * F_A_STAGE, src + (size_t)ks * F_A_STAGE, F_A_STAGE, full + stage);
                }
              }
              __syncwarp();
              if (++stage == FSTAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
      P_MARK(w_o);
      if (prb_ && lane == 0) { a.probe[0] = w_h; a.probe[1] = w_e; a.probe[2] = w_o; }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    {
      const uint32_t idesc = idesc_f16_f32(128, LBN);
      uint32_t stage = 0, phase = 0, it = 0, wphase = 0;
      int cur_dir = -1;
      // descriptors advance by adding to the 14-bit (address >> 4) field: shared memory is < 256 KB, no carry out
      const uint64_t da0 = smem_desc_kb8(smem_u32(sA), 2048, 128);
      const uint64_t db0 = smem_desc_kb8(smem_u32(sW), LBN * 16, 128);
      long long w_a = 0, w_f = 0, w_o = 0;
      P_DECL(true);
      for (int g = cid; g < ngroups; g += ncl) {
        const Group G = group_of(a, g);
        if (G.d != cur_dir) {
          mbar_wait(w_full, wphase);
          wphase ^= 1;
          cur_dir = G.d;
        }
        const int nitems = a.steps * G.nact;
        int slot = 0;
        for (int i = 0; i < nitems; ++i, ++it) {
          const uint32_t buf = it & 1;
          P_MARK(w_o);
          mbar_wait(acc_empty + buf, ((it >> 1) & 1) ^ 1);
          P_MARK(w_a);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * LACC;
#pragma unroll
          for (int ks = 0; ks < FNST; ++ks) {
            P_MARK(w_o);
            mbar_wait(full + stage, phase);
            P_MARK(w_f);
            tc_fence_after();
            if (elect_one()) {
              const uint64_t da = da0 + (uint64_t)(stage * (F_A_STAGE >> 4));
              const uint64_t db = db0 + (uint64_t)(ks * FKS * ((LBN * 16) >> 4));
#pragma unroll
              for (int jk = 0; jk < FKS / 2; ++jk)
                mma_f16_ss(d_tmem, da + (uint64_t)(jk * ((2 * 2048) >> 4)), db + (uint64_t)(jk * ((2 * LBN * 16) >> 4)), idesc,
                           (ks | jk) != 0);
              if (MC) mma_commit_multicast(empty + stage, (uint16_t)((1u << LCL) - 1));
              else mma_commit(empty + stage);
              if (ks == FNST - 1) mma_commit(acc_full + slot);
            }
            __syncwarp();
            if (++stage == FSTAGES) { stage = 0; phase ^= 1; }
          }
          if (++slot == G.nact) slot = 0;
        }
      }
      P_MARK(w_o);
      if (prb_ && lane == 0) { a.probe[4] = w_a; a.probe[5] = w_f; a.probe[6] = w_o; }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ publisher
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    for (int g = cid; g < ngroups; g += ncl) {
      const Group G = group_of(a, g);
      for (int s = 0; s + 1 < a.steps; ++s) {
        for (int k = 0; k < G.nact; ++k) {
          // completes once the 4 epilogue warps of slot k have stored their h_t slices
          if (V5) asm volatile("bar.sync %0, %1;" ::"r"(1 + k), "n"(384 + 32) : "memory");
          else asm volatile("bar.sync %0, %1;" ::"r"(1 + k), "n"(128 + 32) : "memory");
          if (GF) {
            if (lane == 0) {
              fence_proxy_async_global();
              red_release_gpu_add(flag_of(a, cid, k), 1u);     // release: the CTA's h stores (ordered by the bar.sync) first
            }
          } else if (lane < LCL) {
            fence_proxy_async_global();
            fence_acq_rel_cluster();
            mbar_arrive_cluster_relaxed(h_ready + k, lane);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 3) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");      // idle 4th warp of warpgroup 0
  } else {
    // ------------------------------------------------------------------ epilogue: 3 slots x 4 warps
    asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
    const int k = (warp - 4) >> 2, quad = warp & 3;
#define BSRNN_EPI_CASE(QQ)                                                                                   \
  case QQ:                                                                                                   \
    if (VER == 8 || VER == 9) epilogue5_role<QQ, 2>(a, tmem_base, k, quad, lane, cid, ncl, acc_full, acc_empty, w_free);  \
    else if (MC) epilogue5_role<QQ, 1>(a, tmem_base, k, quad, lane, cid, ncl, acc_full, acc_empty, w_free);  \
    else if (V5) epilogue5_role<QQ, 0>(a, tmem_base, k, quad, lane, cid, ncl, acc_full, acc_empty, w_free);  \
    else epilogue_role<QQ>(a, tmem_base, k, quad, lane, cid, ncl, acc_full, acc_empty, w_free);              \
    break;
    switch (q) {
      BSRNN_EPI_CASE(0) BSRNN_EPI_CASE(1) BSRNN_EPI_CASE(2) BSRNN_EPI_CASE(3)
      BSRNN_EPI_CASE(4) BSRNN_EPI_CASE(5) BSRNN_EPI_CASE(6) BSRNN_EPI_CASE(7)
    }
#undef BSRNN_EPI_CASE
  }
  tc_fence_before();
  __syncthreads();
  if (!GF) cluster_sync();              // no CTA exits while peers may still arrive on its barriers
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * LACC);
  }
}

template <int VER>
__global__ void __cluster_dims__(LCL, 1, 1) __launch_bounds__(LTHREADS, 1) lstm_tc_kernel(const LstmTcArgs a) {
  lstm_tc_body<VER, false>(a);
}
// flag-group launch: ordinary grid (no cluster), groups of 8 CTAs by arrival ticket; all CTAs must be co-resident
// (grid <= SMs x 1 CTA/SM, checked by the host) because the groups spin on each other's counters.
__global__ void __launch_bounds__(LTHREADS, 1) lstm_tc_flag_kernel(const LstmTcArgs a) { lstm_tc_body<8, true>(a); }

// ================================================================================================ v7: CTA pairs
// Same decomposition as v5, but the cluster has 16 CTAs = 8 CTA PAIRS and every MMA is a cta_group::2 instruction
// (M = 256, N = 208): pair P = rank/2 owns hidden units [49P, 49P+49); its even CTA serves sequence tile 2j, its odd
// CTA tile 2j+1 of the slot's tile pair j, and each keeps only HALF of the pair's W_hh slice (104 x 400 fp16 = 83 KB)
// in shared memory.  What that buys over v5 (profiles/r01/call20 ncu: tensor pipe 53 % of active cycles, shared-memory
// tensor wavefronts at 81 % of the tensor-active share, producer blocked on a 60 KB ring):
//   * the tensor core of each SM reads A (4 KB) + half of B (3.3 KB) per MMA instead of 4 + 6.6 KB,
//   * the h ring grows from 3 to 6 stages (120 KB > one whole 100 KB tile), so the next item's tile streams in while
//     the current one is consumed and the L2 latency of a refill is no longer exposed.
// Protocol differences: the even CTA (leader) issues all MMAs and needs BOTH CTAs' ring stages — warp 1 of the odd CTA
// relays its `full` completions to the leader's `pfull` barriers (remote arrive); tcgen05.commit multicasts to both
// CTAs' `empty` / `acc_full`; both CTAs' epilogue warps arrive on the LEADER's acc_empty (count 24); h_ready counts the
// 8 CTAs of the same parity (they exchange the h of the same sequence tile).  An odd tile count leaves the odd CTAs of
// the last pair without a tile: they feed zeros, skip loads and stores, and keep the barrier protocol.
constexpr int PCL = 16;
constexpr int PBH = LBN / 2;                               // 104 rows of B per CTA
constexpr int PSTAGES = 6;
constexpr uint32_t P_W_BYTES = LKC * PBH * 16;             // 83200
constexpr int P_NBARS = 3 * PSTAGES + LNS + 2 + LNS + 3;
constexpr size_t P_SMEM = P_W_BYTES + PSTAGES * L_A_STAGE + P_NBARS * 8 + 16;
static_assert(P_SMEM <= 232448, "exceeds the 227 KB per-CTA shared memory limit");

__device__ __forceinline__ Group group_of_p(const LstmTcArgs& a, int g) {      // j0 / nact in tile PAIRS
  const int ptiles = (a.seq_tiles + 1) >> 1;
  Group r;
  r.d = g / a.gpd;
  const int gi = g - r.d * a.gpd;
  r.j0 = (int)(((long)gi * ptiles) / a.gpd);
  r.nact = (int)(((long)(gi + 1) * ptiles) / a.gpd) - r.j0;
  return r;
}

template <int Q>
__device__ __forceinline__ void epilogue7_role(const LstmTcArgs& a, uint32_t tmem_base, int T, int quad, int lane, int cid,
                                               int ncl, int e, uint32_t leader, uint64_t* acc_full, uint64_t* acc_empty,
                                               uint64_t* w_free) {
  const int r = quad * 32 + lane;
  const bool last_third = T == 2;
  const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16) + 64 * T;
  const uint32_t t_lane48 = tmem_base + ((uint32_t)(quad * 32) << 16) + 192;
  const int ngroups = 2 * a.gpd;
  const size_t y_tile = (size_t)LKC * 128 * 8;                 // halves per (step, tile, dir) of y
  const size_t g_tile = (size_t)2 * LCL * LGC * 128 * 8;       // halves per (step, tile) of gates_x
  const int pf_idx = (T * 4 + quad) * 32 + lane;               // 0..383: share of the next step's L2 prefetch
  uint32_t it0 = 0, nfull = 0;
  float c0[17], c1[17], c2[17];
  for (int g = cid; g < ngroups; g += ncl) {
    const Group G = group_of_p(a, g);
#pragma unroll
    for (int i = 0; i < 17; ++i) c0[i] = c1[i] = c2[i] = 0.f;
    for (int s = 0; s < a.steps; ++s) {
      const int p = G.d == 0 ? s : a.steps - 1 - s;
#pragma unroll
      for (int k = 0; k < LNS; ++k) {
        if (k < G.nact) {
          const int j = 2 * (G.j0 + k) + e;
          const bool valid = j < a.seq_tiles;
          const size_t tile = (size_t)p * a.seq_tiles + (valid ? j : 0);
          const __half* gbase = a.gates_x + tile * g_tile + (size_t)(G.d * LCL + Q) * (LGC * 128 * 8);
          __half* ycore = a.y + (tile * 2 + G.d) * y_tile + (size_t)(6 * Q + 2 * T) * (128 * 8) + (size_t)r * 8;
          // a missing tile (odd tile count) is computed on tile 0's input projection and never stored
          uint4 gg[8];
          uint2 g48 = make_uint2(0u, 0u);
          const uint4* gp = reinterpret_cast<const uint4*>(gbase) + (size_t)(8 * T) * 128 + r;
#pragma unroll
          for (int i = 0; i < 8; ++i) gg[i] = __ldg(gp + i * 128);
          if (last_third) g48 = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint4*>(gbase) + 24 * 128 + r));
          if (s + 1 < a.steps) {                         // next step's input projection (53 KB = 416 lines) -> L2
            const long step_off = (G.d == 0 ? 1 : -1) * (long)a.seq_tiles * (long)g_tile;
            const char* nx = reinterpret_cast<const char*>(gbase + step_off);
            prefetch_l2(nx + pf_idx * 128);
            if (pf_idx < LGC * 16 - 384) prefetch_l2(nx + (384 + pf_idx) * 128);
          }
          const uint32_t it = it0 + (uint32_t)(s * G.nact + k);
          const uint32_t buf = it & 1;
          mbar_wait(acc_full + k, (nfull >> k) & 1);
          nfull ^= 1u << k;
          tc_fence_after();
          if (k == 0) epi5_item<Q>(t_lane + buf * LACC, last_third, t_lane48 + buf * LACC, gg, g48, ycore, c0, valid);
          else if (k == 1) epi5_item<Q>(t_lane + buf * LACC, last_third, t_lane48 + buf * LACC, gg, g48, ycore, c1, valid);
          else epi5_item<Q>(t_lane + buf * LACC, last_third, t_lane48 + buf * LACC, gg, g48, ycore, c2, valid);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(acc_empty + buf, leader);   // the pair's MMA issuer counts both CTAs' warps
          if (s + 1 < a.steps) asm volatile("bar.arrive %0, %1;" ::"r"(1 + k), "n"(384 + 32) : "memory");
        }
      }
    }
    const int gn = g + ncl;                          // next group of this cluster switches direction?
    if (gn < ngroups && gn / a.gpd != G.d) {
      __syncwarp();
      if (lane == 0) mbar_arrive(w_free);            // this warp consumed the last accumulator: W may go
    }
    it0 += (uint32_t)(a.steps * G.nact);
  }
}

__global__ void __launch_bounds__(LTHREADS, 1) lstm_tc2_kernel(const LstmTcArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sW = smem;
  uint8_t* sA = smem + P_W_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + PSTAGES * L_A_STAGE);
  uint64_t* full = bars;                       // [PSTAGES] this CTA's ring stage landed
  uint64_t* empty = full + PSTAGES;            // [PSTAGES] MMAs that read the stage retired (multicast commit)
  uint64_t* pfull = empty + PSTAGES;           // [PSTAGES] leader only: the odd CTA's stage landed (relayed)
  uint64_t* acc_full = pfull + PSTAGES;        // [LNS]
  uint64_t* acc_empty = acc_full + LNS;        // [2]   leader only: 24 epilogue warps of the pair
  uint64_t* h_ready = acc_empty + 2;           // [LNS] 8 arrivals: the CTAs of this parity
  uint64_t* w_full = h_ready + LNS;
  uint64_t* w_free = w_full + 1;
  uint64_t* pw_full = w_free + 1;              // leader only: the odd CTA's W half landed (relayed)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pw_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t q = rank >> 1;                // pair = unit slice
  const int e = (int)(rank & 1u);              // tile parity served by this CTA
  const uint32_t leader = rank & ~1u;
  const int cid = cluster_id_x(), ncl = num_clusters_x();

  if (threadIdx.x == 0) {
    for (int i = 0; i < PSTAGES; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); mbar_init(pfull + i, 1); }
    for (int i = 0; i < LNS; ++i) mbar_init(acc_full + i, 1);
    for (int i = 0; i < 2; ++i) mbar_init(acc_empty + i, 24);
    for (int i = 0; i < LNS; ++i) mbar_init(h_ready + i, PCL / 2);
    mbar_init(w_full, 1);
    mbar_init(w_free, 12);
    mbar_init(pw_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc2(tmem_slot, 2 * LACC);
  tc_fence_before();
  __syncthreads();
  cluster_sync();                       // every CTA's barriers are initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int ngroups = 2 * a.gpd;
  const size_t y_tile = (size_t)LKC * 128 * 8;         // halves per (step, tile, dir)

  if (warp == 0) {
    // ------------------------------------------------------------------ producer: W half + this CTA's h tiles
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, hphase = 0, wfphase = 0;
      int cur_dir = -1;
      long long w_h = 0, w_e = 0, w_o = 0;
      P_DECL(e == 0);
      for (int g = cid; g < ngroups; g += ncl) {
        const Group G = group_of_p(a, g);
        if (G.d != cur_dir) {
          if (cur_dir >= 0) {
            mbar_wait(w_free, wfphase);
            wfphase ^= 1;
          }
          mbar_expect_tx(w_full, P_W_BYTES);
          // rows [104e, 104e+104) of every k-core of the pair's packed slice [50][208][8]
          const uint8_t* src = reinterpret_cast<const uint8_t*>(a.w_pack) + ((size_t)G.d * LCL + q) * L_W_BYTES + (size_t)e * (PBH * 16);
          for (int kc = 0; kc < LKC; ++kc) bulk_g2s(sW + kc * (PBH * 16), src + (size_t)kc * (LBN * 16), PBH * 16, w_full);
          cur_dir = G.d;
        }
        for (int s = 0; s < a.steps; ++s) {
          const int p_prev = G.d == 0 ? s - 1 : a.steps - s;        // position whose h feeds this step
          for (int k = 0; k < G.nact; ++k) {
            const int j = 2 * (G.j0 + k) + e;
            const uint8_t* src = reinterpret_cast<const uint8_t*>(a.zero_tile);
            if (s > 0) {
              P_MARK(w_o);
              mbar_wait_cluster(h_ready + k, (hphase >> k) & 1);
              P_MARK(w_h);
              hphase ^= 1u << k;
              fence_proxy_async_global();
              if (j < a.seq_tiles)
                src = reinterpret_cast<const uint8_t*>(a.y + (((size_t)p_prev * a.seq_tiles + j) * 2 + G.d) * y_tile);
            }
            for (int ks = 0; ks < LNST; ++ks) {
              P_MARK(w_o);
              mbar_wait(empty + stage, phase ^ 1);
              P_MARK(w_e);
              mbar_expect_tx(full + stage, L_A_STAGE);
              bulk_g2s(sA + stage * L_A_STAGE, src + (size_t)ks * L_A_STAGE, L_A_STAGE, full + stage);
              if (++stage == PSTAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
      P_MARK(w_o);
      if (prb_) { a.probe[0] = w_h; a.probe[1] = w_e; a.probe[2] = w_o; }
    }
  } else if (warp == 1) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (lane == 0 && e == 0) {
      // ---------------------------------------------------------------- MMA issuer (leader CTA of the pair)
      const uint32_t idesc = idesc_f16_f32(256, LBN);
      const uint16_t pair_mask = (uint16_t)(3u << rank);
      uint32_t stage = 0, phase = 0, it = 0, wphase = 0;
      int cur_dir = -1;
      const uint32_t sw = smem_u32(sW);
      long long w_a = 0, w_f = 0, w_o = 0;
      P_DECL(e == 0);
      for (int g = cid; g < ngroups; g += ncl) {
        const Group G = group_of_p(a, g);
        if (G.d != cur_dir) {
          mbar_wait(w_full, wphase);
          mbar_wait_cluster(pw_full, wphase);
          wphase ^= 1;
          cur_dir = G.d;
        }
        const int nitems = a.steps * G.nact;
        for (int i = 0; i < nitems; ++i, ++it) {
          const uint32_t buf = it & 1;
          P_MARK(w_o);
          mbar_wait_cluster(acc_empty + buf, ((it >> 1) & 1) ^ 1);
          P_MARK(w_a);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * LACC;
          for (int ks = 0; ks < LNST; ++ks) {
            P_MARK(w_o);
            mbar_wait(full + stage, phase);
            mbar_wait_cluster(pfull + stage, phase);
            P_MARK(w_f);
            tc_fence_after();
            const uint32_t sa = smem_u32(sA + stage * L_A_STAGE);
#pragma unroll
            for (int jk = 0; jk < LKS / 2; ++jk) {
              const uint64_t da = smem_desc_kb8(sa + jk * 2 * 2048, 2048, 128);
              const uint64_t db = smem_desc_kb8(sw + (ks * LKS + jk * 2) * (PBH * 16), PBH * 16, 128);
              mma_f16_ss_2cta(d_tmem, da, db, idesc, (ks | jk) != 0);
            }
            mma_commit2_multicast(empty + stage, pair_mask);
            if (++stage == PSTAGES) { stage = 0; phase ^= 1; }
          }
          mma_commit2_multicast(acc_full + (i % G.nact), pair_mask);
        }
      }
      P_MARK(w_o);
      if (prb_) { a.probe[4] = w_a; a.probe[5] = w_f; a.probe[6] = w_o; }
    } else if (lane == 0) {
      // ---------------------------------------------------------------- relay (odd CTA): my stages -> leader's pfull
      uint32_t stage = 0, phase = 0, wphase = 0;
      int cur_dir = -1;
      for (int g = cid; g < ngroups; g += ncl) {
        const Group G = group_of_p(a, g);
        if (G.d != cur_dir) {
          mbar_wait(w_full, wphase);
          wphase ^= 1;
          mbar_arrive_cluster(pw_full, leader);
          cur_dir = G.d;
        }
        const int nfills = a.steps * G.nact * LNST;
        for (int i = 0; i < nfills; ++i) {
          mbar_wait(full + stage, phase);
          mbar_arrive_cluster(pfull + stage, leader);
          if (++stage == PSTAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ publisher
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    for (int g = cid; g < ngroups; g += ncl) {
      const Group G = group_of_p(a, g);
      for (int s = 0; s + 1 < a.steps; ++s) {
        for (int k = 0; k < G.nact; ++k) {
          asm volatile("bar.sync %0, %1;" ::"r"(1 + k), "n"(384 + 32) : "memory");
          if (lane < PCL / 2) {
            fence_proxy_async_global();
            fence_acq_rel_cluster();
            mbar_arrive_cluster_relaxed(h_ready + k, 2 * lane + e);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 3) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");      // idle 4th warp of warpgroup 0
  } else {
    // ------------------------------------------------------------------ epilogue: 12 warps, every item
    asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
    const int k = (warp - 4) >> 2, quad = warp & 3;
#define BSRNN_EPI7_CASE(QQ) \
  case QQ: epilogue7_role<QQ>(a, tmem_base, k, quad, lane, cid, ncl, e, leader, acc_full, acc_empty, w_free); break;
    switch (q) {
      BSRNN_EPI7_CASE(0) BSRNN_EPI7_CASE(1) BSRNN_EPI7_CASE(2) BSRNN_EPI7_CASE(3)
      BSRNN_EPI7_CASE(4) BSRNN_EPI7_CASE(5) BSRNN_EPI7_CASE(6) BSRNN_EPI7_CASE(7)
    }
#undef BSRNN_EPI7_CASE
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();                       // no CTA exits (or frees TMEM) while peers may still arrive / read
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 2 * LACC);
  }
}

static cudaError_t launch_v7(const LstmTcArgs& a, int ncl, cudaStream_t st, int* occupancy) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e1 = cudaFuncSetAttribute(lstm_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P_SMEM);
    if (e1 != cudaSuccess) return e1;
    e1 = cudaFuncSetAttribute(lstm_tc2_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e1 != cudaSuccess) return e1;
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((occupancy ? 16 : ncl) * PCL);
  cfg.blockDim = dim3(LTHREADS);
  cfg.dynamicSmemBytes = P_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = PCL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  if (occupancy) return cudaOccupancyMaxActiveClusters(occupancy, lstm_tc2_kernel, &cfg);
  return cudaLaunchKernelEx(&cfg, lstm_tc2_kernel, a);
}
static int max_active_clusters_v7() {
  int n = 0;
  LstmTcArgs dummy{};
  if (launch_v7(dummy, 0, nullptr, &n) != cudaSuccess) { cudaGetLastError(); return -1; }
  return n;
}

template <int V5>
static int max_active_clusters_t() {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(LCL * 64);
  cfg.blockDim = dim3(LTHREADS);
  cfg.dynamicSmemBytes = F_SMEM;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = LCL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int n = 0;
  if (cudaFuncSetAttribute(lstm_tc_kernel<V5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F_SMEM) != cudaSuccess) return -1;
  if (cudaOccupancyMaxActiveClusters(&n, lstm_tc_kernel<V5>, &cfg) != cudaSuccess) return -1;
  return n;
}
static int max_active_clusters() {
  const int a6 = max_active_clusters_t<6>(), a5 = max_active_clusters_t<5>(), a4 = max_active_clusters_t<4>();
  const int a8 = max_active_clusters_t<8>() < max_active_clusters_t<9>() ? max_active_clusters_t<8>() : max_active_clusters_t<9>();
  const int m = (a5 < a4 ? a5 : a4) < a8 ? (a5 < a4 ? a5 : a4) : a8;
  return a6 < m ? a6 : m;
}

}  // namespace bsrnn
using namespace bsrnn;

static long long* g_lstm_probe = nullptr;
static int g_lstm_probe_cid = 0;
extern "C" void bsrnn_debug_set_lstm_probe(void* p, int cid) {
  g_lstm_probe = reinterpret_cast<long long*>(p);
  g_lstm_probe_cid = cid;
}

// Recurrence schedule: 4 / 5 / 6 / 8 (8-CTA clusters, see lstm_tc_kernel) or 7 (CTA pairs, lstm_tc2_kernel); < 0 = take
// BSRNN_LSTM_VER from the environment at the next call (default 8).
static int g_lstm_ver = -1;
extern "C" void bsrnn_debug_set_lstm_schedule(int ver) { g_lstm_ver = (ver >= 4 && ver <= 9) ? ver : -1; }

// slots: sequence tiles a cluster interleaves (1..3; <= 0 = 3).
extern "C" int bsrnn_blstm_recurrence_tc_ex(const void* gates_x, const void* w_pack, const void* zero_tile, void* y, int R,
                                            int steps, int seq_tiles, int max_clusters, int slots, void* stream) {
  BSRNN_CHECK_ARG(gates_x && w_pack && zero_tile && y, "blstm_recurrence_tc: null pointer");
  BSRNN_CHECK_ARG(R > 0 && steps > 0 && (long)seq_tiles * 128 >= R, "blstm_recurrence_tc: bad dims");
  if (slots <= 0 || slots > LNS) slots = LNS;
  if (slots > seq_tiles) slots = seq_tiles;
  LstmTcArgs a{reinterpret_cast<const __half*>(gates_x), reinterpret_cast<const __half*>(w_pack),
               reinterpret_cast<const __half*>(zero_tile), reinterpret_cast<__half*>(y), R, steps, seq_tiles,
               (seq_tiles + slots - 1) / slots, g_lstm_probe, g_lstm_probe_cid, nullptr};
  static int max_active = -1;
  if (max_active < 0) {
    const int n = max_active_clusters();
    if (n <= 0) {
      cudaGetLastError();
      set_error("blstm_recurrence_tc: no co-resident 8-CTA cluster (512 threads, %zu B shared memory)", F_SMEM);
      return 2;
    }
    max_active = n;
  }
  if (g_lstm_ver < 0) { const char* e = getenv("BSRNN_LSTM_VER"); g_lstm_ver = (e && e[0] >= '4' && e[0] <= '9') ? e[0] - '0' : 8; }
  const int ver = g_lstm_ver;
  if (ver == 7) {
    // CTA-pair schedule: units are (direction, PAIR of sequence tiles); 16-CTA clusters
    static int max7 = -2;
    if (max7 == -2) max7 = max_active_clusters_v7();
    if (max7 <= 0) {
      set_error("blstm_recurrence_tc: no co-resident 16-CTA cluster (512 threads, %zu B shared memory)", P_SMEM);
      return 2;
    }
    const int ptiles = (seq_tiles + 1) / 2;
    int sl = slots > ptiles ? ptiles : slots;
    a.gpd = (ptiles + sl - 1) / sl;
    int want = max7;                                 // spread over all co-resident clusters when there are few units
    if (max_clusters > 0 && want > max_clusters) want = max_clusters;
    if (2 * a.gpd < want) { a.gpd = want / 2 < ptiles ? want / 2 : ptiles; if (a.gpd < 1) a.gpd = 1; }
    int ncl7 = 2 * a.gpd;
    if (ncl7 > want) ncl7 = want;
    cudaError_t le = launch_v7(a, ncl7, (cudaStream_t)stream, nullptr);
    if (le != cudaSuccess) { set_error("blstm_recurrence_tc (pair schedule): %s", cudaGetErrorString(le)); return 3; }
    BSRNN_LAUNCH_OK();
    return 0;
  }
  int ncl = 2 * a.gpd;
  if (ncl > max_active) ncl = max_active;
  if (max_clusters > 0 && ncl > max_clusters) ncl = max_clusters;
  // BSRNN_LSTM_VER=4|5|6|8 selects the schedule for A/B timing.  Default 8 (profiles/r01/call30, BASELINE config 2):
  // v5 7.49 / 6.51 ms (time / band axis), v8 7.16 / 6.31 ms.  The multicast variant (6) gained nothing over 5
  // (call22): the h ring (60 KB in flight / L2 latency) and the MUFU-heavy epilogue pace the kernel, not L2 traffic.
  // 9 = 8 + multicast h loads: 7.08 / 5.99 ms against 6.84 / 6.02 ms for 8 in the same run (call59): still no gain.
  if (ver == 4) lstm_tc_kernel<4><<<ncl * LCL, LTHREADS, F_SMEM, (cudaStream_t)stream>>>(a);
  else if (ver == 5) lstm_tc_kernel<5><<<ncl * LCL, LTHREADS, F_SMEM, (cudaStream_t)stream>>>(a);
  else if (ver == 8) lstm_tc_kernel<8><<<ncl * LCL, LTHREADS, F_SMEM, (cudaStream_t)stream>>>(a);
  else if (ver == 9) lstm_tc_kernel<9><<<ncl * LCL, LTHREADS, F_SMEM, (cudaStream_t)stream>>>(a);
  else lstm_tc_kernel<6><<<ncl * LCL, LTHREADS, F_SMEM, (cudaStream_t)stream>>>(a);
  BSRNN_LAUNCH_OK();
  return 0;
}

// Flag-group schedule (lstm_tc_flag_kernel): up to floor(SMs / 8) = 18 groups of 8 CTAs on a B200, each group owning
// `slots` interleaved sequence tiles.  sync_ws: BSRNN_LSTM_SYNC_BYTES of device memory owned by the caller (zeroed
// here, stream-ordered, before every launch).  slots <= 0: as few interleaved tiles as still cover all units with the
// co-resident groups (2 for the time axis of BASELINE config 2: 34 units on 18 groups; 3 when units abound).
static int flag_max_groups() {
  static int cached = -1;
  if (cached >= 0) return cached;
  int dev = 0, sms = 0, occ = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  if (cudaFuncSetAttribute(lstm_tc_flag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F_SMEM) != cudaSuccess) return -1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lstm_tc_flag_kernel, LTHREADS, F_SMEM) != cudaSuccess) return -1;
  cached = (sms * occ) / LCL;
  if (cached > 32) cached = 32;                       // L_SYNC_WORDS holds 32 groups
  return cached;
}

extern "C" int bsrnn_blstm_recurrence_tc_flag(const void* gates_x, const void* w_pack, const void* zero_tile, void* y, int R,
                                              int steps, int seq_tiles, int max_groups, int slots, void* sync_ws,
                                              void* stream) {
  BSRNN_CHECK_ARG(gates_x && w_pack && zero_tile && y && sync_ws, "blstm_recurrence_tc_flag: null pointer");
  BSRNN_CHECK_ARG(R > 0 && steps > 0 && (long)seq_tiles * 128 >= R, "blstm_recurrence_tc_flag: bad dims");
  int cap = flag_max_groups();
  if (cap <= 0) {
    cudaGetLastError();
    set_error("blstm_recurrence_tc_flag: kernel does not fit (512 threads, %zu B shared memory)", F_SMEM);
    return 2;
  }
  if (max_groups > 0 && cap > max_groups) cap = max_groups;
  if (slots <= 0) {
    slots = (2 * seq_tiles + cap - 1) / cap;
    if (slots < 1) slots = 1;
  }
  if (slots > LNS) slots = LNS;
  if (slots > seq_tiles) slots = seq_tiles;
  LstmTcArgs a{reinterpret_cast<const __half*>(gates_x), reinterpret_cast<const __half*>(w_pack),
               reinterpret_cast<const __half*>(zero_tile), reinterpret_cast<__half*>(y), R, steps, seq_tiles,
               (seq_tiles + slots - 1) / slots, g_lstm_probe, g_lstm_probe_cid, reinterpret_cast<unsigned*>(sync_ws)};
  int ncl = 2 * a.gpd;
  if (ncl > cap) ncl = cap;
  cudaStream_t st = (cudaStream_t)stream;
  BSRNN_CUDA_OK(cudaMemsetAsync(sync_ws, 0, (size_t)L_SYNC_WORDS * sizeof(unsigned), st));
  lstm_tc_flag_kernel<<<ncl * LCL, LTHREADS, F_SMEM, st>>>(a);
  BSRNN_LAUNCH_OK();
  return 0;
}
extern "C" int bsrnn_blstm_tc_flag_max_groups(void) { return flag_max_groups(); }
extern "C" int bsrnn_blstm_tc_sync_bytes(void) { return (int)(L_SYNC_WORDS * sizeof(unsigned)); }

extern "C" int bsrnn_blstm_recurrence_tc(const void* gates_x, const void* w_pack, const void* zero_tile, void* y, int R,
                                         int steps, int seq_tiles, int max_clusters, void* stream) {
  return bsrnn_blstm_recurrence_tc_ex(gates_x, w_pack, zero_tile, y, R, steps, seq_tiles, max_clusters, 0, stream);
}

extern "C" int bsrnn_blstm_tc_max_clusters(void) { return max_active_clusters(); }
extern "C" int bsrnn_blstm_tc_max_pair_clusters(void) { return max_active_clusters_v7(); }
