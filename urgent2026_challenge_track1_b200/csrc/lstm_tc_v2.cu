// lstm_tc.cu — persistent BLSTM recurrence on tcgen05 tensor cores for H = 392 (BSRNN_baseline, N = 196).
//
// Replaces the recurrent half of nn.LSTM(N, 2N, bidirectional) [reference bsrnn_flowse.py:226-238, called at
// :296-297 (time axis: 2176 sequences x 1001 steps at BASELINE config 2) and :303-304 (band axis: 64064 x 34)].
//
// Decomposition.  A work unit is (direction d, tile j of 128 sequences).  A thread-block CLUSTER of 8 CTAs owns a
// unit for all of its steps; CTA q of the cluster owns hidden units [49q, 49q+49) i.e. 196 of the 1568 gate
// columns, and keeps that slice of W_hh (208 x 400 fp16, 166 KB, UMMA KB8 layout) resident in shared memory for the
// whole launch.  Per step every CTA
//   (1) bulk-copies the full h_{t-1} tile (128 x 400 fp16 = 100 KB) from the y buffer in L2 through a 3-stage ring,
//   (2) issues 25 tcgen05.mma (M=128, N=208, K=16) accumulating the recurrent pre-activations in TMEM,
//   (3) epilogue warps read the accumulator (tcgen05.ld), add the precomputed input projection (fp16, from HBM),
//       apply the gates with MUFU tanh, update c (kept in REGISTERS for the whole sequence: 49 f32 per thread) and
//       write their 49-unit slice of h_t into y (which is at once the layer output consumed by the Linear GEMM and
//       the exchange buffer for the other 7 CTAs),
//   (4) release-arrive on the h_ready mbarrier of all 8 CTAs (cluster scope); the producers acquire it before
//       fetching h_t for the next step.
// No grid-wide synchronisation exists: clusters are independent, rows never mix.
//
// y layout (fp16): [step][seq_tile][dir][50 k-cores][128 rows][8]   (k-core 49 = zero padding, K = 400 per dir);
//                  a (step, seq_tile) block is therefore a 128 x 800 KB8 operand tile for the Linear(4N->N) GEMM.
// gates_x (fp16) : [token][dir][q][208]  column c = 4*u_local + gate (i,f,g,o), 196 real + 12 pad
// w_hh pack      : [dir][q][50 k-cores][208][8]
#include "common.cuh"
#include "umma.cuh"
#include <cuda_fp16.h>

namespace bsrnn { namespace v2 {
using namespace umma;

constexpr int LH = 392;            // hidden size
constexpr int LCL = 8;             // cluster size
constexpr int LU = LH / LCL;       // 49 hidden units per CTA
constexpr int LBN = 208;           // gate columns per CTA (4*49 = 196, padded to a multiple of 16)
constexpr int LKC = 50;            // k-cores of the recurrent operand (K = 400)
constexpr int LKS = 10;            // k-cores per A stage
constexpr int LNST = LKC / LKS;    // 5 stages per step
constexpr int LSTAGES = 3;
constexpr int LTHREADS = 320;        // producer warp + MMA warp + 8 epilogue warps
constexpr uint32_t L_W_BYTES = LKC * LBN * 16;          // 166400
constexpr uint32_t L_A_STAGE = LKS * 128 * 16;          // 20480
constexpr size_t L_SMEM = L_W_BYTES + LSTAGES * L_A_STAGE + 16 * 8 + 16;
static_assert(L_SMEM <= 232448, "exceeds the 227 KB per-CTA shared memory limit");

struct LstmTcArgs {
  const __half* gates_x;
  const __half* w_pack;
  __half* y;
  int R, steps, seq_tiles;
  long seq_inner, seq_outer, seq_inner_stride, step_stride;
  long long* trace;      // optional (debug): [step][8] SM-clock stamps written by cluster 0 / CTA 0
};

#define LSTM_TRACE(slot, step)                                                         \
  do {                                                                                 \
    if (a.trace && cid == 0 && q == 0 && (step) < 64) a.trace[(step) * 8 + (slot)] = clock64(); \
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

// Epilogue work split: 8 warps; warp (quadrant, half) owns 32 rows x units [U0, U0+NU) with
//   half 0: units [0,24)  = accumulator columns [0,96)    (3 chunks of 32 columns)
//   half 1: units [24,49) = accumulator columns [96,196)  (3 chunks of 32 + one of 4)
// Unit i of CTA Q is h column k = 49Q + i -> k-core 6Q + (Q+i)/8, slot (Q+i)%8 of the y tile.
template <int Q, int U0, int NU>
__device__ __forceinline__ void store_h(__half* ytile_row /* &y[...][kc=0][r][0] */, const float (&h)[25]) {
#pragma unroll
  for (int jj = 0; jj < 7; ++jj) {
    const int lo = 8 * jj - Q;                       // unit index sitting in slot 0 of this core
    if (lo + 8 <= U0 || lo >= U0 + NU) continue;     // core holds none of our units
    __half* dst = ytile_row + (size_t)(6 * Q + jj) * 128 * 8;
    if (lo >= U0 && lo + 8 <= U0 + NU) {
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
        if (i >= U0 && i < U0 + NU) dst[sl] = __float2half_rn(h[i - U0]);
      }
    }
  }
}

struct GxRegs {            // this thread's slice of the precomputed input projection for one step (fp16)
  uint4 v[12];
  uint2 tail;
};

template <int HALF>
__device__ __forceinline__ void load_gx(const __half* gx, bool row_ok, GxRegs& g) {
  if (row_ok) {
    const uint4* p = reinterpret_cast<const uint4*>(gx + HALF * 96);
#pragma unroll
    for (int i = 0; i < 12; ++i) g.v[i] = __ldg(p + i);
    if (HALF == 1) g.tail = __ldg(reinterpret_cast<const uint2*>(gx + 192));
  } else {
#pragma unroll
    for (int i = 0; i < 12; ++i) g.v[i] = make_uint4(0, 0, 0, 0);
    g.tail = make_uint2(0, 0);
  }
}

template <int Q, int HALF>
__device__ __forceinline__ void epilogue_step(uint32_t t_addr, bool have_acc, const GxRegs& g, __half* ytile_row,
                                              float (&c)[25]) {
  constexpr int U0 = HALF == 0 ? 0 : 24;
  constexpr int NU = HALF == 0 ? 24 : 25;
  float h[25];
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    uint32_t acc[32];
    if (have_acc) {
      tmem_ld_x32(t_addr + HALF * 96 + ch * 32, acc);
      tmem_ld_wait();
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] = 0u;
    }
    const __half2* gh = reinterpret_cast<const __half2*>(&g.v[ch * 4]);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float2 g01 = __half22float2(gh[2 * u]), g23 = __half22float2(gh[2 * u + 1]);
      gate_update(__uint_as_float(acc[4 * u]) + g01.x, __uint_as_float(acc[4 * u + 1]) + g01.y,
                  __uint_as_float(acc[4 * u + 2]) + g23.x, __uint_as_float(acc[4 * u + 3]) + g23.y, c[ch * 8 + u],
                  h[ch * 8 + u]);
    }
  }
  if (HALF == 1) {
    uint32_t acc[4];
    if (have_acc) {
      tmem_ld_x4(t_addr + 192, acc);
      tmem_ld_wait();
    } else {
      acc[0] = acc[1] = acc[2] = acc[3] = 0u;
    }
    const __half2* gh = reinterpret_cast<const __half2*>(&g.tail);
    const float2 g01 = __half22float2(gh[0]), g23 = __half22float2(gh[1]);
    gate_update(__uint_as_float(acc[0]) + g01.x, __uint_as_float(acc[1]) + g01.y, __uint_as_float(acc[2]) + g23.x,
                __uint_as_float(acc[3]) + g23.y, c[24], h[24]);
  }
  store_h<Q, U0, NU>(ytile_row, h);
}

// The epilogue role for one (Q, HALF): loops over this cluster's work units and steps.
template <int Q, int HALF>
__device__ __forceinline__ void epilogue_role(const LstmTcArgs& a, uint32_t tmem_base, int warp, int lane, int cid, int ncl,
                                              uint64_t* acc_full, uint64_t* acc_empty, uint64_t* h_ready, uint64_t* w_free) {
  const int quad = warp & 3;
  const int q = Q;
  const bool tracer = warp == 2 && lane == 0;
  const int r = quad * 32 + lane;
  const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
  const int units = 2 * a.seq_tiles;
  const size_t tile_elems = (size_t)LKC * 128 * 8;
  uint32_t fphase = 0;
  for (int w = cid; w < units; w += ncl) {
    const int d = w / a.seq_tiles, j = w - d * a.seq_tiles;
    const long seq = (long)j * 128 + r;
    const bool row_ok = seq < a.R;
    const long tok0 = row_ok ? (seq / a.seq_inner) * a.seq_outer + (seq % a.seq_inner) * a.seq_inner_stride : 0;
    float c[25];
#pragma unroll
    for (int i = 0; i < 25; ++i) c[i] = 0.f;
    for (int s = 0; s < a.steps; ++s) {
      const int p = d == 0 ? s : a.steps - 1 - s;
      const long token = tok0 + (long)p * a.step_stride;
      const __half* gx = a.gates_x + (token * 2 + d) * (LCL * LBN) + Q * LBN;
      __half* ytile_row = a.y + (((size_t)p * a.seq_tiles + j) * 2 + d) * tile_elems + (size_t)r * 8;
      GxRegs g;
      load_gx<HALF>(gx, row_ok, g);                 // issued before the wait: HBM latency hides behind the MMAs
      const bool have_acc = s > 0;
      if (have_acc) {
        mbar_wait(acc_full, fphase);
        fphase ^= 1;
        tc_fence_after();
        if (tracer) LSTM_TRACE(5, s);
      }
      epilogue_step<Q, HALF>(t_addr, have_acc, g, ytile_row, c);
      if (tracer) LSTM_TRACE(6, s);
      if (have_acc) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty);
      }
      if (s + 1 == a.steps) {
        const int wn = w + ncl;                     // next unit of this cluster switches direction?
        if (wn < units && wn / a.seq_tiles != d) {
          __syncwarp();
          if (lane == 0) mbar_arrive(w_free);
        }
      } else {
        // publish h_t: all 128 rows x 49 units are written once every epilogue thread passed the named barrier;
        // one fence (cumulative over the CTA's stores) then 8 relaxed remote arrives issued by 8 lanes in parallel.
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (warp == 2 && lane < LCL) {
          fence_proxy_async_global();
          fence_acq_rel_cluster();
          mbar_arrive_cluster_relaxed(h_ready, lane);
          if (lane == 0) LSTM_TRACE(7, s);
        }
      }
    }
  }
}

__global__ void __cluster_dims__(LCL, 1, 1) __launch_bounds__(LTHREADS, 1) lstm_tc_v2_kernel(const LstmTcArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sW = smem;
  uint8_t* sA = smem + L_W_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + LSTAGES * L_A_STAGE);
  uint64_t* full = bars;               // [3]
  uint64_t* empty = bars + 3;          // [3]
  uint64_t* acc_full = bars + 6;
  uint64_t* acc_empty = bars + 7;
  uint64_t* h_ready = bars + 8;
  uint64_t* w_full = bars + 9;
  uint64_t* w_free = bars + 10;        // epilogue -> producer: the resident W slice may be overwritten
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t q = cluster_ctarank();
  const int cid = cluster_id_x(), ncl = num_clusters_x();

  if (threadIdx.x == 0) {
    for (int i = 0; i < LSTAGES; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 8);
    mbar_init(h_ready, LCL);
    mbar_init(w_full, 1);
    mbar_init(w_free, 8);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  cluster_sync();                       // every CTA's barriers are initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int units = 2 * a.seq_tiles;
  const size_t tile_elems = (size_t)LKC * 128 * 8;     // halves per (step, tile, dir)

  if (warp == 0) {
    // ------------------------------------------------------------------ producer: W slice + h tiles
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, hphase = 0, wfphase = 0;
      int cur_dir = -1;
      for (int w = cid; w < units; w += ncl) {
        const int d = w / a.seq_tiles, j = w - d * a.seq_tiles;
        if (d != cur_dir) {
          // The epilogue warps signal w_free after consuming the last accumulator of the previous direction, i.e.
          // after every MMA that read the old slice has retired.
          if (cur_dir >= 0) {
            mbar_wait(w_free, wfphase);
            wfphase ^= 1;
          }
          mbar_expect_tx(w_full, L_W_BYTES);
          const uint8_t* src = reinterpret_cast<const uint8_t*>(a.w_pack) + ((size_t)d * LCL + q) * L_W_BYTES;
          // bulk copies are limited in size only by the mbarrier tx-count; split to be conservative
          for (uint32_t off = 0; off < L_W_BYTES; off += 33280) bulk_g2s(sW + off, src + off, 33280, w_full);
          cur_dir = d;
        }
        for (int s = 1; s < a.steps; ++s) {
          const int p_prev = d == 0 ? s - 1 : a.steps - s;        // position whose h feeds this step
          mbar_wait_cluster(h_ready, hphase);
          hphase ^= 1;
          LSTM_TRACE(0, s);
          fence_proxy_async_global();
          const uint8_t* src = reinterpret_cast<const uint8_t*>(
              a.y + (((size_t)p_prev * a.seq_tiles + j) * 2 + d) * tile_elems);
          for (int ks = 0; ks < LNST; ++ks) {
            mbar_wait(empty + stage, phase ^ 1);
            mbar_expect_tx(full + stage, L_A_STAGE);
            bulk_g2s(sA + stage * L_A_STAGE, src + (size_t)ks * L_A_STAGE, L_A_STAGE, full + stage);
            if (++stage == LSTAGES) { stage = 0; phase ^= 1; }
          }
          LSTM_TRACE(1, s);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = idesc_f16_f32(128, LBN);
      uint32_t stage = 0, phase = 0, aphase = 0, wphase = 0;
      int cur_dir = -1;
      const uint32_t sw = smem_u32(sW);
      for (int w = cid; w < units; w += ncl) {
        const int d = w / a.seq_tiles;
        if (d != cur_dir) {
          mbar_wait(w_full, wphase);
          wphase ^= 1;
          cur_dir = d;
        }
        for (int s = 1; s < a.steps; ++s) {
          mbar_wait(acc_empty, aphase ^ 1);
          aphase ^= 1;
          tc_fence_after();
          for (int ks = 0; ks < LNST; ++ks) {
            mbar_wait(full + stage, phase);
            if (ks == 0) LSTM_TRACE(2, s);
            if (ks == LNST - 1) LSTM_TRACE(3, s);
            tc_fence_after();
            const uint32_t sa = smem_u32(sA + stage * L_A_STAGE);
#pragma unroll
            for (int jk = 0; jk < LKS / 2; ++jk) {
              const uint64_t da = smem_desc_kb8(sa + jk * 2 * 2048, 2048, 128);
              const uint64_t db = smem_desc_kb8(sw + (ks * LKS + jk * 2) * (LBN * 16), LBN * 16, 128);
              mma_f16_ss(tmem_base, da, db, idesc, (ks | jk) != 0);
            }
            mma_commit(empty + stage);
            if (++stage == LSTAGES) { stage = 0; phase ^= 1; }
          }
          mma_commit(acc_full);
          LSTM_TRACE(4, s);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: 8 warps (see epilogue_role)
    const int half = (warp - 2) >> 2;
#define BSRNN_EPI_CASE(QQ)                                                                                        \
  case QQ:                                                                                                        \
    if (half == 0) epilogue_role<QQ, 0>(a, tmem_base, warp, lane, cid, ncl, acc_full, acc_empty, h_ready, w_free); \
    else epilogue_role<QQ, 1>(a, tmem_base, warp, lane, cid, ncl, acc_full, acc_empty, h_ready, w_free);           \
    break;
    switch (q) {
      BSRNN_EPI_CASE(0) BSRNN_EPI_CASE(1) BSRNN_EPI_CASE(2) BSRNN_EPI_CASE(3)
      BSRNN_EPI_CASE(4) BSRNN_EPI_CASE(5) BSRNN_EPI_CASE(6) BSRNN_EPI_CASE(7)
    }
#undef BSRNN_EPI_CASE
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();                       // no CTA exits while peers may still arrive on its barriers
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

} }  // namespace bsrnn::v2
using namespace bsrnn; using namespace bsrnn::v2;

static long long* g_lstm_trace = nullptr;
extern "C" void bsrnn_debug_set_lstm_trace_v2(void* p) { g_lstm_trace = reinterpret_cast<long long*>(p); }

extern "C" int bsrnn_blstm_recurrence_tc_v2(const void* gates_x, const void* w_pack, void* y, int R, int steps,
                                         int seq_tiles, long seq_inner, long seq_outer, long seq_inner_stride,
                                         long step_stride, int max_clusters, void* stream) {
  BSRNN_CHECK_ARG(gates_x && w_pack && y, "blstm_recurrence_tc: null pointer");
  BSRNN_CHECK_ARG(R > 0 && steps > 0 && seq_tiles * 128 >= R && seq_inner > 0, "blstm_recurrence_tc: bad dims");
  LstmTcArgs a{reinterpret_cast<const __half*>(gates_x), reinterpret_cast<const __half*>(w_pack),
               reinterpret_cast<__half*>(y), R, steps, seq_tiles, seq_inner, seq_outer, seq_inner_stride, step_stride,
               g_lstm_trace};
  cudaStream_t st = (cudaStream_t)stream;
  BSRNN_CUDA_OK(cudaFuncSetAttribute(lstm_tc_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L_SMEM));
  static int max_active = -1;
  if (max_active < 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(LCL * 64);
    cfg.blockDim = dim3(LTHREADS);
    cfg.dynamicSmemBytes = L_SMEM;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = LCL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, lstm_tc_v2_kernel, &cfg) != cudaSuccess || n <= 0) n = 16;
    max_active = n;
  }
  int ncl = 2 * seq_tiles;
  if (ncl > max_active) ncl = max_active;
  if (max_clusters > 0 && ncl > max_clusters) ncl = max_clusters;
  lstm_tc_v2_kernel<<<ncl * LCL, LTHREADS, L_SMEM, st>>>(a);
  BSRNN_LAUNCH_OK();
  return 0;
}

extern "C" int bsrnn_blstm_tc_max_clusters_v2(void) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(LCL * 64);
  cfg.blockDim = dim3(LTHREADS);
  cfg.dynamicSmemBytes = L_SMEM;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = LCL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int n = 0;
  cudaFuncSetAttribute(lstm_tc_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L_SMEM);
  if (cudaOccupancyMaxActiveClusters(&n, lstm_tc_v2_kernel, &cfg) != cudaSuccess) return -1;
  return n;
}
