// logreg_tc.cu -- fused batched logistic loss + gradient on the 5th-gen tensor cores (sm_100a).
//
// One launch evaluates f, g of every active (candidate x fold) column, i.e. replaces, for all
// columns at once, the two fp32 sgemv passes + pointwise loop that each reference task runs per
// L-BFGS evaluation (SK/linear_model/_linear_loss.py:291-379; ref search.py:230):
//
//     Z  = W  X^T     (slots x rows)   GEMM1   tcgen05.mma, accumulators in TMEM
//     G  = dloss(Z, y) masked by fold  epilogue (CUDA cores): tcgen05.ld -> math -> tcgen05.st
//     dW = G  X       (slots x d)      GEMM2   tcgen05.mma, A operand = G straight from TMEM
//
// X is read ONCE per tile by TMA into 128B-swizzled shared memory and used by both GEMMs: as
// the K-major B operand of GEMM1 (K = features) and, through a second descriptor over the same
// bytes, as the MN-major B operand of GEMM2 (K = rows).
//
// fp32 fidelity on tensor cores (there is no fp32 MMA): every operand is split into two fp16
// numbers, v = hi + lo (22+ mantissa bits after exact power-of-two pre-scaling per feature /
// per column), and each product is three MMAs hi*hi + hi*lo + lo*hi accumulated in fp32.
//
// Work decomposition: a group = 128 slots (MMA M, the TMEM lanes); its rows are cut into P parts;
// one persistent CTA per (group, part) item.  Per CTA: W_hi stationary in shared memory, W_lo
// stationary in TMEM, gradient accumulators (128 x d fp32) stationary in TMEM, X streamed in
// tiles of 64 rows through a ring of 5 half-tile (hi or lo) slots.
//   warp 0      : TMA producer        warp 1 : MMA issuer (+ TMEM allocation)
//   warps 2..17 : epilogue, four warps per TMEM lane quadrant (lane = slot), 16 rows of the tile each
// TMEM columns : [0,256) grad accumulator | [256,384) W_lo (packed fp16 pairs) |
//                [384,448) Z/G buffer 0 | [448,512) Z/G buffer 1
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "skd_internal.h"
#include "tc_ptx.h"

namespace skd {

constexpr int TC_BC = 128;       // slots per group
constexpr int TC_R = 64;         // rows per tile
constexpr int TC_NS = 5;         // ring slots (half tiles)
constexpr int TC_NCH = 147;      // fixed row chunks per group (one partial sum per (chunk, slot)); 147 = 3 * 7 * 7
                                 // divides evenly over 7 / 21 / 49 CTAs per group (20, 7 or 3 groups per GPU)
constexpr int TC_EPI_WARPS = 16;    // four per TMEM lane quadrant
constexpr int TC_EPI_THREADS = TC_EPI_WARPS * 32;
constexpr int TC_THREADS = 64 + TC_EPI_THREADS;
constexpr uint32_t TM_GRAD = 0, TM_WLO = 256, TM_Z0 = 384;
constexpr float XSCALE_TARGET_EXP = 13.f;   // column max scaled into [2^13, 2^14)
constexpr float GSCALE = 16384.f;           // 2^14

__device__ __forceinline__ void epi_bar_sync() {   // named barrier 1: the epilogue warps only
  asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI_THREADS) : "memory");
}

// ---------------------------------------------------------------------------------------------
// data preparation kernels
// ---------------------------------------------------------------------------------------------
// per-feature max |x|
__global__ void tc_colmax_kernel(const float* __restrict__ X, int64_t n, int ldx, int d,
                                 unsigned int* __restrict__ colmax_bits) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= d) return;
  int64_t r0 = (int64_t)blockIdx.y * 4096;
  int64_t r1 = r0 + 4096 < n ? r0 + 4096 : n;
  float m = 0.f;
  for (int64_t r = r0; r < r1; ++r) m = fmaxf(m, fabsf(X[r * ldx + k]));
  atomicMax(&colmax_bits[k], __float_as_uint(m));  // non-negative floats order as uints
}

// scale[k] = 2^(13 - floor(log2(colmax))) ; gscale[k] = 1 / (scale[k] * 2^14)
__global__ void tc_scale_kernel(const unsigned int* __restrict__ colmax_bits, int d, int dpad,
                                float* __restrict__ xscale, double* __restrict__ gscale) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= dpad) return;
  float s = 1.f;
  if (k < d) {
    float m = __uint_as_float(colmax_bits[k]);
    if (m > 0.f && isfinite(m)) {
      int e;
      frexpf(m, &e);  // m = f * 2^e, f in [0.5, 1)  -> floor(log2 m) = e - 1
      s = ldexpf(1.f, (int)XSCALE_TARGET_EXP - (e - 1));
    }
  }
  xscale[k] = s;
  gscale[k] = 1.0 / ((double)s * (double)GSCALE);
}

// Xh, Xl [npad x dpad] fp16 (zero padded), rowmeta[npad] = (fold << 24) | class id (0xFF fold: pad)
__global__ void tc_split_kernel(const float* __restrict__ X, int64_t n, int64_t npad, int ldx, int d,
                                int dpad, const float* __restrict__ xscale,
                                __half* __restrict__ Xh, __half* __restrict__ Xl) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = npad * dpad;
  if (idx >= total) return;
  int64_t r = idx / dpad;
  int k = (int)(idx - r * dpad);
  float v = 0.f;
  if (r < n && k < d) v = X[r * ldx + k] * xscale[k];
  __half h = __float2half_rn(v);
  __half l = __float2half_rn(v - __half2float(h));
  Xh[idx] = h;
  Xl[idx] = l;
}

__global__ void tc_rowmeta_kernel(const int32_t* __restrict__ ycls, const int8_t* __restrict__ fold,
                                  int64_t n, int64_t npad, uint32_t* __restrict__ rowmeta) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= npad) return;
  uint32_t m = 0xFF000000u;
  if (r < n) {
    uint32_t f = fold ? (uint32_t)(uint8_t)fold[r] : 0u;
    m = (f << 24) | (ycls ? ((uint32_t)ycls[r] & 0x00FFFFFFu) : 0u);
  }
  rowmeta[r] = m;
}

// rowsg[li][r] = -y * 2^14 for the rows list li trains on (li < n_lists - 1: every row outside
// fold li; last list: every row), 0 for held-out and padding rows.  y = +1 iff class id == pos.
__global__ void tc_rowsg_kernel(const int32_t* __restrict__ ycls, const int8_t* __restrict__ fold,
                                int64_t n, int64_t npad, int n_lists, int pos, float* __restrict__ rowsg) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= npad) return;
  const bool in = r < n;
  const float s = in ? (ycls[r] == pos ? -GSCALE : GSCALE) : 0.f;
  const int f = (in && fold) ? (int)(uint8_t)fold[r] : -1;
  for (int li = 0; li < n_lists; ++li)
    rowsg[(size_t)li * npad + r] = (li < n_lists - 1 && f == li) ? 0.f : s;
}

// Export of the active slots' iterates in tensor-core form (replaces lb_export_kernel):
// w32 = (float)x (SK/_linear_loss.py:216), W' = w32 / xscale * t, t = 2^(13 - floor(log2 max|w32/xscale|)),
// Wh/Wl [slots_pad x dpad] fp16, wmeta[slot] = {1/t, bias, fold, pos}
struct TcSlotParam {
  float inv_t;
  float bias;
  int32_t fold;
  int32_t pos;
  int32_t neg1;   // SlotMeta::pad (0 = one-vs-rest, k + 1 = pair with class k)
  int32_t col;    // column id of the slot in the caller's batch (-1: padding slot); indexes the row bit matrices
};

__global__ void __launch_bounds__(128)
tc_export_kernel(const double* __restrict__ vec, size_t vec_stride, const SlotMeta* __restrict__ slot,
                 const int32_t* __restrict__ n_act, int d, int dpad, const float* __restrict__ xscale,
                 __half* __restrict__ Wh, __half* __restrict__ Wl, TcSlotParam* __restrict__ sp,
                 const double* __restrict__ xin /* optional: explicit points [slots x (d+1)] */,
                 int fit_intercept) {
  __shared__ float red[4];
  const int s = blockIdx.x;
  if (s >= *n_act) return;
  const SlotMeta sm = slot[s];
  if (sm.col < 0) {   // padding slot of the fold-grouped layout: zero weights, keeps its segment's fold
    for (int k = threadIdx.x; k < dpad; k += 128) {
      Wh[(size_t)s * dpad + k] = __float2half_rn(0.f);
      Wl[(size_t)s * dpad + k] = __float2half_rn(0.f);
    }
    if (threadIdx.x == 0) { TcSlotParam p; p.inv_t = 1.f; p.bias = 0.f; p.fold = sm.fold; p.pos = -1; p.neg1 = 0; p.col = -1; sp[s] = p; }
    return;
  }
  const double* x = xin ? xin + (size_t)s * (d + 1) : vec + (size_t)sm.col * vec_stride;
  float m = 0.f;
  for (int k = threadIdx.x; k < d; k += 128) m = fmaxf(m, fabsf((float)x[k] / xscale[k]));
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  float t = 1.f;
  if (m > 0.f && isfinite(m)) {
    int e;
    frexpf(m, &e);
    t = ldexpf(1.f, (int)XSCALE_TARGET_EXP - (e - 1));
  }
  for (int k = threadIdx.x; k < dpad; k += 128) {
    float v = 0.f;
    if (k < d) v = ((float)x[k] / xscale[k]) * t;   // power-of-two scalings: exact
    __half h = __float2half_rn(v);
    __half l = __float2half_rn(v - __half2float(h));
    Wh[(size_t)s * dpad + k] = h;
    Wl[(size_t)s * dpad + k] = l;
  }
  if (threadIdx.x == 0) {
    TcSlotParam p;
    p.inv_t = 1.f / t;
    p.bias = fit_intercept ? (float)x[d] : 0.f;
    p.fold = sm.fold;
    p.pos = sm.pos;
    p.neg1 = sm.pad;
    p.col = sm.col;
    sp[s] = p;
  }
}

// ---------------------------------------------------------------------------------------------
// the fused kernel
// ---------------------------------------------------------------------------------------------
struct TcParams {
  const __half* Wl;          // [slots_pad x dpad]
  const TcSlotParam* sp;     // [slots]
  const uint32_t* rowmeta;   // [npad]
  const float* yreal;        // [npad] regression targets (TC_R2)
  double* lossp;             // [nz x n_act]        (fit)
  double* gsump;             // [nz x n_act]        (fit)
  float* gradp;              // [nz x n_act x ldw]  (fit)
  unsigned long long* correct;  // [n_act]          (score)
  unsigned long long* count;    // [n_act]          (score)
  int n_act;                 // host value: stride of the partial arrays, upper bound of the live slots
  const int32_t* n_act_dev;  // live slot count on the device (nullptr: n_act is exact)
  int groups;
  int n_tiles;               // npad / 64
  int ldw;                   // leading dimension of gradp (== dpad)
  const int32_t* tilelist;   // per-fold lists of tiles with training rows (nullptr: every tile)
  const int32_t* tilecnt;
  int n_lists, n_tiles_ld;
  const float* rowsg;        // TC_FIT_UNI: [n_lists x npad] per-row -y * 2^14 (0 = not a training row of that list)
  long long rowsg_ld;
  const uint32_t* ybits;     // TC_FIT: per-column row label bits (nullptr: class id == pos), see LogregWork
  const uint32_t* mbits;     // TC_FIT: per-column training-row bits (nullptr: every row of the training folds)
  long long rb_words;
  int g_passes;              // MMA passes of the gradient product: 3 = G_hi X_hi + G_lo X_hi + G_hi X_lo, 2 = without G_lo X_hi
  int debug;                 // SKDIST_B200_TC_DEBUG (timing experiments only): 2 = no GEMM2, 3 = no MMAs, 4 = no epilogue work
};

struct __align__(8) TcBarriers {
  uint64_t full[TC_NS];
  uint64_t empty[TC_NS];
  uint64_t z_full[2];
  uint64_t g_full[2];
  uint64_t w_full;
  uint64_t w_free;      // committed by the MMA warp when it leaves a group: W_hi may be overwritten
  uint64_t wl_full;
  uint64_t acc_done;
  uint64_t acc_free;
  uint32_t tmem_base;
  uint32_t pad;
};

// TC_FIT_UNI: fit where every slot of a group shares (held-out fold, positive class): the row's
// sign/mask comes from a precomputed per-fold array instead of being decoded per element
enum { TC_FIT = 0, TC_SCORE = 1, TC_R2 = 2, TC_FIT_UNI = 3 };

template <int NCHUNK, int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_eval_kernel(const __grid_constant__ CUtensorMap map_xh, const __grid_constant__ CUtensorMap map_xl,
               const __grid_constant__ CUtensorMap map_wh, const TcParams prm) {
  constexpr bool IS_FIT = MODE == TC_FIT || MODE == TC_FIT_UNI;
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // carve shared memory (1024-byte aligned for the 128B swizzle atoms)
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr uint32_t WH_CHUNK = TC_BC * 128;     // [128 slots x 64 fp16]
  constexpr uint32_t X_CHUNK = TC_R * 128;       // [64 rows x 64 fp16]
  constexpr uint32_t SLOT_BYTES = NCHUNK * X_CHUNK;
  uint8_t* s_wh = base;
  uint8_t* s_ring = s_wh + NCHUNK * WH_CHUNK;
  TcBarriers* bars = reinterpret_cast<TcBarriers*>(s_ring + TC_NS * SLOT_BYTES);

  if (threadIdx.x == 0) {
    for (int i = 0; i < TC_NS; ++i) { mbar_init(&bars->full[i], 1); mbar_init(&bars->empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bars->z_full[i], 1); mbar_init(&bars->g_full[i], TC_EPI_THREADS); }
    mbar_init(&bars->w_full, 1);
    mbar_init(&bars->w_free, 1);
    mbar_init(&bars->wl_full, TC_EPI_THREADS);
    mbar_init(&bars->acc_done, 1);
    mbar_init(&bars->acc_free, TC_EPI_THREADS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&bars->tmem_base)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  // Work split.  A group's tile list (all tiles, or the tiles with training rows of the group's
  // fold) is cut into TC_NCH fixed chunks; a unit = (group, chunk); the groups * TC_NCH units are
  // dealt to the CTAs in contiguous ranges.  A chunk is always accumulated by ONE CTA from a zeroed
  // accumulator and written to partial slot `chunk`, so the partial sums -- and with them every
  // fitted coefficient -- do not depend on the grid, on how many columns share the batch, or on how
  // many GPUs the columns were dealt to.  (The host picks grid = groups * parts so that all groups
  // stream the same rows at the same time and X comes from HBM about once.)
  const int n_live = prm.n_act_dev ? min(prm.n_act, (int)*prm.n_act_dev) : prm.n_act;
  const long long units = (long long)prm.groups * TC_NCH;
  const long long u_begin = (long long)blockIdx.x * units / gridDim.x;
  const long long u_end = (long long)(blockIdx.x + 1) * units / gridDim.x;
  struct TcItem { int g, z, t0, t1; const int32_t* tl; };
  auto get_item = [&](long long u, TcItem& it) -> bool {
    it.g = (int)(u / TC_NCH);
    it.z = (int)(u % TC_NCH);
    if (it.g * TC_BC >= n_live) return false;   // group emptied since the host last looked
    int cnt = prm.n_tiles;
    it.tl = nullptr;
    if (prm.tilelist) {
      const int f = prm.sp[it.g * TC_BC].fold;
      const int li = (f >= 0 && f < prm.n_lists - 1) ? f : prm.n_lists - 1;
      it.tl = prm.tilelist + (size_t)li * prm.n_tiles_ld;
      cnt = prm.tilecnt[li];
    }
    it.t0 = (int)((long long)cnt * it.z / TC_NCH);
    it.t1 = (int)((long long)cnt * (it.z + 1) / TC_NCH);
    return it.t1 > it.t0;       // empty chunks (fewer tiles than chunks) are skipped by every role alike
  };

  if (warp == 0) {
    // ================================ TMA producer ==========================================
    if (lane == 0) {
      uint32_t h = 0;  // running half-tile counter (ring position), persists across items
      int it_local = 0, g_prev = -1, w_loads = 0;
      for (long long u = u_begin; u < u_end; ++u) {
        TcItem it;
        if (!get_item(u, it)) continue;
        const int g = it.g, t0 = it.t0, t1 = it.t1;
        const int32_t* tlist = it.tl;
        if (g != g_prev) {
          // W_hi of a new group: the MMA warp signals w_free once per group it leaves, after all its
          // MMAs on that group (one phase per reload, so the parity wait cannot alias; acc_done
          // completes one phase per ITEM and this thread may be several one-tile items ahead)
          if (w_loads > 0) mbar_wait(&bars->w_free, (w_loads - 1) & 1, 100);
          ++w_loads;
          mbar_expect_tx(&bars->w_full, NCHUNK * WH_CHUNK);
#pragma unroll
          for (int c = 0; c < NCHUNK; ++c)
            tma_load_2d(s_wh + c * WH_CHUNK, &map_wh, c * 64, g * TC_BC, &bars->w_full);
          g_prev = g;
        }
        ++it_local;
        for (int t = t0; t < t1; ++t) {
          const int tile = tlist ? tlist[t] : t;
#pragma unroll
          for (int half = 0; half < 2; ++half, ++h) {
            const uint32_t sl = h % TC_NS, ph = (h / TC_NS) & 1;
            mbar_wait(&bars->empty[sl], ph ^ 1, 101);
            mbar_expect_tx(&bars->full[sl], SLOT_BYTES);
            const CUtensorMap* mp = half == 0 ? &map_xh : &map_xl;
#pragma unroll
            for (int c = 0; c < NCHUNK; ++c)
              tma_load_2d(s_ring + sl * SLOT_BYTES + c * X_CHUNK, mp, c * 64, tile * TC_R, &bars->full[sl]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ============================================
    // The whole warp runs the control flow (waits, counters) with warp-uniform values; the MMAs of
    // a tile are issued inside one `elect.sync` branch, which lets ptxas keep the descriptor
    // arithmetic in the uniform datapath (one UIADD3 per MMA instead of an elect/broadcast loop).
    {
      constexpr uint32_t idesc1 = make_idesc(TC_BC, TC_R, 0);            // GEMM1: N = 64 rows
      constexpr uint32_t idesc2 = make_idesc(TC_BC, NCHUNK * 64, 1);     // GEMM2: N = dpad, B MN-major
      constexpr int KS1 = NCHUNK * 4;                                    // K = dpad, 16 per MMA
      // descriptors: only the start-address field (low word, >>4 units) changes between MMAs
      const uint64_t a_base = make_desc(smem_u32(s_wh), 16, 1024);
      const uint64_t bk_base = make_desc(smem_u32(s_ring), 16, 1024);          // K-major view
      const uint64_t bmn_base = make_desc(smem_u32(s_ring), X_CHUNK, 1024);    // MN-major view
      uint32_t h = 0;       // half-tile counter, mirrors the producer
      uint32_t tcount = 0;  // tile counter (Z buffer = tcount & 1)
      int it_local = 0, w_loads = 0, g_prev = -1;
      for (long long u = u_begin; u < u_end; ++u, ++it_local) {
        TcItem it;
        if (!get_item(u, it)) { --it_local; continue; }
        const int nt = it.t1 - it.t0;
        if (it.g != g_prev) {       // weights of a new group (W_hi by TMA, W_lo staged by the epilogue warps)
          if (g_prev >= 0) {        // every MMA issued so far belongs to earlier groups: their completion frees W_hi
            if (elect_one()) tc_commit(&bars->w_free);
            __syncwarp();
          }
          mbar_wait(&bars->w_full, w_loads & 1, 200);
          mbar_wait(&bars->wl_full, w_loads & 1, 201);
          ++w_loads;
          g_prev = it.g;
        }
        if (it_local > 0) mbar_wait(&bars->acc_free, (it_local - 1) & 1, 202);
        tc_fence_after();

        auto issue_g1 = [&](uint32_t hh, uint32_t tc) {
          const uint32_t zcol = tmem + TM_Z0 + (tc & 1) * 64;
          {  // X_hi half tile: W_hi * X_hi + W_lo * X_hi
            const uint32_t sl = hh % TC_NS, ph = (hh / TC_NS) & 1;
            mbar_wait(&bars->full[sl], ph, 210);
            tc_fence_after();
            const uint64_t bb = bk_base + (uint64_t)((sl * SLOT_BYTES) >> 4);
            if (elect_one()) {
              if (prm.debug != 3) {
#pragma unroll
              for (int ks = 0; ks < KS1; ++ks) {
                const uint32_t aoff = ((ks >> 2) * WH_CHUNK + (ks & 3) * 32) >> 4;
                const uint32_t boff = ((ks >> 2) * X_CHUNK + (ks & 3) * 32) >> 4;
                mma_ss(zcol, a_base + aoff, bb + boff, idesc1, ks > 0 ? 1u : 0u);
                mma_ts(zcol, tmem + TM_WLO + ks * 8, bb + boff, idesc1, 1u);
              }
              }
            }
            __syncwarp();
          }
          {  // X_lo half tile: W_hi * X_lo
            const uint32_t sl = (hh + 1) % TC_NS, ph = ((hh + 1) / TC_NS) & 1;
            mbar_wait(&bars->full[sl], ph, 211);
            tc_fence_after();
            const uint64_t bb = bk_base + (uint64_t)((sl * SLOT_BYTES) >> 4);
            if (elect_one()) {
              if (prm.debug != 3) {
#pragma unroll
              for (int ks = 0; ks < KS1; ++ks) {
                const uint32_t aoff = ((ks >> 2) * WH_CHUNK + (ks & 3) * 32) >> 4;
                const uint32_t boff = ((ks >> 2) * X_CHUNK + (ks & 3) * 32) >> 4;
                mma_ss(zcol, a_base + aoff, bb + boff, idesc1, 1u);
              }
              }
              tc_commit(&bars->z_full[tc & 1]);
              if (!IS_FIT) {  // no GEMM2: the ring slots are free once GEMM1 has read them
                tc_commit(&bars->empty[hh % TC_NS]);
                tc_commit(&bars->empty[(hh + 1) % TC_NS]);
              }
            }
            __syncwarp();
          }
        };
        auto issue_g2 = [&](uint32_t hh, uint32_t tc, bool first_tile) {
          const uint32_t gcol = tmem + TM_Z0 + (tc & 1) * 64;
          mbar_wait(&bars->g_full[tc & 1], (tc >> 1) & 1, 220);
          tc_fence_after();
          const uint32_t sl_h = hh % TC_NS, sl_l = (hh + 1) % TC_NS;
          // B operand MN-major: N (features) contiguous within a chunk, chunks LBO apart;
          // K (rows) in groups of 8 rows SBO = 1024 B apart; one MMA covers 16 rows = 2048 B
          const uint64_t bh = bmn_base + (uint64_t)((sl_h * SLOT_BYTES) >> 4);
          const uint64_t bl = bmn_base + (uint64_t)((sl_l * SLOT_BYTES) >> 4);
          if (elect_one()) {
            if (prm.debug < 2) {
#pragma unroll
            for (int ks = 0; ks < TC_R / 16; ++ks) {
              mma_ts(tmem + TM_GRAD, gcol + ks * 16, bh + ks * 128, idesc2, (first_tile && ks == 0) ? 0u : 1u);
              if (prm.g_passes >= 3) mma_ts(tmem + TM_GRAD, gcol + ks * 16 + 8, bh + ks * 128, idesc2, 1u);
            }
            }
            tc_commit(&bars->empty[sl_h]);
            if (prm.debug < 2) {
#pragma unroll
            for (int ks = 0; ks < TC_R / 16; ++ks)
              mma_ts(tmem + TM_GRAD, gcol + ks * 16, bl + ks * 128, idesc2, 1u);
            }
            tc_commit(&bars->empty[sl_l]);
          }
          __syncwarp();
        };

        if (IS_FIT) {
          if (nt > 0) issue_g1(h, tcount);
          for (int i = 0; i < nt; ++i) {
            if (i + 1 < nt) issue_g1(h + 2 * (i + 1), tcount + i + 1);
            issue_g2(h + 2 * i, tcount + i, i == 0);
          }
        } else {
          // score: GEMM1 only; Z buffer b may be overwritten once the epilogue has consumed it
          for (int i = 0; i < nt; ++i) {
            const uint32_t tc = tcount + i;
            if (tc >= 2) { mbar_wait(&bars->g_full[tc & 1], ((tc >> 1) - 1) & 1, 230); tc_fence_after(); }
            issue_g1(h + 2 * i, tc);
          }
        }
        h += 2 * nt;
        tcount += nt;
        if (elect_one()) tc_commit(&bars->acc_done);
        __syncwarp();
      }
      __syncwarp();
    }
  } else {
    // ================================ epilogue warps ========================================
    // 16 warps: warps with the same (warp & 3) share TMEM lane quadrant q and take one 16-row
    // chunk of the tile each (qt = 0..3).  Four warps per scheduler hide the MUFU / TMEM latency;
    // the per-element instruction count, not the tensor pipe, was the limiter with 8 warps.
    const int q = warp & 3;                       // TMEM lane quadrant this warp may access
    const int qt = (warp - 2) >> 2;               // 16-row chunk of every tile
    const int lane_in_group = q * 32 + lane;      // slot within the group == TMEM lane
    const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16);
    constexpr float INV_G = 1.f / GSCALE;
    uint32_t tcount = 0;
    int it_local = 0, g_prev = -1;
    for (long long u = u_begin; u < u_end; ++u, ++it_local) {
      TcItem it;
      if (!get_item(u, it)) { --it_local; continue; }
      const int g = it.g, t0 = it.t0, t1 = it.t1;
      const int32_t* tlist = it.tl;
      const int nt = t1 - t0;
      const int z_part = it.z;
      const bool new_group = g != g_prev;
      g_prev = g;
      const int slot = g * TC_BC + lane_in_group;
      const bool valid = slot < n_live;
      TcSlotParam sp;
      sp.inv_t = 1.f; sp.bias = 0.f; sp.fold = -1; sp.pos = -1; sp.neg1 = 0; sp.col = -1;
      if (valid) sp = prm.sp[slot];
      // fit: work on z / 2^14 so that (row sign * 2^14) * z' = +-z and (row sign * 2^14) * sigma is
      // the scaled gradient entry; all power-of-two factors, results identical to the unscaled form
      const float zi = IS_FIT ? sp.inv_t * INV_G : sp.inv_t;
      const float zb0 = IS_FIT ? sp.bias * INV_G : sp.bias;
      const float* rsg = nullptr;
      if (MODE == TC_FIT_UNI) {
        const int f = prm.sp[g * TC_BC].fold;
        const int li = (f >= 0 && f < prm.n_lists - 1) ? f : prm.n_lists - 1;
        rsg = prm.rowsg + (size_t)li * prm.rowsg_ld;
      }
      if (new_group) {  // W_lo row of this slot -> TMEM (packed fp16 pairs), 16 columns (32 values) at a time.
         // The previous item's MMAs are done with W_lo: its acc_done was waited for below.
        const uint32_t* src = reinterpret_cast<const uint32_t*>(prm.Wl + (size_t)slot * (NCHUNK * 64));
#pragma unroll 1
        for (int c16 = qt; c16 < NCHUNK * 2; c16 += 4) {
          uint32_t r[16];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 v = valid ? __ldg(reinterpret_cast<const uint4*>(src + c16 * 16) + j) : make_uint4(0, 0, 0, 0);
            r[4 * j] = v.x; r[4 * j + 1] = v.y; r[4 * j + 2] = v.z; r[4 * j + 3] = v.w;
          }
          tmem_st16(tl + TM_WLO + c16 * 16, r);
        }
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&bars->wl_full);
      }
      // per-item sums as compensated fp32 pairs (TwoSum): the fp64 pipe is slow enough that two
      // DADDs per tile showed up as 16 % of the epilogue's stall samples
      float ls_hi = 0.f, ls_lo = 0.f, gs_hi = 0.f, gs_lo = 0.f;
      unsigned long long n_ok = 0, n_all = 0;
      for (int i = 0; i < nt; ++i, ++tcount) {
        const int t = tlist ? tlist[t0 + i] : t0 + i;
        const uint32_t zb = tl + TM_Z0 + (tcount & 1) * 64;
        // per-row data of this warp's 16 rows: issue the loads before waiting for the MMA so their
        // latency is hidden behind the wait
        uint32_t rm[16];
        {
          const uint4* rm4 = MODE == TC_FIT_UNI
                                 ? reinterpret_cast<const uint4*>(rsg + (size_t)t * TC_R + qt * 16)
                                 : reinterpret_cast<const uint4*>(prm.rowmeta + (size_t)t * TC_R + qt * 16);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 v = __ldg(rm4 + j);
            rm[4 * j] = v.x; rm[4 * j + 1] = v.y; rm[4 * j + 2] = v.z; rm[4 * j + 3] = v.w;
          }
        }
        float yv[16];
        if (MODE == TC_R2) {
          const float4* y4 = reinterpret_cast<const float4*>(prm.yreal + (size_t)t * TC_R + qt * 16);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float4 v = __ldg(y4 + j);
            yv[4 * j] = v.x; yv[4 * j + 1] = v.y; yv[4 * j + 2] = v.z; yv[4 * j + 3] = v.w;
          }
        }
        uint32_t yb16 = 0, mb16 = 0xFFFFu;
        if (MODE == TC_FIT && (prm.ybits || prm.mbits) && sp.col >= 0) {
          const size_t wi = (size_t)sp.col * prm.rb_words + (size_t)t * 2 + (qt >> 1);
          const unsigned sh = (qt & 1) * 16;
          if (prm.ybits) yb16 = (__ldg(prm.ybits + wi) >> sh) & 0xFFFFu;
          if (prm.mbits) mb16 = (__ldg(prm.mbits + wi) >> sh) & 0xFFFFu;
        }
        mbar_wait(&bars->z_full[tcount & 1], (tcount >> 1) & 1, 300);
        tc_fence_after();
        if (prm.debug == 4) { tc_fence_before(); mbar_arrive(&bars->g_full[tcount & 1]); continue; }
        float lt = 0.f, gt = 0.f;
        int ok_t = 0, all_t = 0;
        uint32_t zr[16];
        tmem_ld16(zb + qt * 16, zr);
        tmem_wait_ld();
        if (IS_FIT) {
          float gv[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float sg;                       // -y * 2^14 for a training row, 0 otherwise
            if (MODE == TC_FIT_UNI) {
              sg = __uint_as_float(rm[j]);
            } else {
              const uint32_t m = rm[j];
              const int fr = (int)(m >> 24);
              const int cls = (int)(m & 0x00FFFFFFu);
              bool yb = cls == sp.pos;
              bool train = (fr != 0xFF) && (fr != sp.fold) && (sp.neg1 == 0 || yb || cls == sp.neg1 - 1);
              if (MODE == TC_FIT) {             // staged row bit matrices (multilabel targets, sampled negatives)
                if (prm.ybits) yb = (yb16 >> j) & 1u;
                if (prm.mbits) train = train && ((mb16 >> j) & 1u);
              }
              sg = train ? (yb ? -GSCALE : GSCALE) : 0.f;
            }
            const float zp = fmaf(__uint_as_float(zr[j]), zi, zb0);
            const float u = sg * zp;        // = -y * z
            const float e = ex2_approx(-fabsf(u) * 1.4426950408889634f);
            const float s1 = 1.f + e;
            const float loss = fmaf(lg2_approx(s1), 0.6931471805599453f, fmaxf(u, 0.f));
            const float r = rcp_approx(s1);
            const float sig = (u >= 0.f) ? r : e * r;
            gv[j] = sg * sig;               // 2^14 * dloss/dz
            lt = fmaf(fabsf(sg), loss, lt); // 2^14 * loss
            gt += gv[j];
          }
          uint32_t out[16];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float a = gv[2 * j], b = gv[2 * j + 1];
            const uint32_t hi = pack_f16x2(a, b);
            const float2 hf2 = unpack_f16x2(hi);
            out[j] = hi;
            out[8 + j] = pack_f16x2(a - hf2.x, b - hf2.y);
          }
          tmem_st16(zb + qt * 16, out);
        } else if (MODE == TC_R2) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const uint32_t m = rm[j];
            const float z = fmaf(__uint_as_float(zr[j]), zi, zb0);
            const int fr = (int)(m >> 24);
            const bool in = (fr != 0xFF) && (sp.fold == -2 || (sp.fold >= 0 && fr == sp.fold) ||
                                             (sp.fold <= -3 && fr != (-3 - sp.fold)));
            const float r = yv[j] - z;
            lt += in ? r * r : 0.f;
            all_t += in ? 1 : 0;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const uint32_t m = rm[j];
            const float z = fmaf(__uint_as_float(zr[j]), zi, zb0);
            const int fr = (int)(m >> 24);
            const bool yb = (int)(m & 0x00FFFFFFu) == sp.pos;
            // fold code: f >= 0 rows of fold f; -2 all rows; -3-f rows NOT in fold f
            const bool in = (fr != 0xFF) && (sp.fold == -2 || (sp.fold >= 0 && fr == sp.fold) ||
                                             (sp.fold <= -3 && fr != (-3 - sp.fold)));
            ok_t += (in && ((z > 0.f) == yb)) ? 1 : 0;
            all_t += in ? 1 : 0;
          }
        }
        if (IS_FIT) tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&bars->g_full[tcount & 1]);
        {
          float sl = ls_hi + lt, bb = sl - ls_hi;
          ls_lo += (ls_hi - (sl - bb)) + (lt - bb);
          ls_hi = sl;
          float sg2 = gs_hi + gt, cc = sg2 - gs_hi;
          gs_lo += (gs_hi - (sg2 - cc)) + (gt - cc);
          gs_hi = sg2;
        }
        n_ok += ok_t;
        n_all += all_t;
      }
      // end of item: all MMAs done -> flush
      mbar_wait(&bars->acc_done, it_local & 1, 310);
      tc_fence_after();
      if (IS_FIT || MODE == TC_R2) {
        // the four warps of a quadrant hold partial sums of the same slots: exchange them through
        // the (now idle) Z buffer and let warp qt = 0 add them in a fixed order (deterministic)
        uint32_t v4[4] = {__float_as_uint(ls_hi), __float_as_uint(ls_lo), __float_as_uint(gs_hi),
                          __float_as_uint(gs_lo)};
        tmem_st4(tl + TM_Z0 + qt * 4, v4);
        tmem_wait_st();
        tc_fence_before();
        epi_bar_sync();
        tc_fence_after();
      }
      if (IS_FIT) {
        float* dst = prm.gradp + ((size_t)z_part * prm.n_act + slot) * prm.ldw;
#pragma unroll 1
        for (int c16 = qt; c16 < NCHUNK * 4; c16 += 4) {
          uint32_t r[16];
          tmem_ld16(tl + TM_GRAD + c16 * 16, r);
          tmem_wait_ld();
          if (valid && nt > 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<uint4*>(dst + c16 * 16 + 4 * j) =
                  make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
          }
        }
      }
      if ((IS_FIT || MODE == TC_R2) && qt == 0) {
        uint32_t r[16];
        tmem_ld16(tl + TM_Z0, r);
        tmem_wait_ld();
        double ls = 0.0, gs = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          ls += (double)__uint_as_float(r[4 * k]) + (double)__uint_as_float(r[4 * k + 1]);
          gs += (double)__uint_as_float(r[4 * k + 2]) + (double)__uint_as_float(r[4 * k + 3]);
        }
        if (IS_FIT) {
          if (valid && nt > 0) {   // partial (z_part, slot) has exactly one writer
            prm.lossp[(size_t)z_part * prm.n_act + slot] = ls * (double)INV_G;
            prm.gsump[(size_t)z_part * prm.n_act + slot] = gs * (double)INV_G;
          }
        } else if (valid && nt > 0) {
          atomicAdd(prm.lossp + slot, ls);   // sum of squared residuals (parts add up)
        }
      }
      if (MODE == TC_R2) {
        if (valid && n_all > 0) atomicAdd(prm.count + slot, n_all);
      } else if (MODE == TC_SCORE && valid && n_all > 0) {
        atomicAdd(prm.correct + slot, n_ok);
        atomicAdd(prm.count + slot, n_all);
      }
      tc_fence_before();
      mbar_arrive(&bars->acc_free);
    }
  }

  // teardown
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D fp16 row-major [rows x cols] tensor, box = [box_rows x 64 cols], 128B swizzle
int tc_make_map_2d(Ctx* c, CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(c, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * sizeof(__half)};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char b[128];
    snprintf(b, sizeof(b), "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return fail(c, b);
  }
  return 0;
}

bool tc_supported(const Ctx* c) { return c->d >= 1 && c->d <= 256; }

void tc_free(Ctx* c) {
  TcData& t = c->tc;
  if (t.Xh) cudaFree(t.Xh);
  if (t.Xl) cudaFree(t.Xl);
  if (t.rowmeta) cudaFree(t.rowmeta);
  if (t.yreal_pad) cudaFree(t.yreal_pad);
  if (t.tilelist) cudaFree(t.tilelist);
  if (t.tilecnt) cudaFree(t.tilecnt);
  if (t.rowsg) cudaFree(t.rowsg);
  if (t.xscale) cudaFree(t.xscale);
  if (t.gscale) cudaFree(t.gscale);
  t = TcData();
}

// Build (or refresh) the fp16-split copy of X and the per-row metadata.
int tc_prepare(Ctx* c) {
  TcData& t = c->tc;
  const int64_t n = c->n;
  const int d = (int)c->d, ldx = (int)c->ldx;
  const int dpad = (d + 63) / 64 * 64;
  const int64_t npad = (n + TC_R - 1) / TC_R * TC_R;
  if (!t.x_valid) {
    if (t.Xh && (t.dpad != dpad || t.npad != npad)) tc_free(c);   // same shape restaged: keep the buffers
    if (!t.Xh) {
      t.dpad = dpad;
      t.npad = npad;
      SKD_CUDA(c, cudaMalloc((void**)&t.Xh, (size_t)npad * dpad * sizeof(__half)));
      SKD_CUDA(c, cudaMalloc((void**)&t.Xl, (size_t)npad * dpad * sizeof(__half)));
      SKD_CUDA(c, cudaMalloc((void**)&t.rowmeta, (size_t)npad * sizeof(uint32_t)));
      SKD_CUDA(c, cudaMalloc((void**)&t.xscale, (size_t)dpad * sizeof(float)));
      SKD_CUDA(c, cudaMalloc((void**)&t.gscale, (size_t)dpad * sizeof(double)));
    }
    unsigned int* colmax;
    SKD_CUDA(c, cudaMalloc((void**)&colmax, (size_t)dpad * sizeof(unsigned int)));
    SKD_CUDA(c, cudaMemsetAsync(colmax, 0, (size_t)dpad * sizeof(unsigned int), c->stream));
    dim3 g1((d + 127) / 128, (unsigned)((n + 4095) / 4096));
    tc_colmax_kernel<<<g1, 128, 0, c->stream>>>(c->X, n, ldx, d, colmax);
    tc_scale_kernel<<<(dpad + 127) / 128, 128, 0, c->stream>>>(colmax, d, dpad, t.xscale, t.gscale);
    int64_t total = npad * dpad;
    tc_split_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(
        c->X, n, npad, ldx, d, dpad, t.xscale, (__half*)t.Xh, (__half*)t.Xl);
    c->launches += 3;
    SKD_CUDA(c, cudaGetLastError());
    SKD_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaFree(colmax);
    if (tc_make_map_2d(c, &t.map_xh, t.Xh, (uint64_t)npad, (uint64_t)dpad, TC_R)) return 1;
    if (tc_make_map_2d(c, &t.map_xl, t.Xl, (uint64_t)npad, (uint64_t)dpad, TC_R)) return 1;
    t.x_valid = true;
    t.meta_valid = false;
  }
  if (!t.meta_valid) {
    if (!c->ycls && !c->yreal) return fail(c, "tc_prepare: neither labels nor targets staged");
    tc_rowmeta_kernel<<<(unsigned)((npad + 255) / 256), 256, 0, c->stream>>>(c->ycls, c->fold, n, npad,
                                                                            t.rowmeta);
    c->launches += 1;
    {  // per-fold tile lists: list f = tiles holding at least one row NOT in fold f; last list = all
      const int n_tiles = (int)(npad / TC_R);
      const int nf = (c->fold && c->n_folds <= 32 && (int64_t)c->h_fold.size() == n) ? c->n_folds : 0;
      std::vector<int32_t> hl((size_t)(nf + 1) * n_tiles), hc(nf + 1, 0);
      for (int tt = 0; tt < n_tiles; ++tt) {
        uint32_t mask = 0;
        const int64_t r1 = std::min<int64_t>(n, (int64_t)(tt + 1) * TC_R);
        if (nf) for (int64_t r = (int64_t)tt * TC_R; r < r1; ++r) mask |= 1u << c->h_fold[r];
        for (int f = 0; f < nf; ++f)
          if (mask & ~(1u << f)) hl[(size_t)f * n_tiles + hc[f]++] = tt;
        hl[(size_t)nf * n_tiles + hc[nf]++] = tt;
      }
      if (t.tilelist) { cudaFree(t.tilelist); t.tilelist = nullptr; }
      if (t.tilecnt) { cudaFree(t.tilecnt); t.tilecnt = nullptr; }
      SKD_CUDA(c, cudaMalloc((void**)&t.tilelist, hl.size() * 4));
      SKD_CUDA(c, cudaMalloc((void**)&t.tilecnt, hc.size() * 4));
      SKD_CUDA(c, cudaMemcpyAsync(t.tilelist, hl.data(), hl.size() * 4, cudaMemcpyHostToDevice, c->stream));
      SKD_CUDA(c, cudaMemcpyAsync(t.tilecnt, hc.data(), hc.size() * 4, cudaMemcpyHostToDevice, c->stream));
      SKD_CUDA(c, cudaStreamSynchronize(c->stream));
      t.min_list_tiles = n_tiles;
      for (int f = 0; f <= nf; ++f) t.min_list_tiles = std::min(t.min_list_tiles, (int)hc[f]);
      if (t.rowsg && t.n_lists != nf + 1) { cudaFree(t.rowsg); t.rowsg = nullptr; }
      t.n_lists = nf + 1;
      t.rowsg_valid = false;
    }
    if (c->yreal) {
      if (!t.yreal_pad) SKD_CUDA(c, cudaMalloc((void**)&t.yreal_pad, (size_t)npad * sizeof(float)));
      SKD_CUDA(c, cudaMemsetAsync(t.yreal_pad, 0, (size_t)npad * sizeof(float), c->stream));
      SKD_CUDA(c, cudaMemcpyAsync(t.yreal_pad, c->yreal, (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
    }
    SKD_CUDA(c, cudaGetLastError());
    t.meta_valid = true;
  }
  return 0;
}

int tc_export(Ctx* c, LogregWork& w, int n_act_upper, const double* xin, int fit_intercept) {
  TcData& t = c->tc;
  tc_export_kernel<<<n_act_upper, 128, 0, c->stream>>>(w.vec, w.vec_stride, w.slot, w.n_act, (int)c->d,
                                                       t.dpad, t.xscale, (__half*)w.Wh, (__half*)w.Wl,
                                                       (TcSlotParam*)w.sp, xin, fit_intercept);
  c->launches += 1;
  SKD_CUDA(c, cudaGetLastError());
  return 0;
}

size_t tc_slot_param_bytes() { return sizeof(TcSlotParam); }
int tc_partials_per_slot() { return TC_NCH; }

template <int MODE>
static cudaError_t tc_launch(int nchunk, int grid, size_t smem, cudaStream_t st, const CUtensorMap& xh,
                             const CUtensorMap& xl, const CUtensorMap& wh, const TcParams& prm) {
  switch (nchunk) {
#define TC_CASE(N)                                                                                   \
  case N: {                                                                                          \
    static bool attr = false;                                                                        \
    if (!attr) {                                                                                     \
      cudaError_t e = cudaFuncSetAttribute(tc_eval_kernel<N, MODE>,                                  \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);     \
      if (e != cudaSuccess) return e;                                                                \
      attr = true;                                                                                   \
    }                                                                                                \
    tc_eval_kernel<N, MODE><<<grid, TC_THREADS, smem, st>>>(xh, xl, wh, prm);                        \
    break;                                                                                           \
  }
    TC_CASE(1) TC_CASE(2) TC_CASE(3) TC_CASE(4)
#undef TC_CASE
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

static int tc_run(Ctx* c, LogregWork& w, int n_act, int mode, int* nz_used, unsigned long long* dcorrect,
                  unsigned long long* dcount) {
  TcData& t = c->tc;
  if (nz_used) *nz_used = 0;
  if (n_act <= 0) return 0;
  const int nchunk = t.dpad / 64;
  const int groups = (n_act + TC_BC - 1) / TC_BC;
  const int n_tiles = (int)(t.npad / TC_R);
  // Grid: units = (group, chunk) pairs (see the kernel).  With few groups every group gets the same
  // number of CTAs ("parts", at most one per chunk), so all groups stream the same rows at the same
  // time and the other groups' reads hit L2; with many groups the units are simply dealt evenly.
  const long long units = (long long)groups * TC_NCH;
  int grid = c->sm_count;
  int parts = groups <= c->sm_count ? c->sm_count / groups : 0;
  if (parts > TC_NCH) parts = TC_NCH;
  if (parts > 0) grid = groups * parts;
  if ((long long)grid > units) grid = (int)units;
  const int nz = TC_NCH;
  if (mode == TC_FIT) {
    if ((int64_t)nz * n_act > w.cap_sc) return fail(c, "tc_eval: partial buffer too small");
    // every (chunk, slot) partial has exactly one writer; only chunks without tiles (fewer tiles
    // than chunks in some list) are never written and must read as zero
    if (t.min_list_tiles < TC_NCH) {
      SKD_CUDA(c, cudaMemsetAsync(w.lossp, 0, (size_t)nz * n_act * sizeof(double), c->stream));
      SKD_CUDA(c, cudaMemsetAsync(w.gsump, 0, (size_t)nz * n_act * sizeof(double), c->stream));
      SKD_CUDA(c, cudaMemsetAsync(w.gradp, 0, (size_t)nz * n_act * w.ldw * sizeof(float), c->stream));
    }
  }
  CUtensorMap map_wh;
  if (tc_make_map_2d(c, &map_wh, w.Wh, (uint64_t)w.slots_pad_cap, (uint64_t)t.dpad, TC_BC)) return 1;
  TcParams prm;
  prm.Wl = (const __half*)w.Wl;
  prm.sp = (const TcSlotParam*)w.sp;
  prm.rowmeta = t.rowmeta;
  prm.yreal = t.yreal_pad;
  prm.lossp = w.lossp;
  prm.gsump = w.gsump;
  prm.gradp = w.gradp;
  prm.correct = dcorrect;
  prm.count = dcount;
  prm.n_act = n_act;
  prm.n_act_dev = (mode == TC_FIT) ? w.n_act : nullptr;
  prm.groups = groups;
  prm.n_tiles = n_tiles;
  prm.ldw = w.ldw;
  { const char* dbg = getenv("SKDIST_B200_TC_DEBUG"); prm.debug = dbg ? atoi(dbg) : 0; }
  { const char* gp = getenv("SKDIST_B200_TC_GPASSES"); prm.g_passes = gp ? atoi(gp) : 3; }
  prm.ybits = mode == TC_FIT ? w.ybits : nullptr;
  prm.mbits = mode == TC_FIT ? w.mbits : nullptr;
  prm.rb_words = w.rb_words;
  const bool uni = mode == TC_FIT && w.grouped && w.uni_pos >= 0 && c->ycls && !w.ybits && !w.mbits;
  if (uni && (!t.rowsg_valid || t.rowsg_pos != w.uni_pos)) {
    if (!t.rowsg) SKD_CUDA(c, cudaMalloc((void**)&t.rowsg, (size_t)t.n_lists * t.npad * sizeof(float)));
    tc_rowsg_kernel<<<(unsigned)((t.npad + 255) / 256), 256, 0, c->stream>>>(c->ycls, c->fold, c->n, t.npad,
                                                                          t.n_lists, w.uni_pos, t.rowsg);
    c->launches += 1;
    SKD_CUDA(c, cudaGetLastError());
    t.rowsg_pos = w.uni_pos;
    t.rowsg_valid = true;
  }
  prm.rowsg = uni ? t.rowsg : nullptr;
  prm.rowsg_ld = t.npad;
  const bool lists = mode == TC_FIT && w.grouped && t.tilelist;
  prm.tilelist = lists ? t.tilelist : nullptr;
  prm.tilecnt = lists ? t.tilecnt : nullptr;
  prm.n_lists = t.n_lists;
  prm.n_tiles_ld = n_tiles;
  const size_t smem = 1024 + (size_t)nchunk * (TC_BC * 128) + (size_t)TC_NS * nchunk * (TC_R * 128) +
                      sizeof(TcBarriers) + 64;
  if (mode == TC_R2 && !t.yreal_pad) return fail(c, "tc_r2: targets not staged");
  cudaError_t e = uni ? tc_launch<TC_FIT_UNI>(nchunk, grid, smem, c->stream, t.map_xh, t.map_xl, map_wh, prm)
                  : mode == TC_FIT ? tc_launch<TC_FIT>(nchunk, grid, smem, c->stream, t.map_xh, t.map_xl, map_wh, prm)
                  : mode == TC_SCORE ? tc_launch<TC_SCORE>(nchunk, grid, smem, c->stream, t.map_xh, t.map_xl, map_wh, prm)
                                     : tc_launch<TC_R2>(nchunk, grid, smem, c->stream, t.map_xh, t.map_xl, map_wh, prm);
  c->launches += 1;
  if (e != cudaSuccess) return fail(c, std::string("tc_eval launch: ") + cudaGetErrorString(e));
  if (nz_used) *nz_used = nz;
  return 0;
}

int tc_eval(Ctx* c, LogregWork& w, int n_act, int* nz_used) {
  return tc_run(c, w, n_act, TC_FIT, nz_used, nullptr, nullptr);
}

// Accuracy counts of n_act slots whose weights were exported with tc_export (sp.fold = scoring code).
int tc_score(Ctx* c, LogregWork& w, int n_act, int64_t* dcorrect, int64_t* dcount) {
  return tc_run(c, w, n_act, TC_SCORE, nullptr, (unsigned long long*)dcorrect, (unsigned long long*)dcount);
}

// Sum of squared residuals / row counts of n_act regression slots (sp.fold = scoring code).
int tc_r2(Ctx* c, LogregWork& w, int n_act, double* dsse, int64_t* dcount) {
  double* keep = w.lossp;
  w.lossp = dsse;
  int rc = tc_run(c, w, n_act, TC_R2, nullptr, nullptr, (unsigned long long*)dcount);
  w.lossp = keep;
  return rc;
}

}  // namespace skd
