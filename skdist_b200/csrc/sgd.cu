// sgd.cu -- exact-order column-batched SGD: every one-vs-rest label column at once.
//
// Replaces K invocations of the reference's `_fit_binary` (ref multiclass.py:109-152) with
// estimator = SGDClassifier: SK/linear_model/_stochastic_gradient.py:387-515 (fit_binary) ->
// SK/linear_model/_sgd_fast.pyx.tp:274-640 (_plain_sgd32) with WeightVector32
// (SK/utils/_weight_vector.pyx.tp) and the Fisher-Yates / xorshift shuffle
// (SK/utils/_seq_dataset.pyx.tp:137-145, SK/utils/_random.pxd:20-34).
//
// SGD is sequential in the samples but independent across label columns, and every column
// sees the same shuffled sample order (same seed).  One WARP owns one column: its d float32
// weights live in registers (d/32 per lane) for the whole epoch, the warp walks the shuffled
// rows (next row prefetched while the current one is processed), and reproduces the reference's
// arithmetic operation by operation -- float32 products accumulated in float64, float32 lazy
// scale `wscale`, float64 intercept / objective, weight update w = float(double(w) + double(x)*q)
// -- so hinge-loss fits are bit-identical to scikit-learn.  No tensor cores: the work per sample
// is two length-d vector operations per column, bound by the FP64 pipe and L2 latency.
#include <math.h>
#include <stdio.h>

#include <chrono>
#include <stdlib.h>

#include "skd_internal.h"

namespace skd {

struct SgdState {
  double wscale, sq_norm, intercept, best_objective, t;
  int32_t no_improve, done, n_iter, status;
};

enum { SGD_HINGE = 0, SGD_LOG = 1 };

// per-sample learning rate and weight-decay factor of one epoch (class independent)
__global__ void sgd_schedule_kernel(int64_t n, double t0, double alpha, double optimal_init,
                                    int lr_type, double eta0, double power_t, double* __restrict__ eta,
                                    float* __restrict__ cfac) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double t = t0 + (double)i;
  double e;
  if (lr_type == 0) e = 1.0 / (alpha * (optimal_init + t - 1.0));   // "optimal"
  else if (lr_type == 1) e = eta0;                                   // "constant"
  else e = eta0 / pow(t, power_t);                                   // "invscaling"
  eta[i] = e;
  cfac[i] = (float)fmax(0.0, __dsub_rn(1.0, __dmul_rn(e, alpha)));  // w.scale(max(0, 1 - eta*alpha)) arg as float
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int DPL, int LOSS>
__global__ void __launch_bounds__(128)
sgd_epoch_kernel(const float* __restrict__ X, int ldx, int d, const int32_t* __restrict__ ycls,
                 const int32_t* __restrict__ order, const double* __restrict__ eta,
                 const float* __restrict__ cfac, int64_t n, const int32_t* __restrict__ active,
                 int n_active, const int32_t* __restrict__ col_pos, float* __restrict__ W, int ldw,
                 SgdState* __restrict__ state, double alpha, int fit_intercept, double tol,
                 int n_iter_no_change) {
  const int lane = threadIdx.x & 31;
  const int a = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (a >= n_active) return;
  const int col = active[a];
  const int pos = col_pos[col];
  SgdState st = state[col];
  float w[DPL];
#pragma unroll
  for (int j = 0; j < DPL; ++j) {
    const int k = lane + 32 * j;
    w[j] = k < d ? W[(size_t)col * ldw + k] : 0.f;
  }
  double wscale = st.wscale, sq_norm = st.sq_norm, intercept = st.intercept;
  double objective_sum = 0.0;

  // software pipeline: everything sample i+1 needs (row index, features, label, step size) is
  // requested while sample i is processed, so the per-sample dependency chain never waits on L2
  float xn[DPL];
  int row = order[0];
  int r_nxt = n > 1 ? order[1] : 0;
#pragma unroll
  for (int j = 0; j < DPL; ++j) {
    const int k = lane + 32 * j;
    xn[j] = k < ldx ? __ldg(X + (size_t)row * ldx + k) : 0.f;
  }
  int yc_n = ycls[row];
  double e_n = eta[0];
  float c_n = cfac[0];
  for (int64_t i = 0; i < n; ++i) {
    float x[DPL];
#pragma unroll
    for (int j = 0; j < DPL; ++j) x[j] = xn[j];
    const double y = (yc_n == pos) ? 1.0 : -1.0;
    const double e = e_n;
    const float c = c_n;
    if (i + 1 < n) {
      row = r_nxt;
      r_nxt = i + 2 < n ? order[i + 2] : 0;
#pragma unroll
      for (int j = 0; j < DPL; ++j) {
        const int k = lane + 32 * j;
        xn[j] = k < ldx ? __ldg(X + (size_t)row * ldx + k) : 0.f;
      }
      yc_n = ycls[row];
      e_n = eta[i + 1];
      c_n = cfac[i + 1];
    }
    // p = w.dot(x) + intercept          (WeightVector32.dot)
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < DPL; ++j) acc += (double)__fmul_rn(w[j], x[j]);
    acc = warp_sum(acc);
    const double p = (double)(float)(acc * wscale) + intercept;
    // loss / gradient
    double cur_loss, dloss;
    if (LOSS == SGD_HINGE) {       // Hinge on y in {-1,+1} (SK/linear_model/_sgd_fast.pyx.tp:131-146)
      const double z = p * y;
      if (z <= 1.0) { cur_loss = 1.0 - z; dloss = -y; } else { cur_loss = 0.0; dloss = 0.0; }
    } else {                       // CyHalfBinomialLoss on y in {0,1} (SK/_loss/_loss.pyx.tp:256-266,686-725)
      const double y01 = y > 0.0 ? 1.0 : 0.0;
      double l1p;
      if (p <= -37.0) l1p = exp(p);
      else if (p <= -2.0) l1p = log1p(exp(p));
      else if (p <= 18.0) l1p = log(__dadd_rn(1.0, exp(p)));
      else if (p <= 33.3) l1p = __dadd_rn(p, exp(-p));
      else l1p = p;
      cur_loss = __dsub_rn(l1p, __dmul_rn(y01, p));
      if (p > -37.0) {
        const double et = exp(-p);
        dloss = __ddiv_rn(__dsub_rn(1.0 - y01, __dmul_rn(y01, et)), __dadd_rn(1.0, et));
      } else {
        dloss = __dsub_rn(exp(p), y01);
      }
    }
    const float normf = (float)sqrt(sq_norm);
    objective_sum = __dadd_rn(objective_sum,
                              __dadd_rn(cur_loss, __dmul_rn(alpha, __dmul_rn(0.5, (double)__fmul_rn(normf, normf)))));
    if (dloss < -1e12) dloss = -1e12; else if (dloss > 1e12) dloss = 1e12;
    const double update = -e * dloss;
    // w.scale(c)
    wscale *= (double)c;
    sq_norm *= (double)__fmul_rn(c, c);
    if (wscale < 1e-6) {          // reset_wscale(): sscal by float(wscale)
      const float wf = (float)wscale;
#pragma unroll
      for (int j = 0; j < DPL; ++j) w[j] = __fmul_rn(w[j], wf);
      wscale = 1.0;
    }
    if (update != 0.0) {           // w.add(x, update)
      const float cf = (float)update, wsf = (float)wscale;
      const double q = (double)__fdiv_rn(cf, wsf);
      double acc2 = 0.0;
#pragma unroll
      for (int j = 0; j < DPL; ++j) {
        w[j] = (float)fma((double)x[j], q, (double)w[j]);
        acc2 += (double)__fmul_rn(w[j], w[j]);
      }
      acc2 = warp_sum(acc2);
      sq_norm = acc2 * (double)__fmul_rn(wsf, wsf);
      if (fit_intercept) intercept += update;
    }
  }
  // end of epoch (SK/linear_model/_sgd_fast.pyx.tp:570-628)
  bool finite = isfinite(intercept);
#pragma unroll
  for (int j = 0; j < DPL; ++j) finite = finite && isfinite(w[j]);
  finite = __all_sync(0xffffffffu, finite);
#pragma unroll
  for (int j = 0; j < DPL; ++j) {
    const int k = lane + 32 * j;
    if (k < d) W[(size_t)col * ldw + k] = w[j];
  }
  if (lane == 0) {
    st.wscale = wscale; st.sq_norm = sq_norm; st.intercept = intercept;
    st.t += (double)n;
    st.n_iter += 1;
    if (!finite) { st.done = 1; st.status = 5; }
    else {
      const double obj = objective_sum / (double)n;
      if (tol > -INFINITY && obj > st.best_objective - tol) st.no_improve += 1; else st.no_improve = 0;
      if (obj < st.best_objective) st.best_objective = obj;
      if (st.no_improve >= n_iter_no_change) { st.done = 1; st.status = 1; }
    }
    state[col] = st;
  }
}

// Hinge loss, speculative blocks.  With hinge loss a sample whose margin y*p exceeds 1 changes
// nothing but the lazy scale (wscale, sq_norm) and the objective sum -- scalars.  The warp therefore
// computes the dot products of the next T = 16 samples against the CURRENT weights at once
// (independent FMAs, all loads in flight together), reduces them with a butterfly, and then walks
// the 16 samples in order doing only the scalar recurrence; the first margin violator (or a
// reset_wscale) applies its weight update exactly as the sequential kernel does and the block
// restarts behind it.  Every value is computed by the same operations in the same order as in
// sgd_epoch_kernel, so the result stays bit-identical to scikit-learn; only the waiting changes.
template <int DPL>
__global__ void __launch_bounds__(128)
sgd_epoch_spec_kernel(const float* __restrict__ X, int ldx, int d, const int32_t* __restrict__ ycls,
                      const int32_t* __restrict__ order, const double* __restrict__ eta,
                      const float* __restrict__ cfac, int64_t n, const int32_t* __restrict__ active,
                      int n_active, const int32_t* __restrict__ col_pos, float* __restrict__ W, int ldw,
                      SgdState* __restrict__ state, double alpha, int fit_intercept, double tol,
                      int n_iter_no_change) {
  constexpr int T = 16;                                                        // samples per block
  constexpr int S = (64 / DPL) < 2 ? 2 : ((64 / DPL) > 8 ? 8 : (64 / DPL));   // samples per load batch
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int a = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (a >= n_active) return;
  const int col = active[a];
  const int pos = col_pos[col];
  SgdState st = state[col];
  float w[DPL];
#pragma unroll
  for (int j = 0; j < DPL; ++j) {
    const int k = lane + 32 * j;
    w[j] = k < d ? W[(size_t)col * ldw + k] : 0.f;
  }
  double wscale = st.wscale, sq_norm = st.sq_norm, intercept = st.intercept;
  double objective_sum = 0.0;

  int64_t i0 = 0;
  while (i0 < n) {
    const int Te = (int)((n - i0) < T ? (n - i0) : T);
    // lane t < Te owns the metadata of sample i0 + t
    int row_l = 0;
    double y_l = 0.0, e_l = 0.0;
    float c_l = 0.f;
    if (lane < Te) {
      row_l = order[i0 + lane];
      e_l = eta[i0 + lane];
      c_l = cfac[i0 + lane];
      y_l = (ycls[row_l] == pos) ? 1.0 : -1.0;
    }
    // dot products of all block samples with the current weights
    double part[T];
#pragma unroll
    for (int t = 0; t < T; ++t) part[t] = 0.0;
    float xb[2][S][DPL];
#pragma unroll
    for (int s2 = 0; s2 < S; ++s2) {
      const int r = __shfl_sync(FULL, row_l, s2);
#pragma unroll
      for (int j = 0; j < DPL; ++j) {
        const int k = lane + 32 * j;
        xb[0][s2][j] = k < ldx ? __ldg(X + (size_t)r * ldx + k) : 0.f;
      }
    }
#pragma unroll
    for (int b = 0; b < T / S; ++b) {
      if (b + 1 < T / S) {
#pragma unroll
        for (int s2 = 0; s2 < S; ++s2) {
          const int r = __shfl_sync(FULL, row_l, (b + 1) * S + s2);
#pragma unroll
          for (int j = 0; j < DPL; ++j) {
            const int k = lane + 32 * j;
            xb[(b + 1) & 1][s2][j] = k < ldx ? __ldg(X + (size_t)r * ldx + k) : 0.f;
          }
        }
      }
#pragma unroll
      for (int s2 = 0; s2 < S; ++s2)
#pragma unroll
        for (int j = 0; j < DPL; ++j) part[b * S + s2] += (double)__fmul_rn(w[j], xb[b & 1][s2][j]);
    }
    // butterfly reduce-scatter: afterwards lane L holds the full sum of sample L >> 1
#pragma unroll
    for (int half = T / 2, m = 16; half >= 1; half >>= 1, m >>= 1) {
      const bool up = (lane & m) != 0;
#pragma unroll
      for (int q = 0; q < half; ++q) {
        const double send = up ? part[q] : part[q + half];
        const double keep = up ? part[q + half] : part[q];
        part[q] = keep + __shfl_xor_sync(FULL, send, m);
      }
    }
    const double mysum = part[0] + __shfl_xor_sync(FULL, part[0], 1);

    // ---- ordered walk over the block, restructured so that only what is inherently serial is serial:
    //  A. the lazy-scale recurrences wscale *= c_t, sq_norm *= c_t^2 (two chains of 16 dependent
    //     multiplies, evaluated redundantly by every lane; lane t keeps the values BEFORE sample t);
    //  B. lane t evaluates sample t: prediction, margin, loss and objective term;
    //  C. the first event (margin violator or reset_wscale) is found with one ballot;
    //  D. the objective terms of samples 0..event are added in order;
    //  E. the event's weight update is applied exactly as in the sequential kernel.
    float c_all[T];
#pragma unroll
    for (int q = 0; q < T; ++q) c_all[q] = __shfl_sync(FULL, c_l, q);
    double my_ws = wscale, my_sq = sq_norm;        // state before this lane's sample
    double my_ws_after = wscale, my_sq_after = sq_norm;
    {
      double ws = wscale, sq = sq_norm;
#pragma unroll
      for (int q = 0; q < T; ++q) {
        if (lane == q) { my_ws = ws; my_sq = sq; }
        ws *= (double)c_all[q];
        sq *= (double)__fmul_rn(c_all[q], c_all[q]);
        if (lane == q) { my_ws_after = ws; my_sq_after = sq; }
      }
    }
    const double acc_l = __shfl_sync(FULL, mysum, (2 * lane) & 31);   // lane t < 16: dot of sample t
    const double p_l = (double)(float)(acc_l * my_ws) + intercept;
    const double z_l = p_l * y_l;
    const bool viol_l = lane < Te && z_l <= 1.0;
    const bool reset_l = lane < Te && my_ws_after < 1e-6;
    const double loss_l = viol_l ? 1.0 - z_l : 0.0;
    const float normf_l = (float)sqrt(my_sq);
    const double term_l = __dadd_rn(loss_l, __dmul_rn(alpha, __dmul_rn(0.5, (double)__fmul_rn(normf_l, normf_l))));
    const unsigned evmask = __ballot_sync(FULL, viol_l || reset_l);
    const int ev = evmask ? __ffs(evmask) - 1 : -1;        // first event sample, -1: none in this block
    const int last = ev >= 0 ? ev : Te - 1;                // samples 0..last are consumed
    {
      double terms[T];
#pragma unroll
      for (int q = 0; q < T; ++q) terms[q] = __shfl_sync(FULL, term_l, q);
#pragma unroll
      for (int q = 0; q < T; ++q)
        if (q <= last) objective_sum = __dadd_rn(objective_sum, terms[q]);
    }
    wscale = __shfl_sync(FULL, my_ws_after, last);
    sq_norm = __shfl_sync(FULL, my_sq_after, last);
    if (ev >= 0) {
      const bool is_reset = (evmask >> ev) & 1u ? __shfl_sync(FULL, (int)reset_l, ev) != 0 : false;
      const bool is_viol = __shfl_sync(FULL, (int)viol_l, ev) != 0;
      if (is_reset) {                 // reset_wscale(): sscal by float(wscale)
        const float wf = (float)wscale;
#pragma unroll
        for (int j2 = 0; j2 < DPL; ++j2) w[j2] = __fmul_rn(w[j2], wf);
        wscale = 1.0;
      }
      if (is_viol) {
        const double y = __shfl_sync(FULL, y_l, ev);
        const double e = __shfl_sync(FULL, e_l, ev);
        const double update = -e * (-y);
        if (update != 0.0) {           // w.add(x, update)
          const int r = __shfl_sync(FULL, row_l, ev);
          const float cf = (float)update, wsf = (float)wscale;
          const double qd = (double)__fdiv_rn(cf, wsf);
          double acc2 = 0.0;
#pragma unroll
          for (int j2 = 0; j2 < DPL; ++j2) {
            const int k = lane + 32 * j2;
            const float xv = k < ldx ? __ldg(X + (size_t)r * ldx + k) : 0.f;
            w[j2] = (float)fma((double)xv, qd, (double)w[j2]);
            acc2 += (double)__fmul_rn(w[j2], w[j2]);
          }
          acc2 = warp_sum(acc2);
          sq_norm = acc2 * (double)__fmul_rn(wsf, wsf);
          if (fit_intercept) intercept += update;
        }
      }
    }
    const int t = last + 1;
    i0 += t;
  }
  // end of epoch (SK/linear_model/_sgd_fast.pyx.tp:570-628)
  bool finite = isfinite(intercept);
#pragma unroll
  for (int j = 0; j < DPL; ++j) finite = finite && isfinite(w[j]);
  finite = __all_sync(FULL, finite);
#pragma unroll
  for (int j = 0; j < DPL; ++j) {
    const int k = lane + 32 * j;
    if (k < d) W[(size_t)col * ldw + k] = w[j];
  }
  if (lane == 0) {
    st.wscale = wscale; st.sq_norm = sq_norm; st.intercept = intercept;
    st.t += (double)n;
    st.n_iter += 1;
    if (!finite) { st.done = 1; st.status = 5; }
    else {
      const double obj = objective_sum / (double)n;
      if (tol > -INFINITY && obj > st.best_objective - tol) st.no_improve += 1; else st.no_improve = 0;
      if (obj < st.best_objective) st.best_objective = obj;
      if (st.no_improve >= n_iter_no_change) { st.done = 1; st.status = 1; }
    }
    state[col] = st;
  }
}

// w.reset_wscale() at the end of _plain_sgd, then export
__global__ void sgd_finish_kernel(const float* __restrict__ W, int ldw, int d, const SgdState* __restrict__ state,
                                  int B, float* __restrict__ coef, double* __restrict__ intercept,
                                  int32_t* __restrict__ n_iter, double* __restrict__ t_out,
                                  int32_t* __restrict__ status) {
  const int col = blockIdx.x;
  if (col >= B) return;
  const float wf = (float)state[col].wscale;
  for (int k = threadIdx.x; k < d; k += blockDim.x) coef[(size_t)col * d + k] = __fmul_rn(W[(size_t)col * ldw + k], wf);
  if (threadIdx.x == 0) {
    intercept[col] = state[col].intercept;
    n_iter[col] = state[col].n_iter;
    t_out[col] = state[col].t;
    status[col] = state[col].status;
  }
}

static inline uint32_t xorshift_rand_r(uint32_t* seed) {   // SK/utils/_random.pxd:20-34
  if (*seed == 0) *seed = 1;
  *seed ^= (uint32_t)(*seed << 13);
  *seed ^= (uint32_t)(*seed >> 17);
  *seed ^= (uint32_t)(*seed << 5);
  return *seed % ((uint32_t)2147483647 + 1);
}

template <int LOSS>
static cudaError_t launch_epoch(int dpl, int grid, cudaStream_t st, const float* X, int ldx, int d,
                                const int32_t* ycls, const int32_t* order, const double* eta,
                                const float* cfac, int64_t n, const int32_t* active, int n_active,
                                const int32_t* col_pos, float* W, int ldw, SgdState* state, double alpha,
                                int fit_intercept, double tol, int nnc) {
  // SKDIST_B200_SGD_SPEC=0 falls back to the one-sample-at-a-time kernel (A/B timing)
  static const bool spec = !(getenv("SKDIST_B200_SGD_SPEC") && getenv("SKDIST_B200_SGD_SPEC")[0] == '0');
#define SGD_CASE(D)                                                                                       \
  case D:                                                                                                 \
    if (LOSS == SGD_HINGE && spec)                                                                        \
      sgd_epoch_spec_kernel<D><<<grid, 128, 0, st>>>(X, ldx, d, ycls, order, eta, cfac, n, active, n_active, \
                                                     col_pos, W, ldw, state, alpha, fit_intercept, tol, nnc); \
    else                                                                                                  \
      sgd_epoch_kernel<D, LOSS><<<grid, 128, 0, st>>>(X, ldx, d, ycls, order, eta, cfac, n, active, n_active, \
                                                      col_pos, W, ldw, state, alpha, fit_intercept, tol, nnc); \
    break;
  switch (dpl) {
    SGD_CASE(1) SGD_CASE(2) SGD_CASE(4) SGD_CASE(8) SGD_CASE(16) SGD_CASE(32)
    default: return cudaErrorInvalidValue;
  }
#undef SGD_CASE
  return cudaGetLastError();
}

int sgd_fit_batch(Ctx* c, int B, const int32_t* col_pos, int loss, double alpha, int fit_intercept,
                  int max_iter, double tol, int shuffle, uint32_t seed, int lr_type, double eta0,
                  double power_t, double optimal_init, int n_iter_no_change, float* coef_out,
                  double* intercept_out, int32_t* n_iter_out, double* t_out, int32_t* status_out) {
  const int64_t n = c->n;
  const int d = (int)c->d, ldx = (int)c->ldx;
  if (d > 1024) return fail(c, "sgd: device path supports d <= 1024");
  if (sgd_tc_supported(c, loss, shuffle))     // hinge: blocked-exact on the tensor cores (sgd_tc.cu), same results
    return sgd_fit_batch_tc(c, B, col_pos, alpha, fit_intercept, max_iter, tol, shuffle, seed, lr_type, eta0, power_t,
                            optimal_init, n_iter_no_change, coef_out, intercept_out, n_iter_out, t_out, status_out);
  int dpl = 1;
  while (dpl * 32 < d) dpl *= 2;
  const int ldw = dpl * 32;
  Scratch sx(c);
  float* W; SgdState* state; int32_t *order, *active, *dpos; double* eta; float* cfac;
  float* dcoef; double *dint, *dt; int32_t *dniter, *dstatus;
  SKD_CUDA(c, sx.alloc(&W, (size_t)B * ldw));
  SKD_CUDA(c, sx.alloc(&state, (size_t)B));
  SKD_CUDA(c, sx.alloc(&order, (size_t)n));
  SKD_CUDA(c, sx.alloc(&active, (size_t)B));
  SKD_CUDA(c, sx.alloc(&dpos, (size_t)B));
  SKD_CUDA(c, sx.alloc(&eta, (size_t)n));
  SKD_CUDA(c, sx.alloc(&cfac, (size_t)n));
  SKD_CUDA(c, sx.alloc(&dcoef, (size_t)B * d));
  SKD_CUDA(c, sx.alloc(&dint, (size_t)B));
  SKD_CUDA(c, sx.alloc(&dt, (size_t)B));
  SKD_CUDA(c, sx.alloc(&dniter, (size_t)B));
  SKD_CUDA(c, sx.alloc(&dstatus, (size_t)B));
  SKD_CUDA(c, cudaMemsetAsync(W, 0, (size_t)B * ldw * sizeof(float), c->stream));
  std::vector<SgdState> hs(B);
  for (auto& s : hs) { s.wscale = 1.0; s.sq_norm = 0.0; s.intercept = 0.0; s.best_objective = INFINITY; s.t = 1.0;
                       s.no_improve = 0; s.done = 0; s.n_iter = 0; s.status = 3; }
  SKD_CUDA(c, cudaMemcpyAsync(state, hs.data(), (size_t)B * sizeof(SgdState), cudaMemcpyHostToDevice, c->stream));
  SKD_CUDA(c, cudaMemcpyAsync(dpos, col_pos, (size_t)B * 4, cudaMemcpyHostToDevice, c->stream));
  std::vector<int32_t> hact(B), hord(n);
  for (int j = 0; j < B; ++j) hact[j] = j;
  for (int64_t i = 0; i < n; ++i) hord[i] = (int32_t)i;
  int n_active = B;
  const char* trace_env = getenv("SKDIST_B200_TRACE");
  const bool trace = trace_env && trace_env[0] == '2';
  for (int epoch = 0; epoch < max_iter && n_active > 0; ++epoch) {
    auto tw0 = std::chrono::steady_clock::now();
    if (shuffle) {   // Fisher-Yates with the SAME seed every epoch, applied to the evolving order
      uint32_t s = seed;
      for (int64_t i = 0; i < n - 1; ++i) {
        int64_t j = i + xorshift_rand_r(&s) % (uint32_t)(n - i);
        std::swap(hord[i], hord[j]);
      }
    }
    if (shuffle || epoch == 0)
      SKD_CUDA(c, cudaMemcpyAsync(order, hord.data(), (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
    SKD_CUDA(c, cudaMemcpyAsync(active, hact.data(), (size_t)n_active * 4, cudaMemcpyHostToDevice, c->stream));
    auto tw1 = std::chrono::steady_clock::now();
    const double t0 = 1.0 + (double)epoch * (double)n;
    sgd_schedule_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(n, t0, alpha, optimal_init, lr_type, eta0,
                                                                           power_t, eta, cfac);
    cudaError_t e = loss == SGD_HINGE
        ? launch_epoch<SGD_HINGE>(dpl, (n_active + 3) / 4, c->stream, c->X, ldx, d, c->ycls, order, eta, cfac, n, active,
                                  n_active, dpos, W, ldw, state, alpha, fit_intercept, tol, n_iter_no_change)
        : launch_epoch<SGD_LOG>(dpl, (n_active + 3) / 4, c->stream, c->X, ldx, d, c->ycls, order, eta, cfac, n, active,
                                n_active, dpos, W, ldw, state, alpha, fit_intercept, tol, n_iter_no_change);
    c->launches += 2;
    if (e != cudaSuccess) return fail(c, std::string("sgd epoch launch: ") + cudaGetErrorString(e));
    SKD_CUDA(c, cudaMemcpyAsync(hs.data(), state, (size_t)B * sizeof(SgdState), cudaMemcpyDeviceToHost, c->stream));
    SKD_CUDA(c, cudaStreamSynchronize(c->stream));
    c->h2d += n * 4; c->d2h += (int64_t)B * sizeof(SgdState);
    if (trace) {
      auto tw2 = std::chrono::steady_clock::now();
      fprintf(stderr, "[skd trace] sgd epoch %3d active %5d host shuffle %7.2f ms device %8.2f ms\n", epoch, n_active,
              std::chrono::duration<double, std::milli>(tw1 - tw0).count(),
              std::chrono::duration<double, std::milli>(tw2 - tw1).count());
    }
    n_active = 0;
    for (int j = 0; j < B; ++j)
      if (!hs[j].done) hact[n_active++] = j;
  }
  sgd_finish_kernel<<<B, 128, 0, c->stream>>>(W, ldw, d, state, B, dcoef, dint, dniter, dt, dstatus);
  c->launches += 1;
  SKD_CUDA(c, cudaGetLastError());
  SKD_CUDA(c, cudaMemcpyAsync(coef_out, dcoef, (size_t)B * d * 4, cudaMemcpyDeviceToHost, c->stream));
  SKD_CUDA(c, cudaMemcpyAsync(intercept_out, dint, (size_t)B * 8, cudaMemcpyDeviceToHost, c->stream));
  SKD_CUDA(c, cudaMemcpyAsync(n_iter_out, dniter, (size_t)B * 4, cudaMemcpyDeviceToHost, c->stream));
  SKD_CUDA(c, cudaMemcpyAsync(t_out, dt, (size_t)B * 8, cudaMemcpyDeviceToHost, c->stream));
  SKD_CUDA(c, cudaMemcpyAsync(status_out, dstatus, (size_t)B * 4, cudaMemcpyDeviceToHost, c->stream));
  SKD_CUDA(c, cudaStreamSynchronize(c->stream));
  c->d2h += (int64_t)B * (d * 4 + 24);
  return 0;
}

}  // namespace skd
