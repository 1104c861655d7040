// skd_internal.h -- context object and helpers shared by the .cu translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <utility>
#include <vector>

#include "lbfgs_core.h"

namespace skd {

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
};

// fp16-split copy of the staged X for the tensor-core path (logreg_tc.cu)
struct TcData {
  void* Xh = nullptr;          // [npad x dpad] fp16, X * xscale rounded to fp16
  void* Xl = nullptr;          // [npad x dpad] fp16, remainder
  uint32_t* rowmeta = nullptr; // [npad] (fold << 24) | class id ; fold 0xFF = padding row
  float* yreal_pad = nullptr;  // [npad] regression targets (zero padded), when staged
  int32_t* tilelist = nullptr; // [(n_lists) x n_tiles] tiles that contain at least one TRAINING row of fold f
  int32_t* tilecnt = nullptr;  // [n_lists]; list index f for fold f, n_lists-1 = every tile (no held-out fold)
  int n_lists = 0;
  int min_list_tiles = 0;      // shortest tile list (chunks without tiles need zeroed partials)
  float* rowsg = nullptr;      // [n_lists x npad] -y * 2^14 per row for positive class rowsg_pos; 0 = not a training row
  int32_t rowsg_pos = -1;
  bool rowsg_valid = false;
  float* xscale = nullptr;     // [dpad] power-of-two per-feature scale
  double* gscale = nullptr;    // [dpad] 1 / (xscale * 2^14): un-scales the gradient partials
  int dpad = 0;
  int64_t npad = 0;
  bool x_valid = false, meta_valid = false;
  CUtensorMap map_xh, map_xl;
};

// binned copy of the staged X for the forest builder (forest.cu)
struct ForestData {
  uint8_t* xbin = nullptr;   // [d][n] bin code of every value, feature-major
  float* binval = nullptr;   // [d][256] distinct values of each feature, ascending (+inf padded)
  uint8_t* xrow = nullptr;   // [n][dp] the same codes row-major, dp = d rounded up to 16 (forest_fast.cu)
  int dp = 0;
  bool well_separated = false;   // every feature's adjacent distinct values are more than 1e-7 apart
  std::vector<float> h_binval;   // host copy of binval (thresholds of compact node records are formed on the host)
  bool valid = false;
};

// host view of one finished tree (arrays of node_count entries; value is [node_count][n_classes]).
// compact != nullptr: the tree comes as compact 32-byte node records (forest_fast.cu) and the arrays
// are null; `binval` ([d][256] distinct feature values) lets the consumer form the thresholds.
struct SkdTreeView {
  const uint32_t* compact = nullptr;
  const float* binval = nullptr;
  int32_t node_count, max_depth, n_classes;
  const int32_t *left, *right, *feature, *n_node_samples;
  const uint8_t* missing_go_to_left;
  const double *threshold, *impurity, *weighted_n_node_samples, *value;
};
typedef void (*ForestSink)(void* arg, int tree_index, const SkdTreeView* view);

struct Ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  int sm_count = 148;
  int kernel_choice = 0;  // 0 auto, 1 simt, 2 tcgen05
  // staged data
  float* X = nullptr;       // [n x ldx] fp32 row-major
  int64_t n = 0, d = 0, ldx = 0;
  int64_t x_cap_rows = 0;   // allocated rows of X (>= n; the sliced staging pads to a multiple of the world size)
  int64_t pend_n = 0, pend_d = 0;   // shape announced by skd_stage_x_begin, committed by skd_stage_x_commit
  int32_t* ycls = nullptr;  // [n]
  float* yreal = nullptr;   // [n]
  int8_t* fold = nullptr;   // [n]; nullptr = no folds staged (points into fold_store otherwise)
  int8_t* fold_store = nullptr;
  std::vector<int32_t> h_ycls; // host copy of the class ids (training-set sizes of one-vs-one pair columns)
  std::vector<int8_t> h_fold;  // host copy of the fold ids (tile lists for fold-aware tile skipping)
  int32_t n_folds = 0;
  std::vector<int64_t> fold_count;  // rows per fold id
  TcData tc;
  ForestData forest;
  int64_t ycls_cap = 0, yreal_cap = 0, fold_cap = 0;   // allocated rows of the staged vectors (reused when large enough)
  int64_t vec_n = 0;        // row count the staged labels / targets / folds belong to (dropped when X changes it)
  // per-column feature masks staged for the next skd_logreg_fit_batch (skd_stage_column_masks)
  std::vector<uint8_t> h_fmask;
  int32_t fmask_cols = 0;
  // per-column row bit matrices staged for the next skd_logreg_fit_batch (skd_stage_row_bits):
  // label of row r in column j / row r trains column j; packed little-endian, rb_words 32-bit words per column
  std::vector<uint32_t> h_ybits, h_mbits;
  int32_t rb_cols = 0;
  int64_t rb_words = 0;
  // scratch pool: device blocks released by finished calls, reused by the next ones (Scratch below)
  std::vector<std::pair<void*, size_t>> pool_free;
  size_t pool_bytes = 0;
  // pinned bounce buffers for staging pageable host arrays (api.cu: stage_rows_h2d)
  std::vector<void*> pin_bufs;
  size_t pin_bytes = 0;
  double forest_kernel_ms = 0.0;            // device time of the tree-builder kernels of the last forest_fit (events)
  void* pin_tree[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // pinned ring for finished trees (forest.cu)
  size_t pin_tree_bytes = 0;
  // counters
  int64_t launches = 0, h2d = 0, d2h = 0;
  // optional per-evaluation timing (bench.py roofline): CUDA events on `stream` around every
  // evaluation launch of skd_logreg_fit_batch
  bool prof = false;
  double prof_eval_ms = 0.0;      // summed device time of the evaluation kernels
  double prof_eval_flops = 0.0;   // algorithmic FLOPs of those launches (4 * n_train * d per column)
  int64_t prof_eval_launches = 0; // evaluation launches (one per L-BFGS round)
  int64_t prof_rounds = 0;
  std::vector<cudaEvent_t> prof_events;
  cudaEvent_t timer[2] = {nullptr, nullptr};
};

// error plumbing ---------------------------------------------------------------------
void set_global_error(const std::string& s);
inline int fail(Ctx* c, const std::string& s) {
  if (c) c->err = s;
  set_global_error(s);
  return 1;
}

#define SKD_CUDA(ctx, call)                                                              \
  do {                                                                                   \
    cudaError_t _e = (call);                                                             \
    if (_e != cudaSuccess) {                                                             \
      char _b[512];                                                                      \
      snprintf(_b, sizeof(_b), "%s:%d: %s -> %s", __FILE__, __LINE__, #call,             \
               cudaGetErrorString(_e));                                                  \
      return skd::fail(ctx, _b);                                                         \
    }                                                                                    \
  } while (0)

// Simple RAII device allocation tied to a stream-ordered free at scope exit.
// Scratch device memory of one API call.  Blocks come from (and go back to) a per-context pool so
// that repeated calls (search steps, refits, scoring passes) do not pay cudaMalloc / cudaFree each
// time; every API call synchronises the stream before returning, so a pooled block is idle.
struct Scratch {
  Ctx* ctx;
  std::vector<std::pair<void*, size_t>> blocks;
  explicit Scratch(Ctx* c) : ctx(c) {}
  ~Scratch() {
    for (auto& b : blocks) { ctx->pool_free.push_back(b); ctx->pool_bytes += b.second; }
    // keep the pool bounded: drop the largest idle blocks beyond 6 GB
    while (ctx->pool_bytes > ((size_t)6 << 30) && !ctx->pool_free.empty()) {
      size_t bi = 0;
      for (size_t i = 1; i < ctx->pool_free.size(); ++i)
        if (ctx->pool_free[i].second > ctx->pool_free[bi].second) bi = i;
      cudaFree(ctx->pool_free[bi].first);
      ctx->pool_bytes -= ctx->pool_free[bi].second;
      ctx->pool_free.erase(ctx->pool_free.begin() + bi);
    }
  }
  template <class T>
  cudaError_t alloc(T** out, size_t count) {
    const size_t need = (count * sizeof(T) + 256 + 511) / 512 * 512;
    size_t best = (size_t)-1;
    for (size_t i = 0; i < ctx->pool_free.size(); ++i) {
      const size_t b = ctx->pool_free[i].second;
      if (b >= need && b <= 2 * need + ((size_t)1 << 20) &&
          (best == (size_t)-1 || b < ctx->pool_free[best].second))
        best = i;
    }
    if (best != (size_t)-1) {
      blocks.push_back(ctx->pool_free[best]);
      ctx->pool_bytes -= ctx->pool_free[best].second;
      *out = (T*)ctx->pool_free[best].first;
      ctx->pool_free.erase(ctx->pool_free.begin() + best);
      return cudaSuccess;
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, need);
    if (e != cudaSuccess && !ctx->pool_free.empty()) {   // out of memory: give the idle blocks back and retry
      for (auto& b : ctx->pool_free) cudaFree(b.first);
      ctx->pool_free.clear();
      ctx->pool_bytes = 0;
      cudaGetLastError();
      e = cudaMalloc(&p, need);
    }
    if (e == cudaSuccess) { blocks.push_back({p, need}); *out = (T*)p; }
    return e;
  }
};

// ---- logistic-regression batch solver pieces (logreg_simt.cu / lbfgs_dev.cu) -----------

// Per-slot metadata of the active evaluation batch.
struct SlotMeta {
  int32_t col;     // column id in the caller's batch
  int32_t fold;    // held-out fold id (-1: none)
  int32_t pos;     // positive class id
  int32_t pad;     // neg1: 0 = every other class is a negative (one-vs-rest); k + 1 = only rows of class
                   // `pos` or class k take part (one-vs-one pair)
};

// Workspace of one skd_logreg_fit_batch call (device pointers).
struct LogregWork {
  int32_t B = 0;        // columns in the batch
  int32_t dp = 0;       // d + 1 (intercept slot always present; pinned to 0 if !fit_intercept)
  int32_t nz = 0;       // max row chunks per evaluation (partials per slot)
  int64_t cap_sc = 0;   // capacity of the partial buffers in (chunk, slot) pairs
  int32_t ldg = 0;      // leading dimension of G (columns, padded)
  // per column (indexed by col)
  LbfgsScalars* sc = nullptr;
  double* vec = nullptr;       // per column block of (5 + 2m) * dp + 2m doubles
  size_t vec_stride = 0;
  double* l2 = nullptr;        // [B] l2 strength
  double* inv_n = nullptr;     // [B] 1 / n_train
  int32_t* col_fold = nullptr; // [B]
  int32_t* col_pos = nullptr;  // [B]
  int32_t* col_neg1 = nullptr; // [B] or nullptr (see SlotMeta::pad)
  uint8_t* fmask = nullptr;    // [B x d] or nullptr: 1 = feature takes part in the column's fit
  const uint32_t* ybits = nullptr;  // [B x rb_words] or nullptr: bit r of column j = its label of row r (instead of class id == pos)
  const uint32_t* mbits = nullptr;  // [B x rb_words] or nullptr: bit r of column j = row r trains the column
  int64_t rb_words = 0;
  int32_t* n_evals = nullptr;  // [B]
  // per slot (active batch)
  SlotMeta* slot = nullptr;    // [B]
  float* Wact = nullptr;       // [B x ldx] fp32 weights of the active slots, then bias[B]
  float* G = nullptr;          // [n x ldg] pointwise gradients (SIMT path)
  double* lossp = nullptr;     // [cap_sc]        indexed z * n_act + slot
  double* gsump = nullptr;     // [cap_sc]
  float* gradp = nullptr;      // [cap_sc x ldx]
  double* gradr = nullptr;     // [slots x ldw] partials reduced in chunk order (tensor-core path)
  int32_t* n_act = nullptr;    // device scalar
  // tensor-core path (logreg_tc.cu)
  bool use_tc = false;
  int32_t ldw = 0;             // leading dimension of gradp rows (ldx for SIMT, dpad for TC)
  const double* gscale = nullptr;  // per-feature un-scaling of gradp (TC) or nullptr
  void* Wh = nullptr;          // [slots_pad_cap x dpad] fp16
  void* Wl = nullptr;          // [slots_pad_cap x dpad] fp16
  void* sp = nullptr;          // [slots_pad_cap] TcSlotParam
  int32_t slots_pad_cap = 0;
  // fold-grouped slot layout (TC path): every group of 128 slots holds columns of ONE fold, padded
  // with col = -1 entries, so a group can skip the tiles made only of its held-out rows
  bool grouped = false;
  int32_t slot_cap = 0;        // slots incl. padding at the start of the solve
  int32_t* n_run = nullptr;    // device scalar: columns still running
  int32_t uni_pos = -1;        // >= 0: every column of the batch has this positive class (grouped layout only)
};

// forward (Z = X W^T, pointwise loss / gradient on training rows) + backward (G^T X)
int simt_eval(Ctx* c, LogregWork& w, int n_act, int* nz_used);
// scoring / decision kernels
int simt_score(Ctx* c, int B, const float* dW, const SlotMeta* dslot, int64_t* dcorrect,
               int64_t* dcount);
int simt_decision(Ctx* c, int B, const float* dW, float* dout);
int simt_r2(Ctx* c, int B, const float* dW, const SlotMeta* dslot, double* dsse, int64_t* dcount);
int sgd_fit_batch(Ctx* c, int B, const int32_t* col_pos, int loss, double alpha, int fit_intercept,
                  int max_iter, double tol, int shuffle, uint32_t seed, int lr_type, double eta0,
                  double power_t, double optimal_init, int n_iter_no_change, float* coef_out,
                  double* intercept_out, int32_t* n_iter_out, double* t_out, int32_t* status_out);
// 2-D fp16 row-major [rows x cols] TMA descriptor, box = [box_rows x 64 cols], 128B swizzle (logreg_tc.cu)
int tc_make_map_2d(Ctx* c, CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows);
bool sgd_tc_supported(const Ctx* c, int loss, int shuffle);
int sgd_fit_batch_tc(Ctx* c, int B, const int32_t* col_pos, double alpha, int fit_intercept, int max_iter, double tol,
                     int shuffle, uint32_t seed, int lr_type, double eta0, double power_t, double optimal_init,
                     int n_iter_no_change, float* coef_out, double* intercept_out, int32_t* n_iter_out, double* t_out,
                     int32_t* status_out);
void forest_free(Ctx* c);
int forest_fit(Ctx* c, int n_trees, const uint8_t* counts, const uint32_t* rand_states, int n_classes,
               int max_features, int max_depth, int min_samples_split, int min_samples_leaf,
               double min_weight_leaf, double min_impurity_decrease, int random_split, const double* h_yreal,
               ForestSink sink, void* sink_arg);
int predict_device(Ctx* c, const float* dX, int64_t m, int ldx, int d, int B, const float* dW, float* dout);
int forest_predict_device(Ctx* c, const float* dX, int64_t m, int ldx, int n_trees, const int64_t* d_off,
                          const void* d_node, const double* d_thr, const double* d_val, int C, double* d_out);
int ridge_fit_batch(Ctx* c, int B, const double* alpha, const int32_t* hold, int fit_intercept,
                    float* coef_out, int32_t* status_out);

// ---- multinomial logistic regression (logreg_multi.cu / lbfgs_dev.cu) -------------------
// One optimiser problem per candidate with K * (d + 1) variables, variable (k, j) at k * dp + j;
// the evaluation sees K slots per active candidate (slot = a * K + k for active index a).
struct MultiWork {
  int32_t B = 0, K = 0, dp = 0;  // candidates of this solve, classes, d + 1
  int32_t nz = 0;                // row chunks per evaluation (fixed by n alone)
  int64_t rpc = 0;               // rows per chunk
  int32_t ldg = 0;               // leading dimension of G
  LbfgsScalars* sc = nullptr;    // [B]
  double* vec = nullptr;         // [B x vec_stride]
  size_t vec_stride = 0;
  double* l2 = nullptr;          // [B]
  double* inv_n = nullptr;       // [B]
  int32_t* n_evals = nullptr;    // [B]
  SlotMeta* cand = nullptr;      // [B] active candidates: col = candidate, fold = held-out fold (-1 none)
  float* W = nullptr;            // [B*K x ldx] weights of the active slots, then bias [B*K]
  float* G = nullptr;            // [n x ldg] raw predictions, overwritten by the pointwise gradients
  double* lossp = nullptr;       // [nz x B]    indexed z * n_act + a
  double* gsump = nullptr;       // [nz x B*K]  indexed z * n_slots + slot
  float* gradp = nullptr;        // [nz x B*K x ldx]
  int32_t* n_act = nullptr;      // device scalar: active candidates
  uint8_t* fmask = nullptr;      // [B x d] or nullptr: 1 = feature takes part in the candidate's fit
};
int multi_lbfgs_init(Ctx* c, MultiWork& w, const int32_t* d_col_fold, double tol, int max_iter);
int multi_lbfgs_enqueue(Ctx* c, MultiWork& w, int n_act_in, int fit_intercept, int32_t* hist);
int multi_lbfgs_finish(Ctx* c, MultiWork& w, float* dcoef, int32_t* dniter, int32_t* dstatus, double* dloss);
int multi_fit(Ctx* c, int B, int K, const double* C, const int32_t* col_fold, int fit_intercept, double tol,
              int max_iter, const uint8_t* fmask /*[B x d] or nullptr*/, float* coef_out, int32_t* n_iter_out,
              int32_t* status_out, double* loss_out, int32_t* n_evals_out);
int multi_score(Ctx* c, int B, int K, const float* coef, const int32_t* col_fold, int64_t* conf_out);
int logloss_batch(Ctx* c, int B, int K, const float* coef, const int32_t* col_fold, const int32_t* col_pos,
                  double* loss_sum_out, int64_t* count_out);
// raw predictions / backward product of the fp32 CUDA-core path on arbitrary slot matrices (logreg_simt.cu)
int simt_raw_prediction(Ctx* c, int n_slots, const float* dW, const float* dbias, float* dout, int ldd);
int simt_backward(Ctx* c, const float* G, int ldg, int n_slots, int nz, int64_t rpc, float* gradp);

// ROC-AUC counts of linear binary classifiers (auc.cu): 2U, n_pos, n_neg per column
int auc_batch(Ctx* c, int B, const float* coef, const int32_t* col_fold, const int32_t* col_pos, int64_t* u2_out,
              int64_t* n_pos_out, int64_t* n_neg_out);
int simt_decision(Ctx* c, int B, const float* dW, float* dout);

// tensor-core evaluation (logreg_tc.cu)
bool tc_supported(const Ctx* c);
void tc_free(Ctx* c);
int tc_prepare(Ctx* c);
int tc_export(Ctx* c, LogregWork& w, int n_act_upper, const double* xin, int fit_intercept);
int tc_eval(Ctx* c, LogregWork& w, int n_act, int* nz_used);
int tc_score(Ctx* c, LogregWork& w, int n_act, int64_t* dcorrect, int64_t* dcount);
int tc_r2(Ctx* c, LogregWork& w, int n_act, double* dsse, int64_t* dcount);
size_t tc_slot_param_bytes();
int tc_partials_per_slot();

// device L-BFGS (lbfgs_dev.cu)
int lbfgs_dev_init(Ctx* c, LogregWork& w, int fit_intercept, double tol, int max_iter);
int lbfgs_dev_enqueue(Ctx* c, LogregWork& w, int n_act_in, int nz_used, int fit_intercept, int32_t* hist);
int lbfgs_dev_readback(Ctx* c, LogregWork& w, int* n_act_out, int* n_run_out);  // advance + compact + export (synchronises)
int lbfgs_dev_gather(Ctx* c, LogregWork& w, int n_act, int nz_used, int fit_intercept,
                     const double* dx, double* df, double* dg);
int lbfgs_dev_finish(Ctx* c, LogregWork& w, float* dcoef, int32_t* dniter, int32_t* dstatus,
                     double* dloss);

}  // namespace skd
