// api.cu -- the C-ABI declared in include/skdist_b200.h.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <thread>
#include <atomic>
#include <chrono>

#include "../../include/skdist_b200.h"
#include "skd_internal.h"

namespace skd {
static std::mutex g_err_mu;
static std::string g_err;
void set_global_error(const std::string& s) {
  std::lock_guard<std::mutex> l(g_err_mu);
  g_err = s;
}
}  // namespace skd

using namespace skd;

struct skd_ctx {
  Ctx c;
};

struct skd_lbfgs {
  LbfgsScalars s;
  LbfgsVectors v;
  std::vector<double> buf;
};

static inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

// Which evaluation kernel serves this batch (SIMT fp32 now; tcgen05 once it lands).
static int eval_dispatch(Ctx* c, LogregWork& w, int n_act, int* nz_used) {
  if (w.use_tc) return tc_eval(c, w, n_act, nz_used);
  return simt_eval(c, w, n_act, nz_used);
}

// 0 = auto: tensor cores whenever the staged shape is supported (d <= 256), else SIMT fp32.
static bool want_tc(const Ctx* c) {
  if (c->kernel_choice == 1) return false;
  if (c->kernel_choice == 2) return true;
  return tc_supported(c) && !getenv("SKDIST_B200_FORCE_SIMT");
}

// Decide the evaluation path for a batch and allocate its evaluation buffers.
static int alloc_eval_buffers(Ctx* c, Scratch& sx, LogregWork& w, int B, int slot_cap = 0) {
  const int64_t n = c->n, ldx = c->ldx;
  if (c->kernel_choice == 2 && !tc_supported(c))
    return fail(c, "tcgen05 path requested but the staged shape is unsupported (needs d <= 256)");
  w.use_tc = want_tc(c);
  if (slot_cap < B) slot_cap = B;
  w.slot_cap = slot_cap;
  // partial sums per slot: row chunks of the SIMT grid, or the fixed chunk count of the tensor-core kernel
  w.cap_sc = w.use_tc ? (int64_t)tc_partials_per_slot() * round_up(slot_cap, 128)
                      : (int64_t)4 * c->sm_count * 64 + slot_cap + 64;
  w.nz = 1024;
  SKD_CUDA(c, sx.alloc(&w.lossp, (size_t)w.cap_sc));
  SKD_CUDA(c, sx.alloc(&w.gsump, (size_t)w.cap_sc));
  if (w.use_tc) {
    if (tc_prepare(c)) return 1;
    w.ldw = c->tc.dpad;
    w.gscale = c->tc.gscale;
    w.slots_pad_cap = (int)round_up(slot_cap, 128);
    size_t wbytes = (size_t)w.slots_pad_cap * c->tc.dpad * 2;
    SKD_CUDA(c, sx.alloc((uint8_t**)&w.Wh, wbytes));
    SKD_CUDA(c, sx.alloc((uint8_t**)&w.Wl, wbytes));
    SKD_CUDA(c, sx.alloc((uint8_t**)&w.sp, (size_t)w.slots_pad_cap * tc_slot_param_bytes()));
    SKD_CUDA(c, cudaMemsetAsync(w.Wh, 0, wbytes, c->stream));
    SKD_CUDA(c, cudaMemsetAsync(w.Wl, 0, wbytes, c->stream));
    SKD_CUDA(c, sx.alloc(&w.gradp, (size_t)w.cap_sc * w.ldw));
    SKD_CUDA(c, sx.alloc(&w.gradr, (size_t)w.slots_pad_cap * w.ldw));
  } else {
    w.ldw = (int)ldx;
    w.gscale = nullptr;
    w.ldg = (int)round_up(B, 64);
    if ((double)n * w.ldg * 4.0 > 120e9)
      return fail(c, "skd_logreg_fit_batch: batch too large for one SIMT call (n * B * 4 bytes > 120 GB); split the batch");
    SKD_CUDA(c, sx.alloc(&w.Wact, (size_t)B * ldx + B));
    SKD_CUDA(c, sx.alloc(&w.G, (size_t)n * w.ldg));
    SKD_CUDA(c, sx.alloc(&w.gradp, (size_t)w.cap_sc * ldx));
  }
  return 0;
}

// Non-finite scan of the staged matrix: scikit-learn's estimators reject NaN / infinity in X
// (check_array -> "Input X contains NaN."), so the staging call does too -- at HBM speed, on the
// copy that is already on the device.
__global__ void finite_check_kernel(const float* __restrict__ X, int64_t total, int* __restrict__ flag) {
  int bad = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = X[i];
    bad |= !(fabsf(v) <= 3.402823466e38f);
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

extern "C" {

int skd_version(void) { return 100; }

int skd_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

const char* skd_last_error(skd_ctx* ctx) {
  if (ctx) return ctx->c.err.c_str();
  std::lock_guard<std::mutex> l(g_err_mu);
  static thread_local std::string copy;
  copy = g_err;
  return copy.c_str();
}

int skd_ctx_create(int device, skd_ctx** out) {
  if (!out) return fail(nullptr, "skd_ctx_create: out is NULL");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(nullptr, std::string("skd_ctx_create: no CUDA device available (") +
                             cudaGetErrorString(e) + "); this library has no CPU fallback");
  }
  if (device < 0 || device >= ndev) return fail(nullptr, "skd_ctx_create: bad device index");
  SKD_CUDA(nullptr, cudaSetDevice(device));
  cudaDeviceProp prop;
  SKD_CUDA(nullptr, cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    char b[256];
    snprintf(b, sizeof(b), "skd_ctx_create: device %d is sm_%d%d; this library is built for sm_100a only",
             device, prop.major, prop.minor);
    return fail(nullptr, b);
  }
  skd_ctx* h = new skd_ctx();
  h->c.device = device;
  h->c.sm_count = prop.multiProcessorCount;
  SKD_CUDA(nullptr, cudaStreamCreateWithFlags(&h->c.stream, cudaStreamNonBlocking));
  *out = h;
  return 0;
}

// SKDIST_B200_TRACE=1: host wall-clock per phase of a call (stream synchronised at each mark), to stderr
struct Trace {
  Ctx* c; const char* call; bool on; std::chrono::steady_clock::time_point t;
  Trace(Ctx* c_, const char* call_) : c(c_), call(call_) {
    const char* e = getenv("SKDIST_B200_TRACE");
    on = e && *e && *e != '0';
    t = std::chrono::steady_clock::now();
  }
  void mark(const char* what) {
    if (!on) return;
    cudaStreamSynchronize(c->stream);
    auto n = std::chrono::steady_clock::now();
    fprintf(stderr, "[skd trace] %s %-10s %9.3f ms\n", call, what, std::chrono::duration<double, std::milli>(n - t).count());
    t = n;
  }
};

static void free_staged(Ctx& c) {
  if (c.X) cudaFree(c.X);
  if (c.ycls) cudaFree(c.ycls);
  if (c.yreal) cudaFree(c.yreal);
  if (c.fold_store) cudaFree(c.fold_store);
  c.X = nullptr; c.ycls = nullptr; c.yreal = nullptr; c.fold = nullptr; c.fold_store = nullptr;
}

int skd_ctx_destroy(skd_ctx* ctx) {
  if (!ctx) return 0;
  cudaSetDevice(ctx->c.device);
  cudaStreamSynchronize(ctx->c.stream);
  free_staged(ctx->c);
  for (void* p : ctx->c.pin_bufs) cudaFreeHost(p);
  ctx->c.pin_bufs.clear();
  for (void* pb : ctx->c.pin_tree) if (pb) cudaFreeHost(pb);
  for (auto& b : ctx->c.pool_free) cudaFree(b.first);
  ctx->c.pool_free.clear();
  tc_free(&ctx->c);
  forest_free(&ctx->c);
  cudaStreamDestroy(ctx->c.stream);
  delete ctx;
  return 0;
}

// Host -> device copy of an [n x d] fp32 matrix (row pitch ldx_src) into dst (row pitch ldx).
// Pinned sources go straight to the copy engine.  Pageable sources (plain numpy arrays) would be
// bounced by the driver through one small internal buffer at a few GB/s; instead T host threads
// each copy their own row blocks into pinned bounce buffers (double-buffered per thread) and
// issue the DMA on their own stream, so memcpy and PCIe transfers of different blocks overlap.
static int stage_rows_h2d(Ctx* c, float* dst, int64_t ldx, const float* src, int64_t n, int64_t d,
                          int64_t ldx_src) {
  cudaPointerAttributes attr;
  bool pinned = cudaPointerGetAttributes(&attr, src) == cudaSuccess && attr.type == cudaMemoryTypeHost;
  cudaGetLastError();
  const size_t total = (size_t)n * d * sizeof(float);
  // rows wider than one bounce block (d * 4 > 8 MiB) cannot go through the threaded bounce: plain copy
  if (pinned || total < ((size_t)8 << 20) || (size_t)d * sizeof(float) > ((size_t)8 << 20)) {
    SKD_CUDA(c, cudaMemcpy2DAsync(dst, ldx * sizeof(float), src, ldx_src * sizeof(float), d * sizeof(float), n,
                                  cudaMemcpyHostToDevice, c->stream));
    SKD_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
  }
  const int T = 8, NB = 2;
  const size_t buf_bytes = (size_t)8 << 20;
  if (c->pin_bufs.size() != (size_t)T * NB || c->pin_bytes != buf_bytes) {
    for (void* p : c->pin_bufs) cudaFreeHost(p);
    c->pin_bufs.clear();
    for (int i = 0; i < T * NB; ++i) {
      void* p = nullptr;
      SKD_CUDA(c, cudaHostAlloc(&p, buf_bytes, cudaHostAllocDefault));
      c->pin_bufs.push_back(p);
    }
    c->pin_bytes = buf_bytes;
  }
  const int64_t rows_per_blk = std::max<int64_t>(1, (int64_t)(buf_bytes / ((size_t)d * sizeof(float))));
  const int64_t n_blk = (n + rows_per_blk - 1) / rows_per_blk;
  std::atomic<int> err_code{0};
  std::atomic<int64_t> next_blk{0};
  auto worker = [&](int tid) {
    if (cudaSetDevice(c->device) != cudaSuccess) { err_code = 1; return; }
    cudaStream_t st;
    cudaEvent_t ev[NB];
    if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) { err_code = 1; return; }
    for (int b = 0; b < NB; ++b) cudaEventCreateWithFlags(&ev[b], cudaEventDisableTiming);
    int used = 0;
    for (;;) {
      const int64_t blk = next_blk.fetch_add(1);
      if (blk >= n_blk || err_code) break;
      const int b = used % NB;
      if (used >= NB && cudaEventSynchronize(ev[b]) != cudaSuccess) { err_code = 2; break; }
      float* pb = (float*)c->pin_bufs[(size_t)tid * NB + b];
      const int64_t r0 = blk * rows_per_blk, r1 = std::min(n, r0 + rows_per_blk);
      if (ldx_src == d) {
        memcpy(pb, src + (size_t)r0 * ldx_src, (size_t)(r1 - r0) * d * sizeof(float));
      } else {
        for (int64_t r = r0; r < r1; ++r)
          memcpy(pb + (size_t)(r - r0) * d, src + (size_t)r * ldx_src, (size_t)d * sizeof(float));
      }
      if (cudaMemcpy2DAsync(dst + (size_t)r0 * ldx, ldx * sizeof(float), pb, d * sizeof(float), d * sizeof(float),
                            (size_t)(r1 - r0), cudaMemcpyHostToDevice, st) != cudaSuccess) { err_code = 3; break; }
      cudaEventRecord(ev[b], st);
      ++used;
    }
    if (cudaStreamSynchronize(st) != cudaSuccess) err_code = 4;
    for (int b = 0; b < NB; ++b) cudaEventDestroy(ev[b]);
    cudaStreamDestroy(st);
  };
  SKD_CUDA(c, cudaStreamSynchronize(c->stream));   // the memset of the padding (if any) is ordered before the copies
  std::vector<std::thread> th;
  for (int t = 0; t < T; ++t) th.emplace_back(worker, t);
  for (auto& t : th) t.join();
  if (err_code) {
    char b[128];
    snprintf(b, sizeof(b), "skd_stage_x: threaded host-to-device staging failed (step %d): %s", (int)err_code,
             cudaGetErrorString(cudaGetLastError()));
    return fail(c, b);
  }
  return 0;
}

// device buffer of the staged matrix: kept when the row pitch is unchanged and it is large enough
// (repeated fits on same-sized data); zeroed when the pitch pads the rows
static int ensure_x_buffer(Ctx* c, int64_t n, int64_t d, int64_t n_alloc) {
  const int64_t ldx = round_up(d, 16);
  if (c->X && !(c->ldx == ldx && c->x_cap_rows >= n_alloc && c->x_cap_rows <= n_alloc + n_alloc / 8 + 64)) {
    cudaFree(c->X); c->X = nullptr; c->n = 0; c->x_cap_rows = 0;
  }
  if (!c->X) {
    SKD_CUDA(c, cudaMalloc((void**)&c->X, (size_t)n_alloc * ldx * sizeof(float)));
    c->x_cap_rows = n_alloc;
    c->ldx = ldx;
  }
  if (ldx != d) SKD_CUDA(c, cudaMemsetAsync(c->X, 0, (size_t)c->x_cap_rows * ldx * sizeof(float), c->stream));
  (void)n;
  return 0;
}

// Labels, targets and fold ids describe the rows of one staged X: a matrix with another row count makes them
// stale (a later call must not read n rows out of a vector staged for fewer), so they are dropped with it.
static void drop_stale_row_vectors(Ctx* c, int64_t n_new) {
  if (c->vec_n == n_new) return;
  if (c->ycls) { cudaFree(c->ycls); c->ycls = nullptr; c->ycls_cap = 0; }
  if (c->yreal) { cudaFree(c->yreal); c->yreal = nullptr; c->yreal_cap = 0; }
  if (c->fold_store) { cudaFree(c->fold_store); c->fold_store = nullptr; c->fold_cap = 0; }
  c->fold = nullptr;
  c->n_folds = 0;
  c->fold_count.clear();
  c->h_fold.clear();
  c->h_ycls.clear();
  c->rb_cols = 0;
  c->tc.meta_valid = false;
  c->vec_n = n_new;
}

static int finite_check_staged(Ctx* c, int64_t n, int64_t ldx) {
  int* dflag;
  Scratch sx(c);
  SKD_CUDA(c, sx.alloc(&dflag, 1));
  SKD_CUDA(c, cudaMemsetAsync(dflag, 0, sizeof(int), c->stream));
  finite_check_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(c->X, n * ldx, dflag);
  c->launches += 1;
  int hflag = 0;
  SKD_CUDA(c, cudaMemcpyAsync(&hflag, dflag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  SKD_CUDA(c, cudaStreamSynchronize(c->stream));
  if (hflag) {
    cudaFree(c->X); c->X = nullptr; c->n = 0; c->x_cap_rows = 0;
    return fail(c, "Input X contains NaN or infinity.");
  }
  return 0;
}

static int stage_x_common(Ctx* c, const float* src, int64_t n, int64_t d, int64_t ldx_src,
                          cudaMemcpyKind kind) {
  if (!src || n <= 0 || d <= 0 || ldx_src < d) return fail(c, "skd_stage_x: bad arguments");
  SKD_CUDA(c, cudaSetDevice(c->device));
  Trace tr(c, "stage_x");
  int64_t ldx = round_up(d, 16);
  if (ensure_x_buffer(c, n, d, n)) return 1;
  tr.mark("alloc");
  if (kind == cudaMemcpyHostToDevice) {
    if (stage_rows_h2d(c, c->X, ldx, src, n, d, ldx_src)) return 1;
  } else {
    SKD_CUDA(c, cudaMemcpy2DAsync(c->X, ldx * sizeof(float), src, ldx_src * sizeof(float),
                                  d * sizeof(float), n, kind, c->stream));
    SKD_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  tr.mark("copy");
  if (finite_check_staged(c, n, ldx)) return 1;
  tr.mark("check");
  drop_stale_row_vectors(c, n);
  c->n = n; c->d = d; c->ldx = ldx;
  c->tc.x_valid = false;
  c->forest.valid = false;
  if (kind == cudaMemcpyHostToDevice) c->h2d += (int64_t)n * d * sizeof(float);
  return 0;
}

int skd_stage_x(skd_ctx* ctx, const float* X, int64_t n, int64_t d, int64_t ldx) {
  if (!ctx) return fail(nullptr, "skd_stage_x: ctx is NULL");
  return stage_x_common(&ctx->c, X, n, d, ldx, cudaMemcpyHostToDevice);
}

int skd_stage_x_device(skd_ctx* ctx, const float* dX, int64_t n, int64_t d, int64_t ldx) {
  if (!ctx) return fail(nullptr, "skd_stage_x_device: ctx is NULL");
  return stage_x_common(&ctx->c, dX, n, d, ldx, cudaMemcpyDeviceToDevice);
}

// Sliced staging for several ranks that all hold X on the host: every rank copies its own row slice
// (1/N of the matrix through its own PCIe link), the caller all-gathers the slices in place over
// NVLink (the buffer holds n_alloc >= n rows so that the slices can be equal-sized), then commits.
int skd_stage_x_begin(skd_ctx* ctx, int64_t n, int64_t d, int64_t n_alloc, const float** dX, int64_t* ldx_out) {
  if (!ctx) return fail(nullptr, "skd_stage_x_begin: ctx is NULL");
  Ctx* c = &ctx->c;
  if (n <= 0 || d <= 0 || n_alloc < n) return fail(c, "skd_stage_x_begin: bad arguments");
  SKD_CUDA(c, cudaSetDevice(c->device));
  c->n = 0;                              // nothing is staged until the commit
  if (ensure_x_buffer(c, n, d, n_alloc)) return 1;
  SKD_CUDA(c, cudaStreamSynchronize(c->stream));
  c->pend_n = n; c->pend_d = d;
  if (dX) *dX = c->X;
  if (ldx_out) *ldx_out = c->ldx;
  return 0;
}

int skd_stage_x_rows(skd_ctx* ctx, const float* X_rows, int64_t ld, int64_t row0, int64_t n_rows) {
  if (!ctx) return fail(nullptr, "skd_stage_x_rows: ctx is NULL");
  Ctx* c = &ctx->c;
  if (!c->X || c->pend_n <= 0) return fail(c, "skd_stage_x_rows: call skd_stage_x_begin first");
  if (n_rows == 0) return 0;
  if (!X_rows || row0 < 0 || n_rows < 0 || row0 + n_rows > c->pend_n || ld < c->pend_d)
    return fail(c, "skd_stage_x_rows: bad arguments");
  SKD_CUDA(c, cudaSetDevice(c->device));
  if (stage_rows_h2d(c, c->X + (size_t)row0 * c->ldx, c->ldx, X_rows, n_rows, c->pend_d, ld)) return 1;
  c->h2d += n_rows * c->pend_d * (int64_t)sizeof(float);
  return 0;
}

int skd_stage_x_commit(skd_ctx* ctx) {
  if (!ctx) return fail(nullptr, "skd_stage_x_commit: ctx is NULL");
  Ctx* c = &ctx->c;
  if (!c->X || c->pend_n <= 0) return fail(c, "skd_stage_x_commit: call skd_stage_x_begin first");
  SKD_CUDA(c, cudaSetDevice(c->device));
  const int64_t n = c->pend_n, d = c->pend_d;
  c->pend_n = 0;
  if (finite_check_staged(c, n, c->ldx)) return 1;
  drop_stale_row_vectors(c, n);
  c->n = n; c->d = d;
  c->tc.x_valid = false;
  c->forest.valid = false;
  return 0;
}

int skd_staged_x(skd_ctx* ctx, const float** dX, int64_t* n, int64_t* d, int64_t* ldx) {
  if (!ctx) return fail(nullptr, "skd_staged_x: ctx is NULL");
  Ctx* c = &ctx->c;
  if (!c->X) return fail(c, "skd_staged_x: nothing staged");
  if (dX) *dX = c->X;
  if (n) *n = c->n;
  if (d) *d = c->d;
  if (ldx) *ldx = c->ldx;
  return 0;
}

int skd_stage_labels(skd_ctx* ctx, const int32_t* y, int64_t n) {
  if (!ctx) return fail(nullptr, "skd_stage_labels: ctx is NULL");
  Ctx* c = &ctx->c;
  if (!y || n != c->n) return fail(c, "skd_stage_labels: n does not match the staged X");
  SKD_CUDA(c, cudaSetDevice(c->device));
  if (c->ycls && c->ycls_cap < n) { cudaFree(c->ycls); c->ycls = nullptr; }
  if (!c->ycls) { SKD_CUDA(c, cudaMalloc((void**)&c->ycls, (size_t)n * sizeof(int32_t))); c->ycls_cap = n; }
  SKD_CUDA(c, cudaMemcpyAsync(c->ycls, y, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  SKD_CUDA(c, cudaStreamSynchronize(c->stream));
  c->h_ycls.assign(y, y + n);
  c->h2d += n * (int64_t)sizeof(int32_t);
  c->tc.meta_valid = false;
  return 0;
}

int skd_stage_targets(skd_ctx* ctx, const float* y, int64_t n) {
  if (!ctx) return fail(nullptr, "skd_stage_targets: ctx is NULL");
  Ctx* c = &ctx->c;
  if (!y || n != c->n) return fail(c, "skd_stage_targets: n does not match the staged X");
  SKD_CUDA(c, cudaSetDevice(c->device));
  if (c->yreal && c->yreal_cap < n) { cudaFree(c->yreal); c->yreal = nullptr; }
  if (!c->yreal) { SKD_CUDA(c, cudaMalloc((void**)&c->yreal, (size_t)n * sizeof(float))); c->yreal_cap = n; }
  SKD_CUDA(c, cudaMemcpyAsync(c->yreal, y, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  SKD_CUDA(c, cudaStreamSynchronize(c->stream));
  c->h2d += n * (int64_t)sizeof(float);
  c->tc.meta_valid = false;
  return 0;
}

int skd_stage_folds(skd_ctx* ctx, const int8_t* fold_id, int64_t n, int32_t n_folds) {
  if (!ctx) return fail(nullptr, "skd_stage_folds: ctx is NULL");
  Ctx* c = &ctx->c;
  SKD_CUDA(c, cudaSetDevice(c->device));
  c->fold = nullptr;
  c->tc.meta_valid = false;
  c->n_folds = 0;
  c->fold_count.clear();
  c->h_fold.clear();
  if (!fold_id) return 0;  // cleared
  if (n != c->n || n_folds <= 0 || n_folds > 127) return fail(c, "skd_stage_folds: bad arguments");
  c->fold_count.assign(n_folds, 0);
  c->h_fold.assign(fold_id, fold_id + n);
  for (int64_t i = 0; i < n; ++i) {
    int f = fold_id[i];
    if (f < 0 || f >= n_folds) return fail(c, "skd_stage_folds: fold id out of range");
    c->fold_count[f] += 1;
  }
  if (c->fold_store && c->fold_cap < n) { cudaFree(c->fold_store); c->fold_store = nullptr; }
  if (!c->fold_store) { SKD_CUDA(c, cudaMalloc((void**)&c->fold_store, (size_t)n)); c->fold_cap = n; }
  c->fold = c->fold_store;
  SKD_CUDA(c, cudaMemcpyAsync(c->fold, fold_id, (size_t)n, cudaMemcpyHostToDevice, c->stream));
  SKD_CUDA(c, cudaStreamSynchronize(c->stream));
  c->n_folds = n_folds;
  c->h2d += n;
  return 0;
}

int skd_stage_column_masks(skd_ctx* ctx, int32_t B, const uint8_t* mask) {
  if (!ctx) return fail(nullptr, "skd_stage_column_masks: ctx is NULL");
  Ctx* c = &ctx->c;
  c->fmask_cols = 0;
  c->h_fmask.clear();
  if (!mask || B <= 0) return 0;   // cleared
  if (!c->X) return fail(c, "skd_stage_column_masks: stage X first");
  c->h_fmask.assign(mask, mask + (size_t)B * c->d);
  c->fmask_cols = B;
  return 0;
}

int skd_stage_row_bits(skd_ctx* ctx, int32_t B, const uint8_t* label_bits, const uint8_t* train_bits,
                       int64_t bytes_per_col) {
  if (!ctx) return fail(nullptr, "skd_stage_row_bits: ctx is NULL");
  Ctx* c = &ctx->c;
  c->rb_cols = 0; c->rb_words = 0;
  c->h_ybits.clear(); c->h_mbits.clear();
  if (B <= 0 || (!label_bits && !train_bits)) return 0;   // cleared
  if (!c->X) return fail(c, "skd_stage_row_bits: stage X first");
  if (bytes_per_col * 8 < c->n) return fail(c, "skd_stage_row_bits: fewer bits per column than staged rows");
  const int64_t words = (c->n + 63) / 64 * 2;              // whole 64-row tiles of the tensor-core path
  const int64_t use = std::min<int64_t>(bytes_per_col, (c->n + 7) / 8);
  auto pack = [&](const uint8_t* src, std::vector<uint32_t>& dst) {
    dst.assign((size_t)B * words, 0u);
    for (int j = 0; j < B; ++j) {
      memcpy(dst.data() + (size_t)j * words, src + (size_t)j * bytes_per_col, (size_t)use);
      if (c->n & 7) {   // bits beyond the last row must read as 0
        uint8_t* last = reinterpret_cast<uint8_t*>(dst.data() + (size_t)j * words) + (c->n >> 3);
        *last &= (uint8_t)((1u << (c->n & 7)) - 1u);
      }
    }
  };
  if (label_bits) pack(label_bits, c->h_ybits);
  if (train_bits) pack(train_bits, c->h_mbits);
  c->rb_cols = B;
  c->rb_words = words;
  return 0;
}

int skd_set_kernel(skd_ctx* ctx, int32_t which) {
  if (!ctx) return -1;
  int prev = ctx->c.kernel_choice;
  ctx->c.kernel_choice = which;
  return prev;
}

int skd_profile(skd_ctx* ctx, int32_t enable, double* eval_ms, double* eval_flops,
                int64_t* eval_launches, int64_t* rounds) {
  if (!ctx) return fail(nullptr, "skd_profile: ctx is NULL");
  Ctx* c = &ctx->c;
  if (eval_ms) *eval_ms = c->prof_eval_ms;
  if (eval_flops) *eval_flops = c->prof_eval_flops;
  if (eval_launches) *eval_launches = c->prof_eval_launches;
  if (rounds) *rounds = c->prof_rounds;
  if (enable >= 0) {
    c->prof = enable != 0;
    c->prof_eval_ms = 0.0; c->prof_eval_flops = 0.0; c->prof_eval_launches = 0; c->prof_rounds = 0;
  }
  return 0;
}

int skd_timer_start(skd_ctx* ctx) {
  if (!ctx) return fail(nullptr, "skd_timer_start: ctx is NULL");
  Ctx* c = &ctx->c;
  SKD_CUDA(c, cudaSetDevice(c->device));
  if (!c->timer[0]) {
    SKD_CUDA(c, cudaEventCreate(&c->timer[0]));
    SKD_CUDA(c, cudaEventCreate(&c->timer[1]));
  }
  SKD_CUDA(c, cudaStreamSynchronize(c->stream));
  SKD_CUDA(c, cudaEventRecord(c->timer[0], c->stream));
  return 0;
}

int skd_timer_stop(skd_ctx* ctx, double* ms_out) {
  if (!ctx) return fail(nullptr, "skd_timer_stop: ctx is NULL");
  Ctx* c = &ctx->c;
  if (!c->timer[0] || !ms_out) return fail(c, "skd_timer_stop: timer not started");
  SKD_CUDA(c, cudaEventRecord(c->timer[1], c->stream));
  SKD_CUDA(c, cudaEventSynchronize(c->timer[1]));
  float ms = 0.f;
  SKD_CUDA(c, cudaEventElapsedTime(&ms, c->timer[0], c->timer[1]));
  *ms_out = ms;
  return 0;
}

int skd_get_counters(skd_ctx* ctx, int64_t* launches, int64_t* h2d, int64_t* d2h) {
  if (!ctx) return fail(nullptr, "skd_get_counters: ctx is NULL");
  if (launches) *launches = ctx->c.launches;
  if (h2d) *h2d = ctx->c.h2d;
  if (d2h) *d2h = ctx->c.d2h;
  return 0;
}

int skd_logreg_fit_batch(skd_ctx* ctx, int32_t B, const double* C, const int32_t* col_fold,
                         const int32_t* col_pos, const int32_t* col_neg, int32_t fit_intercept, double tol,
                         int32_t max_iter, float* coef_out, int32_t* n_iter_out,
                         int32_t* status_out, double* loss_out, int32_t* n_evals_out,
                         double* gpu_seconds_out) {
  if (!ctx) return fail(nullptr, "skd_logreg_fit_batch: ctx is NULL");
  Ctx* c = &ctx->c;
  if (!c->X || !c->ycls) return fail(c, "skd_logreg_fit_batch: stage X and labels first");
  if (B <= 0 || !C || !col_fold || !col_pos || !coef_out || !n_iter_out || !status_out)
    return fail(c, "skd_logreg_fit_batch: bad arguments");
  // staged column masks are one-shot: whatever happens in this call, they do not outlive it
  struct MaskGuard {
    Ctx* c; std::vector<uint8_t> mask; int32_t cols;
    explicit MaskGuard(Ctx* c_) : c(c_), cols(c_->fmask_cols) { mask.swap(c_->h_fmask); c_->fmask_cols = 0; }
  } staged_masks(c);
  struct BitsGuard {     // staged row bit matrices are one-shot as well
    std::vector<uint32_t> y, m; int32_t cols; int64_t words;
    explicit BitsGuard(Ctx* c_) : cols(c_->rb_cols), words(c_->rb_words) { y.swap(c_->h_ybits); m.swap(c_->h_mbits); c_->rb_cols = 0; c_->rb_words = 0; }
  } staged_bits(c);
  if (staged_bits.cols > 0) {
    if (staged_bits.cols != B) return fail(c, "skd_logreg_fit_batch: staged row bit matrices do not match this batch");
    for (int j = 0; j < B; ++j)
      if (col_fold[j] >= 0 || (col_neg && col_neg[j] >= 0))
        return fail(c, "skd_logreg_fit_batch: row bit matrices cannot be combined with folds or pair columns");
  }
  if (max_iter < 1) return fail(c, "skd_logreg_fit_batch: max_iter must be >= 1");
  SKD_CUDA(c, cudaSetDevice(c->device));
  const int64_t n = c->n, d = c->d, ldx = c->ldx;
  const int dp = (int)d + 1, m = 10;

  // host-side per-column constants
  std::vector<double> l2(B), inv_n(B);
  std::vector<int64_t> pair_counts;
  int max_cls = -1;
  double mean_ntrain = 0.0;
  for (int j = 0; j < B; ++j) {
    int f = col_fold[j];
    int64_t ntrain = n;
    if (f >= 0) {
      if (!c->fold || f >= c->n_folds) return fail(c, "skd_logreg_fit_batch: col_fold refers to an unstaged fold");
      ntrain = n - c->fold_count[f];
    }
    if (col_neg && col_neg[j] >= 0) {   // pair column: only rows of class col_pos[j] or col_neg[j] train
      if ((int64_t)c->h_ycls.size() != n) return fail(c, "skd_logreg_fit_batch: labels not staged");
      if (col_neg[j] == col_pos[j]) return fail(c, "skd_logreg_fit_batch: col_neg equals col_pos");
      if (pair_counts.empty()) {         // rows per (class, fold) once per call
        for (int64_t i = 0; i < n; ++i) if (c->h_ycls[i] > max_cls) max_cls = c->h_ycls[i];
        pair_counts.assign((size_t)(max_cls + 1) * (c->n_folds + 1), 0);
        for (int64_t i = 0; i < n; ++i) {
          const int fi = c->h_fold.empty() ? 0 : (int)c->h_fold[i];
          if (c->h_ycls[i] >= 0) pair_counts[(size_t)c->h_ycls[i] * (c->n_folds + 1) + (c->h_fold.empty() ? 0 : fi)] += 1;
        }
      }
      ntrain = 0;
      for (int cls : {col_pos[j], col_neg[j]}) {
        if (cls < 0 || cls > max_cls) continue;
        for (int ff = 0; ff < (c->h_fold.empty() ? 1 : c->n_folds); ++ff)
          if (ff != f) ntrain += pair_counts[(size_t)cls * (c->n_folds + 1) + ff];
      }
    }
    if (staged_bits.cols > 0 && !staged_bits.m.empty()) {      // training rows of the column = set bits of its mask
      ntrain = 0;
      const uint32_t* mw = staged_bits.m.data() + (size_t)j * staged_bits.words;
      for (int64_t q = 0; q < staged_bits.words; ++q) ntrain += __builtin_popcount(mw[q]);
    }
    if (ntrain <= 0) return fail(c, "skd_logreg_fit_batch: empty training set");
    if (!(C[j] > 0.0)) return fail(c, "skd_logreg_fit_batch: C must be positive");
    l2[j] = 1.0 / (C[j] * (double)ntrain);  // SK/linear_model/_logistic.py:580
    inv_n[j] = 1.0 / (double)ntrain;
    mean_ntrain += (double)ntrain / B;
  }

  Trace tr(c, "logreg_fit");
  Scratch sx(c);
  LogregWork w;
  w.B = B; w.dp = dp;
  w.vec_stride = (size_t)(5 + 2 * m) * dp + 2 * m;
  // Fold-grouped slot layout for the tensor-core path: columns sorted by held-out fold, every fold
  // segment padded to a multiple of 128 slots (one MMA group = one fold -> tile skipping).
  std::vector<SlotMeta> hslots;
  const bool grouped = want_tc(c) && tc_supported(c);
  if (grouped) {
    std::vector<int> order(B);
    for (int j = 0; j < B; ++j) order[j] = j;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return col_fold[a] < col_fold[b]; });
    for (int i = 0; i < B;) {
      int f = col_fold[order[i]], i0 = i;
      if (f > 127) return fail(c, "skd_logreg_fit_batch: fold id above 127");
      for (; i < B && col_fold[order[i]] == f; ++i) {
        SlotMeta sm; sm.col = order[i]; sm.fold = f < 0 ? -1 : f; sm.pos = col_pos[order[i]];
        sm.pad = (col_neg && col_neg[order[i]] >= 0) ? col_neg[order[i]] + 1 : 0;
        hslots.push_back(sm);
      }
      (void)i0;
      while (hslots.size() % 128) { SlotMeta sm; sm.col = -1; sm.fold = f < 0 ? -1 : f; sm.pos = -1; sm.pad = 0; hslots.push_back(sm); }
    }
  }
  if (alloc_eval_buffers(c, sx, w, B, grouped ? (int)hslots.size() : B)) return 1;
  w.grouped = grouped && w.use_tc;
  if (w.grouped) {
    w.uni_pos = col_pos[0];
    for (int j = 1; j < B; ++j)
      if (col_pos[j] != col_pos[0]) { w.uni_pos = -1; break; }
    if (col_neg)
      for (int j = 0; j < B; ++j)
        if (col_neg[j] >= 0) { w.uni_pos = -1; break; }   // pair masks are per column: general epilogue
  }
  SKD_CUDA(c, sx.alloc(&w.sc, (size_t)B));
  SKD_CUDA(c, sx.alloc(&w.vec, (size_t)B * w.vec_stride));
  SKD_CUDA(c, sx.alloc(&w.l2, (size_t)B));
  SKD_CUDA(c, sx.alloc(&w.inv_n, (size_t)B));
  SKD_CUDA(c, sx.alloc(&w.col_fold, (size_t)B));
  SKD_CUDA(c, sx.alloc(&w.col_pos, (size_t)B));
  std::vector<int32_t> hneg1;
  if (col_neg) {
    hneg1.resize(B);
    for (int j = 0; j < B; ++j) hneg1[j] = col_neg[j] >= 0 ? col_neg[j] + 1 : 0;
    SKD_CUDA(c, sx.alloc(&w.col_neg1, (size_t)B));
  }
  SKD_CUDA(c, sx.alloc(&w.n_evals, (size_t)B));
  const bool use_fmask = staged_masks.cols > 0;
  if (use_fmask) {
    if (staged_masks.cols != B || (int64_t)staged_masks.mask.size() != (int64_t)B * d)
      return fail(c, "skd_logreg_fit_batch: staged column masks do not match this batch (B x d)");
    SKD_CUDA(c, sx.alloc(&w.fmask, (size_t)B * d));
  }
  if (staged_bits.cols > 0) {
    w.rb_words = staged_bits.words;
    if (!staged_bits.y.empty()) {
      uint32_t* dy;
      SKD_CUDA(c, sx.alloc(&dy, staged_bits.y.size()));
      SKD_CUDA(c, cudaMemcpyAsync(dy, staged_bits.y.data(), staged_bits.y.size() * 4, cudaMemcpyHostToDevice, c->stream));
      w.ybits = dy;
      c->h2d += (int64_t)staged_bits.y.size() * 4;
    }
    if (!staged_bits.m.empty()) {
      uint32_t* dm;
      SKD_CUDA(c, sx.alloc(&dm, staged_bits.m.size()));
      SKD_CUDA(c, cudaMemcpyAsync(dm, staged_bits.m.data(), staged_bits.m.size() * 4, cudaMemcpyHostToDevice, c->stream));
      w.mbits = dm;
      c->h2d += (int64_t)staged_bits.m.size() * 4;
    }
    SKD_CUDA(c, cudaStreamSynchronize(c->stream));
    w.uni_pos = -1;
  }
  SKD_CUDA(c, sx.alloc(&w.slot, (size_t)w.slot_cap));
  SKD_CUDA(c, sx.alloc(&w.n_act, 1));
  SKD_CUDA(c, sx.alloc(&w.n_run, 1));
  float* dcoef; int32_t *dniter, *dstatus; double* dloss;
  SKD_CUDA(c, sx.alloc(&dcoef, (size_t)B * dp));
  SKD_CUDA(c, sx.alloc(&dniter, (size_t)B));
  SKD_CUDA(c, sx.alloc(&dstatus, (size_t)B));
  SKD_CUDA(c, sx.alloc(&dloss, (size_t)B));

  tr.mark("alloc");
  cudaEvent_t e0, e1;
  SKD_CUDA(c, cudaEventCreate(&e0));
  SKD_CUDA(c, cudaEventCreate(&e1));
  SKD_CUDA(c, cudaEventRecord(e0, c->stream));
  SKD_CUDA(c, cudaMemcpyAsync(w.l2, l2.data(), B * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  SKD_CUDA(c, cudaMemcpyAsync(w.inv_n, inv_n.data(), B * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  SKD_CUDA(c, cudaMemcpyAsync(w.col_fold, col_fold, B * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  SKD_CUDA(c, cudaMemcpyAsync(w.col_pos, col_pos, B * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  if (col_neg) SKD_CUDA(c, cudaMemcpyAsync(w.col_neg1, hneg1.data(), B * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  if (use_fmask) {
    SKD_CUDA(c, cudaMemcpyAsync(w.fmask, staged_masks.mask.data(), (size_t)B * d, cudaMemcpyHostToDevice, c->stream));
    SKD_CUDA(c, cudaStreamSynchronize(c->stream));
    c->h2d += (int64_t)B * d;
  }
  c->h2d += (int64_t)B * 24;

  if (w.grouped) {
    int32_t ns = (int32_t)hslots.size();
    SKD_CUDA(c, cudaMemcpyAsync(w.slot, hslots.data(), hslots.size() * sizeof(SlotMeta), cudaMemcpyHostToDevice, c->stream));
    SKD_CUDA(c, cudaMemcpyAsync(w.n_act, &ns, sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  }
  if (lbfgs_dev_init(c, w, fit_intercept, tol, max_iter)) return 1;
  tr.mark("init");
  int n_act = w.grouped ? (int)hslots.size() : B;
  int n_run = B;
  const long max_rounds = (long)max_iter * 52 + 16;
  long rounds = 0;
  const long force_rounds = getenv("SKDIST_B200_FORCE_ROUNDS") ? atol(getenv("SKDIST_B200_FORCE_ROUNDS")) : 0;
  std::vector<double> round_flops;
  std::vector<int> round_act, round_run;
  size_t ev_used = 0;
  // Several optimiser rounds are enqueued per host round trip: the kernels read the live slot count
  // on the device, the host's n_act is only an upper bound that sizes the grids and strides.  Each
  // round records {slots, running} in `hist` so the profile below uses the true counts.
  const int rounds_per_sync = 4;
  int32_t* hist = nullptr;
  const long hist_cap = max_rounds + rounds_per_sync + 4;
  SKD_CUDA(c, sx.alloc(&hist, (size_t)2 * hist_cap));
  std::vector<int> ev_round;              // round index of every profiled evaluation
  while (n_run > 0) {
    for (int k = 0; k < rounds_per_sync; ++k) {
      int nz_used = 0;
      if (c->prof) {
        if (c->prof_events.size() < ev_used + 2) {
          cudaEvent_t a, b;
          SKD_CUDA(c, cudaEventCreate(&a));
          SKD_CUDA(c, cudaEventCreate(&b));
          c->prof_events.push_back(a);
          c->prof_events.push_back(b);
        }
        SKD_CUDA(c, cudaEventRecord(c->prof_events[ev_used], c->stream));
      }
      if (eval_dispatch(c, w, n_act, &nz_used)) return 1;
      if (c->prof) {
        SKD_CUDA(c, cudaEventRecord(c->prof_events[ev_used + 1], c->stream));
        ev_used += 2;
        ev_round.push_back((int)rounds);
      }
      if (force_rounds > 0) {   // timing experiments: repeat the evaluation of the initial point
        if (++rounds >= force_rounds) break;
        continue;
      }
      if (lbfgs_dev_enqueue(c, w, n_act, nz_used, fit_intercept, hist + 2 * rounds)) return 1;
      if (++rounds > max_rounds) return fail(c, "skd_logreg_fit_batch: round limit exceeded (internal error)");
    }
    if (force_rounds > 0) { if (rounds >= force_rounds) break; continue; }
    int n_next = 0, r_next = 0;
    if (lbfgs_dev_readback(c, w, &n_next, &r_next)) return 1;
    n_act = n_next;
    n_run = r_next;
  }
  // true per-round counts (columns evaluated in round r = running after round r - 1)
  std::vector<int32_t> hhist((size_t)2 * std::max<long>(rounds, 1), 0);
  if (force_rounds == 0 && rounds > 0)
    SKD_CUDA(c, cudaMemcpyAsync(hhist.data(), hist, (size_t)2 * rounds * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  SKD_CUDA(c, cudaStreamSynchronize(c->stream));
  long live_rounds = 0;
  for (size_t e = 0; e < ev_round.size(); ++e) {
    const int r = ev_round[e];
    const int running = (force_rounds > 0 || r == 0) ? B : hhist[2 * (r - 1) + 1];
    const int slots = (force_rounds > 0 || r == 0) ? (w.grouped ? (int)hslots.size() : B) : hhist[2 * (r - 1)];
    round_flops.push_back(4.0 * (double)d * (double)running * mean_ntrain);
    round_act.push_back(slots);
    round_run.push_back(running);
  }
  for (long r = 0; r < rounds; ++r)
    if (force_rounds > 0 || r == 0 || hhist[2 * (r - 1) + 1] > 0) ++live_rounds;
  tr.mark("rounds");
  if (c->prof) {
    SKD_CUDA(c, cudaStreamSynchronize(c->stream));
    for (size_t i = 0; i + 1 < ev_used; i += 2) {
      float ms = 0.f;
      SKD_CUDA(c, cudaEventElapsedTime(&ms, c->prof_events[i], c->prof_events[i + 1]));
      if (tr.on && getenv("SKDIST_B200_TRACE")[0] == '2')
        fprintf(stderr, "[skd trace] round %3d slots %5d running %5d eval %7.3f ms\n", (int)(i / 2), round_act[i / 2],
                round_run[i / 2], ms);
      if (round_run[i / 2] <= 0) continue;      // enqueued past convergence: the kernels returned at once
      c->prof_eval_ms += ms;
      c->prof_eval_flops += round_flops[i / 2];
      c->prof_eval_launches += 1;
    }
    c->prof_rounds += live_rounds;
  }
  if (lbfgs_dev_finish(c, w, dcoef, dniter, dstatus, dloss)) return 1;
  SKD_CUDA(c, cudaMemcpyAsync(coef_out, dcoef, (size_t)B * dp * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  SKD_CUDA(c, cudaMemcpyAsync(n_iter_out, dniter, (size_t)B * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  SKD_CUDA(c, cudaMemcpyAsync(status_out, dstatus, (size_t)B * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  if (loss_out)
    SKD_CUDA(c, cudaMemcpyAsync(loss_out, dloss, (size_t)B * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (n_evals_out)
    SKD_CUDA(c, cudaMemcpyAsync(n_evals_out, w.n_evals, (size_t)B * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  SKD_CUDA(c, cudaEventRecord(e1, c->stream));
  SKD_CUDA(c, cudaStreamSynchronize(c->stream));
  c->d2h += (int64_t)B * (dp * 4 + 20);
  float ms = 0.f;
  SKD_CUDA(c, cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (gpu_seconds_out) *gpu_seconds_out = ms * 1e-3;
  tr.mark("finish");
  return 0;
}

int skd_logreg_loss_grad(skd_ctx* ctx, int32_t B, const double* w_in, const double* C,
                         const int32_t* col_fold, const int32_t* col_pos, int32_t fit_intercept,
                         double* loss_out, double* grad_out) {
  if (!ctx) return fail(nullptr, "skd_logreg_loss_grad: ctx is NULL");
  Ctx* c = &ctx->c;
  if (!c->X || !c->ycls) return fail(c, "skd_logreg_loss_grad: stage X and labels first");
  if (B <= 0 || !w_in || !C || !col_fold || !col_pos || !loss_out || !grad_out)
    return fail(c, "skd_logreg_loss_grad: bad arguments");
  SKD_CUDA(c, cudaSetDevice(c->device));
  const int64_t n = c->n, d = c->d, ldx = c->ldx;
  const int dp = (int)d + 1;
  std::vector<double> l2(B), inv_n(B);
  std::vector<SlotMeta> hs(B);
  std::vector<float> hw((size_t)B * ldx + B, 0.f);
  for (int j = 0; j < B; ++j) {
    int f = col_fold[j];
    int64_t ntrain = n;
    if (f >= 0) {
      if (!c->fold || f >= c->n_folds) return fail(c, "skd_logreg_loss_grad: unstaged fold");
      ntrain = n - c->fold_count[f];
    }
    l2[j] = 1.0 / (C[j] * (double)ntrain);
    inv_n[j] = 1.0 / (double)ntrain;
    hs[j].col = j; hs[j].fold = f; hs[j].pos = col_pos[j]; hs[j].pad = 0;
    for (int k = 0; k < d; ++k) hw[(size_t)j * ldx + k] = (float)w_in[(size_t)j * dp + k];
    hw[(size_t)B * ldx + j] = fit_intercept ? (float)w_in[(size_t)j * dp + d] : 0.f;
  }
  Scratch sx(c);
  LogregWork w;
  w.B = B; w.dp = dp;
  if (alloc_eval_buffers(c, sx, w, B)) return 1;
  SKD_CUDA(c, sx.alloc(&w.l2, (size_t)B));
  SKD_CUDA(c, sx.alloc(&w.inv_n, (size_t)B));
  SKD_CUDA(c, sx.alloc(&w.slot, (size_t)B));
  SKD_CUDA(c, sx.alloc(&w.n_act, 1));
  double *dx, *df, *dg;
  SKD_CUDA(c, sx.alloc(&dx, (size_t)B * dp));
  SKD_CUDA(c, sx.alloc(&df, (size_t)B));
  SKD_CUDA(c, sx.alloc(&dg, (size_t)B * dp));
  SKD_CUDA(c, cudaMemcpyAsync(w.l2, l2.data(), B * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  SKD_CUDA(c, cudaMemcpyAsync(w.inv_n, inv_n.data(), B * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  SKD_CUDA(c, cudaMemcpyAsync(w.slot, hs.data(), B * sizeof(SlotMeta), cudaMemcpyHostToDevice, c->stream));
  SKD_CUDA(c, cudaMemcpyAsync(dx, w_in, (size_t)B * dp * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  if (w.use_tc) {
    int32_t nb = B;
    SKD_CUDA(c, cudaMemcpyAsync(w.n_act, &nb, sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    if (tc_export(c, w, B, dx, fit_intercept)) return 1;
  } else {
    SKD_CUDA(c, cudaMemcpyAsync(w.Wact, hw.data(), hw.size() * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  }
  int nz_used = 0;
  if (eval_dispatch(c, w, B, &nz_used)) return 1;
  if (lbfgs_dev_gather(c, w, B, nz_used, fit_intercept, dx, df, dg)) return 1;
  SKD_CUDA(c, cudaMemcpyAsync(loss_out, df, B * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  SKD_CUDA(c, cudaMemcpyAsync(grad_out, dg, (size_t)B * dp * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  SKD_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}

// Pack caller coefficients [B x (d+1)] into the kernel layout: weights [B x ldx] + bias [B].
static int pack_coef(Ctx* c, Scratch& sx, int B, const float* coef, float** dW) {
  const int64_t d = c->d, ldx = c->ldx;
  std::vector<float> h((size_t)B * ldx + B, 0.f);
  for (int j = 0; j < B; ++j) {
    memcpy(&h[(size_t)j * ldx], coef + (size_t)j * (d + 1), d * sizeof(float));
    h[(size_t)B * ldx + j] = coef[(size_t)j * (d + 1) + d];
  }
  SKD_CUDA(c, sx.alloc(dW, h.size()));
  SKD_CUDA(c, cudaMemcpyAsync(*dW, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  SKD_CUDA(c, cudaStreamSynchronize(c->stream));
  c->h2d += (int64_t)h.size() * 4;
  return 0;
}

int skd_linear_score_batch(skd_ctx* ctx, int32_t B, const float* coef, const int32_t* col_fold,
                           const int32_t* col_pos, int64_t* correct_out, int64_t* count_out) {
  if (!ctx) return fail(nullptr, "skd_linear_score_batch: ctx is NULL");
  Ctx* c = &ctx->c;
  if (!c->X || !c->ycls) return fail(c, "skd_linear_score_batch: stage X and labels first");
  if (B <= 0 || !coef || !col_fold || !col_pos || !correct_out || !count_out)
    return fail(c, "skd_linear_score_batch: bad arguments");
  SKD_CUDA(c, cudaSetDevice(c->device));
  Trace tr(c, "score");
  Scratch sx(c);
  std::vector<SlotMeta> hs(B);
  for (int j = 0; j < B; ++j) {
    int f = col_fold[j] >= 0 ? col_fold[j] : (col_fold[j] <= -3 ? -3 - col_fold[j] : -1);
    if (col_fold[j] == -1) return fail(c, "skd_linear_score_batch: col_fold -1 is not a scoring code");
    if (f >= 0 && (!c->fold || f >= c->n_folds))
      return fail(c, "skd_linear_score_batch: col_fold refers to an unstaged fold");
    hs[j].col = j; hs[j].fold = col_fold[j]; hs[j].pos = col_pos[j]; hs[j].pad = 0;
  }
  SlotMeta* dslot; int64_t *dcorrect, *dcount;
  SKD_CUDA(c, sx.alloc(&dslot, (size_t)B));
  SKD_CUDA(c, sx.alloc(&dcorrect, (size_t)B));
  SKD_CUDA(c, sx.alloc(&dcount, (size_t)B));
  SKD_CUDA(c, cudaMemcpyAsync(dslot, hs.data(), B * sizeof(SlotMeta), cudaMemcpyHostToDevice, c->stream));
  SKD_CUDA(c, cudaMemsetAsync(dcorrect, 0, B * sizeof(int64_t), c->stream));
  SKD_CUDA(c, cudaMemsetAsync(dcount, 0, B * sizeof(int64_t), c->stream));
  if (c->kernel_choice == 2 && !tc_supported(c))
    return fail(c, "tcgen05 path requested but the staged shape is unsupported (needs d <= 256)");
  if (want_tc(c)) {
    // tensor-core GEMM1-only pass with a counting epilogue (logreg_tc.cu, TC_SCORE)
    if (tc_prepare(c)) return 1;
    LogregWork w;
    w.B = B; w.dp = (int)c->d + 1; w.use_tc = true; w.ldw = c->tc.dpad;
    w.slot = dslot;
    w.slots_pad_cap = (int)round_up(B, 128);
    size_t wbytes = (size_t)w.slots_pad_cap * c->tc.dpad * 2;
    SKD_CUDA(c, sx.alloc((uint8_t**)&w.Wh, wbytes));
    SKD_CUDA(c, sx.alloc((uint8_t**)&w.Wl, wbytes));
    SKD_CUDA(c, sx.alloc((uint8_t**)&w.sp, (size_t)w.slots_pad_cap * tc_slot_param_bytes()));
    SKD_CUDA(c, sx.alloc(&w.n_act, 1));
    SKD_CUDA(c, cudaMemsetAsync(w.Wh, 0, wbytes, c->stream));
    SKD_CUDA(c, cudaMemsetAsync(w.Wl, 0, wbytes, c->stream));
    std::vector<double> hx((size_t)B * w.dp);
    for (size_t i = 0; i < hx.size(); ++i) hx[i] = (double)coef[i];
    double* dx;
    SKD_CUDA(c, sx.alloc(&dx, hx.size()));
    int32_t nb = B;
    SKD_CUDA(c, cudaMemcpyAsync(dx, hx.data(), hx.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    SKD_CUDA(c, cudaMemcpyAsync(w.n_act, &nb, sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    SKD_CUDA(c, cudaStreamSynchronize(c->stream));
    c->h2d += (int64_t)hx.size() * 8;
    tr.mark("setup");
    if (tc_export(c, w, B, dx, 1)) return 1;
    if (tc_score(c, w, B, dcorrect, dcount)) return 1;
    tr.mark("kernel");
  } else {
    float* dW;
    if (pack_coef(c, sx, B, coef, &dW)) return 1;
    if (simt_score(c, B, dW, dslot, dcorrect, dcount)) return 1;
  }
  SKD_CUDA(c, cudaMemcpyAsync(correct_out, dcorrect, B * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
  SKD_CUDA(c, cudaMemcpyAsync(count_out, dcount, B * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
  SKD_CUDA(c, cudaStreamSynchronize(c->stream));
  c->d2h += (int64_t)B * 16;
  return 0;
}

int skd_logreg_multinomial_fit_batch(skd_ctx* ctx, int32_t B, int32_t n_classes, const double* C,
                                     const int32_t* col_fold, int32_t fit_intercept, double tol, int32_t max_iter,
                                     float* coef_out, int32_t* n_iter_out, int32_t* status_out, double* loss_out,
                                     int32_t* n_evals_out, double* gpu_seconds_out) {
  if (!ctx) return fail(nullptr, "skd_logreg_multinomial_fit_batch: ctx is NULL");
  Ctx* c = &ctx->c;
  if (!c->X || !c->ycls) return fail(c, "skd_logreg_multinomial_fit_batch: stage X and labels first");
  if (B <= 0 || n_classes < 2 || !C || !col_fold || !coef_out || !n_iter_out || !status_out)
    return fail(c, "skd_logreg_multinomial_fit_batch: bad arguments");
  if (max_iter < 1) return fail(c, "skd_logreg_multinomial_fit_batch: max_iter must be >= 1");
  for (int j = 0; j < B; ++j) {
    if (col_fold[j] >= 0 && (!c->fold || col_fold[j] >= c->n_folds))
      return fail(c, "skd_logreg_multinomial_fit_batch: col_fold refers to an unstaged fold");
    if (!(C[j] > 0.0)) return fail(c, "skd_logreg_multinomial_fit_batch: C must be positive");
  }
  // staged column masks are one-shot: whatever happens in this call, they do not outlive it
  std::vector<uint8_t> masks;
  masks.swap(c->h_fmask);
  const int32_t mask_cols = c->fmask_cols;
  c->fmask_cols = 0;
  if (mask_cols > 0 && (mask_cols != B || (int64_t)masks.size() != (int64_t)B * c->d))
    return fail(c, "skd_logreg_multinomial_fit_batch: staged column masks do not match the batch");
  SKD_CUDA(c, cudaSetDevice(c->device));
  Trace tr(c, "multinomial_fit");
  cudaEvent_t e0, e1;
  SKD_CUDA(c, cudaEventCreate(&e0));
  SKD_CUDA(c, cudaEventCreate(&e1));
  SKD_CUDA(c, cudaEventRecord(e0, c->stream));
  const int rc = multi_fit(c, B, n_classes, C, col_fold, fit_intercept, tol, max_iter, mask_cols > 0 ? masks.data() : nullptr,
                           coef_out, n_iter_out, status_out, loss_out, n_evals_out);
  float ms = 0.f;
  if (!rc) {
    cudaEventRecord(e1, c->stream);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (gpu_seconds_out) *gpu_seconds_out = ms * 1e-3;
  return rc;
}

static int multinomial_check(Ctx* c, const char* who, int32_t B, int32_t n_classes, const float* coef,
                             const int32_t* col_fold) {
  if (!c->X || !c->ycls) return fail(c, std::string(who) + ": stage X and labels first");
  if (B <= 0 || n_classes < 2 || !coef || !col_fold) return fail(c, std::string(who) + ": bad arguments");
  for (int j = 0; j < B; ++j) {
    const int f = col_fold[j] >= 0 ? col_fold[j] : (col_fold[j] <= -3 ? -3 - col_fold[j] : -1);
    if (col_fold[j] == -1) return fail(c, std::string(who) + ": col_fold -1 is not a scoring code");
    if (f >= 0 && (!c->fold || f >= c->n_folds))
      return fail(c, std::string(who) + ": col_fold refers to an unstaged fold");
  }
  return 0;
}

int skd_multinomial_confusion_batch(skd_ctx* ctx, int32_t B, int32_t n_classes, const float* coef,
                                    const int32_t* col_fold, int64_t* confusion_out) {
  if (!ctx) return fail(nullptr, "skd_multinomial_confusion_batch: ctx is NULL");
  Ctx* c = &ctx->c;
  if (multinomial_check(c, "skd_multinomial_confusion_batch", B, n_classes, coef, col_fold)) return 1;
  if (!confusion_out) return fail(c, "skd_multinomial_confusion_batch: bad arguments");
  SKD_CUDA(c, cudaSetDevice(c->device));
  Trace tr(c, "multinomial_confusion");
  return multi_score(c, B, n_classes, coef, col_fold, confusion_out);
}

int skd_multinomial_score_batch(skd_ctx* ctx, int32_t B, int32_t n_classes, const float* coef,
                                const int32_t* col_fold, int64_t* correct_out, int64_t* count_out) {
  if (!ctx) return fail(nullptr, "skd_multinomial_score_batch: ctx is NULL");
  Ctx* c = &ctx->c;
  if (multinomial_check(c, "skd_multinomial_score_batch", B, n_classes, coef, col_fold)) return 1;
  if (!correct_out || !count_out) return fail(c, "skd_multinomial_score_batch: bad arguments");
  SKD_CUDA(c, cudaSetDevice(c->device));
  Trace tr(c, "multinomial_score");
  const size_t KK = (size_t)n_classes * n_classes;
  std::vector<int64_t> conf((size_t)B * KK);
  if (multi_score(c, B, n_classes, coef, col_fold, conf.data())) return 1;
  for (int j = 0; j < B; ++j) {
    int64_t tot = 0, diag = 0;
    for (int a = 0; a < n_classes; ++a)
      for (int b = 0; b < n_classes; ++b) {
        const int64_t v = conf[(size_t)j * KK + (size_t)a * n_classes + b];
        tot += v;
        if (a == b) diag += v;
      }
    correct_out[j] = diag;
    count_out[j] = tot;
  }
  return 0;
}

int skd_linear_auc_batch(skd_ctx* ctx, int32_t B, const float* coef, const int32_t* col_fold,
                         const int32_t* col_pos, int64_t* u2_out, int64_t* n_pos_out, int64_t* n_neg_out) {
  if (!ctx) return fail(nullptr, "skd_linear_auc_batch: ctx is NULL");
  Ctx* c = &ctx->c;
  if (!c->X || !c->ycls) return fail(c, "skd_linear_auc_batch: stage X and labels first");
  if (B <= 0 || !coef || !col_fold || !col_pos || !u2_out || !n_pos_out || !n_neg_out)
    return fail(c, "skd_linear_auc_batch: bad arguments");
  for (int j = 0; j < B; ++j) {
    const int f = col_fold[j] >= 0 ? col_fold[j] : (col_fold[j] <= -3 ? -3 - col_fold[j] : -1);
    if (col_fold[j] == -1) return fail(c, "skd_linear_auc_batch: col_fold -1 is not a scoring code");
    if (f >= 0 && (!c->fold || f >= c->n_folds))
      return fail(c, "skd_linear_auc_batch: col_fold refers to an unstaged fold");
  }
  SKD_CUDA(c, cudaSetDevice(c->device));
  Trace tr(c, "auc");
  return auc_batch(c, B, coef, col_fold, col_pos, u2_out, n_pos_out, n_neg_out);
}

int skd_linear_logloss_batch(skd_ctx* ctx, int32_t B, int32_t n_classes, const float* coef, const int32_t* col_fold,
                             const int32_t* col_pos, double* loss_sum_out, int64_t* count_out) {
  if (!ctx) return fail(nullptr, "skd_linear_logloss_batch: ctx is NULL");
  Ctx* c = &ctx->c;
  if (n_classes == 2) return fail(c, "skd_linear_logloss_batch: n_classes is 1 (binary columns) or > 2");
  if (multinomial_check(c, "skd_linear_logloss_batch", B, n_classes == 1 ? 2 : n_classes, coef, col_fold)) return 1;
  if (!loss_sum_out || !count_out || (n_classes == 1 && !col_pos))
    return fail(c, "skd_linear_logloss_batch: bad arguments");
  SKD_CUDA(c, cudaSetDevice(c->device));
  Trace tr(c, "logloss");
  return logloss_batch(c, B, n_classes, coef, col_fold, col_pos, loss_sum_out, count_out);
}

int skd_ridge_fit_batch(skd_ctx* ctx, int32_t B, const double* alpha, const int32_t* col_fold,
                        int32_t fit_intercept, float* coef_out, int32_t* status_out,
                        double* gpu_seconds_out) {
  if (!ctx) return fail(nullptr, "skd_ridge_fit_batch: ctx is NULL");
  Ctx* c = &ctx->c;
  if (!c->X || !c->yreal) return fail(c, "skd_ridge_fit_batch: stage X and targets first");
  if (B <= 0 || !alpha || !col_fold || !coef_out || !status_out) return fail(c, "skd_ridge_fit_batch: bad arguments");
  SKD_CUDA(c, cudaSetDevice(c->device));
  const int n_folds = c->fold ? c->n_folds : 1;
  std::vector<int32_t> hold(B);
  for (int j = 0; j < B; ++j) {
    if (!(alpha[j] >= 0.0)) return fail(c, "skd_ridge_fit_batch: alpha must be non-negative");
    if (col_fold[j] >= 0) {
      if (!c->fold || col_fold[j] >= c->n_folds) return fail(c, "skd_ridge_fit_batch: col_fold refers to an unstaged fold");
      hold[j] = col_fold[j];
    } else {
      hold[j] = n_folds;   // hold out nothing
    }
  }
  cudaEvent_t e0, e1;
  SKD_CUDA(c, cudaEventCreate(&e0));
  SKD_CUDA(c, cudaEventCreate(&e1));
  SKD_CUDA(c, cudaEventRecord(e0, c->stream));
  int rc = ridge_fit_batch(c, B, alpha, hold.data(), fit_intercept, coef_out, status_out);
  float ms = 0.f;
  if (!rc) {
    cudaEventRecord(e1, c->stream);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (gpu_seconds_out) *gpu_seconds_out = ms * 1e-3;
  return rc;
}

int skd_linear_r2_batch(skd_ctx* ctx, int32_t B, const float* coef, const int32_t* col_fold,
                        double* sse_out, int64_t* count_out) {
  if (!ctx) return fail(nullptr, "skd_linear_r2_batch: ctx is NULL");
  Ctx* c = &ctx->c;
  if (!c->X || !c->yreal) return fail(c, "skd_linear_r2_batch: stage X and targets first");
  if (B <= 0 || !coef || !col_fold || !sse_out || !count_out) return fail(c, "skd_linear_r2_batch: bad arguments");
  SKD_CUDA(c, cudaSetDevice(c->device));
  Scratch sx(c);
  std::vector<SlotMeta> hs(B);
  for (int j = 0; j < B; ++j) {
    int f = col_fold[j] >= 0 ? col_fold[j] : (col_fold[j] <= -3 ? -3 - col_fold[j] : -1);
    if (col_fold[j] == -1) return fail(c, "skd_linear_r2_batch: col_fold -1 is not a scoring code");
    if (f >= 0 && (!c->fold || f >= c->n_folds)) return fail(c, "skd_linear_r2_batch: col_fold refers to an unstaged fold");
    hs[j].col = j; hs[j].fold = col_fold[j]; hs[j].pos = 0; hs[j].pad = 0;
  }
  SlotMeta* dslot; double* dsse; int64_t* dcount;
  SKD_CUDA(c, sx.alloc(&dslot, (size_t)B));
  SKD_CUDA(c, sx.alloc(&dsse, (size_t)B));
  SKD_CUDA(c, sx.alloc(&dcount, (size_t)B));
  SKD_CUDA(c, cudaMemcpyAsync(dslot, hs.data(), B * sizeof(SlotMeta), cudaMemcpyHostToDevice, c->stream));
  SKD_CUDA(c, cudaMemsetAsync(dsse, 0, B * sizeof(double), c->stream));
  SKD_CUDA(c, cudaMemsetAsync(dcount, 0, B * sizeof(int64_t), c->stream));
  if (c->kernel_choice == 2 && !tc_supported(c))
    return fail(c, "tcgen05 path requested but the staged shape is unsupported (needs d <= 256)");
  if (want_tc(c)) {
    if (tc_prepare(c)) return 1;
    LogregWork w;
    w.B = B; w.dp = (int)c->d + 1; w.use_tc = true; w.ldw = c->tc.dpad;
    w.slot = dslot;
    w.slots_pad_cap = (int)round_up(B, 128);
    size_t wbytes = (size_t)w.slots_pad_cap * c->tc.dpad * 2;
    SKD_CUDA(c, sx.alloc((uint8_t**)&w.Wh, wbytes));
    SKD_CUDA(c, sx.alloc((uint8_t**)&w.Wl, wbytes));
    SKD_CUDA(c, sx.alloc((uint8_t**)&w.sp, (size_t)w.slots_pad_cap * tc_slot_param_bytes()));
    SKD_CUDA(c, sx.alloc(&w.n_act, 1));
    SKD_CUDA(c, cudaMemsetAsync(w.Wh, 0, wbytes, c->stream));
    SKD_CUDA(c, cudaMemsetAsync(w.Wl, 0, wbytes, c->stream));
    std::vector<double> hx((size_t)B * w.dp);
    for (size_t i = 0; i < hx.size(); ++i) hx[i] = (double)coef[i];
    double* dx;
    SKD_CUDA(c, sx.alloc(&dx, hx.size()));
    int32_t nb = B;
    SKD_CUDA(c, cudaMemcpyAsync(dx, hx.data(), hx.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    SKD_CUDA(c, cudaMemcpyAsync(w.n_act, &nb, sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    SKD_CUDA(c, cudaStreamSynchronize(c->stream));
    c->h2d += (int64_t)hx.size() * 8;
    if (tc_export(c, w, B, dx, 1)) return 1;
    if (tc_r2(c, w, B, dsse, dcount)) return 1;
  } else {
    float* dW;
    if (pack_coef(c, sx, B, coef, &dW)) return 1;
    if (simt_r2(c, B, dW, dslot, dsse, dcount)) return 1;
  }
  SKD_CUDA(c, cudaMemcpyAsync(sse_out, dsse, B * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  SKD_CUDA(c, cudaMemcpyAsync(count_out, dcount, B * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
  SKD_CUDA(c, cudaStreamSynchronize(c->stream));
  c->d2h += (int64_t)B * 16;
  return 0;
}

int skd_sgd_fit_batch(skd_ctx* ctx, int32_t B, const int32_t* col_pos, int32_t loss, double alpha,
                      int32_t fit_intercept, int32_t max_iter, double tol, int32_t shuffle,
                      uint32_t seed, int32_t lr_type, double eta0, double power_t, double optimal_init,
                      int32_t n_iter_no_change, float* coef_out, double* intercept_out,
                      int32_t* n_iter_out, double* t_out, int32_t* status_out, double* gpu_seconds_out) {
  if (!ctx) return fail(nullptr, "skd_sgd_fit_batch: ctx is NULL");
  Ctx* c = &ctx->c;
  if (!c->X || !c->ycls) return fail(c, "skd_sgd_fit_batch: stage X and labels first");
  if (B <= 0 || !col_pos || !coef_out || !intercept_out || !n_iter_out || !t_out || !status_out)
    return fail(c, "skd_sgd_fit_batch: bad arguments");
  if (loss < 0 || loss > 1 || lr_type < 0 || lr_type > 2 || !(alpha > 0.0) || max_iter < 1)
    return fail(c, "skd_sgd_fit_batch: unsupported loss / learning rate / alpha / max_iter");
  SKD_CUDA(c, cudaSetDevice(c->device));
  cudaEvent_t e0, e1;
  SKD_CUDA(c, cudaEventCreate(&e0));
  SKD_CUDA(c, cudaEventCreate(&e1));
  SKD_CUDA(c, cudaEventRecord(e0, c->stream));
  int rc = sgd_fit_batch(c, B, col_pos, loss, alpha, fit_intercept, max_iter, tol, shuffle, seed, lr_type, eta0,
                         power_t, optimal_init, n_iter_no_change, coef_out, intercept_out, n_iter_out, t_out,
                         status_out);
  float ms = 0.f;
  if (!rc) {
    cudaEventRecord(e1, c->stream);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (gpu_seconds_out) *gpu_seconds_out = ms * 1e-3;
  return rc;
}

struct skd_forest {
  struct Tree {
    int32_t max_depth = 0, n_classes = 0, node_count = 0;
    std::vector<int32_t> left, right, feature, nsamp;
    std::vector<uint8_t> mgl;
    std::vector<double> thr, imp, wn, val;
    std::vector<uint32_t> compact;     // 8 words per node (throughput builder); expanded by skd_forest_tree_copy
  };
  std::vector<Tree> trees;
  std::vector<float> binval;           // [d][256] distinct feature values (thresholds of compact records)
};

static void forest_sink(void* arg, int t, const SkdTreeView* v) {
  skd_forest* f = (skd_forest*)arg;
  skd_forest::Tree& tr = f->trees[t];
  const int m = v->node_count;
  tr.max_depth = v->max_depth; tr.n_classes = v->n_classes; tr.node_count = m;
  if (v->compact) {
    tr.compact.assign(v->compact, v->compact + (size_t)m * 8);
    return;
  }
  tr.left.assign(v->left, v->left + m); tr.right.assign(v->right, v->right + m);
  tr.feature.assign(v->feature, v->feature + m); tr.nsamp.assign(v->n_node_samples, v->n_node_samples + m);
  tr.mgl.assign(v->missing_go_to_left, v->missing_go_to_left + m);
  tr.thr.assign(v->threshold, v->threshold + m); tr.imp.assign(v->impurity, v->impurity + m);
  tr.wn.assign(v->weighted_n_node_samples, v->weighted_n_node_samples + m);
  tr.val.assign(v->value, v->value + (size_t)m * v->n_classes);
}

int skd_forest_fit(skd_ctx* ctx, int32_t n_trees, const uint8_t* sample_counts, const uint32_t* rand_states,
                   int32_t n_classes, int32_t max_features, int32_t max_depth, int32_t min_samples_split,
                   int32_t min_samples_leaf, double min_weight_leaf, double min_impurity_decrease,
                   int32_t splitter, const double* y_regression, skd_forest** out, double* gpu_seconds_out) {
  if (!ctx) return fail(nullptr, "skd_forest_fit: ctx is NULL");
  Ctx* c = &ctx->c;
  if (!out) return fail(c, "skd_forest_fit: out is NULL");
  *out = nullptr;
  if (!c->X || (!y_regression && !c->ycls)) return fail(c, "skd_forest_fit: stage X and labels first");
  if (n_trees <= 0 || !rand_states || max_features < 1 || min_samples_split < 2 || min_samples_leaf < 1 ||
      splitter < 0 || splitter > 1)
    return fail(c, "skd_forest_fit: bad arguments");
  SKD_CUDA(c, cudaSetDevice(c->device));
  skd_forest* f = new skd_forest();
  f->trees.resize(n_trees);
  cudaEvent_t e0, e1;
  SKD_CUDA(c, cudaEventCreate(&e0));
  SKD_CUDA(c, cudaEventCreate(&e1));
  SKD_CUDA(c, cudaEventRecord(e0, c->stream));
  int rc = forest_fit(c, n_trees, sample_counts, rand_states, n_classes, max_features, max_depth, min_samples_split,
                      min_samples_leaf, min_weight_leaf, min_impurity_decrease, splitter, y_regression, forest_sink, f);
  if (!rc) f->binval = c->forest.h_binval;
  float ms = 0.f;
  if (!rc) {
    cudaEventRecord(e1, c->stream);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (gpu_seconds_out) *gpu_seconds_out = ms * 1e-3;
  if (rc) { delete f; return rc; }
  *out = f;
  return 0;
}

int skd_forest_kernel_seconds(skd_ctx* ctx, double* seconds_out) {
  if (!ctx || !seconds_out) return fail(nullptr, "skd_forest_kernel_seconds: bad arguments");
  *seconds_out = ctx->c.forest_kernel_ms * 1e-3;
  return 0;
}

int skd_forest_tree_size(skd_forest* f, int32_t tree, int32_t* node_count, int32_t* max_depth) {
  if (!f || tree < 0 || tree >= (int)f->trees.size()) return fail(nullptr, "skd_forest_tree_size: bad arguments");
  if (node_count) *node_count = f->trees[tree].node_count;
  if (max_depth) *max_depth = f->trees[tree].max_depth;
  return 0;
}

int skd_forest_tree_copy(skd_forest* f, int32_t tree, int32_t* left, int32_t* right, int32_t* feature,
                         double* threshold, double* impurity, int32_t* n_node_samples,
                         double* weighted_n_node_samples, uint8_t* missing_go_to_left, double* value) {
  if (!f || tree < 0 || tree >= (int)f->trees.size()) return fail(nullptr, "skd_forest_tree_copy: bad arguments");
  const skd_forest::Tree& t = f->trees[tree];
  if (!t.compact.empty()) {
    // Compact records of the throughput builder: {right child, feature | bin_a << 16 | bin_b << 24,
    // n_node_samples, depth, class sums[4]}, nodes in depth-first order (left child = id + 1).  The
    // float64 fields are formed here with the builder's (= scikit-learn's) operations:
    //   weighted_n = sum_c (double)s_c;  value_c = s_c / weighted_n            (SK/tree/_criterion.pyx:483-486)
    //   impurity   = 1 - (sum_c s_c^2) / (weighted_n * weighted_n)             (Gini, :650-680; what the parent's
    //                children_impurity computed from the same integers)
    //   threshold  = v[a] / 2 + v[b] / 2                                        (SK/tree/_splitter.pyx:459-461)
    const size_t m = (size_t)t.node_count;
    const int C = t.n_classes;
    const uint32_t* r = t.compact.data();
    const float* bv = f->binval.data();
    for (size_t i = 0; i < m; ++i, r += 8) {
      const int32_t rc = (int32_t)r[0];
      const uint32_t code = r[1];
      const bool leaf = (code & 0xFFFFu) == 0xFFFFu;
      if (left) left[i] = leaf ? -1 : (int32_t)i + 1;
      if (right) right[i] = leaf ? -1 : rc;
      if (feature) feature[i] = leaf ? -2 : (int32_t)(code & 0xFFFFu);
      if (threshold) {
        if (leaf) threshold[i] = -2.0;
        else {
          const size_t fo = (size_t)(code & 0xFFFFu) * 256;
          const volatile double ha = (double)bv[fo + ((code >> 16) & 0xFF)] / 2.0;
          const volatile double hb = (double)bv[fo + ((code >> 24) & 0xFF)] / 2.0;
          threshold[i] = ha + hb;
        }
      }
      volatile double w = 0.0, sq = 0.0;
      for (int c = 0; c < C; ++c) {
        const double a = (double)r[4 + c];
        w = w + a;
        const volatile double aa = a * a;      // volatile: no contraction of the product into the sum
        sq = sq + aa;
      }
      if (n_node_samples) n_node_samples[i] = (int32_t)r[2];
      if (weighted_n_node_samples) weighted_n_node_samples[i] = w;
      if (impurity) { const volatile double ww = w * w; const volatile double q = sq / ww; impurity[i] = 1.0 - q; }
      if (value) for (int c = 0; c < C; ++c) value[i * C + c] = (double)r[4 + c] / w;
      if (missing_go_to_left) missing_go_to_left[i] = 0;
    }
    if (missing_go_to_left) {      // n_left > n_right (SK/tree/_splitter.pyx: best_split.missing_go_to_left with no missing values)
      r = t.compact.data();
      for (size_t i = 0; i < m; ++i) {
        const int32_t rc = (int32_t)r[i * 8];
        if ((r[i * 8 + 1] & 0xFFFFu) != 0xFFFFu && rc > 0)
          missing_go_to_left[i] = r[(i + 1) * 8 + 2] > r[(size_t)rc * 8 + 2] ? 1 : 0;
      }
    }
    return 0;
  }
  const size_t m = t.left.size();
  if (left) memcpy(left, t.left.data(), m * 4);
  if (right) memcpy(right, t.right.data(), m * 4);
  if (feature) memcpy(feature, t.feature.data(), m * 4);
  if (threshold) memcpy(threshold, t.thr.data(), m * 8);
  if (impurity) memcpy(impurity, t.imp.data(), m * 8);
  if (n_node_samples) memcpy(n_node_samples, t.nsamp.data(), m * 4);
  if (weighted_n_node_samples) memcpy(weighted_n_node_samples, t.wn.data(), m * 8);
  if (missing_go_to_left) memcpy(missing_go_to_left, t.mgl.data(), m);
  if (value) memcpy(value, t.val.data(), t.val.size() * 8);
  return 0;
}

// The same tree as an array of 64-byte node records {left, right, feature: int64; threshold, impurity:
// float64; n_node_samples: int64; weighted_n_node_samples: float64; missing_go_to_left: uint8 + 7 pad}
// -- scikit-learn's `Node` struct (SK/tree/_tree.pxd:15-25), so that the caller can hand the buffer to
// `Tree.__setstate__` without building it field by field.
int skd_forest_tree_nodes(skd_forest* f, int32_t tree, void* nodes64, double* value) {
  if (!f || tree < 0 || tree >= (int)f->trees.size() || !nodes64) return fail(nullptr, "skd_forest_tree_nodes: bad arguments");
  const skd_forest::Tree& t = f->trees[tree];
  const size_t m = (size_t)t.node_count;
  struct Node64 { int64_t left, right, feature; double threshold, impurity; int64_t n_node_samples; double weighted; uint8_t mgl; uint8_t pad[7]; };
  static_assert(sizeof(Node64) == 64, "node record must be 64 bytes");
  Node64* out = (Node64*)nodes64;
  std::vector<int32_t> l(m), r(m), ft(m), ns(m);
  std::vector<uint8_t> mg(m);
  std::vector<double> th(m), im(m), wn(m);
  if (skd_forest_tree_copy(f, tree, l.data(), r.data(), ft.data(), th.data(), im.data(), ns.data(), wn.data(), mg.data(), value))
    return 1;
  for (size_t i = 0; i < m; ++i) {
    Node64 nd;
    nd.left = l[i]; nd.right = r[i]; nd.feature = ft[i]; nd.threshold = th[i]; nd.impurity = im[i];
    nd.n_node_samples = ns[i]; nd.weighted = wn[i]; nd.mgl = mg[i];
    memset(nd.pad, 0, sizeof(nd.pad));
    out[i] = nd;
  }
  return 0;
}

void skd_forest_free(skd_forest* f) { delete f; }

int skd_predict_linear(skd_ctx* ctx, const float* Xnew, int64_t m, int64_t d, int64_t ld, int32_t B,
                       const float* coef, float* out, double* gpu_seconds_out) {
  if (!ctx) return fail(nullptr, "skd_predict_linear: ctx is NULL");
  Ctx* c = &ctx->c;
  if (!Xnew || m <= 0 || d <= 0 || ld < d || B <= 0 || !coef || !out) return fail(c, "skd_predict_linear: bad arguments");
  SKD_CUDA(c, cudaSetDevice(c->device));
  const int64_t ldx = round_up(d, 4);
  Scratch sx(c);
  // weights [B x ldx] + bias[B]
  std::vector<float> hw((size_t)B * ldx + B, 0.f);
  for (int j = 0; j < B; ++j) {
    memcpy(&hw[(size_t)j * ldx], coef + (size_t)j * (d + 1), d * sizeof(float));
    hw[(size_t)B * ldx + j] = coef[(size_t)j * (d + 1) + d];
  }
  float* dW;
  SKD_CUDA(c, sx.alloc(&dW, hw.size()));
  SKD_CUDA(c, cudaMemcpyAsync(dW, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice, c->stream));
  // row chunks of <= 256 MiB: threaded pinned-bounce H2D (stage_rows_h2d), one pass of the kernel,
  // D2H of the (small) result.  The copy engine is the bottleneck (4*d bytes in per row, 4*B out).
  int64_t rows_per_chunk = std::max<int64_t>(1, ((int64_t)256 << 20) / (ldx * 4));
  if (rows_per_chunk > m) rows_per_chunk = m;
  float *dX, *dO;
  SKD_CUDA(c, sx.alloc(&dX, (size_t)rows_per_chunk * ldx));
  SKD_CUDA(c, sx.alloc(&dO, (size_t)rows_per_chunk * B));
  if (ldx != d) SKD_CUDA(c, cudaMemsetAsync(dX, 0, (size_t)rows_per_chunk * ldx * 4, c->stream));
  SKD_CUDA(c, cudaStreamSynchronize(c->stream));
  auto w0 = std::chrono::steady_clock::now();
  for (int64_t r0 = 0; r0 < m; r0 += rows_per_chunk) {
    int64_t mr = std::min(rows_per_chunk, m - r0);
    if (stage_rows_h2d(c, dX, ldx, Xnew + r0 * ld, mr, d, ld)) return 1;
    if (predict_device(c, dX, mr, (int)ldx, (int)d, B, dW, dO)) return 1;
    SKD_CUDA(c, cudaMemcpyAsync(out + r0 * B, dO, (size_t)mr * B * 4, cudaMemcpyDeviceToHost, c->stream));
    SKD_CUDA(c, cudaStreamSynchronize(c->stream));
    c->h2d += mr * d * 4;
    c->d2h += mr * (int64_t)B * 4;
  }
  if (gpu_seconds_out)
    *gpu_seconds_out = std::chrono::duration<double>(std::chrono::steady_clock::now() - w0).count();
  return 0;
}

int skd_forest_predict(skd_ctx* ctx, const float* Xnew, int64_t m, int64_t d, int64_t ld,
                       int32_t n_trees, const int64_t* tree_offset, const int32_t* left,
                       const int32_t* right, const int32_t* feature, const double* threshold,
                       const double* value, int32_t n_classes, double* proba_out,
                       double* gpu_seconds_out) {
  if (!ctx) return fail(nullptr, "skd_forest_predict: ctx is NULL");
  Ctx* c = &ctx->c;
  if (!Xnew || m <= 0 || d <= 0 || ld < d || n_trees <= 0 || !tree_offset || !left || !right || !feature ||
      !threshold || !value || n_classes <= 0 || !proba_out)
    return fail(c, "skd_forest_predict: bad arguments");
  SKD_CUDA(c, cudaSetDevice(c->device));
  const int64_t total = tree_offset[n_trees];
  if (tree_offset[0] != 0 || total <= 0) return fail(c, "skd_forest_predict: tree_offset must start at 0 and increase");
  // validate the tree arrays once on the host: a malformed child or feature index must not become
  // an out-of-bounds device read
  std::vector<int32_t> hnode((size_t)total * 4);
  for (int t = 0; t < n_trees; ++t) {
    const int64_t b = tree_offset[t], e = tree_offset[t + 1];
    if (e <= b) return fail(c, "skd_forest_predict: empty tree");
    for (int64_t k = b; k < e; ++k) {
      const int32_t l = left[k], r = right[k], f = feature[k];
      if (l != -1 && (l <= 0 || l >= e - b || r <= 0 || r >= e - b || f < 0 || f >= d))
        return fail(c, "skd_forest_predict: malformed tree arrays");
      hnode[(size_t)k * 4] = l; hnode[(size_t)k * 4 + 1] = r; hnode[(size_t)k * 4 + 2] = l == -1 ? 0 : f;
      hnode[(size_t)k * 4 + 3] = 0;
    }
  }
  Scratch sx(c);
  int64_t* d_off; int32_t* d_node; double *d_thr, *d_val;
  SKD_CUDA(c, sx.alloc(&d_off, (size_t)n_trees + 1));
  SKD_CUDA(c, sx.alloc(&d_node, (size_t)total * 4));
  SKD_CUDA(c, sx.alloc(&d_thr, (size_t)total));
  SKD_CUDA(c, sx.alloc(&d_val, (size_t)total * n_classes));
  SKD_CUDA(c, cudaMemcpyAsync(d_off, tree_offset, ((size_t)n_trees + 1) * 8, cudaMemcpyHostToDevice, c->stream));
  SKD_CUDA(c, cudaMemcpyAsync(d_node, hnode.data(), hnode.size() * 4, cudaMemcpyHostToDevice, c->stream));
  SKD_CUDA(c, cudaMemcpyAsync(d_thr, threshold, (size_t)total * 8, cudaMemcpyHostToDevice, c->stream));
  SKD_CUDA(c, cudaMemcpyAsync(d_val, value, (size_t)total * n_classes * 8, cudaMemcpyHostToDevice, c->stream));
  c->h2d += total * (int64_t)(16 + 8 + 8 * n_classes);
  const int64_t ldx = round_up(d, 4);
  int64_t rows_per_chunk = std::max<int64_t>(1, ((int64_t)256 << 20) / (ldx * 4));
  if (rows_per_chunk > m) rows_per_chunk = m;
  float* dX; double* dO;
  SKD_CUDA(c, sx.alloc(&dX, (size_t)rows_per_chunk * ldx));
  SKD_CUDA(c, sx.alloc(&dO, (size_t)rows_per_chunk * n_classes));
  if (ldx != d) SKD_CUDA(c, cudaMemsetAsync(dX, 0, (size_t)rows_per_chunk * ldx * 4, c->stream));
  SKD_CUDA(c, cudaStreamSynchronize(c->stream));
  auto w0 = std::chrono::steady_clock::now();
  for (int64_t r0 = 0; r0 < m; r0 += rows_per_chunk) {
    const int64_t mr = std::min(rows_per_chunk, m - r0);
    if (stage_rows_h2d(c, dX, ldx, Xnew + r0 * ld, mr, d, ld)) return 1;
    if (forest_predict_device(c, dX, mr, (int)ldx, n_trees, d_off, d_node, d_thr, d_val, n_classes, dO)) return 1;
    SKD_CUDA(c, cudaMemcpyAsync(proba_out + r0 * n_classes, dO, (size_t)mr * n_classes * 8, cudaMemcpyDeviceToHost,
                                c->stream));
    SKD_CUDA(c, cudaStreamSynchronize(c->stream));
    c->h2d += mr * d * 4;
    c->d2h += mr * (int64_t)n_classes * 8;
  }
  if (gpu_seconds_out)
    *gpu_seconds_out = std::chrono::duration<double>(std::chrono::steady_clock::now() - w0).count();
  return 0;
}

int skd_linear_decision(skd_ctx* ctx, int32_t B, const float* coef, float* out) {
  if (!ctx) return fail(nullptr, "skd_linear_decision: ctx is NULL");
  Ctx* c = &ctx->c;
  if (!c->X) return fail(c, "skd_linear_decision: stage X first");
  if (B <= 0 || !coef || !out) return fail(c, "skd_linear_decision: bad arguments");
  SKD_CUDA(c, cudaSetDevice(c->device));
  Scratch sx(c);
  float *dW, *dout;
  if (pack_coef(c, sx, B, coef, &dW)) return 1;
  SKD_CUDA(c, sx.alloc(&dout, (size_t)c->n * B));
  if (B <= 16 && c->ldx * 4 * 8 <= 48 * 1024) {
    if (predict_device(c, c->X, c->n, (int)c->ldx, (int)c->d, B, dW, dout)) return 1;
  } else if (simt_decision(c, B, dW, dout)) return 1;
  SKD_CUDA(c, cudaMemcpyAsync(out, dout, (size_t)c->n * B * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  SKD_CUDA(c, cudaStreamSynchronize(c->stream));
  c->d2h += (int64_t)c->n * B * 4;
  return 0;
}

// ---- host-side L-BFGS object (tests) ----------------------------------------------------
skd_lbfgs* skd_lbfgs_create(int32_t n, int32_t m, int32_t maxiter, int32_t maxls, double pgtol,
                            double ftol) {
  skd_lbfgs* h = new skd_lbfgs();
  lbfgs_init(h->s, n, m, maxiter, maxls, pgtol, ftol);
  h->buf.assign((size_t)5 * n + 2 * (size_t)m * n + 2 * m, 0.0);
  double* p = h->buf.data();
  h->v.x = p; p += n;
  h->v.g = p; p += n;
  h->v.t = p; p += n;
  h->v.r = p; p += n;
  h->v.d = p; p += n;
  h->v.S = p; p += (size_t)m * n;
  h->v.Y = p; p += (size_t)m * n;
  h->v.rho = p; p += m;
  h->v.alpha = p;
  return h;
}
double* skd_lbfgs_x(skd_lbfgs* h) { return h->v.x; }
double* skd_lbfgs_g(skd_lbfgs* h) { return h->v.g; }
int skd_lbfgs_advance(skd_lbfgs* h, double f) {
  SeqPar P;
  lbfgs_advance(P, h->s, h->v, f);
  return h->s.status;
}
int skd_lbfgs_nit(skd_lbfgs* h) { return h->s.nit; }
int skd_lbfgs_nfev(skd_lbfgs* h) { return h->s.nfev; }
void skd_lbfgs_free(skd_lbfgs* h) { delete h; }

}  // extern "C"
