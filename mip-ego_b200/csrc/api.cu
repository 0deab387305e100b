// C ABI of libmgp_b200 (see include/mgp.h): handle management and host-side orchestration of
// the CUDA kernels.  No compute happens on the host; without a CUDA device every entry point
// fails with MGP_ECUDA.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "linalg.cuh"
#include "mgp_common.cuh"
#include "mgp_handle.h"
#include "nll.cuh"
#include "ozaki.cuh"
#include "posterior.cuh"

int mgp_fail(mgp_handle *h, int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h) h->err = buf;
  return code;
}
int mgp_cuda_fail(mgp_handle *h, cudaError_t e, const char *what) {
  return mgp_fail(h, e == cudaErrorMemoryAllocation ? MGP_ENOMEM : MGP_ECUDA, "CUDA error '%s' in %s",
                  cudaGetErrorString(e), what);
}
int mgp_ensure(mgp_handle *h, DevBuf &b, size_t bytes) {
  if (b.bytes >= bytes && b.p) return 0;
  if (b.p) {
    cudaFree(b.p);
    b.p = nullptr;
    b.bytes = 0;
  }
  if (bytes == 0) return 0;
  cudaError_t e = cudaMalloc(&b.p, bytes);
  if (e != cudaSuccess) return mgp_cuda_fail(h, e, "cudaMalloc");
  b.bytes = bytes;
  return 0;
}
static void freebuf(DevBuf &b) {
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.bytes = 0;
}
template <typename T>
static inline T *P(const DevBuf &b) { return reinterpret_cast<T *>(b.p); }

struct KTimer {
  mgp_handle *h; cudaStream_t st; int which; cudaEvent_t a = nullptr, b = nullptr;
  KTimer(mgp_handle *h_, cudaStream_t st_, int which_) : h(h_), st(st_), which(which_) {
    if (h->time_kernels) {
      cudaEventCreate(&a); cudaEventCreate(&b);
      cudaEventRecord(a, st);
    }
  }
  ~KTimer() {
    if (a) {
      cudaEventRecord(b, st);
      h->ev_pending.push_back({a, b, which});
    }
  }
};
static void resolve_timers(mgp_handle *h) {
  for (auto &e : h->ev_pending) {
    float ms = 0.f;
    if (cudaEventSynchronize(e.b) == cudaSuccess && cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess) {
      h->kernel_ms[e.which] += ms;
      h->kernel_calls[e.which]++;
    }
    cudaEventDestroy(e.a); cudaEventDestroy(e.b);
  }
  h->ev_pending.clear();
}

// See mgp_handle::ev_ws.  A wait on an event recorded on the same stream is a no-op for the device.
static cudaError_t ws_acquire(mgp_handle *h, cudaStream_t st) {
  if (!h->ws_recorded) return cudaSuccess;
  return cudaStreamWaitEvent(st, h->ev_ws, 0);
}
static cudaError_t ws_release(mgp_handle *h, cudaStream_t st) {
  h->ws_recorded = true;
  return cudaEventRecord(h->ev_ws, st);
}
struct WsScope {   // acquire on construction, release on destruction (every return path)
  mgp_handle *h; cudaStream_t st; cudaError_t err;
  WsScope(mgp_handle *h_, cudaStream_t st_) : h(h_), st(st_), err(ws_acquire(h_, st_)) {}
  ~WsScope() { ws_release(h, st); }
};

#define CHECK_H(h) \
  if (!(h)) return MGP_EINVAL; \
  MGP_CUDA(h, cudaSetDevice((h)->device))

extern "C" {

static void nll_graphs_clear(mgp_handle *h);

int mgp_abi_version(void) { return MGP_ABI_VERSION; }

int mgp_create(int device, mgp_t **out_h) {
  if (!out_h) return MGP_EINVAL;
  *out_h = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0 || device < 0 || device >= count) return MGP_ECUDA;
  if (cudaSetDevice(device) != cudaSuccess) return MGP_ECUDA;
  mgp_handle *h = new mgp_handle();
  h->device = device;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete h; return MGP_ECUDA; }
  if (prop.major != 10) {
    // built for sm_100a only: fail loudly rather than at the first launch
    delete h;
    return MGP_ECUDA;
  }
  h->sm_count = prop.multiProcessorCount;
  {   // MGP_TRMM_I8=1: every handle starts with the INT8 (Ozaki-scheme) K2c switched on (option "trmm_i8")
    const char *e = getenv("MGP_TRMM_I8");
    h->trmm_i8_opt = (e && *e && *e != '0') ? (*e == '7' ? 7 : 8) : 0;
  }
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&h->out_stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete h;
    return MGP_ECUDA;
  }
  for (int i = 0; i < MGP_NLL_STREAMS; i++) {
    // Descending priorities: when batch groups run concurrently the block scheduler then always
    // prefers the pending CTAs of the group that is furthest ahead, so its short kernels (diagonal
    // blocks, panel solves) are not queued behind the wide GEMMs of the other groups
    // (n = 4096, 8 groups: 21.9 -> 21.3 ms per 8 evaluations).
    int pr_lo = 0, pr_hi = 0;
    cudaDeviceGetStreamPriorityRange(&pr_lo, &pr_hi);   // pr_hi is the numerically smallest (greatest priority)
    const int pr = pr_hi + i > pr_lo ? pr_lo : pr_hi + i;
    if (cudaStreamCreateWithPriority(&h->nll_stream[i], cudaStreamNonBlocking, pr) != cudaSuccess)
      cudaStreamCreateWithFlags(&h->nll_stream[i], cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&h->ev_join[i], cudaEventDisableTiming);
    if (cudaStreamCreateWithPriority(&h->nll_side[i], cudaStreamNonBlocking, pr) != cudaSuccess)
      cudaStreamCreateWithFlags(&h->nll_side[i], cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&h->ev_sfork[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_sjoin[i], cudaEventDisableTiming);
  }
  {   // MGP_TRTRI_ROWWISE=0/1 overrides the automatic choice (A/B measurements)
    const char *e = getenv("MGP_TRTRI_ROWWISE");
    if (e && (*e == '0' || *e == '1')) h->rowinv_opt = *e - '0';
  }
  cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&h->ev_ws, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&h->ev_gather, cudaEventDisableTiming);
  for (int i = 0; i < 2; i++) {
    cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_out[i], cudaEventDisableTiming);
  }
  *out_h = h;
  return MGP_OK;
}

int mgp_destroy(mgp_t *h) {
  if (!h) return MGP_OK;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  mgp_comm_destroy(h);
  nll_graphs_clear(h);
  DevBuf *all[] = {&h->dtpart, &h->dXc, &h->dy, &h->dcenter, &h->dtheta, &h->dBop, &h->dan, &h->dXsc, &h->dL, &h->dMinv,
                   &h->dTmp, &h->dMp, &h->dRinvP, &h->dgamma, &h->davec, &h->dFt,
                   &h->dYt, &h->drho, &h->dscal, &h->drP, &h->dmeanpart, &h->dftpart, &h->dqpart,
                   &h->dwP, &h->dXs_stage[0], &h->dXs_stage[1], &h->dout_stage[0],
                   &h->dout_stage[1], &h->dtopk_val, &h->dtopk_idx, &h->dtopk_blk, &h->nR, &h->nL, &h->nM, &h->nT,
                   &h->nLinvDiag, &h->nvec, &h->nscal, &h->npar, &h->ngpart, &h->nFtm, &h->nBhat, &h->ntws,
                   &h->dF, &h->dmu, &h->dbetav, &h->dBhat, &h->dCinv, &h->dFtm, &h->niso, &h->dcomm, &h->dgather, &h->dgather_io, &h->ngio, &h->dMI8, &h->dMscale, &h->dRI8};
  for (DevBuf *b : all) freebuf(*b);
  for (int i = 0; i < 2; i++) {
    if (h->ev_in[i]) cudaEventDestroy(h->ev_in[i]);
    if (h->ev_done[i]) cudaEventDestroy(h->ev_done[i]);
    if (h->ev_out[i]) cudaEventDestroy(h->ev_out[i]);
  }
  for (int i = 0; i < MGP_NLL_STREAMS; i++) {
    if (h->nll_stream[i]) { release_dgemm_stream(h->nll_stream[i]); cudaStreamDestroy(h->nll_stream[i]); }
    if (h->ev_join[i]) cudaEventDestroy(h->ev_join[i]);
    if (h->nll_side[i]) { release_dgemm_stream(h->nll_side[i]); cudaStreamDestroy(h->nll_side[i]); }
    if (h->ev_sfork[i]) cudaEventDestroy(h->ev_sfork[i]);
    if (h->ev_sjoin[i]) cudaEventDestroy(h->ev_sjoin[i]);
  }
  if (h->hpin) cudaFreeHost(h->hpin);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_ws) cudaEventDestroy(h->ev_ws);
  if (h->ev_gather) cudaEventDestroy(h->ev_gather);
  if (h->stream) release_dgemm_stream(h->stream);
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->out_stream) cudaStreamDestroy(h->out_stream);
  delete h;
  return MGP_OK;
}

const char *mgp_last_error(const mgp_t *h) { return h ? h->err.c_str() : "null handle"; }

int mgp_set_option(mgp_t *h, const char *name, int64_t value) {
  if (!h || !name) return MGP_EINVAL;
  if (!strcmp(name, "chunk")) {
    if (value < 0) return mgp_fail(h, MGP_EINVAL, "chunk must be >= 0");
    h->chunk_opt = round_up(value, 128);
    return MGP_OK;
  }
  if (!strcmp(name, "small_path")) {   // 0 disables the matrix-vector path for <= 8 candidates (tests)
    h->no_small_path = value == 0;
    return MGP_OK;
  }
  if (!strcmp(name, "nll_groups")) {
    if (value < 0 || value > MGP_NLL_STREAMS) return mgp_fail(h, MGP_EINVAL, "nll_groups must be in 0..%d", MGP_NLL_STREAMS);
    h->nll_groups_opt = (int)value;
    nll_graphs_clear(h);
    return MGP_OK;
  }
  if (!strcmp(name, "trtri_rowwise")) {   // -1 automatic, 0 recursive inversion after the factorisation, 1 row-wise beside it
    if (value < -1 || value > 1) return mgp_fail(h, MGP_EINVAL, "trtri_rowwise must be -1, 0 or 1");
    h->rowinv_opt = (int)value;
    nll_graphs_clear(h);
    return MGP_OK;
  }
  if (!strcmp(name, "gemm_cta_cap")) {
    h->gemm_cap_opt = (int)value;
    nll_graphs_clear(h);
    return MGP_OK;
  }
  if (!strcmp(name, "trmm_i8")) {   // K2c on the INT8 tensor pipe (Ozaki split): 0 off (FP64 DMMA), 1 or 8: 8 slices, 7: 7 slices
    const int v = value == 0 ? 0 : (value == 7 ? 7 : 8);
    if (v != h->trmm_i8_opt) h->i8_ready = false;      // the packed slices depend on the slice count
    h->trmm_i8_opt = v;
    return MGP_OK;
  }
  if (!strcmp(name, "zero_copy")) {   // 1 (default): calls with <= 4 KB of candidates go through a mapped pinned page (no copy engine)
    h->zero_copy_opt = value != 0;
    return MGP_OK;
  }
  if (!strcmp(name, "trmm_i8_fuse")) {   // 1 (default): K1 emits the int8 slices of r itself where it can; 0: separate slicing pass
    h->trmm_i8_fuse = value != 0;
    return MGP_OK;
  }
  if (!strcmp(name, "trmm_i8_pair")) {   // K2c-i8 as two-CTA clusters sharing the candidate slices by multicast
    h->trmm_i8_pair = value != 0;
    return MGP_OK;
  }
  if (!strcmp(name, "nll_graphs")) {   // 0: every batched likelihood is issued launch by launch
    h->graphs_opt = value != 0;
    if (!h->graphs_opt) nll_graphs_clear(h);
    return MGP_OK;
  }
  if (!strcmp(name, "isotropic")) {   // one theta for all features: n_par shrinks by d - 1
    h->isotropic = value != 0;
    return MGP_OK;
  }
  if (!strcmp(name, "time_kernels")) {
    resolve_timers(h);
    h->time_kernels = value != 0;
    for (int i = 0; i < 4; i++) { h->kernel_ms[i] = 0; h->kernel_calls[i] = 0; }
    return MGP_OK;
  }
  return mgp_fail(h, MGP_EINVAL, "unknown option '%s'", name);
}
int64_t mgp_get_counter(const mgp_t *h, const char *name) {
  if (!h || !name) return -1;
  if (!strcmp(name, "launches")) return h->launches;
  if (!strcmp(name, "npad")) return h->npad;
  if (!strcmp(name, "chunk")) return h->chunk;
  if (!strcmp(name, "sm_count")) return h->sm_count;
  if (!strcmp(name, "small_path")) return h->no_small_path ? 0 : 1;
  if (!strcmp(name, "isotropic")) return h->isotropic ? 1 : 0;
  if (!strcmp(name, "n")) return h->n;
  if (!strcmp(name, "appends")) return h->appends;
  if (!strcmp(name, "trmm_i8")) return h->trmm_i8_opt;
  if (!strcmp(name, "trmm_i8_pair")) return h->trmm_i8_pair;
  if (!strcmp(name, "trmm_i8_fuse")) return h->trmm_i8_fuse;
  if (!strcmp(name, "zero_copy")) return h->zero_copy_opt;
  if (!strcmp(name, "graph_replays")) return h->graph_replays;
  if (!strcmp(name, "nll_graphs")) return (int64_t)h->nll_graphs.size();
  if (!strcmp(name, "world")) return h->world;
  if (!strcmp(name, "rank")) return h->rank;
  if (!strcmp(name, "bcast_bytes")) return h->bcast_bytes;
  static const char *kn[4] = {"rgen", "trmm", "epilogue", "topk"};
  for (int i = 0; i < 4; i++) {
    char buf[64];
    snprintf(buf, sizeof(buf), "%s_ns", kn[i]);
    if (!strcmp(name, buf)) { resolve_timers(const_cast<mgp_handle *>(h)); return (int64_t)(h->kernel_ms[i] * 1e6); }
    snprintf(buf, sizeof(buf), "%s_calls", kn[i]);
    if (!strcmp(name, buf)) { resolve_timers(const_cast<mgp_handle *>(h)); return h->kernel_calls[i]; }
  }
  return -1;
}

// ---------------------------------------------------------------------------------------------
// training data
// ---------------------------------------------------------------------------------------------
int mgp_set_train(mgp_t *h, int n, int d, const double *X, const double *y, int corr, int trend,
                  const double *beta) {
  CHECK_H(h);
  if (n < 1 || d < 1 || !X || !y) return mgp_fail(h, MGP_EINVAL, "bad training data");
  if (d > MGP_MAX_D) return mgp_fail(h, MGP_EINVAL, "d = %d > %d unsupported", d, MGP_MAX_D);
  if (corr != MGP_CORR_SE && corr != MGP_CORR_MATERN32 && corr != MGP_CORR_ABSEXP)
    return mgp_fail(h, MGP_EINVAL, "correlation type %d not supported by the CUDA path", corr);
  if (trend < MGP_TREND_CONST_FIXED || trend > MGP_TREND_LINEAR_ESTIMATED)
    return mgp_fail(h, MGP_EINVAL, "trend type %d unknown", trend);
  const bool linear = trend >= MGP_TREND_LINEAR_FIXED;
  const bool est = trend == MGP_TREND_CONST_ESTIMATED || trend == MGP_TREND_LINEAR_ESTIMATED;
  const int p = linear ? d + 1 : 1;
  if (!est && !beta) return mgp_fail(h, MGP_EINVAL, "a trend with given coefficients needs beta");
  if (est && n <= p) return mgp_fail(h, MGP_EINVAL, "n = %d samples cannot determine %d trend coefficients", n, p);
  h->fitted = false;
  h->rinv_ready = false;
  h->i8_ready = false;
  h->nll_cap = 0;  // workspaces are re-validated against the new padded size
  h->n = n; h->d = d; h->corr = corr; h->trend = trend;
  h->p = p; h->trend_est = est;
  h->beta_in.assign(p, 0.0);
  if (!est) h->beta_in.assign(beta, beta + p);
  h->beta0 = (!est && !linear) ? beta[0] : 0.0;
  h->npad = (int)round_up(n, MGP_TILE);
  h->dpad = (int)round_up(d, 4);
  h->NI = h->npad / MGP_TILE;
  h->hX.assign(X, X + (size_t)n * d);
  h->hy.assign(y, y + n);
  // centring is input marshalling: differences x - x' are translation invariant
  std::vector<double> center(h->dpad, 0.0), Xc((size_t)h->npad * h->dpad, 0.0), yp(h->npad, 0.0);
  for (int i = 0; i < d; i++) {
    double s = 0.0;
    for (int j = 0; j < n; j++) s += X[(size_t)j * d + i];
    center[i] = s / n;
  }
  for (int j = 0; j < n; j++) {
    for (int i = 0; i < d; i++) Xc[(size_t)j * h->dpad + i] = X[(size_t)j * d + i] - center[i];
    yp[j] = y[j];
  }
  WsScope ws(h, h->stream);   // a *_dev call may still be reading the training data on another stream
  MGP_CUDA(h, ws.err);
  h->hcenter = center;
  h->hpar.clear();
  nll_graphs_clear(h);
  MGP_TRY(mgp_ensure(h, h->dXc, Xc.size() * 8));
  MGP_TRY(mgp_ensure(h, h->dy, yp.size() * 8));
  MGP_TRY(mgp_ensure(h, h->dcenter, center.size() * 8));
  MGP_TRY(mgp_ensure(h, h->dscal, 4096));
  MGP_CUDA(h, cudaMemcpyAsync(h->dXc.p, Xc.data(), Xc.size() * 8, cudaMemcpyHostToDevice, h->stream));
  MGP_CUDA(h, cudaMemcpyAsync(h->dy.p, yp.data(), yp.size() * 8, cudaMemcpyHostToDevice, h->stream));
  MGP_CUDA(h, cudaMemcpyAsync(h->dcenter.p, center.data(), center.size() * 8, cudaMemcpyHostToDevice, h->stream));
  // duplicate rows: raise like gpr.py:275-277 (compared on the raw, uncentred coordinates'
  // differences: x_j == x_k  <=>  xc_j == xc_k up to the same subtraction)
  int *flag = P<int>(h->dscal);
  MGP_CUDA(h, cudaMemsetAsync(flag, 0, sizeof(int), h->stream));
  MGP_CUDA(h, launch_dup_check(P<double>(h->dXc), n, h->dpad, flag, h->stream));
  h->launches++;
  int hflag = 0;
  MGP_CUDA(h, cudaMemcpyAsync(&hflag, flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  MGP_CUDA(h, cudaStreamSynchronize(h->stream));
  if (hflag) {
    h->n = 0;
    return mgp_fail(h, MGP_EINVAL, "duplicate rows: Multiple input features cannot have the same target value.");
  }
  // non-constant trends: the basis matrix F (coefficients estimated) or the mean vector F beta
  if (linear) {
    MGP_TRY(mgp_ensure(h, h->dbetav, MGP_PMAX * 8));
    if (est) {
      MGP_TRY(mgp_ensure(h, h->dF, (size_t)h->npad * MGP_TP * 8));
      MGP_CUDA(h, launch_trend_basis(P<double>(h->dXc), P<double>(h->dcenter), n, h->npad, d, h->dpad,
                                     P<double>(h->dF), h->stream));
      // log det(F^T F) (REML, gpr.py:736) with the same normal-equation kernel, on F itself
      MGP_TRY(mgp_ensure(h, h->ntws, trend_ws_doubles() * 8));
      MGP_TRY(mgp_ensure(h, h->dmu, (size_t)h->npad * 8));
      double *sc = P<double>(h->dscal) + 16;
      int *stflag = reinterpret_cast<int *>(P<double>(h->dscal) + 8);
      MGP_CUDA(h, cudaMemsetAsync(stflag, 0, sizeof(int), h->stream));
      MGP_CUDA(h, launch_trend_solve(P<double>(h->dF), P<double>(h->dy), P<double>(h->dmu), n, h->npad, p,
                                     P<double>(h->ntws), sc, stflag, 1, h->stream));
      double ld = 0.0;
      int bad = 0;
      MGP_CUDA(h, cudaMemcpyAsync(&ld, sc + NLL_LOGDETA, 8, cudaMemcpyDeviceToHost, h->stream));
      MGP_CUDA(h, cudaMemcpyAsync(&bad, stflag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
      MGP_CUDA(h, cudaStreamSynchronize(h->stream));
      if (bad) { h->n = 0; return mgp_fail(h, MGP_EINVAL, "the trend basis F is rank deficient"); }
      h->logdetFF = ld;
      h->launches += 2;
    } else {
      MGP_TRY(mgp_ensure(h, h->dmu, (size_t)h->npad * 8));
      std::vector<double> bp(MGP_PMAX, 0.0);
      for (int i = 0; i < p; i++) bp[i] = h->beta_in[i];
      MGP_CUDA(h, cudaMemcpyAsync(h->dbetav.p, bp.data(), MGP_PMAX * 8, cudaMemcpyHostToDevice, h->stream));
      MGP_CUDA(h, launch_trend_mu(P<double>(h->dXc), P<double>(h->dcenter), n, h->npad, d, h->dpad,
                                  P<double>(h->dbetav), P<double>(h->dmu), h->stream));
      MGP_CUDA(h, cudaStreamSynchronize(h->stream));   // bp is a host temporary
      h->launches++;
    }
  }
  return MGP_OK;
}

// ---------------------------------------------------------------------------------------------
// likelihood
// ---------------------------------------------------------------------------------------------
static int nll_alloc(mgp_handle *h, int B) {
  if (B <= h->nll_cap) return 0;
  const size_t mat = (size_t)h->npad * h->npad * 8;
  const int ntiles = h->NI * (h->NI + 1) / 2;
  MGP_TRY(mgp_ensure(h, h->nR, mat * B));
  MGP_TRY(mgp_ensure(h, h->nL, mat * B));
  MGP_TRY(mgp_ensure(h, h->nM, mat * B));
  MGP_TRY(mgp_ensure(h, h->nT, mat * B));
  MGP_TRY(mgp_ensure(h, h->nLinvDiag, (size_t)B * h->NI * 128 * 128 * 8));
  MGP_TRY(mgp_ensure(h, h->nvec, (size_t)5 * B * h->npad * 8));
  MGP_TRY(mgp_ensure(h, h->nscal, (size_t)B * nll_scal_doubles() * 8 + (size_t)B * sizeof(int) + 64));
  MGP_TRY(mgp_ensure(h, h->ngpart, (size_t)B * ntiles * 4 * (h->dpad + 2) * 8));   // four parts per tile (k_nll_grad)
  if (h->p > 1 && h->trend_est) {
    MGP_TRY(mgp_ensure(h, h->nFtm, (size_t)B * h->npad * MGP_TP * 8));
    MGP_TRY(mgp_ensure(h, h->nBhat, (size_t)B * h->npad * MGP_TP * 8));
    MGP_TRY(mgp_ensure(h, h->ntws, (size_t)B * trend_ws_doubles() * 8));
  }
  h->nll_cap = B;
  return 0;
}
static int *nll_status_ptr(mgp_handle *h, int cap) {
  return reinterpret_cast<int *>(P<double>(h->nscal) + (size_t)cap * nll_scal_doubles());
}
static int nll_max_batch(mgp_handle *h) {
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return 1;
  const size_t per = (size_t)h->npad * h->npad * 8 * 4 + (size_t)h->NI * 128 * 128 * 8;
  size_t budget = (free_b + h->nR.bytes + h->nL.bytes + h->nM.bytes + h->nT.bytes) / 2;
  int cap = (int)std::max<size_t>(1, budget / per);
  return std::min(cap, 64);
}

static int check_par(mgp_handle *h, int n_par, int mode) {
  if (h->n == 0) return mgp_fail(h, MGP_ESTATE, "mgp_set_train must be called first");
  const int est = mode & 15;
  if ((mode & ~(15 | MGP_LIK_RESTRICTED)) || est > MGP_MODE_NOISE_ESTIM)
    return mgp_fail(h, MGP_EINVAL, "estimation mode %d not supported", mode);
  const int n_theta = h->isotropic ? 1 : h->d;
  const int want = nll_n_par(mode, h->d) - h->d + n_theta;
  if (n_par != want)
    return mgp_fail(h, MGP_EINVAL, "n_par = %d, expected %d for mode %d (%d theta + %d)", n_par,
                    want, mode, n_theta, want - n_theta);
  return 0;
}
// Isotropic theta (kernel.py:312-317): the kernels always work with d values; a host vector of the
// short layout [theta, extras...] is expanded to [theta x d, extras...].
static void iso_expand_host(const mgp_handle *h, const double *par, int n_par, std::vector<double> &full) {
  full.assign(h->d, par[0]);
  for (int i = 1; i < n_par; i++) full.push_back(par[i]);
}

// Work descriptor for `B` parameter vectors living in workspace slots b_off .. b_off + B - 1 of a
// workspace with capacity `cap`.  d_params / d_llf / d_grad point at the first vector of the slice.
static void fill_work(mgp_handle *h, NllWork &w, int b_off, int B, int cap, int n_par, int mode,
                      double noise_var, int eval_grad, const double *d_params, double *d_llf,
                      double *d_grad) {
  memset(&w, 0, sizeof(w));
  const size_t mat = (size_t)h->npad * h->npad, sec = (size_t)cap * h->npad;
  const int ntiles = h->NI * (h->NI + 1) / 2;
  w.n = h->n; w.npad = h->npad; w.dpad = h->dpad; w.B = B; w.n_par = n_par; w.mode = mode;
  w.corr = h->corr; w.ordinary = h->trend == MGP_TREND_CONST_ESTIMATED; w.eval_grad = eval_grad;
  w.noise_var = noise_var; w.beta0 = h->beta0;
  w.Xc = P<double>(h->dXc); w.y = P<double>(h->dy); w.params = d_params;
  w.R = P<double>(h->nR) + b_off * mat; w.L = P<double>(h->nL) + b_off * mat;
  w.M = P<double>(h->nM) + b_off * mat; w.T = P<double>(h->nT) + b_off * mat;
  w.linv = P<double>(h->nLinvDiag) + (size_t)b_off * h->NI * 128 * 128;
  double *v = P<double>(h->nvec) + (size_t)b_off * h->npad;
  w.Yt = v; w.Ft = v + sec; w.rho = v + 2 * sec; w.gamma = v + 3 * sec; w.avec = v + 4 * sec;
  w.scal = P<double>(h->nscal) + (size_t)b_off * nll_scal_doubles();
  w.gpart = P<double>(h->ngpart) + (size_t)b_off * ntiles * 4 * (h->dpad + 2);
  w.status = nll_status_ptr(h, cap) + b_off;
  w.out_llf = d_llf; w.out_grad = d_grad;
  w.p = h->p;
  w.ordinary = h->trend_est;
  w.mu = (h->p > 1 && !h->trend_est) ? P<double>(h->dmu) : nullptr;
  if (h->p > 1 && h->trend_est) {
    w.F = P<double>(h->dF);
    w.Ftm = P<double>(h->nFtm) + (size_t)b_off * h->npad * MGP_TP;
    w.Bhat = P<double>(h->nBhat) + (size_t)b_off * h->npad * MGP_TP;
    w.tws = P<double>(h->ntws) + (size_t)b_off * trend_ws_doubles();
    w.logdetFF = h->logdetFF;
  }
}

// Number of concurrent batch groups: the factorisation of one group has long phases that occupy
// only a few SMs (diagonal blocks, panel solves); running the groups on separate streams lets the
// wide kernels of one group fill the machine while another group is in such a phase.
static int nll_groups(const mgp_handle *h, int nb) {
  // default: more groups for larger matrices -- for small matrices batching the short kernels of all
  // evaluations into one launch beats overlapping them (measured, 8 evaluations per call, evaluations/s
  // with 1 / 2 / 4 / 8 groups: n = 768: 8205 / 8001 / 7468 / 5990; n = 1024: 6172 / 6145 / 5774 / 4962;
  // n = 1536: 3310 / 3328 / 3292 / 3109; n = 2048: 1899 / 1993 / 2042 / 1953; n = 3072: 755 / 788 / 813 / 825)
  int g = h->nll_groups_opt > 0 ? h->nll_groups_opt
          : (h->npad <= 768 ? 1 : (h->npad <= 1536 ? 2 : (h->npad <= 2560 ? 4 : MGP_NLL_STREAMS)));
  if (g > MGP_NLL_STREAMS) g = MGP_NLL_STREAMS;
  if (g > nb) g = nb;
  return g < 1 ? 1 : g;
}

// Issues the launches of one batched likelihood on `st` (and, for several batch groups, on the
// handle's group streams forked from / joined into `st`).  No workspace ordering, no allocation of
// the main workspace: the caller has done both.  Capturable in a CUDA graph once every lazily
// created resource exists (function attributes, scheduler words: after one eager pass).
static int nll_issue(mgp_handle *h, int B, const double *d_params, int n_par, int mode, double noise_var,
                     int eval_grad, double *d_out_llf, double *d_out_grad, int *d_out_status, cudaStream_t st) {
  // isotropic theta: expand to the d-value layout the kernels use, map the gradient back afterwards
  const int n_short = n_par;
  double *d_grad_user = d_out_grad;
  const bool iso = h->isotropic && h->d > 1;
  if (iso) {
    n_par = nll_n_par(mode, h->d);
    double *pf = P<double>(h->niso);
    MGP_CUDA(h, launch_iso_expand(d_params, B, n_short, h->d, pf, st));
    h->launches++;
    d_params = pf;
    d_out_grad = pf + (size_t)B * n_par;
  }
  const int use = h->nll_cap;
  for (int b0 = 0; b0 < B; b0 += use) {
    const int nb = std::min(use, B - b0);
    const int G = nll_groups(h, nb);
    // with concurrent groups no GEMM launch takes every SM: the short kernels of the other groups
    // (diagonal blocks, panel solves) then start without waiting for a machine-wide GEMM to drain
    const int cta_cap = h->gemm_cap_opt > 0 ? h->gemm_cap_opt : (G > 1 ? h->sm_count * 3 / 4 : 0);
    if (G > 1) MGP_CUDA(h, cudaEventRecord(h->ev_fork, st));
    NllWork works[MGP_NLL_STREAMS];
    int gsize[MGP_NLL_STREAMS], gfirst[MGP_NLL_STREAMS];
    int nph = 0;
    for (int g = 0; g < G; g++) {
      const int g0 = (int)((int64_t)nb * g / G), g1 = (int)((int64_t)nb * (g + 1) / G);
      gfirst[g] = g0; gsize[g] = g1 - g0;
      if (g1 == g0) continue;
      if (G > 1) MGP_CUDA(h, cudaStreamWaitEvent(h->nll_stream[g], h->ev_fork, 0));
      fill_work(h, works[g], g0, g1 - g0, use, n_par, mode, noise_var, eval_grad,
                d_params + (size_t)(b0 + g0) * n_par, d_out_llf + b0 + g0,
                eval_grad ? d_out_grad + (size_t)(b0 + g0) * n_par : nullptr);
      works[g].gemm_cta_cap = cta_cap;
      // small matrices: the step is bound by the serial chain of short launches, so the triangular
      // inversion runs block row by block row on a second stream beside the factorisation
      // (measured on B200 with the 32-row tiles of the inversion's products, ms per call after / beside --
      // profiles/r02_nll_chain.txt: n = 500: 0.350 / 0.315 at 4 vectors, 0.933 / 0.931 at 64; n = 1024: 0.700 / 0.601
      // at 1, 0.990 / 0.876 at 8, 3.79 / 3.66 at 64; n = 2048: 1.60 / 1.46 at 1, 2.12 / 2.01 at 4, 3.33 / 3.32 at 8,
      // 21.64 / 21.68 at 64)
      const bool rowwise = h->NI >= 2 && (h->rowinv_opt == 1 ||
                                          (h->rowinv_opt < 0 && h->NI <= 16));   // by size only: a row of the batch
      // gives the same bits whatever the batch size
      works[g].side = h->nll_side[g];
      works[g].ev_side_fork = h->ev_sfork[g];
      works[g].ev_side_join = h->ev_sjoin[g];
      works[g].rowwise = rowwise ? 1 : 0;
      nph = nll_num_phases(works[g]);
    }
    // phase by phase, group by group: every group has its first kernels queued at once (issuing
    // the ~100 launches of one group after the other delayed the last of 8 groups by milliseconds,
    // which a caller that synchronises after every batch -- the L-BFGS-B driver -- pays in full)
    for (int ph = 0; ph < nph; ph++)
      for (int g = 0; g < G; g++) {
        if (gsize[g] == 0) continue;
        MGP_CUDA(h, nll_pipeline(works[g], G > 1 ? h->nll_stream[g] : st, &h->launches, G > 1 ? ph : -1));
        if (G == 1) ph = nph;   // a single group runs the whole pipeline in one call
      }
    for (int g = 0; g < G; g++) {
      if (gsize[g] == 0) continue;
      cudaStream_t gs = G > 1 ? h->nll_stream[g] : st;
      if (d_out_status)
        MGP_CUDA(h, cudaMemcpyAsync(d_out_status + b0 + gfirst[g], works[g].status, sizeof(int) * gsize[g],
                                    cudaMemcpyDeviceToDevice, gs));
      if (G > 1) {
        MGP_CUDA(h, cudaEventRecord(h->ev_join[g], gs));
        MGP_CUDA(h, cudaStreamWaitEvent(st, h->ev_join[g], 0));
      }
    }
  }
  if (iso && eval_grad) {
    MGP_CUDA(h, launch_iso_gather(d_out_grad, B, n_short, n_par, h->d, mode, d_grad_user, st));
    h->launches++;
  }
  return MGP_OK;
}

static void nll_graphs_clear(mgp_handle *h) {
  for (auto &g : h->nll_graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  h->nll_graphs.clear();
}

// Staging area of the graph path: [params B x n_par | llf B | grad B x n_par | status B ints].
static int nll_staging(mgp_handle *h, int B, int n_par, double **par, double **llf, double **grad, int **status) {
  const size_t need = ((size_t)2 * B * n_par + B) * 8 + (size_t)B * sizeof(int) + 256;
  if (h->ngio.bytes < need) {
    // a larger staging area invalidates the graphs that point into the old one
    nll_graphs_clear(h);
    MGP_TRY(mgp_ensure(h, h->ngio, std::max(need, (size_t)1 << 16)));
  }
  *par = P<double>(h->ngio);
  *grad = *par + (size_t)B * n_par;
  *llf = *grad + (size_t)B * n_par;
  *status = reinterpret_cast<int *>(*llf + B);
  return 0;
}

int mgp_nll_batch_dev(mgp_t *h, int B, const double *d_params, int n_par, int mode,
                      double noise_var, int eval_grad, double *d_out_llf, double *d_out_grad,
                      int *d_out_status, void *stream) {
  CHECK_H(h);
  MGP_TRY(check_par(h, n_par, mode));
  if (B < 1 || !d_params || !d_out_llf || (eval_grad && !d_out_grad))
    return mgp_fail(h, MGP_EINVAL, "bad arguments to mgp_nll_batch");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (B > h->nll_cap) {   // grow the workspace (cudaMemGetInfo is slow: only asked when growing)
    const int cap = std::min(B, nll_max_batch(h));
    if (cap > h->nll_cap) {
      nll_graphs_clear(h);      // the workspace moves
      MGP_TRY(nll_alloc(h, cap));
    }
  }
  if (h->isotropic && h->d > 1) {
    const size_t need = (size_t)2 * B * nll_n_par(mode, h->d) * 8;
    if (h->niso.bytes < need) {
      nll_graphs_clear(h);      // graphs of smaller batches point into the old buffer
      MGP_TRY(mgp_ensure(h, h->niso, need));
    }
  }
  WsScope ws(h, st);
  MGP_CUDA(h, ws.err);
  // ---- CUDA-graph path: shapes that fit the workspace in one pass, on a capturable stream -----------
  const bool graphable = h->graphs_opt && !h->time_kernels && B <= h->nll_cap && st != nullptr &&
                         st != cudaStreamLegacy && st != cudaStreamPerThread;
  if (!graphable)
    return nll_issue(h, B, d_params, n_par, mode, noise_var, eval_grad, d_out_llf, d_out_grad, d_out_status, st);
  const int G = nll_groups(h, B);
  mgp_handle::NllGraph key = {B, n_par, mode, eval_grad, h->n, h->npad, G, h->gemm_cap_opt, h->isotropic ? 1 : 0,
                              noise_var, nullptr, 0, 0};
  mgp_handle::NllGraph *rec = nullptr;
  for (auto &g : h->nll_graphs)
    if (g.B == key.B && g.n_par == key.n_par && g.mode == key.mode && g.eval_grad == key.eval_grad && g.n == key.n &&
        g.npad == key.npad && g.groups == key.groups && g.cap == key.cap && g.iso == key.iso &&
        g.noise_var == key.noise_var) { rec = &g; break; }
  if (!rec) {
    // first sight of this shape: eager pass (creates function attributes, scheduler words, ...)
    if (h->nll_graphs.size() >= 16) nll_graphs_clear(h);
    h->nll_graphs.push_back(key);
    return nll_issue(h, B, d_params, n_par, mode, noise_var, eval_grad, d_out_llf, d_out_grad, d_out_status, st);
  }
  double *sp, *sl, *sg;
  int *ss;
  {
    const size_t before = h->nll_graphs.size();
    MGP_TRY(nll_staging(h, B, n_par, &sp, &sl, &sg, &ss));
    if (h->nll_graphs.size() != before) {      // staging grew: every record is gone, start over for this shape
      h->nll_graphs.push_back(key);
      rec = &h->nll_graphs.back();
    }
  }
  const bool direct = d_params == sp;          // the host-pointer entry point stages into ngio itself
  if (!direct)
    MGP_CUDA(h, cudaMemcpyAsync(sp, d_params, (size_t)B * n_par * 8, cudaMemcpyDeviceToDevice, st));
  if (!rec->exec) {
    const int64_t l0 = h->launches;
    cudaGraph_t graph = nullptr;
    MGP_CUDA(h, cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    const int rc = nll_issue(h, B, sp, n_par, mode, noise_var, eval_grad, sl, sg, ss, st);
    const cudaError_t ce = cudaStreamEndCapture(st, &graph);
    if (rc != 0 || ce != cudaSuccess || !graph) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      h->graphs_opt = 0;         // not capturable here: stay on the eager path
      h->launches = l0;
      if (!direct) {
        MGP_TRY(nll_issue(h, B, d_params, n_par, mode, noise_var, eval_grad, d_out_llf, d_out_grad, d_out_status, st));
        return MGP_OK;
      }
      return nll_issue(h, B, sp, n_par, mode, noise_var, eval_grad, sl, sg, ss, st);
    }
    const cudaError_t ie = cudaGraphInstantiate(&rec->exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) { rec->exec = nullptr; h->graphs_opt = 0; return mgp_cuda_fail(h, ie, "cudaGraphInstantiate"); }
    rec->launches = h->launches - l0;
    h->launches = l0;
  }
  MGP_CUDA(h, cudaGraphLaunch(rec->exec, st));
  rec->uses++;
  h->graph_replays++;
  h->launches += rec->launches;
  if (!direct) {
    MGP_CUDA(h, cudaMemcpyAsync(d_out_llf, sl, (size_t)B * 8, cudaMemcpyDeviceToDevice, st));
    if (eval_grad) MGP_CUDA(h, cudaMemcpyAsync(d_out_grad, sg, (size_t)B * n_par * 8, cudaMemcpyDeviceToDevice, st));
    if (d_out_status) MGP_CUDA(h, cudaMemcpyAsync(d_out_status, ss, (size_t)B * sizeof(int), cudaMemcpyDeviceToDevice, st));
  }
  return MGP_OK;
}

int mgp_nll_batch(mgp_t *h, int B, const double *params, int n_par, int mode, double noise_var,
                  int eval_grad, double *out_llf, double *out_grad, int *out_status) {
  CHECK_H(h);
  MGP_TRY(check_par(h, n_par, mode));
  if (B < 1 || !params || !out_llf || (eval_grad && !out_grad))
    return mgp_fail(h, MGP_EINVAL, "bad arguments to mgp_nll_batch");
  const size_t pb = (size_t)B * n_par * 8;
  // parameters and results go through the staging area the likelihood graphs are wired to
  double *d_par, *d_llf, *d_grad;
  int *d_status;
  MGP_TRY(nll_staging(h, B, n_par, &d_par, &d_llf, &d_grad, &d_status));
  {
    WsScope ws(h, h->stream);     // a graph replay on another stream may still be reading the staging area
    MGP_CUDA(h, ws.err);
    MGP_CUDA(h, cudaMemcpyAsync(d_par, params, pb, cudaMemcpyHostToDevice, h->stream));
  }
  MGP_TRY(mgp_nll_batch_dev(h, B, d_par, n_par, mode, noise_var, eval_grad, d_llf, d_grad, d_status,
                            h->stream));
  MGP_CUDA(h, cudaMemcpyAsync(out_llf, d_llf, (size_t)B * 8, cudaMemcpyDeviceToHost, h->stream));
  if (eval_grad)
    MGP_CUDA(h, cudaMemcpyAsync(out_grad, d_grad, pb, cudaMemcpyDeviceToHost, h->stream));
  if (out_status)
    MGP_CUDA(h, cudaMemcpyAsync(out_status, d_status, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  MGP_CUDA(h, cudaStreamSynchronize(h->stream));
  return MGP_OK;
}

// ---------------------------------------------------------------------------------------------
// fitted state
// ---------------------------------------------------------------------------------------------
int mgp_model_alloc(mgp_handle *h) {
  const size_t mat = (size_t)h->npad * h->npad * 8, vec = (size_t)h->npad * 8;
  MGP_TRY(mgp_ensure(h, h->dL, mat));
  MGP_TRY(mgp_ensure(h, h->dMinv, mat));
  MGP_TRY(mgp_ensure(h, h->dMp, (size_t)8 * h->NI * (h->NI + 1) * 1024 * 8));
  MGP_TRY(mgp_ensure(h, h->dtheta, (size_t)h->dpad * 8));
  MGP_TRY(mgp_ensure(h, h->dBop, (size_t)h->npad * h->dpad * 8));
  MGP_TRY(mgp_ensure(h, h->dan, vec));
  MGP_TRY(mgp_ensure(h, h->dXsc, (size_t)h->npad * grad_row_stride(h->dpad) * 8));
  MGP_TRY(mgp_ensure(h, h->dgamma, vec));
  MGP_TRY(mgp_ensure(h, h->davec, vec));
  MGP_TRY(mgp_ensure(h, h->dFt, vec));
  MGP_TRY(mgp_ensure(h, h->dYt, vec));
  MGP_TRY(mgp_ensure(h, h->drho, vec));
  return 0;
}

// Common tail of finalize_fit / set_state: the NLL workspace slot 0 holds L, M, Yt, Ft, rho, gamma.
static int adopt_slot0(mgp_handle *h, const double *theta_host, bool take_gamma) {
  cudaStream_t st = h->stream;
  const size_t mat = (size_t)h->npad * h->npad * 8, vec = (size_t)h->npad * 8;
  const double *v = P<double>(h->nvec);
  const int B = h->nll_cap;  // section stride of the vector workspace; slot 0 is used
  MGP_CUDA(h, cudaMemcpyAsync(h->dL.p, h->nL.p, mat, cudaMemcpyDeviceToDevice, st));
  MGP_CUDA(h, cudaMemcpyAsync(h->dMinv.p, h->nM.p, mat, cudaMemcpyDeviceToDevice, st));
  MGP_CUDA(h, cudaMemcpyAsync(h->dYt.p, v, vec, cudaMemcpyDeviceToDevice, st));
  MGP_CUDA(h, cudaMemcpyAsync(h->dFt.p, v + (size_t)B * h->npad, vec, cudaMemcpyDeviceToDevice, st));
  MGP_CUDA(h, cudaMemcpyAsync(h->drho.p, v + (size_t)2 * B * h->npad, vec, cudaMemcpyDeviceToDevice, st));
  if (take_gamma)
    MGP_CUDA(h, cudaMemcpyAsync(h->dgamma.p, v + (size_t)3 * B * h->npad, vec, cudaMemcpyDeviceToDevice, st));
  // a = L^-T Ft  (so that Ft^T L^-1 r = a . r)
  MGP_CUDA(h, launch_trmv(P<double>(h->dMinv), h->npad, 0, P<double>(h->dFt), 0, P<double>(h->davec), 0,
                          h->n, 1, 1, st));
  if (h->p > 1) {
    MGP_TRY(mgp_ensure(h, h->dbetav, MGP_PMAX * 8));
    if (h->trend_est) {
      const size_t tb = (size_t)h->npad * MGP_TP * 8;
      MGP_TRY(mgp_ensure(h, h->dBhat, tb));
      MGP_TRY(mgp_ensure(h, h->dFtm, tb));
      MGP_TRY(mgp_ensure(h, h->dCinv, (size_t)MGP_PMAX * MGP_PMAX * 8));
      const double *tw = P<double>(h->ntws);
      MGP_CUDA(h, cudaMemcpyAsync(h->dBhat.p, h->nBhat.p, tb, cudaMemcpyDeviceToDevice, st));
      MGP_CUDA(h, cudaMemcpyAsync(h->dFtm.p, h->nFtm.p, tb, cudaMemcpyDeviceToDevice, st));
      MGP_CUDA(h, cudaMemcpyAsync(h->dCinv.p, tw + TWS_CINV, (size_t)MGP_PMAX * MGP_PMAX * 8, cudaMemcpyDeviceToDevice, st));
      if (take_gamma) {   // coefficients estimated by this run (not adopted from outside)
        MGP_CUDA(h, cudaMemcpyAsync(h->dbetav.p, tw + TWS_BETA, MGP_PMAX * 8, cudaMemcpyDeviceToDevice, st));
        h->hbeta.assign(MGP_PMAX, 0.0);
        MGP_CUDA(h, cudaMemcpyAsync(h->hbeta.data(), tw + TWS_BETA, MGP_PMAX * 8, cudaMemcpyDeviceToHost, st));
      }
    } else {
      h->hbeta = h->beta_in;
      h->hbeta.resize(MGP_PMAX, 0.0);
      MGP_CUDA(h, cudaMemcpyAsync(h->dbetav.p, h->hbeta.data(), MGP_PMAX * 8, cudaMemcpyHostToDevice, st));
    }
  }
  std::vector<double> th(h->dpad, 0.0);
  for (int i = 0; i < h->d; i++) th[i] = theta_host[i];
  MGP_CUDA(h, cudaMemcpyAsync(h->dtheta.p, th.data(), th.size() * 8, cudaMemcpyHostToDevice, st));
  MGP_CUDA(h, launch_prep_model(P<double>(h->dXc), P<double>(h->dtheta), h->npad, h->dpad,
                                P<double>(h->dBop), P<double>(h->dan), P<double>(h->dXsc), h->corr, st));
  MGP_CUDA(h, launch_pack_tri(P<double>(h->dMinv), h->n, h->npad, P<double>(h->dMp), st));
  h->launches += 3;
  MGP_CUDA(h, cudaStreamSynchronize(st));  // th is a host temporary
  h->rinv_ready = false;
  h->i8_ready = false;
  return 0;
}

int mgp_finalize_fit(mgp_t *h, const double *par, int n_par, int mode, double noise_var,
                     double *out_llf) {
  CHECK_H(h);
  if (!par) return mgp_fail(h, MGP_EINVAL, "par is NULL");
  MGP_TRY(check_par(h, n_par, mode));
  std::vector<double> pfull;
  if (h->isotropic && h->d > 1) {
    iso_expand_host(h, par, n_par, pfull);
    par = pfull.data();
    n_par = (int)pfull.size();
  }
  MGP_TRY(mgp_model_alloc(h));
  if (h->nll_cap < 1) MGP_TRY(nll_alloc(h, 1));
  MGP_TRY(mgp_ensure(h, h->npar, (size_t)n_par * 8 + 64));
  cudaStream_t st = h->stream;
  WsScope ws(h, st);
  MGP_CUDA(h, ws.err);
  MGP_CUDA(h, cudaMemcpyAsync(h->npar.p, par, (size_t)n_par * 8, cudaMemcpyHostToDevice, st));
  NllWork w;
  fill_work(h, w, 0, 1, h->nll_cap, n_par, mode, noise_var, 0, P<double>(h->npar), nullptr, nullptr);
  w.want_bhat = 1;
  MGP_CUDA(h, nll_pipeline(w, st, &h->launches));
  std::vector<double> sc(nll_scal_doubles());
  int status = 0;
  MGP_CUDA(h, cudaMemcpyAsync(sc.data(), w.scal, sc.size() * 8, cudaMemcpyDeviceToHost, st));
  MGP_CUDA(h, cudaMemcpyAsync(&status, w.status, sizeof(int), cudaMemcpyDeviceToHost, st));
  MGP_CUDA(h, cudaStreamSynchronize(st));
  if (status > 0) {
    h->fitted = false;
    return mgp_fail(h, MGP_ENOTPD, "correlation matrix not positive definite at pivot %d", status);
  }
  h->mode = mode;
  h->noise_var = sc[NLL_NV];
  h->sigma2 = sc[NLL_SIGMA2];
  h->llf = sc[NLL_LLF];
  h->ftf = sc[NLL_FTF];
  h->G = -sqrt(sc[NLL_FTF]);  // LAPACK's Householder QR of the positive-leading column Ft
  h->beta = sc[NLL_BETA];
  MGP_TRY(adopt_slot0(h, par, true));
  h->hpar.assign(par, par + n_par);
  h->par_noise_var = noise_var;
  h->fitted = true;
  if (out_llf) *out_llf = h->llf;
  return MGP_OK;
}

// Rank-k extension of the fitted model by k appended training rows at the CURRENT hyper-parameters
// (the incremental training gpr.py:387-390 asks for).  With nb = the last multiple of 128 <= n, the
// leading nb x nb blocks L11, M11 = L11^-1 stay; the T = (npad' - nb) / 128 trailing tile rows are
// recomputed:  L21 = R21 M11^T,  S = R22 - L21 L21^T,  L22 = chol(S),  M22 = L22^-1,
// M21 = -M22 (L21 M11)  -- three GEMMs of T*128 x nb x nb, i.e. O(n^2 k) instead of the 2/3 n^3 of a
// rebuild -- then the O(n^2) tail of the likelihood pipeline (Yt, Ft, rho, sigma2, beta, gamma, llf)
// and the packing of the posterior operands.
int mgp_append_train(mgp_t *h, int k, const double *X_add, const double *y_add, double *out_llf) {
  CHECK_H(h);
  if (!h->fitted) return mgp_fail(h, MGP_ESTATE, "model is not fitted");
  if (k < 1 || !X_add || !y_add) return mgp_fail(h, MGP_EINVAL, "bad rows to append");
  if (h->hpar.empty()) return mgp_fail(h, MGP_ESTATE, "append: the model was adopted with mgp_set_state, its hyper-parameters are unknown");
  if (h->p > 1) return mgp_fail(h, MGP_ESTATE, "append: only the constant trend is extended incrementally");
  const int n0 = h->n, npad0 = h->npad, d = h->d, dpad = h->dpad;
  const int n1 = n0 + k, npad1 = (int)round_up(n1, MGP_TILE), NI1 = npad1 / MGP_TILE;
  const int nb = (n0 / MGP_TILE) * MGP_TILE;   // rows [0, nb) are complete tiles and keep their factor
  const int tail = npad1 - nb;                 // recomputed rows (a multiple of 128)
  if (nb < 512 || tail * 4 > nb)
    return mgp_fail(h, MGP_ESTATE, "append: the leading block (%d rows) is too small for %d new rows -- refit instead", nb, k);
  cudaStream_t st = h->stream;
  WsScope ws(h, st);
  MGP_CUDA(h, ws.err);
  // ---- new training buffers (old centring kept: differences are translation invariant) -----------
  DevBuf nXc, nY;
  MGP_TRY(mgp_ensure(h, nXc, (size_t)npad1 * dpad * 8));
  MGP_TRY(mgp_ensure(h, nY, (size_t)npad1 * 8));
  auto drop_new = [&]() { freebuf(nXc); freebuf(nY); };
  std::vector<double> Xa((size_t)k * dpad, 0.0);
  for (int j = 0; j < k; j++)
    for (int i = 0; i < d; i++) Xa[(size_t)j * dpad + i] = X_add[(size_t)j * d + i] - h->hcenter[i];
  cudaError_t e = cudaMemsetAsync(nXc.p, 0, (size_t)npad1 * dpad * 8, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(nY.p, 0, (size_t)npad1 * 8, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(nXc.p, h->dXc.p, (size_t)n0 * dpad * 8, cudaMemcpyDeviceToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(P<double>(nXc) + (size_t)n0 * dpad, Xa.data(), Xa.size() * 8, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(nY.p, h->dy.p, (size_t)n0 * 8, cudaMemcpyDeviceToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(P<double>(nY) + n0, y_add, (size_t)k * 8, cudaMemcpyHostToDevice, st);
  int *flag = P<int>(h->dscal);
  if (e == cudaSuccess) e = cudaMemsetAsync(flag, 0, 2 * sizeof(int), st);
  if (e == cudaSuccess) e = launch_dup_check(P<double>(nXc), n1, dpad, flag, st);
  int hflag = 0;
  if (e == cudaSuccess) e = cudaMemcpyAsync(&hflag, flag, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) { drop_new(); return mgp_cuda_fail(h, e, "append: staging the new rows"); }
  h->launches++;
  if (hflag) { drop_new(); return mgp_fail(h, MGP_EINVAL, "duplicate rows: Multiple input features cannot have the same target value."); }
  // ---- switch the handle to the new geometry (restored if the extension fails) ---------------------
  std::swap(h->dXc, nXc); std::swap(h->dy, nY);
  h->n = n1; h->npad = npad1; h->NI = NI1;
  auto restore = [&]() {
    std::swap(h->dXc, nXc); std::swap(h->dy, nY);
    h->n = n0; h->npad = npad0; h->NI = npad0 / MGP_TILE;
    h->nll_cap = 0;
    drop_new();
  };
  if (npad1 != npad0) h->nll_cap = 0;
  int rc = h->nll_cap < 1 ? nll_alloc(h, 1) : 0;
  if (rc == 0) rc = mgp_ensure(h, h->npar, h->hpar.size() * 8 + 64);
  if (rc != 0) { restore(); return rc; }
  const int n_par = (int)h->hpar.size();
  NllWork w;
  fill_work(h, w, 0, 1, h->nll_cap, n_par, h->mode, h->par_noise_var, 0, P<double>(h->npar), nullptr, nullptr);
  w.have_L = 1; w.have_M = 1; w.want_bhat = 1;
  const size_t ld = (size_t)npad1;
  int *chol_status = flag + 1;
#define AP_CUDA(call)                                                       \
  do {                                                                      \
    cudaError_t e__ = (call);                                               \
    if (e__ != cudaSuccess) { cudaStreamSynchronize(st); restore(); return mgp_cuda_fail(h, e__, #call); } \
  } while (0)
  AP_CUDA(cudaMemcpyAsync(h->npar.p, h->hpar.data(), (size_t)n_par * 8, cudaMemcpyHostToDevice, st));
  // old factor and inverse into the workspace slot at the new leading dimension
  if (npad1 == npad0) {
    AP_CUDA(cudaMemcpyAsync(w.L, h->dL.p, ld * ld * 8, cudaMemcpyDeviceToDevice, st));
    AP_CUDA(cudaMemcpyAsync(w.M, h->dMinv.p, ld * ld * 8, cudaMemcpyDeviceToDevice, st));
  } else {
    AP_CUDA(cudaMemsetAsync(w.L, 0, ld * ld * 8, st));
    AP_CUDA(cudaMemsetAsync(w.M, 0, ld * ld * 8, st));
    AP_CUDA(cudaMemcpy2DAsync(w.L, ld * 8, h->dL.p, (size_t)npad0 * 8, (size_t)nb * 8, nb, cudaMemcpyDeviceToDevice, st));
    AP_CUDA(cudaMemcpy2DAsync(w.M, ld * 8, h->dMinv.p, (size_t)npad0 * 8, (size_t)nb * 8, nb, cudaMemcpyDeviceToDevice, st));
  }
  // R, tile rows >= nb / 128 (lower block triangle)
  AP_CUDA(launch_build_R(w, nb / MGP_TILE, st));
  double *Rt = w.R + (size_t)nb * ld, *Lt = w.L + (size_t)nb * ld, *Mt = w.M + (size_t)nb * ld;
  double *Tt = w.T + (size_t)nb * ld;                       // tail rows of the scratch matrix: L21 M11
  double *Sc = w.T, *L22 = Sc + (size_t)tail * tail, *M22 = L22 + (size_t)tail * tail, *Ts = M22 + (size_t)tail * tail;
  {  // L21 = R21 M11^T   (M11 stored [N][K]: B(k, j) = M11[j][k]; its zero half is multiplied too)
    GemmDesc g = {};
    g.A = Rt; g.lda = (int)ld; g.ta = 0; g.B = w.M; g.ldb = (int)ld; g.tb = 0; g.C = Lt; g.ldc = (int)ld;
    g.M = tail; g.N = nb; g.K = nb; g.alpha = 1.0; g.beta = 0.0; g.n_inner = 1; g.n_outer = 1;
    AP_CUDA(launch_dgemm(g, st));
  }
  {  // S = R22 - L21 L21^T  (lower tiles), in place in R
    GemmDesc g = {};
    g.A = Lt; g.lda = (int)ld; g.ta = 0; g.B = Lt; g.ldb = (int)ld; g.tb = 0; g.C = Rt + nb; g.ldc = (int)ld;
    g.M = tail; g.N = tail; g.K = nb; g.alpha = -1.0; g.beta = 1.0; g.lower_only = 1; g.n_inner = 1; g.n_outer = 1;
    AP_CUDA(launch_dgemm(g, st));
  }
  // L22 = chol(S), M22 = L22^-1 on a compact tail x tail copy
  AP_CUDA(cudaMemcpy2DAsync(Sc, (size_t)tail * 8, Rt + nb, ld * 8, (size_t)tail * 8, tail, cudaMemcpyDeviceToDevice, st));
  AP_CUDA(cudaMemsetAsync(chol_status, 0, sizeof(int), st));
  AP_CUDA(chol_blocked(Sc, L22, tail, 0, w.linv, 0, chol_status, 1, st, &h->launches, true, 0, chol_panels(tail)));
  AP_CUDA(trtri_blocked(L22, M22, Ts, tail, 0, w.linv, 0, 1, st, &h->launches, true));
  AP_CUDA(cudaMemcpy2DAsync(Lt + nb, ld * 8, L22, (size_t)tail * 8, (size_t)tail * 8, tail, cudaMemcpyDeviceToDevice, st));
  AP_CUDA(cudaMemcpy2DAsync(Mt + nb, ld * 8, M22, (size_t)tail * 8, (size_t)tail * 8, tail, cudaMemcpyDeviceToDevice, st));
  {  // Tt = L21 M11   (M11 stored [K][N], lower triangular: k starts at the column block)
    GemmDesc g = {};
    g.A = Lt; g.lda = (int)ld; g.ta = 0; g.B = w.M; g.ldb = (int)ld; g.tb = 1; g.C = Tt; g.ldc = (int)ld;
    g.M = tail; g.N = nb; g.K = nb; g.alpha = 1.0; g.beta = 0.0; g.kmode = 1; g.order = 3; g.n_inner = 1; g.n_outer = 1;
    AP_CUDA(launch_dgemm(g, st));
  }
  {  // M21 = -M22 Tt   (M22 lower triangular: k ends with the row block)
    GemmDesc g = {};
    g.A = Mt + nb; g.lda = (int)ld; g.ta = 0; g.B = Tt; g.ldb = (int)ld; g.tb = 1; g.C = Mt; g.ldc = (int)ld;
    g.M = tail; g.N = nb; g.K = tail; g.alpha = -1.0; g.beta = 0.0; g.kmode = 2; g.order = 2; g.n_inner = 1; g.n_outer = 1;
    AP_CUDA(launch_dgemm(g, st));
  }
  h->launches += 5;
  AP_CUDA(nll_pipeline(w, st, &h->launches));
  std::vector<double> sc(nll_scal_doubles());
  int status = 0, cstatus = 0;
  AP_CUDA(cudaMemcpyAsync(sc.data(), w.scal, sc.size() * 8, cudaMemcpyDeviceToHost, st));
  AP_CUDA(cudaMemcpyAsync(&status, w.status, sizeof(int), cudaMemcpyDeviceToHost, st));
  AP_CUDA(cudaMemcpyAsync(&cstatus, chol_status, sizeof(int), cudaMemcpyDeviceToHost, st));
  AP_CUDA(cudaStreamSynchronize(st));
#undef AP_CUDA
  if (cstatus > 0 || status > 0) {
    restore();
    return mgp_fail(h, MGP_ENOTPD, "append: extended correlation matrix not positive definite at pivot %d",
                    cstatus > 0 ? nb + cstatus : status);
  }
  h->hX.insert(h->hX.end(), X_add, X_add + (size_t)k * d);
  h->hy.insert(h->hy.end(), y_add, y_add + k);
  drop_new();                        // the old training buffers
  h->noise_var = sc[NLL_NV];
  h->sigma2 = sc[NLL_SIGMA2];
  h->llf = sc[NLL_LLF];
  h->ftf = sc[NLL_FTF];
  h->G = -sqrt(sc[NLL_FTF]);
  h->beta = sc[NLL_BETA];
  MGP_TRY(mgp_model_alloc(h));
  MGP_TRY(adopt_slot0(h, h->hpar.data(), true));
  h->appends++;
  if (out_llf) *out_llf = h->llf;
  return MGP_OK;
}

int mgp_nll_env(mgp_t *h, const double *par, int n_par, int mode, double noise_var, double *out_L,
                double *out_rho, double *out_Yt, double *out_Ft, double *out_scalars) {
  CHECK_H(h);
  if (!par) return mgp_fail(h, MGP_EINVAL, "par is NULL");
  MGP_TRY(check_par(h, n_par, mode));
  std::vector<double> pfull;
  if (h->isotropic && h->d > 1) {
    iso_expand_host(h, par, n_par, pfull);
    par = pfull.data();
    n_par = (int)pfull.size();
  }
  if (h->nll_cap < 1) MGP_TRY(nll_alloc(h, 1));
  MGP_TRY(mgp_ensure(h, h->npar, (size_t)n_par * 8 + 64));
  cudaStream_t st = h->stream;
  WsScope ws(h, st);
  MGP_CUDA(h, ws.err);
  MGP_CUDA(h, cudaMemcpyAsync(h->npar.p, par, (size_t)n_par * 8, cudaMemcpyHostToDevice, st));
  NllWork w;
  fill_work(h, w, 0, 1, h->nll_cap, n_par, mode, noise_var, 0, P<double>(h->npar), nullptr, nullptr);
  w.want_bhat = 1;   // exported factor: zero the tiles above the block diagonal
  MGP_CUDA(h, nll_pipeline(w, st, &h->launches));
  std::vector<double> sc(nll_scal_doubles());
  int status = 0;
  const int n = h->n;
  MGP_CUDA(h, cudaMemcpyAsync(sc.data(), w.scal, sc.size() * 8, cudaMemcpyDeviceToHost, st));
  MGP_CUDA(h, cudaMemcpyAsync(&status, w.status, sizeof(int), cudaMemcpyDeviceToHost, st));
  if (out_L)
    MGP_CUDA(h, cudaMemcpy2DAsync(out_L, (size_t)n * 8, w.L, (size_t)h->npad * 8, (size_t)n * 8, n,
                                  cudaMemcpyDeviceToHost, st));
  if (out_rho) MGP_CUDA(h, cudaMemcpyAsync(out_rho, w.rho, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
  if (out_Yt) MGP_CUDA(h, cudaMemcpyAsync(out_Yt, w.Yt, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
  if (out_Ft) {
    if (h->p > 1 && h->trend_est)
      MGP_CUDA(h, cudaMemcpy2DAsync(out_Ft, (size_t)h->p * 8, w.Ftm, (size_t)MGP_TP * 8, (size_t)h->p * 8, n,
                                    cudaMemcpyDeviceToHost, st));
    else
      MGP_CUDA(h, cudaMemcpyAsync(out_Ft, w.Ft, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
  }
  MGP_CUDA(h, cudaStreamSynchronize(st));
  if (status > 0) return mgp_fail(h, MGP_ENOTPD, "correlation matrix not positive definite at pivot %d", status);
  if (out_scalars) {
    out_scalars[0] = sc[NLL_SIGMA2]; out_scalars[1] = sc[NLL_BETA]; out_scalars[2] = -sqrt(sc[NLL_FTF]);
    out_scalars[3] = sc[NLL_LLF]; out_scalars[4] = sc[NLL_NV];
  }
  return MGP_OK;
}

int mgp_get_trend(mgp_t *h, double *out_beta) {
  CHECK_H(h);
  if (!h->fitted) return mgp_fail(h, MGP_ESTATE, "model is not fitted");
  if (!out_beta) return mgp_fail(h, MGP_EINVAL, "out_beta is NULL");
  if (h->p == 1) { out_beta[0] = h->beta; return MGP_OK; }
  MGP_CUDA(h, cudaStreamSynchronize(h->stream));
  for (int i = 0; i < h->p; i++) out_beta[i] = h->hbeta[i];
  return MGP_OK;
}

int mgp_set_state(mgp_t *h, const double *theta, const double *L, const double *gamma,
                  double sigma2, const double *beta) {
  CHECK_H(h);
  if (h->n == 0) return mgp_fail(h, MGP_ESTATE, "mgp_set_train must be called first");
  if (!theta || !L || !gamma || !beta) return mgp_fail(h, MGP_EINVAL, "NULL state array");
  MGP_TRY(mgp_model_alloc(h));
  if (h->nll_cap < 1) MGP_TRY(nll_alloc(h, 1));
  cudaStream_t st = h->stream;
  WsScope ws(h, st);
  MGP_CUDA(h, ws.err);
  const int n = h->n, npad = h->npad;
  // L -> slot 0 (identity on the padding)
  std::vector<double> Lp((size_t)npad * npad, 0.0), gp(npad, 0.0);
  for (int j = 0; j < n; j++)
    for (int k = 0; k <= j; k++) Lp[(size_t)j * npad + k] = L[(size_t)j * n + k];
  for (int j = n; j < npad; j++) Lp[(size_t)j * npad + j] = 1.0;
  for (int j = 0; j < n; j++) gp[j] = gamma[j];
  MGP_CUDA(h, cudaMemcpyAsync(h->nL.p, Lp.data(), Lp.size() * 8, cudaMemcpyHostToDevice, st));
  MGP_CUDA(h, cudaMemcpyAsync(h->dgamma.p, gp.data(), gp.size() * 8, cudaMemcpyHostToDevice, st));
  std::vector<double> dummy(h->d + 1, 1.0);
  MGP_TRY(mgp_ensure(h, h->npar, dummy.size() * 8 + 64));
  MGP_CUDA(h, cudaMemcpyAsync(h->npar.p, dummy.data(), dummy.size() * 8, cudaMemcpyHostToDevice, st));
  NllWork w;
  fill_work(h, w, 0, 1, h->nll_cap, h->d, MGP_MODE_NOISELESS, 0.0, 0, P<double>(h->npar), nullptr, nullptr);
  w.have_L = 1;
  w.want_bhat = 1;
  MGP_CUDA(h, nll_pipeline(w, st, &h->launches));
  std::vector<double> sc(nll_scal_doubles());
  MGP_CUDA(h, cudaMemcpyAsync(sc.data(), w.scal, sc.size() * 8, cudaMemcpyDeviceToHost, st));
  MGP_CUDA(h, cudaStreamSynchronize(st));
  h->sigma2 = sigma2;
  h->beta = beta[0];
  h->ftf = sc[NLL_FTF];
  h->G = -sqrt(sc[NLL_FTF]);
  h->llf = NAN;
  h->hpar.clear();
  MGP_TRY(adopt_slot0(h, theta, false));
  if (h->p > 1) {   // the coefficients travel with the state (they are what the sender fitted)
    h->hbeta.assign(MGP_PMAX, 0.0);
    for (int i = 0; i < h->p; i++) h->hbeta[i] = beta[i];
    MGP_CUDA(h, cudaMemcpyAsync(h->dbetav.p, h->hbeta.data(), MGP_PMAX * 8, cudaMemcpyHostToDevice, st));
    MGP_CUDA(h, cudaStreamSynchronize(st));
  }
  h->fitted = true;
  return MGP_OK;
}

int mgp_get_state(mgp_t *h, double *out_theta, double *out_L, double *out_gamma, double *out_rho,
                  double *out_Yt, double *out_Ft, double *out_scalars) {
  CHECK_H(h);
  if (!h->fitted) return mgp_fail(h, MGP_ESTATE, "model is not fitted");
  cudaStream_t st = h->stream;
  WsScope ws(h, st);
  MGP_CUDA(h, ws.err);
  const int n = h->n;
  if (out_theta) MGP_CUDA(h, cudaMemcpyAsync(out_theta, h->dtheta.p, (size_t)h->d * 8, cudaMemcpyDeviceToHost, st));
  if (out_L)
    MGP_CUDA(h, cudaMemcpy2DAsync(out_L, (size_t)n * 8, h->dL.p, (size_t)h->npad * 8, (size_t)n * 8, n,
                                  cudaMemcpyDeviceToHost, st));
  if (out_gamma) MGP_CUDA(h, cudaMemcpyAsync(out_gamma, h->dgamma.p, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
  if (out_rho) MGP_CUDA(h, cudaMemcpyAsync(out_rho, h->drho.p, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
  if (out_Yt) MGP_CUDA(h, cudaMemcpyAsync(out_Yt, h->dYt.p, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
  if (out_Ft) {
    if (h->p > 1 && h->trend_est)
      MGP_CUDA(h, cudaMemcpy2DAsync(out_Ft, (size_t)h->p * 8, h->dFtm.p, (size_t)MGP_TP * 8, (size_t)h->p * 8, n,
                                    cudaMemcpyDeviceToHost, st));
    else
      MGP_CUDA(h, cudaMemcpyAsync(out_Ft, h->dFt.p, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
  }
  MGP_CUDA(h, cudaStreamSynchronize(st));
  if (out_scalars) {
    out_scalars[0] = h->sigma2; out_scalars[1] = h->beta; out_scalars[2] = h->G; out_scalars[3] = h->llf;
    out_scalars[4] = h->noise_var;
  }
  return MGP_OK;
}

// ---------------------------------------------------------------------------------------------
// posterior pipeline
// ---------------------------------------------------------------------------------------------
struct PostReq {
  int64_t m;
  const double *d_Xs;
  double *d_mean, *d_mse, *d_val, *d_mean_dx, *d_mse_dx, *d_dx;
  int kind, n_sets, minimize, want_dx, topk;
  double plugin;
  const double *params;
  double *d_topv;
  int64_t *d_topi;
  int64_t ld_val;     // elements between parameter sets in d_val (x d in d_dx); 0 -> m
  int64_t idx_base;   // global index of row 0 (top-k over several calls)
  int topk_continue;  // 1: merge into the existing running top-k instead of resetting it
};

static int post_workspace(mgp_handle *h, bool want_dx) {
  // default: eight candidate blocks per SM per pass (amortises launch ramps and the small
  // epilogue / top-k kernels), bounded so that the r scratch stays below 2 GiB
  int64_t chunk = (int64_t)h->sm_count * 128 * 8;
  while (chunk > (int64_t)h->sm_count * 128 && (size_t)h->npad * chunk * 8 > ((size_t)2 << 30)) chunk /= 2;
  if (h->chunk_opt > 0) chunk = h->chunk_opt;
  h->chunk = chunk;
  MGP_TRY(mgp_ensure(h, h->drP, (size_t)h->npad * chunk * 8));
  MGP_TRY(mgp_ensure(h, h->dmeanpart, (size_t)h->NI * chunk * 8));
  MGP_TRY(mgp_ensure(h, h->dftpart, (size_t)h->NI * chunk * 8));
  MGP_TRY(mgp_ensure(h, h->dqpart, (size_t)(h->trmm_i8_opt ? 4 : 2) * h->NI * chunk * 8));
  if (h->trmm_i8_opt) MGP_TRY(mgp_ensure(h, h->dRI8, ozaki_r_bytes(h->npad, chunk, h->trmm_i8_opt)));
  if (want_dx) MGP_TRY(mgp_ensure(h, h->dwP, (size_t)h->npad * chunk * 8));
  if (h->p > 1 && h->trend_est)
    MGP_TRY(mgp_ensure(h, h->dtpart, (size_t)h->NI * rgen_trend_cols(h->dpad) * chunk * 8));
  return 0;
}

static int ensure_rinv(mgp_handle *h, cudaStream_t st) {
  if (h->rinv_ready) return 0;
  const size_t mat = (size_t)h->npad * h->npad * 8;
  MGP_TRY(mgp_ensure(h, h->dTmp, mat));
  MGP_TRY(mgp_ensure(h, h->dRinvP, mat));
  MGP_CUDA(h, lauum_blocked(P<double>(h->dMinv), P<double>(h->dTmp), h->npad, 0, 1, st, &h->launches));
  MGP_CUDA(h, launch_pack_sym(P<double>(h->dTmp), h->n, h->npad, P<double>(h->dRinvP), st));
  MGP_CUDA(h, launch_symmetrize(P<double>(h->dTmp), h->npad, st));   // row-major R^-1 for the small-population path
  h->launches += 2;
  h->rinv_ready = true;
  return 0;
}

// INT8 slices of L^-1 for the Ozaki-scheme K2c (built once per fitted model, on first use)
static int ensure_i8(mgp_handle *h, cudaStream_t st) {
  if (h->i8_ready) return 0;
  MGP_TRY(mgp_ensure(h, h->dMI8, ozaki_M_bytes(h->npad, h->trmm_i8_opt)));
  MGP_TRY(mgp_ensure(h, h->dMscale, (size_t)h->npad * 8));
  MGP_CUDA(h, launch_ozaki_pack_M(P<double>(h->dMinv), h->n, h->npad, P<double>(h->dMscale), P<int8_t>(h->dMI8),
                                  h->trmm_i8_opt, st));
  h->launches += 2;
  h->i8_ready = true;
  return 0;
}

static int posterior_run(mgp_handle *h, const PostReq &q, cudaStream_t st) {
  if (!h->fitted) return mgp_fail(h, MGP_ESTATE, "model is not fitted");
  if (q.m < 0) return mgp_fail(h, MGP_EINVAL, "m < 0");
  if (q.m == 0) return 0;
  const bool acq = q.d_val || q.topk > 0 || q.d_dx;
  if (acq) {
    if (q.kind < MGP_ACQ_EI || q.kind > MGP_ACQ_MGFI) return mgp_fail(h, MGP_EINVAL, "unknown acquisition %d", q.kind);
    if (q.n_sets < 1 || q.n_sets > MGP_MAX_SETS) return mgp_fail(h, MGP_EINVAL, "n_sets must be in 1..%d", MGP_MAX_SETS);
    if (q.kind != MGP_ACQ_EI && !q.params) return mgp_fail(h, MGP_EINVAL, "set_params is NULL");
  }
  if (q.topk > 64) return mgp_fail(h, MGP_EINVAL, "k must be <= 64");
  MGP_TRY(post_workspace(h, q.want_dx));
  WsScope ws(h, st);
  MGP_CUDA(h, ws.err);
  if (q.want_dx) MGP_TRY(ensure_rinv(h, st));
  if (h->trmm_i8_opt && !q.want_dx) MGP_TRY(ensure_i8(h, st));
  const int NJ8 = h->npad / 8;
  const bool ordinary = h->trend_est;
  const bool gen = h->p > 1 && h->trend_est;
  const int PT = rgen_trend_cols(h->dpad);
  double *chunk_val = nullptr;
  int64_t *blk_idx = nullptr;
  const bool top1 = q.topk == 1 && !q.want_dx;   // per-block arg-max fused into the epilogue
  const int64_t nblk_max = (h->chunk + 255) / 256;
  if (q.topk > 0) {
    MGP_TRY(mgp_ensure(h, h->dtopk_val, (size_t)q.n_sets * h->chunk * 8));
    chunk_val = P<double>(h->dtopk_val);
    if (top1) {
      MGP_TRY(mgp_ensure(h, h->dtopk_blk, (size_t)q.n_sets * nblk_max * 8));
      blk_idx = P<int64_t>(h->dtopk_blk);
    }
  }
  for (int64_t c0 = 0; c0 < q.m; c0 += h->chunk) {
    const int64_t mc = std::min<int64_t>(h->chunk, q.m - c0);
    const int64_t ldp = round_up(mc, 128);
    const int nCblk = (int)(ldp / 128);
    const int JS = h->NI;
    const bool small = mc <= 8 && !h->no_small_path;   // matrix-vector path for tiny populations
    const bool mean_only = !q.want_dx && !acq && !q.d_mse;   // predict(eval_MSE=False): K1 alone gives r . gamma
    // (npad <= 16384: a level sum of up to 8 slice products stays below 8 x 16384 x 64 x 128 = 2^30 in int32)
    const bool use_i8 = h->trmm_i8_opt && !q.want_dx && !small && !mean_only && h->npad <= 16384;
    // K1 writes the int8 slices itself where it can (no FP64 r, no separate slicing pass)
    const bool k1_slices = use_i8 && h->trmm_i8_fuse && rgen_can_slice(h->corr, h->dpad, gen, h->trmm_i8_opt);
    RgenParams rp;
    rp.RI8 = k1_slices ? P<uint8_t>(h->dRI8) : nullptr; rp.nks = h->npad / 64;
    rp.Xs = q.d_Xs + c0 * h->d; rp.mc = mc; rp.d = h->d; rp.dpad = h->dpad; rp.NJ8 = NJ8;
    rp.nCblk = nCblk; rp.JS = JS;
    rp.center = P<double>(h->dcenter); rp.theta = P<double>(h->dtheta); rp.Bop = P<double>(h->dBop);
    rp.an = P<double>(h->dan); rp.gamma = P<double>(h->dgamma); rp.avec = P<double>(h->davec);
    rp.Xc = P<double>(h->dXc);
    rp.Bhat = gen ? P<double>(h->dBhat) : nullptr;
    rp.tpart = gen ? P<double>(h->dtpart) : nullptr;
    rp.rP = P<double>(h->drP); rp.meanpart = P<double>(h->dmeanpart); rp.ftpart = P<double>(h->dftpart);
    rp.ldp = ldp;
    { KTimer t(h, st, 0); MGP_CUDA(h, launch_rgen(rp, h->corr, h->sm_count, st)); }
    h->launches++;
    if (!q.want_dx) {
      if (mean_only) {
        // no triangular product needed
      } else if (small) {
        KTimer t(h, st, 1);
        MGP_CUDA(h, launch_small_mv(P<double>(h->dMinv), h->npad, h->n, P<double>(h->drP), (int)mc, 0,
                                    P<double>(h->dqpart), ldp, nullptr, st));
      } else if (use_i8) {
        // K2c on the INT8 tensor pipe: slice r (unless K1 already did), then the tcgen05 product (timed together as "trmm")
        KTimer t(h, st, 1);
        if (!k1_slices) {
          MGP_CUDA(h, launch_ozaki_slice_r(P<double>(h->drP), h->npad, nCblk, P<uint8_t>(h->dRI8), h->trmm_i8_opt, st));
          h->launches++;
        }
        OzakiParams op;
        op.RI8 = P<uint8_t>(h->dRI8); op.MI8 = P<int8_t>(h->dMI8); op.scale = P<double>(h->dMscale);
        op.qpart = P<double>(h->dqpart); op.ldq = ldp; op.NI2 = h->npad / 64; op.nCblk = nCblk; op.nks = h->npad / 64;
        op.cgroup = (int)std::max<int64_t>(1, ((int64_t)32 << 20) / ((int64_t)h->npad * 128 * h->trmm_i8_opt));
        op.pair = h->trmm_i8_pair;
        MGP_CUDA(h, launch_trmm_i8(op, h->trmm_i8_opt, h->sm_count, st));
        h->launches++;
      } else {
        TrmmParams tp;
        tp.Mp = P<double>(h->dMp); tp.rP = P<double>(h->drP); tp.qpart = P<double>(h->dqpart);
        tp.wP = nullptr; tp.NI = h->NI; tp.nCblk = nCblk; tp.NJ8 = NJ8; tp.full = 0; tp.ldq = ldp;
        tp.last_tiles = (h->n - (h->NI - 1) * MGP_TILE + 7) / 8;
        KTimer t(h, st, 1);
        MGP_CUDA(h, launch_trmm(tp, h->sm_count, st));
      }
      EpiParams ep;
      ep.qpart = P<double>(h->dqpart); ep.meanpart = P<double>(h->dmeanpart); ep.ftpart = P<double>(h->dftpart);
      ep.NI = mean_only ? 0 : (small ? h->npad / 32 : (use_i8 ? 4 * h->NI : 2 * h->NI)); ep.JS = JS; ep.ld = ldp; ep.mc = mc;
      ep.beta = h->beta; ep.sigma2 = h->sigma2; ep.G = h->G; ep.ordinary = ordinary;
      ep.p = h->p; ep.PT = PT; ep.d = h->d; ep.trend_est = h->trend_est ? 1 : 0;
      ep.Xs = q.d_Xs + c0 * h->d; ep.betav = P<double>(h->dbetav); ep.Cinv = P<double>(h->dCinv);
      ep.tpart = P<double>(h->dtpart);
      ep.out_mean = q.d_mean ? q.d_mean + c0 : nullptr;
      ep.out_mse = q.d_mse ? q.d_mse + c0 : nullptr;
      ep.out_val = nullptr; ep.ld_val = 0;
      ep.blk_val = nullptr; ep.blk_idx = nullptr; ep.ld_blk = nblk_max; ep.idx_base = q.idx_base + c0;
      if (top1) { ep.blk_val = chunk_val; ep.blk_idx = blk_idx; }
      else if (q.topk > 0) { ep.out_val = chunk_val; ep.ld_val = h->chunk; }
      else if (q.d_val) { ep.out_val = q.d_val + c0; ep.ld_val = q.ld_val ? q.ld_val : q.m; }
      ep.kind = q.kind; ep.n_sets = acq ? q.n_sets : 0; ep.minimize = q.minimize; ep.plugin = q.plugin;
      for (int s = 0; s < MGP_MAX_SETS; s++) ep.params[s] = (acq && q.params && s < q.n_sets) ? q.params[s] : 0.0;
      { KTimer t(h, st, 2); MGP_CUDA(h, launch_epilogue(ep, st)); }
      h->launches += 2;
    } else {
      if (small) {
        KTimer t(h, st, 1);
        MGP_CUDA(h, launch_small_mv(P<double>(h->dTmp), h->npad, h->n, P<double>(h->drP), (int)mc, 1, nullptr, ldp,
                                    P<double>(h->dwP), st));
      } else {
        TrmmParams tp;
        tp.Mp = P<double>(h->dRinvP); tp.rP = P<double>(h->drP); tp.qpart = nullptr;
        tp.wP = P<double>(h->dwP); tp.NI = h->NI; tp.nCblk = nCblk; tp.NJ8 = NJ8; tp.full = 1; tp.ldq = ldp;
        tp.last_tiles = 16;
        KTimer t(h, st, 1);
        MGP_CUDA(h, launch_trmm(tp, h->sm_count, st));
      }
      GradParams gp;
      gp.Xs = q.d_Xs + c0 * h->d; gp.mc = mc; gp.ldp = ldp; gp.n = h->n; gp.d = h->d; gp.dpad = h->dpad;
      gp.NJ8 = NJ8; gp.Xc = P<double>(h->dXc); gp.center = P<double>(h->dcenter); gp.Xsc = P<double>(h->dXsc);
      gp.theta = P<double>(h->dtheta); gp.gamma = P<double>(h->dgamma); gp.avec = P<double>(h->davec);
      gp.rP = P<double>(h->drP); gp.wP = P<double>(h->dwP);
      gp.meanpart = P<double>(h->dmeanpart); gp.ftpart = P<double>(h->dftpart); gp.JS = JS;
      gp.beta = h->beta; gp.sigma2 = h->sigma2; gp.G = h->G; gp.ftf = h->ftf; gp.ordinary = ordinary;
      gp.p = h->p; gp.PT = PT; gp.trend_est = h->trend_est ? 1 : 0;
      gp.betav = P<double>(h->dbetav); gp.Cinv = P<double>(h->dCinv); gp.tpart = P<double>(h->dtpart);
      gp.Bhat = P<double>(h->dBhat);
      gp.wide = small ? 1 : 0;
      gp.out_mean = q.d_mean ? q.d_mean + c0 : nullptr;
      gp.out_mse = q.d_mse ? q.d_mse + c0 : nullptr;
      gp.out_mean_dx = q.d_mean_dx ? q.d_mean_dx + c0 * h->d : nullptr;
      gp.out_mse_dx = q.d_mse_dx ? q.d_mse_dx + c0 * h->d : nullptr;
      gp.out_val = q.d_val ? q.d_val + c0 : nullptr;
      gp.out_dx = q.d_dx ? q.d_dx + c0 * h->d : nullptr;
      gp.ld_val = q.ld_val ? q.ld_val : q.m;
      gp.kind = q.kind; gp.n_sets = acq ? q.n_sets : 0; gp.minimize = q.minimize; gp.plugin = q.plugin;
      for (int s = 0; s < MGP_MAX_SETS; s++) gp.params[s] = (acq && q.params && s < q.n_sets) ? q.params[s] : 0.0;
      { KTimer t(h, st, 2); MGP_CUDA(h, launch_grad(gp, h->corr, st)); }
      h->launches += 2;
    }
    if (q.topk > 0) {
      { KTimer t(h, st, 3);
      if (top1)
        MGP_CUDA(h, launch_topk_merge(chunk_val, blk_idx, nblk_max, (mc + 255) / 256, 0, q.n_sets, 1, q.d_topv,
                                      q.d_topi, c0 == 0 && !q.topk_continue, st));
      else
        MGP_CUDA(h, launch_topk_merge(chunk_val, nullptr, h->chunk, mc, q.idx_base + c0, q.n_sets, q.topk,
                                      q.d_topv, q.d_topi, c0 == 0 && !q.topk_continue, st)); }
      h->launches++;
    }
  }
  return 0;
}

// ---- device-pointer entry points --------------------------------------------------------------
int mgp_predict_dev(mgp_t *h, int64_t m, const double *d_Xs, double *d_out_mean, double *d_out_mse,
                    void *stream) {
  CHECK_H(h);
  if (!d_Xs || !d_out_mean) return mgp_fail(h, MGP_EINVAL, "NULL pointer");
  PostReq q = {};
  q.m = m; q.d_Xs = d_Xs; q.d_mean = d_out_mean; q.d_mse = d_out_mse; q.minimize = 1;
  return posterior_run(h, q, reinterpret_cast<cudaStream_t>(stream));
}

int mgp_acq_dev(mgp_t *h, int64_t m, const double *d_Xs, int kind, int n_sets,
                const double *set_params, double plugin, int minimize, double *d_out_val,
                double *d_out_dx, void *stream) {
  CHECK_H(h);
  if (!d_Xs || !d_out_val) return mgp_fail(h, MGP_EINVAL, "NULL pointer");
  PostReq q = {};
  q.m = m; q.d_Xs = d_Xs; q.d_val = d_out_val; q.d_dx = d_out_dx; q.want_dx = d_out_dx != nullptr;
  q.kind = kind; q.n_sets = n_sets; q.params = set_params; q.plugin = plugin; q.minimize = minimize;
  return posterior_run(h, q, reinterpret_cast<cudaStream_t>(stream));
}

int mgp_acq_topk_dev(mgp_t *h, int64_t m, const double *d_Xs, int kind, int n_sets,
                     const double *set_params, double plugin, int minimize, int k,
                     double *d_out_val, int64_t *d_out_idx, void *stream) {
  CHECK_H(h);
  if (!d_Xs || !d_out_val || !d_out_idx || k < 1) return mgp_fail(h, MGP_EINVAL, "bad arguments");
  PostReq q = {};
  q.m = m; q.d_Xs = d_Xs; q.topk = k; q.d_topv = d_out_val; q.d_topi = d_out_idx;
  q.kind = kind; q.n_sets = n_sets; q.params = set_params; q.plugin = plugin; q.minimize = minimize;
  return posterior_run(h, q, reinterpret_cast<cudaStream_t>(stream));
}

// ---- host-pointer entry points: double-buffered H2D of candidate super-chunks overlapped with
// compute; one D2H of the results at the end. ------------------------------------------------------
struct HostOut {
  double *mean, *mse, *val, *mean_dx, *mse_dx, *dx;
};

static int host_run(mgp_handle *h, int64_t m, const double *Xs, PostReq q, const HostOut &o) {
  if (!h->fitted) return mgp_fail(h, MGP_ESTATE, "model is not fitted");
  if (m < 0 || !Xs) return mgp_fail(h, MGP_EINVAL, "bad candidates");
  if (m == 0) return 0;
  const int d = h->d;
  const int ns = std::max(1, q.n_sets);
  MGP_TRY(post_workspace(h, q.want_dx));
  // staging unit: one compute chunk; inputs and outputs are double-buffered so that the H2D of
  // unit i+1 and the D2H of unit i-1 overlap the kernels of unit i (three streams)
  const int64_t unit = std::min<int64_t>(m, h->chunk);
  size_t per = 0;  // output doubles per staged candidate
  if (o.mean) per += 1;
  if (o.mse) per += 1;
  if (o.val && q.topk == 0) per += ns;
  if (o.mean_dx) per += d;
  if (o.mse_dx) per += d;
  if (o.dx && q.topk == 0) per += (size_t)ns * d;
  for (int i = 0; i < 2; i++) {
    MGP_TRY(mgp_ensure(h, h->dXs_stage[i], (size_t)unit * d * 8));
    MGP_TRY(mgp_ensure(h, h->dout_stage[i], std::max<size_t>(per * unit, 1) * 8));
  }
  if (q.topk > 0) {
    MGP_TRY(mgp_ensure(h, h->dtopk_idx, (size_t)ns * q.topk * 16));
    q.d_topi = P<int64_t>(h->dtopk_idx);
    q.d_topv = reinterpret_cast<double *>(q.d_topi + (size_t)ns * q.topk);
  }
  cudaStream_t st = h->stream, cs = h->copy_stream, os = h->out_stream;
  // Small calls -- the one-point-per-call pattern of the reference's own optimisers (optimizer/utils.py:38-42):
  // no copy engine, no events, one stream, one synchronisation.  The CPU copies the candidates into a pinned,
  // device-mapped page that K1 reads directly, and the result kernels write into the same page.
  if (q.topk == 0 && h->zero_copy_opt && (size_t)m * d * 8 <= 4096 && per * (size_t)m * 8 <= 12288) {
    if (!h->hpin) {
      MGP_CUDA(h, cudaHostAlloc(&h->hpin, 16384, cudaHostAllocMapped));
      MGP_CUDA(h, cudaHostGetDevicePointer(&h->hpin_dev, h->hpin, 0));
    }
    memcpy(h->hpin, Xs, (size_t)m * d * 8);
    double *dout = reinterpret_cast<double *>(static_cast<char *>(h->hpin_dev) + 4096);
    const double *hout = reinterpret_cast<const double *>(static_cast<const char *>(h->hpin) + 4096);
    size_t off = 0;
    auto take = [&](bool on, size_t cnt) { double *r = on ? dout + off : nullptr; if (on) off += cnt; return r; };
    PostReq r = q;
    r.m = m; r.d_Xs = static_cast<const double *>(h->hpin_dev);
    r.d_mean = take(o.mean, m);
    r.d_mse = take(o.mse, m);
    r.d_val = take(o.val, (size_t)ns * m);
    r.d_mean_dx = take(o.mean_dx, (size_t)m * d);
    r.d_mse_dx = take(o.mse_dx, (size_t)m * d);
    r.d_dx = take(o.dx, (size_t)ns * m * d);
    r.ld_val = m;
    r.idx_base = 0;
    r.topk_continue = false;
    MGP_TRY(posterior_run(h, r, st));
    MGP_CUDA(h, cudaStreamSynchronize(st));
    off = 0;
    auto give = [&](double *dst, size_t cnt) { if (dst) { memcpy(dst, hout + off, cnt * 8); off += cnt; } };
    give(o.mean, m);
    give(o.mse, m);
    give(o.val, (size_t)ns * m);
    give(o.mean_dx, (size_t)m * d);
    give(o.mse_dx, (size_t)m * d);
    give(o.dx, (size_t)ns * m * d);
    return MGP_OK;
  }
  // Ramp: the H2D of the first unit is the one copy no kernel can hide, so large populations start
  // with an eighth of a unit, then half a unit (its copy fits under the first unit's kernels), then
  // full units -- the exposed copy shrinks from 12 MB to 1.5 MB at C2 (10^6 candidates, d = 10).
  const int64_t ramp0 = round_up(std::max<int64_t>(unit / 8, (int64_t)h->sm_count * 128), 128);
  const bool ramp = m >= 3 * unit && ramp0 * 4 <= unit;
  int it = 0;
  for (int64_t s0 = 0, ms = 0; s0 < m; s0 += ms, it++) {
    const int64_t want = !ramp ? unit : (it == 0 ? ramp0 : (it == 1 ? round_up(unit / 2, 128) : unit));
    ms = std::min(want, m - s0);
    const int buf = it & 1;
    if (it >= 2) MGP_CUDA(h, cudaStreamWaitEvent(cs, h->ev_done[buf], 0));   // input buffer free
    MGP_CUDA(h, cudaMemcpyAsync(h->dXs_stage[buf].p, Xs + s0 * d, (size_t)ms * d * 8, cudaMemcpyHostToDevice, cs));
    MGP_CUDA(h, cudaEventRecord(h->ev_in[buf], cs));
    MGP_CUDA(h, cudaStreamWaitEvent(st, h->ev_in[buf], 0));
    if (it >= 2) MGP_CUDA(h, cudaStreamWaitEvent(st, h->ev_out[buf], 0));    // output buffer drained
    double *dout = P<double>(h->dout_stage[buf]);
    size_t off = 0;
    auto take = [&](bool on, size_t cnt) { double *r = on ? dout + off : nullptr; if (on) off += cnt; return r; };
    PostReq r = q;
    r.m = ms; r.d_Xs = P<double>(h->dXs_stage[buf]);
    r.d_mean = take(o.mean, ms);
    r.d_mse = take(o.mse, ms);
    r.d_val = take(o.val && q.topk == 0, (size_t)ns * ms);
    r.d_mean_dx = take(o.mean_dx, (size_t)ms * d);
    r.d_mse_dx = take(o.mse_dx, (size_t)ms * d);
    r.d_dx = take(o.dx && q.topk == 0, (size_t)ns * ms * d);
    r.ld_val = ms;
    r.idx_base = s0;
    r.topk_continue = it > 0;
    MGP_TRY(posterior_run(h, r, st));
    MGP_CUDA(h, cudaEventRecord(h->ev_done[buf], st));
    if (per > 0) {
      MGP_CUDA(h, cudaStreamWaitEvent(os, h->ev_done[buf], 0));
      if (r.d_mean) MGP_CUDA(h, cudaMemcpyAsync(o.mean + s0, r.d_mean, (size_t)ms * 8, cudaMemcpyDeviceToHost, os));
      if (r.d_mse) MGP_CUDA(h, cudaMemcpyAsync(o.mse + s0, r.d_mse, (size_t)ms * 8, cudaMemcpyDeviceToHost, os));
      if (r.d_val)
        MGP_CUDA(h, cudaMemcpy2DAsync(o.val + s0, (size_t)m * 8, r.d_val, (size_t)ms * 8, (size_t)ms * 8, ns,
                                      cudaMemcpyDeviceToHost, os));
      if (r.d_mean_dx) MGP_CUDA(h, cudaMemcpyAsync(o.mean_dx + s0 * d, r.d_mean_dx, (size_t)ms * d * 8, cudaMemcpyDeviceToHost, os));
      if (r.d_mse_dx) MGP_CUDA(h, cudaMemcpyAsync(o.mse_dx + s0 * d, r.d_mse_dx, (size_t)ms * d * 8, cudaMemcpyDeviceToHost, os));
      if (r.d_dx)
        MGP_CUDA(h, cudaMemcpy2DAsync(o.dx + s0 * d, (size_t)m * d * 8, r.d_dx, (size_t)ms * d * 8, (size_t)ms * d * 8, ns,
                                      cudaMemcpyDeviceToHost, os));
      MGP_CUDA(h, cudaEventRecord(h->ev_out[buf], os));
    }
  }
  if (q.topk > 0) {
    MGP_CUDA(h, cudaMemcpyAsync(o.val, q.d_topv, (size_t)ns * q.topk * 8, cudaMemcpyDeviceToHost, st));
    MGP_CUDA(h, cudaMemcpyAsync(reinterpret_cast<int64_t *>(o.dx), q.d_topi, (size_t)ns * q.topk * 8, cudaMemcpyDeviceToHost, st));
  }
  MGP_CUDA(h, cudaStreamSynchronize(st));
  MGP_CUDA(h, cudaStreamSynchronize(os));
  MGP_CUDA(h, cudaStreamSynchronize(cs));
  return MGP_OK;
}

int mgp_predict(mgp_t *h, int64_t m, const double *Xs, double *out_mean, double *out_mse) {
  CHECK_H(h);
  if (!out_mean) return mgp_fail(h, MGP_EINVAL, "out_mean is NULL");
  PostReq q = {};
  q.minimize = 1;
  HostOut o = {out_mean, out_mse, nullptr, nullptr, nullptr, nullptr};
  return host_run(h, m, Xs, q, o);
}

int mgp_gradient(mgp_t *h, int64_t m, const double *Xs, double *out_mean_dx, double *out_mse_dx) {
  CHECK_H(h);
  if (!out_mean_dx || !out_mse_dx) return mgp_fail(h, MGP_EINVAL, "NULL output");
  PostReq q = {};
  q.minimize = 1; q.want_dx = 1;
  HostOut o = {nullptr, nullptr, nullptr, out_mean_dx, out_mse_dx, nullptr};
  return host_run(h, m, Xs, q, o);
}

int mgp_acq(mgp_t *h, int64_t m, const double *Xs, int kind, int n_sets, const double *set_params,
            double plugin, int minimize, double *out_val, double *out_dx) {
  CHECK_H(h);
  if (!out_val) return mgp_fail(h, MGP_EINVAL, "out_val is NULL");
  PostReq q = {};
  q.kind = kind; q.n_sets = n_sets; q.params = set_params; q.plugin = plugin; q.minimize = minimize;
  q.want_dx = out_dx != nullptr;
  HostOut o = {nullptr, nullptr, out_val, nullptr, nullptr, out_dx};
  return host_run(h, m, Xs, q, o);
}

int mgp_acq_topk(mgp_t *h, int64_t m, const double *Xs, int kind, int n_sets,
                 const double *set_params, double plugin, int minimize, int k, double *out_val,
                 int64_t *out_idx) {
  CHECK_H(h);
  if (!out_val || !out_idx || k < 1) return mgp_fail(h, MGP_EINVAL, "bad arguments");
  PostReq q = {};
  q.kind = kind; q.n_sets = n_sets; q.params = set_params; q.plugin = plugin; q.minimize = minimize;
  q.topk = k;
  HostOut o = {nullptr, nullptr, out_val, nullptr, nullptr, reinterpret_cast<double *>(out_idx)};
  return host_run(h, m, Xs, q, o);
}

}  // extern "C"
