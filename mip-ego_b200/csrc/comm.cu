// Multi-GPU exchange steps of the GP hot path behind the C ABI: one process
// per GPU, NCCL over NVLink / NVSwitch.
//   * mgp_bcast_model    : the fitted model (theta, L, L^-1, packed posterior operands, gamma, ...)
//                          goes from the root's HBM to every other rank's HBM -- no host bounce, no
//                          re-factorisation on the receivers, bit-identical models everywhere;
//   * mgp_allgather_topk : every rank's k best (value, global index) pairs per parameter set are
//                          all-gathered and merged on the device with the single-GPU order (value
//                          descending, NaN last, ties -> smaller index), identically on every rank;
//   * mgp_allgather_f64  : plain all-gather of doubles (likelihood rows of sharded MLE restarts).
// The path itself needs no collective inside its kernels: candidates and restarts are independent
// units, the payloads here are one model per fit and n_sets * k * 16 bytes per population.
//
// NCCL is bound at run time (dlopen/dlsym): the library has no link-time dependency on it, a
// single-GPU user never loads it, and inside a PyTorch process the NCCL that torch already loaded
// is reused instead of a second copy.
#include <dlfcn.h>
#include <nccl.h>   // types and enums only; every function is resolved with dlsym
#include <string.h>

#include <algorithm>
#include <mutex>
#include <string>
#include <vector>

#include "mgp_handle.h"
#include "posterior.cuh"

namespace {

struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int *) = nullptr;
  std::string where, err;
};

NcclApi *nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    std::vector<std::pair<std::string, int>> tries;
    const char *env = getenv("MGP_NCCL_LIB");
    if (env && *env) tries.push_back({env, RTLD_NOW | RTLD_GLOBAL});
    tries.push_back({"libnccl.so.2", RTLD_NOW | RTLD_NOLOAD});   // the copy this process already loaded (torch's)
    tries.push_back({"libnccl.so.2", RTLD_NOW | RTLD_GLOBAL});
    tries.push_back({"libnccl.so", RTLD_NOW | RTLD_GLOBAL});
    for (auto &t : tries) {
      api.lib = dlopen(t.first.c_str(), t.second);
      if (api.lib) { api.where = t.first + ((t.second & RTLD_NOLOAD) ? " (already loaded)" : ""); break; }
    }
    if (!api.lib) { api.err = "NCCL library not found (set MGP_NCCL_LIB to libnccl.so.2)"; return; }
#define SYM(field, name)                                                       \
  *reinterpret_cast<void **>(&api.field) = dlsym(api.lib, name);               \
  if (!api.field) { api.err = std::string("NCCL symbol missing: ") + name; return; }
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(Broadcast, "ncclBroadcast")
    SYM(AllGather, "ncclAllGather")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(GetErrorString, "ncclGetErrorString")
    SYM(GetVersion, "ncclGetVersion")
#undef SYM
  });
  return &api;
}

int nccl_fail(mgp_handle *h, NcclApi *a, ncclResult_t r, const char *what) {
  return mgp_fail(h, MGP_ENCCL, "NCCL error '%s' in %s", a->GetErrorString ? a->GetErrorString(r) : "?", what);
}
#define MGP_NCCL(h, a, call)                                    \
  do {                                                          \
    ncclResult_t r__ = (call);                                  \
    if (r__ != ncclSuccess) return nccl_fail((h), (a), r__, #call); \
  } while (0)

template <typename T>
inline T *P(const DevBuf &b) { return reinterpret_cast<T *>(b.p); }

// send[0 .. S*k) = values (NaN / empty -> -inf), send[S*k .. 2*S*k) = global indices as int64 bits
__global__ void k_topk_pack(const double *val, const int64_t *idx, int cnt, int64_t offset, int64_t *send) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= cnt) return;
  double v = val[e];
  const int64_t i = idx[e];
  const bool empty = i < 0;
  if (empty || !(v == v)) v = -INFINITY;
  send[e] = __double_as_longlong(v);
  send[cnt + e] = empty ? INT64_MAX : i + offset;
}

const int HDR = 128;  // doubles in the model header: 32 scalars + up to 96 hyper-parameters

}  // namespace

extern "C" {

int mgp_comm_unique_id(char *out_id) {
  if (!out_id) return MGP_EINVAL;
  NcclApi *a = nccl_api();
  if (!a->err.empty()) return MGP_ENCCL;
  ncclUniqueId id;
  if (a->GetUniqueId(&id) != ncclSuccess) return MGP_ENCCL;
  static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
  memcpy(out_id, &id, sizeof(id));
  return MGP_OK;
}

int mgp_comm_init(mgp_t *h, int rank, int world, const char *id_bytes) {
  if (!h) return MGP_EINVAL;
  MGP_CUDA(h, cudaSetDevice(h->device));
  if (!id_bytes || world < 1 || rank < 0 || rank >= world) return mgp_fail(h, MGP_EINVAL, "bad communicator arguments");
  NcclApi *a = nccl_api();
  if (!a->err.empty()) return mgp_fail(h, MGP_ENCCL, "%s", a->err.c_str());
  if (h->comm) { a->CommDestroy(reinterpret_cast<ncclComm_t>(h->comm)); h->comm = nullptr; }
  ncclUniqueId id;
  memcpy(&id, id_bytes, sizeof(id));
  ncclComm_t c = nullptr;
  MGP_NCCL(h, a, a->CommInitRank(&c, world, id, rank));
  h->comm = c;
  h->rank = rank;
  h->world = world;
  return MGP_OK;
}

int mgp_comm_destroy(mgp_t *h) {
  if (!h) return MGP_EINVAL;
  if (h->comm) {
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    nccl_api()->CommDestroy(reinterpret_cast<ncclComm_t>(h->comm));
    h->comm = nullptr;
  }
  h->world = 1;
  h->rank = 0;
  return MGP_OK;
}

const char *mgp_comm_info(const mgp_t *h) {
  static thread_local char buf[256];
  NcclApi *a = nccl_api();
  int v = 0;
  if (a->err.empty() && a->GetVersion) a->GetVersion(&v);
  snprintf(buf, sizeof(buf), "nccl=%s version=%d rank=%d world=%d", a->err.empty() ? a->where.c_str() : a->err.c_str(),
           v, h ? h->rank : 0, h ? h->world : 1);
  return buf;
}

int mgp_bcast_model(mgp_t *h, int root) {
  if (!h) return MGP_EINVAL;
  MGP_CUDA(h, cudaSetDevice(h->device));
  if (!h->comm) return mgp_fail(h, MGP_ESTATE, "mgp_comm_init must be called first");
  if (root < 0 || root >= h->world) return mgp_fail(h, MGP_EINVAL, "root out of range");
  if (h->n == 0) return mgp_fail(h, MGP_ESTATE, "mgp_set_train must be called first (every rank holds X, y)");
  const bool is_root = h->rank == root;
  NcclApi *a = nccl_api();
  ncclComm_t comm = reinterpret_cast<ncclComm_t>(h->comm);
  cudaStream_t st = h->stream;
  if (h->ws_recorded) MGP_CUDA(h, cudaStreamWaitEvent(st, h->ev_ws, 0));   // previous user of the model buffers
  // ---- header: shapes + scalars from the root; every rank checks it against its own training data
  double hdr[HDR] = {0};
  double xsum = 0.0, ysum = 0.0;
  for (double v : h->hX) xsum += v;
  for (double v : h->hy) ysum += v;
  if (is_root) {
    double vals[] = {(double)h->n, (double)h->d, (double)h->corr, (double)h->trend, xsum, ysum, (double)h->mode,
                     h->sigma2, h->beta, h->G, h->ftf, h->llf, h->noise_var, h->logdetFF, (double)(h->isotropic ? 1 : 0),
                     (double)(h->fitted ? 1 : 0)};
    memcpy(hdr, vals, sizeof(vals));
    // the hyper-parameters of the fit travel too: a receiver can then extend the model
    // incrementally (mgp_append_train) exactly like the root
    hdr[16] = (double)h->hpar.size();
    hdr[17] = h->par_noise_var;
    for (size_t i = 0; i < h->hpar.size() && 32 + i < (size_t)HDR; i++) hdr[32 + i] = h->hpar[i];
  }
  MGP_TRY(mgp_ensure(h, h->dcomm, (size_t)(HDR + 2 * h->world + 8) * 8));
  double *dh = P<double>(h->dcomm);
  if (is_root) MGP_CUDA(h, cudaMemcpyAsync(dh, hdr, sizeof(hdr), cudaMemcpyHostToDevice, st));
  MGP_NCCL(h, a, a->Broadcast(dh, dh, HDR, ncclFloat64, root, comm, st));
  MGP_CUDA(h, cudaMemcpyAsync(hdr, dh, sizeof(hdr), cudaMemcpyDeviceToHost, st));
  MGP_CUDA(h, cudaStreamSynchronize(st));
  const bool same = hdr[0] == h->n && hdr[1] == h->d && hdr[2] == h->corr && hdr[3] == h->trend && hdr[4] == xsum &&
                    hdr[5] == ysum && hdr[15] == 1.0;   // [15]: the root's model is fitted
  // agree on the verdict before the large transfers, so that a mismatch fails on EVERY rank instead
  // of leaving the others blocked inside a collective
  double okflag = same ? 1.0 : 0.0;
  double *dflag = dh + HDR, *dall = dh + HDR + 1;
  MGP_CUDA(h, cudaMemcpyAsync(dflag, &okflag, 8, cudaMemcpyHostToDevice, st));
  MGP_NCCL(h, a, a->AllGather(dflag, dall, 1, ncclFloat64, comm, st));
  std::vector<double> oks(h->world, 0.0);
  MGP_CUDA(h, cudaMemcpyAsync(oks.data(), dall, (size_t)h->world * 8, cudaMemcpyDeviceToHost, st));
  MGP_CUDA(h, cudaStreamSynchronize(st));
  for (int r = 0; r < h->world; r++)
    if (oks[r] != 1.0)
      return mgp_fail(h, hdr[15] == 1.0 ? MGP_EINVAL : MGP_ESTATE,
                      hdr[15] == 1.0 ? "rank %d holds different training data (n, d, kernel, trend or X / y) than root %d"
                                     : "the model of root %2$d is not fitted (seen by rank %1$d)", r, root);
  if (!is_root) {
    MGP_TRY(mgp_model_alloc(h));
    h->mode = (int)hdr[6]; h->sigma2 = hdr[7]; h->beta = hdr[8]; h->G = hdr[9]; h->ftf = hdr[10]; h->llf = hdr[11];
    h->noise_var = hdr[12]; h->logdetFF = hdr[13]; h->isotropic = hdr[14] != 0.0;
    const int np = (int)hdr[16];
    h->hpar.assign(hdr + 32, hdr + 32 + (np > 0 && np <= HDR - 32 ? np : 0));
    h->par_noise_var = hdr[17];
  }
  const size_t mat = (size_t)h->npad * h->npad, vec = (size_t)h->npad;
  struct Item { DevBuf *b; size_t count; };
  std::vector<Item> items = {{&h->dtheta, (size_t)h->dpad}, {&h->dL, mat}, {&h->dMinv, mat},
                             {&h->dMp, (size_t)8 * h->NI * (h->NI + 1) * 1024}, {&h->dBop, vec * h->dpad},
                             {&h->dan, vec}, {&h->dXsc, vec * grad_row_stride(h->dpad)}, {&h->dgamma, vec}, {&h->davec, vec}, {&h->dFt, vec}, {&h->dYt, vec},
                             {&h->drho, vec}};
  if (h->p > 1) {
    MGP_TRY(mgp_ensure(h, h->dbetav, MGP_PMAX * 8));
    items.push_back({&h->dbetav, (size_t)MGP_PMAX});
    if (h->trend_est) {
      MGP_TRY(mgp_ensure(h, h->dBhat, vec * MGP_TP * 8));
      MGP_TRY(mgp_ensure(h, h->dFtm, vec * MGP_TP * 8));
      MGP_TRY(mgp_ensure(h, h->dCinv, (size_t)MGP_PMAX * MGP_PMAX * 8));
      items.push_back({&h->dBhat, vec * MGP_TP});
      items.push_back({&h->dFtm, vec * MGP_TP});
      items.push_back({&h->dCinv, (size_t)MGP_PMAX * MGP_PMAX});
    }
  }
  size_t total = 0;
  MGP_NCCL(h, a, a->GroupStart());
  for (auto &it : items) {
    ncclResult_t r = a->Broadcast(it.b->p, it.b->p, it.count, ncclFloat64, root, comm, st);
    if (r != ncclSuccess) { a->GroupEnd(); return nccl_fail(h, a, r, "ncclBroadcast(model)"); }
    total += it.count * 8;
  }
  MGP_NCCL(h, a, a->GroupEnd());
  if (h->p > 1) {
    h->hbeta.assign(MGP_PMAX, 0.0);
    MGP_CUDA(h, cudaMemcpyAsync(h->hbeta.data(), h->dbetav.p, MGP_PMAX * 8, cudaMemcpyDeviceToHost, st));
  }
  h->ws_recorded = true;
  MGP_CUDA(h, cudaEventRecord(h->ev_ws, st));
  MGP_CUDA(h, cudaStreamSynchronize(st));
  if (!is_root) { h->fitted = true; h->rinv_ready = false; }
  h->bcast_bytes = (int64_t)total;
  return MGP_OK;
}

int mgp_allgather_f64_dev(mgp_t *h, const double *d_send, int64_t count, double *d_recv, void *stream) {
  if (!h) return MGP_EINVAL;
  MGP_CUDA(h, cudaSetDevice(h->device));
  if (!h->comm) return mgp_fail(h, MGP_ESTATE, "mgp_comm_init must be called first");
  if (!d_send || !d_recv || count < 1) return mgp_fail(h, MGP_EINVAL, "bad arguments");
  NcclApi *a = nccl_api();
  MGP_NCCL(h, a, a->AllGather(d_send, d_recv, (size_t)count, ncclFloat64, reinterpret_cast<ncclComm_t>(h->comm),
                              reinterpret_cast<cudaStream_t>(stream)));
  return MGP_OK;
}

int mgp_allgather_topk_dev(mgp_t *h, int n_sets, int k, const double *d_val, const int64_t *d_idx,
                           int64_t idx_offset, double *d_out_val, int64_t *d_out_idx, void *stream) {
  if (!h) return MGP_EINVAL;
  MGP_CUDA(h, cudaSetDevice(h->device));
  if (!h->comm) return mgp_fail(h, MGP_ESTATE, "mgp_comm_init must be called first");
  if (n_sets < 1 || n_sets > MGP_MAX_SETS || k < 1 || k > 64 || !d_val || !d_idx || !d_out_val || !d_out_idx)
    return mgp_fail(h, MGP_EINVAL, "bad arguments to mgp_allgather_topk");
  NcclApi *a = nccl_api();
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int cnt = n_sets * k;
  MGP_TRY(mgp_ensure(h, h->dgather, (size_t)(h->world + 1) * 2 * cnt * 8));
  int64_t *send = P<int64_t>(h->dgather), *recv = send + 2 * cnt;
  MGP_CUDA(h, cudaStreamWaitEvent(st, h->ev_gather, 0));   // previous user of the gather buffer
  k_topk_pack<<<(cnt + 127) / 128, 128, 0, st>>>(d_val, d_idx, cnt, idx_offset, send);
  MGP_CUDA(h, cudaGetLastError());
  MGP_NCCL(h, a, a->AllGather(send, recv, (size_t)2 * cnt, ncclInt64, reinterpret_cast<ncclComm_t>(h->comm), st));
  // merge rank by rank, in rank order, with the running top-k kernel of the single-GPU path
  for (int r = 0; r < h->world; r++) {
    const int64_t *blk = recv + (size_t)r * 2 * cnt;
    MGP_CUDA(h, launch_topk_merge(reinterpret_cast<const double *>(blk), blk + cnt, k, k, 0, n_sets, k, d_out_val,
                                  d_out_idx, r == 0, st));
  }
  h->launches += 1 + h->world;
  MGP_CUDA(h, cudaEventRecord(h->ev_gather, st));
  return MGP_OK;
}

int mgp_allgather_topk(mgp_t *h, int n_sets, int k, const double *val, const int64_t *idx, int64_t idx_offset,
                       double *out_val, int64_t *out_idx) {
  if (!h) return MGP_EINVAL;
  MGP_CUDA(h, cudaSetDevice(h->device));
  if (!val || !idx || !out_val || !out_idx || n_sets < 1 || k < 1) return mgp_fail(h, MGP_EINVAL, "bad arguments");
  const size_t cnt = (size_t)n_sets * k;
  MGP_TRY(mgp_ensure(h, h->dgather_io, 4 * cnt * 8));
  double *dv = P<double>(h->dgather_io), *dov = dv + cnt;
  int64_t *di = reinterpret_cast<int64_t *>(dov + cnt), *doi = di + cnt;
  cudaStream_t st = h->stream;
  MGP_CUDA(h, cudaMemcpyAsync(dv, val, cnt * 8, cudaMemcpyHostToDevice, st));
  MGP_CUDA(h, cudaMemcpyAsync(di, idx, cnt * 8, cudaMemcpyHostToDevice, st));
  MGP_TRY(mgp_allgather_topk_dev(h, n_sets, k, dv, di, idx_offset, dov, doi, st));
  MGP_CUDA(h, cudaMemcpyAsync(out_val, dov, cnt * 8, cudaMemcpyDeviceToHost, st));
  MGP_CUDA(h, cudaMemcpyAsync(out_idx, doi, cnt * 8, cudaMemcpyDeviceToHost, st));
  MGP_CUDA(h, cudaStreamSynchronize(st));
  return MGP_OK;
}

}  // extern "C"
