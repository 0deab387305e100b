// Internal handle of libmgp_b200 (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/mgp.h"
#include "linalg.cuh"

#define MGP_MAX_SETS 16
#define MGP_MAX_D 64
#define MGP_NLL_STREAMS 8

struct DevBuf {
  void *p = nullptr;
  size_t bytes = 0;
};

struct mgp_handle {
  int device = 0;
  std::string err;
  cudaStream_t stream = nullptr;   // internal stream for the host-pointer entry points
  cudaStream_t copy_stream = nullptr;  // H2D of staged candidates
  cudaStream_t out_stream = nullptr;   // D2H of staged results
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
  int64_t launches = 0;
  int64_t appends = 0;          // successful mgp_append_train calls
  // Workspace ordering across streams: the posterior / likelihood scratch, the fitted-state buffers
  // and the running top-k are shared by the host-pointer entry points (h->stream) and the *_dev
  // entry points (caller's stream).  Every entry point makes its stream wait on ev_ws (recorded by
  // whoever used the workspaces last) before touching them and records it again when done.
  cudaEvent_t ev_ws = nullptr;
  bool ws_recorded = false;
  // optional per-kernel timing (option "time_kernels"): event pairs resolved lazily
  int time_kernels = 0;
  struct EvPair { cudaEvent_t a, b; int which; };
  std::vector<EvPair> ev_pending;
  double kernel_ms[4] = {0, 0, 0, 0};   // rgen, trmm, epilogue/grad, topk
  int64_t kernel_calls[4] = {0, 0, 0, 0};
  int sm_count = 148;

  // ---- training data -------------------------------------------------------------------
  int n = 0, d = 0, npad = 0, dpad = 0, NI = 0;
  int corr = 0, trend = 0;
  bool isotropic = false;       // one theta shared by all features (kernel.py:312-317); option "isotropic"
  double beta0 = 0.0;
  int p = 1;                    // basis functions of the trend (1 constant, d + 1 linear)
  bool trend_est = false;       // coefficients estimated (ordinary / universal kriging)
  std::vector<double> beta_in;  // given coefficients (p), *_FIXED trends
  std::vector<double> hbeta;    // fitted / adopted coefficients (p)
  double logdetFF = 0.0;        // log det(F^T F), REML
  DevBuf dF;       // npad x MGP_TP basis matrix (p > 1, estimated)
  DevBuf dmu;      // npad: F beta (linear trend with given beta)
  DevBuf dbetav;   // MGP_PMAX: trend coefficients on the device
  DevBuf dBhat;    // npad x MGP_TP: L^-T (Ft C^-T)  (p > 1, estimated)
  DevBuf dCinv;    // MGP_PMAX x MGP_PMAX: C^-1, C = chol(Ft^T Ft)
  DevBuf dFtm;     // npad x MGP_TP: Ft = L^-1 F (state exchange)
  std::vector<double> hX, hy;   // host copies (state exchange, duplicate check)
  DevBuf dXc;      // npad x dpad, centred, zero padded
  DevBuf dy;       // npad
  DevBuf dcenter;  // dpad
  // ---- fitted model --------------------------------------------------------------------
  bool fitted = false;
  int mode = 0;
  double sigma2 = 0, beta = 0, G = 0, ftf = 0, llf = 0, noise_var = 0;
  // hyper-parameters of the fitted model in the kernels' layout (d theta values + extras), kept by
  // mgp_finalize_fit for mgp_append_train; empty after mgp_set_state (adopted from outside)
  std::vector<double> hpar;
  double par_noise_var = 0.0;
  std::vector<double> hcenter;   // centring offsets of dXc
  DevBuf dtheta;   // dpad
  DevBuf dBop;     // npad x dpad : -2 theta_i xc_ji
  DevBuf dan;      // npad        : sum_i theta_i xc_ji^2
  DevBuf dXsc;     // npad x grad_row_stride(dpad): sqrt(theta_i) xc_ji (gradient kernel)
  DevBuf dL;       // npad x npad row-major lower Cholesky factor
  DevBuf dMinv;    // npad x npad row-major L^-1
  DevBuf dTmp;     // npad x npad scratch
  DevBuf dMp;      // packed lower-triangular L^-1 (posterior A operand)
  DevBuf dRinvP;   // packed full R^-1 (gradient path), built lazily
  // INT8 (Ozaki-scheme) variant of K2c, option "trmm_i8": slices of L^-1 (built lazily per fit), row
  // scales, slices of the current chunk's r
  DevBuf dMI8, dMscale, dRI8;
  bool i8_ready = false;
  int trmm_i8_opt = 0;
  // small host-pointer calls: pinned page mapped into the device address space (candidates in, results out)
  void *hpin = nullptr, *hpin_dev = nullptr;
  int zero_copy_opt = 1;
  int trmm_i8_fuse = 1;        // K1 writes the int8 slices of r directly (SE / Matern, constant trend, 8 slices, d <= 40)
  int trmm_i8_pair = 0;        // K2c-i8 as clusters of two CTAs with multicast candidate slices (no gain measured: off)
  bool rinv_ready = false;
  DevBuf dgamma, davec, dFt, dYt, drho;  // npad each
  DevBuf dscal;    // small device scalar area
  // ---- posterior workspace ---------------------------------------------------------------
  int64_t chunk = 0;       // candidates per pass (multiple of 128)
  int64_t chunk_opt = 0;   // user override
  bool no_small_path = false;  // option "small_path" = 0
  DevBuf drP, dmeanpart, dftpart, dqpart, dwP, dtpart;
  DevBuf dXs_stage[2], dout_stage[2];
  DevBuf dtopk_val, dtopk_idx, dtopk_blk;
  // ---- likelihood workspace --------------------------------------------------------------
  int nll_cap = 0;  // batch capacity of the buffers below
  int nll_groups_opt = 0;  // 0 = default
  int gemm_cap_opt = 0;    // user override of the per-launch CTA cap (0 = automatic)
  cudaStream_t nll_stream[MGP_NLL_STREAMS] = {};  // batch groups
  cudaEvent_t ev_fork = nullptr, ev_join[MGP_NLL_STREAMS] = {};
  // second stream per batch group: L^-1 block row by block row while the factorisation runs (RowInv)
  cudaStream_t nll_side[MGP_NLL_STREAMS] = {};
  cudaEvent_t ev_sfork[MGP_NLL_STREAMS] = {}, ev_sjoin[MGP_NLL_STREAMS] = {};
  int rowinv_opt = -1;     // option "trtri_rowwise": -1 automatic (few small matrices), 0 never, 1 always
  DevBuf nR, nL, nM, nT, nLinvDiag, nvec, nscal, npar, ngpart;
  DevBuf nFtm, nBhat, ntws;   // general-trend workspaces (allocated when p > 1)
  DevBuf niso;                // isotropic theta: expanded parameters | full gradient
  // CUDA graphs of the batched likelihood (one per shape): the ~100 launches of a batch on up to 8
  // streams are captured once and replayed with ONE cudaGraphLaunch.  The graph reads its parameters
  // from / writes its results to the handle's own staging area (ngio), so it does not depend on the
  // caller's pointers.
  struct NllGraph {
    int B, n_par, mode, eval_grad, n, npad, groups, cap, iso;
    double noise_var;
    cudaGraphExec_t exec;      // nullptr: shape seen once (eager run), captured at the next call
    int64_t launches;          // kernels in the graph
    int64_t uses;
  };
  std::vector<NllGraph> nll_graphs;
  int graphs_opt = 1;         // option "nll_graphs": 0 disables capture
  int64_t graph_replays = 0;
  DevBuf ngio;                // staging: params | llf | grad | status
  // ---- multi-GPU (comm.cu) -----------------------------------------------------------------
  void *comm = nullptr;       // ncclComm_t
  int rank = 0, world = 1;
  DevBuf dcomm, dgather, dgather_io;
  cudaEvent_t ev_gather = nullptr;
  int64_t bcast_bytes = 0;    // payload of the last mgp_bcast_model
};

int mgp_fail(mgp_handle *h, int code, const char *fmt, ...);
int mgp_cuda_fail(mgp_handle *h, cudaError_t e, const char *what);
int mgp_ensure(mgp_handle *h, DevBuf &b, size_t bytes);
extern "C" int mgp_model_alloc(mgp_handle *h);   // fitted-model buffers of a handle that has training data (api.cu)

#define MGP_CUDA(h, call)                                          \
  do {                                                             \
    cudaError_t e__ = (call);                                      \
    if (e__ != cudaSuccess) return mgp_cuda_fail((h), e__, #call); \
  } while (0)
#define MGP_TRY(call)          \
  do {                         \
    int rc__ = (call);         \
    if (rc__ != 0) return rc__; \
  } while (0)
