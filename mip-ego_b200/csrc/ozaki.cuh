// K2c-i8: the triangular product T = L^-1 r with fused column sums of squares evaluated on the INT8
// tensor pipe (tcgen05.mma kind::i8, TMEM accumulators) by an Ozaki-scheme split of both FP64
// operands (ozaki.cu).  Opt-in through the handle option "trmm_i8".
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct OzakiParams {
  const uint8_t *RI8;   // r slices of the current chunk: [cand block][K stage][S slices][8 KB]
  const int8_t *MI8;    // L^-1 slices: [64-row block I2][K stage <= I2][S slices][4 KB]
  const double *scale;  // npad row scales 2^ea_i of L^-1
  double *qpart;        // [2 NI2][ldq] partial sums of squares per 32-row half block
  int64_t ldq;
  int NI2;              // npad / 64
  int nCblk;            // candidate blocks of 128
  int nks;              // K stages per candidate block (npad / 64)
  int cgroup;           // candidate blocks per L2-resident group
  int pair;             // 1: clusters of two CTAs share (multicast) the candidate slices of a stage
};
// S: slices per operand, 8 (default; 56 bits) or 7 (49 bits, faster)
size_t ozaki_M_bytes(int npad, int S);
size_t ozaki_r_bytes(int npad, int64_t chunk, int S);
// per fit: row scales + int8 slices of the row-major L^-1 (npad x npad, lower)
cudaError_t launch_ozaki_pack_M(const double *Minv, int n, int npad, double *scale, int8_t *MI8, int S, cudaStream_t st);
// per chunk: int8 slices of the packed FP64 r that K1 wrote
cudaError_t launch_ozaki_slice_r(const double *rP, int npad, int nCblk, uint8_t *RI8, int S, cudaStream_t st);
cudaError_t launch_trmm_i8(const OzakiParams &p, int S, int sm_count, cudaStream_t st);
