// expand.cuh -- shared declarations of the mask-expand kernels (csrc/unmold.cu: generic
// kernel + C ABI; csrc/expand_team.cu: the default team kernel; csrc/expand_ws4.cu: the
// warp-specialised generation 4).
#pragma once

#include "common.cuh"

namespace mrx {

// Per kept box, written by unmold_prologue_kernel: the row-invariant constants of the
// horizontal source coordinate  src = (mw*(2*(x-x1)+1) - bw) / (2*bw).
struct __align__(16) BoxAux {
  int D;        // 2 * box width; 0 marks a box the expand kernels must skip (outside the canvas)
  float invD;   // 1 / D
  int stepQ;    // (64*mw) / D : source-column advance per 32 canvas columns
  int stepR;    // (64*mw) % D
};

struct ExpandParams {
  const float *tiles;           // [B,R,mh,mw]
  const int4 *boxes;            // [B,R] (y1,x1,y2,x2)
  const BoxAux *aux;            // [B,R]
  const int *counts;            // [B]
  const int *geom;              // [B,8]
  const long long *canvas_off;  // [B]
  unsigned char *canvas;
  unsigned int *job_counter;
  int B, R, mh, mw, chunk_bytes;
  int flags;                    // development switches (MRX_EXPAND_FLAGS)
};

// Launch the warp-specialised kernel (one persistent CTA per SM).  Returns MRX_OK or an error
// code with mrx_last_error() set.  Requires mw <= 30.
int launch_expand_ws4(const ExpandParams &prm, int sms, int max_smem_optin, cudaStream_t st);   // gen 4
// Generation 6 (default): teams of warps build 2-D tiles.  want_buf = upper bound of a team's
// tile buffer in bytes (0 = as large as fits).  MRX_E_UNSUPPORTED when R does not fit a buffer.
int launch_expand_team(const ExpandParams &prm, int sms, int max_smem_optin, int want_buf,
                       cudaStream_t st);

}  // namespace mrx
