// expand.cuh -- shared declarations of the mask-expand kernels (csrc/unmold.cu: generic
// kernel + C ABI; csrc/expand_team.cu: the default team kernel; csrc/expand_bits.cu: the
// bit-packed output).
#pragma once

#include "common.cuh"

namespace mrx {

struct ExpandParams {
  const float *tiles;           // [B,R,mh,mw]
  const int *tile_index;        // [B,R] tile of kept instance k = tiles[b][tile_index[b][k]]; NULL: k
  const int4 *boxes;            // [B,R] (y1,x1,y2,x2)
  const int *counts;            // [B]
  const int *geom;              // [B,8]
  const long long *canvas_off;  // [B]
  unsigned char *canvas;
  unsigned int *job_counter;    // [0] tile ticket, [1] teams / CTAs retired; zero between launches
  float *values;                // test instantiation only: pre-threshold samples, indexed like canvas
  int B, R, mh, mw, chunk_bytes;
  int flags;                    // development switches (MRX_EXPAND_FLAGS; -DMRX_DEV builds only)
};

// The default kernel: teams of warps build 2-D tiles (one persistent CTA per SM).  want_buf =
// upper bound of a team's tile buffer in bytes (0 = as large as fits).  Returns MRX_OK, an
// error code with mrx_last_error() set, or MRX_E_UNSUPPORTED when R does not fit a buffer or
// mw > 30 (the caller then takes the generic kernel).
int launch_expand_team(const ExpandParams &prm, const DevInfo &dev, int want_buf, cudaStream_t st);

}  // namespace mrx
