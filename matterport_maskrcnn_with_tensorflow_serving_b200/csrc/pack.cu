// pack.cu -- bit-packed mask output (compact mask transport).
// An EXTENSION, not the reference layout: unmold_detections returns bool [H,W,N] (1 byte per
// element, 105 MB per 1024x1024x100 image) and this is what mrx_mask_expand writes.  Callers
// that can take packed masks get   packed[n][y][xb] = np.packbits(masks[:, :, n], axis=1)
// (8 pixels per byte, most significant bit first, rows padded to whole bytes): 8x less
// device -> host traffic; np.unpackbits(packed, axis=-1, count=W).transpose(1, 2, 0) restores
// the reference array exactly.
//
// One CTA = 256 consecutive pixels of one canvas row: their 256*N canvas bytes are contiguous
// (N innermost) and are staged into shared memory; then, per instance n, warp w reads the bytes
// of its 32 pixels (stride N: conflict-free for N = 100), ballots them into a 32-bit word,
// reverses it into packbits order and stores 4 bytes; the 8 warps of the CTA fill one 32-byte
// sector of packed[n][y][.] per instance.  HBM-read bound: N bytes per pixel in, N/8 out.
#include "common.cuh"

namespace mrx {

constexpr int kPackThreads = 256;

__global__ void __launch_bounds__(kPackThreads)
pack_masks_kernel(const unsigned char *__restrict__ canvas, const long long *__restrict__ canvas_off,
                  const int *__restrict__ counts, const int *__restrict__ geom,
                  unsigned char *__restrict__ packed, const long long *__restrict__ packed_off) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int b = blockIdx.z;
  const int H = geom[b * MRX_GEOM_INTS + 0], W = geom[b * MRX_GEOM_INTS + 1];
  const int y = blockIdx.y;
  const int x0 = blockIdx.x * kPackThreads;
  const int N = counts[b];
  if (y >= H || x0 >= W || N <= 0) return;
  const int npx = min(kPackThreads, W - x0);
  const int t = threadIdx.x;
  const int lane = t & 31, warp = t >> 5;

  // stage the npx * N bytes of these pixels: 16-byte loads when the run is 16-byte aligned
  const unsigned char *src = canvas + canvas_off[b] + (static_cast<long long>(y) * W + x0) * N;
  const int nbytes = npx * N;
  if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0u) {
    const int n16 = nbytes >> 4;
    const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
    uint4 *d4 = reinterpret_cast<uint4 *>(smem);
    for (int i = t; i < n16; i += kPackThreads) d4[i] = __ldg(s4 + i);
    for (int i = (n16 << 4) + t; i < nbytes; i += kPackThreads) smem[i] = src[i];
  } else {
    for (int i = t; i < nbytes; i += kPackThreads) smem[i] = src[i];
  }
  __syncthreads();

  const int wb = (W + 7) >> 3;                       // bytes per packed row
  const int px = warp * 32 + lane;                   // this lane's pixel inside the CTA
  const bool valid = px < npx;
  const int bx = (x0 >> 3) + warp * 4;               // first packed byte of this warp's 32 pixels
  const int nb = min(4, wb - bx);                    // bytes of it inside the row (<= 0: none)
  unsigned char *dst = packed + packed_off[b] + (static_cast<long long>(y) * wb + bx);
  const long long plane = static_cast<long long>(H) * wb;
  const unsigned char *mine = smem + static_cast<size_t>(px) * N;
  // one 4-byte store per (warp, instance) when the warp's four bytes exist and every plane's
  // copy of them is 4-byte aligned; byte stores otherwise (row ends, odd row pitches)
  const bool word_ok = nb == 4 && ((reinterpret_cast<uintptr_t>(dst) | static_cast<uintptr_t>(plane)) & 3u) == 0u;
  for (int n = 0; n < N; ++n) {
    const unsigned bit = valid ? (mine[n] != 0) : 0u;
    const unsigned bal = __ballot_sync(0xffffffffu, bit);
    // ballot bit l = pixel l; packbits puts pixel 0 in the most significant bit of byte 0
    const unsigned rev = __brev(bal);                // bit 31 = pixel 0
    if (word_ok) {
      // memory order byte0..byte3 = rev bits 31..24, 23..16, 15..8, 7..0
      if (lane == 0) *reinterpret_cast<unsigned *>(dst + n * plane) = __byte_perm(rev, 0u, 0x0123);
    } else if (lane < nb) {
      dst[n * plane + lane] = static_cast<unsigned char>(rev >> (24 - 8 * lane));
    }
  }
}

}  // namespace mrx

extern "C" int mrx_pack_masks(const unsigned char *d_canvas, const long long *d_canvas_off,
                              const int *d_counts, const int *d_geom, unsigned char *d_packed,
                              const long long *d_packed_off, int B, int R, int max_h, int max_w,
                              void *stream) {
  using namespace mrx;
  MRX_CHECK_ARG(d_canvas && d_canvas_off && d_counts && d_geom && d_packed && d_packed_off,
                "mrx_pack_masks: null pointer");
  MRX_CHECK_ARG(B >= 0 && B <= MRX_MAX_BATCH && R >= 1 && max_h >= 0 && max_w >= 0,
                "mrx_pack_masks: bad sizes B=%d R=%d", B, R);
  if (B == 0 || max_h == 0 || max_w == 0) return MRX_OK;
  MRX_CHECK_SUPPORTED(max_h <= 65535, "mrx_pack_masks: image taller than 65535 rows");
  int dev = 0, max_optin = 0;
  MRX_CUDA(cudaGetDevice(&dev));
  MRX_CUDA(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  const size_t smem = static_cast<size_t>(kPackThreads) * R + 16;
  MRX_CHECK_SUPPORTED(smem <= static_cast<size_t>(max_optin),
                      "mrx_pack_masks: R=%d needs %zu B of shared memory (limit %d)", R, smem,
                      max_optin);
  MRX_CUDA(cudaFuncSetAttribute(pack_masks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(smem)));
  dim3 grid((max_w + kPackThreads - 1) / kPackThreads, max_h, B);
  pack_masks_kernel<<<grid, kPackThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      d_canvas, d_canvas_off, d_counts, d_geom, d_packed, d_packed_off);
  MRX_LAUNCH_CHECK("pack_masks_kernel");
  return MRX_OK;
}
