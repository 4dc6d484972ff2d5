// pack.cu -- bit-pack an existing [H,W,N] byte canvas (compact mask transport).
// An EXTENSION, not the reference layout: unmold_detections returns bool [H,W,N] (1 byte per
// element, 105 MB per 1024x1024x100 image) and this is what mrx_mask_expand writes.  Callers
// that can take packed masks get   packed[n][y][xb] = np.packbits(masks[:, :, n], axis=1)
// (8 pixels per byte, most significant bit first, rows padded to whole bytes): 8x less
// device -> host traffic; np.unpackbits(packed, axis=-1, count=W).transpose(1, 2, 0) restores
// the reference array exactly.  (mrx_mask_expand_packed, expand_bits.cu, produces the same
// bytes without ever writing the byte canvas; this kernel serves callers that hold one.)
//
// HBM-read bound: N bytes per pixel in, N/8 out.  One CTA = 256 consecutive pixels of one canvas
// row: their 256*N canvas bytes are contiguous (N innermost) and are staged into shared memory
// with 16-byte loads, several in flight per thread.  The transpose to bit planes is byte
// arithmetic, no ballots: a thread takes 8 consecutive pixels x 4 consecutive instances -- eight
// 32-bit shared-memory words, one per pixel, each holding the 0/1 bytes of the four instances --
// and folds them with   acc |= (word & 0x01010101) << (7 - k)   into four output bytes, one
// per instance plane.  Lanes are laid out 4 pixel groups x 8 instance quads so that the eight
// words of a step fall into 32 different banks (pixel-group stride 8*N bytes = 8 banks at
// N = 100, instance-quad stride one bank).  Shapes with N % 4 != 0 or rows that are not
// 4-byte aligned take the same walk one instance (one byte) at a time.
#include "common.cuh"

namespace mrx {

constexpr int kPackThreads = 256;
constexpr int kPackPixels = 256;

__global__ void __launch_bounds__(kPackThreads)
pack_masks_kernel(const unsigned char *__restrict__ canvas, const long long *__restrict__ canvas_off,
                  const int *__restrict__ counts, const int *__restrict__ geom,
                  unsigned char *__restrict__ packed, const long long *__restrict__ packed_off) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int b = blockIdx.z;
  const int H = geom[b * MRX_GEOM_INTS + 0], W = geom[b * MRX_GEOM_INTS + 1];
  const int y = blockIdx.y;
  const int x0 = blockIdx.x * kPackPixels;
  const int N = counts[b];
  if (y >= H || x0 >= W || N <= 0) return;
  const int npx = min(kPackPixels, W - x0);
  const int t = threadIdx.x;

  // ---- stage the npx * N bytes of these pixels at their global address mod 16
  const unsigned char *src = canvas + canvas_off[b] + (static_cast<long long>(y) * W + x0) * N;
  const int nbytes = npx * N;
  const int a = static_cast<int>(reinterpret_cast<uintptr_t>(src) & 15u);
  {
    // every canvas slot starts 16-byte aligned and holds whole 16-byte words (see mrx.h): the
    // first and last word of the run may include neighbouring pixels' bytes, never unmapped memory
    const uint4 *s4 = reinterpret_cast<const uint4 *>(src - a);
    uint4 *d4 = reinterpret_cast<uint4 *>(smem);
    const int n16 = (a + nbytes + 15) >> 4;
    int i = t;
    for (; i + 3 * kPackThreads < n16; i += 4 * kPackThreads) {
      const uint4 v0 = __ldg(s4 + i), v1 = __ldg(s4 + i + kPackThreads),
                  v2 = __ldg(s4 + i + 2 * kPackThreads), v3 = __ldg(s4 + i + 3 * kPackThreads);
      d4[i] = v0;
      d4[i + kPackThreads] = v1;
      d4[i + 2 * kPackThreads] = v2;
      d4[i + 3 * kPackThreads] = v3;
    }
    for (; i < n16; i += kPackThreads) d4[i] = __ldg(s4 + i);
  }
  __syncthreads();
  const unsigned char *px0 = smem + a;   // byte of (pixel x0, instance 0)

  const int wb = (W + 7) >> 3;                        // bytes per packed row
  const long long plane = static_cast<long long>(H) * wb;
  unsigned char *dst = packed + packed_off[b] + static_cast<long long>(y) * wb + (x0 >> 3);
  const int ngroups = (npx + 7) >> 3;                 // 8-pixel groups = output bytes per plane
  const int lane = t & 31, warp = t >> 5;
  const int gs = lane >> 3, qs = lane & 7;            // 4 pixel groups x 8 instance slots per warp step
  if (((N | a) & 3) == 0) {
    // ---- four instances at a time: 32-bit words
    const int nquads = N >> 2;
    const int qblocks = (nquads + 7) >> 3;
    const int gblocks = (ngroups + 3) >> 2;
    for (int step = warp; step < gblocks * qblocks; step += kPackThreads / 32) {
      const int gb = step / qblocks, qb = step - gb * qblocks;
      const int g = gb * 4 + gs, q = qb * 8 + qs;
      if (g >= ngroups || q >= nquads) continue;
      const uint32_t *wp = reinterpret_cast<const uint32_t *>(px0 + static_cast<size_t>(8 * g) * N) + q;
      const int live = min(8, npx - 8 * g);           // pixels of the group inside the row
      const int stride = N >> 2;
      uint32_t acc = 0u;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t w = (k < live) ? wp[k * stride] : 0u;
        acc |= (w & 0x01010101u) << (7 - k);
      }
      unsigned char *o = dst + g + static_cast<long long>(4 * q) * plane;
      o[0] = static_cast<unsigned char>(acc);
      o[plane] = static_cast<unsigned char>(acc >> 8);
      o[2 * plane] = static_cast<unsigned char>(acc >> 16);
      o[3 * plane] = static_cast<unsigned char>(acc >> 24);
    }
  } else {
    // ---- any N, any alignment: one instance per slot
    const int nblocks = (N + 7) >> 3;
    const int gblocks = (ngroups + 3) >> 2;
    for (int step = warp; step < gblocks * nblocks; step += kPackThreads / 32) {
      const int gb = step / nblocks, nb = step - gb * nblocks;
      const int g = gb * 4 + gs, n = nb * 8 + qs;
      if (g >= ngroups || n >= N) continue;
      const unsigned char *bp = px0 + static_cast<size_t>(8 * g) * N + n;
      const int live = min(8, npx - 8 * g);
      unsigned acc = 0u;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const unsigned v = (k < live) ? bp[k * N] : 0u;
        acc |= (v & 1u) << (7 - k);
      }
      dst[g + static_cast<long long>(n) * plane] = static_cast<unsigned char>(acc);
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
  DevInfo dev;
  if (int rc = current_device_info(&dev)) return rc;
  const size_t smem = static_cast<size_t>(kPackPixels) * R + 32;
  MRX_CHECK_SUPPORTED(smem <= static_cast<size_t>(dev.max_smem_optin),
                      "mrx_pack_masks: R=%d needs %zu B of shared memory (limit %d)", R, smem,
                      dev.max_smem_optin);
  static SmemCache cache;
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void *>(pack_masks_kernel), &cache,
                                   dev.device, static_cast<int>(smem)))
    return rc;
  dim3 grid((max_w + kPackPixels - 1) / kPackPixels, max_h, B);
  pack_masks_kernel<<<grid, kPackThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      d_canvas, d_canvas_off, d_counts, d_geom, d_packed, d_packed_off);
  MRX_LAUNCH_CHECK("pack_masks_kernel");
  return MRX_OK;
}
