// pack.cu -- bit-pack an existing [H,W,N] byte canvas (compact mask transport).
// An EXTENSION, not the reference layout: unmold_detections returns bool [H,W,N] (1 byte per
// element, 105 MB per 1024x1024x100 image) and this is what mrx_mask_expand writes.  Callers
// that can take packed masks get   packed[n][y][xb] = np.packbits(masks[:, :, n], axis=1)
// (8 pixels per byte, most significant bit first, rows padded to whole bytes): 8x less
// device -> host traffic; np.unpackbits(packed, axis=-1, count=W).transpose(1, 2, 0) restores
// the reference array exactly.  (mrx_mask_expand_packed, expand_bits.cu, produces the same
// bytes without ever writing the byte canvas; this kernel serves callers that hold one.)
//
// HBM-read bound: N bytes per pixel in, N/8 out.  Two kernels, each image goes to one of them:
//
// pack_quads_kernel (N % 4 == 0, e.g. the full 100 instances): no shared memory at all.  A thread
// owns 32 consecutive pixels x 4 consecutive instances: thirty-two independent 4-byte loads
// (the 0/1 bytes of its four instances at each pixel), folded with
//     acc[b] |= (word & 0x01010101) << (7 - k)        pixel 8*b + k of the thread's 32
// into one output byte per (8 pixels, instance), transposed with eight byte permutes into one
// 32-bit word per instance plane and stored.  Lanes are laid out 4 pixel runs x 8 instance
// quads, so every load instruction of a warp reads four fully used 32-byte sectors and every
// store writes 16 contiguous bytes per plane.  (A first version staged 256 pixels in shared
// memory and re-read them: 2 x 25.6 KB of shared-memory traffic per 25.6 KB of canvas put it at
// the shared-memory bandwidth, 0.42 of the HBM roofline.)
//
// pack_bytes_kernel (any N, any alignment; ragged instance counts): 256 consecutive pixels of a
// row staged in shared memory at their global address mod 16, then one instance per lane slot.
#include "common.cuh"

namespace mrx {

constexpr int kPackThreads = 256;
constexpr int kPackPixels = 256;

__global__ void __launch_bounds__(kPackThreads)
pack_quads_kernel(const unsigned char *__restrict__ canvas, const long long *__restrict__ canvas_off,
                  const int *__restrict__ counts, const int *__restrict__ geom,
                  unsigned char *__restrict__ packed, const long long *__restrict__ packed_off) {
  const int b = blockIdx.z;
  const int H = geom[b * MRX_GEOM_INTS + 0], W = geom[b * MRX_GEOM_INTS + 1];
  const int y = blockIdx.y;
  const int N = counts[b];
  if (y >= H || N <= 0 || (N & 3) != 0) return;   // other N: pack_bytes_kernel
  const int lane = threadIdx.x & 31;
  const int nquads = N >> 2;
  const int qblocks = (nquads + 7) >> 3;             // 8 instance quads per warp
  const int pblocks = (W + 127) >> 7;                // 4 runs of 32 pixels per warp
  const int wid = blockIdx.x * (kPackThreads / 32) + (threadIdx.x >> 5);
  if (wid >= pblocks * qblocks) return;
  const int pb = wid / qblocks, qb = wid - pb * qblocks;
  const int run = pb * 4 + (lane >> 3);              // which run of 32 pixels of the row
  const int q = qb * 8 + (lane & 7);
  const int x0 = run << 5;
  if (x0 >= W || q >= nquads) return;
  const int npx = min(32, W - x0);
  // N % 4 == 0 and 16-byte aligned slots: every (pixel, quad) word is 4-byte aligned
  const uint32_t *src = reinterpret_cast<const uint32_t *>(
                            canvas + canvas_off[b] + (static_cast<long long>(y) * W + x0) * N) + q;
  uint32_t acc[4];
  const unsigned stride = static_cast<unsigned>(nquads);   // words per pixel (32-bit index math)
  const unsigned stride_bytes = static_cast<unsigned>(N);
  const unsigned char *srcb = reinterpret_cast<const unsigned char *>(src);
  if (npx == 32) {
    // whole run inside the row (always, when W % 32 == 0): 32 unpredicated loads
#pragma unroll
    for (int bq = 0; bq < 4; ++bq) {
      uint32_t w[8];
#pragma unroll
      for (int k = 0; k < 8; ++k)   // one 32x32+64 multiply-add per address
        w[k] = __ldg(reinterpret_cast<const uint32_t *>(
            srcb + static_cast<unsigned long long>(8 * bq + k) * static_cast<unsigned long long>(stride_bytes)));
      uint32_t a = 0u;
#pragma unroll
      for (int k = 0; k < 8; ++k) a |= (w[k] & 0x01010101u) << (7 - k);
      acc[bq] = a;
    }
  } else {
#pragma unroll
    for (int bq = 0; bq < 4; ++bq) {
      uint32_t a = 0u;
      for (int k = 0; k < 8; ++k) {
        const int px = 8 * bq + k;
        const uint32_t w = px < npx ? __ldg(src + static_cast<unsigned>(px) * stride) : 0u;
        a |= (w & 0x01010101u) << (7 - k);
      }
      acc[bq] = a;
    }
  }
  // acc[bq] byte j = packed byte bq of instance 4q + j  ->  word of instance j = bytes 0..3
  const uint32_t t0 = __byte_perm(acc[0], acc[1], 0x5140), t1 = __byte_perm(acc[2], acc[3], 0x5140);
  const uint32_t t2 = __byte_perm(acc[0], acc[1], 0x7362), t3 = __byte_perm(acc[2], acc[3], 0x7362);
  const uint32_t out[4] = {__byte_perm(t0, t1, 0x5410), __byte_perm(t0, t1, 0x7632),
                           __byte_perm(t2, t3, 0x5410), __byte_perm(t2, t3, 0x7632)};
  const int wb = (W + 7) >> 3;
  const long long plane = static_cast<long long>(H) * wb;
  unsigned char *dst = packed + packed_off[b] + (static_cast<long long>(4 * q) * H + y) * wb + (x0 >> 3);
  const int nb = min(4, wb - (x0 >> 3));             // bytes of this run inside the packed row
  const bool word_ok = nb == 4 && ((reinterpret_cast<uintptr_t>(dst) | static_cast<uintptr_t>(plane)) & 3u) == 0u;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    unsigned char *o = dst + j * plane;
    if (word_ok) {
      *reinterpret_cast<uint32_t *>(o) = out[j];
    } else {
      for (int t = 0; t < nb; ++t) o[t] = static_cast<unsigned char>(out[j] >> (8 * t));
    }
  }
}

// A fixed number of CTAs walks every image: an image packed by pack_quads_kernel (N % 4 == 0)
// costs each CTA one test, not one empty CTA per 256 pixels.
__global__ void __launch_bounds__(kPackThreads)
pack_bytes_kernel(const unsigned char *__restrict__ canvas, const long long *__restrict__ canvas_off,
                  const int *__restrict__ counts, const int *__restrict__ geom,
                  unsigned char *__restrict__ packed, const long long *__restrict__ packed_off,
                  int B) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int t = threadIdx.x;
  // jobs (256 pixels of one row) are numbered across all the images this kernel packs and dealt
  // round robin to the CTAs: no per-image tail
  int base = 0;
  int job = blockIdx.x;
  for (int b = 0; b < B; ++b) {
  const int N = counts[b];
  if (N <= 0 || (N & 3) == 0) continue;   // N % 4 == 0: pack_quads_kernel
  const int H = geom[b * MRX_GEOM_INTS + 0], W = geom[b * MRX_GEOM_INTS + 1];
  const int xblocks = (W + kPackPixels - 1) / kPackPixels;
  const int njobs = H * xblocks;
  for (; job < base + njobs; job += gridDim.x) {
  const int local = job - base;
  const int y = local / xblocks;
  const int x0 = (local - y * xblocks) * kPackPixels;
  const int npx = min(kPackPixels, W - x0);
  __syncthreads();   // the previous job's readers are done with the staging buffer

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
  const int gs = lane >> 3, qs = lane & 7;            // 4 pixel groups x 8 instances per warp step
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
  }   // jobs of image b
  base += njobs;
  }   // images
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
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void *>(pack_bytes_kernel), &cache,
                                   dev.device, static_cast<int>(smem)))
    return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // each image is packed by exactly one of the two kernels (N % 4 == 0 or not); the other one's
  // CTAs for that image return at once
  {
    const int max_quads = (R + 3) >> 2;
    const int warps_per_row = ((max_w + 127) >> 7) * ((max_quads + 7) >> 3);
    dim3 grid((warps_per_row + kPackThreads / 32 - 1) / (kPackThreads / 32), max_h, B);
    pack_quads_kernel<<<grid, kPackThreads, 0, st>>>(d_canvas, d_canvas_off, d_counts, d_geom,
                                                     d_packed, d_packed_off);
    MRX_LAUNCH_CHECK("pack_quads_kernel");
  }
  int occ = 0;
  MRX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pack_bytes_kernel, kPackThreads, smem));
  pack_bytes_kernel<<<dev.sms * max(occ, 1), kPackThreads, smem, st>>>(d_canvas, d_canvas_off, d_counts,
                                                                    d_geom, d_packed, d_packed_off, B);
  MRX_LAUNCH_CHECK("pack_bytes_kernel");
  return MRX_OK;
}
