// expand_ws.cu -- generation 5 of the hot kernel (selected with MRX_EXPAND_IMPL=v5; the default
// is generation 4 in expand_ws4.cu, which is faster this round: 0.79 vs 0.85 ms on config 2).
// Same job as generation 4 -- the hot kernel of api_utils.unmold_detections
// (/root/reference/serve.py:147-154): per-instance 28x28 -> box bilinear resize (zero border,
// half-pixel centres), >= 0.5, paste, np.stack(axis=-1), fused with the canvas zero fill so
// that HBM sees one write per output byte.
//
// One persistent CTA per SM (kWarps warps, see the constants below), warp-specialised;
// hand-offs through mbarriers.  With the constants of this round: 6 consumer groups x 4
// warps, 3 store warps, 3 producers.
//
//   producers    (highest)  fetch work units from a global counter and publish JOB
//                           DESCRIPTORS (one per chunk of canvas): geometry only
//   store warps             when a chunk is complete: fence.proxy.async, ONE bulk (TMA) store
//                           of the chunk shared -> HBM, wait_group.read, re-zero the buffer
//                           (this is the canvas zero fill), hand the buffer back
//   consumers    (lowest)   kGroups groups; a group takes every kGroups-th job.  Each warp owns a
//                           slice of the image's boxes: tests them against the job's rows and
//                           columns (one box per lane), stages the two tile rows each hit
//                           interpolates between with its own 1-D TMA bulk copy
//                           (cp.async.bulk.shared::cluster.global -> warp-private slot +
//                           mbarrier), blends them vertically into ONE REGISTER PER LANE
//                           (lane l = column l-1, zero pad at both ends), then walks the
//                           span 32 columns at a time: exact integer source coordinate
//                           (incremental), horizontal lerp with two warp shuffles, >= 0.5,
//                           st.shared.u8 into the chunk
//
// A job is `chunk_bytes` consecutive bytes of one image's [H,W,N] canvas (N innermost: a pixel
// is N consecutive bytes).  When the row size W*N is a multiple of 16 the chunks are row
// segments ("strip" units: 32 consecutive rows of one segment); otherwise they are flat
// chunks that may start mid-pixel and span rows (units of 8 chunks).  Consumers treat both
// the same way.
//
// Rings in shared memory:
//   descriptors (kNS)      full[s] (producer -> group)      empty[s] (group's warps -> producer)
//   chunk buffers (kNB)    done[b] (last consumer -> store) free[b]  (store warp -> producer)
// Tickets number the jobs CTA-wide: ticket t uses descriptor stage t % kNS, chunk buffer
// t % kNB and consumer group t % kGroups.  Several producers can hold tickets one ring apart
// and an mbarrier parity wait is only unambiguous for the oldest waiter, so every ring slot
// has a generation counter that admits one waiter at a time.
#include <stdlib.h>

#include "expand.cuh"

namespace mrx {

constexpr int kGroups = 6;
constexpr int kGroupWarps = 4;
constexpr int kConsumerWarps = kGroups * kGroupWarps;   // the lowest warp ids
constexpr int kStoreWarps = 3;                          // store warp j owns buffers b % 3 == j
constexpr int kProducers = 3;
constexpr int kWarps = kConsumerWarps + kStoreWarps + kProducers;
constexpr int kThreads = kWarps * 32;
constexpr int kFirstStoreWarp = kConsumerWarps;
constexpr int kFirstProducerWarp = kConsumerWarps + kStoreWarps;
constexpr int kNS = 12;       // descriptor stages (a multiple of kGroups)
constexpr int kMaxNB = 8;     // chunk buffers: template parameter kNB in [2, kMaxNB]
constexpr int kSlots = 4;     // staged tile-row pairs per consumer warp
constexpr int kBandRows = 32;
constexpr int kFlatGroup = 8;
static_assert(kWarps <= 32 && kNS % kGroups == 0, "warp roles / stage ring");

struct __align__(16) JobDesc {
  int valid, buf, len, N;
  int W, g0, g1, r0;      // g0/g1: first/last canvas pixel touched, r0/r1: first/last row
  int r1, sub, img, per;  // sub: bytes of pixel g0 before the chunk; per: boxes per consumer warp
};

struct __align__(16) StoreRec {
  unsigned char *dst;
  int len16;
  int pad_;
};

// number of work units of an image
__device__ __forceinline__ int units_of(int H, int W, int N, int chunk) {
  if (N <= 0) return 0;
  const long long RW = static_cast<long long>(W) * N;
  if ((RW % 16) == 0) {
    const int S = static_cast<int>((RW + chunk - 1) / chunk);
    return S * ((H + kBandRows - 1) / kBandRows);
  }
  const long long jobs = (RW * H + chunk - 1) / chunk;
  return static_cast<int>((jobs + kFlatGroup - 1) / kFlatGroup);
}

template <int kNB>
__global__ void __launch_bounds__(kThreads, 1)
mask_expand_ws_kernel(const ExpandParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int mh = p.mh, mw = p.mw;
  const int slot_floats = 2 * mw;             // two consecutive tile rows
  const uint32_t slot_bytes = slot_floats * 4;
  const int chunk = p.chunk_bytes;

  // ---- carve shared memory
  unsigned char *s_out = smem;                                                 // kNB * chunk
  float *s_slots = reinterpret_cast<float *>(smem + static_cast<size_t>(kNB) * chunk);
  JobDesc *s_desc = reinterpret_cast<JobDesc *>(s_slots + kConsumerWarps * kSlots * slot_floats);
  int *s_uprefix = reinterpret_cast<int *>(s_desc + kNS);                      // B + 1

  __shared__ uint64_t s_full[kNS], s_empty[kNS], s_done[kMaxNB], s_free[kMaxNB];
  __shared__ uint64_t s_wbar[kConsumerWarps];
  __shared__ StoreRec s_store[kMaxNB];
  __shared__ volatile int s_stage_gen[kNS];    // tickets that passed the empty[] wait, per stage
  __shared__ volatile int s_buf_gen[kMaxNB];   // tickets that passed the free[] wait, per buffer
  __shared__ int s_pending[kMaxNB];            // consumer warps still working on the buffer's job
  __shared__ int s_total, s_ticket, s_fin_count;
  __shared__ volatile int s_stop_job;

  // ---- work-unit table, barrier init, canvas zero fill part 1
  if (warp == kFirstProducerWarp) {
    int carry = 0;
    for (int base = 0; base < p.B; base += 32) {
      const int b = base + lane;
      int v = 0;
      if (b < p.B)
        v = units_of(p.geom[b * MRX_GEOM_INTS + 0], p.geom[b * MRX_GEOM_INTS + 1], p.counts[b],
                     chunk);
      int incl = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
      }
      if (b < p.B) s_uprefix[b + 1] = carry + incl;
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) {
      s_uprefix[0] = 0;
      s_total = carry;
    }
  }
  if (warp == kFirstStoreWarp && lane == 0) {
    for (int s = 0; s < kNS; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], kGroupWarps);
      s_stage_gen[s] = 0;
    }
    for (int b = 0; b < kNB; ++b) {
      mbar_init(&s_done[b], 1);
      mbar_init(&s_free[b], 1);
      s_pending[b] = 0;
      s_buf_gen[b] = 0;
    }
    for (int w = 0; w < kConsumerWarps; ++w) mbar_init(&s_wbar[w], 1);
    s_stop_job = -1;
    s_ticket = 0;
    s_fin_count = 0;
    fence_mbar_init();
  }
  {
    // every chunk buffer starts all-zero; the store warps re-zero a buffer after each store
    uint4 *o4 = reinterpret_cast<uint4 *>(s_out);
    const int n16 = (kNB * chunk) >> 4;
    for (int i = tid; i < n16; i += kThreads) o4[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();
  const int total_units = s_total;
  const uint32_t a_full = smem_u32(&s_full[0]), a_empty = smem_u32(&s_empty[0]);
  const uint32_t a_done = smem_u32(&s_done[0]), a_free = smem_u32(&s_free[0]);

  if (warp >= kFirstProducerWarp) {
    // ================================================================= producers
    // Publish one descriptor: take the next ticket, wait for its chunk buffer and descriptor
    // stage, write the record, release it to the consumer group t % kGroups.
    auto publish = [&](bool is_job, unsigned char *dst, JobDesc d) {
      int t = 0;
      if (lane == 0) t = atomicAdd(&s_ticket, 1);
      t = __shfl_sync(0xffffffffu, t, 0);
      const int s = t % kNS, gen = t / kNS;
      const int buf = t % kNB, bgen = t / kNB;
      if (is_job) {
        while (s_buf_gen[buf] != bgen) __nanosleep(32);     // older ticket on this buffer is past its wait
        mbar_wait_a(a_free + 8 * buf, (bgen & 1) ^ 1);      // buffer stored, drained and re-zeroed
      }
      while (s_stage_gen[s] != gen) __nanosleep(32);
      mbar_wait_a(a_empty + 8 * s, (gen & 1) ^ 1);          // consumers are done with the stage
      __syncwarp();
      if (lane == 0) {
        s_stage_gen[s] = gen + 1;
        if (is_job) {
          s_buf_gen[buf] = bgen + 1;
          s_store[buf].dst = dst;
          s_store[buf].len16 = (d.len + 15) & ~15;
          s_pending[buf] = kGroupWarps;
        }
        d.buf = buf;
        s_desc[s] = d;
        mbar_arrive_a(a_full + 8 * s);
      }
      __syncwarp();
      return t;
    };

    int cur_b = 0;
    int unit = 0;
    if (lane == 0) unit = static_cast<int>(atomicAdd(p.job_counter, 1u));
    unit = __shfl_sync(0xffffffffu, unit, 0);
    while (unit < total_units) {
      int next_unit = 0;
      if (lane == 0) next_unit = static_cast<int>(atomicAdd(p.job_counter, 1u));
      while (unit >= s_uprefix[cur_b + 1]) ++cur_b;
      const int b = cur_b;
      const int u_local = unit - s_uprefix[b];
      const int H = p.geom[b * MRX_GEOM_INTS + 0];
      const int W = p.geom[b * MRX_GEOM_INTS + 1];
      const int N = p.counts[b];
      const unsigned RW = static_cast<unsigned>(W) * N;
      const unsigned L = RW * H;                                  // host guarantees < 2^31
      unsigned char *canvas_b = p.canvas + p.canvas_off[b];
      JobDesc d;
      d.valid = 1;
      d.buf = 0;
      d.N = N;
      d.W = W;
      d.img = b;
      d.per = (N + kGroupWarps - 1) / kGroupWarps;

      if ((RW % 16u) == 0u) {
        // ------------------------------------------------------------- strip unit
        const int S = static_cast<int>((RW + chunk - 1) / chunk);
        const int seg_bytes = ((static_cast<int>((RW + S - 1) / S)) + 15) & ~15;
        const int sgm = u_local % S;
        const int band = u_local / S;
        const int ya = band * kBandRows;
        const int yb = min(H, ya + kBandRows);
        const int seg_off = sgm * seg_bytes;
        const int seg_end = min(static_cast<int>(RW), seg_off + seg_bytes);
        const int len = seg_end - seg_off;            // a multiple of 16 (RW and seg_off are)
        if (len > 0) {
          const int xlo = seg_off / N;                // first pixel touched
          const int xhi = (seg_end - 1) / N;          // last pixel touched
          d.len = len;
          d.sub = seg_off - xlo * N;                  // bytes of pixel xlo before the segment
          for (int y = ya; y < yb; ++y) {
            d.g0 = y * W + xlo;
            d.g1 = y * W + xhi;
            d.r0 = y;
            d.r1 = y;
            publish(true, canvas_b + static_cast<unsigned>(y) * RW + seg_off, d);
          }
        }
      } else {
        // ------------------------------------------------------------- flat unit
        const int jobs_b = static_cast<int>((static_cast<unsigned long long>(L) + chunk - 1) / chunk);
        const int j_end = min(jobs_b, (u_local + 1) * kFlatGroup);
        for (int j = u_local * kFlatGroup; j < j_end; ++j) {
          const unsigned c0 = static_cast<unsigned>(j) * chunk;
          d.len = static_cast<int>(min(static_cast<unsigned>(chunk), L - c0));
          d.g0 = static_cast<int>(c0 / N);
          d.g1 = static_cast<int>((c0 + d.len - 1) / N);
          d.r0 = d.g0 / W;
          d.r1 = d.g1 / W;
          d.sub = static_cast<int>(c0 - static_cast<unsigned>(d.g0) * N);
          publish(true, canvas_b + c0, d);
        }
      }
      unit = __shfl_sync(0xffffffffu, next_unit, 0);
    }
    // the last producer to run dry publishes one stop record per consumer group (consecutive
    // tickets reach every group) and tells the store warps how many jobs exist
    int fin = 0;
    if (lane == 0) fin = atomicAdd(&s_fin_count, 1);
    fin = __shfl_sync(0xffffffffu, fin, 0);
    if (fin == kProducers - 1) {
      JobDesc d;
      d.valid = 0;
      d.buf = d.len = d.N = d.W = d.g0 = d.g1 = d.r0 = d.r1 = d.sub = d.img = d.per = 0;
      for (int g = 0; g < kGroups; ++g) {
        const int t = publish(false, nullptr, d);
        if (g == 0 && lane == 0) s_stop_job = t;   // tickets >= t are not jobs
      }
    }
  } else if (warp >= kFirstStoreWarp) {
    // ================================================================= store warps
    // store warp j owns the chunk buffers b with b % kStoreWarps == j (every done[] barrier has
    // a single waiter that sees each of its phases): store, drain and re-zero run in parallel
    // on different buffers
    const int me = warp - kFirstStoreWarp;
    for (int k = 0;; ++k) {
      const int b = k % kNB;
      if ((b % kStoreWarps) != me) continue;
      // wait for job k, or learn that it does not exist (real jobs always complete done[])
      bool have = false;
#pragma unroll 1
      while (true) {
        have = mbar_try_wait_a(a_done + 8 * b, (k / kNB) & 1, 1000u);
        if (have) break;
        const int stop = s_stop_job;
        if (stop >= 0 && k >= stop) break;
      }
      if (!have) break;
      const StoreRec rec = s_store[b];
      unsigned char *buf = s_out + static_cast<size_t>(b) * chunk;
      if (lane == 0) {
        fence_proxy_async_smem();
        if (!(p.flags & 0x400)) {
          bulk_s2g(rec.dst, buf, static_cast<uint32_t>(rec.len16));
          bulk_commit();
          bulk_wait_read<0>();                   // the chunk has left shared memory
        }
      }
      __syncwarp();
      // canvas zero fill, part 2: the buffer goes back to the pool all-zero
      if (!(p.flags & 0x200)) {
        uint4 *o4 = reinterpret_cast<uint4 *>(buf);
        const int n16 = rec.len16 >> 4;
        int i = lane;
        for (; i + 96 < n16; i += 128) {
          o4[i] = make_uint4(0u, 0u, 0u, 0u);
          o4[i + 32] = make_uint4(0u, 0u, 0u, 0u);
          o4[i + 64] = make_uint4(0u, 0u, 0u, 0u);
          o4[i + 96] = make_uint4(0u, 0u, 0u, 0u);
        }
        for (; i < n16; i += 32) o4[i] = make_uint4(0u, 0u, 0u, 0u);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive_a(a_free + 8 * b);
    }
    if (lane == 0) bulk_wait_all<0>();
  } else {
    // ================================================================= consumers
    const int grp = warp / kGroupWarps;
    const int gw = warp - grp * kGroupWarps;
    float *my_slots = s_slots + warp * kSlots * slot_floats;
    const uint32_t my_slots_a = smem_u32(my_slots);
    const uint32_t my_bar = smem_u32(&s_wbar[warp]);
    const unsigned lt_mask = (1u << lane) - 1u;
    uint32_t wpar = 0;
    int c_img = -1, c_base = -1, c_hi = -1;   // which boxes the registers below hold
    int4 bx = make_int4(0, 0, 0, 0);
    BoxAux ax;
    ax.D = 0; ax.invD = 0.f; ax.stepQ = 0; ax.stepR = 0;
    int s = grp - kGroups;        // this group's tickets: grp, grp + kGroups, ...
    uint32_t fpar = 0;
    while (true) {
      s += kGroups;
      if (s >= kNS) {
        s -= kNS;
        fpar ^= 1;
      }
      mbar_wait_a(a_full + 8 * s, fpar);
      const JobDesc d = s_desc[s];
      if (!d.valid) break;
      const uint32_t out_addr = smem_u32(s_out + static_cast<size_t>(d.buf) * chunk);
      const float *tiles_b = p.tiles + static_cast<size_t>(d.img) * p.R * mh * mw;
      const int4 *boxes_b = p.boxes + static_cast<size_t>(d.img) * p.R;
      const BoxAux *aux_b = p.aux + static_cast<size_t>(d.img) * p.R;
      const unsigned ulen = static_cast<unsigned>(d.len);
      const unsigned ostep = 32u * d.N;
      const int n_lo = gw * d.per;
      const int n_hi = (p.flags & 0x100) ? 0 : min(d.N, n_lo + d.per);
      for (int base = n_lo; base < n_hi; base += 32) {
        // this warp's boxes, one per lane
        const int n = base + lane;
        const bool have = n < n_hi;
        if (d.img != c_img || base != c_base || n_hi != c_hi) {
          // (re)load this warp's boxes; consecutive jobs of an image reuse the registers
          bx = make_int4(0, 0, 0, 0);
          ax.D = 0;
          if (have) {
            bx = __ldg(boxes_b + n);
            const int4 raw = __ldg(reinterpret_cast<const int4 *>(aux_b + n));
            ax.D = raw.x;
            ax.invD = __int_as_float(raw.y);
            ax.stepQ = raw.z;
            ax.stepR = raw.w;
          }
          c_img = d.img;
          c_base = base;
          c_hi = n_hi;
        }
        for (int row = d.r0; row <= d.r1; ++row) {
          const int xlo = max(0, d.g0 - row * d.W);
          const int xhi = min(d.W, d.g1 + 1 - row * d.W);
          const int xa = max(xlo, bx.y);
          const int xb = min(xhi, bx.w);
          const bool valid = have && ax.D > 0 && row >= bx.x && row < bx.z && xa < xb;
          // vertical source coordinate (lane-parallel, one box per lane):
          //   src = (mh*(2*(row-y1)+1) - bh) / (2*bh), j0 = floor, wy = fraction
          int jc = 0, otop = -1, obot = -1;
          float wy = 0.f;
          if (valid) {
            const int bh = bx.z - bx.x;
            const int Dy = 2 * bh;
            const int Ay = mh * (2 * (row - bx.x) + 1) - bh;
            int j0 = __float2int_rd(static_cast<float>(Ay) * __frcp_rn(static_cast<float>(Dy)));
            int remy = Ay - j0 * Dy;
            if (remy < 0) {
              --j0;
              remy += Dy;
            } else if (remy >= Dy) {
              ++j0;
              remy -= Dy;
            }
            jc = min(max(j0, 0), mh - 2);              // staged rows: jc, jc + 1
            wy = __fdiv_rn(static_cast<float>(remy), static_cast<float>(Dy));
            otop = (j0 < 0) ? -1 : (j0 - jc) * mw;
            obot = (j0 + 1 > mh - 1) ? -1 : (j0 + 1 - jc) * mw;
          }
          const int obase = (row * d.W - d.g0) * d.N + n - d.sub;
          unsigned mask = __ballot_sync(0xffffffffu, valid);
          while (mask) {
            // ---- stage up to kSlots hits: each lane issues its own 1-D TMA copy
            const int rank = __popc(mask & lt_mask);
            const bool mine = ((mask >> lane) & 1u) && rank < kSlots;
            const int cnt = min(__popc(mask), kSlots);
            if (!(p.flags & 0x2000)) {
              if (lane == 0) mbar_arrive_expect_tx(&s_wbar[warp], cnt * slot_bytes);
              if (mine)
                bulk_g2s_a(my_slots_a + rank * slot_bytes, tiles_b + (n * mh + jc) * mw, slot_bytes,
                           my_bar);
              mbar_wait_a(my_bar, wpar);
              wpar ^= 1;
            }
            // ---- sample them in lane order
            for (int k = 0; k < cnt; ++k) {
              const int src = __ffs(mask) - 1;
              mask &= mask - 1;
              const int e_x1 = __shfl_sync(0xffffffffu, bx.y, src);
              const int e_xa = __shfl_sync(0xffffffffu, xa, src);
              const int e_xb = __shfl_sync(0xffffffffu, xb, src);
              const int e_D = __shfl_sync(0xffffffffu, ax.D, src);
              const float e_invD = __shfl_sync(0xffffffffu, ax.invD, src);
              const int e_stepQ = __shfl_sync(0xffffffffu, ax.stepQ, src);
              const int e_stepR = __shfl_sync(0xffffffffu, ax.stepR, src);
              const float e_wy = __shfl_sync(0xffffffffu, wy, src);
              const int e_otop = __shfl_sync(0xffffffffu, otop, src);
              const int e_obot = __shfl_sync(0xffffffffu, obot, src);
              const int e_obase = __shfl_sync(0xffffffffu, obase, src);
              const float *slot = my_slots + k * slot_floats;
              // vertical blend of the two staged rows, one column per lane:
              // lane l holds B[l] with B[0] = 0, B[1+i] = blend(i), B[mw+1] = 0   (mw <= 30)
              float bl = 0.f;
              {
                const int i = lane - 1;
                if (i >= 0 && i < mw) {
                  const float top = (e_otop >= 0) ? slot[e_otop + i] : 0.f;
                  const float bot = (e_obot >= 0) ? slot[e_obot + i] : 0.f;
                  bl = fmaf(e_wy, bot - top, top);
                }
              }
              // exact source column of this lane's first pixel, then 32 pixels per step
              int x = e_xa + lane;
              int idx, rem;
              {
                const int A = mw * (2 * (x - e_x1) + 1) - (e_D >> 1);
                int i0 = __float2int_rd(static_cast<float>(A) * e_invD);
                rem = A - i0 * e_D;
                if (rem < 0) {
                  --i0;
                  rem += e_D;
                } else if (rem >= e_D) {
                  ++i0;
                  rem -= e_D;
                }
                idx = i0 + 1;   // B[idx], B[idx+1] are the two taps
              }
              unsigned off = static_cast<unsigned>(e_obase + x * d.N);
              for (int xs = e_xa; xs < e_xb; xs += 32) {   // warp-uniform trip count (shuffles inside)
                const float wx = static_cast<float>(rem) * e_invD;
                const float a = __shfl_sync(0xffffffffu, bl, idx);
                const float bq = __shfl_sync(0xffffffffu, bl, idx + 1);
                const float v = fmaf(wx, bq - a, a);
                if (v >= 0.5f && x < e_xb && off < ulen)
                  asm volatile("st.shared.u8 [%0], %1;" ::"r"(out_addr + off), "r"(1u) : "memory");
                x += 32;
                off += ostep;
                rem += e_stepR;
                idx += e_stepQ;
                if (rem >= e_D) {
                  rem -= e_D;
                  ++idx;
                }
              }
            }
            __syncwarp();   // all lanes are done with the slots before the next batch lands
          }
        }
      }
      fence_proxy_async_smem();   // chunk bytes must be visible to the bulk store
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_a(a_empty + 8 * s);
        // whoever reports last on the job releases the chunk to its store warp
        if (atomicSub(&s_pending[d.buf], 1) == 1) mbar_arrive_a(a_done + 8 * d.buf);
      }
    }
  }
}

int launch_expand_ws(const ExpandParams &prm, int sms, int max_smem_optin, cudaStream_t st) {
  const size_t fixed = static_cast<size_t>(kConsumerWarps) * kSlots * 2 * prm.mw * sizeof(float) +
                       kNS * sizeof(JobDesc) + static_cast<size_t>(prm.B + 1) * sizeof(int);
  const size_t reserve = 2048;   // static shared memory (barriers, records) + alignment
  MRX_CHECK_SUPPORTED(fixed + reserve + 2 * static_cast<size_t>(prm.chunk_bytes) <=
                          static_cast<size_t>(max_smem_optin),
                      "mrx_mask_expand: chunk_bytes %d too large for %d B of shared memory",
                      prm.chunk_bytes, max_smem_optin);
  int nb = static_cast<int>((static_cast<size_t>(max_smem_optin) - fixed - reserve) /
                            prm.chunk_bytes);
  if (nb > kMaxNB) nb = kMaxNB;
  {
    const char *e = getenv("MRX_EXPAND_NB");
    if (e && atoi(e) >= 2 && atoi(e) < nb) nb = atoi(e);
  }
  const size_t smem = static_cast<size_t>(nb) * prm.chunk_bytes + fixed;
  void (*kern)(const ExpandParams) = nullptr;
  switch (nb) {
    case 2: kern = mask_expand_ws_kernel<2>; break;
    case 3: kern = mask_expand_ws_kernel<3>; break;
    case 4: kern = mask_expand_ws_kernel<4>; break;
    case 5: kern = mask_expand_ws_kernel<5>; break;
    case 6: kern = mask_expand_ws_kernel<6>; break;
    case 7: kern = mask_expand_ws_kernel<7>; break;
    default: kern = mask_expand_ws_kernel<8>; break;
  }
  MRX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(smem)));
  kern<<<sms, kThreads, smem, st>>>(prm);
  MRX_LAUNCH_CHECK("mask_expand_ws_kernel");
  return MRX_OK;
}

}  // namespace mrx
