// expand_ws4.cu -- warp-specialised mask-expand kernel, generation 4 (the default):
// producers list the (box,row) entries of every chunk and stage their tile rows by TMA into
// a ring of item stages ahead of the consumers; see the block comment below.
// Kept selectable (MRX_EXPAND_IMPL=v4) for comparison with the default team kernel
// (expand_team.cu); generation 5 (descriptor-only producers, 0.85 ms) was removed.
#include <stdlib.h>
#include <string.h>

#include "expand.cuh"

namespace mrx {

namespace ws4 {

// =====================================================================================
// mask expand, warp-specialised: ONE persistent CTA per SM, 32 warps
// =====================================================================================
//
//   warps 29-31 producers: fetch work units, list the (box,row) entries of each chunk, issue
//                          the 1-D TMA loads of their tile rows, cut spans into units
//   warps 27,28 store    : when a chunk is complete, one bulk (TMA) store shared -> HBM,
//                          then re-zeroes the buffer (this is the canvas zero fill)
//   warps 0..26 consumers: kGroups groups of kGroupWarps; a group takes every kGroups-th
//                          item; a warp takes span units: vertical blend of the two staged
//                          tile rows into registers (one column per lane), horizontal lerp
//                          by warp shuffle, >= 0.5, byte store into the shared chunk
//
// Rings in shared memory, all hand-offs through mbarriers (no __syncthreads in steady state):
//   item stages  (kNS): staging rows + entries + units + header     full[s] / empty[s]
//   chunk buffers(kNB): the bytes of one job                        done[b] / free[b]
// A job is one chunk of a canvas; an item is one pass over <= kWsEMax entries and <= kUMax
// span units of a job (first item zero-fills the chunk, last item releases it to the store
// warp).
//
// Work units handed out by the global counter:
//   strip mode (row bytes W*N a multiple of 16): a unit is kBandRows consecutive rows of one
//     row segment.  The boxes that can touch the segment are listed once per unit (active
//     list, spans clipped once); per row only the vertical coordinate changes.
//   flat mode (any shape): a unit is kFlatGroup consecutive flat chunks; every chunk tests
//     all (row, box) pairs.  Chunks may start mid-pixel and span rows.
constexpr int kWsWarps = 32;   // 27 consumers + 2 store + 3 producers
constexpr int kWsThreads = kWsWarps * 32;
constexpr int kGroups = 3;                          // consumer groups
constexpr int kGroupWarps = 9;                      // warps per consumer group
constexpr int kWsConsumerWarps = kGroups * kGroupWarps;
constexpr int kStoreWarps = 2;                      // each owns the chunk buffers b % kStoreWarps
constexpr int kProducers = 3;                       // producer warps (highest warp ids)
constexpr int kWsEMax = 32;   // entries per item (ws kernel)
constexpr int kMaxNB = 6;   // chunk buffers: template parameter kNB in [2, kMaxNB]
constexpr int kNS = 6;   // item stages: two per consumer group
constexpr int kUMax = 256;
constexpr int kMaxUnitsPerEntry = kUMax / 32;
constexpr int kBoxCache = 128;   // per-producer box / aux / active tables
constexpr int kBandRows = 8;
constexpr int kFlatGroup = 8;
constexpr int kMinUnitShift = 3;   // span units are at least 256 columns (8 sampling steps)
constexpr int kFirstProducerWarp = kWsWarps - kProducers;   // highest warp ids: favoured by the issue arbiter
constexpr int kFirstStoreWarp = kFirstProducerWarp - kStoreWarps;
static_assert(kNS % kGroups == 0 && kNS >= kProducers + 1, "stage ring must interleave the groups");
static_assert(kWsConsumerWarps + kStoreWarps + kProducers == kWsWarps, "warp roles");

struct __align__(16) WsEntry {
  int obase;    // byte offset of (row, x=0, n) relative to the chunk start
  int xb;       // span end (exclusive)
  int x1;       // box left
  int D;        // 2 * box width
  float invD;   // 1 / D
  int stepQ;    // (64*mw) / D
  int stepR;    // (64*mw) % D
  float wy;     // vertical weight of the lower source row
  int otop;     // float offset of the upper source row inside the staging slot, -1 = outside
  int obot;     // same for the lower source row
  int pad0_, pad1_;
};

struct __align__(16) WsItem {
  int valid, buf, first, last;
  int E, U, UL, len;
  int len16, N, pad0_, pad1_;
};

struct __align__(16) StoreRec {
  unsigned char *dst;
  int len16;
  int pad_;
};


struct __align__(16) ActBox {
  int n, xa, xb, obase;
};

__device__ __forceinline__ BoxAux make_aux(const int4 bx, int mw) {
  BoxAux a;
  a.D = 2 * (bx.w - bx.y);
  if (a.D <= 0) a.D = 2;   // never used: such boxes fail the validity test
  a.invD = __fdiv_rn(1.0f, static_cast<float>(a.D));
  a.stepQ = (64 * mw) / a.D;
  a.stepR = (64 * mw) - a.stepQ * a.D;
  return a;
}

// ---- debug watchdog (MRX_DEBUG=1): a wait that gives up after ~2^22 polls and records who
// was waiting for what, so that a protocol bug shows up as a report instead of a hang
__device__ int g_ws_debug[64];

__device__ __forceinline__ void mbar_wait_wd(uint32_t bar, uint32_t parity, int code, int a0,
                                             int a1, bool enabled) {
#ifndef MRX_WATCHDOG
  (void)code; (void)a0; (void)a1; (void)enabled;
  mbar_wait_a(bar, parity);
#else
  if (!enabled) {
    mbar_wait_a(bar, parity);
    return;
  }
#pragma unroll 1
  for (int it = 0; it < (1 << 16); ++it) {
    if (mbar_try_wait_a(bar, parity, 1000u)) return;
  }
  if ((threadIdx.x & 31) == 0) {
    const int slot = atomicAdd(&g_ws_debug[0], 1);
    if (slot < 12) {
      int *d = &g_ws_debug[4 + slot * 5];
      d[0] = code;
      d[1] = static_cast<int>(blockIdx.x);
      d[2] = static_cast<int>(threadIdx.x >> 5);
      d[3] = a0;
      d[4] = a1;
    }
  }
#endif
}

__host__ __device__ constexpr size_t ws_stage_bytes(int mw) {
  return static_cast<size_t>(kWsEMax) * 2 * mw * sizeof(float) + kWsEMax * sizeof(WsEntry) +
         kUMax * sizeof(uint32_t) + sizeof(WsItem);
}

// number of work units of an image (see the mode description above)
__device__ __forceinline__ int ws_units_of(int H, int W, int N, int chunk) {
  if (N <= 0) return 0;
  const long long RW = static_cast<long long>(W) * N;
  if ((RW % 16) == 0 && N <= kBoxCache) {
    const int S = static_cast<int>((RW + chunk - 1) / chunk);
    return S * ((H + kBandRows - 1) / kBandRows);
  }
  const long long jobs = (RW * H + chunk - 1) / chunk;
  return static_cast<int>((jobs + kFlatGroup - 1) / kFlatGroup);
}

template <int kNB>
__global__ void __launch_bounds__(kWsThreads, 1)
mask_expand_ws_kernel(const ExpandParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int mh = p.mh, mw = p.mw;
  const int slot_floats = 2 * mw;
  const uint32_t slot_bytes = slot_floats * 4;
  const int chunk = p.chunk_bytes;

  // ---- carve shared memory
  unsigned char *s_out = smem;                                   // kNB * chunk
  unsigned char *stage_base = smem + static_cast<size_t>(kNB) * chunk;
  const size_t stage_bytes = ws_stage_bytes(mw);
  auto stage_rows = [&](int s) { return reinterpret_cast<float *>(stage_base + s * stage_bytes); };
  auto stage_entries = [&](int s) {
    return reinterpret_cast<WsEntry *>(stage_base + s * stage_bytes +
                                       static_cast<size_t>(kWsEMax) * slot_bytes);
  };
  auto stage_units = [&](int s) {
    return reinterpret_cast<uint32_t *>(stage_base + s * stage_bytes +
                                        static_cast<size_t>(kWsEMax) * slot_bytes +
                                        kWsEMax * sizeof(WsEntry));
  };
  auto stage_item = [&](int s) {
    return reinterpret_cast<WsItem *>(stage_base + s * stage_bytes +
                                      static_cast<size_t>(kWsEMax) * slot_bytes +
                                      kWsEMax * sizeof(WsEntry) + kUMax * sizeof(uint32_t));
  };
  unsigned char *after = stage_base + kNS * stage_bytes;
  constexpr size_t kTableBytes = kBoxCache * (sizeof(int4) + sizeof(BoxAux) + sizeof(ActBox));
  const int prod = (warp >= kFirstProducerWarp) ? warp - kFirstProducerWarp : 0;
  int4 *s_box = reinterpret_cast<int4 *>(after + prod * kTableBytes);   // producer-private
  BoxAux *s_aux = reinterpret_cast<BoxAux *>(s_box + kBoxCache);
  ActBox *s_act = reinterpret_cast<ActBox *>(s_aux + kBoxCache);
  int *s_uprefix = reinterpret_cast<int *>(after + kProducers * kTableBytes);   // B + 1 prefix

  __shared__ uint64_t s_full[kNS], s_empty[kNS], s_done[kNB], s_free[kNB];
  __shared__ StoreRec s_store[kNB];
  __shared__ int s_total;
  __shared__ volatile int s_stop_job;
  __shared__ int s_item_ticket, s_job_ticket, s_fin_count;
  // admission control for the parity waits: with several producers, the holders of tickets t
  // and t + ring_size could otherwise wait on the same barrier for different phases, and a
  // parity wait is only unambiguous for the oldest of them
  __shared__ volatile int s_stage_gen[kNS];    // items that passed the empty[] wait, per stage
  __shared__ volatile int s_buf_gen[kMaxNB];   // jobs that passed the free[] wait, per chunk buffer
  __shared__ int s_pending[kMaxNB];   // per chunk buffer: outstanding consumer-warp arrivals + producer token

  // ---- work-unit table (first producer warp) and barrier init (store warp)
  if (warp == kFirstProducerWarp) {
    int carry = 0;
    for (int base = 0; base < p.B; base += 32) {
      const int b = base + lane;
      int v = 0;
      if (b < p.B)
        v = ws_units_of(p.geom[b * MRX_GEOM_INTS + 0], p.geom[b * MRX_GEOM_INTS + 1], p.counts[b],
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
    s_stop_job = -1;
    s_item_ticket = 0;
    s_job_ticket = 0;
    s_fin_count = 0;
    fence_mbar_init();
  }
  {
    // canvas zero fill, part 1: every chunk buffer starts all-zero (the store warp re-zeroes
    // a buffer after each bulk store)
    uint4 *o4 = reinterpret_cast<uint4 *>(s_out);
    const int n16 = (kNB * chunk) >> 4;
    for (int i = tid; i < n16; i += kWsThreads) o4[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();
  const int total_units = s_total;
  const uint32_t a_full = smem_u32(&s_full[0]), a_empty = smem_u32(&s_empty[0]);
  const uint32_t a_done = smem_u32(&s_done[0]), a_free = smem_u32(&s_free[0]);
  const bool wd = (p.flags & 0x8000) != 0;   // debug watchdog

  if (warp >= kFirstProducerWarp) {
    // ================================================================= producers
    // Items and jobs are numbered by CTA-wide tickets so that several producer warps can
    // build different items concurrently; consumers take items, the store warp takes jobs,
    // in ticket order.
    int st_s = 0, st_E = 0, st_U = 0;
    float *st_rows = nullptr;
    WsEntry *st_entries = nullptr;
    uint32_t *st_units = nullptr;

    auto open_item = [&]() {
      int item_idx = 0;
      if (lane == 0) item_idx = atomicAdd(&s_item_ticket, 1);
      item_idx = __shfl_sync(0xffffffffu, item_idx, 0);
      st_s = item_idx % kNS;
      const int gen = item_idx / kNS;
      while (s_stage_gen[st_s] != gen) __nanosleep(32);      // previous user of the stage is past its wait
      mbar_wait_wd(a_empty + 8 * st_s, (gen & 1) ^ 1, 1, item_idx, st_s, wd);   // consumers are done with the stage
      __syncwarp();
      if (lane == 0) s_stage_gen[st_s] = gen + 1;
      st_rows = stage_rows(st_s);
      st_entries = stage_entries(st_s);
      st_units = stage_units(st_s);
      st_E = 0;
      st_U = 0;
    };
    auto publish_item = [&](int buf, bool first, bool last, int UL, int len, int len16, int N) {
      if (lane == 0) {
        WsItem it;
        it.valid = 1;
        it.buf = buf;
        it.first = first ? 1 : 0;
        it.last = last ? 1 : 0;
        it.E = st_E;
        it.U = st_U;
        it.UL = UL;
        it.len = len;
        it.len16 = len16;
        it.N = N;
        it.pad0_ = 0;
        it.pad1_ = 0;
        *stage_item(st_s) = it;
        // kGroupWarps consumer warps will report on this item; the producer's own token
        // (taken in begin_job) is returned after the job's last item
        const int add = kGroupWarps - (last ? 1 : 0);
        const int old = atomicAdd(&s_pending[buf], add);
        (void)old;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive_a(a_full + 8 * st_s);
    };
    auto begin_job = [&](unsigned char *dst, int len16) -> int {
      int q = 0;
      if (lane == 0) q = atomicAdd(&s_job_ticket, 1);
      q = __shfl_sync(0xffffffffu, q, 0);
      const int buf = q % kNB;
      const int bgen = q / kNB;
      while (s_buf_gen[buf] != bgen) __nanosleep(32);        // previous job on the buffer is past its wait
      mbar_wait_wd(a_free + 8 * buf, (bgen & 1) ^ 1, 2, q, buf, wd);      // chunk buffer drained
      __syncwarp();
      if (lane == 0) s_buf_gen[buf] = bgen + 1;
      if (lane == 0) {
        s_store[buf].dst = dst;
        s_store[buf].len16 = len16;
        s_pending[buf] = 1;   // producer token: the job cannot complete before its last item is out
      }
      return buf;
    };
    // One round = up to 32 candidate (box,row) pairs, one per lane.  Returns false (warp
    // uniformly) when the open item cannot take the round; the caller publishes and retries.
    auto emit_round = [&](bool valid, int n, int row, int xa, int xb, int obase, const int4 bx,
                          const BoxAux ax, int ul_shift, const float *tiles_b) -> bool {
      const unsigned bal = __ballot_sync(0xffffffffu, valid);
      const int ecount = __popc(bal);
      const int nu = valid ? ((xb - xa + (32 << ul_shift) - 1) >> (5 + ul_shift)) : 0;
      int incl = nu;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
      }
      const int utotal = __shfl_sync(0xffffffffu, incl, 31);
      if (st_E + ecount > kWsEMax || st_U + utotal > kUMax) return false;
      if (ecount > 0) {
        if (lane == 0) mbar_expect_tx_a(a_full + 8 * st_s, ecount * slot_bytes);
        __syncwarp();
        if (valid) {
          const int slot = st_E + __popc(bal & ((1u << lane) - 1u));
          const int bh = bx.z - bx.x;
          const int Dy = 2 * bh;
          const int Ay = mh * (2 * (row - bx.x) + 1) - bh;
          // j0 = floor(Ay / Dy) in [-1, mh-1]: float estimate + exact integer correction
          int j0 = __float2int_rd(static_cast<float>(Ay) * __frcp_rn(static_cast<float>(Dy)));
          int remy = Ay - j0 * Dy;
          if (remy < 0) {
            --j0;
            remy += Dy;
          } else if (remy >= Dy) {
            ++j0;
            remy -= Dy;
          }
          const int jc = min(max(j0, 0), mh - 2);
          WsEntry e;
          e.obase = obase;
          e.xb = xb;
          e.x1 = bx.y;
          e.D = ax.D;
          e.invD = ax.invD;
          e.stepQ = ax.stepQ;
          e.stepR = ax.stepR;
          e.wy = __fdiv_rn(static_cast<float>(remy), static_cast<float>(Dy));
          e.otop = (j0 < 0) ? -1 : (j0 - jc) * mw;
          e.obot = (j0 + 1 > mh - 1) ? -1 : (j0 + 1 - jc) * mw;
          e.pad0_ = 0;
          e.pad1_ = 0;
          st_entries[slot] = e;
          bulk_g2s_a(smem_u32(st_rows) + slot * slot_bytes,
                     tiles_b + (n * mh + jc) * mw, slot_bytes, a_full + 8 * st_s);
          uint32_t *up = st_units + st_U + (incl - nu);
          for (int j = 0; j < nu; ++j)
            up[j] = (static_cast<uint32_t>(slot) << 24) |
                    static_cast<uint32_t>(xa + (j << (5 + ul_shift)));
        }
      }
      st_E += ecount;
      st_U += utotal;
      return true;
    };
    // smallest shift such that a span of `span_max` columns has <= kMaxUnitsPerEntry units
    auto unit_shift = [&](int span_max) -> int {
      int sh = ((p.flags >> 4) & 7) ? ((p.flags >> 4) & 7) : kMinUnitShift;
      while ((32 << sh) * kMaxUnitsPerEntry < span_max) ++sh;
      return sh;
    };

    int cur_b = 0, cached_b = -1;
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
      const float *tiles_b = p.tiles + static_cast<size_t>(b) * p.R * mh * mw;
      const int4 *boxes_b = p.boxes + static_cast<size_t>(b) * p.R;
      unsigned char *canvas_b = p.canvas + p.canvas_off[b];
      const bool cached = N <= kBoxCache;
      if (cached && b != cached_b) {
        for (int n = lane; n < N; n += 32) {
          const int4 bx = __ldg(boxes_b + n);
          s_box[n] = bx;
          s_aux[n] = make_aux(bx, mw);
        }
        cached_b = b;
        __syncwarp();
      }

      if ((RW % 16u) == 0u && cached) {
        // ------------------------------------------------------------- strip mode
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
          const int sub = seg_off - xlo * N;          // bytes of pixel xlo before the segment
          const int ul_shift = unit_shift(xhi - xlo + 1);
          // active list: boxes whose clipped span is non-empty and whose rows meet the band
          int n_act = 0;
          for (int n0 = 0; n0 < N; n0 += 32) {
            const int n = n0 + lane;
            bool act = false;
            int xa = 0, xb = 0;
            if (n < N) {
              const int4 bx = s_box[n];
              xa = max(xlo, bx.y);
              xb = min(xhi + 1, bx.w);
              const bool sane = bx.x >= 0 && bx.y >= 0 && bx.z <= H && bx.w <= W;
              act = sane && xa < xb && bx.x < yb && bx.z > ya && bx.z > bx.x;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, act);
            if (act) {
              ActBox a;
              a.n = n;
              a.xa = xa;
              a.xb = xb;
              a.obase = n - sub - xlo * N;            // (row*W - g0)*N + n - sub with g0 = row*W + xlo
              s_act[n_act + __popc(bal & ((1u << lane) - 1u))] = a;
            }
            n_act += __popc(bal);
          }
          __syncwarp();
          for (int y = ya; y < yb; ++y) {
            const unsigned c0 = static_cast<unsigned>(y) * RW + seg_off;
            const int buf = begin_job(canvas_b + c0, len);
            open_item();
            bool first = true;
            for (int a0 = 0; a0 < n_act; a0 += 32) {
              const int ai = a0 + lane;
              bool valid = false;
              ActBox a;
              a.n = 0; a.xa = 0; a.xb = 0; a.obase = 0;
              int4 bx = make_int4(0, 0, 1, 1);
              BoxAux ax;
              ax.D = 2; ax.invD = 0.5f; ax.stepQ = 0; ax.stepR = 0;
              if (ai < n_act) {
                a = s_act[ai];
                bx = s_box[a.n];
                ax = s_aux[a.n];
                valid = y >= bx.x && y < bx.z;
              }
              while (!emit_round(valid, a.n, y, a.xa, a.xb, a.obase, bx, ax, ul_shift, tiles_b)) {
                publish_item(buf, first, false, 32 << ul_shift, len, len, N);
                first = false;
                open_item();
              }
            }
            publish_item(buf, first, true, 32 << ul_shift, len, len, N);
          }
        }
      } else {
        // ------------------------------------------------------------- flat mode
        const int jobs_b = static_cast<int>((static_cast<unsigned long long>(L) + chunk - 1) / chunk);
        const int j_end = min(jobs_b, (u_local + 1) * kFlatGroup);
        for (int j = u_local * kFlatGroup; j < j_end; ++j) {
          const unsigned c0 = static_cast<unsigned>(j) * chunk;
          const int len = static_cast<int>(min(static_cast<unsigned>(chunk), L - c0));
          const int len16 = (len + 15) & ~15;
          const int g0 = static_cast<int>(c0 / N);               // first pixel touched
          const int g1 = static_cast<int>((c0 + len - 1) / N);   // last pixel touched
          const int r0 = g0 / W;
          const int r1 = g1 / W;
          const int sub = static_cast<int>(c0 - static_cast<unsigned>(g0) * N);
          const int ul_shift = unit_shift(min(W, g1 - g0 + 1));
          const int buf = begin_job(canvas_b + c0, len16);
          open_item();
          bool first = true;
          for (int row = r0; row <= r1; ++row) {
            const int xlo = max(0, g0 - row * W);
            const int xhi = min(W, g1 + 1 - row * W);
            for (int n0 = 0; n0 < N; n0 += 32) {
              const int n = n0 + lane;
              bool valid = false;
              int4 bx = make_int4(0, 0, 1, 1);
              BoxAux ax;
              ax.D = 2; ax.invD = 0.5f; ax.stepQ = 0; ax.stepR = 0;
              int xa = 0, xb = 0;
              if (n < N) {
                if (cached) {
                  bx = s_box[n];
                  ax = s_aux[n];
                } else {
                  bx = __ldg(boxes_b + n);
                  ax = make_aux(bx, mw);
                }
                xa = max(xlo, bx.y);
                xb = min(xhi, bx.w);
                const bool sane = bx.x >= 0 && bx.y >= 0 && bx.z <= H && bx.w <= W;
                valid = sane && row >= bx.x && row < bx.z && xa < xb;
              }
              const int obase = (row * W - g0) * N + n - sub;
              while (!emit_round(valid, n, row, xa, xb, obase, bx, ax, ul_shift, tiles_b)) {
                publish_item(buf, first, false, 32 << ul_shift, len, len16, N);
                first = false;
                open_item();
              }
            }
          }
          publish_item(buf, first, true, 32 << ul_shift, len, len16, N);
        }
      }
      unit = __shfl_sync(0xffffffffu, next_unit, 0);
    }
    // the last producer to run dry publishes the sentinel item: it tells the consumers
    // (and through them the store warp) to stop
    int fin = 0;
    if (lane == 0) fin = atomicAdd(&s_fin_count, 1);
    fin = __shfl_sync(0xffffffffu, fin, 0);
    if (fin == kProducers - 1) {
      int q = 0;
      if (lane == 0) q = atomicAdd(&s_job_ticket, 0);   // jobs started by all producers
      q = __shfl_sync(0xffffffffu, q, 0);
      for (int g = 0; g < kGroups; ++g) {   // consecutive tickets reach every group once
        open_item();
        if (lane == 0) {
          WsItem it;
          it.valid = 0;
          it.buf = 0;
          it.first = it.last = 0;
          it.E = it.U = it.UL = it.len = it.len16 = it.N = 0;
          it.pad0_ = it.pad1_ = 0;
          *stage_item(st_s) = it;
          s_stop_job = q;
          mbar_arrive_a(a_full + 8 * st_s);
        }
        __syncwarp();
      }
    }
  } else if (warp >= kFirstStoreWarp) {
    // ================================================================= store warps
    // store warp j owns the chunk buffers b with b % kStoreWarps == j (every done[] barrier has
    // a single waiter that sees each of its phases): store, drain and re-zero run in parallel
    // on different buffers
    const int me = warp - kFirstStoreWarp;
    const int n_store = ((p.flags >> 12) & 3) ? ((p.flags >> 12) & 3) : kStoreWarps;
    for (int k = 0; me < n_store; ++k) {
      const int b = k % kNB;
      if ((b % n_store) != me) continue;
      // wait for job k, or learn that it does not exist (the last producer publishes the
      // number of jobs in s_stop_job; real jobs always complete their done[] phase)
      bool have = false;
#pragma unroll 1
      for (int it = 0;; ++it) {
        have = mbar_try_wait_a(a_done + 8 * b, (k / kNB) & 1, 1000u);
        if (have) break;
        const int stop = s_stop_job;
        if (stop >= 0 && k >= stop) break;
#ifdef MRX_WATCHDOG
        if (wd && it > (1 << 16)) {
          if (lane == 0) {
            const int slot = atomicAdd(&g_ws_debug[0], 1);
            if (slot < 12) {
              int *d = &g_ws_debug[4 + slot * 5];
              d[0] = 4; d[1] = static_cast<int>(blockIdx.x); d[2] = warp; d[3] = k; d[4] = s_pending[b] * 1000 + stop;
            }
          }
          break;
        }
#endif
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
    const int grp = warp / kGroupWarps;      // consumers are warps 0 .. kWsConsumerWarps-1
    const int gw = warp - grp * kGroupWarps;
    int s = grp - kGroups;                   // this group's items: tickets grp, grp + kGroups, ...
    uint32_t full_parity = 0;
    while (true) {
      s += kGroups;
      if (s >= kNS) {
        s -= kNS;
        full_parity ^= 1;
      }
      mbar_wait_wd(a_full + 8 * s, full_parity, 3, s, static_cast<int>(full_parity), wd);
      const WsItem it = *stage_item(s);
      if (!it.valid) break;
      const uint32_t out_addr = smem_u32(s_out + static_cast<size_t>(it.buf) * chunk);
      const float *rows = stage_rows(s);
      const WsEntry *entries = stage_entries(s);
      const uint32_t *units = stage_units(s);
      const unsigned ulen = static_cast<unsigned>(it.len);
      const unsigned ostep = 32u * it.N;
      const int n_units = (p.flags & 0x100) ? 0 : it.U;
      for (int u = gw; u < n_units; u += kGroupWarps) {
        const uint32_t unit = units[u];
        const int ei = static_cast<int>(unit >> 24);
        const int x0 = static_cast<int>(unit & 0xffffffu);
        const WsEntry e = entries[ei];
        const float *slot = rows + ei * slot_floats;
        // vertical blend of the two staged tile rows, one column per lane:
        // lane l holds B[l] with B[0] = 0, B[1+i] = blend(i), B[mw+1] = 0   (mw <= 30)
        float bl = 0.f;
        {
          const int i = lane - 1;
          if (i >= 0 && i < mw) {
            const float top = (e.otop >= 0) ? slot[e.otop + i] : 0.f;
            const float bot = (e.obot >= 0) ? slot[e.obot + i] : 0.f;
            bl = fmaf(e.wy, bot - top, top);
          }
        }
        const int xend = min(x0 + it.UL, e.xb);
        int x = x0 + lane;
        // exact source coordinate of this lane's first column, then 32 columns per step
        int idx, rem;
        {
          const int A = mw * (2 * (x - e.x1) + 1) - (e.D >> 1);
          int i0 = __float2int_rd(static_cast<float>(A) * e.invD);
          rem = A - i0 * e.D;
          if (rem < 0) {
            --i0;
            rem += e.D;
          } else if (rem >= e.D) {
            ++i0;
            rem -= e.D;
          }
          idx = i0 + 1;   // B[idx], B[idx+1] are the two taps
        }
        unsigned off = static_cast<unsigned>(e.obase + x * it.N);
#pragma unroll 2
        for (int xs = x0; xs < xend; xs += 32) {   // warp-uniform trip count (shuffles inside)
          const float wx = static_cast<float>(rem) * e.invD;
          const float a = __shfl_sync(0xffffffffu, bl, idx);
          const float bq = __shfl_sync(0xffffffffu, bl, idx + 1);
          const float v = fmaf(wx, bq - a, a);
          if (v >= 0.5f && x < xend && off < ulen)
            asm volatile("st.shared.u8 [%0], %1;" ::"r"(out_addr + off), "r"(1u));
          x += 32;
          off += ostep;
          rem += e.stepR;
          idx += e.stepQ;
          if (rem >= e.D) {
            rem -= e.D;
            ++idx;
          }
        }
      }
      fence_proxy_async_smem();   // chunk bytes must be visible to the bulk store
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_a(a_empty + 8 * s);
        // report on the job; whoever returns the last outstanding arrival releases the chunk
        if (atomicSub(&s_pending[it.buf], 1) == 1) mbar_arrive_a(a_done + 8 * it.buf);
      }
    }
  }
}


}  // namespace ws4

int launch_expand_ws4(const ExpandParams &prm_in, int sms, int max_optin, cudaStream_t st) {
  using namespace ws4;
  ExpandParams prm = prm_in;
  const int mw = prm.mw;
  const int chunk_bytes = prm.chunk_bytes;
  const int B = prm.B;
  {
    const size_t fixed = kNS * ws_stage_bytes(mw) +
                         static_cast<size_t>(kProducers) * kBoxCache * (sizeof(int4) + sizeof(BoxAux) + sizeof(ActBox)) +
                         static_cast<size_t>(B + 1) * sizeof(int) + 1024;
    int nb = static_cast<int>((static_cast<size_t>(max_optin) - fixed) / chunk_bytes);
    if (nb > kMaxNB) nb = kMaxNB;
    {
      const char *e = getenv("MRX_EXPAND_NB");
      if (e && atoi(e) >= 2 && atoi(e) < nb) nb = atoi(e);
    }
    MRX_CHECK_SUPPORTED(nb >= 2 && fixed < static_cast<size_t>(max_optin),
                        "mrx_mask_expand: chunk_bytes %d too large for %d B of shared memory",
                        chunk_bytes, max_optin);
    const size_t smem = static_cast<size_t>(nb) * chunk_bytes + fixed - 1024;
    void (*kern)(const ExpandParams) = nullptr;
    switch (nb) {
      case 2: kern = mask_expand_ws_kernel<2>; break;
      case 3: kern = mask_expand_ws_kernel<3>; break;
      case 4: kern = mask_expand_ws_kernel<4>; break;
      case 5: kern = mask_expand_ws_kernel<5>; break;
      default: kern = mask_expand_ws_kernel<6>; break;
    }
    MRX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem)));
#ifdef MRX_WATCHDOG
    const bool dbg = getenv("MRX_DEBUG") != nullptr;
#else
    const bool dbg = false;   // rebuild with -DMRX_WATCHDOG to get the stuck-wait report
#endif
    if (dbg) {
      int zero[64] = {0};
      MRX_CUDA(cudaMemcpyToSymbol(g_ws_debug, zero, sizeof(zero)));
      prm.flags |= 0x8000;
    }
    kern<<<sms, kWsThreads, smem, st>>>(prm);
    if (dbg) {
      MRX_CUDA(cudaStreamSynchronize(st));
      int h[64];
      MRX_CUDA(cudaMemcpyFromSymbol(h, g_ws_debug, sizeof(h)));
      if (h[0] > 0) {
        fprintf(stderr, "[mrx ws watchdog] %d stuck waits (nb=%d):\n", h[0], nb);
        for (int i = 0; i < h[0] && i < 12; ++i)
          fprintf(stderr, "  code=%d (1=empty 2=free 3=full 4=done) cta=%d warp=%d a0=%d a1=%d\n",
                  h[4 + i * 5], h[5 + i * 5], h[6 + i * 5], h[7 + i * 5], h[8 + i * 5]);
      }
    }
    MRX_LAUNCH_CHECK("mask_expand_ws4_kernel");
  }
  return MRX_OK;
}

}  // namespace mrx
