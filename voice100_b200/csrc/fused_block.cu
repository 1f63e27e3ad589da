// Block-level fusion of InvertedResidual (voice100/models/asr.py:45-59): pointwise EXPAND (1x1 conv + BN + ReLU6)
// and DEPTHWISE (k-tap conv + BN + ReLU6) in one kernel.  The 4x-wide expand output never goes to HBM: it is drained
// from tensor memory as bf16 into a shared-memory sliding window and the depthwise FIR reads it from there.
// (Per 512->2048 block at 256 x 751 that removes a 788 MB write and a 788 MB read, ~42 % of the block's traffic; the
// project GEMM that follows is the unchanged conv_gemm kernel.)
//
// Why expand->depthwise and not depthwise->project or all three: the depthwise stage needs (k-1) columns of halo.
// Fusing it BEHIND the expand GEMM lets a CTA pair walk one utterance's time tiles in order and keep the last 88
// columns in shared memory -- no halo is ever recomputed or re-read -- while the accumulator tile stays the 256-wide
// one the tensor cores want.  Feeding the project GEMM from the depthwise stage would need the hidden tensor of a
// 256-column tile for ALL 2048 channels on chip, or a 512-column TMEM accumulator per output slice on top of the
// expand accumulator: tensor memory (512 columns) has room for one of the two, and narrower tiles are L2-bound.
//
// Structure = conv_gemm.cu's CTA-pair kernel (TMA producer warp, one MMA-issuing warp, TMEM accumulators double
// buffered, cta_group::2, 256 hidden channels x 256 time steps per tile, 128 channels per CTA) with a new epilogue:
//   work unit  = (utterance b, 256-channel slice); its time tiles are processed in order;
//   drain      : TMEM -> scale*acc+shift -> ReLU6 -> bf16 (zero for t >= T) -> window[channel][88 halo + 256 new];
//   depthwise  : per channel, the Toeplitz-block FIR of dwconv.cu on mma.sync (data = the small B operand, read as
//                conflict-free 64-bit words; filter blocks = A, loaded per channel from a packed pair table in
//                global memory, L1/L2 resident), 2 x 128 outputs per tile, BN + ReLU6, 32-byte-sector stores;
//   halo       : the last 88 columns of the row move to the front for the next tile; after the last tile one more
//                128-output pass over zero data flushes the outputs whose taps reach past the end of the clip.
// Output positions lag the accumulator tile by p = (k-1)/2 columns, so consecutive tiles write disjoint ranges.
#include "common.cuh"
#include "host.h"

#include <cstdlib>

namespace v100 {

namespace {

constexpr int kM = 128;                    // hidden channels per CTA (256 per pair)
constexpr int kN = 256;                    // time steps per tile
constexpr int kK = 64;                     // input channels per pipeline stage
constexpr int kABytes = kM * kK * 2;       // 16 KB of weights per stage and CTA
constexpr int kBAtom = kK * 64 * 2;        // 8 KB: 64 k-rows x 64 time steps
constexpr int kBBytes = 2 * kBAtom;        // this CTA's 128 of the tile's 256 time columns
constexpr int kStage = kABytes + kBBytes;  // 32 KB
constexpr int kStages = 3;
constexpr int kHalo = 88;                  // window columns kept from the previous tile (>= 2p + 2 for k <= 83)
constexpr int kWinCols = kHalo + kN;       // 344
constexpr int kWinPitch = kWinCols * 2;    // 688 B = 43 x 16: rows 16 B apart mod 128 -> conflict-free 16-byte stores
constexpr int kWinBytes = kM * kWinPitch + 128;   // + zeroed slack: the zero-weight tail of the last row's window
constexpr int kEpiWarps = 16;               // four per TMEM lane quadrant: 64 columns to drain, 8 channels to filter each
constexpr int kThreads = (4 + kEpiWarps) * 32;
constexpr int kPairTab = 128;              // packed filter pairs per channel
constexpr int kTabBufs = 3;                // per-warp ring of staged pair tables (prefetch distance 2)
constexpr int kTabBytes = kEpiWarps * kTabBufs * kPairTab * 4;
constexpr int kSmem = 1024 + kStages * kStage + kWinBytes + kTabBytes + 256;
static_assert(kSmem <= 232448, "exceeds 227 KB of shared memory");

struct FusedParams {
  int C_in, H, B, T;
  int m_tiles, t_tiles, num_units, k_blocks;
  int p, e1;
  const float *scale1, *shift1, *scale2, *shift2;
  const uint32_t* pairs;          // [H][kPairTab]: (wz[i], wz[i+1]) with wz[16 + e1 + j] = w[j]
  unsigned short* y;
  long long y_pitch;
  int dtype;
};

template <int DT>
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if constexpr (DT == DT_F16) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  } else {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
}

// explicit shared-window accesses: the dynamic shared memory base is re-aligned through integer arithmetic, after which
// the compiler no longer knows the address space and would emit generic loads (measured: 172 M generic LD per launch,
// every fragment and data word of the FIR) -- these keep them LDS / STS
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t a) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <bool B>
struct BoolC { static constexpr bool value = B; };

}  // namespace

template <int Q, int DT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
expand_dw_kernel(const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_x, const FusedParams p) {
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int unit0 = blockIdx.x / 2, unit_step = gridDim.x / 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* win = smem + kStages * kStage;                    // [128 channels][344 columns] 16-bit
  uint32_t* tabs = reinterpret_cast<uint32_t*>(win + kWinBytes);   // [16 warps][3][128] staged pair tables
  uint64_t* bars = reinterpret_cast<uint64_t*>(win + kWinBytes + kTabBytes);
  uint64_t* full_bar = bars;                 // [kStages]
  uint64_t* empty_bar = bars + kStages;      // [kStages]
  uint64_t* tmem_full = bars + 2 * kStages;  // [2]
  uint64_t* tmem_empty = tmem_full + 2;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;

  pdl_trigger();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_w);
    tma_prefetch_desc(&tm_x);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], kEpiWarps * 2);   // the leader's copy collects both CTAs' epilogue warps
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_cg2(tmem_slot, 2 * kN);
    tmem_relinquish_cg2();
  }
  // the window starts as zeros: its bits are later only ever ReLU6 outputs, i.e. finite -- the depthwise MMAs read a few
  // columns past what their non-zero taps need (into the next row / the slack) and 0 x finite must stay 0
  for (int i = threadIdx.x; i < kWinBytes / 16; i += kThreads) reinterpret_cast<uint4*>(win)[i] = make_uint4(0u, 0u, 0u, 0u);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer =====================
    const bool issuer = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    for (int unit = unit0; unit < p.num_units; unit += unit_step) {
      const int m_tile = unit % p.m_tiles, b = unit / p.m_tiles;
      const int m0 = (m_tile * 2 + int(cta_rank)) * kM;
      for (int tt = 0; tt < p.t_tiles; ++tt) {
        const int t_in0 = tt * kN + int(cta_rank) * (kN / 2);
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kStage;
          uint8_t* sb = sa + kABytes;
          if (issuer) {
            if (leader) mbar_expect_tx(&full_bar[stage], 2 * kStage);
            const uint32_t fb = mapa_u32(smem_u32(&full_bar[stage]), 0);
            tma_load_2d_cg2(sa, &tm_w, fb, kb * kK, m0);
            tma_load_3d_cg2(sb, &tm_x, fb, t_in0, kb * kK, b);
            tma_load_3d_cg2(sb + kBAtom, &tm_x, fb, t_in0 + 64, kb * kK, b);
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ===================== MMA issuer =====================
      const bool issuer = elect_one();
      const uint32_t fmt = p.dtype == DT_F16 ? 0u : 1u;
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (0u << 15) | (1u << 16) |
                             (uint32_t(kN >> 3) << 17) | (uint32_t((2 * kM) >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int iter = 0;
      for (int unit = unit0; unit < p.num_units; unit += unit_step) {
        for (int tt = 0; tt < p.t_tiles; ++tt, ++iter) {
          const int accbuf = iter & 1;
          mbar_wait(&tmem_empty[accbuf], ((iter >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + accbuf * kN;
          for (int kb = 0; kb < p.k_blocks; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(smem + stage * kStage);
            const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
            for (int k = 0; k < kK / 16; ++k) {
              const uint64_t da = umma_desc(a_addr + k * 32, 16, 1024);
              const uint64_t db = umma_desc(b_addr + k * 2048, kBAtom, 1024);
              if (issuer) umma_bf16_cg2(d_tmem, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
            if (issuer) umma_commit_cg2(&empty_bar[stage]);
            __syncwarp();
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
          if (issuer) umma_commit_cg2(&tmem_full[accbuf]);
          __syncwarp();
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: drain -> depthwise -> store =====================
    const int qd = warp & 3;              // TMEM lane quadrant = 32 channels
    const int g = (warp - 4) >> 2;        // which 64 columns this warp drains / which 8 channels it filters
    const int row = qd * 32 + lane;       // channel row this thread drains
    const uint32_t lane_addr = tmem_base + (uint32_t(qd * 32) << 16);
    const uint32_t tmem_empty_leader = mapa_u32(smem_u32(&tmem_empty[0]), 0);
    const int gq = lane >> 2, tg = lane & 3;
    const bool even = (gq & 1) == 0;
    const int pos0 = 32 * tg + (even ? gq : gq + 7);      // after the pair exchange: outputs (pos0, pos0+1), (pos0+16, +17)
    const int cb = kHalo - 2 * p.p - 2 * p.e1;            // window column where a tile's data window starts (multiple of 4)
    constexpr int kRows = kM / kEpiWarps;                 // 8 channels per warp
    const int r0 = qd * 32 + g * kRows;                   // this warp's first channel row
    const uint32_t win_s = smem_u32(win);
    const uint32_t my_row_s = win_s + row * kWinPitch + kHalo * 2;             // where this thread drains to
    const uint32_t rows_s = win_s + r0 * kWinPitch;                            // first of the 8 rows this warp filters
    const uint32_t tabs_s = smem_u32(tabs) + (warp - 4) * kTabBufs * kPairTab * 4;
    const uint32_t frag_off = (16 + 4 * tg - gq) * 4;     // byte offset of this lane's first fragment word in a table
    const uint32_t data_off = cb * 2 + lane * 8;          // byte offset of this lane's first data word in a row

    // Pair tables are staged two channels ahead with cp.async (16 bytes per lane = one 512-byte table per warp-wide
    // copy) so that their L2 latency is never on the critical path.  Table n of this warp lives in ring slot n % 3;
    // one commit group per request, empty when there is nothing left to request.
    auto request_table = [&](int ch, int slot) {
      if (ch >= 0)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tabs_s + slot * (kPairTab * 4) + lane * 16),
                     "l"(p.pairs + static_cast<long long>(ch) * kPairTab + lane * 4) : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // Filter this warp's 8 channels: N_MMA x 128 outputs per channel starting at time tau0.  `next_first` = the first
    // channel this warp filters in the pass that follows (the unit's next tile / flush, or the next unit), -1 if none.
    // INTERIOR: every output of the pass lies in [0, T) -- no store predicates.
    int slot = 0;                                         // ring slot of the table the next channel uses
    auto filter_rows = [&](auto two_tiles, auto interior, int first_ch, int b, int tau0, bool move_halo, int next_first) {
      constexpr bool kTwo = decltype(two_tiles)::value, kInterior = decltype(interior)::value;
      uint32_t rp = rows_s;
      unsigned short* yrow = p.y + (static_cast<long long>(b) * p.H + first_ch) * p.y_pitch + tau0 + pos0;
      const int t_lane = tau0 + pos0;
#pragma unroll 1
      for (int i = 0; i < kRows; ++i) {
        asm volatile("cp.async.wait_group 1;" ::: "memory");   // this channel's table has landed (the next may be in flight)
        __syncwarp();                                           // ... for every lane; and the slot two ahead is no longer read
        {
          const int i2 = i + 2;
          const int nch = i2 < kRows ? first_ch + i2 : (next_first >= 0 ? next_first + (i2 - kRows) : -1);
          request_table(nch, slot >= 1 ? slot - 1 : kTabBufs - 1);      // (slot + 2) % 3
        }
        const uint32_t pt = tabs_s + slot * (kPairTab * 4) + frag_off;
        slot = slot == kTabBufs - 1 ? 0 : slot + 1;
        uint32_t af[Q][4];
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          af[q][0] = lds32(pt + 64 * q);          // (m = g    , kk = 4tg, 4tg+1)
          af[q][1] = lds32(pt + 64 * q - 32);     // (m = g + 8, kk = 4tg, 4tg+1)
          af[q][2] = lds32(pt + 64 * q + 8);      // (m = g    , kk = 4tg+2, 4tg+3)
          af[q][3] = lds32(pt + 64 * q - 24);     // (m = g + 8, kk = 4tg+2, 4tg+3)
        }
        const int ch = first_ch + i;
        const float sc = p.scale2 ? __ldg(p.scale2 + ch) : 1.0f;
        const float sh = __ldg(p.shift2 + ch);
        const uint32_t xw = rp + data_off;
        auto finish = [&](float (&acc)[4], int dt) {       // dt = 0 / 128: which of the pass's tiles
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[j] = fmaf(acc[j], sc, sh);
          const float x0 = __shfl_xor_sync(0xffffffffu, even ? acc[2] : acc[0], 4);
          const float x1 = __shfl_xor_sync(0xffffffffu, even ? acc[3] : acc[1], 4);
          const uint32_t o0 = even ? pack2_relu6<DT>(acc[0], x0) : pack2_relu6<DT>(x0, acc[2]);
          const uint32_t o1 = even ? pack2_relu6<DT>(acc[1], x1) : pack2_relu6<DT>(x1, acc[3]);
          if constexpr (kInterior) {
            *reinterpret_cast<uint32_t*>(yrow + dt) = o0;
            *reinterpret_cast<uint32_t*>(yrow + dt + 16) = o1;
          } else {
            const int t = t_lane + dt;
            if (t >= 0 && t < p.T) *reinterpret_cast<uint32_t*>(yrow + dt) = o0;
            if (t + 16 >= 0 && t + 16 < p.T) *reinterpret_cast<uint32_t*>(yrow + dt + 16) = o1;
          }
        };
        if constexpr (kTwo) {
          float acc0[4] = {0.0f, 0.0f, 0.0f, 0.0f}, acc1[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
          for (int q = 0; q < Q; ++q) {
            const uint2 b0 = lds64(xw + 32 * q);
            const uint2 b1 = lds64(xw + 256 + 32 * q);
            mma_16816<DT>(acc0, af[q], b0.x, b0.y);
            mma_16816<DT>(acc1, af[q], b1.x, b1.y);
          }
          finish(acc0, 0);
          finish(acc1, 128);
        } else {
          float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
          for (int q = 0; q < Q; ++q) {
            const uint2 bq = lds64(xw + 32 * q);
            mma_16816<DT>(acc, af[q], bq.x, bq.y);
          }
          finish(acc, 0);
        }
        __syncwarp();                                   // every lane is past its reads of this row
        if (lane < kHalo * 2 / 16) {
          // columns [256, 344) become the next tile's [0, 88); at the end of a unit the next one gets a zero halo
          const uint4 h = move_halo ? lds128(rp + kN * 2 + lane * 16) : make_uint4(0u, 0u, 0u, 0u);
          sts128(rp + lane * 16, h);
        }
        rp += kWinPitch;
        yrow += p.y_pitch;
      }
      __syncwarp();
    };

    int iter = 0;
    const bool tail = p.T > p.t_tiles * kN - p.p - p.e1;       // outputs [256 nT - p - e1, T) need a flush pass
    if (unit0 < p.num_units) {                                 // tables of the first pass's first two channels
      const int first = ((unit0 % p.m_tiles) * 2 + int(cta_rank)) * kM + r0;
      request_table(first, 0);
      request_table(first + 1, 1);
    }
    for (int unit = unit0; unit < p.num_units; unit += unit_step) {
      const int m_tile = unit % p.m_tiles, b = unit / p.m_tiles;
      const int ch0 = (m_tile * 2 + int(cta_rank)) * kM;
      const int first = ch0 + r0;
      const int unit_next = unit + unit_step;
      const int first_next = unit_next < p.num_units ? ((unit_next % p.m_tiles) * 2 + int(cta_rank)) * kM + r0 : -1;
      const float sc1 = p.scale1 ? __ldg(p.scale1 + ch0 + row) : 1.0f;
      const float sh1 = __ldg(p.shift1 + ch0 + row);
      for (int tt = 0; tt < p.t_tiles; ++tt, ++iter) {
        const int accbuf = iter & 1;
        mbar_wait(&tmem_full[accbuf], (iter >> 1) & 1);
        tc_fence_after();
        // ---- drain this warp's 64 columns of its quadrant's 32 channels into the window ----
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          uint32_t v[32];
          tmem_ld32(lane_addr + accbuf * kN + g * 64 + cc * 32, v);
          tmem_ld_wait();
          if (cc == 1) {   // this warp is done reading the accumulator buffer
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tmem_empty_leader + accbuf * 8);
          }
          const int t0 = tt * kN + g * 64 + cc * 32;    // time of v[0]; columns at t >= T are the conv's zero padding
          const uint32_t dst = my_row_s + (g * 64 + cc * 32) * 2;
          uint32_t w[16];
#pragma unroll
          for (int e = 0; e < 16; ++e)
            w[e] = pack2_relu6<DT>(fmaf(__uint_as_float(v[2 * e]), sc1, sh1), fmaf(__uint_as_float(v[2 * e + 1]), sc1, sh1));
          if (t0 + 32 > p.T) {                          // (warp-uniform) the tile straddles the end of the clip
#pragma unroll
            for (int e = 0; e < 16; ++e)
              if (t0 + 2 * e + 1 >= p.T) w[e] = (t0 + 2 * e >= p.T) ? 0u : (w[e] & 0xFFFFu);
          }
#pragma unroll
          for (int k16 = 0; k16 < 4; ++k16) sts128(dst + 16 * k16, make_uint4(w[4 * k16], w[4 * k16 + 1], w[4 * k16 + 2], w[4 * k16 + 3]));
        }
        named_bar_sync(1 + qd, 128);     // all four column slices of these 32 channels are in the window
        const bool last_pass = tt == p.t_tiles - 1 && !tail;
        const int tau0 = tt * kN - p.p - p.e1;
        const int nxt = last_pass ? first_next : first;
        if (tau0 >= 0 && tau0 + kN <= p.T) filter_rows(BoolC<true>{}, BoolC<true>{}, first, b, tau0, !last_pass, nxt);
        else filter_rows(BoolC<true>{}, BoolC<false>{}, first, b, tau0, !last_pass, nxt);
        named_bar_sync(1 + qd, 128);     // all four warps are done reading: the next drain may overwrite columns [88, 344)
      }
      // ---- flush: the outputs whose taps reach past the last tile see zero data there ----
      if (tail) {
        uint32_t rp = rows_s + kHalo * 2;
        for (int i = 0; i < kRows; ++i, rp += kWinPitch)
          if (lane < 8) sts128(rp + lane * 16, make_uint4(0u, 0u, 0u, 0u));     // 64 zero columns >= p
        __syncwarp();
        filter_rows(BoolC<false>{}, BoolC<false>{}, first, b, p.t_tiles * kN - p.p - p.e1, false, first_next);
        named_bar_sync(1 + qd, 128);
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_cg2(tmem_base, 2 * kN);
  }
}

// pairs[c][i] = (wz[i], wz[i+1]) as one 32-bit word, wz[16 + e1 + j] = w[c][j] (zero elsewhere), i < 128
__global__ void __launch_bounds__(128)
dw_pack_pairs_kernel(const unsigned short* __restrict__ w, uint32_t* __restrict__ pairs, int C, int k, int e1) {
  const int c = blockIdx.x, i = threadIdx.x;
  if (c >= C) return;
  auto wz = [&](int idx) -> uint32_t {
    const int j = idx - 16 - e1;
    return (j >= 0 && j < k) ? uint32_t(w[static_cast<long long>(c) * k + j]) : 0u;
  };
  pairs[static_cast<long long>(c) * kPairTab + i] = wz(i) | (wz(i + 1) << 16);
}

int dw_pack_pairs(const void* w, uint32_t* pairs, int C, int k, cudaStream_t stream) {
  if (w == nullptr || pairs == nullptr) return fail(V100_E_INVALID, "dw_pack_pairs: null pointer");
  if (C <= 0 || k <= 0 || (k & 1) == 0 || k > 83) return fail(V100_E_UNSUPPORTED, "dw_pack_pairs: kernel size %d (odd, <= 83)", k);
  const int p = (k - 1) / 2;
  dw_pack_pairs_kernel<<<C, kPairTab, 0, stream>>>(static_cast<const unsigned short*>(w), pairs, C, k, p & 1);
  V100_CUDA(cudaGetLastError());
  return 0;
}

template <int Q>
static int launch_expand_dw(const CUtensorMap& tw, const CUtensorMap& tx, const FusedParams& p, cudaStream_t stream) {
  auto launch = [&](auto kern) -> int {
    V100_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));   // (not a stream operation)
    int pairs = num_sms() / 2;
    if (p.num_units < pairs) pairs = p.num_units;
    V100_CUDA(launch_pdl(kern, dim3(2 * pairs), dim3(kThreads), kSmem, stream, tw, tx, p));
    return 0;
  };
  return p.dtype == DT_F16 ? launch(expand_dw_kernel<Q, DT_F16>) : launch(expand_dw_kernel<Q, DT_BF16>);
}

int expand_dw(const void* x, int64_t x_pitch, const void* W1, const float* scale1, const float* shift1,
              const uint32_t* dw_pairs, const float* scale2, const float* shift2, void* y, int64_t y_pitch, int B,
              int C_in, int H, int T, int k, int dtype, cudaStream_t stream) {
  if (dtype != DT_BF16 && dtype != DT_F16) return fail(V100_E_INVALID, "expand_dw: dtype must be V100_DTYPE_BF16 or V100_DTYPE_F16");
  if (x == nullptr || W1 == nullptr || shift1 == nullptr || dw_pairs == nullptr || shift2 == nullptr || y == nullptr)
    return fail(V100_E_INVALID, "expand_dw: null pointer");
  if (B <= 0 || C_in <= 0 || H <= 0 || T <= 0) return fail(V100_E_INVALID, "expand_dw: non-positive size");
  if (C_in % 8 != 0) return fail(V100_E_UNSUPPORTED, "expand_dw: C_in=%d must be a multiple of 8", C_in);
  if (H % (2 * kM) != 0) return fail(V100_E_UNSUPPORTED, "expand_dw: hidden width %d must be a multiple of 256", H);
  if (k < 1 || (k & 1) == 0 || k > 83) return fail(V100_E_UNSUPPORTED, "expand_dw: kernel size %d (odd, <= 83)", k);
  if (x_pitch < T || (x_pitch & 7) || y_pitch < T || (y_pitch & 7) || (reinterpret_cast<uintptr_t>(x) & 15) ||
      (reinterpret_cast<uintptr_t>(y) & 15))
    return fail(V100_E_INVALID, "expand_dw: pitches must be multiples of 8 and >= T, bases 16-byte aligned");
  const CUtensorMapDataType tt = dtype == DT_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap tw, tx;
  if (int e = make_tmap_2d(&tw, tt, W1, C_in, H, int64_t(C_in) * 2, 64, 128)) return e;
  if (int e = make_tmap_3d(&tx, tt, x, T, C_in, B, x_pitch * 2, int64_t(C_in) * x_pitch * 2, 64, 64)) return e;
  FusedParams p{};
  p.C_in = C_in; p.H = H; p.B = B; p.T = T;
  p.m_tiles = H / (2 * kM);
  p.t_tiles = (T + kN - 1) / kN;
  p.num_units = p.m_tiles * B;
  p.k_blocks = (C_in + kK - 1) / kK;
  p.p = (k - 1) / 2;
  p.e1 = p.p & 1;
  p.scale1 = scale1; p.shift1 = shift1; p.scale2 = scale2; p.shift2 = shift2;
  p.pairs = dw_pairs;
  p.y = static_cast<unsigned short*>(y);
  p.y_pitch = y_pitch;
  p.dtype = dtype;
  const int Q = (k + 15 + p.e1 + 15) / 16;
  switch (Q) {
    case 1: return launch_expand_dw<1>(tw, tx, p, stream);
    case 2: return launch_expand_dw<2>(tw, tx, p, stream);
    case 3: return launch_expand_dw<3>(tw, tx, p, stream);
    case 4: return launch_expand_dw<4>(tw, tx, p, stream);
    case 5: return launch_expand_dw<5>(tw, tx, p, stream);
    case 6: return launch_expand_dw<6>(tw, tx, p, stream);
    default: return launch_expand_dw<7>(tw, tx, p, stream);
  }
}

}  // namespace v100
