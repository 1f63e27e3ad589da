// Depthwise Conv1d + folded BatchNorm + ReLU6 on NCW bf16 activations.
//
// Two kernels:
//  * dw_mma_kernel (stride 1, the hot one).  With k up to 83 taps the depthwise stage of
//    ConvVoiceEncoder costs 72 MFLOP per audio-second -- on fp32 CUDA cores that is ~2.5x MORE time
//    than its HBM traffic, so the FIR is evaluated on the tensor cores instead, as Toeplitz blocks of the
//    filter times 16-sample columns of the time series (mma.sync.m16n8k16, bf16 x bf16 -> fp32,
//    2 Q MMAs fed by Q + 1 shared-memory fragment loads per 256 outputs, Q = ceil((k+15+e1)/16) <= 7;
//    geometry in the comment above the kernel).  The Toeplitz blocks live in registers while a warp walks 8
//    batch rows of its channel; rows are double-buffered in shared memory with cp.async.  8 channels per CTA.
//  * dw_s2_kernel: stride 2 (the first encoder block, 0.4 % of the depthwise FLOPs, HBM-bound): CUDA cores
//    over a shared-memory staged row.
//  * dw_simt_kernel: any stride / any k, plain CUDA cores; last resort and exported for cross-checking.
#include "common.cuh"
#include "host.h"

#include <cstdlib>

namespace v100 {

constexpr int kDwChunk = 1024;            // outputs per CTA along time (4 double-tiles of 256)
// The staged row is kept as two arrays, even and odd 16-sample blocks: up to 5 double tiles of 16 blocks + Q blocks of
// halo = 87 blocks (44 even, 43 odd).  The odd array starts 736 samples = 1472 B = 64 (mod 128) bytes in, so that the
// 16-byte cp.async chunks of one quarter-warp (two even-block halves, two odd-block halves, ...) fall into distinct
// banks: with the arrays a multiple of 128 bytes apart ncu showed the staging copies costing 2.3x the wavefronts of the
// FIR's own data reads.
constexpr int kDwHalf = 736;
constexpr int kDwRow = kDwHalf + 43 * 16; // 1424 samples staged per row
constexpr int kDwWarps = 8;

template <int DT>
__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                          uint32_t b0, uint32_t b1) {
  if constexpr (DT == DT_F16) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  } else {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
}

constexpr int kDwRowsPerWarp = 8;   // batch rows a warp walks through with one set of Toeplitz fragments
// The staged row starts (p rounded up to V100_DW_PAD + 1) samples before the chunk.  16 samples = 32 bytes: the warp-wide
// cp.async footprint then starts on a 32-byte SECTOR boundary.  With 8-sample (16-byte) rounding the filters whose
// (p rounded up to 8) is an odd multiple of 8 (k = 35, 67, 75) staged from a half-sector offset and ran 16-22 % slower
// (same-box A/B: k = 67/75 372 -> 313 us, k = 35 194 -> 152 us; the others unchanged).
#ifndef V100_DW_PAD
#define V100_DW_PAD 15
#endif
constexpr int kDwPad = V100_DW_PAD;

// Staged-row layout: sample i of the staged row (i = 0 at x[tcA]) lives in 16-sample block i >> 4; even blocks are
// packed into xs[0 .. 704), odd blocks into xs[kDwHalf .. kDwRow).  Eight blocks of ONE parity are then 256
// contiguous bytes -- what one mma's data operand reads (see the kernel comment).
__device__ __forceinline__ int dw_map(int i) { return ((i >> 4) & 1) * kDwHalf + ((i >> 5) << 4) + (i & 15); }

// Stage x[b][c][tcA .. tcA + 8 n_chunks) with zeros outside [0, T).  16-byte chunks never straddle 0 (tcA % 8 == 0);
// a chunk straddling T is copied whole and its tail zeroed by `dw_fix_tail` afterwards.
__device__ __forceinline__ void dw_stage_row(unsigned short* xs, const unsigned short* xrow, int tcA, int T, int lane,
                                             int n_chunks) {
  // chunk v = lane + 32 i: t = t_lane + 256 i and dw_map(8 v) = d_lane + 128 i (block parity = bit 1 of the lane, block
  // pair = 8 i + lane / 4, half block = lane & 1), so the loop is two adds per copy instead of the full index arithmetic
  int t = tcA + 8 * lane;
  unsigned short* dst = xs + ((lane >> 1) & 1) * kDwHalf + ((lane >> 2) << 4) + 8 * (lane & 1);
  for (int v = lane; v < n_chunks; v += 32, t += 256, dst += 128) {
    const bool ok = t >= 0 && t < T;
    cp_async_16(dst, xrow + (ok ? t : 0), ok);
  }
  cp_async_commit();
}
__device__ __forceinline__ void dw_fix_tail(unsigned short* xs, int tcA, int T, int lane) {
  const int i0 = T - tcA;               // first staged index that is past the end of the clip
  if ((T & 7) != 0 && i0 > 0 && i0 < 87 * 16) {
    const int i = i0 + lane;
    if (lane < 8 - (T & 7)) xs[dw_map(i)] = 0;
  }
}

// Toeplitz-on-tensor-cores depthwise FIR.  A "double tile" is 256 consecutive outputs of one (batch, channel) row, seen
// as sixteen 16-sample blocks; accumulator A takes the even blocks, accumulator B the odd ones:
//   A[m][n] = out[tau + m + 32 n],   B[m][n] = out[tau + 16 + m + 32 n],          m < 16, n < 8
//   A = sum_c W_c     * X_c,   c = 0 .. Q-1
//   B = sum_c W_(c-1) * X_c,   c = 1 .. Q          X_c[kk][n] = xs[16 (2 n + c) + kk]   (data block 2 n + c)
//   W_q[m][kk] = wz[16 q + kk - m],  wz[i] = w[i - e1]   (Toeplitz blocks of the filter = the "A" operand, in registers)
// Because the two accumulators are one block apart, data fragment X_c serves BOTH of them: Q + 1 shared-memory
// fragment loads feed 2 Q MMAs (the round-1 kernel gave each of its two tiles private fragments: 2 Q loads, and ncu
// showed it bound by exactly these wavefronts -- L1/shared 89 % busy at 50 % of DRAM bandwidth).  X_c is eight blocks
// of one parity, which the staged layout keeps contiguous: with the mma's kk axis permuted (physical rows
// {2j, 2j+1, 2j+8, 2j+9} carry logical kk = 4j .. 4j+3) a thread's (b0, b1) pair is ONE aligned 64-bit load at uint2
// index 4 (c >> 1) + lane of the parity-(c & 1) array: consecutive lanes read consecutive 8-byte words, 2 wavefronts.
// One xor-shuffle pair regroups an accumulator into bf16x2 pairs that are stored as full 32-byte sectors.
// Alignment: the staged row starts at x[tc0 - pl8] (32-byte aligned in global memory); tiles start
// s = e - (e & 1) outputs before tc0 (e = pl8 - p <= 15), leaving e1 = e & 1 to fold into the zero-extended filter;
// Q = ceil((k + 15 + e1) / 16) <= 7 for k <= 83.
template <int Q, bool RELU6, int DT>
__global__ void __launch_bounds__(kDwWarps * 32)
dw_mma_kernel(const unsigned short* __restrict__ x, long long x_pitch, const unsigned short* __restrict__ w,
              const float* __restrict__ scale, const float* __restrict__ shift, unsigned short* __restrict__ y,
              long long y_pitch, int B, int C, int T, int k) {
  __shared__ __align__(128) unsigned short xs_all[kDwWarps][2][kDwRow];
  __shared__ __align__(16) unsigned short ws_all[kDwWarps][16 * Q + 16];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.y * kDwWarps + warp;
  const int b0 = blockIdx.z * kDwRowsPerWarp;
  const int nb = min(kDwRowsPerWarp, B - b0);
  const int tc0 = blockIdx.x * kDwChunk;
  const int p = (k - 1) >> 1;
  const int pl8 = (p + kDwPad) & ~kDwPad;   // staged row starts at x[tc0 - pl8] so global 16-byte chunks stay aligned
  const int e = pl8 - p;
  const int e1 = e & 1;           // folded into the zero-extended filter
  const int s = e - e1;           // tiles start s outputs before tc0
  const int tcA = tc0 - pl8;
  unsigned short* ws = ws_all[warp];
  const unsigned short* xbase = x + static_cast<long long>(c) * x_pitch;
  const long long xbstride = static_cast<long long>(C) * x_pitch;

  const int len = min(kDwChunk, T - tc0);            // outputs this CTA owns: [tc0, tc0 + len)
  const int n_dt = (len + s + 255) / 256;            // double tiles
  const int n_chunks = min(2 * 87, 2 * (16 * n_dt + Q));       // 16-byte chunks the tiles actually read
  pdl_trigger();
  pdl_wait();        // (x is the previous kernel's output)
  dw_stage_row(xs_all[warp][0], xbase + b0 * xbstride, tcA, T, lane, n_chunks);   // first row in flight

  // zero-extended filter: ws[16 + i] = w[i - e1] for 0 <= i - e1 < k; Toeplitz fragments stay in registers
  for (int i = lane; i < 16 * Q + 16; i += 32) {
    const int j = i - 16 - e1;
    ws[i] = (j >= 0 && j < k) ? w[static_cast<long long>(c) * k + j] : static_cast<unsigned short>(0);
  }
  __syncwarp();
  const int g = lane >> 2, tg = lane & 3;
  uint32_t af[Q][4];
  {
    const unsigned short* wsu = ws;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const int i0 = 16 + 16 * q + 4 * tg - g;   // wz[16q + kk - m] at m = g, kk = 4 tg
      af[q][0] = uint32_t(wsu[i0]) | (uint32_t(wsu[i0 + 1]) << 16);          // (m = g    , kk = 4tg, 4tg+1)
      af[q][1] = uint32_t(wsu[i0 - 8]) | (uint32_t(wsu[i0 - 7]) << 16);      // (m = g + 8, kk = 4tg, 4tg+1)
      af[q][2] = uint32_t(wsu[i0 + 2]) | (uint32_t(wsu[i0 + 3]) << 16);      // (m = g    , kk = 4tg+2, 4tg+3)
      af[q][3] = uint32_t(wsu[i0 - 6]) | (uint32_t(wsu[i0 - 5]) << 16);      // (m = g + 8, kk = 4tg+2, 4tg+3)
    }
  }
  const float sc = scale ? scale[c] : 1.0f;
  const float sh = shift[c];
  const bool even = (g & 1) == 0;
  // after the pair exchange this lane stores outputs (t, t+1) and (t+32, t+33) of an accumulator, t relative to tc0:
  const int pos0 = 64 * tg + (even ? g : g + 7) - s;

  for (int r = 0; r < nb; ++r) {
    unsigned short* xs = xs_all[warp][r & 1];
    if (r + 1 < nb) {   // prefetch the next batch row of this channel into the other buffer
      dw_stage_row(xs_all[warp][(r + 1) & 1], xbase + (b0 + r + 1) * xbstride, tcA, T, lane, n_chunks);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncwarp();
    dw_fix_tail(xs, tcA, T, lane);
    __syncwarp();
    const uint2* xe = reinterpret_cast<const uint2*>(xs) + lane;             // even blocks
    const uint2* xo = reinterpret_cast<const uint2*>(xs + kDwHalf) + lane;   // odd blocks
    unsigned short* yp = y + (static_cast<long long>(b0 + r) * C + c) * y_pitch + tc0 + pos0;
    int pos = pos0;
    auto finish = [&](float (&acc)[4], unsigned short* yq, int posq) {
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fmaf(acc[i], sc, sh);
      // acc = out[tau + g + 64 tg + {0, 32}], out[tau + g + 8 + 64 tg + {0, 32}]; trade with lane g^1 so that even
      // g holds (g, g+1) of the first pair of rows and odd g holds (g-1+8, g+8) of the second
      const float r0 = __shfl_xor_sync(0xffffffffu, even ? acc[2] : acc[0], 4);
      const float r1 = __shfl_xor_sync(0xffffffffu, even ? acc[3] : acc[1], 4);
      const float lo0 = even ? acc[0] : r0, hi0 = even ? r0 : acc[2];
      const float lo1 = even ? acc[1] : r1, hi1 = even ? r1 : acc[3];
      uint32_t o0, o1;
      if (RELU6) {
        o0 = pack2_relu6<DT>(lo0, hi0);
        o1 = pack2_relu6<DT>(lo1, hi1);
      } else {
        o0 = pack2<DT>(lo0, hi0);
        o1 = pack2<DT>(lo1, hi1);
      }
      if (posq >= 0 && posq < len) *reinterpret_cast<uint32_t*>(yq) = o0;
      if (posq + 32 >= 0 && posq + 32 < len) *reinterpret_cast<uint32_t*>(yq + 32) = o1;
    };
#pragma unroll 1
    for (int d = 0; d < n_dt; ++d) {
      float accA[4] = {0.0f, 0.0f, 0.0f, 0.0f}, accB[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
      for (int cq = 0; cq <= Q; ++cq) {
        const uint2 f = (cq & 1) ? xo[4 * (cq >> 1)] : xe[4 * (cq >> 1)];
        if (cq < Q) mma_16816<DT>(accA, af[cq][0], af[cq][1], af[cq][2], af[cq][3], f.x, f.y);
        if (cq > 0) mma_16816<DT>(accB, af[cq - 1][0], af[cq - 1][1], af[cq - 1][2], af[cq - 1][3], f.x, f.y);
      }
      finish(accA, yp, pos);
      finish(accB, yp + 16, pos + 16);
      xe += 32;      // eight blocks of each parity per double tile
      xo += 32;
      yp += 256;
      pos += 256;
    }
    __syncwarp();   // everyone is done reading xs before it is refilled two rows from now
  }
}

// ------------------------------------------------------------------------------------------------------------------
// dw_bulk_kernel: the same Toeplitz FIR with the row staged by ONE bulk copy (cp.async.bulk global -> shared, mbarrier
// completion) instead of ~110 16-byte cp.async chunks per row.  ncu on dw_mma_kernel (k = 83): the L1 data pipe is 84 %
// busy, and 37 % of its wavefronts are the staging copies (LDGSTS: 30 per row into shared memory + 19 on the global side),
// which a bulk copy does not send through the LSU at all.  A bulk copy cannot de-interleave the row by block parity,
// so the staged row is contiguous and the two accumulators of a double tile are the two HALVES of its 256 outputs:
//   A[m][n] = out[tau + m + 16 n],   B[m][n] = out[tau + 128 + m + 16 n],          m < 16, n < 8
//   A = sum_c W_c * X_c,  X_c[kk][n] = xs[16 (n + c) + kk]:  a lane's (b0, b1) is the aligned 64-bit word lane + 4 c
// (consecutive lanes, consecutive words: conflict-free), 2 Q fragment loads per double tile (dw_mma_kernel: Q + 1, plus
// the staging).  Everything outside the copied span is zeroed once per buffer -- every row of a CTA has the same span.
// The copy moves whole 16-byte groups up to T & ~7; the last T & 7 samples of a clip come by plain loads one row ahead
// (generic-proxy stores into bytes no bulk copy ever writes, so the row loop needs no proxy fence).
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int Q, bool RELU6, int DT>
__global__ void __launch_bounds__(kDwWarps * 32, 4)
dw_bulk_kernel(const unsigned short* __restrict__ x, long long x_pitch, const unsigned short* __restrict__ w,
               const float* __restrict__ scale, const float* __restrict__ shift, unsigned short* __restrict__ y,
               long long y_pitch, int B, int C, int T, int k) {
  __shared__ __align__(128) unsigned short xs_all[kDwWarps][2][kDwRow];
  __shared__ __align__(16) unsigned short ws_all[kDwWarps][16 * Q + 16];
  __shared__ __align__(8) uint64_t bars[kDwWarps][2];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.y * kDwWarps + warp;
  const int b0 = blockIdx.z * kDwRowsPerWarp;
  const int nb = min(kDwRowsPerWarp, B - b0);
  const int tc0 = blockIdx.x * kDwChunk;
  const int p = (k - 1) >> 1;
  const int pl8 = (p + kDwPad) & ~kDwPad;   // the staged row starts at x[tc0 - pl8], a multiple of 16 samples
  const int e = pl8 - p;
  const int e1 = e & 1;           // folded into the zero-extended filter
  const int s = e - e1;           // tiles start s outputs before tc0
  const int tcA = tc0 - pl8;
  unsigned short* ws = ws_all[warp];
  const unsigned short* xbase = x + static_cast<long long>(c) * x_pitch;
  const long long xbstride = static_cast<long long>(C) * x_pitch;

  const int len = min(kDwChunk, T - tc0);            // outputs this CTA owns: [tc0, tc0 + len)
  const int n_dt = (len + s + 255) / 256;            // double tiles
  const int need = 16 * (16 * n_dt + Q - 1);         // staged samples the tiles read
  const int T8 = T & ~7;                             // rows are copied in whole 16-byte groups
  const int cs = max(tcA, 0), ce = max(min(tcA + need, T8), cs);
  const uint32_t bytes = uint32_t(ce - cs) * 2u;
  const int dst_off = cs - tcA;
  pdl_trigger();

  // zero both buffers once (halo left of the clip, everything right of the copied span), barriers
  for (int i = lane; i < 2 * kDwRow / 8; i += 32)
    reinterpret_cast<uint4*>(xs_all[warp][0])[i] = make_uint4(0u, 0u, 0u, 0u);
  if (lane == 0) {
    mbar_init(&bars[warp][0], 1);
    mbar_init(&bars[warp][1], 1);
    fence_mbar_init();
  }
  fence_proxy_async();     // the zeros (generic proxy) are ordered before the bulk copies (async proxy) into the same rows
  __syncwarp();
  pdl_wait();              // (x is the previous kernel's output)
  // the clip's last T & 7 samples (staged index tail0 + lane), when the tiles read them
  const int tail0 = T8 - tcA;
  const bool has_tail = lane < (T & 7) && tail0 >= 0 && tail0 + lane < need;
  if (lane == 0 && bytes > 0) {
    mbar_expect_tx(&bars[warp][0], bytes);
    bulk_load(xs_all[warp][0] + dst_off, xbase + b0 * xbstride + cs, bytes, &bars[warp][0]);
  }
  if (has_tail) xs_all[warp][0][tail0 + lane] = xbase[b0 * xbstride + T8 + lane];

  // zero-extended filter: ws[16 + i] = w[i - e1] for 0 <= i - e1 < k; Toeplitz fragments stay in registers
  for (int i = lane; i < 16 * Q + 16; i += 32) {
    const int j = i - 16 - e1;
    ws[i] = (j >= 0 && j < k) ? w[static_cast<long long>(c) * k + j] : static_cast<unsigned short>(0);
  }
  __syncwarp();
  const int g = lane >> 2, tg = lane & 3;
  uint32_t af[Q][4];
  {
    const unsigned short* wsu = ws;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const int i0 = 16 + 16 * q + 4 * tg - g;   // wz[16q + kk - m] at m = g, kk = 4 tg
      af[q][0] = uint32_t(wsu[i0]) | (uint32_t(wsu[i0 + 1]) << 16);          // (m = g    , kk = 4tg, 4tg+1)
      af[q][1] = uint32_t(wsu[i0 - 8]) | (uint32_t(wsu[i0 - 7]) << 16);      // (m = g + 8, kk = 4tg, 4tg+1)
      af[q][2] = uint32_t(wsu[i0 + 2]) | (uint32_t(wsu[i0 + 3]) << 16);      // (m = g    , kk = 4tg+2, 4tg+3)
      af[q][3] = uint32_t(wsu[i0 - 6]) | (uint32_t(wsu[i0 - 5]) << 16);      // (m = g + 8, kk = 4tg+2, 4tg+3)
    }
  }
  const float sc = scale ? scale[c] : 1.0f;
  const float sh = shift[c];
  const bool even = (g & 1) == 0;
  // after the pair exchange this lane stores outputs (t, t+1) and (t+16, t+17) of an accumulator, t relative to tc0:
  const int pos0 = 32 * tg + (even ? g : g + 7) - s;

  for (int r = 0; r < nb; ++r) {
    unsigned short* xs = xs_all[warp][r & 1];
    unsigned short tail_next = 0;
    if (r + 1 < nb) {   // the other buffer was last read in iteration r - 1, which ended with __syncwarp
      if (lane == 0 && bytes > 0) {
        mbar_expect_tx(&bars[warp][(r + 1) & 1], bytes);
        bulk_load(xs_all[warp][(r + 1) & 1] + dst_off, xbase + (b0 + r + 1) * xbstride + cs, bytes, &bars[warp][(r + 1) & 1]);
      }
      if (has_tail) tail_next = xbase[(b0 + r + 1) * xbstride + T8 + lane];   // stored after this row's tiles
    }
    if (bytes > 0) {
      const uint32_t parity = (r >> 1) & 1;
      int spins = 0;
      while (!mbar_try_wait(&bars[warp][r & 1], parity))
        if (++spins > (1 << 28)) __trap();   // a protocol bug must fail the launch, not hang the GPU
    }
    __syncwarp();
    const uint2* x2 = reinterpret_cast<const uint2*>(xs) + lane;
    unsigned short* yp = y + (static_cast<long long>(b0 + r) * C + c) * y_pitch + tc0 + pos0;
    int pos = pos0;
    auto finish = [&](float (&acc)[4], unsigned short* yq, int posq) {
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fmaf(acc[i], sc, sh);
      // acc = out[tau + g + 32 tg + {0, 16}], out[tau + g + 8 + 32 tg + {0, 16}]; trade with lane g^1 so that even
      // g holds (g, g+1) of the first pair of rows and odd g holds (g-1+8, g+8) of the second
      const float r0 = __shfl_xor_sync(0xffffffffu, even ? acc[2] : acc[0], 4);
      const float r1 = __shfl_xor_sync(0xffffffffu, even ? acc[3] : acc[1], 4);
      const float lo0 = even ? acc[0] : r0, hi0 = even ? r0 : acc[2];
      const float lo1 = even ? acc[1] : r1, hi1 = even ? r1 : acc[3];
      uint32_t o0, o1;
      if (RELU6) {
        o0 = pack2_relu6<DT>(lo0, hi0);
        o1 = pack2_relu6<DT>(lo1, hi1);
      } else {
        o0 = pack2<DT>(lo0, hi0);
        o1 = pack2<DT>(lo1, hi1);
      }
      if (posq >= 0 && posq < len) *reinterpret_cast<uint32_t*>(yq) = o0;
      if (posq + 16 >= 0 && posq + 16 < len) *reinterpret_cast<uint32_t*>(yq + 16) = o1;
    };
#pragma unroll 1
    for (int d = 0; d < n_dt; ++d) {
      float accA[4] = {0.0f, 0.0f, 0.0f, 0.0f}, accB[4] = {0.0f, 0.0f, 0.0f, 0.0f};
      if (256 * d + 128 - s < len) {
#pragma unroll
        for (int cq = 0; cq < Q; ++cq) {
          const uint2 fa = x2[4 * cq], fb = x2[4 * cq + 32];
          mma_16816<DT>(accA, af[cq][0], af[cq][1], af[cq][2], af[cq][3], fa.x, fa.y);
          mma_16816<DT>(accB, af[cq][0], af[cq][1], af[cq][2], af[cq][3], fb.x, fb.y);
        }
        finish(accA, yp, pos);
        finish(accB, yp + 128, pos + 128);
      } else {   // the row ends in the first half of this double tile (short rows: T = 100 text tokens, ...)
#pragma unroll
        for (int cq = 0; cq < Q; ++cq) {
          const uint2 fa = x2[4 * cq];
          mma_16816<DT>(accA, af[cq][0], af[cq][1], af[cq][2], af[cq][3], fa.x, fa.y);
        }
        finish(accA, yp, pos);
      }
      x2 += 64;      // sixteen blocks per double tile
      yp += 256;
      pos += 256;
    }
    if (has_tail && r + 1 < nb) xs_all[warp][(r + 1) & 1][tail0 + lane] = tail_next;
    __syncwarp();   // everyone is done reading xs before it is refilled two rows from now
  }
}

// ------------------------------------------------------------------------------------------------------------------
// dw_rows_kernel: short rows (the TTS models: 100 text tokens, ~290 aligned frames).  A bulk copy costs the TMA unit a
// fixed ~65 clk whatever its size, so with 200-600 byte rows dw_bulk_kernel is bound by the NUMBER of copies (120 us for a
// 210 MB layer; eight copies in flight per warp instead of two changed nothing).  Here ONE TMA tensor load brings the
// staged rows of a channel for all eight utterances of the warp: x seen as a (T, C, B) tensor, box = [256 steps x 1
// channel x 8 utterances], unswizzled, out-of-bounds elements zero-filled -- which also supplies the filter halo left of
// the clip and everything right of its end, so there is no zero-initialisation and no tail handling.  One or two boxes
// per warp (staged span <= 512 samples); the tile arithmetic is dw_bulk_kernel's, with the fragment address split into
// (box, block in box) because a fragment's eight blocks may straddle the two boxes.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kDwBoxT = 256;     // steps per TMA box
template <int Q, bool RELU6, int DT>
__global__ void __launch_bounds__(kDwWarps * 32)
dw_rows_kernel(const __grid_constant__ CUtensorMap tm_x, const unsigned short* __restrict__ w,
               const float* __restrict__ scale, const float* __restrict__ shift, unsigned short* __restrict__ y,
               long long y_pitch, int B, int C, int T, int k, int n_boxes) {
  extern __shared__ __align__(128) unsigned char rows_smem[];
  __shared__ __align__(16) unsigned short ws_all[kDwWarps][16 * Q + 16];
  __shared__ __align__(8) uint64_t bars[kDwWarps];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.y * kDwWarps + warp;
  const int b0 = blockIdx.z * kDwRowsPerWarp;
  const int nb = min(kDwRowsPerWarp, B - b0);
  const int p = (k - 1) >> 1;
  const int pl8 = (p + kDwPad) & ~kDwPad;   // the staged rows start at x[-pl8] (a multiple of 16 samples; zeros by OOB fill)
  const int e = pl8 - p;
  const int e1 = e & 1;           // folded into the zero-extended filter
  const int s = e - e1;           // tiles start s outputs before 0
  unsigned short* ws = ws_all[warp];
  const int len = T;                                 // one chunk: the whole row
  const int n_dt = (len + s + 255) / 256;            // double tiles
  const uint32_t box_bytes = kDwRowsPerWarp * kDwBoxT * 2;   // 4 KB: [8 utterances][256 steps]
  unsigned char* xs_w = rows_smem + size_t(warp) * n_boxes * box_bytes;
  pdl_trigger();
  if (lane == 0) {
    mbar_init(&bars[warp], 1);
    fence_mbar_init();
    tma_prefetch_desc(&tm_x);
  }
  __syncwarp();
  pdl_wait();              // (x is the previous kernel's output)
  if (lane == 0) {
    mbar_expect_tx(&bars[warp], uint32_t(n_boxes) * box_bytes);   // out-of-bounds (zero-filled) elements count too
    for (int j = 0; j < n_boxes; ++j)
      tma_load_3d(xs_w + j * box_bytes, &tm_x, &bars[warp], -pl8 + j * kDwBoxT, c, b0);
  }

  // zero-extended filter: ws[16 + i] = w[i - e1] for 0 <= i - e1 < k; Toeplitz fragments stay in registers
  for (int i = lane; i < 16 * Q + 16; i += 32) {
    const int j = i - 16 - e1;
    ws[i] = (j >= 0 && j < k) ? w[static_cast<long long>(c) * k + j] : static_cast<unsigned short>(0);
  }
  __syncwarp();
  const int g = lane >> 2, tg = lane & 3;
  uint32_t af[Q][4];
  {
    const unsigned short* wsu = ws;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const int i0 = 16 + 16 * q + 4 * tg - g;   // wz[16q + kk - m] at m = g, kk = 4 tg
      af[q][0] = uint32_t(wsu[i0]) | (uint32_t(wsu[i0 + 1]) << 16);          // (m = g    , kk = 4tg, 4tg+1)
      af[q][1] = uint32_t(wsu[i0 - 8]) | (uint32_t(wsu[i0 - 7]) << 16);      // (m = g + 8, kk = 4tg, 4tg+1)
      af[q][2] = uint32_t(wsu[i0 + 2]) | (uint32_t(wsu[i0 + 3]) << 16);      // (m = g    , kk = 4tg+2, 4tg+3)
      af[q][3] = uint32_t(wsu[i0 - 6]) | (uint32_t(wsu[i0 - 5]) << 16);      // (m = g + 8, kk = 4tg+2, 4tg+3)
    }
  }
  const float sc = scale ? scale[c] : 1.0f;
  const float sh = shift[c];
  const bool even = (g & 1) == 0;
  // after the pair exchange this lane stores outputs (t, t+1) and (t+16, t+17) of an accumulator:
  const int pos0 = 32 * tg + (even ? g : g + 7) - s;
  {
    int spins = 0;
    while (!mbar_try_wait(&bars[warp], 0))
      if (++spins > (1 << 28)) __trap();   // a protocol bug must fail the launch, not hang the GPU
  }
  __syncwarp();
  // block `blk` of utterance r: uint2 index  (blk >> 4) * (8 rows * 64) + r * 64 + (blk & 15) * 4 + tg   (64 uint2 = 256 steps)
  const uint2* xw = reinterpret_cast<const uint2*>(xs_w) + tg;
  auto frag = [&](int r, int blk) -> uint2 {
    return xw[(blk >> 4) * (kDwRowsPerWarp * 64) + r * 64 + ((blk & 15) << 2)];
  };

  for (int r = 0; r < nb; ++r) {
    unsigned short* yp = y + (static_cast<long long>(b0 + r) * C + c) * y_pitch + pos0;
    int pos = pos0;
    auto finish = [&](float (&acc)[4], unsigned short* yq, int posq) {
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fmaf(acc[i], sc, sh);
      const float r0 = __shfl_xor_sync(0xffffffffu, even ? acc[2] : acc[0], 4);
      const float r1 = __shfl_xor_sync(0xffffffffu, even ? acc[3] : acc[1], 4);
      const float lo0 = even ? acc[0] : r0, hi0 = even ? r0 : acc[2];
      const float lo1 = even ? acc[1] : r1, hi1 = even ? r1 : acc[3];
      uint32_t o0, o1;
      if (RELU6) {
        o0 = pack2_relu6<DT>(lo0, hi0);
        o1 = pack2_relu6<DT>(lo1, hi1);
      } else {
        o0 = pack2<DT>(lo0, hi0);
        o1 = pack2<DT>(lo1, hi1);
      }
      if (posq >= 0 && posq < len) *reinterpret_cast<uint32_t*>(yq) = o0;
      if (posq + 16 >= 0 && posq + 16 < len) *reinterpret_cast<uint32_t*>(yq + 16) = o1;
    };
#pragma unroll 1
    for (int d = 0; d < n_dt; ++d) {
      float accA[4] = {0.0f, 0.0f, 0.0f, 0.0f}, accB[4] = {0.0f, 0.0f, 0.0f, 0.0f};
      const int blkA = 16 * d + g;             // this lane's first block: accumulator A column n = g
      if (256 * d + 128 - s < len) {
#pragma unroll
        for (int cq = 0; cq < Q; ++cq) {
          const uint2 fa = frag(r, blkA + cq), fb = frag(r, blkA + 8 + cq);
          mma_16816<DT>(accA, af[cq][0], af[cq][1], af[cq][2], af[cq][3], fa.x, fa.y);
          mma_16816<DT>(accB, af[cq][0], af[cq][1], af[cq][2], af[cq][3], fb.x, fb.y);
        }
        finish(accA, yp, pos);
        finish(accB, yp + 128, pos + 128);
      } else {   // the row ends in the first half of this double tile
#pragma unroll
        for (int cq = 0; cq < Q; ++cq) {
          const uint2 fa = frag(r, blkA + cq);
          mma_16816<DT>(accA, af[cq][0], af[cq][1], af[cq][2], af[cq][3], fa.x, fa.y);
        }
        finish(accA, yp, pos);
      }
      yp += 256;
      pos += 256;
    }
  }
}

// Stride-2 depthwise conv on the same tensor-core FIR (the first encoder block, asr.py:68), as two polyphase
// stride-1 filters:   out[o] = sum_j w[j] x[2 o + j - p]
//                            = sum_a w_e[a] x_e[o + a - p_e]  +  sum_a w_o[a] x_o[o + a - p_o],
// x_e[i] = x[2 i], x_o[i] = x[2 i + 1];  w_e = the taps j = p (mod 2) (they meet even samples), w_o = the others;
// p_e = floor(p / 2), p_o = ceil(p / 2).  The row is staged de-interleaved twice over -- by sample parity (a PRMT pair
// per 32-bit word on the way from the 16-byte global loads to shared memory) and by 16-sample block parity (the
// layout of dw_mma_kernel) -- and both filters accumulate into the same two accumulators of a 256-output double tile.
// Their zero extensions z_e = z + (p & 1), z_o = z give both the same delay D = z_o + p_o (even), so the tile geometry
// and the store pattern are those of dw_mma_kernel with s = P / 2 - D.
// The CUDA-core version of this kernel (dw_s2_kernel below, kept for long filters) ran at 0.30 of the HBM roofline with
// the issue slots 88 % busy on fp32 FMAs and bf16 unpacking.
template <int Q, int DT, int NV>
__global__ void __launch_bounds__(kDwWarps * 32, NV <= 7 ? 3 : 2)
dw_s2_mma_kernel(const unsigned short* __restrict__ x, long long x_pitch, const unsigned short* __restrict__ w,
                 const float* __restrict__ scale, const float* __restrict__ shift, unsigned short* __restrict__ y,
                 long long y_pitch, int B, int C, int T_in, int T_out, int k, int act) {
  __shared__ __align__(128) unsigned short xs_all[kDwWarps][2][kDwRow];     // [phase e / o][block-parity layout]
  __shared__ __align__(16) unsigned short ws_all[kDwWarps][2][16 * Q + 16];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.y * kDwWarps + warp;
  const int b0 = blockIdx.z * kDwRowsPerWarp;
  const int nb = min(kDwRowsPerWarp, B - b0);
  const int oc0 = blockIdx.x * kDwChunk;          // first output of this CTA
  const int p = (k - 1) >> 1;
  const int p_o = (p + 1) >> 1;                   // taps of the odd-phase filter before its centre
  const int z = p_o & 1;                          // makes the common delay D even
  const int D = z + p_o;
  const int P2 = (D + 7) & ~7;                    // the staged phase arrays start at x_e / x_o index oc0 - P2
  const int s = P2 - D;                           // tiles start s outputs before oc0 (even)
  const int iA = oc0 - P2;                        // phase index of staged sample 0 (x index 2 iA, a multiple of 16)
  pdl_trigger();
  if (c >= C) return;

  const int len = min(kDwChunk, T_out - oc0);
  const int n_dt = (len + s + 255) / 256;
  const int n_vec = min(4 * 87, 4 * (16 * n_dt + Q));   // 16-byte global loads per row: 8 samples = 4 per phase each

  // zero-extended polyphase filters: ws[ph][16 + z_ph + a] = w_ph[a]
  for (int i = lane; i < 2 * (16 * Q + 16); i += 32) {
    const int ph = i / (16 * Q + 16), ii = i - ph * (16 * Q + 16);
    const int first = ph == 0 ? (p & 1) : 1 - (p & 1);           // w_e takes taps j = p (mod 2)
    const int a = ii - 16 - (ph == 0 ? z + (p & 1) : z);
    const int j = first + 2 * a;
    ws_all[warp][ph][ii] = (a >= 0 && j < k) ? w[static_cast<long long>(c) * k + j] : static_cast<unsigned short>(0);
  }
  __syncwarp();
  const int g = lane >> 2, tg = lane & 3;
  uint32_t af[2][Q][4];
#pragma unroll
  for (int ph = 0; ph < 2; ++ph) {
    const unsigned short* wsu = ws_all[warp][ph];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const int i0 = 16 + 16 * q + 4 * tg - g;
      af[ph][q][0] = uint32_t(wsu[i0]) | (uint32_t(wsu[i0 + 1]) << 16);
      af[ph][q][1] = uint32_t(wsu[i0 - 8]) | (uint32_t(wsu[i0 - 7]) << 16);
      af[ph][q][2] = uint32_t(wsu[i0 + 2]) | (uint32_t(wsu[i0 + 3]) << 16);
      af[ph][q][3] = uint32_t(wsu[i0 - 6]) | (uint32_t(wsu[i0 - 5]) << 16);
    }
  }
  const float sc = scale ? scale[c] : 1.0f;
  const float sh = shift[c];
  const bool even = (g & 1) == 0;
  const int pos0 = 64 * tg + (even ? g : g + 7) - s;
  const bool relu6 = act == V100_ACT_RELU6;
  unsigned short* xs_e = xs_all[warp][0];
  unsigned short* xs_o = xs_all[warp][1];
  pdl_wait();

  // The row passes through registers on its way to shared memory (cp.async cannot split 16-bit samples by parity), so
  // the NEXT row's 16-byte loads are issued before this row's MMAs and are in flight while they run: with the loads
  // issued and consumed back to back the kernel sat at 0.44 of the HBM roofline waiting on them.
  constexpr int kS2Vec = NV;   // 16-byte loads per lane and row: 11 covers a full 1024-output chunk, 7 three double tiles
  uint4 pre[kS2Vec];
  // Load i of a lane is vector v = 32 i + lane: x index t = t_lane + 256 i, staged word d = d_lane + 64 i (dw_map(4 v) with
  // the lane part taken out: block parity = bit 2 of the lane, block pair = 4 i + lane / 8, word in block = lane & 3), so the
  // loop below is immediate offsets from two per-lane constants (the generic form cost ~25 integer instructions per load).
  const int t_lane = 2 * iA + 8 * lane;
  const int d_lane = ((lane >> 2) & 1) * kDwHalf + ((lane >> 3) << 4) + 4 * (lane & 3);
  auto load_row = [&](int r) {
    const unsigned short* xrow = x + (static_cast<long long>(b0 + r) * C + c) * x_pitch + t_lane;
#pragma unroll
    for (int i = 0; i < kS2Vec; ++i) {
      const int t = t_lane + 256 * i;                               // x index of the first sample (a multiple of 8)
      pre[i] = make_uint4(0u, 0u, 0u, 0u);
      if (32 * i + lane < n_vec && t >= 0 && t < T_in) pre[i] = __ldg(reinterpret_cast<const uint4*>(xrow + 256 * i));
    }
  };
  load_row(0);
  for (int r = 0; r < nb; ++r) {
    // ---- stage: 8 consecutive samples -> 4 even-phase + 4 odd-phase samples ----
#pragma unroll
    for (int i = 0; i < kS2Vec; ++i) {
      if (32 * i + lane < n_vec) {
        uint4 val = pre[i];
        const int nvalid = T_in - (t_lane + 256 * i);               // samples of this vector inside the clip
        if (nvalid > 0 && nvalid < 8) {                             // the one vector that straddles the end of the clip
          uint32_t* u = reinterpret_cast<uint32_t*>(&val);
#pragma unroll
          for (int wd = 0; wd < 4; ++wd) {
            const int cnt = nvalid - 2 * wd;
            u[wd] &= cnt >= 2 ? 0xFFFFFFFFu : (cnt == 1 ? 0x0000FFFFu : 0u);
          }
        }
        const uint2 ev = make_uint2(__byte_perm(val.x, val.y, 0x5410), __byte_perm(val.z, val.w, 0x5410));
        const uint2 od = make_uint2(__byte_perm(val.x, val.y, 0x7632), __byte_perm(val.z, val.w, 0x7632));
        *reinterpret_cast<uint2*>(xs_e + d_lane + 64 * i) = ev;     // phase samples 4v .. 4v+3: one 8-byte word
        *reinterpret_cast<uint2*>(xs_o + d_lane + 64 * i) = od;
      }
    }
    __syncwarp();
    if (r + 1 < nb) load_row(r + 1);
    unsigned short* yp = y + (static_cast<long long>(b0 + r) * C + c) * y_pitch + oc0 + pos0;
    int pos = pos0;
    auto finish = [&](float (&acc)[4], unsigned short* yq, int posq) {
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fmaf(acc[i], sc, sh);
      const float r0 = __shfl_xor_sync(0xffffffffu, even ? acc[2] : acc[0], 4);
      const float r1 = __shfl_xor_sync(0xffffffffu, even ? acc[3] : acc[1], 4);
      const float lo0 = even ? acc[0] : r0, hi0 = even ? r0 : acc[2];
      const float lo1 = even ? acc[1] : r1, hi1 = even ? r1 : acc[3];
      const uint32_t o0 = relu6 ? pack2_relu6<DT>(lo0, hi0) : pack2<DT>(lo0, hi0);
      const uint32_t o1 = relu6 ? pack2_relu6<DT>(lo1, hi1) : pack2<DT>(lo1, hi1);
      // (a pair that starts on the row's last column also writes the pitch padding next to it: posq is even, so that
      //  column exists whenever the row length is odd -- the pitch is a multiple of 8)
      if (posq >= 0 && posq < len) *reinterpret_cast<uint32_t*>(yq) = o0;
      if (posq + 32 >= 0 && posq + 32 < len) *reinterpret_cast<uint32_t*>(yq + 32) = o1;
    };
#pragma unroll 1
    for (int d = 0; d < n_dt; ++d) {
      float accA[4] = {0.0f, 0.0f, 0.0f, 0.0f}, accB[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
      for (int ph = 0; ph < 2; ++ph) {
        const uint2* xe = reinterpret_cast<const uint2*>(ph == 0 ? xs_e : xs_o) + 32 * d + lane;
        const uint2* xo = reinterpret_cast<const uint2*>((ph == 0 ? xs_e : xs_o) + kDwHalf) + 32 * d + lane;
#pragma unroll
        for (int cq = 0; cq <= Q; ++cq) {
          const uint2 f = (cq & 1) ? xo[4 * (cq >> 1)] : xe[4 * (cq >> 1)];
          if (cq < Q) mma_16816<DT>(accA, af[ph][cq][0], af[ph][cq][1], af[ph][cq][2], af[ph][cq][3], f.x, f.y);
          if (cq > 0) mma_16816<DT>(accB, af[ph][cq - 1][0], af[ph][cq - 1][1], af[ph][cq - 1][2], af[ph][cq - 1][3], f.x, f.y);
        }
      }
      finish(accA, yp, pos);
      finish(accB, yp + 16, pos + 16);
      yp += 256;
      pos += 256;
    }
    __syncwarp();   // everyone is done reading the staged row before the next one overwrites it
  }
}

// Stride-2 depthwise conv (the first encoder block, asr.py:68): one warp per (batch, channel) row,
// the row segment staged in shared memory with coalesced 16-byte loads, each lane producing outputs
// o = oc0 + 32 r + lane so that the 32-bit input words (x[2o+2m], x[2o+2m+1]) are consecutive across lanes.
constexpr int kS2Chunk = 512;                 // outputs per CTA pass
constexpr int kS2Row = 2 * kS2Chunk + 96;     // staged inputs (halo <= 48 each side)
constexpr int kS2MaxWords = 48;               // (k + 1) / 2 + 1 <= 48  ->  k <= 93

template <int DT>
__global__ void __launch_bounds__(kDwWarps * 32)
dw_s2_kernel(const unsigned short* __restrict__ x, long long x_pitch, const unsigned short* __restrict__ w,
             const float* __restrict__ scale, const float* __restrict__ shift, unsigned short* __restrict__ y,
             long long y_pitch, int C, int T_in, int T_out, int k, int act) {
  __shared__ __align__(16) unsigned short xs_all[kDwWarps][kS2Row];
  __shared__ __align__(8) float ws_all[kDwWarps][2 * kS2MaxWords];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.y * kDwWarps + warp;
  pdl_trigger();
  if (c >= C) return;
  const int b = blockIdx.z;
  const int oc0 = blockIdx.x * kS2Chunk;
  const int p = (k - 1) >> 1;
  const int p8 = (p + 7) & ~7;
  const int e = p8 - p;                       // xs[2*ol + j + e] = x[2*(oc0+ol) + j - p]
  unsigned short* xs = xs_all[warp];
  float* ws = ws_all[warp];
  pdl_wait();
  const unsigned short* xrow = x + (static_cast<long long>(b) * C + c) * x_pitch;
  const int ia = 2 * oc0 - p8;
  for (int v = lane; v < kS2Row / 8; v += 32) {
    const int t = ia + v * 8;
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (t >= 0 && t < T_in) {
      val = *reinterpret_cast<const uint4*>(xrow + t);
      if (t + 8 > T_in) {
        uint32_t* u = reinterpret_cast<uint32_t*>(&val);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (t + i >= T_in) u[i >> 1] &= (i & 1) ? 0x0000FFFFu : 0xFFFF0000u;
      }
    }
    *reinterpret_cast<uint4*>(xs + v * 8) = val;
  }
  const int n_words = (k + e + 1) >> 1;       // taps j' = j + e in [e, k+e) -> words [0, n_words)
  for (int i = lane; i < 2 * n_words; i += 32) {
    const int j = i - e;
    ws[i] = (j >= 0 && j < k) ? h2f<DT>(w[static_cast<long long>(c) * k + j]) : 0.0f;
  }
  __syncwarp();
  const float sc = scale ? scale[c] : 1.0f, sh = shift[c];
  const uint32_t* xw = reinterpret_cast<const uint32_t*>(xs);
  const float2* w2 = reinterpret_cast<const float2*>(ws);
  unsigned short* yrow = y + (static_cast<long long>(b) * C + c) * y_pitch;
  for (int r0 = 0; r0 < kS2Chunk / 32 && oc0 + r0 * 32 < T_out; r0 += 4) {
    float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    for (int m = 0; m < n_words; ++m) {
      const float2 wm = w2[m];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t v = xw[(r0 + u) * 32 + lane + m];
        acc[u] = fmaf(wm.x, unpack_lo<DT>(v), acc[u]);
        acc[u] = fmaf(wm.y, unpack_hi<DT>(v), acc[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int o = oc0 + (r0 + u) * 32 + lane;
      float v = fmaf(acc[u], sc, sh);
      if (act == V100_ACT_RELU6) v = fminf(fmaxf(v, 0.0f), 6.0f);
      if (o < T_out) yrow[o] = f2h<DT>(v);
    }
  }
}

template <int DT>
__global__ void __launch_bounds__(128)
dw_simt_kernel(const unsigned short* __restrict__ x, long long x_pitch, const unsigned short* __restrict__ w,
               const float* __restrict__ scale, const float* __restrict__ shift, unsigned short* __restrict__ y,
               long long y_pitch, int C, int T_in, int T_out, int k, int stride, int act) {
  const int c = blockIdx.y, b = blockIdx.z;
  const int o0 = (blockIdx.x * 128 + threadIdx.x) * 8;
  pdl_trigger();
  pdl_wait();
  if (o0 >= T_out) return;
  const int p = (k - 1) >> 1;
  const unsigned short* xrow = x + (static_cast<long long>(b) * C + c) * x_pitch;
  const unsigned short* wrow = w + static_cast<long long>(c) * k;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int j = 0; j < k; ++j) {
    const float wj = h2f<DT>(wrow[j]);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int i = (o0 + r) * stride + j - p;
      if (i >= 0 && i < T_in) acc[r] = fmaf(wj, h2f<DT>(xrow[i]), acc[r]);
    }
  }
  const float sc = scale ? scale[c] : 1.0f, sh = shift[c];
  uint32_t o[4];
#pragma unroll
  for (int r = 0; r < 8; r += 2) {
    float v0 = fmaf(acc[r], sc, sh), v1 = fmaf(acc[r + 1], sc, sh);
    if (act == V100_ACT_RELU6) {
      v0 = fminf(fmaxf(v0, 0.0f), 6.0f);
      v1 = fminf(fmaxf(v1, 0.0f), 6.0f);
    }
    o[r >> 1] = pack2<DT>(v0, v1);
  }
  // o0 % 8 == 0 and pitch % 8 == 0, so the 16-byte store stays inside the row's pitch
  *reinterpret_cast<uint4*>(y + (static_cast<long long>(b) * C + c) * y_pitch + o0) = make_uint4(o[0], o[1], o[2], o[3]);
}

template <int Q, int DT>
static void launch_dw_mma_dt(const void* x, int64_t x_pitch, const void* w, const float* scale, const float* shift,
                             void* y, int64_t y_pitch, int B, int C, int T, int k, int act, cudaStream_t stream) {
  dim3 grid((T + kDwChunk - 1) / kDwChunk, C / kDwWarps, (B + kDwRowsPerWarp - 1) / kDwRowsPerWarp);
  auto xp = static_cast<const unsigned short*>(x);
  auto wp = static_cast<const unsigned short*>(w);
  auto yp = static_cast<unsigned short*>(y);
  const long long xpl = x_pitch, ypl = y_pitch;
  // Same-box A/B (tools/dw_time.py, B = 256, T = 751): the bulk-staged kernel wins where staging is a large share of the
  // row's shared-memory traffic -- k = 19/27 155/146 -> 131-137 us, k = 35 152 -> 140, k = 51 156 -> 152, k = 59 303 -> 295 --
  // and loses where its 2 Q fragment loads (3 wavefronts each unless 128-byte aligned) dominate: k = 67/75 314 -> 324,
  // k = 83 339 -> 348.  V100_DW_BULK: 0 = never, 2 = always (A/B runs).
  static const int bulk = getenv("V100_DW_BULK") ? atoi(getenv("V100_DW_BULK")) : 1;
  if (bulk == 2 || (bulk == 1 && Q <= 5)) {
    if (act == V100_ACT_RELU6)
      launch_pdl(dw_bulk_kernel<Q, true, DT>, grid, dim3(kDwWarps * 32), 0, stream, xp, xpl, wp, scale, shift, yp, ypl, B, C, T, k);
    else
      launch_pdl(dw_bulk_kernel<Q, false, DT>, grid, dim3(kDwWarps * 32), 0, stream, xp, xpl, wp, scale, shift, yp, ypl, B, C, T, k);
    return;
  }
  if (act == V100_ACT_RELU6)
    launch_pdl(dw_mma_kernel<Q, true, DT>, grid, dim3(kDwWarps * 32), 0, stream, xp, xpl, wp, scale, shift, yp, ypl, B, C, T, k);
  else
    launch_pdl(dw_mma_kernel<Q, false, DT>, grid, dim3(kDwWarps * 32), 0, stream, xp, xpl, wp, scale, shift, yp, ypl, B, C, T, k);
}

template <int Q>
static void launch_dw_mma(const void* x, int64_t x_pitch, const void* w, const float* scale, const float* shift,
                          void* y, int64_t y_pitch, int B, int C, int T, int k, int act, int dtype, cudaStream_t stream) {
  if (dtype == DT_F16) launch_dw_mma_dt<Q, DT_F16>(x, x_pitch, w, scale, shift, y, y_pitch, B, C, T, k, act, stream);
  else launch_dw_mma_dt<Q, DT_BF16>(x, x_pitch, w, scale, shift, y, y_pitch, B, C, T, k, act, stream);
}

int dwconv1d(const void* x, int64_t x_pitch, const void* w, const float* scale, const float* shift, void* y,
             int64_t y_pitch, int B, int C, int T_in, int k, int stride, int act, int dtype, int force_simt,
             cudaStream_t stream) {
  if (dtype != DT_BF16 && dtype != DT_F16) return fail(V100_E_INVALID, "dwconv: dtype must be V100_DTYPE_BF16 or V100_DTYPE_F16");
  if (B <= 0 || C <= 0 || T_in <= 0 || k <= 0 || stride <= 0) return fail(V100_E_INVALID, "dwconv: non-positive size");
  if ((k & 1) == 0) return fail(V100_E_UNSUPPORTED, "dwconv: kernel size %d must be odd", k);
  if (x == nullptr || w == nullptr || shift == nullptr || y == nullptr) return fail(V100_E_INVALID, "dwconv: null pointer");
  const int T_out = (T_in - 1) / stride + 1;
  if (x_pitch < T_in || (x_pitch & 7) || y_pitch < T_out || (y_pitch & 7) ||
      (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 15))
    return fail(V100_E_INVALID, "dwconv: pitches must be multiples of 8 and >= T, bases 16-byte aligned");
  if (B > 65535 || C > 65535 * kDwWarps) return fail(V100_E_UNSUPPORTED, "dwconv: B or C too large for the grid");
  const int p = (k - 1) / 2;
  const int e1 = (((p + kDwPad) & ~kDwPad) - p) & 1;
  const int Q = (k + 15 + e1 + 15) / 16;
  auto xp = static_cast<const unsigned short*>(x);
  auto wp = static_cast<const unsigned short*>(w);
  auto yp = static_cast<unsigned short*>(y);
  // short rows: one TMA tensor load per (channel, eight utterances) -- see dw_rows_kernel
  {
    static const int rows_on = getenv("V100_DW_ROWS") ? atoi(getenv("V100_DW_ROWS")) : 1;   // A/B runs
    const int pl8 = (p + kDwPad) & ~kDwPad, s = (pl8 - p) - e1;
    const int n_dt = (T_in + s + 255) / 256;
    const bool last_half = 256 * (n_dt - 1) + 128 - s >= T_in;
    const int need = 16 * (16 * n_dt - (last_half ? 8 : 0) + Q - 1);
    if (rows_on && !force_simt && stride == 1 && (C % kDwWarps) == 0 && Q <= 7 && need <= 2 * kDwBoxT) {
      const int n_boxes = need <= kDwBoxT ? 1 : 2;
      CUtensorMap tm;
      if (int e = make_tmap_rows(&tm, dtype == DT_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, x,
                                 T_in, C, B, x_pitch * 2, kDwBoxT, kDwRowsPerWarp)) return e;
      const size_t smem = size_t(kDwWarps) * n_boxes * kDwRowsPerWarp * kDwBoxT * 2;
      dim3 grid(1, C / kDwWarps, (B + kDwRowsPerWarp - 1) / kDwRowsPerWarp);
      const long long ypl = y_pitch;
#define V100_ROWS(QQ, RL, DTT)                                                                                              \
      do {                                                                                                                  \
        auto kern = dw_rows_kernel<QQ, RL, DTT>;                                                                            \
        V100_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));                      \
        V100_CUDA(launch_pdl(kern, grid, dim3(kDwWarps * 32), smem, stream, tm, wp, scale, shift, yp, ypl, B, C, T_in, k, n_boxes)); \
      } while (0)
#define V100_ROWS_Q(QQ)                                                                       \
      do {                                                                                    \
        if (dtype == DT_F16) { if (act == V100_ACT_RELU6) V100_ROWS(QQ, true, DT_F16); else V100_ROWS(QQ, false, DT_F16); } \
        else { if (act == V100_ACT_RELU6) V100_ROWS(QQ, true, DT_BF16); else V100_ROWS(QQ, false, DT_BF16); }               \
      } while (0)
      switch (Q) {
        case 1: V100_ROWS_Q(1); break;
        case 2: V100_ROWS_Q(2); break;
        case 3: V100_ROWS_Q(3); break;
        case 4: V100_ROWS_Q(4); break;
        case 5: V100_ROWS_Q(5); break;
        case 6: V100_ROWS_Q(6); break;
        default: V100_ROWS_Q(7); break;
      }
#undef V100_ROWS_Q
#undef V100_ROWS
      return 0;
    }
  }
  if (!force_simt && stride == 1 && (C % kDwWarps) == 0 && Q <= 7) {
    switch (Q) {
      case 1: launch_dw_mma<1>(x, x_pitch, w, scale, shift, y, y_pitch, B, C, T_in, k, act, dtype, stream); break;
      case 2: launch_dw_mma<2>(x, x_pitch, w, scale, shift, y, y_pitch, B, C, T_in, k, act, dtype, stream); break;
      case 3: launch_dw_mma<3>(x, x_pitch, w, scale, shift, y, y_pitch, B, C, T_in, k, act, dtype, stream); break;
      case 4: launch_dw_mma<4>(x, x_pitch, w, scale, shift, y, y_pitch, B, C, T_in, k, act, dtype, stream); break;
      case 5: launch_dw_mma<5>(x, x_pitch, w, scale, shift, y, y_pitch, B, C, T_in, k, act, dtype, stream); break;
      case 6: launch_dw_mma<6>(x, x_pitch, w, scale, shift, y, y_pitch, B, C, T_in, k, act, dtype, stream); break;
      default: launch_dw_mma<7>(x, x_pitch, w, scale, shift, y, y_pitch, B, C, T_in, k, act, dtype, stream); break;
    }
  } else if (!force_simt && stride == 2 && k <= 59) {
    // polyphase tensor-core kernel: each phase filter has <= 30 taps -> Q <= 3
    const int pp = (k - 1) / 2, p_o = (pp + 1) / 2, zz = p_o & 1;
    const int ke = (k - (pp & 1) + 1) / 2, ko = (k - (1 - (pp & 1)) + 1) / 2;
    const int q_e = (ke + 15 + zz + (pp & 1) + 15) / 16, q_o = (ko + 15 + zz + 15) / 16;
    const int Qs = q_e > q_o ? q_e : q_o;
    dim3 grid((T_out + kDwChunk - 1) / kDwChunk, (C + kDwWarps - 1) / kDwWarps, (B + kDwRowsPerWarp - 1) / kDwRowsPerWarp);
    const long long xpl = x_pitch, ypl = y_pitch;
    // 16-byte loads per row of the longest chunk (the kernel's n_vec), which sizes the register prefetch
    const int len0 = T_out < kDwChunk ? T_out : kDwChunk;
    const bool small = 4 * (16 * ((len0 + 14 + 255) / 256) + 3) <= 7 * 32;
#define V100_S2(QQ, NV)                                                                                                     \
    do {                                                                                                                    \
      if (dtype == DT_F16)                                                                                                  \
        launch_pdl(dw_s2_mma_kernel<QQ, DT_F16, NV>, grid, dim3(kDwWarps * 32), 0, stream, xp, xpl, wp, scale, shift, yp, ypl, B, C, T_in, T_out, k, act); \
      else                                                                                                                  \
        launch_pdl(dw_s2_mma_kernel<QQ, DT_BF16, NV>, grid, dim3(kDwWarps * 32), 0, stream, xp, xpl, wp, scale, shift, yp, ypl, B, C, T_in, T_out, k, act); \
    } while (0)
    if (Qs <= 2) { if (small) V100_S2(2, 7); else V100_S2(2, 11); }
    else { if (small) V100_S2(3, 7); else V100_S2(3, 11); }
#undef V100_S2
  } else if (!force_simt && stride == 2 && k <= 2 * kS2MaxWords - 3) {
    dim3 grid((T_out + kS2Chunk - 1) / kS2Chunk, (C + kDwWarps - 1) / kDwWarps, B);
    const long long xpl = x_pitch, ypl = y_pitch;
    if (dtype == DT_F16)
      launch_pdl(dw_s2_kernel<DT_F16>, grid, dim3(kDwWarps * 32), 0, stream, xp, xpl, wp, scale, shift, yp, ypl, C, T_in, T_out, k, act);
    else
      launch_pdl(dw_s2_kernel<DT_BF16>, grid, dim3(kDwWarps * 32), 0, stream, xp, xpl, wp, scale, shift, yp, ypl, C, T_in, T_out, k, act);
  } else {
    dim3 grid((T_out + 1023) / 1024, C, B);
    const long long xpl = x_pitch, ypl = y_pitch;
    if (dtype == DT_F16)
      launch_pdl(dw_simt_kernel<DT_F16>, grid, dim3(128), 0, stream, xp, xpl, wp, scale, shift, yp, ypl, C, T_in, T_out, k, stride, act);
    else
      launch_pdl(dw_simt_kernel<DT_BF16>, grid, dim3(128), 0, stream, xp, xpl, wp, scale, shift, yp, ypl, C, T_in, T_out, k, stride, act);
  }
  V100_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace v100
