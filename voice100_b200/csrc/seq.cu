// Kernels of the v2 models (voice100/models/_layers_v2.py, _asr_v2.py, _align_v2.py, _tts_v2.py):
//   * dense Conv1d (k = 3/5, stride 1/2) = tap-stacking copy + the tcgen05 GEMM of conv_gemm.cu
//   * LayerNorm over channels + exact GELU on the NCW layout
//   * NCW <-> time-major layout changes around the recurrent layers
//   * one bidirectional LSTM layer as a persistent tcgen05 kernel
//
// Time-major layout ("TM"): x[c][t * Bp + b], Bp = batch rounded up to 8.  One LSTM step then touches one
// contiguous run of columns, and the input projection W_ih x of ALL steps is one plain 1x1-conv GEMM over
// T*Bp columns (v100_conv1x1 with B = 1).
//
// LSTM recurrence.  Per direction the batch is cut into groups of 128 utterances (the UMMA M dimension) and the
// hidden units into slices of 16; one CTA owns (direction, group, slice) for the whole sequence:
//   - its 64 rows of W_hh (4 gates x 16 units, all H columns) stay in shared memory for all T steps;
//   - every step it TMA-loads the group's h_{t-1} [128 x H] from a small exchange buffer in global memory (L2
//     resident), runs  acc[128 x 64] = h_{t-1} W_slice^T  on the tensor core (accumulator in TMEM),
//     adds the precomputed input projection, applies the gates (cell state lives in registers: thread = one
//     utterance, 16 units), writes its 16 units of h_t to the exchange buffer and to the layer output;
//   - the H/16 CTAs of a (direction, group) synchronise once per step through a monotonically increasing
//     counter in global memory (release add / acquire poll).  All CTAs of a launch are co-resident (grid <= SMs,
//     one CTA per SM), which is what makes the spin wait safe; the wait is bounded and traps on a protocol bug.
#include "common.cuh"
#include "host.h"

#include <cstdlib>

namespace v100 {

// ------------------------------------------------------------------------------------------------
// tap stacking for the dense Conv1d:  xs[b][j*C + c][to] = x[b][c][to*S + j - pad]  (zero outside [0, T_in))
// ------------------------------------------------------------------------------------------------
template <int K, int S>
__global__ void __launch_bounds__(256)
tap_stack_kernel(const unsigned short* __restrict__ x, long long x_pitch, unsigned short* __restrict__ xs,
                 long long xs_pitch, int C, int T_in, int pad) {
  const int b = blockIdx.z, c = blockIdx.y;
  const int to0 = (blockIdx.x * 256 + threadIdx.x) * 8;
  if (to0 >= xs_pitch) return;
  const unsigned short* row = x + (static_cast<long long>(b) * C + c) * x_pitch;
  constexpr int W = 7 * S + K;  // input window of 8 outputs
  unsigned short v[W];
  const int ti0 = to0 * S - pad;
#pragma unroll
  for (int i = 0; i < W; ++i) {
    const int t = ti0 + i;
    v[i] = (t >= 0 && t < T_in) ? row[t] : (unsigned short)0;
  }
#pragma unroll
  for (int j = 0; j < K; ++j) {
    uint4 o;
    uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
    for (int i = 0; i < 4; ++i) ow[i] = uint32_t(v[(2 * i) * S + j]) | (uint32_t(v[(2 * i + 1) * S + j]) << 16);
    *reinterpret_cast<uint4*>(xs + ((static_cast<long long>(b) * K + j) * C + c) * xs_pitch + to0) = o;
  }
}

int conv1d(const void* x, int64_t x_pitch, const void* Wp, const float* bias, void* workspace, void* y,
           int64_t y_pitch, int B, int C_in, int C_out, int T_in, int k, int stride, int pad, int dtype,
           cudaStream_t stream) {
  if (B <= 0 || C_in <= 0 || C_out <= 0 || T_in <= 0 || pad < 0) return fail(V100_E_INVALID, "conv1d: bad size");
  if (!((k == 3 || k == 5) && (stride == 1 || stride == 2)))
    return fail(V100_E_UNSUPPORTED, "conv1d: kernel_size %d / stride %d (supported: k 3 or 5, stride 1 or 2)", k, stride);
  if (B > 65535 || C_in > 65535) return fail(V100_E_UNSUPPORTED, "conv1d: B or C_in too large for the grid");
  if (T_in + 2 * pad < k) return fail(V100_E_INVALID, "conv1d: input shorter than the kernel");
  if (x == nullptr || workspace == nullptr || (reinterpret_cast<uintptr_t>(workspace) & 15) != 0)
    return fail(V100_E_INVALID, "conv1d: null x or null/unaligned workspace");
  if (x_pitch < T_in) return fail(V100_E_INVALID, "conv1d: x pitch < T_in");
  const int T_out = (T_in + 2 * pad - k) / stride + 1;
  if (y_pitch < T_out || (y_pitch & 7) != 0) return fail(V100_E_INVALID, "conv1d: y pitch must be >= T_out=%d and a multiple of 8", T_out);
  dim3 grid((unsigned)((y_pitch / 8 + 255) / 256), C_in, B);
  const unsigned short* xi = static_cast<const unsigned short*>(x);
  unsigned short* xs = static_cast<unsigned short*>(workspace);
  if (k == 3 && stride == 1) tap_stack_kernel<3, 1><<<grid, 256, 0, stream>>>(xi, x_pitch, xs, y_pitch, C_in, T_in, pad);
  else if (k == 3) tap_stack_kernel<3, 2><<<grid, 256, 0, stream>>>(xi, x_pitch, xs, y_pitch, C_in, T_in, pad);
  else if (stride == 1) tap_stack_kernel<5, 1><<<grid, 256, 0, stream>>>(xi, x_pitch, xs, y_pitch, C_in, T_in, pad);
  else tap_stack_kernel<5, 2><<<grid, 256, 0, stream>>>(xi, x_pitch, xs, y_pitch, C_in, T_in, pad);
  V100_CUDA(cudaGetLastError());
  return conv1x1(workspace, y_pitch, Wp, nullptr, bias, nullptr, y, y_pitch, B, k * C_in, C_out, T_out, V100_ACT_NONE,
                 dtype, stream);
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over channels + GELU(erf), NCW -> NCW.  One CTA = one utterance x 64 time columns; the [C x 64]
// tile is staged in shared memory once, statistics are two-pass (mean, then centred sum of squares) in fp32.
// ------------------------------------------------------------------------------------------------
// gelu(x) = x * Phi(x), Phi(x) = 0.5 * erfc(-x / sqrt(2)).  erfc by Abramowitz & Stegun 7.1.26 (|abs err| < 1.5e-7),
// evaluated on the side without cancellation: Phi(x) = 0.5 * q for x < 0 and 1 - 0.5 * q for x >= 0 with
// q = poly(t) * exp(-x^2 / 2), t = 1 / (1 + p |x| / sqrt(2)).  A third of the instructions of erff(); the result is
// rounded to a 16-bit type right after.
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float q = poly * t * __expf(-z * z);
  const float phi = x < 0.0f ? 0.5f * q : 1.0f - 0.5f * q;
  return x * phi;
}

template <int DT>
__global__ void __launch_bounds__(256)
layernorm_gelu_kernel(const uint32_t* __restrict__ x, long long x_pitch, const float* __restrict__ gamma,
                      const float* __restrict__ beta, float eps, uint32_t* __restrict__ y, long long y_pitch, int C,
                      int T) {
  extern __shared__ uint32_t ln_tile[];  // [C][32] column pairs
  __shared__ float red[8][64];
  __shared__ float stat[2][64];
  const int b = blockIdx.y, t0 = blockIdx.x * 64;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long xcol = (t0 >> 1) + lane;  // in 32-bit units
  const bool live = t0 + 2 * lane < x_pitch;
  const uint32_t* xb = x + static_cast<long long>(b) * C * (x_pitch >> 1);
  float s0 = 0.0f, s1 = 0.0f;
#pragma unroll 8
  for (int c = w; c < C; c += 8) {
    const uint32_t v = live ? __ldg(xb + c * (x_pitch >> 1) + xcol) : 0u;
    ln_tile[c * 32 + lane] = v;
    s0 += unpack_lo<DT>(v);
    s1 += unpack_hi<DT>(v);
  }
  red[w][2 * lane] = s0;
  red[w][2 * lane + 1] = s1;
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
    stat[0][threadIdx.x] = s / float(C);
  }
  __syncthreads();
  const float m0 = stat[0][2 * lane], m1 = stat[0][2 * lane + 1];
  s0 = 0.0f;
  s1 = 0.0f;
  for (int c = w; c < C; c += 8) {
    const uint32_t v = ln_tile[c * 32 + lane];
    const float d0 = unpack_lo<DT>(v) - m0, d1 = unpack_hi<DT>(v) - m1;
    s0 = fmaf(d0, d0, s0);
    s1 = fmaf(d1, d1, s1);
  }
  red[w][2 * lane] = s0;
  red[w][2 * lane + 1] = s1;
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
    stat[1][threadIdx.x] = rsqrtf(s / float(C) + eps);
  }
  __syncthreads();
  const float r0 = stat[1][2 * lane], r1 = stat[1][2 * lane + 1];
  if (t0 + 2 * lane >= y_pitch) return;
  uint32_t* yb = y + static_cast<long long>(b) * C * (y_pitch >> 1) + xcol;
#pragma unroll 4
  for (int c = w; c < C; c += 8) {
    const uint32_t v = ln_tile[c * 32 + lane];
    const float g = __ldg(gamma + c), be = __ldg(beta + c);
    float a0 = fmaf((unpack_lo<DT>(v) - m0) * r0, g, be);
    float a1 = fmaf((unpack_hi<DT>(v) - m1) * r1, g, be);
    a0 = gelu_erf(a0);
    a1 = gelu_erf(a1);
    // columns past T hold whatever the producer left in the row padding: keep them finite
    if (t0 + 2 * lane >= T) a0 = 0.0f;
    if (t0 + 2 * lane + 1 >= T) a1 = 0.0f;
    yb[c * (y_pitch >> 1)] = pack2<DT>(a0, a1);
  }
}

int layernorm_gelu(const void* x, int64_t x_pitch, const float* gamma, const float* beta, float eps, void* y,
                   int64_t y_pitch, int B, int C, int T, int dtype, cudaStream_t stream) {
  if (x == nullptr || y == nullptr || gamma == nullptr || beta == nullptr) return fail(V100_E_INVALID, "layernorm_gelu: null pointer");
  if (dtype != DT_BF16 && dtype != DT_F16) return fail(V100_E_INVALID, "layernorm_gelu: bad dtype");
  if (B <= 0 || C <= 0 || T <= 0 || B > 65535) return fail(V100_E_INVALID, "layernorm_gelu: bad size");
  if (x_pitch < T || y_pitch < T || (x_pitch & 1) || (y_pitch & 1)) return fail(V100_E_INVALID, "layernorm_gelu: pitch must be even and >= T");
  const size_t smem = size_t(C) * 32 * 4;
  if (smem > 200 * 1024) return fail(V100_E_UNSUPPORTED, "layernorm_gelu: C=%d too large (max 1600)", C);
  auto kern = dtype == DT_F16 ? layernorm_gelu_kernel<DT_F16> : layernorm_gelu_kernel<DT_BF16>;
  V100_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  dim3 grid((T + 63) / 64, B);
  kern<<<grid, 256, smem, stream>>>(static_cast<const uint32_t*>(x), x_pitch, gamma, beta, eps,
                                    static_cast<uint32_t*>(y), y_pitch, C, T);
  V100_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// NCW [B][C][pitch]  <->  time-major [C][T*Bp]   (16-bit elements, 64 x 64 tiles through shared memory)
// ------------------------------------------------------------------------------------------------
// tile = 64 time steps x 64 utterances of one channel; both the reads (64 steps of an utterance row) and the
// writes (64 utterances of a step) are 128-byte segments
__global__ void __launch_bounds__(256)
ncw_to_tm_kernel(const unsigned short* __restrict__ x, long long pitch, unsigned short* __restrict__ y, int B, int C,
                 int T, int Bp) {
  __shared__ uint32_t tile[64][33];  // [utterance][pair of time steps]
  const int c = blockIdx.z, t0 = blockIdx.x * 64, b0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 64; r += 8) {
    const int b = b0 + r, t = t0 + 2 * tx;
    uint32_t v = 0;
    if (b < B && t < pitch) {  // pitch is even; columns in [T, pitch) are masked below
      v = *reinterpret_cast<const uint32_t*>(x + (static_cast<long long>(b) * C + c) * pitch + t);
      if (t >= T) v = 0;
      else if (t + 1 >= T) v &= 0xFFFFu;
    }
    tile[r][tx] = v;
  }
  __syncthreads();
  for (int r = ty; r < 64; r += 8) {  // r = time step within the tile, tx = pair of utterances
    const int t = t0 + r, b = b0 + 2 * tx;
    if (t < T && b < Bp) {
      const uint32_t lo = tile[2 * tx][r >> 1], hi = tile[2 * tx + 1][r >> 1];
      const uint32_t v = (r & 1) ? ((lo >> 16) | (hi & 0xFFFF0000u)) : ((lo & 0xFFFFu) | (hi << 16));
      *reinterpret_cast<uint32_t*>(y + static_cast<long long>(c) * T * Bp + static_cast<long long>(t) * Bp + b) = v;
    }
  }
}

__global__ void __launch_bounds__(256)
tm_to_ncw_kernel(const unsigned short* __restrict__ x, unsigned short* __restrict__ y, long long pitch, int B, int C,
                 int T, int Bp) {
  __shared__ uint32_t tile[64][33];  // [time step][pair of utterances]
  const int c = blockIdx.z, t0 = blockIdx.x * 64, b0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 64; r += 8) {
    const int t = t0 + r, b = b0 + 2 * tx;
    uint32_t v = 0;
    if (t < T && b < Bp)  // Bp is even; utterances in [B, Bp) are zero in a TM tensor
      v = *reinterpret_cast<const uint32_t*>(x + static_cast<long long>(c) * T * Bp + static_cast<long long>(t) * Bp + b);
    tile[r][tx] = v;
  }
  __syncthreads();
  for (int r = ty; r < 64; r += 8) {  // r = utterance within the tile, tx = pair of time steps
    const int b = b0 + r, t = t0 + 2 * tx;
    if (b < B && t < pitch) {
      const uint32_t lo = tile[2 * tx][r >> 1], hi = tile[2 * tx + 1][r >> 1];
      const uint32_t v = (r & 1) ? ((lo >> 16) | (hi & 0xFFFF0000u)) : ((lo & 0xFFFFu) | (hi << 16));
      *reinterpret_cast<uint32_t*>(y + (static_cast<long long>(b) * C + c) * pitch + t) = v;
    }
  }
}

int ncw_to_tm(const void* x, int64_t x_pitch, void* y, int B, int C, int T, int Bp, cudaStream_t stream) {
  if (x == nullptr || y == nullptr) return fail(V100_E_INVALID, "ncw_to_tm: null pointer");
  if (B <= 0 || C <= 0 || T <= 0 || Bp < B || (Bp & 7) != 0 || x_pitch < T || C > 65535)
    return fail(V100_E_INVALID, "ncw_to_tm: bad sizes (Bp must be a multiple of 8, >= B)");
  if ((x_pitch & 1) != 0) return fail(V100_E_INVALID, "ncw_to_tm: pitch must be even");
  dim3 grid((T + 63) / 64, (Bp + 63) / 64, C);
  ncw_to_tm_kernel<<<grid, 256, 0, stream>>>(static_cast<const unsigned short*>(x), x_pitch,
                                             static_cast<unsigned short*>(y), B, C, T, Bp);
  V100_CUDA(cudaGetLastError());
  return 0;
}

int tm_to_ncw(const void* x, void* y, int64_t y_pitch, int B, int C, int T, int Bp, cudaStream_t stream) {
  if (x == nullptr || y == nullptr) return fail(V100_E_INVALID, "tm_to_ncw: null pointer");
  if (B <= 0 || C <= 0 || T <= 0 || Bp < B || (Bp & 7) != 0 || y_pitch < T || C > 65535)
    return fail(V100_E_INVALID, "tm_to_ncw: bad sizes (Bp must be a multiple of 8, >= B)");
  if ((y_pitch & 1) != 0) return fail(V100_E_INVALID, "tm_to_ncw: pitch must be even");
  dim3 grid((unsigned)((y_pitch + 63) / 64), (B + 63) / 64, C);
  tm_to_ncw_kernel<<<grid, 256, 0, stream>>>(static_cast<const unsigned short*>(x), static_cast<unsigned short*>(y),
                                             y_pitch, B, C, T, Bp);
  V100_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// bidirectional LSTM layer
// ------------------------------------------------------------------------------------------------
constexpr int kLstmUnits = 16;    // hidden units per CTA
constexpr int kLstmN = 64;        // gate columns per CTA = 4 gates x 16 units (UMMA N)
constexpr int kLstmRows = 128;    // utterances per group (UMMA M)
constexpr int kLstmThreads = 160; // warps 0-3: gates (TMEM lane quadrant = warp), warp 4: TMA + MMA issue

struct LstmParams {
  int H, T, B, Bp;
  int slices;        // H / 16
  int groups_total;  // ceil(B / 128)
  int group0;        // first group of this launch
  int groups;        // groups in this launch
  long long n_cols;  // T * Bp
  const int32_t* lengths;
  unsigned short* y;      // [2H][n_cols]
  unsigned short* hx;     // exchange buffer [2 dirs][groups_total][2][H/64][128][64]
  unsigned int* counters; // [2 dirs][groups_total]
  int dtype;
  int cluster;            // CTAs per cluster (consecutive slices of one (direction, group)); 1 = no multicast
};

__device__ __forceinline__ float tanh_fast(float x) {
  float r;
  asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(tanh_fast(0.5f * x), 0.5f, 0.5f); }

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add_u32(unsigned int* p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// generic-proxy global writes <-> async-proxy (TMA) global reads
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// TMA load delivered to the same shared-memory offset (and signalled on the mbarrier at the same offset) of
// every CTA in `mask` of this cluster
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
      "[%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}

#ifdef V100_LSTM_PROF
// profiling build only (tools/lstm_prof.py): globaltimer stamps of block 0 for a few steps
__device__ unsigned long long g_lstm_prof[64 * 12];
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define LSTM_STAMP(step, slot) \
  do { if (blockIdx.x == 0 && (step) >= 100 && (step) < 164) g_lstm_prof[((step) - 100) * 12 + (slot)] = gtime(); } while (0)
#else
#define LSTM_STAMP(step, slot) do {} while (0)
#endif

template <int DT>
__global__ void __launch_bounds__(kLstmThreads, 1)
lstm_layer_kernel(const __grid_constant__ CUtensorMap tm_gx, const __grid_constant__ CUtensorMap tm_w,
                  const __grid_constant__ CUtensorMap tm_h, const LstmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int KB = p.H / 64;
  uint8_t* sA = smem;                                   // h tile: KB x [128 rows x 64 k] (16 KB each)
  uint8_t* sW = sA + KB * 16384;                        // W slice: KB x [64 rows x 64 k] (8 KB each)
  unsigned short* sG = reinterpret_cast<unsigned short*>(sW + KB * 8192);  // 2 x [64 gate rows][128 utterances]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sG) + 2 * kLstmN * kLstmRows * 2);
  uint64_t* h_full = bars;         // [8] one per 64-wide k block of the h tile
  uint64_t* acc_full = bars + 8;   // [1]
  uint64_t* gx_full = bars + 9;    // [2]
  uint64_t* w_full = bars + 11;    // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slice = blockIdx.x % p.slices;
  const int gl = (blockIdx.x / p.slices) % p.groups;  // group within the launch
  const int dir = blockIdx.x / (p.slices * p.groups);
  const int grp = p.group0 + gl;
  const int dg = dir * p.groups_total + grp;
  unsigned int* counter = p.counters + dg;
  const int u0 = slice * kLstmUnits;

  auto load_gx = [&](int k) {  // control thread: the four gate blocks of step k -> buffer k & 1
    const int t = dir ? p.T - 1 - k : k;
    const int buf = k & 1;
    mbar_expect_tx(&gx_full[buf], kLstmN * kLstmRows * 2);
#pragma unroll
    for (int q = 0; q < 4; ++q)
      tma_load_2d(sG + (buf * kLstmN + q * kLstmUnits) * kLstmRows, &tm_gx, &gx_full[buf],
                  t * p.Bp + grp * kLstmRows, dir * 4 * p.H + q * p.H + u0);
  };

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tm_gx);
    tma_prefetch_desc(&tm_w);
    tma_prefetch_desc(&tm_h);
    for (int kb = 0; kb < 8; ++kb) mbar_init(&h_full[kb], 1);
    mbar_init(acc_full, 1);
    mbar_init(&gx_full[0], 1);
    mbar_init(&gx_full[1], 1);
    mbar_init(w_full, 1);
    fence_mbar_init();
  }
  if (warp == 4) {
    tmem_alloc(tmem_slot, kLstmN);
    tmem_relinquish();
  }
  tc_fence_before();
  // peers multicast into this CTA's shared memory and signal its barriers: they must be initialised first
  if (p.cluster > 1) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int crank = p.cluster > 1 ? int(cluster_ctarank()) : 0;
  const uint16_t cmask = static_cast<uint16_t>((1u << p.cluster) - 1u);

  if (warp == 4) {
    // ===================== control warp: TMA + MMA issue =====================
    // lane 0 polls the step counter, waits for the tile and issues the MMAs; lanes 0..KB-1 each issue one
    // 16 KB box of the h tile (one warp instruction instead of a serial loop -- the serial issue of 13 TMA
    // operations by one thread was 0.74 us of every 5.4 us step); lane 8 prefetches the next step's Gx.
    if (lane == 0) {
      mbar_expect_tx(w_full, uint32_t(kLstmN) * p.H * 2);
      for (int kb = 0; kb < KB; ++kb)
#pragma unroll
        for (int q = 0; q < 4; ++q)
          tma_load_2d(sW + kb * 8192 + q * kLstmUnits * 128, &tm_w, w_full, kb * 64,
                      dir * 4 * p.H + q * p.H + u0);
      load_gx(0);
      if (p.T > 1) load_gx(1);
      mbar_wait(w_full, 0);
    }
    const uint32_t fmt = DT == DT_F16 ? 0u : 1u;
    // kind::f16, D = f32, A and B K-major, M = 128, N = 64
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (uint32_t(kLstmN >> 3) << 17) |
                           (uint32_t(kLstmRows >> 4) << 24);
    const uint32_t a_addr = smem_u32(sA), w_addr = smem_u32(sW);
    const unsigned int per_step = p.slices;
    const bool loads_box = lane < KB && (lane % p.cluster) == crank;
    for (int k = 1; k < p.T; ++k) {
      if (lane == 0) {
        // arm the tile barriers before the wait: the bytes can only arrive after it anyway
        for (int kb = 0; kb < KB; ++kb) mbar_expect_tx(&h_full[kb], 16384);
        // every slice of this (direction, group) has published h of step k-1
        const unsigned int need = unsigned(k) * per_step;
        LSTM_STAMP(k, 0);
        if (ld_acquire_u32(counter) < need) {
          const long long t_start = clock64();
          while (ld_acquire_u32(counter) < need) {
            if (clock64() - t_start > 8000000000LL) {
              printf("libv100: lstm step wait timed out (block %d step %d have %u need %u)\n", blockIdx.x, k,
                     ld_acquire_u32(counter), need);
              __trap();
            }
          }
        }
        LSTM_STAMP(k, 1);
      }
      __syncwarp();
      // h_{k-1} of the group.  (With clusters the CTAs are slices of the same (direction, group) and all passed
      // the same counter, so each loads 1/cluster of the tile and multicasts it to every peer.)
      if (loads_box) {
        fence_proxy_async_global();
        // exchange buffer = [dir][group][parity][k block][128 rows][64]: one box is 16 KB contiguous in global memory
        const int hrow = ((dg * 2 + ((k - 1) & 1)) * KB + lane) * kLstmRows;
        if (p.cluster > 1) tma_load_2d_mc(sA + lane * 16384, &tm_h, &h_full[lane], 0, hrow, cmask);
        else tma_load_2d(sA + lane * 16384, &tm_h, &h_full[lane], 0, hrow);
      }
      // the gate warps of this CTA are past step k-1 as well: their Gx buffer (k+1)&1 is free again
      if (lane == 8 && k + 1 < p.T) load_gx(k + 1);
      if (lane == 0) {
        LSTM_STAMP(k, 8);
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&h_full[kb], (k - 1) & 1);
          if (kb == 0) LSTM_STAMP(k, 9);
          if (kb == KB - 1) LSTM_STAMP(k, 10);
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t da = umma_desc(a_addr + kb * 16384 + kk * 32, 16, 1024);
            const uint64_t db = umma_desc(w_addr + kb * 8192 + kk * 32, 16, 1024);
            umma_bf16(tmem_base, da, db, idesc, (kb | kk) ? 1u : 0u);
          }
        }
        umma_commit(acc_full);
        LSTM_STAMP(k, 2);
      }
      __syncwarp();
    }
  } else {
    // ===================== gates: one thread = one utterance, 16 hidden units =====================
    const int row = warp * 32 + lane;
    const int b = grp * kLstmRows + row;
    const int len = b < p.B ? __ldg(p.lengths + b) : 0;
    const uint32_t lane_addr = tmem_base + (uint32_t(warp * 32) << 16);
    float c[kLstmUnits];
#pragma unroll
    for (int u = 0; u < kLstmUnits; ++u) c[u] = 0.0f;
    for (int k = 0; k < p.T; ++k) {
      const int t = dir ? p.T - 1 - k : k;
      uint32_t acc[kLstmN];
      if (k > 0) {
        mbar_wait(acc_full, (k - 1) & 1);
        tc_fence_after();
        tmem_ld32(lane_addr, *reinterpret_cast<uint32_t(*)[32]>(acc));
        tmem_ld32(lane_addr + 32, *reinterpret_cast<uint32_t(*)[32]>(acc + 32));
        tmem_ld_wait();
        tc_fence_before();
        if (threadIdx.x == 0) LSTM_STAMP(k, 3);
      } else {
#pragma unroll
        for (int i = 0; i < kLstmN; ++i) acc[i] = 0u;
      }
      mbar_wait(&gx_full[k & 1], (k >> 1) & 1);
      const unsigned short* g = sG + (k & 1) * kLstmN * kLstmRows + row;
      const bool live = t < len;
      uint32_t hw[kLstmUnits / 2];
      float hprev = 0.0f;
#pragma unroll
      for (int u = 0; u < kLstmUnits; ++u) {
        const float ai = __uint_as_float(acc[u]) + h2f<DT>(g[(0 * kLstmUnits + u) * kLstmRows]);
        const float af = __uint_as_float(acc[kLstmUnits + u]) + h2f<DT>(g[(1 * kLstmUnits + u) * kLstmRows]);
        const float ag = __uint_as_float(acc[2 * kLstmUnits + u]) + h2f<DT>(g[(2 * kLstmUnits + u) * kLstmRows]);
        const float ao = __uint_as_float(acc[3 * kLstmUnits + u]) + h2f<DT>(g[(3 * kLstmUnits + u) * kLstmRows]);
        const float cn = fmaf(sigmoid_fast(af), c[u], sigmoid_fast(ai) * tanh_fast(ag));
        const float hn = sigmoid_fast(ao) * tanh_fast(cn);
        c[u] = live ? cn : 0.0f;
        const float h = live ? hn : 0.0f;
        if (u & 1) hw[u >> 1] = pack2<DT>(hprev, h);
        else hprev = h;
      }
      // h_t for the next step's MMA: exchange buffer parity k&1, row = utterance, 16 units = 32 bytes
      uint4* hdst = reinterpret_cast<uint4*>(
          p.hx + (static_cast<long long>((dg * 2 + (k & 1)) * KB + (u0 >> 6)) * kLstmRows + row) * 64 + (u0 & 63));
      hdst[0] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
      hdst[1] = make_uint4(hw[4], hw[5], hw[6], hw[7]);
      // publish: every thread orders its stores against the async proxy (the consumers read them with TMA),
      // the barrier collects the four warps, one release-add makes the whole tile visible (cumulativity)
      if (threadIdx.x == 0) LSTM_STAMP(k, 4);
      fence_proxy_async_global();
      if (threadIdx.x == 0) LSTM_STAMP(k, 5);
      named_bar_sync(1, 128);
      if (threadIdx.x == 0) LSTM_STAMP(k, 6);
      if (threadIdx.x == 0) red_release_add_u32(counter, 1u);
      if (threadIdx.x == 0) LSTM_STAMP(k, 7);
      // layer output (off the critical path): y[dir*H + u0 + u][t*Bp + b]
      if (b < p.Bp) {
        unsigned short* yp = p.y + (static_cast<long long>(dir) * p.H + u0) * p.n_cols + static_cast<long long>(t) * p.Bp + b;
#pragma unroll
        for (int u = 0; u < kLstmUnits; ++u)
          yp[u * p.n_cols] = static_cast<unsigned short>((u & 1) ? (hw[u >> 1] >> 16) : (hw[u >> 1] & 0xFFFFu));
      }
    }
  }

  tc_fence_before();
  // no CTA may exit while a peer can still multicast into its shared memory
  if (p.cluster > 1) cluster_sync_all(); else __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kLstmN);
  }
}

#ifdef V100_LSTM_PROF
extern "C" int v100_debug_lstm_prof(unsigned long long* out) {
  return static_cast<int>(cudaMemcpyFromSymbol(out, g_lstm_prof, sizeof(g_lstm_prof)));
}
#endif

size_t lstm_workspace_bytes(int B, int H) {
  const size_t groups = (size_t(B) + kLstmRows - 1) / kLstmRows;
  return 2 * groups * 2 * kLstmRows * size_t(H) * 2 + 256 + 2 * groups * sizeof(unsigned int);
}

int lstm_layer(const void* gx, const void* w_hh, const int32_t* lengths, void* y, void* workspace, int B, int Bp,
               int T, int H, int dtype, cudaStream_t stream) {
  if (gx == nullptr || w_hh == nullptr || lengths == nullptr || y == nullptr || workspace == nullptr)
    return fail(V100_E_INVALID, "lstm_layer: null pointer");
  if (dtype != DT_BF16 && dtype != DT_F16) return fail(V100_E_INVALID, "lstm_layer: bad dtype");
  if (B <= 0 || T <= 0 || Bp < B || (Bp & 7) != 0) return fail(V100_E_INVALID, "lstm_layer: bad sizes (Bp must be a multiple of 8, >= B)");
  if (H < 64 || H % 64 != 0 || H > 512) return fail(V100_E_UNSUPPORTED, "lstm_layer: hidden size %d (supported: multiples of 64 up to 512)", H);
  if ((reinterpret_cast<uintptr_t>(workspace) & 1023) != 0) return fail(V100_E_INVALID, "lstm_layer: workspace must be 1024-byte aligned");
  if (static_cast<long long>(T) * Bp > 2147483647LL - 256) return fail(V100_E_UNSUPPORTED, "lstm_layer: T*Bp too large");
  const CUtensorMapDataType tt = dtype == DT_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  LstmParams p{};
  p.H = H; p.T = T; p.B = B; p.Bp = Bp;
  p.slices = H / kLstmUnits;
  p.groups_total = (B + kLstmRows - 1) / kLstmRows;
  p.n_cols = static_cast<long long>(T) * Bp;
  p.lengths = lengths;
  p.y = static_cast<unsigned short*>(y);
  p.hx = static_cast<unsigned short*>(workspace);
  const size_t hx_bytes = size_t(2) * p.groups_total * 2 * kLstmRows * H * 2;
  p.counters = reinterpret_cast<unsigned int*>(static_cast<uint8_t*>(workspace) + hx_bytes);
  p.dtype = dtype;
  V100_CUDA(cudaMemsetAsync(p.counters, 0, 2 * p.groups_total * sizeof(unsigned int), stream));
  CUtensorMap tm_gx, tm_w, tm_h;
  if (int e = make_tmap_2d_plain(&tm_gx, tt, gx, p.n_cols, int64_t(8) * H, p.n_cols * 2, kLstmRows, kLstmUnits)) return e;
  if (int e = make_tmap_2d(&tm_w, tt, w_hh, H, int64_t(8) * H, int64_t(H) * 2, 64, kLstmUnits)) return e;
  if (int e = make_tmap_2d(&tm_h, tt, workspace, 64, int64_t(2) * p.groups_total * 2 * (H / 64) * kLstmRows, 128, 64, kLstmRows)) return e;
  const int KB = H / 64;
  const size_t smem = 1024 + size_t(KB) * (16384 + 8192) + 2 * kLstmN * kLstmRows * 2 + 128;
  auto kern = dtype == DT_F16 ? lstm_layer_kernel<DT_F16> : lstm_layer_kernel<DT_BF16>;
  V100_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  // all CTAs of a launch must be co-resident: at most floor(SMs / (2 * slices)) groups per launch
  const int max_groups = num_sms() / (2 * p.slices);
  if (max_groups < 1) return fail(V100_E_UNSUPPORTED, "lstm_layer: device too small for H=%d", H);
  static const int force_cluster = getenv("V100_LSTM_CLUSTER") ? atoi(getenv("V100_LSTM_CLUSTER")) : 0;  // A/B runs
  for (int g0 = 0; g0 < p.groups_total; g0 += max_groups) {
    p.group0 = g0;
    p.groups = p.groups_total - g0 < max_groups ? p.groups_total - g0 : max_groups;
    const int grid = 2 * p.groups * p.slices;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kLstmThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // V100_LSTM_CLUSTER=n: the n CTAs of a cluster each load 1/n of the h tile and multicast it to their peers
    // (largest n <= k blocks, <= 8, whose clusters are all co-resident).  Measured on B200 (256 x 751 steps,
    // H = 512): 3.99 ms per layer without clusters, 4.03 / 4.11 / 4.12 ms with n = 2 / 4 / 8 -- the step is bound
    // by each SM's own ingest of the 128 KB tile, not by L2 read traffic, so the default is no cluster.
    int cs = 1;
    if (force_cluster > 1) cs = force_cluster < KB ? force_cluster : KB;
    if (cs > 8) cs = 8;
    while (cs & (cs - 1)) --cs;
    for (; cs > 1; cs >>= 1) {
      attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      int active = 0;
      if (cudaOccupancyMaxActiveClusters(&active, kern, &cfg) == cudaSuccess && active * cs >= grid) break;
      (void)cudaGetLastError();
    }
    attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    p.cluster = cs;
    V100_CUDA(cudaLaunchKernelEx(&cfg, kern, tm_gx, tm_w, tm_h, p));
  }
  return 0;
}

}  // namespace v100
