// Kernels of the v2 models (voice100/models/_layers_v2.py, _asr_v2.py, _align_v2.py, _tts_v2.py):
//   * dense Conv1d (k = 3/5, stride 1/2) = tap-stacking copy + the tcgen05 GEMM of conv_gemm.cu
//   * LayerNorm over channels + exact GELU on the NCW layout
//   * NCW <-> time-major layout changes around the recurrent layers
//   (the recurrent layer itself is in lstm.cu)
//
// Time-major layout ("TM"): x[c][t * Bp + b], Bp = batch rounded up to 8.  One LSTM step then touches one
// contiguous run of columns, and the input projection W_ih x of ALL steps is one plain 1x1-conv GEMM over
// T*Bp columns (v100_conv1x1 with B = 1).
//
#include "common.cuh"
#include "host.h"


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

}  // namespace v100
