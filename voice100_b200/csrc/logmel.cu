// Fused log-mel front end: framing (reflect padding at the clip's own ends) + periodic Hann(400) in a
// 512 frame + 512-point real FFT + |.|^2 + sparse 64-filter HTK mel bank + log(. + offset), one kernel.
//
// Replaces reflect-pad + unfold + window + cuFFT R2C + abs/pow + sgemm + add/log + transpose
// (>= 6 library launches, and a 257-bin complex spectrum written to and read back from HBM).
// Sixteen threads transform one frame: 256 = 16 x 16, two 16-point FFTs in registers with one transpose
// through shared memory in between (logmel_core.h); a CTA of 8 warps produces a [64 mel x 64 frame] tile so
// that the NCW output rows are written as full 128-byte lines.  The kernel is bound by FFT arithmetic and
// shared-memory traffic, not by HBM (64 KB in + 12.8 KB out per audio-second); see DESIGN.md.
#include "common.cuh"
#include "host.h"
#include "logmel_core.h"

namespace v100 {

constexpr int kMelWarps = 8;
constexpr int kMelFrames = 64;  // frames per CTA
constexpr int kNMels = 64;
constexpr int kNFft = 512, kWin = 400, kHop = 160, kWinLeft = (kNFft - kWin) / 2;  // data_modules.py:266-269
constexpr int kMelMaxTaps = 24;   // longest mel filter (bins); the HTK bank at 512/16 kHz/64 needs 20
// shared memory: twiddles, window pairs, mel weights, output tile, and per half-warp an exchange buffer (16 x 17
// complex, reused for the 257 power bins) and the packed spectrum Z (256 complex)
constexpr int kMelExch = 16 * 17 + 8;  // + 8: consecutive half-warps' buffers start 16 banks apart, so the 32-bit
                                       // power-spectrum accesses of the two frames of a warp do not collide
constexpr int kMelSmemBytes = 512 * 8 + 256 * 8 + kMelMaxTaps * kNMels * 4 + kNMels * (kMelFrames + 1) * 4 +
                              2 * kMelWarps * (kMelExch + 256) * 8;

__global__ void __launch_bounds__(kMelWarps * 32, 2)
logmel_kernel(const float* __restrict__ wav, const int32_t* __restrict__ len, long long wav_pitch,
              const int32_t* __restrict__ fb_start, const int32_t* __restrict__ fb_count,
              const int32_t* __restrict__ fb_off, const float* __restrict__ fb_w, float log_offset, void* out, int T,
              long long out_pitch, int out_mode) {
  extern __shared__ __align__(16) uint8_t mel_smem[];
  cpx* tw = reinterpret_cast<cpx*>(mel_smem);                                  // [512] exp(-2 pi i k / 512)
  float2* win2 = reinterpret_cast<float2*>(tw + 512);                          // [256] window of samples 2n, 2n+1
  float (*melw)[kNMels] = reinterpret_cast<float (*)[kNMels]>(win2 + 256);      // [kMelMaxTaps][64] tap-major
  float (*tile)[kMelFrames + 1] = reinterpret_cast<float (*)[kMelFrames + 1]>(melw + kMelMaxTaps);  // [64][65]
  cpx* exch = reinterpret_cast<cpx*>(tile + kNMels);                           // [16 half-warps][16 x 17]
  cpx* zbuf = exch + 2 * kMelWarps * kMelExch;                                 // [16 half-warps][256]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int f0 = blockIdx.x * kMelFrames;

  for (int k = threadIdx.x; k < 512; k += blockDim.x) {
    float sn, cs;
    sincospif(-float(k) / 256.0f, &sn, &cs);  // exp(-2*pi*i*k/512)
    tw[k] = cpx{cs, sn};
  }
  for (int n = threadIdx.x; n < 256; n += blockDim.x) {   // periodic Hann(400) centred in the 512 frame
    float w2[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int m = 2 * n + h - kWinLeft;
      w2[h] = (m >= 0 && m < kWin) ? 0.5f - 0.5f * cospif(float(2 * m) / float(kWin)) : 0.0f;
    }
    win2[n] = make_float2(w2[0], w2[1]);
  }
  for (int i = threadIdx.x; i < kMelMaxTaps * kNMels; i += blockDim.x) {   // melw[tap][filter], zero padded
    const int tap = i / kNMels, m = i - tap * kNMels;
    melw[tap][m] = tap < __ldg(fb_count + m) ? __ldg(fb_w + __ldg(fb_off + m) + tap) : 0.0f;
  }
  __syncthreads();

  // ---- sixteen threads per frame (logmel_core.h, "register-resident variant"); a warp works on two frames ----
  const int q = lane & 15, hw = warp * 2 + (lane >> 4);
  cpx* E = exch + hw * kMelExch;
  cpx* Z = zbuf + hw * 256;
  float* P = reinterpret_cast<float*>(E);  // the power spectrum reuses the exchange buffer (257 <= 544 floats)
  cpx twq[16];                             // W256^(q k1), constant per thread
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) twq[k1] = tw[(2 * q * k1) & 511];
  const cpx wq = tw[q];                    // exp(-2 pi i (q + 16 k2) / 512) = wq * exp(-2 pi i k2 / 32)
  int s0[4], cmax[4];                      // this thread's mel filters q + 16 i
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    s0[i] = __ldg(fb_start + q + 16 * i);
    int c = __ldg(fb_count + q + 16 * i);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) c = max(c, __shfl_xor_sync(0xffffffffu, c, o));
    cmax[i] = c;                           // trip count of the 16 filters handled together
  }

  const int L = len[b];
  const int n_frames = 1 + L / kHop;
  const float* x = wav + static_cast<long long>(b) * wav_pitch;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(x) & 7) == 0);
  const float blank = out_mode == V100_MEL_POWER_F32_NCW ? 0.0f : logf(log_offset);

  for (int fi = 0; fi < kMelFrames / (2 * kMelWarps); ++fi) {
    // the two half-warps of a warp take frames 16 apart: their tile[mel][frame] stores land 16 banks apart
    const int fl = 16 * (2 * (fi >> 1) + (hw & 1)) + 2 * (hw >> 1) + (fi & 1);
    const int t = f0 + fl;
    const bool valid = t < n_frames;
    if (!__any_sync(0xffffffffu, valid)) {
#pragma unroll
      for (int i = 0; i < 4; ++i) tile[q + 16 * i][fl] = blank;
      continue;
    }
    // frame t covers reflect-padded samples [160 t, 160 t + 512) = clip samples 160 t - 256 + m;
    // this thread takes the complex points n = q + 16 r (samples 2n, 2n+1); the window is zero for n < 28, n >= 228
    cpx v[16];
    const int base = kHop * t - kNFft / 2;
    v[0] = cpx{0.0f, 0.0f};
    v[15] = cpx{0.0f, 0.0f};
    if (valid && vec_ok && base + kWinLeft >= 0 && base + kWinLeft + kWin <= L) {
#pragma unroll
      for (int r = 1; r < 15; ++r) {
        const int n = q + 16 * r;
        float2 sv = make_float2(0.0f, 0.0f);
        if (n >= kWinLeft / 2 && n < (kWinLeft + kWin) / 2) sv = __ldg(reinterpret_cast<const float2*>(x + base + 2 * n));
        const float2 w2 = win2[n];
        v[r] = cpx{sv.x * w2.x, sv.y * w2.y};
      }
    } else {
#pragma unroll
      for (int r = 1; r < 15; ++r) {
        const int n = q + 16 * r;
        const float2 w2 = win2[n];
        float sv[2] = {0.0f, 0.0f};
        if (valid && n >= kWinLeft / 2 && n < (kWinLeft + kWin) / 2) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            int i = base + 2 * n + h;
            i = i < 0 ? -i : i;
            i = i >= L ? 2 * (L - 1) - i : i;
            sv[h] = __ldg(x + i);
          }
        }
        v[r] = cpx{sv[0] * w2.x, sv[1] * w2.y};
      }
    }
    fft16(v);                                              // over r:  Y[q][k1]
    E[q] = v[0];
#pragma unroll
    for (int k1 = 1; k1 < 16; ++k1) E[k1 * 17 + q] = cmul(v[k1], twq[k1]);
    __syncwarp();
#pragma unroll
    for (int qq = 0; qq < 16; ++qq) v[qq] = E[q * 17 + qq]; // this thread is now k1 = q
    fft16(v);                                              // over q:  v[k2] = Z[q + 16 k2]
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) Z[q + 16 * k2] = v[k2];
    __syncwarp();                                          // (every lane is past its reads of E: P may overwrite it)
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) {
      float ck, sk;
      sincospif(-float(k2) / 16.0f, &sk, &ck);             // exp(-2 pi i 16 k2 / 512): compile-time constants
      const int k = q + 16 * k2;
      P[k] = rfft512_power_pair(v[k2], Z[(256 - k) & 255], cmul(wq, cpx{ck, sk}));
    }
    if (q == 0) P[256] = rfft512_power_pair(v[0], v[0], cpx{-1.0f, 0.0f});
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = q + 16 * i;
      float acc = 0.0f;                                    // melw is zero past each filter's own length; P stays in range
      for (int tap = 0; tap < cmax[i]; ++tap) acc = fmaf(melw[tap][m], P[min(s0[i] + tap, 256)], acc);
      tile[m][fl] = !valid ? blank : (out_mode == V100_MEL_POWER_F32_NCW ? acc : logf(acc + log_offset));
    }
    __syncwarp();                                          // P (= E) is free for the next frame
  }
  __syncthreads();

  const int tid = threadIdx.x;
  if (out_mode == V100_MEL_LOG_BF16_NCW || out_mode == V100_MEL_LOG_F16_NCW) {
    const int m = tid >> 2, fs = (tid & 3) * 16;
    unsigned short* row = static_cast<unsigned short*>(out) + (static_cast<long long>(b) * kNMels + m) * out_pitch;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int t = f0 + fs + 8 * h;
      if (t < out_pitch) {
        const float* s = &tile[m][fs + 8 * h];
        uint4 v;
        if (out_mode == V100_MEL_LOG_F16_NCW) {
          v.x = pack_f16x2(s[0], s[1]); v.y = pack_f16x2(s[2], s[3]);
          v.z = pack_f16x2(s[4], s[5]); v.w = pack_f16x2(s[6], s[7]);
        } else {
          v.x = pack_bf16x2(s[0], s[1]); v.y = pack_bf16x2(s[2], s[3]);
          v.z = pack_bf16x2(s[4], s[5]); v.w = pack_bf16x2(s[6], s[7]);
        }
        *reinterpret_cast<uint4*>(row + t) = v;
      }
    }
  } else if (out_mode == V100_MEL_POWER_F32_NCW) {
    const int m = tid >> 2, fs = (tid & 3) * 16;
    float* row = static_cast<float*>(out) + (static_cast<long long>(b) * kNMels + m) * out_pitch;
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const int t = f0 + fs + 4 * h;
      if (t < out_pitch) {
        const float* s = &tile[m][fs + 4 * h];
        *reinterpret_cast<float4*>(row + t) = make_float4(s[0], s[1], s[2], s[3]);
      }
    }
  } else {  // V100_MEL_LOG_F32_NTC: out[b][t][64]
    const int fl = tid >> 2, ms = (tid & 3) * 16;
    const int t = f0 + fl;
    if (t < T) {
      float* row = static_cast<float*>(out) + (static_cast<long long>(b) * T + t) * kNMels + ms;
#pragma unroll
      for (int h = 0; h < 4; ++h)
        *reinterpret_cast<float4*>(row + 4 * h) =
            make_float4(tile[ms + 4 * h][fl], tile[ms + 4 * h + 1][fl], tile[ms + 4 * h + 2][fl], tile[ms + 4 * h + 3][fl]);
    }
  }
}

int logmel(const float* wav, const int32_t* len, int B, int64_t wav_pitch, const int32_t* fb_start,
           const int32_t* fb_count, const int32_t* fb_off, const float* fb_w, float log_offset, void* out, int T,
           int64_t out_pitch, int out_mode, cudaStream_t stream) {
  if (wav == nullptr || len == nullptr || out == nullptr || fb_start == nullptr || fb_count == nullptr ||
      fb_off == nullptr || fb_w == nullptr)
    return fail(V100_E_INVALID, "logmel: null pointer");
  if (B <= 0 || T <= 0 || B > 65535) return fail(V100_E_INVALID, "logmel: bad B=%d or T=%d", B, T);
  if (out_mode == V100_MEL_LOG_BF16_NCW || out_mode == V100_MEL_LOG_F16_NCW) {
    if (out_pitch < T || (out_pitch & 7) || (reinterpret_cast<uintptr_t>(out) & 15))
      return fail(V100_E_INVALID, "logmel: 16-bit NCW pitch must be >= T and a multiple of 8, base 16B aligned");
  } else if (out_mode == V100_MEL_POWER_F32_NCW) {
    if (out_pitch < T || (out_pitch & 3) || (reinterpret_cast<uintptr_t>(out) & 15))
      return fail(V100_E_INVALID, "logmel: fp32 NCW pitch must be >= T and a multiple of 4, base 16B aligned");
  } else if (out_mode == V100_MEL_LOG_F32_NTC) {
    if (reinterpret_cast<uintptr_t>(out) & 15) return fail(V100_E_INVALID, "logmel: output base must be 16B aligned");
  } else {
    return fail(V100_E_INVALID, "logmel: unknown out_mode %d", out_mode);
  }
  static thread_local int configured_dev = -1;
  int dev = 0;
  V100_CUDA(cudaGetDevice(&dev));
  if (configured_dev != dev) {
    V100_CUDA(cudaFuncSetAttribute(logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMelSmemBytes));
    configured_dev = dev;
  }
  dim3 grid((T + kMelFrames - 1) / kMelFrames, B);
  logmel_kernel<<<grid, kMelWarps * 32, kMelSmemBytes, stream>>>(wav, len, wav_pitch, fb_start, fb_count, fb_off, fb_w,
                                                     log_offset, out, T, out_pitch, out_mode);
  V100_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace v100
