// Fused log-mel front end: framing (reflect padding at the clip's own ends) + periodic Hann(400) in a
// 512 frame + 512-point real FFT + |.|^2 + sparse 64-filter HTK mel bank + log(. + offset), one kernel.
//
// Replaces reflect-pad + unfold + window + cuFFT R2C + abs/pow + sgemm + add/log + transpose
// (>= 6 library launches, and a 257-bin complex spectrum written to and read back from HBM).
// One warp transforms one frame (radix-4 Stockham FFT in shared memory, see logmel_core.h); a CTA of
// 8 warps produces a [64 mel x 64 frame] tile so that the NCW output rows are written as full
// 128-byte lines.  The kernel is bound by FFT arithmetic and shared-memory traffic, not by HBM
// (64 KB in + 12.8 KB out per audio-second); see DESIGN.md for its roofline discussion.
#include "common.cuh"
#include "host.h"
#include "logmel_core.h"

namespace v100 {

constexpr int kMelWarps = 8;
constexpr int kMelFrames = 64;  // frames per CTA
constexpr int kNMels = 64;
constexpr int kNFft = 512, kWin = 400, kHop = 160, kWinLeft = (kNFft - kWin) / 2;  // data_modules.py:266-269
constexpr int kMelMaxTaps = 24;   // longest mel filter (bins); the HTK bank at 512/16 kHz/64 needs 20
constexpr int kMelSmemBytes = 512 * 8 + 2 * kMelWarps * 256 * 8 + kNMels * (kMelFrames + 1) * 4 + kWin * 4 +
                              kMelMaxTaps * kNMels * 4;

__global__ void __launch_bounds__(kMelWarps * 32, 3)
logmel_kernel(const float* __restrict__ wav, const int32_t* __restrict__ len, long long wav_pitch,
              const int32_t* __restrict__ fb_start, const int32_t* __restrict__ fb_count,
              const int32_t* __restrict__ fb_off, const float* __restrict__ fb_w, float log_offset, void* out, int T,
              long long out_pitch, int out_mode) {
  extern __shared__ __align__(16) uint8_t mel_smem[];
  cpx* tw = reinterpret_cast<cpx*>(mel_smem);                                  // [512]
  cpx (*bufA)[256] = reinterpret_cast<cpx (*)[256]>(tw + 512);                  // [kMelWarps][256]
  cpx (*bufB)[256] = bufA + kMelWarps;                                          // [kMelWarps][256]
  float (*tile)[kMelFrames + 1] = reinterpret_cast<float (*)[kMelFrames + 1]>(bufB + kMelWarps);  // [64][65]
  float* win = reinterpret_cast<float*>(tile + kNMels);                        // [kWin]
  float (*melw)[kNMels] = reinterpret_cast<float (*)[kNMels]>(win + kWin);      // [kMelMaxTaps][64] tap-major

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int f0 = blockIdx.x * kMelFrames;

  for (int k = threadIdx.x; k < 512; k += blockDim.x) {
    float s, c;
    sincospif(-float(k) / 256.0f, &s, &c);  // exp(-2*pi*i*k/512)
    tw[k] = cpx{c, s};
  }
  for (int i = threadIdx.x; i < kWin; i += blockDim.x) win[i] = 0.5f - 0.5f * cospif(float(2 * i) / float(kWin));
  for (int i = threadIdx.x; i < kMelMaxTaps * kNMels; i += blockDim.x) {   // melw[tap][filter], zero padded
    const int tap = i / kNMels, m = i - tap * kNMels;
    melw[tap][m] = tap < __ldg(fb_count + m) ? __ldg(fb_w + __ldg(fb_off + m) + tap) : 0.0f;
  }
  __syncthreads();

  // ---- per-lane constants, reused for the 8 frames this warp transforms ----
  const tw3 t1a = fft256_twiddles(1, lane, tw), t1b = fft256_twiddles(1, lane + 32, tw);
  const tw3 t2a = fft256_twiddles(4, lane, tw), t2b = fft256_twiddles(4, lane + 32, tw);
  const cpx wlane = tw[lane];                  // exp(-2*pi*i*(lane + 32 i)/512) = wlane * tw[32 i]
  float wv[16];                                // window taps of the samples this lane loads (0 outside 56..455)
#pragma unroll
  for (int u = 0; u < 8; ++u)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int m = 2 * (lane + 32 * u) + h;
      wv[2 * u + h] = (m >= kWinLeft && m < kWinLeft + kWin) ? win[m - kWinLeft] : 0.0f;
    }
  const int s0a = __ldg(fb_start + lane), s0b = __ldg(fb_start + lane + 32);
  const int cnta = __ldg(fb_count + lane), cntb = __ldg(fb_count + lane + 32);
  // filters 0..31 are short (<= 6 bins), filters 32..63 long (<= 20): separate trip counts
  const int cnt_max_a = __reduce_max_sync(0xffffffffu, cnta), cnt_max_b = __reduce_max_sync(0xffffffffu, cntb);

  const int L = len[b];
  const int n_frames = 1 + L / kHop;
  const float* x = wav + static_cast<long long>(b) * wav_pitch;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(x) & 7) == 0);
  const float blank = out_mode == V100_MEL_POWER_F32_NCW ? 0.0f : logf(log_offset);
  cpx* A = bufA[warp];
  cpx* Bf = bufB[warp];
  float* P = reinterpret_cast<float*>(Bf);  // power spectrum reuses the ping-pong buffer (257 <= 512 floats)

  for (int fi = 0; fi < kMelFrames / kMelWarps; ++fi) {
    const int fl = warp * (kMelFrames / kMelWarps) + fi;
    const int t = f0 + fl;
    if (t >= n_frames) {
      tile[lane][fl] = blank;
      tile[lane + 32][fl] = blank;
      continue;
    }
    // frame t covers reflect-padded samples [160 t, 160 t + 512) = clip samples 160 t - 256 + m
    const int base = kHop * t - kNFft / 2;
    if (vec_ok && base + kWinLeft >= 0 && base + kWinLeft + kWin <= L) {
      // interior frame: the 400 windowed samples are in range, 8-byte aligned pairs
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int n = lane + 32 * u;
        float2 v = make_float2(0.0f, 0.0f);
        if (2 * n >= kWinLeft && 2 * n < kWinLeft + kWin) v = __ldg(reinterpret_cast<const float2*>(x + base + 2 * n));
        A[fswz(n)] = cpx{v.x * wv[2 * u], v.y * wv[2 * u + 1]};
      }
    } else {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int n = lane + 32 * u;
        float v[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int m = 2 * n + h;
          float sv = 0.0f;
          if (m >= kWinLeft && m < kWinLeft + kWin) {
            int i = base + m;
            i = i < 0 ? -i : i;
            i = i >= L ? 2 * (L - 1) - i : i;
            sv = __ldg(x + i) * wv[2 * u + h];
          }
          v[h] = sv;
        }
        A[fswz(n)] = cpx{v[0], v[1]};
      }
    }
    __syncwarp();
    fft256_butterfly(A, Bf, 256, 1, lane, t1a);   fft256_butterfly(A, Bf, 256, 1, lane + 32, t1b);   __syncwarp();
    fft256_butterfly(Bf, A, 64, 4, lane, t2a);    fft256_butterfly(Bf, A, 64, 4, lane + 32, t2b);    __syncwarp();
    // pass 3 has only 4 distinct twiddle sets per warp: broadcast shared loads instead of 12 more registers
    fft256_butterfly(A, Bf, 16, 16, lane, fft256_twiddles(16, lane, tw));
    fft256_butterfly(A, Bf, 16, 16, lane + 32, fft256_twiddles(16, lane + 32, tw));
    __syncwarp();
    fft256_butterfly_last(Bf, A, lane);           fft256_butterfly_last(Bf, A, lane + 32);           __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      // exp(-2*pi*i*(32 i)/512) = exp(-i*pi*i/8): compile-time constants
      float ci, si;
      sincospif(-float(i) / 8.0f, &si, &ci);
      P[lane + 32 * i] = rfft512_power(A, lane + 32 * i, cmul(wlane, cpx{ci, si}));
    }
    if (lane == 0) P[256] = rfft512_power(A, 256, cpx{-1.0f, 0.0f});
    __syncwarp();
    float acca = 0.0f, accb = 0.0f;   // melw is zero past each filter's own length; P stays in range
    for (int i = 0; i < cnt_max_a; ++i) acca = fmaf(melw[i][lane], P[min(s0a + i, 256)], acca);
    for (int i = 0; i < cnt_max_b; ++i) accb = fmaf(melw[i][lane + 32], P[min(s0b + i, 256)], accb);
    tile[lane][fl] = out_mode == V100_MEL_POWER_F32_NCW ? acca : logf(acca + log_offset);
    tile[lane + 32][fl] = out_mode == V100_MEL_POWER_F32_NCW ? accb : logf(accb + log_offset);
    __syncwarp();
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
