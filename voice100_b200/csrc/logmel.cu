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
constexpr int kMelSmemBytes = 512 * 8 + 2 * kMelWarps * 256 * 8 + kNMels * (kMelFrames + 1) * 4 + kWin * 4;

__global__ void __launch_bounds__(kMelWarps * 32)
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

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int f0 = blockIdx.x * kMelFrames;

  for (int k = threadIdx.x; k < 512; k += blockDim.x) {
    float s, c;
    sincospif(-float(k) / 256.0f, &s, &c);  // exp(-2*pi*i*k/512)
    tw[k] = cpx{c, s};
  }
  for (int i = threadIdx.x; i < kWin; i += blockDim.x) win[i] = 0.5f - 0.5f * cospif(float(2 * i) / float(kWin));
  __syncthreads();

  const int L = len[b];
  const int n_frames = 1 + L / kHop;
  const float* x = wav + static_cast<long long>(b) * wav_pitch;
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
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int n = lane + 32 * u;
      float v[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int m = 2 * n + h;
        float s = 0.0f;
        if (m >= kWinLeft && m < kWinLeft + kWin) {
          int i = base + m;
          i = i < 0 ? -i : i;
          i = i >= L ? 2 * (L - 1) - i : i;
          s = __ldg(x + i) * win[m - kWinLeft];
        }
        v[h] = s;
      }
      A[n] = cpx{v[0], v[1]};
    }
    __syncwarp();
    fft256_butterfly(A, Bf, 256, 1, lane, tw);  fft256_butterfly(A, Bf, 256, 1, lane + 32, tw);  __syncwarp();
    fft256_butterfly(Bf, A, 64, 4, lane, tw);   fft256_butterfly(Bf, A, 64, 4, lane + 32, tw);   __syncwarp();
    fft256_butterfly(A, Bf, 16, 16, lane, tw);  fft256_butterfly(A, Bf, 16, 16, lane + 32, tw);  __syncwarp();
    fft256_butterfly(Bf, A, 4, 64, lane, tw);   fft256_butterfly(Bf, A, 4, 64, lane + 32, tw);   __syncwarp();
    for (int k = lane; k <= 256; k += 32) P[k] = rfft512_power(A, k, tw);
    __syncwarp();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int m = lane + 32 * h;
      const int s0 = __ldg(fb_start + m), cnt = __ldg(fb_count + m), off = __ldg(fb_off + m);
      float acc = 0.0f;
      for (int i = 0; i < cnt; ++i) acc = fmaf(__ldg(fb_w + off + i), P[s0 + i], acc);
      tile[m][fl] = out_mode == V100_MEL_POWER_F32_NCW ? acc : logf(acc + log_offset);
    }
    __syncwarp();
  }
  __syncthreads();

  const int tid = threadIdx.x;
  if (out_mode == V100_MEL_LOG_BF16_NCW) {
    const int m = tid >> 2, fs = (tid & 3) * 16;
    __nv_bfloat16* row = static_cast<__nv_bfloat16*>(out) + (static_cast<long long>(b) * kNMels + m) * out_pitch;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int t = f0 + fs + 8 * h;
      if (t < out_pitch) {
        const float* s = &tile[m][fs + 8 * h];
        uint4 v;
        v.x = pack_bf16x2(s[0], s[1]); v.y = pack_bf16x2(s[2], s[3]);
        v.z = pack_bf16x2(s[4], s[5]); v.w = pack_bf16x2(s[6], s[7]);
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
  if (out_mode == V100_MEL_LOG_BF16_NCW) {
    if (out_pitch < T || (out_pitch & 7) || (reinterpret_cast<uintptr_t>(out) & 15))
      return fail(V100_E_INVALID, "logmel: bf16 NCW pitch must be >= T and a multiple of 8, base 16B aligned");
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
