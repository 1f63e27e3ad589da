// Arithmetic of the 512-point real FFT used by the log-mel kernel, written so that the SAME code
// compiles for the device (logmel.cu, sixteen threads per frame) and for the host
// (tests/test_logmel_core.py compiles it with g++ and checks it against numpy.fft.rfft).
//
// A 512-sample real frame is packed into 256 complex points z[n] = x[2n] + i x[2n+1], transformed
// as 16 x 16 (two 16-point FFTs in registers around one transpose), and unpacked to the 257
// one-sided bins.
#pragma once

#ifndef V100_HD
#ifdef __CUDACC__
#define V100_HD __host__ __device__ __forceinline__
#else
#define V100_HD inline
#endif
#endif

namespace v100 {

struct cpx {
  float x, y;
};

V100_HD cpx cmul(cpx a, cpx b) { return cpx{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }

// ------------------------------------------------------------------------------------------------
// 256 = 16 x 16.  Sixteen threads transform one frame, each holding 16 complex values:
//   n = q + 16 r,  k = k1 + 16 k2,
//   X[k1 + 16 k2] = sum_q W16^(q k2) * [ W256^(q k1) * sum_r z[q + 16 r] W16^(r k1) ]
// thread q: 16-point FFT over r (in registers) -> twiddle W256^(q k1) -> 16 x 16 transpose through shared
// memory -> thread k1: 16-point FFT over q -> X[k1 + 16 k2], k2 = 0..15.  (The first version of the kernel ran
// four radix-4 Stockham passes through shared memory with one warp per frame: 0.68 ms for 256 x 15 s against
// 0.51 ms for this one.)
// ------------------------------------------------------------------------------------------------

// 4-point forward DFT (W4 = -i), in place on (a, b, c, d) = inputs 0..3 -> outputs 0..3
V100_HD void dft4(cpx& a, cpx& b, cpx& c, cpx& d) {
  const cpx apc{a.x + c.x, a.y + c.y}, amc{a.x - c.x, a.y - c.y};
  const cpx bpd{b.x + d.x, b.y + d.y}, bmd{b.x - d.x, b.y - d.y};
  a = cpx{apc.x + bpd.x, apc.y + bpd.y};
  b = cpx{amc.x + bmd.y, amc.y - bmd.x};   // (a - c) - i (b - d)
  c = cpx{apc.x - bpd.x, apc.y - bpd.y};
  d = cpx{amc.x - bmd.y, amc.y + bmd.x};   // (a - c) + i (b - d)
}

// 16-point forward DFT in place, natural order in and out: v[k] = sum_r v[r] exp(-2 pi i r k / 16).
// r = 4 r1 + r0, k = k0 + 4 k1: DFT4 over r1, twiddle W16^(r0 k0), DFT4 over r0.  Every index is a compile-time
// constant once unrolled, so the array stays in registers.
V100_HD void fft16(cpx* v) {
  // W16^m = exp(-2 pi i m / 16), m = 1, 2, 3, 4, 6, 9
  const float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, r2 = 0.70710678118654752f;
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int r0 = 0; r0 < 4; ++r0) dft4(v[r0], v[4 + r0], v[8 + r0], v[12 + r0]);  // v[4 k0 + r0] = B[r0][k0]
  // twiddles W16^(r0 k0) on v[4 k0 + r0]
  v[5] = cmul(v[5], cpx{c1, -s1});     // r0 1, k0 1: m = 1
  v[6] = cmul(v[6], cpx{r2, -r2});     // r0 2, k0 1: m = 2
  v[7] = cmul(v[7], cpx{s1, -c1});     // r0 3, k0 1: m = 3
  v[9] = cmul(v[9], cpx{r2, -r2});     // r0 1, k0 2: m = 2
  v[10] = cpx{v[10].y, -v[10].x};      // r0 2, k0 2: m = 4 -> -i
  v[11] = cmul(v[11], cpx{-r2, -r2});  // r0 3, k0 2: m = 6
  v[13] = cmul(v[13], cpx{s1, -c1});   // r0 1, k0 3: m = 3
  v[14] = cmul(v[14], cpx{-r2, -r2});  // r0 2, k0 3: m = 6
  v[15] = cmul(v[15], cpx{-c1, s1});   // r0 3, k0 3: m = 9
#if defined(__CUDACC__)
#pragma unroll
#endif
  for (int k0 = 0; k0 < 4; ++k0) dft4(v[4 * k0], v[4 * k0 + 1], v[4 * k0 + 2], v[4 * k0 + 3]);  // -> A[k0 + 4 k1] at v[4 k0 + k1]
  // v[4 k0 + k1] holds A[k0 + 4 k1]: transpose the 4 x 4 index grid to natural order
  cpx t;
  t = v[1]; v[1] = v[4]; v[4] = t;
  t = v[2]; v[2] = v[8]; v[8] = t;
  t = v[3]; v[3] = v[12]; v[12] = t;
  t = v[6]; v[6] = v[9]; v[9] = t;
  t = v[7]; v[7] = v[13]; v[13] = t;
  t = v[11]; v[11] = v[14]; v[14] = t;
}

// |X[k]|^2 of the 512-point real transform from Z[k] and Z[(256 - k) & 255] of the packed 256-point transform
V100_HD float rfft512_power_pair(cpx zk, cpx zn, cpx wk) {
  const cpx e{0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y)};
  const cpx o{0.5f * (zk.y + zn.y), -0.5f * (zk.x - zn.x)};
  const cpx wo = cmul(wk, o);
  const float re = e.x + wo.x, im = e.y + wo.y;
  return re * re + im * im;
}

// Both halves of one conjugate pair at once: with E = (Z[k] + conj Z[N-k]) / 2 and O = -i (Z[k] - conj Z[N-k]) / 2,
// X[k] = E + W^k O and X[256 - k] = conj(E - W^k O), so one complex multiply serves the two power bins.
V100_HD void rfft512_power_both(cpx zk, cpx zn, cpx wk, float* pk, float* pn) {
  const cpx e{0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y)};
  const cpx o{0.5f * (zk.y + zn.y), -0.5f * (zk.x - zn.x)};
  const cpx wo = cmul(wk, o);
  const float re = e.x + wo.x, im = e.y + wo.y;
  const float rn = e.x - wo.x, in = e.y - wo.y;
  *pk = re * re + im * im;
  *pn = rn * rn + in * in;
}

}  // namespace v100
