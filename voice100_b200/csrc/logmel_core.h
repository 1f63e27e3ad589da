// Index math of the 512-point real FFT used by the log-mel kernel, written so that the SAME code
// compiles for the device (logmel.cu, one warp per frame: `idx` = lane, lane+32) and for the host
// (tests/test_logmel_core.py compiles it with g++ and checks it against numpy.fft.rfft).
//
// A 512-sample real frame is packed into 256 complex points z[n] = x[2n] + i x[2n+1], transformed
// with a radix-4 Stockham autosort FFT (4 passes, natural-order output, ping-pong buffers), and
// unpacked to the 257 one-sided bins.
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

// One radix-4 butterfly of the pass that takes length-n sub-transforms at stride s (n*s == 256).
// idx in [0,64).  tw[k] = exp(-2*pi*i*k/512), k < 512.
V100_HD void fft256_butterfly(const cpx* X, cpx* Y, int n, int s, int idx, const cpx* tw) {
  const int m = n >> 2;
  const int p = idx / s, q = idx - p * s;
  const cpx a = X[q + s * p], b = X[q + s * (p + m)], c = X[q + s * (p + 2 * m)], d = X[q + s * (p + 3 * m)];
  const cpx apc{a.x + c.x, a.y + c.y}, amc{a.x - c.x, a.y - c.y};
  const cpx bpd{b.x + d.x, b.y + d.y};
  const cpx jbmd{-(b.y - d.y), b.x - d.x};  // i*(b-d)
  // exp(-2*pi*i*p/n) = tw[2*p*s] because n*s == 256
  const cpx w1 = tw[2 * p * s], w2 = tw[4 * p * s], w3 = tw[6 * p * s];
  Y[q + s * (4 * p + 0)] = cpx{apc.x + bpd.x, apc.y + bpd.y};
  Y[q + s * (4 * p + 1)] = cmul(w1, cpx{amc.x - jbmd.x, amc.y - jbmd.y});
  Y[q + s * (4 * p + 2)] = cmul(w2, cpx{apc.x - bpd.x, apc.y - bpd.y});
  Y[q + s * (4 * p + 3)] = cmul(w3, cpx{amc.x + jbmd.x, amc.y + jbmd.y});
}

// |X[k]|^2 of the 512-point real transform from the 256-point complex one, k in [0,256].
V100_HD float rfft512_power(const cpx* Z, int k, const cpx* tw) {
  const cpx zk = Z[k & 255], zn = Z[(256 - k) & 255];
  const cpx e{0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y)};   // spectrum of the even samples
  const cpx o{0.5f * (zk.y + zn.y), -0.5f * (zk.x - zn.x)};  // spectrum of the odd samples
  const cpx wo = cmul(tw[k], o);
  const float re = e.x + wo.x, im = e.y + wo.y;
  return re * re + im * im;
}

}  // namespace v100
