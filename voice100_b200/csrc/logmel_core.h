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

// Storage swizzle of the 256-point buffers: the Stockham passes read at strides 64/16/4/1 and write at
// strides 1/4/16/64 (in 8-byte complex units); XOR-ing the index with bits [2,6) makes every one of those
// warp-wide 64-bit accesses bank-conflict free (checked exhaustively in tests/test_logmel_core.py).
V100_HD int fswz(int i) { return i ^ ((i >> 2) & 15); }

// The three twiddles of butterfly idx in the pass (n, s): exp(-2*pi*i*j*p/n), j = 1..3, p = idx / s.
// tw[k] = exp(-2*pi*i*k/512), k < 512; exp(-2*pi*i*p/n) = tw[2*p*s] because n*s == 256.
struct tw3 {
  cpx w1, w2, w3;
};
V100_HD tw3 fft256_twiddles(int s, int idx, const cpx* tw) {
  const int p = idx / s;
  return tw3{tw[2 * p * s], tw[4 * p * s], tw[6 * p * s]};
}

// One radix-4 butterfly of the pass that takes length-n sub-transforms at stride s (n*s == 256), idx in [0,64).
V100_HD void fft256_butterfly(const cpx* X, cpx* Y, int n, int s, int idx, const tw3& t) {
  const int m = n >> 2;
  const int p = idx / s, q = idx - p * s;
  const cpx a = X[fswz(q + s * p)], b = X[fswz(q + s * (p + m))], c = X[fswz(q + s * (p + 2 * m))],
            d = X[fswz(q + s * (p + 3 * m))];
  const cpx apc{a.x + c.x, a.y + c.y}, amc{a.x - c.x, a.y - c.y};
  const cpx bpd{b.x + d.x, b.y + d.y};
  const cpx jbmd{-(b.y - d.y), b.x - d.x};  // i*(b-d)
  Y[fswz(q + s * (4 * p + 0))] = cpx{apc.x + bpd.x, apc.y + bpd.y};
  Y[fswz(q + s * (4 * p + 1))] = cmul(t.w1, cpx{amc.x - jbmd.x, amc.y - jbmd.y});
  Y[fswz(q + s * (4 * p + 2))] = cmul(t.w2, cpx{apc.x - bpd.x, apc.y - bpd.y});
  Y[fswz(q + s * (4 * p + 3))] = cmul(t.w3, cpx{amc.x + jbmd.x, amc.y + jbmd.y});
}
// Last pass (n = 4, s = 64): p == 0, all twiddles are 1.
V100_HD void fft256_butterfly_last(const cpx* X, cpx* Y, int idx) {
  const cpx a = X[fswz(idx)], b = X[fswz(idx + 64)], c = X[fswz(idx + 128)], d = X[fswz(idx + 192)];
  const cpx apc{a.x + c.x, a.y + c.y}, amc{a.x - c.x, a.y - c.y};
  const cpx bpd{b.x + d.x, b.y + d.y};
  const cpx jbmd{-(b.y - d.y), b.x - d.x};
  Y[fswz(idx)] = cpx{apc.x + bpd.x, apc.y + bpd.y};
  Y[fswz(idx + 64)] = cpx{amc.x - jbmd.x, amc.y - jbmd.y};
  Y[fswz(idx + 128)] = cpx{apc.x - bpd.x, apc.y - bpd.y};
  Y[fswz(idx + 192)] = cpx{amc.x + jbmd.x, amc.y + jbmd.y};
}

// |X[k]|^2 of the 512-point real transform from the 256-point complex one, k in [0,256];
// wk = exp(-2*pi*i*k/512).
V100_HD float rfft512_power(const cpx* Z, int k, cpx wk) {
  const cpx zk = Z[fswz(k & 255)], zn = Z[fswz((256 - k) & 255)];
  const cpx e{0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y)};   // spectrum of the even samples
  const cpx o{0.5f * (zk.y + zn.y), -0.5f * (zk.x - zn.x)};  // spectrum of the odd samples
  const cpx wo = cmul(wk, o);
  const float re = e.x + wo.x, im = e.y + wo.y;
  return re * re + im * im;
}

}  // namespace v100
