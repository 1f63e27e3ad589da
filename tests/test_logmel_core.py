"""The FFT index math used by the CUDA log-mel kernel (csrc/logmel_core.h compiles for host and device):
built here with g++ and checked against numpy.fft.rfft, plus the full host restatement of one frame
against the oracle's mel pipeline."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HARNESS = r"""
#include "logmel_core.h"
#include <cmath>
using namespace v100;
extern "C" void rfft512_power_host(const float* frame, float* power) {
  static cpx tw[512];
  for (int k = 0; k < 512; ++k) { tw[k].x = (float)cos(-2.0*M_PI*k/512.0); tw[k].y = (float)sin(-2.0*M_PI*k/512.0); }
  cpx A[256], B[256];
  for (int n = 0; n < 256; ++n) { A[fswz(n)].x = frame[2*n]; A[fswz(n)].y = frame[2*n+1]; }
  for (int i = 0; i < 64; ++i) fft256_butterfly(A, B, 256, 1, i, fft256_twiddles(1, i, tw));
  for (int i = 0; i < 64; ++i) fft256_butterfly(B, A, 64, 4, i, fft256_twiddles(4, i, tw));
  for (int i = 0; i < 64; ++i) fft256_butterfly(A, B, 16, 16, i, fft256_twiddles(16, i, tw));
  for (int i = 0; i < 64; ++i) fft256_butterfly_last(B, A, i);
  for (int k = 0; k <= 256; ++k) power[k] = rfft512_power(A, k, tw[k]);
}
extern "C" int fswz_host(int i) { return fswz(i); }
"""


@pytest.fixture(scope="module")
def host_fft():
    d = tempfile.mkdtemp()
    src = os.path.join(d, "h.cpp")
    with open(src, "w") as f:
        f.write(HARNESS)
    so = os.path.join(d, "h.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "voice100_b200", "csrc"), "-o", so, src])
    lib = ctypes.CDLL(so)

    def run(frame):
        frame = np.ascontiguousarray(frame, np.float32)
        out = np.zeros(257, np.float32)
        lib.rfft512_power_host(frame.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p))
        return out
    return run


def test_rfft512_power_matches_numpy(host_fft):
    rng = np.random.default_rng(0)
    for trial in range(4):
        x = rng.standard_normal(512).astype(np.float32) * (10.0 ** (trial - 2))
        ref = np.abs(np.fft.rfft(x.astype(np.float64))) ** 2
        got = host_fft(x)
        assert np.max(np.abs(got - ref)) <= 3e-6 * ref.max()
    imp = np.zeros(512, np.float32)
    imp[3] = 1.0
    np.testing.assert_allclose(host_fft(imp), np.ones(257), rtol=1e-5)            # flat spectrum
    tone = np.cos(2 * np.pi * 37 * np.arange(512) / 512).astype(np.float32)
    p = host_fft(tone)
    assert p.argmax() == 37 and abs(p[37] - 256.0 ** 2) < 1e-2 * 256.0 ** 2       # one bin, N/2 amplitude


def test_swizzle_is_a_conflict_free_bijection(host_fft):
    import glob
    so = glob.glob(os.path.join(tempfile.gettempdir(), "*", "h.so"))
    lib = ctypes.CDLL(sorted(so, key=os.path.getmtime)[-1])
    f = [lib.fswz_host(i) for i in range(256)]
    assert sorted(f) == list(range(256))

    def wavefronts(addrs):           # 64-bit accesses: two half-warps, 16 bank pairs
        tot = 0
        for half in (addrs[:16], addrs[16:]):
            banks = {}
            for a in set(half):
                banks.setdefault(a % 16, set()).add(a)
            tot += max(len(v) for v in banks.values())
        return tot
    n, s = 256, 1
    while n > 1:
        m = n // 4
        for base in (0, 32):
            idx = [base + l for l in range(32)]
            for j in range(4):
                assert wavefronts([f[i % s + s * (i // s + j * m)] for i in idx]) == 2       # reads
                assert wavefronts([f[i % s + s * (4 * (i // s) + j)] for i in idx]) == 2     # writes
        n //= 4
        s *= 4


def test_frame_pipeline_matches_oracle(host_fft):
    import torch
    import v100_oracle as orc
    from voice100_b200 import synth
    from voice100_b200.data_modules import mel_filterbank, sparse_filterbank
    w = synth.harmonic_waveform(1, 4000, seed=3)[0]
    L = len(w)
    start, count, off, fw = sparse_filterbank(mel_filterbank(16000, 512, 64))
    win = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(400) / 400)
    ref = orc.logmel_clip(torch.from_numpy(w)).numpy()
    for t in (0, 1, 7, L // 160):                                  # first frames reflect left, last reflects right
        frame = np.zeros(512, np.float32)
        for m in range(56, 456):
            i = 160 * t - 256 + m
            i = -i if i < 0 else i
            i = 2 * (L - 1) - i if i >= L else i
            frame[m] = w[i] * win[m - 56]
        p = host_fft(frame)
        mel = np.array([np.dot(fw[off[m]:off[m] + count[m]], p[start[m]:start[m] + count[m]]) for m in range(64)])
        np.testing.assert_allclose(np.log(mel + 1e-6), ref[t], rtol=0, atol=2e-3)
