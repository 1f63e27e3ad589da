"""Where does an LSTM step go?  Builds a profiling variant of libv100 (-DV100_LSTM_PROF: globaltimer stamps of
block 0 for steps 100..163) next to the shipped library and prints the median phase durations.

    python tools/lstm_prof.py --build      # here (nvcc), writes voice100_b200/libv100_prof.so
    V100_LIB=voice100_b200/libv100_prof.so python tools/lstm_prof.py      # on the GPU box
"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if "--build" in sys.argv:
    from voice100_b200 import build
    out = os.path.join(ROOT, "voice100_b200", "libv100_prof.so")
    cmd = ["/usr/local/cuda/bin/nvcc"] + build.NVCC_FLAGS + ["-DV100_LSTM_PROF", "-o", out] + build.SOURCES
    subprocess.run(cmd, cwd=build.CSRC, check=True)
    print(out)
    sys.exit(0)

import numpy as np
import torch
from voice100_b200 import _lib, kernels as K

B, T, H = int(os.environ.get("B", 256)), 751, int(os.environ.get("H", 512))
dev = "cuda"
x = K.Tm(torch.randn(H, T * K.pitch_of(B), device=dev).to(torch.bfloat16), B, T, K.pitch_of(B))
w_ih = (torch.randn(8 * H, H, device=dev) / H ** 0.5).to(torch.bfloat16)
w_hh = (torch.randn(2, 4 * H, H, device=dev) / H ** 0.5).to(torch.bfloat16)
bias = torch.zeros(8 * H, device=dev)
lens = torch.full((B,), T, dtype=torch.int32, device=dev)
for _ in range(3):
    K.lstm_layer(x, w_ih, bias, w_hh, lens)
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 768)()
fn = _lib.lib().v100_debug_lstm_prof
fn.argtypes = [ctypes.c_void_p]
assert fn(buf) == 0
st = np.array(buf, dtype=np.int64).reshape(64, 12)[2:62]
names = ["ctl: wait counter", "ctl: TMA h + MMA issue + commit", "gate: commit -> acc in registers",
         "gate: Gx wait + gate math + h store", "gate: proxy fence", "gate: bar.sync", "gate: red.release",
         "next step: publish -> control thread starts waiting"]
d = [st[:, 1] - st[:, 0], st[:, 2] - st[:, 1], st[:, 3] - st[:, 2], st[:, 4] - st[:, 3], st[:, 5] - st[:, 4],
     st[:, 6] - st[:, 5], st[:, 7] - st[:, 6]]
step = np.median(st[1:, 0] - st[:-1, 0])
print(f"B={B} H={H}: median step {step:.0f} ns")
for n, v in zip(names, d):
    print(f"  {n:50s} median {np.median(v):7.0f} ns   p90 {np.percentile(v, 90):7.0f}")
print(f"  {'red.release done -> own counter wait satisfied':50s} median {np.median(st[1:, 1] - st[:-1, 7]):7.0f} ns")
print(f"  {'  counter seen -> all TMA loads issued':50s} median {np.median(st[:, 8] - st[:, 1]):7.0f} ns")
print(f"  {'  loads issued -> first 16 KB box landed':50s} median {np.median(st[:, 9] - st[:, 8]):7.0f} ns")
print(f"  {'  first box -> last box landed':50s} median {np.median(st[:, 10] - st[:, 9]):7.0f} ns")
print(f"  {'  last box -> MMAs issued + commit':50s} median {np.median(st[:, 2] - st[:, 10]):7.0f} ns")
