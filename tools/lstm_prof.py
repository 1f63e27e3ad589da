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
# slots: 0 poll start, 1 counter seen, 8 TMA issued, 10 tile landed, 2 MMAs committed, 3 accumulator in registers,
# 9 activations written, 11 exchange barrier passed, 4 cell update + h stores done, 5 fence, 6 barrier, 7 released
seq = [("counter seen -> TMA issued", (1, 8)), ("TMA issued -> both halves landed", (8, 10)),
       ("MMAs issued + commit", (10, 2)), ("commit -> accumulator in registers", (2, 3)),
       ("+ Gx, activation -> exchange buffer", (3, 9)), ("exchange barrier", (9, 11)),
       ("cell update + h stores", (11, 4)), ("proxy fence", (4, 5)), ("barrier", (5, 6)), ("red.release", (6, 7))]
print(f"B={B} H={H}: median step {np.median(st[1:, 0] - st[:-1, 0]):.0f} ns")
print(f"  {'release done -> counter seen (next step)':45s} {np.median(st[1:, 1] - st[:-1, 7]):7.0f} ns")
for name, ab in seq:
    print(f"  {name:45s} {np.median(st[:, ab[1]] - st[:, ab[0]]):7.0f} ns")
