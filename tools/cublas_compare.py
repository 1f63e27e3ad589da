"""What cuBLAS / cuDNN make of the SAME per-layer problems (library comparison point, GPU box only; never on the
product path): bf16 batched GEMM W[Co,Ci] @ x[b][Ci,T] for the pointwise layers, F.conv1d(groups=C) for the depthwise
ones, against v100_conv1x1 / v100_dwconv1d.  Prints one line per layer."""
import math, sys, torch
import torch.nn.functional as F
sys.path.insert(0, "/root/repo")
from voice100_b200 import kernels as K
dev = "cuda"
B, T = 256, 751

def timeit(fn, reps=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for Ci, Co, act in ((512, 2048, 1), (2048, 512, 0), (256, 1024, 1), (1024, 256, 0), (1024, 512, 0)):
    x = K.Ncw(torch.randn(B, Ci, K.pitch_of(T), device=dev).to(torch.bfloat16), T)
    W = (torch.randn(Co, Ci, device=dev) / math.sqrt(Ci)).to(torch.bfloat16)
    s, b = torch.rand(Co, device=dev) + 0.5, torch.rand(Co, device=dev)
    ours = timeit(lambda: K.conv1x1(x, W, s, b, act))
    lib = timeit(lambda: torch.matmul(W, x.data))                      # GEMM only: no BN, no ReLU6, no residual
    fl = 2.0 * B * T * Ci * Co
    print(f"pointwise {Ci}->{Co}: libv100 (GEMM+BN+act) {ours*1e3:.1f} us = {fl/ours/1e9:.0f} TFLOP/s | "
          f"cuBLAS bmm alone {lib*1e3:.1f} us = {fl/lib/1e9:.0f} TFLOP/s")
for C, k in ((2048, 83), (2048, 59), (1024, 35), (1024, 19)):
    x = K.Ncw(torch.randn(B, C, K.pitch_of(T), device=dev).to(torch.bfloat16), T)
    w = (torch.randn(C, k, device=dev) / math.sqrt(k)).to(torch.bfloat16)
    s, b = torch.rand(C, device=dev) + 0.5, torch.rand(C, device=dev)
    ours = timeit(lambda: K.dwconv(x, w, s, b, k, 1, K.ACT_RELU6))
    xd = x.data[:, :, :T].contiguous()
    lib = timeit(lambda: F.conv1d(xd, w[:, None, :], padding=(k - 1) // 2, groups=C), reps=3)   # conv only
    gb = 2.0 * B * C * T * 2
    print(f"depthwise C={C} k={k}: libv100 (conv+BN+ReLU6) {ours*1e3:.1f} us = {gb/ours/1e6:.0f} GB/s | "
          f"cuDNN conv alone {lib*1e3:.1f} us")
