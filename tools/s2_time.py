"""Times the stride-2 depthwise layer of asr_en_base (C = 256, T_in = 1501, k = 11) through v100_dwconv1d."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from voice100_b200 import kernels as K
dev = "cuda"
B, C, T, k = 256, 256, 1501, 11
x = K.empty_ncw(B, C, T, dev); x.data.normal_()
w = (torch.randn(C, k, device=dev) / k ** 0.5).to(torch.bfloat16)
s, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
for _ in range(3): y = K.dwconv(x, w, s, b, k, 2, K.ACT_RELU6)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): y = K.dwconv(x, w, s, b, k, 2, K.ACT_RELU6)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 50
gb = B * C * (T + (T - 1) // 2 + 1) * 2 / 1e9
print(f"{os.environ.get('V100_LIB', 'default')}: stride-2 dw C={C} T_in={T} k={k}: {ms*1e3:.1f} us, {gb/ms*1e3:.0f} GB/s = {gb/ms*1e3/6555.5:.2f} of HBM")
