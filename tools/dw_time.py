import os, sys, torch
sys.path.insert(0, "/root/repo")
from voice100_b200 import kernels as K
dev = "cuda"
def run(C, B=256, T=751, k=83, reps=20):
    x = K.empty_ncw(B, C, T, dev); x.data.normal_()
    w = (torch.randn(C, k, device=dev) / k ** 0.5).to(torch.bfloat16)
    s, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    for _ in range(3): y = K.dwconv(x, w, s, b, k, 1, K.ACT_RELU6)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): y = K.dwconv(x, w, s, b, k, 1, K.ACT_RELU6)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gb = 2 * B * C * T * 2 / 1e9
    print(f"{os.environ.get('V100_LIB', 'default')}: C={C} B={B} T={T} k={k}: {ms*1e3:.1f} us, {gb/ms*1e3/1e3:.2f} TB/s")
for k in (83, 75, 67, 59):
    run(2048, k=k)
for k in (51, 35, 27, 19):
    run(1024, k=k)
