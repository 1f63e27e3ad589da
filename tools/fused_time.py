"""Per-block timing of the fused expand+depthwise kernel against the two kernels it replaces (GPU box).
    python tools/fused_time.py [--once]     (--once: one launch per shape, for ncu)"""
import math, sys, torch
sys.path.insert(0, "/root/repo")
from voice100_b200 import kernels as K
dev = "cuda"
B, T = 256, 751
once = "--once" in sys.argv
shapes = [(256, 1024, 19), (256, 1024, 35), (256, 1024, 51), (512, 2048, 59), (512, 2048, 83)]
if once:
    shapes = [(256, 1024, 35), (512, 2048, 83)]
for C_in, H, k in shapes:
    x = K.Ncw(torch.randn(B, C_in, K.pitch_of(T), device=dev).to(torch.bfloat16), T)
    W1 = (torch.randn(H, C_in, device=dev) / math.sqrt(C_in)).to(torch.bfloat16)
    wd = (torch.randn(H, k, device=dev) / math.sqrt(k)).to(torch.bfloat16)
    s1, b1 = torch.rand(H, device=dev) + 0.5, torch.rand(H, device=dev)
    pairs = K.dw_pack_pairs(wd)
    def fused():
        return K.expand_dw(x, W1, s1, b1, pairs, s1, b1, k)
    def split():
        return K.dwconv(K.conv1x1(x, W1, s1, b1, K.ACT_RELU6), wd, s1, b1, k, 1, K.ACT_RELU6)
    if once:
        fused(); torch.cuda.synchronize(); continue
    out = {}
    for name, fn in (("fused", fused), ("split", split)):
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        out[name] = e0.elapsed_time(e1) / 10
    fl = 2.0 * B * T * C_in * H
    print(f"C_in={C_in} H={H} k={k}: fused {out['fused']*1e3:.1f} us ({fl/out['fused']/1e9:.0f} TFLOP/s on the GEMM part)  "
          f"split {out['split']*1e3:.1f} us  -> x{out['split']/out['fused']:.2f}")
