"""Does a tensor-bound GEMM on HALF of the SMs keep its per-SM speed while the depthwise FIR runs on the other half?
(The full-chip GEMM is power-limited: profiles/r02_gemm_stalls.md.)  Run with V100_GEMM_PAIRS=37 (or another cap): the
project GEMM then occupies 2 x cap SMs with one 200 KB CTA each, and the depthwise CTAs (48 KB) can only land on the rest.
Prints the time of N GEMM launches alone, N depthwise launches alone, and both streams together."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from voice100_b200 import kernels as K
dev = "cuda"
B, T, N = 256, 751, 10
x = K.empty_ncw(B, 2048, T, dev); x.data.normal_().clamp_(0, 6)
W = (torch.randn(512, 2048, device=dev) / 45).to(torch.bfloat16)
s512, b512 = torch.ones(512, device=dev), torch.zeros(512, device=dev)
res = K.empty_ncw(B, 512, T, dev); res.data.normal_()
xd = K.empty_ncw(B, 2048, T, dev); xd.data.normal_().clamp_(0, 6)
wd = (torch.randn(2048, 83, device=dev) / 9).to(torch.bfloat16)
s2k, b2k = torch.ones(2048, device=dev), torch.zeros(2048, device=dev)
st_g, st_d = torch.cuda.Stream(), torch.cuda.Stream()

def gemm(): K.conv1x1(x, W, s512, b512, 0, res)
def dw(): K.dwconv(xd, wd, s2k, b2k, 83, 1, K.ACT_RELU6)

def run(do_g, do_d):
    torch.cuda.synchronize()
    ev = {k: (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for k in "gd"}
    if do_g:
        with torch.cuda.stream(st_g):
            ev["g"][0].record()
            for _ in range(N): gemm()
            ev["g"][1].record()
    if do_d:
        with torch.cuda.stream(st_d):
            ev["d"][0].record()
            for _ in range(N): dw()
            ev["d"][1].record()
    torch.cuda.synchronize()
    return (ev["g"][0].elapsed_time(ev["g"][1]) / N * 1e3 if do_g else 0.0,
            ev["d"][0].elapsed_time(ev["d"][1]) / N * 1e3 if do_d else 0.0)

for _ in range(2): run(True, True)
for rep in range(2):
    g_alone, _ = run(True, False)
    _, d_alone = run(False, True)
    g_both, d_both = run(True, True)
    print(f"pairs cap {os.environ.get('V100_GEMM_PAIRS', '-')}: GEMM 2048->512 alone {g_alone:.0f} us, dw k=83 alone {d_alone:.0f} us | "
          f"together: GEMM {g_both:.0f} us, dw {d_both:.0f} us per launch (serial sum {g_alone + d_alone:.0f} us)")
