"""Launch each hot kernel once at the benchmark's shapes (asr_en_base, 256 x 15 s) so that
`ncu --set full -k regex:...` can capture them without replaying a whole step.
    ncu --set full --clock-control none --import-source on -k regex:"dw_mma|conv_gemm|logmel" -o gpurun_out/prof python tools/profile_kernels.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from voice100_b200 import kernels as K
from voice100_b200.data_modules import MelSpectrogramAudioTransform

DEV = "cuda"
B = int(os.environ.get("PROF_B", "256"))
T = 751
which = set((os.environ.get("PROF_WHICH") or "dw,gemm,logmel").split(","))   # also: v2


def ncw(C, T):
    x = K.empty_ncw(B, C, T, DEV)
    x.data.normal_()
    return x


if "dw" in which:
    for C, k in ((2048, 83), (2048, 67), (2048, 59), (1024, 35)):
        x = ncw(C, T)
        w = (torch.randn(C, k, device=DEV) / k ** 0.5).to(torch.bfloat16)
        s, b = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
        for _ in range(2):
            K.dwconv(x, w, s, b, k, 1, K.ACT_RELU6)
    x = ncw(256, 1501)
    w = (torch.randn(256, 11, device=DEV) / 3).to(torch.bfloat16)
    K.dwconv(x, w, torch.ones(256, device=DEV), torch.zeros(256, device=DEV), 11, 2, K.ACT_RELU6)
if "gemm" in which:
    for Ci, Co, res in ((512, 2048, False), (2048, 512, True), (256, 1024, False), (1024, 256, True)):
        x = ncw(Ci, T)
        W = (torch.randn(Co, Ci, device=DEV) / Ci ** 0.5).to(torch.bfloat16)
        s, b = torch.ones(Co, device=DEV), torch.zeros(Co, device=DEV)
        r = ncw(Co, T) if res else None
        for _ in range(2):
            K.conv1x1(x, W, s, b, K.ACT_NONE if res else K.ACT_RELU6, r)
if "logmel" in which:
    tr = MelSpectrogramAudioTransform().to(DEV)
    wav = 0.1 * torch.randn(B, 240000, device=DEV)
    ln = torch.full((B,), 240000, dtype=torch.int32, device=DEV)
    for _ in range(2):
        tr.logmel_batch(wav, ln, ncw_bf16=True)
if "v2" in which:
    # the v2 kernels at the asr_v2 benchmark shapes; the LSTM over PROF_LSTM_T steps (ncu replays the kernel)
    x = ncw(512, T)
    g, be = torch.ones(512, device=DEV), torch.zeros(512, device=DEV)
    wp = (torch.randn(512, 5 * 512, device=DEV) / 50).to(torch.bfloat16)
    for _ in range(2):
        y = K.conv1d(x, wp, be, 5, 1, 2)
        K.layernorm_gelu(y, g, be, 1e-5)
        tm = K.ncw_to_tm(y)
    Tl = int(os.environ.get("PROF_LSTM_T", "96"))
    H = 512
    xt = K.Tm(torch.randn(H, Tl * K.pitch_of(B), device=DEV).to(torch.bfloat16), B, Tl, K.pitch_of(B))
    w_ih = (torch.randn(8 * H, H, device=DEV) / H ** 0.5).to(torch.bfloat16)
    w_hh = (torch.randn(2, 4 * H, H, device=DEV) / H ** 0.5).to(torch.bfloat16)
    lens = torch.full((B,), Tl, dtype=torch.int32, device=DEV)
    for _ in range(2):
        K.lstm_layer(xt, w_ih, torch.zeros(8 * H, device=DEV), w_hh, lens)
torch.cuda.synchronize()
print("profile_kernels done")
