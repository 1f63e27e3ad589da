"""BASELINE.json configs[3]: log-mel front-end sweep (n_fft 512 / win 400 / hop 160 / 64 mels), clip length x
batch, fused CUDA kernel vs the oracle's torch/torchaudio-equivalent CPU path.  Prints a markdown table.
    python tools/logmel_sweep.py            (on the GPU box)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch

import v100_oracle as orc
from voice100_b200 import MelSpectrogramAudioTransform

dev = "cuda"
tr = MelSpectrogramAudioTransform().to(dev)
torch.set_num_threads(os.cpu_count() or 1)
print(f"| clip s | batch | GPU audio-s/s | GPU GB/s (alg.) | CPU audio-s/s ({torch.get_num_threads()} thr) | max abs diff (fp32 out) |")
print("|---|---|---|---|---|---|")
for sec in (1, 5, 15, 30, 60):
    for B in (1, 16, 256, 4096):
        L = sec * 16000
        if B * L > 2 ** 30:
            continue
        wav = 0.1 * torch.randn(B, L, device=dev)
        ln = torch.full((B,), L, dtype=torch.int32, device=dev)
        for _ in range(3):
            tr.logmel_batch(wav, ln, ncw_bf16=True)
        torch.cuda.synchronize()
        reps = 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            tr.logmel_batch(wav, ln, ncw_bf16=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        gbs = (4.0 * B * L + 2.0 * 64 * B * (1 + L // 160)) / (ms * 1e-3) / 1e9
        # CPU: bounded sample (<= 64 clips), scaled
        Bc = min(B, max(1, 64 // sec))
        wc = wav[:Bc].cpu()
        orc.mel_power(wc[:1])
        t0 = time.perf_counter()
        ref = torch.log(orc.mel_power(wc).transpose(1, 2) + 1e-6)
        cpu_s = time.perf_counter() - t0
        got, _ = tr.logmel_batch(wav[:Bc].contiguous(), ln[:Bc].contiguous())
        diff = float((got.cpu() - ref).abs().max())
        print(f"| {sec} | {B} | {B * sec / (ms * 1e-3):,.0f} | {gbs:,.0f} | {Bc * sec / cpu_s:,.0f} | {diff:.1e} |")
