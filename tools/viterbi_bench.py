"""Forced alignment: device kernel vs the oracle's numpy DP (the reference's algorithm) on one host core."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import v100_oracle as orc
import voice100_b200 as v
from voice100_b200 import synth
B, T, L, V = 256, 751, 120, 29
lp = torch.zeros(B, T, V); text = torch.zeros(B, L, dtype=torch.int64)
for b in range(B):
    a, lab = synth.viterbi_inputs(T, L, V, 7000 + b)
    lp[b], text[b] = torch.from_numpy(a), torch.from_numpy(lab)
lp_d, text_d = lp.cuda(), text.cuda()
n_f, n_t = torch.full((B,), T), torch.full((B,), L)
for _ in range(3): out = v.ctc_best_path_batch(lp_d, n_f, text_d, n_t)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): out = v.ctc_best_path_batch(lp_d, n_f, text_d, n_t)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
t0 = time.perf_counter()
n_cpu = 8
for b in range(n_cpu):
    s, p, l = orc.ctc_best_path(lp[b].numpy(), text[b].numpy())
    assert np.array_equal(p, out[1][b].cpu().numpy())
cpu = (time.perf_counter() - t0) / n_cpu
print(f"ctc_best_path: B={B} T={T} L={L}: GPU {ms:.3f} ms/batch = {B / (ms * 1e-3):,.0f} utt/s; "
      f"numpy DP {cpu * 1e3:.1f} ms/utt = {1 / cpu:,.1f} utt/s on one core; paths identical")
