"""Eager launches vs CUDA-graph replay of the whole ASR path at the benchmark shape (same box, alternating)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import voice100_b200 as v
from voice100_b200 import synth
dev = "cuda"
cfg = dict(audio_size=64, embed_size=512, vocab_size=29, hidden_size=512)
model = v.AudioToTextCTC(**cfg)
model.load_state_dict({k: torch.from_numpy(np.asarray(t)) for k, t in synth.asr_state_dict(**cfg).items()})
model = model.to(dev).eval()
pipe = v.AsrPipeline(v.MelSpectrogramAudioTransform().to(dev), model)
B, L = int(os.environ.get("B", 256)), 240000
wav = 0.1 * torch.randn(B, L, device=dev)
ln = torch.full((B,), L, dtype=torch.int32, device=dev)
run = pipe.graphed(B, L)
def timeit(fn, n=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for rep in range(3):
    t0 = time.perf_counter(); pipe(wav, ln); cpu_ms = (time.perf_counter() - t0) * 1e3
    print(f"rep {rep}: eager {timeit(lambda: pipe(wav, ln)):.3f} ms  graph {timeit(lambda: run.graph.replay()):.3f} ms  (CPU enqueue of one eager step {cpu_ms:.2f} ms)")
