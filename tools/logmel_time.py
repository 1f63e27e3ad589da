import os, sys, torch
sys.path.insert(0, "/root/repo")
import voice100_b200 as v
tr = v.MelSpectrogramAudioTransform().to("cuda")
wav = 0.1 * torch.randn(256, 240000, device="cuda")
ln = torch.full((256,), 240000, dtype=torch.int32, device="cuda")
for _ in range(3): tr.logmel_batch(wav, ln, ncw_dtype=torch.bfloat16)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): tr.logmel_batch(wav, ln, ncw_dtype=torch.bfloat16)
e1.record(); torch.cuda.synchronize()
print(os.environ.get("V100_LIB", "current"), "logmel 256x15s: %.1f us" % (e0.elapsed_time(e1) / 20 * 1e3))
