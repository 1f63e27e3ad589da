import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import voice100_b200 as v
from voice100_b200 import _lib, synth
dev = torch.device("cuda")
B, L, V, H = 256, 100, 29, 512
amodel = v.TextToAlignTextModel(V, H); amodel.load_state_dict({k: torch.from_numpy(np.asarray(t)) for k, t in synth.align_state_dict(V, H, seed=1234).items()})
vmodel = v.AlignTextToAudioModel(V, H); vmodel.load_state_dict({k: torch.from_numpy(np.asarray(t)) for k, t in synth.audio_state_dict(V, H, seed=1234).items()})
amodel, vmodel = amodel.to(dev).eval(), vmodel.to(dev).eval()
text = torch.from_numpy(synth.text_tokens(B, L, V, seed=1234)); align = torch.from_numpy(synth.synthetic_alignment(B, L, seed=1234))
at, at_len = v.align_batch(text, align)
text_d, at_d = text.to(dev), at.to(dev)
for _ in range(3): amodel(text_d); vmodel.predict(at_d)
class Tr:
    def __init__(s): s.ev = []
    def before(s, n): s._s = torch.cuda.Event(enable_timing=True); s._s.record()
    def after(s, n):
        e = torch.cuda.Event(enable_timing=True); e.record(); s.ev.append((n, s._s, e))
_lib.tracer = Tr()
amodel(text_d); vmodel.predict(at_d)
torch.cuda.synchronize()
ev, _lib.tracer = _lib.tracer.ev, None
tot = 0
for n, s, e in ev:
    ms = s.elapsed_time(e); tot += ms
    print(f"{n.replace('v100_',''):24s} {ms*1e3:8.1f} us")
print("sum", tot, "aligned T", at.shape)
