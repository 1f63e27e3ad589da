#!/usr/bin/env python
"""Error budget of the 16-bit storage contract (CPU only; TEST INFRASTRUCTURE -- imports the oracle).

Evaluates ConvVoiceEncoder + head (voice100/models/asr.py:62-116) in fp32 with ONE family of stored tensors
rounded at a time, and with candidate mixed contracts, against the fp32 oracle:

  feat   log-mel features entering block 0          w    all conv weights
  h1     expand output (4x hidden, post ReLU6)      h2   depthwise output (4x hidden, post ReLU6)
  out    block outputs / residual stream            head head weights

    python tools/error_budget.py [--model small|base] [--batch 4] [--seconds 5]

Prints a table (max|err|/sigma, rms/sigma, raw greedy agreement) that DESIGN.md section 4 quotes.
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import v100_oracle as orc  # noqa: E402
from voice100_b200 import synth  # noqa: E402

SITES = ("feat", "w", "h1", "h2", "out", "head")


def q(t, dtype):
    return t if dtype is None else t.to(dtype).float()


def forward(audio, sd, site_dtype):
    """site_dtype: dict site -> torch dtype or None (fp32)."""
    g = lambda s: site_dtype.get(s)
    x = q(audio.transpose(1, 2), g("feat"))
    for i, (k, s, r) in enumerate(orc._asr_blocks(sd)):
        p = f"encoder.layers.{i}"
        s1, b1 = orc._fold(sd, p + ".conv.0.1")
        s2, b2 = orc._fold(sd, p + ".conv.1.1")
        s3, b3 = orc._fold(sd, p + ".conv.3")
        h = F.conv1d(x, q(sd[p + ".conv.0.0.weight"], g("w")))
        h = q((h * s1 + b1).clamp(0.0, 6.0), g("h1"))
        h = F.conv1d(h, q(sd[p + ".conv.1.0.weight"], g("w")), stride=s, padding=(k - 1) // 2, groups=h.shape[1])
        h = q((h * s2 + b2).clamp(0.0, 6.0), g("h2"))
        y = F.conv1d(h, q(sd[p + ".conv.2.weight"], g("w"))) * s3 + b3
        x = q(x + y if r else y, g("out"))
    logits = F.conv1d(x, q(sd["decoder.layers.1.weight"], g("head")), sd["decoder.layers.1.bias"])
    return logits.transpose(1, 2)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="small", choices=["small", "base"])
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--seconds", type=float, default=5.0)
    ap.add_argument("--seed", type=int, default=7)
    args = ap.parse_args()
    hidden = 256 if args.model == "small" else 512
    cfg = dict(audio_size=64, embed_size=hidden, vocab_size=29, hidden_size=hidden)
    L = int(16000 * args.seconds)
    wav = torch.from_numpy(synth.noise_waveform(args.batch, L, seed=args.seed))
    sd = orc.to_torch_sd(synth.asr_state_dict(**cfg, seed=args.seed, randomize_bn=True))
    audio, _ = orc.logmel_batch(wav, [L] * args.batch)
    sd = orc.calibrate_asr(sd, audio)
    bf, hf = torch.bfloat16, torch.float16
    with torch.no_grad():
        ref = orc.asr_forward(audio, sd)
        rows = []

        def run(name, sites):
            got = forward(audio, sd, sites)
            rep = orc.parity_report(ref, got)
            agree = float((ref.argmax(-1) == got.argmax(-1)).float().mean())
            rows.append((name, rep["max_abs_rel_std"], rep["rms_rel_std"], agree))

        run("fp32 everywhere (sanity)", {})
        for s in SITES:
            run(f"bf16 only at {s}", {s: bf})
        run("bf16 everywhere (round-1 contract)", {s: bf for s in SITES})
        run("fp16 everywhere", {s: hf for s in SITES})
        run("bf16, fp32 residual stream (out)", {s: bf for s in SITES if s != "out"})
        run("bf16, fp16 residual stream (out)", dict({s: bf for s in SITES}, out=hf))
        run("bf16, feat as bf16 hi+lo pair (~fp32)", {s: bf for s in SITES if s != "feat"})
        run("bf16, fp16 out+feat", dict({s: bf for s in SITES}, out=hf, feat=hf))
        run("bf16 w/head, fp16 activations (feat,h1,h2,out)", dict(w=bf, head=bf, feat=hf, h1=hf, h2=hf, out=hf))
        run("bf16 w/head/h1/h2, fp16 feat+out", dict(w=bf, head=bf, h1=bf, h2=bf, feat=hf, out=hf))
        run("bf16 w/head/h2, fp16 feat+h1+out", dict(w=bf, head=bf, h2=bf, h1=hf, feat=hf, out=hf))
        run("fp16 w + activations, bf16 nothing", {s: hf for s in SITES})
        run("fp16 weights only", dict(w=hf, head=hf))
    print(f"model asr_en_{args.model} hidden={hidden}, {args.batch} x {args.seconds:g} s, sigma(logits)={float(ref.std()):.4f}")
    print(f"{'contract':<52s} {'max/sigma':>10s} {'rms/sigma':>10s} {'greedy agree':>13s}")
    for name, mx, rms, ag in rows:
        print(f"{name:<52s} {mx:10.4f} {rms:10.4f} {ag:13.4f}")


if __name__ == "__main__":
    main()
