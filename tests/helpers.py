"""Shared fixture plumbing: rebuild the exact state dicts / inputs behind tests/golden/*.npz."""
import os

import numpy as np
import torch

from voice100_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def with_bn(sd_np, g, prefix="bn/"):
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in sd_np.items()}
    n = 0
    for k in g.files:
        if k.startswith(prefix):
            sd[k[len(prefix):]] = torch.from_numpy(g[k])
            n += 1
    assert n > 0
    return sd


def asr_case(name):
    """-> (state_dict, waveform[B,L] fp32, lengths, golden npz)"""
    g = golden(name)
    audio_size, embed, vocab, hidden, batch, samples, seed = [int(x) for x in g["cfg"]]
    sd = with_bn(synth.asr_state_dict(audio_size, embed, vocab, hidden, seed=seed, randomize_bn=True), g)
    wav = torch.from_numpy(synth.noise_waveform(batch, samples, seed=seed))
    return sd, wav, [int(x) for x in g["lengths"]], g


def tts_case():
    g = golden("tts_en_base")
    V, H, B, L, seed = [int(x) for x in g["cfg"]]
    sd_a = with_bn(synth.align_state_dict(V, H, seed=seed, randomize_bn=True), g, "a/bn/")
    sd_v = with_bn(synth.audio_state_dict(V, H, seed=seed, randomize_bn=True, randomize_norm=True), g, "v/bn/")
    text = torch.from_numpy(synth.text_tokens(B, L, V, seed=seed))
    align = synth.synthetic_alignment(B, L, seed=seed)
    return sd_a, sd_v, text, align, g


def asr_v2_case(name):
    """-> (state_dict, waveform, lengths, encoder settings, golden npz) for an AudioToAlignText fixture"""
    g = golden(name)
    audio_size, hidden, vocab, batch, samples, seed = [int(x) for x in g["cfg"]]
    settings = synth.ASR_V2_BASE_ENCODER if hidden == 512 else synth.ASR_V2_SMALL_ENCODER
    sd = {k: torch.from_numpy(v) for k, v in synth.asr_v2_state_dict(
        audio_size, settings, 2, hidden, vocab, seed=seed, randomize_ln=True, gain=2.0).items()}
    wav = torch.from_numpy(synth.noise_waveform(batch, samples, seed=seed))
    return sd, wav, [int(x) for x in g["lengths"]], settings, g


def tts_v2_case():
    g = golden("tts_v2_en_base")
    V, B, L, seed = [int(x) for x in g["cfg"]]
    sd_a = {k: torch.from_numpy(v) for k, v in synth.align_v2_state_dict(V, 2, 256, 2, seed=seed, gain=2.0).items()}
    sd_v = {k: torch.from_numpy(v) for k, v in synth.audio_v2_state_dict(
        V, seed=seed, randomize_ln=True, randomize_norm=True, gain=2.0).items()}
    text = torch.from_numpy(synth.text_tokens(B, L, V, seed=seed))
    align = synth.synthetic_alignment(B, L, seed=seed)
    return sd_a, sd_v, text, align, g
