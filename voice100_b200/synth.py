"""Deterministic synthetic weights and inputs (numpy PCG64, no torch RNG).

There is no network for checkpoints or datasets, so every test, the oracle,
the golden-vector generator and bench.py build their inputs from this one
module.  Everything is a pure function of (spec, seed): the same call yields
the same bytes in this container and on the GPU box, which is what lets the
golden fixtures under tests/golden/ store OUTPUTS only.

The `state_dict` key layout is the reference's own (probed by importing
/root/reference with the shim in oracle/_shim):
  ASR   voice100/models/asr.py:27-94     encoder.layers.{i}.conv.{0.0,0.1,1.0,1.1,2,3}.*,
                                         decoder.layers.1.{weight,bias}
  align voice100/models/tts.py:67-77     embedding.weight, layers.{0-3}.conv.*, layers.4.*
  audio voice100/models/tts.py:13-29,152-170  embedding.weight, decoder.layers.{0-3,5-7}.conv.*,
                                         decoder.layers.4.* (ConvTranspose1d [C_in,C_out,5]),
                                         decoder.layers.8.*, norm.{f0,logspc,codeap}_{mean,std}
  v2    voice100/models/_layers_v2.py:29-103, _asr_v2.py:31-37, _align_v2.py:18-27, _tts_v2.py:35-43
                                         {encoder,decoder}.{i}.{conv,layer_norm}.*, lstm.{weight,bias}_{ih,hh}_l{n}[_reverse],
                                         dense.* / projection.*, embedding.weight, norm.*
"""
from __future__ import annotations

import zlib
from typing import Dict, List, Tuple

import numpy as np

# (in, out, kernel, stride, residual) per InvertedResidual -- voice100/models/asr.py:67-76
def asr_encoder_blocks(audio_size: int, embed_size: int, hidden_size: int):
    half = hidden_size // 2
    return [
        (audio_size, half, 11, 2, False),
        (half, half, 19, 1, True),
        (half, half, 27, 1, True),
        (half, half, 35, 1, True),
        (half, hidden_size, 51, 1, False),
        (hidden_size, hidden_size, 59, 1, True),
        (hidden_size, hidden_size, 67, 1, True),
        (hidden_size, hidden_size, 75, 1, True),
        (hidden_size, embed_size, 83, 1, False),
    ]


# voice100/models/tts.py:72-76
ALIGN_KERNELS = (5, 11, 17, 29)
# voice100/models/tts.py:17-25
VOICE_DECODER_PRE_KERNELS = (65, 33, 17, 11)
VOICE_DECODER_POST_KERNELS = (33, 11, 7)


def _rng(seed: int, tag: str) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64([int(seed), zlib.crc32(tag.encode())]))


def _uniform(seed, tag, shape, lo, hi):
    u = _rng(seed, tag).random(size=shape)  # float64 in [0,1): stream-stable across numpy versions
    return (lo + (hi - lo) * u).astype(np.float32)


def _bn(sd, seed, prefix, c, randomize_bn):
    if randomize_bn:
        sd[prefix + ".weight"] = _uniform(seed, prefix + ".weight", (c,), 0.5, 1.5)
        sd[prefix + ".bias"] = _uniform(seed, prefix + ".bias", (c,), -0.5, 0.5)
        sd[prefix + ".running_mean"] = _uniform(seed, prefix + ".running_mean", (c,), -0.2, 0.2)
        sd[prefix + ".running_var"] = _uniform(seed, prefix + ".running_var", (c,), 0.5, 1.5)
    else:  # torch defaults (BatchNorm1d.reset_parameters)
        sd[prefix + ".weight"] = np.ones((c,), np.float32)
        sd[prefix + ".bias"] = np.zeros((c,), np.float32)
        sd[prefix + ".running_mean"] = np.zeros((c,), np.float32)
        sd[prefix + ".running_var"] = np.ones((c,), np.float32)
    sd[prefix + ".num_batches_tracked"] = np.zeros((), np.int64)


def _conv(sd, seed, key, shape, fan_in, gain=1.0):
    b = gain / np.sqrt(fan_in)  # torch Conv default: kaiming_uniform(a=sqrt(5)) == U(+-1/sqrt(fan_in))
    sd[key] = _uniform(seed, key, shape, -b, b)


def _inverted_residual(sd, seed, prefix, c_in, c_out, k, randomize_bn, gain):
    h = c_in * 4  # expand_ratio=4, voice100/models/asr.py:41-43
    _conv(sd, seed, f"{prefix}.conv.0.0.weight", (h, c_in, 1), c_in, gain)
    _bn(sd, seed, f"{prefix}.conv.0.1", h, randomize_bn)
    _conv(sd, seed, f"{prefix}.conv.1.0.weight", (h, 1, k), k, gain)
    _bn(sd, seed, f"{prefix}.conv.1.1", h, randomize_bn)
    _conv(sd, seed, f"{prefix}.conv.2.weight", (c_out, h, 1), h, gain)
    _bn(sd, seed, f"{prefix}.conv.3", c_out, randomize_bn)


def asr_state_dict(audio_size=64, embed_size=512, vocab_size=29, hidden_size=512,
                   seed=1234, randomize_bn=False, gain=1.0) -> Dict[str, np.ndarray]:
    """Weights for AudioToTextCTC(audio_size, embed_size, vocab_size, hidden_size)."""
    sd: Dict[str, np.ndarray] = {}
    for i, (ci, co, k, _s, _r) in enumerate(asr_encoder_blocks(audio_size, embed_size, hidden_size)):
        _inverted_residual(sd, seed, f"encoder.layers.{i}", ci, co, k, randomize_bn, gain)
    _conv(sd, seed, "decoder.layers.1.weight", (vocab_size, embed_size, 1), embed_size, gain)
    _conv(sd, seed, "decoder.layers.1.bias", (vocab_size,), embed_size, gain)
    return sd


def align_state_dict(vocab_size=29, hidden_size=512, seed=1234, randomize_bn=False, gain=1.0):
    """Weights for TextToAlignTextModel(vocab_size, hidden_size)."""
    sd: Dict[str, np.ndarray] = {}
    sd["embedding.weight"] = _rng(seed, "embedding.weight").standard_normal(
        (vocab_size, hidden_size)).astype(np.float32)
    for i, k in enumerate(ALIGN_KERNELS):
        _inverted_residual(sd, seed, f"layers.{i}", hidden_size, hidden_size, k, randomize_bn, gain)
    _conv(sd, seed, "layers.4.weight", (2, hidden_size, 1), hidden_size, gain)
    _conv(sd, seed, "layers.4.bias", (2,), hidden_size, gain)
    return sd


def audio_state_dict(vocab_size=29, hidden_size=512, seed=1234, randomize_bn=False, gain=1.0,
                     randomize_norm=False, logspc_size=257):
    """Weights for AlignTextToAudioModel(vocab_size, hidden_size, use_mcep = (logspc_size == 25))."""
    half = hidden_size // 2
    out_ch = 1 + 1 + logspc_size + 1  # hasf0, f0, logspc, codeap -- voice100/models/tts.py:160-167
    sd: Dict[str, np.ndarray] = {}
    sd["embedding.weight"] = _rng(seed, "embedding.weight").standard_normal(
        (vocab_size, hidden_size)).astype(np.float32)
    for i, k in enumerate(VOICE_DECODER_PRE_KERNELS):
        _inverted_residual(sd, seed, f"decoder.layers.{i}", hidden_size, hidden_size, k, randomize_bn, gain)
    # ConvTranspose1d weight layout is [C_in, C_out, k]; torch fan_in for it is C_out*k
    _conv(sd, seed, "decoder.layers.4.weight", (hidden_size, half, 5), half * 5, gain)
    _conv(sd, seed, "decoder.layers.4.bias", (half,), half * 5, gain)
    for j, k in enumerate(VOICE_DECODER_POST_KERNELS):
        _inverted_residual(sd, seed, f"decoder.layers.{5 + j}", half, half, k, randomize_bn, gain)
    _conv(sd, seed, "decoder.layers.8.weight", (out_ch, half, 1), half, gain)
    _conv(sd, seed, "decoder.layers.8.bias", (out_ch,), half, gain)
    if randomize_norm:
        sd["norm.f0_std"] = _uniform(seed, "norm.f0_std", (1,), 20.0, 60.0)
        sd["norm.f0_mean"] = _uniform(seed, "norm.f0_mean", (1,), 100.0, 200.0)
        sd["norm.logspc_std"] = _uniform(seed, "norm.logspc_std", (logspc_size,), 0.5, 2.0)
        sd["norm.logspc_mean"] = _uniform(seed, "norm.logspc_mean", (logspc_size,), -8.0, -2.0)
        sd["norm.codeap_std"] = _uniform(seed, "norm.codeap_std", (1,), 0.5, 2.0)
        sd["norm.codeap_mean"] = _uniform(seed, "norm.codeap_mean", (1,), -3.0, 0.0)
    else:  # voice100/models/_layers_v1.py:100-117 defaults
        sd["norm.f0_std"] = np.ones((1,), np.float32)
        sd["norm.f0_mean"] = np.zeros((1,), np.float32)
        sd["norm.logspc_std"] = np.ones((logspc_size,), np.float32)
        sd["norm.logspc_mean"] = np.zeros((logspc_size,), np.float32)
        sd["norm.codeap_std"] = np.ones((1,), np.float32)
        sd["norm.codeap_mean"] = np.zeros((1,), np.float32)
    return sd


# ----------------------------------------------------------------------------
# v2 models (LayerNorm + GELU conv blocks, bidirectional LSTM)
# ----------------------------------------------------------------------------

# config/asr_en_base.yaml:16-18, config/asr_en_small.yaml:16-18, config/tts_en_base.yaml:20-23
# rows: (out_channels, transpose, kernel_size, stride, padding, bias)
ASR_V2_BASE_ENCODER = ((512, False, 5, 2, 2, False), (512, False, 5, 1, 2, False))
ASR_V2_SMALL_ENCODER = ((256, False, 3, 2, 1, False), (256, False, 3, 1, 1, False))
TTS_V2_BASE_DECODER = ((512, False, 5, 1, 2, False), (512, True, 5, 2, 2, False), (512, False, 5, 1, 2, False))


def _conv_layers_v2(sd, seed, prefix, in_channels, settings, randomize_ln, gain):
    c = in_channels
    for i, (co, transpose, k, _stride, _pad, bias) in enumerate(settings):
        p = f"{prefix}.{i}"
        if randomize_ln:
            sd[p + ".layer_norm.weight"] = _uniform(seed, p + ".layer_norm.weight", (co,), 0.5, 1.5)
            sd[p + ".layer_norm.bias"] = _uniform(seed, p + ".layer_norm.bias", (co,), -0.5, 0.5)
        else:
            sd[p + ".layer_norm.weight"] = np.ones((co,), np.float32)
            sd[p + ".layer_norm.bias"] = np.zeros((co,), np.float32)
        # Conv1d weight [C_out, C_in, k] (fan_in C_in*k); ConvTranspose1d weight [C_in, C_out, k] (fan_in C_out*k)
        shape, fan_in = ((c, co, k), co * k) if transpose else ((co, c, k), c * k)
        _conv(sd, seed, p + ".conv.weight", shape, fan_in, gain)
        if bias:
            _conv(sd, seed, p + ".conv.bias", (co,), fan_in, gain)
        c = co
    return c


def _lstm(sd, seed, prefix, input_size, hidden_size, num_layers, gain):
    # torch.nn.LSTM.reset_parameters: every tensor U(+-1/sqrt(hidden_size)); gate row order i, f, g, o
    for layer in range(num_layers):
        isz = input_size if layer == 0 else 2 * hidden_size
        for sfx in ("", "_reverse"):
            _conv(sd, seed, f"{prefix}.weight_ih_l{layer}{sfx}", (4 * hidden_size, isz), hidden_size, gain)
            _conv(sd, seed, f"{prefix}.weight_hh_l{layer}{sfx}", (4 * hidden_size, hidden_size), hidden_size, gain)
            _conv(sd, seed, f"{prefix}.bias_ih_l{layer}{sfx}", (4 * hidden_size,), hidden_size, gain)
            _conv(sd, seed, f"{prefix}.bias_hh_l{layer}{sfx}", (4 * hidden_size,), hidden_size, gain)


def asr_v2_state_dict(audio_size=64, encoder_settings=ASR_V2_BASE_ENCODER, decoder_num_layers=2,
                      decoder_hidden_size=512, vocab_size=29, seed=1234, randomize_ln=False, gain=1.0):
    """Weights for AudioToAlignText (voice100/models/_asr_v2.py:21-37)."""
    sd: Dict[str, np.ndarray] = {}
    c = _conv_layers_v2(sd, seed, "encoder", audio_size, encoder_settings, randomize_ln, gain)
    assert c == decoder_hidden_size
    _lstm(sd, seed, "lstm", decoder_hidden_size, decoder_hidden_size, decoder_num_layers, gain)
    _conv(sd, seed, "dense.weight", (vocab_size, 2 * decoder_hidden_size), 2 * decoder_hidden_size, gain)
    _conv(sd, seed, "dense.bias", (vocab_size,), 2 * decoder_hidden_size, gain)
    return sd


def align_v2_state_dict(vocab_size=29, num_layers=2, hidden_size=256, num_outputs=2, seed=1234, gain=1.0):
    """Weights for TextToAlignText (voice100/models/_align_v2.py:17-27)."""
    sd: Dict[str, np.ndarray] = {}
    sd["embedding.weight"] = _rng(seed, "embedding.weight").standard_normal(
        (vocab_size, hidden_size)).astype(np.float32)
    _lstm(sd, seed, "lstm", hidden_size, hidden_size, num_layers, gain)
    _conv(sd, seed, "dense.weight", (num_outputs, 2 * hidden_size), 2 * hidden_size, gain)
    _conv(sd, seed, "dense.bias", (num_outputs,), 2 * hidden_size, gain)
    return sd


def audio_v2_state_dict(vocab_size=29, logspc_size=257, codeap_size=1, encoder_num_layers=2,
                        encoder_hidden_size=512, decoder_settings=TTS_V2_BASE_DECODER, seed=1234,
                        randomize_ln=False, randomize_norm=False, gain=1.0):
    """Weights for AlignTextToAudio (voice100/models/_tts_v2.py:13-46)."""
    sd: Dict[str, np.ndarray] = {}
    sd["embedding.weight"] = _rng(seed, "embedding.weight").standard_normal(
        (vocab_size, encoder_hidden_size)).astype(np.float32)
    _lstm(sd, seed, "lstm", encoder_hidden_size, encoder_hidden_size, encoder_num_layers, gain)
    c = _conv_layers_v2(sd, seed, "decoder", 2 * encoder_hidden_size, decoder_settings, randomize_ln, gain)
    audio_size = 2 + logspc_size + 2 * codeap_size
    _conv(sd, seed, "projection.weight", (audio_size, c), c, gain)
    _conv(sd, seed, "projection.bias", (audio_size,), c, gain)
    if randomize_norm:
        sd["norm.f0_std"] = _uniform(seed, "norm.f0_std", (1,), 20.0, 60.0)
        sd["norm.f0_mean"] = _uniform(seed, "norm.f0_mean", (1,), 100.0, 200.0)
        sd["norm.logspc_std"] = _uniform(seed, "norm.logspc_std", (logspc_size,), 0.5, 2.0)
        sd["norm.logspc_mean"] = _uniform(seed, "norm.logspc_mean", (logspc_size,), -8.0, -2.0)
        sd["norm.codeap_std"] = _uniform(seed, "norm.codeap_std", (codeap_size,), 0.5, 2.0)
        sd["norm.codeap_mean"] = _uniform(seed, "norm.codeap_mean", (codeap_size,), -3.0, 0.0)
    else:
        sd["norm.f0_std"] = np.ones((1,), np.float32)
        sd["norm.f0_mean"] = np.zeros((1,), np.float32)
        sd["norm.logspc_std"] = np.ones((logspc_size,), np.float32)
        sd["norm.logspc_mean"] = np.zeros((logspc_size,), np.float32)
        sd["norm.codeap_std"] = np.ones((codeap_size,), np.float32)
        sd["norm.codeap_mean"] = np.zeros((codeap_size,), np.float32)
    return sd


def bn_prefixes(sd: Dict[str, np.ndarray]) -> List[str]:
    """BatchNorm module prefixes of a state dict, in forward order."""
    return [k[: -len(".running_mean")] for k in sd if k.endswith(".running_mean")]


# ----------------------------------------------------------------------------
# inputs
# ----------------------------------------------------------------------------

def noise_waveform(batch: int, samples: int, seed=1234, amp=0.1) -> np.ndarray:
    """`amp * N(0,1)` fp32 audio, the survey's default probe input."""
    return (amp * _rng(seed, f"noise{batch}x{samples}").standard_normal((batch, samples))).astype(np.float32)


def harmonic_waveform(batch: int, samples: int, seed=1234, sample_rate=16000) -> np.ndarray:
    """Speech-like synthetic audio: sum of 1-8 harmonics of a random f0 in [100,1000) Hz with
    amplitudes U*exp(-0.2 i).  Same recipe as the reference's own test fixture
    (tests/test_datasets.py:70-83, make_random_audio), re-expressed on numpy PCG64."""
    g = _rng(seed, f"harm{batch}x{samples}")
    t = np.arange(samples, dtype=np.float64) / sample_rate
    out = np.zeros((batch, samples), np.float64)
    for b in range(batch):
        f0 = 100.0 + 900.0 * g.random()
        n = 1 + int(g.integers(0, 8))
        for i in range(n):
            a = g.random() * np.exp(-0.2 * i)
            out[b] += a * np.sin(2 * np.pi * f0 * (i + 1) * t + 2 * np.pi * g.random())
        out[b] /= max(1.0, np.abs(out[b]).max())
    return out.astype(np.float32)


def ragged_lengths(batch: int, lo: int, hi: int, seed=1234) -> np.ndarray:
    """Seeded clip lengths in samples, U[lo, hi]; the last one is pinned to `hi` so the batch
    maximum is known."""
    ln = _rng(seed, f"len{batch}").integers(lo, hi + 1, size=batch).astype(np.int32)
    ln[-1] = hi
    return ln


def text_tokens(batch: int, length: int, vocab_size=29, seed=1234) -> np.ndarray:
    return _rng(seed, f"text{batch}x{length}").integers(1, vocab_size, size=(batch, length)).astype(np.int64)


def synthetic_alignment(batch: int, length: int, seed=1234) -> np.ndarray:
    """Seeded stand-in for exp(pred)-1 (random-init models give negative gaps):
    gap ~ U[0,1), duration ~ U[0,4) frames per token."""
    g = _rng(seed, f"align{batch}x{length}")
    a = np.empty((batch, length, 2), np.float32)
    a[..., 0] = g.random((batch, length))
    a[..., 1] = 4.0 * g.random((batch, length))
    return a


def viterbi_inputs(T: int, L: int, vocab_size: int = 29, seed=1234):
    """Seeded forced-alignment inputs: fp32 log-probabilities [T, V] (log-softmax of N(0, 2)) and a label
    sequence [L] of non-blank ids with some immediate repeats (the case a naive CTC topology gets wrong)."""
    g = _rng(seed, f"viterbi{T}x{L}")
    z = (2.0 * g.standard_normal((T, vocab_size))).astype(np.float32)
    z = z - z.max(axis=1, keepdims=True)
    lp = (z - np.log(np.exp(z).sum(axis=1, keepdims=True))).astype(np.float32)
    labels = g.integers(1, vocab_size, size=L).astype(np.int64)
    rep = g.random(L) < 0.2
    labels[1:][rep[1:]] = labels[:-1][rep[1:]]
    return lp, labels
