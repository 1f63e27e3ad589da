"""CPU oracle for the Voice100 batched-inference hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU in fp32, the arithmetic of the reference path named in
BASELINE.json:north_star.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import it; the product package `voice100_b200` never does (it
fails loudly when the CUDA library is missing -- there is no CPU fallback).

Where the reference's arithmetic lives
--------------------------------------
* The reference is pure Python; its conv / batch-norm / ReLU6 / transposed-conv / embedding
  arithmetic is `torch` (pinned torch 1.13.1, poetry.lock:1755-1756; this image has 2.11.0) and
  its STFT + mel arithmetic is `torchaudio.transforms.MelSpectrogram` (pinned torchaudio 0.13.1,
  poetry.lock:1796-1797; this image has 2.11.0).  Neither is vendored under /root/reference.
  The functions below restate the published algorithms with `torch.nn.functional` primitives
  (the same ATen kernels the reference dispatches to on CPU) and, for the front end, with an
  explicit reflect-pad -> frame -> Hann -> rFFT -> |.|^2 -> mel -> log pipeline plus an
  independent numpy version (`logmel_numpy`).
* Nothing here is a module tree copied from the reference: the network is evaluated
  functionally from a flat `state_dict` (reference key layout, see voice100_b200/synth.py).

Pinning
-------
The reference's own tests pin NO numeric result on this path, so the
oracle is pinned against outputs of the UNMODIFIED reference modules run in the build
container: oracle/gen_golden.py imports /root/reference (with the pytorch_lightning stand-in
in oracle/_shim), runs them on seeded inputs and writes tests/golden/*.npz;
tests/test_oracle_golden.py checks every function here against those vectors.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# voice100/data_modules.py:23-26
MELSPEC_DIM = 64
LOG_OFFSET = 1e-6
BLANK_AUDIO = math.log(LOG_OFFSET)  # -13.815510557964274
BN_EPS = 1e-5  # torch.nn.BatchNorm1d default, used by voice100/models/asr.py:36,52

SD = Dict[str, torch.Tensor]


def to_torch_sd(sd) -> SD:
    return {k: (v if isinstance(v, torch.Tensor) else torch.from_numpy(np.asarray(v))) for k, v in sd.items()}


# ----------------------------------------------------------------------------------------------
# Front end: voice100/data_modules.py:262-292 (MelSpectrogramAudioTransform), arithmetic from
# torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_fft=512, win_length=400,
# hop_length=160, n_mels=64) with torchaudio defaults f_min=0, f_max=sr/2, power=2, center=True,
# pad_mode="reflect", window=hann(periodic), norm=None, mel_scale="htk".
# ----------------------------------------------------------------------------------------------

def mel_filterbank(sample_rate=16000, n_fft=512, n_mels=MELSPEC_DIM, f_min=0.0, f_max=None) -> torch.Tensor:
    """torchaudio.functional.melscale_fbanks (htk, norm=None): fb[n_freqs, n_mels], fp32."""
    f_max = float(sample_rate // 2) if f_max is None else f_max
    n_freqs = n_fft // 2 + 1
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + f_min / 700.0)
    m_max = 2595.0 * math.log10(1.0 + f_max / 700.0)
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up))


def mel_power(waveform: torch.Tensor, sample_rate=16000, n_fft=512, win_length=400, hop_length=160,
              n_mels=MELSPEC_DIM) -> torch.Tensor:
    """`MelSpectrogramAudioTransform.melspec(waveform[..., L]) -> [..., n_mels, 1 + L//hop]`
    (voice100/data_modules.py:276-281,290).  Same call sequence as torchaudio's
    Spectrogram/MelScale (functional.spectrogram -> torch.stft(center, reflect) -> |.|^2 ->
    matmul(spec^T, fb)^T), so it is bit-identical to the reference on the same host."""
    x = waveform.to(torch.float32)
    lead = x.shape[:-1]
    x = x.reshape(-1, x.shape[-1])
    spec = torch.stft(x, n_fft=n_fft, hop_length=hop_length, win_length=win_length,
                      window=torch.hann_window(win_length, periodic=True, device=x.device), center=True,
                      pad_mode="reflect", normalized=False, onesided=True, return_complex=True)
    power = spec.abs().pow(2.0)                                    # [N, n_fft/2+1, T]
    fb = mel_filterbank(sample_rate, n_fft, n_mels).to(x.device)    # (device-aware only for bench.py's GPU-eager leg)
    mel = torch.matmul(power.transpose(-1, -2), fb).transpose(-1, -2)
    return mel.reshape(*lead, n_mels, -1)


def mel_power_explicit(waveform: torch.Tensor, sample_rate=16000, n_fft=512, win_length=400,
                       hop_length=160, n_mels=MELSPEC_DIM) -> torch.Tensor:
    """The same transform spelled out step by step (what the CUDA kernel implements): frame t
    covers reflect-padded samples [hop*t, hop*t + n_fft); the periodic Hann(win_length) window
    sits centred in the frame (taps 56..455 for 400/512).  Differs from `mel_power` only by fp32
    FFT round-off (~1e-7 of the frame's peak power)."""
    x = waveform.to(torch.float32)
    lead = x.shape[:-1]
    x = x.reshape(-1, x.shape[-1])
    pad = n_fft // 2
    xp = F.pad(x.unsqueeze(1), (pad, pad), mode="reflect").squeeze(1)
    frames = xp.unfold(-1, n_fft, hop_length)                      # [N, T, n_fft]
    left = (n_fft - win_length) // 2
    win = torch.zeros(n_fft)
    win[left:left + win_length] = torch.hann_window(win_length, periodic=True)
    spec = torch.fft.rfft(frames * win, dim=-1)                    # [N, T, n_fft/2+1]
    power = spec.abs().pow(2.0)
    mel = torch.matmul(power, mel_filterbank(sample_rate, n_fft, n_mels))   # [N, T, n_mels]
    return mel.transpose(-1, -2).reshape(*lead, n_mels, -1)


def logmel_clip(waveform: torch.Tensor, log_offset=LOG_OFFSET, **kw) -> torch.Tensor:
    """One clip: `log(melspec(w).T + log_offset) -> [T, 64]` (voice100/data_modules.py:290-291)."""
    return torch.log(mel_power(waveform, **kw).transpose(-1, -2) + log_offset)


def logmel_batch(waveform: torch.Tensor, lengths: Sequence[int], **kw) -> Tuple[torch.Tensor, torch.Tensor]:
    """Ragged batch the way the reference builds one: features are computed per clip on the
    clip's own samples (reflect at the clip's own end), then padded with BLANK_AUDIO
    (voice100/data_modules.py:446-455, padding_value at :453).  Returns
    (audio[B, T_max, 64], audio_len[B] int32)."""
    feats = [logmel_clip(waveform[i, : int(n)], **kw) for i, n in enumerate(lengths)]
    audio_len = torch.tensor([f.shape[0] for f in feats], dtype=torch.int32)
    audio = torch.nn.utils.rnn.pad_sequence(feats, batch_first=True, padding_value=BLANK_AUDIO)
    return audio, audio_len


def maskaudio(audio: torch.Tensor, audio_len: torch.Tensor, log_offset=LOG_OFFSET) -> torch.Tensor:
    """voice100/audio.py:106-108 (BatchSpectrogramAugumentation.maskaudio): frames at or past an utterance's own length
    become log(log_offset) = BLANK_AUDIO; valid frames are floored there.  audio fp32 [B, T, C], audio_len [B]."""
    mask = (torch.arange(audio.shape[1])[None, :, None] < audio_len[:, None, None]).float()
    return torch.log(torch.clamp(torch.exp(audio) * mask, min=log_offset))


def logmel_numpy(waveform: np.ndarray, sample_rate=16000, n_fft=512, win_length=400, hop_length=160,
                 n_mels=MELSPEC_DIM, log_offset=LOG_OFFSET) -> np.ndarray:
    """Independent numpy restatement of the same front end for ONE clip -> [T, n_mels] fp32.
    Uses float64 internally; agrees with the fp32 pipeline to ~1e-5 in the log domain."""
    x = np.asarray(waveform, np.float64)
    L = x.shape[0]
    pad = n_fft // 2
    idx = np.arange(-pad, L + pad)
    idx = np.where(idx < 0, -idx, idx)
    idx = np.where(idx >= L, 2 * (L - 1) - idx, idx)
    xp = x[idx]
    T = 1 + L // hop_length
    left = (n_fft - win_length) // 2
    win = np.zeros(n_fft)
    win[left:left + win_length] = 0.5 * (1.0 - np.cos(2.0 * np.pi * np.arange(win_length) / win_length))
    frames = np.stack([xp[hop_length * t: hop_length * t + n_fft] for t in range(T)]) * win
    power = np.abs(np.fft.rfft(frames, axis=-1)) ** 2
    fb = mel_filterbank(sample_rate, n_fft, n_mels).numpy().astype(np.float64)
    return np.log(power @ fb + log_offset).astype(np.float32)


# ----------------------------------------------------------------------------------------------
# Model blocks: voice100/models/asr.py:27-59
# ----------------------------------------------------------------------------------------------

def _bn_eval(x: torch.Tensor, sd: SD, p: str, calib: Optional[dict]) -> torch.Tensor:
    """BatchNorm1d in eval mode: gamma*(x-mu_run)/sqrt(var_run+eps)+beta (asr.py:36,52).
    With `calib` set, first overwrite the running stats with this batch's statistics (what one
    training-mode forward with momentum=1 would store: biased->unbiased variance), so randomly
    initialised networks keep O(1) activations through all nine blocks."""
    if calib is not None:
        n = x.shape[0] * x.shape[2]
        sd[p + ".running_mean"] = x.mean(dim=(0, 2)).clone()
        sd[p + ".running_var"] = (x.var(dim=(0, 2), unbiased=False) * (n / max(1, n - 1))).clone()
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"],
                        sd[p + ".bias"], training=False, eps=BN_EPS)


def conv_bn_relu6(x, sd, p, kernel_size, stride=1, groups=1, calib=None):
    """ConvBNActivate (asr.py:27-37): Conv1d(bias=False, padding=(k-1)//2) -> BN -> ReLU6."""
    y = F.conv1d(x, sd[p + ".0.weight"], None, stride=stride, padding=(kernel_size - 1) // 2, groups=groups)
    return F.hardtanh(_bn_eval(y, sd, p + ".1", calib), 0.0, 6.0)


def inverted_residual(x, sd, p, kernel_size, stride=1, use_residual=True, calib=None):
    """InvertedResidual (asr.py:40-59): 1x1 expand (x4) -> depthwise k -> 1x1 project (+BN, linear),
    `x + conv(x)` iff use_residual."""
    h = sd[p + ".conv.0.0.weight"].shape[0]
    y = conv_bn_relu6(x, sd, p + ".conv.0", 1, calib=calib)
    y = conv_bn_relu6(y, sd, p + ".conv.1", kernel_size, stride=stride, groups=h, calib=calib)
    y = F.conv1d(y, sd[p + ".conv.2.weight"])
    y = _bn_eval(y, sd, p + ".conv.3", calib)
    return x + y if use_residual else y


def _asr_blocks(sd: SD):
    half = sd["encoder.layers.0.conv.2.weight"].shape[0]
    ks = (11, 19, 27, 35, 51, 59, 67, 75, 83)                       # asr.py:67-76
    return [(k, 2 if i == 0 else 1, i not in (0, 4, 8)) for i, k in enumerate(ks)]


def asr_encoder(x_ncw: torch.Tensor, sd: SD, calib=None) -> torch.Tensor:
    """ConvVoiceEncoder.forward on [B, 64, T] -> [B, embed, (T+1)//2] (asr.py:62-79)."""
    for i, (k, s, r) in enumerate(_asr_blocks(sd)):
        x_ncw = inverted_residual(x_ncw, sd, f"encoder.layers.{i}", k, s, r, calib)
    return x_ncw


def asr_forward(audio: torch.Tensor, sd: SD, calib=None) -> torch.Tensor:
    """AudioToTextCTC.forward(audio[B,T,64]) -> logits[B,(T+1)//2,V] (asr.py:110-116).
    Dropout(0.2) in LinearCharDecoder (asr.py:89-91) is the identity in eval mode."""
    x = asr_encoder(audio.transpose(1, 2), sd, calib)
    logits = F.conv1d(x, sd["decoder.layers.1.weight"], sd["decoder.layers.1.bias"])
    return logits.transpose(1, 2)


def _q(t: torch.Tensor, dtype) -> torch.Tensor:
    return t if dtype is None else t.to(dtype).float()


def _fold(sd: SD, p: str):
    scale = sd[p + ".weight"] / torch.sqrt(sd[p + ".running_var"] + BN_EPS)
    return scale[None, :, None], (sd[p + ".bias"] - sd[p + ".running_mean"] * sd[p + ".weight"] /
                                  torch.sqrt(sd[p + ".running_var"] + BN_EPS))[None, :, None]


def inverted_residual_storage_model(x, sd, p, kernel_size, stride, use_residual, dtype):
    """The same block with every tensor that the CUDA path STORES rounded to `dtype` (weights, the two
    hidden activations, the block output) while all arithmetic in between stays fp32 (fp32 accumulation,
    fp32 folded BN, fp32 residual add before the output rounding).  This is the numerical contract of
    libv100; comparing the kernels with it separates 'bf16 storage error' (inherent, stated in the tests)
    from implementation error (must be ~0)."""
    s1, b1 = _fold(sd, p + ".conv.0.1")
    s2, b2 = _fold(sd, p + ".conv.1.1")
    s3, b3 = _fold(sd, p + ".conv.3")
    h = F.conv1d(x, _q(sd[p + ".conv.0.0.weight"], dtype))
    h = _q((h * s1 + b1).clamp(0.0, 6.0), dtype)
    h = F.conv1d(h, _q(sd[p + ".conv.1.0.weight"], dtype), stride=stride, padding=(kernel_size - 1) // 2,
                 groups=h.shape[1])
    h = _q((h * s2 + b2).clamp(0.0, 6.0), dtype)
    y = F.conv1d(h, _q(sd[p + ".conv.2.weight"], dtype)) * s3 + b3
    return _q(x + y if use_residual else y, dtype)


def asr_forward_storage_model(audio: torch.Tensor, sd: SD, dtype=torch.bfloat16) -> torch.Tensor:
    """asr_forward with libv100's storage roundings (features, weights, activations in `dtype`)."""
    x = _q(audio.transpose(1, 2), dtype)
    for i, (k, s, r) in enumerate(_asr_blocks(sd)):
        x = inverted_residual_storage_model(x, sd, f"encoder.layers.{i}", k, s, r, dtype)
    logits = F.conv1d(x, _q(sd["decoder.layers.1.weight"], dtype), sd["decoder.layers.1.bias"])
    return logits.transpose(1, 2)


def asr_output_length(audio_len: torch.Tensor) -> torch.Tensor:
    """asr.py:81-82,118-122."""
    return torch.div(audio_len + 1, 2, rounding_mode="trunc")


def ctc_greedy(logits: torch.Tensor) -> torch.Tensor:
    """tests/test_onnx.py:40: `logits.argmax(-1)` -> int64 [B, T]."""
    return logits.argmax(-1)


DEFAULT_CHARACTERS = "_ abcdefghijklmnopqrstuvwxyz'"   # voice100/text.py:14


def ctc_collapse_text(tokens: Sequence[int], vocab: str = DEFAULT_CHARACTERS) -> str:
    """CharTokenizer.decode + merge_repeated (voice100/text.py:93-104): drop ids outside the
    vocab, collapse runs of the same character, drop the blank '_' and a lone space."""
    import re
    text = "".join(vocab[int(t)] for t in tokens if 0 <= int(t) < len(vocab))
    text = re.sub(r"(.)\1+", r"\1", text).replace("_", "")
    return "" if text == " " else text


# ----------------------------------------------------------------------------------------------
# TTS: voice100/models/tts.py:13-29,67-110,152-201 and voice100/models/_layers_v1.py:96-138
# ----------------------------------------------------------------------------------------------

def align_forward(text: torch.Tensor, sd: SD, calib=None) -> torch.Tensor:
    """TextToAlignTextModel.forward(text[B,L] int64) -> [B,L,2] = log(align+1) (tts.py:79-87)."""
    x = F.embedding(text, sd["embedding.weight"]).transpose(1, 2)
    for i, k in enumerate((5, 11, 17, 29)):
        x = inverted_residual(x, sd, f"layers.{i}", k, 1, True, calib)
    x = F.conv1d(x, sd["layers.4.weight"], sd["layers.4.bias"])
    return x.transpose(1, 2)


def align_text(text: Sequence[int], align: np.ndarray, head=5, tail=5) -> np.ndarray:
    """TextToAlignTextModel.align for one utterance (tts.py:89-110): token i occupies frames
    [round(t+gap_i), round(t+gap_i+dur_i)), at least one frame; python round() = half-to-even;
    total length head + int(sum(align)) + tail; writes past the end are an IndexError in the
    reference, so inputs here must not produce them."""
    a = torch.as_tensor(np.asarray(align))
    n = head + int(torch.sum(a)) + tail
    out = np.zeros((n,), dtype=np.int64)
    t = head
    for i in range(a.shape[0]):
        t += a[i, 0].item()
        s = round(t)
        t += a[i, 1].item()
        e = round(t)
        if s == e:
            e = max(0, e + 1)
        out[s:e] = int(text[i])
    return out


def voice_decoder(x: torch.Tensor, sd: SD, p="decoder", calib=None) -> torch.Tensor:
    """VoiceDecoder.forward [B,H,T] -> [B,260,2T-1] (tts.py:13-29)."""
    for i, k in enumerate((65, 33, 17, 11)):
        x = inverted_residual(x, sd, f"{p}.layers.{i}", k, 1, True, calib)
    x = F.conv_transpose1d(x, sd[f"{p}.layers.4.weight"], sd[f"{p}.layers.4.bias"], stride=2, padding=2)
    for j, k in enumerate((33, 11, 7)):
        x = inverted_residual(x, sd, f"{p}.layers.{5 + j}", k, 1, True, calib)
    return F.conv1d(x, sd[f"{p}.layers.8.weight"], sd[f"{p}.layers.8.bias"])


def audio_forward(aligntext: torch.Tensor, sd: SD, calib=None):
    """AlignTextToAudioModel.forward -> (hasf0_logits[B,T'], f0_hat[B,T'], logspc_hat[B,T',S],
    codeap_hat[B,T',1]), T' = 2T-1 (tts.py:172-190); S = 257 bins, or 25 mel-cepstra with use_mcep (tts.py:164),
    read off the head's width."""
    x = F.embedding(aligntext, sd["embedding.weight"]).transpose(1, 2)
    x = voice_decoder(x, sd, "decoder", calib).transpose(1, 2)
    hasf0, f0, logspc, codeap = torch.split(x, [1, 1, x.shape[2] - 3, 1], dim=2)
    return hasf0[:, :, 0], f0[:, :, 0], logspc, codeap


def audio_predict(aligntext: torch.Tensor, sd: SD):
    """AlignTextToAudioModel.predict (tts.py:192-201): WORLDNorm.unnormalize
    (_layers_v1.py:131-138: std*x+mean) then f0 := 0 where hasf0_logit < 0."""
    hasf0, f0, logspc, codeap = audio_forward(aligntext, sd)
    f0 = sd["norm.f0_std"] * f0 + sd["norm.f0_mean"]
    logspc = sd["norm.logspc_std"] * logspc + sd["norm.logspc_mean"]
    codeap = sd["norm.codeap_std"] * codeap + sd["norm.codeap_mean"]
    f0 = torch.where(hasf0 < 0, torch.zeros((1,), dtype=f0.dtype), f0)
    return f0, logspc, codeap


# ----------------------------------------------------------------------------------------------
# v2 models: voice100/models/_layers_v2.py:29-103, _asr_v2.py:21-49, _align_v2.py:17-48, _tts_v2.py:13-91
# ----------------------------------------------------------------------------------------------
LN_EPS = 1e-5  # torch.nn.LayerNorm default (_layers_v2.py:40,71)


def conv_layer_block(x: torch.Tensor, sd: SD, p: str, transpose: bool, stride: int, padding: int) -> torch.Tensor:
    """ConvLayerBlock / ConvTransposeLayerBlock.forward on [B,C,T] (_layers_v2.py:51-57,82-88):
    conv -> LayerNorm over channels -> exact (erf) GELU."""
    w, b = sd[p + ".conv.weight"], sd.get(p + ".conv.bias")
    x = (F.conv_transpose1d if transpose else F.conv1d)(x, w, b, stride=stride, padding=padding)
    x = F.layer_norm(x.transpose(1, 2), (x.shape[1],), sd[p + ".layer_norm.weight"], sd[p + ".layer_norm.bias"],
                     LN_EPS).transpose(1, 2)
    return F.gelu(x)


def conv_layers_v2(x: torch.Tensor, sd: SD, prefix: str, settings) -> torch.Tensor:
    """get_conv_layers(...) as a function (_layers_v2.py:91-103)."""
    for i, (_co, transpose, _k, stride, padding, _bias) in enumerate(settings):
        x = conv_layer_block(x, sd, f"{prefix}.{i}", bool(transpose), stride, padding)
    return x


def lstm_bidirectional(x: torch.Tensor, lengths: Sequence[int], sd: SD, prefix: str, num_layers: int) -> torch.Tensor:
    """torch.nn.LSTM(bidirectional=True) in eval mode over a packed batch, written out step by step
    (the reference calls it through pack_padded_sequence/pad_packed_sequence: _asr_v2.py:45-47,
    _align_v2.py:39-41, _tts_v2.py:56-58).  x [B,T,I] -> [B,T,2H]; every utterance runs over its own
    length (the backward direction starts at its last valid step with zero state), rows past the
    length are zero.  Gate order i, f, g, o; c' = f*c + i*g; h' = o*tanh(c')."""
    B, T, _ = x.shape
    lens = torch.as_tensor(list(lengths), dtype=torch.long)
    inp = x
    for layer in range(num_layers):
        outs = []
        for sfx in ("", "_reverse"):
            w_ih, w_hh = sd[f"{prefix}.weight_ih_l{layer}{sfx}"], sd[f"{prefix}.weight_hh_l{layer}{sfx}"]
            bias = sd[f"{prefix}.bias_ih_l{layer}{sfx}"] + sd[f"{prefix}.bias_hh_l{layer}{sfx}"]
            H = w_hh.shape[1]
            gx = inp @ w_ih.T + bias  # [B,T,4H]
            h = torch.zeros(B, H, dtype=x.dtype)
            c = torch.zeros(B, H, dtype=x.dtype)
            out = torch.zeros(B, T, H, dtype=x.dtype)
            steps = range(T - 1, -1, -1) if sfx else range(T)
            for t in steps:
                a = gx[:, t] + h @ w_hh.T
                i, f, g, o = a[:, :H], a[:, H:2 * H], a[:, 2 * H:3 * H], a[:, 3 * H:]
                c_new = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
                h_new = torch.sigmoid(o) * torch.tanh(c_new)
                live = (t < lens)[:, None]
                c = torch.where(live, c_new, torch.zeros_like(c))
                h = torch.where(live, h_new, torch.zeros_like(h))
                out[:, t] = h
            outs.append(out)
        inp = torch.cat(outs, dim=2)
    return inp


def asr_v2_forward(audio: torch.Tensor, audio_len: Sequence[int], sd: SD, encoder_settings, num_layers=2):
    """AudioToAlignText.forward (_asr_v2.py:39-49): audio [B,T,64] -> (logits [T', B, V] time-major,
    lengths (audio_len+1)//2), T' = max length.  Padding frames take part in the convolutions exactly as
    in the reference (no masking before the LSTM)."""
    x = conv_layers_v2(audio.transpose(1, 2), sd, "encoder", encoder_settings).transpose(1, 2)
    x_len = [(int(n) + 1) // 2 for n in audio_len]
    x = lstm_bidirectional(x, x_len, sd, "lstm", num_layers)[:, :max(x_len)]
    y = x @ sd["dense.weight"].T + sd["dense.bias"]
    return y.transpose(0, 1).contiguous(), torch.as_tensor(x_len)


def align_v2_forward(text: torch.Tensor, text_len: Sequence[int], sd: SD, num_layers=2):
    """TextToAlignText.forward (_align_v2.py:29-43): text [B,L] -> ([B, max_len, 2], lengths)."""
    x = F.embedding(text, sd["embedding.weight"])
    x = lstm_bidirectional(x, text_len, sd, "lstm", num_layers)[:, :max(int(n) for n in text_len)]
    return x @ sd["dense.weight"].T + sd["dense.bias"], torch.as_tensor([int(n) for n in text_len])


def align_text_v2(text: Sequence[int], align: np.ndarray, head=5, tail=5) -> np.ndarray:
    """TextToAlignText.align (_align_v2.py:54-82): truncating (int()) frame positions, a gap before every token
    but the first, at least one blank frame between tokens, total head + int(sum(align) - align[0,0]) + tail."""
    a = torch.as_tensor(np.asarray(align))
    n = head + int(torch.sum(a) - a[0, 0]) + tail
    out = np.zeros((n,), dtype=np.int64)
    t, u = head, 0
    for i in range(a.shape[0]):
        if i > 0:
            t += a[i, 0].item()
        s = max(int(t), u)
        u = s + 1
        t += a[i, 1].item()
        e = max(int(t), u)
        u = e
        out[s:e] = int(text[i])
    return out


def audio_v2_forward(aligntext: torch.Tensor, aligntext_len: Sequence[int], sd: SD, decoder_settings,
                     num_layers=2, logspc_size=257, codeap_size=1):
    """AlignTextToAudio.forward (_tts_v2.py:48-78) -> (hasf0_logits[B,T'], f0_hat[B,T'], logspc_hat[B,T',S],
    hascodeap_logits[B,T',A], codeap_hat[B,T',A])."""
    x = F.embedding(aligntext, sd["embedding.weight"])
    x = lstm_bidirectional(x, aligntext_len, sd, "lstm", num_layers)[:, :max(int(n) for n in aligntext_len)]
    x = conv_layers_v2(x.transpose(1, 2), sd, "decoder", decoder_settings).transpose(1, 2)
    x = x @ sd["projection.weight"].T + sd["projection.bias"]
    hasf0, f0, logspc, hascodeap, codeap = torch.split(x, [1, 1, logspc_size, codeap_size, codeap_size], dim=2)
    return hasf0[:, :, 0], f0[:, :, 0], logspc, hascodeap, codeap


def audio_v2_predict(aligntext, aligntext_len, sd: SD, decoder_settings, **kw):
    """AlignTextToAudio.predict (_tts_v2.py:80-91): unnormalize (_layers_v2.py:199-206), f0 := 0 where
    hasf0 < 0, codeap := 0 where hascodeap < 0."""
    hasf0, f0, logspc, hascodeap, codeap = audio_v2_forward(aligntext, aligntext_len, sd, decoder_settings, **kw)
    f0 = sd["norm.f0_std"] * f0 + sd["norm.f0_mean"]
    logspc = sd["norm.logspc_std"] * logspc + sd["norm.logspc_mean"]
    codeap = sd["norm.codeap_std"] * codeap + sd["norm.codeap_mean"]
    f0 = torch.where(hasf0 < 0, torch.zeros((1,), dtype=f0.dtype), f0)
    codeap = torch.where(hascodeap < 0, torch.zeros((1, 1), dtype=codeap.dtype), codeap)
    return f0, logspc, codeap


# ----------------------------------------------------------------------------------------------
# mel-cepstrum -> log spectrum: voice100/vocoder.py:115-145 (create_mc2sp_matrix, freqt; PySPTK conventions),
# applied after predict() by the export wrapper (voice100/export_onnx.py:81-97) and the data module
# (voice100/data_modules.py:229-231) when the model was trained on 25 mel-cepstral coefficients.
# ----------------------------------------------------------------------------------------------

def freqt_vector(c: np.ndarray, out_order: int, alpha: float) -> np.ndarray:
    """SPTK `freqt` on one cepstrum vector: all-pass frequency warping by `alpha` (Oppenheim recursion,
    float64).  Input taken last coefficient first; g' = alpha*g shifted by the recurrence
    g'[0] = c_i + a g[0];  g'[1] = (1 - a^2) g[0] + a g[1];  g'[j] = g[j-1] + a (g[j] - g'[j-1])."""
    g = np.zeros(out_order + 1)
    for ci in np.asarray(c, np.float64)[::-1]:
        prev = g
        g = np.empty_like(prev)
        g[0] = ci + alpha * prev[0]
        if out_order >= 1:
            g[1] = (1.0 - alpha * alpha) * prev[0] + alpha * prev[1]
        for j in range(2, out_order + 1):
            g[j] = prev[j - 1] + alpha * (prev[j] - g[j - 1])
    return g


def mc2sp_matrix(fftlen=512, order=24, alpha=0.410) -> np.ndarray:
    """[order+1, fftlen/2+1] matrix M with logspc = mcep @ M (vocoder.py:115-123): un-warp each unit cepstrum to
    fftlen/2 + 1 coefficients, double c0, extend to the even sequence [c0..cN, cN..c1] and take the real DFT."""
    n = fftlen // 2
    rows = []
    for k in range(order + 1):
        e = np.zeros(order + 1)
        e[k] = 1.0
        c = freqt_vector(e, n, -alpha)
        c[0] *= 2.0
        rows.append(np.fft.rfft(np.concatenate([c, c[:0:-1]])).real)
    return np.stack(rows)


# ----------------------------------------------------------------------------------------------
# Forced alignment: voice100/models/align.py:18-66 (ctc_best_path, max_move = 3), the per-utterance numpy DP
# behind AudioToAlignText.ctc_best_path (voice100/models/_asr_v2.py:100-119)
# ----------------------------------------------------------------------------------------------

def ctc_best_path(logprob: np.ndarray, labels: np.ndarray):
    """Viterbi best path of `labels` through `logprob[T, V]` over the blank-expanded state sequence
    (blank, l0, blank, l1, ..., blank).  Restated as a dense recurrence over all S = 2L+1 states with -inf
    for states outside the active prefix (which starts at 2 states and grows by 2 per frame):
        cand_j[v] = score[v - j] + logprob[i, lab[v]],  j = 0 (stay), 1 (advance), 2 (skip; not onto a blank)
    first maximum over j wins (np.argmax), ending in the better of the last two states (ties -> S-2).
    -> (best_score, best_path[T] state indices, best_labels[T]); fp32 arithmetic like the reference."""
    logprob = np.asarray(logprob)
    T = logprob.shape[0]
    lab = np.zeros(2 * len(labels) + 1, dtype=np.asarray(labels).dtype)
    lab[1::2] = labels
    S = lab.shape[0]
    if S < 2:
        raise IndexError("empty text (the reference indexes labels[1], align.py:31)")
    NEG = np.float32(-np.inf)
    score = np.full(S, NEG, dtype=logprob.dtype)
    score[:2] = logprob[0, lab[:2]]
    active = 2
    back = np.zeros((T, S), dtype=np.int32)
    for i in range(1, T):
        nxt_active = min(active + 2, S)
        emit = logprob[i, lab]
        cands = np.full((3, S), NEG, dtype=logprob.dtype)
        srcs = np.zeros((3, S), dtype=np.int32)
        for j in range(3):
            v = np.arange(j, min(active + j, S))
            cands[j, v] = score[v - j] + emit[v]
            srcs[j, v] = v - j
        cands[2, lab == 0] = NEG
        pick = np.argmax(cands[:, :nxt_active], axis=0)
        cols = np.arange(nxt_active)
        score = np.full(S, NEG, dtype=logprob.dtype)
        score[:nxt_active] = cands[pick, cols]
        back[i, :nxt_active] = srcs[pick, cols]
        active = nxt_active
    # align.py:57-58 works on the `active` live scores: j = S + (-1 if scores[-1] > scores[-2] else -2); scores[j]
    # raises IndexError when j >= active, i.e. always with two or more states missing, and with exactly one state
    # missing only if the last live score is the larger one
    j = S - 1 if score[active - 1] > score[active - 2] else S - 2
    if j >= active:
        raise IndexError("too few frames to reach the end of the text (same failure as the reference)")
    best = score[j]
    path = np.zeros(T, dtype=np.int32)
    for i in range(T - 1, -1, -1):
        path[i] = j
        j = back[i, j]
    return best, path, lab[path]


# ----------------------------------------------------------------------------------------------
# helpers shared by tests / bench
# ----------------------------------------------------------------------------------------------

def calibrate_asr(sd: SD, audio: torch.Tensor) -> SD:
    """Return a copy of `sd` whose BN running stats are those of `audio` (see _bn_eval)."""
    sd = {k: v.clone() for k, v in sd.items()}
    with torch.no_grad():
        asr_forward(audio, sd, calib={})
    return sd


def calibrate_align(sd: SD, text: torch.Tensor) -> SD:
    sd = {k: v.clone() for k, v in sd.items()}
    with torch.no_grad():
        align_forward(text, sd, calib={})
    return sd


def calibrate_audio(sd: SD, aligntext: torch.Tensor) -> SD:
    sd = {k: v.clone() for k, v in sd.items()}
    with torch.no_grad():
        audio_forward(aligntext, sd, calib={})
    return sd


def parity_report(ref: torch.Tensor, got: torch.Tensor) -> Dict[str, float]:
    """max-abs and RMS error normalised by the reference's std."""
    ref = ref.double().flatten()
    got = got.double().flatten()
    std = float(ref.std()) or 1.0
    d = (ref - got).abs()
    return {"max_abs": float(d.max()), "rms": float(d.pow(2).mean().sqrt()), "ref_std": std,
            "max_abs_rel_std": float(d.max()) / std, "rms_rel_std": float(d.pow(2).mean().sqrt()) / std}


def token_agreement(ref_logits: torch.Tensor, got_tokens: torch.Tensor, margin: float):
    """CTC greedy agreement: (raw, gated, gated_fraction).  `gated` counts only frames whose
    fp32 top-1/top-2 logit margin exceeds `margin` (ties below bf16 resolution are not
    decidable by any bf16 implementation)."""
    top2 = ref_logits.float().topk(2, dim=-1).values
    gate = (top2[..., 0] - top2[..., 1]) > margin
    same = ref_logits.argmax(-1) == got_tokens
    raw = float(same.float().mean())
    gated = float(same[gate].float().mean()) if gate.any() else 1.0
    return raw, gated, float(gate.float().mean())
