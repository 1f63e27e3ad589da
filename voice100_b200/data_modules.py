"""Log-mel front end with the reference's interface (voice100/data_modules.py:23-26,262-292).

`MelSpectrogramAudioTransform` keeps the reference's constructor arguments, `audio_size` property and
`.melspec(waveform[..., L]) -> [..., n_mels, 1 + L//hop]` call; the arithmetic is one fused CUDA kernel
(v100_logmel) instead of torchaudio's reflect-pad + cuFFT + sgemm chain.  `logmel_batch` is the batched
form of what `EncodedCacheDataset` + `generate_audio_text_batch` (data_modules.py:210-234,446-455) build
per utterance on the CPU: per-clip features padded with BLANK_AUDIO.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np
import torch
from torch import nn

from . import kernels as K
from ._lib import V100Error

MELSPEC_DIM = 64
LOG_OFFSET = 1e-6
BLANK_AUDIO = math.log(LOG_OFFSET)


def _hz_to_mel_htk(f: float) -> float:
    return 2595.0 * math.log10(1.0 + f / 700.0)


def mel_filterbank(sample_rate: int, n_fft: int, n_mels: int) -> np.ndarray:
    """HTK triangular filters, norm=None, f_min=0, f_max=sr/2 -> fb[n_fft//2+1, n_mels] fp32.  A host-side
    constant table; built with the same fp32 torch ops, in the same order, as
    torchaudio.functional.melscale_fbanks so the weights are bit-identical to the reference's."""
    n_freqs = n_fft // 2 + 1
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_pts = torch.linspace(_hz_to_mel_htk(0.0), _hz_to_mel_htk(float(sample_rate // 2)), n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up)).numpy()


def sparse_filterbank(fb: np.ndarray):
    """Column-compressed form consumed by v100_logmel: per filter (first bin, count, offset) + weights."""
    start, count, off, w = [], [], [], []
    for m in range(fb.shape[1]):
        nz = np.nonzero(fb[:, m])[0]
        s, e = (int(nz[0]), int(nz[-1]) + 1) if len(nz) else (0, 0)
        start.append(s); count.append(e - s); off.append(len(w))
        w.extend(fb[s:e, m].tolist())
    return (np.asarray(start, np.int32), np.asarray(count, np.int32), np.asarray(off, np.int32),
            np.asarray(w, np.float32))


class MelSpectrogramAudioTransform(nn.Module):
    def __init__(self, sample_rate: int = 16000, n_fft: int = 512, win_length: int = 400, hop_length: int = 160,
                 n_mels: int = MELSPEC_DIM, log_offset: float = LOG_OFFSET) -> None:
        super().__init__()
        # (512, 400, 160, 64) -- the only configuration the reference ever builds, data_modules.py:266-269 -- runs the tuned
        # kernel; any other one the generic kernel (v100_logmel_generic)
        if n_fft < 8 or n_fft > 2048 or (n_fft & (n_fft - 1)) != 0:
            raise V100Error(f"n_fft={n_fft}: powers of two in [8, 2048] are supported")
        if not (0 < win_length <= n_fft) or hop_length <= 0 or n_mels <= 0:
            raise V100Error("need 0 < win_length <= n_fft, hop_length > 0, n_mels > 0")
        self.sample_rate, self.n_fft, self.win_length, self.hop_length = sample_rate, n_fft, win_length, hop_length
        self.n_mels, self.log_offset = n_mels, log_offset
        fb = mel_filterbank(sample_rate, n_fft, n_mels)
        for name, arr in zip(("fb_start", "fb_count", "fb_off", "fb_w"), sparse_filterbank(fb)):
            self.register_buffer(name, torch.from_numpy(arr), persistent=False)

    @property
    def audio_size(self) -> int:
        return self.n_mels

    @property
    def _config(self):
        return (self.n_fft, self.win_length, self.hop_length, self.n_mels)

    def _fb(self, device):
        if self.fb_w.device != device:
            self.to(device)
        return (self.fb_start, self.fb_count, self.fb_off, self.fb_w)

    def num_frames(self, num_samples):
        return 1 + num_samples // self.hop_length

    def melspec(self, waveform: torch.Tensor) -> torch.Tensor:
        """Mel power spectrogram, `[..., L] -> [..., 64, 1 + L//160]` fp32 (what the reference's
        `self.melspec` torchaudio module returns, data_modules.py:290)."""
        if not waveform.is_cuda:
            raise V100Error("MelSpectrogramAudioTransform runs only on CUDA tensors (no CPU path)")
        lead, L = waveform.shape[:-1], waveform.shape[-1]
        if L <= self.n_fft // 2:
            raise V100Error(f"clips must be longer than {self.n_fft // 2} samples (reflect padding)")
        wav = self._samples(waveform.reshape(-1, L))
        lengths = torch.full((wav.shape[0],), L, dtype=torch.int32, device=wav.device)
        T = self.num_frames(L)
        out, _ = K.logmel(wav, lengths, self._fb(wav.device), self.log_offset, T, K.MEL_POWER_F32_NCW, self._config)
        return out.valid().reshape(*lead, self.n_mels, T)

    @staticmethod
    def _samples(waveform: torch.Tensor) -> torch.Tensor:
        """fp32 samples, or int16 PCM passed through untouched (the kernel scales by 1/32768, which is what
        torchaudio.load returns for a 16-bit WAV -- data_modules.py:288 -- so both forms give identical features)."""
        if waveform.dtype == torch.int16:
            return waveform.contiguous()
        return waveform.to(torch.float32).contiguous()

    def logmel_batch(self, waveform: torch.Tensor, lengths: torch.Tensor, ncw_bf16: bool = False, ncw_dtype=None):
        """waveform fp32 (or int16 PCM) [B, L_max], lengths [B] samples ->
        (audio fp32 [B, T_max, 64] padded with BLANK_AUDIO  |  16-bit Ncw [B, 64, pitch] when `ncw_dtype`
         (torch.bfloat16 / torch.float16) or `ncw_bf16` is given,  audio_len int32 [B] = 1 + len // 160).
        Host-resident `lengths` are validated like torchaudio does (reflect padding needs > n_fft/2 samples, 0 =
        an empty filler slot); device-resident ones are clamped to [0, L_max] by the kernel."""
        if ncw_bf16 and ncw_dtype is None:
            ncw_dtype = torch.bfloat16
        if not waveform.is_cuda:
            raise V100Error("MelSpectrogramAudioTransform runs only on CUDA tensors (no CPU path)")
        wav = self._samples(waveform)
        if not lengths.is_cuda and lengths.numel():
            lo, hi = int(lengths.min()), int(lengths.max())
            if hi > wav.shape[1] or ((lengths > 0) & (lengths <= self.n_fft // 2)).any():
                raise V100Error(f"clip lengths must be 0 (empty slot) or in ({self.n_fft // 2}, {wav.shape[1]}]; got [{lo}, {hi}]")
        lengths = lengths.to(device=wav.device, dtype=torch.int32).contiguous()
        T = self.num_frames(wav.shape[1])
        mode = K.MEL_LOG_F32_NTC if ncw_dtype is None else (K.MEL_LOG_F16_NCW if K.dt(ncw_dtype) == 1 else K.MEL_LOG_BF16_NCW)
        return K.logmel(wav, lengths, self._fb(wav.device), self.log_offset, T, mode, self._config)

    def forward(self, waveform: torch.Tensor) -> torch.Tensor:
        """One clip `[L] -> [T, 64]` log-mel features.  (The reference's forward takes a file path and does
        load + resample first, data_modules.py:287-289; file I/O is outside the accelerated path.)"""
        return torch.log(self.melspec(waveform).transpose(-1, -2) + self.log_offset)
