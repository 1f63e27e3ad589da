"""Batched CTC forced alignment on the device.

`ctc_best_path_batch` is the batched form of what `AudioToAlignText.ctc_best_path` (voice100/models/_asr_v2.py:
100-119) does one utterance at a time through `.cpu().numpy()` and the numpy DP in voice100/models/align.py:18-66.
It returns the same four things: (score, hist = state index per frame, path = expanded label per frame, lengths).
"""
from __future__ import annotations

from typing import Tuple

import torch

from . import kernels as K
from ._lib import V100Error


def ctc_best_path_batch(logprob: torch.Tensor, logit_len: torch.Tensor, text: torch.Tensor,
                        text_len: torch.Tensor, normalize: bool = False
                        ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """logprob fp32 [B, T, V] (log-softmax of the logits; or the raw logits with normalize=True, the kernel then
    applies log_softmax itself), logit_len [B], text int64 [B, L] (no blanks inside), text_len [B] ->
    (score fp32 [B], hist int32 [B, T], path int64 [B, T], logit_len).  Utterances for which the reference raises
    IndexError (too few frames for the text, empty text, a label outside [0, V)) come back with score NaN and
    hist -1; host-resident `text` is additionally range-checked here and raises like numpy would."""
    if not logprob.is_cuda:
        raise V100Error("ctc_best_path_batch runs only on CUDA tensors (no CPU path)")
    if not text.is_cuda and text.numel() and (int(text.min()) < 0 or int(text.max()) >= logprob.shape[2]):
        raise IndexError(f"text label outside [0, {logprob.shape[2]})")
    score, hist, path = K.ctc_best_path(logprob.float().contiguous(), logit_len, text.to(logprob.device).contiguous(),
                                        text_len, normalize)
    return score, hist, path, logit_len
