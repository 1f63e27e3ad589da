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
                        text_len: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """logprob fp32 [B, T, V] (log-softmax of the logits), logit_len [B], text int64 [B, L] (no blanks inside),
    text_len [B] -> (score fp32 [B], hist int32 [B, T], path int64 [B, T], logit_len).  Utterances with too few
    frames for their text (an IndexError in the reference) come back with score NaN and hist -1."""
    if not logprob.is_cuda:
        raise V100Error("ctc_best_path_batch runs only on CUDA tensors (no CPU path)")
    score, hist, path = K.ctc_best_path(logprob.float().contiguous(), logit_len, text.contiguous(), text_len)
    return score, hist, path, logit_len
