"""The one piece of the reference's feature-side batch processing that is not augmentation noise:
`BatchSpectrogramAugumentation.maskaudio` (voice100/audio.py:106-108), which floors a padded feature batch at
BLANK_AUDIO and blanks every frame at or past an utterance's own length.  The random augmentations of that class
(pitch / amplitude shift, time / frequency masks, noise and utterance mixing) are training-side and out of scope."""
import torch

from . import kernels as K
from .data_modules import LOG_OFFSET


def maskaudio(audio: torch.Tensor, audio_len: torch.Tensor, log_offset: float = LOG_OFFSET) -> torch.Tensor:
    """audio fp32 [B, T, 64] (CUDA), audio_len [B] -> log(clamp(exp(audio) * mask, min=log_offset)),
    mask = t < audio_len[b].  Same call as the reference method; runs `v100_maskaudio`."""
    return K.maskaudio(audio, audio_len, log_offset)
