"""Loading the reference's checkpoints into the drop-in modules.

A Lightning `.ckpt` written by the reference's trainers (ModelCheckpoint, voice100/train_asr.py:33) is a
pickled dict with `state_dict` (the key layout of section 8b) and `hyper_parameters` (what
`save_hyperparameters()` captured, voice100/models/asr.py:101, tts.py:70,157).  WORLD statistics may also
come as a separate plain `state_dict` file (voice100/calc_stat.py:59-68 -> tts.py:258-261)."""
from __future__ import annotations

from typing import Optional

import torch

from ._lib import V100Error
from .asr import AudioToTextCTC
from .tts import AlignTextToAudioModel, TextToAlignTextModel

_CLASSES = {
    "AudioToTextCTC": (AudioToTextCTC, ("audio_size", "embed_size", "vocab_size", "hidden_size")),
    "TextToAlignTextModel": (TextToAlignTextModel, ("vocab_size", "hidden_size")),
    "AlignTextToAudioModel": (AlignTextToAudioModel, ("vocab_size", "hidden_size")),
}
# keys that belong to training-only sub-modules of the reference classes
_IGNORED_PREFIXES = ("criterion.", "batch_augment.")


def _guess_class(state_dict) -> str:
    keys = state_dict.keys()
    if any(k.startswith("encoder.layers.") for k in keys):
        return "AudioToTextCTC"
    if any(k.startswith("decoder.layers.") for k in keys) and "embedding.weight" in keys:
        return "AlignTextToAudioModel"
    if "embedding.weight" in keys and any(k.startswith("layers.") for k in keys):
        return "TextToAlignTextModel"
    raise V100Error("checkpoint does not look like one of the v1 CNN models this package accelerates "
                    "(AudioToTextCTC / TextToAlignTextModel / AlignTextToAudioModel)")


def _infer_hparams(cls_name: str, sd) -> dict:
    if cls_name == "AudioToTextCTC":
        return dict(audio_size=sd["encoder.layers.0.conv.0.0.weight"].shape[1],
                    hidden_size=sd["encoder.layers.4.conv.2.weight"].shape[0],
                    embed_size=sd["encoder.layers.8.conv.2.weight"].shape[0],
                    vocab_size=sd["decoder.layers.1.weight"].shape[0])
    return dict(vocab_size=sd["embedding.weight"].shape[0], hidden_size=sd["embedding.weight"].shape[1])


def load_checkpoint(path: str, model_class: Optional[str] = None, audio_stat: Optional[str] = None,
                    device: str = "cuda", storage_dtype=torch.bfloat16):
    """-> an eval-mode drop-in module with the checkpoint's weights on `device`."""
    ckpt = torch.load(path, map_location="cpu", weights_only=True)
    sd = ckpt["state_dict"] if isinstance(ckpt, dict) and "state_dict" in ckpt else ckpt
    sd = {k: v for k, v in sd.items() if not k.startswith(_IGNORED_PREFIXES)}
    name = model_class or _guess_class(sd)
    if name not in _CLASSES:
        raise V100Error(f"unknown model class {name!r}; expected one of {sorted(_CLASSES)}")
    cls, arg_names = _CLASSES[name]
    hp = dict(ckpt.get("hyper_parameters", {})) if isinstance(ckpt, dict) else {}
    try:
        inferred = _infer_hparams(name, sd)
    except KeyError as e:
        raise V100Error(f"checkpoint has no {e.args[0]!r}: it is not a {name}") from None
    kwargs = {a: int(hp.get(a, inferred[a])) for a in arg_names}
    for a in arg_names:   # the tensors are authoritative (the legacy CLI stored sizes as floats, asr.py:185-186)
        if kwargs[a] != inferred[a]:
            raise V100Error(f"hyper-parameter {a}={kwargs[a]} disagrees with the weights ({inferred[a]})")
    if name == "AlignTextToAudioModel" and bool(hp.get("use_mcep", False)):
        raise V100Error("use_mcep=True checkpoints are not on the accelerated path")
    model = cls(**kwargs)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    if unexpected or [m for m in missing if "num_batches_tracked" not in m]:
        raise V100Error(f"checkpoint/key mismatch: missing {missing[:4]} unexpected {unexpected[:4]}")
    if audio_stat is not None:
        model.norm.load_state_dict(torch.load(audio_stat, map_location="cpu", weights_only=True))
    return model.to(device).eval().set_storage_dtype(storage_dtype)
