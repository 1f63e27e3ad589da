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
from .v2 import AlignTextToAudio, AudioToAlignText, TextToAlignText

_CLASSES = {
    "AudioToTextCTC": (AudioToTextCTC, ("audio_size", "embed_size", "vocab_size", "hidden_size")),
    "TextToAlignTextModel": (TextToAlignTextModel, ("vocab_size", "hidden_size")),
    "AlignTextToAudioModel": (AlignTextToAudioModel, ("vocab_size", "hidden_size")),
}
# v2 classes (voice100/models/_asr_v2.py, _align_v2.py, _tts_v2.py): list-valued hyper-parameters
_V2_CLASSES = {"AudioToAlignText": AudioToAlignText, "TextToAlignText": TextToAlignText,
               "AlignTextToAudio": AlignTextToAudio}
# keys that belong to training-only sub-modules of the reference classes
_IGNORED_PREFIXES = ("criterion.", "batch_augment.")


def _guess_class(state_dict) -> str:
    keys = state_dict.keys()
    if "lstm.weight_ih_l0" in keys:
        if "encoder.0.conv.weight" in keys:
            return "AudioToAlignText"
        if "projection.weight" in keys:
            return "AlignTextToAudio"
        if "embedding.weight" in keys and "dense.weight" in keys:
            return "TextToAlignText"
    if any(k.startswith("encoder.layers.") for k in keys):
        return "AudioToTextCTC"
    if any(k.startswith("decoder.layers.") for k in keys) and "embedding.weight" in keys:
        return "AlignTextToAudioModel"
    if "embedding.weight" in keys and any(k.startswith("layers.") for k in keys):
        return "TextToAlignTextModel"
    raise V100Error("checkpoint does not look like one of the models this package accelerates (AudioToTextCTC / "
                    "TextToAlignTextModel / AlignTextToAudioModel / AudioToAlignText / TextToAlignText / "
                    "AlignTextToAudio)")


def _conv_settings(sd, prefix: str, hp_settings):
    """Rows (out_channels, transpose, kernel_size, stride, padding, bias) of a get_conv_layers stack.  Stride,
    padding and the transpose flag are not recoverable from the tensors, so they come from the checkpoint's
    hyper-parameters; what the tensors do say is cross-checked."""
    n = 0
    while f"{prefix}.{n}.conv.weight" in sd:
        n += 1
    if hp_settings is None:
        raise V100Error(f"the checkpoint stores no hyper_parameters: pass hparams={{'{prefix}_settings': [...]}} "
                        f"(rows as in config/*.yaml) to load_checkpoint")
    rows = [list(r) for r in hp_settings]
    if len(rows) != n:
        raise V100Error(f"{prefix}_settings has {len(rows)} rows but the checkpoint has {n} conv blocks")
    for i, (co, transpose, k, _s, _p, bias) in enumerate(rows):
        w = sd[f"{prefix}.{i}.conv.weight"]
        if w.shape[1 if transpose else 0] != co or w.shape[2] != k or (f"{prefix}.{i}.conv.bias" in sd) != bool(bias):
            raise V100Error(f"{prefix}_settings row {i} disagrees with the weights {tuple(w.shape)}")
    return rows


def _num_lstm_layers(sd) -> int:
    n = 0
    while f"lstm.weight_ih_l{n}" in sd:
        n += 1
    return n


def _build_v2(name: str, sd, hp: dict):
    H = sd["lstm.weight_hh_l0"].shape[1]
    if name == "AudioToAlignText":
        return AudioToAlignText(audio_size=sd["encoder.0.conv.weight"].shape[1],
                                encoder_settings=_conv_settings(sd, "encoder", hp.get("encoder_settings")),
                                decoder_num_layers=_num_lstm_layers(sd), decoder_hidden_size=H,
                                vocab_size=sd["dense.weight"].shape[0])
    if name == "TextToAlignText":
        return TextToAlignText(vocab_size=sd["embedding.weight"].shape[0], num_layers=_num_lstm_layers(sd),
                               hidden_size=H, num_outputs=sd["dense.weight"].shape[0])
    codeap = sd["norm.codeap_std"].shape[0] if "norm.codeap_std" in sd else int(hp.get("codeap_size", 1))
    logspc = sd["projection.weight"].shape[0] - 2 - 2 * codeap
    return AlignTextToAudio(vocab_size=sd["embedding.weight"].shape[0], logspc_size=logspc, codeap_size=codeap,
                            encoder_num_layers=_num_lstm_layers(sd), encoder_hidden_size=H,
                            decoder_settings=_conv_settings(sd, "decoder", hp.get("decoder_settings")))


def _infer_hparams(cls_name: str, sd) -> dict:
    if cls_name == "AudioToTextCTC":
        return dict(audio_size=sd["encoder.layers.0.conv.0.0.weight"].shape[1],
                    hidden_size=sd["encoder.layers.4.conv.2.weight"].shape[0],
                    embed_size=sd["encoder.layers.8.conv.2.weight"].shape[0],
                    vocab_size=sd["decoder.layers.1.weight"].shape[0])
    return dict(vocab_size=sd["embedding.weight"].shape[0], hidden_size=sd["embedding.weight"].shape[1])


def load_checkpoint(path: str, model_class: Optional[str] = None, audio_stat: Optional[str] = None,
                    device: str = "cuda", storage_dtype=torch.bfloat16, hparams: Optional[dict] = None):
    """-> an eval-mode drop-in module with the checkpoint's weights on `device`.  `hparams` overrides / supplies
    hyper-parameters the file does not carry (a bare state_dict of a v2 model needs its `*_settings` rows)."""
    ckpt = torch.load(path, map_location="cpu", weights_only=True)
    sd = ckpt["state_dict"] if isinstance(ckpt, dict) and "state_dict" in ckpt else ckpt
    sd = {k: v for k, v in sd.items() if not k.startswith(_IGNORED_PREFIXES)}
    name = model_class or _guess_class(sd)
    hp = dict(ckpt.get("hyper_parameters", {})) if isinstance(ckpt, dict) else {}
    hp.update(hparams or {})
    if name in _V2_CLASSES:
        try:
            model = _build_v2(name, sd, hp)
        except KeyError as e:
            raise V100Error(f"checkpoint has no {e.args[0]!r}: it is not a {name}") from None
        missing, unexpected = model.load_state_dict(sd, strict=False)
        if unexpected or missing:
            raise V100Error(f"checkpoint/key mismatch: missing {missing[:4]} unexpected {unexpected[:4]}")
        if audio_stat is not None:
            model.norm.load_state_dict(torch.load(audio_stat, map_location="cpu", weights_only=True))
        return model.to(device).eval().set_storage_dtype(storage_dtype)
    if name not in _CLASSES:
        raise V100Error(f"unknown model class {name!r}; expected one of {sorted(_CLASSES) + sorted(_V2_CLASSES)}")
    cls, arg_names = _CLASSES[name]
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
