"""Generate tests/golden/*.npz by running the UNMODIFIED reference (kaiidams/voice100).

Run in the build container only (the GPU box has no /root/reference):

    python oracle/gen_golden.py            # writes tests/golden/*.npz

The reference imports from /root/reference; `pytorch_lightning` (not installable here) is
replaced by the stand-in in oracle/_shim.  Weights come from voice100_b200.synth (numpy PCG64,
reproducible anywhere) and are loaded with the reference's own `load_state_dict`; every
BatchNorm then gets data-calibrated running statistics from one training-mode forward with
momentum=1 so that BN folding is genuinely exercised.  The fixtures store the calibrated BN
statistics and the reference OUTPUTS; inputs and all other weights are regenerated from seeds.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "_shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

import numpy as np
import torch

from voice100_b200 import synth

from voice100.data_modules import (BLANK_AUDIO, MelSpectrogramAudioTransform,  # noqa: E402
                                   generate_audio_text_batch)
from voice100.models.asr import AudioToTextCTC  # noqa: E402
from voice100.models.tts import AlignTextToAudioModel, TextToAlignTextModel  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
torch.set_num_threads(4)


def load(model, sd_np):
    sd = {k: torch.from_numpy(v) for k, v in sd_np.items()}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    # criterion / augmentation buffers are not part of the inference path
    assert all(m.startswith(("criterion", "batch_augment")) for m in missing), missing


def calibrate(model, *inputs):
    bns = [m for m in model.modules() if isinstance(m, torch.nn.BatchNorm1d)]
    for m in bns:
        m.momentum = 1.0
    model.train()
    with torch.no_grad():
        model(*inputs)
    model.eval()
    out = {}
    for k, v in model.state_dict().items():
        if k.endswith("running_mean") or k.endswith("running_var"):
            out["bn/" + k] = v.numpy().copy()
    return out


def gen_logmel():
    tr = MelSpectrogramAudioTransform()
    w_noise = torch.from_numpy(synth.noise_waveform(1, 16000, seed=11))[0]
    w_harm = torch.from_numpy(synth.harmonic_waveform(1, 12345, seed=12))[0]
    w_short = torch.from_numpy(synth.noise_waveform(1, 400, seed=13))[0]
    feats = []
    with torch.no_grad():
        for w in (w_noise, w_harm, w_short):
            # voice100/data_modules.py:290-291 (file load + resample at :288-289 are out of scope)
            feats.append(torch.log(tr.melspec(w).T + tr.log_offset))
        mel_power = tr.melspec(w_harm)
    (audio, audio_len), _ = generate_audio_text_batch([(f, torch.zeros(1, dtype=torch.long)) for f in feats])
    np.savez_compressed(
        os.path.join(OUT, "logmel.npz"),
        noise_16000=feats[0].numpy(), harm_12345=feats[1].numpy(), noise_400=feats[2].numpy(),
        harm_12345_melpower=mel_power.numpy(),
        batch_audio=audio.numpy(), batch_audio_len=audio_len.numpy(),
        blank_audio=np.float64(BLANK_AUDIO), audio_size=np.int64(tr.audio_size))
    print("logmel", [tuple(f.shape) for f in feats], tuple(audio.shape))


def gen_asr(name, hidden, embed, vocab, batch, samples, lengths=None, seed=21):
    tr = MelSpectrogramAudioTransform()
    wav = torch.from_numpy(synth.noise_waveform(batch, samples, seed=seed))
    if lengths is None:
        lengths = [samples] * batch
    with torch.no_grad():
        feats = [torch.log(tr.melspec(wav[i, :n]).T + tr.log_offset) for i, n in enumerate(lengths)]
    (audio, audio_len), _ = generate_audio_text_batch([(f, torch.zeros(1, dtype=torch.long)) for f in feats])
    model = AudioToTextCTC(audio_size=64, embed_size=embed, vocab_size=vocab, hidden_size=hidden,
                           learning_rate=1e-3, weight_decay=0.0)
    load(model, synth.asr_state_dict(64, embed, vocab, hidden, seed=seed, randomize_bn=True))
    bn = calibrate(model, audio)
    with torch.no_grad():
        logits = model(audio)
        out_len = model.output_length(audio_len)
    np.savez_compressed(
        os.path.join(OUT, f"{name}.npz"), logits=logits.numpy(), tokens=logits.argmax(-1).numpy(),
        audio_len=audio_len.numpy(), out_len=out_len.numpy(), lengths=np.asarray(lengths, np.int32),
        cfg=np.asarray([64, embed, vocab, hidden, batch, samples, seed], np.int64), **bn)
    print(name, tuple(logits.shape), "logit std %.4f" % float(logits.std()),
          "enc params", sum(p.numel() for p in model.encoder.parameters()),
          "dec params", sum(p.numel() for p in model.decoder.parameters()))


def gen_tts(seed=31):
    B, L, H, V = 2, 24, 512, 29
    text = torch.from_numpy(synth.text_tokens(B, L, V, seed=seed))
    amodel = TextToAlignTextModel(vocab_size=V, hidden_size=H, learning_rate=1e-3)
    load(amodel, synth.align_state_dict(V, H, seed=seed, randomize_bn=True))
    bn_a = calibrate(amodel, text)
    with torch.no_grad():
        pred = amodel(text)
    # host alignment on the seeded synthetic alignment (random-init preds can be negative)
    align = synth.synthetic_alignment(B, L, seed=seed)
    ats = [amodel.align(text[i], torch.from_numpy(align[i])) for i in range(B)]
    aligntext = torch.nn.utils.rnn.pad_sequence(ats, batch_first=True, padding_value=0)
    vmodel = AlignTextToAudioModel(vocab_size=V, hidden_size=H, learning_rate=1e-3)
    load(vmodel, synth.audio_state_dict(V, H, seed=seed, randomize_bn=True, randomize_norm=True))
    bn_v = calibrate(vmodel, aligntext)
    with torch.no_grad():
        hasf0, f0_hat, logspc_hat, codeap_hat = vmodel(aligntext)
        f0, logspc, codeap = vmodel.predict(aligntext)
    np.savez_compressed(
        os.path.join(OUT, "tts_en_base.npz"),
        align_pred=pred.numpy(), aligntext=aligntext.numpy(),
        aligntext_len=np.asarray([len(a) for a in ats], np.int32),
        hasf0_logits=hasf0.numpy(), f0_hat=f0_hat.numpy(), f0=f0.numpy(), logspc=logspc.numpy(),
        codeap=codeap.numpy(), cfg=np.asarray([V, H, B, L, seed], np.int64),
        **{"a/" + k: v for k, v in bn_a.items()}, **{"v/" + k: v for k, v in bn_v.items()})
    print("tts", tuple(pred.shape), tuple(aligntext.shape), tuple(logspc.shape),
          "align params", sum(p.numel() for p in amodel.layers.parameters()),
          "audio dec params", sum(p.numel() for p in vmodel.decoder.parameters()),
          "voiced frac %.2f" % float((hasf0 >= 0).float().mean()))


def gen_viterbi(seed=41):
    """Forced-alignment vectors from the reference's own ctc_best_path (voice100/models/align.py:18-66)."""
    from voice100.models.align import ctc_best_path as ref_best_path
    out = {}
    # the T == L cases reach only S - 1 states: the reference then succeeds or raises IndexError depending on the
    # last two live scores (align.py:57-58); (9, 10) and (1, 1) always/sometimes fail
    cases = [(60, 7), (200, 31), (751, 120), (40, 19), (3, 1), (12, 12), (30, 30), (7, 7), (31, 31), (9, 10), (1, 1),
             (13, 13), (14, 14)]
    for ci, (T, L) in enumerate(cases):
        lp, labels = viterbi_case(T, L, seed + ci)
        try:
            score, path, best_labels = ref_best_path(lp, labels)
            out[f"c{ci}_fail"] = np.int64(0)
        except IndexError:
            score, path, best_labels = np.nan, np.full(T, -1), np.zeros(T)
            out[f"c{ci}_fail"] = np.int64(1)
        out[f"c{ci}_score"] = np.float32(score)
        out[f"c{ci}_path"] = path.astype(np.int32)
        out[f"c{ci}_labels"] = best_labels.astype(np.int64)
    out["cases"] = np.asarray(cases, np.int64)
    out["seed"] = np.int64(seed)
    np.savez_compressed(os.path.join(OUT, "viterbi.npz"), **out)
    print("viterbi", cases, "IndexError:", [int(out[f"c{ci}_fail"]) for ci in range(len(cases))])


def gen_asr_v2(name, settings, hidden, vocab, batch, samples, lengths, seed):
    """AudioToAlignText (voice100/models/_asr_v2.py:18-49), the architecture of the shipped asr_*.yaml configs."""
    from voice100.models._asr_v2 import AudioToAlignText
    tr = MelSpectrogramAudioTransform()
    wav = torch.from_numpy(synth.noise_waveform(batch, samples, seed=seed))
    with torch.no_grad():
        feats = [torch.log(tr.melspec(wav[i, :n]).T + tr.log_offset) for i, n in enumerate(lengths)]
    (audio, audio_len), _ = generate_audio_text_batch([(f, torch.zeros(1, dtype=torch.long)) for f in feats])
    model = AudioToAlignText(audio_size=64, encoder_settings=[list(r) for r in settings], decoder_num_layers=2,
                             decoder_hidden_size=hidden, vocab_size=vocab)
    load(model, synth.asr_v2_state_dict(64, settings, 2, hidden, vocab, seed=seed, randomize_ln=True, gain=2.0))
    model.eval()
    with torch.no_grad():
        logits, logits_len = model(audio, audio_len)
    np.savez_compressed(
        os.path.join(OUT, f"{name}.npz"), logits=logits.numpy(), logits_len=logits_len.numpy(),
        audio_len=audio_len.numpy(), lengths=np.asarray(lengths, np.int32),
        cfg=np.asarray([64, hidden, vocab, batch, samples, seed], np.int64))
    print(name, tuple(logits.shape), logits_len.tolist(), "logit std %.4f" % float(logits.std()),
          "params", sum(p.numel() for p in model.parameters()))


def gen_tts_v2(seed=51):
    """TextToAlignText + AlignTextToAudio (voice100/models/_align_v2.py, _tts_v2.py; config/align_en_base.yaml,
    config/tts_en_base.yaml)."""
    from voice100.models._align_v2 import TextToAlignText
    from voice100.models._tts_v2 import AlignTextToAudio
    B, L, V = 3, 20, 29
    text = torch.from_numpy(synth.text_tokens(B, L, V, seed=seed))
    text_len = torch.tensor([20, 13, 7])
    amodel = TextToAlignText(vocab_size=V, num_layers=2, hidden_size=256, num_outputs=2, learning_rate=1e-3)
    load(amodel, synth.align_v2_state_dict(V, 2, 256, 2, seed=seed, gain=2.0))
    amodel.eval()
    with torch.no_grad():
        pred, pred_len = amodel(text, text_len)
    align = synth.synthetic_alignment(B, L, seed=seed)
    ats = [amodel.align(text[i, :n], torch.from_numpy(align[i, :n])) for i, n in enumerate(text_len.tolist())]
    aligntext = torch.nn.utils.rnn.pad_sequence(ats, batch_first=True, padding_value=0)
    aligntext_len = torch.tensor([len(a) for a in ats])
    vmodel = AlignTextToAudio(vocab_size=V, logspc_size=257, codeap_size=1, encoder_num_layers=2,
                              encoder_hidden_size=512, decoder_settings=[list(r) for r in synth.TTS_V2_BASE_DECODER])
    load(vmodel, synth.audio_v2_state_dict(V, seed=seed, randomize_ln=True, randomize_norm=True, gain=2.0))
    vmodel.eval()
    with torch.no_grad():
        hasf0, f0_hat, logspc_hat, hascodeap, codeap_hat = vmodel(aligntext, aligntext_len)
        f0, logspc, codeap = vmodel.predict(aligntext, aligntext_len)
    np.savez_compressed(
        os.path.join(OUT, "tts_v2_en_base.npz"),
        align_pred=pred.numpy(), align_pred_len=pred_len.numpy(), text_len=text_len.numpy(),
        aligntext=aligntext.numpy(), aligntext_len=aligntext_len.numpy(),
        hasf0_logits=hasf0.numpy(), f0_hat=f0_hat.numpy(), hascodeap_logits=hascodeap.numpy(),
        f0=f0.numpy(), logspc=logspc.numpy(), codeap=codeap.numpy(),
        cfg=np.asarray([V, B, L, seed], np.int64))
    print("tts_v2", tuple(pred.shape), tuple(aligntext.shape), aligntext_len.tolist(), tuple(logspc.shape),
          "align params", sum(p.numel() for p in amodel.parameters()),
          "audio params", sum(p.numel() for p in vmodel.parameters() if p.requires_grad))


def gen_mcep(seed=61):
    """mel-cepstrum TTS head: create_mc2sp_matrix (voice100/vocoder.py:115-123) and the export wrapper
    AlignTextToAudioPredict (voice100/export_onnx.py:81-97) over a logspc_size=25 model, as config/tts_en_base.yaml
    trains it.  voice100.vocoder imports pyworld at module level; oracle/_shim/pyworld.py stands in for it."""
    from voice100.vocoder import create_mc2sp_matrix
    from voice100.export_onnx import AlignTextToAudioPredict
    from voice100.models._tts_v2 import AlignTextToAudio
    V, B = 29, 2
    aligntext = torch.from_numpy(synth.text_tokens(B, 40, V, seed=seed))
    aligntext_len = torch.tensor([40, 23])
    aligntext[1, 23:] = 0
    model = AlignTextToAudio(vocab_size=V, logspc_size=25, codeap_size=1, encoder_num_layers=2,
                             encoder_hidden_size=512, decoder_settings=[list(r) for r in synth.TTS_V2_BASE_DECODER])
    load(model, synth.audio_v2_state_dict(V, 25, 1, 2, 512, synth.TTS_V2_BASE_DECODER, seed=seed, randomize_ln=True,
                                          randomize_norm=True, gain=2.0))
    model.eval()
    with torch.no_grad():
        f0, logspc, codeap = AlignTextToAudioPredict(model)(aligntext, aligntext_len)
        _, mcep, _ = model.predict(aligntext, aligntext_len)
    np.savez_compressed(
        os.path.join(OUT, "tts_v2_mcep.npz"), mc2sp_matrix=create_mc2sp_matrix(512, 24, 0.410).astype(np.float32),
        aligntext=aligntext.numpy(), aligntext_len=aligntext_len.numpy(), f0=f0.numpy(), logspc=logspc.numpy(),
        mcep=mcep.numpy(), codeap=codeap.numpy(), cfg=np.asarray([V, B, seed], np.int64))
    print("mcep", tuple(mcep.shape), "->", tuple(logspc.shape), "logspc std %.3f" % float(logspc.std()))


def gen_tts_v1_mcep(seed=71):
    """AlignTextToAudioModel(use_mcep=True): 25 mel-cepstrum outputs (voice100/models/tts.py:153,164)."""
    B, T, H, V = 2, 30, 512, 29
    aligntext = torch.from_numpy(synth.text_tokens(B, T, V, seed=seed))
    model = AlignTextToAudioModel(vocab_size=V, hidden_size=H, learning_rate=1e-3, use_mcep=True)
    load(model, synth.audio_state_dict(V, H, seed=seed, randomize_bn=True, randomize_norm=True, logspc_size=25))
    bn = calibrate(model, aligntext)
    with torch.no_grad():
        hasf0, f0_hat, mcep_hat, codeap_hat = model(aligntext)
        f0, mcep, codeap = model.predict(aligntext)
    np.savez_compressed(
        os.path.join(OUT, "tts_v1_mcep.npz"), hasf0_logits=hasf0.numpy(), f0_hat=f0_hat.numpy(),
        mcep_hat=mcep_hat.numpy(), codeap_hat=codeap_hat.numpy(), f0=f0.numpy(), mcep=mcep.numpy(),
        codeap=codeap.numpy(), cfg=np.asarray([V, H, B, T, seed], np.int64), **bn)
    print("tts_v1_mcep", tuple(mcep.shape), "audio_size", model.audio_size)


def gen_tokenizer(seed=81):
    """ids -> decode -> merge_repeated through the reference's CharTokenizer / BasicTokenizer (voice100/text.py:
    74-145) on CTC-like id sequences (runs, blanks, out-of-vocabulary ids)."""
    import json
    from voice100.text import BasicTokenizer, CharTokenizer
    rng = np.random.default_rng(seed)
    out = {}
    for name, tok in (("char", CharTokenizer()), ("en", BasicTokenizer("en")), ("ja", BasicTokenizer("ja"))):
        cases = []
        for _ in range(40):
            ids = []
            for _ in range(int(rng.integers(0, 25))):
                t = int(rng.choice([0, 0, int(rng.integers(0, tok.vocab_size)), int(rng.integers(-2, tok.vocab_size + 3))]))
                ids += [t] * int(rng.integers(1, 5))
            decoded = tok.decode(torch.tensor(ids, dtype=torch.long))
            cases.append(dict(ids=ids, decoded=decoded, merged=tok.merge_repeated(decoded),
                              reencoded=tok.encode(decoded).tolist()))
        out[name] = dict(vocab_size=tok.vocab_size, cases=cases)
    with open(os.path.join(OUT, "tokenizer.json"), "w") as f:
        json.dump(out, f)
    print("tokenizer", {k: len(v["cases"]) for k, v in out.items()})


def gen_align_edges(seed=91):
    """TextToAlignTextModel.align (voice100/models/tts.py:89-110) on alignments with negative gaps / durations: frame
    indices that wrap around from the end of the tensor, non-monotone starts, and the IndexError cases."""
    import json
    model = TextToAlignTextModel(vocab_size=29, hidden_size=64, learning_rate=1e-3)
    rng = np.random.default_rng(seed)
    cases = []
    for trial in range(60):
        L = int(rng.integers(1, 12))
        text = rng.integers(1, 29, L)
        al = rng.normal(1.0, 2.0, (L, 2)).astype(np.float32)
        if trial % 3 == 0:
            al = np.abs(al)
        if trial % 10 == 9:                     # a big negative duration early on: later frames fall outside the tensor
            al[0, 1] = -float(np.abs(al).sum()) - 3.0
        try:
            out = model.align(torch.from_numpy(text), torch.from_numpy(al)).tolist()
        except (IndexError, RuntimeError, ValueError) as exc:
            out = type(exc).__name__
        cases.append(dict(text=text.tolist(), align=[[float(x) for x in row] for row in al], out=out))
    with open(os.path.join(OUT, "align_edges.json"), "w") as f:
        json.dump(cases, f)
    print("align_edges", len(cases), "errors", sum(isinstance(c["out"], str) for c in cases))


GENERIC_MEL_CONFIGS = ((16000, 256, 200, 80, 40), (22050, 1024, 800, 256, 80), (8000, 64, 64, 16, 13), (16000, 512, 320, 100, 64))


def gen_logmel_generic():
    """MelSpectrogramAudioTransform with constructor arguments other than the 512 / 400 / 160 / 64 the reference's code
    builds (voice100/data_modules.py:263-281 takes all of them): log-mel features of a noise and a harmonic clip."""
    out = {}
    w_noise = torch.from_numpy(synth.noise_waveform(1, 6000, seed=111))[0]
    w_harm = torch.from_numpy(synth.harmonic_waveform(1, 4321, seed=112))[0]
    for i, (sr, n_fft, win, hop, n_mels) in enumerate(GENERIC_MEL_CONFIGS):
        tr = MelSpectrogramAudioTransform(sample_rate=sr, n_fft=n_fft, win_length=win, hop_length=hop, n_mels=n_mels)
        with torch.no_grad():
            for name, w in (("noise", w_noise), ("harm", w_harm)):
                out[f"c{i}_{name}"] = torch.log(tr.melspec(w).T + tr.log_offset).numpy()
        out[f"c{i}_cfg"] = np.asarray([sr, n_fft, win, hop, n_mels])
    np.savez_compressed(os.path.join(OUT, "logmel_generic.npz"), **out)
    print("logmel_generic", {k: v.shape for k, v in out.items() if not k.endswith("cfg")})


def gen_maskaudio(seed=101):
    """BatchSpectrogramAugumentation.maskaudio (voice100/audio.py:106-108) on a ragged log-mel-like batch: values over the
    whole range the front end produces (BLANK_AUDIO .. +12), lengths 0 / 1 / interior / full."""
    from voice100.audio import BatchSpectrogramAugumentation
    rng = np.random.Generator(np.random.PCG64(seed))
    B, T, C = 5, 37, 64
    audio = rng.uniform(BLANK_AUDIO - 1.0, 12.0, size=(B, T, C)).astype(np.float32)
    audio[0, :3] = BLANK_AUDIO
    audio_len = np.asarray([37, 0, 1, 20, 36], dtype=np.int32)
    aug = BatchSpectrogramAugumentation()
    with torch.no_grad():
        out = aug.maskaudio(torch.from_numpy(audio), torch.from_numpy(audio_len))
    np.savez_compressed(os.path.join(OUT, "maskaudio.npz"), cfg=np.asarray([B, T, C, seed]), audio_len=audio_len,
                        out=out.numpy())
    print("maskaudio", out.shape, float(out.min()), float(out.max()))


def viterbi_case(T, L, seed):
    from voice100_b200.synth import viterbi_inputs
    return viterbi_inputs(T, L, 29, seed)


if __name__ == "__main__":
    gen_viterbi()
    gen_logmel()
    gen_asr("asr_en_small", hidden=256, embed=256, vocab=29, batch=2, samples=32000)
    gen_asr("asr_ja_phone_ragged", hidden=128, embed=128, vocab=44, batch=3, samples=24000,
            lengths=[9000, 17777, 24000], seed=22)
    gen_tts()
    gen_asr_v2("asr_v2_en_small_ragged", synth.ASR_V2_SMALL_ENCODER, 256, 29, batch=3, samples=24000,
               lengths=[24000, 9000, 17777], seed=23)
    gen_asr_v2("asr_v2_en_base", synth.ASR_V2_BASE_ENCODER, 512, 29, batch=2, samples=16000,
               lengths=[16000, 16000], seed=24)
    gen_asr("asr_ja_phone_base_ragged", hidden=512, embed=512, vocab=44, batch=3, samples=24000,
            lengths=[9000, 17777, 24000], seed=25)        # BASELINE.json configs[4] at its real width
    gen_tts_v2()
    gen_mcep()
    gen_tts_v1_mcep()
    gen_tokenizer()
    gen_align_edges()
    gen_maskaudio()
    gen_logmel_generic()
