"""Pin the CPU oracle against outputs of the unmodified reference (tests/golden, made by
oracle/gen_golden.py) and against the known-answer anchors of SURVEY.md section 8c."""
import math

import numpy as np
import pytest
import torch

import v100_oracle as orc
from voice100_b200 import synth
from helpers import asr_case, golden, tts_case


def test_constants_and_shapes():
    g = golden("logmel")
    assert float(g["blank_audio"]) == orc.BLANK_AUDIO == math.log(1e-6)
    assert abs(orc.BLANK_AUDIO - (-13.815510557964274)) < 1e-12
    assert int(g["audio_size"]) == orc.MELSPEC_DIM == 64           # reference tests/test_datasets.py:273
    fb = orc.mel_filterbank()
    assert fb.shape == (257, 64) and int((fb > 0).sum()) == 500   # SURVEY 8c anchor
    for sec, frames in ((1, 101), (10, 1001), (15, 1501), (60, 6001)):
        assert 1 + (sec * 16000) // 160 == frames
    assert [int(x) for x in orc.asr_output_length(torch.tensor([1501, 1001, 100, 1]))] == [751, 501, 50, 1]


def test_param_counts_match_readme():
    n = lambda sd, p: sum(v.size for k, v in sd.items() if k.startswith(p) and "num_batches" not in k
                          and "running" not in k)
    asr = synth.asr_state_dict(64, 512, 29, 512)
    assert n(asr, "encoder.") == 11_606_784 and n(asr, "decoder.") == 14_877     # README.md:135-147
    al = synth.align_state_dict(29, 512)
    assert n(al, "layers.") == 8_553_474 and n(al, "embedding.") == 14_848       # README.md:59-69
    au = synth.audio_state_dict(29, 512)
    assert n(au, "decoder.") == 11_044_868 and n(au, "norm.") == 518             # README.md:73-85


def test_logmel_matches_reference():
    g = golden("logmel")
    cases = {"noise_16000": synth.noise_waveform(1, 16000, seed=11)[0],
             "harm_12345": synth.harmonic_waveform(1, 12345, seed=12)[0],
             "noise_400": synth.noise_waveform(1, 400, seed=13)[0]}
    for k, w in cases.items():
        got = orc.logmel_clip(torch.from_numpy(w)).numpy()
        assert got.shape == g[k].shape == (1 + len(w) // 160, 64)
        np.testing.assert_allclose(got, g[k], rtol=0, atol=2e-5, err_msg=k)
        # step-by-step restatements: fp32 FFT round-off shows up in near-silent mel bins only
        ex = torch.log(orc.mel_power_explicit(torch.from_numpy(w)).T + orc.LOG_OFFSET).numpy()
        np.testing.assert_allclose(ex, g[k], rtol=0, atol=2e-3, err_msg=k + " explicit")
        np.testing.assert_allclose(orc.logmel_numpy(w), g[k], rtol=0, atol=2e-3, err_msg=k + " numpy")
    mp = orc.mel_power(torch.from_numpy(cases["harm_12345"])).numpy()
    np.testing.assert_allclose(mp, g["harm_12345_melpower"], rtol=1e-5, atol=1e-7)
    # ragged batch: per-clip features padded with BLANK_AUDIO (data_modules.py:446-455)
    L = max(len(w) for w in cases.values())
    wav = np.zeros((3, L), np.float32)
    for i, w in enumerate(cases.values()):
        wav[i, :len(w)] = w
    audio, audio_len = orc.logmel_batch(torch.from_numpy(wav), [len(w) for w in cases.values()])
    assert [int(x) for x in audio_len] == [int(x) for x in g["batch_audio_len"]] == [101, 78, 3]
    np.testing.assert_allclose(audio.numpy(), g["batch_audio"], rtol=0, atol=2e-5)
    assert float(audio[2, 3:].max()) == float(np.float32(orc.BLANK_AUDIO))


def _asr(name):
    sd, wav, lengths, g = asr_case(name)
    audio, audio_len = orc.logmel_batch(wav, lengths)
    with torch.no_grad():
        logits = orc.asr_forward(audio, sd)
    assert logits.shape == g["logits"].shape
    np.testing.assert_allclose(logits.numpy(), g["logits"], rtol=0, atol=2e-4)
    assert [int(x) for x in audio_len] == [int(x) for x in g["audio_len"]]
    assert [int(x) for x in orc.asr_output_length(audio_len)] == [int(x) for x in g["out_len"]]
    agree = (orc.ctc_greedy(logits).numpy() == g["tokens"]).mean()
    assert agree > 0.999, agree


def test_asr_en_small_matches_reference():
    _asr("asr_en_small")


def test_asr_ja_ragged_matches_reference():
    _asr("asr_ja_phone_ragged")


def test_asr_ja_base_ragged_matches_reference():
    """BASELINE.json configs[4] at its real width: AudioToTextCTC(64, 512, 44, 512), ragged clips."""
    _asr("asr_ja_phone_base_ragged")


def test_tts_v1_mcep_matches_reference():
    """AlignTextToAudioModel(use_mcep=True), 25 mel-cepstrum outputs (tts.py:153,164)."""
    from helpers import tts_v1_mcep_case
    sd, aligntext, g = tts_v1_mcep_case()
    with torch.no_grad():
        hasf0, f0_hat, mcep_hat, codeap_hat = orc.audio_forward(aligntext, sd)
        f0, mcep, codeap = orc.audio_predict(aligntext, sd)
    assert mcep.shape == g["mcep"].shape == (aligntext.shape[0], 2 * aligntext.shape[1] - 1, 25)
    np.testing.assert_allclose(hasf0.numpy(), g["hasf0_logits"], rtol=0, atol=2e-4)
    np.testing.assert_allclose(mcep_hat.numpy(), g["mcep_hat"], rtol=0, atol=2e-4)
    np.testing.assert_allclose(mcep.numpy(), g["mcep"], rtol=0, atol=1e-3)
    np.testing.assert_allclose(codeap.numpy(), g["codeap"], rtol=0, atol=1e-3)
    safe = np.abs(g["hasf0_logits"]) > 1e-3
    assert ((f0.numpy() == 0) == (g["f0"] == 0))[safe].all()


def test_tokenizers_match_reference():
    """voice100_b200.text (host tail of greedy decoding) against the reference's CharTokenizer / BasicTokenizer
    outputs stored by oracle/gen_golden.py:gen_tokenizer (voice100/text.py:74-145)."""
    import json
    import os
    from helpers import GOLDEN
    from voice100_b200.text import BasicTokenizer, CharTokenizer
    with open(os.path.join(GOLDEN, "tokenizer.json")) as f:
        g = json.load(f)
    toks = {"char": CharTokenizer(), "en": BasicTokenizer("en"), "ja": BasicTokenizer("ja")}
    assert {k: t.vocab_size for k, t in toks.items()} == {"char": 29, "en": 71, "ja": 44}
    for name, tok in toks.items():
        assert tok.vocab_size == g[name]["vocab_size"]
        for case in g[name]["cases"]:
            decoded = tok.decode(case["ids"])
            assert decoded == case["decoded"], (name, case["ids"])
            assert tok.merge_repeated(decoded) == case["merged"], (name, decoded)
            assert tok.encode(decoded).tolist() == case["reencoded"]
    with pytest.raises(ValueError):
        BasicTokenizer("fr")


def test_align_edge_cases_match_reference():
    """TextToAlignTextModel.align / align_batch on alignments with negative entries: wrap-around frame indices,
    non-monotone starts and the IndexError cases, against the reference's outputs (oracle/gen_golden.py:gen_align_edges).
    Pure host code: no GPU involved."""
    import json
    import os
    from helpers import GOLDEN
    import voice100_b200 as v
    with open(os.path.join(GOLDEN, "align_edges.json")) as f:
        cases = json.load(f)
    model = v.TextToAlignTextModel(29, 64)
    n_err = 0
    for c in cases:
        text = torch.tensor(c["text"], dtype=torch.int64)
        al = torch.tensor(c["align"], dtype=torch.float32)
        if isinstance(c["out"], str):
            n_err += 1
            with pytest.raises((IndexError, RuntimeError, ValueError)):
                model.align(text, al)
            with pytest.raises((IndexError, RuntimeError, ValueError)):
                v.align_batch(text[None], al[None])
            continue
        assert model.align(text, al).tolist() == c["out"]
        got, n = v.align_batch(text[None], al[None])
        assert got[0, : int(n[0])].tolist() == c["out"]
    assert 0 < n_err < len(cases)


def test_tts_matches_reference():
    sd_a, sd_v, text, align, g = tts_case()
    with torch.no_grad():
        pred = orc.align_forward(text, sd_a)
    np.testing.assert_allclose(pred.numpy(), g["align_pred"], rtol=0, atol=2e-4)
    ats = [orc.align_text(text[i].tolist(), align[i]) for i in range(text.shape[0])]
    assert [len(a) for a in ats] == [int(x) for x in g["aligntext_len"]]
    at = np.zeros_like(g["aligntext"])
    for i, a in enumerate(ats):
        at[i, :len(a)] = a
    assert (at == g["aligntext"]).all()
    with torch.no_grad():
        hasf0, f0_hat, _, _ = orc.audio_forward(torch.from_numpy(at), sd_v)
        f0, logspc, codeap = orc.audio_predict(torch.from_numpy(at), sd_v)
    T = at.shape[1]
    assert logspc.shape == (at.shape[0], 2 * T - 1, 257) and codeap.shape == (at.shape[0], 2 * T - 1, 1)
    np.testing.assert_allclose(hasf0.numpy(), g["hasf0_logits"], rtol=0, atol=2e-4)
    np.testing.assert_allclose(logspc.numpy(), g["logspc"], rtol=0, atol=1e-3)
    np.testing.assert_allclose(codeap.numpy(), g["codeap"], rtol=0, atol=1e-3)
    # f0 is gated by the sign of hasf0: compare away from the decision boundary
    safe = np.abs(g["hasf0_logits"]) > 1e-3
    np.testing.assert_allclose(f0.numpy()[safe], g["f0"][safe], rtol=0, atol=2e-2)
    assert ((f0.numpy() == 0) == (g["f0"] == 0))[safe].all()


def test_ctc_collapse_known_answers():
    # reference tests/test_text.py:57-59 style: repeats merged, blank dropped
    enc = lambda s: [orc.DEFAULT_CHARACTERS.index(c) for c in s]
    assert orc.ctc_collapse_text(enc("__hh_eel_ll__oo  w_")) == "hello w"
    assert orc.ctc_collapse_text(enc("___")) == ""
    assert orc.ctc_collapse_text(enc("_ _")) == ""
    assert orc.ctc_collapse_text([0, 99, -1, 2, 2, 0, 2]) == "aa"


def test_viterbi_matches_reference():
    """oracle ctc_best_path vs voice100/models/align.py:18-66 run by oracle/gen_golden.py: bit-exact."""
    g = golden("viterbi")
    for ci, (T, L) in enumerate(g["cases"]):
        lp, labels = synth.viterbi_inputs(int(T), int(L), 29, int(g["seed"]) + ci)
        if int(g[f"c{ci}_fail"]):                                 # the reference raised IndexError (align.py:57-58)
            with pytest.raises(IndexError):
                orc.ctc_best_path(lp, labels)
            continue
        score, path, best_labels = orc.ctc_best_path(lp, labels)
        assert np.float32(score) == g[f"c{ci}_score"]
        assert np.array_equal(path, g[f"c{ci}_path"]) and np.array_equal(best_labels, g[f"c{ci}_labels"])
        # structural properties of any valid alignment
        assert path[0] in (0, 1) and path[-1] in (2 * L - 1, 2 * L) and (np.diff(path) >= 0).all() and (np.diff(path) <= 2).all()
    assert sum(int(g[f"c{ci}_fail"]) for ci in range(len(g["cases"]))) >= 3     # both outcomes of T == L are pinned


# ---- v2 models (LayerNorm/GELU conv blocks + bidirectional LSTM) ----

def _asr_v2(name):
    from helpers import asr_v2_case
    sd, wav, lengths, settings, g = asr_v2_case(name)
    audio, audio_len = orc.logmel_batch(wav, lengths)
    assert [int(x) for x in audio_len] == [int(x) for x in g["audio_len"]]
    logits, out_len = orc.asr_v2_forward(audio, audio_len, sd, settings)
    assert logits.shape == g["logits"].shape                      # [T', B, V], time-major (_asr_v2.py:47-49)
    assert [int(x) for x in out_len] == [int(x) for x in g["logits_len"]]
    np.testing.assert_allclose(logits.numpy(), g["logits"], rtol=0, atol=5e-5)
    return logits, out_len, g


def test_asr_v2_small_ragged_matches_reference():
    logits, out_len, g = _asr_v2("asr_v2_en_small_ragged")
    # rows past an utterance's length are the LSTM's zero padding through the dense layer: the bias alone
    from helpers import asr_v2_case
    sd = asr_v2_case("asr_v2_en_small_ragged")[0]
    np.testing.assert_allclose(logits[int(out_len[1]):, 1].numpy(),
                               np.broadcast_to(sd["dense.bias"].numpy(), logits[int(out_len[1]):, 1].shape), atol=1e-6)


def test_asr_v2_base_matches_reference():
    _asr_v2("asr_v2_en_base")


def test_v2_param_counts():
    n = lambda sd: sum(v.size for k, v in sd.items() if not k.startswith("norm."))
    assert n(synth.asr_v2_state_dict()) == 12_008_477              # printed by the reference modules (gen_golden.py)
    assert n(synth.asr_v2_state_dict(64, synth.ASR_V2_SMALL_ENCODER, 2, 256, 29)) == 2_891_293
    assert n(synth.align_v2_state_dict()) == 2_638_082
    assert n(synth.audio_v2_state_dict()) == 15_896_837


def test_tts_v2_matches_reference():
    from helpers import tts_v2_case
    sd_a, sd_v, text, align, g = tts_v2_case()
    text_len = [int(x) for x in g["text_len"]]
    pred, pred_len = orc.align_v2_forward(text, text_len, sd_a)
    assert [int(x) for x in pred_len] == [int(x) for x in g["align_pred_len"]]
    np.testing.assert_allclose(pred.numpy(), g["align_pred"], rtol=0, atol=5e-5)
    ats = [orc.align_text_v2(text[i, :n].numpy(), align[i, :n]) for i, n in enumerate(text_len)]
    assert [len(a) for a in ats] == [int(x) for x in g["aligntext_len"]]
    for i, a in enumerate(ats):
        np.testing.assert_array_equal(a, g["aligntext"][i, :len(a)])
    aligntext = torch.from_numpy(g["aligntext"])
    lens = [int(x) for x in g["aligntext_len"]]
    hasf0, f0_hat, _logspc_hat, hascodeap, _codeap_hat = orc.audio_v2_forward(
        aligntext, lens, sd_v, synth.TTS_V2_BASE_DECODER)
    assert hasf0.shape == g["hasf0_logits"].shape == (3, 2 * max(lens) - 1)
    np.testing.assert_allclose(hasf0.numpy(), g["hasf0_logits"], rtol=0, atol=1e-4)
    np.testing.assert_allclose(f0_hat.numpy(), g["f0_hat"], rtol=0, atol=1e-4)
    np.testing.assert_allclose(hascodeap.numpy(), g["hascodeap_logits"], rtol=0, atol=1e-4)
    f0, logspc, codeap = orc.audio_v2_predict(aligntext, lens, sd_v, synth.TTS_V2_BASE_DECODER)
    # the sign of a near-zero voicing logit may flip between two fp32 summation orders: compare where it is decisive
    sure = np.abs(g["hasf0_logits"]) > 1e-3
    np.testing.assert_allclose(f0.numpy()[sure], g["f0"][sure], rtol=1e-4, atol=1e-2)
    np.testing.assert_allclose(logspc.numpy(), g["logspc"], rtol=0, atol=1e-3)
    sure_c = np.abs(g["hascodeap_logits"]) > 1e-3
    np.testing.assert_allclose(codeap.numpy()[sure_c], g["codeap"][sure_c], rtol=0, atol=1e-3)


def test_mc2sp_matches_reference():
    """mel-cepstrum -> log-spectrum matrix and the export wrapper's output (vocoder.py:115-123, export_onnx.py:81-97)."""
    from voice100_b200.vocoder import create_mc2sp_matrix
    g = golden("tts_v2_mcep")
    ref = g["mc2sp_matrix"]
    assert ref.shape == (25, 257)
    np.testing.assert_allclose(orc.mc2sp_matrix(512, 24, 0.410), ref, rtol=0, atol=1e-6)
    np.testing.assert_allclose(create_mc2sp_matrix(512, 24, 0.410), ref, rtol=0, atol=1e-6)
    np.testing.assert_allclose(ref[0], 2.0, atol=1e-6)                  # c0 alone is a flat spectrum of 2 c0
    V, B, seed = [int(x) for x in g["cfg"]]
    sd = {k: torch.from_numpy(v) for k, v in synth.audio_v2_state_dict(
        V, 25, 1, 2, 512, synth.TTS_V2_BASE_DECODER, seed=seed, randomize_ln=True, randomize_norm=True, gain=2.0).items()}
    lens = [int(x) for x in g["aligntext_len"]]
    f0, mcep, codeap = orc.audio_v2_predict(torch.from_numpy(g["aligntext"]), lens, sd, synth.TTS_V2_BASE_DECODER,
                                            logspc_size=25)
    np.testing.assert_allclose(mcep.numpy(), g["mcep"], rtol=0, atol=2e-4)
    logspc = mcep.double().numpy() @ orc.mc2sp_matrix(512, 24, 0.410)
    np.testing.assert_allclose(logspc, g["logspc"], rtol=0, atol=5e-3)  # |logspc| ~ 1e2: fp32 matmul round-off


def test_maskaudio_matches_reference_golden():
    """oracle.maskaudio restates BatchSpectrogramAugumentation.maskaudio (voice100/audio.py:106-108); the fixture is the
    reference method's own output (oracle/gen_golden.py:gen_maskaudio)."""
    g = golden("maskaudio")
    B, T, C, seed = [int(x) for x in g["cfg"]]
    rng = np.random.Generator(np.random.PCG64(seed))
    audio = rng.uniform(orc.BLANK_AUDIO - 1.0, 12.0, size=(B, T, C)).astype(np.float32)
    audio[0, :3] = orc.BLANK_AUDIO
    out = orc.maskaudio(torch.from_numpy(audio), torch.from_numpy(g["audio_len"]))
    assert np.array_equal(out.numpy(), g["out"])                      # same ATen ops in the same order: bit-identical
    for b, n in enumerate(g["audio_len"]):
        assert np.all(g["out"][b, int(n):] == np.float32(orc.BLANK_AUDIO))


def test_logmel_generic_configs_match_reference():
    """The oracle front end with constructor arguments other than 512 / 400 / 160 / 64, against the reference class built with
    the same arguments (oracle/gen_golden.py:gen_logmel_generic)."""
    g = golden("logmel_generic")
    clips = {"noise": synth.noise_waveform(1, 6000, seed=111)[0], "harm": synth.harmonic_waveform(1, 4321, seed=112)[0]}
    i = 0
    while f"c{i}_cfg" in g.files:
        sr, n_fft, win, hop, n_mels = [int(x) for x in g[f"c{i}_cfg"]]
        for name, w in clips.items():
            got = orc.logmel_clip(torch.from_numpy(w), sample_rate=sr, n_fft=n_fft, win_length=win, hop_length=hop,
                                  n_mels=n_mels).numpy()
            ref = g[f"c{i}_{name}"]
            assert got.shape == ref.shape == (1 + len(w) // hop, n_mels)
            np.testing.assert_allclose(got, ref, rtol=0, atol=2e-5, err_msg=f"config {i} {name}")
        i += 1
    assert i == 4
