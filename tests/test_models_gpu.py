"""End-to-end parity of the drop-in modules against the CPU oracle (and through it the reference's
golden vectors).  16-bit tensor-core compute vs fp32 reference: the tolerances are stated once, in
tests/helpers.py (BF16_VS_FP32 / F16_VS_FP32 / VS_STORAGE_MODEL / GATE_MARGIN_REL_STD), as fractions of the
reference output's standard deviation; DESIGN.md section 4 has the error budget behind them."""
import numpy as np
import pytest
import torch

import v100_oracle as orc
import voice100_b200 as v
from voice100_b200 import synth
from helpers import (BF16_VS_FP32, BF16_VS_FP32_NARROW, F16_VS_FP32, asr_case, check_asr_parity, tts_case, tts_v1_mcep_case)

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _load(model, sd):
    model.load_state_dict({k: (t if isinstance(t, torch.Tensor) else torch.from_numpy(np.asarray(t))) for k, t in sd.items()})
    return model.to(DEV).eval()


def _i32(x):
    return torch.as_tensor(x, dtype=torch.int32, device=DEV)


@pytest.mark.parametrize("name", ["asr_en_small", "asr_ja_phone_ragged", "asr_ja_phone_base_ragged"])
def test_asr_matches_golden(name):
    """Reference-generated logits (oracle/gen_golden.py).  asr_ja_phone_base_ragged is BASELINE.json configs[4] at
    its real width: AudioToTextCTC(64, 512, 44, 512) on ragged clips padded with BLANK_AUDIO."""
    sd, wav, lengths, g = asr_case(name)
    audio_size, embed, vocab, hidden = [int(x) for x in g["cfg"][:4]]
    model = _load(v.AudioToTextCTC(audio_size, embed, vocab, hidden), sd)
    tr = v.MelSpectrogramAudioTransform().to(DEV)
    audio, audio_len = tr.logmel_batch(wav.to(DEV), _i32(lengths))
    logits = model(audio)
    tokens, out_len = v.AsrPipeline(tr, model)(wav.to(DEV), _i32(lengths))
    torch.cuda.synchronize()
    ref = torch.from_numpy(g["logits"])
    assert logits.shape == ref.shape
    assert audio_len.cpu().tolist() == g["audio_len"].tolist() and out_len.cpu().tolist() == g["out_len"].tolist()
    assert out_len.dtype == torch.int32 and model.output_length(audio_len).cpu().tolist() == g["out_len"].tolist()
    audio_ref, _ = orc.logmel_batch(wav, lengths)
    with torch.no_grad():
        model_ref = orc.asr_forward_storage_model(audio_ref, sd, torch.bfloat16)
    check_asr_parity(name, ref, logits.cpu(), tokens.cpu(), model_ref, BF16_VS_FP32_NARROW if hidden < 256 else BF16_VS_FP32)
    assert (logits.argmax(-1) == tokens).float().mean() > 0.999     # the fused argmax is the logits' argmax


def test_asr_fp16_storage_is_8x_tighter():
    """The same kernels with fp16 storage (3 more mantissa bits): the end-to-end error drops by the expected
    factor, which pins the bf16 numbers on the storage format and not on the kernels."""
    sd, wav, lengths, g = asr_case("asr_en_small")
    audio_size, embed, vocab, hidden = [int(x) for x in g["cfg"][:4]]
    ref = torch.from_numpy(g["logits"])
    tr = v.MelSpectrogramAudioTransform().to(DEV)
    audio_ref, _ = orc.logmel_batch(wav, lengths)
    reps = {}
    for dtype, tol in ((torch.bfloat16, BF16_VS_FP32), (torch.float16, F16_VS_FP32)):
        model = _load(v.AudioToTextCTC(audio_size, embed, vocab, hidden), sd).set_storage_dtype(dtype)
        tokens, _ = v.AsrPipeline(tr, model)(wav.to(DEV), _i32(lengths))
        audio, _ = tr.logmel_batch(wav.to(DEV), _i32(lengths))
        logits = model(audio).cpu()
        with torch.no_grad():
            model_ref = orc.asr_forward_storage_model(audio_ref, sd, dtype)
        reps[dtype] = check_asr_parity(str(dtype), ref, logits, tokens.cpu(), None, tol)
        rep2 = orc.parity_report(model_ref, logits)
        assert rep2["rms_rel_std"] < (0.005 if dtype == torch.float16 else 0.03), rep2
    assert reps[torch.float16]["rms_rel_std"] < reps[torch.bfloat16]["rms_rel_std"] / 4


def test_asr_en_small_config0_live_oracle():
    """BASELINE.json configs[0] at its full size: asr_en_small, 8 x 10 s -> log-mel -> ConvVoiceEncoder -> CTC greedy,
    every clip against the fp32 oracle run live on the host (and against the same-storage oracle), bf16 and fp16."""
    B, L = 8, 160000
    wav = torch.from_numpy(synth.noise_waveform(B, L, seed=91))
    lengths = [L] * B
    sd = orc.to_torch_sd(synth.asr_state_dict(64, 256, 29, 256, seed=91, randomize_bn=True))
    audio_ref, _ = orc.logmel_batch(wav, lengths)
    sd = orc.calibrate_asr(sd, audio_ref[:2])
    with torch.no_grad():
        ref = orc.asr_forward(audio_ref, sd)
    tr = v.MelSpectrogramAudioTransform().to(DEV)
    for dtype, tol in ((torch.bfloat16, BF16_VS_FP32), (torch.float16, F16_VS_FP32)):
        model = _load(v.AudioToTextCTC(64, 256, 29, 256), sd).set_storage_dtype(dtype)
        tokens, out_len = v.AsrPipeline(tr, model)(wav.to(DEV), _i32(lengths))
        audio, _ = tr.logmel_batch(wav.to(DEV), _i32(lengths))
        logits = model(audio).cpu()
        assert tokens.shape == (B, 501) and out_len.cpu().tolist() == [501] * B
        with torch.no_grad():
            model_ref = orc.asr_forward_storage_model(audio_ref, sd, dtype)
        check_asr_parity(f"configs[0] 8x10s {dtype}", ref, logits, tokens.cpu(), model_ref, tol)


def test_tts_fp16_storage():
    sd_a, sd_v, text, align, g = tts_case()
    V, H = int(g["cfg"][0]), int(g["cfg"][1])
    vmodel = _load(v.AlignTextToAudioModel(V, H), sd_v).set_storage_dtype(torch.float16)
    amodel = _load(v.TextToAlignTextModel(V, H), sd_a).set_storage_dtype(torch.float16)
    rep = orc.parity_report(torch.from_numpy(g["align_pred"]), amodel(text.to(DEV)).cpu())
    assert rep["rms_rel_std"] < 0.004, rep
    f0, logspc, codeap = vmodel.predict(torch.from_numpy(g["aligntext"]).to(DEV))
    rep = orc.parity_report(torch.from_numpy(g["logspc"]), logspc.cpu())
    print("fp16 logspc", rep)
    assert rep["rms_rel_std"] < 0.004 and rep["max_abs_rel_std"] < 0.03, rep


def test_asr_base_live_oracle():
    """asr_en_base size, BN calibrated on the batch, CUDA path vs the oracle run live on the host."""
    B, L = 4, 16000 * 3
    wav = torch.from_numpy(synth.noise_waveform(B, L, seed=77))
    lengths = [L, L - 5000, L - 16000, L]
    sd = orc.to_torch_sd(synth.asr_state_dict(64, 512, 29, 512, seed=77, randomize_bn=True))
    audio_ref, _ = orc.logmel_batch(wav, lengths)
    sd = orc.calibrate_asr(sd, audio_ref)
    with torch.no_grad():
        ref = orc.asr_forward(audio_ref, sd)
        model_ref = orc.asr_forward_storage_model(audio_ref, sd, torch.bfloat16)
    model = _load(v.AudioToTextCTC(64, 512, 29, 512), sd)
    tr = v.MelSpectrogramAudioTransform().to(DEV)
    tokens, _ = v.AsrPipeline(tr, model)(wav.to(DEV), _i32(lengths))
    audio, _ = tr.logmel_batch(wav.to(DEV), _i32(lengths))
    logits = model(audio).cpu()
    check_asr_parity("asr_en_base 4x3s", ref, logits, tokens.cpu(), model_ref)
    # sub-module API on NCW tensors (asr.py:78,93)
    enc = model.encoder(audio.transpose(1, 2).contiguous())
    with torch.no_grad():
        enc_ref = orc.asr_encoder(audio_ref.transpose(1, 2), sd)
    assert enc.shape == enc_ref.shape
    # (enc goes through one extra bf16 rounding of the fp32 log-mel features on entry)
    assert orc.parity_report(enc_ref, enc.cpu())["rms_rel_std"] < BF16_VS_FP32["rms_rel_std"] + 0.02


def test_tts_matches_golden():
    sd_a, sd_v, text, align, g = tts_case()
    V, H = int(g["cfg"][0]), int(g["cfg"][1])
    amodel = _load(v.TextToAlignTextModel(V, H), sd_a)
    vmodel = _load(v.AlignTextToAudioModel(V, H), sd_v)
    pred = amodel(text.to(DEV)).cpu()
    rep = orc.parity_report(torch.from_numpy(g["align_pred"]), pred)
    print("align", rep)
    assert rep["max_abs_rel_std"] < 0.08 and rep["rms_rel_std"] < 0.02, rep
    ats = [amodel.align(text[i], torch.from_numpy(align[i])) for i in range(text.shape[0])]
    at = torch.nn.utils.rnn.pad_sequence(ats, batch_first=True, padding_value=0)
    assert torch.equal(at, torch.from_numpy(g["aligntext"]))
    hasf0, f0_hat, logspc_hat, codeap_hat = vmodel(at.to(DEV))
    f0, logspc, codeap = vmodel.predict(at.to(DEV))
    torch.cuda.synchronize()
    assert logspc.shape == g["logspc"].shape and codeap.shape == g["codeap"].shape and f0.shape == g["f0"].shape
    for name, got, ref in (("hasf0", hasf0, g["hasf0_logits"]), ("logspc", logspc, g["logspc"]),
                           ("codeap", codeap, g["codeap"]), ("f0_hat", f0_hat, g["f0_hat"])):
        rep = orc.parity_report(torch.from_numpy(ref), got.cpu())
        print(name, rep)
        assert rep["max_abs_rel_std"] < 0.10 and rep["rms_rel_std"] < 0.02, (name, rep)
    # voiced/unvoiced gate agrees wherever the fp32 logit is not within bf16 error of zero
    safe = np.abs(g["hasf0_logits"]) > 2.5 * float(np.abs(hasf0.cpu().numpy() - g["hasf0_logits"]).max())
    assert ((f0.cpu().numpy() == 0) == (g["f0"] == 0))[safe].all()
    voiced = safe & (g["f0"] != 0)
    rep = orc.parity_report(torch.from_numpy(g["f0"][voiced]), f0.cpu()[torch.from_numpy(voiced)])
    assert rep["max_abs_rel_std"] < 0.10, rep
    # an id outside the vocabulary: nn.Embedding raises IndexError, and so does the drop-in
    bad = at.clone()
    bad[0, 3] = V
    with pytest.raises(IndexError):
        vmodel(bad.to(DEV))
    vmodel(at.to(DEV))                                   # the sticky flag was cleared by the failure above


def test_tts_v1_use_mcep_matches_golden():
    """AlignTextToAudioModel(use_mcep=True): 25 mel-cepstrum outputs (tts.py:153,164), reference-generated golden."""
    sd, aligntext, g = tts_v1_mcep_case()
    V, H = int(g["cfg"][0]), int(g["cfg"][1])
    model = _load(v.AlignTextToAudioModel(V, H, use_mcep=True), sd)
    assert model.logspc_size == 25 and model.audio_size == 28
    hasf0, f0_hat, mcep_hat, codeap_hat = model(aligntext.to(DEV))
    f0, mcep, codeap = model.predict(aligntext.to(DEV))
    torch.cuda.synchronize()
    assert mcep.shape == g["mcep"].shape and codeap.shape == g["codeap"].shape and f0.shape == g["f0"].shape
    for name, got, ref in (("hasf0", hasf0, g["hasf0_logits"]), ("mcep_hat", mcep_hat, g["mcep_hat"]),
                           ("mcep", mcep, g["mcep"]), ("codeap", codeap, g["codeap"]), ("f0_hat", f0_hat, g["f0_hat"])):
        rep = orc.parity_report(torch.from_numpy(ref), got.cpu())
        print("use_mcep", name, rep)
        assert rep["max_abs_rel_std"] < 0.10 and rep["rms_rel_std"] < 0.02, (name, rep)
    safe = np.abs(g["hasf0_logits"]) > 2.5 * float(np.abs(hasf0.cpu().numpy() - g["hasf0_logits"]).max())
    assert ((f0.cpu().numpy() == 0) == (g["f0"] == 0))[safe].all()


def test_tts_config2_full_size_sampled_oracle():
    """BASELINE.json configs[2] at the bench shape: tts_en_base, 256 x 100 tokens -> alignment -> aligned text ->
    WORLD parameters.  The CPU oracle cannot run 256 utterances in seconds: the whole batch is checked for
    determinism and batch invariance, the host alignment against the oracle's loop on every utterance, and the
    models against the oracle on three sampled utterances."""
    B, Ltxt, V, H = 256, 100, 29, 512
    text = torch.from_numpy(synth.text_tokens(B, Ltxt, V, seed=55))
    align = synth.synthetic_alignment(B, Ltxt, seed=55)
    sd_a = orc.to_torch_sd(synth.align_state_dict(V, H, seed=55, randomize_bn=True))
    sd_v = orc.to_torch_sd(synth.audio_state_dict(V, H, seed=55, randomize_bn=True, randomize_norm=True))
    sd_a = orc.calibrate_align(sd_a, text[:2])
    aligntext, at_len = v.align_batch(text, torch.from_numpy(align))
    for i in range(0, B, 17):
        ref_at = orc.align_text(text[i].tolist(), align[i])
        assert aligntext[i, : len(ref_at)].tolist() == list(ref_at) and int(at_len[i]) == len(ref_at)
    sd_v = orc.calibrate_audio(sd_v, aligntext[:2])
    amodel, vmodel = _load(v.TextToAlignTextModel(V, H), sd_a), _load(v.AlignTextToAudioModel(V, H), sd_v)
    pred = amodel(text.to(DEV))
    f0, logspc, codeap = vmodel.predict(aligntext.to(DEV))
    f0b, logspcb, _ = vmodel.predict(aligntext.to(DEV))
    torch.cuda.synchronize()
    T_out = 2 * aligntext.shape[1] - 1
    assert pred.shape == (B, Ltxt, 2) and logspc.shape == (B, T_out, 257) and f0.shape == (B, T_out)
    assert torch.equal(logspc, logspcb) and torch.equal(f0, f0b)                     # deterministic
    _, sub, _ = vmodel.predict(aligntext[96:104].contiguous().to(DEV))
    assert torch.equal(sub, logspc[96:104])                                          # batch-invariant
    for i in (0, 101, 255):
        with torch.no_grad():
            ref_pred = orc.align_forward(text[i:i + 1], sd_a)
            ref_f0, ref_logspc, ref_codeap = orc.audio_predict(aligntext[i:i + 1], sd_v)
            ref_hasf0 = orc.audio_forward(aligntext[i:i + 1], sd_v)[0]
        rep = orc.parity_report(ref_pred, pred[i:i + 1].cpu())
        assert rep["max_abs_rel_std"] < 0.08 and rep["rms_rel_std"] < 0.02, rep
        rep = orc.parity_report(ref_logspc, logspc[i:i + 1].cpu())
        print("configs[2] utt", i, "logspc", rep)
        assert rep["max_abs_rel_std"] < 0.10 and rep["rms_rel_std"] < 0.02, rep
        rep = orc.parity_report(ref_codeap, codeap[i:i + 1].cpu())
        assert rep["max_abs_rel_std"] < 0.10 and rep["rms_rel_std"] < 0.02, rep
        safe = ref_hasf0.abs() > 0.25 * ref_hasf0.std()
        assert ((f0[i:i + 1].cpu() == 0) == (ref_f0 == 0))[safe].all()


def test_asr_full_size_properties():
    """BASELINE.json configs[1] at full size (asr_en_base, 256 x 15 s, ragged lengths): the CPU oracle cannot
    run all of it in seconds, so check size-independent properties -- determinism, batch-composition
    invariance (utterances are independent: the sharding assumption of the multi-GPU path), the length
    formula, int16-PCM input == fp32 input -- plus the oracle on a sample of utterances."""
    B, L = 256, 240000
    g = torch.Generator(device=DEV).manual_seed(4321)
    pcm = torch.randint(-3277, 3277, (B, L), device=DEV, generator=g, dtype=torch.int32).to(torch.int16)
    wav = pcm.float() / 32768.0                                                  # what torchaudio.load returns
    lengths = torch.from_numpy(synth.ragged_lengths(B, 32000, L, seed=4321))
    # weights: BN calibrated by the oracle on a small batch so activations stay O(1)
    sd = orc.to_torch_sd(synth.asr_state_dict(64, 512, 29, 512, seed=4321, randomize_bn=True))
    calib_audio, _ = orc.logmel_batch(wav[:2, :48000].cpu(), [48000, 48000])
    sd = orc.calibrate_asr(sd, calib_audio)
    model = _load(v.AudioToTextCTC(64, 512, 29, 512), sd)
    pipe = v.AsrPipeline(v.MelSpectrogramAudioTransform().to(DEV), model)
    len_d = lengths.to(DEV)
    tokens, out_len = pipe(wav, len_d)
    tokens2, _ = pipe(wav, len_d)
    tokens_pcm, out_len_pcm = pipe(pcm, len_d)
    torch.cuda.synchronize()
    assert tokens.shape == (B, 751) and torch.equal(tokens, tokens2)                       # deterministic
    assert torch.equal(tokens, tokens_pcm) and torch.equal(out_len, out_len_pcm)           # int16 PCM == fp32 samples
    assert out_len.cpu().tolist() == [((1 + int(n) // 160) + 1) // 2 for n in lengths]     # asr.py:81-82
    sub, _ = pipe(wav[40:48].contiguous(), len_d[40:48].contiguous())
    assert torch.equal(sub, tokens[40:48])                                                 # batch-invariant
    # host path (pinned buffers, chunked H2D, async D2H) returns the same tokens
    tok_h, len_h = pipe.transcribe_host(wav.cpu().pin_memory(), lengths.pin_memory(), device=DEV)
    assert torch.equal(tok_h, tokens.cpu()) and len_h.tolist() == out_len.cpu().tolist()
    # streaming form: whole-batch graphs, two batches in flight on alternating buffer sets; fp32 and int16 PCM
    wav_h, len_p = wav.cpu().pin_memory(), lengths.pin_memory()
    pcm_r = pcm.flip(0).cpu().pin_memory()
    len_r = lengths.flip(0).contiguous().pin_memory()
    t1 = pipe.submit_host(wav_h, len_p, device=DEV, chunks=1)
    t2 = pipe.submit_host(pcm_r, len_r, device=DEV, chunks=1)
    t3 = pipe.submit_host(wav_h, len_p, device=DEV, chunks=1)
    assert torch.equal(t2.result()[0], tokens.flip(0).cpu())
    assert torch.equal(t3.result()[0], tokens.cpu()) and t3.result()[1].tolist() == out_len.cpu().tolist()
    del t1
    # oracle on a sample of utterances (features padded to the batch's frame count with BLANK_AUDIO)
    for i in (0, 131, 255):
        n = int(lengths[i])
        feat = orc.logmel_clip(wav[i, :n].cpu())
        audio = torch.full((1, 1501, 64), orc.BLANK_AUDIO)
        audio[0, : feat.shape[0]] = feat
        with torch.no_grad():
            ref = orc.asr_forward(audio, sd)
            model_ref = orc.asr_forward_storage_model(audio, sd, torch.bfloat16)
        logits = model(audio.to(DEV)).cpu()
        check_asr_parity(f"configs[1] utt {i}", ref, logits, tokens[i:i + 1].cpu(), model_ref, valid=[int(out_len[i])])


def test_asr_ja_phone_base_config4_full_size():
    """BASELINE.json configs[4]: asr_ja_phone_base (V = 44, hidden 512), variable-length padded batch of 256 clips
    U[2 s, 15 s]: determinism, batch invariance, BLANK_AUDIO padding, the oracle on sampled utterances, and the
    phone-vocabulary text tail (device CTC collapse + BasicTokenizer('ja'))."""
    B, L, V = 256, 240000, 44
    g = torch.Generator(device=DEV).manual_seed(99)
    wav = 0.1 * torch.randn((B, L), device=DEV, generator=g)
    lengths = torch.from_numpy(synth.ragged_lengths(B, 32000, L, seed=99))
    sd = orc.to_torch_sd(synth.asr_state_dict(64, 512, V, 512, seed=99, randomize_bn=True))
    calib_audio, _ = orc.logmel_batch(wav[:2, :48000].cpu(), [48000, 48000])
    sd = orc.calibrate_asr(sd, calib_audio)
    model = _load(v.AudioToTextCTC(64, 512, V, 512), sd)
    tr = v.MelSpectrogramAudioTransform().to(DEV)
    pipe = v.AsrPipeline(tr, model)
    len_d = lengths.to(DEV)
    tokens, out_len = pipe(wav, len_d)
    sub, _ = pipe(wav[200:208].contiguous(), len_d[200:208].contiguous())
    feats, audio_len = tr.logmel_batch(wav, len_d)
    torch.cuda.synchronize()
    assert tokens.shape == (B, 751) and int(tokens.max()) < V and torch.equal(sub, tokens[200:208])
    for i in (3, 77):                                     # frames past the clip's own length are BLANK_AUDIO
        n = int(audio_len[i])
        assert n == 1 + int(lengths[i]) // 160
        assert torch.all(feats[i, n:] == torch.tensor(orc.BLANK_AUDIO, device=DEV))
    for i in (5, 250):
        n = int(lengths[i])
        feat = orc.logmel_clip(wav[i, :n].cpu())
        audio = torch.full((1, 1501, 64), orc.BLANK_AUDIO)
        audio[0, : feat.shape[0]] = feat
        with torch.no_grad():
            ref = orc.asr_forward(audio, sd)
            model_ref = orc.asr_forward_storage_model(audio, sd, torch.bfloat16)
        logits = model(audio.to(DEV)).cpu()
        check_asr_parity(f"configs[4] utt {i}", ref, logits, tokens[i:i + 1].cpu(), model_ref, valid=[int(out_len[i])])
    ids, counts = pipe.transcribe_ids(wav[:8].contiguous(), len_d[:8].contiguous())
    tok = v.BasicTokenizer("ja")
    for b in range(8):
        n = int(out_len[b])
        want = tok.merge_repeated(tok.decode(tokens[b, :n].cpu().tolist()))
        assert tok.decode(ids[b, : int(counts[b])].cpu().tolist()) == want


def test_cuda_graph_replay_matches_eager():
    """BASELINE.json configs[0] shape (asr_en_small, 8 x 10 s) captured into a CUDA graph."""
    B, L = 8, 160000
    model = _load(v.AudioToTextCTC(64, 256, 29, 256), synth.asr_state_dict(64, 256, 29, 256, seed=5, randomize_bn=True))
    pipe = v.AsrPipeline(v.MelSpectrogramAudioTransform().to(DEV), model)
    run = pipe.graphed(B, L, device=DEV)
    for seed in (1, 2):
        wav = torch.from_numpy(synth.noise_waveform(B, L, seed=seed)).to(DEV)
        lengths = torch.from_numpy(synth.ragged_lengths(B, 20000, L, seed=seed)).to(DEV)
        tok_e, len_e = pipe(wav, lengths)
        tok_g, len_g = run(wav, lengths)
        torch.cuda.synchronize()
        assert tok_g.shape == (B, 501)
        assert torch.equal(tok_e, tok_g) and torch.equal(len_e, len_g)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_model_on_a_non_current_device():
    """A model and its inputs on cuda:1 while cuda:0 is the current device: every call runs on the tensors' device and
    that device's current stream (no torch.cuda.set_device needed); operands on two devices are refused."""
    from voice100_b200 import V100Error, kernels as K
    d1 = torch.device("cuda", 1)
    assert torch.cuda.current_device() == 0
    sd = synth.asr_state_dict(64, 128, 29, 128, seed=8, randomize_bn=True)
    wav = torch.from_numpy(synth.noise_waveform(2, 16000, seed=8))
    lens = torch.tensor([16000, 12000], dtype=torch.int32)
    m0 = _load(v.AudioToTextCTC(64, 128, 29, 128), sd)
    m1 = v.AudioToTextCTC(64, 128, 29, 128)
    m1.load_state_dict({k: torch.from_numpy(np.asarray(t)) for k, t in sd.items()})
    m1 = m1.to(d1).eval()
    t0, _ = v.AsrPipeline(v.MelSpectrogramAudioTransform().to(DEV), m0)(wav.to(DEV), lens.to(DEV))
    p1 = v.AsrPipeline(v.MelSpectrogramAudioTransform().to(d1), m1)
    t1, _ = p1(wav.to(d1), lens.to(d1))
    assert t1.device == d1 and torch.equal(t0.cpu(), t1.cpu())
    tok_h, _ = p1.transcribe_host(wav.pin_memory(), lens.pin_memory(), device=d1, chunks=1)
    assert torch.equal(tok_h, t0.cpu())
    assert torch.cuda.current_device() == 0
    with pytest.raises(V100Error):
        K.conv1x1(K.empty_ncw(1, 64, 16, d1), torch.zeros(8, 64, device=DEV, dtype=torch.bfloat16), None,
                  torch.zeros(8, device=DEV), 0)
