"""v2 models (LayerNorm/GELU conv blocks + bidirectional LSTM) on libv100 against the CPU oracle and the
reference's golden vectors.  16-bit storage vs fp32 reference: tolerances stated per check."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import v100_oracle as orc
import voice100_b200 as v
from voice100_b200 import kernels as K, synth, v2
from helpers import asr_v2_case, tts_v2_case

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rng_t(seed, *shape, scale=1.0):
    g = np.random.Generator(np.random.PCG64(seed))
    return torch.from_numpy((scale * g.standard_normal(shape)).astype(np.float32))


def _q(t, dtype=torch.bfloat16):
    return t.to(dtype).float()


# ---------------------------------------------------------------- kernels

@pytest.mark.parametrize("k,stride,pad", [(3, 1, 1), (3, 2, 1), (5, 1, 2), (5, 2, 2), (5, 1, 0)])
@pytest.mark.parametrize("B,C_in,C_out,T", [(3, 64, 256, 301), (2, 512, 512, 130), (1, 8, 24, 9)])
def test_conv1d_matches_torch(k, stride, pad, B, C_in, C_out, T):
    x = _q(_rng_t(1, B, C_in, T))
    w = _q(_rng_t(2, C_out, C_in, k, scale=1.0 / np.sqrt(C_in * k)))
    bias = _rng_t(3, C_out, scale=0.1)
    ref = F.conv1d(x, w, bias, stride=stride, padding=pad)
    xn = K.ncw_from_f32(x.to(DEV), torch.bfloat16)
    wp = w.permute(0, 2, 1).reshape(C_out, k * C_in).to(DEV, torch.bfloat16).contiguous()
    y = K.conv1d(xn, wp, bias.to(DEV), k, stride, pad)
    got = K.ncw_to_f32(y).cpu()
    assert got.shape == ref.shape
    # same bf16 inputs, fp32 accumulation, output rounded to bf16: |err| <= 2^-8 |ref| + accumulation-order noise
    err = (got - ref).abs()
    assert float((err - ref.abs() * 2.0 ** -8).max()) < 2e-3, float(err.max())


@pytest.mark.parametrize("k", [1, 3, 5])
@pytest.mark.parametrize("B,C_in,C_out,T", [(3, 64, 256, 77), (20, 512, 512, 41), (1, 128, 24, 9)])
def test_conv1d_time_major_matches_torch(k, B, C_in, C_out, T):
    """Dense conv evaluated on the time-major tensor: taps are column offsets of (j - pad) * Bp."""
    x = _q(_rng_t(11, B, C_in, T))
    w = _q(_rng_t(12, C_out, C_in, k, scale=1.0 / np.sqrt(C_in * k)))
    bias = _rng_t(13, C_out, scale=0.1)
    ref = F.conv1d(x, w, bias, stride=1, padding=(k - 1) // 2)
    tm = K.ncw_to_tm(K.ncw_from_f32(x.to(DEV), torch.bfloat16))
    wp = w.permute(0, 2, 1).reshape(C_out, k * C_in).to(DEV, torch.bfloat16).contiguous()
    y = K.conv1d_tm(tm, wp, bias.to(DEV), k)
    got = K.ncw_to_f32(K.tm_to_ncw(y)).cpu()
    assert got.shape == ref.shape
    err = (got - ref).abs()
    assert float((err - ref.abs() * 2.0 ** -8).max()) < 2e-3, float(err.max())
    # the padding utterances [B, Bp) see zero input: conv(0) = bias
    if tm.Bp > B:
        pad = y.data.float().view(C_out, T, tm.Bp)[:, :, B:].cpu()
        np.testing.assert_allclose(pad, bias.to(torch.bfloat16).float()[:, None, None].expand_as(pad), atol=1e-6)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("B,C,T", [(3, 512, 131), (2, 256, 64), (1, 24, 7)])
def test_layernorm_gelu_matches_torch(dtype, B, C, T):
    x = _q(_rng_t(4, B, C, T, scale=2.0), dtype)
    gamma, beta = 1.0 + 0.3 * _rng_t(5, C), 0.2 * _rng_t(6, C)
    ref = F.gelu(F.layer_norm(x.transpose(1, 2), (C,), gamma, beta, 1e-5).transpose(1, 2))
    xn = K.ncw_from_f32(x.to(DEV), dtype)
    got = K.ncw_to_f32(K.layernorm_gelu(xn, gamma.to(DEV), beta.to(DEV), 1e-5)).cpu()
    tol = 2.0 ** -8 if dtype == torch.bfloat16 else 2.0 ** -11
    err = (got - ref).abs()
    assert float((err - ref.abs() * tol).max()) < 1e-4, float(err.max())
    # padding columns are written as zeros (they feed later TMA boxes)
    assert float(xn.data[:, :, T:].float().abs().max() if xn.pitch > T else 0.0) == 0.0


@pytest.mark.parametrize("B,C,T", [(3, 64, 45), (33, 16, 70), (130, 8, 3)])
def test_time_major_round_trip(B, C, T):
    x = _q(_rng_t(7, B, C, T))
    xn = K.ncw_from_f32(x.to(DEV), torch.bfloat16)
    tm = K.ncw_to_tm(xn)
    assert tm.Bp % 8 == 0 and tm.data.shape == (C, T * tm.Bp)
    ref = torch.zeros(C, T, tm.Bp)
    ref[:, :, :B] = x.permute(1, 2, 0)
    assert torch.equal(tm.data.float().cpu().view(C, T, tm.Bp), ref)          # bit exact, pad columns zero
    back = K.tm_to_ncw(tm)
    assert torch.equal(K.ncw_to_f32(back).cpu(), x)


def _lstm_case(B, T, I, H, lengths, seed, dtype=torch.bfloat16):
    sd = {}
    synth._lstm(sd, seed, "lstm", I, H, 1, 2.0)
    sd = {k: _q(torch.from_numpy(val), dtype) if "weight" in k else torch.from_numpy(val) for k, val in sd.items()}
    x = _q(_rng_t(seed + 1, B, T, I), dtype)
    ref = orc.lstm_bidirectional(x, lengths, sd, "lstm", 1)                        # [B, T, 2H]
    xn = K.ncw_from_f32(x.transpose(1, 2).contiguous().to(DEV), dtype)
    w_ih = torch.cat([sd["lstm.weight_ih_l0"], sd["lstm.weight_ih_l0_reverse"]], 0).to(DEV, dtype).contiguous()
    w_hh = torch.stack([sd["lstm.weight_hh_l0"], sd["lstm.weight_hh_l0_reverse"]], 0).to(DEV, dtype).contiguous()
    bias = (torch.cat([sd["lstm.bias_ih_l0"], sd["lstm.bias_ih_l0_reverse"]], 0) +
            torch.cat([sd["lstm.bias_hh_l0"], sd["lstm.bias_hh_l0_reverse"]], 0)).to(DEV).contiguous()
    y = K.lstm_layer(K.ncw_to_tm(xn), w_ih, bias, w_hh, torch.tensor(lengths, dtype=torch.int32, device=DEV))
    got = K.ncw_to_f32(K.tm_to_ncw(y)).cpu().transpose(1, 2)                       # [B, T, 2H]
    return ref, got


# STATED TOLERANCE for one LSTM layer (|h| <= 1): the input projection and every h_t are rounded to the
# storage type (bf16: 2^-9 relative) and the gates use tanh.approx (2^-11); the recurrence damps old errors
# through the forget gate, so the error stays a small multiple of one rounding step.
LSTM_MAX_ABS = {torch.bfloat16: 0.01, torch.float16: 0.002}
LSTM_RMS = {torch.bfloat16: 0.0015, torch.float16: 0.0003}


@pytest.mark.parametrize("B,T,I,H,ragged", [(3, 20, 64, 64, True), (5, 37, 128, 256, True), (130, 12, 64, 128, True),
                                           (2, 60, 512, 512, False), (260, 9, 64, 512, True), (1, 1, 64, 64, False),
                                           (65, 2, 64, 192, True), (7, 33, 192, 320, True), (70, 5, 64, 448, True),
                                           (600, 3, 64, 128, True)])
def test_lstm_layer_matches_oracle(B, T, I, H, ragged):
    g = np.random.Generator(np.random.PCG64(B * 1000 + T))
    lengths = [int(n) for n in g.integers(1, T + 1, size=B)] if ragged else [T] * B
    lengths[0] = T
    ref, got = _lstm_case(B, T, I, H, lengths, seed=100 + B)
    err = (got - ref).abs()
    print("lstm", (B, T, I, H), "max", float(err.max()), "rms", float(err.pow(2).mean().sqrt()), "ref std", float(ref.std()))
    for b, n in enumerate(lengths):                       # pad_packed_sequence: exact zeros past the length
        assert float(got[b, n:].abs().max() if n < T else 0.0) == 0.0
    assert float(err.max()) < LSTM_MAX_ABS[torch.bfloat16] and float(err.pow(2).mean().sqrt()) < LSTM_RMS[torch.bfloat16]


def test_lstm_layer_fp16_storage():
    ref, got = _lstm_case(4, 30, 128, 256, [30, 11, 25, 1], seed=77, dtype=torch.float16)
    err = (got - ref).abs()
    print("lstm fp16 max", float(err.max()), "rms", float(err.pow(2).mean().sqrt()))
    assert float(err.max()) < LSTM_MAX_ABS[torch.float16] and float(err.pow(2).mean().sqrt()) < LSTM_RMS[torch.float16]


def test_lstm_is_deterministic_and_reusable():
    """Same inputs -> same bits on a second call with the same workspace (step counters are reset per call)."""
    a = _lstm_case(9, 25, 64, 128, [25, 3, 9, 25, 1, 7, 20, 13, 2], seed=5)[1]
    b = _lstm_case(9, 25, 64, 128, [25, 3, 9, 25, 1, 7, 20, 13, 2], seed=5)[1]
    assert torch.equal(a, b)


def test_v2_argument_errors():
    x = K.ncw_from_f32(torch.zeros(1, 8, 16, device=DEV), torch.bfloat16)
    w = torch.zeros(8, 7 * 8, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(v.V100Error, match="kernel_size 7"):
        K.conv1d(x, w, torch.zeros(8, device=DEV), 7, 1, 3)
    tm = K.ncw_to_tm(K.ncw_from_f32(torch.zeros(2, 96, 4, device=DEV), torch.bfloat16))
    with pytest.raises(v.V100Error, match="hidden size"):
        K.lstm_layer(tm, torch.zeros(8 * 96, 96, device=DEV, dtype=torch.bfloat16), torch.zeros(8 * 96, device=DEV),
                     torch.zeros(2, 4 * 96, 96, device=DEV, dtype=torch.bfloat16),
                     torch.ones(2, dtype=torch.int32, device=DEV))


# ---------------------------------------------------------------- models vs the reference's golden vectors

# STATED TOLERANCES (bf16 storage vs the fp32 reference), as fractions of the output's standard deviation.
V2_LOGIT_MAX_REL_STD = 0.10
V2_LOGIT_RMS_REL_STD = 0.025


def _load(model, sd):
    model.load_state_dict(sd, strict=False)
    return model.to(DEV).eval()


@pytest.mark.parametrize("name", ["asr_v2_en_small_ragged", "asr_v2_en_base"])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_asr_v2_matches_golden(name, dtype):
    sd, wav, lengths, settings, g = asr_v2_case(name)
    audio_size, hidden, vocab = [int(x) for x in g["cfg"][:3]]
    model = _load(v2.AudioToAlignText(audio_size, [list(r) for r in settings], 2, hidden, vocab), sd)
    model.set_storage_dtype(dtype)
    tr = v.MelSpectrogramAudioTransform().to(DEV)
    lens = torch.tensor(lengths, dtype=torch.int32, device=DEV)
    audio, audio_len = tr.logmel_batch(wav.to(DEV), lens)
    logits, logits_len = model(audio, audio_len)
    ref = torch.from_numpy(g["logits"])
    assert logits.shape == ref.shape and logits_len.tolist() == g["logits_len"].tolist()
    rep = orc.parity_report(ref, logits.cpu())
    print(name, dtype, rep)
    scale = 1.0 if dtype == torch.bfloat16 else 0.25
    assert rep["max_abs_rel_std"] < V2_LOGIT_MAX_REL_STD * scale and rep["rms_rel_std"] < V2_LOGIT_RMS_REL_STD * scale, rep
    # greedy tokens through the fused pipeline (log-mel on the device, argmax without materialising logits)
    pipe = v2.AsrV2Pipeline(tr, model)
    tokens, out_len = pipe(wav.to(DEV), lens)                      # [B, T'] like AsrPipeline
    assert out_len.cpu().tolist() == g["logits_len"].tolist()
    tok = tokens.t()[: ref.shape[0]].cpu()
    # pinned-host streaming form (CUDA graphs on alternating buffer sets) returns the same tokens
    tok_h, len_h = pipe.transcribe_host(wav.pin_memory(), lens.cpu().pin_memory(), device=DEV, chunks=1)
    assert torch.equal(tok_h, tokens.cpu()) and len_h.tolist() == g["logits_len"].tolist()
    margin = 2.5 * rep["max_abs"]
    top2 = ref.topk(2, dim=-1).values
    sure = (top2[..., 0] - top2[..., 1]) > margin
    valid = torch.arange(ref.shape[0])[:, None] < torch.from_numpy(g["logits_len"])[None, :]
    agree = (tok == ref.argmax(-1))
    print(name, "token agreement raw %.4f, decisive frames %.2f" % (float(agree[valid].float().mean()), float(sure[valid].float().mean())))
    assert bool(agree[sure & valid].all())


def test_tts_v2_matches_golden():
    sd_a, sd_v, text, align, g = tts_v2_case()
    V = int(g["cfg"][0])
    amodel = _load(v2.TextToAlignText(V, 2, 256, 2), sd_a)
    text_len = torch.from_numpy(g["text_len"])
    pred, pred_len = amodel(text.to(DEV), text_len)
    ref = torch.from_numpy(g["align_pred"])
    assert pred.shape == ref.shape and pred_len.tolist() == g["align_pred_len"].tolist()
    rep = orc.parity_report(ref, pred.cpu())
    print("align v2", rep)
    assert rep["max_abs_rel_std"] < V2_LOGIT_MAX_REL_STD and rep["rms_rel_std"] < V2_LOGIT_RMS_REL_STD, rep
    # host alignment: integer work, bit exact against the reference's own output
    for i, n in enumerate(text_len.tolist()):
        at = v2.TextToAlignText.align(text[i, :n], torch.from_numpy(align[i, :n]))
        np.testing.assert_array_equal(at.numpy(), g["aligntext"][i, :len(at)])
        assert len(at) == int(g["aligntext_len"][i])
    vmodel = _load(v2.AlignTextToAudio(V, 257, 1, 2, 512, [list(r) for r in synth.TTS_V2_BASE_DECODER]), sd_v)
    aligntext, at_len = torch.from_numpy(g["aligntext"]).to(DEV), torch.from_numpy(g["aligntext_len"])
    hasf0, f0_hat, logspc_hat, hascodeap, codeap_hat = vmodel(aligntext, at_len)
    for name, got, key in (("hasf0", hasf0, "hasf0_logits"), ("f0_hat", f0_hat, "f0_hat"),
                           ("hascodeap", hascodeap, "hascodeap_logits")):
        ref = torch.from_numpy(g[key])
        assert got.shape == ref.shape
        rep = orc.parity_report(ref, got.cpu())
        print("tts v2", name, rep)
        assert rep["max_abs_rel_std"] < V2_LOGIT_MAX_REL_STD and rep["rms_rel_std"] < V2_LOGIT_RMS_REL_STD, rep
    f0, logspc, codeap = vmodel.predict(aligntext, at_len)
    ref = torch.from_numpy(g["logspc"])
    rep = orc.parity_report(ref, logspc.cpu())
    print("tts v2 logspc", rep)
    assert logspc.shape == ref.shape and rep["rms_rel_std"] < V2_LOGIT_RMS_REL_STD
    sure = np.abs(g["hasf0_logits"]) > 2.5 * 0.25 * float(np.std(g["hasf0_logits"]))
    np.testing.assert_allclose(f0.cpu().numpy()[sure], g["f0"][sure], rtol=0, atol=0.25 * float(np.std(g["f0"])))
    assert codeap.shape == g["codeap"].shape


def test_tts_v2_mcep_head_matches_golden():
    """AlignTextToAudioPredict: predict + mel-cepstrum -> log-spectrum folded into the projection GEMM."""
    from helpers import golden
    from voice100_b200.vocoder import AlignTextToAudioPredict
    g = golden("tts_v2_mcep")
    V, B, seed = [int(x) for x in g["cfg"]]
    sd = {k: torch.from_numpy(val) for k, val in synth.audio_v2_state_dict(
        V, 25, 1, 2, 512, synth.TTS_V2_BASE_DECODER, seed=seed, randomize_ln=True, randomize_norm=True, gain=2.0).items()}
    model = _load(v2.AlignTextToAudio(V, 25, 1, 2, 512, [list(r) for r in synth.TTS_V2_BASE_DECODER]), sd)
    wrap = AlignTextToAudioPredict(model).to(DEV)
    aligntext, lens = torch.from_numpy(g["aligntext"]).to(DEV), torch.from_numpy(g["aligntext_len"])
    f0, logspc, codeap = wrap(aligntext, lens)
    ref = torch.from_numpy(g["logspc"])
    assert logspc.shape == ref.shape == (B, 79, 257) and f0.shape == (B, 79) and codeap.shape == (B, 79, 1)
    # the spectrum carries a large per-bin offset (the un-normalised c0): compare the variation around the bin means
    centred = lambda t: t - ref.mean(dim=(0, 1), keepdim=True)
    rep = orc.parity_report(centred(ref), centred(logspc.cpu()))
    print("tts v2 mcep->logspc", rep)
    assert rep["max_abs_rel_std"] < V2_LOGIT_MAX_REL_STD and rep["rms_rel_std"] < V2_LOGIT_RMS_REL_STD, rep
    # and the un-fused route agrees with the fused head: predict() -> 25 mel-cepstra -> matrix on the host
    _, mcep, _ = model.predict(aligntext, lens)
    unfused = mcep.double().cpu() @ torch.from_numpy(orc.mc2sp_matrix(512, 24, 0.410))
    rep2 = orc.parity_report(centred(ref), centred(unfused.float()))
    print("  un-fused route", rep2)
    assert rep2["rms_rel_std"] < V2_LOGIT_RMS_REL_STD


def test_asr_v2_forced_alignment_matches_oracle():
    """AudioToAlignText.ctc_best_path: model forward + log-softmax + batched Viterbi vs the oracle's restatement of
    the reference DP run on the oracle's fp32 logits (paths compared where the two logit sets agree on the winner)."""
    sd, wav, lengths, settings, g = asr_v2_case("asr_v2_en_small_ragged")
    audio_size, hidden, vocab = [int(x) for x in g["cfg"][:3]]
    model = _load(v2.AudioToAlignText(audio_size, [list(r) for r in settings], 2, hidden, vocab), sd)
    model.set_storage_dtype(torch.float16)
    tr = v.MelSpectrogramAudioTransform().to(DEV)
    audio, audio_len = tr.logmel_batch(wav.to(DEV), torch.tensor(lengths, dtype=torch.int32, device=DEV))
    text = torch.from_numpy(synth.text_tokens(3, 9, vocab, seed=7)).to(DEV)
    text_len = torch.tensor([9, 4, 6])
    score, hist, path, logits_len = model.ctc_best_path(audio, audio_len, text, text_len)
    assert path.shape == (3, 76) and logits_len.tolist() == g["logits_len"].tolist()
    ref_lp = torch.log_softmax(torch.from_numpy(g["logits"]), dim=-1)      # [T', B, V] from the reference
    for b in range(3):
        n, m = int(logits_len[b]), int(text_len[b])
        rs, rh, rp = orc.ctc_best_path(ref_lp[:n, b].numpy(), text[b, :m].cpu().numpy())
        assert abs(float(score[b]) - float(rs)) < 0.05 * n * 0.01 + 0.05    # fp16 logits vs fp32 logits
        # every frame's label is one of the text's labels or blank, in order, and the path decodes to the text
        p = path[b, :n].cpu().numpy()
        collapsed = [int(x) for i, x in enumerate(p) if x != 0 and (i == 0 or x != p[i - 1] or hist[b, i] != hist[b, i - 1])]
        assert collapsed == text[b, :m].cpu().tolist()
        assert (p == rp).mean() > 0.9                                        # same alignment up to near-ties
    toks = model.ctc_best_path(audio, audio_len)
    assert toks.shape == (76, 3)
