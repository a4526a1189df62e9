"""The front of the flow step (token embedding, upsampling Conformer encoder, encoder_proj, speaker projection)
on a B200 against its CPU oracle (oracle/flow_enc_ref.py; parity unpinned: the engine is not vendored).  Both sides hold the
same seeded weights and get the same tokens."""
import numpy as np
import pytest
import torch

from conftest import snr_db
from oracle import flow_enc_ref as ER
from oracle import flow_ref as FR

pytestmark = pytest.mark.gpu

# (max-abs, SNR dB) of mu against the fp32 oracle, at about 2x what the kernels deliver (measured on B200: tf32 1.4e-3 /
# 64.5 dB, bf16 1.2e-2 / 46.4 dB at |mu| up to 2.4).  Operands are rounded to 10 (tf32) / 7 (bf16) mantissa bits through
# ~45 GEMMs; LayerNorm and the softmax run in fp32 on both paths.
TOL = {"tf32": (3e-3, 58.0), "bf16": (2.5e-2, 40.0)}


@pytest.fixture(scope="module")
def sd():
    return ER.random_state_dict(0)


@pytest.fixture(scope="module")
def oracle(sd):
    return ER.load_front(sd)


@pytest.fixture(scope="module")
def fronts(lib, cuda_device, sd):
    from gonova_tts_b200.flow_front import B200FlowFront

    cache = {}

    def get(dtype):
        if dtype not in cache:
            cache[dtype] = B200FlowFront(sd, device=cuda_device, dtype=dtype)
        return cache[dtype]

    return get


def _check(got, want, dtype, what):
    err, snr = np.abs(got - want).max(), snr_db(got, want)
    print(f"[parity] flow front {what} {dtype}: max-abs {err:.3e}  SNR {snr:.1f} dB  (|mu| max {np.abs(want).max():.2f})")
    assert err <= TOL[dtype][0] and snr >= TOL[dtype][1], (err, snr)


@pytest.mark.parametrize("dtype", ["tf32", "bf16"])
def test_tokens_to_mu_and_spks(fronts, oracle, cuda_device, dtype):
    B, L = 2, 50
    tokens, token_len, emb = ER.synthetic_tokens(B, L, seed=5)
    with torch.inference_mode():
        want = oracle.encode(tokens, token_len).numpy()
        want_s = oracle.speaker(emb).numpy()
    mu, spks = fronts(dtype).encode(tokens.to(cuda_device), token_len.to(cuda_device), emb.to(cuda_device))
    assert mu.shape == (B, 80, 2 * L) and spks.shape == (B, 80)
    _check(mu.cpu().numpy(), want, dtype, f"B={B} L={L}")
    assert np.abs(spks.cpu().numpy() - want_s).max() <= 1e-5          # fp32 on both sides


@pytest.mark.parametrize("dtype", ["tf32", "bf16"])
def test_ragged_batch_rows_equal_the_utterance_alone(fronts, oracle, cuda_device, dtype):
    """Lengths around the tile sizes (16-row attention blocks, 32-key tiles, the 3-token look-ahead), an empty utterance, and
    a workspace poisoned with NaNs: every utterance must come out as if it were encoded alone, frames past its length zero."""
    B, L = 5, 70
    lengths = [70, 33, 1, 0, 48]
    tokens, token_len, emb = ER.synthetic_tokens(B, L, seed=9, lengths=lengths)
    with torch.inference_mode():
        want = oracle.encode(tokens, token_len).numpy()
    f = fronts(dtype)
    f._workspace(B, L).view(torch.float32)[: f.workspace_bytes(B, L) // 4].fill_(float("nan"))
    mu, _ = f.encode(tokens.to(cuda_device), token_len.to(cuda_device), emb.to(cuda_device))
    got = mu.cpu().numpy()
    assert np.isfinite(got).all()
    for b, n in enumerate(lengths):
        assert not got[b, :, 2 * n:].any(), f"utterance {b}: frames past its length are not zero"
    _check(got, want, dtype, f"ragged {lengths}")
    # and bit-identical to that utterance in a batch of its own
    alone, _ = f.encode(tokens[1:2, :33].to(cuda_device).contiguous(), None, None)
    assert torch.equal(alone[0], mu[1, :, :66])


def test_long_utterance_bf16(fronts, oracle, cuda_device):
    """20 s of speech tokens (25 Hz): positions far from the origin in the relative-position table."""
    B, L = 1, 500
    tokens, token_len, emb = ER.synthetic_tokens(B, L, seed=11)
    with torch.inference_mode():
        want = oracle.encode(tokens, token_len).numpy()
    mu, _ = fronts("bf16").encode(tokens.to(cuda_device), None, None)
    _check(mu.cpu().numpy(), want, "bf16", f"B={B} L={L}")


def test_flow_inference_call_shape(lib, cuda_device, sd, oracle):
    """Upstream's `flow.inference(token, token_len, prompt_token, prompt_token_len, prompt_feat, prompt_feat_len, embedding,
    finalize)`: tokens -> mel through the encoder AND the ten-step CFM decoder, against the two oracles chained."""
    from gonova_tts_b200.flow_front import B200FlowInference

    est_sd = FR.random_state_dict(0)
    full = dict(sd)
    full.update({"decoder.estimator." + k: v for k, v in est_sd.items()})
    cfm = FR.CausalConditionalCFM(FR.load_estimator(est_sd), noise_seed=0)
    g = torch.Generator().manual_seed(21)
    token = torch.randint(0, ER.VOCAB, (1, 30), generator=g, dtype=torch.int32)
    prompt_token = torch.randint(0, ER.VOCAB, (1, 12), generator=g, dtype=torch.int32)
    prompt_feat = torch.randn(1, 24, 80, generator=g) * 0.5
    emb = torch.randn(1, 192, generator=g)
    with torch.inference_mode():
        want = ER.flow_inference(oracle, cfm, token, prompt_token, prompt_feat, emb).numpy()
    flow = B200FlowInference(full, device=cuda_device, dtype="tf32", noise_seed=0)
    got, _ = flow.inference(token=token.to(cuda_device), token_len=torch.tensor([30]), prompt_token=prompt_token.to(cuda_device),
                            prompt_token_len=torch.tensor([12]), prompt_feat=prompt_feat.to(cuda_device),
                            prompt_feat_len=torch.tensor([24]), embedding=emb.to(cuda_device), finalize=True)
    got = got.cpu().numpy()
    assert got.shape == want.shape == (1, 80, 2 * 42 - 24)
    err, snr = np.abs(got - want).max(), snr_db(got, want)
    print(f"[parity] flow.inference tf32, tokens -> mel: max-abs {err:.3e}  SNR {snr:.1f} dB")
    assert err <= 5e-3 and snr >= 59.0, (err, snr)        # measured 1.9e-3 / 65.7 dB
    # finalize=False (a chunk in the middle of a token stream): the last 3 tokens' 6 frames are held back before the decoder
    with torch.inference_mode():
        want_open = ER.flow_inference(oracle, cfm, token, prompt_token, prompt_feat, emb, finalize=False).numpy()
    got_open, _ = flow.inference(token=token.to(cuda_device), token_len=torch.tensor([30]), prompt_token=prompt_token.to(cuda_device),
                                 prompt_token_len=torch.tensor([12]), prompt_feat=prompt_feat.to(cuda_device),
                                 prompt_feat_len=torch.tensor([24]), embedding=emb.to(cuda_device), finalize=False)
    got_open = got_open.cpu().numpy()
    assert got_open.shape == want_open.shape == (1, 80, 2 * 42 - 24 - 6)
    err, snr = np.abs(got_open - want_open).max(), snr_db(got_open, want_open)
    print(f"[parity] flow.inference tf32, finalize=False: max-abs {err:.3e}  SNR {snr:.1f} dB")
    assert err <= 5e-3 and snr >= 59.0, (err, snr)


@pytest.mark.parametrize("dtype", ["bf16", "tf32"])
def test_tensor_core_attention_equals_the_cuda_core_kernel(fronts, cuda_device, tmp_path, dtype):
    """Both handles run the relative-position attention on mma.sync (bf16 m16n8k16 / tf32 m16n8k8: scores from two MMAs,
    rel_shift as a skewed read-back); GONOVA_ENC_ATTN_MMA=0 (read once per process, hence a child process) keeps the fp32
    CUDA-core kernel.  Lengths around the 64-query / 64-key tiles."""
    import os
    import subprocess
    import sys

    B, L = 4, 130
    lengths = [130, 64, 65, 7]
    tokens, token_len, emb = ER.synthetic_tokens(B, L, seed=13, lengths=lengths)
    mu, _ = fronts(dtype).encode(tokens.to(cuda_device), token_len.to(cuda_device), None)
    got = mu.cpu().numpy()
    out = tmp_path / "simt.npy"
    code = (
        "import sys, numpy as np, torch\n"
        f"sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})\n"
        "from oracle import flow_enc_ref as ER\n"
        "from gonova_tts_b200.flow_front import B200FlowFront\n"
        f"tokens, token_len, emb = ER.synthetic_tokens({B}, {L}, seed=13, lengths={lengths})\n"
        f"f = B200FlowFront(ER.random_state_dict(0), device='cuda:0', dtype={dtype!r})\n"
        "mu, _ = f.encode(tokens.to('cuda:0'), token_len.to('cuda:0'), None)\n"
        f"np.save({str(out)!r}, mu.cpu().numpy())\n"
    )
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, GONOVA_ENC_ATTN_MMA="0"), capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    ref = np.load(out)
    snr = snr_db(got, ref)
    print(f"[parity] flow front {dtype}: mma.sync attention against the CUDA-core kernel: max-abs {np.abs(got - ref).max():.3e}  SNR {snr:.1f} dB")
    assert snr >= (40.0 if dtype == "bf16" else 60.0)


@pytest.mark.parametrize("dtype", ["tf32", "bf16"])
@pytest.mark.parametrize("L", [1, 2, 3, 5, 251])
def test_tiny_and_odd_token_counts(fronts, oracle, cuda_device, dtype, L):
    """Fewer tokens than the look-ahead (3) and than the conv kernels' taps; one token = a 1 x 1 attention with a one-row
    position table; 251 tokens = 502 frames, not a multiple of any tile."""
    tokens, token_len, emb = ER.synthetic_tokens(2, L, seed=30 + L, lengths=[L, max(1, L - 1)])
    with torch.inference_mode():
        want = oracle.encode(tokens, token_len).numpy()
    mu, _ = fronts(dtype).encode(tokens.to(cuda_device), token_len.to(cuda_device), None)
    _check(mu.cpu().numpy(), want, dtype, f"L={L}")


def test_argument_and_weight_errors(lib, cuda_device, sd, fronts):
    """The C ABI's error behaviour: a missing or mis-shaped weight names the tensor; a too-small or unaligned workspace, a
    spks / embedding mismatch and non-positive sizes are refused before anything is launched."""
    import ctypes as C

    from gonova_tts_b200 import _cabi
    from gonova_tts_b200.flow_front import B200FlowFront

    bad = dict(sd)
    del bad["encoder.up_layer.conv.weight"]
    with pytest.raises(RuntimeError, match="encoder.up_layer.conv.weight"):
        B200FlowFront(bad, device=cuda_device, dtype="bf16")
    bad = dict(sd)
    bad["encoder.encoders.2.self_attn.pos_bias_u"] = torch.zeros(8, 32)
    with pytest.raises(RuntimeError, match="encoder.encoders.2.self_attn"):
        B200FlowFront(bad, device=cuda_device, dtype="bf16")
    f = fronts("bf16")
    B, L = 2, 16
    tok = torch.zeros(B, L, dtype=torch.int32, device=cuda_device)
    mu = torch.empty(B, 80, 2 * L, device=cuda_device)
    spks = torch.empty(B, 80, device=cuda_device)
    ws = torch.empty(f.workspace_bytes(B, L) + 2048, dtype=torch.uint8, device=cuda_device)
    base = ws.data_ptr() + (-ws.data_ptr()) % 1024
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def call(tokens=tok, emb=None, spk=None, b=B, l=L, wptr=base, wbytes=f.workspace_bytes(B, L)):
        return lib.gnv_flow_encode(f._h, C.c_void_p(tokens.data_ptr()) if tokens is not None else None, None,
                                   C.c_void_p(emb.data_ptr()) if emb is not None else None, b, l, C.c_void_p(mu.data_ptr()),
                                   C.c_void_p(spk.data_ptr()) if spk is not None else None, C.c_void_p(wptr), wbytes, st)

    assert call() == 0
    for kw, msg in ((dict(tokens=None), "NULL"), (dict(spk=spks), "go together"), (dict(b=0), "positive"), (dict(l=0), "positive"),
                    (dict(wptr=base + 8), "aligned"), (dict(wbytes=1024), "too small")):
        assert call(**kw) != 0, kw
        assert msg in _cabi.last_error(None), (kw, _cabi.last_error(None))
    torch.cuda.synchronize()


def test_tokens_to_waveform_like_s3gen_inference(lib, cuda_device, sd):
    """B200Token2Wav.inference(speech_tokens, ref_dict=...) = upstream S3Token2Wav.inference: the mel against the two flow
    oracles chained, and the waveform against the vocoder oracle's decode of THAT mel with the source the call returned
    (the NSF source is stochastic), times trim_fade."""
    from gonova_tts_b200 import B200Token2Wav, random_state_dict
    from oracle import hift_ref as R

    est_sd = FR.random_state_dict(0)
    hift_sd = random_state_dict(0, False)
    full = {"flow." + k: v for k, v in sd.items()}
    full.update({"flow.decoder.estimator." + k: v for k, v in est_sd.items()})
    full.update({"mel2wav." + k: v for k, v in hift_sd.items()})
    t2w = B200Token2Wav.from_state_dict(full, device=cuda_device, dtype="tf32", noise_seed=0)
    g = torch.Generator().manual_seed(33)
    tokens = torch.randint(0, ER.VOCAB, (25,), generator=g, dtype=torch.int32)          # 1-D like the engine's call
    ref = {"prompt_token": torch.randint(0, ER.VOCAB, (1, 10), generator=g, dtype=torch.int32), "prompt_token_len": torch.tensor([10]),
           "prompt_feat": torch.randn(1, 20, 80, generator=g) * 0.5, "prompt_feat_len": None, "embedding": torch.randn(1, 192, generator=g)}
    dev_ref = {k: (v.to(cuda_device) if isinstance(v, torch.Tensor) and k != "prompt_token_len" else v) for k, v in ref.items()}
    wav, src = t2w.inference(tokens.to(cuda_device), ref_dict=dev_ref)
    assert wav.shape == (1, 480 * 50) and src.shape == (1, 1, 480 * 50)
    front = ER.load_front(sd)
    cfm = FR.CausalConditionalCFM(FR.load_estimator(est_sd), noise_seed=0)
    with torch.inference_mode():
        mel_want = ER.flow_inference(front, cfm, tokens.unsqueeze(0), ref["prompt_token"], ref["prompt_feat"], ref["embedding"])
        mel_got = t2w.flow_inference(tokens.to(cuda_device), dev_ref).cpu()
        err_m = float((mel_got - mel_want).abs().max())
        hift = R.load_model(hift_sd)
        want = hift.decode(mel_got, src.cpu()).clone()
        want[:, :960] *= R.trim_fade_window()
    err = float((wav.cpu() - want).abs().max())
    print(f"[parity] tokens -> waveform tf32: mel max-abs {err_m:.3e}; wav vs oracle decode of that mel + source {err:.3e}")
    assert err_m <= 5e-3 and err <= 1e-3
    assert not wav[:, :480].any()                                                     # trim_fade silences the first 20 ms
    with pytest.raises(ValueError, match="ref_dict"):
        t2w.inference(tokens.to(cuda_device))


def test_batched_requests_equal_single_requests(lib, cuda_device, sd):
    """B200Token2Wav.flow_inference_batch / inference_batch: three requests with different token counts, prompts and voices as
    ONE ragged batch; every mel equals that request's own flow_inference (masking, not approximation), and the waveforms have
    each request's own length and trim_fade."""
    from gonova_tts_b200 import B200Token2Wav, random_state_dict

    est_sd = FR.random_state_dict(0)
    full = {"flow." + k: v for k, v in sd.items()}
    full.update({"flow.decoder.estimator." + k: v for k, v in est_sd.items()})
    full.update({"mel2wav." + k: v for k, v in random_state_dict(0, False).items()})
    t2w = B200Token2Wav.from_state_dict(full, device=cuda_device, dtype="tf32", noise_seed=0)
    g = torch.Generator().manual_seed(44)
    reqs = []
    for n_tok, n_prompt in ((31, 10), (12, 20), (50, 5)):
        ref = {"prompt_token": torch.randint(0, ER.VOCAB, (1, n_prompt), generator=g, dtype=torch.int32).to(cuda_device),
               "prompt_token_len": torch.tensor([n_prompt]), "prompt_feat": (torch.randn(1, 2 * n_prompt, 80, generator=g) * 0.5).to(cuda_device),
               "prompt_feat_len": None, "embedding": torch.randn(1, 192, generator=g).to(cuda_device)}
        reqs.append((torch.randint(0, ER.VOCAB, (n_tok,), generator=g, dtype=torch.int32).to(cuda_device), ref))
    mels = t2w.flow_inference_batch(reqs)
    for (tokens, ref), mel in zip(reqs, mels):
        alone = t2w.flow_inference(tokens, ref)
        assert mel.shape == alone.shape == (1, 80, 2 * tokens.numel())
        snr = snr_db(mel.cpu().numpy(), alone.cpu().numpy())
        print(f"[parity] batched request ({tokens.numel()} tokens) against the same request alone: SNR {snr:.1f} dB")
        assert snr >= 50.0
    wavs = t2w.inference_batch(reqs)
    for (tokens, _), w in zip(reqs, wavs):
        assert w.shape == (1, 480 * 2 * tokens.numel()) and torch.isfinite(w).all()
        assert not w[:, :480].any() and w[:, 960:].abs().max() > 0
    assert t2w.inference_batch([]) == []


def test_request_batcher_over_tokens_to_pcm(lib, cuda_device, sd):
    """batching.for_token2wav: concurrent (tokens, ref_dict) requests come back as each caller's own waveform."""
    from gonova_tts_b200 import B200Token2Wav, random_state_dict
    from gonova_tts_b200.batching import for_token2wav

    full = {"flow." + k: v for k, v in sd.items()}
    full.update({"flow.decoder.estimator." + k: v for k, v in FR.random_state_dict(0).items()})
    full.update({"mel2wav." + k: v for k, v in random_state_dict(0, False).items()})
    t2w = B200Token2Wav.from_state_dict(full, device=cuda_device, dtype="bf16", noise_seed=0)
    g = torch.Generator().manual_seed(45)
    ref = {"prompt_token": torch.randint(0, ER.VOCAB, (1, 8), generator=g, dtype=torch.int32).to(cuda_device), "prompt_token_len": None,
           "prompt_feat": (torch.randn(1, 16, 80, generator=g) * 0.5).to(cuda_device), "prompt_feat_len": None,
           "embedding": torch.randn(1, 192, generator=g).to(cuda_device)}
    rb = for_token2wav(t2w, max_batch=8)
    counts = [20, 7, 33, 12, 25]
    futs = [rb.submit((torch.randint(0, ER.VOCAB, (n,), generator=g, dtype=torch.int32).to(cuda_device), ref)) for n in counts]
    for n, f in zip(counts, futs):
        w = f.result(120)
        assert w.shape == (1, 480 * 2 * n) and torch.isfinite(w).all() and not w[:, :480].any()
    rb.close()
    assert rb.metrics["requests"] == 5 and rb.metrics["batches"] <= 5
