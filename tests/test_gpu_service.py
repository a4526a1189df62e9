"""The drop-in itself, driven the way the reference service drives it:
engine object -> service.install() -> generate()-shaped calls -> PcmSink, against the oracle sitting in the same
place.  The engine is tests/fake_engine.py (the real one is not installable here)."""
import asyncio
import os
import sys

import numpy as np
import pytest
import torch

from conftest import snr_db
from fake_engine import FakeEngine
from oracle import hift_ref as R

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def sd():
    from gonova_tts_b200 import random_state_dict

    return random_state_dict(0, False)


@pytest.fixture(scope="module")
def engines(lib, cuda_device, sd):
    """(engine with the B200 decoder installed, reference engine with the oracle vocoder on the CPU)."""
    from gonova_tts_b200 import B200HiFT
    from gonova_tts_b200.service import install

    eng = FakeEngine(sd, device=cuda_device)
    old = eng.s3gen.mel2wav
    dec = install(eng, dtype="tf32")
    assert isinstance(dec, B200HiFT) and eng.s3gen.mel2wav is dec and dec is not old
    assert "mel2wav" in dict(eng.s3gen.named_children())           # a registered child, like upstream's
    return eng, FakeEngine(sd, device="cpu"), dec


def test_install_then_generate_matches_the_oracle_engine(engines, cuda_device):
    """Upstream's keyword call `mel2wav.inference(speech_feat=..., cache_source=torch.zeros(1, 1, 0))` + trim_fade.  The
    NSF source is stochastic, so the waveform is compared through the returned source fed to the oracle's decode."""
    eng, ref, dec = engines
    text = "Hello, this is a warmup test."
    wav = eng.generate(text, audio_prompt_path=None, exaggeration=0.5, cfg_weight=0.5, temperature=0.8)
    assert wav.shape == (1, 480 * FakeEngine.frames_for(text)) and wav.device.type == "cuda"
    mel = ref.mel_for(text)
    _, src = dec.inference(speech_feat=mel.to(cuda_device), cache_source=torch.zeros(1, 1, 0, device=cuda_device), seed=5)
    got, _ = dec.inference(speech_feat=mel.to(cuda_device), cache_source=src)       # full-length cache: deterministic
    with torch.inference_mode():
        want = ref.s3gen.mel2wav.decode(mel, src.cpu())
    err, snr = float((got.cpu() - want).abs().max()), snr_db(got.cpu().numpy(), want.numpy())
    print(f"[parity] installed decoder vs oracle engine: max-abs {err:.3e} SNR {snr:.1f} dB")
    assert err <= 2e-4 and snr >= 58.0
    assert float(wav[0, :480].abs().max()) == 0.0                   # trim_fade zeroed the first 20 ms
    # the service's own tail (synthesizer.py:352-357): squeeze().cpu().numpy() float32
    audio = wav.squeeze().cpu().numpy()
    assert audio.dtype == np.float32 and audio.ndim == 1


def test_pcm_sink_f32_is_byte_identical_to_the_reference_tail(engines, cuda_device):
    """server.py:150-155 sends `audio_chunk.tobytes()` of `.cpu().numpy().astype(float32)`: PcmSink(fmt='f32') must
    produce exactly those bytes, as bytes and as a view of pinned memory."""
    from gonova_tts_b200.service import PcmSink

    eng, _, _ = engines
    wav = eng.generate("The quick brown fox.")
    want = wav.squeeze().cpu().numpy().astype(np.float32).tobytes()
    sink = PcmSink(cuda_device, max_samples=1000, fmt="f32")        # too small on purpose: grows
    assert sink.to_bytes(wav) == want
    mv = sink.to_memoryview(wav)
    assert isinstance(mv, memoryview) and mv.readonly and mv.nbytes == len(want) and bytes(mv) == want
    mv2 = sink.to_memoryview(wav * 0.5)                             # rotating buffers: the first view is still intact
    assert bytes(mv) == want and bytes(mv2) != want
    # trim_fade inside the sink == upstream's in-place multiply
    raw, _ = eng.s3gen.mel2wav.inference(eng.last_mel, seed=9)
    a = raw.clone()
    a[:, :960] *= R.trim_fade_window().to(cuda_device)
    assert sink.to_bytes(raw, trim_fade=True) == a.squeeze().cpu().numpy().tobytes()
    # int16 / mu-law sinks agree with the numpy definitions
    from oracle import tail_ref as TR

    _, want_i = TR.pcm_tail(raw.cpu().numpy(), None, None, 0.99)
    assert PcmSink(cuda_device, fmt="i16").to_bytes(raw) == want_i.tobytes()
    assert PcmSink(cuda_device, fmt="mulaw").to_bytes(raw) == TR.mulaw_encode(want_i.reshape(-1)).tobytes()


def test_chunk_tap_streams_inside_generate(engines, cuda_device):
    """service.chunk_tap: the engine's single mel2wav.inference call hands out 2 s chunks while it runs; their
    concatenation is the returned waveform, and that is the chunked decode (== StreamingDecoder) of the sentence."""
    from gonova_tts_b200 import StreamingDecoder
    from gonova_tts_b200.service import chunk_tap

    eng, ref, dec = engines
    text = "x" * 60                                                  # 324 frames: 4 chunks
    got = []
    with chunk_tap(dec, lambda pcm, cid, last: got.append((bytes(pcm), cid, last)), fmt="f32", as_memoryview=True):
        wav = eng.generate(text)
    assert [c for _, c, _ in got] == [0, 1, 2, 3] and [l for _, _, l in got] == [False, False, False, True]
    stream = np.concatenate([np.frombuffer(b, dtype=np.float32) for b, _, _ in got])
    assert stream.shape == (480 * 324,)
    # generate() applied trim_fade once more to what the tap already faded (documented): undo nothing, compare tails
    np.testing.assert_array_equal(stream[960:], wav.squeeze().cpu().numpy()[960:])
    assert np.all(stream[:480] == 0)
    # outside the tap the decoder is back to one call = one decode
    n_before = len(got)
    eng.generate("Hello.")
    assert len(got) == n_before
    # the stream is the offline chunked decode of the same mel + source, bit for bit
    with chunk_tap(dec, lambda *a: None):
        w2, src = dec.inference(ref.mel_for(text).to(cuda_device), seed=77)
    _, f32 = StreamingDecoder(dec, trim_fade=True).decode_all(ref.mel_for(text).to(cuda_device), src, want_i16=False)
    assert torch.equal(w2, f32)
    # and inside the fp32 tolerance of the oracle's single-shot decode
    with torch.inference_mode():
        want = ref.s3gen.mel2wav.decode(ref.mel_for(text), src.cpu())
    want = want.clone()
    want[:, :960] *= R.trim_fade_window()
    err, snr = float((w2.cpu() - want).abs().max()), snr_db(w2.cpu().numpy(), want.numpy())
    print(f"[parity] chunk_tap stream vs oracle single-shot decode: max-abs {err:.3e} SNR {snr:.1f} dB")
    assert err <= 2e-4 and snr >= 58.0


def test_patched_synthesizer_yields_chunks_and_stats(engines, cuda_device):
    """examples/patched_synthesizer.py: the patched `_generate_sentence` (synthesizer.py:296-321) yields float32 numpy
    chunks of <= 2 s while generate() runs; get_stats() carries the decoder's counters (synthesizer.py:411-420)."""
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    from patched_synthesizer import StreamingSynthesizerPatch
    from gonova_tts_b200.service import patch_get_stats

    eng, _, dec = engines

    class Synth(StreamingSynthesizerPatch):
        """The slice of the reference's StreamingSynthesizer the patch touches."""

        def __init__(self):
            self.model, self.device, self.decoder = eng, cuda_device, dec
            self.stats = {"syntheses": 0}

        def get_stats(self):
            return self.stats.copy()

        def _synthesize_sync(self, text, voice_embedding=None, exaggeration=0.25):       # synthesizer.py:327-359
            audio = self.model.generate(text, audio_prompt_path=None, exaggeration=exaggeration, cfg_weight=0.5,
                                        temperature=0.8)
            return audio.squeeze().cpu().numpy().astype(np.float32)

    synth = Synth()
    patch_get_stats(synth, dec)
    before = synth.get_stats()["decoder"]

    async def run():
        return [c async for c in synth._generate_sentence("y" * 50, None, 0.25)]

    chunks = asyncio.run(run())
    assert [c.shape[0] for c in chunks] == [48000, 48000, 480 * 274 - 96000]
    assert all(c.dtype == np.float32 for c in chunks)
    after = synth.get_stats()["decoder"]
    assert after["frames"] > before["frames"] and after["calls"] > before["calls"]
    assert after["plans"]["built"] >= 1 and after["dtype"] == "tf32" and after["audio_seconds"] == after["frames"] / 50

    async def failing():
        synth._synthesize_sync = lambda *a: (_ for _ in ()).throw(RuntimeError("engine failed"))
        return [c async for c in synth._generate_sentence("z", None, 0.25)]

    with pytest.raises(RuntimeError, match="engine failed"):
        asyncio.run(failing())


def test_parent_load_state_dict_rebuilds_the_decoder(lib, cuda_device, sd):
    """`S3Gen.load_state_dict(strict=False)` with `mel2wav.*` entries reaches B200HiFT._load_from_state_dict: the
    handle is rebuilt from them (here: new weights change the output; loading the old ones restores it)."""
    from gonova_tts_b200 import random_state_dict
    from gonova_tts_b200.service import install

    eng = FakeEngine(sd, device=cuda_device)
    dec = install(eng, dtype="bf16", max_frames=64)
    mel = R.synthetic_mel(1, 16, seed=3).to(cuda_device)
    s = torch.zeros(1, 1, 16 * 480, device=cuda_device)
    a = dec.decode(mel, s).clone()
    other = {"mel2wav." + k: v for k, v in random_state_dict(1, False).items()}
    res = eng.s3gen.load_state_dict(other, strict=False)
    assert not res.unexpected_keys
    b = dec.decode(mel, s).clone()
    assert not torch.equal(a, b)
    eng.s3gen.load_state_dict({"mel2wav." + k: v for k, v in sd.items()}, strict=False)
    assert torch.equal(dec.decode(mel, s), a)
    assert dec.state_dict() == {} and list(dec.parameters()) == []
    with pytest.raises(RuntimeError, match="missing|weight"):
        eng.s3gen.load_state_dict({"mel2wav.conv_pre.bias": torch.zeros(512)}, strict=False)


def test_install_flow_swaps_the_cfm_decoder(lib, cuda_device):
    """service.install_flow on an engine-shaped object: `s3gen.flow.decoder` (estimator + noise buffer + Euler loop) becomes a
    B200Flow that answers upstream's keyword call with the oracle module's mel (same weights, same noise buffer)."""
    from types import SimpleNamespace

    from gonova_tts_b200 import B200Flow
    from gonova_tts_b200.service import install_flow
    from oracle import flow_ref as FR

    class FakeFlowHolder(torch.nn.Module):
        def __init__(self, cfm):
            super().__init__()
            self.decoder = cfm

    cfm = FR.CausalConditionalCFM(FR.make_estimator(2), noise_seed=11)
    s3gen = torch.nn.Module()
    s3gen.flow = FakeFlowHolder(cfm)
    eng = SimpleNamespace(s3gen=s3gen)
    new = install_flow(eng, dtype="tf32", device=cuda_device)
    assert isinstance(new, B200Flow) and eng.s3gen.flow.decoder is new
    assert "decoder" in dict(eng.s3gen.flow.named_children())
    B, T = 1, 45
    _, mu, mask, spks, cond = FR.synthetic_inputs(B, T, seed=4)
    want, _ = cfm(mu=mu, mask=mask, spks=spks, cond=cond, n_timesteps=10)
    dev = cuda_device
    got, none = eng.s3gen.flow.decoder(mu=mu.to(dev), mask=mask.to(dev), spks=spks.to(dev), cond=cond.to(dev), n_timesteps=10)
    assert none is None
    err, snr = float((got.cpu() - want).abs().max()), snr_db(got.cpu().numpy(), want.numpy())
    print(f"[parity] installed flow decoder vs oracle CFM module: max-abs {err:.3e} SNR {snr:.1f} dB")
    assert err <= 1e-2 and snr >= 55.0


def test_install_flow_rebinds_flow_inference(lib, cuda_device):
    """An engine whose flow module holds the whole front (input_embedding / encoder / encoder_proj / spk_embed_affine_layer,
    upstream's CausalMaskedDiffWithXvec): install_flow also rebinds `flow.inference`, so the engine's own keyword call runs
    tokens -> mel on this library and returns what the oracle modules return."""
    from types import SimpleNamespace

    from gonova_tts_b200.service import install_flow
    from oracle import flow_enc_ref as ER
    from oracle import flow_ref as FR

    class FakeFlow(ER.FlowFront):
        def __init__(self, cfm):
            super().__init__()
            self.decoder = cfm

        def inference(self, token, token_len, prompt_token, prompt_token_len, prompt_feat, prompt_feat_len, embedding, finalize):
            return ER.flow_inference(self, self.decoder, token, prompt_token, prompt_feat, embedding), None

    flow = FakeFlow(FR.CausalConditionalCFM(FR.make_estimator(3), noise_seed=5))
    flow.load_state_dict(ER.random_state_dict(4), strict=False)
    flow.eval()
    s3gen = torch.nn.Module()
    s3gen.flow = flow
    eng = SimpleNamespace(s3gen=s3gen)
    g = torch.Generator().manual_seed(8)
    token = torch.randint(0, ER.VOCAB, (1, 40), generator=g, dtype=torch.int32)
    prompt_token = torch.randint(0, ER.VOCAB, (1, 10), generator=g, dtype=torch.int32)
    prompt_feat = torch.randn(1, 20, 80, generator=g) * 0.5
    emb = torch.randn(1, 192, generator=g)
    kw = dict(token=token, token_len=torch.tensor([40]), prompt_token=prompt_token, prompt_token_len=torch.tensor([10]),
              prompt_feat=prompt_feat, prompt_feat_len=torch.tensor([20]), embedding=emb, finalize=True)
    with torch.inference_mode():
        want, _ = eng.s3gen.flow.inference(**kw)
    new = install_flow(eng, dtype="tf32", device=cuda_device)
    assert eng.s3gen.flow.decoder is new and new.flow_inference.front is not None
    eng.s3gen.eval()                                     # the module tree stays a tree (no cycle through flow_inference)
    assert sum(1 for _ in eng.s3gen.modules()) < 400 and "decoder.flow_inference" not in dict(eng.s3gen.flow.named_modules())
    got, none = eng.s3gen.flow.inference(**{k: (v.to(cuda_device) if k in ("token", "prompt_token", "prompt_feat", "embedding") else v)
                                            for k, v in kw.items()})
    assert none is None and got.shape == want.shape == (1, 80, 80)
    err, snr = float((got.cpu() - want).abs().max()), snr_db(got.cpu().numpy(), want.numpy())
    print(f"[parity] installed flow.inference (tokens -> mel) vs the oracle modules: max-abs {err:.3e} SNR {snr:.1f} dB")
    assert err <= 5e-3 and snr >= 59.0
