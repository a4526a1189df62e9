"""Intra-sentence streaming: mel arrives in pieces, PCM leaves in chunks.

The bar: the incremental stream is EXACTLY the offline chunked stream of the whole utterance (which
tests/test_gpu_decode.py ties to the oracle), for any way of cutting the mel into pieces — int16 PCM bit for bit."""
import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

from oracle import hift_ref as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hift(lib, cuda_device):
    from gonova_tts_b200 import B200HiFT, random_state_dict

    return B200HiFT(random_state_dict(0, True), device=cuda_device, dtype="bf16")   # "corners": voiced and unvoiced f0


def _pieces(T, cuts):
    edges = sorted({0, T, *[c for c in cuts if 0 < c < T]})
    return list(zip(edges[:-1], edges[1:]))


def test_source_stream_pieces_concatenate_to_the_whole(hift, cuda_device):
    B, T = 2, 300
    mel = R.synthetic_mel(B, T, seed=3).to(cuda_device)
    f0 = hift.predict_f0(mel)
    whole = hift.source_from_f0(f0, seed=17)
    parts, acc = [], None
    for a, b in _pieces(T, [1, 2, 9, 100, 101, 250]):
        s, acc = hift.source_stream(f0[:, a:b].contiguous(), 17, a, acc)
        parts.append(s)
    got = torch.cat(parts, dim=2)
    assert torch.equal(got, whole)
    assert float(whole.abs().max()) > 0.01


@settings(max_examples=15, deadline=None)
@given(T=st.integers(1, 460), cuts=st.lists(st.integers(1, 459), max_size=8), seed=st.integers(0, 99))
def test_incremental_stream_equals_offline_stream(hift, cuda_device, T, cuts, seed):
    from gonova_tts_b200 import IncrementalDecoder, StreamingDecoder

    B = 2
    mel = R.synthetic_mel(B, T, seed=seed).to(cuda_device)
    want_i16, want_f32 = StreamingDecoder(hift).decode_all(mel, seed=5)
    inc = IncrementalDecoder(hift, B=B, seed=5, want_f32=True)
    chunks = []
    for a, b in _pieces(T, cuts):
        chunks += inc.push(mel[:, :, a:b])
        assert inc._mel.shape[2] <= 100 + 2 * 16 + 5 + 1 + (b - a)      # history stays bounded
    chunks += inc.finish()
    got_i16 = torch.cat([c[0] for c in chunks], dim=1)
    got_f32 = torch.cat([c[1] for c in chunks], dim=1)
    assert got_i16.shape == want_i16.shape == (B, T * 480)
    assert torch.equal(got_i16, want_i16), (T, cuts)
    assert torch.equal(got_f32, want_f32)
    assert len(chunks) == -(-T // 100)


def test_incremental_latency_is_the_lookahead_and_errors(hift, cuda_device):
    from gonova_tts_b200 import IncrementalDecoder

    inc = IncrementalDecoder(hift, B=1, seed=1)
    mel = R.synthetic_mel(1, 260, seed=1).to(cuda_device)
    assert inc.push(mel[:, :, :121]) == []                           # 100 + 1 + 16 + 5 = 122 frames release chunk 0
    out = inc.push(mel[:, :, 121:122])
    assert len(out) == 1 and out[0][0].shape == (1, 100 * 480) and out[0][0].dtype == torch.int16
    assert len(inc.push(mel[:, :, 122:260])) == 1                    # chunk 1 (needs 222 frames); chunk 2 waits for the end
    rest = inc.finish()
    assert len(rest) == 1 and rest[0][0].shape == (1, 60 * 480)
    with pytest.raises(RuntimeError):
        inc.push(mel[:, :, :1])
    with pytest.raises(ValueError):
        IncrementalDecoder(hift, B=1).push(torch.zeros(2, 80, 3, device=cuda_device))
