"""The streaming-tail definitions (oracle/tail_ref.py) against known answers and properties."""
import os

import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import tail_ref as TR

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_pack_known_answers():
    z = np.load(os.path.join(GOLD, "tail_kat.npz"))
    np.testing.assert_array_equal(TR.pack_i16(z["x"]), z["x_i16"])
    # hand-checked: ties go to even, +-0.99 -> +-32439, saturation at the int16 rails
    x = np.array([0.5 / 32767, 1.5 / 32767, 2.5 / 32767, 0.99, -0.99, 1.0, -1.0, 2.0, -2.0], dtype=np.float32)
    np.testing.assert_array_equal(TR.pack_i16(x), [0, 2, 2, 32439, -32439, 32767, -32767, 32767, -32768])


def test_tail_known_answers():
    z = np.load(os.path.join(GOLD, "tail_kat.npz"))
    f, i = TR.pcm_tail(z["cur"], z["prev"], z["w"], 0.99)
    np.testing.assert_array_equal(f, z["f_cf"])
    np.testing.assert_array_equal(i, z["i_cf"])
    f, i = TR.pcm_tail(z["cur"], None, TR.trim_fade_window(), 0.99)
    np.testing.assert_array_equal(f, z["f_tf"])
    np.testing.assert_array_equal(i, z["i_tf"])
    assert np.all(f[:, :480] == 0)


@settings(max_examples=50, deadline=None)
@given(st.lists(st.floats(-4, 4, width=32), min_size=1, max_size=64))
def test_pack_properties(xs):
    x = np.array(xs, dtype=np.float32)
    q = TR.pack_i16(x)
    assert q.dtype == np.int16
    order = np.argsort(x, kind="stable")
    assert np.all(np.diff(q[order].astype(np.int32)) >= 0)             # monotone
    unsat = np.abs(x) < 1.0
    np.testing.assert_array_equal(TR.pack_i16(-x)[unsat], -q[unsat])                                # odd below the rails
    inside = np.abs(x) <= 0.99
    assert np.all(np.abs(q[inside].astype(np.float64) - x[inside].astype(np.float64) * 32767.0) <= 0.5 + 1e-3)


def test_crossfade_is_identity_on_equal_signals_and_clamps():
    w = TR.fade_window(480)
    assert w[0] == 0.0 and abs(w[-1] - 1.0) < 1e-7 and np.all(np.diff(w) >= 0)
    rng = np.random.default_rng(1)
    cur = (rng.standard_normal((1, 960)) * 0.2).astype(np.float32)
    f, _ = TR.pcm_tail(cur, cur[:, :480].copy(), w, 0.99)
    np.testing.assert_allclose(f, cur, atol=1e-7)
    big = np.full((1, 16), 3.0, dtype=np.float32)
    f, i = TR.pcm_tail(big, None, None, 0.99)
    assert np.all(f == np.float32(0.99)) and np.all(i == 32439)


@settings(max_examples=30, deadline=None)
@given(st.integers(1, 700), st.integers(1, 150), st.integers(0, 20))
def test_chunk_plan_covers_every_frame_once(T, chunk, halo):
    owned = []
    for own_lo, own_hi, lo, hi, last in TR.chunk_plan(T, chunk, halo):
        assert 0 <= lo <= own_lo < own_hi <= hi <= T
        assert own_lo - lo <= halo and (last or hi >= min(T, own_hi + 1))
        owned += list(range(own_lo, own_hi))
    assert owned == list(range(T))


def test_stream_decode_ref_with_a_local_decoder_is_exact():
    # with a decode_fn that has no context dependence the chunked result equals the one-shot result
    rng = np.random.default_rng(2)
    T = 230
    mel = rng.standard_normal((1, 80, T)).astype(np.float32)
    s = (rng.standard_normal((1, 1, 480 * T)) * 0.3).astype(np.float32)

    def decode_fn(m, src):
        return src[:, 0, :] * np.float32(0.5)

    f, i = TR.stream_decode_ref(decode_fn, mel, s, chunk=100, halo=16)
    want = np.clip(s[:, 0, :] * np.float32(0.5), -0.99, 0.99)
    np.testing.assert_allclose(f, want, atol=1e-7)
    np.testing.assert_array_equal(i, TR.pack_i16(f))


def test_mulaw_matches_audioop_on_every_int16():
    """The mu-law oracle against a real reference implementation (CPython's audioop, G.711)."""
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore", DeprecationWarning)
        audioop = __import__("audioop")
    x = np.arange(-32768, 32768, dtype=np.int32).astype(np.int16)
    want = np.frombuffer(audioop.lin2ulaw(x.tobytes(), 2), dtype=np.uint8)
    np.testing.assert_array_equal(TR.mulaw_encode(x), want)
    # known answers: silence is 0xFF, full scale saturates at 0x80 / 0x00
    np.testing.assert_array_equal(TR.mulaw_encode(np.array([0, 32767, -32768, 4, -4], dtype=np.int16)),
                                  [0xFF, 0x80, 0x00, 0xFE, 0x7E])
